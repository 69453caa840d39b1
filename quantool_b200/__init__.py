"""quantool_b200 — B200-native implementation of quantool's per-layer quantization math
behind quantool's method-registry plugin API."""
import os as _os

# One decoder layer runs its four distinct inputs on four streams, each with a high-priority look-ahead stream and
# a "far" stream of its own (csrc/linalg.cu), next to the copy and NCCL streams.  With the default 8 hardware work
# queues two of those streams alias to one queue and serialise behind each other (measured: the fourth input's
# whole chain started only when the K = 14336 chain's look-ahead stream had drained, 29 ms into a 72 ms phase;
# profiles/r02_timeline_n1_before.json).  Takes effect only if set before the CUDA context exists.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .core import BaseQuantizer, QuantizerRegistry, TemplateQuantizationCard

__all__ = ["BaseQuantizer", "QuantizerRegistry", "TemplateQuantizationCard"]
__version__ = "0.1.0"
