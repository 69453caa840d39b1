"""SmoothQuant for one decoder layer: per-channel activation min/max, smoothing scales, fold.

Mirrors UPSTREAM llmcompressor SmoothQuantModifier, first modifier of the
reference's recipe at ref/src/quantool/methods/llm_compressor/smoothquant/smoothquant.py:77-84.
Host glue over the C-ABI kernels (qt_channel_minmax, qt_smooth_scales, qt_scale_matrix).
"""
from typing import Dict, List, Optional

import torch

from .. import cabi
from . import llama

# Llama default mappings: balance layers <- smooth layer, and the captured input name
MAPPINGS = [
    ("input_layernorm", ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"], "attn_in"),
    ("post_attention_layernorm", ["mlp.gate_proj", "mlp.up_proj"], "mlp_in"),
]


def compute_scales(act_min: torch.Tensor, act_max: torch.Tensor, balance_weights: List[torch.Tensor],
                   alpha: float) -> torch.Tensor:
    """s = (max-min)^alpha / (2 max|W|)^(1-alpha), evaluated in the weights' dtype; fp32 [K]."""
    K = act_min.shape[0]
    wmn, wmx = cabi.new_minmax(K, act_min.device)
    for w in balance_weights:
        cabi.channel_minmax(w, wmn, wmx)
    return cabi.smooth_scales(act_min, act_max, wmn, wmx, alpha, balance_weights[0].dtype)


def apply_scales(smooth_weight: torch.Tensor, balance_weights: List[torch.Tensor], s: torch.Tensor) -> None:
    for w in balance_weights:
        cabi.scale_matrix_(w, s)                       # W *= s[None, :]
    cabi.scale_matrix_(smooth_weight, s, divide=True)  # norm weight /= s


def smooth_layer(shape: llama.LlamaShape, w: Dict[str, torch.Tensor], calib, cos=None, sin=None, alpha: float = 0.5,
                 chunk_samples: int = 8, dist=None) -> Dict[str, torch.Tensor]:
    """One calibration pass over this rank's samples, min/max all-reduce, fold in place into `w`.
    `calib`: a pipeline.CalibSet, or hidden states [n, seq, hidden] together with `cos`/`sin`.
    Returns the smoothing scales per mapping (keyed by smooth-layer name)."""
    from .pipeline import CalibSet
    if isinstance(calib, torch.Tensor):
        calib = CalibSet.from_hidden(calib, cos, sin)
    dev = calib.groups[0].device if calib.groups else next(iter(w.values())).device
    dtype = w["input_layernorm.weight"].dtype
    dims = shape.input_dims()
    names = [m[2] for m in MAPPINGS]
    stats = {n: cabi.new_minmax(dims[n], dev) for n in names}
    chunk_tokens = max(1, chunk_samples) * max(calib.max_len, 1)
    cap = {n: torch.empty((chunk_tokens, k), dtype=dtype, device=dev) for n, k in dims.items()}
    for gi, a, b in calib.chunks(chunk_tokens):
        hb = calib.groups[gi][a:b]
        rows = hb.shape[0] * hb.shape[1]
        llama.layer_forward(shape, w, hb, *calib.ropes[gi], capture=cap, row0=0, stop_after="mlp_in")
        for n in names:
            cabi.channel_minmax(cap[n][:rows], *stats[n])
    out = {}
    for smooth_name, balance, inp in MAPPINGS:
        mn, mx = stats[inp]
        if dist is not None and dist.on:
            dist.all_reduce_min(mn)
            dist.all_reduce_max(mx)
        bw = [w[f"{b}.weight"] for b in balance]
        s = compute_scales(mn, mx, bw, alpha)
        apply_scales(w[f"{smooth_name}.weight"], bw, s)
        out[smooth_name] = s
    return out
