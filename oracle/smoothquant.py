"""ORACLE — test infrastructure only.  CPU (torch) restatement of llm-compressor's
SmoothQuantModifier, which the reference puts in front of GPTQ at
ref/src/quantool/methods/llm_compressor/smoothquant/smoothquant.py:77-84.

llm-compressor is not installed and has no source on this box: restated from SURVEY.md §C
(`_calculate_smoothing_scales`, `_apply_smoothing`, the per-channel min/max forward hook).
All arithmetic runs in the tensors' own dtype exactly as torch would (bf16 tensors -> bf16 ops).
Property pinned in tests/test_oracle_cpu.py: folding the scales leaves the block's function unchanged (fp64).
PARITY UNPINNED: the reference's tests hold no SmoothQuant vectors.
"""
from typing import List, Optional, Tuple

import torch

MINIMUM_SMOOTHING_SCALE = 1e-5


def update_channel_minmax(out: torch.Tensor, mn: Optional[torch.Tensor], mx: Optional[torch.Tensor]):
    """Forward hook on the smooth layer's OUTPUT: running per-channel min and max."""
    o = out.reshape(-1, out.shape[-1])
    lo = torch.min(o, dim=0)[0]
    hi = torch.max(o, dim=0)[0]
    if mn is None:
        return lo, hi
    return torch.minimum(mn, lo), torch.maximum(mx, hi)


def calculate_smoothing_scales(balance_weights: List[torch.Tensor], activation_scales: torch.Tensor,
                               smoothing_strength: float) -> torch.Tensor:
    weight_scales = []
    for w in balance_weights:
        weight_scales.append(w.abs().max(dim=0, keepdim=True)[0])
    weight_scales = 2.0 * torch.cat(weight_scales, dim=0).max(dim=0)[0]
    scales = activation_scales.pow(smoothing_strength) / weight_scales.pow(1 - smoothing_strength)
    scales = torch.where(weight_scales > 0.0, scales, activation_scales)
    return scales


def smoothing_scales(mn: torch.Tensor, mx: torch.Tensor, balance_weights: List[torch.Tensor],
                     smoothing_strength: float = 0.5) -> torch.Tensor:
    activation_scales = mx - mn          # dynamic range, NOT absmax
    scales = calculate_smoothing_scales(balance_weights, activation_scales, smoothing_strength)
    return torch.maximum(scales, torch.Tensor([MINIMUM_SMOOTHING_SCALE]).to(scales.device))


def apply_smoothing(smooth_weight: torch.Tensor, balance_weights: List[torch.Tensor], scales: torch.Tensor):
    """In place: balance W *= s[None, :];  smooth (norm) weight /= s."""
    for w in balance_weights:
        w.mul_(scales.view(1, -1))
    if smooth_weight.ndim == 1:
        smooth_weight.div_(scales)
    else:
        smooth_weight.div_(scales.view(-1, 1))
