"""ORACLE — test infrastructure only.  CPU (torch) restatement of llm-compressor's AWQModifier,
which the reference builds at ref/src/quantool/methods/llm_compressor/awq/awq.py:81.

llm-compressor is not installed and has no source on this box: restated from SURVEY.md §B.1-§B.4
(`_accumulate_mean`, `_compute_best_scale`, `_pseudo_quantize_tensor`, `_compute_loss`, smoothing
application).  Tensor dtypes are what torch would produce: weights and `w_mean` in the model
dtype, `x_mean` and the candidate scales in fp32, losses as python floats.
Property pinned in tests/test_oracle_cpu.py: folding the scales leaves the block's function unchanged (fp64).
PARITY UNPINNED: the reference's tests hold no AWQ vectors; dtype choices that
SURVEY.md does not state (x_mean fp32, scales fp32) are this restatement's and are documented in
DESIGN.md.
"""
from typing import Callable, List, Optional, Tuple

import torch

N_GRID = 20


def accumulate_mean(x: torch.Tensor, state: Optional[Tuple[torch.Tensor, int]]):
    """forward hook on balance_layers[0]: per-channel sum of |x| over all tokens, and the count."""
    x = x.reshape(-1, x.shape[-1])
    s = x.abs().to(torch.float32).sum(dim=0)
    n = x.shape[0]
    if state is None:
        return s, n
    return state[0] + s, state[1] + n


def weight_mean(balance_weights: List[torch.Tensor], group_size: int) -> torch.Tensor:
    weight = torch.cat(balance_weights, dim=0)
    org_shape = weight.shape
    gs = group_size if group_size and group_size > 0 else org_shape[1]
    w = weight.view(-1, gs).abs()
    w = w / (w.amax(dim=1, keepdim=True) + 1e-6)
    w = w.view(org_shape)
    return w.mean(0)


def pseudo_quantize_tensor(w: torch.Tensor, symmetric: bool, bit_width: int, group_size: int) -> torch.Tensor:
    org_w_shape = w.shape
    if group_size and group_size > 0:
        w = w.reshape(-1, group_size)
    if not symmetric:
        max_val = w.amax(dim=1, keepdim=True)
        min_val = w.amin(dim=1, keepdim=True)
        max_int = 2 ** bit_width - 1
        min_int = 0
        scales = (max_val - min_val).clamp(min=1e-5) / max_int
        zeros = (-torch.round(min_val / scales)).clamp_(min_int, max_int)
        w = (torch.clamp(torch.round(w / scales) + zeros, min_int, max_int) - zeros) * scales
    else:
        max_val = w.abs().amax(dim=1, keepdim=True)
        max_val = max_val.clamp(min=1e-5)
        max_int = 2 ** (bit_width - 1) - 1
        min_int = -(2 ** (bit_width - 1))
        scales = max_val / max_int
        w = torch.clamp(torch.round(w / scales), min_int, max_int) * scales
    return w.reshape(org_w_shape)


def candidate_scales(x_mean: torch.Tensor, w_mean: torch.Tensor, ratio: float, duo_scaling: bool = True):
    if duo_scaling:
        scales = (x_mean.pow(ratio) / (w_mean.pow(1 - ratio) + 1e-4)).clamp(min=1e-4)
    else:
        scales = x_mean.pow(ratio).clamp(min=1e-4).view(-1)
    scales = scales / (scales.max() * scales.min()).sqrt()
    scales[torch.isinf(scales)] = 1
    scales[torch.isnan(scales)] = 1
    return scales


def compute_loss(fp16_outputs: List[torch.Tensor], int_w_outputs: List[torch.Tensor]) -> float:
    loss = 0.0
    num_elements = 0
    for a, b in zip(fp16_outputs, int_w_outputs):
        loss += (a - b).float().pow(2).sum().item()
        num_elements += a.numel()
    return loss / num_elements


def gram_loss_single_linear(x_batches: List[torch.Tensor], w: torch.Tensor, w_candidate: torch.Tensor) -> float:
    """The loss `compute_loss` gives for a parent that is ONE Linear (F.linear(x, w)), written through the Gram matrix
    G = X^T X:  sum ||x (Wc - W)^T||^2 / numel = tr(D G D^T) / numel with D = Wc - W.  This is the form the CUDA path
    evaluates (qt_awq_gram_loss); here in fp64 so tests can pin both the identity and the size of what it leaves out
    (upstream rounds each parent output to the model dtype before taking the difference; the Gram form does not)."""
    K = w.shape[1]
    G = torch.zeros((K, K), dtype=torch.float64)
    numel = 0
    for x in x_batches:
        x2 = x.reshape(-1, K).double()
        G += x2.t() @ x2
        numel += x2.shape[0] * w.shape[0]
    D = w_candidate.double() - w.double()
    return float(((D @ G) * D).sum()) / numel


def compute_best_scale(x_mean: torch.Tensor, w_mean: torch.Tensor, balance_weights: List[torch.Tensor],
                       parent_forward: Callable[[List[torch.Tensor]], List[torch.Tensor]], fp16_outputs,
                       symmetric: bool, bit_width: int, group_size: int, n_grid: int = N_GRID,
                       duo_scaling: bool = True):
    """parent_forward(list of patched balance weights) -> list of parent outputs (one per batch).
    Returns (best_scales, best_ratio, history)."""
    history = []
    best_ratio = -1
    best_scales = None
    best_error = float("inf")
    org = [w.clone() for w in balance_weights]
    for grid_idx in range(n_grid):
        ratio = grid_idx / n_grid
        scales = candidate_scales(x_mean, w_mean, ratio, duo_scaling)
        sv = scales.view(1, -1)
        patched = []
        for w in org:
            ws = w.clone()
            ws.mul_(sv)                                             # model dtype *= fp32 scales
            q = pseudo_quantize_tensor(ws, symmetric, bit_width, group_size) / sv   # fp32
            wq = torch.empty_like(w)
            wq.copy_(q)                                             # update_offload_parameter -> model dtype
            patched.append(wq)
        outs = parent_forward(patched)
        loss = compute_loss(fp16_outputs, outs)
        history.append(loss)
        if loss < best_error:
            best_error = loss
            best_ratio = ratio
            best_scales = scales.clone()
    assert best_ratio != -1 and not torch.isnan(best_scales).any()
    return best_scales, best_ratio, history


def apply_scales(smooth_weight: torch.Tensor, balance_weights: List[torch.Tensor], scales: torch.Tensor):
    """In place: balance W *= s[None, :]; smooth layer: 1-D (norm) weight /= s, 2-D (Linear)
    weight[-len(s):] /= s[:, None]."""
    for w in balance_weights:
        w.mul_(scales.view(1, -1))
    if smooth_weight.ndim == 1:
        smooth_weight.div_(scales)
    else:
        smooth_weight[-scales.size(0):].div_(scales.view(-1, 1))
