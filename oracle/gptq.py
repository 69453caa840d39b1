"""ORACLE — test infrastructure only.  CPU (torch fp32) restatement of llm-compressor's GPTQ.

The reference builds `GPTQModifier(...)` at ref/src/quantool/methods/llm_compressor/gptq/gptq.py:86
and runs it through `llmcompressor.oneshot` at ref/src/quantool/methods/llm_compressor/base.py:159-161.
llm-compressor (pin `>=0.8.1`, ref/pyproject.toml:49-51) is NOT installed and has no source on
this box, so `accumulate_hessian` / `quantize_weight` are restated from SURVEY.md §A.1-§A.6.
Every quantization primitive is the INSTALLED compressed-tensors 0.15.0.1 code called directly
(`calculate_qparams`, `fake_quantize`, `quantize`, `pack_to_int32`,
`PackedQuantizationCompressor.compress`), so those parts are pinned to the real thing.

PARITY UNPINNED for the GPTQ driver itself: the reference's tests hold no GPTQ vectors
 and upstream cannot be run here.  Pins in tests/test_oracle_cpu.py:
H against an fp64 X^T X, Hinv against fp64 linalg, the GPTQ objective (layer output error
must beat round-to-nearest), and - against the PUBLISHED algorithm - the artifact codes of this
driver equal those of an independently written fp64 column-by-column OBQ recursion (explicit H^-1,
one Gaussian elimination step per column; no Cholesky, no blocks) for sym / asym, with / without
act_order (`test_gptq_oracle_equals_the_published_obq_recursion`: 100 % measured, >= 99.9 % asserted).

One stated deviation: `torch.argsort(..., descending=True)` is made `stable=True` so that ties
on diag(H) order identically on CPU and GPU.
"""
import math
from typing import Optional

import torch
from compressed_tensors.quantization import (ActivationOrdering, QuantizationArgs, QuantizationStrategy,
                                             fake_quantize, preset_name_to_scheme)
from compressed_tensors.quantization.lifecycle.forward import quantize as ct_quantize
from compressed_tensors.quantization.utils import calculate_qparams

GPTQ_PRECISION = torch.float32


def scheme_weight_args(level: str) -> QuantizationArgs:
    """Preset scheme name -> weight QuantizationArgs (CT/quantization/quant_scheme.py:257-428)."""
    return preset_name_to_scheme(level, ["Linear"]).weights


# ---- §A.1 ------------------------------------------------------------------------------------
def make_empty_hessian(K: int) -> torch.Tensor:
    return torch.zeros((K, K), dtype=GPTQ_PRECISION)


def accumulate_hessian(inp: torch.Tensor, H: torch.Tensor, num_samples: int):
    """inp: [B, S, K] (or [S, K] = one sample).  Running-mean form; returns (H, num_samples)."""
    inp = inp.to(H.device)
    if inp.ndim == 2:
        inp = inp.unsqueeze(0)
    num_added = inp.shape[0]
    inp = inp.reshape((-1, inp.shape[-1])).t()
    H *= num_samples / (num_samples + num_added)
    num_samples += num_added
    inp = inp.to(dtype=GPTQ_PRECISION)
    inp = math.sqrt(2 / num_samples) * inp
    H += inp.matmul(inp.t())
    return H, num_samples


# ---- observer (MinMax, averaging_constant=1.0 => plain min/max) -------------------------------
def minmax_qparams(W: torch.Tensor, args: QuantizationArgs):
    """scale, zero_point of W [N, K] for CHANNEL ([N,1]) or GROUP ([N, K/gs]) strategy."""
    if args.strategy == QuantizationStrategy.CHANNEL:
        mn = torch.amin(W, dim=1, keepdim=True)
        mx = torch.amax(W, dim=1, keepdim=True)
        return calculate_qparams(mn, mx, args)
    if args.strategy == QuantizationStrategy.GROUP:
        gs = args.group_size
        N, K = W.shape
        Wg = W.reshape(N, K // gs, gs)
        return calculate_qparams(torch.amin(Wg, dim=2), torch.amax(Wg, dim=2), args)
    raise NotImplementedError(args.strategy)


def _channel_args(args: QuantizationArgs) -> QuantizationArgs:
    d = args.model_dump()
    d.update(strategy=QuantizationStrategy.CHANNEL, group_size=None, actorder=None)
    return QuantizationArgs(**d)


# ---- §A.2 - §A.5 -----------------------------------------------------------------------------
def quantize_weight(weight: torch.Tensor, H: torch.Tensor, args: QuantizationArgs, blocksize: int = 128,
                    percdamp: float = 0.01, return_hinv: bool = False, perm_override: Optional[torch.Tensor] = None):
    """Returns (loss, W_q [model dtype], scale [model dtype], zero_point int8, g_idx or None).
    `perm_override` (test hook, not in upstream): use this act_order permutation instead of argsort(diag H) -
    lets a test separate "the two fp32 diagonals order near-ties differently" from every other source of
    disagreement (tests/test_gptq_gpu.py::test_gptq_baseline_widths_own_hessian)."""
    final_dtype = weight.dtype
    W = weight.clone().to(GPTQ_PRECISION)
    H = H.clone()
    N, K = W.shape
    strategy = args.strategy
    actorder = args.actorder
    perm = None
    if strategy == QuantizationStrategy.GROUP:
        gs = args.group_size
        g_idx = torch.arange(K, dtype=torch.int) // gs
        if actorder == ActivationOrdering.GROUP:
            W, H, perm = _apply_activation_ordering(W, H, perm_override)
            scale, zero_point = minmax_qparams(W, args)
        elif actorder == ActivationOrdering.WEIGHT:
            scale, zero_point = minmax_qparams(W, args)
            W, H, perm = _apply_activation_ordering(W, H, perm_override)
            g_idx = g_idx[perm]
        else:
            scale, zero_point = minmax_qparams(W, args)
    else:
        g_idx = None
        scale, zero_point = minmax_qparams(W, args)

    dead = torch.diag(H) == 0
    H[dead, dead] = 1
    W[:, dead] = 0

    losses = torch.zeros(N)
    damp = percdamp * torch.mean(torch.diag(H))
    diag = torch.arange(K)
    H[diag, diag] += damp
    try:
        H = torch.linalg.cholesky(H)
        H = torch.cholesky_inverse(H)
        H = torch.linalg.cholesky(H, upper=True)
        Hinv = H
    except torch._C._LinAlgError:
        Hinv = H = torch.eye(K, dtype=H.dtype)

    ch_args = _channel_args(args)
    for i1 in range(0, K, blocksize):
        i2 = min(i1 + blocksize, K)
        count = i2 - i1
        W1 = W[:, i1:i2].clone()
        Q1 = torch.zeros_like(W1)
        Err1 = torch.zeros_like(W1)
        losses1 = torch.zeros_like(W1)
        Hinv1 = Hinv[i1:i2, i1:i2]
        for i in range(count):
            w = W1[:, i]
            d = Hinv1[i, i]
            q = w.clone()
            if strategy == QuantizationStrategy.CHANNEL:
                q = fake_quantize(q, scale[:, 0], zero_point[:, 0], args)
            else:
                column_idx = i1 + i
                group_index = int(g_idx[column_idx])
                if actorder != ActivationOrdering.WEIGHT and column_idx % args.group_size == 0:
                    grp = W[:, g_idx == group_index]
                    _s, _z = calculate_qparams(torch.amin(grp, dim=1, keepdim=True),
                                               torch.amax(grp, dim=1, keepdim=True), args)
                    scale[:, group_index] = _s[:, 0]
                    zero_point[:, group_index] = _z[:, 0]
                q = fake_quantize(q, scale[:, group_index], zero_point[:, group_index], ch_args)
            Q1[:, i] = q
            losses1[:, i] = (w - q) ** 2 / d**2
            err1 = (w - q) / d
            W1[:, i:] -= err1.unsqueeze(1).matmul(Hinv1[i, i:].unsqueeze(0))
            Err1[:, i] = err1
        W[:, i1:i2] = Q1
        losses += torch.sum(losses1, 1) / 2
        W[:, i2:] -= Err1.matmul(Hinv[i1:i2, i2:])

    has_gidx = False
    if strategy == QuantizationStrategy.GROUP:
        if actorder == ActivationOrdering.WEIGHT:
            invperm = torch.argsort(perm)
            W = W[:, invperm]
        elif actorder == ActivationOrdering.GROUP:
            invperm = torch.argsort(perm)
            W = W[:, invperm]
            g_idx = g_idx[invperm]
            has_gidx = True
    loss = torch.sum(losses).item()
    out = (loss, W.to(final_dtype), scale.to(final_dtype), zero_point.to(torch.int8),
           g_idx if has_gidx else None)
    if return_hinv:
        return out + (Hinv, perm)
    return out


def _apply_activation_ordering(W: torch.Tensor, H: torch.Tensor, perm_override: Optional[torch.Tensor] = None):
    perm = torch.argsort(torch.diag(H), descending=True, stable=True) if perm_override is None \
        else perm_override.to(torch.int64)
    return W[:, perm], H[perm][:, perm], perm


# ---- §A.6: what save_pretrained(save_compressed=True) stores --------------------------------
def compress_packed(Wq: torch.Tensor, scale: torch.Tensor, zero_point: Optional[torch.Tensor],
                    g_idx: Optional[torch.Tensor], args: QuantizationArgs):
    """int8 codes + int32 packing exactly as PackedQuantizationCompressor.compress does
    (CT/compressors/pack_quantized/base.py:35-77): codes are re-derived from the saved
    (model-dtype) weight and scale."""
    from compressed_tensors.compressors.pack_quantized.helpers import pack_to_int32
    codes = ct_quantize(x=Wq, scale=scale, zero_point=zero_point, g_idx=g_idx, args=args, dtype=torch.int8)
    packed = pack_to_int32(codes, args.num_bits)
    packed_zp = None
    if not args.symmetric and zero_point is not None:
        packed_zp = pack_to_int32(zero_point, args.num_bits, packed_dim=0).contiguous()
    return codes, packed, packed_zp


def compress_int8(Wq: torch.Tensor, scale: torch.Tensor, zero_point: Optional[torch.Tensor],
                  args: QuantizationArgs) -> torch.Tensor:
    """IntQuantizationCompressor (CT/compressors/naive_quantized/base.py:122-134): int8 `weight`."""
    return ct_quantize(x=Wq, scale=scale, zero_point=zero_point, args=args, dtype=torch.int8)


def rtn_quantize(weight: torch.Tensor, args: QuantizationArgs):
    """Round-to-nearest baseline (observer qparams + fake_quantize in the weight's dtype): what
    QuantizationMixin does without error feedback."""
    scale, zp = minmax_qparams(weight, args)
    return fake_quantize(weight, scale, zp, args), scale, zp


def layer_error(W: torch.Tensor, Wq: torch.Tensor, X: torch.Tensor) -> float:
    """||W X^T - Wq X^T||_F with X [T, K] (fp64 accumulate): the quantity GPTQ minimises."""
    D = (W.double() - Wq.double())
    return float(torch.linalg.norm(D @ X.double().t()))
