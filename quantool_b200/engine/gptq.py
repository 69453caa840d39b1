"""GPTQ for one Linear on one B200: Hessian accumulation, inverse-Hessian factor, blocked column
quantization, artifact codes.  Host glue over the C-ABI kernels; no arithmetic in Python.

Mirrors UPSTREAM llmcompressor `accumulate_hessian` / `quantize_weight`,
which the reference reaches through GPTQModifier at
ref/src/quantool/methods/llm_compressor/gptq/gptq.py:86.
"""
from dataclasses import dataclass
from typing import Optional

import torch

from .. import cabi
from .schemes import WeightArgs


class HessianAccumulator:
    """H = (2 / n_samples) * sum_b X_b^T X_b  (SURVEY §A.1).  Raw fp32 sums are kept on the device
    (upper-triangle tiles only) and scaled/mirrored once in `finalize`.  The diagonal is accumulated separately
    in fp32 round-to-nearest (the tensor-core accumulator truncates, which would reorder argsort(diag H));
    `sync_diagonal` writes it into H - call it before reducing H across ranks."""

    def __init__(self, K: int, device):
        self.K = K
        self.H = torch.zeros((K, K), dtype=torch.float32, device=device)
        self.diag = torch.zeros((K,), dtype=torch.float32, device=device)
        self._scratch = torch.empty((32 * K,), dtype=torch.float32, device=device)
        self.n_samples = 0
        self._final = False
        self._diag_synced = True

    def add(self, x: torch.Tensor, n_samples: Optional[int] = None, syrk_events=None) -> None:
        """x: [B, S, K] or [T, K]; n_samples defaults to B (a 2-D input counts as one sample).
        syrk_events = (start, end) CUDA events recorded around the tensor-core SYRK launch alone."""
        assert not self._final
        if n_samples is None:
            n_samples = x.shape[0] if x.dim() == 3 else 1
        x2 = x.reshape(-1, x.shape[-1]).contiguous()      # bf16 or fp16, as the model produces it (no cast)
        if syrk_events is not None:
            syrk_events[0].record()
        cabi.hessian_accumulate(x2, self.H)
        if syrk_events is not None:
            syrk_events[1].record()
        cabi.hessian_diag_accumulate(x2, self.diag, self._scratch)
        self._diag_synced = False
        self.n_samples += int(n_samples)

    def reset(self) -> None:
        """Reuse the buffers for the next decoder layer (stream-ordered after the previous consumers)."""
        self.H.zero_()
        self.diag.zero_()
        self.n_samples = 0
        self._final = False
        self._diag_synced = True

    def sync_diagonal(self) -> None:
        if not self._diag_synced:
            cabi.hessian_set_diagonal(self.H, self.diag)
            self._diag_synced = True

    def finalize(self, total_samples: Optional[int] = None) -> torch.Tensor:
        n = total_samples if total_samples is not None else self.n_samples
        if not self._final:
            self.sync_diagonal()
            cabi.hessian_finalize(self.H, 2.0 / max(n, 1))
            self._final = True
        return self.H


@dataclass
class GPTQResult:
    weight: torch.Tensor             # fake-quantized weight, model dtype, original column order
    scale: torch.Tensor              # [N, G] model dtype
    zero_point: torch.Tensor         # [N, G] int8
    g_idx: Optional[torch.Tensor]    # [K] int32 (actorder="group" only)
    losses: torch.Tensor             # [N] fp32 per-row GPTQ loss (device)
    info: torch.Tensor               # device int32: 0 ok, else failing Cholesky pivot (identity fallback used)
    perm: Optional[torch.Tensor] = None   # [K] int32 act_order permutation that was applied (None without act_order)


def quantize_linear(weight: torch.Tensor, H: torch.Tensor, args: WeightArgs, blocksize: int = 128,
                    percdamp: float = 0.01, check_info: bool = True, tensor_core_lazy: bool = True,
                    tensor_core_chain: Optional[bool] = None) -> GPTQResult:
    """weight [N, K] (CUDA, model dtype), H finalized [K, K] fp32.  H is not modified.
    tensor_core_chain: None = tensor-core inverse-Hessian chain whenever K allows (K % 256 == 0)."""
    if blocksize != 128:
        raise ValueError("the sm_100a GPTQ kernel is specialised for block_size=128 (upstream default)")
    N, K = weight.shape
    dev = weight.device
    final_dtype = weight.dtype
    weight = weight.contiguous()
    perm = inv_perm = None
    if args.actorder in ("group", "weight"):
        perm = torch.argsort(torch.diagonal(H), descending=True, stable=True).to(torch.int32)
        inv_perm = torch.argsort(perm).to(torch.int32)

    g_idx_perm = None
    if args.strategy == "channel":
        mode = cabi.GPTQ_MODE_CHANNEL
        scale, zp = cabi.minmax_qparams(weight.float(), 0, args.num_bits, args.symmetric)
        gs = 0
    else:
        gs = args.group_size
        if K % gs:
            raise ValueError(f"tensor column shape must be divisble by the given group_size {gs} but got {K}")
        if args.actorder == "weight":
            mode = cabi.GPTQ_MODE_STATIC_GIDX
            scale, zp = cabi.minmax_qparams(weight.float(), gs, args.num_bits, args.symmetric)
            g_idx_perm = (torch.arange(K, device=dev, dtype=torch.int32) // gs)[perm.long()].contiguous()
        else:
            mode = cabi.GPTQ_MODE_GROUP_REFIT
            scale = torch.empty((N, K // gs), dtype=torch.float32, device=dev)
            zp = torch.empty((N, K // gs), dtype=torch.float32, device=dev)

    Hf, dead = cabi.gptq_prepare_hessian(H, perm, percdamp)
    U, info = cabi.gptq_hinv_factor(Hf, tensor_core=tensor_core_chain)
    if check_info and int(info.item()) != 0:
        cabi.set_identity(U)   # upstream: on LinAlgError, Hinv = eye(K)
    wp = cabi.gptq_permute_in(weight, perm, dead)
    U_split = cabi.split_tf32_transpose(U) if (tensor_core_lazy and K > 128) else None
    losses = cabi.gptq_quantize_weight(wp, U, scale, zp, g_idx_perm, gs, args.num_bits, args.symmetric, mode,
                                       U_split=U_split)
    wq = cabi.gptq_permute_out(wp, inv_perm, final_dtype)
    g_idx = None
    if args.strategy == "group" and args.actorder == "group":
        g_idx = (torch.arange(K, device=dev, dtype=torch.int32) // gs)[inv_perm.long()].contiguous()
    return GPTQResult(wq, scale.to(final_dtype), zp.to(torch.int8), g_idx, losses, info, perm)


def compress_linear(wq: torch.Tensor, scale: torch.Tensor, zero_point: Optional[torch.Tensor],
                    g_idx: Optional[torch.Tensor], args: WeightArgs, fmt: str = "pack-quantized"):
    """Artifact tensors of one Linear as compressed-tensors stores them:
    codes are re-derived from the saved model-dtype weight and scale."""
    N, K = wq.shape
    gs = args.group_size if args.strategy == "group" else 0
    zpf = zero_point.float() if zero_point is not None else None
    codes, _ = cabi.quantize_codes(wq.contiguous(), scale.contiguous(), zpf, g_idx, gs, args.num_bits)
    out = {}
    if fmt == "pack-quantized":
        out["weight_packed"] = cabi.pack_int32(codes, args.num_bits)
        out["weight_scale"] = scale
        out["weight_shape"] = torch.tensor([N, K], dtype=torch.int64)
        if not args.symmetric:
            # pack_to_int32(zero_point, num_bits, packed_dim=0): [ceil(N/pf), G]
            out["weight_zero_point"] = cabi.pack_int32(zero_point.t().contiguous(), args.num_bits).t().contiguous()
        if g_idx is not None:
            out["weight_g_idx"] = g_idx
    elif fmt == "int-quantized":
        out["weight"] = codes
        out["weight_scale"] = scale
        if not args.symmetric:
            out["weight_zero_point"] = zero_point
    else:
        raise ValueError(fmt)
    return out, codes
