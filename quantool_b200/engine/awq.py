"""AWQ for one decoder layer: activation means, w_mean, n_grid scale search, smoothing, RTN qparams.

Mirrors UPSTREAM llmcompressor AWQModifier, built by the reference at
ref/src/quantool/methods/llm_compressor/awq/awq.py:81.  The per-candidate weight
(scale -> quantize -> dequantize -> unscale) is ONE fused kernel pass (qt_awq_scale_qdq).  For a parent that
is a single Linear (down_proj; o_proj where v -> o applies) the reconstruction loss is evaluated in Gram form,
tr(D G D^T) with G = X^T X, by one tcgen05 GEMM with a reducing epilogue per grid point
(qt_awq_scale_qdq_delta + qt_awq_gram_loss): no forward, no materialised output.  The self_attn / mlp parents run
their forwards through torch (cuBLAS / SDPA) and the loss is reduced on the device (qt_sq_err_sum).  20 losses
come back to the host once per mapping.
"""
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .. import cabi
from . import llama
from .schemes import WeightArgs

N_GRID = 20
USE_GRAM_LOSS = True     # single-Linear parents: tr(D G D^T) on the tensor cores instead of 20 forwards (see search_mapping)


@dataclass
class Mapping:
    smooth: str              # weight key of the smooth layer
    balance: List[str]       # Linear names
    inp: str                 # captured input feeding the balance layers
    parent: str              # "self_attn" | "mlp" | "linear"


def llama_mappings(shape: llama.LlamaShape) -> List[Mapping]:
    """Default Llama mappings: v_proj -> o_proj is skipped when the shapes differ (GQA)."""
    m = [Mapping("input_layernorm", ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"], "attn_in",
                 "self_attn")]
    if shape.kv_dim == shape.q_dim and shape.q_dim == shape.hidden_size:
        m.append(Mapping("self_attn.v_proj", ["self_attn.o_proj"], "o_in", "linear"))
    m.append(Mapping("post_attention_layernorm", ["mlp.gate_proj", "mlp.up_proj"], "mlp_in", "mlp"))
    m.append(Mapping("mlp.up_proj", ["mlp.down_proj"], "down_in", "linear"))
    return m


def candidate_scales(x_mean: torch.Tensor, w_mean: torch.Tensor, ratio: float, duo_scaling: bool = True):
    """[K]-vector bookkeeping of the grid search (negligible work; torch ops on the device)."""
    if duo_scaling:
        scales = (x_mean.pow(ratio) / (w_mean.pow(1 - ratio) + 1e-4)).clamp(min=1e-4)
    else:
        scales = x_mean.pow(ratio).clamp(min=1e-4).view(-1)
    scales = scales / (scales.max() * scales.min()).sqrt()
    scales[torch.isinf(scales)] = 1
    scales[torch.isnan(scales)] = 1
    return scales.to(torch.float32).contiguous()


def _parent_forward(shape, kind: str, w: Dict[str, torch.Tensor], lin: Optional[str], x: torch.Tensor, cos, sin):
    if kind == "self_attn":
        return llama.attention_forward(shape, w, x, cos, sin)
    if kind == "mlp":
        return llama.mlp_forward(w, x)
    return F.linear(x, w[f"{lin}.weight"])


def search_mapping(shape, w: Dict[str, torch.Tensor], mp: Mapping, x_all: torch.Tensor, x_mean: torch.Tensor,
                   args: WeightArgs, chunks, dist=None, n_grid: int = N_GRID, duo_scaling: bool = True):
    """x_all: [T_local, K] cached inputs of the parent; chunks: [(row0, B, S, cos, sin)] views into it (samples of
    one chunk share a length).  Returns (best_scales, best_ratio, losses)."""
    dev = x_all.device
    gs = args.group_size if args.strategy == "group" else 0
    bw = [w[f"{b}.weight"] for b in mp.balance]
    w_mean = cabi.awq_wmean(bw, gs)
    lin = mp.balance[0] if mp.parent == "linear" else None

    def xin(c):
        r0, B, S, _, _ = c
        return x_all[r0: r0 + B * S].view(B, S, -1)

    losses_dev = torch.zeros((n_grid,), dtype=torch.float64, device=dev)
    cand = []
    gram = (mp.parent == "linear" and USE_GRAM_LOSS and len(bw) == 1 and x_all.dtype in (torch.bfloat16, torch.float16)
            and cabi.awq_gram_ok(x_all.shape[1], gs))
    if gram:
        # Single-Linear parent: ||X D^T||^2 = tr(D G D^T) with G = X^T X built once (tcgen05 SYRK, summed over the
        # ranks) - per grid point one [N, K] x [K, K] GEMM with a reducing epilogue instead of a [T, K] x [K, N]
        # forward, a materialised output and a second pass over it (T = 64 K tokens vs K = 14 K: 4.6x fewer FLOP).
        K = x_all.shape[1]
        H = torch.zeros((K, K), dtype=torch.float32, device=dev)
        cabi.hessian_accumulate(x_all, H)
        if dist is not None and dist.on:
            dist.all_reduce_hessian(H)
        cabi.hessian_finalize(H, 1.0)
        G = H.to(torch.bfloat16)
        del H
        # G is the same on every rank, and tr(D G D^T) is a sum over the rows of D: each rank takes a row slice of
        # the weight (the losses are summed below) - otherwise all ranks would repeat the same 20 products
        wl = bw[0]
        if dist is not None and dist.on:
            from .pipeline import row_split
            sizes = row_split(wl.shape[0], dist.world, 16)
            r0 = sum(sizes[: dist.rank])
            wl = wl[r0: r0 + sizes[dist.rank]]
        if wl.shape[0] > 0:
            d16 = torch.empty(wl.shape, dtype=torch.bfloat16, device=dev)
            d32 = torch.empty(wl.shape, dtype=torch.float32, device=dev)
        for gi in range(n_grid):
            s = candidate_scales(x_mean, w_mean, gi / n_grid, duo_scaling)
            cand.append(s)
            if wl.shape[0] > 0:
                cabi.awq_scale_qdq_delta(wl, s, gs, args.num_bits, args.symmetric, d16, d32)
                cabi.awq_gram_loss(d16, d32, G, losses_dev[gi:gi + 1])
        numel = x_all.shape[0] * bw[0].shape[0]
        del G
    else:
        ref_out = [_parent_forward(shape, mp.parent, w, lin, xin(c), c[3], c[4]) for c in chunks]
        numel = sum(o.numel() for o in ref_out)
        patched = dict(w)
        bufs = [torch.empty_like(t) for t in bw]
        for gi in range(n_grid):
            ratio = gi / n_grid
            s = candidate_scales(x_mean, w_mean, ratio, duo_scaling)
            cand.append(s)
            for name, t, buf in zip(mp.balance, bw, bufs):
                cabi.awq_scale_qdq(t, s, gs, args.num_bits, args.symmetric, out=buf)
                patched[f"{name}.weight"] = buf
            for c, ro in zip(chunks, ref_out):
                out = _parent_forward(shape, mp.parent, patched, lin, xin(c), c[3], c[4])
                cabi.sq_err_sum(ro, out, losses_dev[gi:gi + 1])
    tot = torch.tensor([float(numel)], dtype=torch.float64, device=dev)
    if dist is not None and dist.on:
        dist.all_reduce_sum(losses_dev)   # forward form: token shards; Gram form: row slices of D against the summed G
        dist.all_reduce_sum(tot)
    losses = (losses_dev / tot).cpu().tolist()
    best, best_err = -1, float("inf")
    for gi, l in enumerate(losses):
        if l < best_err:          # strict <: first minimum wins
            best_err, best = l, gi
    if best < 0:
        raise RuntimeError("AWQ scale search produced no finite loss")
    return cand[best], best / n_grid, losses


def apply_mapping(w: Dict[str, torch.Tensor], mp: Mapping, s: torch.Tensor) -> None:
    for b in mp.balance:
        cabi.scale_matrix_(w[f"{b}.weight"], s)                       # W *= s[None, :]
    sw = w[f"{mp.smooth}.weight"]
    if sw.dim() == 1:
        cabi.scale_matrix_(sw, s, divide=True)                        # norm weight /= s
    else:
        tail = sw[sw.shape[0] - s.numel():]
        cabi.scale_matrix_(tail, s, divide=True, by_row=True)         # weight[-len(s):] /= s[:, None]


def awq_layer(shape: llama.LlamaShape, w: Dict[str, torch.Tensor], calib, cos=None, sin=None, args: WeightArgs = None,
              chunk_samples: int = 32, dist=None, n_grid: int = N_GRID, duo_scaling: bool = True):
    """Calibrate + smooth one decoder layer in place.  `calib`: a pipeline.CalibSet, or hidden states
    [n, seq, hidden] with `cos`/`sin`.  Returns {smooth-name: (scales, ratio, losses)}."""
    from .pipeline import CalibSet
    if isinstance(calib, torch.Tensor):
        calib = CalibSet.from_hidden(calib, cos, sin)
    dev = next(iter(w.values())).device
    dtype = w["input_layernorm.weight"].dtype
    dims = shape.input_dims()
    T = calib.tokens_local
    cache = {n: torch.empty((T, k), dtype=dtype, device=dev) for n, k in dims.items()}
    chunks, row0 = [], 0
    for gi, a, b in calib.chunks(max(1, chunk_samples) * max(calib.max_len, 1)):
        hb = calib.groups[gi][a:b]
        c, s_ = calib.ropes[gi]
        llama.layer_forward(shape, w, hb, c, s_, capture=cache, row0=row0, stop_after="down_in")
        chunks.append((row0, hb.shape[0], hb.shape[1], c, s_))
        row0 += hb.shape[0] * hb.shape[1]
    out = {}
    for mp in llama_mappings(shape):
        x_all = cache[mp.inp]
        acc = torch.zeros((x_all.shape[1],), dtype=torch.float32, device=dev)
        cabi.channel_abs_sum(x_all, acc)
        cnt = torch.tensor([float(T)], dtype=torch.float32, device=dev)
        if dist is not None and dist.on:
            dist.all_reduce_sum(acc)
            dist.all_reduce_sum(cnt)
        x_mean = acc / cnt
        s, ratio, losses = search_mapping(shape, w, mp, x_all, x_mean, args, chunks, dist, n_grid, duo_scaling)
        apply_mapping(w, mp, s)
        out[mp.smooth] = (s, ratio, losses)
    return out


def rtn_qparams(weight: torch.Tensor, args: WeightArgs):
    """Final weight qparams: minmax observer evaluated in the weight's dtype."""
    gs = args.group_size if args.strategy == "group" else 0
    scale, zp = cabi.minmax_qparams(weight.contiguous(), gs, args.num_bits, args.symmetric)
    return scale.to(weight.dtype), zp.to(torch.int8)
