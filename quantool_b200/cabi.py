"""ctypes binding of the C-ABI library (include/quantool_b200.h).

Host glue only: unwraps torch tensors into device pointers + sizes and passes torch's current
CUDA stream.  There is NO CPU fallback: if the library is missing or the tensor is not on a
CUDA device the call raises.
"""
import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libquantool_b200.so")

QT_F32, QT_F16, QT_BF16 = 0, 1, 2
_DT = {torch.float32: QT_F32, torch.float16: QT_F16, torch.bfloat16: QT_BF16}

GGML = {"Q4_0": 2, "Q4_1": 3, "Q5_0": 6, "Q5_1": 7, "Q8_0": 8, "Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14, "IQ4_NL": 20}

_i64, _i32, _vp, _f32 = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_float

# name -> argtypes; every function returns int unless listed in _RESTYPE
_SIGS = {
    "qt_abi_version": [],
    "qt_device_sm_count": [],
    "qt_launch_count": [],
    "qt_gguf_block_elems": [_i32],
    "qt_gguf_block_bytes": [_i32],
    "qt_gguf_quantize": [_i32, _vp, _i32, _i32, _i64, _i64, _vp, _vp],
    "qt_gguf_batch_table_bytes": [_i32],
    "qt_gguf_quantize_batch": [_i32, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp],
    "qt_gguf_dequantize": [_i32, _vp, _i64, _i64, _vp, _vp],
    "qt_minmax_qparams": [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp],
    "qt_quantize_codes": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp],
    "qt_pack_int32": [_vp, _i32, _i32, _i32, _vp, _vp],
    "qt_unpack_int32": [_vp, _i32, _i32, _i32, _vp, _vp],
    "qt_hessian_set_splits": [_i32],
    "qt_hessian_reserve_sms": [_i32],
    "qt_hessian_accumulate": [_vp, _i32, _i64, _i32, _vp, _vp],
    "qt_hessian_finalize": [_vp, _i32, _f32, _vp],
    "qt_hessian_diag_accumulate": [_vp, _i32, _i64, _i32, _vp, _vp, _vp],
    "qt_hessian_set_diagonal": [_vp, _i32, _vp, _vp],
    "qt_tri_packed_elems": [_i32, _i32],
    "qt_tri_pack": [_vp, _i32, _i32, _vp, _vp],
    "qt_tri_unpack": [_vp, _i32, _i32, _i32, _vp, _vp],
    "qt_gptq_prepare_hessian": [_vp, _vp, _i32, _f32, _vp, _vp, _vp, _vp],
    "qt_gptq_hinv_factor": [_vp, _vp, _vp, _i32, _vp, _vp],
    "qt_gptq_hinv_factor_tc": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp],
    "qt_set_identity": [_vp, _i32, _vp],
    "qt_sgemm": [_i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _i32, _i32, _vp],
    "qt_sgemm_ex": [_i32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _i32, _i32, _i32, _i32, _vp],
    "qt_tri_chain_block": [_vp, _vp, _vp, _i32, _i32, _vp, _vp],
    "qt_flip_upper": [_vp, _vp, _i32, _vp],
    "qt_gptq_permute_in": [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp],
    "qt_gptq_permute_out": [_vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "qt_fill_f32": [_vp, _i32, _f32, _vp],
    "qt_channel_minmax": [_vp, _i32, _i64, _i32, _vp, _vp, _vp],
    "qt_channel_abs_sum": [_vp, _i32, _i64, _i32, _vp, _vp],
    "qt_smooth_scales": [_vp, _vp, _vp, _vp, _f32, _f32, _i32, _vp, _i32, _vp],
    "qt_scale_matrix": [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _vp],
    "qt_gemm_tf32x3": [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "qt_rms_norm": [_vp, _vp, _vp, _i32, _i64, _i32, _f32, _vp],
    "qt_rope_inplace": [_vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _vp],
    "qt_silu_mul": [_vp, _vp, _vp, _i32, _i64, _vp],
    "qt_awq_wmean_accumulate": [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp],
    "qt_awq_scale_qdq": [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp],
    "qt_sq_err_sum": [_vp, _vp, _i32, _i64, _vp, _vp],
    "qt_awq_scale_qdq_delta": [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp],
    "qt_awq_gram_loss": [_vp, _vp, _vp, _i32, _i32, _vp, _vp],
    "qt_gptq_quantize_weight": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "qt_gptq_set_block_kernel": [_i32],
    "qt_split_tf32": [_vp, _vp, _vp, _i64, _vp],
    "qt_split_tf32_transpose": [_vp, _vp, _vp, _i32, _vp],
}
_RESTYPE = {"qt_last_error": ctypes.c_char_p, "qt_launch_count": ctypes.c_ulonglong,
            "qt_gguf_batch_table_bytes": ctypes.c_int64, "qt_tri_packed_elems": ctypes.c_int64}

_lib: Optional[ctypes.CDLL] = None


class QtError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the C-ABI library; fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QtError(f"{LIB_PATH} not found: build it with `python -m quantool_b200.csrc.build` "
                          "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.qt_last_error.restype = ctypes.c_char_p
        L.qt_last_error.argtypes = []
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, ctypes.c_int)
        _lib = L
    return _lib


def launch_count() -> int:
    return int(lib().qt_launch_count())


def exported_symbols():
    return ["qt_last_error"] + list(_SIGS)


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().qt_last_error()
        raise QtError(f"{what} failed: rc={rc} ({msg.decode() if msg else ''})")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise QtError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise QtError(f"{name} must be contiguous")
    return t


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


# ---------------------------------------------------------------------------------------
# GGUF
# ---------------------------------------------------------------------------------------
def gguf_block_elems(t: str) -> int:
    return lib().qt_gguf_block_elems(GGML[t])


def gguf_block_bytes(t: str) -> int:
    return lib().qt_gguf_block_bytes(GGML[t])


def gguf_quantize(x: torch.Tensor, qtype: str, round_via_f16: bool = True,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [nrows, ncols] (f32/f16/bf16, CUDA) -> uint8 [nrows, ncols/be*bb] packed blocks."""
    _dev(x, "x")
    assert x.dim() == 2
    nrows, ncols = x.shape
    be, bb = gguf_block_elems(qtype), gguf_block_bytes(qtype)
    if ncols % be:
        raise QtError(f"ncols={ncols} is not a multiple of the {qtype} block size {be}")
    if out is None:
        out = torch.empty((nrows, ncols // be * bb), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib().qt_gguf_quantize(GGML[qtype], _p(x), _DT[x.dtype], int(round_via_f16), nrows, ncols,
                                      _p(out), _stream()), "qt_gguf_quantize")
    return out


def gguf_quantize_batch(xs, qtype: str, round_via_f16: bool = True, outs=None):
    """Many 2-D tensors (same dtype, same device) -> packed blocks of one type in ONE kernel launch."""
    if not xs:
        return []
    be, bb = gguf_block_elems(qtype), gguf_block_bytes(qtype)
    dev, dt = xs[0].device, xs[0].dtype
    for x in xs:
        _dev(x, "x")
        if x.dim() != 2 or x.device != dev or x.dtype != dt:
            raise QtError("gguf_quantize_batch takes 2-D tensors of one dtype on one device")
        if x.shape[1] % be:
            raise QtError(f"ncols={x.shape[1]} is not a multiple of the {qtype} block size {be}")
    n = len(xs)
    sizes = [x.numel() // be * bb for x in xs]
    if outs is None:
        # one allocation for the whole group; the per-tensor views are made after the launch, while it runs
        offs, total = [], 0
        for sz in sizes:
            offs.append(total)
            total += (sz + 15) & ~15
        buf = torch.empty((max(total, 1),), dtype=torch.uint8, device=dev)
        base = buf.data_ptr()
        dst = (ctypes.c_void_p * n)(*[base + o for o in offs])
    else:
        buf = None
        dst = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
    src = (ctypes.c_void_p * n)(*[x.data_ptr() for x in xs])
    nel = (ctypes.c_int64 * n)(*[x.numel() for x in xs])
    table = torch.empty((int(lib().qt_gguf_batch_table_bytes(n)),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _check(lib().qt_gguf_quantize_batch(GGML[qtype], n, src, nel, dst, _DT[dt], int(round_via_f16), _p(table),
                                            _stream()), "qt_gguf_quantize_batch")
    if buf is not None:
        outs = [buf[o:o + sz].view(x.shape[0], -1) if sz else torch.empty((x.shape[0], x.shape[1] // be * bb),
                                                                           dtype=torch.uint8, device=dev)
                for x, o, sz in zip(xs, offs, sizes)]
    return outs


def gguf_quantize_many(plan, round_via_f16: bool = True):
    """plan: list of (tensor, qtype).  One launch per (qtype, dtype, device) group; results in plan order."""
    groups = {}
    for i, (x, qt) in enumerate(plan):
        groups.setdefault((qt, x.dtype, x.device), []).append(i)
    out = [None] * len(plan)
    for (qt, _, _), idx in groups.items():
        ys = gguf_quantize_batch([plan[i][0] for i in idx], qt, round_via_f16)
        for i, y in zip(idx, ys):
            out[i] = y
    return out


def gguf_dequantize(y: torch.Tensor, qtype: str, ncols: int) -> torch.Tensor:
    _dev(y, "y")
    nrows = y.shape[0]
    out = torch.empty((nrows, ncols), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _check(lib().qt_gguf_dequantize(GGML[qtype], _p(y), nrows, ncols, _p(out), _stream()),
               "qt_gguf_dequantize")
    return out


# ---------------------------------------------------------------------------------------
# compressed-tensors primitives (observer, codes, int32 packing)
# ---------------------------------------------------------------------------------------
def minmax_qparams(w: torch.Tensor, group_size: int, num_bits: int, symmetric: bool):
    """MinMax observer + calculate_qparams, evaluated in w.dtype.  Returns (scale, zp) fp32 [N, G];
    zp holds integer values.  group_size <= 0 means one scale per row (CHANNEL)."""
    _dev(w, "w")
    N, K = w.shape
    G = K // group_size if group_size > 0 else 1
    scale = torch.empty((N, G), dtype=torch.float32, device=w.device)
    zp = torch.empty((N, G), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _check(lib().qt_minmax_qparams(_p(w), _DT[w.dtype], N, K, max(group_size, 0), num_bits, int(symmetric),
                                       _p(scale), _p(zp), _stream()), "qt_minmax_qparams")
    return scale, zp


def quantize_codes(w: torch.Tensor, scale: torch.Tensor, zp: Optional[torch.Tensor], g_idx: Optional[torch.Tensor],
                   group_size: int, num_bits: int, want_codes: bool = True, want_dq: bool = False):
    """CT `quantize` (-> int8 codes) and optionally `dequantize`, in w.dtype arithmetic.
    scale must have w's dtype ([N, G]); zp fp32 integers [N, G] or None; g_idx int32 [K] or None."""
    _dev(w, "w"); _dev(scale, "scale")
    if scale.dtype != w.dtype:
        raise QtError("scale must have the weight's dtype (the artifact stores weight_scale in model dtype)")
    N, K = w.shape
    G = scale.shape[1]
    codes = torch.empty((N, K), dtype=torch.int8, device=w.device) if want_codes else None
    dq = torch.empty_like(w) if want_dq else None
    if zp is not None:
        zp = _dev(zp.to(torch.float32), "zp")
    if g_idx is not None:
        g_idx = _dev(g_idx.to(torch.int32), "g_idx")
    with torch.cuda.device(w.device):
        _check(lib().qt_quantize_codes(_p(w), _p(scale), _p(zp), _p(g_idx), _DT[w.dtype], N, K, G,
                                       max(group_size, 0), num_bits, _p(codes), _p(dq), _stream()),
               "qt_quantize_codes")
    return codes, dq


def pack_int32(codes: torch.Tensor, num_bits: int) -> torch.Tensor:
    _dev(codes, "codes")
    assert codes.dtype == torch.int8
    N, K = codes.shape
    pf = 32 // num_bits
    out = torch.empty((N, (K + pf - 1) // pf), dtype=torch.int32, device=codes.device)
    with torch.cuda.device(codes.device):
        _check(lib().qt_pack_int32(_p(codes), N, K, num_bits, _p(out), _stream()), "qt_pack_int32")
    return out


def unpack_int32(packed: torch.Tensor, num_bits: int, K: int) -> torch.Tensor:
    _dev(packed, "packed")
    N = packed.shape[0]
    out = torch.empty((N, K), dtype=torch.int8, device=packed.device)
    with torch.cuda.device(packed.device):
        _check(lib().qt_unpack_int32(_p(packed), N, K, num_bits, _p(out), _stream()), "qt_unpack_int32")
    return out


# ---------------------------------------------------------------------------------------
# GPTQ
# ---------------------------------------------------------------------------------------
def hessian_accumulate(x: torch.Tensor, H: torch.Tensor) -> None:
    """H (fp32 [K,K], raw sums, upper tiles) += x^T x.  x: [T, K] bf16 or fp16 (fed to the tensor cores as is)."""
    _dev(x, "x"); _dev(H, "H")
    if x.dtype not in (torch.bfloat16, torch.float16):
        raise QtError(f"hessian_accumulate takes bf16 or fp16 activations, got {x.dtype} (load the model in its "
                      "16-bit dtype; there is no fp32 tensor-core Hessian and no silent down-cast)")
    T, K = x.shape
    assert H.shape == (K, K) and H.dtype == torch.float32
    with torch.cuda.device(x.device):
        _check(lib().qt_hessian_accumulate(_p(x), _DT[x.dtype], T, K, _p(H), _stream()), "qt_hessian_accumulate")


def hessian_diag_accumulate(x: torch.Tensor, diag: torch.Tensor, scratch: torch.Tensor) -> None:
    """diag (fp32 [K], raw sums) += column sums of squares of x [T, K] bf16/fp16, fp32 round-to-nearest."""
    _dev(x, "x"); _dev(diag, "diag"); _dev(scratch, "scratch")
    T, K = x.shape
    assert x.dtype in (torch.bfloat16, torch.float16) and diag.dtype == torch.float32 and scratch.numel() >= 32 * K
    with torch.cuda.device(x.device):
        _check(lib().qt_hessian_diag_accumulate(_p(x), _DT[x.dtype], T, K, _p(diag), _p(scratch), _stream()),
               "qt_hessian_diag_accumulate")


def hessian_set_diagonal(H: torch.Tensor, diag: torch.Tensor) -> None:
    with torch.cuda.device(H.device):
        _check(lib().qt_hessian_set_diagonal(_p(_dev(H, "H")), H.shape[0], _p(_dev(diag, "diag")), _stream()),
               "qt_hessian_set_diagonal")


def hessian_finalize(H: torch.Tensor, factor: float) -> None:
    """Scale the accumulated upper triangle by `factor` (2 / n_samples) and mirror it."""
    _dev(H, "H")
    with torch.cuda.device(H.device):
        _check(lib().qt_hessian_finalize(_p(H), H.shape[0], float(factor), _stream()), "qt_hessian_finalize")


def tri_packed_elems(K: int, align: int) -> int:
    n = int(lib().qt_tri_packed_elems(K, align))
    if n < 0:
        raise QtError("qt_tri_packed_elems: invalid arguments")
    return n


def tri_pack(M: torch.Tensor, align: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Packed upper block-triangle of a square fp32 matrix (align 256: raw Hessian sums, 128: the factor U)."""
    _dev(M, "M")
    K = M.shape[0]
    n = tri_packed_elems(K, align)
    if out is None:
        out = torch.empty((n,), dtype=torch.float32, device=M.device)
    assert out.numel() >= n and out.dtype == torch.float32
    with torch.cuda.device(M.device):
        _check(lib().qt_tri_pack(_p(M), K, align, _p(out), _stream()), "qt_tri_pack")
    return out[:n]


def tri_unpack(packed: torch.Tensor, M: torch.Tensor, align: int, zero_below: bool = False) -> torch.Tensor:
    _dev(M, "M")
    _dev(packed, "packed")
    with torch.cuda.device(M.device):
        _check(lib().qt_tri_unpack(_p(packed), M.shape[0], align, int(zero_below), _p(M), _stream()), "qt_tri_unpack")
    return M


def gptq_prepare_hessian(H: torch.Tensor, perm: Optional[torch.Tensor], percdamp: float):
    """Returns (Hf, dead): the damped Hessian, act_order-permuted and index-reversed (lower
    triangle), and the dead-column mask of the ORIGINAL column order."""
    _dev(H, "H")
    K = H.shape[0]
    Hf = torch.empty_like(H)
    dead = torch.empty((K,), dtype=torch.uint8, device=H.device)
    damp = torch.empty((1,), dtype=torch.float32, device=H.device)
    if perm is not None:
        perm = _dev(perm.to(torch.int32), "perm")
    with torch.cuda.device(H.device):
        _check(lib().qt_gptq_prepare_hessian(_p(H), _p(perm), K, float(percdamp), _p(Hf), _p(dead), _p(damp),
                                             _stream()), "qt_gptq_prepare_hessian")
    return Hf, dead


def hinv_tensor_core_ok(K: int) -> bool:
    """The tensor-core chain needs K to be a multiple of its 256-wide tiles; other K run the FFMA chain."""
    return K % 256 == 0


def gptq_hinv_factor(Hf: torch.Tensor, scratch_x: Optional[torch.Tensor] = None,
                     scratch_w: Optional[torch.Tensor] = None, tensor_core: Optional[bool] = None,
                     workspace=None, info: Optional[torch.Tensor] = None, stages: int = 3):
    """In place: Hf (flipped damped H) -> U = cholesky(H^-1, upper).  Returns (U, info) where
    info is a device int32 (0 = ok, else 1-based failing pivot; caller decides when to read it).
    tensor_core (default: whenever K allows) runs the big products as 3xTF32 tcgen05 GEMMs and needs four
    more K x K fp32 workspaces (`workspace`, allocated here when not given)."""
    _dev(Hf, "Hf")
    K = Hf.shape[0]
    X = scratch_x if scratch_x is not None else torch.empty_like(Hf)
    W = scratch_w if scratch_w is not None else torch.empty_like(Hf)
    info = torch.zeros((1,), dtype=torch.int32, device=Hf.device) if info is None else info
    tc = hinv_tensor_core_ok(K) if tensor_core is None else tensor_core
    with torch.cuda.device(Hf.device):
        if tc:
            ws = workspace if workspace is not None else [torch.empty_like(Hf) for _ in range(4)]
            _check(lib().qt_gptq_hinv_factor_tc(_p(Hf), _p(X), _p(W), _p(ws[0]), _p(ws[1]), _p(ws[2]), _p(ws[3]), K,
                                                _p(info), int(stages), _stream()), "qt_gptq_hinv_factor_tc")
        else:
            _check(lib().qt_gptq_hinv_factor(_p(Hf), _p(X), _p(W), K, _p(info), _stream()), "qt_gptq_hinv_factor")
    return Hf, info


def set_identity(U: torch.Tensor) -> None:
    with torch.cuda.device(U.device):
        _check(lib().qt_set_identity(_p(_dev(U, "U")), U.shape[0], _stream()), "qt_set_identity")


def sgemm(A, B, C, alpha=1.0, beta=0.0, b_is_nk=False, lower_tiles_only=False, a_lower_tri=False,
          b_lower_tri=False, tri_row_offset=0):
    """C = alpha * A @ (B^T if b_is_nk else B) + beta * C on fp32 views with unit inner stride (row
    strides are the leading dimensions, so sub-blocks of larger buffers work in place)."""
    M, Kd = A.shape
    N = C.shape[1]
    assert A.stride(1) == 1 and B.stride(1) == 1 and C.stride(1) == 1
    with torch.cuda.device(A.device):
        _check(lib().qt_sgemm_ex(int(b_is_nk), _p(A), _p(B), _p(C), M, N, Kd, A.stride(0), B.stride(0), C.stride(0),
                                 float(alpha), float(beta), int(lower_tiles_only), int(a_lower_tri), int(b_lower_tri),
                                 int(tri_row_offset), _stream()), "qt_sgemm_ex")
    return C


def tri_chain_block(A, X, W, n: int, info) -> None:
    """Cholesky + triangular inverse of the leading n x n block of views A/X/W (same leading dimension)."""
    assert A.stride(0) == X.stride(0) == W.stride(0) and A.stride(1) == 1
    with torch.cuda.device(A.device):
        _check(lib().qt_tri_chain_block(_p(A), _p(X), _p(W), n, A.stride(0), _p(info), _stream()), "qt_tri_chain_block")


def flip_upper(X, U) -> None:
    with torch.cuda.device(X.device):
        _check(lib().qt_flip_upper(_p(X), _p(U), X.shape[0], _stream()), "qt_flip_upper")


def gptq_permute_in(w: torch.Tensor, perm: Optional[torch.Tensor], dead: Optional[torch.Tensor]) -> torch.Tensor:
    _dev(w, "w")
    N, K = w.shape
    out = torch.empty((N, K), dtype=torch.float32, device=w.device)
    if perm is not None:
        perm = _dev(perm.to(torch.int32), "perm")
    with torch.cuda.device(w.device):
        _check(lib().qt_gptq_permute_in(_p(w), _DT[w.dtype], _p(perm), _p(dead), _p(out), N, K, _stream()),
               "qt_gptq_permute_in")
    return out


def gptq_permute_out(wp: torch.Tensor, inv_perm: Optional[torch.Tensor], dtype: torch.dtype) -> torch.Tensor:
    _dev(wp, "wp")
    N, K = wp.shape
    out = torch.empty((N, K), dtype=dtype, device=wp.device)
    if inv_perm is not None:
        inv_perm = _dev(inv_perm.to(torch.int32), "inv_perm")
    with torch.cuda.device(wp.device):
        _check(lib().qt_gptq_permute_out(_p(wp), _p(inv_perm), _p(out), _DT[dtype], N, K, _stream()),
               "qt_gptq_permute_out")
    return out


GPTQ_MODE_GROUP_REFIT, GPTQ_MODE_STATIC_GIDX, GPTQ_MODE_CHANNEL = 0, 1, 2


def split_tf32(x: torch.Tensor, hi: Optional[torch.Tensor] = None, lo: Optional[torch.Tensor] = None):
    """x = hi + lo with hi exactly representable in tf32 (operands of the 3xTF32 tensor-core GEMM)."""
    _dev(x, "x")
    hi = hi if hi is not None else torch.empty_like(x)
    lo = lo if lo is not None else torch.empty_like(x)
    with torch.cuda.device(x.device):
        _check(lib().qt_split_tf32(_p(x), _p(hi), _p(lo), x.numel(), _stream()), "qt_split_tf32")
    return hi, lo


def split_tf32_transpose(U: torch.Tensor, hi: Optional[torch.Tensor] = None, lo: Optional[torch.Tensor] = None):
    """(hi, lo) with hi + lo = U^T, hi tf32-exact: the B operand of the tensor-core lazy update."""
    _dev(U, "U")
    K = U.shape[0]
    hi = hi if hi is not None else torch.empty_like(U)
    lo = lo if lo is not None else torch.empty_like(U)
    with torch.cuda.device(U.device):
        _check(lib().qt_split_tf32_transpose(_p(U), _p(hi), _p(lo), K, _stream()), "qt_split_tf32_transpose")
    return hi, lo


def gptq_quantize_weight(wp: torch.Tensor, U: torch.Tensor, scale: torch.Tensor, zp: torch.Tensor,
                         g_idx: Optional[torch.Tensor], group_size: int, num_bits: int, symmetric: bool,
                         mode: int, err_scratch: Optional[torch.Tensor] = None,
                         U_split: Optional[tuple] = None) -> torch.Tensor:
    """Blocked GPTQ column loop, in place on wp (fp32 [N,K], permuted order).  Returns per-row losses.
    U_split = split_tf32_transpose(U) runs the lazy-batch update on the tensor cores."""
    _dev(wp, "wp"); _dev(U, "U"); _dev(scale, "scale"); _dev(zp, "zp")
    N, K = wp.shape
    G = scale.shape[1]
    losses = torch.zeros((N,), dtype=torch.float32, device=wp.device)
    nerr = 2 * 512 if U_split is not None else 128      # tensor-core path keeps Err of a 512-column outer block
    err = err_scratch if err_scratch is not None else torch.empty((nerr * N,), dtype=torch.float32, device=wp.device)
    assert err.numel() >= nerr * N
    if g_idx is not None:
        g_idx = _dev(g_idx.to(torch.int32), "g_idx")
    uh, ul = U_split if U_split is not None else (None, None)
    with torch.cuda.device(wp.device):
        _check(lib().qt_gptq_quantize_weight(_p(wp), _p(U), _p(uh), _p(ul), _p(err), _p(scale), _p(zp), _p(g_idx),
                                             _p(losses), N, K, G, max(group_size, 0), num_bits, int(symmetric), mode,
                                             _stream()), "qt_gptq_quantize_weight")
    return losses


# ---------------------------------------------------------------------------------------
# SmoothQuant / AWQ reductions and folds
# ---------------------------------------------------------------------------------------
FLT_MAX = 3.4028234663852886e38


def new_minmax(K: int, device):
    mn = torch.full((K,), FLT_MAX, dtype=torch.float32, device=device)
    mx = torch.full((K,), -FLT_MAX, dtype=torch.float32, device=device)
    return mn, mx


def channel_minmax(x: torch.Tensor, mn: torch.Tensor, mx: torch.Tensor) -> None:
    """Running per-channel min/max over the rows of x [T, K] into fp32 mn/mx [K]."""
    _dev(x, "x")
    x2 = x.reshape(-1, x.shape[-1])
    T, K = x2.shape
    with torch.cuda.device(x.device):
        _check(lib().qt_channel_minmax(_p(x2), _DT[x2.dtype], T, K, _p(mn), _p(mx), _stream()), "qt_channel_minmax")


def channel_abs_sum(x: torch.Tensor, acc: torch.Tensor) -> None:
    """acc [K] fp32 += sum over rows of |x|."""
    _dev(x, "x")
    x2 = x.reshape(-1, x.shape[-1])
    T, K = x2.shape
    with torch.cuda.device(x.device):
        _check(lib().qt_channel_abs_sum(_p(x2), _DT[x2.dtype], T, K, _p(acc), _stream()), "qt_channel_abs_sum")


def smooth_scales(amin, amax, wmin, wmax, alpha: float, compute_dtype: torch.dtype) -> torch.Tensor:
    K = amin.shape[0]
    out = torch.empty((K,), dtype=torch.float32, device=amin.device)
    with torch.cuda.device(amin.device):
        _check(lib().qt_smooth_scales(_p(amin), _p(amax), _p(wmin), _p(wmax), float(alpha), float(1 - alpha),
                                      _DT[compute_dtype], _p(out), K, _stream()), "qt_smooth_scales")
    return out


def scale_matrix_(w: torch.Tensor, s: torch.Tensor, divide: bool = False, by_row: bool = False) -> None:
    """In place, in w's dtype: w[n][c] *= s[c] (or /=; by_row: s[n]).  1-D w is treated as one row."""
    _dev(w, "w"); _dev(s, "s")
    assert s.dtype == torch.float32
    N, K = (1, w.shape[0]) if w.dim() == 1 else w.shape
    assert s.numel() == (N if by_row else K)
    with torch.cuda.device(w.device):
        for r0 in range(0, N, 65535):
            n = min(65535, N - r0)
            ptr = ctypes.c_void_p(w.data_ptr() + r0 * K * w.element_size())
            sp = ctypes.c_void_p(s.data_ptr() + (r0 * 4 if by_row else 0))
            _check(lib().qt_scale_matrix(ptr, _DT[w.dtype], n, K, sp, int(divide), int(by_row), _stream()),
                   "qt_scale_matrix")


def gemm_tf32x3(a_split, b_split, C: torch.Tensor, negate=False, accumulate=False, lower_tiles_only=False,
                a_tri: int = 0, b_tri: int = 0) -> torch.Tensor:
    """C (op)= +-A @ B.T on the tensor cores; a_split / b_split = (hi, lo) pairs from split_tf32 (views allowed)."""
    (ah, al), (bh, bl) = a_split, b_split
    for t in (ah, al, bh, bl, C):
        if not t.is_cuda:
            raise QtError("operand must be a CUDA tensor (no CPU fallback)")
        assert t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1
    M, Kd = ah.shape
    N = bh.shape[0]
    assert al.shape == ah.shape and bl.shape == bh.shape and bh.shape[1] == Kd and C.shape == (M, N)
    assert al.stride(0) == ah.stride(0) and bl.stride(0) == bh.stride(0)
    flags = int(negate) | (int(accumulate) << 1) | (int(lower_tiles_only) << 2) | (a_tri << 4) | (b_tri << 6)
    with torch.cuda.device(C.device):
        _check(lib().qt_gemm_tf32x3(_p(ah), _p(al), _p(bh), _p(bl), _p(C), M, N, Kd, ah.stride(0), bh.stride(0),
                                    C.stride(0), flags, _stream()), "qt_gemm_tf32x3")
    return C


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LlamaRMSNorm over the last dim of a contiguous x, one pass."""
    _dev(x, "x"); _dev(weight, "weight")
    assert x.is_contiguous() and weight.dtype == x.dtype and weight.numel() == x.shape[-1]
    H = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    assert out.is_contiguous() and out.shape == x.shape and out.dtype == x.dtype
    with torch.cuda.device(x.device):
        _check(lib().qt_rms_norm(_p(x), _p(weight), _p(out), _DT[x.dtype], x.numel() // H, H, float(eps), _stream()),
               "qt_rms_norm")
    return out


def rope_(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, seq: int, n_heads: int, head_dim: int) -> torch.Tensor:
    """In place rotary embedding on the projection output x[..., n_heads*head_dim] (token-major, contiguous)."""
    _dev(x, "x"); _dev(cos, "cos"); _dev(sin, "sin")
    assert x.is_contiguous() and cos.is_contiguous() and sin.is_contiguous()
    assert cos.dtype == x.dtype and sin.dtype == x.dtype and cos.shape == (seq, head_dim) and sin.shape == cos.shape
    assert x.shape[-1] == n_heads * head_dim
    T = x.numel() // (n_heads * head_dim)
    with torch.cuda.device(x.device):
        _check(lib().qt_rope_inplace(_p(x), _p(cos), _p(sin), _DT[x.dtype], T, seq, n_heads, head_dim, _stream()),
               "qt_rope_inplace")
    return x


def silu_mul(gate: torch.Tensor, up: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _dev(gate, "gate"); _dev(up, "up")
    assert gate.is_contiguous() and up.is_contiguous() and gate.shape == up.shape and gate.dtype == up.dtype
    out = torch.empty_like(gate) if out is None else out
    assert out.is_contiguous() and out.shape == gate.shape and out.dtype == gate.dtype
    with torch.cuda.device(gate.device):
        _check(lib().qt_silu_mul(_p(gate), _p(up), _p(out), _DT[gate.dtype], gate.numel(), _stream()), "qt_silu_mul")
    return out


def awq_wmean(weights, group_size: int) -> torch.Tensor:
    """Column mean of the group-normalised |W| over the concatenated balance weights; result in
    the weights' dtype (torch: bf16 ops, fp32-accumulated mean)."""
    K = weights[0].shape[1]
    dev = weights[0].device
    acc = torch.zeros((K,), dtype=torch.float32, device=dev)
    total = 0
    gs = group_size if group_size and group_size > 0 else 0
    with torch.cuda.device(dev):
        for w in weights:
            _dev(w, "w")
            N = w.shape[0]
            scratch = torch.empty((N, K // gs if gs else 1), dtype=torch.float32, device=dev)
            _check(lib().qt_awq_wmean_accumulate(_p(w), _DT[w.dtype], N, K, gs, _p(scratch), _p(acc), _stream()),
                   "qt_awq_wmean_accumulate")
            total += N
    return (acc / total).to(weights[0].dtype)


def awq_scale_qdq(w: torch.Tensor, s: torch.Tensor, group_size: int, num_bits: int, symmetric: bool,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """W' = pseudo_quant(W * s) / s in one pass (AWQ grid-search candidate weight), w's dtype."""
    _dev(w, "w"); _dev(s, "s")
    assert s.dtype == torch.float32
    N, K = w.shape
    if out is None:
        out = torch.empty_like(w)
    with torch.cuda.device(w.device):
        _check(lib().qt_awq_scale_qdq(_p(w), _DT[w.dtype], N, K, _p(s), max(group_size or 0, 0), num_bits,
                                      int(symmetric), _p(out), _stream()), "qt_awq_scale_qdq")
    return out


def awq_scale_qdq_delta(w: torch.Tensor, s: torch.Tensor, group_size: int, num_bits: int, symmetric: bool,
                        d_bf16: torch.Tensor, d_f32: torch.Tensor) -> None:
    """D = pseudo_quant(W * s) / s - W as bf16 (MMA operand) and fp32 (epilogue operand) for `awq_gram_loss`."""
    _dev(w, "w"); _dev(s, "s"); _dev(d_bf16, "d_bf16"); _dev(d_f32, "d_f32")
    assert s.dtype == torch.float32 and d_bf16.dtype == torch.bfloat16 and d_f32.dtype == torch.float32
    assert d_bf16.shape == w.shape and d_f32.shape == w.shape
    N, K = w.shape
    with torch.cuda.device(w.device):
        _check(lib().qt_awq_scale_qdq_delta(_p(w), _DT[w.dtype], N, K, _p(s), int(group_size), num_bits, int(symmetric),
                                            _p(d_bf16), _p(d_f32), _stream()), "qt_awq_scale_qdq_delta")


def awq_gram_loss(d_bf16: torch.Tensor, d_f32: torch.Tensor, g_bf16: torch.Tensor, acc: torch.Tensor) -> None:
    """acc (device float64 scalar) += tr(D G D^T) = ||X D^T||_F^2 with G = X^T X: one tcgen05 GEMM with a reducing
    epilogue (nothing is written but the scalar)."""
    _dev(d_bf16, "d_bf16"); _dev(d_f32, "d_f32"); _dev(g_bf16, "g_bf16")
    M, K = d_bf16.shape
    assert g_bf16.shape == (K, K) and g_bf16.dtype == torch.bfloat16 and acc.dtype == torch.float64
    with torch.cuda.device(d_bf16.device):
        _check(lib().qt_awq_gram_loss(_p(d_bf16), _p(d_f32), _p(g_bf16), M, K, _p(acc), _stream()), "qt_awq_gram_loss")


def awq_gram_ok(K: int, group_size: int) -> bool:
    """Shapes the Gram-form loss kernel takes (K-major bf16 tiles of 64, 16-column lanes)."""
    return K % 64 == 0 and group_size in (32, 64, 128)


def sq_err_sum(a: torch.Tensor, b: torch.Tensor, acc: torch.Tensor) -> None:
    """acc (device float64 scalar) += sum((a - b)^2), difference rounded to the tensors' dtype."""
    _dev(a, "a"); _dev(b, "b")
    assert a.dtype == b.dtype and a.numel() == b.numel() and acc.dtype == torch.float64
    with torch.cuda.device(a.device):
        _check(lib().qt_sq_err_sum(_p(a), _p(b), _DT[a.dtype], a.numel(), _p(acc), _stream()), "qt_sq_err_sum")
