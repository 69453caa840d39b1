"""GPU parity: GPTQ kernels (tcgen05 Hessian, fp32 Cholesky chain, blocked column loop, codes)
through the C-ABI against the CPU oracle (oracle/gptq.py, built on installed compressed-tensors)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _acts(T, K, seed=7, outliers=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((T, K), generator=g)
    if outliers:
        idx = torch.randperm(K, generator=g)[: max(1, K // 200)]
        x[:, idx] *= 20.0
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("T,K", [(512, 256), (1000, 576), (4096, 1024), (3000, 2048)])
def test_hessian_tcgen05_vs_fp64(T, K):
    """H within 1e-3 relative (north_star) of the fp64 value and of the oracle's running mean."""
    from quantool_b200 import cabi
    from oracle import gptq as og
    x = _acts(T, K)
    n_samples = 8
    H = torch.zeros((K, K), dtype=torch.float32, device="cuda")
    xc = x.cuda()
    half = (T // 2 // 8) * 8
    cabi.hessian_accumulate(xc[:half].contiguous(), H)      # two batches: accumulation path
    cabi.hessian_accumulate(xc[half:].contiguous(), H)
    cabi.hessian_finalize(H, 2.0 / n_samples)
    ref = (2.0 / n_samples) * (x.double().t() @ x.double())
    got = H.cpu().double()
    rel = torch.linalg.norm(got - ref) / torch.linalg.norm(ref)
    assert rel < 1e-5, rel
    assert torch.equal(H, H.t())
    # oracle (SURVEY §A.1 running mean, fp32)
    Ho, n = og.make_empty_hessian(K), 0
    for xb in x.reshape(n_samples, T // n_samples, K) if T % n_samples == 0 else [x]:
        Ho, n = og.accumulate_hessian(xb.unsqueeze(0) if T % n_samples == 0 else xb, Ho, n)
    if T % n_samples != 0:
        Ho = Ho * (1.0 / n_samples)   # one 2-D batch counted as 1 sample -> rescale to n_samples
    rel_o = torch.linalg.norm(got - Ho.double()) / torch.linalg.norm(Ho.double())
    assert rel_o < 1e-3, rel_o


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_hessian_full_width_vs_device_fp64(dtype):
    """K = 4096 (Llama-3-8B hidden) against an fp64 X^T X on the same device, for both 16-bit activation dtypes
    (an fp16 checkpoint's activations go to the tensor cores as fp16: no silent bf16 down-cast)."""
    from quantool_b200 import cabi
    T, K = 2048, 4096
    x = _acts(T, K).float()
    x = (x * (0.05 if dtype == torch.float16 else 1.0)).to(dtype).cuda()      # keep the x20 outliers inside fp16
    H = torch.zeros((K, K), dtype=torch.float32, device="cuda")
    cabi.hessian_accumulate(x, H)
    cabi.hessian_finalize(H, 1.0)
    R = x.double().t() @ x.double()
    rel = (torch.linalg.norm(H.double() - R) / torch.linalg.norm(R)).item()
    assert rel < 1e-5, rel
    if dtype == torch.float16:
        # what the old path did: round the activations to bf16 first -> 3 mantissa bits lost per element
        xb = x.to(torch.bfloat16).double()
        rel_cast = (torch.linalg.norm(xb.t() @ xb - R) / torch.linalg.norm(R)).item()
        assert rel < 0.05 * rel_cast, (rel, rel_cast)


def test_hessian_accumulator_fp16_exact_diagonal():
    from quantool_b200.engine import gptq as eg
    K, T = 512, 4096
    x = (_acts(T, K, seed=5).float() * 0.05).to(torch.float16)
    acc = eg.HessianAccumulator(K, "cuda")
    for xb in x.reshape(4, T // 4, K):
        acc.add(xb.unsqueeze(0).cuda())
    H = acc.finalize()
    d64 = (2.0 / 4) * (x.double() ** 2).sum(0)
    assert ((torch.diagonal(H).cpu().double() - d64).abs() / d64).max().item() < 5e-7
    with pytest.raises(Exception):
        eg.HessianAccumulator(K, "cuda").add(x.float().cuda())           # fp32 activations are rejected, not cast


@pytest.mark.parametrize("M,N,Kd,nk", [(300, 256, 128, False), (257, 384, 200, True), (128, 128, 64, False)])
def test_sgemm(M, N, Kd, nk):
    from quantool_b200 import cabi
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn((M, Kd), device="cuda", generator=g)
    B = torch.randn((N, Kd) if nk else (Kd, N), device="cuda", generator=g)
    C = torch.randn((M, N), device="cuda", generator=g)
    ref = -1.0 * (A.double() @ (B.double().t() if nk else B.double())) + C.double()
    cabi.sgemm(A, B, C, alpha=-1.0, beta=1.0, b_is_nk=nk)
    assert (C.double() - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("K,tc", [(128, False), (384, False), (576, False), (1024, False), (2048, False),
                                  (256, True), (1024, True), (2048, True), (2560, True), (3072, True)])
@pytest.mark.parametrize("act", [False, True])
def test_hinv_factor_vs_fp64(K, tc, act):
    """tc = the tensor-core (3xTF32) chain; the FFMA chain serves K that are not multiples of 256."""
    from quantool_b200 import cabi
    T = 4 * K
    x = _acts(T, K, seed=K)
    H = ((2.0 / 16) * (x.double().t() @ x.double()))
    Hc = H.float().cuda()
    perm = torch.argsort(torch.diagonal(Hc), descending=True, stable=True).to(torch.int32) if act else None
    Hf, dead = cabi.gptq_prepare_hessian(Hc, perm, 0.01)
    U, info = cabi.gptq_hinv_factor(Hf, tensor_core=tc)
    assert int(info.item()) == 0
    Hd = H.clone()
    if act:
        p = perm.cpu().long()
        Hd = Hd[p][:, p]
    Hd += 0.01 * torch.mean(torch.diag(Hd)) * torch.eye(K, dtype=torch.float64)
    ref = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hd)), upper=True)
    got = U.cpu().double()
    assert torch.equal(torch.tril(got, -1), torch.zeros_like(got))
    rel = torch.linalg.norm(got - ref) / torch.linalg.norm(ref)
    assert rel < 2e-4, rel


def test_hinv_not_pd_reports_info():
    from quantool_b200 import cabi
    K = 256
    H = -torch.eye(K, device="cuda")
    Hf, _ = cabi.gptq_prepare_hessian(H, None, 0.0)
    _, info = cabi.gptq_hinv_factor(Hf, tensor_core=False)
    assert int(info.item()) != 0
    Hf, _ = cabi.gptq_prepare_hessian(H, None, 0.0)
    _, info = cabi.gptq_hinv_factor(Hf, tensor_core=True)
    assert int(info.item()) != 0


CASES = [
    ("W4A16", None, 96, 512),
    ("W4A16", "group", 96, 512),
    ("W4A16", "weight", 64, 384),
    ("W4A16_ASYM", "group", 80, 256),
    ("W8A16", None, 72, 320),
    ("W8A8", None, 64, 576),
]


@pytest.mark.parametrize("level,actorder,N,K", CASES)
def test_gptq_quantize_vs_oracle(level, actorder, N, K):
    """north_star bar: artifact integer codes agree on >= 99.9 % of entries, ||WX - QX|| within 1 %."""
    from quantool_b200.engine import gptq as eg, schemes
    from oracle import gptq as og
    from compressed_tensors.quantization import ActivationOrdering
    g = torch.Generator().manual_seed(N * 1000 + K)
    W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
    T = 8 * K
    x = _acts(T, K, seed=K + 1)
    # oracle
    oargs = og.scheme_weight_args(level)
    if actorder is not None:
        oargs.actorder = ActivationOrdering.GROUP if actorder == "group" else ActivationOrdering.WEIGHT
    Ho, n = og.make_empty_hessian(K), 0
    for xb in x.reshape(8, T // 8, K):
        Ho, n = og.accumulate_hessian(xb.unsqueeze(0), Ho, n)
    loss_o, Wq_o, s_o, z_o, gi_o = og.quantize_weight(W, Ho, oargs)
    # CUDA path
    args = schemes.resolve(level, actorder)
    acc = eg.HessianAccumulator(K, "cuda")
    for xb in x.reshape(8, T // 8, K):
        acc.add(xb.unsqueeze(0).cuda())
    H = acc.finalize()
    relH = (torch.linalg.norm(H.cpu() - Ho) / torch.linalg.norm(Ho)).item()
    assert relH < 1e-3, relH
    res = eg.quantize_linear(W.cuda(), H, args)
    assert int(res.info.item()) == 0
    # artifact codes (what save_compressed stores)
    if args.num_bits == 4:
        codes_o, packed_o, pzp_o = og.compress_packed(Wq_o, s_o, z_o if not args.symmetric else None, gi_o, oargs)
        art, codes = eg.compress_linear(res.weight, res.scale, res.zero_point, res.g_idx, args)
    else:
        codes_o = og.compress_int8(Wq_o, s_o, None, oargs)
        art, codes = eg.compress_linear(res.weight, res.scale, res.zero_point, res.g_idx, args, fmt="int-quantized")
    agree = (codes.cpu() == codes_o).float().mean().item()
    assert agree >= 0.999, f"code agreement {agree:.5f}"
    if gi_o is not None:
        assert torch.equal(res.g_idx.cpu().to(torch.int64), gi_o.to(torch.int64))
    e_o = og.layer_error(W, Wq_o, x.float())
    e_c = og.layer_error(W, res.weight.cpu(), x.float())
    assert abs(e_c - e_o) <= 0.01 * e_o, (e_c, e_o)
    # and GPTQ must beat plain round-to-nearest on its own objective
    Wr, _, _ = og.rtn_quantize(W, og.scheme_weight_args(level))
    assert e_c < og.layer_error(W, Wr, x.float())
    loss_c = res.losses.sum().item()
    assert abs(loss_c - loss_o) <= 0.02 * abs(loss_o) + 1e-12, (loss_c, loss_o)


@pytest.mark.parametrize("level", ["W4A16", "W4A16_ASYM", "W8A16"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_compress_codes_bit_exact_vs_compressed_tensors(level, dtype):
    """Row a6: codes/packing from a saved (model dtype) weight + scale are integer work: bit-exact."""
    from quantool_b200 import cabi
    from quantool_b200.engine import gptq as eg, schemes
    from oracle import gptq as og
    N, K = 70, 768
    g = torch.Generator().manual_seed(5)
    W = (torch.randn((N, K), generator=g) * 0.05).to(dtype)
    oargs = og.scheme_weight_args(level)
    args = schemes.resolve(level)
    Wq_o, s_o, z_o = og.rtn_quantize(W, oargs)
    # observer parity (evaluated in the weight's dtype)
    s_c, z_c = cabi.minmax_qparams(W.cuda(), args.group_size or 0, args.num_bits, args.symmetric)
    assert torch.equal(s_c.cpu().to(dtype), s_o)
    assert torch.equal(z_c.cpu().to(torch.int8), z_o.to(torch.int8))
    gi = None
    if args.strategy == "group":
        gi = (torch.arange(K) // args.group_size)[torch.randperm(K, generator=g)].to(torch.int32)
    if args.num_bits == 4:
        codes_o, packed_o, pzp_o = og.compress_packed(W, s_o, z_o if not args.symmetric else None, gi, oargs)
        art, codes = eg.compress_linear(W.cuda(), s_o.cuda(), z_o.to(torch.int8).cuda(),
                                        gi.cuda() if gi is not None else None, args)
        assert torch.equal(codes.cpu(), codes_o)
        assert torch.equal(art["weight_packed"].cpu(), packed_o)
        if pzp_o is not None:
            assert torch.equal(art["weight_zero_point"].cpu(), pzp_o)
        back = cabi.unpack_int32(art["weight_packed"], 4, K)
        assert torch.equal(back, codes)
    else:
        codes_o = og.compress_int8(W, s_o, None, oargs)
        art, codes = eg.compress_linear(W.cuda(), s_o.cuda(), z_o.to(torch.int8).cuda(), None, args,
                                        fmt="int-quantized")
        assert torch.equal(codes.cpu(), codes_o)


def test_pack_kat():
    """SURVEY §8c KAT: pack_to_int32([-8,-7,0,7,1,2,3,4, 5,6,7,-1,-2,-3,-4,-5], 4)."""
    from quantool_b200 import cabi
    v = torch.tensor([[-8, -7, 0, 7, 1, 2, 3, 4, 5, 6, 7, -1, -2, -3, -4, -5]], dtype=torch.int8).cuda()
    p = cabi.pack_int32(v, 4).cpu().numpy().astype(np.uint32)
    assert p.tolist() == [[0xCBA9F810, 0x34567FED]]


@pytest.mark.parametrize("M,K,i1,i2", [(256, 1024, 128, 256), (300, 2048, 0, 128), (96, 640, 256, 384)])
def test_lazy_update_tensor_core_tf32x3(M, K, i1, i2):
    """3xTF32 tcgen05 lazy-batch update is fp32-faithful: compare with fp64 and with the FFMA GEMM."""
    from quantool_b200 import cabi
    g = torch.Generator(device="cuda").manual_seed(M + K)
    W = torch.randn((M, K), device="cuda", generator=g)
    U = torch.triu(torch.randn((K, K), device="cuda", generator=g)) * 0.1
    err = torch.randn((M, 128), device="cuda", generator=g)
    ref = W.double().clone()
    ref[:, i2:] -= err.double() @ U.double()[i1:i1 + 128, i2:]
    W_simt = W.clone()
    cabi.sgemm(err, U[i1:i1 + 128, i2:], W_simt[:, i2:], alpha=-1.0, beta=1.0)
    uh, ul = cabi.split_tf32_transpose(U)
    eh, el = cabi.split_tf32(err)
    # both halves are tf32 values (low 13 mantissa bits clear) and together carry x to 2^-22 relative
    for part in (uh, ul, eh, el):
        assert torch.equal((part.view(torch.int32) & 0x1FFF), torch.zeros_like(part, dtype=torch.int32))
    assert ((uh + ul - U.t()).abs() <= 2.0 ** -22 * U.t().abs()).all()
    W_tc = W.clone()
    # the product path's call (csrc/gptq.cu): C[:, i2:] -= Err * (U^T[i2:, i1:i1+128])^T, TMA reduce-add epilogue
    cabi.gemm_tf32x3((eh, el), (uh[i2:, i1:i1 + 128], ul[i2:, i1:i1 + 128]), W_tc[:, i2:], negate=True, accumulate=True)
    assert torch.equal(W_tc[:, :i2], W[:, :i2])                       # untouched columns
    scale = (err.double().abs() @ U.double().abs()[i1:i1 + 128, i2:]).max().item()
    e_tc = (W_tc.double() - ref).abs().max().item() / scale
    e_simt = (W_simt.double() - ref).abs().max().item() / scale
    assert e_tc < 2e-6, (e_tc, e_simt)
    assert e_tc < 20 * e_simt + 1e-7, (e_tc, e_simt)


@pytest.mark.parametrize("level,actorder,N,K", [("W4A16", "group", 160, 768), ("W8A8", None, 96, 640)])
def test_gptq_tensor_core_path_vs_oracle(level, actorder, N, K):
    """Same parity bar as the FFMA path, with the lazy-batch update on the tensor cores."""
    from quantool_b200 import cabi
    from quantool_b200.engine import gptq as eg, schemes
    from oracle import gptq as og
    from compressed_tensors.quantization import ActivationOrdering
    g = torch.Generator().manual_seed(K)
    W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
    x = _acts(8 * K, K, seed=K + 5)
    oargs = og.scheme_weight_args(level)
    if actorder:
        oargs.actorder = ActivationOrdering.GROUP
    Ho, n = og.make_empty_hessian(K), 0
    for xb in x.reshape(8, K, K):
        Ho, n = og.accumulate_hessian(xb.unsqueeze(0), Ho, n)
    _, Wq_o, s_o, z_o, gi_o = og.quantize_weight(W, Ho, oargs)
    args = schemes.resolve(level, actorder)
    res = eg.quantize_linear(W.cuda(), Ho.cuda(), args, tensor_core_lazy=True)
    if args.num_bits == 4:
        codes_o, _, _ = og.compress_packed(Wq_o, s_o, None, gi_o, oargs)
        _, codes = eg.compress_linear(res.weight, res.scale, res.zero_point, res.g_idx, args)
    else:
        codes_o = og.compress_int8(Wq_o, s_o, None, oargs)
        _, codes = eg.compress_linear(res.weight, res.scale, res.zero_point, res.g_idx, args, fmt="int-quantized")
    agree = (codes.cpu() == codes_o).float().mean().item()
    assert agree >= 0.999, agree
    e_o = og.layer_error(W, Wq_o, x.float())
    e_c = og.layer_error(W, res.weight.cpu(), x.float())
    assert abs(e_c - e_o) <= 0.01 * e_o, (e_c, e_o)


def test_dead_columns_and_identity_fallback_vs_oracle():
    """Edge cases upstream handles explicitly: a calibration column that is always zero
    (diag(H) == 0 -> H[d,d] = 1, W[:, d] = 0) and a Hessian whose factorisation fails (Hinv = I)."""
    from quantool_b200 import cabi
    from quantool_b200.engine import gptq as eg, schemes
    from oracle import gptq as og
    N, K = 64, 256
    g = torch.Generator().manual_seed(3)
    W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
    x = _acts(4 * K, K, seed=9)
    x[:, 17] = 0
    x[:, 200] = 0
    oargs = og.scheme_weight_args("W4A16")
    Ho, n = og.make_empty_hessian(K), 0
    for xb in x.reshape(4, K, K):
        Ho, n = og.accumulate_hessian(xb.unsqueeze(0), Ho, n)
    assert Ho[17, 17] == 0 and Ho[200, 200] == 0
    _, Wq_o, s_o, z_o, _ = og.quantize_weight(W, Ho, oargs)
    args = schemes.resolve("W4A16")
    res = eg.quantize_linear(W.cuda(), Ho.cuda(), args)
    assert int(res.info.item()) == 0
    assert torch.all(res.weight[:, 17] == 0) and torch.all(res.weight[:, 200] == 0)
    codes_o, _, _ = og.compress_packed(Wq_o, s_o, None, None, oargs)
    _, codes = eg.compress_linear(res.weight, res.scale, res.zero_point, None, args)
    assert (codes.cpu() == codes_o).float().mean().item() >= 0.999
    # not positive definite: upstream catches LinAlgError and uses Hinv = I (plain RTN with no feedback)
    Hbad = -torch.eye(K)
    _, Wq_b, s_b, _, _ = og.quantize_weight(W, Hbad.clone(), oargs)
    res_b = eg.quantize_linear(W.cuda(), Hbad.cuda(), args)
    assert int(res_b.info.item()) != 0
    codes_ob, _, _ = og.compress_packed(Wq_b, s_b, None, None, oargs)
    _, codes_b = eg.compress_linear(res_b.weight, res_b.scale, res_b.zero_point, None, args)
    assert (codes_b.cpu() == codes_ob).float().mean().item() >= 0.999


def test_hessian_empty_and_unaligned_errors():
    from quantool_b200 import cabi
    H = torch.zeros((64, 64), device="cuda")
    cabi.hessian_accumulate(torch.empty((0, 64), device="cuda", dtype=torch.bfloat16), H)   # empty batch: no-op
    assert float(H.abs().sum()) == 0.0
    with pytest.raises(cabi.QtError):
        cabi.hessian_accumulate(torch.zeros((8, 60), device="cuda", dtype=torch.bfloat16),
                                torch.zeros((60, 60), device="cuda"))                        # K % 8 != 0
    with pytest.raises(cabi.QtError):
        cabi.hessian_accumulate(torch.zeros((8, 64), device="cuda"), H)                      # fp32 activations


def test_hessian_exact_diagonal():
    """Row a1: the tensor-core accumulator truncates; the diagonal (which decides the act_order permutation) is
    accumulated separately in fp32 round-to-nearest and written into H."""
    from quantool_b200 import cabi
    from quantool_b200.engine import gptq as eg
    K, T = 1024, 16384
    x = _acts(T, K, seed=3)
    acc = eg.HessianAccumulator(K, "cuda")
    for xb in x.reshape(4, T // 4, K):
        acc.add(xb.cuda(), 2)
    raw_tc = torch.diagonal(acc.H).clone()                   # tensor-core sums, before the exact diagonal is written
    H = acc.finalize()
    d64 = (2.0 / 8) * (x.double() ** 2).sum(0)
    got = torch.diagonal(H).cpu().double()
    assert ((got - d64).abs() / d64).max().item() < 5e-7
    tc = (2.0 / 8) * raw_tc.cpu().double()
    assert ((tc - d64) / d64).mean().item() < 0            # the bias this pass removes: truncation -> too small
    assert torch.equal(H, H.t())
    # direct C-ABI use: accumulate twice = double
    diag = torch.zeros((K,), device="cuda")
    scr = torch.empty((32 * K,), device="cuda")
    xb = x[:4096].cuda().contiguous()
    cabi.hessian_diag_accumulate(xb, diag, scr)
    one = diag.clone()
    cabi.hessian_diag_accumulate(xb, diag, scr)
    assert torch.allclose(diag, 2 * one, rtol=1e-6)
    with pytest.raises(cabi.QtError):
        cabi.hessian_diag_accumulate(xb.cpu(), diag, scr)


@pytest.mark.parametrize("actorder,own_h,floor", [(None, True, 0.9999), ("group", False, 0.999)])
def test_gptq_larger_k_vs_oracle(actorder, own_h, floor):
    """K = 2048 (tensor-core chain + two-level lazy update).  Without act_order the codes match the oracle to the
    last entry with the GPU's own Hessian.  With act_order the comparison is made on the oracle's H: argsort(diag H)
    breaks near-ties differently for two fp32 evaluations of the same diagonal, which moves group boundaries
    (DESIGN.md §6 "Parity at larger K")."""
    from quantool_b200.engine import gptq as eg, schemes
    from oracle import gptq as og
    from compressed_tensors.quantization import ActivationOrdering
    N, K = 48, 2048
    g = torch.Generator().manual_seed(77)
    W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
    T = 4 * K
    x = _acts(T, K, seed=K + 9)
    oargs = og.scheme_weight_args("W4A16")
    if actorder:
        oargs.actorder = ActivationOrdering.GROUP
    Ho, n = og.make_empty_hessian(K), 0
    for xb in x.reshape(8, T // 8, K):
        Ho, n = og.accumulate_hessian(xb.unsqueeze(0), Ho, n)
    loss_o, Wq_o, s_o, z_o, gi_o = og.quantize_weight(W, Ho.clone(), oargs)
    codes_o, _, _ = og.compress_packed(Wq_o, s_o, None, gi_o, oargs)
    args = schemes.resolve("W4A16", actorder)
    if own_h:
        acc = eg.HessianAccumulator(K, "cuda")
        for xb in x.reshape(8, T // 8, K):
            acc.add(xb.unsqueeze(0).cuda())
        H = acc.finalize()
    else:
        H = Ho.cuda()
    res = eg.quantize_linear(W.cuda(), H, args)
    assert int(res.info.item()) == 0
    _, codes = eg.compress_linear(res.weight, res.scale, res.zero_point, res.g_idx, args)
    agree = (codes.cpu() == codes_o).float().mean().item()
    assert agree >= floor, agree
    e_o = og.layer_error(W, Wq_o, x.float())
    e_c = og.layer_error(W, res.weight.cpu(), x.float())
    assert abs(e_c - e_o) <= 0.01 * e_o


# ---- BASELINE widths, the path's OWN Hessian, actorder=group (VERDICT r01 "Next" 1a) --------------------
def _fp64_diag(x, n_samples):
    d = torch.zeros(x.shape[1], dtype=torch.float64)
    for xb in x.split(2048):
        d += (xb.double() ** 2).sum(0)
    return d * (2.0 / n_samples)


@pytest.mark.parametrize("N,K,T", [(64, 4096, 32768), (64, 8192, 65536), (32, 14336, 65536), (64, 4096, 8192)])
def test_gptq_baseline_widths_own_hessian(N, K, T):
    """Rows a1-a6 end to end at the widths of BASELINE's 8B / 1B configs (hidden 4096, intermediate 8192 /
    14336), W4A16 g128 actorder=group, everything from the CUDA path's own tensor-core Hessian.

    Asserted: H within 1e-3 (north_star) of the oracle's; ||WX - QX|| within 1 % of the oracle's; integer codes
    >= 99.9 % equal to the oracle's when the oracle is given the SAME act_order permutation; and that the two
    permutations differ only where diag(H) is tied to fp32 resolution (both are argsort of an fp32 evaluation
    of the same sums: the oracle's running mean, the GPU's two-stage column reduction).  Against the oracle's own
    permutation the codes differ more, because every swapped pair of near-tied columns that straddles a group
    boundary changes that group's members; the figure is printed and floored at 99 %.

    Conditioning.  BASELINE calibrates with 128 x 2048 tokens, i.e. T/K = 64 (K = 4096) and 18 (K = 14336); the
    first three cases use T/K = 8 / 8 / 4.6 (what the CPU oracle finishes in seconds) and must reach 99.9 %.  The
    last case (T/K = 2) is deliberately ill-conditioned: there GPTQ's error feedback amplifies fp32 rounding of H
    itself - the oracle run on two fp32 evaluations of the same H (relative difference 5e-7, the same distance the
    CUDA Hessian has from the oracle's) agrees with itself on only 99.2-99.9 % of the codes at these widths
    (profiles/r02_noise_floor.json) - so the assertion is made against that measured floor."""
    from quantool_b200.engine import gptq as eg, schemes
    from oracle import gptq as og
    from compressed_tensors.quantization import ActivationOrdering
    ns = 8
    g = torch.Generator().manual_seed(K + N)
    W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
    x = _acts(T, K, seed=K + 11)
    oargs = og.scheme_weight_args("W4A16")
    oargs.actorder = ActivationOrdering.GROUP
    Ho, n = og.make_empty_hessian(K), 0
    for xb in x.reshape(ns, T // ns, K):
        Ho, n = og.accumulate_hessian(xb.unsqueeze(0), Ho, n)
    args = schemes.resolve("W4A16", "group")
    acc = eg.HessianAccumulator(K, "cuda")
    for xb in x.reshape(ns, T // ns, K):
        acc.add(xb.unsqueeze(0).cuda())
    H = acc.finalize()
    relH = (torch.linalg.norm(H.cpu() - Ho) / torch.linalg.norm(Ho)).item()
    assert relH < 1e-3, relH
    res = eg.quantize_linear(W.cuda(), H, args)
    assert int(res.info.item()) == 0
    _, codes = eg.compress_linear(res.weight, res.scale, res.zero_point, res.g_idx, args)
    codes = codes.cpu()
    perm_g = res.perm.cpu().long()

    # (1) the permutations: equal except inside fp32 near-ties of the diagonal
    perm_o = torch.argsort(torch.diag(Ho), descending=True, stable=True)
    d64 = _fp64_diag(x, ns)
    same = (perm_g == perm_o).float().mean().item()
    tie = ((d64[perm_g] - d64[perm_o]).abs() / d64[perm_o]).max().item()
    inv_g = int((d64[perm_g][1:] > d64[perm_g][:-1]).sum())       # adjacent inversions vs the fp64 diagonal
    inv_o = int((d64[perm_o][1:] > d64[perm_o][:-1]).sum())
    print(f"[{N},{K}] T={T}: relH {relH:.2e}; perm positions equal {same:.4f}, max rel. diag gap at a differing "
          f"position {tie:.2e}; adjacent inversions vs fp64 diag: gpu {inv_g}, oracle {inv_o}")
    assert tie < 4e-6, tie                      # every disagreement is a tie at fp32 resolution
    assert inv_g <= inv_o + 2                   # the GPU diagonal orders at least as faithfully as the oracle's
    if K <= 4096:
        assert same >= 0.98, same

    # (2) codes with the permutation pinned: >= 99.9 %
    _, Wq_p, s_p, z_p, gi_p = og.quantize_weight(W, Ho.clone(), oargs, perm_override=perm_g)
    codes_p, _, _ = og.compress_packed(Wq_p, s_p, None, gi_p, oargs)
    assert torch.equal(gi_p.to(torch.int32), res.g_idx.cpu())
    agree_p = (codes == codes_p).float().mean().item()
    # (3) objective and codes against the oracle's own permutation
    _, Wq_o, s_o, z_o, gi_o = og.quantize_weight(W, Ho.clone(), oargs)
    codes_o, _, _ = og.compress_packed(Wq_o, s_o, None, gi_o, oargs)
    agree_o = (codes == codes_o).float().mean().item()
    xs = x[:4096].float()
    e_o = og.layer_error(W, Wq_o, xs)
    e_p = og.layer_error(W, Wq_p, xs)
    e_c = og.layer_error(W, res.weight.cpu(), xs)
    print(f"    codes equal: same perm {agree_p:.5f}, oracle's perm {agree_o:.5f}; ||WX-QX|| gpu {e_c:.5f} "
          f"oracle {e_o:.5f} oracle(same perm) {e_p:.5f}")
    if T >= 4 * K:
        assert agree_p >= 0.999, agree_p
    else:
        # the restated reference's own sensitivity to fp32 rounding of H: same sums accumulated in fp64, rounded once
        H64 = torch.zeros((K, K), dtype=torch.float64, device="cuda")
        for xb in x.cuda().split(4096):
            H64 += xb.double().t() @ xb.double()
        H2 = (H64 * (2.0 / ns)).float().cpu()
        _, Wq_2, s_2, _, gi_2 = og.quantize_weight(W, H2, oargs, perm_override=perm_g)
        codes_2, _, _ = og.compress_packed(Wq_2, s_2, None, gi_2, oargs)
        floor = (codes_2 == codes_p).float().mean().item()
        print(f"    ill-conditioned case: oracle(H) vs oracle(H rounded differently) agree on {floor:.5f}")
        assert agree_p >= min(floor, 0.999) - 0.004, (agree_p, floor)
    # vs the oracle's OWN permutation: 6 % of 14336 positions are fp32 near-ties (asserted above to be ties), and each
    # swapped pair that straddles a group boundary changes that group's members, hence its scale and its 128 codes
    assert agree_o >= (0.99 if K <= 8192 else 0.95), agree_o
    assert abs(e_c - e_o) <= 0.01 * e_o
    assert abs(e_c - e_p) <= 0.01 * e_p


@pytest.mark.parametrize("K,align", [(1024, 256), (1024, 128), (2816, 256), (2816, 128), (200, 128), (333, 256)])
def test_tri_pack_round_trip(K, align):
    """Packed upper block-triangle (what the all-reduce of H and the broadcast of U move): every element of the
    stored region survives the round trip bit for bit, nothing outside it is touched (H) or it is zeroed (U)."""
    from quantool_b200 import cabi
    M = torch.randn((K, K), device="cuda")
    r = torch.arange(K, device="cuda")
    c0 = ((r // 128) * 128 // align) * align
    stored = r[None, :] >= c0[:, None]
    n = cabi.tri_packed_elems(K, align)
    assert n == int(stored.sum().item())
    p = cabi.tri_pack(M, align)
    assert p.numel() == n
    assert torch.equal(p, M[stored])                      # row-major within a row block == row-major overall per block
    out = torch.full((K, K), 7.0, device="cuda")
    cabi.tri_unpack(p, out, align)
    assert torch.equal(out[stored], M[stored]) and bool((out[~stored] == 7.0).all())
    out2 = torch.full((K, K), 7.0, device="cuda")
    cabi.tri_unpack(p, out2, align, zero_below=True)
    assert torch.equal(out2, torch.where(stored, M, torch.zeros_like(M)))


@pytest.mark.parametrize("level,actorder,N,K", [("W4A16", "group", 200, 1024), ("W4A16", "weight", 96, 512),
                                                ("W8A16", None, 77, 768), ("W4A16_ASYM", "group", 130, 1024),
                                                ("W4A16", "group", 4100, 256)])
def test_lean_block_kernel_equals_generic(level, actorder, N, K):
    """The lean full-block column kernel (division through a refined reciprocal, magic-constant rounding, group
    parameters fitted before the walk) must give the same bits as the generic kernel: fake-quantized weight,
    scales, zero points and per-row losses."""
    from quantool_b200 import cabi
    from quantool_b200.engine import gptq as eg, schemes
    g = torch.Generator().manual_seed(N + K)
    W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16).cuda()
    W[3, :] = 0                                           # an all-zero row (scale -> eps)
    W[5, 17] = 3.0                                        # an outlier
    x = _acts(2048, K, seed=K)
    acc = eg.HessianAccumulator(K, "cuda")
    acc.add(x.cuda())
    H = acc.finalize()
    args = schemes.resolve(level, actorder)
    outs = []
    try:
        for generic in (1, 0):
            cabi.lib().qt_gptq_set_block_kernel(generic)
            r = eg.quantize_linear(W, H, args)
            outs.append((r.weight.clone(), r.scale.clone(), r.zero_point.clone(), r.losses.clone()))
    finally:
        cabi.lib().qt_gptq_set_block_kernel(0)
    for a, b in zip(*outs):
        assert torch.equal(a, b)
