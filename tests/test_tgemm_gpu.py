"""GPU parity: the 3xTF32 tensor-core GEMM of the inverse-Hessian chain vs an fp64 product."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, Kd, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn((M, Kd), generator=g, device="cuda")
    B = torch.randn((N, Kd), generator=g, device="cuda")
    return A, B


def _err(C, ref, A, B):
    # error relative to sum |a||b| (what an fp32 dot product is measured against)
    bound = A.abs().double() @ B.abs().double().T
    return ((C.double() - ref).abs() / bound.clamp_min(1e-30)).max().item()


@pytest.mark.parametrize("M,N,Kd", [(128, 256, 32), (300, 500, 160), (1024, 768, 1024), (129, 40, 64)])
@pytest.mark.parametrize("negate,accumulate", [(False, False), (True, False), (True, True)])
def test_gemm_tf32x3_vs_fp64(M, N, Kd, negate, accumulate):
    from quantool_b200 import cabi
    A, B = _mk(M, N, Kd)
    C0 = torch.randn((M, N), device="cuda")
    C = C0.clone()
    cabi.gemm_tf32x3(cabi.split_tf32(A), cabi.split_tf32(B), C, negate=negate, accumulate=accumulate)
    ref = A.double() @ B.double().T
    ref = (-ref if negate else ref) + (C0.double() if accumulate else 0)
    assert _err(C, ref, A, B) < 2e-6


def test_gemm_tf32x3_strided_views_and_guard_region():
    from quantool_b200 import cabi
    A, B = _mk(256, 512, 96, seed=3)
    big = torch.full((400, 700), 7.0, device="cuda")
    C = big[16:16 + 256, 32:32 + 512]
    ah, al = cabi.split_tf32(A)
    cabi.gemm_tf32x3((ah, al), cabi.split_tf32(B), C)
    assert _err(C, A.double() @ B.double().T, A, B) < 2e-6
    big[16:16 + 256, 32:32 + 512] = 7.0
    assert (big == 7.0).all()                         # nothing outside the sub-block was written
    # operands as column slices of wider arrays (ld > Kd)
    wide = torch.randn((256, 300), device="cuda")
    wh, wl = cabi.split_tf32(wide)
    C2 = torch.empty((256, 512), device="cuda")
    cabi.gemm_tf32x3((wh[:, 64:160], wl[:, 64:160]), cabi.split_tf32(B), C2)
    assert _err(C2, wide[:, 64:160].double() @ B.double().T, wide[:, 64:160], B) < 2e-6


def test_gemm_tf32x3_syrk_lower_and_triangular_operands():
    from quantool_b200 import cabi
    n, Kd = 1024, 512
    A, _ = _mk(n, n, Kd, seed=5)
    sp = cabi.split_tf32(A)
    C0 = torch.randn((n, n), device="cuda")
    C = C0.clone()
    cabi.gemm_tf32x3(sp, sp, C, negate=True, accumulate=True, lower_tiles_only=True)
    ref = C0.double() - A.double() @ A.double().T
    # strictly below the diagonal: sums of mixed sign, fp32-faithful.  ON the diagonal the sum of squares shows the
    # truncating fp32 accumulation of the tensor cores (~1e-5 relative) - the Cholesky recomputes those entries.
    low = torch.tril(torch.ones((n, n), device="cuda", dtype=torch.bool), -1)
    bound = A.abs().double() @ A.abs().double().T
    rel = (C.double() - ref).abs() / bound
    assert rel[low].max().item() < 2e-6
    assert torch.diagonal(rel).max().item() < 5e-5
    # tiles entirely above the diagonal are untouched: tile (tm, tn) covers rows 128tm.., cols 256tn..
    assert torch.equal(C[0:128, 256:], C0[0:128, 256:]) and torch.equal(C[256:384, 512:], C0[256:384, 512:])
    # triangular operands: same result as the dense product of the (explicitly zeroed) matrices
    L = torch.tril(torch.randn((n, n), device="cuda"))
    U = torch.triu(torch.randn((n, n), device="cuda"))
    Bm = torch.randn((768, n), device="cuda")
    for T, tri in ((L, 1), (U, 2)):
        out = torch.empty((n, 768), device="cuda")
        cabi.gemm_tf32x3(cabi.split_tf32(T), cabi.split_tf32(Bm), out, a_tri=tri)
        assert _err(out, T.double() @ Bm.double().T, T, Bm) < 2e-6
        out = torch.empty((768, n), device="cuda")
        cabi.gemm_tf32x3(cabi.split_tf32(Bm), cabi.split_tf32(T), out, b_tri=tri)
        assert _err(out, Bm.double() @ T.double().T, Bm, T) < 2e-6


def test_gemm_tf32x3_rejects_bad_shapes():
    from quantool_b200 import cabi
    A, B = _mk(128, 256, 40)
    with pytest.raises(cabi.QtError):
        cabi.gemm_tf32x3(cabi.split_tf32(A), cabi.split_tf32(B), torch.empty((128, 256), device="cuda"))
