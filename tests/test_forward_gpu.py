"""GPU parity: one-pass calibration-forward kernels (csrc/forward.cu) vs the plain-torch restatement
of the HF modules (oracle/llama_forward.py) on the same device tensors."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ulp_close(a, b, frac=0.99):
    """Elementwise equal except for rare 1-ulp flips caused by a different fp32 summation order."""
    a32, b32 = a.float(), b.float()
    same = (a32 == b32).float().mean().item()
    tol = 2.0 ** (-6 if a.dtype == torch.bfloat16 else -9 if a.dtype == torch.float16 else -20)   # 2 ulp: two roundings
    if a.dtype != torch.float32:          # in fp32 a different summation order shows in the last bit of most elements
        assert same >= frac, same
    assert ((a32 - b32).abs() <= tol * b32.abs() + 1e-30).all()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("H", [576, 4096])
def test_rms_norm_vs_torch(dtype, H):
    from quantool_b200 import cabi
    from oracle import llama_forward as lf
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn((3, 37, H), generator=g, device="cuda") * 2).to(dtype)
    x[0, 0] = 0                                        # zero row: rsqrt(eps)
    w = (1 + 0.1 * torch.randn((H,), generator=g, device="cuda")).to(dtype)
    _ulp_close(cabi.rms_norm(x, w, 1e-5), lf.rms_norm(x, w, 1e-5))
    out = torch.empty_like(x)
    assert cabi.rms_norm(x, w, 1e-5, out=out) is out


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("heads,hd", [(9, 64), (32, 128)])
def test_rope_bit_exact_vs_torch(dtype, heads, hd):
    from quantool_b200 import cabi
    from quantool_b200.engine import llama
    from oracle import llama_forward as lf
    B, S = 2, 50
    shape = llama.LlamaShape(heads * hd, 128, 1, heads, heads, 64, head_dim=hd)
    cos, sin = llama.rope_tables(shape, S, "cuda", dtype)
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn((B, S, heads * hd), generator=g, device="cuda").to(dtype)
    qt = q.view(B, S, heads, hd).transpose(1, 2)
    ref = (qt * cos[None, None] + lf._rot_half(qt) * sin[None, None]).transpose(1, 2).reshape(B, S, heads * hd)
    got = cabi.rope_(q.clone(), cos, sin, S, heads, hd)
    if dtype == torch.float32:
        # torch contracts a*c + r*s into an fma in fp32; the kernel rounds each product like the 16-bit path
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6)
    else:
        assert torch.equal(got, ref)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_silu_mul_vs_torch(dtype):
    from quantool_b200 import cabi
    g = torch.Generator(device="cuda").manual_seed(2)
    a = (torch.randn((5, 33, 1536), generator=g, device="cuda") * 3).to(dtype)
    b = torch.randn((5, 33, 1536), generator=g, device="cuda").to(dtype)
    ref = torch.nn.functional.silu(a) * b
    got = cabi.silu_mul(a, b)
    _ulp_close(got, ref, frac=0.999)
    with pytest.raises(cabi.QtError):
        cabi.silu_mul(a.cpu(), b.cpu())                # no CPU fallback


def test_layer_forward_matches_restatement_and_captures():
    from quantool_b200.engine import llama
    from oracle import llama_forward as lf
    shape = llama.LlamaShape(256, 512, 1, 4, 2, 64)
    w = llama.random_layer_weights(shape, 0, "cuda", seed=5)
    g = torch.Generator(device="cuda").manual_seed(3)
    B, S = 3, 40
    h = torch.randn((B, S, 256), generator=g, device="cuda").to(torch.bfloat16)
    cos, sin = llama.rope_tables(shape, S, "cuda", h.dtype)
    dims = shape.input_dims()
    cap_a = {n: torch.zeros((B * S + 7, k), dtype=h.dtype, device="cuda") for n, k in dims.items()}
    cap_b = {n: torch.zeros((B * S + 7, k), dtype=h.dtype, device="cuda") for n, k in dims.items()}
    out_a = llama.layer_forward(shape, w, h, cos, sin, capture=cap_a, row0=7)
    out_b = lf.layer_forward(shape, w, h, cos, sin, capture=cap_b, row0=7)
    for n in dims:
        assert (cap_a[n][:7] == 0).all()
        a, b = cap_a[n][7:].float(), cap_b[n][7:].float()
        assert (a - b).norm() <= 1e-2 * b.norm(), n
    assert (out_a.float() - out_b.float()).norm() <= 1e-2 * out_b.float().norm()
    # statistics passes stop early and still fill the captures they need
    cap_c = {n: torch.zeros((B * S, k), dtype=h.dtype, device="cuda") for n, k in dims.items()}
    assert llama.layer_forward(shape, w, h, cos, sin, capture=cap_c, stop_after="down_in") is None
    assert torch.equal(cap_c["down_in"], cap_a["down_in"][7:])
    cap_d = {n: torch.zeros((B * S, k), dtype=h.dtype, device="cuda") for n, k in dims.items()}
    assert llama.layer_forward(shape, w, h, cos, sin, capture=cap_d, stop_after="mlp_in") is None
    assert torch.equal(cap_d["mlp_in"], cap_a["mlp_in"][7:]) and (cap_d["down_in"] == 0).all()
    # parents used by the AWQ search
    x = cap_a["attn_in"][7:].view(B, S, -1)
    assert torch.allclose(llama.attention_forward(shape, w, x, cos, sin).float(),
                          lf.attention_forward(shape, w, x, cos, sin).float(), rtol=2e-2, atol=2e-2)
    x = cap_a["mlp_in"][7:].view(B, S, -1)
    assert torch.allclose(llama.mlp_forward(w, x).float(), lf.mlp_forward(w, x).float(), rtol=2e-2, atol=2e-2)
