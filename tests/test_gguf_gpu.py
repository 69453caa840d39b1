"""GPU parity: GGUF packers / unpackers through the C-ABI vs the C oracle and gguf-py."""
import hashlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ALL = ["Q8_0", "Q4_0", "Q4_1", "Q5_0", "Q5_1", "IQ4_NL", "Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"]
KAT = {"Q8_0": "d8b3256c9fd9ae4c", "Q4_0": "d153bbca62836335", "Q5_0": "c9154c30f4008bbe",
       "Q4_1": "f1844724bdbc3ca3", "Q5_1": "2cce57372f71162a"}


def _edge_input(rng, nrows, ncols):
    x = rng.standard_normal((nrows, ncols)).astype(np.float32)
    x[0, :256] = 0.0                      # all-zero super-block
    x[1, :32] = 1.5                       # constant sub-block
    x[2, :] *= 1e-3
    x[3, 5] = 40.0                        # outlier
    x[4, :256] = np.abs(x[4, :256])       # all-positive (min clamps to 0)
    x[5, :256] = 1e-20                    # below GROUP_MAX_EPS
    x[6, 0] = -x[6, 1]                    # |v| tie with opposite sign: first wins
    return x


@pytest.mark.parametrize("qtype", ALL)
def test_kat_seed0(qtype):
    """SURVEY §8c known-answer vectors (gguf-py on seed-0 input) through the CUDA path."""
    from quantool_b200 import cabi
    from oracle import ggml_quants as oq
    x = np.random.default_rng(0).standard_normal((8, 512)).astype(np.float32)
    y = cabi.gguf_quantize(torch.from_numpy(x).cuda(), qtype, round_via_f16=False).cpu().numpy()
    assert np.array_equal(y, oq.quantize(x, qtype))
    if qtype in KAT:
        assert hashlib.sha256(y.tobytes()).hexdigest()[:16] == KAT[qtype]


@pytest.mark.parametrize("qtype", ALL)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_pack_bit_exact_vs_oracle(qtype, dtype):
    from quantool_b200 import cabi
    from oracle import ggml_quants as oq
    rng = np.random.default_rng(11)
    x = _edge_input(rng, 67, 1280)        # 67 rows: ragged vs the 256-block CTA tile
    xt = torch.from_numpy(x).to(dtype)
    via = dtype != torch.float16
    xin = xt.float().numpy()
    if via:
        xin = oq.round_f16(xin)
    ref = oq.quantize(xin, qtype)
    y = cabi.gguf_quantize(xt.cuda(), qtype, round_via_f16=via).cpu().numpy()
    assert y.shape == ref.shape
    bad = np.argwhere(y != ref)
    assert bad.size == 0, f"{qtype}/{dtype}: {len(bad)} byte mismatches, first at {bad[:5].tolist()}"


@pytest.mark.parametrize("qtype", ALL)
def test_dequant_bit_exact_vs_gguf_py(qtype):
    from gguf import GGMLQuantizationType as T
    from gguf import quants as gq
    from quantool_b200 import cabi
    from oracle import ggml_quants as oq
    rng = np.random.default_rng(5)
    x = _edge_input(rng, 19, 768)
    packed = oq.quantize(x, qtype)
    ref = gq.dequantize(packed, getattr(T, qtype))
    got = cabi.gguf_dequantize(torch.from_numpy(packed).cuda(), qtype, 768).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), ref.astype(np.float32).view(np.uint32))
    assert np.array_equal(got.view(np.uint32), oq.dequantize(packed, qtype, 768).view(np.uint32))


@pytest.mark.parametrize("qtype", ["Q8_0", "Q4_K", "Q6_K"])
def test_empty_and_errors(qtype):
    from quantool_b200 import cabi
    be = cabi.gguf_block_elems(qtype)
    y = cabi.gguf_quantize(torch.empty((0, be), device="cuda"), qtype)
    assert y.shape[0] == 0
    with pytest.raises(cabi.QtError):
        cabi.gguf_quantize(torch.zeros((2, be + 8), device="cuda"), qtype)
    with pytest.raises(cabi.QtError):
        cabi.gguf_quantize(torch.zeros((2, be)), qtype)  # CPU tensor: no fallback


@pytest.mark.parametrize("qtype", ["Q8_0", "Q4_0", "Q2_K", "Q3_K", "Q4_K", "Q6_K"])
def test_full_size_roundtrip_property(qtype):
    """BASELINE config 1 scale (SmolLM2 ffn_down 576x1536 x 30 layers at once): size-independent
    properties — pack is deterministic, dequant(pack(x)) is within the format's step of x, and
    re-packing the dequantized tensor reproduces the same codes for the 8-bit format."""
    from quantool_b200 import cabi
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn((30 * 576, 1536), generator=g, device="cuda") * 0.02).half()
    y1 = cabi.gguf_quantize(x, qtype)
    y2 = cabi.gguf_quantize(x, qtype)
    assert torch.equal(y1, y2)
    d = cabi.gguf_dequantize(y1, qtype, 1536)
    err = (d - x.float()).abs().max().item()
    amax = x.float().abs().max().item()
    tol = {"Q8_0": 1 / 127, "Q4_0": 1 / 7, "Q2_K": 0.8, "Q3_K": 0.4, "Q4_K": 1 / 7, "Q6_K": 1 / 30}[qtype]
    assert err <= tol * amax, (err, amax)
    if qtype == "Q8_0":
        y3 = cabi.gguf_quantize(d, qtype, round_via_f16=False)
        # scales may differ by fp16 rounding of d, codes must survive the round trip
        a = y1.view(-1, 34)[:, 2:]
        b = y3.view(-1, 34)[:, 2:]
        assert (a != b).float().mean().item() < 1e-3


@pytest.mark.parametrize("qtype", ["Q8_0", "Q4_0", "Q5_1", "IQ4_NL", "Q3_K", "Q4_K", "Q6_K"])
def test_batch_launch_equals_per_tensor(qtype):
    """Many tensors in one launch (segment table) == one launch per tensor, including an empty tensor, tensors
    smaller than a CTA tile and tensors that end in the middle of one."""
    from quantool_b200 import cabi
    be = cabi.gguf_block_elems(qtype)
    g = torch.Generator(device="cuda").manual_seed(9)
    shapes = [(3, be), (0, 2 * be), (17, 5 * be), (1, be), (300, 9 * be), (64, 4 * be)]
    xs = [(torch.randn(s, generator=g, device="cuda") * 0.05).half() for s in shapes]
    ys = cabi.gguf_quantize_batch(xs, qtype)
    for x, y in zip(xs, ys):
        assert torch.equal(y, cabi.gguf_quantize(x, qtype))
    many = cabi.gguf_quantize_many([(x, qtype) for x in xs] + [(xs[2].float(), "Q8_0")])
    assert torch.equal(many[2], ys[2]) and torch.equal(many[-1], cabi.gguf_quantize(xs[2].float(), "Q8_0"))
    with pytest.raises(cabi.QtError):
        cabi.gguf_quantize_batch([xs[0], xs[2].float()], qtype)      # mixed dtypes
