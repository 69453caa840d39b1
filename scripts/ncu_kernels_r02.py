"""Round-2 ncu set: one or two launches of every kernel added or reworked this round, at the shapes the product path
runs them at.  Not a benchmark: numbers printed under ncu are never bench values.
    ncu --set full --clock-control none --import-source on -k regex:'potrf_inv|panel_top|sgemm_nt_k128|gptq_block128|tgemm|scale_qdq16|dequant_staged|pack_k45|pack_q6k|pack_iq4nl|tri_pack' \
        -o gpurun_out/r02_ncu python scripts/ncu_kernels_r02.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi

dev = "cuda"
torch.manual_seed(0)

# 1. inverse-Hessian chain at K = 512 (4 panels): potrf_inv, panel_top, the skinny Kd = 128 sgemm, tgemm (small)
K = 512
x = torch.randn((4096, K), device=dev, dtype=torch.bfloat16)
H = torch.zeros((K, K), device=dev)
cabi.hessian_accumulate(x, H)
cabi.hessian_finalize(H, 2.0 / 8)
Hf, dead = cabi.gptq_prepare_hessian(H, None, 0.01)
U, info = cabi.gptq_hinv_factor(Hf, tensor_core=True)
# the skinny sgemm at a mid-chain TRSM shape of K = 14336 (7168 rows below the panel)
P = torch.randn((7168, 128), device=dev)
Xi = torch.tril(torch.randn((128, 128), device=dev))
out = torch.empty_like(P)
cabi.sgemm(P, Xi, out, b_is_nk=True)

# 2. the lean GPTQ block kernel: N = 4096 (latency regime) and N = 14336 (issue regime), one 128-column block + lazy update
for N in (4096, 14336):
    wp = torch.randn((N, 256), device=dev) * 0.02
    Hs, _ = cabi.gptq_prepare_hessian(H[:256, :256].contiguous(), None, 0.01)
    Us, _ = cabi.gptq_hinv_factor(Hs)
    scale = torch.empty((N, 2), device=dev)
    zp = torch.empty_like(scale)
    cabi.gptq_quantize_weight(wp, Us, scale, zp, None, 128, 4, True, 0, U_split=cabi.split_tf32_transpose(Us))

# 3. error-feedback GEMM (3xTF32) at the column loop's shapes: Kd = 512 deferred update, Kd = 128 in-block update
M, Kfull = 4096, 14336
for Kd, Ncols in ((512, Kfull - 512), (128, 384)):
    A = torch.randn((M, Kd), device=dev)
    Bt = torch.randn((Ncols, Kd), device=dev)
    C = torch.zeros((M, Ncols), device=dev)
    cabi.gemm_tf32x3(cabi.split_tf32(A), cabi.split_tf32(Bt), C, negate=True, accumulate=True)
del A, Bt, C

# 4. AWQ: 16-columns-per-lane scale -> qdq kernel and the Gram-form loss GEMM (bf16 tcgen05, reducing epilogue)
W = (torch.randn((4096, 14336), device=dev) * 0.02).to(torch.bfloat16)
s = torch.rand((14336,), device=dev) + 0.5
cabi.awq_scale_qdq(W, s, 128, 4, True)
d16 = torch.empty(W.shape, dtype=torch.bfloat16, device=dev)
d32 = torch.empty(W.shape, dtype=torch.float32, device=dev)
cabi.awq_scale_qdq_delta(W, s, 128, 4, True, d16, d32)
G = torch.randn((14336, 14336), device=dev).to(torch.bfloat16)
acc = torch.zeros((1,), dtype=torch.float64, device=dev)
cabi.awq_gram_loss(d16, d32, G, acc)
del W, d16, d32, G

# 5. GGUF: K-quant packers after the op-count reduction, IQ4_NL with the table in shared memory, staged dequantize
xg = (torch.randn((4096, 14336), device=dev) * 0.02).half()
for t in ("Q4_K", "Q6_K", "IQ4_NL", "Q8_0", "Q4_0"):
    y = cabi.gguf_quantize(xg, t)
    cabi.gguf_dequantize(y, t, 14336)

# 6. packed triangle (what crosses NVLink)
Hb = torch.randn((14336, 14336), device=dev)
p = cabi.tri_pack(Hb, 256)
cabi.tri_unpack(p, Hb, 256)
torch.cuda.synchronize()
print("ok", int(info.item()), float(acc.item()) != 0.0)
