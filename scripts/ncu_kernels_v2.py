"""Second ncu set (kernels added after the first capture): the 3xTF32 chain GEMM, the warp-autonomous GGUF
packers, Q2_K/Q3_K, the calibration-forward kernels, the 36-row GPTQ block kernel, the exact-diagonal pass.
Not a benchmark: numbers printed under ncu are never bench values."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
dev = "cuda"
torch.manual_seed(0)
# chain GEMM: triangular-inverse style product (Kd = 4096, store) and a deferred Cholesky update (Kd = 512, lower, +=)
n = 8192
A = torch.randn((n, 4096), device=dev); B = torch.randn((n, 4096), device=dev); C = torch.zeros((n, n), device=dev)
sa, sb = cabi.split_tf32(A), cabi.split_tf32(B)
cabi.gemm_tf32x3(sa, sb, C)
P = (sa[0][:, :512], sa[1][:, :512])
cabi.gemm_tf32x3(P, P, C, negate=True, accumulate=True, lower_tiles_only=True)
del A, B, C, sa, sb, P
# GPTQ block kernel at N = 14336 (36-row CTAs) and the exact-diagonal pass
K = 4096
x = torch.randn((32768, K), device=dev, dtype=torch.bfloat16)
H = torch.zeros((K, K), device=dev); diag = torch.zeros((K,), device=dev); scr = torch.empty((32 * K,), device=dev)
cabi.hessian_accumulate(x, H); cabi.hessian_diag_accumulate(x, diag, scr); cabi.hessian_set_diagonal(H, diag)
cabi.hessian_finalize(H, 2.0 / 16)
Hf, dead = cabi.gptq_prepare_hessian(H[:256, :256].contiguous(), None, 0.01)
Us, info = cabi.gptq_hinv_factor(Hf)
wp = torch.randn((14336, 256), device=dev) * 0.02
scale = torch.empty((14336, 2), device=dev); zp = torch.empty_like(scale)
cabi.gptq_quantize_weight(wp, Us, scale, zp, None, 128, 4, True, 0)
# calibration-forward kernels on a 16 x 2048-token chunk of the 8B shape
h = torch.randn((16, 2048, 4096), device=dev).to(torch.bfloat16)
w = torch.ones((4096,), device=dev, dtype=torch.bfloat16)
cabi.rms_norm(h, w, 1e-5)
from quantool_b200.engine import llama
cos, sin = llama.rope_tables(llama.SHAPES["llama-3-8b"], 2048, dev, torch.bfloat16)
cabi.rope_(h, cos, sin, 2048, 32, 128)
g = torch.randn((16, 2048, 14336), device=dev).to(torch.bfloat16)
cabi.silu_mul(g, g)
del h, g, x
# GGUF: reworked 4/5-bit packers and the new K types, > L2 input
xg = (torch.randn((16384, 14336), device=dev) * 0.02).half()
for t in ("Q4_0", "Q5_0", "Q4_1", "Q2_K", "Q3_K"):
    y = cabi.gguf_quantize(xg, t)
    if t in ("Q2_K", "Q3_K"):
        cabi.gguf_dequantize(y, t, 14336)
torch.cuda.synchronize()
print("ok", int(info.item()))
