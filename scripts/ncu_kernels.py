"""One or two launches of every hot kernel at a representative (L2-exceeding where it matters) size,
for `ncu --set full`.  Not a benchmark: numbers printed under ncu are never bench values."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
dev = "cuda"
torch.manual_seed(0)
# Hessian SYRK (K = 4096, T = 65536: 512 MB of X)
K, T = 4096, 65536
x = torch.randn((T, K), device=dev, dtype=torch.bfloat16)
H = torch.zeros((K, K), device=dev)
cabi.hessian_accumulate(x, H); cabi.hessian_finalize(H, 2.0 / 32)
del x
# lazy-batch update on the tensor cores, one full-width block step (M = 4096 rows, K = 4096)
U = torch.triu(torch.randn((K, K), device=dev)) * 0.05
uh, ul = cabi.split_tf32_transpose(U)
err = torch.randn((4096, 128), device=dev); eh, el = cabi.split_tf32(err)
Wf = torch.randn((4096, K), device=dev)
cabi.gptq_lazy_update_tf32x3(eh, el, uh, ul, Wf, 0, 128)
cabi.sgemm(err, U[:128, 128:], Wf[:, 128:], alpha=-1.0, beta=1.0)          # the FFMA GEMM it replaces
# small chain + column loop: 2 potrf launches, 2 block launches (N = 4096 rows)
Ks = 256
Hs = H[:Ks, :Ks].contiguous()
Hf, dead = cabi.gptq_prepare_hessian(Hs, None, 0.01)
Us, info = cabi.gptq_hinv_factor(Hf)
wp = torch.randn((4096, Ks), device=dev) * 0.02
scale = torch.empty((4096, Ks // 128), device=dev); zp = torch.empty_like(scale)
cabi.gptq_quantize_weight(wp, Us, scale, zp, None, 128, 4, True, 0)
wq = torch.randn((4096, K), device=dev).to(torch.bfloat16) * 0.02
sc = (torch.rand((4096, K // 128), device=dev) * 0.01 + 0.002).to(torch.bfloat16)
codes, _ = cabi.quantize_codes(wq, sc, None, None, 128, 4)
cabi.pack_int32(codes, 4)
# HBM streaming kernels on > L2 inputs
big = torch.randn((65536, 4096), device=dev, dtype=torch.bfloat16)       # 512 MB
mn, mx = cabi.new_minmax(4096, dev); cabi.channel_minmax(big, mn, mx)
acc = torch.zeros((4096,), device=dev); cabi.channel_abs_sum(big, acc)
del big
xg = (torch.randn((16384, 14336), device=dev) * 0.02).half()            # 470 MB
for t in ("Q8_0", "Q4_0", "Q5_0", "Q4_K", "Q6_K"):
    y = cabi.gguf_quantize(xg, t)
    if t in ("Q8_0", "Q4_K"):
        cabi.gguf_dequantize(y, t, 14336)
wt = (torch.randn((14336, 4096), device=dev) * 0.02).to(torch.bfloat16)
s = torch.rand((4096,), device=dev) + 0.5
cabi.awq_scale_qdq(wt, s, 128, 4, True)
cabi.scale_matrix_(wt, s)
torch.cuda.synchronize()
print("ok", int(info.item()))
