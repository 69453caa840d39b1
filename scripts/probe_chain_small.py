import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
K = int(os.environ.get("K", 4096))
x = torch.randn((8192, K), device="cuda", dtype=torch.bfloat16)
H = torch.zeros((K,K), device="cuda"); cabi.hessian_accumulate(x, H); cabi.hessian_finalize(H, 2/128)
X = torch.empty_like(H); W = torch.empty_like(H)
for _ in range(2):
    Hf, dead = cabi.gptq_prepare_hessian(H, None, 0.01); U, info = cabi.gptq_hinv_factor(Hf, X, W)
torch.cuda.synchronize()
N=4096
w = (torch.randn((N,K), device="cuda")*0.02).to(torch.bfloat16)
wp = cabi.gptq_permute_in(w, None, dead); scale = torch.empty((N,K//128), device="cuda"); zp = torch.empty_like(scale)
cabi.gptq_quantize_weight(wp, U, scale, zp, None, 128, 4, True, 0)
torch.cuda.synchronize()
print("ok", int(info.item()))
