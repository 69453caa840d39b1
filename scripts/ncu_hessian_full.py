"""The bench's roofline kernel at its exact shape (X[262144,14336] bf16 -> H) for one `ncu --set full` capture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
K, T = 14336, 262144
x = torch.randn((T, K), device="cuda", dtype=torch.bfloat16)
H = torch.zeros((K, K), device="cuda")
for _ in range(2):
    cabi.hessian_accumulate(x, H)
torch.cuda.synchronize()
print("ok")
