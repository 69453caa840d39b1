"""One GGUF packer launch loop for ncu: python scripts/probe_gguf_one.py Q4_0 [quantize|dequantize]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
t = sys.argv[1] if len(sys.argv) > 1 else "Q4_0"
op = sys.argv[2] if len(sys.argv) > 2 else "quantize"
n, k = 16384, 14336
x = (torch.randn((n, k), device="cuda") * 0.02).half()
y = cabi.gguf_quantize(x, t)
for _ in range(4):
    if op == "quantize": cabi.gguf_quantize(x, t, out=y)
    else: cabi.gguf_dequantize(y, t, k)
torch.cuda.synchronize()
print("ok")
