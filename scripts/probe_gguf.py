import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
n, k = 16384, 14336
x = (torch.randn((n, k), device="cuda") * 0.02).half()
for t in ("Q8_0", "Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"):
    be, bb = cabi.gguf_block_elems(t), cabi.gguf_block_bytes(t)
    y = torch.empty((n, k // be * bb), dtype=torch.uint8, device="cuda")
    for _ in range(3): cabi.gguf_quantize(x, t, out=y)
    torch.cuda.synchronize(); ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cabi.gguf_quantize(x, t, out=y); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts)//2]; byts = x.numel()*2 + y.numel()
    print(f"{t}: {ms:.3f} ms  {byts/ms/1e6:.0f} GB/s  ({byts/ms/1e6/6533.5*100:.0f}% of measured HBM peak)")
for t in ("Q8_0", "Q4_0", "Q5_0", "Q2_K", "Q3_K", "Q4_K", "Q6_K"):
    y = cabi.gguf_quantize(x, t)
    for _ in range(2): cabi.gguf_dequantize(y, t, k)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); cabi.gguf_dequantize(y, t, k); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts)//2]; byts = x.numel()*4 + y.numel()
    print(f"dequant {t}: {ms:.3f} ms  {byts/ms/1e6:.0f} GB/s  ({byts/ms/1e6/6533.5*100:.0f}% of measured HBM peak)")
