"""Calibration-forward timing at the 8B shape: fused kernels vs plain torch ops (one 16x2048 chunk)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200.engine import llama
from oracle import llama_forward as lf
shape = llama.SHAPES["llama-3-8b"]
w = llama.random_layer_weights(shape, 0, "cuda")
B, S = 16, 2048
h = torch.randn((B, S, shape.hidden_size), device="cuda").to(torch.bfloat16)
cos, sin = llama.rope_tables(shape, S, "cuda", h.dtype)
cap = {n: torch.empty((B * S, k), dtype=h.dtype, device="cuda") for n, k in shape.input_dims().items()}
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
print("torch  pass1 (capture, full layer):", t(lambda: lf.layer_forward(shape, w, h, cos, sin, capture=cap)))
print("fused  pass1 (capture, stop at down_in):", t(lambda: llama.layer_forward(shape, w, h, cos, sin, capture=cap, stop_after="down_in")))
print("torch  pass2:", t(lambda: lf.layer_forward(shape, w, h, cos, sin)))
print("fused  pass2:", t(lambda: llama.layer_forward(shape, w, h, cos, sin)))
x = h
print("gemm only (q,k,v,o,gate,up,down):", t(lambda: [torch.nn.functional.linear(x, w[f"{n}.weight"]) for n in llama.LINEARS if "down" not in n] + [torch.nn.functional.linear(cap["down_in"].view(B, S, -1), w["mlp.down_proj.weight"])]))
