import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for K in (4096, 14336):
    x = torch.randn((8192 if K==4096 else 16384, K), device="cuda", dtype=torch.bfloat16)
    x[:, :20] *= 20
    H = torch.zeros((K,K), device="cuda"); cabi.hessian_accumulate(x, H); cabi.hessian_finalize(H, 2/128); del x
    perm = torch.argsort(torch.diagonal(H), descending=True, stable=True).to(torch.int32)
    X = torch.empty_like(H); W = torch.empty_like(H)
    def chain():
        Hf, dead = cabi.gptq_prepare_hessian(H, perm, 0.01); cabi.gptq_hinv_factor(Hf, X, W)
    print(f"hinv chain K={K}: {timeit(chain):.2f} ms", flush=True)
    N = 4096
    Hf, dead = cabi.gptq_prepare_hessian(H, perm, 0.01); U, info = cabi.gptq_hinv_factor(Hf, X, W)
    w = (torch.randn((N,K), device="cuda")*0.02).to(torch.bfloat16)
    wp = cabi.gptq_permute_in(w, perm, dead); scale = torch.empty((N,K//128), device="cuda"); zp = torch.empty_like(scale)
    def loop():
        wq = wp.clone(); cabi.gptq_quantize_weight(wq, U, scale, zp, None, 128, 4, True, 0)
    print(f"gptq loop (FFMA lazy) N={N} K={K}: {timeit(loop):.2f} ms", flush=True)
    us = cabi.split_tf32_transpose(U, X, W)
    print(f"  split_transpose: {timeit(lambda: cabi.split_tf32_transpose(U, X, W)):.2f} ms")
    def loop_tc():
        wq = wp.clone(); cabi.gptq_quantize_weight(wq, U, scale, zp, None, 128, 4, True, 0, U_split=us)
    print(f"gptq loop (tcgen05 tf32x3 lazy) N={N} K={K}: {timeit(loop_tc):.2f} ms", flush=True)
    if K == 4096:
        N2 = 14336
        w2 = (torch.randn((N2,K), device="cuda")*0.02).to(torch.bfloat16)
        wp2 = cabi.gptq_permute_in(w2, perm, dead); scale2 = torch.empty((N2,K//128), device="cuda"); zp2 = torch.empty_like(scale2)
        def loop_tc2():
            wq = wp2.clone(); cabi.gptq_quantize_weight(wq, U, scale2, zp2, None, 128, 4, True, 0, U_split=us)
        print(f"gptq loop (tcgen05 tf32x3 lazy) N={N2} K={K}: {timeit(loop_tc2):.2f} ms", flush=True)
    del H, X, W, Hf, U; torch.cuda.empty_cache()
