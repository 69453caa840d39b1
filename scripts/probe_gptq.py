"""Kernel-level timing probe at Llama-3-8B shapes (CUDA events, L2-cold inputs > 126 MB)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
from quantool_b200.engine import gptq as eg, schemes

def timeit(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)

out = {}
T = int(os.environ.get("T", 262144))
for K in (4096, 14336):
    x = torch.randn((T, K), device="cuda", dtype=torch.bfloat16)
    H = torch.zeros((K, K), device="cuda")
    for S in ([0] if K == 14336 else [0, 8, 16, 37]):
        cabi.lib().qt_hessian_set_splits(S)
        best, avg = timeit(lambda: cabi.hessian_accumulate(x, H), n=3)
        out[f"hessian_K{K}_S{S}"] = dict(ms=best, tflops_ref=2 * T * K * K / best / 1e9, tflops_exec=None)
        print(f"hessian K={K} T={T} S={S}: {best:.2f} ms  ref-equivalent {2*T*K*K/best/1e9:.0f} TFLOP/s", flush=True)
    cabi.lib().qt_hessian_set_splits(0)
    best, _ = timeit(lambda: torch.matmul(x.t(), x), n=3)
    print(f"  cuBLAS x^T x bf16 (full GEMM): {best:.2f} ms {2*T*K*K/best/1e9:.0f} TFLOP/s", flush=True)
    cabi.hessian_finalize(H, 2.0 / 128)
    del x
    # hinv chain
    perm = torch.argsort(torch.diagonal(H), descending=True, stable=True).to(torch.int32)
    X = torch.empty_like(H); W = torch.empty_like(H)
    def chain():
        Hf, dead = cabi.gptq_prepare_hessian(H, perm, 0.01)
        cabi.gptq_hinv_factor(Hf, X, W)
    best, _ = timeit(chain, n=2)
    print(f"hinv chain K={K}: {best:.2f} ms  ({(2/3)*K**3/best/1e9:.1f} TFLOP/s fp32 on (2/3)K^3)", flush=True)
    out[f"hinv_K{K}"] = dict(ms=best)
    Hf, dead = cabi.gptq_prepare_hessian(H, perm, 0.01)
    U, info = cabi.gptq_hinv_factor(Hf, X, W)
    print("  info", int(info.item()))
    def torch_chain():
        Hd = H + 0.01 * torch.diagonal(H).mean() * torch.eye(K, device="cuda")
        L = torch.linalg.cholesky(Hd); Hi = torch.cholesky_inverse(L); torch.linalg.cholesky(Hi, upper=True)
    best, _ = timeit(torch_chain, n=2)
    print(f"  torch/cuSOLVER chain: {best:.2f} ms", flush=True)
    del X, W
    args = schemes.resolve("W4A16", "group")
    for N in ((4096, 1024) if K == 4096 else (4096,)):
        w = (torch.randn((N, K), device="cuda") * 0.02).to(torch.bfloat16)
        wp = cabi.gptq_permute_in(w, perm, dead)
        scale = torch.empty((N, K // 128), device="cuda"); zp = torch.empty_like(scale)
        def loop():
            wq = wp.clone()
            cabi.gptq_quantize_weight(wq, U, scale, zp, None, 128, 4, True, 0)
        best, _ = timeit(loop, n=2)
        print(f"gptq column loop N={N} K={K}: {best:.2f} ms ({N*K*K/best/1e9:.1f} TFLOP/s on N K^2)", flush=True)
        out[f"loop_N{N}_K{K}"] = dict(ms=best)
    del H, U, Hf
    torch.cuda.empty_cache()
# sgemm peak
M = 8192
A = torch.randn((M, M), device="cuda"); B = torch.randn((M, M), device="cuda"); C = torch.zeros((M, M), device="cuda")
best, _ = timeit(lambda: cabi.sgemm(A, B, C), n=3)
print(f"sgemm NN 8192^3: {best:.2f} ms {2*M**3/best/1e9:.1f} TFLOP/s")
best, _ = timeit(lambda: cabi.sgemm(A, B, C, b_is_nk=True), n=3)
print(f"sgemm NT 8192^3: {best:.2f} ms {2*M**3/best/1e9:.1f} TFLOP/s")
torch.backends.cuda.matmul.allow_tf32 = False
best, _ = timeit(lambda: torch.matmul(A, B, out=C), n=3)
print(f"cuBLAS fp32 8192^3: {best:.2f} ms {2*M**3/best/1e9:.1f} TFLOP/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_gptq.json", "w"), indent=1)
