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
T=262144
for K in (4096, 14336):
    x = torch.randn((T, K), device="cuda", dtype=torch.bfloat16); H = torch.zeros((K,K), device="cuda")
    NI, NJ = (K+127)//128, (K+255)//256
    nt = sum(min(2*j+2, NI) for j in range(NJ))
    for S in ([0, 16, 24] if K==14336 else [0, 4, 8]):
        cabi.lib().qt_hessian_set_splits(S)
        ms = timeit(lambda: cabi.hessian_accumulate(x, H))
        print(f"hessian K={K} S={S}: {ms:.2f} ms exec {2*T*nt*128*256/ms/1e9:.0f} TF/s ref-eq {2*T*K*K/ms/1e9:.0f}", flush=True)
    cabi.lib().qt_hessian_set_splits(0)
    # correctness spot check vs reference on a slice
    xs = x[:4096].contiguous(); H1 = torch.zeros((K,K), device="cuda"); cabi.hessian_accumulate(xs, H1); cabi.hessian_finalize(H1, 1.0)
    ref = xs.float().t() @ xs.float()
    print("  rel err vs torch fp32:", (torch.linalg.norm(H1-ref)/torch.linalg.norm(ref)).item())
    del x, H, H1, ref
M=8192
A = torch.randn((M, M), device="cuda"); B = torch.randn((M, M), device="cuda"); C = torch.zeros((M, M), device="cuda")
print(f"sgemm NN 8192^3: {2*M**3/timeit(lambda: cabi.sgemm(A, B, C))/1e9:.1f} TFLOP/s; NT: {2*M**3/timeit(lambda: cabi.sgemm(A, B, C, b_is_nk=True))/1e9:.1f}")
