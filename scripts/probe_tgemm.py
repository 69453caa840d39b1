import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200 import cabi
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for M, N, Kd in ((8192, 8192, 8192), (8192, 8192, 512), (14336, 14336, 512), (14336, 14336, 128)):
    A = torch.randn((M, Kd), device="cuda"); B = torch.randn((N, Kd), device="cuda"); C = torch.zeros((M, N), device="cuda")
    sa, sb = cabi.split_tf32(A), cabi.split_tf32(B)
    for acc in (False, True):
        ms = t(lambda: cabi.gemm_tf32x3(sa, sb, C, negate=True, accumulate=acc))
        print(f"tgemm {M}x{N}x{Kd} acc={acc}: {ms:.3f} ms  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s fp32-equivalent")
    ms = t(lambda: cabi.gemm_tf32x3(sa, sa, C, negate=True, accumulate=True, lower_tiles_only=True)) if M == N else 0
    print(f"  syrk lower: {ms:.3f} ms")
    ms = t(lambda: cabi.sgemm(A, B, C, b_is_nk=True) if hasattr(cabi, 'sgemm') else None)
    print(f"  ffma sgemm: {ms:.3f} ms  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s")
    ms = t(lambda: torch.matmul(A, B.T, out=C))
    print(f"  cublas fp32: {ms:.3f} ms  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s")
