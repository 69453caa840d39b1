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
for K in (4096, 14336):
    g = torch.Generator(device="cuda").manual_seed(K)
    X = torch.randn((4 * K if K < 8192 else 2 * K, K), generator=g, device="cuda")
    X[:, ::97] *= 8
    H = (X.t() @ X) * (2.0 / 128); del X
    perm = torch.argsort(torch.diagonal(H), descending=True, stable=True).to(torch.int32)
    Xs, W = torch.empty_like(H), torch.empty_like(H)
    ws = [torch.empty_like(H) for _ in range(4)]
    res = {}
    for tc in (False, 1, 2, True):
        def run():
            Hf, dead = cabi.gptq_prepare_hessian(H, perm, 0.01)
            U, info = cabi.gptq_hinv_factor(Hf, Xs, W, tensor_core=bool(tc), workspace=ws, stages=3 if tc is True else int(tc))
            return U, info
        ms = t(run)
        U, info = run()
        res[tc] = U.clone()
        print(f"K={K} tensor_core={tc}: {ms:.2f} ms info={int(info.item())}  rel diff vs FFMA {((U - res[False]).norm() / res[False].norm()).item():.3e}")
    if K <= 4096:
        p = perm.cpu().long()
        Hd = H.cpu().double()[p][:, p]
        Hd += 0.01 * torch.mean(torch.diag(Hd)) * torch.eye(K, dtype=torch.float64)
        ref = torch.linalg.cholesky(torch.cholesky_inverse(torch.linalg.cholesky(Hd)), upper=True)
        for tc in (False, 1, 2, True):
            print(f"  vs fp64, tensor_core={tc}: {((res[tc].cpu().double() - ref).norm() / ref.norm()).item():.3e}")
