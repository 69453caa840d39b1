"""Where does the act_order disagreement with the oracle at K=4096 come from: H (perm ties) or the solve?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200.engine import gptq as eg, schemes
from oracle import gptq as og
from compressed_tensors.quantization import ActivationOrdering
N, K = 32, 4096
g = torch.Generator().manual_seed(N * 1000 + K)
W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
T = 4 * K
x = torch.randn((T, K), generator=g); x[:, ::97] *= 8; x = x.to(torch.bfloat16)
oargs = og.scheme_weight_args("W4A16"); oargs.actorder = ActivationOrdering.GROUP
Ho, n = og.make_empty_hessian(K), 0
for xb in x.reshape(8, T // 8, K):
    Ho, n = og.accumulate_hessian(xb.unsqueeze(0), Ho, n)
H64 = (2.0 / 8) * (x.double().t() @ x.double())
loss_o, Wq_o, s_o, z_o, gi_o = og.quantize_weight(W, Ho.clone(), oargs)
codes_o, _, _ = og.compress_packed(Wq_o, s_o, None, gi_o, oargs)
args = schemes.resolve("W4A16", "group")
acc = eg.HessianAccumulator(K, "cuda")
for xb in x.reshape(8, T // 8, K):
    acc.add(xb.unsqueeze(0).cuda())
H = acc.finalize()
def rel(a, b): return ((a.double() - b).norm() / b.norm()).item()
print("H rel err vs fp64: gpu", rel(H.cpu(), H64), " oracle(cpu fp32)", rel(Ho, H64))
dg, do, d64 = torch.diagonal(H.cpu()).double(), torch.diagonal(Ho).double(), torch.diagonal(H64)
print("diag max rel err: gpu", ((dg - d64).abs() / d64).max().item(), " oracle", ((do - d64).abs() / d64).max().item(),
      " gpu mean signed", ((dg - d64) / d64).mean().item())
pg = torch.argsort(dg, descending=True, stable=True); po = torch.argsort(do, descending=True, stable=True)
print("perm positions that differ:", (pg != po).sum().item(), "of", K)
for name, Hin in (("gpu H", H), ("oracle H", Ho.cuda())):
    r = eg.quantize_linear(W.cuda(), Hin, args)
    _, codes = eg.compress_linear(r.weight, r.scale, r.zero_point, r.g_idx, args)
    print(f"{name}: code agreement with oracle {(codes.cpu() == codes_o).float().mean().item():.5f}")
