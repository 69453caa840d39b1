"""Where do GPTQ code disagreements with the CPU oracle come from at BASELINE widths?  For each (N, K, T):
  floor      oracle(H) vs oracle(H'), H' = the same sums accumulated in fp64 and rounded once (rel. diff ~5e-7):
             the restated reference's own sensitivity to fp32 rounding of H
  gpu        CUDA path, own tensor-core Hessian, default (3xTF32 chain + 3xTF32 lazy update)
  gpu_ffma   CUDA path, own Hessian, FFMA chain + FFMA lazy update (no tensor-core products after the Hessian)
  gpu_Ho     CUDA path fed the oracle's H (default chain)
All with the act_order permutation pinned to the GPU's.  Writes gpurun_out/r02_noise_floor.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200.engine import gptq as eg, schemes
from oracle import gptq as og
from compressed_tensors.quantization import ActivationOrdering


def acts(T, K, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((T, K), generator=g)
    idx = torch.randperm(K, generator=g)[: max(1, K // 200)]
    x[:, idx] *= 20.0
    return x.to(torch.bfloat16)


cases = [(64, 4096, 8192), (64, 4096, 32768), (64, 8192, 16384), (64, 8192, 65536), (32, 14336, 16384), (32, 14336, 65536)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in c.split(",")) for c in sys.argv[1:]]
out = []
for N, K, T in cases:
    t0 = time.time()
    ns = 8
    g = torch.Generator().manual_seed(K + N)
    W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
    x = acts(T, K, K + 11)
    oargs = og.scheme_weight_args("W4A16"); oargs.actorder = ActivationOrdering.GROUP
    args = schemes.resolve("W4A16", "group")
    Ho, n = og.make_empty_hessian(K), 0
    for xb in x.reshape(ns, T // ns, K):
        Ho, n = og.accumulate_hessian(xb.unsqueeze(0), Ho, n)
    xc = x.cuda()
    H64 = torch.zeros((K, K), dtype=torch.float64, device="cuda")
    for xb in xc.split(4096):
        H64 += xb.double().t() @ xb.double()
    H2 = (H64 * (2.0 / ns)).float().cpu()
    del H64
    acc = eg.HessianAccumulator(K, "cuda")
    for xb in xc.reshape(ns, T // ns, K):
        acc.add(xb.unsqueeze(0))
    H = acc.finalize()
    relH = (torch.linalg.norm(H.cpu() - Ho) / torch.linalg.norm(Ho)).item()

    def gpu_codes(Hin, **kw):
        r = eg.quantize_linear(W.cuda(), Hin, args, **kw)
        _, c = eg.compress_linear(r.weight, r.scale, r.zero_point, r.g_idx, args)
        return c.cpu(), r.perm.cpu().long(), r.weight.cpu()

    c_gpu, perm_g, wq_g = gpu_codes(H)
    c_ffma, perm_f, _ = gpu_codes(H, tensor_core_lazy=False, tensor_core_chain=False)
    c_Ho, perm_h, _ = gpu_codes(Ho.cuda())

    def oracle_codes(Hin, perm):
        _, Wq, s, z, gi = og.quantize_weight(W, Hin.clone(), oargs, perm_override=perm)
        c, _, _ = og.compress_packed(Wq, s, None, gi, oargs)
        return c, Wq

    c_o, Wq_o = oracle_codes(Ho, perm_g)
    c_o2, _ = oracle_codes(H2, perm_g)
    c_oh, _ = oracle_codes(Ho, perm_h)
    xs = x[:4096].float()
    rec = {"N": N, "K": K, "T": T, "relH_gpu_vs_oracle": relH,
           "floor_oracle_vs_oracle": (c_o == c_o2).float().mean().item(),
           "gpu": (c_gpu == c_o).float().mean().item(),
           "gpu_ffma": (c_ffma == c_o).float().mean().item() if torch.equal(perm_f, perm_g) else None,
           "gpu_Ho": (c_Ho == c_oh).float().mean().item(),
           "gpu_vs_gpu_ffma": (c_gpu == c_ffma).float().mean().item(),
           "err_gpu": og.layer_error(W, wq_g, xs), "err_oracle": og.layer_error(W, Wq_o, xs),
           "seconds": round(time.time() - t0, 1)}
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del xc, acc, H
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r02_noise_floor.json", "w"), indent=1)
