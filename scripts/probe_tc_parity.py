"""Code agreement with the CPU oracle at larger K: FFMA chain vs tensor-core chain."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200.engine import gptq as eg, schemes
from oracle import gptq as og
from compressed_tensors.quantization import ActivationOrdering
for N, K, act in ((64, 2048, "group"), (32, 4096, "group"), (32, 4096, None)):
    g = torch.Generator().manual_seed(N * 1000 + K)
    W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
    T = 4 * K
    x = torch.randn((T, K), generator=g); x[:, ::97] *= 8; x = x.to(torch.bfloat16)
    oargs = og.scheme_weight_args("W4A16")
    if act: oargs.actorder = ActivationOrdering.GROUP
    t0 = time.time()
    Ho, n = og.make_empty_hessian(K), 0
    for xb in x.reshape(8, T // 8, K):
        Ho, n = og.accumulate_hessian(xb.unsqueeze(0), Ho, n)
    loss_o, Wq_o, s_o, z_o, gi_o = og.quantize_weight(W, Ho, oargs)
    codes_o, _, _ = og.compress_packed(Wq_o, s_o, None, gi_o, oargs)
    print(f"N={N} K={K} act={act}: oracle {time.time()-t0:.1f}s")
    args = schemes.resolve("W4A16", act)
    acc = eg.HessianAccumulator(K, "cuda")
    for xb in x.reshape(8, T // 8, K):
        acc.add(xb.unsqueeze(0).cuda())
    H = acc.finalize()
    e_o = og.layer_error(W, Wq_o, x.float())
    res = {}
    for tc in (False, True):
        r = eg.quantize_linear(W.cuda(), H, args, tensor_core_chain=tc)
        _, codes = eg.compress_linear(r.weight, r.scale, r.zero_point, r.g_idx, args)
        res[tc] = codes.cpu()
        e_c = og.layer_error(W, r.weight.cpu(), x.float())
        print(f"  tensor_core_chain={tc}: agreement with oracle {(res[tc] == codes_o).float().mean().item():.5f}  objective rel diff {(e_c-e_o)/e_o:+.2e}")
    print(f"  TC vs FFMA agreement {(res[True] == res[False]).float().mean().item():.5f}")
# full-size agreement between the two chains (no oracle): down_proj-like slice
for N, K in ((256, 14336),):
    g = torch.Generator(device="cuda").manual_seed(7)
    W = (torch.randn((N, K), generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    x = torch.randn((2 * K, K), generator=g, device="cuda"); x[:, ::97] *= 8; x = x.to(torch.bfloat16)
    acc = eg.HessianAccumulator(K, "cuda"); acc.add(x.unsqueeze(0)); H = acc.finalize()
    args = schemes.resolve("W4A16", "group")
    res = {}
    for tc in (False, True):
        r = eg.quantize_linear(W, H, args, tensor_core_chain=tc)
        _, codes = eg.compress_linear(r.weight, r.scale, r.zero_point, r.g_idx, args)
        res[tc] = (codes, r.losses.sum().item())
    print(f"N={N} K={K}: TC vs FFMA code agreement {(res[True][0] == res[False][0]).float().mean().item():.5f}  loss {res[False][1]:.6e} vs {res[True][1]:.6e}")
