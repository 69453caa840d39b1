"""NCCL sharded-vs-unsharded parity of the CUDA GPTQ path (VERDICT r01 "Next" 1b).  Launch with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tests/dist_parity_nccl.py [--out profiles/xxx.json]
(tests/test_dist_nccl_gpu.py does that when the box has >= 2 GPUs.)

Two checks, each on a tiny-but-wide Llama (hidden 1024, intermediate 2816: both tensor-core-chain widths):
  A. layer level, IDENTICAL all-reduced H on both sides: `GPTQLayerQuantizer.quantize_layer` with rows sharded over
     the ranks (owner-computes chain + broadcast, all-gathered rows) must give bit-identical `weight_packed`,
     `weight_scale` and `weight_g_idx` to the unsharded run of the same kernels on every rank.
  B. whole model through `quantize_model_gptq`: samples sharded + NCCL all-reduce(H) vs one rank doing everything.
     H now differs in fp32 summation order: without act_order >= 99.5 % of layer 0's codes must agree (later layers
     see inputs that already went through slightly different codes: objective only); with actorder=group
     (permutation = argsort of a nearly flat diagonal on random-init weights) the summed GPTQ loss must agree to 1 %.
"""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    world = dist.get_world_size()
    from quantool_b200.engine import llama, pipeline, schemes
    from quantool_b200.engine.gptq import compress_linear

    shape = llama.LlamaShape(1024, 2816, 2, 8, 4, 2048)
    args = schemes.resolve("W4A16", "group")
    sharded, single = pipeline.Dist(), pipeline.Dist(enabled=False)
    report = {"world": world, "shape": "hidden 1024 / intermediate 2816 / 2 layers", "scheme": "W4A16 g128 actorder=group"}

    # ---- A: one layer, identical H ---------------------------------------------------------------------
    n_total, seq = 8 * world, 512
    per = pipeline.row_split(n_total, world)
    dims = shape.input_dims()
    acts = {}
    for i, (n, k) in enumerate(dims.items()):
        g = torch.Generator(device=dev).manual_seed(1000 * rank + i)
        x = torch.randn((per[rank] * seq, k), device=dev, generator=g, dtype=torch.float32)
        x[:, :: 97] *= 12.0
        acts[n] = x.to(torch.bfloat16)
    hess = pipeline.accumulate_layer_hessians(acts, per[rank], n_total, sharded)       # all-reduced: same on all ranks
    w = llama.random_layer_weights(shape, 0, dev)
    res_sh = pipeline.GPTQLayerQuantizer(args, dist=sharded).quantize_layer(w, hess)
    res_1 = pipeline.GPTQLayerQuantizer(args, dist=single).quantize_layer(w, hess)
    ident = {}
    for lin in llama.LINEARS:
        a_sh, _ = compress_linear(res_sh[lin].weight, res_sh[lin].scale, res_sh[lin].zero_point, res_sh[lin].g_idx, args)
        a_1, _ = compress_linear(res_1[lin].weight, res_1[lin].scale, res_1[lin].zero_point, res_1[lin].g_idx, args)
        ident[lin] = all(torch.equal(a_sh[k].cpu(), a_1[k].cpu()) for k in ("weight_packed", "weight_scale", "weight_g_idx"))
    flags = torch.tensor([int(all(ident.values()))], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    report["layer_identical_given_same_H"] = {"per_linear": ident, "all_ranks": bool(flags.item())}

    # ---- B: whole model ---------------------------------------------------------------------------------
    # Samples sharded + NCCL all-reduce(H) vs one rank doing everything: H differs in fp32 summation order.  Without
    # act_order that moves a few codes (asserted >= 99 % equal over both layers).  With actorder=group the
    # permutation is argsort(diag H): on random-init weights the diagonal is nearly flat, every near-tie can swap,
    # every swap across a group boundary changes that group's scale - so there the OBJECTIVE is asserted (summed
    # GPTQ loss within 1 %), the code agreement is reported.
    host_sd = llama.random_state_dict(shape, seed=3)
    g = torch.Generator().manual_seed(1234)
    ids = torch.randint(0, shape.vocab_size, (n_total, seq), generator=g)

    def both(a_):
        r_sh = pipeline.quantize_model_gptq(shape, host_sd, ids, a_, dev, dist=sharded)
        out = None
        if rank == 0:
            r_1 = pipeline.quantize_model_gptq(shape, host_sd, ids, a_, dev, dist=single)
            same, tot = {}, {}
            h_sh, h_1 = hashlib.sha256(), hashlib.sha256()
            for k in sorted(r_1.tensors):
                if k.endswith("weight_packed"):
                    p1, p2 = r_1.tensors[k], r_sh.tensors[k]
                    lay = int(k.split(".")[2])
                    for sft in range(0, 32, 4):
                        same[lay] = same.get(lay, 0) + int((((p1 >> sft) & 15) == ((p2 >> sft) & 15)).sum())
                    tot[lay] = tot.get(lay, 0) + p1.numel() * 8
                    h_1.update(p1.numpy().tobytes())
                    h_sh.update(p2.numpy().tobytes())
            out = {"code_agreement": sum(same.values()) / sum(tot.values()),
                   "code_agreement_per_layer": [same[l] / tot[l] for l in sorted(tot)],
                   "artifact_sha_sharded": h_sh.hexdigest()[:16],
                   "artifact_sha_single": h_1.hexdigest()[:16],
                   "gptq_loss_sharded_vs_single": [sum(r_sh.losses.values()), sum(r_1.losses.values())]}
        dist.barrier()
        return out

    rb_plain = both(schemes.resolve("W4A16"))
    rb_act = both(args)
    if rank == 0:
        report["model_no_actorder"] = rb_plain
        report["model_actorder_group"] = rb_act
        print(json.dumps(report), flush=True)
        if a.out:
            with open(os.path.join(ROOT, a.out), "w") as f:
                json.dump(report, f, indent=1)
    dist.barrier()
    ok = report["layer_identical_given_same_H"]["all_ranks"]
    if rank == 0:
        l_sh, l_1 = rb_act["gptq_loss_sharded_vs_single"]
        p_sh, p_1 = rb_plain["gptq_loss_sharded_vs_single"]
        # layer 0 sees identical inputs on both sides (H differs by summation order only); from layer 1 on the inputs
        # themselves differ (they went through layer 0's slightly different codes), so only the objective is asserted
        ok = (ok and rb_plain["code_agreement_per_layer"][0] >= 0.995 and abs(p_sh - p_1) <= 0.01 * abs(p_1)
              and abs(l_sh - l_1) <= 0.01 * abs(l_1))
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(f"sharded / unsharded parity FAILED: {report}")


if __name__ == "__main__":
    main()
