"""Per-layer kernel timeline of the GPTQ hot path (bench.py's `value` step) from CUPTI (torch.profiler sees
every kernel of the process, including the ones launched through the C-ABI).  Works single-GPU or under torchrun.

Writes gpurun_out/<tag>_timeline_rank<r>.json:
  span_ms            first kernel start -> last kernel end of ONE decoder layer
  busy_ms            union of all kernel intervals (GPU not idle)
  by_kernel          name -> {n, total_ms, share_of_span}
  by_stream          stream -> busy_ms
  concurrency        time with exactly 1 / 2 / 3+ kernels in flight
  phases             hessian / chain+loop split (first potrf or gather_flip launch marks the boundary)
  critical           the kernels on the longest stream, i.e. what the layer waits for
Numbers here are taken under the profiler: they say WHERE the time goes, never how fast the step is.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # see quantool_b200/__init__.py
import torch  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from quantool_b200.engine import llama, pipeline, schemes
    from quantool_b200.engine.gptq import compress_linear
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    shape = llama.SHAPES[os.environ.get("MODEL", "llama-3-8b")]
    d = pipeline.Dist()
    args = schemes.resolve("W4A16", "group")
    samples, seq = 128, 2048
    n_local = pipeline.row_split(samples, d.world)[d.rank]
    dims = shape.input_dims()
    acts = {n: bench.synth_acts(n_local * seq, k, dev, 7 + i + 100 * d.rank) for i, (n, k) in enumerate(dims.items())}
    w = llama.random_layer_weights(shape, 0, dev)
    lq = pipeline.GPTQLayerQuantizer(args, dist=d)

    from quantool_b200.engine.gptq import HessianAccumulator
    accs = {n: HessianAccumulator(k, dev) for n, k in dims.items()}

    def layer():
        done = pipeline.accumulate_layer_sums(acts, n_local, accs, dist=d)
        res = lq.quantize_layer(w, None, accs=accs, n_total=samples, acc_events=done)
        for lin, r in res.items():
            compress_linear(r.weight, r.scale, r.zero_point, r.g_idx, args)

    for _ in range(2):
        layer()
    torch.cuda.synchronize()
    if d.on:
        d.dist.barrier()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        layer()
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.time_range is not None:
            evs.append((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_index", 0)))
    # stream ids are only in the chrome trace: export and parse it
    os.makedirs("gpurun_out", exist_ok=True)
    trace = f"gpurun_out/{tag}_trace_rank{rank}.json"
    prof.export_chrome_trace(trace)
    tr = json.load(open(trace))
    ks = [e for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ks.sort(key=lambda e: e["ts"])
    if not ks:
        print("no kernels captured")
        return
    t0 = ks[0]["ts"]
    t1 = max(e["ts"] + e["dur"] for e in ks)
    span = (t1 - t0) / 1e3

    def short(n):
        n = n.split("(")[0]
        n = n.replace("void ", "").replace("qt::", "")
        return n[:70]

    by_k, by_s = {}, {}
    for e in ks:
        n = short(e["name"])
        a = by_k.setdefault(n, {"n": 0, "total_ms": 0.0})
        a["n"] += 1
        a["total_ms"] += e["dur"] / 1e3
        s = str(e.get("args", {}).get("stream", "?"))
        by_s.setdefault(s, []).append((e["ts"], e["ts"] + e["dur"], n))
    # union + concurrency
    pts = []
    for e in ks:
        pts.append((e["ts"], 1))
        pts.append((e["ts"] + e["dur"], -1))
    pts.sort()
    conc = {0: 0.0, 1: 0.0, 2: 0.0, 3: 0.0}
    cur, last = 0, pts[0][0]
    for t, dlt in pts:
        conc[min(cur, 3)] += (t - last) / 1e3
        cur += dlt
        last = t
    # phase boundary: first chain kernel
    chain_names = ("gather_flip", "potrf_inv", "prepare")
    tb = next((e["ts"] for e in ks if any(c in e["name"] for c in chain_names)), t1)
    streams = {}
    for s, iv in by_s.items():
        busy = sum(b - a for a, b, _ in iv) / 1e3
        first, lastt = (iv[0][0] - t0) / 1e3, (max(b for _, b, _ in iv) - t0) / 1e3
        top = {}
        for a, b, n in iv:
            top[n] = top.get(n, 0.0) + (b - a) / 1e3
        streams[s] = {"busy_ms": round(busy, 3), "first_ms": round(first, 3), "last_ms": round(lastt, 3), "kernels": len(iv),
                      "top": {k: round(v, 3) for k, v in sorted(top.items(), key=lambda kv: -kv[1])[:8]}}
    # the stream that finishes last is the critical one
    crit = max(streams.items(), key=lambda kv: kv[1]["last_ms"])[0]
    out = {"tag": tag, "rank": rank, "world": world, "span_ms": round(span, 3),
           "busy_ms": round(span - conc[0], 3), "idle_ms": round(conc[0], 3),
           "concurrency_ms": {"1": round(conc[1], 3), "2": round(conc[2], 3), "3+": round(conc[3], 3)},
           "hessian_phase_ms": round((tb - t0) / 1e3, 3), "chain_loop_phase_ms": round((t1 - tb) / 1e3, 3),
           "critical_stream": crit, "streams": streams,
           "by_kernel": {k: {"n": v["n"], "total_ms": round(v["total_ms"], 3), "share_of_span": round(v["total_ms"] / span, 3)}
                         for k, v in sorted(by_k.items(), key=lambda kv: -kv[1]["total_ms"])[:40]},
           "note": "taken under the CUPTI profiler: shares, not speeds"}
    json.dump(out, open(f"gpurun_out/{tag}_timeline_rank{rank}.json", "w"), indent=1)
    if os.path.getsize(trace) > 24 << 20 or rank > 1:
        os.remove(trace)
    if rank == 0:
        print(json.dumps({k: out[k] for k in ("span_ms", "busy_ms", "idle_ms", "concurrency_ms", "hessian_phase_ms",
                                              "chain_loop_phase_ms", "critical_stream")}))
        for k, v in list(out["by_kernel"].items())[:25]:
            print(f"  {k:70s} n={v['n']:5d} {v['total_ms']:9.3f} ms  {v['share_of_span']:.3f}")
        for s, v in streams.items():
            print("  stream", s, v["busy_ms"], v["first_ms"], v["last_ms"], v["kernels"])
    if d.on:
        d.dist.barrier()
        d.dist.destroy_process_group()


if __name__ == "__main__":
    main()
