#!/usr/bin/env python
"""bench.py - headline benchmark of the quantization hot path (BASELINE.json metric:
"GPTQ/AWQ s per 8B model at 1/2/4/8 B200; GGUF pack GB/s vs HBM peak").

One "step" = one complete GPTQ W4A16 g128 act_order pass over a random-init Llama-3-8B-shaped
model (32 decoder layers x 7 Linears, 4 distinct Hessians per layer) with 128 x 2048-token
synthetic calibration activations:
    value  : seconds per model, activations and weights already resident in HBM
             (Hessian SYRK -> NCCL all-reduce -> inverse-Hessian factor -> column loop -> codes/pack)
    e2e    : seconds per model through the plugin-level entry (`quantize_model_gptq`): HOST
             (pinned) weights and token ids in, HOST packed tensors out; includes every H2D/D2H
             copy and the decoder-layer forwards that produce the activations.
The same JSON line carries the other halves of the metric at the same N:
    awq          AWQ W4A16 n_grid=20, 128x512 (BASELINE config 3): s/model device-resident + e2e
    smoothquant  alpha=0.5 scales + fold (config 4)
    gguf_pack    GB/s per block type vs measured HBM peak;  gguf_smollm2: config 1 device-resident and file->file
    dequant      GB/s per type, vLLM's ggml_dequantize timed beside it
    tgemm        the error-feedback GEMM at the column loop's shapes
    parity       layer-0 code agreement of this (sharded) run vs an unsharded run + artifact sha
Multi-GPU (torchrun): calibration samples and output rows are sharded, H is all-reduced; total
work is fixed => "strong" scaling.  Timing: CUDA events, barrier + synchronize on both sides, max
over ranks.  `--impl reference` times the CPU oracle port on a bounded sample (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # see quantool_b200/__init__.py (must precede CUDA init)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--layers", type=int, default=0, help="debug: limit decoder layers (0 = all)")
    ap.add_argument("--samples", type=int, default=128)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--level", default="W4A16")
    ap.add_argument("--actorder", default="group")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gguf", action="store_true")
    ap.add_argument("--no-vllm", action="store_true", help="skip the vLLM ggml_dequantize comparison child process")
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--no-side", action="store_true", help="skip the awq / smoothquant / tgemm blocks of the default line")
    ap.add_argument("--workload", default="gptq", choices=["gptq", "awq", "smoothquant", "gguf", "probes"],
                    help="gptq = the full headline line (with awq / smoothquant / gguf / dequant / tgemm blocks); "
                         "the others run one block alone")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synth_acts(T, K, device, seed):
    """N(0,1) bf16 with 0.5 % outlier channels x20 (SURVEY §8d synthetic inputs)."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn((T, K), generator=g, device=device, dtype=torch.bfloat16)
    idx = torch.randperm(K, generator=g, device=device)[: max(1, K // 200)]
    x[:, idx] *= 20.0
    return x


# ------------------------------------------------------------------------------------------
def cpu_reference_sample(shape, level, actorder, n_samples, seq, threads=None, hess_samples=2):
    """Times the CPU oracle (torch CPU fp32, all host threads) on ONE FULL DECODER LAYER of the workload
    (SURVEY.md 8d): the Hessians of its 4 distinct inputs - accumulated over `hess_samples` of the n_samples
    calibration samples, then scaled by n_samples / hess_samples (the accumulation is exactly linear in the sample
    count) - and `quantize_weight` of all 7 Linears at their full shapes.  The model figure is that layer x L."""
    from oracle import gptq as og
    from compressed_tensors.quantization import ActivationOrdering
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    args = og.scheme_weight_args(level)
    if actorder and actorder != "none":
        args.actorder = ActivationOrdering.GROUP if actorder == "group" else ActivationOrdering.WEIGHT
    g = torch.Generator().manual_seed(0)
    L = shape.num_hidden_layers
    t_hess = t_quant = 0.0
    hs = {}
    for name, K in shape.input_dims().items():
        x = torch.randn((hess_samples, seq, K), generator=g).to(torch.bfloat16)
        H, n = og.make_empty_hessian(K), 0
        t0 = time.perf_counter()
        for b in range(hess_samples):
            H, n = og.accumulate_hessian(x[b:b + 1], H, n)
        t_hess += time.perf_counter() - t0
        hs[name] = H
        del x
    from quantool_b200.engine import llama
    for lin, (N, K) in shape.linear_shapes().items():
        W = (torch.randn((N, K), generator=g) * 0.02).to(torch.bfloat16)
        t0 = time.perf_counter()
        og.quantize_weight(W, hs[llama.INPUT_OF[lin]], args)
        t_quant += time.perf_counter() - t0
    layer_s = t_hess * (n_samples / hess_samples) + t_quant
    sample = (f"oracle port (torch CPU fp32, restated llm-compressor GPTQ) on one full decoder layer of {L}: "
              f"4 Hessians over {hess_samples} of {n_samples} samples x {seq} tokens ({t_hess:.2f} s, scaled x"
              f"{n_samples / hess_samples:.0f}: exactly linear in samples) + quantize_weight of all 7 Linears "
              f"({t_quant:.2f} s); model = layer x {L} (extrapolation, stated)")
    return {"value": layer_s * L, "unit": "s", "cores": threads, "kind": "port", "sample": sample,
            "sample_seconds": t_hess + t_quant, "layer_seconds_extrapolated": layer_s}


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def ev_time(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def med_time(fn, warm=2, n=5):
    for _ in range(warm):
        fn()
    ts = sorted(ev_time(fn) for _ in range(n))
    return ts[len(ts) // 2]


def l2_flush(dev):
    """Write a buffer larger than the 126 MB L2 between timed iterations of the small-input probes."""
    if not hasattr(l2_flush, "buf"):
        l2_flush.buf = torch.empty((256 << 20,), dtype=torch.uint8, device=dev)
    l2_flush.buf.fill_(1)


def gguf_probe(device, hbm):
    """GGUF pack GB/s vs HBM peak (second half of BASELINE's metric): one [16384, 14336] fp16 tensor (470 MB, larger
    than L2) per type, algorithmic bytes = fp16 in + packed out."""
    from quantool_b200 import cabi
    out = {}
    n, k = 4096 * 4, 14336
    x = (torch.randn((n, k), device=device) * 0.02).half()
    for t in ("Q8_0", "Q4_0", "Q5_0", "IQ4_NL", "Q4_K", "Q6_K"):
        be, bb = cabi.gguf_block_elems(t), cabi.gguf_block_bytes(t)
        y = torch.empty((n, k // be * bb), dtype=torch.uint8, device=device)
        ms = med_time(lambda: cabi.gguf_quantize(x, t, out=y), warm=3, n=5)
        byts = x.numel() * 2 + y.numel()
        out[t] = {"ms": round(ms, 4), "GBps": round(byts / ms / 1e6, 1), "bytes_per_elem": round(byts / x.numel(), 4),
                  "Gelem_per_s": round(x.numel() / ms / 1e6, 2), "frac_of_hbm_peak": round(byts / ms / 1e6 / hbm, 3)}
    # the K-quant packers are bound by fp32 instruction issue (strictly ordered candidate search), not by HBM:
    # their roofline is the issue rate, ops/element counted from SASS (profiles/r02_kquant_ops.json)
    try:
        ops = json.load(open(os.path.join(ROOT, "profiles", "r02_kquant_ops.json")))
        for t in ("Q4_K", "Q6_K", "IQ4_NL"):
            if t in ops:
                o = ops[t]
                ach = out[t]["Gelem_per_s"] * 1e9 * o["issue_slots_per_element"]
                out[t]["alu_roofline"] = {"bound": "fp32 issue", "issue_slots_per_element": o["issue_slots_per_element"],
                                          "achieved_Gslots_per_s": round(ach / 1e9, 1), "peak_Gslots_per_s": o["peak_Gslots_per_s"],
                                          "frac": round(ach / 1e9 / o["peak_Gslots_per_s"], 3), "source": o["source"]}
    except Exception:
        pass
    return out


def dequant_probe(device, hbm, vllm=True):
    """GGUF dequantize GB/s (row a16), with vLLM's CUDA `ggml_dequantize` (the bar to beat) timed beside
    it in a child process (so that vLLM's extension is not loaded into this one)."""
    from quantool_b200 import cabi
    n, k = 8192, 14336
    x = (torch.randn((n, k), device=device) * 0.02).half()
    out = {}
    for t in ("Q8_0", "Q4_0", "Q5_0", "Q4_K", "Q6_K"):
        y = cabi.gguf_quantize(x, t)
        ms = med_time(lambda: cabi.gguf_dequantize(y, t, k), warm=2, n=5)
        byts = y.numel() + n * k * 4
        out[t] = {"ms": round(ms, 4), "GBps": round(byts / ms / 1e6, 1), "frac_of_hbm_peak": round(byts / ms / 1e6 / hbm, 3),
                  "bytes": "packed in + fp32 out"}
    code = r"""
import json, torch
from vllm import _custom_ops as ops
n, k = 8192, 14336
T = {"Q8_0": (8, 32, 34), "Q4_0": (2, 32, 18), "Q5_0": (6, 32, 22), "Q4_K": (12, 256, 144), "Q6_K": (14, 256, 210)}
res = {}
for name, (tid, be, bb) in T.items():
    W = torch.randint(0, 255, (n, k // be * bb), dtype=torch.uint8, device="cuda")
    f = lambda: ops.ggml_dequantize(W, tid, n, k, torch.float16)
    for _ in range(3): f()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[2]
    res[name] = {"ms": round(ms, 4), "GBps": round((W.numel() + n * k * 2) / ms / 1e6, 1), "bytes": "packed in + fp16 out"}
print("VLLM_JSON " + json.dumps(res))
"""
    if not vllm:
        return out
    try:
        env = dict(os.environ)
        idx = device.index or 0
        vis = env.get("CUDA_VISIBLE_DEVICES")
        env["CUDA_VISIBLE_DEVICES"] = vis.split(",")[idx] if vis else str(idx)
        for k_ in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k_, None)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240, env=env)
        got = [l for l in r.stdout.splitlines() if l.startswith("VLLM_JSON ")]
        if got:
            v = json.loads(got[-1][len("VLLM_JSON "):])
            for t, d in v.items():
                out[t]["vllm_ggml_dequantize"] = d
                out[t]["speedup_vs_vllm_time"] = round(d["ms"] / out[t]["ms"], 2)
        else:
            out["vllm_error"] = (r.stderr or r.stdout)[-300:]
    except Exception as e:      # vLLM missing / too slow to import: our numbers stand alone
        out["vllm_error"] = repr(e)[:300]
    return out


def tgemm_probe(device, peaks):
    """The error-feedback GEMM (`tgemm_kernel`, 3xTF32 tcgen05) at the shapes the column loop runs it at:
    W[:, i2:] -= Err[M, Kd] * U[i1:i2, i2:]  with M = rows of the Linear, Kd = 128 (inside a 512-column outer block)
    or 512 (everything right of it).  TFLOP/s counts the three tf32 products (what the tensor pipe executes);
    `fp32_equivalent` is 2*M*N*Kd."""
    from quantool_b200 import cabi
    tf32_peak = (peaks.get("bf16_tflops") or 1675.0) / 2.0        # dense tf32 = half the bf16 rate
    out = {}
    for M, K in ((4096, 14336), (14336, 4096), (4096, 4096)):
        for Kd in (128, 512):
            N = K - Kd
            A = torch.randn((M, Kd), device=device)
            Bt = torch.randn((K, Kd), device=device)
            C = torch.zeros((M, K), device=device)
            sa, sb = cabi.split_tf32(A), cabi.split_tf32(Bt)
            fn = lambda: cabi.gemm_tf32x3(sa, (sb[0][Kd:], sb[1][Kd:]), C[:, Kd:], negate=True, accumulate=True)
            ms = med_time(fn, warm=2, n=7)
            fl = 2.0 * M * N * Kd
            out[f"M{M}_N{N}_Kd{Kd}"] = {"ms": round(ms, 4), "fp32_equivalent_tflops": round(fl / ms / 1e9, 1),
                                        "tensor_tflops_tf32": round(3 * fl / ms / 1e9, 1),
                                        "frac_of_tf32_peak": round(3 * fl / ms / 1e9 / tf32_peak, 3),
                                        "c_traffic_GBps": round(M * N * 4 * 2 / ms / 1e6, 1)}
            del A, Bt, C, sa, sb
    out["peak"] = {"tf32_tflops": tf32_peak, "source": "half of MEASURED_PEAKS bf16_tflops (burst; kernel timed alone)"}
    return out


def gguf_smollm2(dev, world, rank):
    """BASELINE config 1: GGUF Q8_0 and Q4_K_M of a random-init SmolLM2-135M-shaped Llama.
    (a) device-resident: every 2-D tensor packed on the GPU vs the C oracle with all host threads, bit-exact check;
    (b) file -> file through the plugin: `GGUF.quantize(model=<HF dir>, level=[Q8_0, Q4_K_M])`, i.e. safetensors on
        disk in, two .gguf files on disk out (convert to the f16 base + quantize, all H2D/D2H inside)."""
    import shutil
    import tempfile
    from quantool_b200 import cabi
    from quantool_b200.engine import gguf_file, llama
    from oracle import ggml_quants as oq
    shape = llama.SHAPES["smollm2-135m"]
    sd = llama.random_state_dict(shape, seed=0, device=dev)
    tensors = []
    for name, t in sd.items():
        g = gguf_file.hf_to_gguf_name(name, shape.num_hidden_layers)
        if g is not None and t.dim() == 2:
            tensors.append((g, t.half().contiguous()))
    out = {}
    for ftype in ("Q8_0", "Q4_K_M"):
        plan = [(x, gguf_file.tensor_type(g, tuple(x.shape), ftype, shape.num_hidden_layers, False)) for g, x in tensors]
        ys = cabi.gguf_quantize_many(plan)
        ms = med_time(lambda: cabi.gguf_quantize_many(plan), warm=2, n=5)
        nelem = sum(x.numel() for x, _ in plan)
        byts = sum(x.numel() * 2 for x, _ in plan) + sum(y.numel() for y in ys)
        host = [(x.float().cpu().numpy(), qt) for x, qt in plan]
        t0 = time.perf_counter()
        refs = [oq.quantize(xh, qt) for xh, qt in host]
        cpu_s = time.perf_counter() - t0
        exact = all((y.cpu().numpy() == r).all() for y, r in zip(ys, refs))
        types = {}
        for _, qt in plan:
            types[qt] = types.get(qt, 0) + 1
        out[ftype] = {"gpu_ms_device_resident": round(ms, 3), "elements": nelem, "GBps": round(byts / ms / 1e6, 1),
                      "cpu_oracle_s": round(cpu_s, 3), "cpu_threads": oq.get_threads(), "bit_exact_vs_oracle": bool(exact),
                      "tensor_types": types}
    # (b) file -> file
    tmp = tempfile.mkdtemp(prefix="qt_gguf_")
    try:
        from safetensors.torch import save_file
        import quantool_b200.methods  # noqa: F401
        from quantool_b200 import QuantizerRegistry
        mdir = os.path.join(tmp, "smollm2-135m-rand")
        os.makedirs(mdir)
        save_file({k: v.cpu() for k, v in sd.items()}, os.path.join(mdir, "model.safetensors"), metadata={"format": "pt"})
        json.dump(shape.to_hf_config(), open(os.path.join(mdir, "config.json"), "w"))
        q = QuantizerRegistry.create("gguf", model_id="bench/smollm2-135m-rand")
        q.quantize(model=mdir, level=["Q8_0"], output_dir=os.path.join(tmp, "warm"), require_tokenizer=False)
        t0 = time.perf_counter()
        files = q.quantize(model=mdir, level=["Q8_0", "Q4_K_M"], output_dir=os.path.join(tmp, "out"), require_tokenizer=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        in_b = os.path.getsize(os.path.join(mdir, "model.safetensors"))
        out["file_to_file"] = {"seconds": round(dt, 3), "call": "GGUF.quantize(model=<dir>, level=['Q8_0','Q4_K_M'])",
                               "input_safetensors_bytes": in_b, "output_bytes": [os.path.getsize(f) for f in files],
                               "includes": "safetensors read, f16 base GGUF write+read, H2D, CUDA packers, D2H, two GGUF writes",
                               "gpus_used": torch.cuda.device_count()}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def smoothquant_block(a, shape, dev, d, hbm):
    """BASELINE config 4: SmoothQuant alpha=0.5 scales + fold on the random-init 8B shape (128x512 calibration),
    samples sharded over the ranks (min/max all-reduce)."""
    from quantool_b200 import cabi
    from quantool_b200.engine import llama, pipeline, smoothquant as esq
    n, seq = a.samples, 512
    per = pipeline.row_split(n, d.world)[d.rank]
    g = torch.Generator(device=dev).manual_seed(5 + d.rank)
    h = torch.randn((max(per, 1), seq, shape.hidden_size), device=dev, generator=g).to(torch.bfloat16)[:per]
    cos, sin = llama.rope_tables(shape, seq, dev, h.dtype)
    L = shape.num_hidden_layers
    ws = [llama.random_layer_weights(shape, l, dev) for l in range(2)]

    def step():
        for l in range(L):
            w = {k: v.clone() for k, v in ws[l % 2].items()}      # the fold is in place
            esq.smooth_layer(shape, w, h, cos, sin, 0.5, 8, d)
    step()
    if d.on:
        d.dist.barrier()
    ms = torch.tensor([ev_time(step)], device=dev)
    d.all_reduce_max(ms)
    x = torch.cat([h.reshape(-1, shape.hidden_size)] * max(1, (2 << 30) // max(1, h.numel() * 2)))
    mn, mx = cabi.new_minmax(shape.hidden_size, dev)
    kms = med_time(lambda: cabi.channel_minmax(x, mn, mx), warm=1, n=5)
    gbs = x.numel() * 2 / kms / 1e6
    wt = ws[0]["mlp.gate_proj.weight"].clone()
    s = torch.rand((wt.shape[1],), device=dev) + 0.5
    fms = med_time(lambda: cabi.scale_matrix_(wt, s), warm=1, n=5)
    return {"value": ms.item() / 1e3, "unit": "s per 8B model", "n_gpus": d.world, "calibration": f"{n}x{seq}",
            "includes": "calibration forward up to post_attention_layernorm, channel min/max, scales, fold",
            "roofline": {"kernel": "col_reduce_kernel<MINMAX>", "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                         "frac": gbs / hbm, "traffic": None},
            "fold_kernel": {"kernel": "scale_kernel", "GBps": wt.numel() * 4 / fms / 1e6, "frac": wt.numel() * 4 / fms / 1e6 / hbm}}


def awq_block(a, shape, dev, d, hbm, peaks):
    """BASELINE config 3: AWQ W4A16 g128 n_grid=20 on the random-init 8B shape, 128x512 calibration; samples sharded
    over the ranks (x_mean and the 20 losses all-reduced).  value = device-resident, e2e = quantize_model_awq with
    pinned host weights in / packed host tensors out."""
    from quantool_b200 import cabi
    from quantool_b200.engine import awq as eawq, llama, pipeline, schemes
    args = schemes.resolve("W4A16")
    n, seq = a.samples, 512
    per = pipeline.row_split(n, d.world)[d.rank]
    g = torch.Generator(device=dev).manual_seed(5 + d.rank)
    h = torch.randn((per, seq, shape.hidden_size), device=dev, generator=g).to(torch.bfloat16)
    cos, sin = llama.rope_tables(shape, seq, dev, h.dtype)
    L = shape.num_hidden_layers
    ws = [llama.random_layer_weights(shape, l, dev) for l in range(2)]

    def run(layers):
        for l in range(layers):
            w = {k: v.clone() for k, v in ws[l % 2].items()}
            eawq.awq_layer(shape, w, h, cos, sin, args, 32, d)
    run(1)
    if d.on:
        d.dist.barrier()
    l0 = cabi.launch_count()
    ms = torch.tensor([ev_time(lambda: run(L))], device=dev)
    launches = cabi.launch_count() - l0
    d.all_reduce_max(ms)
    out = {"value": ms.item() / 1e3, "unit": "s per 8B model", "n_gpus": d.world, "steps": 1, "warmup": "1 layer",
           "calibration": f"{n}x{seq}", "n_grid": 20, "gpu_launches": launches,
           "flops_per_model": eawq.model_flops(shape, n * seq, 20) if hasattr(eawq, "model_flops") else None}
    if out["flops_per_model"]:
        tf = out["flops_per_model"] / d.world / ms.item() / 1e9
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        out["roofline"] = {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                           "what": "all GEMM FLOPs of the 20-point search per rank / search time (bf16 dense)"}
    wt = ws[0]["mlp.gate_proj.weight"]
    s = torch.rand((wt.shape[1],), device=dev) + 0.5
    o = torch.empty_like(wt)
    kms = med_time(lambda: cabi.awq_scale_qdq(wt, s, 128, 4, True, out=o), warm=1, n=5)
    out["scale_qdq_kernel"] = {"GBps": wt.numel() * 4 / kms / 1e6, "frac_of_hbm_peak": wt.numel() * 4 / kms / 1e6 / hbm,
                               "algorithmic": "4 B/elem (bf16 read + bf16 write)"}
    del h, ws
    torch.cuda.empty_cache()
    if not a.no_e2e:
        host_sd = {}
        for l in range(L):
            for k, v in llama.random_layer_weights(shape, l % 2, "cpu").items():
                host_sd[f"model.layers.{l}.{k}"] = v.pin_memory() if l < 2 else host_sd[f"model.layers.{l % 2}.{k}"]
        host_sd["model.embed_tokens.weight"] = (torch.randn((shape.vocab_size, shape.hidden_size)) * 0.02).to(torch.bfloat16).pin_memory()
        ids = torch.randint(0, shape.vocab_size, (n, seq), generator=torch.Generator().manual_seed(1234)).pin_memory()
        if d.on:
            d.dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = pipeline.quantize_model_awq(shape, host_sd, ids, args, dev, dist=d)
        torch.cuda.synchronize()
        if d.on:
            d.dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        d.all_reduce_max(dt)
        out["e2e"] = {"value": dt.item(), "unit": "s", "h2d_bytes_per_step": r.h2d_bytes, "d2h_bytes_per_step": r.d2h_bytes,
                      "path": "pipeline.quantize_model_awq (what AWQ.quantize() runs)"}
    return out


def parity_block(shape, args, host_sd, token_ids, dev, d, sharded_result, n_samples, seq):
    """Every multi-GPU bench run doubles as a parity check.
    A. Layer level at the workload's own shapes, IDENTICAL all-reduced H on both sides: `quantize_layer` with rows
       sharded over the ranks (owner-computes chain on its NCCL lane, packed broadcast of U, gathered rows) must give
       bit-identical `weight_packed` / `weight_scale` / `weight_g_idx` to the unsharded run of the same kernels.
    B. Rank 0 re-runs decoder layer 0 of the e2e workload UNSHARDED and compares with what the sharded run produced:
       H differs by fp32 summation order only.  Reported: summed GPTQ loss of layer 0 on both sides (the objective)
       and the code agreement - which on random-init weights is low BY CONSTRUCTION with actorder=group: diag(H) is
       nearly flat, so argsort(diag H) is decided by the last bits and every swap moves group boundaries."""
    import hashlib
    from quantool_b200.engine import llama, pipeline
    from quantool_b200.engine.gptq import compress_linear
    out = {}
    # ---- A
    per = pipeline.row_split(min(n_samples, 16), d.world)
    acts = {n: synth_acts(max(per[d.rank], 1) * seq, k, dev, 900 + i + 100 * d.rank)[: per[d.rank] * seq]
            for i, (n, k) in enumerate(shape.input_dims().items())}
    hess = pipeline.accumulate_layer_hessians(acts, per[d.rank], sum(per), d)
    del acts
    w = llama.random_layer_weights(shape, 0, dev)
    lq_sh = pipeline.GPTQLayerQuantizer(args, dist=d)
    res_sh = lq_sh.quantize_layer(w, hess)
    lq_sh.drop_scratch()
    lq_1 = pipeline.GPTQLayerQuantizer(args, dist=pipeline.Dist(enabled=False))
    res_1 = lq_1.quantize_layer(w, hess)
    lq_1.drop_scratch()
    same = True
    sha = hashlib.sha256()
    for lin in llama.LINEARS:
        a_sh, _ = compress_linear(res_sh[lin].weight, res_sh[lin].scale, res_sh[lin].zero_point, res_sh[lin].g_idx, args)
        a_1, _ = compress_linear(res_1[lin].weight, res_1[lin].scale, res_1[lin].zero_point, res_1[lin].g_idx, args)
        for k in a_1:
            same = same and torch.equal(a_sh[k], a_1[k])
        sha.update(a_sh["weight_packed"].cpu().numpy().tobytes())
    flag = torch.tensor([int(same)], device=dev)
    d.all_reduce_min(flag)
    out["layer_identical_given_same_H"] = bool(flag.item())
    out["layer_weight_packed_sha"] = sha.hexdigest()[:16]
    del hess, res_sh, res_1, w
    torch.cuda.empty_cache()
    # ---- B
    if d.rank == 0:
        sha = hashlib.sha256()
        for k in sorted(sharded_result.tensors):
            if k.endswith("weight_packed"):
                sha.update(sharded_result.tensors[k].numpy().tobytes())
        out["artifact_sha"] = sha.hexdigest()[:16]
        one = llama.LlamaShape(**{**shape.__dict__, "num_hidden_layers": 1})
        sd1 = {k: v for k, v in host_sd.items() if not k.startswith("model.layers.") or k.startswith("model.layers.0.")}
        r1 = pipeline.quantize_model_gptq(one, sd1, token_ids[:n_samples], args, dev, dist=pipeline.Dist(enabled=False))
        same = tot = 0
        for k, p1 in r1.tensors.items():
            if k.endswith("weight_packed"):
                p2 = sharded_result.tensors[k]
                for sft in range(0, 32, 4):
                    same += int((((p1 >> sft) & 15) == ((p2 >> sft) & 15)).sum())
                tot += p1.numel() * 8
        l_1 = sum(v for k, v in r1.losses.items())
        l_sh = sum(v for k, v in sharded_result.losses.items() if k.startswith("model.layers.0."))
        out["layer0_gptq_loss_sharded_vs_unsharded"] = [l_sh, l_1]
        out["layer0_loss_rel_diff"] = abs(l_sh - l_1) / max(abs(l_1), 1e-30)
        out["code_agreement_layer0_vs_unsharded"] = same / max(tot, 1)
        out["note"] = ("world=1: the same path run twice (determinism check)" if d.world == 1 else
                       f"{d.world} ranks (samples + rows sharded, NCCL all-reduce of H) vs one rank doing everything; "
                       "random-init weights give a nearly flat diag(H), so with actorder=group the permutation - and with "
                       "it the codes - follow the last bits of H: the loss and check A are the parity figures")
    if d.on:
        d.dist.barrier()
    return out


# ------------------------------------------------------------------------------------------
def main():
    a = parse()
    from quantool_b200.engine import llama
    shape = llama.SHAPES[a.model]
    if a.layers:
        shape = llama.LlamaShape(**{**shape.__dict__, "num_hidden_layers": a.layers})
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = (f"gptq {a.level} g128 actorder={a.actorder} on random-init {a.model} "
                f"({shape.num_hidden_layers} layers), {a.samples}x{a.seq}-token synthetic calibration")
    config = {"workload": workload, "level": a.level, "actorder": a.actorder, "calibration": f"{a.samples}x{a.seq}",
              "layers": shape.num_hidden_layers, "parallelism": f"samples+rows sharded x{world}, H all-reduce",
              "l2": "inputs larger than L2"}
    _t_rank = -(-a.samples // world) * a.seq
    config["l2"] = ("inputs larger than L2 (%.1f GB activations per rank + %.2f GB weights per layer per step; 126 MB L2)"
                    % (sum(_t_rank * k * 2 for k in shape.input_dims().values()) / 1e9,
                       sum(n_ * k_ * 2 for n_, k_ in shape.linear_shapes().values()) / 1e9))

    if a.impl == "reference":
        if rank != 0:
            return
        # one "step" = one model, measured on a bounded sample (one full decoder layer) and extrapolated; the number
        # of samples actually run is bounded so that the whole arm ends within a few minutes whatever K is
        budget_s = 200.0
        runs, t_start = [], time.perf_counter()
        for i in range(max(1, a.steps)):
            runs.append(cpu_reference_sample(shape, a.level, a.actorder, a.samples, a.seq))
            if time.perf_counter() - t_start + runs[-1]["sample_seconds"] > budget_s:
                break
        cb = dict(runs[-1])
        val = sum(r["value"] for r in runs) / len(runs)
        cb["value"] = val
        cb["samples_timed"] = len(runs)
        line = {"impl": "reference", "metric": "gptq_seconds_per_8b_model", "value": val, "unit": "s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": val * 1e3,
                "step_definition": "one model; value = (one full decoder layer timed on the host cores) x layers, "
                                   f"mean of {len(runs)} timed samples of {cb['sample_seconds']:.1f} s each",
                "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from quantool_b200 import cabi
    from quantool_b200.engine import pipeline, schemes
    from quantool_b200.engine.gptq import compress_linear
    d = pipeline.Dist()
    peaks = load_peaks()
    hbm = peaks.get("hbm_gbs") or 6650.0

    def finish(obj):
        if rank == 0:
            print(json.dumps(obj))
        if d.on:
            d.dist.destroy_process_group()

    if a.workload == "gguf":
        return finish({"metric": "gguf_pack_smollm2_135m", "n_gpus": world, "results": gguf_smollm2(dev, world, rank) if rank == 0 else None})
    if a.workload == "awq":
        return finish({"metric": "awq_seconds_per_8b_model", "higher_is_better": False, **awq_block(a, shape, dev, d, hbm, peaks)})
    if a.workload == "smoothquant":
        return finish({"metric": "smoothquant_seconds_per_8b_model", "higher_is_better": False,
                       **smoothquant_block(a, shape, dev, d, hbm)})
    if a.workload == "probes":
        return finish({"gguf_pack": gguf_probe(dev, hbm), "dequant": dequant_probe(dev, hbm), "tgemm": tgemm_probe(dev, peaks)}
                      if rank == 0 else None)

    args = schemes.resolve(a.level, a.actorder if a.actorder != "none" else None)
    per = pipeline.row_split(a.samples, d.world)
    n_local = per[d.rank]
    T_local = n_local * a.seq
    dims = shape.input_dims()
    acts = {n: synth_acts(T_local, k, dev, 7 + i + 100 * d.rank) for i, (n, k) in enumerate(dims.items())}
    L = shape.num_hidden_layers
    weights = [llama.random_layer_weights(shape, l, dev) for l in range(L)]
    lq = pipeline.GPTQLayerQuantizer(args, dist=d)
    Kmax = max(dims.values())
    hess_events = []

    from quantool_b200.engine.gptq import HessianAccumulator
    accs = {n: HessianAccumulator(k, dev) for n, k in dims.items()}

    def hot_step(record=False):
        for l in range(L):
            ev = {}
            if record:
                ev = {n: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                      for n, x in acts.items() if x.shape[-1] == Kmax}
                hess_events.extend(ev.values())
            # raw sums per rank; the NCCL all-reduce of each H, its 2/n scaling, the chain and the column loops run
            # per distinct input on that input's stream and communicator lane (what quantize_model_gptq does)
            done = pipeline.accumulate_layer_sums(acts, n_local, accs, syrk_events=ev, dist=d)
            res = lq.quantize_layer(weights[l], None, accs=accs, n_total=a.samples, acc_events=done)
            for lin, r in res.items():
                compress_linear(r.weight, r.scale, r.zero_point, r.g_idx, args)

    def barrier():
        if d.on:
            d.dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        hot_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(a.steps):
        hot_step(record=True)
    e1.record()
    barrier()
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    d.all_reduce_max(ms_total)
    launches = cabi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total.item() / a.steps
    hess_ms = sum(x.elapsed_time(y) for x, y in hess_events) / max(1, len(hess_events))
    del hess_events[:]

    # roofline of the dominant single kernel: the tcgen05 Hessian SYRK at K = intermediate size
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)" \
        if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    NI, NJ = (Kmax + 127) // 128, (Kmax + 255) // 256
    ntiles = sum(min(2 * j + 2, NI) for j in range(NJ))
    exec_flops = 2.0 * T_local * ntiles * 128 * 256
    ref_flops = 2.0 * T_local * Kmax * Kmax
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "hessian_syrk_traffic.json")))
        if tj.get("T") == T_local and tj.get("K") == Kmax:
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    roofline = {"kernel": "hessian_syrk_kernel", "bound": "tensor", "achieved": exec_flops / hess_ms / 1e9,
                "peak": peak, "unit": "TFLOP/s", "frac": exec_flops / hess_ms / 1e9 / peak, "traffic": traffic,
                "traffic_unit": "bytes/launch (dram read + write)", "traffic_source": traffic_src,
                "peak_source": peak_src, "ms_per_launch": hess_ms, "shape": f"X[{T_local},{Kmax}] bf16 -> H[{Kmax},{Kmax}] fp32",
                "algorithmic": "executed upper-triangle tile FLOPs 2*T*ntiles*128*256",
                "reference_equivalent_tflops": ref_flops / hess_ms / 1e9}

    # ---- e2e through the plugin-level entry with host buffers -------------------------------
    e2e = parity = None
    del acts, accs
    lq.drop_scratch()
    if not a.no_e2e:
        host_sd = {}
        for l in range(L):
            for k, v in weights[l].items():
                host_sd[f"model.layers.{l}.{k}"] = v.cpu().pin_memory()
        g = torch.Generator().manual_seed(1234)
        emb = (torch.randn((shape.vocab_size, shape.hidden_size), device=dev) * 0.02).to(torch.bfloat16)
        host_sd["model.embed_tokens.weight"] = emb.cpu().pin_memory()
        del emb
        token_ids = torch.randint(0, shape.vocab_size, (a.samples, a.seq), generator=g).pin_memory()
        del weights
        torch.cuda.empty_cache()
        fmt = "pack-quantized"

        def e2e_step():
            return pipeline.quantize_model_gptq(shape, host_sd, token_ids, args, dev, fmt=fmt, dist=d)

        e2e_step()   # warm-up (allocator, cuBLAS heuristics)
        barrier()
        t0 = time.perf_counter()
        r = None
        for _ in range(a.e2e_steps):
            r = e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / a.e2e_steps], device=dev)
        d.all_reduce_max(dt)
        e2e = {"value": dt.item(), "unit": "s", "h2d_bytes_per_step": r.h2d_bytes, "d2h_bytes_per_step": r.d2h_bytes,
               "steps": a.e2e_steps, "warmup": 1,
               "path": "quantool_b200.engine.pipeline.quantize_model_gptq (what GPTQ.quantize() runs): pinned host "
                       "weights + token ids -> layer forwards -> Hessians -> GPTQ -> packed host tensors"}
        try:
            parity = parity_block(shape, args, host_sd, token_ids, dev, d, r, a.samples, a.seq)
            if e2e is not None and parity:
                e2e["artifact_sha"] = parity.get("artifact_sha")
                e2e["code_agreement"] = parity.get("code_agreement_layer0_vs_unsharded")
                e2e["layer_identical_given_same_H"] = parity.get("layer_identical_given_same_H")
                e2e["layer0_loss_rel_diff"] = parity.get("layer0_loss_rel_diff")
        except Exception as ex:
            parity = {"error": repr(ex)[:300]}
        del host_sd, r
    else:
        del weights
    torch.cuda.empty_cache()

    # ---- the other halves of the metric, same run, same N --------------------------------------------
    def guarded(fn, *xs):
        try:
            return fn(*xs)
        except Exception as ex:          # a failing side block must not take the headline line down
            import traceback
            traceback.print_exc()
            return {"error": repr(ex)[:300]}

    awq = smooth = None
    if not a.no_side and a.model == "llama-3-8b" and not a.layers:
        awq = guarded(awq_block, a, shape, dev, d, hbm, peaks)
        torch.cuda.empty_cache()
        smooth = guarded(smoothquant_block, a, shape, dev, d, hbm)
        torch.cuda.empty_cache()
    if rank != 0:
        if d.on:
            d.dist.barrier()
            d.dist.destroy_process_group()
        return
    gg = None if a.no_gguf else guarded(gguf_probe, dev, hbm)
    dq = None if a.no_gguf else guarded(dequant_probe, dev, hbm, not a.no_vllm)
    gs = None if a.no_gguf else guarded(gguf_smollm2, dev, world, rank)
    tg = None if a.no_side else guarded(tgemm_probe, dev, peaks)
    cb = None if a.no_cpu else cpu_reference_sample(shape, a.level, a.actorder, a.samples, a.seq)
    line = {"metric": "gptq_seconds_per_8b_model", "value": ms_step / 1e3, "unit": "s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "dtype_detail": "bf16 activations on the tensor cores -> f32 accumulate; f32 (3xTF32) inverse-Hessian solve and column loop",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cb, "parity": parity, "awq": awq, "smoothquant": smooth,
            "gguf_pack": gg, "gguf_smollm2": gs, "dequant": dq, "tgemm": tg}
    print(json.dumps(line))
    if d.on:
        d.dist.barrier()
        d.dist.destroy_process_group()


def _run_with_clean_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on
    communicator creation), so file descriptor 1 is pointed at stderr for the whole run and the JSON line is
    written to the saved descriptor at the end."""
    import io
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    py_stdout, sys.stdout = sys.stdout, buf
    try:
        main()
    finally:
        sys.stdout = py_stdout
        os.dup2(real, 1)
        os.close(real)
        out = buf.getvalue()
        if out:
            os.write(1, out.encode())


if __name__ == "__main__":
    _run_with_clean_stdout()
