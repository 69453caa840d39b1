"""Where does the host-to-host GPTQ time go?  8 layers of the 8B shape, per-layer CPU issue time vs wall time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200.engine import llama, pipeline, schemes
shape0 = llama.SHAPES["llama-3-8b"]
L = 8
shape = llama.LlamaShape(**{**shape0.__dict__, "num_hidden_layers": L})
dev = torch.device("cuda:0")
host_sd = {}
for l in range(L):
    for k, v in llama.random_layer_weights(shape, l, dev).items():
        host_sd[f"model.layers.{l}.{k}"] = v.cpu().pin_memory()
host_sd["model.embed_tokens.weight"] = (torch.randn((shape.vocab_size, shape.hidden_size), device=dev) * 0.02).to(torch.bfloat16).cpu().pin_memory()
token_ids = torch.randint(0, shape.vocab_size, (128, 2048)).pin_memory()
args = schemes.resolve("W4A16", "group")
for rep in range(3):
    marks = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = pipeline.quantize_model_gptq(shape, host_sd, token_ids, args, dev, progress=lambda l: marks.append(time.perf_counter()))
    t1 = time.perf_counter()
    per = [marks[0] - t0] + [b - a for a, b in zip(marks, marks[1:])]
    print(f"rep {rep}: total {t1 - t0:.3f} s ({(t1 - t0) / L * 32:.2f} s per 32 layers); per-layer host time ms:", [round(p * 1e3) for p in per], "tail", round((t1 - marks[-1]) * 1e3))
import subprocess
print(subprocess.run(["bash", "-c", "nproc; uptime; grep MHz /proc/cpuinfo | sort | uniq -c | sort -rn | head -3"], capture_output=True, text=True).stdout)
