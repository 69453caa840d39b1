import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantool_b200.engine import awq as eawq, llama, schemes
shape = llama.SHAPES["llama-3-8b"]; dev = "cuda"
g = torch.Generator(device=dev).manual_seed(5)
h = torch.randn((128, 512, shape.hidden_size), device=dev, generator=g).to(torch.bfloat16)
cos, sin = llama.rope_tables(shape, 512, dev, h.dtype)
args = schemes.resolve("W4A16")
for chunk in (8, 32, 128):
    for rep in range(2):
        w = llama.random_layer_weights(shape, 0, dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eawq.awq_layer(shape, w, h, cos, sin, args, chunk); e1.record(); torch.cuda.synchronize()
    print(f"chunk_samples={chunk}: {e0.elapsed_time(e1):.1f} ms / layer")
