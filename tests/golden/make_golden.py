"""Generates tests/golden/*.npz from the INSTALLED third-party halves of the reference's path:
gguf-py 0.19 (`gguf.quants.quantize` for Q8_0/Q4_0/Q4_1/Q5_0/Q5_1, `dequantize` for all types) and
compressed-tensors 0.15.0.1 (`calculate_qparams`, `quantize`, `pack_to_int32`).  The reference repo
itself holds no numeric fixtures, and llm-compressor / llama.cpp cannot be run here,
so these are the only externally pinned vectors.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    from gguf import GGMLQuantizationType as T
    from gguf import quants as gq
    rng = np.random.default_rng(20251018)
    x = rng.standard_normal((6, 512)).astype(np.float32)
    x[0, :32] = 0.0
    x[1, 3] = 25.0
    x[2, 0] = -x[2, 1]
    out = {"x": x}
    for name in ("Q8_0", "Q4_0", "Q4_1", "Q5_0", "Q5_1"):
        out[f"packed_{name}"] = gq.quantize(x, getattr(T, name))
    np.savez_compressed(os.path.join(HERE, "gguf_simple_types.npz"), **out)

    from oracle import ggml_quants as oq
    kq = {"x": x}
    for name in ("IQ4_NL", "Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K"):
        packed = oq.quantize(x, name)                  # oracle bytes (parity unpinned vs llama.cpp)
        kq[f"packed_{name}"] = packed
        kq[f"dequant_{name}"] = gq.dequantize(packed, getattr(T, name)).astype(np.float32)   # gguf-py reading them
    np.savez_compressed(os.path.join(HERE, "gguf_k_quants.npz"), **kq)

    from compressed_tensors.compressors.pack_quantized.helpers import pack_to_int32
    from compressed_tensors.quantization import preset_name_to_scheme
    from compressed_tensors.quantization.lifecycle.forward import quantize
    from compressed_tensors.quantization.utils import calculate_qparams
    g = torch.Generator().manual_seed(7)
    W = (torch.randn((24, 256), generator=g) * 0.05).to(torch.bfloat16)
    ct = {"W_bf16_bits": W.view(torch.int16).numpy()}
    for level in ("W4A16", "W4A16_ASYM", "W8A16"):
        a = preset_name_to_scheme(level, ["Linear"]).weights
        gs = a.group_size or 256
        Wg = W.reshape(24, 256 // gs, gs)
        s, z = calculate_qparams(torch.amin(Wg, dim=2), torch.amax(Wg, dim=2), a)
        codes = quantize(x=W, scale=s, zero_point=z, args=a, dtype=torch.int8)
        ct[f"{level}_scale_bits"] = s.view(torch.int16).numpy()
        ct[f"{level}_zp"] = z.to(torch.int8).numpy()
        ct[f"{level}_codes"] = codes.numpy()
        if a.num_bits == 4:
            ct[f"{level}_packed"] = pack_to_int32(codes, 4).numpy()
    np.savez_compressed(os.path.join(HERE, "ct_quantize.npz"), **ct)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, hashlib.sha256(open(os.path.join(HERE, f), "rb").read()).hexdigest()[:16])


if __name__ == "__main__":
    main()
