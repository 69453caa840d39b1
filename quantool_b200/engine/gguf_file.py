"""GGUF file handling around the block packers: HF -> f16 GGUF, per-tensor type selection,
f16 GGUF -> quantized GGUF.

Replaces the two child processes of the reference's GGUF plugin
(ref/src/quantool/methods/llama_cpp/llama_cpp.py:121-178): `convert_hf_to_gguf.py --outtype f16`
and `llama-quantize <in> <out> <LEVEL>`.  The container work (GGUF v3 layout, 32-byte alignment,
HF -> GGUF tensor names) uses the installed gguf-py writer/reader; the
per-tensor type table restates llama.cpp `llama_tensor_get_type`; all
block arithmetic runs in the CUDA packers (csrc/gguf.cu) with tensors distributed over the GPUs
of the node by size.
"""
import json
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import cabi

# default tensor type of each file type in the reference's QuantType list (llama_cpp.py:15-31)
BASE_TYPE = {"Q4_0": "Q4_0", "Q4_1": "Q4_1", "Q5_0": "Q5_0", "Q5_1": "Q5_1", "Q8_0": "Q8_0",
             "Q2_K": "Q2_K", "Q3_K_S": "Q3_K", "Q3_K_M": "Q3_K", "Q3_K_L": "Q3_K",
             "Q4_K_S": "Q4_K", "Q4_K_M": "Q4_K", "Q5_K_S": "Q5_K", "Q5_K_M": "Q5_K", "Q6_K": "Q6_K"}
# row length not a multiple of 256: llama.cpp falls back per tensor (Q2_K / Q3_K -> IQ4_NL)
FALLBACK = {"Q4_K": "Q5_0", "Q5_K": "Q5_1", "Q6_K": "Q8_0", "Q2_K": "IQ4_NL", "Q3_K": "IQ4_NL"}


def use_more_bits(i_layer: int, n_layers: int) -> bool:
    return i_layer < n_layers // 8 or i_layer >= 7 * n_layers // 8 or (i_layer - n_layers // 8) % 3 == 2


def tensor_type(name: str, shape: Tuple[int, ...], ftype: str, n_layers: int, has_output: bool,
                n_head: int = 0, n_head_kv: int = 0) -> str:
    """ggml type of one tensor under file type `ftype` (llama arch, no imatrix, default options).
    Returns "F32" for tensors that are not quantized."""
    if len(shape) < 2 or not name.endswith("weight") or "norm" in name:
        return "F32"
    new = BASE_TYPE[ftype]
    layer = int(name.split(".")[1]) if name.startswith("blk.") else -1
    ncols = shape[-1]
    is_output = name == "output.weight" or (name == "token_embd.weight" and not has_output)
    if is_output:
        if ncols % 256 != 0:
            new = "Q8_0"
        elif new != "Q8_0":
            new = "Q6_K"
    elif name.endswith("attn_v.weight"):
        n_gqa = n_head // n_head_kv if n_head and n_head_kv else 1
        if ftype == "Q2_K":
            new = "Q4_K" if n_gqa >= 4 else "Q3_K"
        elif ftype == "Q3_K_M":
            new = "Q5_K" if layer < 2 else "Q4_K"
        elif ftype == "Q3_K_L":
            new = "Q5_K"
        elif ftype in ("Q4_K_M", "Q5_K_M") and use_more_bits(layer, n_layers):
            new = "Q6_K"
        elif ftype == "Q4_K_S" and layer < 4:
            new = "Q5_K"
        if n_gqa >= 8 and n_layers == 80 and new in ("Q3_K", "Q4_K"):
            new = "Q5_K"   # 70B: 8 heads share attn_v
    elif name.endswith("ffn_down.weight"):
        if ftype == "Q2_K":
            new = "Q3_K"
        elif ftype == "Q3_K_M":
            new = "Q5_K" if layer < n_layers // 16 else "Q4_K"
        elif ftype == "Q3_K_L":
            new = "Q5_K"
        elif ftype in ("Q4_K_M", "Q5_K_M") and use_more_bits(layer, n_layers):
            new = "Q6_K"
        elif ftype == "Q4_K_S" and layer < n_layers // 8:
            new = "Q5_K"
    elif name.endswith("attn_output.weight"):
        new = {"Q2_K": "Q3_K", "Q3_K_M": "Q4_K", "Q3_K_L": "Q5_K"}.get(ftype, new)
    be = 256 if new.endswith("_K") else 32
    if ncols % be != 0:
        new = FALLBACK.get(new, new)
        if ncols % 32 != 0:
            new = "F16"
    return new


def hf_to_gguf_name(name: str, n_layers: int) -> Optional[str]:
    import gguf
    tm = gguf.get_tensor_name_map(gguf.MODEL_ARCH.LLAMA, n_layers)
    return tm.get_name(name, try_suffixes=(".weight", ".bias"))


def _permute_qk(w: torch.Tensor, n_head: int) -> torch.Tensor:
    """convert_hf_to_gguf LlamaModel.permute: undo HF's rotary half-split layout for llama.cpp."""
    out_dim = w.shape[0]
    return (w.reshape(n_head, 2, out_dim // n_head // 2, *w.shape[1:]).swapaxes(1, 2).reshape(w.shape))


def load_hf_model(model_path: str):
    """(hf config dict, {name: host tensor}) from a local HF directory with safetensors shards."""
    from safetensors.torch import load_file
    with open(os.path.join(model_path, "config.json")) as f:
        cfg = json.load(f)
    sd = {}
    files = sorted(f for f in os.listdir(model_path) if f.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no .safetensors files under {model_path}")
    for fn in files:
        sd.update(load_file(os.path.join(model_path, fn)))
    return cfg, sd


def convert_hf_to_f16_gguf(model_path: str, out_file: str, outtype: str = "f16") -> str:
    """HF checkpoint -> GGUF with 2-D tensors in fp16 (or fp32) and 1-D tensors in fp32."""
    import gguf
    cfg, sd = load_hf_model(model_path)
    n_layers = cfg["num_hidden_layers"]
    n_head = cfg["num_attention_heads"]
    n_kv = cfg.get("num_key_value_heads", n_head)
    w = gguf.GGUFWriter(out_file, "llama")
    w.add_name(os.path.basename(os.path.normpath(model_path)))
    w.add_block_count(n_layers)
    w.add_context_length(cfg.get("max_position_embeddings", 2048))
    w.add_embedding_length(cfg["hidden_size"])
    w.add_feed_forward_length(cfg["intermediate_size"])
    w.add_head_count(n_head)
    w.add_head_count_kv(n_kv)
    w.add_layer_norm_rms_eps(cfg.get("rms_norm_eps", 1e-5))
    rope = cfg.get("rope_theta") or (cfg.get("rope_parameters") or {}).get("rope_theta") or 10000.0
    w.add_rope_freq_base(float(rope))
    w.add_vocab_size(cfg["vocab_size"])
    w.add_file_type(int(gguf.LlamaFileType.MOSTLY_F16 if outtype == "f16" else gguf.LlamaFileType.ALL_F32))
    tied = cfg.get("tie_word_embeddings", False)
    for name, t in sd.items():
        if tied and name == "lm_head.weight":
            continue
        gname = hf_to_gguf_name(name, n_layers)
        if gname is None:
            continue
        if name.endswith("q_proj.weight"):
            t = _permute_qk(t, n_head)
        elif name.endswith("k_proj.weight"):
            t = _permute_qk(t, n_kv)
        if t.dim() >= 2 and outtype == "f16":
            arr = t.to(torch.float16).numpy()
        else:
            arr = t.to(torch.float32).numpy()
        w.add_tensor(gname, np.ascontiguousarray(arr))
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return out_file


def assign_devices(sizes: List[int], n_dev: int) -> List[int]:
    """Greedy largest-first placement of tensors on devices (GGUF tensors are
    independent, so they are spread over the GPUs with no collective)."""
    load = [0] * n_dev
    out = [0] * len(sizes)
    for i in sorted(range(len(sizes)), key=lambda i: -sizes[i]):
        d = min(range(n_dev), key=lambda d: load[d])
        out[i] = d
        load[d] += sizes[i]
    return out


def quantize_gguf(input_gguf: str, out_file: str, ftype: str, devices: Optional[List[int]] = None) -> str:
    """f16 GGUF -> `ftype` GGUF.  Every 2-D weight goes host -> HBM -> CUDA packer -> host."""
    import gguf
    if ftype not in BASE_TYPE:
        raise ValueError(f"unknown GGUF level {ftype}")
    if not torch.cuda.is_available():
        raise RuntimeError("GGUF quantization runs on CUDA kernels; no CUDA device is visible (no CPU fallback)")
    devices = devices if devices is not None else list(range(torch.cuda.device_count()))
    r = gguf.GGUFReader(input_gguf)
    arch = "llama"
    w = gguf.GGUFWriter(out_file, arch)
    n_layers = n_head = n_kv = 0
    for key, field in r.fields.items():
        if key.startswith("GGUF.") or key == "general.architecture":
            continue
        if key == "general.file_type":
            continue
        val = field.contents()
        vt = field.types[0]
        if key.endswith(".block_count"):
            n_layers = int(val)
        if key.endswith(".attention.head_count"):
            n_head = int(val)
        if key.endswith(".attention.head_count_kv"):
            n_kv = int(val)
        if vt == gguf.GGUFValueType.ARRAY:
            w.add_array(key, val)
        else:
            w.add_key_value(key, val, vt)
    w.add_file_type(int(getattr(gguf.LlamaFileType, "MOSTLY_" + ftype)))
    names = [t.name for t in r.tensors]
    has_output = "output.weight" in names
    plan = []
    for t in r.tensors:
        shape = tuple(int(x) for x in reversed(t.shape))      # gguf stores dims innermost-first
        plan.append((t, shape, tensor_type(t.name, shape, ftype, n_layers, has_output, n_head, n_kv)))
    sizes = [int(np.prod(s)) for _, s, _ in plan]
    place = assign_devices(sizes, len(devices))
    pending = [None] * len(plan)
    groups: Dict[tuple, list] = {}            # (device, type, dtype) -> [(index, device tensor)]: one launch per group
    for i, ((t, shape, qt), d) in enumerate(zip(plan, place)):
        data = np.asarray(t.data)
        if qt == "F32":
            out = data.astype(np.float32) if data.dtype != np.float32 else data
            pending[i] = (t.name, out.reshape(shape), None, None)
        elif qt == "F16":
            pending[i] = (t.name, data.astype(np.float16).reshape(shape), None, None)
        else:
            dev = torch.device("cuda", devices[d])
            src = torch.from_numpy(np.ascontiguousarray(data.reshape(-1, shape[-1]))).to(dev, non_blocking=True)
            groups.setdefault((dev, qt, src.dtype), []).append((i, src))
    for (dev, qt, dt), items in groups.items():
        with torch.cuda.device(dev):
            ys = cabi.gguf_quantize_batch([x for _, x in items], qt, round_via_f16=(dt != torch.float16))
        for (i, _), y in zip(items, ys):
            t, shape, _ = plan[i]
            pending[i] = (t.name, y, shape, getattr(gguf.GGMLQuantizationType, qt))
    for name, arr, shape, raw in pending:
        if raw is None:
            w.add_tensor(name, np.ascontiguousarray(arr))
        else:
            packed = arr.cpu().numpy()
            byte_shape = tuple(shape[:-1]) + (packed.shape[-1],)
            w.add_tensor(name, packed.reshape(byte_shape), raw_dtype=raw)
    w.write_header_to_file()
    w.write_kv_data_to_file()
    w.write_tensors_to_file()
    w.close()
    return out_file
