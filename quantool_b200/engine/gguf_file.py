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
    """(hf config dict, {name: host tensor}) from a local HF directory: `model.safetensors`, the shards listed in
    `model.safetensors.index.json` (or every `*.safetensors` file when there is no index), else `pytorch_model*.bin`."""
    from safetensors.torch import load_file
    with open(os.path.join(model_path, "config.json")) as f:
        cfg = json.load(f)
    sd = {}
    index = os.path.join(model_path, "model.safetensors.index.json")
    if os.path.exists(index):
        with open(index) as f:
            files = sorted(set(json.load(f)["weight_map"].values()))
    else:
        files = sorted(f for f in os.listdir(model_path) if f.endswith(".safetensors"))
    for fn in files:
        sd.update(load_file(os.path.join(model_path, fn)))
    if not files:
        bins = sorted(f for f in os.listdir(model_path) if f.startswith("pytorch_model") and f.endswith(".bin"))
        if not bins:
            raise FileNotFoundError(f"no .safetensors or pytorch_model*.bin files under {model_path}")
        for fn in bins:
            sd.update(torch.load(os.path.join(model_path, fn), map_location="cpu", weights_only=True))
    return cfg, sd


def _tokenizer_pre(model_path: str) -> str:
    """`tokenizer.ggml.pre` (which pre-tokenizer regex llama.cpp applies).  convert_hf_to_gguf.py identifies it by
    hashing the tokenization of a fixed probe text against a table; that table is not on this box, so the family is
    read off the pre-tokenizer description in tokenizer.json instead (the regex IS what the hash fingerprints)."""
    try:
        with open(os.path.join(model_path, "tokenizer.json")) as f:
            pre = json.load(f).get("pre_tokenizer") or {}
    except Exception:
        return "default"
    kinds, patterns, digits = [], [], False

    def walk(node):
        nonlocal digits
        if isinstance(node, dict):
            if "type" in node:
                kinds.append(node["type"])
            if node.get("type") == "Digits" and node.get("individual_digits"):
                digits = True
            pat = node.get("pattern")
            if isinstance(pat, dict):
                patterns.extend(str(v) for v in pat.values())
            for v in node.values():
                walk(v)
        elif isinstance(node, list):
            for v in node:
                walk(v)
    walk(pre)
    rx = " ".join(patterns)
    if "\\p{N}{1,3}" in rx and "(?i:" in rx:
        return "llama-bpe"                    # Llama-3 family
    if digits and "ByteLevel" in kinds:
        return "smollm"                       # SmolLM / SmolLM2: Digits(individual) + ByteLevel
    if "ByteLevel" in kinds:
        return "gpt-2"
    return "default"


def write_vocab(w, model_path: str, cfg: dict, require_tokenizer: bool = True) -> bool:
    """tokenizer.ggml.* metadata, as convert_hf_to_gguf.py LlamaModel.set_vocab writes it: sentencepiece
    (`tokenizer.model`) -> model "llama" with scores; byte-level BPE (`tokenizer.json`) -> model "gpt2" with merges;
    then the special-token ids / chat template through gguf.SpecialVocab.  llama.cpp refuses to load a file without
    `tokenizer.ggml.model`, so a missing tokenizer is an error unless the caller opts out (synthetic benchmarks)."""
    import gguf
    from gguf import vocab as gv
    n_vocab = int(cfg["vocab_size"])
    if os.path.exists(os.path.join(model_path, "tokenizer.model")):
        v = gv.SentencePieceVocab(gv.Path(model_path))
        tokens, scores, types = [], [], []
        for text, score, tt in v.all_tokens():
            tokens.append(text); scores.append(score); types.append(int(tt))
        while len(tokens) < n_vocab:
            tokens.append(f"[PAD{len(tokens)}]".encode()); scores.append(-1000.0); types.append(int(gguf.TokenType.UNUSED))
        w.add_tokenizer_model("llama")
        w.add_tokenizer_pre("default")
        w.add_token_list(tokens)
        w.add_token_scores(scores)
        w.add_token_types(types)
        gguf.SpecialVocab(model_path, n_vocab=len(tokens)).add_to_gguf(w)
        return True
    if os.path.exists(os.path.join(model_path, "tokenizer.json")):
        from transformers import AutoTokenizer
        tok = AutoTokenizer.from_pretrained(model_path)
        vocab = tok.get_vocab()
        rev = {i: t for t, i in vocab.items()}
        added = tok.get_added_vocab()
        dec = tok.added_tokens_decoder
        tokens, types = [], []
        for i in range(max(n_vocab, len(rev))):
            if i not in rev:
                tokens.append(f"[PAD{i}]"); types.append(int(gguf.TokenType.UNUSED))
                continue
            t = rev[i]
            if t in added:
                special = (i in dec and dec[i].special) or (t.startswith("<|") and t.endswith("|>"))
                types.append(int(gguf.TokenType.CONTROL if special else gguf.TokenType.USER_DEFINED))
            else:
                types.append(int(gguf.TokenType.NORMAL))
            tokens.append(t)
        w.add_tokenizer_model("gpt2")
        w.add_tokenizer_pre(_tokenizer_pre(model_path))
        w.add_token_list(tokens)
        w.add_token_types(types)
        gguf.SpecialVocab(model_path, load_merges=True, n_vocab=len(tokens)).add_to_gguf(w)
        return True
    if require_tokenizer:
        raise FileNotFoundError(f"{model_path} holds neither tokenizer.model nor tokenizer.json: the GGUF would lack "
                                f"tokenizer.ggml.* and llama.cpp could not load it (the reference's "
                                f"convert_hf_to_gguf.py fails here too); pass require_tokenizer=False only for "
                                f"synthetic weight-packing runs")
    return False


def _rope_freq_factors(cfg: dict, head_dim: int, rope_theta: float):
    """llama3 rope scaling travels as the `rope_freqs.weight` tensor (convert_hf_to_gguf LlamaModel.generate_extra_tensors)."""
    rs = cfg.get("rope_scaling") or {}
    if rs.get("rope_type", rs.get("type")) != "llama3":
        return None
    import math
    factor = float(rs.get("factor", 8.0))
    low, high = float(rs.get("low_freq_factor", 1.0)), float(rs.get("high_freq_factor", 4.0))
    old = float(rs.get("original_max_position_embeddings", 8192))
    out = []
    for k in range(0, head_dim, 2):
        freq = 1.0 / (rope_theta ** (k / head_dim))
        wavelen = 2 * math.pi / freq
        if wavelen < old / high:
            out.append(1.0)
        elif wavelen > old / low:
            out.append(factor)
        else:
            smooth = (old / wavelen - low) / (high - low)
            out.append(1.0 / ((1 - smooth) / factor + smooth))
    return np.asarray(out, dtype=np.float32)


def convert_hf_to_f16_gguf(model_path: str, out_file: str, outtype: str = "f16", require_tokenizer: bool = True) -> str:
    """HF checkpoint -> GGUF with 2-D tensors in fp16 (or fp32) and 1-D tensors in fp32, plus the metadata
    llama.cpp needs to load it (architecture hyper-parameters, rope, tokenizer)."""
    import gguf
    cfg, sd = load_hf_model(model_path)
    if cfg.get("model_type", "llama") != "llama":
        raise ValueError(f"convert: model_type={cfg.get('model_type')!r} is not supported (llama only)")
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
    head_dim = cfg.get("head_dim") or cfg["hidden_size"] // n_head
    w.add_rope_dimension_count(int(head_dim))
    rs = cfg.get("rope_scaling") or {}
    if rs.get("rope_type", rs.get("type")) == "linear" and "factor" in rs:
        w.add_rope_scaling_type(gguf.RopeScalingType.LINEAR)
        w.add_rope_scaling_factor(float(rs["factor"]))
    w.add_vocab_size(cfg["vocab_size"])
    w.add_file_type(int(gguf.LlamaFileType.MOSTLY_F16 if outtype == "f16" else gguf.LlamaFileType.ALL_F32))
    write_vocab(w, model_path, cfg, require_tokenizer)
    ff = _rope_freq_factors(cfg, int(head_dim), float(rope))
    if ff is not None:
        w.add_tensor("rope_freqs.weight", ff)
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
    if "general.quantization_version" not in r.fields:
        w.add_quantization_version(gguf.GGML_QUANT_VERSION)       # llama-quantize stamps every quantized file
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
