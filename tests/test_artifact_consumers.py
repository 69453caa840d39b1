"""Consumer side of the artifact contract (SURVEY.md 8b "Artifact contract", 8f row 1), on the CPU: what the
host-side writers put on disk is read back by the code that consumes such artifacts downstream -

  * transformers' `CompressedTensorsHfQuantizer` + the installed compressed-tensors decompressors for the
    llm-compressor family (`model.safetensors` + `config.json["quantization_config"]`),
  * transformers' GGUF loader (`modeling_gguf_pytorch_utils.load_gguf_checkpoint`, gguf-py underneath) for the
    f16 base file `convert_hf_to_f16_gguf` writes (tensor names, q/k permutation, llama.* keys, tokenizer.ggml.*),
  * vLLM's own `gptq_pack` / `awq_pack` (VLLM/model_executor/layers/quantization/utils/quant_utils.py:785-813)
    for the AutoGPTQ / AutoAWQ `qweight` views.

The tensors fed to the writers here come from the CPU oracle (these tests run without a GPU);
tests/test_zz_consumers_gpu.py repeats the first two with artifacts produced by the CUDA path."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tiny_shape(tied=True):
    from quantool_b200.engine import llama
    return llama.LlamaShape(256, 512, 2, 4, 2, 512, rope_theta=10000.0, tie_word_embeddings=tied)


def oracle_artifact(shape, sd, level, actorder):
    """Artifact tensors of every Linear from the CPU oracle (GPTQ on a random well-conditioned Hessian), in the key
    names of rows a6/a7; returns (tensors, {weight key: fake-quantized weight})."""
    from compressed_tensors.quantization import ActivationOrdering
    from oracle import gptq as og
    oargs = og.scheme_weight_args(level)
    if actorder == "group":
        oargs.actorder = ActivationOrdering.GROUP
    g = torch.Generator().manual_seed(0)
    tensors, fake = {}, {}
    for k, v in sd.items():
        if not k.endswith("proj.weight"):
            tensors[k] = v
            continue
        N, K = v.shape
        x = torch.randn((1, 4 * K, K), generator=g).to(torch.bfloat16)
        H, _ = og.accumulate_hessian(x, og.make_empty_hessian(K), 0)
        _, Wq, s, z, gi = og.quantize_weight(v, H, oargs)
        base = k[: -len(".weight")]
        zp = None if oargs.symmetric else z
        if level in ("W4A16", "W4A16_ASYM", "W8A16"):
            _, packed, zp_packed = og.compress_packed(Wq, s, zp, gi, oargs)
            tensors[base + ".weight_packed"] = packed
            tensors[base + ".weight_shape"] = torch.tensor([N, K], dtype=torch.int64)
            if zp is not None:
                tensors[base + ".weight_zero_point"] = zp_packed
            if gi is not None:
                tensors[base + ".weight_g_idx"] = gi.to(torch.int32)
        else:
            tensors[base + ".weight"] = og.compress_int8(Wq, s, zp, oargs)
        tensors[base + ".weight_scale"] = s
        fake[k] = Wq
    return tensors, fake


def check_loads_in_transformers(out_dir, reference_weights, tol, vocab):
    """`out_dir` loads through transformers' compressed-tensors quantizer both ways (kept compressed / decompressed
    on load); the decompressed Linear weights are `reference_weights` up to `tol` (relative, Frobenius) and both
    models produce the same logits."""
    from transformers import AutoModelForCausalLM
    from transformers.utils.quantization_config import CompressedTensorsConfig
    kept = AutoModelForCausalLM.from_pretrained(out_dir, dtype=torch.bfloat16)
    dec = AutoModelForCausalLM.from_pretrained(out_dir, dtype=torch.bfloat16,
                                               quantization_config=CompressedTensorsConfig(run_compressed=False))
    worst = 0.0
    for k, want in reference_weights.items():
        w = dec.get_submodule(k[: -len(".weight")]).weight.detach().float()
        assert w.shape == want.shape, k
        worst = max(worst, (torch.linalg.norm(w - want.float()) / torch.linalg.norm(want.float())).item())
    assert worst < tol, worst
    ids = torch.randint(0, vocab, (2, 16), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        a, b = kept(ids).logits.float(), dec(ids).logits.float()
    assert torch.isfinite(a).all()
    assert (a - b).abs().max().item() <= 1e-2 * b.abs().max().item()
    return worst


@pytest.mark.parametrize("level,actorder", [("W4A16", "group"), ("W4A16_ASYM", None), ("W8A16", None), ("W8A8", None)])
def test_compressed_tensors_artifact_loads_in_transformers(tmp_path, level, actorder):
    """`QuantizedModel.save_pretrained` (engine/artifacts.py; what `plugin.save_pretrained` calls, reference
    ref/src/quantool/methods/llm_compressor/base.py:188) writes a directory the HF loader accepts: the
    quantization_config block parses, every tensor key is consumed, and decompression returns the fake-quantized
    weights (to the bf16 rounding of the stored scale)."""
    from quantool_b200.engine import artifacts, llama, schemes
    shape = tiny_shape()
    sd = llama.random_state_dict(shape, seed=3)
    tensors, fake = oracle_artifact(shape, sd, level, actorder)
    a = schemes.resolve(level, actorder)
    fmt = artifacts.artifact_format(a.num_bits, level)
    qm = artifacts.QuantizedModel(shape.to_hf_config(), tensors, artifacts.quantization_config(level, actorder, fmt))
    out = str(tmp_path / "out")
    qm.save_pretrained(out)
    qc = json.load(open(os.path.join(out, "config.json")))["quantization_config"]
    assert qc["format"] == fmt and qc["quant_method"] == "compressed-tensors"
    check_loads_in_transformers(out, fake, 1e-2, shape.vocab_size)


@pytest.mark.parametrize("level,actorder", [("W4A16", "group"), ("W4A16", "weight"), ("W4A16_ASYM", None), ("W8A8", None),
                                            ("W8A16", None), ("W4A8", None)])
def test_quantization_config_block_equals_compressed_tensors_own_writer(tmp_path, level, actorder):
    """The `quantization_config` block against the installed compressed-tensors' own writer: a tiny HF Llama gets the
    preset scheme applied, is compressed by `ModelCompressor.compress_model` and described by
    `ModelCompressor.update_config` - what `save_pretrained(save_compressed=True)` runs for the reference
    (ref/src/quantool/methods/llm_compressor/base.py:188) - and the resulting block must be ours, key for key."""
    from compressed_tensors.compressors import ModelCompressor
    from compressed_tensors.quantization import (QuantizationConfig, QuantizationStatus, apply_quantization_config,
                                                 preset_name_to_scheme)
    from transformers import LlamaConfig, LlamaForCausalLM
    from quantool_b200.engine import artifacts, llama, schemes
    shape = llama.LlamaShape(128, 256, 1, 2, 1, 64, rope_theta=10000.0, tie_word_embeddings=True)
    cfg = LlamaConfig(**{k: v for k, v in shape.to_hf_config().items() if k not in ("model_type", "architectures")})
    model = LlamaForCausalLM(cfg).to(torch.bfloat16)
    scheme = preset_name_to_scheme(level, ["Linear"])
    if actorder:
        scheme.weights.actorder = "static" if actorder == "weight" else actorder
    apply_quantization_config(model, QuantizationConfig(config_groups={"group_0": scheme}, ignore=["lm_head"]))
    for mod in model.modules():
        if hasattr(mod, "quantization_status"):
            mod.quantization_status = QuantizationStatus.FROZEN
    compressor = ModelCompressor.from_pretrained_model(model)
    model.config.save_pretrained(str(tmp_path))
    compressor.compress_model(model)
    compressor.update_config(str(tmp_path))
    want = json.load(open(tmp_path / "config.json"))["quantization_config"]
    a = schemes.resolve(level, actorder)
    fmt = artifacts.artifact_format(a.num_bits, level)
    assert artifacts.quantization_config(level, actorder, fmt) == want


def test_sharded_artifact_loads_in_transformers(tmp_path):
    """Above the shard limit (5 GB by default, the reference's transformers 4.56.2 default; 200 kB here) the writer
    emits `model-0000i-of-0000n.safetensors` + `model.safetensors.index.json` like `save_pretrained` does, and the
    HF loader reassembles the packed tensors from the shards."""
    from quantool_b200.engine import artifacts, llama, schemes
    shape = tiny_shape()
    sd = llama.random_state_dict(shape, seed=3)
    tensors, fake = oracle_artifact(shape, sd, "W4A16", "group")
    fmt = artifacts.artifact_format(schemes.resolve("W4A16", "group").num_bits, "W4A16")
    qm = artifacts.QuantizedModel(shape.to_hf_config(), tensors, artifacts.quantization_config("W4A16", "group", fmt))
    out = str(tmp_path / "out")
    qm.save_pretrained(out, max_shard_size=200_000)
    files = sorted(os.listdir(out))
    shards = [f for f in files if f.startswith("model-") and f.endswith(".safetensors")]
    assert len(shards) >= 2 and "model.safetensors" not in files and "model.safetensors.index.json" in files
    assert shards[0] == f"model-00001-of-{len(shards):05d}.safetensors"
    index = json.load(open(os.path.join(out, "model.safetensors.index.json")))
    assert set(index["weight_map"]) == set(tensors) and set(index["weight_map"].values()) == set(shards)
    assert index["metadata"]["total_size"] == sum(t.numel() * t.element_size() for t in tensors.values())
    check_loads_in_transformers(out, fake, 1e-2, shape.vocab_size)
    # and the loader of this repo reads a sharded directory back
    from quantool_b200.engine import gguf_file
    _, back = gguf_file.load_hf_model(out)
    assert set(back) == set(tensors) and all(torch.equal(back[k], tensors[k]) for k in tensors)


def test_recipe_yaml_is_written_next_to_the_weights(tmp_path):
    import yaml
    from quantool_b200.engine import artifacts
    from quantool_b200.methods.llm_compressor.base import Modifier
    qm = artifacts.QuantizedModel({"model_type": "llama"}, {"w": torch.zeros(4)}, {"quant_method": "compressed-tensors"})
    qm.recipe = [Modifier(kind="smoothquant", smoothing_strength=0.8), Modifier(kind="gptq", scheme="W8A8", dampening_frac=0.05)]
    qm.save_pretrained(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["config.json", "model.safetensors", "recipe.yaml"]
    mods = yaml.safe_load(open(tmp_path / "recipe.yaml"))["default_stage"]["default_modifiers"]
    assert list(mods) == ["SmoothQuantModifier", "GPTQModifier"] and mods["SmoothQuantModifier"] == {"smoothing_strength": 0.8}
    assert mods["GPTQModifier"] == {"targets": ["Linear"], "ignore": ["lm_head"], "scheme": "W8A8", "block_size": 128,
                                    "dampening_frac": 0.05}


def unpermute_qk(w, n_head):
    """Inverse of gguf_file._permute_qk (what a GGUF consumer applies to attn_q / attn_k)."""
    out_dim = w.shape[0]
    return w.reshape(n_head, out_dim // n_head // 2, 2, *w.shape[1:]).swapaxes(1, 2).reshape(w.shape)


def check_quantized_gguf_in_transformers(path, sd, shape, ftype):
    """A quantized GGUF file read by transformers' GGUF loader: every tensor name and block type is accepted, and the
    dequantized values are the C oracle's pack -> dequantize of the fp16-rounded source weight (rope permutation of
    q/k applied before packing and undone by the consumer)."""
    import numpy as np
    from transformers import LlamaConfig, LlamaForCausalLM
    from transformers.modeling_gguf_pytorch_utils import load_gguf_checkpoint
    from quantool_b200.engine import gguf_file
    from oracle import ggml_quants as oq
    cfg = load_gguf_checkpoint(path, return_tensors=False)["config"]
    with torch.device("meta"):
        skeleton = LlamaForCausalLM(LlamaConfig(**{k: v for k, v in cfg.items() if k != "model_type"}))
    got = load_gguf_checkpoint(path, return_tensors=True, model_to_load=skeleton)["tensors"]
    tied = shape.tie_word_embeddings
    want_keys = {k for k in sd if not (tied and k == "lm_head.weight")}
    assert set(got) == want_keys
    types = set()
    for k in sorted(want_keys):
        w = sd[k]
        heads = shape.num_attention_heads if k.endswith("q_proj.weight") else \
            shape.num_key_value_heads if k.endswith("k_proj.weight") else 0
        if heads:
            w = gguf_file._permute_qk(w, heads)
        gname = gguf_file.hf_to_gguf_name(k, shape.num_hidden_layers)
        qt = gguf_file.tensor_type(gname, tuple(w.shape), ftype, shape.num_hidden_layers, not tied,
                                   shape.num_attention_heads, shape.num_key_value_heads)
        types.add(qt)
        if qt == "F32":
            ref = w.float()
        elif qt == "F16":
            ref = w.half().float()
        else:
            x = oq.round_f16(w.float().numpy())
            ref = torch.from_numpy(np.asarray(oq.dequantize(oq.quantize(x, qt), qt, x.shape[-1]), dtype=np.float32))
            ref = ref.reshape(w.shape)
        if heads:
            ref = unpermute_qk(ref, heads)
        assert torch.equal(torch.as_tensor(got[k]).float(), ref), (k, qt)
    return types


@pytest.mark.parametrize("tied", [True, False])
def test_f16_gguf_loads_in_transformers(tmp_path, tied):
    """The f16 base GGUF (`convert_hf_to_f16_gguf`, standing in for llama.cpp's convert_hf_to_gguf.py at
    ref/src/quantool/methods/llama_cpp/llama_cpp.py:126-161) read by transformers' GGUF loader: the config it
    reconstructs from the llama.* keys is the source config, every tensor maps back to its HF name with the q/k
    rope permutation undone and fp16-exact values, and the tokenizer rebuilt from tokenizer.ggml.* encodes like the
    source tokenizer."""
    from safetensors.torch import save_file
    from transformers import AutoTokenizer, LlamaConfig, LlamaForCausalLM
    from transformers.modeling_gguf_pytorch_utils import load_gguf_checkpoint
    from quantool_b200.engine import gguf_file, llama
    from _tiny import write_tiny_tokenizer
    shape = tiny_shape(tied)
    sd = llama.random_state_dict(shape, seed=3)
    mdir = tmp_path / "tiny-llama"
    mdir.mkdir()
    save_file(sd, str(mdir / "model.safetensors"), metadata={"format": "pt"})
    json.dump(shape.to_hf_config(), open(mdir / "config.json", "w"))
    write_tiny_tokenizer(str(mdir))
    gdir = tmp_path / "gguf"
    gdir.mkdir()
    f = gguf_file.convert_hf_to_f16_gguf(str(mdir), str(gdir / "tiny-llama-F16.gguf"), "f16")
    meta = load_gguf_checkpoint(f, return_tensors=False)
    cfg = meta["config"]
    src = shape.to_hf_config()
    for k in ("hidden_size", "intermediate_size", "num_hidden_layers", "num_attention_heads", "num_key_value_heads",
              "vocab_size"):
        assert cfg[k] == src[k], k
    assert cfg["model_type"] == "llama" and cfg["tie_word_embeddings"] == tied
    assert abs(cfg["rope_theta"] - src["rope_theta"]) < 1e-3 and abs(cfg["rms_norm_eps"] - src["rms_norm_eps"]) < 1e-9
    with torch.device("meta"):
        skeleton = LlamaForCausalLM(LlamaConfig(**{k: v for k, v in cfg.items() if k != "model_type"}))
    got = load_gguf_checkpoint(f, return_tensors=True, model_to_load=skeleton)["tensors"]
    want = {k: v for k, v in sd.items() if not (tied and k == "lm_head.weight")}
    assert set(got) == set(want)
    for k, v in want.items():
        assert torch.equal(torch.as_tensor(got[k]).float(), v.half().float() if v.dim() == 2 else v.float()), k
    t_gguf = AutoTokenizer.from_pretrained(str(gdir), gguf_file=os.path.basename(f))
    t_src = AutoTokenizer.from_pretrained(str(mdir))
    text = "the thin thing in the inn"
    ids = t_gguf(text, add_special_tokens=False).input_ids
    assert ids == t_src(text, add_special_tokens=False).input_ids and len(ids) < len(text)   # merges applied
    # special-token ids as the file carries them (transformers 5.5 rebuilds a CodeLlamaTokenizer from a llama-arch
    # GGUF and assigns its own eos there, so the ids are read from the parsed tokenizer.ggml.* block)
    assert meta["tokenizer"]["bos_token_id"] == t_src.bos_token_id == t_gguf.bos_token_id
    assert meta["tokenizer"]["eos_token_id"] == t_src.eos_token_id
    assert meta["tokenizer"]["tokens"][t_src.eos_token_id] == t_src.eos_token


def test_autogptq_autoawq_views_equal_vllm_packers(tmp_path):
    """`autogptq_view` / `autoawq_view` (engine/artifacts.py) against the packers vLLM's own kernel tests use to build
    GPTQ / AWQ checkpoints.  vLLM is imported in a child process so that its extension modules stay out of this one."""
    from compressed_tensors.compressors.pack_quantized.helpers import pack_to_int32
    from quantool_b200.engine import artifacts
    N, K, gs = 64, 512, 128
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int8)
    scale = torch.rand((N, K // gs), generator=g).to(torch.bfloat16)
    zp = torch.randint(-8, 8, (N, K // gs), generator=g, dtype=torch.int8)
    packed = pack_to_int32(codes, 4)
    gv = artifacts.autogptq_view(packed, scale, zp, None, 4, K, gs)
    av = artifacts.autoawq_view(packed, scale, zp, K)
    torch.save({"u": (codes.to(torch.int32) + 8).t().contiguous(), "zu": (zp.to(torch.int32) + 8).t().contiguous(),
                "gptq_qweight": gv["qweight"], "gptq_qzeros": gv["qzeros"], "awq_qweight": av["qweight"],
                "awq_qzeros": av["qzeros"]}, tmp_path / "views.pt")
    code = r"""
import sys, torch
from vllm.model_executor.layers.quantization.utils import quant_utils as q
print("VLLM_IMPORTED")
d = torch.load(sys.argv[1])
u, zu = d["u"], d["zu"]                       # [K, N] and [G, N] unsigned 4-bit values
K, N = u.shape
G = zu.shape[0]
assert torch.equal(q.gptq_pack(u, 4, K, N), d["gptq_qweight"]), "gptq qweight"
assert torch.equal(q.pack_cols(zu, 4, G, N), d["gptq_qzeros"]), "gptq qzeros"
assert torch.equal(q.awq_pack(u, 4, K, N), d["awq_qweight"]), "awq qweight"
assert torch.equal(q.awq_pack(zu, 4, G, N), d["awq_qzeros"]), "awq qzeros"
print("VIEWS_OK")
"""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, "-c", code, str(tmp_path / "views.pt")], capture_output=True, text=True,
                       timeout=600, env=env, cwd=str(tmp_path))
    if "VLLM_IMPORTED" not in r.stdout:
        pytest.skip("vLLM is not importable here: " + (r.stderr or "").strip().splitlines()[-1][:200] if r.stderr else "vLLM missing")
    assert "VIEWS_OK" in r.stdout, (r.stdout + r.stderr)[-2000:]


@pytest.mark.parametrize("method,level,mk", [("gptq", "W4A16", {"actorder": "group"}), ("smoothquant", "W8A8", {}),
                                             ("awq", "W4A16_ASYM", {})])
def test_plugin_host_path_end_to_end_with_a_stub_engine(tmp_path, monkeypatch, method, level, mk):
    """Everything `Plugin.quantize()` does on the HOST - key routing, recipe, checkpoint and calibration loading, the
    call into the engine, quantization_config, artifact directory (weights, config, recipe.yaml, tokenizer files),
    model card, `save_pretrained` - with only `engine.pipeline.quantize_model_*` replaced: here (no GPU) a stub that
    returns artifact tensors computed by the CPU oracle.  The directory it leaves must load in transformers.  The
    GPU twin of this test (tests/test_zz_consumers_gpu.py) runs the same flow on the CUDA engine."""
    if torch.cuda.is_available():
        pytest.skip("CPU stand-in for the GPU twin")
    import yaml
    from safetensors.torch import save_file
    import quantool_b200.methods  # noqa: F401
    from quantool_b200 import QuantizerRegistry
    from quantool_b200.engine import llama, pipeline
    from _tiny import write_tiny_tokenizer
    shape = tiny_shape()
    sd = llama.random_state_dict(shape, seed=3)
    mdir = tmp_path / "tiny-llama"
    mdir.mkdir()
    save_file(sd, str(mdir / "model.safetensors"), metadata={"format": "pt"})
    json.dump(shape.to_hf_config(), open(mdir / "config.json", "w"))
    write_tiny_tokenizer(str(mdir))
    calls = []

    def engine(kind):
        def run(shape_, host_sd, token_ids, args, dev, fmt="pack-quantized", **kw):
            calls.append((kind, fmt, kw, token_ids))
            assert set(host_sd) == set(sd) and args.num_bits == (8 if level == "W8A8" else 4)
            tensors, _ = oracle_artifact(shape_, host_sd, level, mk.get("actorder"))
            return pipeline.ModelQuantResult(tensors=tensors)
        return run
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(pipeline, "quantize_model_gptq", engine("gptq"))
    monkeypatch.setattr(pipeline, "quantize_model_awq", engine("awq"))
    ids = torch.randint(0, shape.vocab_size, (12, 128), generator=torch.Generator().manual_seed(1234))
    q = QuantizerRegistry.create(method, model_id="org/tiny-llama")
    out = q.quantize(model=str(mdir), level=level, dataset=ids, num_calibration_samples=8, max_seq_length=96,
                     output_dir=str(tmp_path / "out"), method_kwargs=mk)
    monkeypatch.undo()                       # the consumers below see the real (CUDA-less) torch again
    (kind, fmt, kw, token_ids), = calls
    assert kind == ("awq" if method == "awq" else "gptq") and tuple(token_ids.shape) == (8, 96)
    assert fmt == ("int-quantized" if level == "W8A8" else "pack-quantized")
    if method == "smoothquant":
        assert kw["smooth_strength"] == 0.5
    files = sorted(os.listdir(out))
    assert {"config.json", "model.safetensors", "recipe.yaml", "tokenizer.json", "tokenizer_config.json"} <= set(files)
    mods = yaml.safe_load(open(os.path.join(out, "recipe.yaml")))["default_stage"]["default_modifiers"]
    assert list(mods) == {"gptq": ["GPTQModifier"], "awq": ["AWQModifier"], "smoothquant": ["SmoothQuantModifier", "GPTQModifier"]}[method]
    check_loads_in_transformers(out, {}, 1.0, shape.vocab_size)
    q.save_model_card(str(tmp_path / "saved"))
    q.save_pretrained(str(tmp_path / "saved"))
    assert {"README.md", "config.json", "model.safetensors", "recipe.yaml"} <= set(os.listdir(tmp_path / "saved"))


def test_only_the_writer_rank_touches_the_output_directory(tmp_path):
    """Multi-process runs: every rank builds a QuantizedModel, rank 0 alone holds the tensors and writes."""
    from quantool_b200.engine import artifacts
    qm = artifacts.QuantizedModel({"model_type": "llama"}, {}, {"quant_method": "compressed-tensors"}, writer=False)
    qm.recipe = []
    qm.save_pretrained(str(tmp_path / "out"))
    assert not os.path.exists(tmp_path / "out")
