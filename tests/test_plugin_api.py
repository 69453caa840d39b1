"""Plugin / registry API: same names, attributes, argument routing and error behaviour as the
reference (ref/src/quantool/core/{base,registry}.py, methods/*).  The GGUF flow
test mirrors ref/tests/quantool/methods/test_llama_cpp.py."""
from pathlib import Path
from unittest.mock import patch

import pytest
import torch

import quantool_b200.methods  # noqa: F401  (registers the plugins)
from quantool_b200 import BaseQuantizer, QuantizerRegistry, TemplateQuantizationCard
from quantool_b200.core.registry import Registry
from quantool_b200.methods.llama_cpp.llama_cpp import GGUF, QuantType
from quantool_b200.methods.llm_compressor.base import Modifier


def test_four_methods_registered_with_reference_attributes():
    assert set(QuantizerRegistry.list()) == {"gguf", "gptq", "awq", "smoothquant"}
    gptq = QuantizerRegistry.create("gptq", model_id="org/model")
    assert gptq.supported_levels == ["W4A16", "W8A8", "INT8", "W8A16", "W4A16_ASYM", "W4A8"]
    assert QuantizerRegistry.create("awq", model_id="m").supported_levels == ["W4A16", "W4A16_ASYM", "W8A16"]
    assert QuantizerRegistry.create("smoothquant", model_id="m").supported_levels == ["W8A8", "INT8", "W4A8"]
    assert gptq.supports_multiple_levels is False and gptq.require_calibration() is True
    with patch.object(GGUF, "_check_dependencies"):
        g = QuantizerRegistry.create("gguf", model_id="m")
    assert g.supports_multiple_levels is True and g.require_calibration() is False
    assert [q.value for q in QuantType][-3:] == ["Q8_0", "f16", "f32"] and len(list(QuantType)) == 16
    assert isinstance(gptq.template_card, TemplateQuantizationCard)


def test_registry_errors():
    r = Registry()

    class NoName(BaseQuantizer):
        def quantize(self, model, level, **kw):
            return ""
    with pytest.raises(ValueError):
        r.register(NoName)

    class A(BaseQuantizer):
        name = "a"

        def quantize(self, model, level, **kw):
            super().quantize(model, level, **kw)
            return "ok"
    r.register(A)
    with pytest.raises(KeyError):
        r.register(A)
    with pytest.raises(KeyError):
        r.create("missing")
    a = r.create("a", model_id="x")
    assert a.model_id == "x" and r.list() == ["a"]
    with pytest.raises(ValueError):
        a.quantize("m", ["L1", "L2"])          # list level on a single-level method


def test_kwargs_routing_like_the_reference(tmp_path):
    q = QuantizerRegistry.create("gptq", model_id="org/model")
    seen = {}

    def fake_oneshot(**kw):
        seen.update(kw)
        return object()
    ids = torch.zeros((2, 8), dtype=torch.long)
    with patch.object(q, "_oneshot", side_effect=fake_oneshot):
        out = q.quantize(model="some/path", level="W4A16", dataset=ids, num_calibration_samples=2, max_seq_length=8,
                         output_dir=str(tmp_path / "o"), method_kwargs__dampening_frac=0.05,
                         method_kwargs={"actorder": "group"}, targets="ignored-top-level", llama_cpp_path="ignored")
    assert out == str((tmp_path / "o").resolve()) and q.last_output_dir == (tmp_path / "o").resolve()
    assert seen["num_calibration_samples"] == 2 and seen["max_seq_length"] == 8 and seen["model"] == "some/path"
    assert seen["save_compressed"] is True and seen["trust_remote_code_model"] is True
    rec = seen["recipe"]
    assert isinstance(rec, Modifier) and rec.kind == "gptq" and rec.scheme == "W4A16"
    assert rec.dampening_frac == 0.05 and rec.actorder == "group"
    assert rec.targets == "Linear" and rec.ignore == ["lm_head"]          # top-level `targets` is dropped


def test_default_output_dir_and_errors(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    q = QuantizerRegistry.create("gptq", model_id="org/model")
    with patch.object(q, "_oneshot", return_value=object()):
        out = q.quantize(model="p", level="W8A8", dataset=torch.zeros((1, 4), dtype=torch.long))
    assert out.endswith("output/gptq_org_model_W8A8")
    with pytest.raises(ValueError, match="calibration data"):
        q.quantize(model="p", level="W4A16")
    with pytest.raises(ValueError, match="not a valid compressed-tensors preset scheme"):
        q.quantize(model="p", level="W3A16", dataset=torch.zeros((1, 4), dtype=torch.long))
    with pytest.raises(ValueError, match="does not support multiple"):
        q.quantize(model="p", level=["W4A16", "W8A8"], dataset=torch.zeros((1, 4), dtype=torch.long))
    with pytest.raises(RuntimeError):
        q2 = QuantizerRegistry.create("awq", model_id="m")
        q2.save_pretrained(str(tmp_path / "s"))         # nothing quantized yet


def test_smoothquant_recipe_is_smoothing_then_gptq():
    q = QuantizerRegistry.create("smoothquant", model_id="m")
    rec, scheme = q._build_recipe(None, {"smoothing_strength": 0.8})
    assert scheme == "W8A8" and [m.kind for m in rec] == ["smoothquant", "gptq"]
    assert rec[0].smoothing_strength == 0.8 and rec[1].scheme == "W8A8"


def test_gguf_multiple_levels_returns_list(tmp_path):
    with patch.object(GGUF, "_check_dependencies"):
        quantizer = QuantizerRegistry.create("gguf", model_id="test/model")
    output_dir = tmp_path / "gguf"
    levels = ["Q4_K_M", QuantType.Q3_K_S]
    convert_calls, quantize_calls = [], []

    def fake_convert(model_path, out_path, outtype):
        convert_calls.append((model_path, out_path, outtype))
        assert outtype == "f16"
        return "base_f16.gguf"

    def fake_quantize(base_file, out_path, quant_level):
        quantize_calls.append((base_file, out_path, quant_level))
        return str(Path(out_path) / f"model-{quant_level}.gguf")

    with (patch.object(quantizer, "_ensure_output_directory", return_value=output_dir) as mock_ensure,
          patch.object(quantizer, "_convert_hf", side_effect=fake_convert) as mock_convert,
          patch.object(quantizer, "_quantize_gguf", side_effect=fake_quantize) as mock_quantize):
        artifacts = quantizer.quantize(model="hf-repo/model", level=levels)
    expected = [str(output_dir / "model-Q4_K_M.gguf"), str(output_dir / "model-Q3_K_S.gguf")]
    assert artifacts == expected and quantizer.last_gguf == expected
    assert mock_ensure.called
    mock_convert.assert_called_once()
    assert convert_calls == [("hf-repo/model", output_dir, "f16")]
    assert mock_quantize.call_count == 2
    assert quantize_calls == [("base_f16.gguf", output_dir, "Q4_K_M"), ("base_f16.gguf", output_dir, "Q3_K_S")]


def test_gguf_q8_0_goes_through_the_f16_base_and_bad_level_defaults(tmp_path):
    with patch.object(GGUF, "_check_dependencies"):
        q = QuantizerRegistry.create("gguf", model_id="org/My-Model")
    calls = []
    with (patch.object(q, "_convert_hf", side_effect=lambda m, o, t: calls.append(("c", t)) or "b.gguf"),
          patch.object(q, "_quantize_gguf", side_effect=lambda b, o, l: calls.append(("q", l)) or f"x-{l}.gguf")):
        q.quantize(model="p", level="Q8_0", output_dir=str(tmp_path))
        q.quantize(model="p", level="nonsense", output_dir=str(tmp_path))
        q.quantize(model="p", level="f16", output_dir=str(tmp_path))
    assert calls == [("c", "f16"), ("q", "Q8_0"), ("c", "f16"), ("q", "Q4_K_M"), ("c", "f16")]


def test_gguf_tensor_type_table_smollm_and_llama():
    """SmolLM2-135M (h=576) under Q4_K_M."""
    from quantool_b200.engine.gguf_file import tensor_type, use_more_bits
    more = [i for i in range(30) if use_more_bits(i, 30)]
    assert more == [0, 1, 2, 5, 8, 11, 14, 17, 20, 23, 26, 27, 28, 29]
    t = lambda n, s, f="Q4_K_M", L=30, out=False: tensor_type(n, s, f, L, out)
    assert t("blk.3.attn_q.weight", (576, 576)) == "Q5_0"            # Q4_K falls back (576 % 256 != 0)
    assert t("blk.0.attn_v.weight", (192, 576)) == "Q8_0"            # Q6_K falls back
    assert t("blk.3.attn_v.weight", (192, 576)) == "Q5_0"
    assert t("blk.0.ffn_down.weight", (576, 1536)) == "Q6_K"
    assert t("blk.3.ffn_down.weight", (576, 1536)) == "Q4_K"
    assert t("token_embd.weight", (49152, 576)) == "Q8_0"            # tied: acts as output, Q6_K -> Q8_0
    assert t("blk.0.attn_norm.weight", (576,)) == "F32"
    assert t("output.weight", (128256, 4096), L=32, out=True) == "Q6_K"
    assert t("token_embd.weight", (128256, 4096), L=32, out=True) == "Q4_K"
    assert t("blk.1.ffn_gate.weight", (14336, 4096), "Q8_0", 32, True) == "Q8_0"
    assert t("output.weight", (128256, 4096), "Q8_0", 32, True) == "Q8_0"


def test_gguf_tensor_type_table_low_bit_levels():
    """Q2_K / Q3_K_{S,M,L} per-tensor choices of llama.cpp llama_tensor_get_type (llama arch, no imatrix)."""
    from quantool_b200.engine.gguf_file import tensor_type
    t = lambda n, s, f, L=32, kv=8, h=32: tensor_type(n, s, f, L, True, h, kv)
    for f, base in (("Q2_K", "Q2_K"), ("Q3_K_S", "Q3_K"), ("Q3_K_M", "Q3_K"), ("Q3_K_L", "Q3_K")):
        assert t("blk.5.attn_q.weight", (4096, 4096), f) == base
        assert t("blk.5.ffn_gate.weight", (14336, 4096), f) == base
        assert t("output.weight", (128256, 4096), f) == "Q6_K"
        assert t("token_embd.weight", (128256, 4096), f) == base
    assert t("blk.5.attn_v.weight", (1024, 4096), "Q2_K") == "Q4_K"            # n_gqa = 4
    assert t("blk.5.attn_v.weight", (4096, 4096), "Q2_K", kv=32) == "Q3_K"
    assert t("blk.5.ffn_down.weight", (4096, 14336), "Q2_K") == "Q3_K"
    assert t("blk.5.attn_output.weight", (4096, 4096), "Q2_K") == "Q3_K"
    assert t("blk.5.attn_v.weight", (1024, 4096), "Q3_K_S") == "Q3_K"
    assert t("blk.1.attn_v.weight", (1024, 4096), "Q3_K_M") == "Q5_K"
    assert t("blk.2.attn_v.weight", (1024, 4096), "Q3_K_M") == "Q4_K"
    assert t("blk.1.ffn_down.weight", (4096, 14336), "Q3_K_M") == "Q5_K"       # i_layer < n_layer/16
    assert t("blk.2.ffn_down.weight", (4096, 14336), "Q3_K_M") == "Q4_K"
    assert t("blk.9.attn_output.weight", (4096, 4096), "Q3_K_M") == "Q4_K"
    assert t("blk.9.attn_v.weight", (1024, 4096), "Q3_K_L") == "Q5_K"
    assert t("blk.9.ffn_down.weight", (4096, 14336), "Q3_K_L") == "Q5_K"
    assert t("blk.9.attn_output.weight", (4096, 4096), "Q3_K_L") == "Q5_K"
    assert t("blk.9.attn_v.weight", (1024, 8192), "Q3_K_S", L=80, h=64) == "Q5_K"    # 70B: shared attn_v
    assert t("blk.0.attn_q.weight", (576, 576), "Q3_K_S", L=30, kv=3) == "IQ4_NL"    # 576 % 256 != 0
    assert t("blk.0.ffn_down.weight", (576, 1536), "Q2_K", L=30, kv=3) == "Q3_K"


def test_autogptq_and_autoawq_views_are_pure_repacks_cpu():
    """the AutoGPTQ / AutoAWQ key layouts are index shuffles of the compressed-tensors codes."""
    import torch
    from compressed_tensors.compressors.pack_quantized.helpers import pack_to_int32
    from quantool_b200.engine import artifacts
    N, K, gs = 32, 256, 128
    g = torch.Generator().manual_seed(0)
    codes = torch.randint(-8, 8, (N, K), generator=g, dtype=torch.int8)
    packed = pack_to_int32(codes, 4)
    scale = torch.rand((N, K // gs), generator=g).to(torch.bfloat16)
    zp = torch.randint(-8, 8, (N, K // gs), generator=g, dtype=torch.int8)
    u = codes.to(torch.int32) + 8
    v = artifacts.autogptq_view(packed, scale, zp, None, 4, K, gs)
    qw = v["qweight"]
    assert qw.shape == (K // 8, N)
    back = torch.stack([(qw >> (4 * i)) & 0xF for i in range(8)], dim=1).reshape(K, N).t()
    assert torch.equal(back, u)
    a = artifacts.autoawq_view(packed, scale, zp, K)
    assert a["qweight"].shape == (K, N // 8) and a["qzeros"].shape == (K // gs, N // 8) and a["scales"].shape == (K // gs, N)
    rev = [0, 4, 1, 5, 2, 6, 3, 7]                     # vLLM's AWQ unpack order
    nib = torch.stack([(a["qweight"] >> (4 * i)) & 0xF for i in range(8)], dim=2)      # [K, N/8, 8] in packed order
    assert torch.equal(nib[:, :, rev].reshape(K, N).t(), u)
    znib = torch.stack([(a["qzeros"] >> (4 * i)) & 0xF for i in range(8)], dim=2)
    assert torch.equal(znib[:, :, rev].reshape(K // gs, N).t(), zp.to(torch.int32) + 8)


# ---- round 2: config keys are honoured or rejected, never dropped (ADVICE) -------------------
def test_artifact_format_matches_compressed_tensors():
    """The format per preset is what the installed compressed-tensors infers for a Linear with that scheme
    (W8A16 is pack-quantized, W4A8 is int-quantized - not a function of num_bits alone)."""
    import torch
    from compressed_tensors.compressors.format import infer_module_format
    from compressed_tensors.quantization import preset_name_to_scheme
    from quantool_b200.engine import artifacts, schemes
    for level in ("W4A16", "W4A16_ASYM", "W8A16", "W8A8", "INT8", "W4A8"):
        want = infer_module_format(torch.nn.Linear, preset_name_to_scheme(level, ["Linear"])).value
        assert artifacts.artifact_format(schemes.resolve(level).num_bits, level) == want, level
    assert artifacts.artifact_format(8, "W8A16") == "pack-quantized"
    assert artifacts.artifact_format(4, "W4A8") == "int-quantized"


def test_unsupported_recipe_keys_raise():
    import pytest
    from quantool_b200.methods.llm_compressor.base import LLMCompressorQuantizer, Modifier
    chk = LLMCompressorQuantizer._check_supported
    chk([Modifier(kind="gptq", scheme="W4A16")], {})
    chk([Modifier(kind="gptq", scheme="W4A16", targets=["Linear"], sequential_targets=["LlamaDecoderLayer"],
                  ignore=["lm_head", "re:.*down_proj"])], {"sequential_targets": "LlamaDecoderLayer"})
    for bad in (dict(targets=["re:.*q_proj"]), dict(sequential_targets=["LlamaAttention"]), dict(mappings=[["a"], "b"])):
        with pytest.raises(ValueError):
            chk([Modifier(kind="gptq", scheme="W4A16", **bad)], {})
    with pytest.raises(ValueError):
        chk([Modifier(kind="awq", scheme="W4A16", ignore=["lm_head", "re:.*down_proj"])], {})
    for kw in ({"calibration_dataloader": object()}, {"pad_to_max_length": True}, {"sequential_targets": ["X"]}):
        with pytest.raises(ValueError):
            chk([Modifier(kind="gptq", scheme="W4A16")], kw)


def test_oneshot_keys_that_cannot_be_honoured_raise():
    import pytest
    from quantool_b200.methods.llm_compressor.base import LLMCompressorQuantizer, Modifier
    chk = LLMCompressorQuantizer._check_supported
    m = [Modifier(kind="gptq", scheme="W4A16")]
    chk(m, {"precision": "auto", "streaming": False, "batch_size": 4, "text_column": "body"})
    for bad in ({"data_collator": object()}, {"streaming": True}, {"precision": "float16"}, {"stage": "s"},
                {"recipe_args": {"a": 1}}, {"processor": object()}, {"pad_to_max_length": True}):
        with pytest.raises(ValueError):
            chk(m, bad)


def test_ignore_patterns():
    import pytest
    from quantool_b200.engine.pipeline import _ignored_linears
    assert _ignored_linears(("lm_head",)) == set()
    assert _ignored_linears(["lm_head", "re:.*down_proj"]) == {"mlp.down_proj"}
    assert _ignored_linears(["model.layers.3.self_attn.o_proj"]) == {"model.layers.3.self_attn.o_proj"}
    for bad in ("re:.*layers\\.0\\..*q_proj", "model.norm", "re:.*nothing"):
        with pytest.raises(ValueError):
            _ignored_linears([bad])


def test_calibration_rows_keep_their_own_length(tmp_path):
    """A short sample does not truncate the others; n defaults to 512; shuffled (seeded); dataset_path files load."""
    import json
    import torch
    from quantool_b200.methods.llm_compressor.gptq import GPTQ
    q = GPTQ("org/m")
    rows = [list(range(1, 1 + n)) for n in (5, 40, 40, 17, 40)]
    out = q._token_ids({"dataset": rows, "shuffle_calibration_samples": False}, None)
    assert isinstance(out, list) and [int(r.numel()) for r in out] == [5, 40, 40, 17, 40]
    out = q._token_ids({"dataset": rows, "max_seq_length": 16, "num_calibration_samples": 3,
                        "shuffle_calibration_samples": False}, None)
    assert [int(r.numel()) for r in out] == [5, 16, 16]
    t = torch.arange(600 * 8).reshape(600, 8)
    out = q._token_ids({"dataset": t}, None)
    assert isinstance(out, torch.Tensor) and out.shape == (512, 8)
    assert not torch.equal(out, t[:512]) and torch.equal(out, q._token_ids({"dataset": t}, None))   # shuffled, repeatable
    p = tmp_path / "calib.jsonl"
    with open(p, "w") as f:
        for r in rows:
            f.write(json.dumps({"input_ids": r}) + "\n")
    out = q._token_ids({"dataset_path": str(p), "shuffle_calibration_samples": False}, None)
    assert [int(r.numel()) for r in out] == [5, 40, 40, 17, 40]
    import pytest
    with pytest.raises(ValueError):
        q._token_ids({"dataset_path": "org/some-hub-dataset"}, None)


def test_llama_config_validation_and_llama3_rope():
    import pytest
    import torch
    from quantool_b200.engine import llama
    base = llama.SHAPES["llama-3.2-1b"].to_hf_config()
    assert llama.LlamaShape.from_hf_config(base).rope_scaling is None
    for bad in ({"model_type": "qwen2"}, {"architectures": ["MistralForCausalLM"]}, {"attention_bias": True},
                {"mlp_bias": True}, {"sliding_window": 4096}, {"rope_scaling": {"rope_type": "yarn", "factor": 4.0}},
                {"hidden_act": "gelu"}):
        with pytest.raises(ValueError):
            llama.LlamaShape.from_hf_config({**base, **bad})
    rs = {"factor": 32.0, "high_freq_factor": 4.0, "low_freq_factor": 1.0, "original_max_position_embeddings": 8192,
          "rope_type": "llama3"}
    shape = llama.LlamaShape.from_hf_config({**base, "rope_scaling": rs})
    assert shape.rope_scaling["rope_type"] == "llama3" and shape.to_hf_config()["rope_scaling"]["factor"] == 32.0
    cos, sin = llama.rope_tables(shape, 64, "cpu", torch.float32)
    # reference: transformers' own llama3 frequency computation
    from transformers import LlamaConfig
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS
    hf = LlamaConfig(**{k: v for k, v in base.items() if k not in ("architectures", "torch_dtype")}, rope_scaling=rs)
    inv, _ = ROPE_INIT_FUNCTIONS["llama3"](hf, "cpu")
    f = torch.outer(torch.arange(64, dtype=torch.float32), inv)
    want = torch.cat((f, f), dim=-1)
    assert torch.allclose(cos, want.cos(), atol=1e-6) and torch.allclose(sin, want.sin(), atol=1e-6)
    plain, _ = llama.rope_tables(llama.SHAPES["llama-3.2-1b"], 64, "cpu", torch.float32)
    assert not torch.allclose(cos, plain, atol=1e-3)      # the scaling matters below 2048 positions too


def test_yaml_recipes_round_trip_and_unknown_fields_raise(tmp_path):
    """`recipe=` accepts llm-compressor's YAML layout (and the `recipe.yaml` this engine writes); nothing in it is
    silently dropped."""
    import pytest
    from quantool_b200.engine import artifacts
    from quantool_b200.methods.llm_compressor.base import Modifier, parse_recipe
    mods = [Modifier(kind="smoothquant", smoothing_strength=0.8),
            Modifier(kind="gptq", scheme="W8A8", dampening_frac=0.05, actorder="group", ignore=["lm_head", "re:.*down_proj"])]
    text = artifacts.recipe_yaml(mods)
    assert parse_recipe(text) == mods
    p = tmp_path / "recipe.yaml"
    p.write_text(text)
    assert parse_recipe(str(p)) == mods and parse_recipe(mods) == mods and parse_recipe(mods[1]) == [mods[1]]
    upstream_style = """
quant_stage:
  quant_modifiers:
    GPTQModifier:
      targets: [Linear]
      ignore: [lm_head]
      scheme: W4A16
      dampening_frac: 0.1
"""
    (m,) = parse_recipe(upstream_style)
    assert m.kind == "gptq" and m.scheme == "W4A16" and m.dampening_frac == 0.1 and m.targets == "Linear"
    with pytest.raises(ValueError, match="unsupported recipe fields"):
        parse_recipe("s:\n  g_modifiers:\n    GPTQModifier:\n      scheme: W4A16\n      offload_hessians: true\n")
    with pytest.raises(ValueError, match="not implemented"):
        parse_recipe("s:\n  g_modifiers:\n    SparseGPTModifier:\n      sparsity: 0.5\n")
    with pytest.raises(TypeError):
        parse_recipe(object())
