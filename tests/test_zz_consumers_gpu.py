"""GPU twin of tests/test_artifact_consumers.py: the artifacts the CUDA path writes through the plugin API are read
back by their downstream consumers (transformers' compressed-tensors quantizer; transformers' GGUF loader).
Runs last in the suite (file name) - it exercises third-party loaders, not kernels."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_model(tmp_path, shape):
    from safetensors.torch import save_file
    from quantool_b200.engine import llama
    from _tiny import write_tiny_tokenizer
    sd = llama.random_state_dict(shape, seed=3)
    d = tmp_path / "tiny-llama"
    d.mkdir()
    save_file(sd, str(d / "model.safetensors"), metadata={"format": "pt"})
    json.dump(shape.to_hf_config(), open(d / "config.json", "w"))
    write_tiny_tokenizer(str(d))
    return sd, str(d)


@pytest.mark.parametrize("method,level,mk", [("gptq", "W4A16", {"actorder": "group"}), ("smoothquant", "W8A8", {})])
def test_plugin_artifact_loads_in_transformers(tmp_path, method, level, mk):
    """GPTQ.quantize() / SmoothQuant.quantize() output directory -> AutoModelForCausalLM.from_pretrained, kept
    compressed and decompressed on load.  GPTQ: the decompressed weights are the source weights up to the 4-bit
    quantization error (SmoothQuant stores folded weights W*s, so only loading and the two forwards are checked)."""
    import quantool_b200.methods  # noqa: F401
    from quantool_b200 import QuantizerRegistry
    from test_artifact_consumers import check_loads_in_transformers, tiny_shape
    shape = tiny_shape()
    sd, path = _write_model(tmp_path, shape)
    ids = torch.randint(0, shape.vocab_size, (12, 128), generator=torch.Generator().manual_seed(1234))
    q = QuantizerRegistry.create(method, model_id="org/tiny-llama")
    out = q.quantize(model=path, level=level, dataset=ids, num_calibration_samples=8, max_seq_length=96,
                     output_dir=str(tmp_path / "out"), method_kwargs=mk)
    ref = {k: v for k, v in sd.items() if k.endswith("proj.weight")} if method == "gptq" else {}
    worst = check_loads_in_transformers(out, ref, 0.25, shape.vocab_size)
    assert method != "gptq" or worst > 1e-3          # it really is the quantized weight that came back


def test_plugin_gguf_loads_in_transformers(tmp_path):
    """GGUF.quantize(level=[Q8_0, Q4_K_M]) files -> transformers' GGUF loader: dequantized tensors equal the C
    oracle's pack -> dequantize of the same weights, for every block type the level's table assigns."""
    import quantool_b200.methods  # noqa: F401
    from quantool_b200 import QuantizerRegistry
    from quantool_b200.engine import llama
    from test_artifact_consumers import check_quantized_gguf_in_transformers
    shape = llama.LlamaShape(512, 1024, 2, 8, 4, 512, rope_theta=10000.0, tie_word_embeddings=True)
    sd, path = _write_model(tmp_path, shape)
    q = QuantizerRegistry.create("gguf", model_id="org/tiny-llama")
    outs = q.quantize(model=path, level=["Q8_0", "Q4_K_M"], output_dir=str(tmp_path / "gg"))
    assert check_quantized_gguf_in_transformers(outs[0], sd, shape, "Q8_0") >= {"Q8_0", "F32"}
    assert check_quantized_gguf_in_transformers(outs[1], sd, shape, "Q4_K_M") >= {"Q4_K", "Q6_K", "F32"}
