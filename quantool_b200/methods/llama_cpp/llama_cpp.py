"""GGUF plugin (drop-in for ref/src/quantool/methods/llama_cpp/llama_cpp.py:15-267).

Same class, `QuantType` enum, constructor, `quantize()` control flow (one f16 base reused for a
list of levels, file naming `{model_name}-{LEVEL}.gguf`, `last_gguf`) and `_save_model_files`.
`_convert_hf` / `_quantize_gguf` run in process on the GPU (quantool_b200.engine.gguf_file)
instead of spawning `convert_hf_to_gguf.py` / `llama-quantize`.
"""
import shutil
import tempfile
from enum import Enum
from pathlib import Path
from typing import List, Optional, Union

from ...core.base import BaseQuantizer
from ...core.meta import TemplateQuantizationCard
from ...core.registry import QuantizerRegistry


class QuantType(Enum):
    Q2_K = "Q2_K"
    Q3_K_S = "Q3_K_S"
    Q3_K_M = "Q3_K_M"
    Q3_K_L = "Q3_K_L"
    Q4_0 = "Q4_0"
    Q4_1 = "Q4_1"
    Q4_K_S = "Q4_K_S"
    Q4_K_M = "Q4_K_M"
    Q5_0 = "Q5_0"
    Q5_1 = "Q5_1"
    Q5_K_S = "Q5_K_S"
    Q5_K_M = "Q5_K_M"
    Q6_K = "Q6_K"
    Q8_0 = "Q8_0"
    F16 = "f16"
    F32 = "f32"

    def __str__(self):
        return self.value


@QuantizerRegistry.register
class GGUF(BaseQuantizer):
    name = "gguf"
    supported_levels = list(QuantType)
    supports_multiple_levels = True

    template_card = TemplateQuantizationCard(
        title="GGUF",
        description="GGUF quantization using llama.cpp for efficient CPU and GPU inference",
        hyperparameters={"format": "gguf", "method": "gguf", "quantization_type": "Q4_K_M", "context_length": 2048},
        intended_use="Efficient inference on CPU and GPU with llama.cpp",
        limitations="Requires llama.cpp conversion tools and specific model architectures",
        citations=["https://github.com/ggml-org/llama.cpp"],
    )

    # lowercase on purpose, as in the reference (llama_cpp.py:57): "Q8_0" is NOT in this set, so
    # Q8_0 always goes f16 GGUF -> block quantizer (SURVEY §3.2 note)
    CONVERT_OUTTYPES = {"f32", "f16", "bf16", "q8_0", "tq1_0", "tq2_0", "auto"}

    def __init__(self, *args, llama_cpp_path: Optional[Union[str, Path]] = None, **kwargs) -> None:
        # The CLI hands the whole `quantization_config` to the constructor AND to quantize() (ref cli.py:201-203,
        # 345-350).  In the reference any key other than `llama_cpp_path` falls through to object.__init__ and raises
        # TypeError (tests/golden/cli_traces.json, case gguf_levels_and_ignored_calibration); here such keys - e.g.
        # `output_dir`, which quantize() consumes - are left to quantize().
        if "model_id" in kwargs:
            args = (kwargs.pop("model_id"),) + tuple(args)
        super().__init__(*args)
        if kwargs:
            self.logger.info(f"constructor ignores quantization_config keys {sorted(kwargs)} (quantize() receives them)")
        self.llama_cpp_path = Path(llama_cpp_path) if llama_cpp_path else None   # accepted, unused
        self.use_module_import = True
        self._check_dependencies()

    def _check_dependencies(self):
        """The reference looks for convert_hf_to_gguf.py and the llama-quantize binary and raises
        RuntimeError when missing (llama_cpp.py:71-114); here the dependency is the CUDA library."""
        from ... import cabi
        try:
            cabi.lib()
        except Exception as e:
            raise RuntimeError(f"Could not load the quantool_b200 CUDA library: {e}")
        self.convert_script = None
        self.quantize_bin = cabi.LIB_PATH

    def _ensure_output_directory(self, output_dir: Optional[Union[str, Path]]) -> Path:
        output_path = Path(output_dir or tempfile.mkdtemp())
        output_path.mkdir(parents=True, exist_ok=True)
        return output_path

    def _convert_hf(self, model_path: str, output_path: Union[str, Path], outtype: str) -> str:
        from ...engine import gguf_file
        if outtype not in ("f16", "f32"):
            raise NotImplementedError(f"direct conversion to outtype {outtype!r} is not implemented (f16, f32 are)")
        out_file = Path(output_path) / f"model.{outtype}.gguf"
        self.logger.info(f"Converting {model_path} -> {out_file}")
        return gguf_file.convert_hf_to_f16_gguf(model_path, str(out_file), outtype,
                                                require_tokenizer=getattr(self, "_require_tokenizer", True))

    def _quantize_gguf(self, input_gguf: Union[str, Path], output_path: Union[str, Path], quant: str) -> str:
        from ...engine import gguf_file
        model_name = Path(self.model_id).name if isinstance(self.model_id, str) else str(self.model_id)
        out_file = Path(output_path) / f"{model_name}-{quant}.gguf"
        self.logger.info(f"Quantizing {input_gguf} -> {out_file} ({quant})")
        return gguf_file.quantize_gguf(str(input_gguf), str(out_file), quant)

    def _validate_and_convert_level(self, level) -> QuantType:
        if not isinstance(level, QuantType):
            try:
                level = QuantType[level]
            except KeyError:
                try:
                    level = QuantType(level)
                except ValueError:
                    self.logger.warning(f"Invalid quantization level '{level}', defaulting to Q4_K_M.")
                    level = QuantType.Q4_K_M
        return level

    def quantize(self, model: Union[str, Path], level: Union[str, QuantType, List[Union[str, QuantType]]] = QuantType.Q4_K_M,
                 output_dir: Optional[Union[str, Path]] = None, **kwargs) -> Union[str, List[str]]:
        output_path = self._ensure_output_directory(output_dir)
        model_path = getattr(model, "name_or_path", str(model))
        # additive key: synthetic weight-only directories (benchmarks) may skip the tokenizer metadata; real models
        # always carry it, as the reference's convert_hf_to_gguf.py requires
        self._require_tokenizer = bool(kwargs.pop("require_tokenizer", True))
        self.source_model = model
        if isinstance(level, list):
            base_gguf = self._convert_hf(model_path, output_path, "f16")
            results = []
            for lvl in level:
                lvl_str = str(self._validate_and_convert_level(lvl))
                if lvl_str in self.CONVERT_OUTTYPES:
                    final = self._convert_hf(model_path, output_path, lvl_str)
                else:
                    final = self._quantize_gguf(base_gguf, output_path, lvl_str)
                results.append(final)
            self.last_gguf = results
            return results
        lvl = str(self._validate_and_convert_level(level))
        if lvl in self.CONVERT_OUTTYPES:
            final = self._convert_hf(model_path, output_path, lvl)
        else:
            base = self._convert_hf(model_path, output_path, "f16")
            final = self._quantize_gguf(base, output_path, lvl)
        self.last_gguf = final
        return final

    def _save_model_files(self, save_directory: Union[str, Path]):
        if hasattr(self, "last_gguf"):
            files = self.last_gguf if isinstance(self.last_gguf, list) else [self.last_gguf]
            for f in files:
                shutil.copy(f, save_directory)
        else:
            self.logger.warning("No GGUF file found, using default quantization method.")
            self.quantize(self.source_model, QuantType(self.template_card.hyperparameters["quantization_type"]),
                          save_directory)
