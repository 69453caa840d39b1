"""CLI / YAML schema: the same seven dataclasses and field names as the reference
(ref/src/quantool/args/quantization_args.py:5-171, common_args.py:5-48) so that its YAML
configs parse unchanged (`HfArgumentParser.parse_yaml_file(..., allow_extra_keys=False)`)."""
from dataclasses import dataclass, field
from typing import Any, Optional


@dataclass
class ModelArguments:
    model_id: str = field(metadata={"help": "Path, HF repo ID or alias of the pretrained model."})
    tokenizer_name: Optional[str] = field(default=None, metadata={"help": "Tokenizer name/path if different."})
    cache_dir: Optional[str] = field(default=None, metadata={"help": "Cache directory for model and tokenizer."})
    use_auth_token: bool = field(default=False, metadata={"help": "Use the HF auth token."})
    revision: Optional[str] = field(default=None, metadata={"help": "Model Git revision."})


@dataclass
class QuantizationArguments:
    method: str = field(default="gguf", metadata={"help": "Quantization method: gptq, awq, smoothquant, gguf."})
    quant_level: Optional[Any] = field(default=None, metadata={"help": "Level label or list of labels."})
    quantization_config: dict = field(default_factory=dict, metadata={"help": "Method specific configuration."})


@dataclass
class EvaluationArguments:
    enable_evaluation: bool = field(default=True, metadata={"help": "Enable evaluation after quantization."})
    eval_dataset: Optional[str] = field(default=None, metadata={"help": "Dataset to use for evaluation."})
    metrics: Any = field(default_factory=lambda: ["perplexity"], metadata={"help": "Metrics to compute."})


@dataclass
class ExportArguments:
    output_path: str = field(default="quantized_model", metadata={"help": "Path for the exported model."})
    push_to_hub: bool = field(default=False, metadata={"help": "Push the quantized model to the Hub."})
    repo_id: Optional[str] = field(default=None, metadata={"help": "Repository ID on the Hub."})
    private: Optional[bool] = field(default=None, metadata={"help": "Whether the Hub repository is private."})


@dataclass
class CalibrationArguments:
    dataset_id: Optional[str] = field(default=None, metadata={"help": "HF datasets id used for calibration."})
    dataset_path: Optional[str] = field(default=None, metadata={"help": "Local dataset path for calibration."})
    dataset_config: Optional[str] = field(default=None, metadata={"help": "Dataset config name."})
    split: Optional[str] = field(default="train", metadata={"help": "Dataset split."})
    dataset_cache_dir: Optional[str] = field(default=None, metadata={"help": "Cache directory for datasets."})
    sample_size: Optional[int] = field(default=1024, metadata={"help": "Number of calibration examples."})
    shuffle: bool = field(default=True, metadata={"help": "Shuffle before sampling."})
    dataset_seed: int = field(default=42, metadata={"help": "Seed for sampling/shuffling."})
    load_in_pipeline: bool = field(default=False, metadata={"help": "Load the dataset in the pipeline."})
    preprocess_fn: Optional[str] = field(default=None, metadata={"help": "module.func run on examples."})
    calibration_config: dict = field(default_factory=dict, metadata={"help": "Method specific calibration config."})


@dataclass
class CommonArguments:
    seed: int = field(default=42, metadata={"help": "Random seed."})
    device: Optional[str] = field(default=None, metadata={"help": "Device."})
    verbose: bool = field(default=False, metadata={"help": "Verbose output."})


@dataclass
class LoggingArguments:
    report_to: Optional[Any] = field(default=None, metadata={"help": "Experiment trackers."})
    experiment_name: Optional[str] = field(default=None, metadata={"help": "Experiment name."})
    log_level: str = field(default="INFO", metadata={"help": "Logging level."})
    save_logs: bool = field(default=True, metadata={"help": "Save logs to a file."})
    log_dir: Optional[str] = field(default=None, metadata={"help": "Log directory."})

    def __post_init__(self):
        if isinstance(self.report_to, str):
            self.report_to = [self.report_to]


ALL = (ModelArguments, QuantizationArguments, CalibrationArguments, EvaluationArguments, ExportArguments,
       CommonArguments, LoggingArguments)
