"""Export / calibration mixins.

Same method names and call order as ref/src/quantool/core/helpers/export_mixin.py:62-139
and calibration_mixin.py:4-30.  Hub upload needs network and is outside the hot path
(SURVEY.md §2 row 8); ``push_to_hub`` keeps the signature and imports huggingface_hub lazily.
"""
import os
import tempfile
from typing import Optional, Union


class ExportMixin:
    def _save_model_files(self, save_directory: Union[str, os.PathLike]):
        raise NotImplementedError("Subclasses must implement _save_model_files method")

    def _save_model_card(self, save_directory: Union[str, os.PathLike]):
        """README.md exactly as the reference writes it (ref core/helpers/export_mixin.py:27-39 ->
        modelcard_generator.py:6-20): huggingface_hub's default model-card template with the template card's fields
        as YAML front matter - `name`, `tags: [quantization]`, `description`, `metrics` (the hyperparameters),
        `intended_use`, `limitations`, `citations` - which is what the Hub indexes."""
        card = getattr(self, "template_card", None)
        if card is None:
            self.logger.warning("No template_card attribute found, skipping model card generation")
            return
        from huggingface_hub import ModelCard, ModelCardData
        data = ModelCardData(name=card.title, tags=["quantization"], description=card.description,
                             metrics=card.hyperparameters, intended_use=card.intended_use,
                             limitations=card.limitations, citations=card.citations)
        path = os.path.join(save_directory, "README.md")
        ModelCard.from_template(card_data=data, template_name=card.title).save(path)
        self.logger.info(f"Model card saved to {path}")

    def save_pretrained(self, save_directory: Union[str, os.PathLike]):
        os.makedirs(save_directory, exist_ok=True)
        self._save_model_files(save_directory)

    def save_model_card(self, save_directory: Union[str, os.PathLike]):
        os.makedirs(save_directory, exist_ok=True)
        self._save_model_card(save_directory)

    def push_to_hub(self, repo_id: Optional[str] = None, commit_message: Optional[str] = None,
                    private: Optional[bool] = None, token: Optional[str] = None,
                    create_pr: bool = False, safe_serialization: bool = False,
                    variant: Optional[str] = None):
        from huggingface_hub import create_repo, upload_folder
        if repo_id is None:
            repo_id = getattr(self, "repo_id", None) or getattr(self, "name", None)
            if repo_id is None:
                raise ValueError("repo_id must be specified if the model doesn't have a name attribute")
        token = token if token is not None else os.environ.get("HF_TOKEN", None)
        with tempfile.TemporaryDirectory() as tmpdir:
            self.save_pretrained(tmpdir)
            self.save_model_card(tmpdir)
            create_repo(repo_id=repo_id, token=token, private=private, exist_ok=True)
            return upload_folder(folder_path=tmpdir, path_in_repo=".", repo_id=repo_id,
                                 repo_type="model", token=token,
                                 commit_message=commit_message or f"Upload {self.__class__.__name__} model",
                                 create_pr=create_pr)


class CalibrationMixin:
    def require_calibration(self) -> bool:
        return False

    def prepare_calibration_data(self, dataset, tokenizer=None):
        return dataset

    def run_calibration(self):
        return None
