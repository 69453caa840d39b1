"""Chat-template rendering of conversational calibration rows.

The reference's `LLMCompressorQuantizer.prepare_calibration_data(dataset, tokenizer=...)`
(ref/src/quantool/methods/llm_compressor/base.py:257-345) maps every row through its `convert_row` helper
(ref/src/quantool/utils/dataset_textifier.py:178-260) before it makes sure a `text` column exists.  This module is
the engine-side counterpart with the same observable behaviour:

  row shape (lists of {"role", "content"} dicts)        rendered keys
  {"messages"}                                         {"text"}
  {"prompt"}                                           {"prompt"}
  {"prompt", "completion"[, "label"]}                  {"prompt", "completion"[, "label"]}
  {"prompt", "chosen", "rejected"}                     {"prompt", "chosen", "rejected"}
  {"chosen", "rejected"}                               {"chosen", "rejected"}

A prompt ending on a user turn is rendered with the generation prompt appended, one ending on an assistant turn
as a continued final message; a response is what the template adds after the text it shares with the rendered
prompt.  Rows that are not conversational, tokenizers without a chat template, and rows on which the template
raises come back unchanged; a conversational row with an unsupported key combination raises KeyError.
"""
from typing import Any, Dict, List, Optional

_ROLE_KEYS = ("prompt", "chosen", "rejected", "completion", "messages")
_KEY_SETS = (
    frozenset({"messages"}),
    frozenset({"prompt"}),
    frozenset({"prompt", "completion"}),
    frozenset({"prompt", "chosen", "rejected"}),
    frozenset({"chosen", "rejected"}),
    frozenset({"prompt", "completion", "label"}),
)


def has_chat_template(tokenizer: Any, verify: bool = False) -> bool:
    """True when `tokenizer` (or processor) can render conversations: it has `apply_chat_template` and a non-empty
    `chat_template` string - or, with verify=True, rendering a one-turn conversation does not raise."""
    if not hasattr(tokenizer, "apply_chat_template"):
        return False
    template = getattr(tokenizer, "chat_template", None)
    if isinstance(template, str) and template.strip():
        return True
    if not verify:
        return False
    try:
        tokenizer.apply_chat_template([{"role": "user", "content": "ping"}], tokenize=False, add_generation_prompt=False)
    except Exception:
        return False
    return True


def is_conversational(row: Dict[str, Any]) -> bool:
    """A row is conversational when one of its prompt / response / messages fields is a list of role-content dicts."""
    present = {k for k in row.keys() if k in _ROLE_KEYS}
    if not present:
        return False
    turns = row[present.pop()]
    if not isinstance(turns, list) or not turns:
        return False
    first = turns[0]
    return isinstance(first, dict) and "role" in first and "content" in first


def _shared_prefix_len(a: str, b: str) -> int:
    n = 0
    for x, y in zip(a, b):
        if x != y:
            break
        n += 1
    return n


def render_chat_row(row: Dict[str, Any], tokenizer: Any, tools: Optional[List[Any]] = None, **template_kwargs) -> Dict[str, Any]:
    """One dataset row -> rendered strings (table in the module docstring)."""
    if not is_conversational(row) or not has_chat_template(tokenizer):
        return row
    keys = frozenset(k for k in row.keys() if k in _ROLE_KEYS or k == "label")
    if keys not in _KEY_SETS:
        raise KeyError(f"Invalid keys in the example: {set(keys)}")
    kw = {**(row.get("chat_template_kwargs") or {}), **template_kwargs}

    def render(turns, **extra):
        return tokenizer.apply_chat_template(turns, tools=tools, tokenize=False, **extra, **kw)

    try:
        if "messages" in row:
            return {"text": render(row["messages"], add_generation_prompt=False)}
        out: Dict[str, Any] = {}
        if "prompt" in row:
            last = row["prompt"][-1].get("role")
            if last not in ("user", "assistant"):
                raise ValueError(f"Invalid role in the last message: {last}")
            prompt = render(row["prompt"], add_generation_prompt=(last == "user"),
                            continue_final_message=(last == "assistant"))
            for field in ("chosen", "rejected", "completion"):
                if field in row:
                    full = render(row["prompt"] + row[field])
                    cut = _shared_prefix_len(prompt, full)
                    prompt, out[field] = full[:cut], full[cut:]
            out["prompt"] = prompt
        else:
            for field in ("chosen", "rejected"):
                if field in row:
                    out[field] = render(row[field])
        if "label" in row:
            out["label"] = row["label"]
        return out
    except Exception:
        return row      # the template could not be applied: leave the row for the plain-text path
