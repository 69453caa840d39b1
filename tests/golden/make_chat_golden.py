"""Writes tests/golden/chat_rows.json: outputs of the REFERENCE's own `convert_row`
(/root/reference/src/quantool/utils/dataset_textifier.py:178-260, loaded by file path because
`quantool.utils/__init__` pulls in packages this image lacks) on a fixed set of calibration rows, rendered with
the tiny test tokenizer and the chat template below.  Run in the build container only (the reference is not on
the GPU box):  python tests/golden/make_chat_golden.py"""
import importlib.util
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

TEMPLATE = ("{% for m in messages %}<|{{ m['role'] }}|>{{ m['content'] }}<|end|>{% endfor %}"
            "{% if add_generation_prompt %}<|assistant|>{% endif %}")
U, A, A2, S = ({"role": "user", "content": "hi there"}, {"role": "assistant", "content": "yo"},
               {"role": "assistant", "content": "no"}, {"role": "system", "content": "be brief"})
ROWS = [
    {"messages": [S, U, A]},
    {"messages": [U]},
    {"prompt": [U]},
    {"prompt": [U, A]},                                  # ends on an assistant turn: continued final message
    {"prompt": [S, U], "completion": [A]},
    {"prompt": [U], "completion": [A], "label": False},
    {"prompt": [U], "chosen": [A], "rejected": [A2]},
    {"chosen": [U, A], "rejected": [U, A2]},
    {"text": "not conversational"},
    {"prompt": "a plain string prompt"},
    {"prompt": [S]},                                      # last role neither user nor assistant: row unchanged
    {"messages": [U, A], "chat_template_kwargs": {"add_generation_prompt": False}},
]


def main():
    spec = importlib.util.spec_from_file_location("ref_textifier", "/root/reference/src/quantool/utils/dataset_textifier.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from transformers import AutoTokenizer
    from _tiny import write_tiny_tokenizer
    with tempfile.TemporaryDirectory() as d:
        write_tiny_tokenizer(d)
        tok = AutoTokenizer.from_pretrained(d)
    tok.chat_template = TEMPLATE
    out = {"template": TEMPLATE, "cases": []}
    for row in ROWS:
        out["cases"].append({"row": row, "rendered": ref.convert_row(dict(row), tok)})
    bad = {"messages": [U], "prompt": [U]}
    try:
        ref.convert_row(dict(bad), tok)
        raised = None
    except Exception as e:
        raised = type(e).__name__
    out["invalid"] = {"row": bad, "raises": raised}
    out["has_chat_template"] = {"with_template": ref.has_chat_template(tok)}
    tok.chat_template = None
    out["has_chat_template"]["without_template"] = ref.has_chat_template(tok)
    out["no_template_row"] = {"row": ROWS[0], "rendered": ref.convert_row(dict(ROWS[0]), tok)}
    with open(os.path.join(HERE, "chat_rows.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
