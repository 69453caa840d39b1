"""Test helper: a tiny byte-level BPE tokenizer (tokenizer.json + config) written next to a synthetic checkpoint, so
the GGUF converter has the tokenizer files a real model directory carries."""


def write_tiny_tokenizer(model_dir: str) -> int:
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    alphabet = sorted(pre_tokenizers.ByteLevel.alphabet())
    vocab = {ch: i for i, ch in enumerate(alphabet)}
    merges = [("Ġ", "t"), ("h", "e"), ("Ġt", "he"), ("i", "n")]
    for a, b in merges:
        vocab[a + b] = len(vocab)
    tok = Tokenizer(models.BPE(vocab=vocab, merges=merges))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, bos_token="<|begin_of_text|>", eos_token="<|end_of_text|>")
    fast.save_pretrained(model_dir)
    return len(fast)
