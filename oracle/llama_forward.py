"""ORACLE — test infrastructure only.  Plain-torch Llama decoder-layer forward, op for op what
`transformers` runs under `llmcompressor.oneshot` (reference call site
ref/src/quantool/methods/llm_compressor/base.py:162): LlamaRMSNorm (fp32 statistics, weight multiply in
the model dtype), apply_rotary_pos_emb (`q*cos + rotate_half(q)*sin`), eager attention through SDPA,
LlamaMLP (`down(silu(gate(x)) * up(x))`).  The product path (quantool_b200/engine/llama.py) runs the same
expression with fused CUDA kernels; tests compare the two.  PINNED against the installed transformers
(5.5; the reference pins 4.56.2): tests/test_oracle_cpu.py::test_forward_restatement_equals_transformers_llama -
in fp32 the hidden states after every layer and the four captured Linear inputs are bit-exact with a
LlamaForCausalLM built from the same config and weights (plain and llama3-scaled rope); in bf16 they differ only
behind the library attention kernel, by rounding of its accumulation.
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (w * xf.to(x.dtype))


def _rot_half(x):
    x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def layer_forward(shape, w: Dict[str, torch.Tensor], h: torch.Tensor, cos, sin,
                  capture: Optional[Dict[str, torch.Tensor]] = None, row0: int = 0) -> torch.Tensor:
    """h: [B, S, hidden].  If `capture` is given, the inputs of the Linears are written into
    capture[name][row0 : row0 + B*S] (preallocated [T, K] buffers)."""
    B, S, _ = h.shape
    nh, nkv, hd = shape.num_attention_heads, shape.num_key_value_heads, shape.head_dim

    def cap(name, t):
        if capture is not None:
            capture[name][row0: row0 + B * S].copy_(t.reshape(B * S, -1))

    x = rms_norm(h, w["input_layernorm.weight"], shape.rms_norm_eps)
    cap("attn_in", x)
    q = F.linear(x, w["self_attn.q_proj.weight"]).view(B, S, nh, hd).transpose(1, 2)
    k = F.linear(x, w["self_attn.k_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
    v = F.linear(x, w["self_attn.v_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
    c, s = cos[None, None], sin[None, None]
    q = q * c + _rot_half(q) * s
    k = k * c + _rot_half(k) * s
    a = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=(nkv != nh))
    a = a.transpose(1, 2).reshape(B, S, nh * hd)
    cap("o_in", a)
    h = h + F.linear(a, w["self_attn.o_proj.weight"])
    x = rms_norm(h, w["post_attention_layernorm.weight"], shape.rms_norm_eps)
    cap("mlp_in", x)
    d = F.silu(F.linear(x, w["mlp.gate_proj.weight"])) * F.linear(x, w["mlp.up_proj.weight"])
    cap("down_in", d)
    return h + F.linear(d, w["mlp.down_proj.weight"])


def attention_forward(shape, w: Dict[str, torch.Tensor], x: torch.Tensor, cos, sin) -> torch.Tensor:
    """self_attn(x) on normed input x [B,S,hidden] -> [B,S,hidden] (AWQ parent module of q/k/v)."""
    B, S, _ = x.shape
    nh, nkv, hd = shape.num_attention_heads, shape.num_key_value_heads, shape.head_dim
    q = F.linear(x, w["self_attn.q_proj.weight"]).view(B, S, nh, hd).transpose(1, 2)
    k = F.linear(x, w["self_attn.k_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
    v = F.linear(x, w["self_attn.v_proj.weight"]).view(B, S, nkv, hd).transpose(1, 2)
    c, s = cos[None, None], sin[None, None]
    q = q * c + _rot_half(q) * s
    k = k * c + _rot_half(k) * s
    a = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=(nkv != nh))
    a = a.transpose(1, 2).reshape(B, S, nh * hd)
    return F.linear(a, w["self_attn.o_proj.weight"])


def mlp_forward(w: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    return F.linear(F.silu(F.linear(x, w["mlp.gate_proj.weight"])) * F.linear(x, w["mlp.up_proj.weight"]),
                    w["mlp.down_proj.weight"])
