"""Llama-family decoder plumbing for the sequential calibration driver.

What upstream's SequentialPipeline gets from `transformers` (SURVEY.md §3.1 "HOT LOOP 1/2"):
run one decoder layer at a time over the calibration batches, exposing the inputs of every
Linear.  This is plumbing around the quantization kernels - library GEMMs (cuBLAS) and SDPA, with the
normalisation / rotary / gated-activation steps between them as one-pass CUDA kernels
(csrc/forward.cu) - written functionally so a layer's weights can be swapped for their quantized
versions between the two passes.  CUDA only: the CPU restatement lives in oracle/llama_forward.py.

Weights follow the HF checkpoint naming (`model.layers.{i}.self_attn.q_proj.weight`, ...).
"""
import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .. import cabi

LINEARS = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "self_attn.o_proj",
           "mlp.gate_proj", "mlp.up_proj", "mlp.down_proj")
# which captured activation feeds which Linear (4 distinct inputs per layer)
INPUT_OF = {"self_attn.q_proj": "attn_in", "self_attn.k_proj": "attn_in", "self_attn.v_proj": "attn_in",
            "self_attn.o_proj": "o_in", "mlp.gate_proj": "mlp_in", "mlp.up_proj": "mlp_in",
            "mlp.down_proj": "down_in"}


@dataclass
class LlamaShape:
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    vocab_size: int
    head_dim: Optional[int] = None
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    tie_word_embeddings: bool = False
    max_position_embeddings: int = 8192
    rope_scaling: Optional[dict] = None      # None or {"rope_type": "llama3", factor, low_freq_factor, ...}

    def __post_init__(self):
        if self.head_dim is None:
            self.head_dim = self.hidden_size // self.num_attention_heads

    @property
    def kv_dim(self):
        return self.num_key_value_heads * self.head_dim

    @property
    def q_dim(self):
        return self.num_attention_heads * self.head_dim

    def linear_shapes(self) -> Dict[str, tuple]:
        h, i = self.hidden_size, self.intermediate_size
        return {"self_attn.q_proj": (self.q_dim, h), "self_attn.k_proj": (self.kv_dim, h),
                "self_attn.v_proj": (self.kv_dim, h), "self_attn.o_proj": (h, self.q_dim),
                "mlp.gate_proj": (i, h), "mlp.up_proj": (i, h), "mlp.down_proj": (h, i)}

    def input_dims(self) -> Dict[str, int]:
        return {"attn_in": self.hidden_size, "o_in": self.q_dim, "mlp_in": self.hidden_size,
                "down_in": self.intermediate_size}

    def to_hf_config(self) -> dict:
        return {"architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": self.hidden_size,
                "intermediate_size": self.intermediate_size, "num_hidden_layers": self.num_hidden_layers,
                "num_attention_heads": self.num_attention_heads, "num_key_value_heads": self.num_key_value_heads,
                "head_dim": self.head_dim, "vocab_size": self.vocab_size, "rms_norm_eps": self.rms_norm_eps,
                "rope_theta": self.rope_theta, "tie_word_embeddings": self.tie_word_embeddings,
                "max_position_embeddings": self.max_position_embeddings, "hidden_act": "silu",
                "torch_dtype": "bfloat16", "attention_bias": False, "mlp_bias": False,
                **({"rope_scaling": dict(self.rope_scaling)} if self.rope_scaling else {})}

    @staticmethod
    def from_hf_config(cfg: dict) -> "LlamaShape":
        """Raises on anything the calibration forward below would compute wrongly (the reference runs the real
        HF module, so a silently different forward would change every Hessian)."""
        mt = cfg.get("model_type")
        archs = cfg.get("architectures") or []
        if mt != "llama" or any(a != "LlamaForCausalLM" for a in archs):
            raise ValueError(f"the sm_100a calibration driver implements the Llama decoder layer only "
                             f"(model_type={mt!r}, architectures={archs})")
        if cfg.get("attention_bias") or cfg.get("mlp_bias"):
            raise ValueError("Llama variants with attention_bias / mlp_bias are not supported (bias tensors would be "
                             "dropped from the forward and from the artifact)")
        if cfg.get("sliding_window") or cfg.get("use_sliding_window"):
            raise ValueError("sliding-window attention is not supported by the calibration forward")
        if cfg.get("hidden_act", "silu") != "silu":
            raise ValueError(f"hidden_act={cfg.get('hidden_act')!r} is not supported (silu only)")
        rope = cfg.get("rope_theta")
        rp = cfg.get("rope_parameters") if isinstance(cfg.get("rope_parameters"), dict) else None
        if rope is None and rp:
            rope = rp.get("rope_theta")
        scaling = cfg.get("rope_scaling")
        if scaling is None and rp and rp.get("rope_type", "default") != "default":
            scaling = rp
        if scaling is not None:
            kind = scaling.get("rope_type", scaling.get("type", "default"))
            if kind == "default":
                scaling = None
            elif kind != "llama3":
                raise ValueError(f"rope_scaling type {kind!r} is not implemented (default and llama3 are)")
            else:
                scaling = {"rope_type": "llama3", "factor": float(scaling["factor"]),
                           "low_freq_factor": float(scaling.get("low_freq_factor", 1.0)),
                           "high_freq_factor": float(scaling.get("high_freq_factor", 4.0)),
                           "original_max_position_embeddings": int(scaling["original_max_position_embeddings"])}
        return LlamaShape(hidden_size=cfg["hidden_size"], intermediate_size=cfg["intermediate_size"],
                          num_hidden_layers=cfg["num_hidden_layers"],
                          num_attention_heads=cfg["num_attention_heads"],
                          num_key_value_heads=cfg.get("num_key_value_heads", cfg["num_attention_heads"]),
                          vocab_size=cfg["vocab_size"], head_dim=cfg.get("head_dim"),
                          rms_norm_eps=cfg.get("rms_norm_eps", 1e-5), rope_theta=rope or 10000.0,
                          tie_word_embeddings=cfg.get("tie_word_embeddings", False),
                          max_position_embeddings=cfg.get("max_position_embeddings", 8192), rope_scaling=scaling)


# the shapes BASELINE.json names (SURVEY.md §8 header)
SHAPES = {
    "smollm2-135m": LlamaShape(576, 1536, 30, 9, 3, 49152, rope_theta=100000.0, tie_word_embeddings=True),
    "llama-3.2-1b": LlamaShape(2048, 8192, 16, 32, 8, 128256, head_dim=64, tie_word_embeddings=True),
    "llama-3-8b": LlamaShape(4096, 14336, 32, 32, 8, 128256),
    "llama-3-70b": LlamaShape(8192, 28672, 80, 64, 8, 128256),
}


def random_layer_weights(shape: LlamaShape, layer: int, device, dtype=torch.bfloat16, seed: int = 0,
                         pin: bool = False) -> Dict[str, torch.Tensor]:
    """Random-init weights of one decoder layer (normal(0, 0.02) like HF `initializer_range`)."""
    g = torch.Generator(device=device).manual_seed(seed * 100003 + layer)
    out = {}
    for name, (n, k) in shape.linear_shapes().items():
        w = (torch.randn((n, k), generator=g, device=device, dtype=torch.float32) * 0.02).to(dtype)
        out[f"{name}.weight"] = w
    out["input_layernorm.weight"] = torch.ones(shape.hidden_size, device=device, dtype=dtype)
    out["post_attention_layernorm.weight"] = torch.ones(shape.hidden_size, device=device, dtype=dtype)
    if pin and str(device) == "cpu":
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def random_state_dict(shape: LlamaShape, dtype=torch.bfloat16, seed: int = 0, device="cpu") -> Dict[str, torch.Tensor]:
    """Whole-model random init in HF key names (host tensors by default)."""
    sd = {}
    g = torch.Generator(device=device).manual_seed(seed * 100003 + 99991)
    sd["model.embed_tokens.weight"] = (torch.randn((shape.vocab_size, shape.hidden_size), generator=g,
                                                   device=device) * 0.02).to(dtype)
    for l in range(shape.num_hidden_layers):
        for k, v in random_layer_weights(shape, l, device, dtype, seed).items():
            sd[f"model.layers.{l}.{k}"] = v
    sd["model.norm.weight"] = torch.ones(shape.hidden_size, device=device, dtype=dtype)
    if not shape.tie_word_embeddings:
        sd["lm_head.weight"] = (torch.randn((shape.vocab_size, shape.hidden_size), generator=g,
                                            device=device) * 0.02).to(dtype)
    return sd


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LlamaRMSNorm (fp32 statistics, weight multiply in the model dtype), one fused pass."""
    return cabi.rms_norm(x, w, eps, out=out)


def rope_tables(shape: LlamaShape, seq: int, device, dtype):
    inv = 1.0 / (shape.rope_theta ** (torch.arange(0, shape.head_dim, 2, device=device, dtype=torch.float32)
                                     / shape.head_dim))
    if shape.rope_scaling:
        # transformers modeling_rope_utils._compute_llama3_parameters: long wavelengths are divided by `factor`,
        # short ones kept, the band in between interpolated - this changes low-frequency dims at EVERY position
        sc = shape.rope_scaling
        old = sc["original_max_position_embeddings"]
        lo_w, hi_w = old / sc["low_freq_factor"], old / sc["high_freq_factor"]
        wavelen = 2 * math.pi / inv
        inv_l = torch.where(wavelen > lo_w, inv / sc["factor"], inv)
        smooth = (old / wavelen - sc["low_freq_factor"]) / (sc["high_freq_factor"] - sc["low_freq_factor"])
        smoothed = (1 - smooth) * inv_l / sc["factor"] + smooth * inv_l
        mid = ~(wavelen < hi_w) & ~(wavelen > lo_w)
        inv = torch.where(mid, smoothed, inv_l)
    t = torch.arange(seq, device=device, dtype=torch.float32)
    f = torch.outer(t, inv)
    emb = torch.cat((f, f), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def _qkv_rope(shape: LlamaShape, w: Dict[str, torch.Tensor], x: torch.Tensor, cos, sin):
    """q/k/v projections + in-place rotary embedding; returns SDPA-layout views [B, heads, S, hd]."""
    B, S, _ = x.shape
    nh, nkv, hd = shape.num_attention_heads, shape.num_key_value_heads, shape.head_dim
    q = F.linear(x, w["self_attn.q_proj.weight"])
    k = F.linear(x, w["self_attn.k_proj.weight"])
    v = F.linear(x, w["self_attn.v_proj.weight"])
    cabi.rope_(q, cos, sin, S, nh, hd)
    cabi.rope_(k, cos, sin, S, nkv, hd)
    return (q.view(B, S, nh, hd).transpose(1, 2), k.view(B, S, nkv, hd).transpose(1, 2),
            v.view(B, S, nkv, hd).transpose(1, 2))


def _attend(shape: LlamaShape, q, k, v, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Causal SDPA; the head-major result is copied once into token-major [B, S, heads*hd] (`out`)."""
    B, nh, S, hd = q.shape
    a = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=(k.shape[1] != nh))
    if out is None:
        out = torch.empty((B, S, nh * hd), dtype=a.dtype, device=a.device)
    out.view(B, S, nh, hd).copy_(a.transpose(1, 2))
    return out


def layer_forward(shape: LlamaShape, w: Dict[str, torch.Tensor], h: torch.Tensor, cos, sin,
                  capture: Optional[Dict[str, torch.Tensor]] = None, row0: int = 0,
                  stop_after: Optional[str] = None) -> Optional[torch.Tensor]:
    """h: [B, S, hidden] contiguous.  If `capture` is given, the inputs of the Linears are produced
    directly in capture[name][row0 : row0 + B*S] (preallocated [T, K] buffers).  `stop_after` names a
    captured input ("mlp_in", "down_in") after which the rest of the layer is skipped (statistics
    passes do not need the layer output); the function then returns None."""
    B, S, _ = h.shape

    def slot(name):
        if capture is None:
            return None
        buf = capture[name]
        return buf[row0: row0 + B * S].view(B, S, buf.shape[1])

    x = rms_norm(h, w["input_layernorm.weight"], shape.rms_norm_eps, out=slot("attn_in"))
    q, k, v = _qkv_rope(shape, w, x, cos, sin)
    a = _attend(shape, q, k, v, out=slot("o_in"))
    del q, k, v
    h = h + F.linear(a, w["self_attn.o_proj.weight"])
    x = rms_norm(h, w["post_attention_layernorm.weight"], shape.rms_norm_eps, out=slot("mlp_in"))
    if stop_after == "mlp_in":
        return None
    g = F.linear(x, w["mlp.gate_proj.weight"])
    u = F.linear(x, w["mlp.up_proj.weight"])
    d = cabi.silu_mul(g, u, out=slot("down_in"))
    del g, u
    if stop_after == "down_in":
        return None
    return h + F.linear(d, w["mlp.down_proj.weight"])


def layer_forward_pre_down(shape: LlamaShape, w: Dict[str, torch.Tensor], h: torch.Tensor, cos, sin,
                           d_out: torch.Tensor) -> torch.Tensor:
    """The layer up to (not including) down_proj: returns the post-attention residual [B, S, hidden] and writes
    silu(gate) * up into d_out [B*S, intermediate].  Same ops, same rounding points as `layer_forward`."""
    B, S, _ = h.shape
    x = rms_norm(h, w["input_layernorm.weight"], shape.rms_norm_eps)
    q, k, v = _qkv_rope(shape, w, x, cos, sin)
    a = _attend(shape, q, k, v)
    del q, k, v
    h = h + F.linear(a, w["self_attn.o_proj.weight"])
    x = rms_norm(h, w["post_attention_layernorm.weight"], shape.rms_norm_eps)
    g = F.linear(x, w["mlp.gate_proj.weight"])
    u = F.linear(x, w["mlp.up_proj.weight"])
    cabi.silu_mul(g, u, out=d_out.view(B, S, d_out.shape[1]))
    return h


def layer_forward_down(w: Dict[str, torch.Tensor], h: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """h [T, hidden] (post-attention residual) + down_proj(d), d [T, intermediate]: the product is rounded to the
    model dtype before the residual add, as in `layer_forward`."""
    return h + F.linear(d, w["mlp.down_proj.weight"])


def attention_forward(shape: LlamaShape, w: Dict[str, torch.Tensor], x: torch.Tensor, cos, sin) -> torch.Tensor:
    """self_attn(x) on normed input x [B,S,hidden] -> [B,S,hidden] (AWQ parent module of q/k/v)."""
    q, k, v = _qkv_rope(shape, w, x, cos, sin)
    return F.linear(_attend(shape, q, k, v), w["self_attn.o_proj.weight"])


def mlp_forward(w: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    g = F.linear(x, w["mlp.gate_proj.weight"])
    u = F.linear(x, w["mlp.up_proj.weight"])
    return F.linear(cabi.silu_mul(g, u), w["mlp.down_proj.weight"])
