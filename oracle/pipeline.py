"""ORACLE — test infrastructure only.  CPU sequential-calibration driver built from the oracle
restatements (oracle/gptq.py, oracle/awq.py, oracle/smoothquant.py): what `llmcompressor.oneshot`
does layer by layer under the reference's call at ref/src/quantool/methods/llm_compressor/base.py:159-161
(SURVEY.md §3.1).  The decoder-layer forward is the oracle's plain-torch restatement
(oracle/llama_forward.py); only shape bookkeeping (LlamaShape, rope tables) is shared with the product.  Parity unpinned (see oracle/gptq.py).
"""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from quantool_b200.engine import llama

from . import awq as oawq
from . import llama_forward as lf
from . import gptq as og
from . import smoothquant as osq


def _layer(sd, l):
    pre = f"model.layers.{l}."
    return {k[len(pre):]: v.clone() for k, v in sd.items() if k.startswith(pre)}


def run_gptq(shape, sd: Dict[str, torch.Tensor], token_ids: torch.Tensor, oargs, percdamp=0.01,
             smooth_strength: Optional[float] = None):
    """Returns {linear key: (Wq, scale, zp, g_idx)} and the final hidden states."""
    h = F.embedding(token_ids, sd["model.embed_tokens.weight"])
    n, seq = token_ids.shape
    cos, sin = llama.rope_tables(shape, seq, h.device, h.dtype)
    dims = shape.input_dims()
    out = {}
    for l in range(shape.num_hidden_layers):
        w = _layer(sd, l)
        if smooth_strength is not None:
            cap = {k: torch.empty((n * seq, d), dtype=h.dtype) for k, d in dims.items()}
            lf.layer_forward(shape, w, h, cos, sin, capture=cap)
            for smooth, balance, inp in (("input_layernorm", ["self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"], "attn_in"),
                                         ("post_attention_layernorm", ["mlp.gate_proj", "mlp.up_proj"], "mlp_in")):
                mn, mx = osq.update_channel_minmax(cap[inp], None, None)
                bw = [w[f"{b}.weight"] for b in balance]
                s = osq.smoothing_scales(mn, mx, bw, smooth_strength)
                osq.apply_smoothing(w[f"{smooth}.weight"], bw, s)
                out[f"model.layers.{l}.{smooth}.smooth_scales"] = s
        cap = {k: torch.empty((n * seq, d), dtype=h.dtype) for k, d in dims.items()}
        lf.layer_forward(shape, w, h, cos, sin, capture=cap)
        H = {}
        for k in dims:
            Hk, cnt = og.make_empty_hessian(dims[k]), 0
            xs = cap[k].view(n, seq, -1)
            for b in range(n):
                Hk, cnt = og.accumulate_hessian(xs[b:b + 1], Hk, cnt)
            H[k] = Hk
        for lin in llama.LINEARS:
            loss, Wq, s, z, gi = og.quantize_weight(w[f"{lin}.weight"], H[llama.INPUT_OF[lin]], oargs, percdamp=percdamp)
            out[f"model.layers.{l}.{lin}"] = (Wq, s, z, gi, w[f"{lin}.weight"].clone(), cap[llama.INPUT_OF[lin]])
            w[f"{lin}.weight"] = Wq
        h = lf.layer_forward(shape, w, h, cos, sin)
    return out, h


def run_awq(shape, sd, token_ids, symmetric: bool, bits: int, group_size: int, n_grid: int = 20):
    h = F.embedding(token_ids, sd["model.embed_tokens.weight"])
    n, seq = token_ids.shape
    cos, sin = llama.rope_tables(shape, seq, h.device, h.dtype)
    dims = shape.input_dims()
    from quantool_b200.engine.awq import llama_mappings
    out = {}
    for l in range(shape.num_hidden_layers):
        w = _layer(sd, l)
        cap = {k: torch.empty((n * seq, d), dtype=h.dtype) for k, d in dims.items()}
        lf.layer_forward(shape, w, h, cos, sin, capture=cap)
        for mp in llama_mappings(shape):
            x_all = cap[mp.inp]
            ssum, cnt = oawq.accumulate_mean(x_all, None)
            x_mean = ssum / cnt
            bw = [w[f"{b}.weight"] for b in mp.balance]
            w_mean = oawq.weight_mean(bw, group_size)
            xin = x_all.view(n, seq, -1)

            def parent(patched, _mp=mp, _w=w, _xin=xin):
                ww = dict(_w)
                for name, t in zip(_mp.balance, patched):
                    ww[f"{name}.weight"] = t
                if _mp.parent == "self_attn":
                    return [lf.attention_forward(shape, ww, _xin, cos, sin)]
                if _mp.parent == "mlp":
                    return [lf.mlp_forward(ww, _xin)]
                return [F.linear(_xin, ww[f"{_mp.balance[0]}.weight"])]

            ref = parent(bw)
            s, ratio, hist = oawq.compute_best_scale(x_mean, w_mean, bw, parent, ref, symmetric, bits, group_size, n_grid)
            oawq.apply_scales(w[f"{mp.smooth}.weight"], bw, s)
            out[f"model.layers.{l}.{mp.smooth}"] = (s, ratio, hist, x_mean, w_mean)
        for lin in llama.LINEARS:
            out[f"model.layers.{l}.{lin}.weight"] = w[f"{lin}.weight"].clone()
        h = lf.layer_forward(shape, w, h, cos, sin)
    return out, h
