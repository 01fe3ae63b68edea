"""TEST INFRASTRUCTURE ONLY — CPU restatement of the two text encoders the reference's Flux embed path calls when prompt
embeddings are not pre-cached:

    prompt_embeds = self.text_encoder_2(text_input_ids, output_hidden_states=False)[0]      # T5EncoderModel (T5 v1.1 XXL)
    pooled        = self.text_encoder(text_input_ids, output_hidden_states=False).pooler_output   # CLIPTextModel (CLIP-L)
    (simpletuner/helpers/models/flux/pipeline.py:1085, 1127-1130; driven by Flux._encode_prompts, flux/model.py:497-520, and
    helpers/caching/text_embeds.py)

The modules themselves are `transformers` classes (T5EncoderModel / CLIPTextModel), a dependency that IS importable in this
container: oracle/make_golden_text.py runs tiny random-weight instances of the real classes and commits their inputs,
weights' seed and outputs as tests/golden/text_golden.pt, which pins this restatement (tests/test_text_oracle.py).
Parameter names are the transformers state-dict names.  Only tests/ may import this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------ T5 encoder
@dataclass
class T5Config:
    vocab_size: int = 32128
    d_model: int = 4096
    d_kv: int = 64
    d_ff: int = 10240
    num_layers: int = 24
    num_heads: int = 64
    relative_attention_num_buckets: int = 32
    relative_attention_max_distance: int = 128
    layer_norm_epsilon: float = 1e-6


def t5_param_shapes(c: T5Config) -> Dict[str, tuple]:
    inner = c.num_heads * c.d_kv
    s = {"shared.weight": (c.vocab_size, c.d_model), "encoder.final_layer_norm.weight": (c.d_model,),
         "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight": (c.relative_attention_num_buckets, c.num_heads)}
    for i in range(c.num_layers):
        p = f"encoder.block.{i}.layer."
        for n in ("q", "k", "v"):
            s[p + f"0.SelfAttention.{n}.weight"] = (inner, c.d_model)
        s[p + "0.SelfAttention.o.weight"] = (c.d_model, inner)
        s[p + "0.layer_norm.weight"] = (c.d_model,)
        s[p + "1.DenseReluDense.wi_0.weight"] = (c.d_ff, c.d_model)
        s[p + "1.DenseReluDense.wi_1.weight"] = (c.d_ff, c.d_model)
        s[p + "1.DenseReluDense.wo.weight"] = (c.d_model, c.d_ff)
        s[p + "1.layer_norm.weight"] = (c.d_model,)
    return s


def init_params(shapes: Dict[str, tuple], seed: int = 0, std: float = 0.05) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        if k.endswith("layer_norm.weight") or k.endswith("layer_norm1.weight") or k.endswith("layer_norm2.weight"):
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            out[k] = 0.02 * torch.randn(shp, generator=g)
        else:
            out[k] = std * torch.randn(shp, generator=g)
    return out


def t5_relative_position_bucket(relative_position: Tensor, num_buckets: int = 32, max_distance: int = 128) -> Tensor:
    """T5Attention._relative_position_bucket, bidirectional (encoder)."""
    num_buckets //= 2
    buckets = (relative_position > 0).to(torch.long) * num_buckets
    rp = torch.abs(relative_position)
    max_exact = num_buckets // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rp, large)


def t5_position_bias(P: Dict[str, Tensor], c: T5Config, S: int) -> Tensor:
    """T5Attention.compute_bias: [H, S, S] (query row, key column); shared by every layer."""
    ctx = torch.arange(S)[:, None]
    mem = torch.arange(S)[None, :]
    bucket = t5_relative_position_bucket(mem - ctx, c.relative_attention_num_buckets, c.relative_attention_max_distance)
    w = P["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"]
    return w[bucket].permute(2, 0, 1)


def t5_layer_norm(x: Tensor, w: Tensor, eps: float) -> Tensor:
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    if w.dtype in (torch.float16, torch.bfloat16):
        x = x.to(w.dtype)
    return w * x


def t5_encoder(P: Dict[str, Tensor], c: T5Config, input_ids: Tensor) -> Tensor:
    """T5EncoderModel(input_ids)[0] in eval mode, no attention mask (pipeline.py:1085 passes none)."""
    B, S = input_ids.shape
    H, hd = c.num_heads, c.d_kv
    h = P["shared.weight"][input_ids]
    bias = t5_position_bias(P, c, S).to(h.dtype)
    for i in range(c.num_layers):
        p = f"encoder.block.{i}.layer."
        n = t5_layer_norm(h, P[p + "0.layer_norm.weight"], c.layer_norm_epsilon)
        q = F.linear(n, P[p + "0.SelfAttention.q.weight"]).view(B, S, H, hd).transpose(1, 2)
        k = F.linear(n, P[p + "0.SelfAttention.k.weight"]).view(B, S, H, hd).transpose(1, 2)
        v = F.linear(n, P[p + "0.SelfAttention.v.weight"]).view(B, S, H, hd).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(3, 2)) + bias[None]          # T5 does not scale the scores
        w = F.softmax(scores.float(), dim=-1).type_as(scores)
        o = torch.matmul(w, v).transpose(1, 2).reshape(B, S, H * hd)
        h = h + F.linear(o, P[p + "0.SelfAttention.o.weight"])
        n = t5_layer_norm(h, P[p + "1.layer_norm.weight"], c.layer_norm_epsilon)
        g = F.gelu(F.linear(n, P[p + "1.DenseReluDense.wi_0.weight"]), approximate="tanh")   # "gelu_new"
        u = g * F.linear(n, P[p + "1.DenseReluDense.wi_1.weight"])
        h = h + F.linear(u, P[p + "1.DenseReluDense.wo.weight"])
    return t5_layer_norm(h, P["encoder.final_layer_norm.weight"], c.layer_norm_epsilon)


# ------------------------------------------------------------------------------------------------ CLIP text model
@dataclass
class CLIPTextConfig:
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    layer_norm_eps: float = 1e-5
    eos_token_id: int = 2          # the original CLIP-L config value: pooled = hidden state at argmax(input_ids)
    hidden_act: str = "quick_gelu"


def clip_param_shapes(c: CLIPTextConfig) -> Dict[str, tuple]:
    D, I = c.hidden_size, c.intermediate_size
    s = {"text_model.embeddings.token_embedding.weight": (c.vocab_size, D),
         "text_model.embeddings.position_embedding.weight": (c.max_position_embeddings, D),
         "text_model.final_layer_norm.weight": (D,), "text_model.final_layer_norm.bias": (D,)}
    for i in range(c.num_hidden_layers):
        p = f"text_model.encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"] = (D, D)
            s[p + f"self_attn.{n}.bias"] = (D,)
        for n in ("layer_norm1", "layer_norm2"):
            s[p + n + ".weight"] = (D,)
            s[p + n + ".bias"] = (D,)
        s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (I, D), (I,)
        s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (D, I), (D,)
    return s


def clip_text_model(P: Dict[str, Tensor], c: CLIPTextConfig, input_ids: Tensor):
    """CLIPTextModel(input_ids) in eval mode -> (last_hidden_state, pooler_output); causal mask, no padding mask."""
    B, S = input_ids.shape
    H = c.num_attention_heads
    hd = c.hidden_size // H
    h = P["text_model.embeddings.token_embedding.weight"][input_ids] + P["text_model.embeddings.position_embedding.weight"][:S][None]
    causal = torch.full((S, S), float("-inf")).triu(1).to(h.dtype)
    act = (lambda x: x * torch.sigmoid(1.702 * x)) if c.hidden_act == "quick_gelu" else (lambda x: F.gelu(x))
    for i in range(c.num_hidden_layers):
        p = f"text_model.encoder.layers.{i}."
        n = F.layer_norm(h, (c.hidden_size,), P[p + "layer_norm1.weight"], P[p + "layer_norm1.bias"], c.layer_norm_eps)
        q = F.linear(n, P[p + "self_attn.q_proj.weight"], P[p + "self_attn.q_proj.bias"]).view(B, S, H, hd).transpose(1, 2)
        k = F.linear(n, P[p + "self_attn.k_proj.weight"], P[p + "self_attn.k_proj.bias"]).view(B, S, H, hd).transpose(1, 2)
        v = F.linear(n, P[p + "self_attn.v_proj.weight"], P[p + "self_attn.v_proj.bias"]).view(B, S, H, hd).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(3, 2)) * (hd ** -0.5) + causal
        w = F.softmax(scores, dim=-1, dtype=torch.float32).to(q.dtype)
        o = torch.matmul(w, v).transpose(1, 2).reshape(B, S, c.hidden_size)
        h = h + F.linear(o, P[p + "self_attn.out_proj.weight"], P[p + "self_attn.out_proj.bias"])
        n = F.layer_norm(h, (c.hidden_size,), P[p + "layer_norm2.weight"], P[p + "layer_norm2.bias"], c.layer_norm_eps)
        h = h + F.linear(act(F.linear(n, P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"])), P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"])
    last = F.layer_norm(h, (c.hidden_size,), P["text_model.final_layer_norm.weight"], P["text_model.final_layer_norm.bias"], c.layer_norm_eps)
    if c.eos_token_id == 2:
        idx = input_ids.to(torch.int).argmax(dim=-1)
    else:
        idx = (input_ids.to(torch.int) == c.eos_token_id).int().argmax(dim=-1)
    return last, last[torch.arange(B), idx]
