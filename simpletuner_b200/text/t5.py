"""`T5EncoderModel` on libstb200 — mirror of transformers' T5EncoderModel as the reference calls it
(`self.text_encoder_2(text_input_ids, output_hidden_states=False)[0]`, simpletuner/helpers/models/flux/pipeline.py:1085):
same constructor config fields, the transformers state-dict names (`shared.weight`, `encoder.block.N.layer.0.SelfAttention.q.weight`
...), eval-mode forward without an attention mask.  T5 v1.1 layout only (gated tanh-GELU feed-forward, no biases).

Per block: T5LayerNorm (`stb_rmsnorm_fwd`) -> fused q|k|v GEMM -> attention with the shared relative-position bias tile, scale 1
(`stb_attn_fwd` BIAS instantiation) -> o-projection GEMM with the residual add in its epilogue -> T5LayerNorm -> wi_0 GEMM with
tanh-GELU epilogue -> wi_1 GEMM with the gated multiply epilogue -> wo GEMM with the residual add.  No CPU fallback.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..flux.transformer import Linear, RMSNormWeight, _Weight


class ModelOutput(tuple):
    """Tuple that also answers attribute access, like transformers' ModelOutput (`out[0]`, `out.last_hidden_state`)."""

    def __new__(cls, **fields):
        self = super().__new__(cls, tuple(fields.values()))
        self._fields = dict(fields)
        return self

    def __getattr__(self, name):
        try:
            return self.__dict__["_fields"][name]
        except KeyError as exc:
            raise AttributeError(name) from exc


class _SelfAttention(nn.Module):
    def __init__(self, d_model, inner, heads, buckets, has_bias, dtype):
        super().__init__()
        self.q = Linear(d_model, inner, bias=False, dtype=dtype)
        self.k = Linear(d_model, inner, bias=False, dtype=dtype)
        self.v = Linear(d_model, inner, bias=False, dtype=dtype)
        self.o = Linear(inner, d_model, bias=False, dtype=dtype)
        if has_bias:
            self.relative_attention_bias = _Weight((buckets, heads), dtype)
            self.relative_attention_bias.weight.requires_grad_(False)


class _AttnLayer(nn.Module):
    def __init__(self, d_model, inner, heads, buckets, has_bias, dtype):
        super().__init__()
        self.SelfAttention = _SelfAttention(d_model, inner, heads, buckets, has_bias, dtype)
        self.layer_norm = RMSNormWeight(d_model, dtype)


class _DenseGated(nn.Module):
    def __init__(self, d_model, d_ff, dtype):
        super().__init__()
        self.wi_0 = Linear(d_model, d_ff, bias=False, dtype=dtype)
        self.wi_1 = Linear(d_model, d_ff, bias=False, dtype=dtype)
        self.wo = Linear(d_ff, d_model, bias=False, dtype=dtype)


class _FFLayer(nn.Module):
    def __init__(self, d_model, d_ff, dtype):
        super().__init__()
        self.DenseReluDense = _DenseGated(d_model, d_ff, dtype)
        self.layer_norm = RMSNormWeight(d_model, dtype)


class _Block(nn.Module):
    def __init__(self, d_model, inner, heads, d_ff, buckets, has_bias, dtype):
        super().__init__()
        self.layer = nn.ModuleList([_AttnLayer(d_model, inner, heads, buckets, has_bias, dtype), _FFLayer(d_model, d_ff, dtype)])


class _Stack(nn.Module):
    def __init__(self, n, d_model, inner, heads, d_ff, buckets, dtype):
        super().__init__()
        self.block = nn.ModuleList([_Block(d_model, inner, heads, d_ff, buckets, i == 0, dtype) for i in range(n)])
        self.final_layer_norm = RMSNormWeight(d_model, dtype)


def relative_position_bucket(relative_position: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """T5Attention._relative_position_bucket (bidirectional): integer math on the [S, S] grid of key - query offsets."""
    num_buckets //= 2
    buckets = (relative_position > 0).to(torch.long) * num_buckets
    rp = torch.abs(relative_position)
    max_exact = num_buckets // 2
    is_small = rp < max_exact
    large = max_exact + (torch.log(rp.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)).to(torch.long)
    large = torch.min(large, torch.full_like(large, num_buckets - 1))
    return buckets + torch.where(is_small, rp, large)


class T5EncoderModel(nn.Module):
    def __init__(self, vocab_size: int = 32128, d_model: int = 4096, d_kv: int = 64, d_ff: int = 10240, num_layers: int = 24,
                 num_heads: int = 64, relative_attention_num_buckets: int = 32, relative_attention_max_distance: int = 128,
                 layer_norm_epsilon: float = 1e-6, feed_forward_proj: str = "gated-gelu", dtype=torch.bfloat16, **unused):
        super().__init__()
        if feed_forward_proj != "gated-gelu":
            raise NotImplementedError("libstb200 T5 encoder implements the T5 v1.1 gated-GELU feed-forward only")
        if d_kv not in (64, 128):
            raise NotImplementedError("libstb200 attention supports head_dim 64 / 128")
        self.config = SimpleNamespace(vocab_size=vocab_size, d_model=d_model, d_kv=d_kv, d_ff=d_ff, num_layers=num_layers,
                                      num_heads=num_heads, relative_attention_num_buckets=relative_attention_num_buckets,
                                      relative_attention_max_distance=relative_attention_max_distance,
                                      layer_norm_epsilon=layer_norm_epsilon, feed_forward_proj=feed_forward_proj)
        self.shared = _Weight((vocab_size, d_model), dtype)
        self.shared.weight.requires_grad_(False)
        self.encoder = _Stack(num_layers, d_model, num_heads * d_kv, num_heads, d_ff, relative_attention_num_buckets, dtype)
        self._plans: Optional[list] = None
        self._bias_cache: Dict[int, torch.Tensor] = {}
        self.eval()

    @property
    def dtype(self):
        return self.shared.weight.dtype

    # transformers ties `encoder.embed_tokens.weight` to `shared.weight`; checkpoints may carry either or both
    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = dict(state_dict)
        tied = sd.pop("encoder.embed_tokens.weight", None)
        if "shared.weight" not in sd and tied is not None:
            sd["shared.weight"] = tied
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._plans, self._bias_cache = None, {}
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._plans, self._bias_cache = None, {}
        return out

    def _position_bias(self, S: int, device) -> torch.Tensor:
        """T5Attention.compute_bias of block 0, shared by every block: bf16 [H, S, S]."""
        b = self._bias_cache.get(S)
        if b is None:
            c = self.config
            ctx = torch.arange(S, device=device)[:, None]
            mem = torch.arange(S, device=device)[None, :]
            bucket = relative_position_bucket(mem - ctx, c.relative_attention_num_buckets, c.relative_attention_max_distance)
            w = self.encoder.block[0].layer[0].SelfAttention.relative_attention_bias.weight
            b = w[bucket].permute(2, 0, 1).contiguous()
            self._bias_cache[S] = b
        return b

    def _build_plans(self):
        plans = []
        for blk in self.encoder.block:
            a = blk.layer[0].SelfAttention
            plans.append(torch.cat([a.q.weight.detach(), a.k.weight.detach(), a.v.weight.detach()], 0).contiguous())
        return plans

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask=None, output_hidden_states: bool = False, return_dict: bool = True,
                **unused):
        if attention_mask is not None:
            raise NotImplementedError("the reference calls the T5 encoder without an attention mask (flux/pipeline.py:1085)")
        if not input_ids.is_cuda:
            from .._lib import StbError
            raise StbError("T5EncoderModel (libstb200) needs CUDA tensors; there is no CPU fallback")
        c = self.config
        B, S = input_ids.shape
        H, hd, eps = c.num_heads, c.d_kv, c.layer_norm_epsilon
        inner = H * hd
        if self._plans is None:
            self._plans = self._build_plans()
        h = F.embedding(input_ids, self.shared.weight).contiguous()
        bias = self._position_bias(S, input_ids.device)
        for blk, w_qkv in zip(self.encoder.block, self._plans):
            att, ff = blk.layer[0], blk.layer[1]
            n = ops.rmsnorm_fwd(h, att.layer_norm.weight, eps)
            qkv = ops.gemm([n], [w_qkv])
            q, k, v = (qkv[:, :, i * inner:(i + 1) * inner].unflatten(-1, (H, hd)) for i in range(3))
            o, _ = ops.attn_fwd(q, k, v, scale=1.0, bias=bias)            # T5 does not scale the scores
            ops.gemm([o.view(B, S, inner)], [att.SelfAttention.o.weight], None, out=h, epi=ops.EPI_ADD_RES, res=h)
            n = ops.rmsnorm_fwd(h, ff.layer_norm.weight, eps)
            d = ff.DenseReluDense
            g = ops.gemm([n], [d.wi_0.weight], None, epi=ops.EPI_GELU)      # "gelu_new" = the tanh approximation
            u = ops.gemm([n], [d.wi_1.weight], None, epi=ops.EPI_MUL, aux=g)
            ops.gemm([u], [d.wo.weight], None, out=h, epi=ops.EPI_ADD_RES, res=h)
        last = ops.rmsnorm_fwd(h, self.encoder.final_layer_norm.weight, eps)
        if not return_dict:
            return (last,)
        return ModelOutput(last_hidden_state=last)
