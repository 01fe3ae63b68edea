"""`CLIPTextModel` on libstb200 — mirror of transformers' CLIPTextModel as the reference calls it
(`self.text_encoder(text_input_ids, output_hidden_states=False).pooler_output`, simpletuner/helpers/models/flux/pipeline.py:
1127-1130): transformers state-dict names (`text_model.encoder.layers.N.self_attn.q_proj.weight` ...), eval-mode forward with
the causal mask only (the reference passes no attention mask).

Per layer: LayerNorm (the LN-modulate kernel with scale = w - 1, shift = b) -> fused q|k|v GEMM (+bias) -> attention with the
causal mask as an additive bias tile, scale hd^-0.5 -> out-projection GEMM with the residual add -> LayerNorm -> fc1 GEMM with
the quick-GELU epilogue -> fc2 GEMM with the residual add.  Pooling: the hidden state at the end-of-text token (both rules of
CLIPTextTransformer.forward: argmax(input_ids) when eos_token_id == 2, first eos position otherwise).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..flux.transformer import Linear, _Weight
from .t5 import ModelOutput


class _LN(nn.Module):
    def __init__(self, dim, dtype):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim, dtype=dtype), requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(dim, dtype=dtype), requires_grad=False)


class _Attn(nn.Module):
    def __init__(self, D, dtype):
        super().__init__()
        self.q_proj, self.k_proj = Linear(D, D, dtype=dtype), Linear(D, D, dtype=dtype)
        self.v_proj, self.out_proj = Linear(D, D, dtype=dtype), Linear(D, D, dtype=dtype)


class _MLP(nn.Module):
    def __init__(self, D, I, dtype):
        super().__init__()
        self.fc1, self.fc2 = Linear(D, I, dtype=dtype), Linear(I, D, dtype=dtype)


class _Layer(nn.Module):
    def __init__(self, D, I, dtype):
        super().__init__()
        self.self_attn = _Attn(D, dtype)
        self.layer_norm1 = _LN(D, dtype)
        self.mlp = _MLP(D, I, dtype)
        self.layer_norm2 = _LN(D, dtype)


class _Encoder(nn.Module):
    def __init__(self, n, D, I, dtype):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(D, I, dtype) for _ in range(n)])


class _Embeddings(nn.Module):
    def __init__(self, vocab, pos, D, dtype):
        super().__init__()
        self.token_embedding = _Weight((vocab, D), dtype)
        self.position_embedding = _Weight((pos, D), dtype)
        for p in self.parameters():
            p.requires_grad_(False)


class _TextTransformer(nn.Module):
    def __init__(self, c, dtype):
        super().__init__()
        self.embeddings = _Embeddings(c.vocab_size, c.max_position_embeddings, c.hidden_size, dtype)
        self.encoder = _Encoder(c.num_hidden_layers, c.hidden_size, c.intermediate_size, dtype)
        self.final_layer_norm = _LN(c.hidden_size, dtype)


class CLIPTextModel(nn.Module):
    def __init__(self, vocab_size: int = 49408, hidden_size: int = 768, intermediate_size: int = 3072, num_hidden_layers: int = 12,
                 num_attention_heads: int = 12, max_position_embeddings: int = 77, layer_norm_eps: float = 1e-5,
                 eos_token_id: int = 2, hidden_act: str = "quick_gelu", dtype=torch.bfloat16, **unused):
        super().__init__()
        if hidden_act != "quick_gelu":
            raise NotImplementedError("libstb200 CLIP text model implements hidden_act = quick_gelu (CLIP-L) only")
        if hidden_size // num_attention_heads not in (64, 128):
            raise NotImplementedError("libstb200 attention supports head_dim 64 / 128")
        self.config = SimpleNamespace(vocab_size=vocab_size, hidden_size=hidden_size, intermediate_size=intermediate_size,
                                      num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                                      max_position_embeddings=max_position_embeddings, layer_norm_eps=layer_norm_eps,
                                      eos_token_id=eos_token_id, hidden_act=hidden_act)
        self.text_model = _TextTransformer(self.config, dtype)
        self._plans: Optional[list] = None
        self.eval()

    @property
    def dtype(self):
        return self.text_model.final_layer_norm.weight.dtype

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        sd = {k: v for k, v in state_dict.items() if not k.endswith("position_ids")}     # (a buffer in older checkpoints)
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._plans = None
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._plans = None
        return out

    def _build_plans(self):
        plans = []
        for lyr in self.text_model.encoder.layers:
            a = lyr.self_attn
            plans.append((torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).detach().contiguous(),
                          torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).detach().contiguous()))
        return plans

    @staticmethod
    def _ln(x, ln: _LN, eps: float):
        B = x.shape[0]
        scale = (ln.weight.float() - 1.0).to(x.dtype).expand(B, -1).contiguous()     # LN(x) * (1 + scale) + shift
        shift = ln.bias.expand(B, -1).contiguous()
        return ops.ln_modulate_fwd(x, shift, scale, eps)

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask=None, position_ids=None, output_hidden_states: bool = False,
                return_dict: bool = True, **unused):
        if attention_mask is not None or position_ids is not None:
            raise NotImplementedError("the reference calls the CLIP text model with input_ids only (flux/pipeline.py:1127)")
        if not input_ids.is_cuda:
            from .._lib import StbError
            raise StbError("CLIPTextModel (libstb200) needs CUDA tensors; there is no CPU fallback")
        c = self.config
        tm = self.text_model
        B, S = input_ids.shape
        D, H = c.hidden_size, c.num_attention_heads
        hd = D // H
        if self._plans is None:
            self._plans = self._build_plans()
        h = (F.embedding(input_ids, tm.embeddings.token_embedding.weight) + tm.embeddings.position_embedding.weight[:S][None]).contiguous()
        causal = torch.full((S, S), float("-inf"), device=h.device, dtype=torch.float32).triu(1).to(h.dtype)[None].contiguous()
        for lyr, (w_qkv, b_qkv) in zip(tm.encoder.layers, self._plans):
            n = self._ln(h, lyr.layer_norm1, c.layer_norm_eps)
            qkv = ops.gemm([n], [w_qkv], b_qkv)
            q, k, v = (qkv[:, :, i * D:(i + 1) * D].unflatten(-1, (H, hd)) for i in range(3))
            o, _ = ops.attn_fwd(q, k, v, scale=hd ** -0.5, bias=causal)
            a = lyr.self_attn
            ops.gemm([o.view(B, S, D)], [a.out_proj.weight], a.out_proj.bias, out=h, epi=ops.EPI_ADD_RES, res=h)
            n = self._ln(h, lyr.layer_norm2, c.layer_norm_eps)
            u = ops.gemm([n], [lyr.mlp.fc1.weight], lyr.mlp.fc1.bias, epi=ops.EPI_QUICK_GELU)
            ops.gemm([u], [lyr.mlp.fc2.weight], lyr.mlp.fc2.bias, out=h, epi=ops.EPI_ADD_RES, res=h)
        last = self._ln(h, tm.final_layer_norm, c.layer_norm_eps)
        ids = input_ids.to(torch.int)
        idx = ids.argmax(dim=-1) if c.eos_token_id == 2 else (ids == c.eos_token_id).int().argmax(dim=-1)
        pooled = last[torch.arange(B, device=last.device), idx]
        if not return_dict:
            return (last, pooled)
        return ModelOutput(last_hidden_state=last, pooler_output=pooled)
