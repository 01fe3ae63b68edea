"""SD3Transformer2DModel on libstb200 — B200-native drop-in for reference
simpletuner/helpers/models/sd3/transformer.py:244-909 (class SD3Transformer2DModel), LoRA-training path.

Same constructor arguments / `.config`, forward kwargs (`hidden_states [B,C,H,W]`, `encoder_hidden_states`,
`pooled_projections`, `timestep`, `return_dict=False`) and `(Tensor[B,C_out,H,W],)` return as the reference;
diffusers / PEFT parameter names (`transformer_blocks.N.attn.to_q.weight`, `…attn2.to_out.0.lora_A.default.weight`,
`pos_embed.proj.weight`, the persistent `pos_embed.pos_embed` buffer …).

The joint blocks run on the shared `DoubleBlockFn` schedule (flux/blocks.py): SD3 has no RoPE, so the joint
attention is order-invariant and the [text | image] joint buffer of the Flux path is reused unchanged; SD3.5's
image-only `attn2` (dual_attention_layers) and the `context_pre_only` last block are flags of that schedule.
The 2x2 stride-2 PatchEmbed conv is a K=64 GEMM on the patchified latents (same (c, dy, dx) feature order as
Flux `pack_latents`) with the cropped positional table added in the epilogue.
Unsupported: TREAD routing, controlnet residuals, token-wise timesteps, flow-map / TwinFlow inputs, full
— they raise so a shim can keep the reference module.  Full fine-tune (model_type=full, BASELINE configs[2]) runs
through sd3/fullft.py after `enable_full_finetune()`.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..flux.blocks import AttnPlan, DoubleBlockFn, MlpPlan, TailFn, _t, _wt
from ..flux.transformer import AttnProcessorAPI, LoraDropoutAPI, Linear, RMSNormWeight, _AdaNorm, _FeedForward, _TimestepEmbedding, _attn_plan, _lora_list, _sinusoid

SD3_LORA_TARGETS = ["to_k", "to_q", "to_v", "to_out.0"]  # SD3.DEFAULT_LORA_TARGET, reference sd3/model.py:122


class _JointAttention(nn.Module):
    def __init__(self, dim, heads, head_dim, joint: bool, pre_only: bool, qk_norm: bool, dtype):
        super().__init__()
        self.to_q = Linear(dim, dim, dtype=dtype)
        self.to_k = Linear(dim, dim, dtype=dtype)
        self.to_v = Linear(dim, dim, dtype=dtype)
        self.to_out = nn.ModuleList([Linear(dim, dim, dtype=dtype), nn.Identity()])
        if qk_norm:
            self.norm_q = RMSNormWeight(head_dim, dtype)
            self.norm_k = RMSNormWeight(head_dim, dtype)
        if joint:
            self.add_q_proj = Linear(dim, dim, dtype=dtype)
            self.add_k_proj = Linear(dim, dim, dtype=dtype)
            self.add_v_proj = Linear(dim, dim, dtype=dtype)
            if not pre_only:
                self.to_add_out = Linear(dim, dim, dtype=dtype)
            if qk_norm:
                self.norm_added_q = RMSNormWeight(head_dim, dtype)
                self.norm_added_k = RMSNormWeight(head_dim, dtype)


class _NoNorm:
    weight = None


def _plan(q, k, v, out, nq, nk, dyn: bool = False) -> AttnPlan:
    w_qkv = torch.cat([q.weight.detach(), k.weight.detach(), v.weight.detach()], 0).contiguous()
    b_qkv = torch.cat([q.bias.detach(), k.bias.detach(), v.bias.detach()], 0).contiguous()
    p = AttnPlan(w_qkv, b_qkv, _wt(w_qkv, dyn), norm_q=None if nq is None else nq.weight.detach(),
                 norm_k=None if nk is None else nk.weight.detach())
    if out is not None:
        p.w_out, p.b_out, p.w_out_t = out.weight.detach(), out.bias.detach(), _wt(out.weight.detach(), dyn)
    return p


class JointTransformerBlock(nn.Module):
    """diffusers JointTransformerBlock as configured at reference sd3/transformer.py:360-367."""

    def __init__(self, dim, heads, head_dim, context_pre_only: bool, qk_norm: Optional[str], use_dual_attention: bool, dtype):
        super().__init__()
        self.dim, self.heads, self.head_dim = dim, heads, head_dim
        self.context_pre_only, self.use_dual_attention, self.qk_norm = context_pre_only, use_dual_attention, bool(qk_norm)
        self.norm1 = _AdaNorm(dim, 9 if use_dual_attention else 6, dtype)
        self.norm1_context = _AdaNorm(dim, 2 if context_pre_only else 6, dtype)
        self.attn = _JointAttention(dim, heads, head_dim, True, context_pre_only, bool(qk_norm), dtype)
        if use_dual_attention:
            self.attn2 = _JointAttention(dim, heads, head_dim, False, False, bool(qk_norm), dtype)
        self.ff = _FeedForward(dim, dtype)
        if not context_pre_only:
            self.ff_context = _FeedForward(dim, dtype)
        self._plans = None

    def plans(self):
        if self._plans is None:
            a = self.attn
            g = lambda m, n: getattr(m, n, None)
            dyn = bool(getattr(self, "_full_ft", False))      # full fine-tune: the weights change every step
            mk = lambda ff: MlpPlan(ff.net[0].proj.weight.detach(), ff.net[0].proj.bias.detach(), _wt(ff.net[0].proj.weight.detach(), dyn),
                                    ff.net[2].weight.detach(), ff.net[2].bias.detach(), _wt(ff.net[2].weight.detach(), dyn))
            pl = {
                "img_attn": _plan(a.to_q, a.to_k, a.to_v, a.to_out[0], g(a, "norm_q"), g(a, "norm_k"), dyn),
                "txt_attn": _plan(a.add_q_proj, a.add_k_proj, a.add_v_proj, g(a, "to_add_out"), g(a, "norm_added_q"), g(a, "norm_added_k"), dyn),
                "img_mlp": mk(self.ff),
            }
            if not self.context_pre_only:
                pl["txt_mlp"] = mk(self.ff_context)
            if self.use_dual_attention:
                b = self.attn2
                pl["img_attn2"] = _plan(b.to_q, b.to_k, b.to_v, b.to_out[0], g(b, "norm_q"), g(b, "norm_k"), dyn)
            self._plans = pl
        return self._plans

    FULL_PARAM_ORDER = None

    def full_param_names(self):
        """Block-local names of every trainable tensor the full fine-tune schedule produces a gradient for (adaLN linears
        are differentiated by torch autograd outside the block Function)."""
        if self.FULL_PARAM_ORDER is None:
            skip = ("norm1.", "norm1_context.")
            self.FULL_PARAM_ORDER = [n for n, _ in self.named_parameters() if not n.startswith(skip) and ".lora_" not in n]
        return self.FULL_PARAM_ORDER

    def forward(self, h, silu_temb, S_txt, lora_scaling):
        mod_img = self.norm1.linear(silu_temb)
        mod_txt = self.norm1_context.linear(silu_temb)
        if getattr(self, "_full_ft", False):
            from .fullft import JointBlockFullFn
            names = self.full_param_names()
            params = dict(self.named_parameters())
            st = {"S_txt": S_txt, "H": self.heads, "hd": self.head_dim, "plans": self.plans(),
                  "context_pre_only": self.context_pre_only, "dual": self.use_dual_attention, "param_names": names}
            return JointBlockFullFn.apply(h, mod_img, mod_txt, st, *[params[n] for n in names])
        st = {"S_txt": S_txt, "H": self.heads, "hd": self.head_dim, "plans": self.plans(), "lora_scaling": lora_scaling,
              "nan_to_num_txt": False, "context_pre_only": self.context_pre_only, "dual": self.use_dual_attention,
              "_n_lora": 24, "lora_drop": getattr(self, "_lora_drop", None)}
        a = self.attn
        lins = [a.to_q, a.to_k, a.to_v, a.to_out[0], a.add_q_proj, a.add_k_proj, a.add_v_proj]
        lora = _lora_list(lins)
        tao = getattr(a, "to_add_out", None)
        lora += list(tao.lora_tensors()) if tao is not None else [None, None]
        if self.use_dual_attention:
            b = self.attn2
            lora += _lora_list([b.to_q, b.to_k, b.to_v, b.to_out[0]])
        else:
            lora += [None] * 8
        return DoubleBlockFn.apply(h, mod_img, mod_txt, None, None, st, *lora)


class _PatchEmbed(nn.Module):
    """diffusers PatchEmbed(patch_size=2, pos_embed_max_size=...): conv `proj` + persistent sincos `pos_embed` buffer."""

    def __init__(self, in_channels, dim, patch_size, max_size, dtype):
        super().__init__()
        self.patch_size, self.max_size = patch_size, max_size
        self.proj = nn.Conv2d(in_channels, dim, kernel_size=patch_size, stride=patch_size, bias=True, dtype=dtype)
        for p in self.proj.parameters():
            p.requires_grad_(False)
        self.register_buffer("pos_embed", torch.zeros(1, max_size * max_size, dim, dtype=torch.float32), persistent=True)


class _TimeTextEmbed(nn.Module):
    def __init__(self, dim, pooled_dim, dtype):
        super().__init__()
        self.timestep_embedder = _TimestepEmbedding(256, dim, dtype)
        self.text_embedder = _TimestepEmbedding(pooled_dim, dim, dtype)


class SD3Transformer2DModel(AttnProcessorAPI, LoraDropoutAPI, nn.Module):
    _no_split_modules = ["JointTransformerBlock"]

    def __init__(self, sample_size: int = 128, patch_size: int = 2, in_channels: int = 16, num_layers: int = 18,
                 attention_head_dim: int = 64, num_attention_heads: int = 18, joint_attention_dim: int = 4096,
                 caption_projection_dim: int = 1152, pooled_projection_dim: int = 2048, out_channels: int = 16,
                 pos_embed_max_size: int = 96, dual_attention_layers: Tuple[int, ...] = (), qk_norm: Optional[str] = None,
                 dtype=torch.bfloat16, **unused):
        super().__init__()
        if patch_size != 2:
            raise NotImplementedError("libstb200 SD3 path supports patch_size=2")
        if attention_head_dim not in (64, 128):
            raise NotImplementedError("libstb200 attention supports head_dim 64 / 128")
        self.config = SimpleNamespace(sample_size=sample_size, patch_size=patch_size, in_channels=in_channels,
                                      num_layers=num_layers, attention_head_dim=attention_head_dim,
                                      num_attention_heads=num_attention_heads, joint_attention_dim=joint_attention_dim,
                                      caption_projection_dim=caption_projection_dim, pooled_projection_dim=pooled_projection_dim,
                                      out_channels=out_channels, pos_embed_max_size=pos_embed_max_size,
                                      dual_attention_layers=tuple(dual_attention_layers), qk_norm=qk_norm)
        self.out_channels = out_channels if out_channels is not None else in_channels
        self.inner_dim = D = num_attention_heads * attention_head_dim
        if caption_projection_dim != D:
            raise ValueError("caption_projection_dim must equal the inner dim (as in every released SD3 checkpoint)")
        self.pos_embed = _PatchEmbed(in_channels, D, patch_size, pos_embed_max_size, dtype)
        self.time_text_embed = _TimeTextEmbed(D, pooled_projection_dim, dtype)
        self.context_embedder = Linear(joint_attention_dim, caption_projection_dim, dtype=dtype)
        self.transformer_blocks = nn.ModuleList([
            JointTransformerBlock(D, num_attention_heads, attention_head_dim, i == num_layers - 1, qk_norm,
                                  i in self.config.dual_attention_layers, dtype) for i in range(num_layers)])
        self.norm_out = _AdaNorm(D, 2, dtype)
        self.proj_out = Linear(D, patch_size * patch_size * self.out_channels, dtype=dtype)
        self._lora_scaling = 1.0
        self._tail_plan = None
        self._pos_cache: Dict[Any, torch.Tensor] = {}
        self.peft_config: Dict[str, Any] = {}
        self.gradient_checkpointing = False

    # ---- reference-facing utilities -------------------------------------------------------------
    def enable_gradient_checkpointing(self):
        """Re-run every joint block in backward (torch.utils.checkpoint, as the reference does); off by default."""
        self.gradient_checkpointing = True

    def disable_gradient_checkpointing(self):
        self.gradient_checkpointing = False

    def _run_block(self, blk, *args):
        if self.gradient_checkpointing and torch.is_grad_enabled():
            from torch.utils.checkpoint import checkpoint
            return checkpoint(blk, *args, use_reentrant=False)
        return blk(*args)

    def invalidate_plans(self):
        for blk in self.transformer_blocks:
            blk._plans = None
        self._tail_plan = None
        self._pos_cache.clear()

    def enable_full_finetune(self, enabled: bool = True):
        """model_type=full (BASELINE configs[2]): every parameter trains.  The joint blocks switch to the full fine-tune
        schedule (sd3/fullft.py); fused / transposed weight layouts are rebuilt after every optimizer step
        (`after_optimizer_step`, called by TrainStep) because the base weights now change."""
        if enabled and self.lora_linears():
            raise NotImplementedError("full fine-tune and LoRA adapters are mutually exclusive on the libstb200 path")
        self._full_ft = bool(enabled)
        for p in self.parameters():
            p.requires_grad_(bool(enabled))
        for blk in self.transformer_blocks:
            blk._full_ft = bool(enabled)
        self.invalidate_plans()
        return self

    def after_optimizer_step(self):
        if getattr(self, "_full_ft", False):
            self.invalidate_plans()

    def before_graph_capture(self):
        """training.step.GraphedTrainStep: in full fine-tune the fused / transposed layouts are rebuilt INSIDE the captured
        step (they depend on weights the optimizer changes between replays)."""
        if getattr(self, "_full_ft", False):
            self.invalidate_plans()

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.invalidate_plans()
        return out

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.invalidate_plans()
        return out

    def lora_linears(self) -> Dict[str, Linear]:
        return {n: m for n, m in self.named_modules() if isinstance(m, Linear) and m.lora_A is not None}

    def add_adapter(self, lora_config=None, adapter_name: str = "default", *, rank: Optional[int] = None,
                    lora_alpha: Optional[float] = None, target_modules: Optional[Sequence[str]] = None,
                    lora_dropout: float = 0.0):
        if lora_config is not None:
            rank = getattr(lora_config, "r", rank)
            lora_alpha = getattr(lora_config, "lora_alpha", lora_alpha)
            target_modules = getattr(lora_config, "target_modules", target_modules)
            lora_dropout = getattr(lora_config, "lora_dropout", lora_dropout)
        lora_dropout = self._check_dropout_p(lora_dropout)
        if not 1 <= rank <= 128:
            raise NotImplementedError("fused LoRA path supports rank 1..128 (one 128-wide rank block per adapted Linear)")
        lora_alpha = float(lora_alpha) if lora_alpha is not None else float(rank)
        targets = list(target_modules) if target_modules is not None else SD3_LORA_TARGETS
        supported = {"to_q", "to_k", "to_v", "to_out.0", "add_q_proj", "add_k_proj", "add_v_proj", "to_add_out"}
        n = 0
        for name, mod in self.named_modules():
            if not isinstance(mod, Linear) or ".attn" not in name:
                continue
            hit = [t for t in targets if name.endswith("." + t)]  # PEFT suffix matching
            if not hit:
                continue
            if hit[0] not in supported:
                raise NotImplementedError(f"LoRA target {hit[0]} is not supported by the fused path")
            mod.add_lora(rank, lora_alpha, adapter_name)
            n += 1
        if n == 0:
            raise ValueError(f"no module matched LoRA targets {targets}")
        self._lora_scaling = lora_alpha / rank
        self._lora_dropout_p = lora_dropout
        self.peft_config[adapter_name] = SimpleNamespace(r=rank, lora_alpha=lora_alpha, target_modules=targets, lora_dropout=lora_dropout)
        return n

    def disable_lora(self):
        for m in self.lora_linears().values():
            m.lora_enabled = False

    def enable_lora(self):
        for m in self.lora_linears().values():
            m.lora_enabled = True

    def trainable_parameters(self):
        return [p for p in self.parameters() if p.requires_grad]

    # ---- forward --------------------------------------------------------------------------------
    def _cropped_pos(self, h: int, w: int, dtype) -> torch.Tensor:
        """diffusers PatchEmbed.cropped_pos_embed: centre crop of the [max, max] table to the (h, w) patch grid."""
        key = (h, w, dtype)
        hit = self._pos_cache.get(key)
        if hit is None:
            mx = self.pos_embed.max_size
            if h > mx or w > mx:
                raise ValueError(f"patch grid {h}x{w} exceeds pos_embed_max_size {mx}")
            top, left = (mx - h) // 2, (mx - w) // 2
            pe = self.pos_embed.pos_embed.reshape(1, mx, mx, -1)[:, top:top + h, left:left + w, :]
            hit = pe.reshape(1, h * w, -1).to(dtype).contiguous()
            self._pos_cache[key] = hit
        return hit

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: torch.Tensor = None,
                pooled_projections: torch.Tensor = None, timestep: torch.Tensor = None, timestep_sign=None, r_timestep=None,
                block_controlnet_hidden_states=None, joint_attention_kwargs: Optional[Dict[str, Any]] = None,
                return_dict: bool = True, skip_layers=None, force_keep_mask=None, hidden_states_buffer=None,
                grounding_kwargs=None, _packed_latents: Optional[torch.Tensor] = None, _packed_output: bool = False):
        for nm, v in (("timestep_sign", timestep_sign), ("r_timestep", r_timestep), ("skip_layers", skip_layers),
                      ("block_controlnet_hidden_states", block_controlnet_hidden_states),
                      ("force_keep_mask", force_keep_mask), ("grounding_kwargs", grounding_kwargs)):
            if v is not None:
                raise NotImplementedError(f"libstb200 SD3 path does not support `{nm}`; use the reference module")
        if timestep.ndim != 1:
            raise NotImplementedError("token-wise timesteps are not supported by the libstb200 SD3 path")
        if not hidden_states.is_cuda:
            from .._lib import StbError
            raise StbError("SD3Transformer2DModel (libstb200) needs CUDA tensors; there is no CPU fallback")
        dt = self.context_embedder.weight.dtype
        B, C, Hh, Ww = hidden_states.shape
        hp, wp = Hh // 2, Ww // 2
        S_img, S_txt, D = hp * wp, encoder_hidden_states.shape[1], self.inner_dim
        dev = hidden_states.device
        if _packed_latents is None:
            x = hidden_states.to(dt).view(B, C, hp, 2, wp, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, S_img, C * 4)
        else:
            x = _packed_latents
        full = getattr(self, "_full_ft", False) and torch.is_grad_enabled()
        pos = self._cropped_pos(hp, wp, dt).expand(B, S_img, D)
        if full:
            from .fullft import EmbedFullFn, TailFullFn
            h = EmbedFullFn.apply(x.contiguous(), encoder_hidden_states.to(dt).contiguous(), pos, self.pos_embed.proj.weight,
                                  self.pos_embed.proj.bias, self.context_embedder.weight, self.context_embedder.bias)
        else:
            h = torch.empty((B, S_txt + S_img, D), device=dev, dtype=dt)
            # PatchEmbed: conv2x2/s2 == GEMM over (c, dy, dx) features; + cropped pos table in the epilogue
            w_pe = self.pos_embed.proj.weight.detach().reshape(D, C * 4)
            ops.gemm([x.contiguous()], [w_pe], self.pos_embed.proj.bias.detach(), out=h[:, S_txt:], epi=ops.EPI_ADD_RES, res=pos)
            ops.gemm([encoder_hidden_states.to(dt).contiguous()], [self.context_embedder.weight], self.context_embedder.bias,
                     out=h[:, :S_txt])
        tte = self.time_text_embed
        temb = tte.timestep_embedder(_sinusoid(timestep.to(dev).float()).to(dt)) + tte.text_embedder(pooled_projections.to(dt).contiguous())
        silu_temb = F.silu(temb).contiguous()
        self._begin_lora_dropout(list(self.transformer_blocks))
        for blk in self.transformer_blocks:
            h = self._run_block(blk, h, silu_temb, S_txt, self._lora_scaling)
        if self._tail_plan is None:
            self._tail_plan = {"w_proj": self.proj_out.weight.detach(), "b_proj": self.proj_out.bias.detach(),
                               "w_proj_t": _wt(self.proj_out.weight.detach(), bool(getattr(self, "_full_ft", False)))}
        mod = self.norm_out.linear(silu_temb)
        if full:
            out = TailFullFn.apply(h, mod, self.proj_out.weight, self.proj_out.bias, S_txt)
        else:
            out = TailFn.apply(h, mod, {"S_txt": S_txt, **self._tail_plan})  # [B, S_img, p*p*C_out] in (dy, dx, c) order
        if not _packed_output:
            Co = self.out_channels
            out = torch.einsum("nhwpqc->nchpwq", out.reshape(B, hp, wp, 2, 2, Co)).reshape(B, Co, hp * 2, wp * 2)
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)
