"""CPU oracle for the SD3 / SD3.5 MMDiT training step — TEST INFRASTRUCTURE ONLY (see flux_oracle.py header).

Restates reference simpletuner/helpers/models/sd3/transformer.py:
  * SD3Transformer2DModel.forward                 :560-909  (default path: no TREAD / controlnet / flow-map)
  * _sd3_apply_joint_transformer_block            :145-241  (adaLN-Zero both streams, joint attention,
        optional image-only `attn2` for layers in `dual_attention_layers`, `context_pre_only` last block)
  * SD3._model_predict_single                     sd3/model.py:540-569 (raw 0..1000 timesteps cast to the
        bf16 weight dtype — quirk Q8 — and latents passed un-packed)
and the diffusers classes it instantiates (JointTransformerBlock, JointAttnProcessor2_0 — joint order
[image, text] —, SD35AdaLayerNormZeroX, PatchEmbed with a persistent `pos_embed` buffer,
CombinedTimestepTextProjEmbeddings, AdaLayerNormContinuous).
PARITY STATUS: **parity unpinned** — the reference's tests hold no golden vector for this forward
(SURVEY.md §4); the shared primitives (RMSNorm, adaLN, attention, flow prep / loss) are the same
restatements the Flux oracle uses.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .flux_oracle import layer_norm_noaffine, linear, rms_norm, sdpa, timestep_embedding

Tensor = torch.Tensor


@dataclass
class SD3Config:
    """Mirror of SD3Transformer2DModel.__init__ (reference sd3/transformer.py:282-300)."""
    sample_size: int = 128
    patch_size: int = 2
    in_channels: int = 16
    num_layers: int = 24
    attention_head_dim: int = 64
    num_attention_heads: int = 24
    joint_attention_dim: int = 4096
    caption_projection_dim: int = 1536
    pooled_projection_dim: int = 2048
    out_channels: int = 16
    pos_embed_max_size: int = 384
    dual_attention_layers: Tuple[int, ...] = tuple(range(13))
    qk_norm: Optional[str] = "rms_norm"

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim


def sd3_param_shapes(cfg: SD3Config) -> Dict[str, Tuple[int, ...]]:
    D, hd = cfg.inner_dim, cfg.attention_head_dim
    sh: Dict[str, Tuple[int, ...]] = {}

    def lin(name, o, i):
        sh[name + ".weight"] = (o, i)
        sh[name + ".bias"] = (o,)

    sh["pos_embed.proj.weight"] = (D, cfg.in_channels, cfg.patch_size, cfg.patch_size)
    sh["pos_embed.proj.bias"] = (D,)
    sh["pos_embed.pos_embed"] = (1, cfg.pos_embed_max_size ** 2, D)
    lin("time_text_embed.timestep_embedder.linear_1", D, 256)
    lin("time_text_embed.timestep_embedder.linear_2", D, D)
    lin("time_text_embed.text_embedder.linear_1", D, cfg.pooled_projection_dim)
    lin("time_text_embed.text_embedder.linear_2", D, D)
    lin("context_embedder", cfg.caption_projection_dim, cfg.joint_attention_dim)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        last = i == cfg.num_layers - 1
        dual = i in cfg.dual_attention_layers
        lin(p + "norm1.linear", (9 if dual else 6) * D, D)
        lin(p + "norm1_context.linear", (2 if last else 6) * D, D)
        names = ["to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0"] + ([] if last else ["to_add_out"])
        for n in names:
            lin(p + "attn." + n, D, D)
        if cfg.qk_norm:
            for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
                sh[p + f"attn.{n}.weight"] = (hd,)
        if dual:
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(p + "attn2." + n, D, D)
            if cfg.qk_norm:
                for n in ("norm_q", "norm_k"):
                    sh[p + f"attn2.{n}.weight"] = (hd,)
        lin(p + "ff.net.0.proj", 4 * D, D)
        lin(p + "ff.net.2", D, 4 * D)
        if not last:
            lin(p + "ff_context.net.0.proj", 4 * D, D)
            lin(p + "ff_context.net.2", D, 4 * D)
    lin("norm_out.linear", 2 * D, D)
    lin("proj_out", cfg.patch_size * cfg.patch_size * cfg.out_channels, D)
    return sh


def init_sd3_params(cfg: SD3Config, seed: int = 0, dtype=torch.float32, std: float = 0.02) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in sd3_param_shapes(cfg).items():
        if "norm_q" in name or "norm_k" in name or "norm_added" in name:
            t = 1.0 + 0.05 * torch.randn(shape, generator=g)
        elif name == "pos_embed.pos_embed":
            t = 0.5 * torch.randn(shape, generator=g)  # stands in for the checkpoint's sincos buffer
        elif name.endswith(".bias"):
            t = 0.5 * std * torch.randn(shape, generator=g)
        else:
            t = std * torch.randn(shape, generator=g)
        out[name] = t.to(dtype)
    return out


SD3_LORA_TARGETS = ("to_k", "to_q", "to_v", "to_out.0")
"""SD3.DEFAULT_LORA_TARGET (reference sd3/model.py:122): PEFT suffix match -> attn.* and attn2.*"""


def lora_target_names(cfg: SD3Config, targets=SD3_LORA_TARGETS):
    names = []
    for name in sd3_param_shapes(cfg):
        if not name.endswith(".weight") or ".attn" not in name:
            continue
        mod = name[: -len(".weight")]
        if any(mod.endswith("." + t) for t in targets):
            names.append(mod)
    return names


def init_lora_params(cfg: SD3Config, rank: int, seed: int = 1, b_std: float = 0.02, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    D = cfg.inner_dim
    out = {}
    bound = D ** -0.5
    for n in lora_target_names(cfg):
        out[n + ".lora_A.weight"] = ((torch.rand((rank, D), generator=g) * 2 - 1) * bound).to(dtype)
        out[n + ".lora_B.weight"] = (b_std * torch.randn((D, rank), generator=g)).to(dtype)
    return out


def cropped_pos_embed(P, cfg: SD3Config, height: int, width: int) -> Tensor:
    """diffusers PatchEmbed.cropped_pos_embed: centre crop of the [max, max] grid to the (H/p, W/p) patch grid."""
    h, w = height // cfg.patch_size, width // cfg.patch_size
    mx = cfg.pos_embed_max_size
    top, left = (mx - h) // 2, (mx - w) // 2
    pe = P["pos_embed.pos_embed"].reshape(1, mx, mx, -1)[:, top:top + h, left:left + w, :]
    return pe.reshape(1, h * w, -1)


def joint_attention(P, cfg, prefix, x, enc, lora, lora_scale, pre_only):
    """diffusers JointAttnProcessor2_0: joint order [image, text]."""
    H, hd = cfg.num_attention_heads, cfg.attention_head_dim
    B = x.shape[0]

    def heads(t):
        return t.view(B, -1, H, hd).transpose(1, 2)

    q = heads(linear(x, P, prefix + "to_q", lora, lora_scale))
    k = heads(linear(x, P, prefix + "to_k", lora, lora_scale))
    v = heads(linear(x, P, prefix + "to_v", lora, lora_scale))
    if cfg.qk_norm:
        q = rms_norm(q, P[prefix + "norm_q.weight"], 1e-6)
        k = rms_norm(k, P[prefix + "norm_k.weight"], 1e-6)
    n_img = x.shape[1]
    if enc is not None:
        eq = heads(linear(enc, P, prefix + "add_q_proj", lora, lora_scale))
        ek = heads(linear(enc, P, prefix + "add_k_proj", lora, lora_scale))
        ev = heads(linear(enc, P, prefix + "add_v_proj", lora, lora_scale))
        if cfg.qk_norm:
            eq = rms_norm(eq, P[prefix + "norm_added_q.weight"], 1e-6)
            ek = rms_norm(ek, P[prefix + "norm_added_k.weight"], 1e-6)
        q = torch.cat([q, eq], dim=2)
        k = torch.cat([k, ek], dim=2)
        v = torch.cat([v, ev], dim=2)
    o = sdpa(q, k, v).transpose(1, 2).reshape(B, -1, H * hd).to(q.dtype)
    if enc is not None:
        o, eo = o[:, :n_img], o[:, n_img:]
        o = linear(o, P, prefix + "to_out.0", lora, lora_scale)
        eo = None if pre_only else linear(eo, P, prefix + "to_add_out", lora, lora_scale)
        return o, eo
    return linear(o, P, prefix + "to_out.0", lora, lora_scale)


def sd3_joint_block(P, cfg, i, x, enc, temb, lora, lora_scale):
    """reference sd3/transformer.py:145-241."""
    p = f"transformer_blocks.{i}."
    last = i == cfg.num_layers - 1
    dual = i in cfg.dual_attention_layers
    mod = F.linear(F.silu(temb), P[p + "norm1.linear.weight"], P[p + "norm1.linear.bias"])
    if dual:
        sh_a, sc_a, g_a, sh_m, sc_m, g_m, sh_a2, sc_a2, g_a2 = mod.chunk(9, dim=1)
    else:
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)
    ln = layer_norm_noaffine(x)
    nx = ln * (1 + sc_a[:, None]) + sh_a[:, None]
    cmod = F.linear(F.silu(temb), P[p + "norm1_context.linear.weight"], P[p + "norm1_context.linear.bias"])
    if last:
        c_scale, c_shift = cmod.chunk(2, dim=1)
        nenc = layer_norm_noaffine(enc) * (1 + c_scale)[:, None, :] + c_shift[:, None, :]
    else:
        csh_a, csc_a, cg_a, csh_m, csc_m, cg_m = cmod.chunk(6, dim=1)
        nenc = layer_norm_noaffine(enc) * (1 + csc_a[:, None]) + csh_a[:, None]
    ao, eo = joint_attention(P, cfg, p + "attn.", nx, nenc, lora, lora_scale, last)
    x = x + g_a.unsqueeze(1) * ao
    if dual:
        nx2 = ln * (1 + sc_a2[:, None]) + sh_a2[:, None]
        ao2 = joint_attention(P, cfg, p + "attn2.", nx2, None, lora, lora_scale, False)
        x = x + g_a2.unsqueeze(1) * ao2
    nx = layer_norm_noaffine(x) * (1 + sc_m[:, None]) + sh_m[:, None]
    ff = F.linear(F.gelu(F.linear(nx, P[p + "ff.net.0.proj.weight"], P[p + "ff.net.0.proj.bias"]), approximate="tanh"),
                  P[p + "ff.net.2.weight"], P[p + "ff.net.2.bias"])
    x = x + g_m.unsqueeze(1) * ff
    if last:
        return None, x
    enc = enc + cg_a.unsqueeze(1) * eo
    nenc = layer_norm_noaffine(enc) * (1 + csc_m[:, None]) + csh_m[:, None]
    cff = F.linear(F.gelu(F.linear(nenc, P[p + "ff_context.net.0.proj.weight"], P[p + "ff_context.net.0.proj.bias"]), approximate="tanh"),
                   P[p + "ff_context.net.2.weight"], P[p + "ff_context.net.2.bias"])
    enc = enc + cg_m.unsqueeze(1) * cff
    return enc, x


def sd3_forward(P, cfg: SD3Config, hidden_states: Tensor, encoder_hidden_states: Tensor, pooled_projections: Tensor,
                timestep: Tensor, lora=None, lora_scale: float = 1.0) -> Tensor:
    """hidden_states [B, C, H, W] latents -> [B, C_out, H, W]."""
    B, _, Hh, Ww = hidden_states.shape
    ps = cfg.patch_size
    x = F.conv2d(hidden_states, P["pos_embed.proj.weight"], P["pos_embed.proj.bias"], stride=ps)
    x = x.flatten(2).transpose(1, 2)  # [B, (H/p)(W/p), D]
    x = (x + cropped_pos_embed(P, cfg, Hh, Ww)).to(x.dtype)
    dt = pooled_projections.dtype
    pre = "time_text_embed."

    def mlp(v, name):
        h = F.silu(F.linear(v, P[pre + name + ".linear_1.weight"], P[pre + name + ".linear_1.bias"]))
        return F.linear(h, P[pre + name + ".linear_2.weight"], P[pre + name + ".linear_2.bias"])

    temb = mlp(timestep_embedding(timestep).to(dt), "timestep_embedder") + mlp(pooled_projections, "text_embedder")
    enc = F.linear(encoder_hidden_states, P["context_embedder.weight"], P["context_embedder.bias"])
    for i in range(cfg.num_layers):
        enc, x = sd3_joint_block(P, cfg, i, x, enc, temb, lora, lora_scale)
    emb = F.linear(F.silu(temb).to(x.dtype), P["norm_out.linear.weight"], P["norm_out.linear.bias"])
    scale, shift = torch.chunk(emb, 2, dim=1)
    x = layer_norm_noaffine(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    x = F.linear(x, P["proj_out.weight"], P["proj_out.bias"])
    h, w = Hh // ps, Ww // ps
    x = x.reshape(B, h, w, ps, ps, cfg.out_channels)
    x = torch.einsum("nhwpqc->nchpwq", x)
    return x.reshape(B, cfg.out_channels, h * ps, w * ps)


def sd3_model_predict(P, cfg, noisy_latents, timesteps, prompt_embeds, pooled, lora=None, lora_scale=1.0,
                      weight_dtype=torch.bfloat16):
    """SD3._model_predict_single (reference sd3/model.py:540-569): raw timesteps cast to the weight dtype (quirk Q8)."""
    t = timesteps.to(weight_dtype).to(torch.float32)
    return sd3_forward(P, cfg, noisy_latents, prompt_embeds, pooled, t, lora, lora_scale)
