"""CPU oracle for the PixArt-Sigma denoiser + epsilon-prediction step — TEST INFRASTRUCTURE ONLY
(see flux_oracle.py header: never imported by the product path).

Restates, in plain fp32 torch:
  * PixArtTransformer2DModel.forward            reference pixart/transformer.py:499-787
  * _prepare_timestep_embeddings (adaLN-single) reference pixart/transformer.py:789-853
  * BasicTransformerBlock(ada_norm_single)      reference pixart/transformer.py:57-145 delegates 1-D timesteps to
    diffusers' BasicTransformerBlock.forward; the token-wise override at :95-145 spells out the same arithmetic
    (scale_shift_table + t -> 6 chunks; LN -> modulate -> attn1 -> gate; attn2 on the un-normed stream; LN ->
    modulate -> FF -> gate) and is what this file follows, with [B, 6, D] modulation instead of [B, S, 6, D].
  * PixartSigma._model_predict_single           reference pixart/model.py:274-319  (`.chunk(2, dim=1)[0]`)
  * _build_added_cond_kwargs fallback           reference pixart/model.py:360-379  (quirk Q4: latent-shape fallback)
  * epsilon prepare_batch + loss                reference common.py:5982-6002, 6376-6398, 6426-6429
Third-party arithmetic (diffusers >= 0.36, not vendored, not installed here) restated from its published
structure: PatchEmbed (conv 2x2/s2 + 2-D sincos table built for the input grid), AdaLayerNormSingle /
PixArtAlphaCombinedTimestepSizeEmbeddings, PixArtAlphaTextProjection (Linear -> GELU(tanh) -> Linear),
Attention/AttnProcessor2_0 (bias=True, no qk-norm, scale = head_dim**-0.5, additive key mask),
FeedForward("gelu-approximate").
PARITY STATUS: **parity unpinned** for the denoiser numerics (the reference's tests pin no
denoiser outputs); the schedule / timestep pieces used by the step are pinned in tests/test_schedule.py and
tests/test_noise.py against the reference's own source.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from .flux_oracle import layer_norm_noaffine, linear, timestep_embedding

Tensor = torch.Tensor

PIXART_LORA_TARGETS = ("to_k", "to_q", "to_v", "to_out.0")  # PixartSigma.DEFAULT_LORA_TARGET, pixart/model.py:59


@dataclass
class PixArtConfig:
    """Constructor defaults of reference pixart/transformer.py:210-236 (PixArt-Sigma-XL-2-1024)."""
    num_attention_heads: int = 16
    attention_head_dim: int = 72
    in_channels: int = 4
    out_channels: int = 8
    num_layers: int = 28
    cross_attention_dim: int = 1152
    sample_size: int = 128
    patch_size: int = 2
    caption_channels: int = 4096
    interpolation_scale: Optional[float] = None
    use_additional_conditions: Optional[bool] = None

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def additional_conditions(self) -> bool:
        if self.use_additional_conditions is None:
            return self.sample_size == 128   # pixart/transformer.py:249-253
        return bool(self.use_additional_conditions)

    @property
    def interp(self) -> float:
        return self.interpolation_scale if self.interpolation_scale is not None else max(self.sample_size // 64, 1)


def pixart_param_shapes(cfg: PixArtConfig) -> Dict[str, Tuple[int, ...]]:
    D = cfg.inner_dim
    sh: Dict[str, Tuple[int, ...]] = {}

    def lin(name, o, i):
        sh[name + ".weight"] = (o, i)
        sh[name + ".bias"] = (o,)

    sh["pos_embed.proj.weight"] = (D, cfg.in_channels, cfg.patch_size, cfg.patch_size)
    sh["pos_embed.proj.bias"] = (D,)
    lin("adaln_single.emb.timestep_embedder.linear_1", D, 256)
    lin("adaln_single.emb.timestep_embedder.linear_2", D, D)
    if cfg.additional_conditions:
        s = D // 3
        for n in ("resolution_embedder", "aspect_ratio_embedder"):
            lin(f"adaln_single.emb.{n}.linear_1", s, 256)
            lin(f"adaln_single.emb.{n}.linear_2", s, s)
    lin("adaln_single.linear", 6 * D, D)
    lin("caption_projection.linear_1", D, cfg.caption_channels)
    lin("caption_projection.linear_2", D, D)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        sh[p + "scale_shift_table"] = (6, D)
        for a, kdim in (("attn1", D), ("attn2", cfg.cross_attention_dim)):
            lin(p + a + ".to_q", D, D)
            lin(p + a + ".to_k", D, kdim)
            lin(p + a + ".to_v", D, kdim)
            lin(p + a + ".to_out.0", D, D)
        lin(p + "ff.net.0.proj", 4 * D, D)
        lin(p + "ff.net.2", D, 4 * D)
    sh["scale_shift_table"] = (2, D)
    lin("proj_out", cfg.patch_size * cfg.patch_size * cfg.out_channels, D)
    return sh


def init_pixart_params(cfg: PixArtConfig, seed: int = 0, std: float = 0.02, dtype=torch.float32) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    D = cfg.inner_dim
    out = {}
    for name, shape in pixart_param_shapes(cfg).items():
        if name.endswith("scale_shift_table"):
            t = torch.randn(shape, generator=g) / D ** 0.5      # nn.Parameter(randn / dim**0.5), transformer.py:336
        elif name.endswith(".bias"):
            t = 0.5 * std * torch.randn(shape, generator=g)
        else:
            t = std * torch.randn(shape, generator=g)
        out[name] = t.to(dtype)
    return out


def lora_target_names(cfg: PixArtConfig, targets=PIXART_LORA_TARGETS):
    return [f"transformer_blocks.{i}.{a}.{n}" for i in range(cfg.num_layers) for a in ("attn1", "attn2") for n in targets]


def init_lora_params(cfg: PixArtConfig, rank: int, seed: int = 1, b_std: float = 0.02, dtype=torch.float32) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    D = cfg.inner_dim
    out = {}
    for n in lora_target_names(cfg):
        k_in = cfg.cross_attention_dim if (".attn2.to_k" in n or ".attn2.to_v" in n) else D
        bound = 1.0 / math.sqrt(k_in)
        out[n + ".lora_A.weight"] = ((torch.rand((rank, k_in), generator=g) * 2 - 1) * bound).to(dtype)
        out[n + ".lora_B.weight"] = (b_std * torch.randn((D, rank), generator=g)).to(dtype)
    return out


# --------------------------------------------------------------------------------------------------
def sincos_pos_embed_2d(dim: int, grid_h: int, grid_w: int, base_size: int, interpolation_scale: float) -> Tensor:
    """diffusers get_2d_sincos_pos_embed(embed_dim, (grid_h, grid_w), base_size=, interpolation_scale=) -> [h*w, dim] fp32.
    Token order is row-major (row, col); the FIRST half of the channels encodes the column coordinate and the
    second half the row coordinate (np.meshgrid(grid_w, grid_h) puts w first), each half = [sin | cos]."""
    gh = torch.arange(grid_h, dtype=torch.float32) / (grid_h / base_size) / interpolation_scale
    gw = torch.arange(grid_w, dtype=torch.float32) / (grid_w / base_size) / interpolation_scale
    col = gw[None, :].expand(grid_h, grid_w).reshape(-1)
    row = gh[:, None].expand(grid_h, grid_w).reshape(-1)

    def one_d(d, pos):
        omega = torch.arange(d // 2, dtype=torch.float64) / (d / 2.0)
        omega = 1.0 / 10000 ** omega
        out = pos.double()[:, None] * omega[None, :]
        return torch.cat([torch.sin(out), torch.cos(out)], dim=1)

    return torch.cat([one_d(dim // 2, col), one_d(dim // 2, row)], dim=1).float()


def patch_embed(P, cfg: PixArtConfig, latents: Tensor) -> Tensor:
    """diffusers PatchEmbed.forward (pos_embed_max_size=None): conv -> flatten -> + sincos table for this grid."""
    x = F.conv2d(latents, P["pos_embed.proj.weight"], P["pos_embed.proj.bias"], stride=cfg.patch_size)
    B, D, h, w = x.shape
    x = x.flatten(2).transpose(1, 2)
    pos = sincos_pos_embed_2d(D, h, w, cfg.sample_size // cfg.patch_size, cfg.interp)
    return x + pos[None].to(x.dtype)


def _ts_embedder(P, name, x):
    return linear(F.silu(linear(x, P, name + ".linear_1")), P, name + ".linear_2")


def adaln_single(P, cfg: PixArtConfig, timestep: Tensor, resolution: Tensor, aspect_ratio: Tensor):
    """AdaLayerNormSingle.forward == pixart/transformer.py:803-853 for 1-D timesteps -> (modulation [B,6D], embedded [B,D])."""
    B = timestep.shape[0]
    emb = _ts_embedder(P, "adaln_single.emb.timestep_embedder", timestep_embedding(timestep))
    if cfg.additional_conditions:
        r = _ts_embedder(P, "adaln_single.emb.resolution_embedder", timestep_embedding(resolution.flatten().float())).reshape(B, -1)
        a = _ts_embedder(P, "adaln_single.emb.aspect_ratio_embedder", timestep_embedding(aspect_ratio.flatten().float())).reshape(B, -1)
        emb = emb + torch.cat([r, a], dim=1)
    return linear(F.silu(emb), P, "adaln_single.linear"), emb


def attention(P, cfg: PixArtConfig, prefix: str, x: Tensor, ctx: Tensor, bias: Optional[Tensor], lora, lora_scale) -> Tensor:
    """diffusers Attention + AttnProcessor2_0: to_q(x), to_k/to_v(ctx), SDPA with additive key bias [B,1,1,Sk], to_out[0]."""
    H, hd = cfg.num_attention_heads, cfg.attention_head_dim
    B, S, _ = x.shape
    q = linear(x, P, prefix + ".to_q", lora, lora_scale).view(B, S, H, hd).transpose(1, 2)
    k = linear(ctx, P, prefix + ".to_k", lora, lora_scale).view(B, -1, H, hd).transpose(1, 2)
    v = linear(ctx, P, prefix + ".to_v", lora, lora_scale).view(B, -1, H, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    if bias is not None:
        s = s + bias[:, None, :, :]
    o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, S, H * hd)
    return linear(o, P, prefix + ".to_out.0", lora, lora_scale)


def pixart_block(P, cfg, i, x, ctx, ctx_bias, t6, lora, lora_scale):
    p = f"transformer_blocks.{i}"
    B, _, D = x.shape
    mod = P[p + ".scale_shift_table"][None] + t6.reshape(B, 6, D)
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
    n = layer_norm_noaffine(x) * (1 + scale_msa) + shift_msa
    x = gate_msa * attention(P, cfg, p + ".attn1", n, n, None, lora, lora_scale) + x
    x = attention(P, cfg, p + ".attn2", x, ctx, ctx_bias, lora, lora_scale) + x
    n = layer_norm_noaffine(x) * (1 + scale_mlp) + shift_mlp
    ff = linear(F.gelu(linear(n, P, p + ".ff.net.0.proj"), approximate="tanh"), P, p + ".ff.net.2")
    return gate_mlp * ff + x


def pixart_forward(P, cfg: PixArtConfig, latents: Tensor, encoder_hidden_states: Tensor, timestep: Tensor,
                   encoder_attention_mask: Optional[Tensor], resolution: Tensor, aspect_ratio: Tensor,
                   lora: Optional[Dict[str, Tensor]] = None, lora_scale: float = 1.0) -> Tensor:
    """-> [B, out_channels, H, W] (reference pixart/transformer.py:556-787)."""
    B, _, Hh, Ww = latents.shape
    ps = cfg.patch_size
    h, w = Hh // ps, Ww // ps
    bias = None
    if encoder_attention_mask is not None:  # :562-564
        bias = ((1 - encoder_attention_mask.to(latents.dtype)) * -10000.0).unsqueeze(1)
    x = patch_embed(P, cfg, latents)
    t6, emb = adaln_single(P, cfg, timestep, resolution, aspect_ratio)
    ctx = linear(F.gelu(linear(encoder_hidden_states, P, "caption_projection.linear_1"), approximate="tanh"), P,
                 "caption_projection.linear_2")
    for i in range(cfg.num_layers):
        x = pixart_block(P, cfg, i, x, ctx, bias, t6, lora, lora_scale)
    shift, scale = (P["scale_shift_table"][None] + emb[:, None]).chunk(2, dim=1)
    x = layer_norm_noaffine(x) * (1 + scale) + shift
    x = linear(x, P, "proj_out")
    x = x.reshape(-1, h, w, ps, ps, cfg.out_channels)
    return torch.einsum("nhwpqc->nchpwq", x).reshape(-1, cfg.out_channels, h * ps, w * ps)


def added_cond_fallback(latent_h: int, latent_w: int, B: int):
    """pixart/model.py:360-379 with no `resolution` in the batch (always, quirk Q4)."""
    return (torch.tensor([[latent_h, latent_w]]).expand(B, -1), torch.tensor([[float(latent_h / latent_w)]]).expand(B, -1))


def pixart_model_predict(P, cfg, noisy_latents, timesteps, prompt_embeds, attention_mask, lora=None, lora_scale=1.0):
    """PixartSigma._model_predict_single: first half of the channels (learned-sigma half dropped), pixart/model.py:313."""
    B, _, Hh, Ww = noisy_latents.shape
    res, ar = added_cond_fallback(Hh, Ww, B)
    out = pixart_forward(P, cfg, noisy_latents, prompt_embeds, timesteps, attention_mask, res, ar, lora, lora_scale)
    return out.chunk(2, dim=1)[0]


def ddpm_add_noise(alphas_cumprod: Tensor, latents: Tensor, noise: Tensor, timesteps: Tensor) -> Tensor:
    """DDPMScheduler.add_noise on fp32 inputs (common.py:5998-6002)."""
    a = alphas_cumprod[timesteps] ** 0.5
    b = (1 - alphas_cumprod[timesteps]) ** 0.5
    return a.view(-1, 1, 1, 1) * latents.float() + b.view(-1, 1, 1, 1) * noise.float()


def eps_loss(pred: Tensor, noise: Tensor, weights: Optional[Tensor] = None) -> Tensor:
    """common.py:6376-6398 + 6426-6429: mse(pred.float(), noise.float()) [* min-SNR weight per sample] -> mean."""
    l = F.mse_loss(pred.float(), noise.float(), reduction="none")
    if weights is not None:
        l = l * weights.view(-1, 1, 1, 1)
    return l.mean(dim=[1, 2, 3]).mean()
