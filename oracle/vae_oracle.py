"""CPU oracle for the VAE latent-encode path — TEST INFRASTRUCTURE ONLY (see flux_oracle.py header).

Restates what reference simpletuner/helpers/caching/vae.py:1293-1355 runs per batch:
    vae.encode(samples)                      (ModelFoundation.encode_with_vae, common.py:2766-2772)
    .latent_dist.sample()                    (vae.py:1337)
    (z - shift_factor) * scaling_factor      (VaeLatentScalingMixin.scale_vae_latents_for_cache,
                                              foundation_mixins.py:68-81; SDXL-style: z * scaling_factor)
`vae` is diffusers' `AutoencoderKL` (third party, diffusers>=0.36.0, not vendored, not installed here); its
encoder is restated from its published structure: conv_in 3x3 -> 4 x DownEncoderBlock2D (2 ResnetBlock2D each:
GroupNorm32 -> SiLU -> conv3x3 -> GroupNorm32 -> SiLU -> conv3x3 (+ 1x1 conv shortcut when channels change);
blocks 0..2 end with a stride-2 3x3 conv after F.pad(0,1,0,1)) -> UNetMidBlock2D (resnet, single-head
attention over H*W tokens with GroupNorm32 + residual, resnet) -> GroupNorm32 -> SiLU -> conv_out 3x3 ->
optional 1x1 quant_conv -> DiagonalGaussianDistribution(mean | logvar clamped to [-30, 20]).
PARITY STATUS: **parity unpinned** (no golden for encode outputs in the reference's tests; tests/test_vae.py
there pins cache file naming only).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class VaeConfig:
    """AutoencoderKL config fields used by the encoder (Flux / SD3 VAE defaults)."""
    in_channels: int = 3
    latent_channels: int = 16
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    use_quant_conv: bool = False
    scaling_factor: float = 0.3611
    shift_factor: Optional[float] = 0.1159
    mid_block_add_attention: bool = True


def vae_encoder_param_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, ...]]:
    sh: Dict[str, Tuple[int, ...]] = {}

    def conv(name, o, i, k):
        sh[name + ".weight"] = (o, i, k, k)
        sh[name + ".bias"] = (o,)

    def norm(name, c):
        sh[name + ".weight"] = (c,)
        sh[name + ".bias"] = (c,)

    def resnet(p, i, o):
        norm(p + "norm1", i)
        conv(p + "conv1", o, i, 3)
        norm(p + "norm2", o)
        conv(p + "conv2", o, o, 3)
        if i != o:
            conv(p + "conv_shortcut", o, i, 1)

    ch = cfg.block_out_channels
    conv("encoder.conv_in", ch[0], cfg.in_channels, 3)
    prev = ch[0]
    for bi, c in enumerate(ch):
        for li in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{bi}.resnets.{li}.", prev if li == 0 else c, c)
        prev = c
        if bi != len(ch) - 1:
            conv(f"encoder.down_blocks.{bi}.downsamplers.0.conv", c, c, 3)
    c = ch[-1]
    resnet("encoder.mid_block.resnets.0.", c, c)
    if cfg.mid_block_add_attention:
        norm("encoder.mid_block.attentions.0.group_norm", c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            sh[f"encoder.mid_block.attentions.0.{n}.weight"] = (c, c)
            sh[f"encoder.mid_block.attentions.0.{n}.bias"] = (c,)
    resnet("encoder.mid_block.resnets.1.", c, c)
    norm("encoder.conv_norm_out", c)
    conv("encoder.conv_out", 2 * cfg.latent_channels, c, 3)
    if cfg.use_quant_conv:
        conv("quant_conv", 2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)
    return sh


def init_vae_params(cfg: VaeConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in vae_encoder_param_shapes(cfg).items():
        if "norm" in name and name.endswith(".weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
        out[name] = t.to(dtype)
    return out


def _gn_silu(x, P, name, groups, eps=1e-6, silu=True):
    y = F.group_norm(x, groups, P[name + ".weight"], P[name + ".bias"], eps)
    return F.silu(y) if silu else y


def _resnet(x, P, p, groups):
    h = F.conv2d(_gn_silu(x, P, p + "norm1", groups), P[p + "conv1.weight"], P[p + "conv1.bias"], padding=1)
    h = F.conv2d(_gn_silu(h, P, p + "norm2", groups), P[p + "conv2.weight"], P[p + "conv2.bias"], padding=1)
    if (p + "conv_shortcut.weight") in P:
        x = F.conv2d(x, P[p + "conv_shortcut.weight"], P[p + "conv_shortcut.bias"])
    return x + h  # output_scale_factor = 1


def vae_encode_moments(P: Dict[str, Tensor], cfg: VaeConfig, x: Tensor) -> Tensor:
    """AutoencoderKL.encode(x).latent_dist.parameters: [B, 2*latent, H/8, W/8]."""
    G = cfg.norm_num_groups
    h = F.conv2d(x, P["encoder.conv_in.weight"], P["encoder.conv_in.bias"], padding=1)
    for bi in range(len(cfg.block_out_channels)):
        for li in range(cfg.layers_per_block):
            h = _resnet(h, P, f"encoder.down_blocks.{bi}.resnets.{li}.", G)
        name = f"encoder.down_blocks.{bi}.downsamplers.0.conv"
        if (name + ".weight") in P:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), P[name + ".weight"], P[name + ".bias"], stride=2)
    h = _resnet(h, P, "encoder.mid_block.resnets.0.", G)
    if cfg.mid_block_add_attention:
        a = "encoder.mid_block.attentions.0."
        B, C, Hh, Ww = h.shape
        t = _gn_silu(h, P, a + "group_norm", G, silu=False).view(B, C, Hh * Ww).transpose(1, 2)
        q = F.linear(t, P[a + "to_q.weight"], P[a + "to_q.bias"])
        k = F.linear(t, P[a + "to_k.weight"], P[a + "to_k.bias"])
        v = F.linear(t, P[a + "to_v.weight"], P[a + "to_v.bias"])
        s = (q.float() @ k.float().transpose(1, 2)) * (C ** -0.5)
        o = (torch.softmax(s, dim=-1) @ v.float()).to(t.dtype)
        o = F.linear(o, P[a + "to_out.0.weight"], P[a + "to_out.0.bias"])
        h = h + o.transpose(1, 2).reshape(B, C, Hh, Ww)
    h = _resnet(h, P, "encoder.mid_block.resnets.1.", G)
    h = _gn_silu(h, P, "encoder.conv_norm_out", G)
    h = F.conv2d(h, P["encoder.conv_out.weight"], P["encoder.conv_out.bias"], padding=1)
    if cfg.use_quant_conv:
        h = F.conv2d(h, P["quant_conv.weight"], P["quant_conv.bias"])
    return h


def gaussian_sample(moments: Tensor, eps: Tensor) -> Tensor:
    """DiagonalGaussianDistribution.sample with externally supplied standard-normal `eps`."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
    return mean + std * eps


def scale_latents(z: Tensor, cfg: VaeConfig) -> Tensor:
    """reference foundation_mixins.py:69-81."""
    if cfg.shift_factor is not None:
        return (z - cfg.shift_factor) * cfg.scaling_factor
    return z * cfg.scaling_factor


def vae_cache_latents(P, cfg: VaeConfig, pixels: Tensor, eps: Tensor) -> Tensor:
    """encode -> sample -> scale, as reference caching/vae.py:1311, 1337, 1355."""
    return scale_latents(gaussian_sample(vae_encode_moments(P, cfg, pixels), eps), cfg)
