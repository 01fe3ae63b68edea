"""VAE latent-encode parity harness: simpletuner_b200 AutoencoderKL encoder (CUDA, bf16) vs the fp32 CPU oracle."""
from __future__ import annotations

import torch

from oracle import vae_oracle as O

LATENT_COS = 0.999   # stated tolerance: cosine of cached latents vs fp32 oracle (bf16 activations through ~25 convs)
LATENT_RMS = 0.06    # rms(latents - ref) / rms(ref)


def small_config(quant=False, attn=True, latent=16):
    return O.VaeConfig(block_out_channels=(64, 128, 256, 256), use_quant_conv=quant, mid_block_add_attention=attn,
                       latent_channels=latent)


def build_cuda_vae(cfg, P, device="cuda"):
    from simpletuner_b200.vae.autoencoder import AutoencoderKL
    m = AutoencoderKL(in_channels=cfg.in_channels, latent_channels=cfg.latent_channels,
                      block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                      norm_num_groups=cfg.norm_num_groups, use_quant_conv=cfg.use_quant_conv,
                      scaling_factor=cfg.scaling_factor, shift_factor=cfg.shift_factor,
                      mid_block_add_attention=cfg.mid_block_add_attention)
    missing, unexpected = m.load_state_dict({k: v.bfloat16() for k, v in P.items()}, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m.to(device)


def run_parity(cfg=None, B=2, H=64, W=96, seed=0, device="cuda"):
    cfg = cfg or small_config()
    P = {k: v.bfloat16().float() for k, v in O.init_vae_params(cfg, seed=seed).items()}
    g = torch.Generator().manual_seed(seed + 1)
    pixels = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).bfloat16()
    eps = torch.randn(B, cfg.latent_channels, H // 8, W // 8, generator=g).bfloat16()
    vae = build_cuda_vae(cfg, P, device)
    dist = vae.encode(pixels.to(device)).latent_dist
    moments = dist.parameters.float().cpu()
    latents = vae.encode_scaled(pixels.to(device), eps.to(device)).float().cpu()
    torch.cuda.synchronize()
    mref = O.vae_encode_moments(P, cfg, pixels.float())
    lref = O.vae_cache_latents(P, cfg, pixels.float(), eps.float())
    cos = torch.nn.functional.cosine_similarity
    return {"moments_cos": float(cos(moments.flatten(), mref.flatten(), dim=0)),
            "latents_cos": float(cos(latents.flatten(), lref.flatten(), dim=0)),
            "latents_rms_rel": float((latents - lref).pow(2).mean().sqrt() / lref.pow(2).mean().sqrt()),
            "finite": bool(torch.isfinite(latents).all()), "shape": tuple(latents.shape),
            "sample_shape": tuple(dist.sample().shape)}


def check(res):
    assert res["finite"], res
    assert res["latents_cos"] >= LATENT_COS, res
    assert res["latents_rms_rel"] <= LATENT_RMS, res
