"""CPU: the VAE oracle's structure against independent torch modules, and the host-side AutoencoderKL mirror."""
import torch
import torch.nn.functional as F

from oracle import vae_oracle as O


def _cfg():
    return O.VaeConfig(block_out_channels=(32, 64, 64, 64), layers_per_block=1, latent_channels=4)


def test_shapes_and_downsampling():
    cfg = _cfg()
    P = O.init_vae_params(cfg)
    x = torch.randn(2, 3, 32, 48)
    m = O.vae_encode_moments(P, cfg, x)
    assert m.shape == (2, 8, 4, 6)
    eps = torch.randn(2, 4, 4, 6)
    z = O.vae_cache_latents(P, cfg, x, eps)
    mean, logvar = m.chunk(2, dim=1)
    ref = ((mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * eps) - cfg.shift_factor) * cfg.scaling_factor
    assert torch.allclose(z, ref, atol=1e-6)


def test_sdxl_style_scaling_has_no_shift():
    cfg = _cfg()
    cfg.shift_factor = None
    z = torch.randn(1, 4, 2, 2)
    assert torch.equal(O.scale_latents(z, cfg), z * cfg.scaling_factor)


def test_resnet_matches_module_composition():
    """The restated ResnetBlock2D equals an nn.Module composition with the same weights."""
    cfg = _cfg()
    P = O.init_vae_params(cfg, seed=3)
    p = "encoder.down_blocks.1.resnets.0."
    x = torch.randn(1, 32, 8, 8)
    gn1 = torch.nn.GroupNorm(32, 32, eps=1e-6); gn1.weight.data, gn1.bias.data = P[p + "norm1.weight"], P[p + "norm1.bias"]
    gn2 = torch.nn.GroupNorm(32, 64, eps=1e-6); gn2.weight.data, gn2.bias.data = P[p + "norm2.weight"], P[p + "norm2.bias"]
    h = F.conv2d(F.silu(gn1(x)), P[p + "conv1.weight"], P[p + "conv1.bias"], padding=1)
    h = F.conv2d(F.silu(gn2(h)), P[p + "conv2.weight"], P[p + "conv2.bias"], padding=1)
    ref = F.conv2d(x, P[p + "conv_shortcut.weight"], P[p + "conv_shortcut.bias"]) + h
    assert torch.allclose(O._resnet(x, P, p, 32), ref, atol=1e-5)


def test_host_mirror_state_dict_names_match_diffusers_layout():
    from simpletuner_b200.vae.autoencoder import AutoencoderKL
    for quant in (False, True):
        cfg = O.VaeConfig(use_quant_conv=quant)
        m = AutoencoderKL(use_quant_conv=quant)
        sd = m.state_dict()
        sh = O.vae_encoder_param_shapes(cfg)
        assert set(sd) == set(sh)
        for k, v in sd.items():
            assert tuple(v.shape) == sh[k], k


def test_tap_major_weight_relayout():
    from simpletuner_b200.vae.autoencoder import _Conv
    c = _Conv(4, 6, 3, torch.float32)
    c.weight.data = torch.randn(6, 4, 3, 3)
    w9 = c.w9()
    assert w9.shape == (6, 36)
    # K index = (dy*3 + dx) * C_in + ci
    assert torch.equal(w9[2, (1 * 3 + 2) * 4 + 3], c.weight[2, 3, 1, 2])
