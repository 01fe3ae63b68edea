"""GPU (-m gpu): AutoencoderKL encode -> sample -> scale on libstb200 vs the fp32 CPU oracle (oracle/vae_oracle.py)."""
import pytest

from tests import vae_parity as VP

pytestmark = pytest.mark.gpu


def test_vae_flux_style():
    res = VP.run_parity()
    VP.check(res)
    assert res["shape"] == (2, 16, 8, 12) and res["sample_shape"] == (2, 16, 8, 12)


def test_vae_sdxl_style_quant_conv_no_shift():
    cfg = VP.small_config(quant=True, latent=4)
    cfg.shift_factor, cfg.scaling_factor = None, 0.13025
    VP.check(VP.run_parity(cfg, B=1, H=128, W=64, seed=3))


def test_vae_no_mid_attention():
    VP.check(VP.run_parity(VP.small_config(attn=False), B=1, H=64, W=64, seed=5))


def test_vae_cpu_input_fails_loudly():
    import torch
    from oracle import vae_oracle as O
    from simpletuner_b200._lib import StbError
    cfg = VP.small_config()
    vae = VP.build_cuda_vae(cfg, O.init_vae_params(cfg))
    with pytest.raises(StbError):
        vae.encode(torch.zeros(1, 3, 64, 64))
