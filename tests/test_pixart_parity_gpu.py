"""GPU (-m gpu): PixArt-Sigma epsilon-prediction LoRA step on libstb200 vs the fp32 CPU oracle."""
import pytest

from tests import pixart_parity as PP

pytestmark = pytest.mark.gpu


def test_pixart_hd72_prefix_mask():
    res = PP.run_parity()
    PP.check(res)
    assert res["n_lora_tensors"] == 2 * 8 * 2


def test_pixart_arbitrary_mask_min_snr():
    PP.check(PP.run_parity(B=3, Hh=24, Ww=16, S_txt=150, seed=4, mask_mode="holes", snr_gamma=5.0))


def test_pixart_hd40_no_size_conditioning_rank4():
    cfg = PP.small_config(layers=2, heads=8, hd=40, sample_size=64)     # sample_size != 128 -> no resolution / AR embedders
    PP.check(PP.run_parity(cfg, B=1, Hh=32, Ww=32, S_txt=77, rank=4, seed=7))
