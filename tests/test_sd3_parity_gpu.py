"""GPU (-m gpu): the libstb200 SD3 / SD3.5 LoRA training step against the fp32 CPU oracle."""
import pytest

from tests import flux_parity as FP
from tests import sd3_parity as SP

pytestmark = pytest.mark.gpu


def _assert(res):
    import inspect
    FP.record(inspect.stack()[1].function, res)
    assert res["noisy_bit_exact"], res
    assert res["loss_rel_err"] <= FP.LOSS_RTOL, res
    assert res["pred_cos"] >= FP.PRED_COS, res
    assert res["grad_cos_min"] >= FP.GRAD_COS, res


def test_sd35_dual_attention_qknorm_step_parity():
    # 3 joint blocks: block 0 has attn2 (dual attention), block 2 is context_pre_only; ragged 8x12 patch grid, 77 text tokens
    _assert(SP.run_parity())


def test_sd3_medium_no_qknorm_step_parity():
    _assert(SP.run_parity(cfg=SP.small_config(layers=2, dual=(), qk_norm=None), Hh=16, Ww=16, S_txt=64, seed=4))
