"""GPU (-m gpu): step parity at BASELINE GEOMETRY (VERDICT r1 "what's weak" #1).  The toy-width cases in
test_{flux,sd3,pixart}_parity_gpu.py exercise every code path; these run the same prepare_batch -> model_predict -> loss ->
backward comparison against the fp32 CPU oracle at the real widths, head counts and sequence lengths, with the depth cut to
what the CPU oracle finishes in about a minute per case on the GPU box's host cores:

  * Flux.1-dev:  D = 3072 = 24 x 128, T5 width 4096, pooled 768, latent 128 x 128 (S_img = 4096) + 512 text tokens,
                 1 double + 1 single block, B = 1                      (BASELINE configs[1] geometry)
  * SD3.5-medium: D = 1536 = 24 x 64, dual attention + QK-norm, latent 64 x 64 (1024 tokens) + 231 text tokens, 2 blocks
                 (one dual, one context_pre_only), B = 2               (BASELINE configs[2] geometry)
  * PixArt-Sigma: D = 1152 = 16 x 72, caption width 4096, 300 caption tokens with random-length masks,
                 latent 128 x 128 (S = 4096, 2 blocks) and 192 x 192 (S = 9216, 1 block)   (configs[4] buckets)
Same stated tolerances as the toy-width tests (tests/flux_parity.py)."""
import pytest

from tests import flux_parity as FP

pytestmark = pytest.mark.gpu


def _assert(tag, res):
    FP.record(tag, res)
    print(f"[fullwidth] {tag}", res)
    assert res.get("noisy_bit_exact", True), res
    assert res["loss_rel_err"] <= FP.LOSS_RTOL, res
    assert res["pred_cos"] >= FP.PRED_COS, res
    assert res["grad_cos_min"] >= FP.GRAD_COS, res


def test_flux_dev_width_one_double_one_single_block():
    from oracle import flux_oracle as O
    cfg = O.FluxConfig(in_channels=64, num_layers=1, num_single_layers=1, attention_head_dim=128, num_attention_heads=24,
                       joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
    _assert("flux_dev_D3072_S4608", FP.run_parity(cfg=cfg, B=1, Hh=128, Ww=128, S_txt=512, rank=16, seed=11))


def test_flux_dev_width_portrait_bucket():
    """A non-square aspect bucket at full width (832 x 1216 px -> latent 152 x 104 -> S_img = 3952, not a tile multiple)."""
    from oracle import flux_oracle as O
    cfg = O.FluxConfig(in_channels=64, num_layers=1, num_single_layers=1, attention_head_dim=128, num_attention_heads=24,
                       joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=True, axes_dims_rope=(16, 56, 56))
    _assert("flux_dev_D3072_S3952+512", FP.run_parity(cfg=cfg, B=1, Hh=152, Ww=104, S_txt=512, rank=16, seed=12))


def test_sd35_medium_width():
    from oracle import sd3_oracle as O
    from tests import sd3_parity as SP
    cfg = O.SD3Config(sample_size=128, num_layers=2, attention_head_dim=64, num_attention_heads=24, joint_attention_dim=4096,
                      caption_projection_dim=1536, pooled_projection_dim=2048, pos_embed_max_size=384,
                      dual_attention_layers=(0,), qk_norm="rms_norm")
    _assert("sd35_medium_D1536_S1024+231", SP.run_parity(cfg=cfg, B=2, Hh=64, Ww=64, S_txt=231, rank=16, seed=13))


@pytest.mark.parametrize("hw,layers", [(128, 2), (192, 1)])
def test_pixart_sigma_width(hw, layers):
    from oracle import pixart_oracle as O
    from tests import pixart_parity as PP
    cfg = O.PixArtConfig(num_attention_heads=16, attention_head_dim=72, num_layers=layers, cross_attention_dim=1152,
                         caption_channels=4096, sample_size=128)
    res = PP.run_parity(cfg, B=1, Hh=hw, Ww=hw, S_txt=300, rank=32, seed=14, mask_mode="prefix")
    FP.record(f"pixart_sigma_D1152_S{(hw // 2) ** 2}", res)
    print("[fullwidth] pixart", hw, res)
    PP.check(res)
