"""GPU (-m gpu): PEFT `lora_dropout` (reference default 0.1, field_registry/sections/lora.py:130-137) on the fused LoRA path.

The CUDA path draws its masks from a counter-based generator keyed by (seed, stream, element index) and regenerates them
in backward (csrc/elementwise.cuh); torch's Philox stream is not reproduced, so parity is pinned by REPLAYING the masks the
CUDA path used into the fp32 oracle (`oracle.flux_oracle.DROPOUT_MASKS`)."""
import pytest
import torch

from tests import flux_parity as FP

pytestmark = pytest.mark.gpu


def test_dropout_kernels_keep_rate_independence_and_backward_consistency():
    from simpletuner_b200 import ops

    B, S, K, p = 2, 300, 512, 0.1
    x = torch.randn(B, S + 4, K + 8, device="cuda").bfloat16()[:, 2:S + 2, :K]          # a strided view
    xm = ops.dropout_expand(x, 3, p, seed=1234, stream0=7)
    assert xm.shape == (3, B, S, K)
    keep = xm != 0
    for m in range(3):
        rate = float(keep[m].float().mean())
        assert abs(rate - (1 - p)) < 5e-3, rate
        kept = xm[m][keep[m]].float()
        want = (x.float() / (1 - p)).bfloat16()[keep[m]].float()      # x * 1/(1-p) rounded once to bf16, like torch dropout
        assert torch.equal(kept, want)
    # independent streams: joint keep rate ~ (1-p)^2, not (1-p)
    joint = float((keep[0] & keep[1]).float().mean())
    assert abs(joint - (1 - p) ** 2) < 5e-3, joint
    # reproducible, and stream-addressed: member 1 of stream0=7 is member 0 of stream0=8
    assert torch.equal(ops.dropout_expand(x, 3, p, 1234, 7), xm)
    assert torch.equal(ops.dropout_expand(x, 1, p, 1234, 8)[0], xm[1])
    assert not torch.equal(ops.dropout_expand(x, 1, p, 1235, 7)[0], xm[0])
    # backward through the same masks: dx += sum_m keep_m / (1 - p) * d_m
    d = torch.randn(3, B, S, K, device="cuda").bfloat16()
    dx0 = torch.randn(B, S, K, device="cuda").bfloat16()
    dx = dx0.clone()
    ops.dropout_accum_(dx, d, p, 1234, 7)
    ref = dx0.float() + sum(keep[m].float() * d[m].float() / (1 - p) for m in range(3))
    assert torch.allclose(dx.float(), ref, rtol=1e-2, atol=1e-2)
    # p = 0 is the identity
    assert torch.equal(ops.dropout_expand(x, 2, 0.0, 1, 0)[1], x.contiguous())


def _masks_for_flux(den, cfg, B, S_img, S_txt, p):
    """Materialise the masks of the last forward as {oracle linear name: mask / (1 - p)} (CPU fp32)."""
    from simpletuner_b200 import ops

    D = cfg.num_attention_heads * cfg.attention_head_dim
    masks = {}

    def grab(shape_S, drop, members, K=D):
        ones = torch.ones(B, shape_S, K, device="cuda", dtype=torch.bfloat16)
        return ops.dropout_expand(ones, members, drop.p, drop.seed, drop.stream).float().cpu()

    adapted = set(den.lora_linears())

    for i, blk in enumerate(den.transformer_blocks):
        drop = blk._lora_drop
        assert drop is not None and drop.p == p
        pre = f"transformer_blocks.{i}.attn."
        m = grab(S_img, drop.at(0), 3)
        masks[pre + "to_q"], masks[pre + "to_k"], masks[pre + "to_v"] = m[0], m[1], m[2]
        masks[pre + "to_out.0"] = grab(S_img, drop.at(6), 1)[0]
        m = grab(S_txt, drop.at(8), 3)
        masks[pre + "add_q_proj"], masks[pre + "add_k_proj"], masks[pre + "add_v_proj"] = m[0], m[1], m[2]
        masks[pre + "to_add_out"] = grab(S_txt, drop.at(14), 1)[0]
        blk_name = f"transformer_blocks.{i}."
        for nm, S_, off, K in (("ff.net.0.proj", S_img, 3, D), ("ff.net.2", S_img, 4, 4 * D),
                               ("ff_context.net.0.proj", S_txt, 11, D), ("ff_context.net.2", S_txt, 12, 4 * D)):
            if blk_name + nm in adapted:
                masks[blk_name + nm] = grab(S_, drop.at(off), 1, K)[0]
    for j, blk in enumerate(den.single_transformer_blocks):
        drop = blk._lora_drop
        pre = f"single_transformer_blocks.{j}.attn."
        m = grab(S_img + S_txt, drop, 3)
        masks[pre + "to_q"], masks[pre + "to_k"], masks[pre + "to_v"] = m[0], m[1], m[2]
        blk_name = f"single_transformer_blocks.{j}."
        if blk_name + "proj_mlp" in adapted:
            masks[blk_name + "proj_mlp"] = grab(S_img + S_txt, drop.at(3), 1)[0]
        if blk_name + "proj_out" in adapted:
            masks[blk_name + "proj_out"] = grab(S_img + S_txt, drop.at(4), 1, 5 * D)[0]
    if "proj_out" in adapted:
        masks["proj_out"] = grab(S_img, den.proj_out._lora_drop, 1)[0]
    if "x_embedder" in adapted:
        masks["x_embedder"] = grab(S_img, den.x_embedder._lora_drop, 1, cfg.in_channels)[0]
    return masks


@pytest.mark.parametrize("target", ["all", "all+ffs+embedder"])
def test_flux_step_parity_with_lora_dropout_replayed_into_the_oracle(target):
    from oracle import flux_oracle as O
    from simpletuner_b200.flux.transformer import FLUX_LORA_TARGETS

    p = 0.1
    cfg = FP.small_config(layers=1, single=1)
    rank, B, Hh, Ww, S_txt = 16, 2, 16, 16, 64
    P = {k: v.bfloat16().float() for k, v in O.init_flux_params(cfg, seed=0).items()}
    L = {k: v.bfloat16().float() for k, v in O.init_lora_params(cfg, rank, seed=1, b_std=0.02,
                                                                targets=tuple(FLUX_LORA_TARGETS[target])).items()}
    batch = FP.make_batch(B, Hh, Ww, S_txt, cfg, seed=2)
    w = FP.build_cuda_model(cfg, P, None, rank, target=target)
    w.config.lora_dropout = p
    w.add_lora_adapter()
    den = w._denoiser()
    assert den.peft_config["default"].lora_dropout == p
    with torch.no_grad():
        for name, lin in den.lora_linears().items():
            lin.lora_A["default"].weight.copy_(L[name + ".lora_A.weight"].bfloat16())
            lin.lora_B["default"].weight.copy_(L[name + ".lora_B.weight"].bfloat16())
    den.train()
    torch.manual_seed(77); torch.cuda.manual_seed(77)
    prepared = w.prepare_batch({k: v.clone() for k, v in batch.items()}, {"global_step": 0})
    out = w.model_predict(prepared)
    loss = w.loss(prepared, out)
    loss.backward()
    torch.cuda.synchronize()
    masks = _masks_for_flux(den, cfg, B, (Hh // 2) * (Ww // 2), S_txt, p)
    lat, noise = prepared["latents"].float().cpu(), prepared["noise"].float().cpu()
    sig = prepared["sigmas"].flatten().float().cpu()
    Lg = {k: v.clone().requires_grad_(True) for k, v in L.items()}
    noisy = O.flow_noisy_latents(lat.bfloat16(), noise.bfloat16(), sig).float()
    O.DROPOUT_MASKS = masks
    try:
        pred_ref = O.flux_model_predict(P, cfg, noisy, sig * 1000.0, batch["prompt_embeds"].float(), batch["add_text_embeds"].float(), 1.0, Lg, 1.0)
        loss_ref = O.flow_loss(pred_ref, O.flow_target(lat.bfloat16(), noise.bfloat16()))
        loss_ref.backward()
        # the same oracle WITHOUT the masks must differ measurably (the test would otherwise be vacuous)
        O.DROPOUT_MASKS = None
        Ln = {k: v.clone().requires_grad_(True) for k, v in L.items()}
        pred_nodrop = O.flux_model_predict(P, cfg, noisy, sig * 1000.0, batch["prompt_embeds"].float(), batch["add_text_embeds"].float(), 1.0, Ln, 1.0)
        O.flow_loss(pred_nodrop, O.flow_target(lat.bfloat16(), noise.bfloat16())).backward()
        pred_nodrop = pred_nodrop.detach()
    finally:
        O.DROPOUT_MASKS = None
    cos = torch.nn.functional.cosine_similarity
    pred = out["model_prediction"].detach().float().cpu()
    res = {"loss_rel_err": abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)),
           "pred_cos": float(cos(pred.flatten(), pred_ref.detach().flatten(), dim=0)),
           "dropout_effect": float((pred_ref.detach() - pred_nodrop).abs().max())}
    gmin, gmin_nomask = 1.0, 1.0
    for name, lin in den.lora_linears().items():
        for which, prm in (("lora_A", lin.lora_A["default"].weight), ("lora_B", lin.lora_B["default"].weight)):
            g = prm.grad.float().cpu().flatten()
            gmin = min(gmin, float(cos(g, Lg[f"{name}.{which}.weight"].grad.flatten(), dim=0)))
            gmin_nomask = min(gmin_nomask, float(cos(g, Ln[f"{name}.{which}.weight"].grad.flatten(), dim=0)))
    res["grad_cos_min"] = gmin
    res["grad_cos_min_vs_unmasked_oracle"] = gmin_nomask
    FP.record(f"flux_lora_dropout_0.1[{target}]", res)
    print("[lora-dropout]", res)
    # the masks matter: against the oracle WITHOUT them the LoRA gradients are visibly off (cos ~ sqrt(1 - p))
    assert res["dropout_effect"] > 1e-4 and res["grad_cos_min_vs_unmasked_oracle"] < 0.99, res
    assert res["loss_rel_err"] <= FP.LOSS_RTOL and res["pred_cos"] >= FP.PRED_COS and res["grad_cos_min"] >= FP.GRAD_COS, res
    # eval mode (like nn.Dropout): no mask, bit-identical to the p = 0 model
    den.eval()
    with torch.no_grad():
        prep2 = dict(prepared); prep2["timesteps"] = prepared["timesteps"] * 1000.0
        a = w.model_predict(prep2)["model_prediction"]
        den._lora_dropout_p = 0.0
        den.train()
        prep3 = dict(prepared); prep3["timesteps"] = prepared["timesteps"] * 1000.0
        b = w.model_predict(prep3)["model_prediction"]
    assert torch.equal(a, b)


def test_sd3_and_pixart_steps_run_with_lora_dropout():
    """The shared LoRA helpers carry dropout for the other families too (dual attention and cross attention groups)."""
    from tests import pixart_parity as PP
    from tests import sd3_parity as SP
    from oracle import pixart_oracle as PO
    from oracle import sd3_oracle as SO

    cfg = SP.small_config()
    P = {k: v.bfloat16().float() for k, v in SO.init_sd3_params(cfg, seed=0).items()}
    w = SP.build_cuda_model(cfg, P, None, 16)
    w.config.lora_dropout = 0.1
    w.add_lora_adapter()
    g = torch.Generator().manual_seed(3)
    batch = {"latent_batch": torch.randn(2, 16, 16, 24, generator=g).bfloat16(),
             "prompt_embeds": torch.randn(2, 77, cfg.joint_attention_dim, generator=g).bfloat16(),
             "add_text_embeds": torch.randn(2, cfg.pooled_projection_dim, generator=g).bfloat16()}
    with torch.no_grad():
        for lin in w._denoiser().lora_linears().values():
            lin.lora_B["default"].weight.normal_(0, 0.02)
    prep = w.prepare_batch(batch, {})
    loss = w.loss(prep, w.model_predict(prep))
    loss.backward()
    assert torch.isfinite(loss) and all(torch.isfinite(p.grad.float()).all() for p in w._denoiser().trainable_parameters())

    pcfg = PP.small_config()
    PPm = {k: v.bfloat16().float() for k, v in PO.init_pixart_params(pcfg, seed=0).items()}
    pw = PP.build_cuda_model(pcfg, PPm, None, 16, lora_dropout=0.1)
    pw.add_lora_adapter()
    with torch.no_grad():
        for lin in pw._denoiser().lora_linears().values():
            lin.lora_B["default"].weight.normal_(0, 0.02)
    mask = torch.ones(2, 40); mask[0, 20:] = 0
    pbatch = {"latent_batch": torch.randn(2, 4, 16, 24, generator=g).bfloat16(),
              "prompt_embeds": torch.randn(2, 40, pcfg.caption_channels, generator=g).bfloat16(), "encoder_attention_mask": mask}
    pprep = pw.prepare_batch(pbatch, {})
    ploss = pw.loss(pprep, pw.model_predict(pprep))
    ploss.backward()
    assert torch.isfinite(ploss) and all(torch.isfinite(p.grad.float()).all() for p in pw._denoiser().trainable_parameters())
