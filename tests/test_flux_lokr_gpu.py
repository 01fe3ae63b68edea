"""GPU (-m gpu): Flux training step with a LyCORIS LoKr adapter (BASELINE config 4) against the fp32 oracle restatement
(oracle/lokr_oracle.py; third-party algorithm, parity unpinned — see that file's header)."""
import pytest
import torch

from tests import flux_parity as FP

pytestmark = pytest.mark.gpu

LYCORIS_CFG = {"algo": "lokr", "multiplier": 1.0, "linear_dim": 10000, "linear_alpha": 1, "factor": 10,
               "apply_preset": {"target_module": ["Attention", "FeedForward"],
                                "module_algo_map": {"Attention": {"factor": 10}, "FeedForward": {"factor": 4}}}}


def _factor_of(name):
    return 4 if (".ff." in name or ".ff_context." in name) else 10


def _run(cfg, lyc, B=2, Hh=16, Ww=16, S_txt=64, seed=0):
    from oracle import flux_oracle as O
    from oracle import lokr_oracle as LO

    P = {k: v.bfloat16().float() for k, v in O.init_flux_params(cfg, seed=seed).items()}
    shapes = O.flux_param_shapes(cfg)
    targets = [k[:-7] for k in shapes if k.endswith(".weight") and len(shapes[k]) == 2 and (".attn." in k or ".ff" in k)]
    K = {k: v.bfloat16().float() for k, v in LO.init_lokr_params({n: shapes[n + ".weight"] for n in targets}, lyc["linear_dim"],
                                                                 _factor_of, seed=seed + 1, w2_std=0.02).items()}
    batch = FP.make_batch(B, Hh, Ww, S_txt, cfg, seed=seed + 2)
    w = FP.build_cuda_model(cfg, P, None, 16)
    w.config.lora_type = "lycoris"
    net = w.add_lycoris_adapter(dict(lyc))
    net.to("cuda")
    assert len(net.loras) == len(targets)
    with torch.no_grad():
        for lora in net.loras:
            oname = next(t for t in targets if "lycoris_" + t.replace(".", "_") == lora.lora_name)
            for pn, prm in lora.named_parameters():
                prm.copy_(K[f"{oname}.{pn}"].bfloat16())
    w._denoiser().invalidate_plans()
    torch.manual_seed(1234); torch.cuda.manual_seed(1234)
    prepared = w.prepare_batch({k: v.clone() for k, v in batch.items()}, {"global_step": 0})
    out = w.model_predict(prepared)
    loss = w.loss(prepared, out)
    loss.backward()
    torch.cuda.synchronize()
    lat, noise = prepared["latents"].float().cpu(), prepared["noise"].float().cpu()
    sig = prepared["sigmas"].flatten().float().cpu()
    Kg = {k: v.clone().requires_grad_(True) for k, v in K.items()}
    noisy = O.flow_noisy_latents(lat.bfloat16(), noise.bfloat16(), sig).float()
    O.LOKR = {"linear_dim": lyc["linear_dim"], "linear_alpha": lyc["linear_alpha"], "multiplier": 1.0}
    try:
        pred_ref = O.flux_model_predict(P, cfg, noisy, sig * 1000.0, batch["prompt_embeds"].float(), batch["add_text_embeds"].float(), 1.0, Kg, 1.0)
        pred_base = O.flux_model_predict(P, cfg, noisy, sig * 1000.0, batch["prompt_embeds"].float(), batch["add_text_embeds"].float(), 1.0, None, 1.0)
    finally:
        O.LOKR = {"linear_dim": 10000, "linear_alpha": 1, "multiplier": 1.0}
    loss_ref = O.flow_loss(pred_ref, O.flow_target(lat.bfloat16(), noise.bfloat16()))
    loss_ref.backward()
    cos = torch.nn.functional.cosine_similarity
    pred = w.unpacked_prediction(out).float().cpu()
    res = {"loss_rel_err": abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)),
           "pred_cos": float(cos(pred.flatten(), pred_ref.detach().flatten(), dim=0)),
           "adapter_effect": float((pred_ref.detach() - pred_base.detach()).abs().max()), "n_adapted": len(net.loras)}
    rows = []
    for lora in net.loras:
        oname = next(t for t in targets if "lycoris_" + t.replace(".", "_") == lora.lora_name)
        for pn, prm in lora.named_parameters():
            assert prm.grad is not None, (oname, pn)
            gref = Kg[f"{oname}.{pn}"].grad
            c = float(cos(prm.grad.float().cpu().flatten(), gref.flatten(), dim=0))
            rows.append((c, f"{oname}.{pn}", float(gref.norm()), float(prm.grad.float().norm())))
    rows.sort()
    gmax = max(r[2] for r in rows)
    # gradients eight orders below the largest one are below the bf16 noise floor of the backward pass (the text stream's q / k
    # projections at toy size, as in tests/test_sd3_fullft_gpu.py); they are reported, not asserted
    live = [r for r in rows if r[2] >= 1e-4 * gmax]
    res["grad_cos_min"], res["grad_worst"] = live[0][0], live[0][1]
    res["grad_rows_worst"] = [(round(c, 5), n, float(f"{a:.3g}"), float(f"{b:.3g}")) for c, n, a, b in rows[:6]]
    res["n_below_noise_floor"] = len(rows) - len(live)
    return res, w, net


@pytest.mark.parametrize("linear_dim", [10000, 8])
def test_flux_lokr_step_parity(linear_dim):
    lyc = dict(LYCORIS_CFG, linear_dim=linear_dim, linear_alpha=(1 if linear_dim == 10000 else 4))
    res, w, net = _run(FP.small_config(layers=2, single=2), lyc)
    FP.record(f"flux_lokr[linear_dim={linear_dim}]", res)
    print("[lokr]", res)
    assert net.loras[0].full_matrix == (linear_dim == 10000)
    assert res["adapter_effect"] > 1e-3, res                       # the adapter does change the prediction
    assert res["loss_rel_err"] <= FP.LOSS_RTOL and res["pred_cos"] >= FP.PRED_COS and res["grad_cos_min"] >= FP.GRAD_COS, res


def test_lokr_train_steps_follow_the_factors_and_round_trip(tmp_path):
    """After an optimizer step the rebuilt projection weights embed the UPDATED factors (TrainStep -> after_optimizer_step);
    set_multiplier(0) restores the base model; save_weights / load_weights round-trip in LyCORIS key layout."""
    from simpletuner_b200.training.optim import AdamWBF16
    from simpletuner_b200.training.step import TrainStep
    res, w, net = _run(FP.small_config(layers=1, single=1), LYCORIS_CFG)
    den = w._denoiser()
    for p in net.parameters():
        p.grad = None
    step = TrainStep(w, AdamWBF16(list(net.parameters()), lr=1e-2, seed=0), max_grad_norm=2.0)
    cfg = FP.small_config(layers=1, single=1)
    batch = {k: v.cuda() for k, v in FP.make_batch(2, 16, 16, 64, cfg, seed=9).items()}
    lin = den.transformer_blocks[0].attn.to_q
    w_before = lin.effective_weight().clone()
    plan_before = den.transformer_blocks[0].plans()["img_attn"].w_qkv.clone()
    step(dict(batch))
    torch.cuda.synchronize()
    assert not torch.equal(lin.effective_weight(), w_before)
    D = lin.out_features
    assert torch.equal(den.transformer_blocks[0].plans()["img_attn"].w_qkv[:D], lin.effective_weight())
    assert not torch.equal(den.transformer_blocks[0].plans()["img_attn"].w_qkv, plan_before)
    net.set_multiplier(0.0)
    assert torch.equal(lin.effective_weight(), lin.weight)
    net.set_multiplier(1.0)
    f = str(tmp_path / "lokr.safetensors")
    net.save_weights(f, torch.bfloat16, {"k": "v"})
    from safetensors.torch import load_file
    sd = load_file(f)
    assert "lycoris_transformer_blocks_0_attn_to_q.lokr_w1" in sd and "lycoris_transformer_blocks_0_ff_net_0_proj.lokr_w2" in sd
    ref = {k: v.clone() for k, v in net.state_dict_lycoris().items() if not k.endswith(".alpha")}   # (alpha is saved in `dtype` too)
    with torch.no_grad():
        for p in net.parameters():
            p.zero_()
    net.load_weights(f)
    assert all(torch.equal(v, ref[k]) for k, v in net.state_dict_lycoris().items() if k in ref) and len(ref) == 2 * len(net.loras)


def test_lycoris_unsupported_options_raise():
    from simpletuner_b200 import lycoris as LY
    for bad in ({"algo": "loha"}, {"algo": "lokr", "bypass_mode": True}, {"algo": "lokr", "dropout": 0.1}, {"algo": "lokr", "weight_decompose": True}):
        with pytest.raises(NotImplementedError):
            LY.validate_lycoris_config(bad)
