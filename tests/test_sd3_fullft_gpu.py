"""GPU (-m gpu): SD3 / SD3.5 FULL fine-tune step (model_type=full, BASELINE configs[2]) on libstb200 vs the fp32 CPU oracle:
prepare_batch -> model_predict -> loss -> backward with a gradient for EVERY parameter (weights and biases of all linears,
adaLN linears, timestep / text embedders, per-head RMSNorm weights, PatchEmbed conv, context embedder, norm_out, proj_out),
then one optimizer step through TrainStep (derived weight layouts must follow the updated weights).
Stated tolerances: loss / prediction as in tests/flux_parity.py; per-parameter gradient cosine >= 0.995 for matrices
(bf16 GEMM chains), >= 0.99 for vectors (biases, norm weights: sums of ~10^4 bf16 products)."""
import pytest
import torch

from oracle import flux_oracle as FO
from oracle import sd3_oracle as O
from tests import flux_parity as FP
from tests import sd3_parity as SP

pytestmark = pytest.mark.gpu

MAT_COS, VEC_COS = 0.995, 0.99


def _run(cfg, B=2, Hh=16, Ww=24, S_txt=77, seed=0, tag="sd3_fullft"):
    P = {k: v.bfloat16().float() for k, v in O.init_sd3_params(cfg, seed=seed).items()}
    g = torch.Generator().manual_seed(seed + 2)
    batch = {"latent_batch": torch.randn(B, 16, Hh, Ww, generator=g).bfloat16(),
             "prompt_embeds": torch.randn(B, S_txt, cfg.joint_attention_dim, generator=g).bfloat16(),
             "add_text_embeds": torch.randn(B, cfg.pooled_projection_dim, generator=g).bfloat16()}
    w = SP.build_cuda_model(cfg, P, None)
    den = w._denoiser()
    den.enable_full_finetune()
    assert all(p.requires_grad for p in den.parameters())
    torch.manual_seed(1234); torch.cuda.manual_seed(1234)
    prepared = w.prepare_batch({k: v.clone() for k, v in batch.items()}, {"global_step": 0})
    out = w.model_predict(prepared)
    loss = w.loss(prepared, out)
    loss.backward()
    torch.cuda.synchronize()
    lat, noise = prepared["latents"].float().cpu(), prepared["noise"].float().cpu()
    sig = prepared["sigmas"].flatten().float().cpu()
    Pg = {k: (v.clone().requires_grad_(True) if k != "pos_embed.pos_embed" else v) for k, v in P.items()}
    noisy = FO.flow_noisy_latents(lat.bfloat16(), noise.bfloat16(), sig).float()
    pred_ref = O.sd3_model_predict(Pg, cfg, noisy, sig * 1000.0, batch["prompt_embeds"].float(), batch["add_text_embeds"].float(), None, 1.0)
    loss_ref = FO.flow_loss(pred_ref, FO.flow_target(lat.bfloat16(), noise.bfloat16()))
    loss_ref.backward()
    cos = torch.nn.functional.cosine_similarity
    pred = out["model_prediction"].detach().float().cpu()
    res = {"loss_rel_err": abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)),
           "pred_cos": float(cos(pred.flatten(), pred_ref.detach().flatten(), dim=0))}
    worst_m, worst_v, missing = (1.0, None), (1.0, None), []
    n = 0
    gmax = max(float(v.grad.norm()) for v in Pg.values() if getattr(v, "grad", None) is not None)
    for name, p in den.named_parameters():
        gref = Pg[name].grad
        if gref is None or float(gref.norm()) < 1e-7 * gmax:
            # unused by this configuration, or at the bf16 noise floor: e.g. the text QUERY projection of the block before a
            # context_pre_only block only reaches the loss through the next block's keys / values (|g| ~ 1e-8 of the largest)
            continue
        if p.grad is None:
            missing.append(name)
            continue
        n += 1
        c = float(cos(p.grad.float().cpu().flatten(), gref.flatten(), dim=0))
        if p.dim() >= 2:
            if c < worst_m[0]:
                worst_m = (c, name)
        elif c < worst_v[0]:
            worst_v = (c, name)
    res.update({"n_param_grads": n, "grad_cos_min_matrix": worst_m[0], "worst_matrix": worst_m[1],
                "grad_cos_min_vector": worst_v[0], "worst_vector": worst_v[1], "grad_cos_min": min(worst_m[0], worst_v[0])})
    FP.record(tag, res)
    print(f"[{tag}]", res)
    assert not missing, missing
    assert res["loss_rel_err"] <= FP.LOSS_RTOL and res["pred_cos"] >= FP.PRED_COS, res
    assert res["grad_cos_min_matrix"] >= MAT_COS and res["grad_cos_min_vector"] >= VEC_COS, res
    return w, batch


def test_sd35_dual_attention_qknorm_full_finetune_parity_and_optimizer_step():
    w, batch = _run(SP.small_config(), tag="sd35_fullft_small")
    # one optimizer step through TrainStep: derived layouts (fused qkv, transposed copies) must follow the new weights
    from simpletuner_b200.training.optim import AdamWBF16
    from simpletuner_b200.training.step import TrainStep
    den = w._denoiser()
    opt = AdamWBF16(list(den.parameters()), lr=1e-3, weight_decay=0.0, seed=3)
    step = TrainStep(w, opt, max_grad_norm=2.0, grad_clip_method="value")
    torch.manual_seed(5); torch.cuda.manual_seed(5)
    l0 = float(step({k: v.clone() for k, v in batch.items()}))
    blk = den.transformer_blocks[0]
    assert blk._plans is None          # invalidated by after_optimizer_step()
    wq = blk.attn.to_q.weight.detach().clone()
    torch.manual_seed(5); torch.cuda.manual_seed(5)
    l1 = float(step({k: v.clone() for k, v in batch.items()}))
    assert torch.equal(blk.plans()["img_attn"].w_qkv[: wq.shape[0]], blk.attn.to_q.weight.detach())
    assert not torch.equal(blk.attn.to_q.weight.detach(), wq)
    assert l1 < l0                      # same batch, same noise: a step of lr 1e-3 on every weight lowers the loss


def test_sd3_medium_no_qknorm_full_finetune_parity():
    _run(SP.small_config(layers=2, dual=(), qk_norm=None), Hh=16, Ww=16, S_txt=64, seed=4, tag="sd3_fullft_noqknorm")


def test_sd35_medium_width_full_finetune_parity():
    cfg = O.SD3Config(sample_size=128, num_layers=2, attention_head_dim=64, num_attention_heads=24, joint_attention_dim=4096,
                      caption_projection_dim=1536, pooled_projection_dim=2048, pos_embed_max_size=384,
                      dual_attention_layers=(0,), qk_norm="rms_norm")
    _run(cfg, B=2, Hh=64, Ww=64, S_txt=231, seed=13, tag="sd35_medium_width_fullft")
