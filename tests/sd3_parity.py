"""SD3 model-level parity harness: simpletuner_b200 SD3 (CUDA, bf16) vs the fp32 CPU oracle."""
from __future__ import annotations

import torch

from oracle import flux_oracle as FO
from oracle import sd3_oracle as O
from tests.flux_parity import GRAD_COS, LOSS_RTOL, PRED_COS  # same stated tolerances


def small_config(layers=3, heads=4, hd=64, joint=192, pooled=96, dual=(0,), qk_norm="rms_norm"):
    D = heads * hd
    return O.SD3Config(sample_size=32, num_layers=layers, attention_head_dim=hd, num_attention_heads=heads,
                       joint_attention_dim=joint, caption_projection_dim=D, pooled_projection_dim=pooled,
                       pos_embed_max_size=48, dual_attention_layers=tuple(dual), qk_norm=qk_norm)


def build_cuda_model(cfg, P, lora, rank=16, device="cuda"):
    from simpletuner_b200.flux.model import default_config
    from simpletuner_b200.sd3.model import SD3
    from simpletuner_b200.sd3.transformer import SD3Transformer2DModel

    m = SD3Transformer2DModel(sample_size=cfg.sample_size, in_channels=cfg.in_channels, num_layers=cfg.num_layers,
                              attention_head_dim=cfg.attention_head_dim, num_attention_heads=cfg.num_attention_heads,
                              joint_attention_dim=cfg.joint_attention_dim, caption_projection_dim=cfg.caption_projection_dim,
                              pooled_projection_dim=cfg.pooled_projection_dim, out_channels=cfg.out_channels,
                              pos_embed_max_size=cfg.pos_embed_max_size, dual_attention_layers=cfg.dual_attention_layers,
                              qk_norm=cfg.qk_norm)
    sd = {k: (v.float() if k == "pos_embed.pos_embed" else v.bfloat16()) for k, v in P.items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    m.to(device)
    w = SD3(default_config(lora_rank=rank), transformer=m, device=torch.device(device))
    if lora is not None:
        w.add_lora_adapter()
        with torch.no_grad():
            for name, lin in m.lora_linears().items():
                lin.lora_A["default"].weight.copy_(lora[name + ".lora_A.weight"].bfloat16())
                lin.lora_B["default"].weight.copy_(lora[name + ".lora_B.weight"].bfloat16())
    return w


def run_parity(cfg=None, B=2, Hh=16, Ww=24, S_txt=77, rank=16, seed=0, device="cuda"):
    cfg = cfg or small_config()
    P = {k: v.bfloat16().float() for k, v in O.init_sd3_params(cfg, seed=seed).items()}
    L = {k: v.bfloat16().float() for k, v in O.init_lora_params(cfg, rank, seed=seed + 1, b_std=0.02).items()}
    g = torch.Generator().manual_seed(seed + 2)
    batch = {"latent_batch": torch.randn(B, 16, Hh, Ww, generator=g).bfloat16(),
             "prompt_embeds": torch.randn(B, S_txt, cfg.joint_attention_dim, generator=g).bfloat16(),
             "add_text_embeds": torch.randn(B, cfg.pooled_projection_dim, generator=g).bfloat16()}
    w = build_cuda_model(cfg, P, L, rank, device)
    torch.manual_seed(1234)
    torch.cuda.manual_seed(1234)
    prepared = w.prepare_batch({k: v.clone() for k, v in batch.items()}, {"global_step": 0})
    out = w.model_predict(prepared)
    loss = w.loss(prepared, out)
    loss.backward()
    torch.cuda.synchronize()
    lat = prepared["latents"].float().cpu()
    noise = prepared["noise"].float().cpu()
    sig = prepared["sigmas"].flatten().float().cpu()
    Lg = {k: v.clone().requires_grad_(True) for k, v in L.items()}
    noisy_ref = FO.flow_noisy_latents(lat.bfloat16(), noise.bfloat16(), sig).float()
    pred_ref = O.sd3_model_predict(P, cfg, noisy_ref, sig * 1000.0, batch["prompt_embeds"].float(), batch["add_text_embeds"].float(), Lg, 1.0)
    loss_ref = FO.flow_loss(pred_ref, FO.flow_target(lat.bfloat16(), noise.bfloat16()))
    loss_ref.backward()
    pred = w.unpacked_prediction(out).detach().float().cpu()
    res = {"noisy_bit_exact": bool(torch.equal(prepared["noisy_latents"].cpu(), noisy_ref.bfloat16())),
           "loss": float(loss.item()), "loss_ref": float(loss_ref.item()),
           "loss_rel_err": abs(float(loss.item()) - float(loss_ref.item())) / abs(float(loss_ref.item())),
           "pred_cos": float(torch.nn.functional.cosine_similarity(pred.flatten(), pred_ref.detach().flatten(), dim=0))}
    cos_min, worst = 1.0, None
    for name, lin in w._denoiser().lora_linears().items():
        for which, p in (("lora_A", lin.lora_A["default"].weight), ("lora_B", lin.lora_B["default"].weight)):
            gref = Lg[f"{name}.{which}.weight"].grad
            assert p.grad is not None, f"no grad for {name}.{which}"
            c = float(torch.nn.functional.cosine_similarity(p.grad.float().cpu().flatten(), gref.flatten(), dim=0))
            if c < cos_min:
                cos_min, worst = c, f"{name}.{which}"
    res.update({"grad_cos_min": cos_min, "grad_worst": worst, "n_lora_tensors": 2 * len(w._denoiser().lora_linears())})
    return res
