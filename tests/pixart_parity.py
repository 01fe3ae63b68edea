"""PixArt model-level parity harness: simpletuner_b200 PixArt-Sigma (CUDA, bf16) vs the fp32 CPU oracle."""
from __future__ import annotations

import torch

from oracle import pixart_oracle as O
from tests.flux_parity import GRAD_COS, LOSS_RTOL, PRED_COS  # same stated tolerances


def small_config(layers=2, heads=4, hd=72, caption=96, sample_size=128):
    D = heads * hd
    return O.PixArtConfig(num_attention_heads=heads, attention_head_dim=hd, num_layers=layers, cross_attention_dim=D,
                          caption_channels=caption, sample_size=sample_size)


def build_cuda_model(cfg, P, lora, rank=16, device="cuda", **cfg_over):
    from simpletuner_b200.pixart.model import PixartSigma, default_config
    from simpletuner_b200.pixart.transformer import PixArtTransformer2DModel

    m = PixArtTransformer2DModel(num_attention_heads=cfg.num_attention_heads, attention_head_dim=cfg.attention_head_dim,
                                 in_channels=cfg.in_channels, out_channels=cfg.out_channels, num_layers=cfg.num_layers,
                                 cross_attention_dim=cfg.cross_attention_dim, sample_size=cfg.sample_size,
                                 caption_channels=cfg.caption_channels, interpolation_scale=cfg.interpolation_scale,
                                 use_additional_conditions=cfg.use_additional_conditions)
    missing, unexpected = m.load_state_dict({k: v.bfloat16() for k, v in P.items()}, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    m.to(device)
    w = PixartSigma(default_config(lora_rank=rank, **cfg_over), transformer=m, device=torch.device(device))
    if lora is not None:
        w.add_lora_adapter()
        with torch.no_grad():
            for name, lin in m.lora_linears().items():
                lin.lora_A["default"].weight.copy_(lora[name + ".lora_A.weight"].bfloat16())
                lin.lora_B["default"].weight.copy_(lora[name + ".lora_B.weight"].bfloat16())
    return w


def run_parity(cfg=None, B=2, Hh=16, Ww=24, S_txt=40, rank=16, seed=0, device="cuda", mask_mode="prefix", snr_gamma=None):
    cfg = cfg or small_config()
    P = {k: v.bfloat16().float() for k, v in O.init_pixart_params(cfg, seed=seed).items()}
    L = {k: v.bfloat16().float() for k, v in O.init_lora_params(cfg, rank, seed=seed + 1).items()}
    g = torch.Generator().manual_seed(seed + 2)
    mask = torch.ones(B, S_txt)
    if mask_mode == "prefix":          # tokenizer padding on the right, a different length per sample
        for b in range(B):
            mask[b, max(1, S_txt // 3 + 7 * b):] = 0
        mask[B - 1] = 1
    elif mask_mode == "holes":         # arbitrary keep-mask, including a fully masked leading tile
        mask = (torch.rand(B, S_txt, generator=g) > 0.4).float()
        mask[0, : S_txt // 2] = 0
        mask[:, -1] = 1
    batch = {"latent_batch": torch.randn(B, 4, Hh, Ww, generator=g).bfloat16(),
             "prompt_embeds": torch.randn(B, S_txt, cfg.caption_channels, generator=g).bfloat16(),
             "encoder_attention_mask": mask}
    w = build_cuda_model(cfg, P, L, rank, device, snr_gamma=snr_gamma)
    torch.manual_seed(1234)
    torch.cuda.manual_seed(1234)
    prepared = w.prepare_batch({k: v.clone() for k, v in batch.items()}, {"global_step": 0})
    out = w.model_predict(prepared)
    loss = w.loss(prepared, out)
    loss.backward()
    torch.cuda.synchronize()
    lat, noise = prepared["latents"].float().cpu(), prepared["noise"].float().cpu()
    t = prepared["timesteps"].cpu().long()
    Lg = {k: v.clone().requires_grad_(True) for k, v in L.items()}
    ac = w.noise_schedule.alphas_cumprod
    noisy_ref = O.ddpm_add_noise(ac, lat, noise, t)
    pred_ref = O.pixart_model_predict(P, cfg, noisy_ref.bfloat16().float(), t, batch["prompt_embeds"].float(), mask, Lg, 1.0)
    wts = None
    if snr_gamma:
        snr = ac[t] / (1 - ac[t])
        wts = torch.minimum(snr, torch.full_like(snr, snr_gamma)) / snr
    loss_ref = O.eps_loss(pred_ref, noise, wts)
    loss_ref.backward()
    pred = w.unpacked_prediction(out).detach().float().cpu()
    cos = torch.nn.functional.cosine_similarity
    res = {"noisy_bit_exact": bool(torch.equal(prepared["noisy_latents"].cpu(), noisy_ref.bfloat16())),
           "timesteps": t.tolist(), "loss": float(loss.item()), "loss_ref": float(loss_ref.item()),
           "loss_rel_err": abs(float(loss.item()) - float(loss_ref.item())) / abs(float(loss_ref.item())),
           "pred_cos": float(cos(pred.flatten(), pred_ref.detach().flatten(), dim=0))}
    cos_min, worst = 1.0, None
    for name, lin in w._denoiser().lora_linears().items():
        for which, p in (("lora_A", lin.lora_A["default"].weight), ("lora_B", lin.lora_B["default"].weight)):
            gref = Lg[f"{name}.{which}.weight"].grad
            assert p.grad is not None, f"no grad for {name}.{which}"
            c = float(cos(p.grad.float().cpu().flatten(), gref.flatten(), dim=0))
            if c < cos_min:
                cos_min, worst = c, f"{name}.{which}"
    res.update({"grad_cos_min": cos_min, "grad_worst": worst, "n_lora_tensors": 2 * len(w._denoiser().lora_linears())})
    return res


def check(res):
    import inspect
    from tests.flux_parity import record
    record("pixart:" + inspect.stack()[1].function, res)
    assert res["noisy_bit_exact"], res
    assert res["loss_rel_err"] <= LOSS_RTOL, res
    assert res["pred_cos"] >= PRED_COS, res
    assert res["grad_cos_min"] >= GRAD_COS, res
