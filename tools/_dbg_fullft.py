import torch, sys
sys.path.insert(0, "/root/repo")
from oracle import flux_oracle as FO, sd3_oracle as O
from tests import sd3_parity as SP
cfg = SP.small_config(layers=2, dual=(), qk_norm=None)
B,Hh,Ww,S_txt,seed=2,16,16,64,4
P = {k: v.bfloat16().float() for k, v in O.init_sd3_params(cfg, seed=seed).items()}
g = torch.Generator().manual_seed(seed + 2)
batch = {"latent_batch": torch.randn(B, 16, Hh, Ww, generator=g).bfloat16(),
         "prompt_embeds": torch.randn(B, S_txt, cfg.joint_attention_dim, generator=g).bfloat16(),
         "add_text_embeds": torch.randn(B, cfg.pooled_projection_dim, generator=g).bfloat16()}
w = SP.build_cuda_model(cfg, P, None); den = w._denoiser(); den.enable_full_finetune()
torch.manual_seed(1234); torch.cuda.manual_seed(1234)
prepared = w.prepare_batch({k: v.clone() for k, v in batch.items()}, {"global_step": 0})
out = w.model_predict(prepared); loss = w.loss(prepared, out); loss.backward(); torch.cuda.synchronize()
lat, noise = prepared["latents"].float().cpu(), prepared["noise"].float().cpu(); sig = prepared["sigmas"].flatten().float().cpu()
Pg = {k: (v.clone().requires_grad_(True) if k != "pos_embed.pos_embed" else v) for k, v in P.items()}
noisy = FO.flow_noisy_latents(lat.bfloat16(), noise.bfloat16(), sig).float()
pred_ref = O.sd3_model_predict(Pg, cfg, noisy, sig * 1000.0, batch["prompt_embeds"].float(), batch["add_text_embeds"].float(), None, 1.0)
FO.flow_loss(pred_ref, FO.flow_target(lat.bfloat16(), noise.bfloat16())).backward()
cos = torch.nn.functional.cosine_similarity
for name, p in den.named_parameters():
    gref = Pg[name].grad
    if gref is None: print(name, "no ref"); continue
    if p.grad is None: print(name, "MISSING"); continue
    a=p.grad.float().cpu().flatten(); b=gref.flatten()
    print(f"{name:60s} cos {float(cos(a,b,dim=0)):+.5f}  norm {float(a.norm()):.3e} ref {float(b.norm()):.3e}")
