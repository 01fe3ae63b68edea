"""CPU: self-consistency of the oracle restatement (fp32) on a tiny Flux configuration."""
import torch

from oracle import flux_oracle as O

CFG = O.FluxConfig(in_channels=16, num_layers=1, num_single_layers=1, attention_head_dim=16, num_attention_heads=2,
                   joint_attention_dim=24, pooled_projection_dim=8, guidance_embeds=True, axes_dims_rope=(4, 6, 6))


def _batch(B=2, C=4, H=4, W=6, S_txt=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"latents": torch.randn(B, C, H, W, generator=g), "noise": torch.randn(B, C, H, W, generator=g),
            "sigmas": torch.tensor([0.3, 0.8][:B]), "prompt_embeds": torch.randn(B, S_txt, CFG.joint_attention_dim, generator=g),
            "pooled": torch.randn(B, CFG.pooled_projection_dim, generator=g)}


def test_param_inventory_counts():
    full = O.FluxConfig()
    n = sum(torch.Size(s).numel() for s in O.flux_param_shapes(full).values())
    assert 11.8e9 < n < 12.0e9  # "12B parameters" (reference flux/model.py:52)
    names = O.lora_target_names(full)
    assert len(names) == 19 * 8 + 38 * 3 == 266
    assert sum(16 * (3072 + 3072) for _ in names) == 26_148_864  # ~26.1 M trainable params at r=16


def test_lora_zero_b_is_identity_and_grads_flow():
    P = O.init_flux_params(CFG, seed=0)
    b = _batch()
    loss0, pred0 = O.flux_train_step_loss(P, CFG, b)
    L = O.init_lora_params(CFG, rank=4, b_std=0.0)
    loss1, pred1 = O.flux_train_step_loss(P, CFG, b, lora=L)
    assert torch.equal(pred0, pred1)
    L = {k: v.clone().requires_grad_(True) for k, v in O.init_lora_params(CFG, rank=4, b_std=0.05).items()}
    loss2, _ = O.flux_train_step_loss(P, CFG, b, lora=L)
    loss2.backward()
    assert all(v.grad is not None and torch.isfinite(v.grad).all() for v in L.values())
    assert any(v.grad.abs().max() > 0 for v in L.values())


def test_single_block_matches_manual_concat_path():
    # proj_out(cat[attn, mlp]) == attn @ W[:, :D].T + mlp @ W[:, D:].T + b  (the K-segment identity the GEMM uses)
    P = O.init_flux_params(CFG, seed=1)
    D = CFG.inner_dim
    g = torch.Generator().manual_seed(3)
    a, m = torch.randn(2, 7, D, generator=g), torch.randn(2, 7, 4 * D, generator=g)
    W, bvec = P["single_transformer_blocks.0.proj_out.weight"], P["single_transformer_blocks.0.proj_out.bias"]
    ref = torch.nn.functional.linear(torch.cat([a, m], 2), W, bvec)
    alt = a @ W[:, :D].t() + m @ W[:, D:].t() + bvec
    torch.testing.assert_close(ref, alt, atol=1e-5, rtol=1e-5)


def test_bf16_emulation_tracks_fp32():
    P = O.init_flux_params(CFG, seed=0)
    b = _batch()
    loss32, _ = O.flux_train_step_loss(P, CFG, b)
    Pb = {k: v.bfloat16() for k, v in P.items()}
    bb = {k: (v.bfloat16() if v.dtype == torch.float32 and k != "sigmas" else v) for k, v in b.items()}
    loss16, _ = O.flux_train_step_loss(Pb, CFG, bb)
    assert abs(loss16.item() - loss32.item()) / loss32.item() < 3e-2
