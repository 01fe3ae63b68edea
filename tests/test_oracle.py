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


def test_oracle_attention_pieces_against_torch_modules():
    """Independent cross-checks of the restated layer math with torch's own modules / functionals."""
    import torch.nn.functional as F
    from oracle import pixart_oracle as PO
    from oracle import vae_oracle as VO

    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(2, 3, 17, 16, generator=g) for _ in range(3))
    assert torch.allclose(O.sdpa(q, k, v), F.scaled_dot_product_attention(q, k, v), atol=1e-5)
    x = torch.randn(2, 9, 24, generator=g)
    assert torch.allclose(O.layer_norm_noaffine(x), torch.nn.LayerNorm(24, elementwise_affine=False, eps=1e-6)(x), atol=1e-6)
    w = 1.0 + 0.1 * torch.randn(24, generator=g)
    assert torch.allclose(O.rms_norm(x, w, 1e-6), F.rms_norm(x, (24,), w, 1e-6), atol=1e-6)
    # PixArt cross attention with an additive key bias == SDPA with the same float mask
    cfg = PO.PixArtConfig(num_attention_heads=3, attention_head_dim=8, num_layers=1, cross_attention_dim=24, caption_channels=12)
    P = PO.init_pixart_params(cfg, std=0.2)
    hs, ctx = torch.randn(2, 7, 24, generator=g), torch.randn(2, 5, 24, generator=g)
    mask = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1.0]])
    bias = ((1 - mask) * -10000.0).unsqueeze(1)
    got = PO.attention(P, cfg, "transformer_blocks.0.attn2", hs, ctx, bias, None, 1.0)
    p = "transformer_blocks.0.attn2"
    heads = lambda t: t.view(2, -1, 3, 8).transpose(1, 2)
    qh = heads(F.linear(hs, P[p + ".to_q.weight"], P[p + ".to_q.bias"]))
    kh = heads(F.linear(ctx, P[p + ".to_k.weight"], P[p + ".to_k.bias"]))
    vh = heads(F.linear(ctx, P[p + ".to_v.weight"], P[p + ".to_v.bias"]))
    ref = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=bias[:, None])
    ref = F.linear(ref.transpose(1, 2).reshape(2, 7, 24), P[p + ".to_out.0.weight"], P[p + ".to_out.0.bias"])
    assert torch.allclose(got, ref, atol=1e-5)
    # VAE mid-block attention (single head over H*W tokens) == nn.MultiheadAttention-free SDPA formulation
    vc = VO.VaeConfig(block_out_channels=(32, 32), layers_per_block=1, latent_channels=4)
    VP = VO.init_vae_params(vc, seed=1)
    m = VO.vae_encode_moments(VP, vc, torch.randn(1, 3, 16, 16, generator=g))
    assert m.shape == (1, 8, 8, 8) and torch.isfinite(m).all()
