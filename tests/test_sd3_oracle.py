"""CPU: SD3 oracle self-consistency and the shared-schedule assumptions the CUDA path relies on."""
import torch

from oracle import sd3_oracle as O

CFG = O.SD3Config(sample_size=16, num_layers=2, attention_head_dim=8, num_attention_heads=2, joint_attention_dim=12,
                  caption_projection_dim=16, pooled_projection_dim=6, pos_embed_max_size=12, dual_attention_layers=(0,))


def _inputs(B=2, H=8, W=12, S=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 16, H, W, generator=g), torch.randn(B, S, 12, generator=g), torch.randn(B, 6, generator=g),
            torch.tensor([250.0, 731.0][:B]))


def test_param_inventory_sd35_medium():
    n = sum(torch.Size(s).numel() for k, s in O.sd3_param_shapes(O.SD3Config()).items() if k != "pos_embed.pos_embed")
    assert 2.2e9 < n < 2.6e9  # SD3.5-medium ~2.5 B
    names = O.lora_target_names(O.SD3Config())
    assert len(names) == 24 * 4 + 13 * 4  # attn.{q,k,v,out} in every block + attn2.* in the 13 dual layers


def test_patch_embed_conv_equals_gemm_on_packed_latents():
    # the CUDA path runs PatchEmbed as a K=64 GEMM over Flux-packed (c, dy, dx) features
    from oracle.flux_oracle import pack_latents
    P = O.init_sd3_params(CFG, seed=0)
    x, *_ = _inputs()
    conv = torch.nn.functional.conv2d(x, P["pos_embed.proj.weight"], P["pos_embed.proj.bias"], stride=2).flatten(2).transpose(1, 2)
    gemm = pack_latents(x, 2, 16, 8, 12) @ P["pos_embed.proj.weight"].reshape(16, 64).t() + P["pos_embed.proj.bias"]
    torch.testing.assert_close(conv, gemm, atol=1e-5, rtol=1e-5)


def test_joint_attention_is_order_invariant():
    # [image, text] (reference) vs [text, image] (CUDA joint buffer): identical up to fp32 summation order
    P = O.init_sd3_params(CFG, seed=1)
    g = torch.Generator().manual_seed(3)
    x, enc = torch.randn(2, 24, 16, generator=g), torch.randn(2, 5, 16, generator=g)
    o1, e1 = O.joint_attention(P, CFG, "transformer_blocks.0.attn.", x, enc, None, 1.0, False)
    # emulate the swapped order by hand
    import oracle.flux_oracle as FO
    H, hd = 2, 8
    heads = lambda t: t.view(2, -1, H, hd).transpose(1, 2)
    p = "transformer_blocks.0.attn."
    q = O.rms_norm(heads(FO.linear(x, P, p + "to_q")), P[p + "norm_q.weight"], 1e-6)
    k = O.rms_norm(heads(FO.linear(x, P, p + "to_k")), P[p + "norm_k.weight"], 1e-6)
    v = heads(FO.linear(x, P, p + "to_v"))
    eq = O.rms_norm(heads(FO.linear(enc, P, p + "add_q_proj")), P[p + "norm_added_q.weight"], 1e-6)
    ek = O.rms_norm(heads(FO.linear(enc, P, p + "add_k_proj")), P[p + "norm_added_k.weight"], 1e-6)
    ev = heads(FO.linear(enc, P, p + "add_v_proj"))
    o = O.sdpa(torch.cat([eq, q], 2), torch.cat([ek, k], 2), torch.cat([ev, v], 2)).transpose(1, 2).reshape(2, -1, 16)
    torch.testing.assert_close(FO.linear(o[:, 5:], P, p + "to_out.0"), o1, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(FO.linear(o[:, :5], P, p + "to_add_out"), e1, atol=1e-5, rtol=1e-5)


def test_forward_shapes_and_lora_identity():
    P = O.init_sd3_params(CFG, seed=2)
    x, enc, pooled, t = _inputs()
    y0 = O.sd3_forward(P, CFG, x, enc, pooled, t)
    assert y0.shape == x.shape
    L = O.init_lora_params(CFG, rank=4, b_std=0.0)
    assert torch.equal(O.sd3_forward(P, CFG, x, enc, pooled, t, lora=L), y0)
    L = {k: v.clone().requires_grad_(True) for k, v in O.init_lora_params(CFG, rank=4, b_std=0.05).items()}
    O.sd3_forward(P, CFG, x, enc, pooled, t, lora=L).square().mean().backward()
    assert all(v.grad is not None for v in L.values())


def test_timestep_bf16_rounding_quirk_q8():
    # raw timesteps are cast to bf16 before the sinusoidal embedding: 731 -> 732
    t = torch.tensor([731.0])
    assert float(t.to(torch.bfloat16)) == 732.0
