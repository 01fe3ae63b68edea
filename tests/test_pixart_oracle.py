"""CPU: PixArt oracle structure, host-side helpers of the PixArt path (padding layouts, sincos table, config)."""
import torch

from oracle import pixart_oracle as O


def _cfg():
    return O.PixArtConfig(num_attention_heads=3, attention_head_dim=8, num_layers=1, cross_attention_dim=24,
                          caption_channels=12, sample_size=128)


def test_masked_keys_do_not_influence_prediction():
    cfg = _cfg()
    P = O.init_pixart_params(cfg, std=0.2)
    x = torch.randn(1, 4, 8, 8)
    enc = torch.randn(1, 6, 12)
    mask = torch.tensor([[1, 1, 1, 0, 0, 0.0]])
    a = O.pixart_model_predict(P, cfg, x, torch.tensor([300]), enc, mask)
    enc2 = enc.clone()
    enc2[:, 3:] = torch.randn(1, 3, 12) * 5
    b = O.pixart_model_predict(P, cfg, x, torch.tensor([300]), enc2, mask)
    assert a.shape == (1, 4, 8, 8) and torch.allclose(a, b, atol=1e-6)
    c = O.pixart_model_predict(P, cfg, x, torch.tensor([300]), enc2, torch.ones(1, 6))
    assert not torch.allclose(a, c, atol=1e-4)


def test_size_conditioning_only_for_sample_size_128():
    assert O.PixArtConfig(sample_size=128).additional_conditions and not O.PixArtConfig(sample_size=64).additional_conditions
    assert "adaln_single.emb.resolution_embedder.linear_1.weight" not in O.pixart_param_shapes(O.PixArtConfig(sample_size=64))
    assert O.PixArtConfig(sample_size=128).interp == 2 and O.PixArtConfig(sample_size=32).interp == 1


def test_sincos_table_matches_numpy_formulation_and_host_mirror():
    import numpy as np
    from simpletuner_b200.pixart.transformer import sincos_pos_embed_2d
    D, gh, gw, base, interp = 16, 3, 5, 4, 2.0
    grid_h = np.arange(gh, dtype=np.float32) / (gh / base) / interp
    grid_w = np.arange(gw, dtype=np.float32) / (gw / base) / interp
    grid = np.stack(np.meshgrid(grid_w, grid_h), axis=0).reshape([2, 1, gw, gh])

    def one_d(d, pos):
        omega = 1.0 / 10000 ** (np.arange(d // 2, dtype=np.float64) / (d / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    ref = np.concatenate([one_d(D // 2, grid[0]), one_d(D // 2, grid[1])], axis=1).astype(np.float32)
    got = O.sincos_pos_embed_2d(D, gh, gw, base, interp)
    assert np.allclose(got.numpy(), ref, atol=1e-6)
    assert torch.equal(sincos_pos_embed_2d(D, gh, gw, base, interp, "cpu"), got)


def test_head_padding_layouts_roundtrip():
    from simpletuner_b200.pixart.transformer import _pad_cols, _pad_rows
    H, hd, hdp = 3, 8, 16
    w = torch.randn(H * hd, 5)
    wp = _pad_rows(w, H, hd, hdp)
    idx = (torch.arange(H)[:, None] * hdp + torch.arange(hd)[None, :]).reshape(-1)
    assert wp.shape == (H * hdp, 5) and torch.equal(wp[idx], w) and int((wp != 0).sum()) == int((w != 0).sum())
    o = torch.randn(7, H * hd)
    wo = torch.randn(4, H * hd)
    op = torch.zeros(7, H * hdp)
    op[:, idx] = o
    assert torch.allclose(op @ _pad_cols(wo, H, hd, hdp).t(), o @ wo.t(), atol=1e-5)


def test_pack_lora_scatter_matches_dense_delta():
    from simpletuner_b200.flux.blocks import pack_lora
    H, hd, hdp, K, r = 2, 8, 16, 24, 4
    idx = (torch.arange(H)[:, None] * hdp + torch.arange(hd)[None, :]).reshape(-1)
    A, Bm = torch.randn(r, K), torch.randn(H * hd, r)
    pk = pack_lora([(A, Bm)], H * hdp, K, 2.0, "cpu", dtype=torch.float32, row_index=idx)
    delta = pk.b_ext @ pk.a_stack                       # [H*hdp, K]
    assert torch.allclose(delta[idx], 2.0 * Bm @ A, atol=1e-5)
    keep = torch.ones(H * hdp, dtype=torch.bool)
    keep[idx] = False
    assert float(delta[keep].abs().max()) == 0.0
    pk2 = pack_lora([(torch.randn(r, H * hd), torch.randn(6, r))], 6, H * hdp, 1.0, "cpu", dtype=torch.float32, col_index=idx)
    assert pk2.a_stack.shape == (8, H * hdp) and float(pk2.a_stack[:, keep].abs().max()) == 0.0


def test_eps_step_pieces():
    from simpletuner_b200.training.noise import make_ddpm_schedule
    ns = make_ddpm_schedule(1000, 0.0001, 0.02, "linear")
    x, e = torch.randn(2, 4, 4, 4), torch.randn(2, 4, 4, 4)
    t = torch.tensor([0, 999])
    n = O.ddpm_add_noise(ns.alphas_cumprod, x, e, t)
    assert torch.allclose(n[0], x[0], atol=6e-2) and torch.allclose(n[1], e[1], atol=6e-2)
    assert torch.allclose(O.eps_loss(e, e), torch.tensor(0.0))
    w = torch.tensor([2.0, 0.0])
    assert torch.allclose(O.eps_loss(x, e, w), ((x[0] - e[0]) ** 2).mean())


def test_sincos_table_matches_the_mae_implementation_shipped_with_transformers():
    """diffusers' get_2d_sincos_pos_embed descends from MAE's; with base_size == grid size and interpolation_scale 1 the
    two coincide.  `transformers` (installed here) ships MAE's version: an implementation we did not write."""
    import numpy as np
    import pytest
    mae = pytest.importorskip("transformers.models.vit_mae.modeling_vit_mae")
    ref = mae.get_2d_sincos_pos_embed(32, 6, add_cls_token=False)
    ref = ref.numpy() if hasattr(ref, "numpy") else np.asarray(ref)
    got = O.sincos_pos_embed_2d(32, 6, 6, base_size=6, interpolation_scale=1.0).numpy()
    assert got.shape == ref.shape == (36, 32) and np.allclose(got, ref, atol=1e-6)
