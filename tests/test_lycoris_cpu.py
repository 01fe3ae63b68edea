"""CPU: host logic of the LyCORIS LoKr mirror (simpletuner_b200/lycoris.py) — factorisation against the oracle restatement,
module shapes for the Flux dimensions with the reference's documented preset, the factor-gradient formulas against autograd
through torch.kron (what the reference differentiates), preset / option validation, state-dict key layout."""
import pytest
import torch

from oracle import lokr_oracle as LO
from simpletuner_b200 import lycoris as LY
from tests.test_shim_cpu import _ref_flux


def test_factorization_matches_the_oracle_and_known_flux_shapes():
    for dim in (64, 72, 128, 768, 1152, 1536, 3072, 4096, 9216, 12288, 15360, 97, 2 * 3 * 5 * 7):
        for f in (-1, 1, 2, 4, 8, 10, 16, 64):
            assert LY.factorization(dim, f) == LO.factorization(dim, f), (dim, f)
            m, n = LY.factorization(dim, f)
            assert m * n == dim and m <= n
    # the reference's documented preset on Flux.1-dev: factor 10 on attention (3072 is not divisible by 10 -> largest divisor <= 10)
    assert LY.factorization(3072, 10) == (8, 384) and LY.factorization(3072, 4) == (4, 768) and LY.factorization(12288, 4) == (4, 3072)


def test_factor_gradients_equal_autograd_through_kron():
    g = torch.Generator().manual_seed(0)
    for (a, b, c, d), scale in (((4, 6, 3, 8), 1.0), ((2, 16, 2, 24), 0.5), ((8, 3, 8, 5), 2.0)):
        w1 = torch.randn(a, c, generator=g, requires_grad=True)
        w2 = torch.randn(b, d, generator=g, requires_grad=True)
        dW = torch.randn(a * b, c * d, generator=g)
        (torch.kron(w1, w2) * scale * dW).sum().backward()          # d loss / d (W + delta W) = dW
        d1, d2 = LY.lokr_factor_grads(dW, w1.detach(), w2.detach(), scale)
        assert torch.allclose(d1, w1.grad, rtol=1e-5, atol=1e-5) and torch.allclose(d2, w2.grad, rtol=1e-5, atol=1e-5)


def test_network_surface_preset_and_key_layout():
    m = _ref_flux()
    LY.LycorisNetwork.apply_preset({"target_module": ["Attention", "FeedForward"],
                                    "module_algo_map": {"Attention": {"factor": 10}, "FeedForward": {"factor": 4}}})
    try:
        net = LY.create_lycoris(m, 1.0, 10000, 1, algo="lokr", factor=10)
        # one double block (8 attention + 4 feed-forward Linears) + one single block (q, k, v): proj_mlp / proj_out are not in
        # `Attention` / `FeedForward` modules and stay un-adapted, like in the reference
        assert len(net.loras) == 15
        names = {l.lora_name for l in net.loras}
        assert "lycoris_transformer_blocks_0_attn_to_q" in names and "lycoris_transformer_blocks_0_ff_context_net_2" in names
        assert not any("proj_mlp" in n or "proj_out" in n for n in names)
        lora = next(l for l in net.loras if l.lora_name.endswith("ff_net_0_proj"))
        assert lora.shape == ((4, 256), (4, 64)) and lora.full_matrix and lora.scale == 1.0          # 1024 x 256 at the toy width
        assert float(lora.lokr_w2.abs().sum()) == 0.0 and float(lora.lokr_w1.abs().sum()) > 0        # LyCORIS init: w2 = 0
        net.apply_to()
        lin = m.transformer_blocks[0].ff.net[0].proj
        assert lin.lokr is lora and torch.equal(lin.effective_weight(), lin.weight)                   # delta W = 0 at init
        with torch.no_grad():
            lora.lokr_w2.fill_(0.01)
        want = lin.weight.float() + torch.kron(lora.lokr_w1.float(), lora.lokr_w2.float())
        assert torch.allclose(lin.effective_weight().float(), want, atol=1e-2)
        net.set_multiplier(0.0)
        assert torch.equal(lin.effective_weight(), lin.weight)
        sd = net.state_dict_lycoris()
        assert {k.split(".")[-1] for k in sd} == {"lokr_w1", "lokr_w2", "alpha"} and len(sd) == 45
        net.restore()
        assert lin.lokr is None
        # low-rank w2 below the full-matrix threshold
        small = LY.create_lycoris(_ref_flux(), 1.0, 8, 4, algo="lokr", factor=4)
        l0 = small.loras[0]
        assert not l0.full_matrix and l0.lokr_w2_a.shape[1] == 8 and l0.scale == 0.5
    finally:
        LY.LycorisNetwork._preset = {"target_module": ["Attention", "FeedForward"], "module_algo_map": {}}
    for bad in ({"algo": "loha"}, {"bypass_mode": True}, {"rank_dropout": 0.1}, {"use_tucker": True}):
        with pytest.raises(NotImplementedError):
            LY.validate_lycoris_config(dict(algo="lokr", **bad) if "algo" not in bad else bad)
    with pytest.raises(NotImplementedError):
        LY.LycorisNetwork.apply_preset({"target_module": ["Attention"], "name_algo_map": {"x": {}}})
