"""CPU: host-side index / schedule logic and the oracle restatements, pinned bit-exactly against
vectors produced by the reference's OWN source (oracle/make_golden.py -> tests/golden/)."""
from types import SimpleNamespace

import torch

from oracle import flux_oracle as O
from simpletuner_b200.flux import functional as FX
from simpletuner_b200.flux.model import prepare_latent_image_ids
from simpletuner_b200.training import schedule as S


def test_pack_unpack_ids_bit_exact(golden):
    lat = golden["pack.in"]
    for mod in (FX, O):
        assert torch.equal(mod.pack_latents(lat, 2, 16, 8, 12), golden["pack.out"])
        assert torch.equal(mod.unpack_latents(golden["pack.out"], 64, 96, 16), golden["unpack.out"])
    assert torch.equal(golden["unpack.out"], lat)  # round trip
    assert torch.equal(prepare_latent_image_ids(8, 12), golden["ids.out_8x12"])
    assert torch.equal(O.prepare_latent_image_ids(8, 12), golden["ids.out_8x12"])


def test_flow_schedule_shift_bit_exact(golden):
    sig = golden["shift.in"]
    c3 = SimpleNamespace(flow_schedule_shift=3.0, flow_schedule_auto_shift=False)
    c1 = SimpleNamespace(flow_schedule_shift=1.0, flow_schedule_auto_shift=False)
    assert torch.equal(S.apply_flow_schedule_shift(c3, None, sig.clone(), None), golden["shift.out_s3"])
    assert torch.equal(S.apply_flow_schedule_shift(c1, None, sig.clone(), None), golden["shift.out_s1"])
    assert torch.equal(O.apply_flow_schedule_shift(sig.clone(), 3.0), golden["shift.out_s3"])
    # known answer quoted in SURVEY.md §8c: [0.1, 0.5, 0.9] -> [0.25, 0.75, 0.9643]
    torch.testing.assert_close(golden["shift.out_s3"][:3], torch.tensor([0.25, 0.75, 0.9643]), atol=1e-4, rtol=0)


def test_auto_shift_matches_formula():
    c = SimpleNamespace(flow_schedule_shift=0.0, flow_schedule_auto_shift=True)
    sched = SimpleNamespace(config=SimpleNamespace(patch_size=2, base_image_seq_len=256, max_image_seq_len=4096,
                                                   base_shift=0.5, max_shift=1.15))
    noise = torch.zeros(1, 16, 128, 128)
    out = S.apply_flow_schedule_shift(c, sched, torch.tensor([0.5]), noise)
    import math
    s = math.exp(1.15)  # seq_len 4096 -> mu = max_shift
    torch.testing.assert_close(out, torch.tensor([0.5 * s / (1 + (s - 1) * 0.5)]))


def test_timestep_weights_bit_exact(golden):
    for strat, kw in (("none", {}), ("later", {}), ("earlier", {}), ("range", dict(timestep_bias_begin=200, timestep_bias_end=500))):
        a = SimpleNamespace(timestep_bias_strategy=strat, timestep_bias_portion=0.25, timestep_bias_multiplier=2.0,
                            timestep_bias_begin=kw.get("timestep_bias_begin", 0), timestep_bias_end=kw.get("timestep_bias_end", 1000))
        assert torch.equal(S.generate_timestep_weights(a, 1000), golden[f"tsw.{strat}"])


def test_segmented_timestep_selection_bit_exact_including_quirk_q1(golden):
    cfg = SimpleNamespace(refiner_training=False, refiner_training_invert_schedule=False, refiner_training_strength=0.2)
    for bsz in (2, 4, 7):
        torch.manual_seed(42)
        w = torch.ones(1000)
        sel = S.segmented_timestep_selection(1000, bsz, w, cfg)
        assert torch.equal(sel, golden[f"segsel.bsz{bsz}"])
        assert sel.dtype == golden[f"segsel.bsz{bsz}"].dtype
        # in-place normalisation of the caller's weights (quirk Q1)
        assert torch.equal(w, golden[f"segsel.bsz{bsz}.weights_after"])
    assert golden["segsel.bsz4"].tolist() == [960, 545, 311, 69]
    # range property pinned by reference tests/test_custom_schedules.py:59-107
    rcfg = SimpleNamespace(refiner_training=True, refiner_training_invert_schedule=True, refiner_training_strength=0.35)
    sel = S.segmented_timestep_selection(1000, 4, torch.ones(1000), rcfg)
    assert all(350 <= int(t) <= 999 for t in sel)
    rcfg = SimpleNamespace(refiner_training=True, refiner_training_invert_schedule=False, refiner_training_strength=0.35)
    sel = S.segmented_timestep_selection(1000, 4, torch.ones(1000), rcfg)
    assert all(0 <= int(t) < 350 for t in sel)


def test_sample_flow_sigmas_bit_exact(golden):
    c = SimpleNamespace(flow_schedule_shift=3.0, flow_schedule_auto_shift=False, flow_sigmoid_scale=1.0)
    torch.manual_seed(42)
    sig, ts = S.sample_flow_sigmas(c, None, 4, torch.zeros(4, 16, 8, 8), "cpu")
    assert torch.equal(sig, golden["sample_sigmas.seed42.sigmas"])
    assert torch.equal(ts, golden["sample_sigmas.seed42.timesteps"])
    torch.manual_seed(42)
    sig2, _ = O.sample_flow_sigmas(4)
    assert torch.equal(sig2, golden["sample_sigmas.seed42.sigmas"])


def test_oracle_noisy_latents_and_target_bit_exact(golden):
    assert golden["noisy.known_answer"].item() == 3.0  # reference tests/test_mixflow.py:76-87
    for tag in ("f32", "bf16"):
        lat, eps, sg = golden[f"noisy.{tag}.latents"], golden[f"noisy.{tag}.noise"], golden[f"noisy.{tag}.sigmas"]
        assert torch.equal(O.flow_noisy_latents(lat, eps, sg), golden[f"noisy.{tag}.out"])
        assert torch.equal(O.flow_target(lat, eps), golden[f"target.{tag}.out"])


def test_oracle_rope_application_bit_exact(golden):
    x, cos, sin = golden["rope.x"], golden["rope.cos"], golden["rope.sin"]
    assert torch.equal(O.apply_rope(x, cos, sin), golden["rope.out"])
    assert torch.equal(O.apply_rope(x.bfloat16(), cos, sin), golden["rope.out_bf16"])


def test_rope_tables_product_equals_oracle():
    from simpletuner_b200.flux.transformer import rope_tables
    ids = torch.cat([torch.zeros(5, 3), O.prepare_latent_image_ids(8, 12)], 0)
    c1, s1 = rope_tables(ids, (16, 56, 56))
    c2, s2 = O.rope_tables(ids, (16, 56, 56))
    assert torch.equal(c1, c2) and torch.equal(s1, s2)
    assert torch.all(c1[:5] == 1) and torch.all(s1[:5] == 0)  # text tokens: identity rotation (quirk Q9)


def test_flux_guidance_modes_follow_reference_draws():
    """flux/model.py:682-706: constant value, or `random.uniform(min, max)` once per sample from python's generator."""
    import random
    from types import SimpleNamespace

    from simpletuner_b200.flux.model import Flux, default_config

    w = Flux.__new__(Flux)
    w.model = SimpleNamespace(config=SimpleNamespace(guidance_embeds=True))
    w.config = default_config(flux_guidance_mode="random-range", flux_guidance_min=1.5, flux_guidance_max=3.5)
    random.seed(11)
    got = w._guidance(3, "cpu")
    random.seed(11)
    want = [random.uniform(1.5, 3.5) for _ in range(3)]
    assert torch.allclose(got, torch.tensor(want, dtype=torch.float32))
    w.config = default_config()
    assert torch.equal(w._guidance(2, "cpu"), torch.ones(2))
    w.model.config.guidance_embeds = False
    assert w._guidance(2, "cpu") is None


# ---- custom timestep lists: the reference's own known answers (reference tests/test_flow_custom_timesteps.py:33-118)
def _sampler(custom, mode, world=(2,), rank=0):
    from types import SimpleNamespace

    from simpletuner_b200.training.schedule import FlowSigmaSampler

    cfg = SimpleNamespace(flow_custom_timesteps=custom, flow_timesteps_mode=mode)
    sizes = list(world)

    def layout(bsz):
        return SimpleNamespace(global_batch_size=sum(sizes), local_batch_offset=sum(sizes[:rank]))

    return FlowSigmaSampler(cfg, None, "cpu", layout_fn=layout)


def test_round_robin_cycles_custom_timesteps():
    s = _sampler("100,200,300", "round-robin", world=(2,))
    assert torch.equal(s.sample(2, None, {})[1], torch.tensor([100.0, 200.0]))
    assert torch.equal(s.sample(2, None, {})[1], torch.tensor([300.0, 100.0]))


def test_round_robin_offsets_distributed_ranks_and_varying_batch_sizes():
    r0, r1 = _sampler("100,200,300,400,500", "round-robin", (2, 2), 0), _sampler("100,200,300,400,500", "round-robin", (2, 2), 1)
    assert torch.equal(r0.sample(2, None, {"global_step": 0})[1], torch.tensor([100.0, 200.0]))
    assert torch.equal(r1.sample(2, None, {"global_step": 0})[1], torch.tensor([300.0, 400.0]))
    assert torch.equal(r0.sample(2, None, {"global_step": 0})[1], torch.tensor([500.0, 100.0]))
    lst = "100,200,300,400,500,600,700,800"
    r0, r1 = _sampler(lst, "round-robin", (1, 3), 0), _sampler(lst, "round-robin", (1, 3), 1)
    assert torch.equal(r0.sample(1, None, {"global_step": 0})[1], torch.tensor([100.0]))
    assert torch.equal(r1.sample(3, None, {"global_step": 0})[1], torch.tensor([200.0, 300.0, 400.0]))
    assert torch.equal(r0.sample(1, None, {"global_step": 0})[1], torch.tensor([500.0]))


def test_round_robin_resume_and_reset():
    s = _sampler("100,200,300,400,500", "round-robin", (2, 2), 0)
    assert torch.equal(s.sample(2, None, {"global_step": 1})[1], torch.tensor([500.0, 100.0]))    # cursor = 1 * 4 % 5
    s = _sampler("100,200,300,400,500", "round-robin", (2, 2), 0)
    s.sample(2, None, {"global_step": 0})
    s.reset_cursor(global_step=1)
    assert torch.equal(s.sample(2, None, {"global_step": 1})[1], torch.tensor([500.0, 100.0]))


def test_custom_list_modes_and_sigma_interpretation():
    import pytest as _pt
    with _pt.raises(ValueError, match="flow_timesteps_mode"):
        _sampler("100,200", "sequential").sample(1, None, {})
    sig, t = _sampler("0.25;0.5", "fixed-list").sample(64, None, {})          # values <= 1 are sigmas
    assert set(sig.tolist()) <= {0.25, 0.5} and torch.equal(t, sig * 1000.0)
    sig, t = _sampler("[750]", "fixed-list").sample(3, None, {})               # JSON list, single entry
    assert torch.equal(t, torch.full((3,), 750.0)) and torch.equal(sig, torch.full((3,), 0.75))


def test_other_flow_schedules_draw_like_the_reference():
    import random
    from types import SimpleNamespace

    from simpletuner_b200.training.schedule import FlowSigmaSampler

    base = dict(flow_custom_timesteps=None, flow_schedule_shift=1.0, flow_schedule_auto_shift=False, flow_sigmoid_scale=1.0,
                flux_fast_schedule=False, flow_use_beta_schedule=False, flow_use_uniform_schedule=False)
    torch.manual_seed(3)
    sig, t = FlowSigmaSampler(SimpleNamespace(**{**base, "flow_use_uniform_schedule": True}), None, "cpu").sample(5, None)
    torch.manual_seed(3)
    assert torch.equal(sig, torch.rand((5,))) and torch.equal(t, sig * 1000.0)
    random.seed(5)
    sig, _ = FlowSigmaSampler(SimpleNamespace(**{**base, "flux_fast_schedule": True}), None, "cpu").sample(6, None)
    random.seed(5)
    assert sig.tolist() == random.choices([1.0] * 7 + [0.75, 0.5, 0.25], k=6)
    torch.manual_seed(9)
    sig, _ = FlowSigmaSampler(SimpleNamespace(**{**base, "flow_use_beta_schedule": True, "flow_beta_schedule_alpha": 2.0,
                                                 "flow_beta_schedule_beta": 2.0}), None, "cpu").sample(4, None)
    torch.manual_seed(9)
    assert torch.equal(sig, torch.distributions.Beta(2.0, 2.0).sample((4,)))
    torch.manual_seed(1)
    sig, _ = FlowSigmaSampler(SimpleNamespace(**{**base, "mixflow_enabled": True}), None, "cpu").sample(4, None)
    torch.manual_seed(1)
    assert torch.equal(sig, 1.0 - torch.sqrt(torch.rand((4,))))
