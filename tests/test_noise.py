"""CPU: ε / v-prediction schedule helpers, pinned bit-exactly to the reference's own source where it owns them."""
import torch

from simpletuner_b200.training import noise as N


def test_compute_snr_bit_exact(golden):
    sched = N.make_ddpm_schedule()
    ts = golden["snr.timesteps"]
    assert torch.equal(N.compute_snr(ts, sched), golden["snr.out"])
    assert torch.equal(N.compute_snr(ts, sched, use_soft_min=True, sigma_data=0.5), golden["snr.out_softmin"])


def test_compute_time_ids_bit_exact(golden):
    a = N.compute_time_ids((1024, 768), (4, 96, 128), torch.float32, crop_coordinates=[0, 0])
    assert torch.equal(a, golden["time_ids.a"])
    assert a.tolist() == [[768.0, 1024.0, 0.0, 0.0, 768.0, 1024.0]]  # value quoted in SURVEY.md §8c
    b = N.compute_time_ids((1536, 640), (4, 80, 192), torch.bfloat16, crop_coordinates=[12, 34])
    assert torch.equal(b, golden["time_ids.b"]) and b.dtype == torch.bfloat16


def test_add_noise_velocity_and_min_snr():
    sched = N.make_ddpm_schedule()
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(3, 4, 8, 8, generator=g), torch.randn(3, 4, 8, 8, generator=g)
    t = torch.tensor([0, 500, 999])
    xt = N.add_noise(sched, x, e, t)
    a = sched.alphas_cumprod[t].sqrt().view(3, 1, 1, 1)
    s = (1 - sched.alphas_cumprod[t]).sqrt().view(3, 1, 1, 1)
    torch.testing.assert_close(xt, a * x + s * e)
    v = N.get_velocity(sched, x, e, t)
    # round trip: x = a * x_t - s * v
    torch.testing.assert_close(a * xt - s * v, x, atol=1e-5, rtol=1e-5)
    w = N.min_snr_loss_weights(t, sched, 5.0, "epsilon")
    snr = N.compute_snr(t, sched)
    torch.testing.assert_close(w, torch.minimum(snr, torch.tensor(5.0)) / snr)
    wv = N.min_snr_loss_weights(t, sched, 5.0, "v_prediction")
    torch.testing.assert_close(wv, torch.minimum(snr, torch.tensor(5.0)) / (snr + 1))


def test_sample_noise_draw_order_and_options():
    """common.py:5936-5967: randn_like -> [random() gate] -> [randn(B,C,1,1)] -> [randn_like]; the loss target keeps
    the unperturbed noise, only `input_noise` carries the perturbation; flow matching never applies offset noise."""
    import random
    from types import SimpleNamespace

    from simpletuner_b200.training.noise import sample_noise

    lat = torch.zeros(3, 4, 5, 6)
    cfg = SimpleNamespace(offset_noise=True, noise_offset=0.1, noise_offset_probability=1.0, input_perturbation=0.2,
                          input_perturbation_steps=100)
    torch.manual_seed(7)
    random.seed(7)
    noise, inp = sample_noise(cfg, lat, {"global_step": 25}, flow_matching=False)
    torch.manual_seed(7)
    n0 = torch.randn_like(lat)
    n1 = n0 + 0.1 * torch.randn(3, 4, 1, 1)
    i1 = n1 + 0.2 * (1.0 - 25 / 100) * torch.randn_like(lat)
    assert torch.equal(noise, n1) and torch.equal(inp, i1)
    torch.manual_seed(7)
    noise_f, inp_f = sample_noise(cfg, lat, {"global_step": 250}, flow_matching=True)   # past the perturbation window
    assert torch.equal(noise_f, n0) and inp_f is noise_f
    cfg.noise_offset_probability = 0.0
    torch.manual_seed(7)
    random.seed(1)
    noise_g, _ = sample_noise(cfg, lat, {"global_step": 250}, flow_matching=False)
    assert torch.equal(noise_g, n0)
