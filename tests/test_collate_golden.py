"""CPU: size-feature gathers and the logged gradient infinity-norm, pinned bit-exactly to the reference's own source executed
verbatim (oracle/make_golden_collate.py -> tests/golden/collate_golden.pt; collate.py:59-98, 487-523; trainer.py:6376-6398)."""
from pathlib import Path
from types import SimpleNamespace

import torch

from simpletuner_b200.training import noise as N

G = torch.load(Path(__file__).parent / "golden" / "collate_golden.pt", weights_only=False)


def test_sdxl_size_features_bit_exact():
    lat = torch.zeros(G["sdxl.latent_shape"])
    for dt, key in ((torch.bfloat16, "sdxl.time_ids.bf16"), (torch.float32, "sdxl.time_ids.f32")):
        got = N.gather_conditional_sdxl_size_features(G["sdxl.examples"], lat, dt)
        assert got.dtype == dt and got.shape == (3, 1, 6) and torch.equal(got, G[key])
    assert G["sdxl.time_ids.f32"][0, 0].tolist() == [768.0, 1024.0, 0.0, 0.0, 768.0, 1024.0]   # [orig_h, orig_w, crop, tgt_h, tgt_w]
    assert float(G["sdxl.time_ids.f32"][2].abs().sum()) == 0.0                                  # dropped conditioning -> zeros
    try:
        N.gather_conditional_sdxl_size_features(G["sdxl.examples"][:2], lat, torch.float32)
        assert False
    except ValueError as e:
        assert "must match" in str(e)


def test_pixart_size_features_bit_exact():
    got = N.gather_conditional_pixart_size_features([0, 1, 2], torch.zeros(3, 4, 160, 96), torch.bfloat16, device="cpu")
    assert torch.equal(got["resolution"], G["pixart.resolution"]) and torch.equal(got["aspect_ratio"], G["pixart.aspect_ratio"])
    assert got["resolution"][0].tolist() == [1280.0, 768.0]


def test_max_grad_value_matches_reference():
    params = [SimpleNamespace(grad=t) for t in G["maxgrad.grads"]] + [SimpleNamespace(grad=None)]
    got = N.max_grad_value(params)
    assert torch.equal(torch.as_tensor(got), torch.as_tensor(G["maxgrad.out"]))
    assert N.max_grad_value([SimpleNamespace(grad=None)]) == G["maxgrad.empty"] == float("-inf")


def test_trainstep_tracks_the_logged_grad_norm():
    from simpletuner_b200.training.step import TrainStep
    from tests.test_step_cpu import Toy, _batch
    w = Toy()
    step = TrainStep(w, torch.optim.SGD(w.model.parameters(), lr=0.0), max_grad_norm=2.0, grad_clip_method="value")
    step.track_grad_norm = True
    step.optimizer.zero_grad = lambda set_to_none=True: None          # keep the gradient for the comparison
    step(_batch(1, scale=3.0))
    assert float(step.grad_norm) > 0   # inf-norm of the UNCLIPPED gradient (trainer.py:7144-7147 runs before the clamp)
    ref = Toy(); b = _batch(1, scale=3.0)
    ref.loss_with_logs(b, ref.model_predict(b))[0].backward()
    assert abs(float(step.grad_norm) - float(ref.model.weight.grad.abs().max())) < 1e-6
