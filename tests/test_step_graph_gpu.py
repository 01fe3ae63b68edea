"""GPU (-m gpu): GraphedTrainStep (CUDA-graph replay of the micro-step) against the eager TrainStep on a deterministic toy
model: per-bucket graphs, gradient accumulation into shared persistent buffers (a capture landing in the middle of an
accumulation window must not disturb it), eager `prepare_batch` mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu


class _Toy:
    """Family-wrapper-shaped toy: y = W x with an MSE loss; prepare_batch is deterministic (scales the input)."""

    def __init__(self, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.model = torch.nn.Linear(16, 16, bias=False).cuda()
        with torch.no_grad():
            self.model.weight.copy_(torch.randn(16, 16, generator=g) * 0.1)
        self.prepared_calls = 0

    def prepare_batch(self, batch, state):
        self.prepared_calls += 1
        return {"x": batch["x"] * 2.0, "y": batch["y"]}

    def model_predict(self, pb):
        return self.model(pb["x"])

    def loss_with_logs(self, pb, out, apply_conditioning_mask=True):
        return ((out - pb["y"]) ** 2).mean(), {}


def _batches(n, shapes):
    g = torch.Generator().manual_seed(1)
    return [{"x": torch.randn(shapes[i % len(shapes)], 16, generator=g).cuda(), "y": torch.randn(shapes[i % len(shapes)], 16, generator=g).cuda()}
            for i in range(n)]


@pytest.mark.parametrize("accum,capture_prepare", [(1, True), (2, True), (3, False)])
def test_graph_replay_matches_eager_step(accum, capture_prepare):
    from simpletuner_b200.training.step import GraphedTrainStep, TrainStep
    batches = _batches(12, shapes=(4, 7, 5))            # three "buckets"; with accum = 2 / 3 a new shape is first seen mid-window
    a, b = _Toy(), _Toy()
    eager = TrainStep(a, torch.optim.SGD(a.model.parameters(), lr=0.05), max_grad_norm=0.0, gradient_accumulation_steps=accum)
    graphed = GraphedTrainStep(TrainStep(b, torch.optim.SGD(b.model.parameters(), lr=0.05), max_grad_norm=0.0,
                                         gradient_accumulation_steps=accum), capture_prepare=capture_prepare)
    for bt in batches:
        le = eager(dict(bt))
        lg = graphed(dict(bt))
        assert torch.allclose(le, lg, rtol=1e-5, atol=1e-6), (float(le), float(lg))
    torch.cuda.synchronize()
    assert eager.state["global_step"] == graphed.state["global_step"] == 12 // accum
    assert torch.allclose(a.model.weight, b.model.weight, rtol=1e-5, atol=1e-6)
    assert len(graphed._graphs) == 3
    if not capture_prepare:
        assert b.prepared_calls == 12                  # prepare_batch ran eagerly on every call, never inside a capture
