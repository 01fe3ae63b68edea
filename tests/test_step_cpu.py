"""CPU: the loop body `TrainStep` (reference trainer.py:6951-7567 reduced to the hot path) on a toy wrapper — gradient
accumulation boundary, default value clip / norm clip, the deferred non-finite check, and DDP (Gloo, world size 2) with
`no_sync` on non-boundary micro-steps."""
import os
import tempfile
import traceback
from datetime import timedelta

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simpletuner_b200.training.step import TrainStep, wrap_ddp


class Toy:
    """Minimal family wrapper: same three methods the Trainer calls (SURVEY.md §8b seam B9)."""

    def __init__(self):
        torch.manual_seed(0)
        self.model = torch.nn.Linear(4, 1, bias=False)
        self.calls = []

    def prepare_batch(self, batch, state):
        self.calls.append(dict(state))
        return batch

    def model_predict(self, pb):
        return {"model_prediction": self.model(pb["x"])}

    def loss_with_logs(self, pb, out, apply_conditioning_mask=True):
        return ((out["model_prediction"] - pb["y"]) ** 2).mean(), None


def _batch(seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return {"x": torch.randn(8, 4, generator=g) * scale, "y": torch.randn(8, 1, generator=g)}


def test_accumulation_boundary_and_mean_gradient():
    w = Toy()
    opt = torch.optim.SGD(w.model.parameters(), lr=0.1)
    step = TrainStep(w, opt, max_grad_norm=0.0, gradient_accumulation_steps=2)
    w0 = w.model.weight.detach().clone()
    b1, b2 = _batch(1), _batch(2)
    step(b1)
    assert torch.equal(w.model.weight, w0) and step.state == {"global_step": 0, "micro_step": 1}
    step(b2)
    assert step.state == {"global_step": 1, "micro_step": 2} and w.model.weight.grad is None     # zero_grad(set_to_none)
    ref = Toy()
    l = sum(ref.loss_with_logs(b, ref.model_predict(b))[0] for b in (b1, b2)) / 2
    l.backward()
    assert torch.allclose(w.model.weight, w0 - 0.1 * ref.model.weight.grad, atol=1e-6)
    assert [c["global_step"] for c in w.calls] == [0, 0]


@pytest.mark.parametrize("method", ["value", "norm"])
def test_gradient_clipping_methods(method):
    w = Toy()
    opt = torch.optim.SGD(w.model.parameters(), lr=1.0)
    step = TrainStep(w, opt, max_grad_norm=0.05, grad_clip_method=method)
    w0 = w.model.weight.detach().clone()
    step(_batch(3, scale=10.0))
    delta = (w0 - w.model.weight).detach()
    if method == "value":     # trainer.py:7209-7213: clip_grad_value_ (element clamp), the reference default
        assert float(delta.abs().max()) <= 0.05 + 1e-7 and float(delta.abs().max()) == pytest.approx(0.05)
    else:                     # grad_clip_method=norm: trainer.py:7201-7205
        assert float(delta.norm()) == pytest.approx(0.05, rel=1e-4)
    with pytest.raises(ValueError):
        TrainStep(w, opt, max_grad_norm=1.0, grad_clip_method="bogus")(_batch(4))


def test_non_finite_loss_is_reported_at_check_time():
    w = Toy()
    step = TrainStep(w, torch.optim.SGD(w.model.parameters(), lr=0.0), max_grad_norm=0.0)
    step(_batch(5))
    step.check_finite()
    bad = _batch(6)
    bad["y"][0] = float("nan")
    step(bad)
    step(_batch(7))
    with pytest.raises(RuntimeError, match="Non-finite loss"):
        step.check_finite()
    step.check_finite()   # flag is cleared


def _ddp_worker(rank, world_size, init_method, q):
    try:
        dist.init_process_group("gloo", init_method=init_method, rank=rank, world_size=world_size, timeout=timedelta(seconds=30))
        w = Toy()
        wrap_ddp(w)
        opt = torch.optim.SGD(w.model.parameters(), lr=0.1)
        step = TrainStep(w, opt, max_grad_norm=0.0, gradient_accumulation_steps=2)
        step(_batch(10 + rank))         # no_sync micro-step
        step(_batch(20 + rank))         # boundary: one all-reduce of the accumulated gradients
        q.put(("ok", rank, w.model.module.weight.detach().flatten().tolist()))
    except BaseException:
        q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_ddp_gloo_accumulation_all_reduces_once_and_averages():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        init = f"file://{os.path.join(d, 'rdv')}"
        procs = [ctx.Process(target=_ddp_worker, args=(r, 2, init, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(90)
        assert all(not p.is_alive() for p in procs)
    res = sorted((q.get(timeout=5) for _ in procs), key=lambda t: t[1])
    assert [r[0] for r in res] == ["ok", "ok"], res
    assert res[0][2] == res[1][2]
    ref = Toy()
    l = sum(ref.loss_with_logs(b, ref.model_predict(b))[0] for b in (_batch(10), _batch(20), _batch(11), _batch(21))) / 4
    w0 = ref.model.weight.detach().clone()
    l.backward()
    assert torch.allclose(torch.tensor(res[0][2]), (w0 - 0.1 * ref.model.weight.grad).flatten(), atol=1e-6)


def _flat_worker(rank, world_size, init_method, q):
    try:
        dist.init_process_group("gloo", init_method=init_method, rank=rank, world_size=world_size, timeout=timedelta(seconds=30))
        from simpletuner_b200.training.dist import FlatGradSync
        w = Toy()
        with torch.no_grad():
            w.model.weight.add_(float(rank))          # ranks start apart: the constructor broadcasts rank 0's weights
        sync = FlatGradSync(w.model.parameters())
        opt = torch.optim.SGD(w.model.parameters(), lr=0.1)
        step = TrainStep(w, opt, max_grad_norm=0.0, gradient_accumulation_steps=2, grad_sync=sync)
        step(_batch(10 + rank))         # accumulates locally, no collective
        step(_batch(20 + rank))         # boundary: one flat all-reduce (mean) of the accumulated gradients
        q.put(("ok", rank, w.model.weight.detach().flatten().tolist()))
    except BaseException:
        q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_flat_grad_sync_gloo_matches_ddp_semantics():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        init = f"file://{os.path.join(d, 'rdv')}"
        procs = [ctx.Process(target=_flat_worker, args=(r, 2, init, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(90)
        assert all(not p.is_alive() for p in procs)
    res = sorted((q.get(timeout=5) for _ in procs), key=lambda t: t[1])
    assert [r[0] for r in res] == ["ok", "ok"], res
    assert res[0][2] == res[1][2]
    ref = Toy()
    l = sum(ref.loss_with_logs(b, ref.model_predict(b))[0] for b in (_batch(10), _batch(20), _batch(11), _batch(21))) / 4
    w0 = ref.model.weight.detach().clone()
    l.backward()
    assert torch.allclose(torch.tensor(res[0][2]), (w0 - 0.1 * ref.model.weight.grad).flatten(), atol=1e-6)


def _chunk_worker(rank, world_size, init_method, q):
    try:
        dist.init_process_group("gloo", init_method=init_method, rank=rank, world_size=world_size, timeout=timedelta(seconds=30))
        from simpletuner_b200.training.dist import FlatGradSync
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(n)) for n in (5, 300, 7, 64, 1000, 3)]
        sync = FlatGradSync(ps, pipeline_chunks=3)
        for i, p in enumerate(ps):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        works = sync.start_chunks()
        seen = []
        for work, chunk in works:
            if work is not None:
                work.wait()
            seen += [id(p) for p in chunk]
        ok = seen == [id(p) for p in ps] and 2 <= len(works) <= 3
        ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(ps))   # mean of ranks 1, 2
        q.put(("ok" if ok else "bad", rank, [len(c) for _, c in works]))
    except BaseException:
        q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_pipelined_gradient_chunks_cover_every_parameter_once_and_average():
    """FlatGradSync(pipeline_chunks=n).start_chunks(): the chunked exchange TrainStep overlaps with the optimizer launches
    (full fine-tune); world size 2 over Gloo."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        init = f"file://{os.path.join(d, 'rdv')}"
        procs = [ctx.Process(target=_chunk_worker, args=(r, 2, init, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(90)
        res = [q.get(timeout=5) for _ in range(2)]
    assert all(r[0] == "ok" for r in res), res


class _ChunkSGD(torch.optim.SGD):
    """SGD with the chunked-step surface of AdamWBF16 (`grad_clamp`, `only`, `salt`) so that the pipelined exchange of
    TrainStep can be exercised on the CPU."""
    supports_chunked_step = True

    @torch.no_grad()
    def step(self, grad_clamp=None, only=None, salt=0, **kw):
        ids = None if only is None else {id(p) for p in only}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None or (ids is not None and id(p) not in ids):
                    continue
                g = p.grad.clamp(-grad_clamp, grad_clamp) if grad_clamp else p.grad
                p.add_(g, alpha=-group["lr"])


class _Toy3(Toy):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.model = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 4), torch.nn.Linear(4, 4))


def _pipe_worker(rank, world_size, init_method, q, chunks):
    try:
        dist.init_process_group("gloo", init_method=init_method, rank=rank, world_size=world_size, timeout=timedelta(seconds=30))
        from simpletuner_b200.training.dist import FlatGradSync
        w = _Toy3()
        params = list(w.model.parameters())
        sync = FlatGradSync(params, pipeline_chunks=chunks)
        step = TrainStep(w, _ChunkSGD(params, lr=0.1), max_grad_norm=0.05, grad_clip_method="value", grad_sync=sync)
        assert step._pipelined_ok() == (chunks > 0)
        for i in range(3):
            step(_batch(10 * i + rank))
        q.put(("ok", rank, torch.cat([p.detach().flatten() for p in params]).tolist()))
    except BaseException:
        q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_pipelined_exchange_and_chunked_optimizer_equal_the_flat_path():
    """TrainStep with FlatGradSync(pipeline_chunks=3) (per-chunk all-reduce -> per-chunk clamp + update) ends at exactly the
    weights of the flat exchange + one optimizer step, on both ranks (Gloo, world size 2)."""
    ctx = mp.get_context("spawn")
    res = {}
    for chunks in (0, 3):
        q = ctx.Queue()
        with tempfile.TemporaryDirectory() as d:
            init = f"file://{os.path.join(d, 'rdv')}"
            procs = [ctx.Process(target=_pipe_worker, args=(r, 2, init, q, chunks)) for r in range(2)]
            for p in procs:
                p.start()
            for p in procs:
                p.join(90)
            out = sorted((q.get(timeout=5) for _ in procs), key=lambda t: t[1])
        assert [o[0] for o in out] == ["ok", "ok"], out
        assert out[0][2] == out[1][2]                      # replicas stay identical
        res[chunks] = torch.tensor(out[0][2])
    assert torch.allclose(res[0], res[3], rtol=0, atol=1e-7)
