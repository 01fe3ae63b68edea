"""CPU, world_size 2 over Gloo: the data-parallel metric collectives (rank-varying batch sizes) —
same scenario and expected values as reference tests/test_distributed_batch_layout.py:48-111."""
import os
import tempfile
import traceback
from datetime import timedelta

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world_size, init_method, q):
    try:
        dist.init_process_group("gloo", init_method=init_method, rank=rank, world_size=world_size, timeout=timedelta(seconds=20))
        from simpletuner_b200.training import dist as D

        n = 1 if rank == 0 else 3
        loss = torch.tensor(2.0 if rank == 0 else 4.0)
        vals = torch.tensor([10.0]) if rank == 0 else torch.tensor([20.0, 30.0, 40.0])
        layout = D.resolve_batch_layout(n)
        w = D.gather_sample_weighted_scalar(loss, n)
        g = D.gather_variable_batch_tensor(vals, layout)
        # round-robin custom timesteps over the REAL collective layout (reference tests/test_flow_custom_timesteps.py:66-82)
        from types import SimpleNamespace
        from simpletuner_b200.training.schedule import FlowSigmaSampler
        cfg = SimpleNamespace(flow_custom_timesteps="100,200,300,400,500,600,700,800", flow_timesteps_mode="round-robin")
        smp = FlowSigmaSampler(cfg, None, "cpu")
        t_first = smp.sample(n, None, {"global_step": 0})[1].tolist()
        t_next = smp.sample(n, None, {"global_step": 0})[1].tolist()
        q.put(("ok", rank, float(w), g.tolist(), layout.global_batch_size, layout.local_batch_offset,
               D.device_seed(42, rank), list(D.shard_units(8, rank, world_size)), t_first, t_next))
    except BaseException:
        q.put(("error", rank, traceback.format_exc()))
        raise
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_gloo_rank_varying_batch_collectives():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        init = f"file://{os.path.join(d, 'rdv')}"
        procs = [ctx.Process(target=_worker, args=(r, 2, init, q)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(60)
        assert all(not p.is_alive() for p in procs)
    res = sorted((q.get(timeout=5) for _ in procs), key=lambda t: t[1])
    assert [r[0] for r in res] == ["ok", "ok"], res
    for _, rank, w, g, total, off, seed, units, t_first, t_next in res:
        assert t_first == ([100.0] if rank == 0 else [200.0, 300.0, 400.0])
        assert t_next == ([500.0] if rank == 0 else [600.0, 700.0, 800.0])
        assert w == 3.5
        assert g == [10.0, 20.0, 30.0, 40.0]
        assert total == 4 and off == (0 if rank == 0 else 1)
        assert seed == 42 + rank
        assert units == ([0, 1, 2, 3] if rank == 0 else [4, 5, 6, 7])


def test_single_process_passthrough():
    from simpletuner_b200.training import dist as D

    assert float(D.gather_sample_weighted_scalar(torch.tensor(2.5), 3)) == 2.5
    lay = D.resolve_batch_layout(3)
    assert (lay.global_batch_size, lay.local_batch_offset, lay.world_size) == (3, 0, 1)
