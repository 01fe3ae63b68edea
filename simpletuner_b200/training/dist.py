"""Data-parallel metric collectives of the training step (host-side logic over torch.distributed).

Behavioural mirror of reference helpers/data_backend/runtime/context_parallel_sync.py:235-348 for
the pure data-parallel case the B200 path runs (one process per GPU, NCCL; Gloo in CPU tests):
ranks may hold different local batch sizes (aspect buckets), so the logged loss is the
sample-weighted world mean, and per-sample tensors are gathered with padding.
These run at logging cadence, not inside the timed step (training/step.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class BatchLayout:
    local_batch_size: int
    global_batch_size: int
    local_batch_offset: int
    rank: int
    world_size: int
    batch_sizes: Tuple[int, ...]


def _world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _all_gather_cat(t: torch.Tensor) -> torch.Tensor:
    _, ws = _world()
    out = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(out, t)
    return torch.cat(out, dim=0)


def resolve_batch_layout(local_batch_size: int, device=None) -> BatchLayout:
    """context_parallel_sync.py:235-295 without model-parallel groups: global count and this rank's prefix offset."""
    local_batch_size = int(local_batch_size)
    if local_batch_size < 1:
        raise ValueError("local_batch_size must be greater than 0.")
    rank, ws = _world()
    if ws == 1:
        return BatchLayout(local_batch_size, local_batch_size, 0, 0, 1, (local_batch_size,))
    count = torch.tensor([local_batch_size], device=device, dtype=torch.long)
    sizes = tuple(int(v) for v in _all_gather_cat(count).cpu().tolist())
    return BatchLayout(local_batch_size, sum(sizes), sum(sizes[:rank]), rank, ws, sizes)


def gather_sample_weighted_scalar(value: torch.Tensor, local_batch_size: int) -> torch.Tensor:
    """context_parallel_sync.py:327-348: sum_r(loss_r * n_r) / sum_r(n_r) from one [2] contribution per rank."""
    local_batch_size = int(local_batch_size)
    if local_batch_size < 1:
        raise ValueError("local_batch_size must be greater than 0.")
    if value.numel() != 1:
        raise ValueError("Sample-weighted scalar gather requires a scalar tensor.")
    value = value.detach().float().reshape(())
    _, ws = _world()
    if ws == 1:
        return value
    n = value.new_tensor(float(local_batch_size))
    gathered = _all_gather_cat(torch.stack((value * n, n))).reshape(-1, 2)
    totals = gathered.sum(dim=0)
    return totals[0] / totals[1]


def gather_variable_batch_tensor(tensor: torch.Tensor, layout: Optional[BatchLayout] = None) -> torch.Tensor:
    """context_parallel_sync.py:298-324: gather per-sample tensors when ranks hold different batch sizes."""
    if tensor.ndim < 1:
        raise ValueError("Variable batch tensor gather requires a batch dimension.")
    if layout is None:
        layout = resolve_batch_layout(tensor.shape[0], tensor.device)
    if layout.world_size == 1:
        return tensor.detach()
    mx = max(layout.batch_sizes)
    padded = tensor.detach()
    if padded.shape[0] < mx:
        padded = torch.cat((padded, padded.new_zeros((mx - padded.shape[0], *padded.shape[1:]))), dim=0)
    g = _all_gather_cat(padded).reshape(layout.world_size, mx, *padded.shape[1:])
    return torch.cat([g[r, :layout.batch_sizes[r]] for r in range(layout.world_size)], dim=0)


def device_seed(seed: int, rank: int, seed_for_each_device: bool = True) -> int:
    """accelerate.utils.set_seed(seed, device_specific=True) as used at trainer.py:2554-2556: rank r seeds with seed + r."""
    return seed + rank if seed_for_each_device else seed


def shard_units(num_units: int, rank: int, world_size: int) -> range:
    """Contiguous split of independent work units (micro-batches) across ranks for weak-scaling runs."""
    per = num_units // world_size
    return range(rank * per, (rank + 1) * per)


class FlatGradSync:
    """Data-parallel gradient exchange as ONE flat all-reduce after backward, on the compute stream.

    The reference wraps the model in torch DDP (accelerator.prepare, trainer.py:4571, 1026-1041): bucketed all-reduces
    launched from autograd hooks on NCCL's side stream, overlapping the backward.  For LoRA training the whole exchange
    is 52 MB (26 M bf16 gradients): < 1 ms over NVLink 5 / NVSwitch, while the overlapped NCCL kernels take SMs away from
    the persistent one-CTA-per-SM GEMM / attention grids of the backward for its whole duration (measured in round 1:
    +36 ms per step from 4 ranks up, SCALE_r01.json).  So the exchange runs after backward instead: cat -> all_reduce
    (mean) -> foreach copy back, three launches plus one collective.  `__call__` is invoked by TrainStep on
    accumulation boundaries only (the reference's `no_sync` on the other micro-steps).  Parameters are broadcast from
    rank 0 at construction, as DDP does."""

    def __init__(self, params, group=None, broadcast: bool = True, pipeline_chunks: int = 0):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        # pipeline_chunks > 0 (large gradient sets: full fine-tune): `start_chunks()` issues one asynchronous coalesced
        # all-reduce per chunk directly on the gradient tensors (no flat copy) and TrainStep runs the optimizer on chunk i
        # while chunk i + 1 is still on the wire (training/step.py::_optimizer_step)
        self.pipeline_chunks = int(pipeline_chunks)
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.sizes = [p.numel() for p in self.params]
        if broadcast and self.world > 1 and self.params:
            with torch.no_grad():
                flat = torch.cat([p.detach().reshape(-1) for p in self.params])
                dist.broadcast(flat, src=0, group=group)
                torch._foreach_copy_([p.data for p in self.params], [c.view_as(p) for c, p in zip(flat.split(self.sizes), self.params)])

    @torch.no_grad()
    def __call__(self) -> None:
        if self.world == 1 or not self.params:
            return
        grads = []
        for p in self.params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads.append(p.grad)
        flat = torch.cat([g.reshape(-1) for g in grads])
        backend = dist.get_backend(self.group)
        if backend == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.div_(self.world)
        torch._foreach_copy_(grads, [c.view_as(g) for c, g in zip(flat.split(self.sizes), grads)])

    def chunks(self):
        """Parameters in `pipeline_chunks` groups of roughly equal size (parameter order kept)."""
        n = max(1, self.pipeline_chunks)
        total = sum(self.sizes)
        target = (total + n - 1) // n
        out, cur, acc = [], [], 0
        for p, sz in zip(self.params, self.sizes):
            cur.append(p)
            acc += sz
            if acc >= target and len(out) < n - 1:
                out.append(cur)
                cur, acc = [], 0
        if cur:
            out.append(cur)
        return out

    @torch.no_grad()
    def start_chunks(self):
        """[(work handle or None, parameters of the chunk)]: the mean all-reduce of every chunk's gradients is in flight (NCCL:
        one coalesced, asynchronous collective per chunk on NCCL's stream) or already done (other backends)."""
        out = []
        for chunk in self.chunks():
            grads = []
            for p in chunk:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                grads.append(p.grad)
            if self.world == 1:
                out.append((None, chunk))
                continue
            if dist.get_backend(self.group) == "nccl":
                with dist._coalescing_manager(group=self.group, async_ops=True) as cm:     # -> one allreduce_coalesced
                    for g in grads:
                        dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.group)
                out.append((cm, chunk))
            else:
                for g in grads:
                    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
                    g.div_(self.world)
                out.append((None, chunk))
        return out
