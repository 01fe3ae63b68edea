"""Latent cache wire format + pinned, batched host->device staging.

Reference: `VAECache` writes ONE `torch.save`d tensor per image under a name derived by `generate_vae_cache_filename`
(helpers/caching/vae.py:678-703), and reads them back per sample: `retrieve_from_cache` -> `.to("cpu").pin_memory()` per latent
(helpers/training/collate.py:177-190) -> `torch.stack` in collate -> a synchronous `.to(device)` in `prepare_batch`
(common.py:5885-5913).  With the step at ~0.4-0.7 s that per-sample pin + stack + blocking copy path is what keeps 8 GPUs from
being fed.

This module keeps the WIRE FORMAT (file naming and `torch.save` payload are unchanged, so caches written by either side are
interchangeable) and replaces the read side:
  * `LatentStager` owns a ring of pre-allocated PINNED batch buffers per bucket shape; a batch's files are read by a small
    thread pool straight into slot rows (no per-sample pin, no `torch.stack`), then ONE async H2D copy per tensor is issued on a
    dedicated copy stream and fenced with an event;
  * `next()` hands the step device tensors plus the event; the consumer waits on the event from its compute stream
    (`torch.cuda.current_stream().wait_event`) — the host never blocks on a copy, and batch i+1 is staged while step i runs.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from hashlib import sha256
from typing import Dict, List, Optional, Sequence, Tuple

import torch


def generate_vae_cache_filename(filepath: str, cache_dir: str, instance_data_dir: Optional[str], hash_filenames: bool = True) -> Tuple[str, str]:
    """helpers/caching/vae.py:678-703 (image files; sample-id backends excluded): (full cache path, base filename).
    Known answers of the reference's own test (tests/test_vae.py:50-109) are reproduced in tests/test_latent_cache.py."""
    if filepath.endswith(".pt"):
        return filepath, os.path.basename(filepath)
    base_filename = os.path.splitext(os.path.basename(filepath))[0]
    if hash_filenames:
        base_filename = str(sha256(str(base_filename).encode()).hexdigest())
    base_filename = str(base_filename) + ".pt"
    subfolders = ""
    if instance_data_dir is not None:
        subfolders = os.path.dirname(filepath).replace(instance_data_dir, "")
        subfolders = subfolders.lstrip(os.sep)
    if len(subfolders) > 0:
        full_filename = os.path.join(cache_dir, subfolders, base_filename)
    else:
        full_filename = os.path.join(cache_dir, base_filename)
    return full_filename, base_filename


def write_latents(filepaths: Sequence[str], latents: Sequence[torch.Tensor]) -> None:
    """`_write_latents_in_batch` (vae.py:1398-1449) for plain image latents: one `torch.save(tensor.clone())` per `.pt` path."""
    for path, lat in zip(filepaths, latents):
        if os.path.splitext(path)[1] != ".pt":
            raise ValueError(f"Cannot write a latent embedding to an image path, {path}")
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        torch.save(lat.detach().to("cpu").clone(), path)


def read_latent(path: str) -> torch.Tensor:
    """`retrieve_from_cache`: the payload is the latent tensor itself (dict payloads carry it under "latents")."""
    obj = torch.load(path, map_location="cpu", weights_only=False)
    return obj["latents"] if isinstance(obj, dict) else obj


class StagedBatch:
    def __init__(self, tensors: Dict[str, torch.Tensor], event, slot: int):
        self.tensors, self.event, self.slot = tensors, event, slot

    def wait(self) -> Dict[str, torch.Tensor]:
        """Make the consumer's CURRENT stream wait for the copies (device-side wait; the host does not block)."""
        if self.event is not None:
            torch.cuda.current_stream().wait_event(self.event)
        return self.tensors


class LatentStager:
    """Pinned ring of pre-stacked batches.  `stage(files, extras)` fills the next slot and starts its H2D copies;
    the returned StagedBatch is valid until `depth` further batches have been staged."""

    def __init__(self, device, depth: int = 3, workers: int = 8, dtype=torch.bfloat16):
        self.device = torch.device(device)
        self.depth, self.dtype = depth, dtype
        self.cuda = self.device.type == "cuda"
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self.stream = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._rings: Dict[Tuple, List[Dict[str, torch.Tensor]]] = {}
        self._dev: Dict[Tuple, List[Dict[str, torch.Tensor]]] = {}
        self._events: Dict[Tuple, List] = {}
        self._cursor: Dict[Tuple, int] = {}

    def _slot(self, key, shapes: Dict[str, Tuple[Tuple[int, ...], torch.dtype]]):
        if key not in self._rings:
            mk = lambda shp, dt: torch.empty(shp, dtype=dt, pin_memory=self.cuda)
            self._rings[key] = [{n: mk(shp, dt) for n, (shp, dt) in shapes.items()} for _ in range(self.depth)]
            self._dev[key] = [{n: torch.empty(shp, dtype=dt, device=self.device) for n, (shp, dt) in shapes.items()}
                              for _ in range(self.depth)]
            self._events[key] = [None] * self.depth
            self._cursor[key] = 0
        i = self._cursor[key]
        self._cursor[key] = (i + 1) % self.depth
        ev = self._events[key][i]
        if ev is not None:
            ev.synchronize()        # the H2D copy that last read this pinned slot is long done (depth batches ago)
        return i

    def stage(self, latent_files: Sequence[str], extras: Optional[Dict[str, torch.Tensor]] = None) -> StagedBatch:
        """latent_files: cache paths of ONE micro-batch (uniform shape, as collate's check_latent_shapes enforces).
        extras: already-stacked host tensors of the batch (text embeds, pooled embeds, masks) to ride the same slot."""
        first = read_latent(latent_files[0])
        B = len(latent_files)
        shapes = {"latent_batch": ((B, *first.shape), self.dtype)}
        for n, t in (extras or {}).items():
            shapes[n] = (tuple(t.shape), t.dtype)
        key = tuple((n, shp, str(dt)) for n, (shp, dt) in sorted(shapes.items()))
        i = self._slot(key, shapes)
        host = self._rings[key][i]

        def load(j):
            t = first if j == 0 else read_latent(latent_files[j])
            if tuple(t.shape) != tuple(first.shape):
                raise ValueError(f"latent shape mismatch inside a micro-batch: {tuple(t.shape)} vs {tuple(first.shape)} ({latent_files[j]})")
            host["latent_batch"][j].copy_(t)      # cast + write straight into the pinned batch row

        list(self.pool.map(load, range(B)))
        for n, t in (extras or {}).items():
            host[n].copy_(t)
        dev = self._dev[key][i]
        ev = None
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))   # the slot's previous consumer has been enqueued
            with torch.cuda.stream(self.stream):
                for n in host:
                    dev[n].copy_(host[n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self._events[key][i] = ev
        else:
            for n in host:
                dev[n].copy_(host[n])
        return StagedBatch(dict(dev), ev, i)

    def close(self):
        self.pool.shutdown(wait=True)
