#!/usr/bin/env python
"""Time attention fwd / bwd at the Flux shape with the library named by $STB200_LIB (experiments)."""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from simpletuner_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


B, H, S, HD = 4, 24, 4608, 128
q, k, v, do = (torch.randn(B, S, H, HD, device="cuda").bfloat16() for _ in range(4))
o, lse = ops.attn_fwd(q, k, v)
dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse)
r = {"lib": os.environ.get("STB200_LIB", "default"), "fwd_ms": round(timeit(lambda: ops.attn_fwd(q, k, v, out=o)), 4),
     "bwd_ms": round(timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, dq=dq, dk=dk, dv=dv)), 4)}
prof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA])
with prof:
    for _ in range(3):
        ops.attn_bwd(q, k, v, o, do, lse, dq=dq, dk=dk, dv=dv)
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "attn" in e.key:
        r[e.key.split("<")[0].split("::")[-1]] = round(e.device_time_total / e.count / 1e3, 4)
print(json.dumps(r))
