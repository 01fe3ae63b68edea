#!/usr/bin/env python
"""Launch a handful of isolated hot kernels (for `ncu --set full -k regex:...`)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from simpletuner_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
torch.manual_seed(0)
if which == "gemm":
    M, N, K = 16384, 3072, 3072
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    for _ in range(4):
        ops.gemm([a], [w])
elif which == "attn":
    q, k, v, do = (torch.randn(1, 4608, 24, 128, device="cuda").bfloat16() for _ in range(4))
    for _ in range(3):
        o, lse = ops.attn_fwd(q, k, v)
        ops.attn_bwd(q, k, v, o, do, lse)
torch.cuda.synchronize()
print("done", which)
