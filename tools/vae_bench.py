#!/usr/bin/env python
"""Time the libstb200 VAE latent-encode path (Flux VAE config) on one GPU; prints a JSON line.
Optionally (--torch) also times the same encoder as plain torch bf16 channels_last convs (cuDNN) for reference."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def conv_flops(cfg, H, W):
    ch = cfg.block_out_channels
    fl = 2 * 27 * ch[0] * H * W
    prev, h, w = ch[0], H, W
    for i, c in enumerate(ch):
        for l in range(cfg.layers_per_block):
            cin = prev if l == 0 else c
            fl += 2 * 9 * cin * c * h * w + 2 * 9 * c * c * h * w + (2 * cin * c * h * w if cin != c else 0)
        prev = c
        if i != len(ch) - 1:
            h, w = h // 2, w // 2
            fl += 2 * 9 * c * c * h * w
    c = ch[-1]
    S = h * w
    fl += 4 * (2 * 9 * c * c * S) + 4 * 2 * c * c * S + 2 * 2 * S * S * c
    fl += 2 * 9 * c * 2 * cfg.latent_channels * S
    return fl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--torch", action="store_true")
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    from oracle import vae_oracle as O  # parameter initialiser only (tools/, not the product path)
    from tests.vae_parity import build_cuda_vae
    cfg = O.VaeConfig()
    P = O.init_vae_params(cfg)
    vae = build_cuda_vae(cfg, P)
    x = (torch.rand(a.batch, 3, a.res, a.res, device="cuda") * 2 - 1).bfloat16()
    eps = torch.randn(a.batch, 16, a.res // 8, a.res // 8, device="cuda").bfloat16()
    for _ in range(2):
        z = vae.encode_scaled(x, eps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        z = vae.encode_scaled(x, eps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    fl = conv_flops(cfg, a.res, a.res) * a.batch
    out = {"workload": f"flux_vae_encode_b{a.batch}_{a.res}", "ms": ms, "images_per_sec": a.batch / ms * 1e3,
           "tflops": fl / ms / 1e9, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
           "finite": bool(torch.isfinite(z).all())}
    if a.profile:
        prof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA])
        with prof:
            vae.encode_scaled(x, eps)
            torch.cuda.synchronize()
        rows = sorted(((e.device_time_total / 1e3, e.count, e.key) for e in prof.key_averages()), reverse=True)
        out["kernels_ms"] = [[round(t, 3), c, k[:70]] for t, c, k in rows[:14]]
    if a.torch:
        import torch.nn.functional as F
        Pc = {k: v.cuda().bfloat16() for k, v in P.items()}
        Pc = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in Pc.items()}
        xc = x.contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            for _ in range(2):
                O.vae_encode_moments(Pc, cfg, xc)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(a.iters):
                O.vae_encode_moments(Pc, cfg, xc)
            e1.record()
            torch.cuda.synchronize()
        out["torch_cudnn_ms"] = e0.elapsed_time(e1) / a.iters
    print(json.dumps(out))


if __name__ == "__main__":
    main()
