"""Generate tests/golden/ema_golden.pt: the reference's own `EMAModel.get_decay` (helpers/training/ema.py:321-349)
executed verbatim over a grid of configurations and steps.  TEST INFRASTRUCTURE ONLY.   python -m oracle.make_golden_ema
"""
from pathlib import Path

import torch

from . import ref_extract as rx

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "ema_golden.pt"


def main():
    assert rx.available()
    Cls = rx.methods("helpers/training/ema.py", "EMAModel", ["get_decay"])
    rows = []
    for cfg in (dict(decay=0.9999, min_decay=0.0, update_after_step=0, warmup_steps=0, use_ema_warmup=False, inv_gamma=1.0, power=2 / 3),
                dict(decay=0.999, min_decay=0.5, update_after_step=5, warmup_steps=0, use_ema_warmup=False, inv_gamma=1.0, power=2 / 3),
                dict(decay=0.9999, min_decay=0.0, update_after_step=0, warmup_steps=0, use_ema_warmup=True, inv_gamma=1.0, power=2 / 3),
                dict(decay=0.99, min_decay=0.0, update_after_step=2, warmup_steps=0, use_ema_warmup=True, inv_gamma=2.0, power=0.75),
                dict(decay=0.995, min_decay=0.0, update_after_step=0, warmup_steps=10, use_ema_warmup=False, inv_gamma=1.0, power=2 / 3)):
        m = Cls.__new__(Cls)
        for k, v in cfg.items():
            setattr(m, k, v)
        m.optimization_step = 0
        steps = [0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 50, 1000, 31623, 1000000]
        rows.append({"cfg": cfg, "steps": steps, "decay": [float(m.get_decay(s)) for s in steps]})
    torch.save(rows, OUT)
    print(f"wrote {OUT} ({len(rows)} configurations)")


if __name__ == "__main__":
    main()
