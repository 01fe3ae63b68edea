"""CPU: EMA host mirror — decay schedule against the reference's own `get_decay` executed verbatim
(tests/golden/ema_golden.pt), and the foreach update / interval / copy_to semantics."""
from pathlib import Path

import torch

from simpletuner_b200.training.ema import EMAModel, should_update_ema

ROWS = torch.load(Path(__file__).parent / "golden" / "ema_golden.pt")


def test_decay_schedule_matches_reference():
    for row in ROWS:
        ema = EMAModel([torch.nn.Parameter(torch.zeros(1))], **row["cfg"])
        got = [ema.get_decay(s) for s in row["steps"]]
        assert got == row["decay"], (row["cfg"], got, row["decay"])


def test_update_rule_interval_and_copy_to():
    p = torch.nn.Parameter(torch.ones(4))
    frozen = torch.nn.Parameter(torch.full((2,), 3.0), requires_grad=False)
    ema = EMAModel([p, frozen], decay=0.5, ema_update_interval=2)
    with torch.no_grad():
        p.add_(1.0)
        frozen.add_(1.0)
    ema.step([p, frozen], global_step=1)                    # 1 % 2 != 0 -> skipped
    assert torch.equal(ema.shadow_params[0], torch.ones(4)) and ema.optimization_step == 0
    ema.step([p, frozen], global_step=2)                    # step = max(0, 2 - 0 - 1) = 1 -> decay = min(2/11, 0.5)
    d = 2 / 11
    assert ema.cur_decay_value == d
    assert torch.allclose(ema.shadow_params[0], torch.full((4,), 1.0 - (1 - d) * (1.0 - 2.0)))
    assert torch.equal(ema.shadow_params[1], torch.full((2,), 4.0))      # non-trainable tensors are copied
    q = torch.nn.Parameter(torch.zeros(4))
    ema.copy_to([q, torch.nn.Parameter(torch.zeros(2))])
    assert torch.equal(q.detach(), ema.shadow_params[0])
    assert should_update_ema(None, 7) and should_update_ema(5, 10) and not should_update_ema(5, 11)
