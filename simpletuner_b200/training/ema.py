"""Exponential moving average of the trainable parameters — host-side mirror of reference
simpletuner/helpers/training/ema.py (class EMAModel): `get_decay` (:321-349, fixed decay after `warmup_steps`, the
`(1 + step) / (10 + step)` ramp, or the `1 - (1 + step / inv_gamma) ** -power` warm-up, clamped to [min_decay, decay]),
`should_update_ema` (:29-38) and the foreach update `shadow -= (1 - decay) * (shadow - param)` (:352-420).
SURVEY.md §8f rank 1 lists it next to the optimizer; it runs on 26 M LoRA parameters after the optimizer step (outside
the benchmarked step unless enabled), as two torch foreach kernels — there is no activation-sized work here.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch


def should_update_ema(ema_update_interval: Optional[int], step: int) -> bool:
    return True if ema_update_interval is None else step % ema_update_interval == 0


class EMAModel:
    def __init__(self, parameters: Iterable[torch.nn.Parameter], decay: float = 0.9999, min_decay: float = 0.0,
                 update_after_step: int = 0, warmup_steps: int = 0, use_ema_warmup: bool = False, inv_gamma: float = 1.0,
                 power: float = 2 / 3, ema_update_interval: Optional[int] = None):
        params = list(parameters)
        self.shadow_params: List[torch.Tensor] = [p.detach().clone() for p in params]
        self._tracked = params        # identity of the tracked tensors: lets the optimizer kernel own the update (fused_plan)
        self.decay, self.min_decay = decay, min_decay
        self.update_after_step = update_after_step
        self.warmup_steps = max(0, int(warmup_steps))
        self.use_ema_warmup, self.inv_gamma, self.power = use_ema_warmup, inv_gamma, power
        self.ema_update_interval = ema_update_interval
        self.optimization_step = 0
        self.cur_decay_value: Optional[float] = None

    def get_decay(self, optimization_step: Optional[int] = None) -> float:
        if optimization_step is None:
            optimization_step = self.optimization_step
        step = max(0, optimization_step - self.update_after_step - 1)
        if self.warmup_steps > 0:   # copy weights through the warm-up, then the configured fixed decay
            return 0.0 if optimization_step < self.warmup_steps else self.decay
        if step <= 0:
            return 0.0
        if self.use_ema_warmup:
            cur = 1 - (1 + step / self.inv_gamma) ** -self.power
        else:
            cur = (1 + step) / (10 + step)
        return max(min(cur, self.decay), self.min_decay)

    def begin_step(self, global_step: Optional[int] = None) -> Optional[float]:
        """Bookkeeping half of `step`: advances the counter and returns this update's decay, or None when
        `ema_update_interval` skips it.  Used by `step` and by AdamWBF16's fused update (the kernel applies the arithmetic)."""
        if global_step is not None and not should_update_ema(self.ema_update_interval, global_step):
            return None   # (global_step None = "always update": the reference would raise TypeError on `None % int`)
        if global_step is not None:      # periodic updates: the counter cannot be trusted (ema.py:381-385)
            self.optimization_step = global_step
        else:
            self.optimization_step += 1
        decay = self.get_decay(self.optimization_step)
        self.cur_decay_value = decay
        return decay

    def fused_plan(self, params: List[torch.Tensor]) -> Optional[List[torch.Tensor]]:
        """Shadows in the order of `params` when the optimizer kernel can own the update: `params` are exactly the tracked
        tensors, all trainable, with contiguous bf16 CUDA shadows.  None = use `step`."""
        by_id = {id(p): s for p, s in zip(self._tracked, self.shadow_params)}
        if len(params) != len(by_id) or any(id(p) not in by_id for p in params):
            return None
        out = [by_id[id(p)] for p in params]
        ok = all(p.requires_grad and s.is_cuda and s.dtype == torch.bfloat16 and s.is_contiguous() and s.shape == p.shape
                 for p, s in zip(params, out))
        return out if ok else None

    @torch.no_grad()
    def step(self, parameters: Iterable[torch.nn.Parameter], global_step: Optional[int] = None):
        params = list(parameters)
        if len(params) != len(self.shadow_params):
            raise RuntimeError(f"EMA tracks {len(self.shadow_params)} parameters but {len(params)} were given.")
        decay = self.begin_step(global_step)
        if decay is None:
            return
        self.apply(decay, params)

    @torch.no_grad()
    def apply(self, decay: float, params: Optional[List[torch.Tensor]] = None) -> None:
        """The arithmetic half of `step` for an already-determined decay (default: the tensors given at construction)."""
        params = self._tracked if params is None else params
        train = [(s, p) for s, p in zip(self.shadow_params, params) if p.requires_grad]
        frozen = [(s, p) for s, p in zip(self.shadow_params, params) if not p.requires_grad]
        if frozen:
            torch._foreach_copy_([s for s, _ in frozen], [p for _, p in frozen], non_blocking=True)
        if train:
            shadows, ps = [s for s, _ in train], [p.detach() for _, p in train]
            torch._foreach_sub_(shadows, torch._foreach_sub(shadows, ps), alpha=1 - decay)

    @torch.no_grad()
    def copy_to(self, parameters: Iterable[torch.nn.Parameter]) -> None:
        params = list(parameters)
        torch._foreach_copy_([p.data for p in params], [s.to(p.device) for s, p in zip(self.shadow_params, params)])

    def state_dict(self):
        return {"decay": self.decay, "min_decay": self.min_decay, "optimization_step": self.optimization_step,
                "update_after_step": self.update_after_step, "warmup_steps": self.warmup_steps,
                "use_ema_warmup": self.use_ema_warmup, "inv_gamma": self.inv_gamma, "power": self.power,
                "shadow_params": self.shadow_params}

    def load_state_dict(self, state_dict: dict) -> None:
        """ema.py `load_state_dict`: restores the schedule fields, the step counter and the shadow weights (resume)."""
        for k in ("decay", "min_decay", "optimization_step", "update_after_step", "warmup_steps", "use_ema_warmup", "inv_gamma",
                  "power"):
            if k in state_dict:
                setattr(self, k, state_dict[k])
        shadow = state_dict.get("shadow_params")
        if shadow is not None:
            shadow = list(shadow)
            if len(shadow) != len(self.shadow_params):
                raise ValueError(f"EMA state has {len(shadow)} shadow tensors, this model tracks {len(self.shadow_params)}.")
            with torch.no_grad():
                for dst, src in zip(self.shadow_params, shadow):
                    if dst.shape != src.shape:
                        raise ValueError(f"EMA shadow shape mismatch: {tuple(src.shape)} vs {tuple(dst.shape)}")
                    dst.copy_(src.to(device=dst.device, dtype=dst.dtype))

    def to(self, device=None, dtype=None, non_blocking: bool = False) -> "EMAModel":
        """ema.py `to`: move (and optionally cast the floating-point) shadow weights."""
        self.shadow_params = [
            s.to(device=device, dtype=dtype, non_blocking=non_blocking) if s.is_floating_point()
            else s.to(device=device, non_blocking=non_blocking) for s in self.shadow_params]
        return self
