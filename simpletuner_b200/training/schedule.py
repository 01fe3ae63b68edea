"""Timestep / sigma sampling and schedule helpers of the training step (host-side logic).

Behavioural mirror of reference simpletuner/helpers/training/custom_schedule.py:18-106, 443-477 and
helpers/models/common.py:5062-5073, 5089-5090.  These are tiny index / RNG computations that the
reference runs with torch on the host or on [B]-sized device tensors; they stay torch here (random
draws must consume torch's generators exactly as the reference does, SURVEY.md §7 "RNG parity").
tests/test_schedule.py pins every function bit-exactly against vectors produced by the reference's
own source (tests/golden/flux_step_golden.pt).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch


def apply_flow_schedule_shift(config, noise_scheduler, sigmas: torch.Tensor, noise: Optional[torch.Tensor]) -> torch.Tensor:
    """custom_schedule.py:443-477.  Static shift when `flow_schedule_shift > 0`; otherwise the
    resolution-dependent shift exp(mu) with mu linear in the packed sequence length."""
    shift = None
    fs = getattr(config, "flow_schedule_shift", None)
    if fs is not None and fs > 0:
        shift = fs
    elif getattr(config, "flow_schedule_auto_shift", False):
        if noise.ndim == 5:
            num_frames, height, width = noise.shape[-3:]
        else:
            num_frames = 1
            height, width = noise.shape[-2:]
        sc = noise_scheduler.config
        patch_size = getattr(sc, "patch_size", 2)
        if patch_size is None or patch_size <= 0:
            patch_size = 2
        seq_len = num_frames * (height // patch_size) * (width // patch_size)
        mu = calculate_shift(seq_len, sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
        shift = math.exp(mu)
    if shift is not None:
        sigmas = (sigmas * shift) / (1 + (shift - 1) * sigmas)
    return sigmas


def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.15) -> float:
    """diffusers pipeline_flux.calculate_shift (imported at reference flux/__init__.py:5)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def generate_timestep_weights(args, num_timesteps: int) -> torch.Tensor:
    """custom_schedule.py:61-106."""
    weights = torch.ones(num_timesteps)
    num_to_bias = int(args.timestep_bias_portion * num_timesteps)
    strategy = args.timestep_bias_strategy
    if strategy == "later":
        bias_indices = slice(-num_to_bias, None)
    elif strategy == "earlier":
        bias_indices = slice(0, num_to_bias)
    elif strategy == "range":
        range_begin, range_end = args.timestep_bias_begin, args.timestep_bias_end
        if range_begin < 0:
            raise ValueError("When using the range strategy for timestep bias, you must provide a beginning timestep greater or equal to zero.")
        if range_end > num_timesteps:
            raise ValueError("When using the range strategy for timestep bias, you must provide an ending timestep smaller than the number of timesteps.")
        bias_indices = slice(range_begin, range_end)
    else:
        return weights
    if args.timestep_bias_multiplier <= 0:
        raise ValueError("The parameter --timestep_bias_multiplier is not intended to be used to disable the training of specific timesteps.")
    weights[bias_indices] *= args.timestep_bias_multiplier
    weights /= weights.sum()
    return weights


def segmented_timestep_selection(actual_num_timesteps: int, bsz: int, weights: torch.Tensor, config,
                                 use_refiner_range: bool = False) -> torch.Tensor:
    """custom_schedule.py:18-58, including quirk Q1: each segment's weights are
    normalised IN PLACE on a view of `weights`, and neighbouring segments share one boundary index, so
    the normalisation of segment i changes the boundary weight segment i+1 sees.  One multinomial
    draw per segment, in order, from the default generator of `weights.device`."""
    num_timesteps = actual_num_timesteps
    if use_refiner_range or config.refiner_training:
        if config.refiner_training_invert_schedule:
            start_timestep = actual_num_timesteps - 1
            end_timestep = int(config.refiner_training_strength * actual_num_timesteps)
        else:
            start_timestep = int(actual_num_timesteps * config.refiner_training_strength) - 1
            end_timestep = 0
        num_timesteps = start_timestep - end_timestep + 1
    else:
        start_timestep = actual_num_timesteps - 1
        end_timestep = 0
    segment_size = max(num_timesteps // bsz, 1)
    selected = []
    for i in range(bsz):
        start = start_timestep - i * segment_size
        end = max(start - segment_size, end_timestep) if i != bsz - 1 else end_timestep
        seg = weights[end:start + 1]
        seg /= seg.sum()  # in place, on the caller's tensor
        idx = torch.multinomial(seg, 1).item()
        selected.append(end + idx)
    return torch.tensor(selected)


def sample_flow_sigmas(config, noise_scheduler, bsz: int, noise: Optional[torch.Tensor], device,
                       timestep_offset: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """common.py:5062-5073, 5089-5090 — default branch (no custom list / beta / uniform / fast schedule):
    sigma = sigmoid(scale * (N(0,1) + offset)), schedule shift, t = 1000 sigma."""
    if getattr(config, "flow_use_uniform_schedule", False):
        sigmas = torch.rand((bsz,), device=device)
    elif getattr(config, "flow_use_beta_schedule", False) or getattr(config, "flux_fast_schedule", False):
        raise NotImplementedError("beta / fast flow schedules are not part of the B200 step yet")
    else:
        normal = torch.randn((bsz,), device=device)
        if timestep_offset:
            normal = normal + timestep_offset
        sigmas = torch.sigmoid(getattr(config, "flow_sigmoid_scale", 1.0) * normal)
    sigmas = apply_flow_schedule_shift(config, noise_scheduler, sigmas, noise)
    return sigmas, sigmas * 1000.0


# --------------------------------------------------------------------------------------------------
# full `sample_flow_sigmas` (every branch except the cubic-spline schedule), with the round-robin cursor state
# --------------------------------------------------------------------------------------------------
def normalize_flow_custom_timesteps(raw_value, device=None) -> Optional[torch.Tensor]:
    """common.py:4799-4838: comma/semicolon separated string, JSON list, array or tensor -> finite 1-D fp32 tensor."""
    import json

    if raw_value is None or (isinstance(raw_value, str) and raw_value in ("", "None")):
        return None
    candidate = raw_value
    if isinstance(candidate, str):
        stripped = candidate.strip()
        if stripped == "":
            return None
        try:
            candidate = json.loads(stripped)
        except Exception:
            segments = [seg for seg in stripped.replace(";", ",").split(",") if seg.strip()]
            try:
                candidate = [float(seg.strip()) for seg in segments]
            except Exception:
                return None
    if hasattr(candidate, "tolist") and not torch.is_tensor(candidate):
        candidate = candidate.tolist()
    try:
        tensor = torch.as_tensor(candidate, device=device, dtype=torch.float32).flatten()
    except Exception:
        return None
    if tensor.numel() == 0:
        return None
    finite = torch.isfinite(tensor)
    if not torch.all(finite):
        tensor = tensor[finite]
    return tensor if tensor.numel() else None


class FlowSigmaSampler:
    """Stateful mirror of `ModelFoundation.sample_flow_sigmas` (common.py:4994-5090) + the round-robin cursor helpers
    (`reset_flow_custom_timestep_cursor`).  `layout_fn(bsz)` returns an object with `global_batch_size` and
    `local_batch_offset` (training.dist.resolve_batch_layout by default), so ranks with different local batch sizes walk
    disjoint slices of the custom list exactly like the reference (tests/test_flow_custom_timesteps.py)."""

    def __init__(self, config, noise_scheduler, device, layout_fn=None):
        self.config, self.noise_scheduler, self.device = config, noise_scheduler, device
        if layout_fn is None:
            from .dist import resolve_batch_layout
            layout_fn = lambda bsz: resolve_batch_layout(bsz, device)
        self._layout_fn = layout_fn
        self._cursor: Optional[int] = None
        self._resume_step: Optional[int] = None
        self._warned = False

    def reset_cursor(self, global_step: Optional[int] = None):
        """common.py `reset_flow_custom_timestep_cursor`: forget the cursor; the next draw re-derives it from the step."""
        self._cursor = None
        self._resume_step = None if global_step is None else int(global_step)

    def sample(self, bsz: int, noise: Optional[torch.Tensor], state: Optional[dict] = None, timestep_offset: float = 0.0):
        import random

        c, dev = self.config, self.device
        state = state or {}
        if getattr(c, "mixflow_enabled", False) is True:
            sigmas = 1.0 - torch.sqrt(torch.rand((bsz,), device=dev))
            sigmas = apply_flow_schedule_shift(c, self.noise_scheduler, sigmas, noise)
            return sigmas, sigmas * 1000.0
        custom = normalize_flow_custom_timesteps(getattr(c, "flow_custom_timesteps", None), dev)
        if custom is not None:
            mode = str(getattr(c, "flow_timesteps_mode", "fixed-list") or "fixed-list").replace("_", "-")
            if mode not in {"fixed-list", "round-robin"}:
                raise ValueError("flow_timesteps_mode must be either 'fixed-list' or 'round-robin'.")
            if torch.max(custom) <= 1.0:     # values <= 1 are sigmas, otherwise timesteps in [0, 1000]
                base_sigmas = custom.clamp(0.0, 1.0)
                base_t = base_sigmas * 1000.0
            else:
                base_t = custom.clamp(0.0, 1000.0)
                base_sigmas = (base_t / 1000.0).clamp(0.0, 1.0)
            n = base_t.numel()
            if n == 1:
                return base_sigmas.expand(bsz), base_t.expand(bsz)
            if mode == "round-robin":
                layout = self._layout_fn(bsz)
                g = int(layout.global_batch_size)
                if n < g and not self._warned:
                    self._warned = True   # the reference logs a warning once: ranks may reuse entries on the same step
                if self._cursor is None:
                    completed = int(self._resume_step if self._resume_step is not None else state.get("global_step", 0) or 0)
                    self._cursor = (completed * g) % n
                    self._resume_step = None
                idx = (torch.arange(bsz, device=dev) + self._cursor + int(layout.local_batch_offset)) % n
                self._cursor = (self._cursor + g) % n
            else:
                idx = torch.randint(0, n, (bsz,), device=dev)
            return base_sigmas[idx], base_t[idx]
        if getattr(c, "flow_cubic_schedule", None) or getattr(c, "flow_cubic_schedule_weights", None):
            raise NotImplementedError("the cubic-spline flow schedule is not part of the B200 step")
        fast = getattr(c, "flux_fast_schedule", False)
        beta = getattr(c, "flow_use_beta_schedule", False)
        uniform = getattr(c, "flow_use_uniform_schedule", False)
        if not fast and not beta and not uniform:
            normal = torch.randn((bsz,), device=dev)
            if timestep_offset:
                normal = normal + timestep_offset
            sigmas = torch.sigmoid(getattr(c, "flow_sigmoid_scale", 1.0) * normal)
            sigmas = apply_flow_schedule_shift(c, self.noise_scheduler, sigmas, noise)
        elif uniform:
            sigmas = torch.rand((bsz,), device=dev)
            sigmas = apply_flow_schedule_shift(c, self.noise_scheduler, sigmas, noise)
        elif beta:
            from torch.distributions import Beta
            sigmas = Beta(c.flow_beta_schedule_alpha, c.flow_beta_schedule_beta).sample((bsz,)).to(device=dev)
            sigmas = apply_flow_schedule_shift(c, self.noise_scheduler, sigmas, noise)
        else:
            sigmas = torch.tensor(random.choices([1.0] * 7 + [0.75, 0.5, 0.25], k=bsz), device=dev)
        return sigmas, sigmas * 1000.0
