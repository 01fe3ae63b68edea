"""View-only layout helpers of the Flux wrapper (index math; no arithmetic).

Behavioural mirror of reference simpletuner/helpers/models/flux/__init__.py:25-44; pinned bit-exactly
against the reference's own source in tests/test_schedule.py.  The training step itself never calls
these: the patchify / unpatchify index math is folded into the flow_prep_pack / flow_mse_loss kernels.
"""
from __future__ import annotations

import torch


def pack_latents(latents: torch.Tensor, batch_size: int, num_channels_latents: int, height: int, width: int) -> torch.Tensor:
    """[B, C, H, W] -> [B, (H/2)(W/2), 4C], token = (row, col) patch, channel = ((c*2 + dy)*2 + dx)."""
    x = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
    return x.permute(0, 2, 4, 1, 3, 5).reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)


def unpack_latents(latents: torch.Tensor, height: int, width: int, vae_scale_factor: int) -> torch.Tensor:
    """Inverse of pack_latents; (height, width) are pixel sizes and vae_scale_factor the pixel/patch ratio
    (the wrapper passes latent*8 and 16, reference flux/model.py:856-861)."""
    b, _, ch = latents.shape
    h, w = height // vae_scale_factor, width // vae_scale_factor
    x = latents.view(b, h, w, ch // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
    return x.reshape(b, ch // 4, h * 2, w * 2)
