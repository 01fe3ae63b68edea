"""The compute half of `FluxPipeline.encode_prompt` (simpletuner/helpers/models/flux/pipeline.py:1139-1235, called by
Flux._encode_prompts, flux/model.py:497-520): token ids in, `(prompt_embeds, pooled_prompt_embeds, text_ids, attention_mask)`
out.  Tokenisation (pipeline.py:1065-1076, 1109-1120) is string processing and stays with the caller."""
from __future__ import annotations

from typing import Optional

import torch


@torch.no_grad()
def encode_token_ids(text_encoder, text_encoder_2, clip_input_ids: torch.Tensor, t5_input_ids: torch.Tensor,
                     t5_attention_mask: Optional[torch.Tensor] = None, num_images_per_prompt: int = 1, t5_padding: str = "unmodified"):
    """Returns (prompt_embeds [B n, S, 4096], pooled_prompt_embeds [B n, 768], text_ids [S, 3], masks) like encode_prompt;
    `t5_padding = "zero"` applies Flux._encode_prompts' zeroing of padded positions (flux/model.py:515-517)."""
    pooled = text_encoder(clip_input_ids, output_hidden_states=False).pooler_output
    pooled = pooled.to(dtype=text_encoder.dtype)
    B = pooled.shape[0]
    pooled = pooled.repeat(1, num_images_per_prompt).view(B * num_images_per_prompt, -1)
    embeds = text_encoder_2(t5_input_ids, output_hidden_states=False)[0].to(dtype=text_encoder_2.dtype)
    S = embeds.shape[1]
    embeds = embeds.repeat(1, num_images_per_prompt, 1).view(B * num_images_per_prompt, S, -1)
    if t5_padding == "zero" and t5_attention_mask is not None:
        m = t5_attention_mask.to(device=embeds.device).repeat_interleave(num_images_per_prompt, dim=0)
        embeds = embeds * m.unsqueeze(-1).expand(embeds.shape)
    text_ids = torch.zeros(S, 3, device=embeds.device, dtype=embeds.dtype)
    return embeds, pooled, text_ids, t5_attention_mask
