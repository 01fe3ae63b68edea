"""Text-embed cache wire format — the other half of SURVEY.md 8f rank 3 (the latent half is training/latent_cache.py).

Host mirror of the parts of reference simpletuner/helpers/caching/text_embeds.py that define WHAT is on disk and under which
name, so that embeddings produced by the libstb200 text encoders (simpletuner_b200.text) land in — and are read back from —
the reference's own cache:

  * file name: md5(normalised key [+ "\\0prompt\\0" + prompt for path-keyed models]).hexdigest() + "-" + model_type + ".pt"
    (`create_hash`, `_normalize_key_value`, `_resolve_cache_key_value`, `hash_prompt_with_path`: text_embeds.py:126-185;
    `canonicalize_data_uri`: helpers/utils/pathing.py:5-9; key kinds: helpers/models/common.py:398-401)
  * one `.pt` per prompt holding the model's formatted dict (Flux: prompt_embeds / pooled_prompt_embeds / time_ids /
    attention_masks, flux/model.py:413-431), sliced out of the encoder's batch by `_slice_batch_output_for_cache`
    (text_embeds.py:226-264: per-sample slices, pooled vectors squeezed, prompt / mask trimmed to the mask's true length)
  * `load_from_cache` (text_embeds.py:445-460): old tuple files are converted through the model's `_format_text_embedding`.
"""
from __future__ import annotations

import hashlib
import os
from enum import Enum
from typing import Any, Callable, Dict, Optional

import torch


class TextEmbedCacheKey(Enum):
    CAPTION = "caption"
    FILENAME = "filename"
    DATASET_AND_FILENAME = "dataset_and_filename"


def canonicalize_data_uri(path: str) -> str:
    """Repair URI identifiers that pathlib normalises to a single slash (helpers/utils/pathing.py:5-9)."""
    if path.startswith("webshart:/") and not path.startswith("webshart://"):
        return f"webshart://{path[len('webshart:/'):]}"
    return path


def normalize_key_value(key_value, key_type: TextEmbedCacheKey = TextEmbedCacheKey.CAPTION) -> str:
    if key_value is None:
        return ""
    normalized = str(key_value)
    if key_type is TextEmbedCacheKey.FILENAME:
        if "://" not in normalized:
            normalized = os.path.normcase(os.path.abspath(os.path.normpath(normalized)))
    elif key_type is TextEmbedCacheKey.DATASET_AND_FILENAME:
        dataset_id, separator, data_path = normalized.partition(":")
        if separator:
            normalized = f"{dataset_id}:{canonicalize_data_uri(data_path)}"
    return normalized


def requires_path_based_keys(key_type: TextEmbedCacheKey) -> bool:
    return key_type in (TextEmbedCacheKey.FILENAME, TextEmbedCacheKey.DATASET_AND_FILENAME)


def create_hash(key_value, model_type: str, prompt=None, key_type: TextEmbedCacheKey = TextEmbedCacheKey.CAPTION) -> str:
    md5_hash = hashlib.md5()
    md5_hash.update(str(normalize_key_value(key_value, key_type)).encode())
    if requires_path_based_keys(key_type) and prompt:       # path keys are shared by multi-line caption alternatives
        md5_hash.update(b"\0prompt\0")
        md5_hash.update(str(prompt).encode())
    return md5_hash.hexdigest() + f"-{model_type}"


def resolve_cache_key_value(prompt_record: Dict[str, Any], key_type: TextEmbedCacheKey = TextEmbedCacheKey.CAPTION) -> str:
    if "key" in prompt_record:              # (empty strings are valid: dropout captions)
        return prompt_record["key"]
    if requires_path_based_keys(key_type):
        raise ValueError("Prompt record is missing 'key' but model requires filename-based text embeddings. "
                         f"Record metadata: {prompt_record.get('metadata')} prompt={prompt_record.get('prompt')}")
    if "prompt" in prompt_record:
        return prompt_record["prompt"]
    raise ValueError("Prompt record is missing both 'key' and 'prompt' values.")


def cache_filename(prompt_record: Dict[str, Any], cache_dir: str, model_type: str,
                   key_type: TextEmbedCacheKey = TextEmbedCacheKey.CAPTION) -> str:
    """`TextEmbeddingCache.hash_prompt_with_path`."""
    key_value = resolve_cache_key_value(prompt_record, key_type)
    return os.path.join(cache_dir, create_hash(key_value, model_type, prompt=prompt_record.get("prompt"), key_type=key_type) + ".pt")


def slice_batch_output_for_cache(text_encoder_output: Dict[str, Any], batch_index: int, batch_size: int) -> Dict[str, Any]:
    """`TextEmbeddingCache._slice_batch_output_for_cache` (the default slicer)."""
    per_sample: Dict[str, Any] = {}
    attention_mask = text_encoder_output.get("attention_mask")
    true_length = None
    if isinstance(attention_mask, torch.Tensor):
        if attention_mask.ndim < 2 or attention_mask.shape[0] != batch_size:
            raise ValueError(f"Batched attention_mask shape {tuple(attention_mask.shape)} does not match batch size {batch_size}.")
        true_length = int(attention_mask[batch_index].sum().item())
    for key, value in text_encoder_output.items():
        if not isinstance(value, torch.Tensor):
            per_sample[key] = value
            continue
        if value.ndim > 0 and value.shape[0] != batch_size and batch_size == 1:
            value = value.unsqueeze(0)
        if value.shape[0] != batch_size:
            raise ValueError(f"Batched text encoder output '{key}' shape {tuple(value.shape)} does not match batch size {batch_size}.")
        sample = value[batch_index:batch_index + 1]
        if key in {"pooled_prompt_embeds", "negative_pooled_prompt_embeds"} and sample.ndim == 2:
            sample = sample.squeeze(0)
        if true_length is not None and key in {"prompt_embeds", "attention_mask"} and sample.ndim >= 2:
            sample = sample[:, :true_length]
        per_sample[key] = sample.clone().contiguous()
    return per_sample


def format_text_embedding_flux(text_embedding) -> Dict[str, Any]:
    """`Flux._format_text_embedding` (flux/model.py:413-431): the tuple `_encode_prompts` returns -> the cached dict."""
    prompt_embeds, pooled_prompt_embeds, time_ids, masks = text_embedding
    return {"prompt_embeds": prompt_embeds, "pooled_prompt_embeds": pooled_prompt_embeds.squeeze(0), "time_ids": time_ids,
            "attention_masks": masks}


def format_text_embedding_pooled(text_embedding) -> Dict[str, Any]:
    """`SD3._format_text_embedding` / `SDXL._format_text_embedding` (sd3/model.py:369-385, sdxl/model.py:117-133)."""
    prompt_embeds, pooled_prompt_embeds = text_embedding
    return {"prompt_embeds": prompt_embeds, "pooled_prompt_embeds": pooled_prompt_embeds.squeeze(0)}


def format_text_embedding_pixart(text_embedding) -> Dict[str, Any]:
    """`PixartSigma._format_text_embedding` (pixart/model.py:176-192)."""
    prompt_embeds, prompt_attention_mask = text_embedding
    return {"prompt_embeds": prompt_embeds, "attention_mask": prompt_attention_mask.squeeze(0)}


FORMATTERS = {"flux": format_text_embedding_flux, "sd3": format_text_embedding_pooled, "sdxl": format_text_embedding_pooled,
              "pixart_sigma": format_text_embedding_pixart}


def write_text_embeds(filename: str, embeddings: Dict[str, Any]) -> None:
    """`save_to_cache` -> `data_backend.torch_save`: one torch.save per prompt."""
    os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
    torch.save({k: (v.detach().to("cpu") if isinstance(v, torch.Tensor) else v) for k, v in embeddings.items()}, filename)


def read_text_embeds(filename: str, format_tuple: Optional[Callable] = format_text_embedding_flux) -> Dict[str, Any]:
    """`load_from_cache`: dict files as they are; old tuple files through the model's `_format_text_embedding`."""
    result = torch.load(filename, map_location="cpu", weights_only=False)
    if isinstance(result, tuple):
        if format_tuple is None:
            raise ValueError(f"{filename} holds the old tuple format and no formatter was given")
        result = format_tuple(result)
    return result


@torch.no_grad()
def encode_and_cache(prompt_records, clip_ids, t5_ids, t5_masks, text_encoder, text_encoder_2, cache_dir: str,
                     model_type: str = "flux", key_type: TextEmbedCacheKey = TextEmbedCacheKey.CAPTION):
    """One batch of `_encode_and_cache_prompt_batch` (text_embeds.py:266-301) with the libstb200 encoders: encode the token
    ids, format like Flux, slice per prompt and write one file each.  Returns the file names."""
    from ..text import encode_token_ids
    embeds, pooled, _text_ids, masks = encode_token_ids(text_encoder, text_encoder_2, clip_ids, t5_ids, t5_masks)
    B = embeds.shape[0]
    out = {"prompt_embeds": embeds, "pooled_prompt_embeds": pooled, "time_ids": None,
           "attention_masks": masks if masks is not None else torch.ones(embeds.shape[:2], dtype=torch.long)}
    files = []
    for i, rec in enumerate(prompt_records):
        fn = cache_filename(rec, cache_dir, model_type, key_type)
        write_text_embeds(fn, slice_batch_output_for_cache(out, i, B))
        files.append(fn)
    return files


def collate_tensors(tensors):
    """`compute_prompt_embeddings._collate_tensors` (helpers/training/collate.py:409-451): 2-D [seq, dim] and 1-D entries are
    stacked, 3-D [1, seq, dim] entries (cached with their batch dimension) concatenated, mixed ranks normalised to 3-D."""
    tensors = [t for t in tensors if t is not None]
    if not tensors:
        return None
    dims = tensors[0].dim()
    all_same = all(t.dim() == dims for t in tensors)
    if dims == 2:
        return torch.stack(tensors)
    if dims == 3 and all_same:
        return torch.cat(tensors, dim=0)
    if dims == 1:
        return torch.stack(tensors)
    normalized = []
    for t in tensors:
        if t.dim() in (1, 2):
            normalized.append(t.unsqueeze(0))
        elif t.dim() == 3:
            normalized.append(t)
        else:
            raise ValueError(f"Unexpected tensor dimension: {t.dim()} with shape {t.shape}")
    return torch.cat(normalized, dim=0)


def collate_prompt_embeds(text_encoder_output) -> Dict[str, Any]:
    """The default branch of `compute_prompt_embeddings` (collate.py:453-483): per-prompt cache dicts -> the batch the step
    consumes (`prompt_embeds`, `pooled_prompt_embeds` -> `add_text_embeds` downstream, `attention_masks`, `time_ids`)."""
    first = text_encoder_output[0]
    out: Dict[str, Any] = {}
    if "prompt_embeds" in first:
        out["prompt_embeds"] = collate_tensors([t["prompt_embeds"] for t in text_encoder_output])
    if "pooled_prompt_embeds" in first:
        out["pooled_prompt_embeds"] = collate_tensors([t["pooled_prompt_embeds"] for t in text_encoder_output])
    for old in ("attention_mask", "prompt_attention_mask", "attention_masks"):       # old styles first, the new key wins
        if old in first:
            out["attention_masks"] = collate_tensors([t[old] for t in text_encoder_output])
    if "time_ids" in first:
        out["time_ids"] = collate_tensors([t["time_ids"] for t in text_encoder_output])
    if not out:
        raise Exception(f"Could not compute text encoder output: {text_encoder_output}")
    return out
