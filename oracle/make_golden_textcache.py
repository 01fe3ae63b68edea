"""Generates tests/golden/textcache_golden.pt by EXECUTING the reference's own methods (lifted verbatim from
simpletuner/helpers/caching/text_embeds.py with oracle/ref_extract.py): file-name hashing for the three key kinds and the
per-sample slicing of a batched encoder output.  Run here: `python -m oracle.make_golden_textcache`."""
import hashlib
import os
from enum import Enum
from pathlib import Path

import torch

from oracle import ref_extract as RX

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "textcache_golden.pt"


class TextEmbedCacheKey(Enum):       # helpers/models/common.py:398-401 (the lifted methods only compare members by identity)
    CAPTION = "caption"
    FILENAME = "filename"
    DATASET_AND_FILENAME = "dataset_and_filename"


def main():
    pathing = RX.functions("helpers/utils/pathing.py", ["canonicalize_data_uri"])
    ns = {"hashlib": hashlib, "os": os, "TextEmbedCacheKey": TextEmbedCacheKey, "PromptCacheRecord": dict,
          "canonicalize_data_uri": pathing["canonicalize_data_uri"]}
    Ref = RX.methods("helpers/caching/text_embeds.py", "TextEmbeddingCache",
                     ["_normalize_key_value", "create_hash", "_resolve_cache_key_value", "hash_prompt_with_path",
                      "_slice_batch_output_for_cache"], extra_ns=ns)
    records = [{"prompt": "a photo of a cat", "key": "a photo of a cat"}, {"prompt": "caption one", "key": "dataset-1:path/to/sample.png"},
               {"prompt": "caption two", "key": "dataset-1:webshart:/0/3/sample.mp4"}, {"prompt": "", "key": "__caption_dropout__"},
               {"prompt": "only a prompt"}, {"prompt": "x", "key": "relative/dir/../img.png"}, {"prompt": "y", "key": "s3://bucket/img.png"}]
    fx = {"records": records, "names": {}}
    for kt in TextEmbedCacheKey:
        for model_type in ("flux", "sd3"):
            c = Ref()
            c.key_type, c.model_type, c.cache_dir, c.model = kt, model_type, "/cache/text", None
            c._requires_path_based_keys = kt in (TextEmbedCacheKey.FILENAME, TextEmbedCacheKey.DATASET_AND_FILENAME)
            names = []
            for r in records:
                try:
                    names.append(c.hash_prompt_with_path(dict(r)))
                except ValueError as e:
                    names.append("ValueError")
            fx["names"][(kt.value, model_type)] = names
    g = torch.Generator().manual_seed(0)
    batch = {"prompt_embeds": torch.randn(3, 12, 8, generator=g), "pooled_prompt_embeds": torch.randn(3, 6, generator=g),
             "attention_mask": torch.tensor([[1] * 12, [1] * 5 + [0] * 7, [1] * 9 + [0] * 3]), "time_ids": None,
             "extra": torch.randn(3, 2, 2, generator=g)}
    c = Ref()
    c.model = None
    fx["slice_in"] = batch
    fx["slice_out"] = [c._slice_batch_output_for_cache(batch, i, 3) for i in range(3)]
    torch.save(fx, OUT)
    print("wrote", OUT, fx["names"][("caption", "flux")][:2])


if __name__ == "__main__":
    main()
