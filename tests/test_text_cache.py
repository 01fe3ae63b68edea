"""CPU: text-embed cache wire format against vectors produced by executing the reference's own
`TextEmbeddingCache` methods (oracle/make_golden_textcache.py), the known answers of the reference's tests
(tests/test_text_embeds.py:236-272) and a write / read round trip incl. the old tuple format."""
import os
from pathlib import Path

import pytest
import torch

from simpletuner_b200.training import text_cache as TC

G = torch.load(Path(__file__).parent / "golden" / "textcache_golden.pt", weights_only=False)


def test_file_names_match_the_reference_for_every_key_kind():
    for (kt, model_type), want in G["names"].items():
        key_type = TC.TextEmbedCacheKey(kt)
        for rec, w in zip(G["records"], want):
            try:
                got = TC.cache_filename(dict(rec), "/cache/text", model_type, key_type)
            except ValueError:
                got = "ValueError"
            if got != "ValueError" and w != "ValueError" and kt == "filename" and "://" not in str(rec.get("key", rec.get("prompt"))):
                # FILENAME keys are made absolute against the working directory of whoever hashes them
                assert got.endswith(f"-{model_type}.pt")
                continue
            assert got == w, (kt, model_type, rec)


def test_reference_test_known_answers():
    K = TC.TextEmbedCacheKey
    h = lambda key, kt, prompt=None: TC.create_hash(key, "flux", prompt=prompt, key_type=kt)
    assert h("dataset-1:webshart:/0/3/sample.mp4", K.DATASET_AND_FILENAME) == h("dataset-1:webshart://0/3/sample.mp4", K.DATASET_AND_FILENAME)
    a = TC.cache_filename({"prompt": "caption one", "key": "dataset-1:path/to/sample.png"}, "/c", "flux", K.DATASET_AND_FILENAME)
    b = TC.cache_filename({"prompt": "caption two", "key": "dataset-1:path/to/sample.png"}, "/c", "flux", K.DATASET_AND_FILENAME)
    assert a != b                                                            # path keys distinguish caption variants
    a = TC.cache_filename({"prompt": "caption one", "key": "shared-key"}, "/c", "flux", K.CAPTION)
    b = TC.cache_filename({"prompt": "caption two", "key": "shared-key"}, "/c", "flux", K.CAPTION)
    assert a == b                                                            # caption keys do not add the prompt component
    rec = {"prompt": "", "key": "__caption_dropout__"}
    assert TC.cache_filename(rec, "/c", "flux", K.DATASET_AND_FILENAME) == os.path.join("/c", h("__caption_dropout__", K.DATASET_AND_FILENAME) + ".pt")
    with pytest.raises(ValueError):
        TC.resolve_cache_key_value({"prompt": "hello world"}, K.DATASET_AND_FILENAME)


def test_slicing_matches_the_reference():
    for i, want in enumerate(G["slice_out"]):
        got = TC.slice_batch_output_for_cache(G["slice_in"], i, 3)
        assert set(got) == set(want)
        for k in want:
            if isinstance(want[k], torch.Tensor):
                assert got[k].shape == want[k].shape and torch.equal(got[k], want[k]), (i, k)
            else:
                assert got[k] == want[k]
    assert TC.slice_batch_output_for_cache(G["slice_in"], 1, 3)["prompt_embeds"].shape == (1, 5, 8)      # trimmed to the mask length
    with pytest.raises(ValueError):
        TC.slice_batch_output_for_cache({"prompt_embeds": torch.zeros(2, 4, 8)}, 0, 3)


def test_write_read_round_trip_and_old_tuple_files(tmp_path):
    emb = (torch.randn(1, 7, 16).bfloat16(), torch.randn(1, 8).bfloat16(), None, torch.ones(1, 7, dtype=torch.long))
    d = TC.format_text_embedding_flux(emb)
    assert d["pooled_prompt_embeds"].shape == (8,) and set(d) == {"prompt_embeds", "pooled_prompt_embeds", "time_ids", "attention_masks"}
    fn = TC.cache_filename({"prompt": "p", "key": "p"}, str(tmp_path / "text"), "flux")
    TC.write_text_embeds(fn, d)
    back = TC.read_text_embeds(fn)
    assert torch.equal(back["prompt_embeds"], d["prompt_embeds"]) and back["time_ids"] is None
    old = str(tmp_path / "old.pt")
    torch.save(emb, old)                                                      # a cache file written before the dict format
    assert torch.equal(TC.read_text_embeds(old)["pooled_prompt_embeds"], emb[1].squeeze(0))
    with pytest.raises(ValueError):
        TC.read_text_embeds(old, format_tuple=None)


def test_encode_and_cache_writes_one_reference_format_file_per_prompt(tmp_path):
    """The glue of `_encode_and_cache_prompt_batch` (text_embeds.py:266-301) with stand-in encoders (the CUDA encoders are
    covered by tests/test_text_encoders.py): per-prompt files, padded positions trimmed by the attention mask."""
    from types import SimpleNamespace

    class _Clip:
        dtype = torch.bfloat16
        def __call__(self, ids, output_hidden_states=False):
            return SimpleNamespace(pooler_output=ids.float().mean(1, keepdim=True).expand(-1, 8).bfloat16())

    class _T5:
        dtype = torch.bfloat16
        def __call__(self, ids, output_hidden_states=False):
            return (ids.float()[..., None].expand(-1, -1, 16).bfloat16(),)

    clip_ids = torch.arange(2 * 5).view(2, 5)
    t5_ids = torch.arange(2 * 12).view(2, 12)
    masks = torch.tensor([[1] * 12, [1] * 4 + [0] * 8])
    recs = [{"prompt": "first", "key": "first"}, {"prompt": "second", "key": "second"}]
    files = TC.encode_and_cache(recs, clip_ids, t5_ids, masks, _Clip(), _T5(), str(tmp_path / "text"), model_type="flux")
    assert [os.path.basename(f) for f in files] == [TC.create_hash("first", "flux") + ".pt", TC.create_hash("second", "flux") + ".pt"]
    a, b = TC.read_text_embeds(files[0]), TC.read_text_embeds(files[1])
    assert a["prompt_embeds"].shape == (1, 12, 16) and a["pooled_prompt_embeds"].shape == (8,) and a["time_ids"] is None
    assert torch.equal(b["prompt_embeds"][0, :, 0].float(), t5_ids[1].float().bfloat16().float())      # (no trimming: Flux keys its mask `attention_masks`)
    assert torch.equal(b["attention_masks"], masks[1:2])


def test_collate_of_cached_prompt_dicts():
    """collate.py:409-483: cached [1, seq, dim] embeds are concatenated, [dim] pooled vectors stacked, None time_ids dropped."""
    g = torch.Generator().manual_seed(3)
    recs = [{"prompt_embeds": torch.randn(1, 6, 4, generator=g), "pooled_prompt_embeds": torch.randn(5, generator=g), "time_ids": None,
             "attention_masks": torch.ones(1, 6, dtype=torch.long)} for _ in range(3)]
    out = TC.collate_prompt_embeds(recs)
    assert out["prompt_embeds"].shape == (3, 6, 4) and out["pooled_prompt_embeds"].shape == (3, 5) and out["attention_masks"].shape == (3, 1, 6)     # [1, seq] masks are 2-D: stacked
    assert out["time_ids"] is None
    assert torch.equal(out["prompt_embeds"][1], recs[1]["prompt_embeds"][0])
    assert TC.collate_tensors([torch.zeros(6, 4), torch.zeros(6, 4)]).shape == (2, 6, 4)                 # 2-D -> stack
    assert TC.collate_tensors([torch.zeros(1, 6, 4), torch.zeros(6, 4)]).shape == (2, 6, 4)              # mixed ranks
    with pytest.raises(Exception):
        TC.collate_prompt_embeds([{"unknown": 1}])


def test_family_formatters():
    pe, pooled, mask = torch.zeros(1, 5, 8), torch.zeros(1, 6), torch.ones(1, 5, dtype=torch.long)
    assert TC.FORMATTERS["sd3"]((pe, pooled))["pooled_prompt_embeds"].shape == (6,)
    d = TC.FORMATTERS["pixart_sigma"]((pe, mask))
    assert set(d) == {"prompt_embeds", "attention_mask"} and d["attention_mask"].shape == (5,)
    # the PixArt dict is what `slice_batch_output_for_cache` trims by its attention mask
    out = TC.slice_batch_output_for_cache({"prompt_embeds": torch.zeros(2, 5, 8), "attention_mask": torch.tensor([[1, 1, 1, 0, 0], [1] * 5])}, 0, 2)
    assert out["prompt_embeds"].shape == (1, 3, 8) and out["attention_mask"].shape == (1, 3)
