"""Latent cache wire format (helpers/caching/vae.py:678-703, 1398-1449) and the pinned staging reader.
CPU: file naming against the known answers of the reference's own test (tests/test_vae.py:50-109) and the round trip through
the staging ring; GPU: the staged device tensors equal the stacked file contents while the copy overlaps a compute stream."""
import os
from hashlib import sha256

import pytest
import torch

from simpletuner_b200.training import latent_cache as LC


def test_filename_mapping_known_answers():
    h = lambda s: sha256(s.encode()).hexdigest()
    cases = [
        ("/data/image1.pt", "/data", "/data/image1.pt"),
        ("/data/image1.png", "/data", "cache/" + h("image1") + ".pt"),
        ("/data/subdir1/subdir2/image2.jpg", "/data", "cache/subdir1/subdir2/" + h("image2") + ".pt"),
        ("data/subdir1/subdir2/image2.jpg", "data", "cache/subdir1/subdir2/" + h("image2") + ".pt"),
        ("/anotherdir/image3.png", None, "cache/" + h("image3") + ".pt"),
        ("/data/image4.png", None, "cache/" + h("image4") + ".pt"),
        ("/image5.png", "/data", "cache/" + h("image5") + ".pt"),
        ("/data/image6.png", "/data", "cache/" + h("image6") + ".pt"),
    ]
    for path, inst, want in cases:
        assert LC.generate_vae_cache_filename(path, "cache", inst, True)[0] == want, path
    assert LC.generate_vae_cache_filename("/data/a/b.png", "cache", "/data", False) == ("cache/a/b.pt", "b.pt")


def _write(tmp_path, n, shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    lats = [torch.randn(*shape, generator=g).bfloat16() for _ in range(n)]
    files = [LC.generate_vae_cache_filename(f"/data/sub/img{i}.png", str(tmp_path / "cache"), "/data")[0] for i in range(n)]
    LC.write_latents(files, lats)
    return files, lats


def test_wire_format_is_a_plain_tensor_and_stager_round_trips_on_cpu(tmp_path):
    files, lats = _write(tmp_path, 5, (16, 8, 12))
    assert all(f.endswith(".pt") and os.path.exists(f) for f in files)
    assert torch.equal(torch.load(files[2], weights_only=False), lats[2])          # what the reference's retrieve_from_cache reads
    with pytest.raises(ValueError, match="image path"):
        LC.write_latents(["x.png"], lats[:1])
    st = LC.LatentStager("cpu", depth=2, workers=2)
    extras = {"prompt_embeds": torch.randn(4, 7, 5).bfloat16(), "add_text_embeds": torch.randn(4, 3).bfloat16()}
    b = st.stage(files[:4], extras).wait()
    assert torch.equal(b["latent_batch"], torch.stack(lats[:4])) and torch.equal(b["prompt_embeds"], extras["prompt_embeds"])
    b2 = st.stage(files[1:5], extras).wait()
    assert torch.equal(b2["latent_batch"], torch.stack(lats[1:5]))
    other, _ = _write(tmp_path / "o", 1, (16, 4, 4), seed=1)
    with pytest.raises(ValueError, match="shape mismatch"):
        st.stage([files[0], other[0]])
    st.close()


@pytest.mark.gpu
def test_staged_batches_on_gpu_overlap_and_match(tmp_path):
    files, lats = _write(tmp_path, 12, (16, 64, 64))
    st = LC.LatentStager("cuda", depth=3, workers=4)
    busy = torch.randn(4096, 4096, device="cuda")
    staged = []
    for k in range(5):
        sel = [files[(k + j) % 12] for j in range(4)]
        s = st.stage(sel, {"add_text_embeds": torch.full((4, 8), float(k)).bfloat16()})
        busy = busy @ busy * 1e-4                      # compute enqueued on the current stream while the copy stream works
        dev = s.wait()
        assert dev["latent_batch"].is_cuda and dev["latent_batch"].dtype == torch.bfloat16
        want = torch.stack([lats[(k + j) % 12] for j in range(4)])
        staged.append((dev["latent_batch"].clone(), want, dev["add_text_embeds"].clone(), k))
    torch.cuda.synchronize()
    for got, want, emb, k in staged:
        assert torch.equal(got.cpu(), want) and float(emb[0, 0]) == float(k)
    st.close()
