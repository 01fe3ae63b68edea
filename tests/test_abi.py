"""CPU: the C-ABI library loads, exports every symbol include/stb200.h declares, and the product path
refuses to run without a CUDA device (no silent CPU fallback)."""
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    txt = (ROOT / "include" / "stb200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(stb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from simpletuner_b200 import _lib

    if _lib.needs_build():
        _lib.build()
    handle = _lib.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(handle, n), f"libstb200.so does not export {n}"
    # and the ctypes table covers the header exactly
    assert sorted(_lib.SYMBOLS) == names
    assert handle.stb_version() == 100


def test_product_path_fails_loudly_without_cuda():
    from simpletuner_b200 import _lib, ops

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    w = torch.zeros(64, 64, dtype=torch.bfloat16)
    with pytest.raises(_lib.StbError):
        ops.gemm([a], [w])
    from simpletuner_b200.flux.transformer import FluxTransformer2DModel

    m = FluxTransformer2DModel(num_layers=1, num_single_layers=1, num_attention_heads=1, joint_attention_dim=64,
                               pooled_projection_dim=32, in_channels=16)
    with pytest.raises(_lib.StbError):
        m(hidden_states=torch.zeros(1, 4, 16), encoder_hidden_states=torch.zeros(1, 4, 64),
          pooled_projections=torch.zeros(1, 32), timestep=torch.zeros(1), img_ids=torch.zeros(4, 3),
          txt_ids=torch.zeros(4, 3), return_dict=False)


def test_no_oracle_import_in_product():
    bad = []
    for p in (ROOT / "simpletuner_b200").rglob("*.py"):
        s = p.read_text()
        if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M):
            bad.append(str(p))
    assert not bad, f"product code must not import the oracle: {bad}"
