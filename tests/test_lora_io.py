"""CPU: LoRA save / load round trip in the reference's key layout (save_hooks.py:862-891) and the fused <-> un-fused
q|k|v adapter conversion (diffusers_overrides.py:133-466)."""
import torch

from simpletuner_b200.training import lora_io as IO
from tests.test_shim_cpu import _ref_flux


def _model(rank=4):
    m = _ref_flux()
    m.add_adapter(rank=rank, lora_alpha=rank)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for lin in m.lora_linears().values():
            lin.lora_B["default"].weight.normal_(0, 0.02, generator=g)
    return m


def test_peft_state_dict_keys_and_safetensors_round_trip(tmp_path):
    m = _model()
    sd = IO.get_peft_model_state_dict(m)
    assert len(sd) == 22 and "transformer_blocks.0.attn.to_q.lora_A.weight" in sd and "single_transformer_blocks.0.attn.to_v.lora_B.weight" in sd
    assert all(".default." not in k for k in sd)
    path = IO.save_lora_weights(tmp_path, sd)
    assert path.endswith("pytorch_lora_weights.safetensors")
    from safetensors.torch import load_file
    raw = load_file(path)
    assert all(k.startswith("transformer.") for k in raw) and len(raw) == 22          # diffusers component prefix
    back = IO.load_lora_weights(tmp_path)
    m2 = _model()
    with torch.no_grad():
        for lin in m2.lora_linears().values():
            lin.lora_A["default"].weight.zero_(); lin.lora_B["default"].weight.zero_()
    IO.set_peft_model_state_dict(m2, back)
    for (n1, p1), (n2, p2) in zip(m.named_parameters(), m2.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2), n1
    try:
        IO.set_peft_model_state_dict(m2, {"nope.lora_A.weight": torch.zeros(1)})
        assert False
    except KeyError:
        pass


def test_fused_qkv_adapter_conversion_is_exact():
    m = _model(rank=4)
    sd = {k: v.float() for k, v in IO.get_peft_model_state_dict(m).items()}
    fused = IO.fuse_qkv_lora(sd)
    assert "transformer_blocks.0.attn.to_qkv.lora_A.weight" in fused and "transformer_blocks.0.attn.add_qkv_proj.lora_B.weight" in fused
    assert not any(k.endswith(".to_q.lora_A.weight") for k in fused) and "transformer_blocks.0.attn.to_out.0.lora_A.weight" in fused
    x = torch.randn(5, 256)
    pre = "transformer_blocks.0.attn."
    want = torch.cat([x @ sd[pre + f"{n}.lora_A.weight"].t() @ sd[pre + f"{n}.lora_B.weight"].t() for n in ("to_q", "to_k", "to_v")], 1)
    got = x @ fused[pre + "to_qkv.lora_A.weight"].t() @ fused[pre + "to_qkv.lora_B.weight"].t()
    assert torch.allclose(got, want, atol=1e-6)
    back = IO.unfuse_qkv_lora(fused)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    # a dense fused B (one rank-R adapter trained on the fused projection) splits into three adapters sharing A
    dense = dict(fused)
    dense[pre + "to_qkv.lora_B.weight"] = torch.randn_like(fused[pre + "to_qkv.lora_B.weight"])
    un = IO.unfuse_qkv_lora(dense)
    gotd = torch.cat([x @ un[pre + f"{n}.lora_A.weight"].t() @ un[pre + f"{n}.lora_B.weight"].t() for n in ("to_q", "to_k", "to_v")], 1)
    assert torch.allclose(gotd, x @ dense[pre + "to_qkv.lora_A.weight"].t() @ dense[pre + "to_qkv.lora_B.weight"].t(), atol=1e-5)
