"""CPU: the shipped reference-side shims (simpletuner_b200/shim) against a stub that reproduces the reference's load /
LoRA-attach hook order.  Plumbing only (no kernels run on CPU): class swap in `post_model_load_setup`, identical
state dict, LoRA attach through a LoraConfig-shaped object, fall-back to the reference class on unsupported options,
the attention-processor API, guard rails that must raise, and the RoPE cache key."""
from types import SimpleNamespace

import pytest
import torch

from simpletuner_b200.flux.model import Flux, default_config, prepare_latent_image_ids
from simpletuner_b200.flux.transformer import B200FusedAttnProcessor, FluxTransformer2DModel
from simpletuner_b200.shim import FAMILIES, make_b200_family
from tests.shim_stub import StubFoundation

TINY = dict(in_channels=64, num_layers=1, num_single_layers=1, attention_head_dim=128, num_attention_heads=2,
            joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 56, 56))


def _ref_flux():
    """Stands in for the diffusers-based reference denoiser: same `.config` fields and state-dict names."""
    torch.manual_seed(0)
    m = FluxTransformer2DModel(**TINY)
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0, 0.02)
    return m


class RefFlux(StubFoundation):
    MODEL_CLASS = staticmethod(_ref_flux)
    LORA_TARGETS = ["to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"]


def _cfg(**over):
    kw = dict(lora_rank=4, model_type="lora", lora_type="standard")
    kw.update(over)
    return default_config(**kw)


def test_families_cover_the_reference_registry_keys():
    assert set(FAMILIES) == {"flux", "sd3", "pixart_sigma"}          # flux/model.py:1504, sd3/model.py:900, pixart/model.py:858
    for spec in FAMILIES.values():
        assert spec.denoiser() is not None and spec.step() is not None


def test_post_model_load_setup_swaps_the_denoiser_after_the_reference_hooks():
    cls = make_b200_family(RefFlux, "flux")
    assert cls.__mro__[1].__name__ == "B200FoundationMixin" and issubclass(cls, RefFlux)
    fam = cls(_cfg(gradient_checkpointing=True, gradient_checkpointing_interval=2), "cpu")
    fam.load_model()
    # hook order of common.py:3543-3548, the shim's hook last
    assert fam.calls == ["configure_chunked_feed_forward", "apply_gradient_checkpointing_settings", "fuse_qkv_projections",
                         "post_model_load_setup"]
    assert fam._b200 is not None and fam._b200_fallback_reason is None
    assert isinstance(fam.model, FluxTransformer2DModel) and fam._b200.model is fam.model
    ref = _ref_flux()
    sd = fam.model.state_dict()
    assert set(sd) == set(ref.state_dict())
    assert all(torch.equal(sd[k], v) for k, v in ref.state_dict().items())
    # settings applied to the reference module before the hook are carried over
    assert fam.model.gradient_checkpointing is True and fam.model.gradient_checkpointing_interval == 2
    # LoRA attach goes through the reference's own add_lora_adapter -> model.add_adapter(LoraConfig)
    fam.add_lora_adapter()
    lin = fam.model.lora_linears()
    assert len(lin) == 8 + 3 and all(l.lora_A["default"].weight.shape[0] == 4 for l in lin.values())
    names = [n for n, p in fam.model.named_parameters() if p.requires_grad]
    assert all(".lora_A.default.weight" in n or ".lora_B.default.weight" in n for n in names) and len(names) == 22


@pytest.mark.parametrize("over,needle", [
    (dict(model_type="full"), "model_type"),
    (dict(lora_type="lycoris"), "LyCORIS"),
    (dict(flux_attention_masked_training=True), "masked"),
    (dict(controlnet=True), "controlnet"),
    (dict(weight_dtype=torch.float32), "bf16"),
])
def test_unsupported_options_fall_back_to_the_reference_class(over, needle):
    fam = make_b200_family(RefFlux, "flux")(_cfg(**over), "cpu")
    fam.load_model()
    assert fam._b200 is None and needle in fam._b200_fallback_reason
    # every step method now runs the REFERENCE implementation
    fam.prepare_batch({"x": 1}, {})
    fam.model_predict({"x": 1})
    fam.loss_with_logs({}, {})
    assert fam.calls[-3:] == ["ref.prepare_batch", "ref.model_predict", "ref.loss_with_logs"]


def test_step_methods_route_to_the_b200_step_and_follow_ddp_wrapping():
    fam = make_b200_family(RefFlux, "flux")(_cfg(), "cpu")
    fam.load_model()
    seen = []
    step = fam._b200
    step.prepare_batch = lambda b, s: seen.append(("prepare", step.model)) or b
    step.model_predict = lambda pb: seen.append(("predict", step.model)) or {"model_prediction": 0}
    step.loss_with_logs = lambda pb, out, apply_conditioning_mask=True: (seen.append(("loss", apply_conditioning_mask)) or (0, None))
    wrapped = SimpleNamespace(module=fam.model)      # what accelerator.prepare leaves in family.model (DDP)
    fam.model = wrapped
    fam.prepare_batch({}, {})
    fam.model_predict(prepared_batch={})
    fam.loss_with_logs(prepared_batch={}, model_output={}, apply_conditioning_mask=True)
    assert [s[0] for s in seen] == ["prepare", "predict", "loss"] and seen[0][1] is wrapped and seen[2][1] is True
    assert "ref.prepare_batch" not in fam.calls
    with pytest.raises(NotImplementedError):
        fam.model_predict(prepared_batch={}, custom_timesteps=[1])


def test_guard_rails_raise_instead_of_ignoring():
    w = Flux(_cfg(), transformer=_ref_flux(), device=torch.device("cpu"))
    with pytest.raises(NotImplementedError, match="conditioning"):
        w.prepare_batch({"latent_batch": torch.zeros(1, 16, 4, 4), "conditioning_packed_latents": torch.zeros(1)}, {})
    w2 = Flux(_cfg(flux_attention_masked_training=True), transformer=_ref_flux(), device=torch.device("cpu"))
    with pytest.raises(NotImplementedError, match="masked"):
        w2.prepare_batch({"latent_batch": torch.zeros(1, 16, 4, 4)}, {})
    for key in ("loss_mask_type", "conditioning_type"):
        for kind in ("mask", "segmentation"):
            with pytest.raises(NotImplementedError, match="masked"):
                w.loss({key: kind}, {"model_prediction": torch.zeros(1, 4, 64)}, apply_conditioning_mask=True)


def test_attn_processor_api():
    m = _ref_flux()
    procs = m.attn_processors
    assert set(procs) == {"transformer_blocks.0.attn.processor", "single_transformer_blocks.0.attn.processor"}
    assert all(isinstance(p, B200FusedAttnProcessor) for p in procs.values())

    class FluxAttnProcessor2_0:   # same NAME as the reference's default processor (flux/transformer.py:116)
        pass

    class IPAdapterFluxAttnProcessor:
        pass

    m.set_attn_processor(FluxAttnProcessor2_0())
    assert all(type(p).__name__ == "FluxAttnProcessor2_0" for p in m.attn_processors.values())
    with pytest.raises(NotImplementedError):
        m.set_attn_processor(IPAdapterFluxAttnProcessor())
    with pytest.raises(ValueError, match="number of processors"):
        m.set_attn_processor({"transformer_blocks.0.attn.processor": FluxAttnProcessor2_0()})
    assert m.fuse_qkv_projections() is None


def test_rope_cache_distinguishes_transposed_buckets():
    """(128, 64) and (64, 128) latent grids have equal id shapes and equal id sums."""
    m = _ref_flux()
    txt = torch.zeros(8, 3)
    a, b = prepare_latent_image_ids(128, 64), prepare_latent_image_ids(64, 128)
    assert a.shape == b.shape and float(a.sum()) == float(b.sum()) and not torch.equal(a, b)
    cos_a, sin_a = m._rope(txt, a, "cpu")[:2]
    cos_b, sin_b = m._rope(txt, b, "cpu")[:2]
    from simpletuner_b200.flux.transformer import rope_tables
    ref_b = rope_tables(torch.cat((txt, b)), m.config.axes_dims_rope)
    assert torch.equal(cos_b, ref_b[0]) and torch.equal(sin_b, ref_b[1])
    assert not torch.equal(cos_a, cos_b)
    assert m._rope(txt, a, "cpu")[0] is cos_a     # cache hit for identical ids


def test_x_prediction_fixup_replaces_the_prediction_and_loss_repacks():
    from simpletuner_b200.training.step import TrainStep

    class W:
        noise_schedule = SimpleNamespace(config=SimpleNamespace(prediction_type="sample"))
        model = torch.nn.Linear(1, 1)

        def model_predict(self, pb):
            p = torch.ones(1, 16, 4, 4)
            return {"model_prediction": p, "_packed_prediction": Flux._pack(p), "_unpacked_ref": p}

    step = TrainStep(W(), torch.optim.SGD(W.model.parameters(), lr=0.1))
    pb = {"noise": torch.full((1, 16, 4, 4), 0.25)}
    out = step.model_predict(pb)
    assert torch.equal(out["model_prediction"], torch.full((1, 16, 4, 4), 0.75))
    f = Flux.__new__(Flux)
    packed = f._packed_for_loss(out)                 # the private packed copy is stale -> re-packed from the new tensor
    assert torch.equal(packed, Flux._pack(out["model_prediction"]))
    untouched = W().model_predict(pb)
    assert f._packed_for_loss(untouched) is untouched["_packed_prediction"]


def test_install_lycoris_routes_b200_denoisers_and_leaves_other_modules_alone():
    """trainer.py:3391-3497 calls `lycoris.create_lycoris(component, multiplier, linear_dim, linear_alpha, **cfg)`; with the
    route installed a libstb200 denoiser gets simpletuner_b200.lycoris, anything else the third-party implementation."""
    import types

    from simpletuner_b200 import lycoris as LY
    from simpletuner_b200.shim import foundation as F

    calls = []

    class _Net:
        @staticmethod
        def apply_preset(p):
            calls.append(("preset", p))

    fake = types.SimpleNamespace(create_lycoris=lambda m, *a, **k: calls.append(("third-party", type(m).__name__)) or "tp",
                                 LycorisNetwork=_Net)
    old = F._LYCORIS_ROUTED
    try:
        F.install_lycoris(fake)
        assert F._LYCORIS_ROUTED
        fake.LycorisNetwork.apply_preset({"target_module": ["Attention"], "module_algo_map": {"Attention": {"factor": 4}}})
        assert calls[-1][0] == "preset" and LY.LycorisNetwork._preset["target_module"] == ["Attention"]
        assert fake.create_lycoris(torch.nn.Linear(4, 4), 1.0, 8, 1, algo="lokr") == "tp" and calls[-1] == ("third-party", "Linear")
        net = fake.create_lycoris(_ref_flux(), 1.0, 10000, 1, algo="lokr", factor=4)
        assert isinstance(net, LY.LycorisNetwork) and len(net.loras) == 11 and net.loras[0].shape == ((4, 64), (4, 64))
        # a LyCORIS run swaps to the B200 classes only when routed, and only LoKr
        fam = make_b200_family(RefFlux, "flux")(_cfg(lora_type="lycoris", lycoris_config={"algo": "lokr", "linear_dim": 10000}), "cpu")
        fam.load_model()
        assert fam._b200 is not None, fam._b200_fallback_reason
        fam = make_b200_family(RefFlux, "flux")(_cfg(lora_type="lycoris", lycoris_config={"algo": "loha"}), "cpu")
        fam.load_model()
        assert fam._b200 is None and "lokr" in fam._b200_fallback_reason
    finally:
        F._LYCORIS_ROUTED = old
        LY.LycorisNetwork._preset = {"target_module": ["Attention", "FeedForward"], "module_algo_map": {}}
