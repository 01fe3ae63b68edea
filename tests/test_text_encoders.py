"""Text-encoder embed path: CPU — state-dict names / shapes are the transformers ones; GPU — the
libstb200 T5 encoder and CLIP text model against outputs of the REAL transformers classes (tests/golden/text_golden.pt) and the
fp32 oracle at a wider config."""
from pathlib import Path

import pytest
import torch

from oracle import text_oracle as TO

G = torch.load(Path(__file__).parent / "golden" / "text_golden.pt")


def _t5(cfg):
    from simpletuner_b200.text import T5EncoderModel
    return T5EncoderModel(**cfg.__dict__)


def _clip(cfg):
    from simpletuner_b200.text import CLIPTextModel
    return CLIPTextModel(**cfg.__dict__)


def test_state_dict_names_are_the_transformers_names():
    tc = TO.T5Config(**G["t5_cfg"])
    m = _t5(tc)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == TO.t5_param_shapes(tc)
    sd = {k: torch.zeros(s) for k, s in TO.t5_param_shapes(tc).items()}
    sd["encoder.embed_tokens.weight"] = sd["shared.weight"]          # transformers' tied alias is accepted
    m.load_state_dict(sd, strict=True)
    cc = TO.CLIPTextConfig(**G["clip_cfg_eos2"])
    c = _clip(cc)
    assert {k: tuple(v.shape) for k, v in c.state_dict().items()} == TO.clip_param_shapes(cc)
    with pytest.raises(NotImplementedError):
        _t5(TO.T5Config(d_kv=32))
    from simpletuner_b200._lib import StbError
    with pytest.raises(StbError):
        m(torch.zeros(1, 8, dtype=torch.long))          # no CPU fallback


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.mark.gpu
@pytest.mark.parametrize("S", [40, 300])
def test_t5_encoder_matches_transformers_golden(S):
    tc = TO.T5Config(**G["t5_cfg"])
    P = {k: v.bfloat16().float() for k, v in TO.init_params(TO.t5_param_shapes(tc), seed=G["t5_seed"]).items()}
    m = _t5(tc)
    m.load_state_dict({k: v.bfloat16() for k, v in P.items()})
    m.cuda()
    out = m(G[f"t5_ids_{S}"].cuda(), output_hidden_states=False)[0].float().cpu()
    ref = G[f"t5_out_{S}"]                       # fp32 transformers output on fp32 weights
    # bf16 activations against the fp32 reference on the SAME (bf16-representable) weights.  A random-init T5 is not a gentle
    # function: un-scaled attention logits make the softmax peaked, so bf16 rounding of the activations is amplified layer by
    # layer (measured 0.99985 at S = 300 while every single op agrees to >= 0.99993, tools/_dbg_t5.py) — stated bound 0.999
    cos = float(torch.nn.functional.cosine_similarity(out.double().flatten(), ref.double().flatten(), dim=0))
    assert cos >= 0.999 and _rel(out, ref) < 5e-2, (cos, _rel(out, ref))


@pytest.mark.gpu
@pytest.mark.parametrize("eos", [2, 199])
def test_clip_text_model_matches_transformers_golden(eos):
    cc = TO.CLIPTextConfig(**G[f"clip_cfg_eos{eos}"])
    P = {k: v.bfloat16().float() for k, v in TO.init_params(TO.clip_param_shapes(cc), seed=G["clip_seed"]).items()}
    m = _clip(cc)
    m.load_state_dict({k: v.bfloat16() for k, v in P.items()})
    m.cuda()
    o = m(G[f"clip_ids_eos{eos}"].cuda(), output_hidden_states=False)
    last, pooled = o.last_hidden_state.float().cpu(), o.pooler_output.float().cpu()
    assert torch.equal(o[0], o.last_hidden_state) and torch.equal(o[1], o.pooler_output)
    cos = torch.nn.functional.cosine_similarity
    assert float(cos(last.double().flatten(), G[f"clip_last_eos{eos}"].double().flatten(), dim=0)) >= 0.999
    assert float(cos(pooled.double().flatten(), G[f"clip_pooled_eos{eos}"].double().flatten(), dim=0)) >= 0.999
    assert _rel(pooled, G[f"clip_pooled_eos{eos}"]) < 5e-2


@pytest.mark.gpu
def test_t5_wide_config_and_encode_token_ids_against_the_oracle():
    """A wider T5 (d_model 1024, 16 heads, S = 512: four key tiles, position offsets beyond max_distance) and CLIP through
    `encode_token_ids` (the compute half of FluxPipeline.encode_prompt) against the fp32 oracle on the same bf16-rounded weights."""
    from simpletuner_b200.text import encode_token_ids
    tc = TO.T5Config(vocab_size=500, d_model=1024, d_kv=64, d_ff=2048, num_layers=3, num_heads=16)
    cc = TO.CLIPTextConfig(vocab_size=500, hidden_size=768, intermediate_size=3072, num_hidden_layers=2, num_attention_heads=12, eos_token_id=2)
    # std 0.01: keeps the (un-scaled) T5 attention logits O(1) so that the comparison measures the kernels, not the chaos of a
    # random deep network with one-hot softmaxes
    Pt = {k: v.bfloat16().float() for k, v in TO.init_params(TO.t5_param_shapes(tc), seed=3, std=0.01).items()}
    Pc = {k: v.bfloat16().float() for k, v in TO.init_params(TO.clip_param_shapes(cc), seed=4, std=0.02).items()}
    t5, clip = _t5(tc), _clip(cc)
    t5.load_state_dict({k: v.bfloat16() for k, v in Pt.items()}); clip.load_state_dict({k: v.bfloat16() for k, v in Pc.items()})
    t5.cuda(); clip.cuda()
    g = torch.Generator().manual_seed(5)
    t5_ids = torch.randint(0, 500, (2, 512), generator=g)
    clip_ids = torch.randint(3, 490, (2, 77), generator=g)
    clip_ids[0, 20:] = 499; clip_ids[1, 76] = 499
    mask = torch.ones(2, 512, dtype=torch.long); mask[0, 100:] = 0
    emb, pooled, text_ids, m = encode_token_ids(clip, t5, clip_ids.cuda(), t5_ids.cuda(), mask, num_images_per_prompt=2, t5_padding="zero")
    assert emb.shape == (4, 512, 1024) and pooled.shape == (4, 768) and text_ids.shape == (512, 3) and float(text_ids.abs().sum()) == 0
    ref = TO.t5_encoder(Pt, tc, t5_ids) * mask[..., None]
    ref_pooled = TO.clip_text_model(Pc, cc, clip_ids)[1]
    cos = torch.nn.functional.cosine_similarity
    e = emb.float().cpu()
    assert torch.equal(e[0], e[1]) and float(e[0, 100:].abs().sum()) == 0          # repeat per image; zeroed padding
    ref = torch.stack([ref[0], ref[1]])
    c1 = float(cos(e[0::2].double().flatten(), ref.double().flatten(), dim=0))
    c2 = float(cos(pooled.double().cpu()[0::2].flatten(), ref_pooled.double().flatten(), dim=0))
    assert c1 >= 0.999 and c2 >= 0.999, (c1, c2)


@pytest.mark.gpu
def test_attention_bias_and_new_epilogues_against_torch():
    from simpletuner_b200 import ops
    torch.manual_seed(0)
    B, S, H, hd = 2, 300, 3, 64
    q, k, v = (torch.randn(B, S, H, hd, device="cuda").bfloat16() for _ in range(3))
    bias = torch.randn(H, S, S, device="cuda").bfloat16()
    bias[:, :, 200:] = float("-inf")                               # a padding-style mask on top
    o, lse = ops.attn_fwd(q, k, v, scale=0.2, bias=bias)
    sc = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) * 0.2 + bias.float()[None]
    ref = torch.einsum("bhqk,bkhd->bqhd", sc.softmax(-1), v.float())
    assert torch.allclose(o.float(), ref, atol=2e-2, rtol=2e-2), float((o.float() - ref).abs().max())
    assert float(torch.nn.functional.cosine_similarity(o.double().flatten(), ref.double().flatten(), dim=0)) >= 0.99995
    assert torch.allclose(lse, torch.logsumexp(sc, -1), atol=2e-2, rtol=1e-3)
    # shared-across-heads causal mask, one ragged key tile
    S2 = 77
    q, k, v = (torch.randn(1, S2, 2, 64, device="cuda").bfloat16() for _ in range(3))
    causal = torch.full((S2, S2), float("-inf"), device="cuda").triu(1).bfloat16()[None]
    o, _ = ops.attn_fwd(q, k, v, bias=causal)
    ref = torch.nn.functional.scaled_dot_product_attention(q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2),
                                                           is_causal=True).transpose(1, 2)
    assert torch.allclose(o.float(), ref, atol=2e-2, rtol=2e-2)
    # epilogues
    x, w, bvec, aux = (torch.randn(2, 100, 96, device="cuda").bfloat16(), (torch.randn(160, 96, device="cuda") * 0.2).bfloat16(),
                       torch.randn(160, device="cuda").bfloat16(), torch.randn(2, 100, 160, device="cuda").bfloat16())
    y = (x.float() @ w.float().t() + bvec.float())
    got = ops.gemm([x], [w], bvec, epi=ops.EPI_MUL, aux=aux)
    assert torch.allclose(got.float(), y.bfloat16().float() * aux.float(), atol=3e-2, rtol=2e-2)
    got = ops.gemm([x], [w], bvec, epi=ops.EPI_QUICK_GELU)
    yb = y.bfloat16().float()
    assert torch.allclose(got.float(), yb * torch.sigmoid(1.702 * yb), atol=3e-2, rtol=2e-2)
    # T5LayerNorm kernel on a strided view
    xs = torch.randn(2, 50, 4096 + 8, device="cuda").bfloat16()[:, :, :4096]
    wn = (1 + 0.1 * torch.randn(4096, device="cuda")).bfloat16()
    got = ops.rmsnorm_fwd(xs, wn, 1e-6)
    assert torch.allclose(got.float(), TO.t5_layer_norm(xs, wn, 1e-6).float(), atol=2e-2, rtol=1e-2)
