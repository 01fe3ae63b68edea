import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.nn.functional as F
from oracle import text_oracle as TO
from simpletuner_b200 import ops
from simpletuner_b200.text import T5EncoderModel
G = torch.load(Path(__file__).resolve().parent.parent / "tests/golden/text_golden.pt")
cos = lambda a, b: float(F.cosine_similarity(a.float().flatten().cpu(), b.float().flatten().cpu(), dim=0))
tc = TO.T5Config(**G["t5_cfg"])
P = {k: v.bfloat16().float() for k, v in TO.init_params(TO.t5_param_shapes(tc), seed=G["t5_seed"]).items()}
for S in (40, 129, 300):
    ids = G["t5_ids_300"][:, :S]
    m = T5EncoderModel(**tc.__dict__); m.load_state_dict({k: v.bfloat16() for k, v in P.items()}); m.cuda()
    B = ids.shape[0]; H, hd = tc.num_heads, tc.d_kv; inner = H * hd
    h_ref = P["shared.weight"][ids]
    bias_ref = TO.t5_position_bias(P, tc, S)
    h = F.embedding(ids.cuda(), m.shared.weight).contiguous()
    bias = m._position_bias(S, h.device)
    print(S, "bias", cos(bias, bias_ref), "bias maxdiff", float((bias.float().cpu() - bias_ref).abs().max()))
    p = "encoder.block.0.layer."
    n_ref = TO.t5_layer_norm(h_ref, P[p + "0.layer_norm.weight"], 1e-6)
    n = ops.rmsnorm_fwd(h, m.encoder.block[0].layer[0].layer_norm.weight, 1e-6)
    print(S, "rmsnorm", cos(n, n_ref))
    w_qkv = m._build_plans()[0]
    qkv = ops.gemm([n], [w_qkv])
    q_ref = F.linear(n_ref, P[p + "0.SelfAttention.q.weight"]); k_ref = F.linear(n_ref, P[p + "0.SelfAttention.k.weight"]); v_ref = F.linear(n_ref, P[p + "0.SelfAttention.v.weight"])
    print(S, "qkv", cos(qkv[..., :inner], q_ref), cos(qkv[..., inner:2 * inner], k_ref), cos(qkv[..., 2 * inner:], v_ref))
    q, k, v = (qkv[:, :, i * inner:(i + 1) * inner].unflatten(-1, (H, hd)) for i in range(3))
    o, _ = ops.attn_fwd(q, k, v, scale=1.0, bias=bias)
    qh, kh, vh = (t.view(B, S, H, hd).transpose(1, 2) for t in (q_ref, k_ref, v_ref))
    sc = qh @ kh.transpose(2, 3) + bias_ref[None]
    o_ref = (sc.softmax(-1) @ vh).transpose(1, 2)
    print(S, "attn", cos(o, o_ref), "per-row-block", [round(cos(o[:, a:a + 64], o_ref[:, a:a + 64]), 4) for a in range(0, S, 64)])
    o_nb, _ = ops.attn_fwd(q, k, v, scale=1.0)
    o_ref_nb = ((qh @ kh.transpose(2, 3)).softmax(-1) @ vh).transpose(1, 2)
    print(S, "attn no-bias", cos(o_nb, o_ref_nb))
    # same inputs on the device in fp32 (isolates the kernel from bf16 input rounding)
    scd = torch.einsum("bqhd,bkhd->bhqk", q.float(), k.float()) + bias.float()[None]
    o_dev = torch.einsum("bhqk,bkhd->bqhd", scd.softmax(-1), v.float())
    print(S, "attn vs device fp32 on same q/k/v/bias", cos(o, o_dev), "max abs", float((o.float() - o_dev).abs().max()), "ref max", float(o_dev.abs().max()))

    # ---- continue through the block with the module's own ops, comparing every stage
    a = m.encoder.block[0].layer[0]; ff = m.encoder.block[0].layer[1]; d = ff.DenseReluDense
    h1 = h.clone()
    ops.gemm([o.view(B, S, inner)], [a.SelfAttention.o.weight], None, out=h1, epi=ops.EPI_ADD_RES, res=h1)
    h1_ref = h_ref + F.linear(o_ref.reshape(B, S, inner), P[p + "0.SelfAttention.o.weight"])
    print(S, "after o-proj (in place)", cos(h1, h1_ref))
    h1b = ops.gemm([o.view(B, S, inner)], [a.SelfAttention.o.weight], None, epi=ops.EPI_ADD_RES, res=h)
    print(S, "after o-proj (out of place)", cos(h1b, h1_ref), "in-place == out-of-place", bool(torch.equal(h1, h1b)))
    n2 = ops.rmsnorm_fwd(h1, ff.layer_norm.weight, 1e-6)
    n2_ref = TO.t5_layer_norm(h1_ref, P[p + "1.layer_norm.weight"], 1e-6)
    print(S, "rmsnorm2", cos(n2, n2_ref))
    g = ops.gemm([n2], [d.wi_0.weight], None, epi=ops.EPI_GELU)
    g_ref = F.gelu(F.linear(n2_ref, P[p + "1.DenseReluDense.wi_0.weight"]), approximate="tanh")
    print(S, "gelu(wi_0)", cos(g, g_ref))
    u = ops.gemm([n2], [d.wi_1.weight], None, epi=ops.EPI_MUL, aux=g)
    u_ref = g_ref * F.linear(n2_ref, P[p + "1.DenseReluDense.wi_1.weight"])
    print(S, "gated", cos(u, u_ref))
    h2 = h1.clone()
    ops.gemm([u], [d.wo.weight], None, out=h2, epi=ops.EPI_ADD_RES, res=h2)
    h2_ref = h1_ref + F.linear(u_ref, P[p + "1.DenseReluDense.wo.weight"])
    print(S, "after FF", cos(h2, h2_ref), "norms", float(h2.float().norm()), float(h2_ref.norm()))
    full = m(ids.cuda())[0]
    full_ref = TO.t5_encoder(P, tc, ids)
    print(S, "FULL", cos(full, full_ref))
