"""Kernel-level parity checks of libstb200 against plain PyTorch fp32 math on the same GPU tensors.

Each check is a function returning a dict {"ok": bool, "max_err": ..., "tol": ..., ...}.  They are used
by the `-m gpu` pytest tests and by tools/run_gpu_checks.py (which runs every check in its own
subprocess so that a trapped kernel cannot poison the rest of the run).
"""
from __future__ import annotations

import math

import torch

from simpletuner_b200 import ops

DEV = "cuda"


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def _report(name, got, ref, atol, rtol, extra=None):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    nbad = int(bad.sum().item())
    out = {
        "name": name,
        "ok": nbad == 0 and bool(torch.isfinite(got).all().item()),
        "max_err": float(err.max().item()),
        "ref_absmax": float(ref.abs().max().item()),
        "n_bad": nbad,
        "numel": got.numel(),
    }
    if nbad:
        idx = torch.nonzero(bad)[:8].tolist()
        out["first_bad"] = idx
        # error structure along the last two dims in 16-wide blocks: reveals tile / descriptor mistakes
        e2 = err.reshape(-1, err.shape[-1])
        rows = e2.shape[0]
        rb = max(1, rows // 8)
        cb = max(1, e2.shape[1] // 8)
        out["row_block_maxerr"] = [round(float(e2[i * rb:(i + 1) * rb].max().item()), 4) for i in range(min(8, math.ceil(rows / rb)))]
        out["col_block_maxerr"] = [round(float(e2[:, j * cb:(j + 1) * cb].max().item()), 4) for j in range(min(8, math.ceil(e2.shape[1] / cb)))]
    if extra:
        out.update(extra)
    return out


# --------------------------------------------------------------------------------------------- GEMM
def gelu_tanh(x):
    return torch.nn.functional.gelu(x, approximate="tanh")


def check_gemm(M=256, N=256, K=128, B=1, bias=False, epi=ops.EPI_STORE, tile=(0, 0), segs=None, strided=False,
               nan_to_num=False, name=None, w_kn=None):
    """segs: list of extra K sizes appended as additional segments.  w_kn[i]: hand segment i's weight over as [K, N]
    (the contraction index is the row — what a dgrad reads from the forward weight) instead of [N, K]."""
    Ks = [K] + list(segs or [])
    S = M
    a_list, w_list = [], []
    for i, k in enumerate(Ks):
        if strided:
            full = _rand(B, S + 3, k + 16, seed=10 + i)
            a_list.append(full[:, 1:S + 1, 8:8 + k])
            wfull = _rand(N, k + 8, scale=0.5, seed=20 + i)
            w_list.append(wfull[:, :k])
        else:
            a_list.append(_rand(B, S, k, seed=10 + i))
            w_list.append(_rand(N, k, scale=0.5, seed=20 + i))
    bias_t = _rand(N, seed=30) if bias else None
    acc = sum(a.float() @ w.float().t() for a, w in zip(a_list, w_list))
    if bias_t is not None:
        acc = acc + bias_t.float()
    gate = res = aux = None
    kw = {}
    if epi == ops.EPI_STORE:
        ref = acc
    elif epi == ops.EPI_GELU:
        aux = torch.zeros(B, S, N, dtype=torch.bfloat16, device=DEV)
        ref = gelu_tanh(acc.bfloat16().float())
    elif epi == ops.EPI_GATE_RES:
        gate = _rand(B, N, seed=31)
        res = _rand(B, S, N, seed=32)
        y = acc.bfloat16().float()
        ref = res.float() + (gate.float()[:, None, :] * y).bfloat16().float()
    elif epi == ops.EPI_MUL_DGELU:
        aux = _rand(B, S, N, seed=33)
        x = aux.float().requires_grad_(True)
        gelu_tanh(x).sum().backward()
        ref = acc * x.grad
    elif epi == ops.EPI_ADD_RES:
        res = _rand(B, S, N, seed=32)
        ref = acc + res.float()
    out = None
    if strided:
        outfull = torch.zeros(B, S + 2, N + 8, dtype=torch.bfloat16, device=DEV)
        out = outfull[:, 1:S + 1, :N]
    w_pass = list(w_list)
    if w_kn is not None:
        for i, flag in enumerate(w_kn):
            if flag:                      # [K, N] storage; with `strided` a column range of a wider matrix (row stride > N)
                if strided:
                    wide = torch.zeros(w_list[i].shape[1], N + 24, dtype=torch.bfloat16, device=DEV)
                    wide[:, 8:8 + N] = w_list[i].t()
                    w_pass[i] = wide[:, 8:8 + N]
                else:
                    w_pass[i] = w_list[i].t().contiguous()
    got = ops.gemm(a_list, w_pass, bias_t, out=out, epi=epi, gate=gate, res=res, aux=aux, nan_to_num=nan_to_num, tile=tile, w_kn=w_kn)
    torch.cuda.synchronize()
    scale = float(ref.abs().max().item())
    r = _report(name or f"gemm_M{M}_N{N}_K{Ks}_B{B}_epi{epi}_tile{tile}", got, ref, atol=scale * 6e-3, rtol=1.6e-2)
    if epi == ops.EPI_GELU:
        r2 = _report("aux", aux, acc, atol=float(acc.abs().max()) * 6e-3, rtol=1.6e-2)
        r["aux_ok"] = r2["ok"]
        r["ok"] = r["ok"] and r2["ok"]
    if strided:
        # nothing outside the view may be touched
        mask = torch.ones_like(outfull, dtype=torch.bool)
        mask[:, 1:S + 1, :N] = False
        r["halo_clean"] = bool((outfull[mask] == 0).all().item())
        r["ok"] = r["ok"] and r["halo_clean"]
    return r


# --------------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, scale):
    # q,k,v [B,S,H,D] -> fp32 reference in [B,S,H,D]
    qf, kf, vf = (t.float().permute(0, 2, 1, 3) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * scale
    lse = torch.logsumexp(s, dim=-1)
    p = torch.softmax(s, dim=-1)
    o = p @ vf
    return o.permute(0, 2, 1, 3), lse


def check_attn_fwd(B=1, H=2, Sq=256, Sk=None, HD=128, strided=False, name=None, qscale=1.0):
    Sk = Sk or Sq
    if strided:
        # q/k/v as slices of a fused [B, S, 3*H*HD] projection buffer (the layout the model uses)
        assert Sk == Sq
        qkv = _rand(B, Sq, 3 * H * HD, seed=1, scale=qscale)
        q = qkv[..., 0 * H * HD:1 * H * HD].unflatten(-1, (H, HD))
        k = qkv[..., 1 * H * HD:2 * H * HD].unflatten(-1, (H, HD))
        v = qkv[..., 2 * H * HD:3 * H * HD].unflatten(-1, (H, HD))
    else:
        q = _rand(B, Sq, H, HD, seed=1, scale=qscale)
        k = _rand(B, Sk, H, HD, seed=2, scale=qscale)
        v = _rand(B, Sk, H, HD, seed=3)
    scale = HD ** -0.5
    o, lse = ops.attn_fwd(q, k, v, scale)
    torch.cuda.synchronize()
    o_ref, lse_ref = _attn_ref(q, k, v, scale)
    r = _report(name or f"attn_fwd_B{B}_H{H}_Sq{Sq}_Sk{Sk}_HD{HD}", o, o_ref, atol=2e-2 * float(o_ref.abs().max()), rtol=2e-2)
    r2 = _report("lse", lse, lse_ref, atol=2e-2, rtol=1e-3)
    r["lse_ok"] = r2["ok"]
    r["lse_max_err"] = r2["max_err"]
    r["ok"] = r["ok"] and r2["ok"]
    return r


def check_attn_bwd(B=1, H=2, Sq=256, Sk=None, HD=128, name=None):
    Sk = Sk or Sq
    q = _rand(B, Sq, H, HD, seed=1)
    k = _rand(B, Sk, H, HD, seed=2)
    v = _rand(B, Sk, H, HD, seed=3)
    d_o = _rand(B, Sq, H, HD, seed=4)
    scale = HD ** -0.5
    o, lse = ops.attn_fwd(q, k, v, scale)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, d_o, lse, scale)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = _attn_ref(qf, kf, vf, scale)
    o_ref.backward(d_o.float())
    res = {}
    ok = True
    for nm, got, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        r = _report(nm, got, ref, atol=3e-2 * float(ref.abs().max()), rtol=3e-2)
        res[nm] = r
        ok = ok and r["ok"]
    return {"name": name or f"attn_bwd_B{B}_H{H}_Sq{Sq}_Sk{Sk}_HD{HD}", "ok": ok,
            "max_err": max(res[n]["max_err"] for n in res), **{f"{n}_detail": res[n] for n in res if not res[n]["ok"]}}


# --------------------------------------------------------------------------------------------- elementwise
def check_ln_modulate(B=2, S=64, D=3072):
    x = _rand(B, S, D, seed=1)
    mod = _rand(B, 6 * D, seed=2, scale=0.3)
    shift, scale = mod[:, :D], mod[:, D:2 * D]
    out = ops.ln_modulate_fwd(x, shift, scale)
    ln = torch.nn.functional.layer_norm(x.float(), (D,), eps=1e-6)
    ref = ln * (1 + scale.float()[:, None]) + shift.float()[:, None]
    r = _report(f"ln_modulate_fwd_D{D}", out, ref, atol=3e-2, rtol=2e-2)
    # backward
    dy = _rand(B, S, D, seed=3)
    add = _rand(B, S, D, seed=4)
    dx = ops.ln_modulate_bwd(dy, x, scale, add=add)
    xf = x.float().requires_grad_(True)
    lnf = torch.nn.functional.layer_norm(xf, (D,), eps=1e-6)
    (lnf * (1 + scale.float().bfloat16().float()[:, None])).backward(dy.float())
    ref_dx = xf.grad + add.float()
    r2 = _report("ln_modulate_bwd", dx, ref_dx, atol=3e-2 * float(ref_dx.abs().max()), rtol=2e-2)
    r["bwd_ok"] = r2["ok"]
    r["bwd_max_err"] = r2["max_err"]
    r["ok"] = r["ok"] and r2["ok"]
    return r


def _rope_tables(S, HD, seed=5):
    g = torch.Generator().manual_seed(seed)
    ang = torch.rand(S, HD // 2, generator=g) * 6.28
    cos = ang.cos().repeat_interleave(2, dim=-1).float().to(DEV).contiguous()
    sin = ang.sin().repeat_interleave(2, dim=-1).float().to(DEV).contiguous()
    return cos, sin


def _rmsnorm_rope_ref(x, w, cos, sin, eps=1e-6):
    # x [B,S,H,HD] fp32
    var = x.pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps) * w
    xr, xi = y.reshape(*y.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return y * cos[None, :, None, :] + rot * sin[None, :, None, :]


def check_qk_rmsnorm_rope(B=2, S=96, H=4, HD=128, s_split=32):
    Cc = 3 * H * HD
    src = _rand(B, S, Cc, seed=1)
    wq, wk, wq1, wk1 = (_rand(HD, seed=10 + i, scale=0.2) + 1 for i in range(4))
    cos, sin = _rope_tables(S, HD)
    q, k = ops.qk_rmsnorm_rope_fwd(src, H * HD, H, HD, wq, wk, wq1, wk1, s_split, cos, sin)
    torch.cuda.synchronize()

    def ref_of(xf):
        xq = xf[..., :H * HD].unflatten(-1, (H, HD))
        xk = xf[..., H * HD:2 * H * HD].unflatten(-1, (H, HD))
        outq = torch.cat([_rmsnorm_rope_ref(xq[:, :s_split], wq1.float(), cos[:s_split], sin[:s_split]),
                          _rmsnorm_rope_ref(xq[:, s_split:], wq.float(), cos[s_split:], sin[s_split:])], 1)
        outk = torch.cat([_rmsnorm_rope_ref(xk[:, :s_split], wk1.float(), cos[:s_split], sin[:s_split]),
                          _rmsnorm_rope_ref(xk[:, s_split:], wk.float(), cos[s_split:], sin[s_split:])], 1)
        return outq, outk

    xf = src.float().requires_grad_(True)
    rq, rk = ref_of(xf)
    r = _report(f"qk_rmsnorm_rope_fwd_HD{HD}", torch.cat([q, k], -1), torch.cat([rq, rk], -1), atol=4e-2, rtol=2e-2)
    dq = _rand(B, S, H, HD, seed=20)
    dk = _rand(B, S, H, HD, seed=21)
    dsrc = torch.zeros_like(src)
    ops.qk_rmsnorm_rope_bwd(dq, dk, src, H * HD, H, HD, wq, wk, wq1, wk1, s_split, cos, sin, dsrc=dsrc)
    torch.cuda.synchronize()
    (rq * dq.float()).sum().backward(retain_graph=True)
    (rk * dk.float()).sum().backward()
    ref = xf.grad[..., :2 * H * HD]
    r2 = _report("qk_rmsnorm_rope_bwd", dsrc[..., :2 * H * HD], ref, atol=3e-2 * float(ref.abs().max()), rtol=3e-2)
    r["bwd_ok"] = r2["ok"]
    r["bwd_max_err"] = r2["max_err"]
    r["v_untouched"] = bool((dsrc[..., 2 * H * HD:] == 0).all().item())
    r["ok"] = r["ok"] and r2["ok"] and r["v_untouched"]
    return r


def check_flow(B=2, Cc=16, Hh=32, Ww=48):
    lat = _rand(B, Cc, Hh, Ww, seed=1)
    noise = _rand(B, Cc, Hh, Ww, seed=2)
    sig = torch.tensor([0.25, 0.8125][:B] + [0.5] * max(0, B - 2), dtype=torch.float32, device=DEV)
    noisy, packed = ops.flow_prep_pack(lat, noise, sig)
    torch.cuda.synchronize()
    # eager bf16 chain exactly as the reference writes it (common.py:4953-4960, 4989-4991)
    grid = sig.view(B, 1, 1, 1).to(torch.bfloat16)
    ref_noisy = (1.0 - grid) * lat + grid * noise
    ref_packed = ref_noisy.view(B, Cc, Hh // 2, 2, Ww // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (Hh // 2) * (Ww // 2), Cc * 4)
    bit_noisy = bool(torch.equal(noisy, ref_noisy))
    bit_packed = bool(torch.equal(packed, ref_packed))
    pred = _rand(B, (Hh // 2) * (Ww // 2), Cc * 4, seed=3)
    loss, dpred = ops.flow_mse_loss(pred, lat, noise)
    torch.cuda.synchronize()
    pf = pred.float().requires_grad_(True)
    unp = pf.view(B, Hh // 2, Ww // 2, Cc, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, Cc, Hh, Ww)
    tgt = (noise - lat).float()
    ref_loss = torch.nn.functional.mse_loss(unp, tgt, reduction="none").mean(dim=(1, 2, 3)).mean()
    ref_loss.backward()
    lerr = abs(float(loss.item()) - float(ref_loss.item()))
    r = _report("flow_loss_grad", dpred, pf.grad, atol=1e-2 * float(pf.grad.abs().max()), rtol=1e-2)
    r.update({"name": "flow_prep_loss", "bit_exact_noisy": bit_noisy, "bit_exact_packed": bit_packed,
              "loss": float(loss.item()), "loss_ref": float(ref_loss.item()), "loss_err": lerr})
    r["ok"] = r["ok"] and bit_noisy and bit_packed and lerr <= 1e-5 * max(1.0, abs(float(ref_loss.item())))
    return r


def check_skinny(B=2, S=300, R=16, N=3072):
    L = _rand(B, S, R, seed=1)
    Rm = _rand(B, S, N, seed=2)
    out = ops.skinny_tn(L, Rm, alpha=0.5)
    torch.cuda.synchronize()
    ref = 0.5 * (L.float().reshape(-1, R).t() @ Rm.float().reshape(-1, N))
    return _report(f"skinny_tn_R{R}_N{N}", out, ref, atol=1e-3 * float(ref.abs().max()), rtol=1e-3)


# --------------------------------------------------------------------------------------------- registry
E = ops
CHECKS = {
    # GEMM: descriptor / pipeline basics first
    "gemm_basic": lambda: check_gemm(256, 256, 128),
    "gemm_k64": lambda: check_gemm(128, 256, 64),
    "gemm_deepk": lambda: check_gemm(256, 512, 1024),
    "gemm_tails": lambda: check_gemm(200, 328, 200, B=2),
    "gemm_bn128": lambda: check_gemm(256, 384, 256, tile=(1, 128)),
    "gemm_bn64": lambda: check_gemm(256, 64, 256, tile=(1, 64)),
    "gemm_mt2": lambda: check_gemm(512, 512, 256, tile=(2, 256)),
    "gemm_mt2_bn128": lambda: check_gemm(384, 256, 256, tile=(2, 128)),
    "gemm_persistent": lambda: check_gemm(4096, 3072, 512),
    "gemm_seg2": lambda: check_gemm(256, 256, 192, segs=[64]),
    "gemm_seg3_lora": lambda: check_gemm(256, 512, 256, segs=[128, 16], bias=True),
    "gemm_strided": lambda: check_gemm(200, 256, 136, B=2, strided=True, bias=True),
    "gemm_bias": lambda: check_gemm(256, 256, 128, bias=True),
    "gemm_gelu": lambda: check_gemm(256, 512, 128, bias=True, epi=E.EPI_GELU),
    "gemm_gate_res": lambda: check_gemm(256, 256, 128, B=2, bias=True, epi=E.EPI_GATE_RES, nan_to_num=True),
    "gemm_mul_dgelu": lambda: check_gemm(256, 256, 128, epi=E.EPI_MUL_DGELU),
    "gemm_add_res": lambda: check_gemm(256, 256, 128, epi=E.EPI_ADD_RES),
    "gemm_flux_shape": lambda: check_gemm(4096, 3072, 3072, B=1, bias=True),
    # attention
    "attn_fwd_256": lambda: check_attn_fwd(1, 2, 256),
    "attn_fwd_128": lambda: check_attn_fwd(1, 1, 128),
    "attn_fwd_512": lambda: check_attn_fwd(2, 3, 512),
    "attn_fwd_ragged": lambda: check_attn_fwd(1, 2, 333, 417),
    "attn_fwd_strided": lambda: check_attn_fwd(2, 4, 384, strided=True),
    "attn_fwd_hd64": lambda: check_attn_fwd(1, 2, 320, HD=64),
    "attn_fwd_long": lambda: check_attn_fwd(1, 2, 4608),
    "attn_fwd_bigscore": lambda: check_attn_fwd(1, 2, 512, qscale=4.0),
    "attn_bwd_256": lambda: check_attn_bwd(1, 2, 256),
    "attn_bwd_128": lambda: check_attn_bwd(1, 1, 128),
    "attn_bwd_ragged": lambda: check_attn_bwd(1, 2, 333, 417),
    "attn_bwd_hd64": lambda: check_attn_bwd(1, 2, 320, HD=64),
    "attn_bwd_long": lambda: check_attn_bwd(1, 2, 2304),
    # elementwise
    "ln_modulate": lambda: check_ln_modulate(),
    "ln_modulate_d1536": lambda: check_ln_modulate(D=1536),
    "qk_rmsnorm_rope": lambda: check_qk_rmsnorm_rope(),
    "qk_rmsnorm_rope_hd64": lambda: check_qk_rmsnorm_rope(HD=64),
    "flow": lambda: check_flow(),
    "skinny": lambda: check_skinny(),
    "skinny_r48": lambda: check_skinny(R=48, N=4096),
    "skinny_flux": lambda: check_skinny(B=4, S=4608, R=48, N=9216),
    "skinny_ragged": lambda: check_skinny(B=3, S=333, R=16, N=200),
    "skinny_cuda_core_path": lambda: check_skinny(B=2, S=100, R=16, N=250),
}
# BASELINE-geometry kernel shapes: the deepest K of the Flux step (fc2 / single-block proj_out, K = 12288 and
# the 2-segment 3072 + 12288 `cat([attn, mlp]) @ W_out`), and the attention backward at the full joint sequence
CHECKS["gemm_k12288"] = lambda: check_gemm(4608, 3072, 12288, B=1, bias=True)
CHECKS["gemm_seg2_3072_12288_gate_res"] = lambda: check_gemm(4608, 3072, 3072, B=1, segs=[12288], bias=True, epi=E.EPI_GATE_RES,
                                                              nan_to_num=True)
CHECKS["gemm_seg3_dgrad_12288_9216_48"] = lambda: check_gemm(4608, 3072, 12288, B=1, segs=[9216, 48])
CHECKS["attn_bwd_s4608"] = lambda: check_attn_bwd(1, 2, 4608)
CHECKS["attn_bwd_s4608_ragged_cross"] = lambda: check_attn_bwd(1, 2, 4096, 4608 - 77)
CHECKS["attn_fwd_hd64_long"] = lambda: check_attn_fwd(2, 3, 1255, HD=64)
CHECKS["attn_bwd_hd64_long"] = lambda: check_attn_bwd(2, 3, 1255, HD=64)
CHECKS["skinny_r96"] = lambda: check_skinny(B=2, S=700, R=96, N=1536)
CHECKS["skinny_r24"] = lambda: check_skinny(B=1, S=300, R=24, N=512)


# --------------------------------------------------------------------------------------------- VAE encode kernels
def check_conv3x3(B=2, H=24, W=40, Ci=64, Co=128, stride=1, bias=True, res=False):
    import torch.nn.functional as F
    x = _rand(B, H, W, Ci, seed=1)
    w = _rand(Co, Ci, 3, 3, scale=(9 * Ci) ** -0.5, seed=2)
    bv = _rand(Co, scale=0.1, seed=3) if bias else None
    Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
    rv = _rand(B, Ho, Wo, Co, seed=4) if res else None
    w9 = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous()
    out = ops.conv3x3_nhwc(x, w9, bv, rv, stride)
    torch.cuda.synchronize()
    xn = x.float().permute(0, 3, 1, 2)
    if stride == 1:
        ref = F.conv2d(xn, w.float(), bv.float() if bias else None, padding=1)
    else:
        ref = F.conv2d(F.pad(xn, (0, 1, 0, 1)), w.float(), bv.float() if bias else None, stride=2)
    ref = ref.permute(0, 2, 3, 1)
    if res:
        ref = ref + rv.float()
    return _report(f"conv3x3_s{stride}_{Ci}to{Co}_{H}x{W}", out, ref, atol=2e-2, rtol=1e-2)


def check_conv_in(B=2, H=20, W=36, C=128):
    import torch.nn.functional as F
    x = _rand(B, 3, H, W, seed=1)
    w = _rand(C, 3, 3, 3, scale=27 ** -0.5, seed=2)
    bv = _rand(C, scale=0.1, seed=3)
    out = ops.conv_in_3ch(x, w, bv)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float(), w.float(), bv.float(), padding=1).permute(0, 2, 3, 1)
    return _report("conv_in_3ch", out, ref, atol=1e-2, rtol=1e-2)


def check_groupnorm(B=2, H=16, W=24, C=128, G=32, silu=True, mean=0.0):
    import torch.nn.functional as F
    x = (_rand(B, H, W, C, seed=1).float() + mean).bfloat16()
    gm = (1.0 + 0.1 * _rand(C, seed=2).float()).bfloat16()
    bt = _rand(C, scale=0.1, seed=3)
    out = ops.groupnorm_nhwc(x, gm, bt, G, 1e-6, silu)
    torch.cuda.synchronize()
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), G, gm.float(), bt.float(), 1e-6)
    if silu:
        ref = F.silu(ref)
    return _report(f"groupnorm_C{C}_silu{int(silu)}_mean{mean}", out, ref.permute(0, 2, 3, 1), atol=2e-2, rtol=1e-2)


def check_softmax_rows(rows=300, cols=1024, scale=0.0442):
    s = _rand(rows, cols, scale=20.0, seed=1)
    ref = torch.softmax(s.float() * scale, dim=-1)
    ops.softmax_rows_(s, scale)
    torch.cuda.synchronize()
    return _report("softmax_rows", s, ref, atol=2e-3, rtol=1e-2)


def check_gaussian_sample(B=2, L=16, h=12, w=20, shift=0.1159):
    m = _rand(B, h, w, 2 * L, seed=1)
    eps = _rand(B, L, h, w, seed=2)
    out = ops.gaussian_sample_scale(m, eps, shift, 0.3611)
    torch.cuda.synchronize()
    mn = m.permute(0, 3, 1, 2)
    mean, logvar = mn[:, :L], mn[:, L:]
    z = mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * eps          # bf16 tensor ops, like diffusers
    ref = (z - shift) * 0.3611 if shift is not None else z * 0.3611
    r = _report("gaussian_sample_scale", out, ref, atol=0.0, rtol=0.0)
    if not r["ok"]:  # exp() ulp differences may move a bf16 rounding; never more than 1 bf16 ulp
        r2 = _report("gaussian_sample_scale", out, ref, atol=1e-3, rtol=2 ** -7)
        r2["bit_exact"] = False
        return r2
    r["bit_exact"] = True
    return r


CHECKS.update({
    "conv3x3_s1": lambda: check_conv3x3(),
    "conv3x3_s1_res": lambda: check_conv3x3(Ci=128, Co=256, res=True),
    "conv3x3_s1_wide": lambda: check_conv3x3(B=1, H=8, W=300, Ci=128, Co=128),
    "conv3x3_s1_512": lambda: check_conv3x3(B=1, H=16, W=16, Ci=512, Co=512, res=True),
    "conv3x3_out32": lambda: check_conv3x3(B=2, H=16, W=16, Ci=512, Co=32),
    "conv3x3_s2": lambda: check_conv3x3(stride=2, Ci=128, Co=128),
    "conv3x3_s2_wide": lambda: check_conv3x3(B=1, H=6, W=600, stride=2, Ci=64, Co=64),
    "conv_in_3ch": lambda: check_conv_in(),
    "groupnorm_silu": lambda: check_groupnorm(),
    "groupnorm_plain_512": lambda: check_groupnorm(C=512, silu=False),
    "groupnorm_256_offset": lambda: check_groupnorm(C=256, mean=3.0),
    "groupnorm_c64": lambda: check_groupnorm(C=64),
    "softmax_rows": lambda: check_softmax_rows(),
    "gaussian_sample": lambda: check_gaussian_sample(),
    "gaussian_sample_noshift": lambda: check_gaussian_sample(shift=None),
})


# --------------------------------------------------------------------------------------------- epsilon-family step kernels
def check_ddpm_prep(B=3, C=4, Hh=16, Ww=24):
    lat, noise = _rand(B, C, Hh, Ww, seed=1), _rand(B, C, Hh, Ww, seed=2)
    ac = torch.cumprod(1 - torch.linspace(1e-4, 0.02, 1000), 0).to(DEV)
    t = torch.tensor([3, 500, 999], device=DEV)[:B]
    ca, cb = (ac[t] ** 0.5).contiguous(), ((1 - ac[t]) ** 0.5).contiguous()
    noisy, packed = ops.ddpm_prep_pack(lat, noise, ca, cb)
    torch.cuda.synchronize()
    ref = (ca.view(-1, 1, 1, 1) * lat.float() + cb.view(-1, 1, 1, 1) * noise.float()).bfloat16()   # eager fp32 chain
    refp = ref.view(B, C, Hh // 2, 2, Ww // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, -1, C * 4)
    r = _report("ddpm_prep_pack", noisy, ref, atol=0.0, rtol=0.0)
    r["packed_exact"] = bool(torch.equal(packed, refp))
    r["ok"] = r["ok"] and r["packed_exact"]
    return r


def check_target_mse(B=3, C=4, Hh=16, Ww=24, weighted=True):
    S = (Hh // 2) * (Ww // 2)
    pred = _rand(B, S, 4 * C, seed=1).requires_grad_(False)
    tgt = _rand(B, C, Hh, Ww, seed=2)
    w = torch.tensor([0.5, 1.0, 2.5], device=DEV)[:B] if weighted else None
    loss, dpred = ops.target_mse_loss(pred, tgt, w, layout=1)
    torch.cuda.synchronize()
    p32 = pred.float().requires_grad_(True)
    un = torch.einsum("nhwpqc->nchpwq", p32.reshape(B, Hh // 2, Ww // 2, 2, 2, C)).reshape(B, C, Hh, Ww)
    l = (un - tgt.float()) ** 2
    if weighted:
        l = l * w.view(-1, 1, 1, 1)
    lref = l.mean(dim=[1, 2, 3]).mean()
    lref.backward()
    r = _report("target_mse_dpred", dpred, p32.grad, atol=1e-6, rtol=1e-2)
    r["loss"], r["loss_ref"] = float(loss.item()), float(lref.item())
    r["ok"] = r["ok"] and abs(r["loss"] - r["loss_ref"]) <= 1e-5 * abs(r["loss_ref"])
    return r


CHECKS.update({
    "ddpm_prep_pack": lambda: check_ddpm_prep(),
    "target_mse_weighted": lambda: check_target_mse(),
    "target_mse_plain": lambda: check_target_mse(weighted=False),
    "attn_fwd_cross_300": lambda: check_attn_fwd(2, 4, 1024, 300),
    "attn_bwd_cross_300": lambda: check_attn_bwd(2, 4, 1024, 300),
})


# --------------------------------------------------------------------------------------------- CTA-pair (cta_group::2) GEMM
CHECKS.update({
    "gemm_pair_basic": lambda: check_gemm(512, 512, 256, tile=(3, 256)),
    "gemm_pair_one_tile": lambda: check_gemm(256, 256, 64, tile=(3, 256)),
    "gemm_pair_tails": lambda: check_gemm(200, 328, 200, B=2, tile=(3, 256)),
    "gemm_pair_half_empty": lambda: check_gemm(100, 256, 128, B=3, tile=(3, 256)),
    "gemm_pair_seg3_lora": lambda: check_gemm(512, 512, 256, segs=[128, 16], bias=True, tile=(3, 256)),
    "gemm_pair_gate_res": lambda: check_gemm(512, 256, 128, B=2, bias=True, epi=E.EPI_GATE_RES, nan_to_num=True, tile=(3, 256)),
    "gemm_pair_gelu": lambda: check_gemm(256, 512, 128, bias=True, epi=E.EPI_GELU, tile=(3, 256)),
    "gemm_pair_persistent": lambda: check_gemm(4096, 3072, 512, tile=(3, 256)),
    "gemm_pair_flux_shape": lambda: check_gemm(4608, 3072, 3072, B=2, bias=True, tile=(3, 256)),
})


# --------------------------------------------------------------------------------------------- fused q/k-prep backward epilogue
def check_attn_bwd_fused_prep(B=2, S=333, H=3, HD=128, s_split=77, rope=True, norm_w=True):
    """attn_bwd(qk_prep=...) == attn_bwd -> qk_rmsnorm_rope_bwd (the two-kernel path) on the same inputs."""
    D = H * HD
    qkv = _rand(B, S, 3 * D, seed=1)
    d_o = _rand(B, S, H, HD, seed=2)
    mk = lambda sd: (1.0 + 0.1 * _rand(HD, seed=sd).float()).bfloat16() if norm_w else None
    wq, wk, wqa, wka = mk(3), mk(4), mk(5), mk(6)
    cos, sin = _rope_tables(S, HD) if rope else (None, None)
    q, k = ops.qk_rmsnorm_rope_fwd(qkv, D, H, HD, wq, wk, wqa, wka, s_split, cos, sin, 1e-6)
    v = qkv[:, :, 2 * D:].unflatten(-1, (H, HD))
    o, lse = ops.attn_fwd(q, k, v)
    # reference: two kernels
    dq, dk, dv = ops.attn_bwd(q, k, v, o, d_o, lse)
    ref = torch.zeros_like(qkv)
    ops.qk_rmsnorm_rope_bwd(dq, dk, qkv, D, H, HD, wq, wk, wqa, wka, s_split, cos, sin, 1e-6, dsrc=ref)
    ref[:, :, 2 * D:] = dv.reshape(B, S, D)
    got = torch.zeros_like(qkv)
    ops.attn_bwd(q, k, v, o, d_o, lse, dq=got[:, :, 0:D].unflatten(-1, (H, HD)), dk=got[:, :, D:2 * D].unflatten(-1, (H, HD)),
                 dv=got[:, :, 2 * D:].unflatten(-1, (H, HD)),
                 qk_prep=dict(src=qkv, k_off=D, wq=wq, wk=wk, wq_added=wqa, wk_added=wka, s_split=s_split, cos=cos, sin=sin, eps=1e-6))
    torch.cuda.synchronize()
    # the two-kernel path rounds dq / dk to bf16 before the norm backward, the fused one does not
    return _report(f"attn_bwd_fused_prep_S{S}_HD{HD}_rope{int(rope)}_w{int(norm_w)}", got, ref,
                   atol=2e-2 * float(ref.float().abs().max()), rtol=2e-2)


CHECKS.update({
    "attn_bwd_fused_prep": lambda: check_attn_bwd_fused_prep(),
    "attn_bwd_fused_prep_hd64_norope": lambda: check_attn_bwd_fused_prep(B=1, S=320, H=4, HD=64, s_split=0, rope=False),
    "attn_bwd_fused_prep_now": lambda: check_attn_bwd_fused_prep(B=1, S=256, H=2, HD=128, s_split=0, rope=True, norm_w=False),
})


# --------------------------------------------------------------------------------------------- CTA-pair implicit-GEMM conv
CHECKS.update({
    "conv3x3_pair_rows2": lambda: check_conv3x3(B=2, H=96, W=128, Ci=128, Co=128),
    "conv3x3_pair_oddrows": lambda: check_conv3x3(B=1, H=149, W=120, Ci=128, Co=128, res=True),
    "conv3x3_pair_wide": lambda: check_conv3x3(B=1, H=80, W=300, Ci=64, Co=256, res=True),
    "conv3x3_pair_512": lambda: check_conv3x3(B=2, H=64, W=128, Ci=512, Co=512),
    "conv3x3_pair_s2": lambda: check_conv3x3(B=2, H=160, W=256, stride=2, Ci=128, Co=128),
    "conv3x3_pair_s2_wide": lambda: check_conv3x3(B=1, H=160, W=600, stride=2, Ci=64, Co=256),
})

CHECKS.update({
    "conv_in_3ch_wide": lambda: check_conv_in(B=2, H=9, W=300, C=128),     # tile tail (300 = 2 * 128 + 44), tensor-core path
    "conv_in_3ch_c64": lambda: check_conv_in(B=1, H=33, W=130, C=64),
    "conv_in_3ch_c8_cuda_core": lambda: check_conv_in(B=1, H=12, W=20, C=8),  # C % 16 != 0 -> CUDA-core fallback kernel
})


# --------------------------------------------------------------------------------------------- full-rank weight gradient
def check_wgrad_full(B=2, S=300, N=256, K=320, strided=False, accumulate=False, alpha=1.0):
    if strided:   # row ranges / column ranges of larger buffers, like the joint hidden buffer and the fused qkv gradient
        dyf = _rand(B, S + 5, N + 64, seed=41)
        xf = _rand(B, S + 9, K + 32, seed=42)
        dy, x = dyf[:, 3:S + 3, 32:32 + N], xf[:, 4:S + 4, 8:8 + K]
    else:
        dy, x = _rand(B, S, N, seed=41), _rand(B, S, K, seed=42)
    ref = alpha * torch.einsum("bsn,bsk->nk", dy.float(), x.float())
    out = None
    if accumulate:
        out = _rand(N, K, seed=43)
        ref = ref + out.float()
    got = ops.wgrad_full(dy, x, out=out, alpha=alpha, accumulate=accumulate)
    torch.cuda.synchronize()
    return _report(f"wgrad_full_B{B}_S{S}_N{N}_K{K}", got, ref, atol=float(ref.abs().max()) * 6e-3, rtol=1.6e-2)


def check_colsum2(B=3, S=333, D=1536):
    dyf, zf = _rand(B, S + 3, D + 16, seed=51), _rand(B, S, D, seed=52)
    dy = dyf[:, 1:S + 1, 8:8 + D]
    sm, dt = ops.colsum2(dy, zf)
    torch.cuda.synchronize()
    r1 = _report("colsum", sm, dy.float().sum(1), atol=0.05, rtol=1e-3)
    r2 = _report("coldot", dt, (dy.float() * zf.float()).sum(1), atol=0.05, rtol=1e-3)
    only = ops.colsum2(dy, None)[0]
    return {"name": f"colsum2_B{B}_S{S}_D{D}", "ok": r1["ok"] and r2["ok"] and bool(torch.allclose(only, sm, atol=1e-3)),
            "max_err": max(r1["max_err"], r2["max_err"])}


CHECKS.update({
    "wgrad_full_basic": lambda: check_wgrad_full(),
    "wgrad_full_strided_accum": lambda: check_wgrad_full(B=3, S=231, N=384, K=200, strided=True, accumulate=True, alpha=0.5),
    "wgrad_full_sd3_qkv": lambda: check_wgrad_full(B=2, S=1255, N=4608, K=1536),
    "wgrad_full_narrow": lambda: check_wgrad_full(B=1, S=1000, N=128, K=64),
    "wgrad_full_flux_mlp": lambda: check_wgrad_full(B=1, S=4608, N=3072, K=12288),
    "colsum2": lambda: check_colsum2(),
    "colsum2_small": lambda: check_colsum2(B=1, S=7, D=64),
})


# --------------------------------------------------------------------------------------------- LyCORIS LoKr + stand-alone GELU
def check_lokr_rebuild(a=8, b=48, c=4, d=40, scale=0.5, fused=True):
    """out = bf16(W + kron(w1, w2) * scale) and its transpose, written into slices of larger (fused q|k|v style) buffers."""
    N, K = a * b, c * d
    W, w1, w2 = _rand(N, K, seed=61) * 0.05, _rand(a, c, seed=62), _rand(b, d, seed=63) * 0.05
    ref = (W.float() + torch.kron(w1.float(), w2.float()) * scale).bfloat16()
    if fused:
        big, big_t = torch.zeros(3 * N, K, device="cuda", dtype=torch.bfloat16), torch.zeros(K, 3 * N, device="cuda", dtype=torch.bfloat16)
        out, out_t = big[N:2 * N], big_t[:, N:2 * N]
    else:
        out, out_t = torch.empty(N, K, device="cuda", dtype=torch.bfloat16), None
    ops.lokr_rebuild(W, w1, w2, scale, out, out_t)
    torch.cuda.synchronize()
    # fp32 association differs (scale * w1 first), so a rare bf16 rounding tie may land on the neighbouring value
    close = lambda x, y: float((x != y).float().mean()) < 2e-3 and bool(torch.allclose(x.float(), y.float(), rtol=2 ** -7, atol=1e-6))
    ok = close(out, ref)
    if fused:
        ok = ok and bool(torch.equal(out_t, out.t())) and float(big[:N].abs().sum() + big[2 * N:].abs().sum() + big_t[:, :N].abs().sum()) == 0.0
    return {"name": f"lokr_rebuild_{a}x{b}_{c}x{d}", "ok": ok, "max_err": float((out.float() - ref.float()).abs().max())}


def check_lokr_factor_grads(a=8, b=48, c=4, d=40, scale=0.5, strided=True):
    N, K = a * b, c * d
    w1, w2 = _rand(a, c, seed=64), _rand(b, d, seed=65)
    g_full = _rand(N + 8, K + 16, seed=66)
    dW = g_full[4:4 + N, 8:8 + K] if strided else g_full[:N, :K].contiguous()
    g4 = dW.float().view(a, b, c, d) if not strided else dW.float().reshape(a, b, c, d)
    r1 = torch.einsum("ajcl,jl->ac", g4, w2.float()) * scale
    r2 = torch.einsum("ajcl,ac->jl", g4, w1.float()) * scale
    d1, d2 = ops.lokr_factor_grads(dW, w1, w2, scale)
    torch.cuda.synchronize()
    x1 = _report("lokr_dw1", d1, r1, atol=float(r1.abs().max()) * 1e-4, rtol=1e-4)
    x2 = _report("lokr_dw2", d2, r2, atol=float(r2.abs().max()) * 1e-4, rtol=1e-4)
    return {"name": f"lokr_factor_grads_{a}x{b}_{c}x{d}", "ok": x1["ok"] and x2["ok"], "max_err": max(x1["max_err"], x2["max_err"])}


def check_gelu_tanh(B=2, S=77, D=640):
    import torch.nn.functional as F
    pf, gf = _rand(B, S + 2, D + 8, seed=71), _rand(B, S, D, seed=72)
    pre = pf[:, 1:S + 1, :D]
    act = ops.gelu_tanh(pre)
    ref = F.gelu(pre.float(), approximate="tanh")
    r1 = _report("gelu", act, ref, atol=2e-2, rtol=1.6e-2)
    x = pre.float().requires_grad_(True)
    (F.gelu(x, approximate="tanh") * gf.float()).sum().backward()
    out = ops.mul_dgelu_tanh(gf, pre)
    r2 = _report("dgelu", out, x.grad, atol=3e-2, rtol=1.6e-2)
    inplace = gf.clone()
    ops.mul_dgelu_tanh(inplace, pre, out=inplace)
    torch.cuda.synchronize()
    return {"name": f"gelu_tanh_B{B}_S{S}_D{D}", "ok": r1["ok"] and r2["ok"] and bool(torch.equal(inplace, out)), "max_err": max(r1["max_err"], r2["max_err"])}


def check_gelu_matches_gemm_epilogue(B=1, S=256, N=512, K=128):
    """ops.gelu_tanh(pre) must be BIT-identical to the activation EPI_GELU wrote next to `pre` (the LoRA-on-fc2 weight gradient
    re-creates the activation from the saved pre-activation)."""
    x, w, bias = _rand(B, S, K, seed=73), _rand(N, K, seed=74) * 0.2, _rand(N, seed=75)
    pre = torch.empty(B, S, N, device="cuda", dtype=torch.bfloat16)
    act = ops.gemm([x], [w], bias, epi=E.EPI_GELU, aux=pre)
    again = ops.gelu_tanh(pre)
    torch.cuda.synchronize()
    return {"name": "gelu_matches_gemm_epilogue", "ok": bool(torch.equal(act, again)), "max_err": float((act.float() - again.float()).abs().max())}


CHECKS.update({
    "lokr_rebuild_fused_slices": lambda: check_lokr_rebuild(),
    "lokr_rebuild_plain_ragged": lambda: check_lokr_rebuild(a=3, b=37, c=5, d=13, scale=1.0, fused=False),
    "lokr_rebuild_flux_attn": lambda: check_lokr_rebuild(a=8, b=384, c=8, d=384, scale=1.0),
    "lokr_factor_grads": lambda: check_lokr_factor_grads(),
    "lokr_factor_grads_flux_ff": lambda: check_lokr_factor_grads(a=4, b=3072, c=4, d=768, scale=1.0, strided=False),
    "lokr_factor_grads_tiny": lambda: check_lokr_factor_grads(a=2, b=3, c=2, d=8, strided=False),
    "gelu_tanh": lambda: check_gelu_tanh(),
    "gelu_matches_gemm_epilogue": lambda: check_gelu_matches_gemm_epilogue(),
})


# --------------------------------------------------------------------------------------------- [K, N] weight segments (dgrad on W itself)
CHECKS.update({
    "gemm_wkn_basic": lambda: check_gemm(256, 256, 128, w_kn=[True]),
    "gemm_wkn_bn64_tails": lambda: check_gemm(200, 72, 200, B=2, w_kn=[True], tile=(1, 64)),
    "gemm_wkn_bn128_ktail": lambda: check_gemm(300, 136, 72, w_kn=[True], tile=(1, 128)),
    "gemm_wkn_strided_bias": lambda: check_gemm(333, 320, 192, B=2, bias=True, strided=True, w_kn=[True]),
    "gemm_wkn_pair": lambda: check_gemm(4096, 3072, 512, w_kn=[True], tile=(3, 256)),
    "gemm_wkn_pair_tails": lambda: check_gemm(700, 328, 456, B=2, w_kn=[True], tile=(3, 256)),
    "gemm_wkn_mixed_segments": lambda: check_gemm(512, 384, 256, B=2, segs=[48, 128], w_kn=[True, False, True], tile=(3, 256)),
    "gemm_wkn_dgelu_epilogue": lambda: check_gemm(256, 512, 128, epi=E.EPI_MUL_DGELU, w_kn=[True]),
    "gemm_wkn_flux_dgrad": lambda: check_gemm(4608, 3072, 9216, w_kn=[True], tile=(3, 256)),
})
