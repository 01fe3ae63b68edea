"""Flux MMDiT blocks as explicit forward/backward kernel schedules (torch.autograd.Function).

Mirrors, for the training path, reference flux/transformer.py:
  * FluxTransformerBlock.forward + _ffn_forward         (:563-687)  -> DoubleBlockFn
  * FluxSingleTransformerBlock.forward + _ffn_forward   (:453-510)  -> SingleBlockFn
  * FluxAttnProcessor2_0.__call__                       (:116-224)  (inlined in both)
  * PEFT lora.Linear on the attention projections       (common.py:1094-1117, flux/model.py:1249-1262)

Every arithmetic step is a libstb200 kernel (ops.*); torch only allocates, slices and carries the
autograd graph between blocks.  Activations saved per block are the minimum the LoRA-only backward
needs (block input, fused pre-norm QKV, attention output + LSE, post-attention stream, MLP
pre-activation, the rank-r LoRA down-projections); LayerNorm-modulate, QK-RMSNorm/RoPE and GELU are
recomputed inside the backward kernels instead of being stored.

The hidden state travels as ONE joint buffer [B, S_txt + S_img, D] (text tokens first, as the
reference concatenates them for attention, :166-168) so no torch.cat / split copies are needed.

Base weights are frozen (LoRA training, BASELINE config 2): gradients are produced for the LoRA
A/B matrices and for the hidden state only; no gradient flows to temb / the adaLN linears because
none of their parameters is trainable in that configuration (full fine-tune is a later row).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops

EPS = 1e-6


# ------------------------------------------------------------------------------------------------
# derived (frozen) weight layouts
# ------------------------------------------------------------------------------------------------
@dataclass
class AttnPlan:
    """Fused / transposed copies of one attention's projection weights (one per stream)."""
    w_qkv: torch.Tensor      # [3D, D]   rows: q | k | v
    b_qkv: torch.Tensor      # [3D]
    w_qkv_t: torch.Tensor    # [D, 3D]   dgrad weight
    w_out: Optional[torch.Tensor] = None    # [D, D]  (None for single blocks: pre_only attention)
    b_out: Optional[torch.Tensor] = None
    w_out_t: Optional[torch.Tensor] = None  # [D, D]
    norm_q: Optional[torch.Tensor] = None   # [hd]
    norm_k: Optional[torch.Tensor] = None


@dataclass
class MlpPlan:
    w1: torch.Tensor     # [4D, D]
    b1: torch.Tensor
    w1_t: torch.Tensor   # [D, 4D]
    w2: torch.Tensor     # [D, 4D]  (single block: proj_out [D, 5D], mlp part = columns D:)
    b2: torch.Tensor
    w2_t: torch.Tensor   # [4D, D]  (single block: [5D, D])


def _t(w: torch.Tensor) -> torch.Tensor:
    return w.t().contiguous()


@dataclass
class LoraPack:
    """Per-step packed view of the LoRA matrices of one fused projection group.

    a_stack  [R, K]   rows: A of each member (R = n_members * r)
    a_stack_t [K, R]
    b_ext    [N, R]   block structured: member m owns rows n_m and columns [m r, (m+1) r); scaling folded in
    b_ext_t  [R, N]
    """
    a_stack: torch.Tensor
    a_stack_t: torch.Tensor
    b_ext: torch.Tensor
    b_ext_t: torch.Tensor
    rank: int
    scaling: float
    members: List[Optional[int]]  # index into the flat (A, B) parameter list, None = member not adapted
    n_out: int                    # rows per member


def pack_lora(params: List[Optional[Tuple[torch.Tensor, torch.Tensor]]], n_out: int, k_in: int, scaling: float,
              device, dtype=torch.bfloat16) -> Optional[LoraPack]:
    """params[m] = (A [r, K], B [n_out, r]) or None for each member of the fused projection."""
    present = [p for p in params if p is not None]
    if not present:
        return None
    r = present[0][0].shape[0]
    M = len(params)
    a_stack = torch.zeros((M * r, k_in), device=device, dtype=dtype)
    b_ext = torch.zeros((M * n_out, M * r), device=device, dtype=dtype)
    members: List[Optional[int]] = []
    for m, p in enumerate(params):
        if p is None:
            members.append(None)
            continue
        a, b = p
        a_stack[m * r:(m + 1) * r].copy_(a.detach())
        blk = b_ext[m * n_out:(m + 1) * n_out, m * r:(m + 1) * r]
        blk.copy_(b.detach())
        if scaling != 1.0:
            blk.mul_(scaling)
        members.append(m)
    return LoraPack(a_stack, _t(a_stack), b_ext, _t(b_ext), r, scaling, members, n_out)


def _lora_down(x: torch.Tensor, pack: LoraPack) -> torch.Tensor:
    """T = x A_stack^T  [B, S, R]."""
    return ops.gemm([x], [pack.a_stack])


def _lora_grads(pack: LoraPack, x: torch.Tensor, t_down: torch.Tensor, dy: torch.Tensor, t_up: torch.Tensor):
    """dA_stack [R, K] = T'^T x  (T' = dy B_ext, scaling already inside);  dB_ext^T [R, N] = T^T dy."""
    R = pack.a_stack.shape[0]
    d_a = ops.skinny_tn(t_up, x)        # [R, K] fp32
    d_bt = ops.skinny_tn(t_down, dy)    # [R, N] fp32
    out = []
    r = pack.rank
    for m, idx in enumerate(pack.members):
        if idx is None:
            out.append((None, None))
            continue
        da = d_a[m * r:(m + 1) * r]
        dbt = d_bt[m * r:(m + 1) * r, m * pack.n_out:(m + 1) * pack.n_out]
        db = dbt.t()
        if pack.scaling != 1.0:
            db = db * pack.scaling
        out.append((da.to(torch.bfloat16), db.to(torch.bfloat16).contiguous()))
    return out


# ------------------------------------------------------------------------------------------------
# shared pieces
# ------------------------------------------------------------------------------------------------
def _linear_lora_fwd(x, w, b, pack: Optional[LoraPack], **kw):
    """y = x W^T + b (+ T B_ext^T as an extra K-segment).  Returns (y, T or None)."""
    if pack is None:
        return ops.gemm([x], [w], b, **kw), None
    t = _lora_down(x, pack)
    return ops.gemm([x, t], [w, pack.b_ext], b, **kw), t


def _linear_lora_dgrad(dy, w_t, pack: Optional[LoraPack], **kw):
    """dx = dy W (+ (dy B_ext) A_stack).  Returns (dx, T' or None)."""
    if pack is None:
        return ops.gemm([dy], [w_t], None, **kw), None
    t_up = ops.gemm([dy], [pack.b_ext_t])
    return ops.gemm([dy, t_up], [w_t, pack.a_stack_t], None, **kw), t_up


# ------------------------------------------------------------------------------------------------
# Double-stream block
# ------------------------------------------------------------------------------------------------
class DoubleBlockFn(torch.autograd.Function):
    """h_out = FluxTransformerBlock(h_in) on the joint [B, S_txt + S_img, D] buffer.

    inputs: h, mod_img [B, 6D], mod_txt [B, 6D] (adaLN vectors, no grad), cos, sin, static (plans /
    sizes), then the flat LoRA tensors in the order
      img: to_q.A, to_q.B, to_k.A, to_k.B, to_v.A, to_v.B, to_out.A, to_out.B,
      txt: add_q.A, add_q.B, add_k.A, ..., to_add_out.A, to_add_out.B      (None where not adapted)
    """

    @staticmethod
    def forward(ctx, h, mod_img, mod_txt, cos, sin, st, *lora):
        B, S, D = h.shape
        S_txt, H, hd = st["S_txt"], st["H"], st["hd"]
        plans: Dict[str, object] = st["plans"]
        scaling = st["lora_scaling"]
        dev = h.device
        streams = (("txt", slice(0, S_txt), mod_txt, 8), ("img", slice(S_txt, S), mod_img, 0))

        def lp(base, n_members, n_out, k_in):
            ps = []
            for m in range(n_members):
                a, b = lora[base + 2 * m], lora[base + 2 * m + 1]
                ps.append(None if a is None else (a, b))
            return pack_lora(ps, n_out, k_in, scaling, dev)

        packs = {}
        qkv = torch.empty((B, S, 3 * D), device=dev, dtype=torch.bfloat16)
        saved_small = {}
        for name, sl, mod, base in streams:
            ap: AttnPlan = plans[name + "_attn"]
            packs[name + "_qkv"] = lp(base, 3, D, D)
            packs[name + "_out"] = lp(base + 6, 1, D, D)
            nh = ops.ln_modulate_fwd(h[:, sl], mod[:, 0:D], mod[:, D:2 * D], EPS)
            _, t = _linear_lora_fwd(nh, ap.w_qkv, ap.b_qkv, packs[name + "_qkv"], out=qkv[:, sl])
            saved_small[name + "_t_qkv"] = t
        ia: AttnPlan = plans["img_attn"]
        ta: AttnPlan = plans["txt_attn"]
        q, k = ops.qk_rmsnorm_rope_fwd(qkv, D, H, hd, ia.norm_q, ia.norm_k, ta.norm_q, ta.norm_k, S_txt, cos, sin, EPS)
        v = qkv[:, :, 2 * D:].unflatten(-1, (H, hd))
        o, lse = ops.attn_fwd(q, k, v)
        del q, k
        o = o.view(B, S, D)
        h1 = torch.empty_like(h)
        h2 = torch.empty_like(h)
        mlp_pre = {}
        for name, sl, mod, base in streams:
            ap = plans[name + "_attn"]
            mp: MlpPlan = plans[name + "_mlp"]
            _, t = _linear_lora_fwd(o[:, sl], ap.w_out, ap.b_out, packs[name + "_out"], out=h1[:, sl],
                                    epi=ops.EPI_GATE_RES, gate=mod[:, 2 * D:3 * D], res=h[:, sl])
            saved_small[name + "_t_out"] = t
            nh2 = ops.ln_modulate_fwd(h1[:, sl], mod[:, 3 * D:4 * D], mod[:, 4 * D:5 * D], EPS)
            pre = torch.empty((B, sl.stop - sl.start, 4 * D), device=dev, dtype=torch.bfloat16)
            act = ops.gemm([nh2], [mp.w1], mp.b1, epi=ops.EPI_GELU, aux=pre)
            del nh2
            ops.gemm([act], [mp.w2], mp.b2, out=h2[:, sl], epi=ops.EPI_GATE_RES, gate=mod[:, 5 * D:6 * D],
                     res=h1[:, sl], nan_to_num=(name == "txt"))
            del act
            mlp_pre[name] = pre
        ctx.st = st
        ctx.packs = packs
        ctx.n_lora = len(lora)
        ctx.lora_present = [x is not None for x in lora]
        ctx.save_for_backward(h, mod_img, mod_txt, cos, sin, qkv, o, lse, h1, mlp_pre["txt"], mlp_pre["img"],
                              *(saved_small[k] if saved_small[k] is not None else h.new_empty(0)
                                for k in ("txt_t_qkv", "img_t_qkv", "txt_t_out", "img_t_out")))
        return h2

    @staticmethod
    def backward(ctx, dh2):
        (h, mod_img, mod_txt, cos, sin, qkv, o, lse, h1, pre_txt, pre_img,
         t_qkv_txt, t_qkv_img, t_out_txt, t_out_img) = ctx.saved_tensors
        st = ctx.st
        packs = ctx.packs
        B, S, D = h.shape
        S_txt, H, hd = st["S_txt"], st["H"], st["hd"]
        plans = st["plans"]
        dev = h.device
        dh2 = dh2.contiguous()
        streams = (("txt", slice(0, S_txt), mod_txt, 8, pre_txt, t_qkv_txt, t_out_txt),
                   ("img", slice(S_txt, S), mod_img, 0, pre_img, t_qkv_img, t_out_img))
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n_lora
        dh1 = torch.empty_like(h)
        d_o = torch.empty_like(o)
        for name, sl, mod, base, pre, t_qkv, t_out in streams:
            ap: AttnPlan = plans[name + "_attn"]
            mp: MlpPlan = plans[name + "_mlp"]
            # ---- MLP branch: h2 = h1 + gate_mlp * fc2(gelu(fc1(LNmod(h1))))
            g2 = ops.gate_mul(dh2[:, sl], mod[:, 5 * D:6 * D])
            d_pre = ops.gemm([g2], [mp.w2_t], None, epi=ops.EPI_MUL_DGELU, aux=pre)
            del g2
            d_nh2 = ops.gemm([d_pre], [mp.w1_t], None)
            del d_pre
            ops.ln_modulate_bwd(d_nh2, h1[:, sl], mod[:, 4 * D:5 * D], add=dh2[:, sl], eps=EPS, out=dh1[:, sl])
            del d_nh2
            # ---- attention output projection: h1 = h + gate_msa * to_out(o)
            g1 = ops.gate_mul(dh1[:, sl], mod[:, 2 * D:3 * D])
            pk = packs[name + "_out"]
            _, t_up = _linear_lora_dgrad(g1, ap.w_out_t, pk, out=d_o[:, sl])
            if pk is not None:
                (da, db), = _lora_grads(pk, o[:, sl], t_out, g1, t_up)
                grads[base + 6], grads[base + 7] = da, db
            del g1
        # ---- attention core
        ia: AttnPlan = plans["img_attn"]
        ta: AttnPlan = plans["txt_attn"]
        q, k = ops.qk_rmsnorm_rope_fwd(qkv, D, H, hd, ia.norm_q, ia.norm_k, ta.norm_q, ta.norm_k, S_txt, cos, sin, EPS)
        v = qkv[:, :, 2 * D:].unflatten(-1, (H, hd))
        d_qkv = torch.empty_like(qkv)
        dq = torch.empty_like(q)
        dk = torch.empty_like(k)
        ops.attn_bwd(q, k, v, o.view(B, S, H, hd), d_o.view(B, S, H, hd), lse, dq=dq, dk=dk,
                     dv=d_qkv[:, :, 2 * D:].unflatten(-1, (H, hd)))
        del q, k
        ops.qk_rmsnorm_rope_bwd(dq, dk, qkv, D, H, hd, ia.norm_q, ia.norm_k, ta.norm_q, ta.norm_k, S_txt, cos, sin,
                                EPS, dsrc=d_qkv)
        del dq, dk
        dh = torch.empty_like(h)
        for name, sl, mod, base, pre, t_qkv, t_out in streams:
            ap = plans[name + "_attn"]
            pk = packs[name + "_qkv"]
            d_nh, t_up = _linear_lora_dgrad(d_qkv[:, sl], ap.w_qkv_t, pk)
            if pk is not None:
                nh = ops.ln_modulate_fwd(h[:, sl], mod[:, 0:D], mod[:, D:2 * D], EPS)
                for m, (da, db) in enumerate(_lora_grads(pk, nh, t_qkv, d_qkv[:, sl], t_up)):
                    grads[base + 2 * m], grads[base + 2 * m + 1] = da, db
                del nh
            ops.ln_modulate_bwd(d_nh, h[:, sl], mod[:, D:2 * D], add=dh1[:, sl], eps=EPS, out=dh[:, sl])
            del d_nh
        for i, present in enumerate(ctx.lora_present):
            if not present:
                grads[i] = None
        return (dh, None, None, None, None, None, *grads)


# ------------------------------------------------------------------------------------------------
# Single-stream block
# ------------------------------------------------------------------------------------------------
class SingleBlockFn(torch.autograd.Function):
    """h_out = FluxSingleTransformerBlock(h_in); LoRA order: to_q.A, to_q.B, to_k.A, to_k.B, to_v.A, to_v.B."""

    @staticmethod
    def forward(ctx, h, mod, cos, sin, st, *lora):
        B, S, D = h.shape
        H, hd = st["H"], st["hd"]
        plans = st["plans"]
        ap: AttnPlan = plans["attn"]
        mp: MlpPlan = plans["mlp"]
        dev = h.device
        ps = []
        for m in range(3):
            a, b = lora[2 * m], lora[2 * m + 1]
            ps.append(None if a is None else (a, b))
        pk = pack_lora(ps, D, D, st["lora_scaling"], dev)
        nh = ops.ln_modulate_fwd(h, mod[:, 0:D], mod[:, D:2 * D], EPS)
        qkv, t_qkv = _linear_lora_fwd(nh, ap.w_qkv, ap.b_qkv, pk)
        q, k = ops.qk_rmsnorm_rope_fwd(qkv, D, H, hd, ap.norm_q, ap.norm_k, None, None, 0, cos, sin, EPS)
        v = qkv[:, :, 2 * D:].unflatten(-1, (H, hd))
        o, lse = ops.attn_fwd(q, k, v)
        del q, k
        o = o.view(B, S, D)
        pre = torch.empty((B, S, 4 * D), device=dev, dtype=torch.bfloat16)
        act = ops.gemm([nh], [mp.w1], mp.b1, epi=ops.EPI_GELU, aux=pre)
        del nh
        # proj_out(cat[attn, mlp]) as two K-segments of one GEMM; gate, residual, nan_to_num in the epilogue
        h_out = ops.gemm([o, act], [mp.w2[:, :D], mp.w2[:, D:]], mp.b2, epi=ops.EPI_GATE_RES,
                         gate=mod[:, 2 * D:3 * D], res=h, nan_to_num=True)
        del act
        ctx.st = st
        ctx.pack = pk
        ctx.lora_present = [x is not None for x in lora]
        ctx.save_for_backward(h, mod, cos, sin, qkv, o, lse, pre, t_qkv if t_qkv is not None else h.new_empty(0))
        return h_out

    @staticmethod
    def backward(ctx, dh_out):
        h, mod, cos, sin, qkv, o, lse, pre, t_qkv = ctx.saved_tensors
        st = ctx.st
        pk: Optional[LoraPack] = ctx.pack
        B, S, D = h.shape
        H, hd = st["H"], st["hd"]
        ap: AttnPlan = st["plans"]["attn"]
        mp: MlpPlan = st["plans"]["mlp"]
        dh_out = dh_out.contiguous()
        g = ops.gate_mul(dh_out, mod[:, 2 * D:3 * D])
        d_o = ops.gemm([g], [mp.w2_t[:D]], None)
        d_pre = ops.gemm([g], [mp.w2_t[D:]], None, epi=ops.EPI_MUL_DGELU, aux=pre)
        del g
        q, k = ops.qk_rmsnorm_rope_fwd(qkv, D, H, hd, ap.norm_q, ap.norm_k, None, None, 0, cos, sin, EPS)
        v = qkv[:, :, 2 * D:].unflatten(-1, (H, hd))
        d_qkv = torch.empty_like(qkv)
        dq = torch.empty_like(q)
        dk = torch.empty_like(k)
        ops.attn_bwd(q, k, v, o.view(B, S, H, hd), d_o.view(B, S, H, hd), lse, dq=dq, dk=dk,
                     dv=d_qkv[:, :, 2 * D:].unflatten(-1, (H, hd)))
        del q, k, d_o
        ops.qk_rmsnorm_rope_bwd(dq, dk, qkv, D, H, hd, ap.norm_q, ap.norm_k, None, None, 0, cos, sin, EPS, dsrc=d_qkv)
        del dq, dk
        grads: List[Optional[torch.Tensor]] = [None] * 6
        # d_nh = d_pre W_mlp + d_qkv W_qkv (+ LoRA)  — one GEMM, two/three K-segments
        if pk is None:
            d_nh = ops.gemm([d_pre, d_qkv], [mp.w1_t, ap.w_qkv_t], None)
        else:
            t_up = ops.gemm([d_qkv], [pk.b_ext_t])
            d_nh = ops.gemm([d_pre, d_qkv, t_up], [mp.w1_t, ap.w_qkv_t, pk.a_stack_t], None)
            nh = ops.ln_modulate_fwd(h, mod[:, 0:D], mod[:, D:2 * D], EPS)
            for m, (da, db) in enumerate(_lora_grads(pk, nh, t_qkv, d_qkv, t_up)):
                grads[2 * m], grads[2 * m + 1] = da, db
            del nh
        del d_pre, d_qkv
        dh = ops.ln_modulate_bwd(d_nh, h, mod[:, D:2 * D], add=dh_out, eps=EPS)
        for i, present in enumerate(ctx.lora_present):
            if not present:
                grads[i] = None
        return (dh, None, None, None, None, *grads)


# ------------------------------------------------------------------------------------------------
# Tail: AdaLayerNormContinuous + proj_out   (reference flux/transformer.py:1503-1506)
# ------------------------------------------------------------------------------------------------
class TailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, mod, st):
        """h [B, S, D] joint buffer; mod [B, 2D] = (scale | shift) (AdaLayerNormContinuous chunk order)."""
        D = h.shape[2]
        S_txt = st["S_txt"]
        x = h[:, S_txt:]
        nx = ops.ln_modulate_fwd(x, mod[:, D:2 * D], mod[:, 0:D], EPS)
        out = ops.gemm([nx], [st["w_proj"]], st["b_proj"])
        ctx.st = st
        ctx.save_for_backward(h, mod)
        return out

    @staticmethod
    def backward(ctx, d_out):
        h, mod = ctx.saved_tensors
        st = ctx.st
        D = h.shape[2]
        S_txt = st["S_txt"]
        d_nx = ops.gemm([d_out.contiguous()], [st["w_proj_t"]], None)
        dh = torch.zeros_like(h) if S_txt > 0 else torch.empty_like(h)
        ops.ln_modulate_bwd(d_nx, h[:, S_txt:], mod[:, 0:D], add=None, eps=EPS, out=dh[:, S_txt:])
        return dh, None, None


class FlowLossFn(torch.autograd.Function):
    """loss = mean_b mean_chw (unpack(pred) - (noise - latents))^2 ; backward = precomputed d loss/d pred."""

    @staticmethod
    def forward(ctx, pred_packed, latents, noise):
        loss, dpred = ops.flow_mse_loss(pred_packed.contiguous(), latents, noise, want_grad=True)
        ctx.save_for_backward(dpred)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        # g is the scalar upstream gradient (1.0 for loss.backward()); keep it on-device
        return (dpred * g.to(dpred.dtype)), None, None
