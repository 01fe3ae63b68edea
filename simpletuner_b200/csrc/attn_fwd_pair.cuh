// simpletuner_b200 — attention forward, CTA-pair (cta_group::2) version for head_dim 128, sm_100a.
//
// Same contract as attn_fwd.cuh (O = softmax(Q K^T * scale) V and the natural-log LSE), different mapping, chosen
// from the measured MMA rates (tools/probe/mma_probe.cu, profiles/r01/mma_probe.log): SS-form MMAs are limited by
// the ~75 B/clk/SM operand fetch from shared memory (128x128x16: 107 cycles against a 64-cycle floor), TS-form
// N = 128 MMAs run at 85-100 % of the floor, and in pair mode the B operand is split between the two SMs.
//
// A 2-CTA cluster (one TPC) owns 256 query rows of one (batch, head); CTA r owns rows [128 r, 128 r + 128):
//   * every 128-key K tile is split by KEY ROWS (CTA r stages keys [64 r, +64)), every V tile by HEAD-DIM COLUMNS
//     (CTA r stages columns [64 r, +64)): exactly the halves of the B operands the pair MMA reads from each SM,
//     so each SM stages 32 KB per key tile (as the single-CTA kernel did for two Q tiles);
//   * the leader CTA issues M = 256 MMAs for both SMs; commits are multicast to both CTAs;
//   * S is double-buffered in TMEM across key tiles (S_{j+1} is computed while the compute warps run the softmax
//     of tile j);
//   * the two compute warpgroups are INDEPENDENT: warpgroup w owns the 64-key half w of every key tile with its own
//     running max / row sum and its own output accumulator O_w (O = O_0 + O_1 are merged with the usual
//     2^(m_w - M) weights in the epilogue).  No per-tile max exchange, so the warpgroups drift apart and one's
//     MUFU-bound exp pass overlaps the other's TMEM reads / max pass; the leader issues PV_w as soon as P_w is ready.
// TMEM (per CTA): S0 [0,128)  S1 [128,256)  O_0 [256,384)  O_1 [384,512).  Q stays in shared memory (SS-form S MMA).
// warps 0-7 compute, warp 8 TMA producer, warp 9 MMA issuer (leader only), warp 10 TMEM allocator.
#pragma once
#include "attn_fwd.cuh"

#ifndef STB_ATTN_DEBUG_SKIP
#define STB_ATTN_DEBUG_SKIP 0
#endif

namespace stb {

struct AttnFwdPairMaps {
  CUtensorMap q128;  // 4-D (d, h, s, b), box (64, 1, 128, 1): this CTA's 128 query rows, one 64-wide d atom
  CUtensorMap k64;   // box (64, 1, 64, 1):  half of a key tile's rows, one 64-wide d atom
  CUtensorMap v128;  // box (64, 1, 128, 1): all keys of a tile, one 64-wide d atom (this CTA's half of the columns)
};

struct AttnFwdPairParams {
  AttnFwdParams base;
};

struct AttnFwdPairCfg {
  static constexpr int HD = 128;
  static constexpr int Q_BYTES = 128 * HD * 2;   // this CTA's Q tile
  static constexpr int K_BYTES = 64 * HD * 2;    // this CTA's half of a key tile (64 keys x 128)
  static constexpr int V_BYTES = 128 * 64 * 2;   // this CTA's half of a value tile (128 keys x 64 columns)
  static constexpr int STAGES = 4;
  static constexpr int SCRATCH = 2 * 2 * 128 * 4;   // (m_w, l_w) exchange for the epilogue merge
  static constexpr int SMEM_BYTES = Q_BYTES + STAGES * (K_BYTES + V_BYTES) + SCRATCH + 1024 + 256;
};

__global__ void __launch_bounds__(384, 1)
attn_fwd_pair_kernel(const __grid_constant__ AttnFwdPairMaps maps, const AttnFwdPairParams pp) {
  using Cfg = AttnFwdPairCfg;
  constexpr int HD = Cfg::HD;
  constexpr int NSTG = Cfg::STAGES;
  const AttnFwdParams& p = pp.base;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  auto k_smem = [&](int s) { return smem_base + Cfg::Q_BYTES + uint32_t(s) * Cfg::K_BYTES; };
  auto v_smem = [&](int s) { return smem_base + Cfg::Q_BYTES + NSTG * Cfg::K_BYTES + uint32_t(s) * Cfg::V_BYTES; };
  const uint32_t scratch = smem_base + Cfg::Q_BYTES + NSTG * (Cfg::K_BYTES + Cfg::V_BYTES);
  const uint32_t bar_base = scratch + Cfg::SCRATCH;
  auto k_full = [&](int s) { return bar_base + 8u * s; };                 // leader's copy is the live one
  auto v_full = [&](int s) { return bar_base + 8u * (NSTG + s); };        // leader
  auto k_empty = [&](int s) { return bar_base + 8u * (2 * NSTG + s); };   // both CTAs (multicast commit)
  auto v_empty = [&](int s) { return bar_base + 8u * (3 * NSTG + s); };   // both
  auto s_full = [&](int u) { return bar_base + 8u * (4 * NSTG + u); };    // both
  auto p_full = [&](int u, int w) { return bar_base + 8u * (4 * NSTG + 2 + 2 * u + w); };  // leader: 4 warps x 2 CTAs
  auto o_done = [&](int w) { return bar_base + 8u * (4 * NSTG + 6 + w); };                 // both
  const uint32_t q_full = bar_base + 8u * (4 * NSTG + 8);                 // leader
  const uint32_t tmem_slot = bar_base + 8u * (4 * NSTG + 9);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* scratch_f = reinterpret_cast<float*>(smem_raw + (scratch - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int q0 = int(blockIdx.x >> 1) * 256 + int(rank) * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.Sk + 127) / 128;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&maps.q128);
    tma_prefetch_desc(&maps.k64);
    tma_prefetch_desc(&maps.v128);
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(v_full(s), 1);
      mbar_init(k_empty(s), 1);
      mbar_init(v_empty(s), 1);
    }
    for (int u = 0; u < 2; ++u) {
      mbar_init(s_full(u), 1);
      mbar_init(p_full(u, 0), 8);
      mbar_init(p_full(u, 1), 8);
      mbar_init(o_done(u), 1);
    }
    mbar_init(q_full, 1);
    fence_mbar_init();
  }
  if (warp == 10) {
    tmem_alloc2(tmem_slot, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  auto S_col = [&](int u) { return tmem_base + uint32_t(u * 128); };
  auto O_col = [&](int w) { return tmem_base + 256 + uint32_t(w * 128); };

  if (warp == 8) {
    // ===================== TMA producer (both CTAs; bytes are accounted on the leader's barriers) =====================
    if (elect_one()) {
      if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * Cfg::Q_BYTES);
      for (int a = 0; a < HD / 64; ++a)
        tma_load_4d_pair(q_smem + a * (128 * 64 * 2), &maps.q128, q_full, a * 64, h, q0, b);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int stg = j % NSTG;
      const uint32_t par = ((j / NSTG) & 1) ^ 1u;
      mbar_wait(k_empty(stg), par, 10);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(k_full(stg), 2 * Cfg::K_BYTES);
        for (int a = 0; a < HD / 64; ++a)
          tma_load_4d_pair(k_smem(stg) + a * (64 * 64 * 2), &maps.k64, k_full(stg), a * 64, h, j * 128 + int(rank) * 64, b);
      }
      __syncwarp();
      mbar_wait(v_empty(stg), par, 11);
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(v_full(stg), 2 * Cfg::V_BYTES);
        tma_load_4d_pair(v_smem(stg), &maps.v128, v_full(stg), int(rank) * 64, h, j * 128, b);
      }
      __syncwarp();
    }
  } else if (warp == 9 && rank == 0) {
    // ===================== MMA issuer (leader CTA, one elected lane) =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(256, 128, 0, 0);  // Q (K-major, own rows) x K (K-major halves)
    constexpr uint32_t idesc_o = make_idesc_bf16(256, HD, 0, 1);   // P_w (TMEM)            x V (MN-major halves)
    auto issue_S = [&](int u, int stg) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk)
        mma_ss2(S_col(u), sdesc_k(q_smem, (kk / 4) * (128 * 64 * 2) + (kk % 4) * 32),
                sdesc_k(k_smem(stg), (kk / 4) * (64 * 64 * 2) + (kk % 4) * 32), idesc_s, kk > 0);
    };
    auto issue_PV = [&](int u, int w, int stg, bool acc) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)   // the 64 keys of half w: P_w packed at S columns [64 w, 64 w + 32)
        mma_ts2(O_col(w), S_col(u) + 64 * w + 8 * kk, sdesc_mn(v_smem(stg), (4 * w + kk) * 2048, 128 * 64 * 2), idesc_o,
                (acc || kk > 0) ? 1u : 0u);
    };
    mbar_wait(q_full, 0, 20);
    mbar_wait(k_full(0), 0, 21);
    tc_fence_after();
    if (elect_one()) {
      issue_S(0, 0);
      tc_commit2(s_full(0));
      tc_commit2(k_empty(0));
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int stg = j % NSTG;
      const uint32_t par = (j / NSTG) & 1;
      const int u = j & 1;
      if (j + 1 < n_kv) {
        const int stg_n = (j + 1) % NSTG;
        mbar_wait(k_full(stg_n), ((j + 1) / NSTG) & 1, 23);
        tc_fence_after();
        if (elect_one()) {   // S_{j+1} into the other buffer (both halves of P_{j-1} were consumed by the PVs issued last iteration)
          issue_S(u ^ 1, stg_n);
          tc_commit2(s_full(u ^ 1));
          tc_commit2(k_empty(stg_n));
        }
        __syncwarp();
      }
      mbar_wait(v_full(stg), par, 22);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        mbar_wait(p_full(u, w), (j >> 1) & 1, 24 + w);   // both CTAs' P_{j, w} is in TMEM
        tc_fence_after();
        if (elect_one()) {
          issue_PV(u, w, stg, j > 0);
          tc_commit2(o_done(w));
          if (w == 1) tc_commit2(v_empty(stg));
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    // ===================== compute warpgroups: warpgroup w owns key half w of every tile, independently =====================
    const int w = warp >> 2;
    const int r = (warp & 3) * 32 + lane;    // row within this CTA's Q tile == TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const int s = q0 + r;
    const bool row_ok = s < p.Sq;
    const float sl2 = p.scale * 1.4426950408889634f;
    const uint32_t o_t = O_col(w) + lane_off;
    float m_used = -INFINITY;  // raw-score max currently baked into O_w and l
    float l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int u = j & 1;
      const uint32_t s_t = S_col(u) + lane_off + 64 * w;
      mbar_wait(s_full(u), (j >> 1) & 1, 30);
      tc_fence_after();
      const int kv_valid = p.Sk - j * 128 - 64 * w;  // valid columns of this half; < 64 only on the ragged last tile
      const bool ragged = kv_valid < 64;
#if STB_ATTN_DEBUG_SKIP >= 2
      l += 1.f;
      if (kv_valid < -1000000 && ragged) m_used = 0.f;
#else
      // ---- single TMEM read of this half (64 fp32 scores stay in registers): the forward is bound by the
      // 64 B/clk TMEM read path, a second pass over S would double the per-tile time (attn_fwd.cuh reads it twice)
      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(s_t, va);
      tmem_ld_32x32b_x32(s_t + 32, vb);
      tc_wait_ld();
      if (ragged) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) va[i] = 0xff800000u;       // -inf
          if (32 + i >= kv_valid) vb[i] = 0xff800000u;
        }
      }
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(va[i]), __uint_as_float(vb[i])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(va[i + 1]), __uint_as_float(vb[i + 1])));
        m2 = fmaxf(m2, fmaxf(__uint_as_float(va[i + 2]), __uint_as_float(vb[i + 2])));
        m3 = fmaxf(m3, fmaxf(__uint_as_float(va[i + 3]), __uint_as_float(vb[i + 3])));
      }
      const float m_tile = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      // ---- lazy rescale of this warpgroup's accumulator (warp-uniform decision; -inf - -inf = NaN compares false)
      const bool need = (m_tile - m_used) * sl2 > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_used, m_tile);
        const float alpha = (m_used == -INFINITY) ? 0.f : ex2f((m_used - m_new) * sl2);
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          mbar_wait(o_done(w), (j - 1) & 1, 31);  // PV_{j-1, w} has landed in O_w
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < HD; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(o_t + c, v);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x32(o_t + c, v);
          }
          tc_wait_st();
        }
      }
      // ---- P = exp2(s*sl2 - m*sl2) -> bf16, packed into the first 32 columns of this half
      const float mb = (m_used == -INFINITY) ? 0.f : m_used * sl2;   // a half with no valid key so far: exp2(-inf) = 0
      float l0 = 0.f, l1 = 0.f;
      {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = ex2f(fmaf(__uint_as_float(va[i]), sl2, -mb));
          const float x1 = ex2f(fmaf(__uint_as_float(va[i + 1]), sl2, -mb));
          l0 += x0;
          l1 += x1;
          pk[i / 2] = pack_bf16x2(x0, x1);
        }
        tmem_st_32x32b_x16(s_t, pk);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = ex2f(fmaf(__uint_as_float(vb[i]), sl2, -mb));
          const float x1 = ex2f(fmaf(__uint_as_float(vb[i + 1]), sl2, -mb));
          l0 += x0;
          l1 += x1;
          pk[i / 2] = pack_bf16x2(x0, x1);
        }
        tmem_st_32x32b_x16(s_t + 16, pk);
      }
      l += l0 + l1;
#endif
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(p_full(u, w), 0u);
    }

    // ---- epilogue: merge the two half-accumulators: O = (O_0 a_0 + O_1 a_1) / (l_0 a_0 + l_1 a_1), a_w = 2^((m_w - M) sl2)
    scratch_f[(w * 2 + 0) * 128 + r] = m_used;
    scratch_f[(w * 2 + 1) * 128 + r] = l;
    named_bar_sync(2, 256);
    const float m_o = scratch_f[((w ^ 1) * 2 + 0) * 128 + r], l_o = scratch_f[((w ^ 1) * 2 + 1) * 128 + r];
    const float M = fmaxf(m_used, m_o);
    const float a_me = (m_used == -INFINITY) ? 0.f : ex2f((m_used - M) * sl2);
    const float a_ot = (m_o == -INFINITY) ? 0.f : ex2f((m_o - M) * sl2);
    const float lsum = l * a_me + l_o * a_ot;
    const float inv_l = 1.f / lsum;
    const float a0 = (w == 0 ? a_me : a_ot) * inv_l, a1 = (w == 0 ? a_ot : a_me) * inv_l;
    mbar_wait(o_done(0), (n_kv - 1) & 1, 32);
    mbar_wait(o_done(1), (n_kv - 1) & 1, 33);
    tc_fence_after();
    __nv_bfloat16* orow = p.O + (long long)b * p.o_b + (long long)s * p.o_s + (long long)h * p.o_h + 64 * w;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {   // this warpgroup stores output columns [64 w, 64 w + 64)
      uint32_t v0[32], v1[32];
      tmem_ld_32x32b_x32(O_col(0) + lane_off + 64 * w + c, v0);
      tmem_ld_32x32b_x32(O_col(1) + lane_off + 64 * w + c, v1);
      tc_wait_ld();
      if (row_ok) {
        uint4* dp = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(v0[q * 8 + i]) * a0 + __uint_as_float(v1[q * 8 + i]) * a1;
          dp[q] = pack8(o);
        }
      }
    }
    if (w == 0 && row_ok && p.lse) p.lse[((long long)b * p.H + h) * p.Sq + s] = M * p.scale + logf(lsum);
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still read our shared memory / barriers / TMEM through the pair MMAs
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace stb
