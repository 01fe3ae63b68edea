// simpletuner_b200 — attention forward, CTA-pair (cta_group::2) version for head_dim 128, sm_100a.
//
// Same contract as attn_fwd.cuh (O = softmax(Q K^T * scale) V and the natural-log LSE), different mapping, chosen
// from the measured MMA rates (tools/probe/mma_probe.cu, profiles/r01/mma_probe.log): SS-form MMAs are limited by
// the ~75 B/clk/SM operand fetch from shared memory (128x128x16: 107 cycles against a 64-cycle floor), TS-form
// N = 128 MMAs run at 85-100 % of the floor, and in pair mode the B operand is split between the two SMs.
//
// A 2-CTA cluster (one TPC) owns 256 query rows of one (batch, head); CTA r owns rows [128 r, 128 r + 128):
//   * Q lives in TMEM (bf16 pairs, staged once by the compute warps) and is the TS-form A operand of S = Q K^T;
//   * every 128-key K tile is split by KEY ROWS (CTA r stages keys [64 r, +64)), every V tile by HEAD-DIM COLUMNS
//     (CTA r stages columns [64 r, +64)): exactly the halves of the B operands the pair MMA reads from each SM,
//     so each SM stages 32 KB per key tile (as the single-CTA kernel did for two Q tiles) and fetches 2 KB per MMA;
//   * the leader CTA issues M = 256 MMAs for both SMs; commits are multicast to both CTAs;
//   * S is double-buffered in TMEM across key tiles (S_{j+1} is computed while the 8 compute warps run the softmax
//     of tile j); both warpgroups work on every tile, each on 64 of its 128 columns, exchanging the row maxima
//     through shared memory, and pack their bf16 P into the first 32 columns of their own half.
// TMEM (per CTA): S0 [0,128)  S1 [128,256)  O [256,384)  Q [384,448).
// warps 0-7 compute, warp 8 TMA producer, warp 9 MMA issuer (leader only), warp 10 TMEM allocator.
#pragma once
#include "attn_fwd.cuh"

#ifndef STB_ATTN_DEBUG_SKIP
#define STB_ATTN_DEBUG_SKIP 0
#endif

namespace stb {

struct AttnFwdPairMaps {
  CUtensorMap k64;   // 4-D (d, h, s, b), box (64, 1, 64, 1):  half of a key tile's rows, one 64-wide d atom
  CUtensorMap v128;  // 4-D (d, h, s, b), box (64, 1, 128, 1): all keys of a tile, one 64-wide d atom (this CTA's half)
};

struct AttnFwdPairParams {
  AttnFwdParams base;
  const __nv_bfloat16* q;   // raw view: the compute warps stage their Q rows into TMEM themselves
  long long q_b, q_s, q_h;
};

struct AttnFwdPairCfg {
  static constexpr int HD = 128;
  static constexpr int K_BYTES = 64 * HD * 2;    // this CTA's half of a key tile (64 keys x 128)
  static constexpr int V_BYTES = 128 * 64 * 2;   // this CTA's half of a value tile (128 keys x 64 columns)
  static constexpr int STAGES = 4;
  static constexpr int SCRATCH = 2 * 2 * 128 * 4 + 2 * 128 * 4;   // row-max exchange (double-buffered) + row-sum exchange
  static constexpr int SMEM_BYTES = STAGES * (K_BYTES + V_BYTES) + SCRATCH + 1024 + 256;
};

__global__ void __launch_bounds__(384, 1)
attn_fwd_pair_kernel(const __grid_constant__ AttnFwdPairMaps maps, const AttnFwdPairParams pp) {
  using Cfg = AttnFwdPairCfg;
  constexpr int HD = Cfg::HD;
  constexpr int NSTG = Cfg::STAGES;
  const AttnFwdParams& p = pp.base;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto k_smem = [&](int s) { return smem_base + uint32_t(s) * Cfg::K_BYTES; };
  auto v_smem = [&](int s) { return smem_base + NSTG * Cfg::K_BYTES + uint32_t(s) * Cfg::V_BYTES; };
  const uint32_t scratch = smem_base + NSTG * (Cfg::K_BYTES + Cfg::V_BYTES);
  const uint32_t bar_base = scratch + Cfg::SCRATCH;
  auto k_full = [&](int s) { return bar_base + 8u * s; };                 // leader's copy is the live one
  auto v_full = [&](int s) { return bar_base + 8u * (NSTG + s); };        // leader
  auto k_empty = [&](int s) { return bar_base + 8u * (2 * NSTG + s); };   // both CTAs (multicast commit)
  auto v_empty = [&](int s) { return bar_base + 8u * (3 * NSTG + s); };   // both
  auto s_full = [&](int u) { return bar_base + 8u * (4 * NSTG + u); };    // both
  auto p_full = [&](int u) { return bar_base + 8u * (4 * NSTG + 2 + u); };  // leader: 8 warps x 2 CTAs
  const uint32_t o_done = bar_base + 8u * (4 * NSTG + 4);                 // both
  const uint32_t q_ready = bar_base + 8u * (4 * NSTG + 5);                // leader: 8 warps x 2 CTAs
  const uint32_t tmem_slot = bar_base + 8u * (4 * NSTG + 6);
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
      mbar_init(p_full(u), 16);
    }
    mbar_init(o_done, 1);
    mbar_init(q_ready, 16);
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
  const uint32_t O_col = tmem_base + 256;
  const uint32_t QT = tmem_base + 384;
  // TMEM column of the packed bf16 P chunk kk (16 keys): chunks 0-3 belong to warpgroup 0 (S columns 0-63, packed
  // at 0-31), chunks 4-7 to warpgroup 1 (S columns 64-127, packed at 64-95)
  auto P_col = [&](int kk) { return uint32_t(kk < 4 ? 8 * kk : 64 + 8 * (kk - 4)); };

  if (warp == 8) {
    // ===================== TMA producer (both CTAs; bytes are accounted on the leader's barriers) =====================
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
    constexpr uint32_t idesc_s = make_idesc_bf16(256, 128, 0, 0);  // Q (TMEM)  x K (K-major halves)
    constexpr uint32_t idesc_o = make_idesc_bf16(256, HD, 0, 1);   // P (TMEM)  x V (MN-major halves)
    auto issue_S = [&](int u, int stg) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk)
        mma_ts2(S_col(u), QT + 8 * kk, sdesc_k(k_smem(stg), (kk / 4) * (64 * 64 * 2) + (kk % 4) * 32), idesc_s, kk > 0);
    };
    auto issue_PV = [&](int u, int stg, bool acc) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)   // 128 keys / 16
        mma_ts2(O_col, S_col(u) + P_col(kk), sdesc_mn(v_smem(stg), kk * 2048, 128 * 64 * 2), idesc_o, (acc || kk > 0) ? 1u : 0u);
    };
    mbar_wait(q_ready, 0, 20);
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
      if (j + 1 < n_kv) {
        const int stg_n = (j + 1) % NSTG;
        mbar_wait(k_full(stg_n), ((j + 1) / NSTG) & 1, 23);
        tc_fence_after();
        if (elect_one()) {   // S_{j+1} into the other buffer (its P_{j-1} was consumed by the PV issued last iteration)
          issue_S((j + 1) & 1, stg_n);
          tc_commit2(s_full((j + 1) & 1));
          tc_commit2(k_empty(stg_n));
        }
        __syncwarp();
      }
      mbar_wait(v_full(stg), par, 22);
      mbar_wait(p_full(j & 1), (j >> 1) & 1, 24);   // both CTAs' P_j is in TMEM
      tc_fence_after();
      if (elect_one()) {
        issue_PV(j & 1, stg, j > 0);
        tc_commit2(o_done);
        tc_commit2(v_empty(stg));
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ===================== compute warpgroups: both work on every key tile, 64 columns each =====================
    const int w = warp >> 2;
    const int r = (warp & 3) * 32 + lane;    // row within this CTA's Q tile == TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const int s = q0 + r;
    const bool row_ok = s < p.Sq;
    const float sl2 = p.scale * 1.4426950408889634f;
    {  // stage this row's Q half (head-dim columns [64 w, 64 w + 64)) into TMEM as packed bf16 pairs
      const __nv_bfloat16* src = pp.q + (long long)b * pp.q_b + (long long)s * pp.q_s + (long long)h * pp.q_h + 64 * w;
      uint32_t v[32];
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const uint4 u = row_ok ? __ldg(reinterpret_cast<const uint4*>(src) + q4) : make_uint4(0u, 0u, 0u, 0u);
        v[4 * q4 + 0] = u.x, v[4 * q4 + 1] = u.y, v[4 * q4 + 2] = u.z, v[4 * q4 + 3] = u.w;
      }
      tmem_st_32x32b_x32(QT + lane_off + 32 * w, v);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(q_ready, 0u);
    }
    float* mx = scratch_f;                 // [2 (j parity)][2 (warpgroup)][128]
    float* lx = scratch_f + 2 * 2 * 128;   // [2 (warpgroup)][128]
    const uint32_t o_t = O_col + lane_off + 64 * w;   // this warpgroup rescales / stores O columns [64 w, +64)
    float m_used = -INFINITY;  // raw-score max currently baked into O and l
    float l = 0.f;             // this thread's partial row sum (its 64 columns of every tile)

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
      // ---- pass 1: max over this half, then exchange with the other warpgroup
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(s_t + c, v);
        tc_wait_ld();
        if (ragged) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i >= kv_valid) v[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[i]));
          m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
      }
      float m_tile = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      mx[(u * 2 + w) * 128 + r] = m_tile;
      named_bar_sync(2, 256);
      m_tile = fmaxf(m_tile, mx[(u * 2 + (w ^ 1)) * 128 + r]);
      // ---- lazy rescale (both warpgroups take the same decision for the same rows)
      const bool need = (m_tile - m_used) * sl2 > 8.0f;  // true on the first tile (m_used = -inf)
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_used, m_tile);
        const float alpha = ex2f((m_used - m_new) * sl2);  // 0 on the first tile
        l *= alpha;
        m_used = m_new;
        if (j > 0) {
          mbar_wait(o_done, (j - 1) & 1, 31);  // PV_{j-1} has landed in O
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
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
      // ---- pass 2: P = exp2(s*sl2 - m*sl2) -> bf16, packed into the first 32 columns of this half
      const float mb = m_used * sl2;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(s_t + c, v);
        tc_wait_ld();
        if (ragged) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c + i >= kv_valid) v[i] = 0xff800000u;  // exp2(-inf) = 0
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = ex2f(fmaf(__uint_as_float(v[i]), sl2, -mb));
          const float x1 = ex2f(fmaf(__uint_as_float(v[i + 1]), sl2, -mb));
          l0 += x0;
          l1 += x1;
          pk[i / 2] = pack_bf16x2(x0, x1);
        }
        tmem_st_32x32b_x16(s_t + c / 2, pk);
      }
      l += l0 + l1;
#endif
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(p_full(u), 0u);
    }

    // ---- epilogue: O / l -> bf16 -> global (this warpgroup's 64 columns); LSE
    lx[w * 128 + r] = l;
    named_bar_sync(2, 256);
    l += lx[(w ^ 1) * 128 + r];
    mbar_wait(o_done, (n_kv - 1) & 1, 32);
    tc_fence_after();
    const float inv_l = 1.f / l;
    __nv_bfloat16* orow = p.O + (long long)b * p.o_b + (long long)s * p.o_s + (long long)h * p.o_h + 64 * w;
#pragma unroll 1
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(o_t + c, v);
      tc_wait_ld();
      if (row_ok) {
        uint4* dp = reinterpret_cast<uint4*>(orow + c);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u4;
          u4.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * inv_l, __uint_as_float(v[q * 8 + 1]) * inv_l);
          u4.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * inv_l, __uint_as_float(v[q * 8 + 3]) * inv_l);
          u4.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * inv_l, __uint_as_float(v[q * 8 + 5]) * inv_l);
          u4.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * inv_l, __uint_as_float(v[q * 8 + 7]) * inv_l);
          dp[q] = u4;
        }
      }
    }
    if (w == 0 && row_ok && p.lse) p.lse[((long long)b * p.H + h) * p.Sq + s] = m_used * p.scale + logf(l);
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still read our shared memory / barriers / TMEM through the pair MMAs
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace stb
