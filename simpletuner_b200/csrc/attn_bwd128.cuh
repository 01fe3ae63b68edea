// simpletuner_b200 — attention backward with 128-wide streamed tiles (round 2), tcgen05 / sm_100a.
//
// Same contract and the same two-kernel, atomics-free split as attn_bwd.cuh (dK/dV per key tile, dQ per query tile), but
// the streamed operand moves in 128-row tiles and each tile's exp / dS math is split into two PHASES so that the single
// score buffer never idles the tensor pipe:
//
//   dkdv128 (TMEM lane = key):  X = S^T = K Q_i^T,  Y = dP^T = V dO_i^T   (SS, N = 128)
//        phase A: P^T = exp2(X*sl2 - lse)   -> bf16 over X   -> dV += P^T dO_i   (TS, B = dO_i MN-major)
//        phase B: dS^T = P^T o (Y - delta)  -> bf16 over Y   -> dK += dS^T Q_i   (TS, B = Q_i  MN-major)
//        tensor-pipe order:  S_0 dP_0 | dV_0 S_1 dK_0 dP_1 | dV_1 S_2 dK_1 dP_2 | ...
//        phase A of tile i+1 needs only S_{i+1} (issued right after dV_i) and runs under dK_i / dP_{i+1};
//        phase B of tile i runs under dV_i / S_{i+1}.
//   dq128 (TMEM lane = query):  S = Q K_j^T,  dP = dO V_j^T   (TS: Q / dO resident in TMEM as bf16 pairs, N = 128)
//        phase A: P = exp2(S*sl2 - lse) kept in registers; the S buffer is released as soon as it has been READ
//        phase B: dS = P o (dP - delta) -> bf16 over dP      -> dQ += dS K_j       (TS, B = K_j MN-major)
//        tensor-pipe order:  S_0 dP_0 | S_1 dQ_0 dP_1 | S_2 dQ_1 dP_2 | ...
//
// Why 128-wide: the MMA probe (tools/probe/mma_probe.cu, profiles/r02) runs the 128-wide instruction mixes of these two
// kernels at 2061 + 1668 cycles per 128 x 128 tile (99 % / 92 % of the tensor-pipe floor) against 4088 + 4217 for the
// 64-wide mixes of attn_bwd.cuh, whose N = 64 MMAs are bound by per-instruction overhead.  TMEM has no room to double-
// buffer 128-wide score tiles next to the accumulators (4 x 128 columns = 512), hence the phase split instead.
// Both compute warpgroups work on every tile (warpgroup w owns streamed columns [64 w, 64 w + 64) and packs its bf16
// result into the first 32 columns of its own fp32 range), thread = TMEM lane.
// warps 0-7 compute, warp 8 TMA producer, warp 9 MMA issuer, warp 10 TMEM allocator.
#pragma once
#include "attn_bwd.cuh"

namespace stb {

template <int HD>
struct AttnBwd128Cfg {
  static constexpr int TILE = 128 * HD * 2;                 // any 128-row operand tile
  static constexpr int KV_STAGES_DKDV = (HD == 128) ? 2 : 4;   // streamed (Q, dO) stages next to the resident K, V tiles
  static constexpr int KV_STAGES_DQ = (HD == 128) ? 3 : 6;     // streamed (K, V) stages (Q / dO live in TMEM)
  static constexpr int STAT_BYTES = 8 * 2 * 128 * 4;       // per compute warp: [2 parities][lse2 64 | delta 64]
  static constexpr int SMEM_DKDV = 2 * TILE + KV_STAGES_DKDV * 2 * TILE + STAT_BYTES + 1024 + 256;
  static constexpr int SMEM_DQ = KV_STAGES_DQ * 2 * TILE + 1024 + 256;
};

// TMEM column (relative to a 128-wide score buffer) of the packed bf16 operand chunk kk (16 of the 128 streamed rows):
// chunks 0-3 are written by compute warpgroup 0 at columns 0-31, chunks 4-7 by warpgroup 1 at columns 64-95.
#define PK128_COL(kk) (uint32_t((kk) < 4 ? 8 * (kk) : 64 + 8 * ((kk) - 4)))

__device__ __forceinline__ void tmem_ld_64(uint32_t taddr, uint32_t (&r)[64]) { tmem_ld_32x32b_x64(taddr, r); }
__device__ __forceinline__ void tmem_st_32(uint32_t taddr, const uint32_t (&r)[32]) { tmem_st_32x32b_x32(taddr, r); }

// ------------------------------------------------------------------------------------------------
// dK / dV, 128-query streamed tiles
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dkdv128_kernel(const __grid_constant__ AttnBwdMaps maps, const AttnBwdParams p) {
  using Cfg = AttnBwd128Cfg<HD>;
  constexpr int ATOMS = HD / 64;
  constexpr int TILE = Cfg::TILE, NSTG = Cfg::KV_STAGES_DKDV;
  constexpr int ATOM128 = 128 * 64 * 2;  // [128 rows x 64] SWIZZLE_128B box

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t k_smem = smem_base;
  const uint32_t v_smem = k_smem + TILE;
  auto q_smem = [&](int s) { return v_smem + TILE + uint32_t(s) * 2 * TILE; };
  auto do_smem = [&](int s) { return q_smem(s) + TILE; };
  const uint32_t stat_smem = smem_base + 2 * TILE + NSTG * 2 * TILE;
  const uint32_t bar_base = stat_smem + Cfg::STAT_BYTES;
  const uint32_t kv_full = bar_base;
  auto qdo_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto qdo_empty = [&](int s) { return bar_base + 8u * (1 + NSTG + s); };
  const uint32_t st_full = bar_base + 8u * (1 + 2 * NSTG);
  const uint32_t dpt_full = st_full + 8u;
  const uint32_t p_full = st_full + 16u;
  const uint32_t ds_full = st_full + 24u;
  const uint32_t acc_done = st_full + 32u;
  const uint32_t tmem_slot = st_full + 40u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* stat_ptr = reinterpret_cast<float*>(smem_raw + (stat_smem - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_q = (p.Sq + 127) / 128;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&maps.q128);
    tma_prefetch_desc(&maps.k128);
    tma_prefetch_desc(&maps.v128);
    tma_prefetch_desc(&maps.do128);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(qdo_full(s), 1);
      mbar_init(qdo_empty(s), 1);
    }
    mbar_init(st_full, 1);
    mbar_init(dpt_full, 1);
    mbar_init(p_full, 8);    // one arrive per compute warp
    mbar_init(ds_full, 8);
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 10) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t X = tmem_base;            // S^T  (P^T packed in place)
  const uint32_t Y = tmem_base + 128;      // dP^T (dS^T packed in place)
  const uint32_t DV = tmem_base + 256;
  const uint32_t DK = tmem_base + 256 + HD;

  if (warp >= 8) {
    if (warp == 8) {
      // ===================== TMA producer =====================
      if (elect_one()) {
        mbar_arrive_expect_tx(kv_full, 2 * TILE);
        for (int a = 0; a < ATOMS; ++a) {
          tma_load_4d(k_smem + a * ATOM128, &maps.k128, kv_full, a * 64, h, kv0, b);
          tma_load_4d(v_smem + a * ATOM128, &maps.v128, kv_full, a * 64, h, kv0, b);
        }
      }
      __syncwarp();
      int stg = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_q; ++i) {
        mbar_wait(qdo_empty(stg), ph ^ 1u, 60);
        if (elect_one()) {
          mbar_arrive_expect_tx(qdo_full(stg), 2 * TILE);
          for (int a = 0; a < ATOMS; ++a) {
            tma_load_4d(q_smem(stg) + a * ATOM128, &maps.q128, qdo_full(stg), a * 64, h, i * 128, b);
            tma_load_4d(do_smem(stg) + a * ATOM128, &maps.do128, qdo_full(stg), a * 64, h, i * 128, b);
          }
        }
        __syncwarp();
        if (++stg == NSTG) {
          stg = 0;
          ph ^= 1u;
        }
      }
    } else if (warp == 9) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_st = make_idesc_bf16(128, 128, 0, 0);  // [kv x q128] = K-major x K-major
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, 0, 1);  // [kv x d]    = TMEM x MN-major
      auto issue_S = [&](int stg) {
#pragma unroll 1
        for (int kk = 0; kk < HD / 16; ++kk)
          mma_ss(X, sdesc_k(k_smem, (kk / 4) * ATOM128 + (kk % 4) * 32), sdesc_k(q_smem(stg), (kk / 4) * ATOM128 + (kk % 4) * 32),
                 idesc_st, kk > 0);
        tc_commit(st_full);
      };
      auto issue_dP = [&](int stg) {
#pragma unroll 1
        for (int kk = 0; kk < HD / 16; ++kk)
          mma_ss(Y, sdesc_k(v_smem, (kk / 4) * ATOM128 + (kk % 4) * 32), sdesc_k(do_smem(stg), (kk / 4) * ATOM128 + (kk % 4) * 32),
                 idesc_st, kk > 0);
        tc_commit(dpt_full);
      };
      mbar_wait(kv_full, 0, 61);
      mbar_wait(qdo_full(0), 0, 62);
      tc_fence_after();
      if (elect_one()) {
        issue_S(0);
        issue_dP(0);
      }
      __syncwarp();
      int stg = 0;
      uint32_t ph_n = 0;   // phase of the stage of tile i+1
      for (int i = 0; i < n_q; ++i) {
        int stg_n = stg + 1;
        if (stg_n == NSTG) stg_n = 0;
        const bool more = (i + 1 < n_q);
        mbar_wait(p_full, i & 1, 63);     // P^T of tile i is in TMEM (and X has been read by every compute warp)
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int kk = 0; kk < 8; ++kk)   // contraction over the 128 query rows of the tile
            mma_ts(DV, X + PK128_COL(kk), sdesc_mn(do_smem(stg), kk * 2048, ATOM128), idesc_acc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        __syncwarp();
        if (more) {
          if (stg_n == 0) ph_n ^= 1u;
          mbar_wait(qdo_full(stg_n), ph_n, 64);
          tc_fence_after();
          if (elect_one()) issue_S(stg_n);     // executes after dV_i (issue order), overwrites X
          __syncwarp();
        }
        mbar_wait(ds_full, i & 1, 65);    // dS^T of tile i is in TMEM
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int kk = 0; kk < 8; ++kk)
            mma_ts(DK, Y + PK128_COL(kk), sdesc_mn(q_smem(stg), kk * 2048, ATOM128), idesc_acc, (i > 0 || kk > 0) ? 1u : 0u);
          tc_commit(qdo_empty(stg));
          if (more) issue_dP(stg_n);           // executes after dK_i, overwrites Y
        }
        __syncwarp();
        stg = stg_n;
      }
      if (elect_one()) tc_commit(acc_done);
      __syncwarp();
    }
  } else {
    // ===================== compute warpgroups =====================
    const int w = warp >> 2;
    const int r = (warp & 3) * 32 + lane;          // key row within the tile == TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const uint32_t col0 = 64u * w;                 // this warpgroup's query columns of every tile
    const float sl2 = p.scale * 1.4426950408889634f;
    const float2 sl2x2 = make_float2(sl2, sl2);
    const float l2e = 1.4426950408889634f;
    const long long stat_base = ((long long)b * p.H + h) * p.Sq;
    float* my_stat = stat_ptr + warp * 256;        // [2 parities][lse2 64 | delta 64], private to this warp
    // lane l fetches the statistics of queries (col0 + l) and (col0 + 32 + l) of a tile
    auto fetch = [&](int tile, float (&s)[4]) {
      const int q0 = tile * 128 + int(col0) + lane;
      const bool ok0 = q0 < p.Sq, ok1 = q0 + 32 < p.Sq;
      // stored NEGATED: they are the addends of the packed FFMA2 / FADD2 below
      s[0] = ok0 ? -p.lse[stat_base + q0] * l2e : -INFINITY;       // invalid query: P = exp2(-inf) = 0
      s[1] = ok1 ? -p.lse[stat_base + q0 + 32] * l2e : -INFINITY;
      s[2] = ok0 ? -p.delta[stat_base + q0] : 0.f;
      s[3] = ok1 ? -p.delta[stat_base + q0 + 32] : 0.f;
    };
    auto stash = [&](int buf, const float (&s)[4]) {
      float* d = my_stat + buf * 128;
      d[lane] = s[0];
      d[32 + lane] = s[1];
      d[64 + lane] = s[2];
      d[96 + lane] = s[3];
    };
    {
      float s0[4];
      fetch(0, s0);
      stash(0, s0);
      __syncwarp();
    }
    for (int i = 0; i < n_q; ++i) {
      const int buf = i & 1;
      const float* lse_s = my_stat + buf * 128;
      const float* del_s = lse_s + 64;
      float nxt[4];
      const bool pre = (i + 1 < n_q);
      if (pre) fetch(i + 1, nxt);                  // consumed at the end of the iteration (latency hidden)
      // ---- phase A: P^T = exp2(S^T * sl2 - lse2)
      uint32_t pv[64];                             // fp32 P, reused by phase B
      mbar_wait(st_full, i & 1, 66);
      tc_fence_after();
      tmem_ld_64(X + lane_off + col0, pv);
      tc_wait_ld();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {   // 32 queries at a time: 16 packed columns per tcgen05.st
        uint32_t pk[16];
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          const int j = hf * 32 + jj;
          const float4 l4 = *reinterpret_cast<const float4*>(lse_s + j);   // smem broadcast (-lse2)
          const float2 a01 = __ffma2_rn(make_float2(__uint_as_float(pv[j + 0]), __uint_as_float(pv[j + 1])), sl2x2, make_float2(l4.x, l4.y));
          const float2 a23 = __ffma2_rn(make_float2(__uint_as_float(pv[j + 2]), __uint_as_float(pv[j + 3])), sl2x2, make_float2(l4.z, l4.w));
          const float x0 = ex2f(a01.x), x1 = ex2f(a01.y), x2 = ex2f(a23.x), x3 = ex2f(a23.y);
          pv[j + 0] = __float_as_uint(x0);
          pv[j + 1] = __float_as_uint(x1);
          pv[j + 2] = __float_as_uint(x2);
          pv[j + 3] = __float_as_uint(x3);
          pk[jj / 2] = pack_bf16x2(x0, x1);
          pk[jj / 2 + 1] = pack_bf16x2(x2, x3);
        }
        tmem_st_32x32b_x16(X + lane_off + col0 + 16 * hf, pk);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- phase B: dS^T = P^T o (dP^T - delta)   (the softmax scale of dS is applied once to the dK accumulator)
      mbar_wait(dpt_full, i & 1, 67);
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {   // 32 queries at a time keeps the live set at P (64) + dP (32) + packed (16) registers
        uint32_t dv[32];
        tmem_ld_32x32b_x32(Y + lane_off + col0 + 32 * hf, dv);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int jj = 0; jj < 32; jj += 4) {
          const int j = hf * 32 + jj;
          const float4 d4 = *reinterpret_cast<const float4*>(del_s + j);   // (-delta)
          const float2 t01 = __fadd2_rn(make_float2(__uint_as_float(dv[jj + 0]), __uint_as_float(dv[jj + 1])), make_float2(d4.x, d4.y));
          const float2 t23 = __fadd2_rn(make_float2(__uint_as_float(dv[jj + 2]), __uint_as_float(dv[jj + 3])), make_float2(d4.z, d4.w));
          const float2 e01 = __fmul2_rn(make_float2(__uint_as_float(pv[j + 0]), __uint_as_float(pv[j + 1])), t01);
          const float2 e23 = __fmul2_rn(make_float2(__uint_as_float(pv[j + 2]), __uint_as_float(pv[j + 3])), t23);
          pk[jj / 2] = pack_bf16x2(e01.x, e01.y);
          pk[jj / 2 + 1] = pack_bf16x2(e23.x, e23.y);
        }
        tmem_st_32x32b_x16(Y + lane_off + col0 + 16 * hf, pk);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      if (pre) {
        stash(buf ^ 1, nxt);   // parity buf^1 was last read in iteration i-1 by this same warp
        __syncwarp();
      }
    }
    // ---- epilogue: warpgroup 0 stores dV, warpgroup 1 stores dK
    mbar_wait(acc_done, 0, 68);
    tc_fence_after();
    const int kv = kv0 + r;
    const bool row_ok = kv < p.Sk;
    if (w == 0)
      store_acc_row<HD>(DV + lane_off, p.dv + (long long)b * p.dv_b + (long long)kv * p.dv_s + (long long)h * p.dv_h, row_ok);
    else
      store_acc_row<HD>(DK + lane_off, p.dk + (long long)b * p.dk_b + (long long)kv * p.dk_s + (long long)h * p.dk_h, row_ok, p.scale);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// dQ, 128-key streamed tiles
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dq128_kernel(const __grid_constant__ AttnBwdMaps maps, const AttnBwdParams p) {
  using Cfg = AttnBwd128Cfg<HD>;
  constexpr int ATOMS = HD / 64;
  constexpr int TILE = Cfg::TILE, NSTG = Cfg::KV_STAGES_DQ;
  constexpr int ATOM128 = 128 * 64 * 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto k_smem = [&](int s) { return smem_base + uint32_t(s) * 2 * TILE; };
  auto v_smem = [&](int s) { return k_smem(s) + TILE; };
  const uint32_t bar_base = smem_base + NSTG * 2 * TILE;
  const uint32_t qdo_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + NSTG + s); };
  const uint32_t s_full = bar_base + 8u * (1 + 2 * NSTG);
  const uint32_t dp_full = s_full + 8u;
  const uint32_t s_read = s_full + 16u;
  const uint32_t ds_full = s_full + 24u;
  const uint32_t dq_done = s_full + 32u;
  const uint32_t tmem_slot = s_full + 40u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.Sk + 127) / 128;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&maps.k128);
    tma_prefetch_desc(&maps.v128);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(qdo_full, 8);
    for (int s = 0; s < NSTG; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(s_read, 8);
    mbar_init(ds_full, 8);
    mbar_init(dq_done, 1);
    fence_mbar_init();
  }
  if (warp == 10) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t Sb = tmem_base;              // S
  const uint32_t Yb = tmem_base + 128;        // dP (dS packed in place)
  const uint32_t DQ = tmem_base + 256;
  const uint32_t QT = tmem_base + 256 + HD;            // Q  [128 x HD] bf16 : HD / 2 columns
  const uint32_t DOT = tmem_base + 256 + HD + HD / 2;  // dO [128 x HD] bf16

  if (warp >= 8) {
    if (warp == 8) {
      int stg = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(kv_empty(stg), ph ^ 1u, 70);
        if (elect_one()) {
          mbar_arrive_expect_tx(kv_full(stg), 2 * TILE);
          for (int a = 0; a < ATOMS; ++a) {
            tma_load_4d(k_smem(stg) + a * ATOM128, &maps.k128, kv_full(stg), a * 64, h, j * 128, b);
            tma_load_4d(v_smem(stg) + a * ATOM128, &maps.v128, kv_full(stg), a * 64, h, j * 128, b);
          }
        }
        __syncwarp();
        if (++stg == NSTG) {
          stg = 0;
          ph ^= 1u;
        }
      }
    } else if (warp == 9) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 0, 1);
      auto issue_S = [&](int stg) {
#pragma unroll 1
        for (int kk = 0; kk < HD / 16; ++kk)
          mma_ts(Sb, QT + 8 * kk, sdesc_k(k_smem(stg), (kk / 4) * ATOM128 + (kk % 4) * 32), idesc_s, kk > 0);
        tc_commit(s_full);
      };
      auto issue_dP = [&](int stg) {
#pragma unroll 1
        for (int kk = 0; kk < HD / 16; ++kk)
          mma_ts(Yb, DOT + 8 * kk, sdesc_k(v_smem(stg), (kk / 4) * ATOM128 + (kk % 4) * 32), idesc_s, kk > 0);
        tc_commit(dp_full);
      };
      mbar_wait(qdo_full, 0, 71);
      mbar_wait(kv_full(0), 0, 72);
      tc_fence_after();
      if (elect_one()) {
        issue_S(0);
        issue_dP(0);
      }
      __syncwarp();
      int stg = 0;
      uint32_t ph_n = 0;
      for (int j = 0; j < n_kv; ++j) {
        int stg_n = stg + 1;
        if (stg_n == NSTG) stg_n = 0;
        const bool more = (j + 1 < n_kv);
        if (more) {
          if (stg_n == 0) ph_n ^= 1u;
          mbar_wait(s_read, j & 1, 73);        // every compute warp holds S_j in registers
          mbar_wait(kv_full(stg_n), ph_n, 74);
          tc_fence_after();
          if (elect_one()) issue_S(stg_n);
          __syncwarp();
        }
        mbar_wait(ds_full, j & 1, 75);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll 1
          for (int kk = 0; kk < 8; ++kk)   // contraction over the 128 keys of the tile
            mma_ts(DQ, Yb + PK128_COL(kk), sdesc_mn(k_smem(stg), kk * 2048, ATOM128), idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
          tc_commit(kv_empty(stg));
          if (more) issue_dP(stg_n);           // executes after dQ_j, overwrites the dP / dS buffer
        }
        __syncwarp();
        stg = stg_n;
      }
      if (elect_one()) tc_commit(dq_done);
      __syncwarp();
    }
  } else {
    const int w = warp >> 2;
    const int r = (warp & 3) * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const uint32_t col0 = 64u * w;         // this warpgroup's key columns of every tile
    const float sl2 = p.scale * 1.4426950408889634f;
    const int qrow = q0 + r;
    const bool row_ok = qrow < p.Sq;
    const long long stat_idx = ((long long)b * p.H + h) * p.Sq + qrow;
    const float lse2 = row_ok ? p.lse[stat_idx] * 1.4426950408889634f : INFINITY;
    const float delta = row_ok ? p.delta[stat_idx] : 0.f;
    const float2 sl2x2 = make_float2(sl2, sl2), nlse2x2 = make_float2(-lse2, -lse2), ndeltax2 = make_float2(-delta, -delta);
    {  // warpgroup 0 stages this thread's Q row, warpgroup 1 its dO row (global -> registers -> TMEM)
      const __nv_bfloat16* src = (w == 0)
          ? p.q + (long long)b * p.q_b + (long long)qrow * p.q_s + (long long)h * p.q_h
          : p.d_o + (long long)b * p.do_b + (long long)qrow * p.do_s + (long long)h * p.do_h;
      const uint32_t dst = (w == 0 ? QT : DOT) + lane_off;
#pragma unroll
      for (int c = 0; c < HD / 2; c += 32) {           // 32 columns = 64 bf16 = 128 bytes per chunk
        uint32_t v[32];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const uint4 u = row_ok ? __ldg(reinterpret_cast<const uint4*>(src + 2 * c) + q4) : make_uint4(0u, 0u, 0u, 0u);
          v[4 * q4 + 0] = u.x, v[4 * q4 + 1] = u.y, v[4 * q4 + 2] = u.z, v[4 * q4 + 3] = u.w;
        }
        tmem_st_32x32b_x32(dst + c, v);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qdo_full);
    }
    for (int j = 0; j < n_kv; ++j) {
      const int kv_valid = p.Sk - j * 128 - int(col0);  // < 64 only on the ragged last tile
      // ---- phase A: read S (then release the buffer), P = exp2(S * sl2 - lse2)
      uint32_t pv[64];
      mbar_wait(s_full, j & 1, 76);
      tc_fence_after();
      tmem_ld_64(Sb + lane_off + col0, pv);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_read);
      if (kv_valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kv_valid) pv[i] = 0xff800000u;  // -inf -> P = 0
      }
#pragma unroll
      for (int i = 0; i < 64; i += 2) {   // packed FFMA2 halves the FMA-pipe instruction count of the phase
        const float2 a = __ffma2_rn(make_float2(__uint_as_float(pv[i]), __uint_as_float(pv[i + 1])), sl2x2, nlse2x2);
        pv[i] = __float_as_uint(ex2f(a.x));
        pv[i + 1] = __float_as_uint(ex2f(a.y));
      }
      // ---- phase B: dS = P o (dP - delta)   (softmax scale applied in the epilogue)
      mbar_wait(dp_full, j & 1, 77);
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t dv[32];
        tmem_ld_32x32b_x32(Yb + lane_off + col0 + 32 * hf, dv);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 t = __fadd2_rn(make_float2(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])), ndeltax2);
          const float2 e = __fmul2_rn(make_float2(__uint_as_float(pv[hf * 32 + i]), __uint_as_float(pv[hf * 32 + i + 1])), t);
          pk[i / 2] = pack_bf16x2(e.x, e.y);
        }
        tmem_st_32x32b_x16(Yb + lane_off + col0 + 16 * hf, pk);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    mbar_wait(dq_done, 0, 78);
    tc_fence_after();
    // both warpgroups cover all 128 lanes: split the HD columns between them
    __nv_bfloat16* dqrow = p.dq + (long long)b * p.dq_b + (long long)qrow * p.dq_s + (long long)h * p.dq_h;
#pragma unroll 1
    for (int c = 0; c < HD / 2; c += 32) {
      const int col = w * (HD / 2) + c;
      uint32_t v[32];
      tmem_ld_32x32b_x32(DQ + lane_off + col, v);
      tc_wait_ld();
      if (row_ok) {
        uint4* dp = reinterpret_cast<uint4*>(dqrow + col);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]) * p.scale, __uint_as_float(v[q * 8 + 1]) * p.scale);
          u.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]) * p.scale, __uint_as_float(v[q * 8 + 3]) * p.scale);
          u.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]) * p.scale, __uint_as_float(v[q * 8 + 5]) * p.scale);
          u.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]) * p.scale, __uint_as_float(v[q * 8 + 7]) * p.scale);
          dp[q] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace stb
