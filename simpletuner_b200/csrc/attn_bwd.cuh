// simpletuner_b200 — attention backward on tcgen05, sm_100a (autograd of the SDPA call at reference
// flux/transformer.py:200-207).
//
//   P = exp(S*scale - LSE),  dV = P^T dO,  dP = dO V^T,  dS = P o (dP - Delta) * scale,
//   dQ = dS K,  dK = dS^T Q,   Delta_i = sum_d dO_id O_id
//
// Round-1 structure: two kernels, no atomics, deterministic.
//   attn_bwd_dkdv_kernel : one CTA per (b, h, 128-key tile), loops over 128-row query tiles.  Works in
//       the transposed domain (TMEM lane = key): S^T = K Q^T, dP^T = V dO^T; P^T and dS^T are written
//       back to TMEM as bf16 and fed as the A operand of  dV += P^T dO,  dK += dS^T Q  (dO / Q tiles
//       re-read from the same smem bytes through MN-major descriptors).
//   attn_bwd_dq_kernel   : one CTA per (b, h, 128-row query tile), loops over key tiles (lane = query
//       row): S = Q K^T (double-buffered), dP = dO V^T, dS -> TMEM bf16, dQ += dS K.
// Both: warps 0-7 = two compute warpgroups that split the 128 tile columns (WG g owns columns
// [64g, 64g+64) and packs its bf16 results into the low half of its own column range), warp 8 = TMA
// producer, warp 9 = MMA issuer, warp 10 = TMEM allocator.
#pragma once
#include "../../include/stb200.h"
#include "common.cuh"

namespace stb {

struct AttnBwdParams {
  int B, H, Sq, Sk;
  float scale;
  const float* lse;    // [B, H, Sq]
  const float* delta;  // [B, H, Sq]
  __nv_bfloat16 *dq, *dk, *dv;
  long long dq_b, dq_s, dq_h, dk_b, dk_s, dk_h, dv_b, dv_s, dv_h;
};

struct AttnBwdMaps {
  CUtensorMap q, k, v, d_o;  // 4-D (d, h, s, b), box (64, 1, 128, 1), SWIZZLE_128B
};

template <int HD>
struct AttnBwdCfg {
  static constexpr int TILE_BYTES = 128 * HD * 2;
  static constexpr int SMEM_BYTES = 6 * TILE_BYTES + 4 * 512 + 1024 + 256;
};

// ------------------------------------------------------------------------------------------------
// Delta = rowsum(dO * O), one warp per (b, s, h)
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256)
attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, long long o_b, long long o_s, long long o_h,
                      const __nv_bfloat16* __restrict__ d_o, long long do_b, long long do_s, long long do_h,
                      float* __restrict__ delta, int B, int H, int S) {
  constexpr int EPL = HD / 32;
  const long long wid = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wid >= (long long)B * S * H) return;
  const int lane = threadIdx.x & 31;
  const int h = int(wid % H);
  const long long r = wid / H;
  const int s = int(r % S);
  const int b = int(r / S);
  const __nv_bfloat16* po = o + b * o_b + s * o_s + h * o_h + lane * EPL;
  const __nv_bfloat16* pg = d_o + b * do_b + s * do_s + h * do_h + lane * EPL;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc += __bfloat162float(po[i]) * __bfloat162float(pg[i]);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) delta[((long long)b * H + h) * S + s] = acc;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// dK / dV
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ AttnBwdMaps maps, const AttnBwdParams p) {
  constexpr int ATOMS = HD / 64;
  constexpr int TILE = 128 * HD * 2;
  constexpr int ATOM_BYTES = 128 * 64 * 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t k_smem = smem_base;
  const uint32_t v_smem = k_smem + TILE;
  auto q_smem = [&](int s) { return v_smem + TILE + uint32_t(s) * 2 * TILE; };
  auto do_smem = [&](int s) { return q_smem(s) + TILE; };
  const uint32_t stat_smem = smem_base + 6 * TILE;  // [2 stages][lse 128 | delta 128] fp32
  const uint32_t bar_base = stat_smem + 4 * 512;
  const uint32_t kv_full = bar_base;
  auto qdo_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto qdo_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  const uint32_t st_full = bar_base + 8u * 5;
  const uint32_t dpt_full = bar_base + 8u * 6;
  const uint32_t p_full = bar_base + 8u * 7;
  const uint32_t ds_full = bar_base + 8u * 8;
  const uint32_t acc_done = bar_base + 8u * 9;
  const uint32_t tmem_slot = bar_base + 8u * 10;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* stat_ptr = reinterpret_cast<float*>(smem_raw + (stat_smem - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_q = (p.Sq + 127) / 128;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    tma_prefetch_desc(&maps.d_o);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(qdo_full(s), 1);
      mbar_init(qdo_empty(s), 1);
    }
    mbar_init(st_full, 1);
    mbar_init(dpt_full, 1);
    mbar_init(p_full, 256);
    mbar_init(ds_full, 256);
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
  const uint32_t X = tmem_base;         // S^T  (P^T packed in place)
  const uint32_t Y = tmem_base + 128;   // dP^T (dS^T packed in place)
  const uint32_t DV = tmem_base + 256;
  const uint32_t DK = tmem_base + 384;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * TILE);
      for (int a = 0; a < ATOMS; ++a) {
        tma_load_4d(k_smem + a * ATOM_BYTES, &maps.k, kv_full, a * 64, h, kv0, b);
        tma_load_4d(v_smem + a * ATOM_BYTES, &maps.v, kv_full, a * 64, h, kv0, b);
      }
      for (int i = 0; i < n_q; ++i) {
        const int stg = i & 1;
        mbar_wait(qdo_empty(stg), ((i >> 1) & 1) ^ 1u, 40);
        mbar_arrive_expect_tx(qdo_full(stg), 2 * TILE);
        for (int a = 0; a < ATOMS; ++a) {
          tma_load_4d(q_smem(stg) + a * ATOM_BYTES, &maps.q, qdo_full(stg), a * 64, h, i * 128, b);
          tma_load_4d(do_smem(stg) + a * ATOM_BYTES, &maps.d_o, qdo_full(stg), a * 64, h, i * 128, b);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc_st = make_idesc_bf16(128, 128, 0, 0);  // [kv x q] = K-major x K-major
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, 0, 1);  // [kv x d] = TMEM x MN-major
      mbar_wait(kv_full, 0, 41);
      for (int i = 0; i < n_q; ++i) {
        const int stg = i & 1;
        mbar_wait(qdo_full(stg), (i >> 1) & 1, 42);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t ad = sdesc_kmajor(k_smem + (kk / 4) * ATOM_BYTES, (kk % 4) * 16);
          const uint64_t bd = sdesc_kmajor(q_smem(stg) + (kk / 4) * ATOM_BYTES, (kk % 4) * 16);
          mma_ss(X, ad, bd, idesc_st, kk > 0);
        }
        tc_commit(st_full);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t ad = sdesc_kmajor(v_smem + (kk / 4) * ATOM_BYTES, (kk % 4) * 16);
          const uint64_t bd = sdesc_kmajor(do_smem(stg) + (kk / 4) * ATOM_BYTES, (kk % 4) * 16);
          mma_ss(Y, ad, bd, idesc_st, kk > 0);
        }
        tc_commit(dpt_full);
        mbar_wait(p_full, i & 1, 43);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // contraction over the 128 query rows of the tile
          const uint64_t bd = sdesc_mnmajor(do_smem(stg), kk * 16, ATOM_BYTES);
          mma_ts(DV, X + 64 * (kk / 4) + 8 * (kk % 4), bd, idesc_acc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        mbar_wait(ds_full, i & 1, 44);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t bd = sdesc_mnmajor(q_smem(stg), kk * 16, ATOM_BYTES);
          mma_ts(DK, Y + 64 * (kk / 4) + 8 * (kk % 4), bd, idesc_acc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(qdo_empty(stg));
      }
      tc_commit(acc_done);
    }
  } else if (warp < 8) {
    // ===================== compute warpgroups =====================
    const int g = warp >> 2;                       // column half
    const int r = (warp & 3) * 32 + lane;          // key row within tile == TMEM lane
    const int ct = threadIdx.x;                    // 0..255
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const float l2e = 1.4426950408889634f;
    const long long stat_base = ((long long)b * p.H + h) * p.Sq;
    // stage the per-query statistics of tile 0
    {
      const int qi = ct & 127;
      const bool ok = qi < p.Sq;
      float v = (ct < 128) ? (ok ? p.lse[stat_base + qi] : INFINITY) : (ok ? p.delta[stat_base + qi] : 0.f);
      stat_ptr[ct] = v;  // [0,128) lse, [128,256) delta of stage 0
    }
    named_bar_sync(1, 256);
    for (int i = 0; i < n_q; ++i) {
      const int stg = i & 1;
      const float* lse_s = stat_ptr + stg * 256;
      const float* del_s = lse_s + 128;
      // prefetch next tile's statistics into a register
      float nxt = 0.f;
      if (i + 1 < n_q) {
        const int qi = (i + 1) * 128 + (ct & 127);
        const bool ok = qi < p.Sq;
        nxt = (ct < 128) ? (ok ? p.lse[stat_base + qi] : INFINITY) : (ok ? p.delta[stat_base + qi] : 0.f);
      }
      mbar_wait(st_full, i & 1, 45);
      tc_fence_after();
      float pv[64];  // P^T for this thread's 64 columns (kept for dS)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(X + lane_off + 64 * g + 32 * c, v);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int q = 64 * g + 32 * c + j;
          float x0 = ex2f(fmaf(__uint_as_float(v[j]), sl2, -lse_s[q] * l2e));
          float x1 = ex2f(fmaf(__uint_as_float(v[j + 1]), sl2, -lse_s[q + 1] * l2e));
          pv[32 * c + j] = x0;
          pv[32 * c + j + 1] = x1;
          pk[j / 2] = pack_bf16x2(x0, x1);
        }
        tmem_st_32x32b_x16(X + lane_off + 64 * g + 16 * c, pk);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_wait(dpt_full, i & 1, 46);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(Y + lane_off + 64 * g + 32 * c, v);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int q = 64 * g + 32 * c + j;
          float d0 = pv[32 * c + j] * (__uint_as_float(v[j]) - del_s[q]) * p.scale;
          float d1 = pv[32 * c + j + 1] * (__uint_as_float(v[j + 1]) - del_s[q + 1]) * p.scale;
          pk[j / 2] = pack_bf16x2(d0, d1);
        }
        tmem_st_32x32b_x16(Y + lane_off + 64 * g + 16 * c, pk);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(ds_full);
      // publish next tile's statistics (other stage; its previous readers finished a tile ago)
      if (i + 1 < n_q) stat_ptr[(stg ^ 1) * 256 + ct] = nxt;
      named_bar_sync(1, 256);
    }
    // ---- epilogue: dV, dK -> bf16 -> global
    mbar_wait(acc_done, 0, 47);
    tc_fence_after();
    const int kv = kv0 + r;
    const bool row_ok = kv < p.Sk;
    __nv_bfloat16* dvrow = p.dv + (long long)b * p.dv_b + (long long)kv * p.dv_s + (long long)h * p.dv_h;
    __nv_bfloat16* dkrow = p.dk + (long long)b * p.dk_b + (long long)kv * p.dk_s + (long long)h * p.dk_h;
#pragma unroll 1
    for (int c = 0; c < HD / 2; c += 32) {
      const int col = g * (HD / 2) + c;
      uint32_t v[32];
      tmem_ld_32x32b_x32(DV + lane_off + col, v);
      tc_wait_ld();
      if (row_ok) {
        uint4* dp = reinterpret_cast<uint4*>(dvrow + col);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
          dp[q] = u;
        }
      }
      tmem_ld_32x32b_x32(DK + lane_off + col, v);
      tc_wait_ld();
      if (row_ok) {
        uint4* dp = reinterpret_cast<uint4*>(dkrow + col);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
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

// ------------------------------------------------------------------------------------------------
// dQ
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(384, 1)
attn_bwd_dq_kernel(const __grid_constant__ AttnBwdMaps maps, const AttnBwdParams p) {
  constexpr int ATOMS = HD / 64;
  constexpr int TILE = 128 * HD * 2;
  constexpr int ATOM_BYTES = 128 * 64 * 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  const uint32_t do_smem = q_smem + TILE;
  auto k_smem = [&](int s) { return do_smem + TILE + uint32_t(s) * 2 * TILE; };
  auto v_smem = [&](int s) { return k_smem(s) + TILE; };
  const uint32_t bar_base = smem_base + 6 * TILE + 4 * 512;
  const uint32_t qdo_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (3 + s); };
  auto s_full = [&](int s) { return bar_base + 8u * (5 + s); };
  const uint32_t dp_full = bar_base + 8u * 7;
  const uint32_t ds_full = bar_base + 8u * 8;
  const uint32_t dq_done = bar_base + 8u * 9;
  const uint32_t tmem_slot = bar_base + 8u * 10;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_kv = (p.Sk + 127) / 128;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.v);
    tma_prefetch_desc(&maps.d_o);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(qdo_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
      mbar_init(s_full(s), 1);
    }
    mbar_init(dp_full, 1);
    mbar_init(ds_full, 256);
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
  auto Sbuf = [&](int s) { return tmem_base + uint32_t(s) * 128; };
  const uint32_t Y = tmem_base + 256;   // dP (dS packed in place)
  const uint32_t DQ = tmem_base + 384;

  if (warp == 8) {
    if (lane == 0) {
      mbar_arrive_expect_tx(qdo_full, 2 * TILE);
      for (int a = 0; a < ATOMS; ++a) {
        tma_load_4d(q_smem + a * ATOM_BYTES, &maps.q, qdo_full, a * 64, h, q0, b);
        tma_load_4d(do_smem + a * ATOM_BYTES, &maps.d_o, qdo_full, a * 64, h, q0, b);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int stg = j & 1;
        mbar_wait(kv_empty(stg), ((j >> 1) & 1) ^ 1u, 50);
        mbar_arrive_expect_tx(kv_full(stg), 2 * TILE);
        for (int a = 0; a < ATOMS; ++a) {
          tma_load_4d(k_smem(stg) + a * ATOM_BYTES, &maps.k, kv_full(stg), a * 64, h, j * 128, b);
          tma_load_4d(v_smem(stg) + a * ATOM_BYTES, &maps.v, kv_full(stg), a * 64, h, j * 128, b);
        }
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 0, 1);
      auto issue_S = [&](int buf, int stg) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t ad = sdesc_kmajor(q_smem + (kk / 4) * ATOM_BYTES, (kk % 4) * 16);
          const uint64_t bd = sdesc_kmajor(k_smem(stg) + (kk / 4) * ATOM_BYTES, (kk % 4) * 16);
          mma_ss(Sbuf(buf), ad, bd, idesc_s, kk > 0);
        }
      };
      auto issue_dP = [&](int stg) {
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t ad = sdesc_kmajor(do_smem + (kk / 4) * ATOM_BYTES, (kk % 4) * 16);
          const uint64_t bd = sdesc_kmajor(v_smem(stg) + (kk / 4) * ATOM_BYTES, (kk % 4) * 16);
          mma_ss(Y, ad, bd, idesc_s, kk > 0);
        }
      };
      mbar_wait(qdo_full, 0, 51);
      mbar_wait(kv_full(0), 0, 52);
      tc_fence_after();
      issue_S(0, 0);
      tc_commit(s_full(0));
      issue_dP(0);
      tc_commit(dp_full);
      for (int j = 0; j < n_kv; ++j) {
        const int stg = j & 1;
        const bool more = j + 1 < n_kv;
        if (more) {
          mbar_wait(kv_full(stg ^ 1), ((j + 1) >> 1) & 1, 53);
          tc_fence_after();
          issue_S(stg ^ 1, stg ^ 1);
          tc_commit(s_full(stg ^ 1));
        }
        mbar_wait(ds_full, j & 1, 54);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // contraction over the 128 keys of the tile
          const uint64_t bd = sdesc_mnmajor(k_smem(stg), kk * 16, ATOM_BYTES);
          mma_ts(DQ, Y + 64 * (kk / 4) + 8 * (kk % 4), bd, idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit(kv_empty(stg));
        if (more) {
          issue_dP(stg ^ 1);
          tc_commit(dp_full);
        }
      }
      tc_commit(dq_done);
    }
  } else if (warp < 8) {
    const int g = warp >> 2;
    const int r = (warp & 3) * 32 + lane;  // query row within tile == TMEM lane
    const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
    const float sl2 = p.scale * 1.4426950408889634f;
    const int qrow = q0 + r;
    const bool row_ok = qrow < p.Sq;
    const long long stat_idx = ((long long)b * p.H + h) * p.Sq + qrow;
    const float lse2 = row_ok ? p.lse[stat_idx] * 1.4426950408889634f : INFINITY;
    const float delta = row_ok ? p.delta[stat_idx] : 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int buf = j & 1;
      const int kv_valid = p.Sk - j * 128;
      mbar_wait(s_full(buf), (j >> 1) & 1, 55);
      mbar_wait(dp_full, j & 1, 56);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32b_x32(Sbuf(buf) + lane_off + 64 * g + 32 * c, sv);
        tmem_ld_32x32b_x32(Y + lane_off + 64 * g + 32 * c, dv);
        tc_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int col = 64 * g + 32 * c + i;
          float p0 = ex2f(fmaf(__uint_as_float(sv[i]), sl2, -lse2));
          float p1 = ex2f(fmaf(__uint_as_float(sv[i + 1]), sl2, -lse2));
          if (col >= kv_valid) p0 = 0.f;
          if (col + 1 >= kv_valid) p1 = 0.f;
          float d0 = p0 * (__uint_as_float(dv[i]) - delta) * p.scale;
          float d1 = p1 * (__uint_as_float(dv[i + 1]) - delta) * p.scale;
          pk[i / 2] = pack_bf16x2(d0, d1);
        }
        tmem_st_32x32b_x16(Y + lane_off + 64 * g + 16 * c, pk);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(ds_full);
    }
    mbar_wait(dq_done, 0, 57);
    tc_fence_after();
    __nv_bfloat16* dqrow = p.dq + (long long)b * p.dq_b + (long long)qrow * p.dq_s + (long long)h * p.dq_h;
#pragma unroll 1
    for (int c = 0; c < HD / 2; c += 32) {
      const int col = g * (HD / 2) + c;
      uint32_t v[32];
      tmem_ld_32x32b_x32(DQ + lane_off + col, v);
      tc_wait_ld();
      if (row_ok) {
        uint4* dp = reinterpret_cast<uint4*>(dqrow + col);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
          u.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
          u.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
          u.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
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
