// simpletuner_b200 — full-rank weight gradient GEMM on tcgen05, sm_100a (round 2: full fine-tune, BASELINE config 3).
//
//   dW[n, k] = alpha * sum_{b,s} dY[b, s, n] * X[b, s, k]   (+ dW_old[n, k])        bf16 out, fp32 accumulation in TMEM
//
// This is the autograd of every nn.Linear weight the reference trains when `model_type=full`
// (reference trainer.py:7126 `accelerator.backward`; SD3 blocks sd3/transformer.py:145-241).  Both operands are contracted
// over their SLOW dimension (the token rows), so both are MN-major for the tensor core — the layout of wgrad.cuh (the
// rank-r LoRA version), generalised to a persistent, tiled kernel:
//   A = dY^T tile [128 n x 64 tokens]  (two   [64 tokens x 64 n] SWIZZLE_128B boxes)
//   B = X^T  tile [BN  k x 64 tokens]  (BN/64 [64 tokens x 64 k] boxes),  D tile [128 x BN] fp32 in TMEM,
// k-blocks run over (batch, 64-token chunk); rows past the end of a batch slab are zero-filled by TMA, so ragged
// sequences and strided [B, S, N] views (a row range of the joint hidden buffer) need no copies.
// Roles (256 threads): warp0 TMA producer, warp1 MMA issuer, warp2 TMEM alloc, warps 4-7 epilogue; TMEM holds two
// accumulator stages so the bf16 store of tile i overlaps the mainloop of tile i+1.
// Algorithmic work: 2 * M * N * K flops (M = B*S tokens) — the same as the forward GEMM of that layer.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace stb {

struct WgradFullParams {
  int S, B, N, K;
  float alpha;
  int accumulate;               // 1: out += (bf16 read-modify-write)
  __nv_bfloat16* out;           // [N, K]
  long long out_row_stride;
};

struct WgradFullMaps {
  CUtensorMap dy;  // 3-D (n, s, b) box (64, 64, 1) SWIZZLE_128B
  CUtensorMap x;   // 3-D (k, s, b) box (64, 64, 1) SWIZZLE_128B
};

template <int BN>
struct WgradFullCfg {
  static constexpr int A_BYTES = 2 * 8192;
  static constexpr int B_BYTES = (BN / 64) * 8192;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int ACC_STAGES = 512 / BN >= 2 ? 2 : 1;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
wgrad_full_kernel(const __grid_constant__ WgradFullMaps maps, const WgradFullParams p) {
  using Cfg = WgradFullCfg<BN>;
  constexpr int STAGES = Cfg::STAGES, ACC_STAGES = Cfg::ACC_STAGES;
  constexpr int A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto accf_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto acce_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (p.N + 127) / 128;
  const int tiles_k = (p.K + BN - 1) / BN;
  const int num_tiles = tiles_n * tiles_k;
  const int chunks = (p.S + 63) / 64;       // 64-token chunks per batch slab
  const int kblocks = chunks * p.B;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.dy);
    tma_prefetch_desc(&maps.x);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(accf_bar(s), 1);
      mbar_init(acce_bar(s), 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      // k-tile fastest: CTAs running side by side share the dY columns (A) of one n-tile through L2
      const int tn = tile / tiles_k, tk = tile - tn * tiles_k;
      const int n0 = tn * 128, k0 = tk * BN;
      for (int kb = 0; kb < kblocks; ++kb) {
        const int b = kb / chunks, s = (kb - b * chunks) * 64;
        mbar_wait(empty_bar(stage), phase ^ 1u, 80);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_3d(sa, &maps.dy, full_bar(stage), n0, s, b);
          tma_load_3d(sa + 8192, &maps.dy, full_bar(stage), n0 + 64, s, b);
#pragma unroll
          for (int a = 0; a < BN / 64; ++a) tma_load_3d(sa + A_BYTES + a * 8192, &maps.x, full_bar(stage), k0 + a * 64, s, b);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(acce_bar(acc), acc_phase ^ 1u, 81);
      tc_fence_after();
      const uint32_t d_base = tmem_base + acc * BN;
      uint32_t accumulate = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(full_bar(stage), phase, 82);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        // tokens past the end of the slab are zeros (TMA fill): whole 16-token MMAs beyond it can be skipped
        const int s = (kb % chunks) * 64;
        const int nk = (min(64, p.S - s) + 15) / 16;
        if (elect_one()) {
          for (int kk = 0; kk < nk; ++kk) {
            mma_ss(d_base, sdesc_mn(sa, kk * 2048, 8192), sdesc_mn(sa, A_BYTES + kk * 2048, 8192), idesc, accumulate);
            accumulate = 1;
          }
          tc_commit(empty_bar(stage));
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) tc_commit(accf_bar(acc));
      __syncwarp();
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tn = tile / tiles_k, tk = tile - tn * tiles_k;
      const int n = tn * 128 + ew * 32 + lane;
      const int k0 = tk * BN;
      mbar_wait(accf_bar(acc), acc_phase, 83);
      tc_fence_after();
      __nv_bfloat16* orow = p.out + (long long)n * p.out_row_stride;
      const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        if (k0 + c >= p.K) break;   // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + c, v);
        tc_wait_ld();
        if (n < p.N) {
          const int k = k0 + c;
          if (k + 32 <= p.K) {
            uint4* dp = reinterpret_cast<uint4*>(orow + k);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = p.alpha * __uint_as_float(v[q * 8 + j]);
              if (p.accumulate) {
                const uint4 o = dp[q];
                f[0] += bf16_lo(o.x); f[1] += bf16_hi(o.x); f[2] += bf16_lo(o.y); f[3] += bf16_hi(o.y);
                f[4] += bf16_lo(o.z); f[5] += bf16_hi(o.z); f[6] += bf16_lo(o.w); f[7] += bf16_hi(o.w);
              }
              uint4 u;
              u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
              u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
              dp[q] = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (k + j < p.K) {
                float f = p.alpha * __uint_as_float(v[j]);
                if (p.accumulate) f += __bfloat162float(orow[k + j]);
                orow[k + j] = __float2bfloat16(f);
              }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acce_bar(acc));
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-(batch, column) reductions over the token axis, fp32 atomics into zeroed outputs:
//   sum[b, d] = sum_s dy[b, s, d]            (gradient of an adaLN shift, bias gradients when summed over b on the host)
//   dot[b, d] = sum_s dy[b, s, d] * z[b, s, d]   (gradient of an adaLN scale with z = LayerNorm(x), of a gate with z = the
//                                                 gated linear output)
// Either output may be null.  HBM-bound: reads dy (and z) once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum2_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_b, long long dy_s, const __nv_bfloat16* __restrict__ z,
               long long z_b, long long z_s, float* __restrict__ sum, float* __restrict__ dot, int B, int S, int D,
               int rows_per_cta) {
  const int vecs = D >> 3;
  const int b = blockIdx.z;
  const int s0 = blockIdx.y * rows_per_cta;
  const int s1 = min(S, s0 + rows_per_cta);
  for (int vi = blockIdx.x * blockDim.x + threadIdx.x; vi < vecs; vi += gridDim.x * blockDim.x) {
    const int c = vi * 8;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int s = s0; s < s1; ++s) {
      float g[8];
      unpack8(*reinterpret_cast<const uint4*>(dy + b * dy_b + s * dy_s + c), g);
      if (dot) {
        float zz[8];
        unpack8(*reinterpret_cast<const uint4*>(z + b * z_b + s * z_s + c), zz);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += g[j] * zz[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += g[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (sum) atomicAdd(sum + (long long)b * D + c + j, a[j]);
      if (dot) atomicAdd(dot + (long long)b * D + c + j, d[j]);
    }
  }
}

}  // namespace stb
