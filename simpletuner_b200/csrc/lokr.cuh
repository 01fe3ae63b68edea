// simpletuner_b200 — LyCORIS LoKr support kernels (BASELINE configs[3], host side: simpletuner_b200/lycoris.py).
// reference call sites: helpers/training/trainer.py:3390-3505 (create_lycoris / apply_to), peft_init.py:34-38 (lokr_w1 / lokr_w2);
// algorithm: lycoris-lora (third-party, setup.py:319): delta W = kron(w1 [a, c], w2 [b, d]) * scale, y = linear(x, W + delta W).
// Both kernels are single streaming passes over an adapted weight (HBM-bound; 2-6 bytes per weight element):
//   lokr_rebuild_kernel : out[n, k] = bf16(W[n, k] + scale * w1[n / b, k / d] * w2[n % b, k % d]), written row-major into
//                         the fused projection layout AND transposed into the dgrad layout in the same pass (64 x 64 tile
//                         through shared memory) — replaces torch.kron + add + cast + cat + transpose (~30 B / element).
//   lokr_factor_grad_kernel : d w1[i, k'] = scale * sum_{j, l} dW[i b + j, k' d + l] * w2[j, l]
//                             d w2[j, l]  = scale * sum_{i, k'} dW[i b + j, k' d + l] * w1[i, k']
//                         each CTA owns 2048 consecutive (j, l) positions of w2, walks all (i, k') blocks of dW (coalesced
//                         16-byte reads along l), keeps its d w2 slice in registers (no atomics) and adds one partial per
//                         (i, k') to d w1.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace stb {

constexpr int LOKR_TILE = 64;

__global__ void __launch_bounds__(256)
lokr_rebuild_kernel(const __nv_bfloat16* __restrict__ W, long long w_rs, const __nv_bfloat16* __restrict__ w1,
                    const __nv_bfloat16* __restrict__ w2, float scale, __nv_bfloat16* __restrict__ out, long long o_rs,
                    __nv_bfloat16* __restrict__ out_t, long long t_rs, int N, int K, int b, int c, int d) {
  __shared__ __nv_bfloat16 tile[LOKR_TILE][LOKR_TILE + 8];
  const int n0 = blockIdx.y * LOKR_TILE, k0 = blockIdx.x * LOKR_TILE;
  // 256 threads: 8 column groups of 8 elements x 32 rows, two row passes
  const int cg = threadIdx.x & 7, r_in = threadIdx.x >> 3;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int r = r_in + pass * 32;
    const int n = n0 + r, k = k0 + cg * 8;
    if (n < N && k < K) {
      float wv[8], o[8];
      const bool full = (k + 8 <= K) && ((w_rs & 7) == 0) && ((o_rs & 7) == 0);
      if (full) {
        unpack8(*reinterpret_cast<const uint4*>(W + (long long)n * w_rs + k), wv);
      } else {
        for (int j = 0; j < 8; ++j) wv[j] = (k + j < K) ? __bfloat162float(W[(long long)n * w_rs + k + j]) : 0.f;
      }
      const int i = n / b, jj = n % b;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kk = min(k + j, K - 1);
        const float f1 = __bfloat162float(w1[i * c + kk / d]);
        const float f2 = __bfloat162float(w2[(long long)jj * d + kk % d]);
        o[j] = fmaf(scale * f1, f2, wv[j]);
      }
      if (full) {
        *reinterpret_cast<uint4*>(out + (long long)n * o_rs + k) = pack8(o);
      } else {
        for (int j = 0; j < 8; ++j)
          if (k + j < K) out[(long long)n * o_rs + k + j] = __float2bfloat16(o[j]);
      }
      if (out_t) {
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[r][cg * 8 + j] = __float2bfloat16(o[j]);
      }
    }
  }
  if (!out_t) return;
  __syncthreads();
  // transposed write: thread -> (k row, 8 consecutive n)
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int kr = r_in + pass * 32;
    const int k = k0 + kr, n = n0 + cg * 8;
    if (k < K && n < N) {
      if (n + 8 <= N && ((t_rs & 7) == 0)) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __bfloat162float(tile[cg * 8 + j][kr]);
        *reinterpret_cast<uint4*>(out_t + (long long)k * t_rs + n) = pack8(o);
      } else {
        for (int j = 0; j < 8 && n + j < N; ++j) out_t[(long long)k * t_rs + n + j] = tile[cg * 8 + j][kr];
      }
    }
  }
}

// dW [a*b, c*d] (row stride g_rs, d % 8 == 0); dw1 [a, c] / dw2 [b, d] fp32, zero-initialised by the caller
__global__ void __launch_bounds__(256)
lokr_factor_grad_kernel(const __nv_bfloat16* __restrict__ dW, long long g_rs, const __nv_bfloat16* __restrict__ w1,
                        const __nv_bfloat16* __restrict__ w2, float scale, float* __restrict__ dw1, float* __restrict__ dw2,
                        int a, int b, int c, int d) {
  extern __shared__ float part[];            // [a * c] partial sums of this CTA
  const int ac = a * c;
  for (int t = threadIdx.x; t < ac; t += blockDim.x) part[t] = 0.f;
  __syncthreads();
  const long long pos = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;   // flat (j, l) position, l fastest
  const bool live = pos < (long long)b * d;
  const int j = live ? int(pos / d) : 0, l = live ? int(pos % d) : 0;
  float w2v[8], acc2[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc2[q] = 0.f;
  if (live) unpack8(*reinterpret_cast<const uint4*>(w2 + (long long)j * d + l), w2v);
  const int lane = threadIdx.x & 31;
  for (int i = 0; i < a; ++i) {
    for (int kq = 0; kq < c; ++kq) {
      float p = 0.f;
      if (live) {
        float g[8];
        unpack8(*reinterpret_cast<const uint4*>(dW + (long long)(i * b + j) * g_rs + (long long)kq * d + l), g);
        const float f1 = __bfloat162float(w1[i * c + kq]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          acc2[q] = fmaf(g[q], f1, acc2[q]);
          p = fmaf(g[q], w2v[q], p);
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
      if (lane == 0) atomicAdd(&part[i * c + kq], p);
    }
  }
  if (live) {
    float4* o = reinterpret_cast<float4*>(dw2 + (long long)j * d + l);
    o[0] = make_float4(acc2[0] * scale, acc2[1] * scale, acc2[2] * scale, acc2[3] * scale);
    o[1] = make_float4(acc2[4] * scale, acc2[5] * scale, acc2[6] * scale, acc2[7] * scale);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ac; t += blockDim.x) atomicAdd(&dw1[t], part[t] * scale);
}

}  // namespace stb
