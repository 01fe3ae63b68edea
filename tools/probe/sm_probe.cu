// SM-side throughput probe (sm_100a) for the attention compute warps: tcgen05.ld / tcgen05.st rates, MUFU ex2 rate and
// candidate softmax inner loops (plain, packed f32x2, polynomial exp2 offload).  One CTA per SM, 4 or 8 compute warps.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I simpletuner_b200/csrc -o tools/probe/sm_probe tools/probe/sm_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "common.cuh"

using namespace stb;

// mode 0: LDTM x32 repeated; 1: LDTM x128; 2: STTM x32; 3: ex2 only; 4..: softmax loops on a 128 x 128 tile per 4 warps
__global__ void __launch_bounds__(256, 1) probe(int mode, int nwarps, int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t lane_off = uint32_t((warp & 3) * 32) << 16;
  const uint32_t col0 = (warp >> 2) * 128;   // second warpgroup works on its own columns
  float acc = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp < nwarps) {
    // zero the columns we read (defined values)
    {
      uint32_t z[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(0.001f * float(i + threadIdx.x));
      for (int c = 0; c < 128; c += 32) tmem_st_32x32b_x32(tm + lane_off + col0 + c, z);
      tc_wait_st();
    }
    __syncwarp();
    t0 = clock64();
    if (mode == 0) {
      for (int it = 0; it < iters; ++it) {
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 128; c += 32) {
          tmem_ld_32x32b_x32(tm + lane_off + col0 + c, v);
          tc_wait_ld();
          acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
        }
      }
    } else if (mode == 1) {
      for (int it = 0; it < iters; ++it) {
        uint32_t v[128];
        tmem_ld_32x32b_x128(tm + lane_off + col0, v);
        tc_wait_ld();
        acc += __uint_as_float(v[0]) + __uint_as_float(v[127]);
      }
    } else if (mode == 2) {
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = i + threadIdx.x;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 128; c += 32) tmem_st_32x32b_x32(tm + lane_off + col0 + c, v);
        tc_wait_st();
      }
    } else if (mode == 3) {
      float x[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = -0.01f * float(i + (threadIdx.x & 7));
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
          for (int i = 0; i < 16; ++i) x[i] = ex2f(x[i]) - 1.0f;   // 128 ex2 per thread per iteration (+ 128 FADD)
        }
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) acc += x[i];
    } else {
      // softmax tile loops: thread = row, 128 columns (mode 4, 5, 6) — the single-read structure
      const float sl2 = 0.1275f, mb = 3.0f;
      float l0 = 0.f, l1 = 0.f;
      for (int it = 0; it < iters; ++it) {
        uint32_t v[128];
        tmem_ld_32x32b_x128(tm + lane_off + col0, v);
        tc_wait_ld();
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(v[i]));
          m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
        const float mt = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2 + mb * 0.f;
        uint32_t pkA[32], pkB[32];
        if (mode == 4) {
#pragma unroll
          for (int i = 0; i < 128; i += 2) {
            const float x0 = ex2f(fmaf(__uint_as_float(v[i]), sl2, -mt));
            const float x1 = ex2f(fmaf(__uint_as_float(v[i + 1]), sl2, -mt));
            l0 += x0; l1 += x1;
            if (i < 64) pkA[i / 2] = pack_bf16x2(x0, x1); else pkB[i / 2 - 32] = pack_bf16x2(x0, x1);
          }
        } else {
          // mode 5: every 4th pair through a degree-3 polynomial on the FMA pipe (25 %); mode 6: every 2nd pair (50 %)
          const int every = (mode == 5) ? 4 : 2;
#pragma unroll
          for (int i = 0; i < 128; i += 2) {
            float a0 = fmaf(__uint_as_float(v[i]), sl2, -mt), a1 = fmaf(__uint_as_float(v[i + 1]), sl2, -mt);
            float x0, x1;
            if (((i / 2) % every) == 0) {
              // 2^a = 2^floor(a) * p(frac):  Cody-Waite with the magic-number floor, Horner degree 3
              a0 = fmaxf(a0, -126.f); a1 = fmaxf(a1, -126.f);
              const float f0 = a0 + 12582912.f, f1 = a1 + 12582912.f;       // round to nearest integer in the mantissa
              const float r0 = a0 - (f0 - 12582912.f), r1 = a1 - (f1 - 12582912.f);   // in [-0.5, 0.5]
              float p0 = fmaf(r0, 0.0555041f, 0.2402265f), p1 = fmaf(r1, 0.0555041f, 0.2402265f);
              p0 = fmaf(p0, r0, 0.6931472f); p1 = fmaf(p1, r1, 0.6931472f);
              p0 = fmaf(p0, r0, 1.0f); p1 = fmaf(p1, r1, 1.0f);
              x0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(f0) << 23));
              x1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(f1) << 23));
            } else {
              x0 = ex2f(a0); x1 = ex2f(a1);
            }
            l0 += x0; l1 += x1;
            if (i < 64) pkA[i / 2] = pack_bf16x2(x0, x1); else pkB[i / 2 - 32] = pack_bf16x2(x0, x1);
          }
        }
        tmem_st_32x32b_x32(tm + lane_off + col0, pkA);
        tmem_st_32x32b_x32(tm + lane_off + col0 + 32, pkB);
        tc_wait_st();
        // restore fp32-looking data so the next iteration computes on sane values
        {
          uint32_t z[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(0.001f * float(i + it));
          tmem_st_32x32b_x32(tm + lane_off + col0, z);
          tmem_st_32x32b_x32(tm + lane_off + col0 + 32, z);
          tc_wait_st();
        }
      }
      acc = l0 + l1;
    }
    t1 = clock64();
  }
  if (acc == 123.456f) sink[threadIdx.x] = acc;
  if (warp < nwarps && (threadIdx.x & 31) == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  long long* out;
  float* sink;
  cudaMallocManaged(&out, sizeof(long long) * nsm * 8);
  cudaMalloc(&sink, 4096);
  const char* names[] = {"LDTM 32x32b.x32 x4 (+wait each)", "LDTM 32x32b.x128", "STTM 32x32b.x32 x4", "MUFU ex2 (128 / thread / iter)",
                         "softmax 128 cols: MUFU only", "softmax 128 cols: 25% poly", "softmax 128 cols: 50% poly"};
  for (int mode = 0; mode <= 6; ++mode) {
    for (int nw : {4, 8}) {
      const int iters = 256;
      for (int rep = 0; rep < 2; ++rep) {
        probe<<<nsm, 256>>>(mode, nw, iters, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      }
      double mx = 0;
      for (int i = 0; i < nsm; ++i)
        for (int w = 0; w < nw; ++w) mx += double(out[i * 8 + w]);
      mx /= double(nsm) * nw * iters;
      // per iteration each warp moves 32 lanes x 128 cols x 4 B = 16 KB (modes 0-2) / handles a 32 x 128 slab (modes 4-6)
      printf("%-36s warps %d : %8.1f cycles / (128-col slab per warp)", names[mode], nw, mx);
      if (mode <= 2) printf("   -> %6.1f B/clk/SM", 16384.0 * nw / mx);
      if (mode == 3) printf("   -> %6.2f ex2/clk/SM", 128.0 * 32 * nw / mx);
      printf("\n");
    }
  }
  return 0;
}
