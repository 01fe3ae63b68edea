// tcgen05.mma throughput probe (sm_100a): one CTA per SM, one thread issues a long dependent-free stream of MMAs of one
// shape / operand form on zeroed operands and reports cycles per MMA.  Used to decide tile shapes.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I simpletuner_b200/csrc -o tools/probe/mma_probe tools/probe/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include "common.cuh"

using namespace stb;

struct Mode {
  const char* name;
  int N;        // MMA N
  int ts;       // A from TMEM
  int b_mn;     // B MN-major
  int nacc;     // number of distinct accumulators rotated (1 = back-to-back accumulate on the same D)
  int a_span;   // distinct A k-steps cycled (smem footprint realism)
};

__global__ void __launch_bounds__(128, 1) probe(int N, int ts, int b_mn, int nacc, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;                 // 128 x 64 bf16 (16 KB) x 2 atoms
  const uint32_t b_smem = base + 32768;         // up to 256 x 64 bf16 (32 KB) x 2 atoms
  const uint32_t bar = base + 32768 + 65536;
  const uint32_t slot = bar + 16;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (threadIdx.x < 32) {   // warp-uniform loop, instructions issued by one elected lane (the product kernels' pattern)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(b_mn) << 16) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tm + uint32_t((it % nacc) * N) % 256u;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t bd = b_mn ? sdesc_mn(b_smem, (kk % 4) * 2048, 64 * 64 * 2) : sdesc_k(b_smem, (kk / 4) * (N * 128) + (kk % 4) * 32);
          if (ts) mma_ts(d, tm + 256 + 8 * kk, bd, idesc, 1u);
          else mma_ss(d, sdesc_k(a_smem, (kk / 4) * 16384 + (kk % 4) * 32), bd, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0, 99);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// One "tile" of MMAs as the attention kernels issue them (operands zeroed, accumulators rotated like the kernels do).
//   mix 1: dkdv, 64-wide streamed tile  : 8+8 SS N=64 (scores), 4+4 TS MN-major-B N=128 (dV, dK)
//   mix 2: dq,   64-wide, Q/dO in TMEM  : 8+8 TS N=64,          4   TS MN-major-B N=128 (dQ)
//   mix 3: dq,   64-wide, Q/dO in smem  : 8+8 SS N=64,          4   TS MN-major-B N=128
//   mix 4: dkdv, 128-wide streamed tile : 8+8 SS N=128,         8+8 TS MN-major-B N=128
//   mix 5: fwd,  128 x 128 tile         : 8   SS N=128 (S),     8   TS MN-major-B N=128 (PV)
//   mix 6: dq,   128-wide, Q/dO in TMEM : 8+8 TS N=128,         8   TS MN-major-B N=128
__global__ void __launch_bounds__(128, 1) probe_mix(int mix, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, a2_smem = base + 32768, b_smem = base + 65536, b2_smem = base + 65536 + 32768;
  const uint32_t bar = base + 131072;
  const uint32_t slot = bar + 16;
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (threadIdx.x < 32) {
    constexpr uint32_t i64 = make_idesc_bf16(128, 64, 0, 0), i128 = make_idesc_bf16(128, 128, 0, 0), i128mn = make_idesc_bf16(128, 128, 0, 1);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t w = it & 1;
      if (elect_one()) {
        const bool wide = (mix == 4 || mix == 5 || mix == 6 || mix == 7);
        const uint32_t X = (mix == 8) ? tm : tm + (wide ? 0u : w * 64u), Y = (mix == 8) ? tm + 64 : tm + 128 + (wide ? 0u : w * 64u);
        const int nscore = (mix == 5 || mix == 7) ? 1 : 2;
        for (int g = 0; g < nscore; ++g) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t d = g ? Y : X;
            if (mix == 2 || mix == 8) mma_ts(d, tm + (mix == 8 ? 128 : 384) + g * 64 + 8 * kk, sdesc_k(g ? b2_smem : b_smem, (kk / 4) * 8192 + (kk % 4) * 32), i64, kk > 0);
            else if (mix == 6 || mix == 7) mma_ts(d, tm + 384 + g * 64 + 8 * kk, sdesc_k(g ? b2_smem : b_smem, (kk / 4) * 16384 + (kk % 4) * 32), i128, kk > 0);
            else if (wide) mma_ss(d, sdesc_k(g ? a2_smem : a_smem, (kk / 4) * 16384 + (kk % 4) * 32), sdesc_k(g ? b2_smem : b_smem, (kk / 4) * 16384 + (kk % 4) * 32), i128, kk > 0);
            else mma_ss(d, sdesc_k(g ? a2_smem : a_smem, (kk / 4) * 16384 + (kk % 4) * 32), sdesc_k(g ? b2_smem : b_smem, (kk / 4) * 8192 + (kk % 4) * 32), i64, kk > 0);
          }
        }
        const int nacc = (mix == 1 || mix == 4 || mix == 8) ? 2 : 1;
        const int ksteps = wide ? 8 : 4;
        for (int g = 0; g < nacc; ++g) {
#pragma unroll 8
          for (int kk = 0; kk < ksteps; ++kk)
            mma_ts(tm + 256 + g * 128, (g ? Y : X) + 8 * kk, sdesc_mn(g ? b_smem : b2_smem, kk * 2048, wide ? 16384 : 8192), i128mn, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0, 98);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// CTA-pair (cta_group::2) mixes, cluster (2,1,1): the leader issues M = 256 MMAs; operands zeroed.
//   pmix 1: fwd pair tile: 8 TS2 N=128 (Q in TMEM x K half, K-major) + 8 TS2 N=128 (P x V half, MN-major)
//   pmix 2: 8 SS2 N=128 (A smem 128 rows, B half) + 8 TS2 MN-major
//   pmix 3: 16 TS2 N=128 K-major only
//   pmix 4: GEMM k-block: 4 SS2 N=256
__global__ void __launch_bounds__(128, 1) probe_pair(int pmix, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base, b_smem = base + 32768, b2_smem = base + 65536;
  const uint32_t bar = base + 98304;
  const uint32_t slot = bar + 16;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc2(slot, 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tm = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x < 32 && rank == 0) {
    constexpr uint32_t i128 = make_idesc_bf16(256, 128, 0, 0), i128mn = make_idesc_bf16(256, 128, 0, 1), i256 = make_idesc_bf16(256, 256, 0, 0);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t u = it & 1;
      if (elect_one()) {
        if (pmix == 4) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) mma_ss2(tm + u * 256, sdesc_k(a_smem, kk * 32), sdesc_k(b_smem, kk * 32), i256, 1u);
        } else {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t bd = sdesc_k(b_smem, (kk / 4) * 8192 + (kk % 4) * 32);
            if (pmix == 2) mma_ss2(tm + u * 128, sdesc_k(a_smem, (kk / 4) * 16384 + (kk % 4) * 32), bd, i128, kk > 0);
            else mma_ts2(tm + u * 128, tm + 384 + 8 * kk, bd, i128, kk > 0);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (pmix == 3) mma_ts2(tm + u * 128, tm + 384 + 8 * kk, sdesc_k(b2_smem, (kk / 4) * 8192 + (kk % 4) * 32), i128, 1u);
            else mma_ts2(tm + 256, tm + (u ^ 1) * 128 + (kk < 4 ? 8 * kk : 64 + 8 * (kk - 4)), sdesc_mn(b2_smem, kk * 2048, 16384), i128mn, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit2(bar);
    __syncwarp();
    mbar_wait(bar, 0, 97);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x >> 1] = t1 - t0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc2(tm, 512); }
}

int main() {
  int nsm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 32768 + 65536 + 1024 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMallocManaged(&out, sizeof(long long) * nsm);
  const Mode modes[] = {
      {"SS K-major  N=256", 256, 0, 0, 1, 8}, {"SS K-major  N=128", 128, 0, 0, 2, 8}, {"SS K-major  N=64 ", 64, 0, 0, 4, 8},
      {"TS K-major  N=256", 256, 1, 0, 1, 8}, {"TS K-major  N=128", 128, 1, 0, 2, 8}, {"TS K-major  N=64 ", 64, 1, 0, 4, 8},
      {"SS B MN-maj N=128", 128, 0, 1, 2, 8}, {"TS B MN-maj N=128", 128, 1, 1, 2, 8}, {"TS B MN-maj N=64 ", 64, 1, 1, 4, 8},
      {"SS K-major  N=64 same D", 64, 0, 0, 1, 8}, {"TS B MN-maj N=128 same D", 128, 1, 1, 1, 8},
  };
  const int iters = 512;
  for (int grid : {nsm}) {
    for (const Mode& m : modes) {
      probe<<<grid, 128, smem>>>(m.N, m.ts, m.b_mn, m.nacc, iters, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", m.name, cudaGetErrorString(e)); return 1; }
      probe<<<grid, 128, smem>>>(m.N, m.ts, m.b_mn, m.nacc, iters, out);
      cudaDeviceSynchronize();
      double avg = 0;
      for (int i = 0; i < grid; ++i) avg += double(out[i]);
      avg /= grid;
      const double per = avg / (iters * 8);
      const double ideal = 128.0 * m.N / 256.0;   // guide: max(M,128) * N / 256 cycles per K=16 dispatch
      printf("grid %3d  %-26s  %7.1f cyc/MMA  (floor %5.1f, eff %5.1f%%)\n", grid, m.name, per, ideal, 100.0 * ideal / per);
    }
  }
  const int smem2 = 131072 + 1024 + 64;
  cudaFuncSetAttribute(probe_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
  const char* names[] = {"", "dkdv 64-wide (SS scores)", "dq 64-wide (TS scores)", "dq 64-wide (SS scores)", "dkdv 128-wide (SS scores)",
                         "fwd 128x128", "dq 128-wide (TS scores)", "fwd 128x128 (TS scores)", "dkdv 64-wide (TS scores)"};
  const double floors[] = {0, 1024, 768, 768, 2048, 1024, 1536, 1024, 1024};
  for (int mix = 1; mix <= 8; ++mix) {
    for (int rep = 0; rep < 2; ++rep) {
      probe_mix<<<nsm, 128, smem2>>>(mix, 512, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mix %d: %s\n", mix, cudaGetErrorString(e)); return 1; }
    }
    double avg = 0;
    for (int i = 0; i < nsm; ++i) avg += double(out[i]);
    avg /= nsm * 512.0;
    printf("grid %3d  mix %d %-28s %8.1f cyc/tile  (MMA floor %6.0f, eff %5.1f%%)\n", nsm, mix, names[mix], avg, floors[mix], 100.0 * floors[mix] / avg);
  }
  {
    const int smem3 = 98304 + 1024 + 64;
    cudaFuncSetAttribute(probe_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
    const char* pn[] = {"", "fwd pair: 8 TS2 K-major + 8 TS2 MN-major (N=128)", "8 SS2 + 8 TS2 MN-major (N=128)", "16 TS2 K-major (N=128)", "GEMM k-block: 4 SS2 N=256"};
    const double pf[] = {0, 1024, 1024, 1024, 512};
    for (int pmix = 1; pmix <= 4; ++pmix) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nsm);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = smem3;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      for (int rep = 0; rep < 2; ++rep) {
        cudaLaunchKernelEx(&cfg, probe_pair, pmix, 512, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pmix %d: %s\n", pmix, cudaGetErrorString(e)); return 1; }
      }
      double avg = 0;
      for (int i = 0; i < nsm / 2; ++i) avg += double(out[i]);
      avg /= (nsm / 2) * 512.0;
      printf("pairs %3d  pmix %d %-52s %8.1f cyc/iter  (MMA floor %6.0f, eff %5.1f%%)\n", nsm / 2, pmix, pn[pmix], avg, pf[pmix], 100.0 * pf[pmix] / avg);
    }
  }
  return 0;
}
