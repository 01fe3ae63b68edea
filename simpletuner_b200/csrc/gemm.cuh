// simpletuner_b200 — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[b, s, n] = epilogue( sum_seg  A_seg[b, s, :] . W_seg[n, :]  + bias[n] )
//
// Every operand is K-major bf16 ("TN": activations [rows, K], nn.Linear weights [N, K]), fp32
// accumulation in TMEM.  Up to three K-segments are chained in one mainloop so that
//   * torch.cat([attn, mlp], -1) @ W^T          (reference flux/transformer.py:460-464)
//   * x @ W^T + (x A^T) B^T   (PEFT LoRA linear, reference common.py:1094-1117)
// become extra k-blocks of the same tile instead of extra kernels / extra HBM round trips.
//
// Roles (256 threads): warp0 = TMA producer, warp1 = MMA issuer (one thread), warp2 = TMEM
// allocator, warps 4..7 = epilogue (TMEM -> registers -> fused epilogue -> global).
// A CTA owns MT x 128 rows and BN columns of D per tile; TMEM holds 512/(MT*BN) accumulator
// stages so the epilogue of tile i overlaps the mainloop of tile i+1 when there are >= 2.
// PAIR = true is the default for large problems: a 2-CTA cluster (cta_group::2) owns a 256 x BN tile, each CTA
// stages half of it, the leader issues M = 256 MMAs (see the comment at the kernel).  CONV = true turns the A
// operand into the shifted NHWC window of a 3x3 convolution (implicit GEMM, VAE encoder).
#pragma once
#include <type_traits>
#include "common.cuh"

namespace stb {

enum GemmEpi : int {
  EPI_STORE = 0,        // D = acc + bias
  EPI_GELU = 1,         // D = gelu_tanh(acc + bias); aux (optional) = acc + bias (pre-activation)
  EPI_GATE_RES = 2,     // D = res + gate[b, n] * (acc + bias); optional nan_to_num; aux (optional) = acc + bias
  EPI_MUL_DGELU = 3,    // D = acc * gelu_tanh'(aux)          (dgrad through the activation)
  EPI_ADD_RES = 4,      // D = acc + bias + res                (gradient accumulation; residual add of the text encoders)
  EPI_MUL = 5,          // D = (acc + bias) * aux              (T5 gated feed-forward: gelu(x wi_0) * (x wi_1))
  EPI_QUICK_GELU = 6,   // D = y * sigmoid(1.702 y), y = bf16(acc + bias)   (CLIP text model activation)
};

struct GemmParams {
  int rows_per_batch;  // S : rows in one batch slab of A / D
  int num_batches;     // B
  int N;
  int nseg;
  int kblocks[3];  // ceil(K_seg / 64)
  int kmmas_last[3];  // UMMA_K=16 steps needed in the last k-block of the segment (1..4)
  int w_kn[3];        // 1: the segment's weight is given as [K, N] row-major (contraction index = row): the B operand is
                      // staged MN-major (64 k-rows x 64 n-columns SWIZZLE_128B boxes) — dgrad reads W itself, no W^T copy
  int epi;
  int nan_to_num;
  __nv_bfloat16* D;
  long long d_batch_stride, d_row_stride;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* gate;
  long long gate_batch_stride;
  const __nv_bfloat16* res;
  long long res_batch_stride, res_row_stride;
  __nv_bfloat16* aux;  // EPI_GELU / EPI_GATE_RES: written (optional); EPI_MUL_DGELU: read
  long long aux_batch_stride, aux_row_stride;
  // CONV mode (3x3 NHWC implicit GEMM): a "batch" is one output image row, a "row" an output pixel x.
  // maps.a[0] is then a 4-D map (c, x, y, img) with box (64, 128, 1, 1) and x element-stride = conv_stride;
  // k-block kb covers tap kb / conv_cblocks (dy = tap / 3, dx = tap % 3) and channels (kb % conv_cblocks) * 64.
  int conv_h_out, conv_stride, conv_pad, conv_cblocks;
  // CONV + pair: 1 = the two CTAs take pixels [s0, s0+128) and [s0+128, s0+256) of one output row (W_out > 128);
  //              2 = they take two consecutive output rows of <= 128 pixels each
  int conv_pair_rows;
};

struct GemmMaps {
  CUtensorMap a[3];  // 3-D (k, s, b), box (64, 128, 1), SWIZZLE_128B
  CUtensorMap w[3];  // 2-D (k, n),   box (64, BN),      SWIZZLE_128B;  w_kn: 2-D (n, k), box (64, 64)
};

template <int MT, int BN>
struct GemmCfg {
  static constexpr int BM = 128 * MT;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = 128 * BK * 2;          // one 128-row A tile
  static constexpr int W_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = MT * A_BYTES + W_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int ACC_COLS = MT * BN;
  static constexpr int ACC_STAGES = (512 / ACC_COLS) >= 2 ? 2 : 1;
  static constexpr int TMEM_COLS = (ACC_COLS * ACC_STAGES) <= 32    ? 32
                                   : (ACC_COLS * ACC_STAGES) <= 64  ? 64
                                   : (ACC_COLS * ACC_STAGES) <= 128 ? 128
                                   : (ACC_COLS * ACC_STAGES) <= 256 ? 256
                                                                    : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// CTA-pair tile: 256 x BN per 2-CTA cluster; per CTA 128 rows of A + BN/2 rows of W per stage, 128 x BN accumulator
template <int BN>
struct GemmPairCfg {
  static constexpr int BM = 256;
  static constexpr int BK = 64;
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int W_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int ACC_COLS = BN;
  static constexpr int ACC_STAGES = (512 / ACC_COLS) >= 2 ? 2 : 1;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

__device__ __forceinline__ float gemm_bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

__device__ __forceinline__ void gemm_tile_coords(int tile, int tiles_m, int tiles_n, int& tm, int& tn) {
  // grouped rasterisation: bands of 8 m-tiles sweep all n-tiles -> square-ish L2 footprint
  constexpr int GM = 8;
  int per_group = GM * tiles_n;
  int g = tile / per_group;
  int first_m = g * GM;
  int gsize = min(GM, tiles_m - first_m);
  int within = tile - g * per_group;
  tm = first_m + within % gsize;
  tn = within / gsize;
}

// PAIR = true: cta_group::2 kernel (launch with cluster dims (2,1,1), MT must be 1).  The cluster owns a 256 x BN
// tile: CTA rank r stages A rows [128 r, 128 r + 128) and W rows [BN/2 r, BN/2 r + BN/2) of every k-block in its own
// shared memory (6 KB -> 4 KB of operand fetch per SM per MMA k-step at BN = 256), the leader issues one M = 256
// tcgen05.mma per k-step for both SMs, and each CTA drains its own 128 accumulator lanes.
template <int MT, int BN, bool CONV = false, bool PAIR = false>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_tn_kernel(const __grid_constant__ GemmMaps maps, const GemmParams p) {
  using Cfg = typename std::conditional<PAIR, GemmPairCfg<BN>, GemmCfg<MT, BN>>::type;
  static_assert(!PAIR || MT == 1, "pair kernel: MT = 1");
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int tile0 = PAIR ? int(blockIdx.x >> 1) : int(blockIdx.x);
  const int tile_step = PAIR ? int(gridDim.x >> 1) : int(gridDim.x);
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC_STAGES = Cfg::ACC_STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto accf_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto acce_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + ACC_STAGES + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC_STAGES);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const bool row_pairs = PAIR && CONV && p.conv_pair_rows == 2;
  const int tiles_per_batch = row_pairs ? 1 : (p.rows_per_batch + Cfg::BM - 1) / Cfg::BM;
  const int tiles_m = row_pairs ? (p.num_batches + 1) / 2 : tiles_per_batch * p.num_batches;
  // (batch, first row) of this CTA's 128-row half of m-tile `tm`
  auto tile_origin = [&](int tm, int& b, int& s0) {
    if (row_pairs) {
      b = 2 * tm + int(rank);
      s0 = 0;
    } else {
      b = tm / tiles_per_batch;
      s0 = (tm - b * tiles_per_batch) * Cfg::BM + (PAIR ? int(rank) * 128 : 0);
    }
  };
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  int kb_total = 0;
  for (int s = 0; s < p.nseg; ++s) kb_total += p.kblocks[s];

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&maps.a[s]);
      tma_prefetch_desc(&maps.w[s]);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      mbar_init(accf_bar(s), 1);
      mbar_init(acce_bar(s), PAIR ? 8 : 4);   // pair: the leader's barrier collects both CTAs' epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform loop, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        int tm, tn;
        gemm_tile_coords(tile, tiles_m, tiles_n, tm, tn);
        int b, s0;
        tile_origin(tm, b, s0);
        const int n0 = tn * BN;
        for (int seg = 0; seg < p.nseg; ++seg) {
          for (int kb = 0; kb < p.kblocks[seg]; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u, 1);
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            if (elect_one()) {
              if constexpr (PAIR) {
                // both CTAs' bytes are accounted on the leader's barrier (the MMA issuer waits there)
                if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
                if constexpr (CONV) {
                  const int tap = kb / p.conv_cblocks;
                  const int c0 = (kb - tap * p.conv_cblocks) * 64;
                  const int img = b / p.conv_h_out, yo = b - img * p.conv_h_out;   // b past the last row -> img out of range -> zeros
                  const int dy = tap / 3, dx = tap - dy * 3;
                  tma_load_4d_pair(sa, &maps.a[0], full_bar(stage), c0, s0 * p.conv_stride + dx - p.conv_pad,
                                   yo * p.conv_stride + dy - p.conv_pad, img);
                } else {
                  tma_load_3d_pair(sa, &maps.a[seg], full_bar(stage), kb * 64, s0, b);
                }
                if (p.w_kn[seg]) {
#pragma unroll
                  for (int g = 0; g < BN / 128; ++g)
                    tma_load_2d_pair(sa + Cfg::A_BYTES + g * 8192, &maps.w[seg], full_bar(stage), n0 + int(rank) * (BN / 2) + g * 64, kb * 64);
                } else {
                  tma_load_2d_pair(sa + Cfg::A_BYTES, &maps.w[seg], full_bar(stage), kb * 64, n0 + int(rank) * (BN / 2));
                }
              } else {
              mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
              if constexpr (CONV) {
                // implicit-GEMM 3x3 conv: shifted input window; TMA zero-fills the halo (x / y out of range)
                const int tap = kb / p.conv_cblocks;
                const int c0 = (kb - tap * p.conv_cblocks) * 64;
                const int img = b / p.conv_h_out, yo = b - img * p.conv_h_out;
                const int dy = tap / 3, dx = tap - dy * 3;
                tma_load_4d(sa, &maps.a[0], full_bar(stage), c0, s0 * p.conv_stride + dx - p.conv_pad,
                            yo * p.conv_stride + dy - p.conv_pad, img);
              } else {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
                  tma_load_3d(sa + mt * Cfg::A_BYTES, &maps.a[seg], full_bar(stage), kb * 64, s0 + mt * 128, b);
              }
              if (p.w_kn[seg]) {
#pragma unroll
                for (int g = 0; g < BN / 64; ++g)
                  tma_load_2d(sa + MT * Cfg::A_BYTES + g * 8192, &maps.w[seg], full_bar(stage), n0 + g * 64, kb * 64);
              } else {
                tma_load_2d(sa + MT * Cfg::A_BYTES, &maps.w[seg], full_bar(stage), kb * 64, n0);
              }
              }
            }
            __syncwarp();
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      constexpr uint32_t idesc_k = make_idesc_bf16(PAIR ? 256 : 128, BN, 0, 0);
      constexpr uint32_t idesc_mn = make_idesc_bf16(PAIR ? 256 : 128, BN, 0, 1);   // B operand MN-major (w_kn segments)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(acce_bar(acc), acc_phase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_base = tmem_base + acc * Cfg::ACC_COLS;
        uint32_t accumulate = 0;
        for (int seg = 0; seg < p.nseg; ++seg) {
          for (int kb = 0; kb < p.kblocks[seg]; ++kb) {
            mbar_wait(full_bar(stage), phase, 3);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint32_t sw = sa + MT * Cfg::A_BYTES;
            const int nk = (kb == p.kblocks[seg] - 1) ? p.kmmas_last[seg] : 4;
            const bool wkn = p.w_kn[seg] != 0;
            const uint32_t idesc = wkn ? idesc_mn : idesc_k;
            if (elect_one()) {
              for (int kk = 0; kk < nk; ++kk) {
                const uint64_t bdesc = wkn ? sdesc_mn(sw, kk * 2048, 8192) : sdesc_k(sw, kk * 32);
                if constexpr (PAIR) {
                  mma_ss2(d_base, sdesc_k(sa, kk * 32), bdesc, idesc, accumulate);
                } else {
#pragma unroll
                  for (int mt = 0; mt < MT; ++mt) {
                    const uint64_t adesc = sdesc_k(sa, mt * Cfg::A_BYTES + kk * 32);
                    mma_ss(d_base + mt * BN, adesc, bdesc, idesc, accumulate);
                  }
                }
                accumulate = 1;
              }
              // frees the smem slot (in both CTAs of a pair) once these MMAs retire
              if constexpr (PAIR) tc_commit2(empty_bar(stage)); else tc_commit(empty_bar(stage));
            }
            __syncwarp();
            accumulate = 1;
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        if (elect_one()) {  // accumulator complete -> epilogue (of both CTAs of a pair)
          if constexpr (PAIR) tc_commit2(accf_bar(acc)); else tc_commit(accf_bar(acc));
        }
        __syncwarp();
        if (++acc == ACC_STAGES) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      int tm, tn;
      gemm_tile_coords(tile, tiles_m, tiles_n, tm, tn);
      int b, s0;
      tile_origin(tm, b, s0);
      const int n0 = tn * BN;
      mbar_wait(accf_bar(acc), acc_phase, 4);
      tc_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        const int s = s0 + mt * 128 + ew * 32 + lane;
        const bool row_ok = s < p.rows_per_batch && b < p.num_batches;
        __nv_bfloat16* drow = p.D + (long long)b * p.d_batch_stride + (long long)s * p.d_row_stride;
        const __nv_bfloat16* rrow =
            p.res ? p.res + (long long)b * p.res_batch_stride + (long long)s * p.res_row_stride : nullptr;
        __nv_bfloat16* xrow =
            p.aux ? p.aux + (long long)b * p.aux_batch_stride + (long long)s * p.aux_row_stride : nullptr;
        const __nv_bfloat16* grow = p.gate ? p.gate + (long long)b * p.gate_batch_stride : nullptr;
        const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + acc * Cfg::ACC_COLS + mt * BN;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          if (n0 + c >= p.N) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + c, v);
          tc_wait_ld();
          const int n = n0 + c;
          const bool full = (n + 32 <= p.N);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
            if (full) {
              const uint4* bp = reinterpret_cast<const uint4*>(p.bias + n);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 u = __ldg(bp + q);
                f[q * 8 + 0] += bf16_lo(u.x); f[q * 8 + 1] += bf16_hi(u.x);
                f[q * 8 + 2] += bf16_lo(u.y); f[q * 8 + 3] += bf16_hi(u.y);
                f[q * 8 + 4] += bf16_lo(u.z); f[q * 8 + 5] += bf16_hi(u.z);
                f[q * 8 + 6] += bf16_lo(u.w); f[q * 8 + 7] += bf16_hi(u.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n + j < p.N) f[j] += __bfloat162float(p.bias[n + j]);
            }
          }
          if (row_ok) {
            if (p.epi == EPI_GELU) {
              if (xrow) {
                if (full) {
                  uint4* xp = reinterpret_cast<uint4*>(xrow + n);
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    uint4 u;
                    u.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
                    u.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
                    u.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
                    u.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
                    xp[q] = u;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (n + j < p.N) xrow[n + j] = __float2bfloat16(f[j]);
                }
              }
              // the activation sees the bf16-rounded pre-activation, as the reference's
              // nn.Linear -> nn.GELU chain does (bf16 tensor between the two modules)
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = gelu_tanh(__bfloat162float(__float2bfloat16(f[j])));
            } else if (p.epi == EPI_GATE_RES) {
              float g[32], r[32];
              if (full) {
                const uint4* gp = reinterpret_cast<const uint4*>(grow + n);
                const uint4* rp = reinterpret_cast<const uint4*>(rrow + n);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 u = __ldg(gp + q);
                  g[q * 8 + 0] = bf16_lo(u.x); g[q * 8 + 1] = bf16_hi(u.x);
                  g[q * 8 + 2] = bf16_lo(u.y); g[q * 8 + 3] = bf16_hi(u.y);
                  g[q * 8 + 4] = bf16_lo(u.z); g[q * 8 + 5] = bf16_hi(u.z);
                  g[q * 8 + 6] = bf16_lo(u.w); g[q * 8 + 7] = bf16_hi(u.w);
                  uint4 w = rp[q];
                  r[q * 8 + 0] = bf16_lo(w.x); r[q * 8 + 1] = bf16_hi(w.x);
                  r[q * 8 + 2] = bf16_lo(w.y); r[q * 8 + 3] = bf16_hi(w.y);
                  r[q * 8 + 4] = bf16_lo(w.z); r[q * 8 + 5] = bf16_hi(w.z);
                  r[q * 8 + 6] = bf16_lo(w.w); r[q * 8 + 7] = bf16_hi(w.w);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  g[j] = (n + j < p.N) ? __bfloat162float(grow[n + j]) : 0.f;
                  r[j] = (n + j < p.N) ? __bfloat162float(rrow[n + j]) : 0.f;
                }
              }
              // reference order: bias -> (bf16 linear output) -> gate * y -> residual + (...)
              // (flux/transformer.py:464-465, 584-586, 652-653); each step is a bf16 tensor there.
              if (xrow) {   // full fine-tune: keep the (bf16) linear output — the gate gradient is sum_s dOut * y
                if (full) {
                  uint4* xp = reinterpret_cast<uint4*>(xrow + n);
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    uint4 u;
                    u.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
                    u.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
                    u.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
                    u.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
                    xp[q] = u;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (n + j < p.N) xrow[n + j] = __float2bfloat16(f[j]);
                }
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float y = __bfloat162float(__float2bfloat16(f[j]));
                float gy = __bfloat162float(__float2bfloat16(g[j] * y));
                float o = r[j] + gy;
                if (p.nan_to_num) {
                  o = __bfloat162float(__float2bfloat16(o));
                  if (o != o) o = 0.f;
                  else if (o == INFINITY) o = 65504.f;
                  else if (o == -INFINITY) o = -65504.f;
                }
                f[j] = o;
              }
            } else if (p.epi == EPI_MUL_DGELU) {
              if (full) {
                const uint4* xp = reinterpret_cast<const uint4*>(xrow + n);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 u = xp[q];
                  f[q * 8 + 0] *= gelu_tanh_grad(bf16_lo(u.x)); f[q * 8 + 1] *= gelu_tanh_grad(bf16_hi(u.x));
                  f[q * 8 + 2] *= gelu_tanh_grad(bf16_lo(u.y)); f[q * 8 + 3] *= gelu_tanh_grad(bf16_hi(u.y));
                  f[q * 8 + 4] *= gelu_tanh_grad(bf16_lo(u.z)); f[q * 8 + 5] *= gelu_tanh_grad(bf16_hi(u.z));
                  f[q * 8 + 6] *= gelu_tanh_grad(bf16_lo(u.w)); f[q * 8 + 7] *= gelu_tanh_grad(bf16_hi(u.w));
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) f[j] *= gelu_tanh_grad(__bfloat162float(xrow[n + j]));
              }
            } else if (p.epi == EPI_MUL) {
              if (full) {
                const uint4* xp = reinterpret_cast<const uint4*>(xrow + n);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 u = xp[q];
                  f[q * 8 + 0] = gemm_bf16r(f[q * 8 + 0]) * bf16_lo(u.x); f[q * 8 + 1] = gemm_bf16r(f[q * 8 + 1]) * bf16_hi(u.x);
                  f[q * 8 + 2] = gemm_bf16r(f[q * 8 + 2]) * bf16_lo(u.y); f[q * 8 + 3] = gemm_bf16r(f[q * 8 + 3]) * bf16_hi(u.y);
                  f[q * 8 + 4] = gemm_bf16r(f[q * 8 + 4]) * bf16_lo(u.z); f[q * 8 + 5] = gemm_bf16r(f[q * 8 + 5]) * bf16_hi(u.z);
                  f[q * 8 + 6] = gemm_bf16r(f[q * 8 + 6]) * bf16_lo(u.w); f[q * 8 + 7] = gemm_bf16r(f[q * 8 + 7]) * bf16_hi(u.w);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) f[j] = gemm_bf16r(f[j]) * __bfloat162float(xrow[n + j]);
              }
            } else if (p.epi == EPI_QUICK_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float y = gemm_bf16r(f[j]);
                f[j] = y / (1.f + __expf(-1.702f * y));
              }
            } else if (p.epi == EPI_ADD_RES) {
              if (full) {
                const uint4* rp = reinterpret_cast<const uint4*>(rrow + n);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 w = rp[q];
                  f[q * 8 + 0] += bf16_lo(w.x); f[q * 8 + 1] += bf16_hi(w.x);
                  f[q * 8 + 2] += bf16_lo(w.y); f[q * 8 + 3] += bf16_hi(w.y);
                  f[q * 8 + 4] += bf16_lo(w.z); f[q * 8 + 5] += bf16_hi(w.z);
                  f[q * 8 + 6] += bf16_lo(w.w); f[q * 8 + 7] += bf16_hi(w.w);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n + j < p.N) f[j] += __bfloat162float(rrow[n + j]);
              }
            }
            if (full) {
              uint4* dp = reinterpret_cast<uint4*>(drow + n);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 u;
                u.x = pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]);
                u.y = pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]);
                u.z = pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]);
                u.w = pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]);
                dp[q] = u;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n + j < p.N) drow[n + j] = __float2bfloat16(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_remote(acce_bar(acc), 0u); else mbar_arrive(acce_bar(acc));
      }
      if (++acc == ACC_STAGES) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();   // pair: the peer may still touch our smem / barriers / TMEM
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace stb
