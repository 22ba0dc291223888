// K3 — 3x3 implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   out[p, n] = epilogue( sum_{tap, c} in[p + tap, c] * wpk[tap][n][c] )
//
// * operands: fp32 in HBM, read as TF32 (kind::tf32), fp32 accumulation in TMEM;
// * A tile = 128 output pixels as a (batch, rows, cols) box of the NHWC input
//   view, fetched by ONE 4-D TMA load per (tap, 32-channel chunk): the tap shift is
//   a coordinate offset and the conv zero padding is TMA's out-of-bounds fill, so
//   there is no im2col buffer and no halo logic; the strided view that implements
//   the nearest down-sample of the segmap is expressed in the tensor map strides;
// * B tile = BN x 32 slab of the packed weights [9][Nout][Cin], 3-D TMA load;
// * both land in 128-byte-swizzled K-major shared tiles consumed by tcgen05.mma
//   (M=128, N=BN, K=8 per instruction, 4 instructions per stage);
// * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc),
//   warps 2..5 = epilogue (tcgen05.ld, one accumulator row = one pixel per thread);
// * persistent CTAs (one per SM) walk a static schedule of work items; 4-6 stage mbarrier
//   ring for the operands and TWO accumulator stages in TMEM, so the epilogue of one work
//   item overlaps the main loop of the next.  Work item = 1 or 2 M tiles x 128 or 256 columns.
#include <stdlib.h>
#include <string.h>
#include "k3_common.cuh"
#include "tc_ptx.cuh"

namespace ag2v {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_THREADS = 192;

struct TcGeom { int Wt, Ht, Bt, tiles_x, tiles_y, tiles_b, mtiles, ntiles, work, ksplit, its_per_split; };

// MT = M tiles (128 pixels each) per work item, BN = output columns per work item.
// L2 -> smem operand traffic per MMA cycle is (MT*16 KB + BN*128 B) / (MT * BN * 2 cycles):
//   MT=1,BN=128: 128 B/cycle     MT=2,BN=128 or MT=1,BN=256: 96 B/cycle
// Two accumulator stages in TMEM (2 * MT * BN <= 512 columns) let the epilogue of one
// work item overlap the main loop of the next.
template <int MT, int BN>
struct TcCfg {
  static constexpr int kA = TC_BM * TC_BK * 4;                 // 16 KB per M tile
  static constexpr int kB = BN * TC_BK * 4;                    // 16 / 32 KB
  static constexpr int kStage = MT * kA + kB;
  static constexpr int kStages = (192 * 1024) / kStage;        // 6 (32 KB stages) or 4 (48 KB stages)
  static constexpr int kAccStages = 2;
  static constexpr int kAccCols = MT * BN;
  static constexpr int kTmemCols = kAccStages * kAccCols;      // 256 or 512
  static constexpr int kStatBytes = 4 * 2 * BN * 4;             // [epilogue warp][sum | sum of squares][BN] staging of the fused BN statistics
  static constexpr int kBytes = kStages * kStage + 1024 /*align*/ + 256 /*barriers*/ + kStatBytes;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

// one 32-column chunk of one accumulator row (= one pixel): v[] holds the fp32 sums
template <int EPI, bool ROUND_OUT>
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, float (&v)[32], int n, long long pp, long long xpix,
                                               float* orow) {
  if (EPI == EPI_SPADE) {
    const int c = n >> 1;                                           // 16 channels: [g8 | b8 | g8 | b8]
    const size_t off = (size_t)pp * p.C + c;
    const size_t xoff = (size_t)xpix * p.C + c;
    const size_t sc = (p.group_pixels > 0 ? (size_t)(pp / p.group_pixels) * p.C : 0) + c;
    float xs[16], mu[16], rs[16], o[16], gm_[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float4*>(xs + 4 * i) = *reinterpret_cast<const float4*>(p.x + xoff + 4 * i);
      *reinterpret_cast<float4*>(mu + 4 * i) = *reinterpret_cast<const float4*>(p.mean + sc + 4 * i);
      *reinterpret_cast<float4*>(rs + 4 * i) = *reinterpret_cast<const float4*>(p.rstd + sc + 4 * i);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int jg = (i >> 3) * 16 + (i & 7), jb = jg + 8;
      float g = v[jg], bb = v[jb];
      if (p.bias) { g += p.bias[n + jg]; bb += p.bias[n + jb]; }
      float r = (xs[i] - mu[i]) * rs[i] * (1.f + g) + bb;
      if (p.slope != 1.f) r = r > 0.f ? r : r * p.slope;
      if (ROUND_OUT) r = tc_round_tf32(r);
      o[i] = r; gm_[i] = g;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      *reinterpret_cast<float4*>(orow + c + 4 * i) = *reinterpret_cast<const float4*>(o + 4 * i);
      if (p.gamma_out) *reinterpret_cast<float4*>(p.gamma_out + off + 4 * i) = *reinterpret_cast<const float4*>(gm_ + 4 * i);
    }
  } else {
    const int ncols = min(32, p.Nout - n);
    if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
      if (EPI == EPI_BIAS && p.scale) {
        const float sc = p.scale[p.group_pixels > 0 ? pp / p.group_pixels : 0];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= sc;
      }
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += p.bias[n + j];
      }
      if (EPI == EPI_BIAS && p.res) {
        const float* rr = p.res + (size_t)pp * p.Nout + n;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (j < ncols) {
            const float4 r4 = *reinterpret_cast<const float4*>(rr + j);
            v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
          }
        }
      }
      if (EPI == EPI_BIAS_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
    } else if (EPI == EPI_GATE) {
      const float* gt = p.gate + (size_t)pp * p.Nout + n;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) {
          const float4 g4 = *reinterpret_cast<const float4*>(gt + j);
          v[j] = g4.x > 0.f ? v[j] : 0.f; v[j + 1] = g4.y > 0.f ? v[j + 1] : 0.f;
          v[j + 2] = g4.z > 0.f ? v[j + 2] : 0.f; v[j + 3] = g4.w > 0.f ? v[j + 3] : 0.f;
        }
      }
    } else if (EPI == EPI_ACCUM) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < ncols) {
          const float4 o4 = *reinterpret_cast<const float4*>(orow + n + j);
          v[j] += o4.x; v[j + 1] += o4.y; v[j + 2] += o4.z; v[j + 3] += o4.w;
        }
      }
    }
    if (ROUND_OUT) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = tc_round_tf32(v[j]);
    }
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      if (j < ncols) *reinterpret_cast<float4*>(orow + n + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
}

// Column sums over the 32 lanes of a warp for 32 columns at once (lane j ends up with the sum of column j): a
// butterfly in which every round halves the columns a lane still carries - 31 shuffles instead of 32 x 5.
template <int STEP>
__device__ __forceinline__ void col_sum_round(float (&v)[32], int lane) {
  const bool up = (lane & STEP) != 0;
#pragma unroll
  for (int i = 0; i < STEP; ++i) {
    const float lo = v[i], hi = v[i + STEP];
    const float send = up ? lo : hi, keep = up ? hi : lo;
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, STEP);
  }
}
__device__ __forceinline__ float col_sum32(float (&v)[32], int lane) {
  col_sum_round<16>(v, lane); col_sum_round<8>(v, lane); col_sum_round<4>(v, lane);
  col_sum_round<2>(v, lane); col_sum_round<1>(v, lane);
  return v[0];
}

// work item w -> (column tile, first M tile); column tile fastest so that CTAs running side by
// side share the activation tile in L2.  tile -> (x0, y0, b0); a tile past the end gives b0 >= B
// (every row masked, TMA fills zeros).
__device__ __forceinline__ void tile_origin(const TcGeom& gm, int tile, int& x0, int& y0, int& b0) {
  const int tx = tile % gm.tiles_x; tile /= gm.tiles_x;
  const int ty = tile % gm.tiles_y; tile /= gm.tiles_y;
  x0 = tx * gm.Wt; y0 = ty * gm.Ht; b0 = tile * gm.Bt;
}

// Persistent: grid = min(work items, SMs).  Work items are handed out DYNAMICALLY: the TMA lane takes the next index
// from a global counter (p.sched[0]) and publishes it to the MMA lane and the four epilogue warps through a 4-deep
// shared-memory ring (mbarriers sfull / sempty).  With a static schedule (w = blockIdx.x + k * gridDim.x) a CTA that
// cannot be placed because another kernel holds its SM - NCCL's gradient all-reduce during the backward pass - keeps
// its share of the tiles waiting and the whole convolution ends only after that other kernel; here the CTAs that run
// take all the work and a late CTA finds the counter exhausted and leaves.  The last CTA to leave resets the counter.
constexpr int TC_SCHED = 4;
template <int MT, int BN, int EPI, bool ROUND_OUT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  ConvParams p, TcGeom gm) {
  using Cfg = TcCfg<MT, BN>;
  constexpr int S = Cfg::kStages, AS = Cfg::kAccStages;
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                         // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bars = base + S * Cfg::kStage;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * S;            // smem ring
  const uint32_t bar_tfull = bars + 16 * S, bar_tempty = bar_tfull + 8 * AS;   // accumulator stages
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_smem_raw + (bars - raw) + 16 * S + 16 * AS);
  static_assert(16 * S + 16 * AS + 8 <= 136, "barrier block layout");
  const uint32_t bar_sfull = bars + 136, bar_sempty = bar_sfull + 8 * TC_SCHED;       // work-item ring
  volatile int* sched_w = reinterpret_cast<volatile int*>(tc_smem_raw + (bars - raw) + 136 + 16 * TC_SCHED);   // 200..216 of 256
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = p.Cin / TC_BK;
  const int total = 9 * KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int a = 0; a < AS; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
    for (int j = 0; j < TC_SCHED; ++j) { mbar_init(bar_sfull + 8 * j, 1); mbar_init(bar_sempty + 8 * j, 5); }   // MMA lane + 4 epilogue warps
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_b); }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g_it = 0;
      for (uint32_t j = 0;; ++j) {
        int w = p.sched ? (int)atomicAdd(p.sched, 1u) : (int)(blockIdx.x + j * gridDim.x);     // null: static round robin
        if (w >= gm.work) w = -1;
        const uint32_t sl = j % TC_SCHED;
        mbar_wait(bar_sempty + 8 * sl, ((j / TC_SCHED) & 1u) ^ 1u);     // all five readers are done with this ring slot
        sched_w[sl] = w;
        mbar_arrive(bar_sfull + 8 * sl);
        if (w < 0) break;
        const int split = w % gm.ksplit, wt = w / gm.ksplit;
        const int n0 = (wt % gm.ntiles) * BN, mt0 = (wt / gm.ntiles) * MT;
        const int it0 = split * gm.its_per_split, it1 = min(total, it0 + gm.its_per_split);
        int x0[MT], y0[MT], b0[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) tile_origin(gm, mt0 + mt, x0[mt], y0[mt], b0[mt]);
        for (int it = it0; it < it1; ++it, ++g_it) {
          const uint32_t s = g_it % S, ph = (g_it / S) & 1u;
          mbar_wait(bar_empty + 8 * s, ph ^ 1u);                        // slot free
          const int tap = it / KC, kc = it - tap * KC;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          const uint32_t a_dst = base + s * Cfg::kStage, b_dst = a_dst + MT * Cfg::kA;
          mbar_expect_tx(bar_full + 8 * s, Cfg::kStage);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
            tma_load_4d(a_dst + mt * Cfg::kA, &map_a, bar_full + 8 * s, kc * TC_BK, x0[mt] + dx, y0[mt] + dy, b0[mt]);
          tma_load_3d(b_dst, &map_b, bar_full + 8 * s, kc * TC_BK, n0, tap);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N=BN, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      uint32_t g_it = 0, a_it = 0;
      for (;; ++a_it) {
        const uint32_t sl = a_it % TC_SCHED;
        mbar_wait(bar_sfull + 8 * sl, (a_it / TC_SCHED) & 1u);
        const int w = sched_w[sl];
        mbar_arrive(bar_sempty + 8 * sl);
        if (w < 0) break;
        const uint32_t as = a_it % AS, aph = (a_it / AS) & 1u;
        const int split = w % gm.ksplit;
        const int it0 = split * gm.its_per_split, it1 = min(total, it0 + gm.its_per_split);
        mbar_wait(bar_tempty + 8 * as, aph ^ 1u);                       // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t acc = tmem_base + as * Cfg::kAccCols;
        for (int it = it0; it < it1; ++it, ++g_it) {
          const uint32_t s = g_it % S, ph = (g_it / S) & 1u;
          mbar_wait(bar_full + 8 * s, ph);                              // TMA bytes landed
          tc_fence_after();
          const uint32_t a_s = base + s * Cfg::kStage, b_s = a_s + MT * Cfg::kA;
          const uint64_t db = make_sw128_kmajor_desc(b_s);
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {                         // +32 bytes along K inside the swizzle atom
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t da = make_sw128_kmajor_desc(a_s + mt * Cfg::kA);
              umma_tf32(acc + (uint32_t)(mt * BN), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it > it0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(bar_empty + 8 * s);                               // frees the smem slot when the MMAs retire
        }
        umma_commit(bar_tfull + 8 * as);                                // accumulators of this work item complete
      }
    }
  } else {
    // ---- epilogue: warps 2..5, TMEM lane quarter = warp % 4 --------------------
    const int q = warp & 3;
    float* stat_s = reinterpret_cast<float*>(tc_smem_raw + (bars - raw) + 256);    // [4][2][BN]
    const bool stats = (EPI == EPI_BIAS) && p.stat_part != nullptr;
    const int m = q * 32 + lane;                                       // accumulator row = pixel within the tile
    const int xt = m % gm.Wt, yt = (m / gm.Wt) % gm.Ht, bt = m / (gm.Wt * gm.Ht);
    for (uint32_t a_it = 0;; ++a_it) {
      const uint32_t sl = a_it % TC_SCHED;
      mbar_wait(bar_sfull + 8 * sl, (a_it / TC_SCHED) & 1u);
      const int w = sched_w[sl];
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty + 8 * sl);
      if (w < 0) break;
      const uint32_t as = a_it % AS, aph = (a_it / AS) & 1u;
      const int split = w % gm.ksplit, wt = w / gm.ksplit;
      const int n0 = (wt % gm.ntiles) * BN, mt0 = (wt / gm.ntiles) * MT;
      mbar_wait(bar_tfull + 8 * as, aph);
      tc_fence_after();
      const uint32_t acc = tmem_base + as * Cfg::kAccCols + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt) {
        int x0, y0, b0;
        tile_origin(gm, mt0 + mt, x0, y0, b0);
        const int x = x0 + xt, y = y0 + yt, b = b0 + bt;
        const bool valid = x < p.Ww && y < p.Hh && b < p.B;
        const long long pp = ((long long)b * p.Hh + y) * p.Ww + x;
        const long long xpix = p.x_up ? ((long long)b * (p.Hh >> 1) + (y >> 1)) * (p.Ww >> 1) + (x >> 1) : pp;
        float* orow = (EPI == EPI_RAW)
            ? p.splitk_ws + ((size_t)split * p.B * p.Hh * p.Ww + (size_t)pp) * p.Nout        // partial sums [split][pixel][n]
            : p.out + (long long)b * p.out_sb + (long long)y * p.out_sy + (long long)x * p.out_sx;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          float v[32];
          tmem_ld32(acc + (uint32_t)(mt * BN + ch * 32), v);
          if (mt == MT - 1 && ch == BN / 32 - 1) {                      // last TMEM read of this stage: hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
          }
          const int n = n0 + ch * 32;
          if (valid && n < p.Nout) {
            if (EPI == EPI_RAW) {
              const int ncols = min(32, p.Nout - n);
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (j < ncols) *reinterpret_cast<float4*>(orow + n + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              epilogue_chunk<EPI, ROUND_OUT>(p, v, n, pp, xpix, orow);
            }
          }
          if (EPI == EPI_BIAS && stats) {
            // BN statistics of what was just stored (v holds the final values): per-column sums over this warp's
            // 32 pixels by warp shuffles, staged per warp; pixels / columns outside the problem contribute zero
            float w2[32];
            const int ncols = n < p.Nout ? min(32, p.Nout - n) : 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (!valid || j >= ncols) v[j] = 0.f;
              w2[j] = v[j] * v[j];
            }
            const float s1 = col_sum32(v, lane), s2 = col_sum32(w2, lane);
            stat_s[(q * 2 + 0) * BN + ch * 32 + lane] = s1;
            stat_s[(q * 2 + 1) * BN + ch * 32 + lane] = s2;
          }
        }
        if (EPI == EPI_BIAS && stats) {
          // the four epilogue warps (128 threads, named barrier 2) combine their quarters in warp order
          asm volatile("bar.sync 2, 128;\n" ::: "memory");
          const int et = threadIdx.x - 64;
          for (int i = et; i < 2 * BN; i += 128) {
            const int which = i / BN, col = i - which * BN;
            if (n0 + col < p.Nout && mt0 + mt < gm.mtiles) {
              const float s = ((stat_s[(0 * 2 + which) * BN + col] + stat_s[(1 * 2 + which) * BN + col]) +
                               stat_s[(2 * 2 + which) * BN + col]) + stat_s[(3 * 2 + which) * BN + col];
              p.stat_part[((size_t)(mt0 + mt) * 2 + which) * p.Nout + n0 + col] = s;
            }
          }
          asm volatile("bar.sync 2, 128;\n" ::: "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::kTmemCols); }
  if (threadIdx.x == 0 && p.sched && atomicAdd(p.sched + 1, 1u) == gridDim.x - 1) {    // every CTA has taken its last (exhausted) index
    p.sched[0] = 0; p.sched[1] = 0;
    __threadfence();
  }
}

// Split-K finish: sum the partial tiles in split order (deterministic) and apply the epilogue.
// One thread per (pixel, 32-column chunk), the granularity of epilogue_chunk.
template <int EPI, bool ROUND_OUT>
__global__ void __launch_bounds__(128) conv_finish_kernel(ConvParams p, int ksplit) {
  const long long P = (long long)p.B * p.Hh * p.Ww;
  const int chunks = (p.Nout + 31) / 32;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * chunks) return;
  const long long pp = i / chunks;
  const int n = (int)(i - pp * chunks) * 32;
  const int ncols = min(32, p.Nout - n);
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 0.f;
  for (int s = 0; s < ksplit; ++s) {
    const float* src = p.splitk_ws + ((size_t)s * P + (size_t)pp) * p.Nout + n;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (j < ncols) {
        const float4 t = *reinterpret_cast<const float4*>(src + j);
        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
      }
    }
  }
  const int b = (int)(pp / ((long long)p.Hh * p.Ww));
  const int rem = (int)(pp - (long long)b * p.Hh * p.Ww);
  const int y = rem / p.Ww, x = rem - y * p.Ww;
  float* orow = p.out + (long long)b * p.out_sb + (long long)y * p.out_sy + (long long)x * p.out_sx;
  const long long xpix = p.x_up ? ((long long)b * (p.Hh >> 1) + (y >> 1)) * (p.Ww >> 1) + (x >> 1) : pp;
  epilogue_chunk<EPI, ROUND_OUT>(p, v, n, pp, xpix, orow);
}

// ---- host side -----------------------------------------------------------------
EncodeTiledFn tma_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static bool tc_geometry(const ConvParams& p, TcGeom* g) {
  if (p.Ww >= TC_BM) { g->Wt = TC_BM; g->Ht = 1; g->Bt = 1; }
  else {
    if (!is_pow2(p.Ww)) return false;
    g->Wt = p.Ww;
    int rows = TC_BM / g->Wt;
    if (p.Hh >= rows) { g->Ht = rows; g->Bt = 1; }
    else { if (!is_pow2(p.Hh)) return false; g->Ht = p.Hh; g->Bt = rows / p.Hh; }
  }
  g->tiles_x = ceil_div(p.Ww, g->Wt); g->tiles_y = ceil_div(p.Hh, g->Ht); g->tiles_b = ceil_div(p.B, g->Bt);
  g->mtiles = g->tiles_x * g->tiles_y * g->tiles_b; g->ntiles = 0; g->work = 0;
  return true;
}

bool conv3x3_tc_supported(const ConvParams& p, int epi) {
  TcGeom g;
  static const bool disabled = getenv("AG2V_DISABLE_TC") != nullptr;   // debugging aid: route everything to mma.sync
  if (disabled) return false;
  if (p.Cin % TC_BK != 0 || p.Cin < TC_BK) return false;
  if (p.Nout % 32 != 0) return false;
  if (epi == EPI_SPADE && p.Nout != 2 * p.C) return false;
  if (!tc_geometry(p, &g)) return false;
  if (p.in_sx % 4 || p.in_sy % 4 || p.in_sb % 4) return false;
  if (p.out_sx % 4 || p.out_sy % 4 || p.out_sb % 4) return false;
  return tma_encode_fn() != nullptr;
}

// Geometry of the fused BN statistics: M tiles of the launch and how many consecutive tiles make up one statistics
// group (a group = B / groups consecutive images); false when a tile would straddle two groups.
bool conv3x3_tc_stat_geometry(const ConvParams& p, int groups, int* mtiles, int* tiles_per_group) {
  TcGeom g;
  if (groups < 1 || p.B % groups || !tc_geometry(p, &g)) return false;
  const int Bg = p.B / groups;
  if (Bg % g.Bt) return false;
  *mtiles = g.mtiles;
  *tiles_per_group = (Bg / g.Bt) * g.tiles_y * g.tiles_x;
  return true;
}

// sums[g][0][c] = sum over the group's tiles of part[tile][0][c] (and [1] = squares), in double, fixed order:
// warp w of a block adds tiles w, w + 8, ... for 32 channels, the 8 warps are combined in warp order.
__global__ void __launch_bounds__(256) conv_stats_reduce_kernel(const float* __restrict__ part, int tiles_per_group, int Nout,
                                                                double* __restrict__ sums) {
  __shared__ double red[8][2][32];
  const int g = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s1 = 0.0, s2 = 0.0;
  if (c < Nout) {
    for (int t = warp; t < tiles_per_group; t += 8) {
      const float* src = part + ((size_t)(g * tiles_per_group + t) * 2) * Nout + c;
      s1 += (double)src[0];
      s2 += (double)src[Nout];
    }
  }
  red[warp][0][lane] = s1; red[warp][1][lane] = s2;
  __syncthreads();
  if (warp == 0 && c < Nout) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += red[w][0][lane]; b += red[w][1][lane]; }
    sums[((size_t)g * 2 + 0) * Nout + c] = a;
    sums[((size_t)g * 2 + 1) * Nout + c] = b;
  }
}

int conv_stats_reduce(const float* part, int tiles_per_group, int Nout, int groups, double* sums, cudaStream_t stream) {
  conv_stats_reduce_kernel<<<dim3(ceil_div(Nout, 32), groups), 256, 0, stream>>>(part, tiles_per_group, Nout, sums);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// Split the reduction (9 taps x Cin/32 chunks) over CTAs when the output tiles alone cannot fill
// the chip: low-resolution layers have 1-16 tiles but up to 576 reduction steps.
static int choose_ksplit(const ConvParams& p, int tiles) {
  const int total = 9 * (p.Cin / TC_BK);
  if (p.splitk_ws == nullptr || tiles * 2 > sm_count() || total < 24) return 1;
  int ks = ceil_div(sm_count(), tiles);
  if (ks > total / 8) ks = total / 8;                       // at least 8 reduction steps per split
  const size_t per = (size_t)p.B * p.Hh * p.Ww * p.Nout;
  while (ks > 1 && (size_t)ks * per > p.splitk_ws_floats) --ks;
  return ks < 1 ? 1 : ks;
}

template <int EPI, bool RO>
static int launch_finish(const ConvParams& p, int ksplit, cudaStream_t stream) {
  const long long items = (long long)p.B * p.Hh * p.Ww * ((p.Nout + 31) / 32);
  conv_finish_kernel<EPI, RO><<<(unsigned)ceil_div_ll(items, 128), 128, 0, stream>>>(p, ksplit);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// Work counters of the dynamic tile scheduler: {next index, CTAs done} pairs, zero between launches (the kernel resets
// its pair).  Launches take the pairs round robin, so two launches that might overlap on different streams never share
// one (a pair comes round again 256 launches later).  One pool per device, allocated on first use.
static unsigned* sched_pair() {
  constexpr int kPairs = 256, kMaxDev = 64;
  static unsigned* pool[kMaxDev] = {};
  static unsigned next[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  if (!pool[dev]) {
    unsigned* q = nullptr;
    if (cudaMalloc(&q, kPairs * 2 * sizeof(unsigned)) != cudaSuccess) return nullptr;
    if (cudaMemset(q, 0, kPairs * 2 * sizeof(unsigned)) != cudaSuccess) return nullptr;
    cudaDeviceSynchronize();
    pool[dev] = q;
  }
  return pool[dev] + 2 * (next[dev]++ % kPairs);
}

template <int MT, int BN, int EPI, bool RO>
static int launch_tc(const ConvParams& p0, const TcGeom& g0, cudaStream_t stream) {
  ConvParams p = p0;
  static const bool static_schedule = getenv("AG2V_TC_STATIC") && getenv("AG2V_TC_STATIC")[0] == '1';   // ablation only
  p.sched = static_schedule ? nullptr : sched_pair();
  AG2V_REQUIRE(static_schedule || p.sched, "conv3x3_tc: cannot allocate the work counters");
  TcGeom g = g0;
  g.ntiles = ceil_div(p.Nout, BN);
  const int tiles = ceil_div(g.mtiles, MT) * g.ntiles;
  const int total = 9 * (p.Cin / TC_BK);
  g.ksplit = (EPI == EPI_RAW) ? choose_ksplit(p, tiles) : 1;
  g.its_per_split = ceil_div(total, g.ksplit);
  g.ksplit = ceil_div(total, g.its_per_split);              // no empty splits
  g.work = tiles * g.ksplit;
  EncodeTiledFn enc = tma_encode_fn();
  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Ww, (cuuint64_t)p.Hh, (cuuint64_t)p.B};
    cuuint64_t strides[3] = {(cuuint64_t)p.in_sx * 4, (cuuint64_t)p.in_sy * 4, (cuuint64_t)p.in_sb * 4};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)g.Wt, (cuuint32_t)g.Ht, (cuuint32_t)g.Bt};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)p.in, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AG2V_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed with %d (Cin=%d %dx%dx%d strides %lld %lld %lld)", (int)r,
                                       p.Cin, p.B, p.Hh, p.Ww, p.in_sb, p.in_sy, p.in_sx);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.Cin, (cuuint64_t)p.Nout, 9};
    cuuint64_t strides[2] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Cin * p.Nout * 4};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)p.wpk, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AG2V_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  const int smem = TcCfg<MT, BN>::kBytes;
  AG2V_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<MT, BN, EPI, RO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = g.work < sm_count() ? g.work : sm_count();
  conv3x3_tc_kernel<MT, BN, EPI, RO><<<grid, TC_THREADS, smem, stream>>>(map_a, map_b, p, g);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

template <int MT, int BN>
static int dispatch_tc(const ConvParams& p, const TcGeom& g, int epi, int round_out, cudaStream_t stream) {
  const int tiles = ceil_div(g.mtiles, MT) * ceil_div(p.Nout, BN);
  const int ks = choose_ksplit(p, tiles);
  if (ks > 1) {                                               // partial tiles, then the finishing pass
    int rc = launch_tc<MT, BN, EPI_RAW, false>(p, g, stream);
    if (rc) return rc;
    const int total = 9 * (p.Cin / TC_BK);
    const int ksplit = ceil_div(total, ceil_div(total, ks));
    switch (epi) {
      case EPI_BIAS: return round_out ? launch_finish<EPI_BIAS, true>(p, ksplit, stream) : launch_finish<EPI_BIAS, false>(p, ksplit, stream);
      case EPI_BIAS_RELU: return round_out ? launch_finish<EPI_BIAS_RELU, true>(p, ksplit, stream) : launch_finish<EPI_BIAS_RELU, false>(p, ksplit, stream);
      case EPI_SPADE: return round_out ? launch_finish<EPI_SPADE, true>(p, ksplit, stream) : launch_finish<EPI_SPADE, false>(p, ksplit, stream);
      case EPI_GATE: return round_out ? launch_finish<EPI_GATE, true>(p, ksplit, stream) : launch_finish<EPI_GATE, false>(p, ksplit, stream);
      case EPI_ACCUM: return launch_finish<EPI_ACCUM, false>(p, ksplit, stream);
    }
  }
  switch (epi) {
    case EPI_BIAS: return round_out ? launch_tc<MT, BN, EPI_BIAS, true>(p, g, stream) : launch_tc<MT, BN, EPI_BIAS, false>(p, g, stream);
    case EPI_BIAS_RELU: return round_out ? launch_tc<MT, BN, EPI_BIAS_RELU, true>(p, g, stream) : launch_tc<MT, BN, EPI_BIAS_RELU, false>(p, g, stream);
    case EPI_SPADE: return round_out ? launch_tc<MT, BN, EPI_SPADE, true>(p, g, stream) : launch_tc<MT, BN, EPI_SPADE, false>(p, g, stream);
    case EPI_GATE: return round_out ? launch_tc<MT, BN, EPI_GATE, true>(p, g, stream) : launch_tc<MT, BN, EPI_GATE, false>(p, g, stream);
    case EPI_ACCUM: return launch_tc<MT, BN, EPI_ACCUM, false>(p, g, stream);
  }
  return fail(AG2V_ERR_ARG, "conv3x3_tc: unknown epilogue %d", epi);
}

int conv3x3_tc(const ConvParams& p, int epi, int round_out, cudaStream_t stream) {
  TcGeom g;
  if (!tc_geometry(p, &g)) return fail(AG2V_ERR_UNSUPPORTED, "conv3x3_tc: unsupported spatial shape %dx%d", p.Hh, p.Ww);
  if (((uintptr_t)p.in & 15) || ((uintptr_t)p.wpk & 15)) return fail(AG2V_ERR_ARG, "conv3x3_tc: operands must be 16-byte aligned");
  // tile shape: widest columns the problem has, two M tiles per CTA once that still fills the chip
  const char* force = getenv("AG2V_TC_TILE");                 // e.g. "2x256": force a tile shape (tests / ablation)
  // 256 columns when the problem has them; otherwise two M tiles per work item once that still
  // leaves every SM at least two work items (TMEM holds 2 stages x 256 columns either way)
  int bn = p.Nout >= 256 ? 256 : 128;
  int mt = (bn == 128 && ceil_div(g.mtiles, 2) * ceil_div(p.Nout, 128) >= 2 * sm_count()) ? 2 : 1;
  if (force) { mt = force[0] == '2' ? 2 : 1; bn = (mt == 1 && strstr(force, "256")) ? 256 : 128; }
  if (mt == 2) return dispatch_tc<2, 128>(p, g, epi, round_out, stream);
  if (bn == 256) return dispatch_tc<1, 256>(p, g, epi, round_out, stream);
  return dispatch_tc<1, 128>(p, g, epi, round_out, stream);
}

}  // namespace ag2v
