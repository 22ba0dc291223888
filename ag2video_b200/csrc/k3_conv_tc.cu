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
// * 3/4-stage mbarrier ring.  Tile shape per launch: 1 or 2 M tiles x 128 or 256
//   columns per CTA (TMEM: 128..512 columns), chosen on the host; the smallest shape
//   runs two CTAs per SM so one CTA's epilogue overlaps the other's main loop.
#include <stdlib.h>
#include <string.h>
#include "k3_common.cuh"
#include "tc_ptx.cuh"

namespace ag2v {

constexpr int TC_BM = 128, TC_BK = 32;
constexpr int TC_THREADS = 192;

struct TcGeom { int Wt, Ht, Bt, tiles_x, tiles_y, tiles_b; };

// MT = M tiles (128 pixels each) per CTA, BN = output columns per CTA.  L2 -> smem
// operand traffic per MMA cycle is (MT*16 KB + BN*128 B) / (MT * BN * 2 cycles):
//   MT=1,BN=128: 128 B/cycle   MT=2,BN=128 or MT=1,BN=256: 96 B/cycle   MT=2,BN=256: 64 B/cycle
// which is what decides the tensor-pipe utilisation of this kernel.
template <int MT, int BN>
struct TcCfg {
  static constexpr int kA = TC_BM * TC_BK * 4;                 // 16 KB per M tile
  static constexpr int kB = BN * TC_BK * 4;                    // 16 / 32 KB
  static constexpr int kStage = MT * kA + kB;
  static constexpr int kStages = (MT == 1 && BN == 128) ? 3 : (kStage <= 48 * 1024 ? 4 : 3);
  static constexpr int kBytes = kStages * kStage + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = MT * BN;                    // 128, 256 or 512
  static constexpr int kMinBlocks = (MT == 1 && BN == 128) ? 2 : 1;
};

template <int MT, int BN, int EPI, bool ROUND_OUT>
__global__ void __launch_bounds__(TC_THREADS, TcCfg<MT, BN>::kMinBlocks)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  ConvParams p, TcGeom gm) {
  using Cfg = TcCfg<MT, BN>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                         // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t bars = base + S * Cfg::kStage;                         // full[S], empty[S], tmem_full, tmem slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tc_smem_raw + (bars - raw) + 8 * (2 * S + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates of the MT consecutive tiles this CTA owns
  int x0[MT], y0[MT], b0[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    int tile = blockIdx.x * MT + mt;
    const int tx = tile % gm.tiles_x; tile /= gm.tiles_x;
    const int ty = tile % gm.tiles_y; tile /= gm.tiles_y;
    x0[mt] = tx * gm.Wt; y0[mt] = ty * gm.Ht; b0[mt] = tile * gm.Bt;   // tile past the end => b0 >= B: all rows masked
  }
  const int n0 = blockIdx.y * BN;
  const int KC = p.Cin / TC_BK;
  const int total = 9 * KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (S + s), 1); }
    mbar_init(bars + 8 * (2 * S), 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_a); prefetch_tmap(&map_b); }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < total; ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        mbar_wait(bars + 8 * (S + s), ph ^ 1u);                        // slot free
        const int tap = it / KC, kc = it - tap * KC;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const uint32_t a_dst = base + s * Cfg::kStage, b_dst = a_dst + MT * Cfg::kA;
        mbar_expect_tx(bars + 8 * s, Cfg::kStage);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
          tma_load_4d(a_dst + mt * Cfg::kA, &map_a, bars + 8 * s, kc * TC_BK, x0[mt] + dx, y0[mt] + dy, b0[mt]);
        tma_load_3d(b_dst, &map_b, bars + 8 * s, kc * TC_BK, n0, tap);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N=BN, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int it = 0; it < total; ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        mbar_wait(bars + 8 * s, ph);                                    // TMA bytes landed
        tc_fence_after();
        const uint32_t a_s = base + s * Cfg::kStage, b_s = a_s + MT * Cfg::kA;
        const uint64_t db = make_sw128_kmajor_desc(b_s);
#pragma unroll
        for (int k = 0; k < TC_BK / 8; ++k) {                           // +32 bytes along K inside the swizzle atom
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t da = make_sw128_kmajor_desc(a_s + mt * Cfg::kA);
            umma_tf32(tmem_acc + (uint32_t)(mt * BN), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(bars + 8 * (S + s));                                // frees the smem slot when the MMAs retire
      }
      umma_commit(bars + 8 * (2 * S));                                  // accumulators complete
    }
  } else {
    // ---- epilogue: warps 2..5, TMEM lane quarter = warp % 4 --------------------
    const int q = warp & 3;
    const int m = q * 32 + lane;                                       // accumulator row = pixel within the tile
    const int xt = m % gm.Wt, yt = (m / gm.Wt) % gm.Ht, bt = m / (gm.Wt * gm.Ht);
    mbar_wait(bars + 8 * (2 * S), 0);
    tc_fence_after();
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
    const int x = x0[mt] + xt, y = y0[mt] + yt, b = b0[mt] + bt;
    const bool valid = x < p.Ww && y < p.Hh && b < p.B;
    const long long pp = ((long long)b * p.Hh + y) * p.Ww + x;
    float* orow = p.out + (long long)b * p.out_sb + (long long)y * p.out_sy + (long long)x * p.out_sx;
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * BN + ch * 32), v);
      const int n = n0 + ch * 32;
      if (!valid || n >= p.Nout) continue;
      if (EPI == EPI_SPADE) {
        const int c = n >> 1;                                           // 16 channels: [g8 | b8 | g8 | b8]
        const size_t off = (size_t)pp * p.C + c;
        float xs[16], mu[16], rs[16], o[16], gm_[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          *reinterpret_cast<float4*>(xs + 4 * i) = *reinterpret_cast<const float4*>(p.x + off + 4 * i);
          *reinterpret_cast<float4*>(mu + 4 * i) = *reinterpret_cast<const float4*>(p.mean + c + 4 * i);
          *reinterpret_cast<float4*>(rs + 4 * i) = *reinterpret_cast<const float4*>(p.rstd + c + 4 * i);
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
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < ncols) v[j] += p.bias[n + j];
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_acc, Cfg::kTmemCols); }
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

template <int MT, int BN, int EPI, bool RO>
static int launch_tc(const ConvParams& p, const TcGeom& g, cudaStream_t stream) {
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
  dim3 grid(ceil_div(g.tiles_x * g.tiles_y * g.tiles_b, MT), ceil_div(p.Nout, BN));
  conv3x3_tc_kernel<MT, BN, EPI, RO><<<grid, TC_THREADS, smem, stream>>>(map_a, map_b, p, g);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

template <int MT, int BN>
static int dispatch_tc(const ConvParams& p, const TcGeom& g, int epi, int round_out, cudaStream_t stream) {
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
  const int mtiles = g.tiles_x * g.tiles_y * g.tiles_b;
  int bn = p.Nout >= 256 ? 256 : 128;
  int mt = (long long)ceil_div(mtiles, 2) * ceil_div(p.Nout, bn) >= sm_count() ? 2 : 1;
  if (force) { mt = force[0] == '2' ? 2 : 1; bn = strstr(force, "256") ? 256 : 128; }
  if (mt == 2 && bn == 256) return dispatch_tc<2, 256>(p, g, epi, round_out, stream);
  if (mt == 2) return dispatch_tc<2, 128>(p, g, epi, round_out, stream);
  if (bn == 256) return dispatch_tc<1, 256>(p, g, epi, round_out, stream);
  return dispatch_tc<1, 128>(p, g, epi, round_out, stream);
}

}  // namespace ag2v
