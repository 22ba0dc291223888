// K3 — weight gradient of the 3x3 convolutions on tcgen05 tensor cores.
//
//   part[split][tap][n][c] = sum_{p in split} dY[p, n] * X[p + tap, c]
//
// Per tap this is a GEMM with M = Nout, N = Cin and the reduction over pixels.  Both
// operands have the reduction index as their ROW in memory ([pixel][channel], NHWC),
// i.e. they are "MN-major" for the tensor core; tcgen05 takes MN-major TF32 operands
// directly, so no transpose is needed: a TMA box {32 channels, 32 pixels} lands as
// 32 rows x 128 B in the "128B swizzle, 32B atom" pattern, the one MN-major layout the
// tensor core accepts for 32-bit operands (eight 4-row atoms per box).
//
// One CTA owns a 128 (n) x 128 (c) tile for the THREE taps of one kernel row
// (dy fixed, dx = -1, 0, +1): the dY chunk is loaded once and multiplied with three
// shifted X chunks into three TMEM accumulators (3 x 128 columns).  The tap shift is a
// TMA coordinate offset and the zero padding is TMA's out-of-bounds fill.  The pixel
// range is split over CTAs (split-K); partials are reduced in split order by
// unpack_dw3x3 (deterministic).
#include "k3_common.cuh"
#include "tc_ptx.cuh"

namespace ag2v {

constexpr int WT_T = 128;            // tile edge (n and c)
constexpr int WT_PIX = 32;           // pixels (reduction rows) per stage
constexpr int WT_STAGES = 3;
constexpr int WT_THREADS = 192;
constexpr int WT_BOX = WT_PIX * 128;                 // one {32 ch, 32 px} box = 4 KB
constexpr int WT_A = 4 * WT_BOX;                     // dY chunk: 4 channel groups   = 16 KB
constexpr int WT_B = 3 * 4 * WT_BOX;                 // X chunks: 3 taps x 4 groups  = 48 KB
constexpr int WT_STAGE = WT_A + WT_B;
constexpr int WT_SMEM = WT_STAGES * WT_STAGE + 1024 + 256;

struct WtGeom { int Wt, Ht, Bt, tiles_x, tiles_y, tiles_b, chunks, cps, nsplit, ntiles, ctiles; };

struct WtParams { int B, Hh, Ww, Nout, Cin; float* part; };

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                   WtParams p, WtGeom gm) {
  extern __shared__ uint8_t wt_smem_raw[];
  const uint32_t raw = smem_u32(wt_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bars = base + WT_STAGES * WT_STAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wt_smem_raw + (bars - raw) + 8 * (2 * WT_STAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = blockIdx.x / gm.ctiles, ctile = blockIdx.x - ntile * gm.ctiles;
  const int n0 = ntile * WT_T, c0 = ctile * WT_T;
  const int dy = (int)blockIdx.y - 1;
  const int split = blockIdx.z;
  const int ch0 = split * gm.cps;
  const int ch1 = min(ch0 + gm.cps, gm.chunks);
  const int nchunks = ch1 - ch0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < WT_STAGES; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (WT_STAGES + s), 1); }
    mbar_init(bars + 8 * (2 * WT_STAGES), 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&map_dy); prefetch_tmap(&map_x); }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < nchunks; ++it) {
        const int s = it % WT_STAGES;
        const uint32_t ph = (uint32_t)(it / WT_STAGES) & 1u;
        mbar_wait(bars + 8 * (WT_STAGES + s), ph ^ 1u);
        int chunk = ch0 + it;
        const int tx = chunk % gm.tiles_x; chunk /= gm.tiles_x;
        const int ty = chunk % gm.tiles_y; chunk /= gm.tiles_y;
        const int x0 = tx * gm.Wt, y0 = ty * gm.Ht, b0 = chunk * gm.Bt;
        const uint32_t a_dst = base + s * WT_STAGE, b_dst = a_dst + WT_A;
        mbar_expect_tx(bars + 8 * s, WT_STAGE);
#pragma unroll
        for (int i = 0; i < 4; ++i) tma_load_4d(a_dst + i * WT_BOX, &map_dy, bars + 8 * s, n0 + 32 * i, x0, y0, b0);
#pragma unroll
        for (int t = 0; t < 3; ++t)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            tma_load_4d(b_dst + (t * 4 + j) * WT_BOX, &map_x, bars + 8 * s, c0 + 32 * j, x0 + t - 1, y0 + dy, b0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = f32, A = B = tf32, both MN-major (bits 15, 16), N = 128, M = 128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(WT_T >> 3) << 17) | ((uint32_t)(WT_T >> 4) << 24);
      for (int it = 0; it < nchunks; ++it) {
        const int s = it % WT_STAGES;
        const uint32_t ph = (uint32_t)(it / WT_STAGES) & 1u;
        mbar_wait(bars + 8 * s, ph);
        tc_fence_after();
        const uint32_t a_s = base + s * WT_STAGE, b_s = a_s + WT_A;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
#pragma unroll
          for (int kk = 0; kk < WT_PIX / 8; ++kk) {        // 8 pixels = two 4-row atoms (1024 B) per instruction
            const uint64_t da = make_sw128b32_mnmajor_desc(a_s + kk * 1024, WT_BOX, 512);
            const uint64_t db = make_sw128b32_mnmajor_desc(b_s + t * 4 * WT_BOX + kk * 1024, WT_BOX, 512);
            umma_tf32(tmem_acc + (uint32_t)(t * WT_T), da, db, idesc, (it > 0 || kk > 0) ? 1u : 0u);
          }
        }
        umma_commit(bars + 8 * (WT_STAGES + s));
      }
      umma_commit(bars + 8 * (2 * WT_STAGES));
    }
  } else {
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    mbar_wait(bars + 8 * (2 * WT_STAGES), 0);
    tc_fence_after();
#pragma unroll 1
    for (int t = 0; t < 3; ++t) {
      const int tap = (dy + 1) * 3 + t;
      float* dst = p.part + (((size_t)split * 9 + tap) * p.Nout + n) * p.Cin + c0;
#pragma unroll 1
      for (int ch = 0; ch < WT_T / 32; ++ch) {
        float v[32];
        tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * WT_T + ch * 32), v);
        if (n >= p.Nout || c0 + ch * 32 >= p.Cin) continue;
        if (nchunks <= 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + ch * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_acc, 512); }
}

static bool wt_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static bool wt_geometry(int B, int Hh, int Ww, int Nout, int Cin, WtGeom* g) {
  if (Nout % 32 || Cin % 32) return false;
  if (Ww >= WT_PIX) { g->Wt = WT_PIX; g->Ht = 1; g->Bt = 1; }
  else {
    if (!wt_pow2(Ww)) return false;
    g->Wt = Ww;
    const int rows = WT_PIX / Ww;
    if (Hh >= rows) { g->Ht = rows; g->Bt = 1; }
    else { if (!wt_pow2(Hh)) return false; g->Ht = Hh; g->Bt = rows / Hh; }
  }
  g->tiles_x = ceil_div(Ww, g->Wt); g->tiles_y = ceil_div(Hh, g->Ht); g->tiles_b = ceil_div(B, g->Bt);
  g->chunks = g->tiles_x * g->tiles_y * g->tiles_b;
  g->ntiles = ceil_div(Nout, WT_T); g->ctiles = ceil_div(Cin, WT_T);
  const int groups = g->ntiles * g->ctiles * 3;
  int ns = sm_count() / groups;          // floor: one CTA per SM, a single wave (ceil would leave a 2nd wave of stragglers)
  if (ns > g->chunks) ns = g->chunks;
  if (ns > 64) ns = 64;
  if (ns < 1) ns = 1;
  g->cps = ceil_div(g->chunks, ns);
  g->nsplit = ceil_div(g->chunks, g->cps);
  return true;
}

bool wgrad3x3_tc_supported(int B, int Hh, int Ww, int Nout, int Cin) {
  static const bool disabled = getenv("AG2V_DISABLE_TC") != nullptr;
  WtGeom g;
  return !disabled && wt_geometry(B, Hh, Ww, Nout, Cin, &g) && tma_encode_fn() != nullptr;
}

int wgrad3x3_tc_nsplit(int B, int Hh, int Ww, int Nout, int Cin) {
  WtGeom g;
  return wt_geometry(B, Hh, Ww, Nout, Cin, &g) ? g.nsplit : -1;
}

int wgrad3x3_tc(const float* dy, int Nout, const float* x, long long x_sb, long long x_sy, long long x_sx, int Cin,
                int B, int Hh, int Ww, float* part, cudaStream_t stream) {
  WtGeom g;
  if (!wt_geometry(B, Hh, Ww, Nout, Cin, &g)) return fail(AG2V_ERR_UNSUPPORTED, "wgrad3x3_tc: unsupported shape");
  EncodeTiledFn enc = tma_encode_fn();
  CUtensorMap map_dy, map_x;
  cuuint32_t box[4] = {32, (cuuint32_t)g.Wt, (cuuint32_t)g.Ht, (cuuint32_t)g.Bt};
  cuuint32_t es[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {(cuuint64_t)Nout, (cuuint64_t)Ww, (cuuint64_t)Hh, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Nout * 4, (cuuint64_t)Nout * Ww * 4, (cuuint64_t)Nout * Ww * Hh * 4};
    CUresult r = enc(&map_dy, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)dy, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AG2V_ERR_CUDA, "cuTensorMapEncodeTiled(dY) failed with %d", (int)r);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Ww, (cuuint64_t)Hh, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)x_sx * 4, (cuuint64_t)x_sy * 4, (cuuint64_t)x_sb * 4};
    CUresult r = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(AG2V_ERR_CUDA, "cuTensorMapEncodeTiled(X) failed with %d", (int)r);
  }
  WtParams p{B, Hh, Ww, Nout, Cin, part};
  AG2V_CUDA(cudaFuncSetAttribute(wgrad3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WT_SMEM));
  dim3 grid(g.ntiles * g.ctiles, 3, g.nsplit);
  wgrad3x3_tc_kernel<<<grid, WT_THREADS, WT_SMEM, stream>>>(map_dy, map_x, p, g);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

}  // namespace ag2v
