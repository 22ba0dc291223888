// K3 — weight gradient of the 3x3 convolutions (NHWC fp32, mma.sync TF32):
//
//   part[split][tap][n][c] = sum_{p in split} dY[p, n] * X[p + tap, c]
//
// i.e. per tap a GEMM with M = Nout, N = Cin and the huge reduction K = B*r*r
// pixels, split over CTAs (deterministic: partials are summed in split order by
// unpack_dw3x3).  CTA tile 128 x 128, 32 pixels per step, 3-stage cp.async ring.
// Both operands are "MN-major" (the reduction index is the row), so the shared
// tiles are [32 pixels][128 + 8 pad]: the pad makes the fragment reads
// conflict-free.
#include "k3_common.cuh"

namespace ag2v {

constexpr int WG_T = 128, WG_K = 32, WG_LD = 136, WG_STAGES = 3;
constexpr int kWgThreads = 256;
constexpr int kWgStageFloats = 2 * WG_K * WG_LD;

__device__ __forceinline__ void wg_cp16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ uint32_t wg_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void wg_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

struct WgradParams {
  const float* dy; int Nout;                                   // [P, Nout] contiguous
  const float* x; long long x_sb, x_sy, x_sx; int Cin;         // input view [B, Hh, Ww, Cin]
  int B, Hh, Ww;
  float* part;                                                 // [nsplit][9][Nout][Cin]
  int nsplit, chunks_per_split, ctiles;
};

template <bool PRECISE>
__global__ void __launch_bounds__(kWgThreads, PRECISE ? 1 : 2) wgrad3x3_mma_kernel(WgradParams p) {
  extern __shared__ __align__(16) float wg_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const int ntile = blockIdx.x / p.ctiles, ctile = blockIdx.x - ntile * p.ctiles;
  const int n0 = ntile * WG_T, c0 = ctile * WG_T;
  const int tap = blockIdx.y, split = blockIdx.z;
  const int dy = tap / 3 - 1, dx = tap % 3 - 1;
  const long long P = (long long)p.B * p.Hh * p.Ww;
  const long long total_chunks = (P + WG_K - 1) / WG_K;
  const long long ch0 = (long long)split * p.chunks_per_split;
  long long ch1 = ch0 + p.chunks_per_split;
  if (ch1 > total_chunks) ch1 = total_chunks;
  const int nchunks = ch1 > ch0 ? (int)(ch1 - ch0) : 0;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(wg_smem);
  const int lrow = tid >> 5, lcol = (tid & 31) * 4;

  auto load_stage = [&](int i, int stage) {
    const long long pbase = (ch0 + i) * WG_K;
    const uint32_t a_s = smem_base + (uint32_t)stage * kWgStageFloats * 4;
    const uint32_t b_s = a_s + WG_K * WG_LD * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = lrow + 8 * j;
      const long long pp = pbase + row;
      const bool pv = pp < P;
      {
        const bool ok = pv && (n0 + lcol) < p.Nout;
        const float* src = ok ? p.dy + (size_t)pp * p.Nout + n0 + lcol : p.dy;
        wg_cp16(a_s + (uint32_t)(row * WG_LD + lcol) * 4, src, ok ? 16 : 0);
      }
      {
        bool ok = pv && (c0 + lcol) < p.Cin;
        const float* src = p.x;
        if (ok) {
          const int b = (int)(pp / ((long long)p.Hh * p.Ww));
          const int rem = (int)(pp - (long long)b * p.Hh * p.Ww);
          const int y = rem / p.Ww + dy, x = rem % p.Ww + dx;
          ok = (unsigned)y < (unsigned)p.Hh && (unsigned)x < (unsigned)p.Ww;
          if (ok) src = p.x + (long long)b * p.x_sb + (long long)y * p.x_sy + (long long)x * p.x_sx + c0 + lcol;
        }
        wg_cp16(b_s + (uint32_t)(row * WG_LD + lcol) * 4, src, ok ? 16 : 0);
      }
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f; }

#pragma unroll
  for (int s = 0; s < WG_STAGES - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    asm volatile("cp.async.commit_group;\n" ::);
  }
  for (int it = 0; it < nchunks; ++it) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(WG_STAGES - 2));
    __syncthreads();
    {
      const int nx = it + WG_STAGES - 1;
      if (nx < nchunks) load_stage(nx, nx % WG_STAGES);
      asm volatile("cp.async.commit_group;\n" ::);
    }
    const float* As = wg_smem + (size_t)(it % WG_STAGES) * kWgStageFloats;   // dY chunk [32][136]
    const float* Bs = As + WG_K * WG_LD;                                       // X  chunk [32][136]
#pragma unroll
    for (int ks = 0; ks < WG_K / 8; ++ks) {
      const int k_lo = (ks * 8 + t) * WG_LD, k_hi = (ks * 8 + t + 4) * WG_LD;
      uint32_t bf[4][2], bl[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = wn * 32 + nt * 8 + g;
        const float b0 = Bs[k_lo + c], b1 = Bs[k_hi + c];
        bf[nt][0] = wg_tf32(b0); bf[nt][1] = wg_tf32(b1);
        if (PRECISE) { bl[nt][0] = wg_tf32(b0 - __uint_as_float(bf[nt][0])); bl[nt][1] = wg_tf32(b1 - __uint_as_float(bf[nt][1])); }
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int n = wm * 64 + mt * 16 + g;
        const float av[4] = {As[k_lo + n], As[k_lo + n + 8], As[k_hi + n], As[k_hi + n + 8]};
        uint32_t af[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { af[i] = wg_tf32(av[i]); if (PRECISE) al[i] = wg_tf32(av[i] - __uint_as_float(af[i])); }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (PRECISE) { wg_mma(acc[mt][nt], al, bf[nt]); wg_mma(acc[mt][nt], af, bl[nt]); }
          wg_mma(acc[mt][nt], af, bf[nt]);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);

  float* dst = p.part + ((size_t)split * 9 + tap) * p.Nout * p.Cin;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int n = n0 + wm * 64 + mt * 16 + g + half * 8;
      if (n >= p.Nout) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = c0 + wn * 32 + nt * 8 + 2 * t;
        if (c >= p.Cin) continue;
        *reinterpret_cast<float2*>(dst + (size_t)n * p.Cin + c) = make_float2(acc[mt][nt][half * 2], acc[mt][nt][half * 2 + 1]);
      }
    }
  }
}

static void wgrad_geometry(long long P, int Nout, int Cin, int* nsplit, int* cps, int* ctiles, int* ntiles) {
  *ntiles = ceil_div(Nout, WG_T);
  *ctiles = ceil_div(Cin, WG_T);
  const long long chunks = ceil_div_ll(P, WG_K);
  const int tiles = *ntiles * *ctiles * 9;
  long long ns = ceil_div_ll(2LL * 2 * sm_count(), tiles);
  if (ns > chunks) ns = chunks;
  if (ns > 64) ns = 64;
  if (ns < 1) ns = 1;
  *cps = (int)ceil_div_ll(chunks, ns);
  *nsplit = (int)ceil_div_ll(chunks, *cps);
}

}  // namespace ag2v

namespace ag2v {
bool wgrad3x3_tc_supported(int B, int Hh, int Ww, int Nout, int Cin);
int wgrad3x3_tc_nsplit(int B, int Hh, int Ww, int Nout, int Cin);
int wgrad3x3_tc(const float* dy, int Nout, const float* x, long long x_sb, long long x_sy, long long x_sx, int Cin,
                int B, int Hh, int Ww, float* part, cudaStream_t stream);
}

using namespace ag2v;

static bool use_tc(int impl, int B, int Hh, int Ww, int Nout, int Cin) {
  return (impl == 0 || impl == 2) && wgrad3x3_tc_supported(B, Hh, Ww, Nout, Cin);
}

// Number of split-K partials (and so the workspace: nsplit * 9 * Nout * Cin floats) for
// the kernel `impl` selects (0 auto, 1 mma.sync, 2 tcgen05, 3 mma.sync 3xTF32).
extern "C" int ag2v_wgrad3x3_nsplit(int B, int Hh, int Ww, int Nout, int Cin, int impl) {
  if (use_tc(impl, B, Hh, Ww, Nout, Cin)) return wgrad3x3_tc_nsplit(B, Hh, Ww, Nout, Cin);
  int nsplit, cps, ct, nt;
  wgrad_geometry((long long)B * Hh * Ww, Nout, Cin, &nsplit, &cps, &ct, &nt);
  return nsplit;
}

// dy [B*Hh*Ww, Nout] contiguous; x is a [B, Hh, Ww, Cin] view with element strides.
// part receives nsplit partial [9][Nout][Cin] blocks (reduce with ag2v_unpack_dw3x3).
extern "C" int ag2v_wgrad3x3(const float* dy, int Nout, const float* x, long long x_sb, long long x_sy,
                             long long x_sx, int Cin, int B, int Hh, int Ww, float* part, int impl,
                             cudaStream_t stream) {
  AG2V_REQUIRE(dy && x && part, "wgrad3x3: null pointer");
  AG2V_REQUIRE(B > 0 && Hh > 0 && Ww > 0 && Nout > 0 && Cin > 0, "wgrad3x3: bad sizes");
  AG2V_REQUIRE(Nout % 4 == 0 && Cin % 4 == 0, "wgrad3x3: Nout and Cin must be multiples of 4 (Nout=%d Cin=%d)", Nout, Cin);
  AG2V_REQUIRE(x_sx % 4 == 0 && x_sy % 4 == 0 && x_sb % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0,
               "wgrad3x3: operands must be 16-byte aligned");
  if (impl == 2 && !wgrad3x3_tc_supported(B, Hh, Ww, Nout, Cin))
    return fail(AG2V_ERR_UNSUPPORTED, "wgrad3x3: shape not supported by the tcgen05 kernel (Nout=%d Cin=%d %dx%d)", Nout, Cin, Hh, Ww);
  if (use_tc(impl, B, Hh, Ww, Nout, Cin)) return wgrad3x3_tc(dy, Nout, x, x_sb, x_sy, x_sx, Cin, B, Hh, Ww, part, stream);
  const int precise = impl == 3;
  WgradParams p;
  p.dy = dy; p.Nout = Nout; p.x = x; p.x_sb = x_sb; p.x_sy = x_sy; p.x_sx = x_sx; p.Cin = Cin;
  p.B = B; p.Hh = Hh; p.Ww = Ww; p.part = part;
  int ntiles;
  wgrad_geometry((long long)B * Hh * Ww, Nout, Cin, &p.nsplit, &p.chunks_per_split, &p.ctiles, &ntiles);
  const size_t smem = (size_t)WG_STAGES * kWgStageFloats * sizeof(float);
  dim3 grid(ntiles * p.ctiles, 9, p.nsplit);
  if (precise) {
    AG2V_CUDA(cudaFuncSetAttribute(wgrad3x3_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad3x3_mma_kernel<true><<<grid, kWgThreads, smem, stream>>>(p);
  } else {
    AG2V_CUDA(cudaFuncSetAttribute(wgrad3x3_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wgrad3x3_mma_kernel<false><<<grid, kWgThreads, smem, stream>>>(p);
  }
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}
