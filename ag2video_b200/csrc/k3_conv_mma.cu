// K3 — 3x3 implicit-GEMM convolution on NHWC fp32, warp-level mma.sync (TF32,
// fp32 accumulate) with a cp.async pipeline.  This is the generic path: any
// Cin % 4 == 0, any Nout, any spatial size, strided input views (the nearest
// down-sample of the segmap, normalization.py:102, is just a strided view), and it
// is the cross-check for the tcgen05 kernel in k3_conv_tc.cu, which takes over
// the shapes it supports.
//
//   out[p, n] = epilogue( sum_{tap, c} in[p + tap, c] * wpk[tap][n][c] )
//
// CTA tile 128 pixels x 128 columns, K step 32 channels of one tap, 3 stages.
// Shared tiles are [row][32 floats] with the 16-byte chunk index XOR-ed by
// (row & 7): conflict-free for both the cp.async writes and the fragment reads.
#include "k3_common.cuh"

namespace ag2v {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3;
constexpr int kConvThreads = 256;
constexpr int kStageFloats = (BM + BN) * BK;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float round_tf32(float x) { return __uint_as_float(to_tf32(x)); }

__device__ __forceinline__ void mma_16n8k8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ int swz(int row, int chunk, int off) { return row * BK + (((chunk ^ (row & 7)) << 2) | off); }

// PRECISE: every product is split hi/lo (3xTF32, fp32-class accuracy).  Validation mode:
// it separates logic errors from TF32 rounding in end-to-end comparisons.
template <int EPI, bool ROUND_OUT, bool PRECISE>
__global__ void __launch_bounds__(kConvThreads, PRECISE ? 1 : 2) conv3x3_mma_kernel(ConvParams p) {
  extern __shared__ __align__(16) float conv_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;
  const long long P = (long long)p.B * p.Hh * p.Ww;
  const long long p0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int KC = (p.Cin + BK - 1) / BK;
  const int total = 9 * KC;

  // per-thread load geometry: 4 rows (tid>>3 + 32 j), one 16-byte chunk column (tid & 7)
  const int cj = tid & 7;
  int ry[4], rx[4];
  long long rbase[4];
  bool rvalid[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long pp = p0 + (tid >> 3) + 32 * j;
    rvalid[j] = pp < P;
    const long long q = rvalid[j] ? pp : 0;
    const int b = (int)(q / ((long long)p.Hh * p.Ww));
    const int rem = (int)(q - (long long)b * p.Hh * p.Ww);
    ry[j] = rem / p.Ww; rx[j] = rem - ry[j] * p.Ww;
    rbase[j] = (long long)b * p.in_sb + (long long)ry[j] * p.in_sy + (long long)rx[j] * p.in_sx;
  }
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(conv_smem);

  auto load_stage = [&](int it, int stage) {
    const int tap = it / KC, kc = it - tap * KC;
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int c = kc * BK + cj * 4;
    const bool cvalid = c < p.Cin;
    const uint32_t a_s = smem_base + (uint32_t)stage * kStageFloats * 4;
    const uint32_t b_s = a_s + BM * BK * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = (tid >> 3) + 32 * j;
      const bool ok = cvalid && rvalid[j] && (unsigned)(ry[j] + dy) < (unsigned)p.Hh && (unsigned)(rx[j] + dx) < (unsigned)p.Ww;
      const float* src = ok ? p.in + rbase[j] + (long long)dy * p.in_sy + (long long)dx * p.in_sx + c : p.in;
      cp_async16(a_s + (uint32_t)swz(row, cj, 0) * 4, src, ok ? 16 : 0);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = (tid >> 3) + 32 * j;
      const bool ok = cvalid && (n0 + row) < p.Nout;
      const float* src = ok ? p.wpk + ((size_t)tap * p.Nout + n0 + row) * p.Cin + c : p.wpk;
      cp_async16(b_s + (uint32_t)swz(row, cj, 0) * 4, src, ok ? 16 : 0);
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f; }

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < total) load_stage(s, s);
    cp_async_commit();
  }
  for (int it = 0; it < total; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nx = it + STAGES - 1;
      if (nx < total) load_stage(nx, nx % STAGES);
      cp_async_commit();
    }
    const float* As = conv_smem + (size_t)(it % STAGES) * kStageFloats;
    const float* Bs = As + BM * BK;
#pragma unroll
    for (int ks = 0; ks < BK / 8; ++ks) {
      uint32_t bf[4][2], bl[4][2];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = wn * 32 + nt * 8 + g;
        const float b0 = Bs[swz(n, 2 * ks, t)], b1 = Bs[swz(n, 2 * ks + 1, t)];
        bf[nt][0] = to_tf32(b0); bf[nt][1] = to_tf32(b1);
        if (PRECISE) { bl[nt][0] = to_tf32(b0 - __uint_as_float(bf[nt][0])); bl[nt][1] = to_tf32(b1 - __uint_as_float(bf[nt][1])); }
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int r = wm * 64 + mt * 16 + g;
        float av[4];
        av[0] = As[swz(r, 2 * ks, t)]; av[1] = As[swz(r + 8, 2 * ks, t)];
        av[2] = As[swz(r, 2 * ks + 1, t)]; av[3] = As[swz(r + 8, 2 * ks + 1, t)];
        uint32_t af[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { af[i] = to_tf32(av[i]); if (PRECISE) al[i] = to_tf32(av[i] - __uint_as_float(af[i])); }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (PRECISE) { mma_16n8k8(acc[mt][nt], al, bf[nt]); mma_16n8k8(acc[mt][nt], af, bl[nt]); }
          mma_16n8k8(acc[mt][nt], af, bf[nt]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // ---- epilogue -------------------------------------------------------------
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const long long pp = p0 + wm * 64 + mt * 16 + g + half * 8;
      if (pp >= P) continue;
      const int b = (int)(pp / ((long long)p.Hh * p.Ww));
      const int rem = (int)(pp - (long long)b * p.Hh * p.Ww);
      const int y = rem / p.Ww, x = rem - y * p.Ww;
      float* orow = p.out + (long long)b * p.out_sb + (long long)y * p.out_sy + (long long)x * p.out_sx;
      if (EPI == EPI_SPADE) {
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const int ng = n0 + wn * 32 + pr * 16 + 2 * t;     // packed column of gamma; beta is ng + 8
          if (ng >= p.Nout) continue;
          const int c = ((ng >> 4) << 3) + (ng & 7);
          float g0 = acc[mt][2 * pr][half * 2 + 0], g1 = acc[mt][2 * pr][half * 2 + 1];
          float b0 = acc[mt][2 * pr + 1][half * 2 + 0], b1 = acc[mt][2 * pr + 1][half * 2 + 1];
          if (p.bias) { g0 += p.bias[ng]; g1 += p.bias[ng + 1]; b0 += p.bias[ng + 8]; b1 += p.bias[ng + 9]; }
          const size_t off = (size_t)pp * p.C + c;
          const long long xpix = p.x_up ? ((long long)b * (p.Hh >> 1) + (y >> 1)) * (p.Ww >> 1) + (x >> 1) : pp;
          const float2 xv = *reinterpret_cast<const float2*>(p.x + (size_t)xpix * p.C + c);
          const size_t sc = (p.group_pixels > 0 ? (size_t)(pp / p.group_pixels) * p.C : 0) + c;
          const float2 mu = *reinterpret_cast<const float2*>(p.mean + sc);
          const float2 rs = *reinterpret_cast<const float2*>(p.rstd + sc);
          float o0 = (xv.x - mu.x) * rs.x * (1.f + g0) + b0;
          float o1 = (xv.y - mu.y) * rs.y * (1.f + g1) + b1;
          if (p.slope != 1.f) { o0 = o0 > 0.f ? o0 : o0 * p.slope; o1 = o1 > 0.f ? o1 : o1 * p.slope; }
          if (ROUND_OUT) { o0 = round_tf32(o0); o1 = round_tf32(o1); }
          *reinterpret_cast<float2*>(orow + c) = make_float2(o0, o1);
          if (p.gamma_out) *reinterpret_cast<float2*>(p.gamma_out + off) = make_float2(g0, g1);
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int n = n0 + wn * 32 + nt * 8 + 2 * t;
          if (n >= p.Nout) continue;
          float v0 = acc[mt][nt][half * 2 + 0], v1 = acc[mt][nt][half * 2 + 1];
          if (EPI == EPI_BIAS || EPI == EPI_BIAS_RELU) {
            if (EPI == EPI_BIAS && p.scale) {
              const float sc = p.scale[p.group_pixels > 0 ? pp / p.group_pixels : 0];
              v0 *= sc; v1 *= sc;
            }
            if (p.bias) { v0 += p.bias[n]; v1 += p.bias[n + 1]; }
            if (EPI == EPI_BIAS && p.res) {
              const float2 r2 = *reinterpret_cast<const float2*>(p.res + (size_t)pp * p.Nout + n);
              v0 += r2.x; v1 += r2.y;
            }
            if (EPI == EPI_BIAS_RELU) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
          } else if (EPI == EPI_GATE) {
            const float2 gt = *reinterpret_cast<const float2*>(p.gate + (size_t)pp * p.Nout + n);
            v0 = gt.x > 0.f ? v0 : 0.f; v1 = gt.y > 0.f ? v1 : 0.f;
          } else if (EPI == EPI_ACCUM) {
            const float2 old = *reinterpret_cast<const float2*>(orow + n);
            v0 += old.x; v1 += old.y;
          }
          if (ROUND_OUT) { v0 = round_tf32(v0); v1 = round_tf32(v1); }
          *reinterpret_cast<float2*>(orow + n) = make_float2(v0, v1);
        }
      }
    }
  }
}

int conv3x3_check(const ConvParams& p, int epi) {
  AG2V_REQUIRE(p.in && p.wpk && p.out, "conv3x3: null pointer");
  AG2V_REQUIRE(p.B > 0 && p.Hh > 0 && p.Ww > 0 && p.Cin > 0 && p.Nout > 0, "conv3x3: bad sizes");
  AG2V_REQUIRE(p.Cin % 4 == 0 && p.Nout % 2 == 0, "conv3x3: Cin %% 4 == 0 and even Nout required (Cin=%d Nout=%d)", p.Cin, p.Nout);
  AG2V_REQUIRE(p.in_sx % 4 == 0 && p.in_sy % 4 == 0 && p.in_sb % 4 == 0 && ((uintptr_t)p.in & 15) == 0, "conv3x3: input view must be 16-byte aligned");
  AG2V_REQUIRE(p.out_sx % 2 == 0 && p.out_sy % 2 == 0 && p.out_sb % 2 == 0, "conv3x3: output view must be 8-byte aligned");
  if (epi == EPI_SPADE) {
    AG2V_REQUIRE(p.x && p.mean && p.rstd && p.C > 0 && p.Nout == 2 * p.C && p.C % 8 == 0, "conv3x3: SPADE epilogue needs x/mean/rstd and Nout == 2C, C %% 8 == 0");
    AG2V_REQUIRE(!p.x_up || (p.Hh % 2 == 0 && p.Ww % 2 == 0), "conv3x3: up-sampled x needs even output sizes");
  }
  if (epi == EPI_GATE) AG2V_REQUIRE(p.gate, "conv3x3: gate epilogue needs a gate tensor");
  return AG2V_OK;
}

template <int EPI, bool RO, bool PR>
static int launch_mma(const ConvParams& p, cudaStream_t stream) {
  const long long P = (long long)p.B * p.Hh * p.Ww;
  dim3 grid((unsigned)ceil_div_ll(P, BM), (unsigned)ceil_div(p.Nout, BN));
  const size_t smem = (size_t)STAGES * kStageFloats * sizeof(float);
  AG2V_CUDA(cudaFuncSetAttribute(conv3x3_mma_kernel<EPI, RO, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv3x3_mma_kernel<EPI, RO, PR><<<grid, kConvThreads, smem, stream>>>(p);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// precise != 0: 3xTF32 products and no output rounding (validation mode)
int conv3x3_mma(const ConvParams& p, int epi, int round_out, int precise, cudaStream_t stream) {
  int rc = conv3x3_check(p, epi);
  if (rc) return rc;
  if (precise) {
    switch (epi) {
      case EPI_BIAS: return launch_mma<EPI_BIAS, false, true>(p, stream);
      case EPI_BIAS_RELU: return launch_mma<EPI_BIAS_RELU, false, true>(p, stream);
      case EPI_SPADE: return launch_mma<EPI_SPADE, false, true>(p, stream);
      case EPI_GATE: return launch_mma<EPI_GATE, false, true>(p, stream);
      case EPI_ACCUM: return launch_mma<EPI_ACCUM, false, true>(p, stream);
    }
  }
  switch (epi) {
    case EPI_BIAS: return round_out ? launch_mma<EPI_BIAS, true, false>(p, stream) : launch_mma<EPI_BIAS, false, false>(p, stream);
    case EPI_BIAS_RELU: return round_out ? launch_mma<EPI_BIAS_RELU, true, false>(p, stream) : launch_mma<EPI_BIAS_RELU, false, false>(p, stream);
    case EPI_SPADE: return round_out ? launch_mma<EPI_SPADE, true, false>(p, stream) : launch_mma<EPI_SPADE, false, false>(p, stream);
    case EPI_GATE: return round_out ? launch_mma<EPI_GATE, true, false>(p, stream) : launch_mma<EPI_GATE, false, false>(p, stream);
    case EPI_ACCUM: return launch_mma<EPI_ACCUM, false, false>(p, stream);
  }
  return fail(AG2V_ERR_ARG, "conv3x3: unknown epilogue %d", epi);
}

}  // namespace ag2v
