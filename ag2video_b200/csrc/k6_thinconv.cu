// K6 — 3x3 convolutions with a handful of output channels at full resolution: the generator's
// conv_img (64 -> 3, between a LeakyReLU(0.2) and a tanh; spade_models/networks/generator.py) and
// the flow head conv_flow (32 -> 2; flows_generator.py).  With 2-3 output channels these are not
// GEMMs: per pixel 9*Ci*Co MACs against Ci*4 bytes of input, i.e. HBM/L1-bound streaming kernels.
// Library implicit-GEMM kernels pad Co up to a tensor-core tile and took ~350 us per launch here
// (forward, input gradient and weight gradient each); these take the time of reading the input.
//
// Layout: x [B,H,W,CI] NHWC, y / dy [B,H,W,CO] NHWC (channels_last tensors), weights either
// [CO][CI][3][3] or channels-last [CO][3][3][CI].  The activation in front (LeakyReLU slope_in)
// and behind (tanh) are folded in: y = act_out(conv(act_in(x)) + bias).
#include "common.cuh"

namespace ag2v {

constexpr int kThinThreads = 128;

struct ThinParams {
  const float* x; const float* w; const float* bias;
  int B, H, W; int w_cl; float slope_in; int act_out;   // act_out: 0 none, 1 tanh
  float* y;
  // backward
  const float* dy; float* dx; float* part; int nparts;
};

template <int CI, int CO>
__device__ __forceinline__ float thin_w(const ThinParams& p, int co, int tap, int ci) {
  return p.w_cl ? p.w[((size_t)co * 9 + tap) * CI + ci] : p.w[((size_t)co * CI + ci) * 9 + tap];
}

// CI/4 threads per output pixel (each owns 4 input channels: one coalesced 16-byte load per tap),
// partial sums reduced over those threads with shuffles; weights [tap][co][ci] in shared memory.
template <int CI, int CO>
__global__ void __launch_bounds__(kThinThreads) thin_conv_fwd_kernel(ThinParams p) {
  __shared__ __align__(16) float ws[9 * CO * CI];
  for (int i = threadIdx.x; i < 9 * CO * CI; i += kThinThreads) {
    const int ci = i % CI, co = (i / CI) % CO, tap = i / (CO * CI);
    ws[i] = thin_w<CI, CO>(p, co, tap, ci);
  }
  __syncthreads();
  constexpr int Q = CI / 4;                          // threads per pixel (8 or 16: a power of two <= 32)
  constexpr int PIX = kThinThreads / Q;              // pixels per CTA pass
  const int sub = threadIdx.x % Q, slot = threadIdx.x / Q;
  const long long P = (long long)p.B * p.H * p.W;
  for (long long p0 = (long long)blockIdx.x * PIX; p0 < P; p0 += (long long)gridDim.x * PIX) {
    const long long pp = p0 + slot;
    const bool live = pp < P;
    float acc[CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[co] = 0.f;
    if (live) {
      const int b = (int)(pp / ((long long)p.H * p.W));
      const int rem = (int)(pp - (long long)b * p.H * p.W);
      const int yy = rem / p.W, xx = rem - yy * p.W;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int y2 = yy + tap / 3 - 1, x2 = xx + tap % 3 - 1;
        if (y2 < 0 || y2 >= p.H || x2 < 0 || x2 >= p.W) continue;
        float4 v = __ldg(reinterpret_cast<const float4*>(p.x + (((size_t)b * p.H + y2) * p.W + x2) * CI) + sub);
        if (p.slope_in != 1.f) {
          v.x = v.x > 0.f ? v.x : v.x * p.slope_in; v.y = v.y > 0.f ? v.y : v.y * p.slope_in;
          v.z = v.z > 0.f ? v.z : v.z * p.slope_in; v.w = v.w > 0.f ? v.w : v.w * p.slope_in;
        }
#pragma unroll
        for (int co = 0; co < CO; ++co) {
          const float4 w4 = *reinterpret_cast<const float4*>(ws + (tap * CO + co) * CI + 4 * sub);
          acc[co] = fmaf(v.x, w4.x, acc[co]); acc[co] = fmaf(v.y, w4.y, acc[co]);
          acc[co] = fmaf(v.z, w4.z, acc[co]); acc[co] = fmaf(v.w, w4.w, acc[co]);
        }
      }
    }
#pragma unroll
    for (int co = 0; co < CO; ++co) {
#pragma unroll
      for (int o = Q / 2; o > 0; o >>= 1) acc[co] += __shfl_xor_sync(0xffffffffu, acc[co], o);
    }
    if (live && sub < CO) {
      float r = (sub == 0 ? acc[0] : (sub == 1 ? acc[1 % CO] : acc[2 % CO])) + (p.bias ? p.bias[sub] : 0.f);
      p.y[(size_t)pp * CO + sub] = p.act_out == 1 ? tanhf(r) : r;
    }
  }
}

// gradient at the convolution output, dyp = dy * act_out'(y): written once into the workspace
__global__ void thin_dyp_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, float* __restrict__ dyp) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float t = y[i];
    dyp[i] = dy[i] * (1.f - t * t);
  }
}

template <int CO>
__device__ __forceinline__ void thin_dyp(const ThinParams& p, size_t pix, float (&g)[CO]) {
#pragma unroll
  for (int co = 0; co < CO; ++co) g[co] = p.dy[pix * CO + co];      // p.dy already holds dyp
}

// input gradient: one thread = (pixel, 4 input channels); dx = act_in'(x) * sum_tap sum_co dyp[p - off] W[co][tap][ci]
template <int CI, int CO>
__global__ void __launch_bounds__(kThinThreads) thin_conv_dgrad_kernel(ThinParams p) {
  __shared__ float ws[9 * CO * CI];                       // [tap][co][ci]
  for (int i = threadIdx.x; i < 9 * CO * CI; i += kThinThreads) {
    const int ci = i % CI, co = (i / CI) % CO, tap = i / (CO * CI);
    ws[i] = thin_w<CI, CO>(p, co, tap, ci);
  }
  __syncthreads();
  constexpr int Q = CI / 4;
  const long long P = (long long)p.B * p.H * p.W;
  const long long gid = (long long)blockIdx.x * kThinThreads + threadIdx.x;
  if (gid >= P * Q) return;
  const long long pp = gid / Q;
  const int c = (int)(gid - pp * Q) * 4;
  const int b = (int)(pp / ((long long)p.H * p.W));
  const int rem = (int)(pp - (long long)b * p.H * p.W);
  const int yy = rem / p.W, xx = rem - yy * p.W;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    // output pixel q = p - off(tap) used input pixel p through tap `tap`
    const int y2 = yy - (tap / 3 - 1), x2 = xx - (tap % 3 - 1);
    if (y2 < 0 || y2 >= p.H || x2 < 0 || x2 >= p.W) continue;
    float g[CO];
    thin_dyp<CO>(p, ((size_t)b * p.H + y2) * p.W + x2, g);
#pragma unroll
    for (int co = 0; co < CO; ++co) {
      const float4 w4 = *reinterpret_cast<const float4*>(ws + (tap * CO + co) * CI + c);
      acc.x = fmaf(g[co], w4.x, acc.x); acc.y = fmaf(g[co], w4.y, acc.y);
      acc.z = fmaf(g[co], w4.z, acc.z); acc.w = fmaf(g[co], w4.w, acc.w);
    }
  }
  if (p.slope_in != 1.f) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)pp * CI + c));
    acc.x = xv.x > 0.f ? acc.x : acc.x * p.slope_in; acc.y = xv.y > 0.f ? acc.y : acc.y * p.slope_in;
    acc.z = xv.z > 0.f ? acc.z : acc.z * p.slope_in; acc.w = xv.w > 0.f ? acc.w : acc.w * p.slope_in;
  }
  *reinterpret_cast<float4*>(p.dx + (size_t)pp * CI + c) = acc;
}

// weight / bias gradient partials: CTA = 256 threads = (256 / CI) pixel lanes x CI channels; every
// lane walks a contiguous run of input pixels and keeps the 3x3 window of dyp around the current
// pixel in registers (3*CO new values per step).  Thread (lane, ci) keeps 9*CO sums.
//   part[cta][ (tap*CO + co)*CI + ci ]  and  part[cta][9*CO*CI + co] (bias)
template <int CI, int CO>
__global__ void __launch_bounds__(256) thin_conv_wgrad_kernel(ThinParams p) {
  constexpr int LANES = 256 / CI;
  constexpr int NV = 9 * CO * CI + CO;
  __shared__ float red[LANES][9 * CO * CI + CO];
  const int ci = threadIdx.x % CI, lane = threadIdx.x / CI;
  const long long P = (long long)p.B * p.H * p.W;
  const long long runs = (long long)gridDim.x * LANES;
  const long long per = (P + runs - 1) / runs;
  const long long q0 = ((long long)blockIdx.x * LANES + lane) * per, q1 = q0 + per < P ? q0 + per : P;
  float acc[9 * CO];
#pragma unroll
  for (int i = 0; i < 9 * CO; ++i) acc[i] = 0.f;
  float bsum[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) bsum[co] = 0.f;
  float win[3][3][CO];                     // dyp at (yy-1..yy+1, xx-1..xx+1), zero outside the image
  auto load_col = [&](int b, int yy, int xc, int slot) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int y2 = yy + r - 1;
      const bool in = y2 >= 0 && y2 < p.H && xc >= 0 && xc < p.W;
#pragma unroll
      for (int co = 0; co < CO; ++co)
        win[r][slot][co] = in ? p.dy[(((size_t)b * p.H + y2) * p.W + xc) * CO + co] : 0.f;
    }
  };
  for (long long q = q0; q < q1; ++q) {
    const int b = (int)(q / ((long long)p.H * p.W));
    const int rem = (int)(q - (long long)b * p.H * p.W);
    const int yy = rem / p.W, xx = rem - yy * p.W;
    if (q == q0 || xx == 0) {
      load_col(b, yy, xx - 1, 0); load_col(b, yy, xx, 1); load_col(b, yy, xx + 1, 2);
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int co = 0; co < CO; ++co) { win[r][0][co] = win[r][1][co]; win[r][1][co] = win[r][2][co]; }
      load_col(b, yy, xx + 1, 2);
    }
    float a = __ldg(p.x + (size_t)q * CI + ci);
    if (p.slope_in != 1.f) a = a > 0.f ? a : a * p.slope_in;
    // output pixel (yy - dy, xx - dx) read input pixel q through tap (dy+1, dx+1)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int r = 1 - (tap / 3 - 1), c = 1 - (tap % 3 - 1);
#pragma unroll
      for (int co = 0; co < CO; ++co) acc[tap * CO + co] = fmaf(win[r][c][co], a, acc[tap * CO + co]);
    }
    if (ci == 0) {
#pragma unroll
      for (int co = 0; co < CO; ++co) bsum[co] += win[1][1][co];
    }
  }
#pragma unroll
  for (int i = 0; i < 9 * CO; ++i) red[lane][i * CI + ci] = acc[i];
  if (ci == 0) {
#pragma unroll
    for (int co = 0; co < CO; ++co) red[lane][9 * CO * CI + co] = bsum[co];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += 256) {
    float s = red[0][i];
#pragma unroll
    for (int l = 1; l < LANES; ++l) s += red[l][i];
    p.part[(size_t)blockIdx.x * NV + i] = s;
  }
}

// dw[co][..] / dbias[co] = sum over CTAs of the partials (fixed order), scattered to the weight's layout
template <int CI, int CO>
__global__ void thin_conv_wreduce_kernel(const float* __restrict__ part, int nparts, int w_cl, float* __restrict__ dw,
                                         float* __restrict__ dbias) {
  constexpr int NV = 9 * CO * CI + CO;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NV) return;
  float s = 0.f;
  for (int c = 0; c < nparts; ++c) s += part[(size_t)c * NV + i];
  if (i >= 9 * CO * CI) {
    if (dbias) dbias[i - 9 * CO * CI] = s;
    return;
  }
  const int ci = i % CI, co = (i / CI) % CO, tap = i / (CI * CO);
  dw[w_cl ? ((size_t)co * 9 + tap) * CI + ci : ((size_t)co * CI + ci) * 9 + tap] = s;
}

static int thin_parts() { return sm_count() * 4; }

template <int CI, int CO>
static int thin_fwd(const ThinParams& p, cudaStream_t stream) {
  const long long P = (long long)p.B * p.H * p.W;
  const long long passes = ceil_div_ll(P, kThinThreads / (CI / 4));
  const int grid = (int)(passes < (long long)sm_count() * 16 ? passes : (long long)sm_count() * 16);
  thin_conv_fwd_kernel<CI, CO><<<grid, kThinThreads, 0, stream>>>(p);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

template <int CI, int CO>
static int thin_bwd(const ThinParams& p0, float* dw, float* dbias, cudaStream_t stream) {
  ThinParams p = p0;
  const long long P = (long long)p.B * p.H * p.W;
  constexpr int NVp = 9 * CO * CI + CO;
  if (p.act_out == 1) {                       // dyp = dy * (1 - y^2) behind the partials in the workspace
    float* dyp = p.part + (size_t)p.nparts * NVp;
    const long long n = P * CO;
    thin_dyp_kernel<<<(unsigned)(ceil_div_ll(n, 256) > 2368 ? 2368 : ceil_div_ll(n, 256)), 256, 0, stream>>>(p.dy, p.y, n, dyp);
    AG2V_LAUNCH_CHECK();
    p.dy = dyp;
  }
  if (p.dx) {
    thin_conv_dgrad_kernel<CI, CO><<<(unsigned)ceil_div_ll(P * (CI / 4), kThinThreads), kThinThreads, 0, stream>>>(p);
    AG2V_LAUNCH_CHECK();
  }
  thin_conv_wgrad_kernel<CI, CO><<<p.nparts, 256, 0, stream>>>(p);
  AG2V_LAUNCH_CHECK();
  constexpr int NV = 9 * CO * CI + CO;
  thin_conv_wreduce_kernel<CI, CO><<<ceil_div(NV, 256), 256, 0, stream>>>(p.part, p.nparts, p.w_cl, dw, dbias);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

}  // namespace ag2v

using namespace ag2v;

extern "C" int ag2v_thin_conv3x3_supported(int CI, int CO) {
  return ((CI == 64 && CO == 3) || (CI == 32 && CO == 2)) ? 1 : 0;
}

extern "C" size_t ag2v_thin_conv3x3_workspace_floats(int B, int H, int W, int CI, int CO) {
  return (size_t)thin_parts() * (9 * CO * CI + CO) + (size_t)B * H * W * CO;
}

extern "C" int ag2v_thin_conv3x3_fwd(const float* x, const float* w, const float* bias, int B, int H, int W, int CI, int CO,
                                     int w_channels_last, float slope_in, int act_out, float* y, cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  AG2V_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0, "thin_conv3x3_fwd: bad arguments");
  AG2V_REQUIRE(ag2v_thin_conv3x3_supported(CI, CO), "thin_conv3x3: (Ci, Co) = (%d, %d) is not instantiated", CI, CO);
  AG2V_REQUIRE(((uintptr_t)x & 15) == 0, "thin_conv3x3: x must be 16-byte aligned");
  ThinParams p{};
  p.x = x; p.w = w; p.bias = bias; p.B = B; p.H = H; p.W = W; p.w_cl = w_channels_last; p.slope_in = slope_in;
  p.act_out = act_out; p.y = y;
  return CI == 64 ? thin_fwd<64, 3>(p, stream) : thin_fwd<32, 2>(p, stream);
}

// dx may be NULL (no input gradient wanted); y is only read for act_out = 1; workspace: see above
extern "C" int ag2v_thin_conv3x3_bwd(const float* x, const float* w, const float* dy, const float* y, int B, int H, int W,
                                     int CI, int CO, int w_channels_last, float slope_in, int act_out, float* dx,
                                     float* workspace, float* dw, float* dbias, cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  AG2V_REQUIRE(x && w && dy && workspace && dw && (act_out == 0 || y), "thin_conv3x3_bwd: bad arguments");
  AG2V_REQUIRE(ag2v_thin_conv3x3_supported(CI, CO), "thin_conv3x3: (Ci, Co) = (%d, %d) is not instantiated", CI, CO);
  AG2V_REQUIRE((((uintptr_t)x | (uintptr_t)dx) & 15) == 0, "thin_conv3x3: x / dx must be 16-byte aligned");
  ThinParams p{};
  p.x = x; p.w = w; p.B = B; p.H = H; p.W = W; p.w_cl = w_channels_last; p.slope_in = slope_in; p.act_out = act_out;
  p.y = const_cast<float*>(y); p.dy = dy; p.dx = dx; p.part = workspace; p.nparts = thin_parts();
  return CI == 64 ? thin_bwd<64, 3>(p, dw, dbias, stream) : thin_bwd<32, 2>(p, dw, dbias, stream);
}
