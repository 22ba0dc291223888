// K5 — spectral normalisation of every convolution weight of a generator call in ONE launch.
//
// Reference: SPADEResnetBlock wraps conv_0 / conv_1 / conv_s in torch's spectral_norm
// (models/spade_models/networks/architecture.py:34-41) and get_nonspade_norm_layer does the
// same for the flow network and conv_dim_in (networks/normalization.py:16-50).  In training
// mode every convolution call runs one power iteration on its weight seen as a
// [Cout, Cin*kh*kw] matrix W (torch/nn/utils/spectral_norm.py, compute_weight):
//     v <- normalize(W^T u),  u <- normalize(W v),  sigma = u . (W v),  weight = W / sigma
// and the backward treats u, v as constants:
//     dW = G / sigma - (<G, W> / sigma^2) u v^T
// Eager PyTorch spends ~14 small kernels per weight forward and ~25 backward; a generator
// call has 38 such weights (≈300 MB), so a train step of three frames launches >4000 tiny
// kernels for them.  Here one cooperative kernel per direction walks a table of up to 48
// weights; the four (three) phases are separated by grid barriers and every phase is spread
// over all SMs.  All reductions use a fixed order (deterministic).
//
// Layout: a weight is either contiguous [Co][Cin][kh][kw] or channels_last
// [Co][kh][kw][Cin]; the kernels work on the physical row (K = Cin*kh*kw contiguous floats)
// and translate to torch's logical column order (ci*T + tap) only when reading / writing the
// module's `weight_v` buffer.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ag2v {

constexpr int kSnMax = 48;        // weights per launch (the host loops over larger sets)
constexpr int kSnThreads = 512;
constexpr int kSnWarps = kSnThreads / 32;
constexpr int kSnChunk = 16384;   // elements per CTA item in the element-wise phases

struct SnTensor {
  const float* w;   // weight_orig
  float* out;       // normalised weight (forward) / dW (backward)
  const float* g;   // backward: gradient w.r.t. the normalised weight
  float* u;         // weight_u buffer [co]
  float* v;         // weight_v buffer [K], logical column order
  int co, cin, taps, cl;
  int koff, roff;   // offsets of this weight's columns / rows in the packed vectors
};

struct SnParams {
  SnTensor t[kSnMax];
  int itemA[kSnMax + 1];   // prefix sums of ceil(K/32)            (phase A warp items)
  int itemC[kSnMax + 1];   // prefix sums of ceil(co*K / kSnChunk)  (element-wise CTA items)
  int n, rows;             // number of weights, total rows
  float* vphys;   // [sum K]  W^T u, then the normalised v in physical column order (saved)
  float* sraw;    // [sum co] W (W^T u)                                           (scratch)
  float* tsq;     // [itemA[n]] per-item sums of squares of W^T u                  (scratch)
  float* sigma;   // [n]                                                           (saved)
  float* usave;   // [sum co] u used in sigma                                      (saved)
  float* partial; // backward: [itemC[n]] partial <G, W>
  int power_iter;
  float eps;
  // sigma mode (frames of a clip batched into one call): `iters` successive power iterations, one
  // per frame group, each saving its sigma / u / v; no normalised weight is written — the caller
  // applies 1/sigma_g to the convolution output of group g instead.
  int iters, sigma_only;
  long long it_sigma, it_u, it_v;   // per-iteration strides of sigma / usave / vphys
};

__device__ __forceinline__ int sn_find(const int* prefix, int n, int item) {
  int t = 0;
  while (t + 1 < n && item >= prefix[t + 1]) ++t;
  return t;
}

__device__ __forceinline__ int sn_logical(const SnTensor& t, int p) {
  if (!t.cl) return p;
  const int tap = p / t.cin, ci = p - tap * t.cin;
  return ci * t.taps + tap;
}

// warp 0 only: fixed-order sum of a[0..n)
__device__ __forceinline__ float sn_warp_reduce(const float* a, int n, int lane) {
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += a[i];
  return warp_sum(s);
}

__global__ void __launch_bounds__(kSnThreads, 1) specnorm_fwd_kernel(const __grid_constant__ SnParams P) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kSnWarps + warp, nwarps = gridDim.x * kSnWarps;
  __shared__ float s_sigma;

  for (int it = 0; it < P.iters; ++it) {
  float* const vphys = P.vphys + it * P.it_v;
  float* const usave = P.usave + it * P.it_u;
  float* const sigma_it = P.sigma + it * P.it_sigma;
  // ---- phase A: t = W^T u (physical column order); lane = column, 32 columns per warp item
  for (int item = gwarp; item < P.itemA[P.n]; item += nwarps) {
    const int ti = sn_find(P.itemA, P.n, item);
    const SnTensor& T = P.t[ti];
    const int K = T.cin * T.taps;
    const int p = (item - P.itemA[ti]) * 32 + lane;
    float acc = 0.f;
    if (p < K) {
      if (P.power_iter) {
        const float* w = T.w + p;
        int r = 0;
        for (; r + 16 <= T.co; r += 16) {
          float x[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) x[j] = __ldg(w + (size_t)(r + j) * K);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc = fmaf(x[j], T.u[r + j], acc);
        }
        for (; r < T.co; ++r) acc = fmaf(__ldg(w + (size_t)r * K), T.u[r], acc);
      } else {
        acc = T.v[sn_logical(T, p)];
      }
      vphys[T.koff + p] = acc;
    }
    const float sq = warp_sum(p < K ? acc * acc : 0.f);
    if (lane == 0) P.tsq[item] = sq;
  }
  grid.sync();

  // ---- phase B: sraw = W t; one warp per row
  for (int row = gwarp; row < P.rows; row += nwarps) {
    int ti = 0;
    while (ti + 1 < P.n && row >= P.t[ti + 1].roff) ++ti;
    const SnTensor& T = P.t[ti];
    const int K = T.cin * T.taps, r = row - T.roff;
    const float* w = T.w + (size_t)r * K;
    const float* t = vphys + T.koff;
    float acc = 0.f;
    int p = lane;
    if ((K & 3) == 0 && (T.koff & 3) == 0) {          // rows and the packed vector are 16-byte aligned
      const float4* w4 = reinterpret_cast<const float4*>(w);
      const float4* t4 = reinterpret_cast<const float4*>(t);
      const int K4 = K >> 2;
      int q = lane;
      for (; q + 96 < K4; q += 128) {
        float4 a[4], b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[j] = __ldg(w4 + q + 32 * j); b[j] = t4[q + 32 * j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc = fmaf(a[j].x, b[j].x, acc); acc = fmaf(a[j].y, b[j].y, acc);
          acc = fmaf(a[j].z, b[j].z, acc); acc = fmaf(a[j].w, b[j].w, acc);
        }
      }
      for (; q < K4; q += 32) {
        const float4 a = __ldg(w4 + q), b = t4[q];
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
      p = K;
    }
    for (; p + 96 < K; p += 128) {
      const float a0 = __ldg(w + p), a1 = __ldg(w + p + 32), a2 = __ldg(w + p + 64), a3 = __ldg(w + p + 96);
      acc = fmaf(a0, t[p], acc);
      acc = fmaf(a1, t[p + 32], acc);
      acc = fmaf(a2, t[p + 64], acc);
      acc = fmaf(a3, t[p + 96], acc);
    }
    for (; p < K; p += 32) acc = fmaf(__ldg(w + p), t[p], acc);
    acc = warp_sum(acc);
    if (lane == 0) P.sraw[row] = acc;
  }
  grid.sync();

  // ---- phase S: per weight (one CTA each): sigma, new u, normalised v
  for (int ti = blockIdx.x; ti < P.n; ti += gridDim.x) {
    const SnTensor& T = P.t[ti];
    const int K = T.cin * T.taps;
    __syncthreads();
    if (warp == 0) {
      float tn = 1.f;
      if (P.power_iter) {
        const float nt = sqrtf(sn_warp_reduce(P.tsq + P.itemA[ti], P.itemA[ti + 1] - P.itemA[ti], lane));
        tn = fmaxf(nt, P.eps);
      }
      const float* sr = P.sraw + T.roff;
      float sigma;
      if (P.power_iter) {
        float q = 0.f;
        for (int r = lane; r < T.co; r += 32) { const float s = sr[r] / tn; q = fmaf(s, s, q); }
        const float un = fmaxf(sqrtf(warp_sum(q)), P.eps);
        float d = 0.f;
        for (int r = lane; r < T.co; r += 32) {
          const float s = sr[r] / tn, un_r = s / un;
          d = fmaf(un_r, s, d);
          usave[T.roff + r] = un_r;
          T.u[r] = un_r;
        }
        sigma = warp_sum(d);
      } else {
        float d = 0.f;
        for (int r = lane; r < T.co; r += 32) {
          const float ur = T.u[r];
          d = fmaf(ur, sr[r], d);
          usave[T.roff + r] = ur;
        }
        sigma = warp_sum(d);
      }
      if (lane == 0) { s_sigma = tn; sigma_it[ti] = sigma; }
    }
    __syncthreads();
    if (P.power_iter) {                        // normalised v: saved (physical) and module buffer (logical)
      const float tn = s_sigma;
      for (int p = threadIdx.x; p < K; p += kSnThreads) {
        const float vn = vphys[T.koff + p] / tn;
        vphys[T.koff + p] = vn;
        T.v[sn_logical(T, p)] = vn;
      }
    }
  }
  grid.sync();

  }   // iterations
  if (P.sigma_only) return;

  // ---- phase C: out = W / sigma; one CTA per 16K-element chunk
  for (int item = blockIdx.x; item < P.itemC[P.n]; item += gridDim.x) {
    const int ti = sn_find(P.itemC, P.n, item);
    const SnTensor& T = P.t[ti];
    const int K = T.cin * T.taps;
    const int chunk = item - P.itemC[ti];
    const float sigma = P.sigma[ti];
    const long long total = (long long)T.co * K;
    const long long i0 = (long long)chunk * kSnChunk;
    const long long i1 = i0 + kSnChunk < total ? i0 + kSnChunk : total;
    const long long v1 = i0 + ((i1 - i0) & ~3LL);
    long long i = i0 + 4 * threadIdx.x;
    for (; i + 12 * kSnThreads < v1; i += 16 * kSnThreads) {
      float4 x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = __ldg(reinterpret_cast<const float4*>(T.w + i + 4 * kSnThreads * j));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        x[j].x /= sigma; x[j].y /= sigma; x[j].z /= sigma; x[j].w /= sigma;
        *reinterpret_cast<float4*>(T.out + i + 4 * kSnThreads * j) = x[j];
      }
    }
    for (; i < v1; i += 4 * kSnThreads) {
      float4 x = __ldg(reinterpret_cast<const float4*>(T.w + i));
      x.x /= sigma; x.y /= sigma; x.z /= sigma; x.w /= sigma;
      *reinterpret_cast<float4*>(T.out + i) = x;
    }
    for (long long i = v1 + threadIdx.x; i < i1; i += kSnThreads) T.out[i] = T.w[i] / sigma;
  }
}

__global__ void __launch_bounds__(kSnThreads, 1) specnorm_bwd_kernel(const __grid_constant__ SnParams P) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ float s_red[kSnWarps];

  // ---- phase 1: partial <G, W> per chunk
  for (int item = blockIdx.x; item < P.itemC[P.n]; item += gridDim.x) {
    const int ti = sn_find(P.itemC, P.n, item);
    const SnTensor& T = P.t[ti];
    const long long total = (long long)T.co * T.cin * T.taps;
    const long long i0 = (long long)(item - P.itemC[ti]) * kSnChunk;
    const long long i1 = i0 + kSnChunk < total ? i0 + kSnChunk : total;
    const long long v1 = i0 + ((i1 - i0) & ~3LL);
    float acc = 0.f;
    long long i = i0 + 4 * threadIdx.x;
    for (; i + 4 * kSnThreads < v1; i += 8 * kSnThreads) {
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(T.g + i));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(T.w + i));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(T.g + i + 4 * kSnThreads));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(T.w + i + 4 * kSnThreads));
      acc = fmaf(a0.x, b0.x, acc); acc = fmaf(a0.y, b0.y, acc); acc = fmaf(a0.z, b0.z, acc); acc = fmaf(a0.w, b0.w, acc);
      acc = fmaf(a1.x, b1.x, acc); acc = fmaf(a1.y, b1.y, acc); acc = fmaf(a1.z, b1.z, acc); acc = fmaf(a1.w, b1.w, acc);
    }
    for (; i < v1; i += 4 * kSnThreads) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(T.g + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(T.w + i));
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
    for (long long j = v1 + threadIdx.x; j < i1; j += kSnThreads) acc = fmaf(T.g[j], T.w[j], acc);
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) s_red[warp] = acc;
    __syncthreads();
    if (warp == 0) {
      const float s = warp_sum(lane < kSnWarps ? s_red[lane] : 0.f);
      if (lane == 0) P.partial[item] = s;
    }
  }
  grid.sync();

  // ---- per weight: coef = <G,W> / sigma^2 (one warp each, fixed order)
  {
    const int gw = blockIdx.x * kSnWarps + warp;
    if (gw < P.n) {
      const float dot = sn_warp_reduce(P.partial + P.itemC[gw], P.itemC[gw + 1] - P.itemC[gw], lane);
      const float sg = P.sigma[gw];
      if (lane == 0) P.partial[P.itemC[P.n] + gw] = dot / (sg * sg);
    }
  }
  grid.sync();

  // ---- phase 2: dW = G / sigma - coef u v^T
  for (int item = blockIdx.x; item < P.itemC[P.n]; item += gridDim.x) {
    const int ti = sn_find(P.itemC, P.n, item);
    const SnTensor& T = P.t[ti];
    const int K = T.cin * T.taps;
    const float coef = P.partial[P.itemC[P.n] + ti], sigma = P.sigma[ti];
    const float* u = P.usave + T.roff;
    const float* v = P.vphys + T.koff;
    const long long total = (long long)T.co * K;
    const long long i0 = (long long)(item - P.itemC[ti]) * kSnChunk;
    const long long i1 = i0 + kSnChunk < total ? i0 + kSnChunk : total;
    if ((K & 3) == 0 && (T.koff & 3) == 0) {          // four consecutive elements share a row
      const long long v1 = i0 + ((i1 - i0) & ~3LL);    // (chunks start at multiples of 4; total % 4 == 0)
      for (long long i = i0 + 4 * threadIdx.x; i < v1; i += 4 * kSnThreads) {
        const int r = (int)(i / K), p = (int)(i - (long long)r * K);
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(T.g + i));
        const float4 v4 = *reinterpret_cast<const float4*>(v + p);
        const float cu = coef * u[r];
        float4 o;
        o.x = g4.x / sigma - cu * v4.x; o.y = g4.y / sigma - cu * v4.y;
        o.z = g4.z / sigma - cu * v4.z; o.w = g4.w / sigma - cu * v4.w;
        *reinterpret_cast<float4*>(T.out + i) = o;
      }
    } else {
      for (long long i = i0 + threadIdx.x; i < i1; i += kSnThreads) {
        const int r = (int)(i / K), p = (int)(i - (long long)r * K);
        T.out[i] = T.g[i] / sigma - coef * u[r] * v[p];
      }
    }
  }
}

// sigma mode backward: dW = sum_it dsigma[it] u_it v_it^T  (the direct path G/sigma reaches
// weight_orig through the convolution itself, which runs on the unnormalised weight)
__global__ void __launch_bounds__(kSnThreads) specnorm_sigma_bwd_kernel(const __grid_constant__ SnParams P,
                                                                        const float* __restrict__ dsigma) {
  for (int item = blockIdx.x; item < P.itemC[P.n]; item += gridDim.x) {
    const int ti = sn_find(P.itemC, P.n, item);
    const SnTensor& T = P.t[ti];
    const int K = T.cin * T.taps;
    const long long total = (long long)T.co * K;
    const long long i0 = (long long)(item - P.itemC[ti]) * kSnChunk;
    const long long i1 = i0 + kSnChunk < total ? i0 + kSnChunk : total;
    if ((K & 3) == 0) {
      for (long long i = i0 + 4 * threadIdx.x; i < i1; i += 4 * kSnThreads) {
        const int r = (int)(i / K), p = (int)(i - (long long)r * K);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int it = 0; it < P.iters; ++it) {
          const float cu = dsigma[it * P.n + ti] * P.usave[it * P.it_u + T.roff + r];
          const float4 v4 = *reinterpret_cast<const float4*>(P.vphys + it * P.it_v + T.koff + p);
          o.x = fmaf(cu, v4.x, o.x); o.y = fmaf(cu, v4.y, o.y); o.z = fmaf(cu, v4.z, o.z); o.w = fmaf(cu, v4.w, o.w);
        }
        *reinterpret_cast<float4*>(T.out + i) = o;
      }
    } else {
      for (long long i = i0 + threadIdx.x; i < i1; i += kSnThreads) {
        const int r = (int)(i / K), p = (int)(i - (long long)r * K);
        float o = 0.f;
        for (int it = 0; it < P.iters; ++it)
          o = fmaf(dsigma[it * P.n + ti] * P.usave[it * P.it_u + T.roff + r], P.vphys[it * P.it_v + T.koff + p], o);
        T.out[i] = o;
      }
    }
  }
}

// sigma mode backward for ONE weight of the list, straight from the gradients of its 1/sigma scales (so that the weight's
// gradient is complete as soon as its own layer has run its backward - the gradient all-reduce of the data-parallel path
// can then start layer by layer instead of after a batched kernel at the end of the backward pass):
//   dW = sum_it coef_it u_it v_it^T,  coef_it = - (dscale_g[it] + sum_j dscale_img[it * ipg + j]) / sigma_it^2
// (scale = 1 / sigma per frame group; dscale_img is the gradient of the per-image copy of it).
__global__ void __launch_bounds__(kSnThreads) specnorm_scale_bwd_one_kernel(const __grid_constant__ SnParams P, int ti,
                                                                            const float* __restrict__ dscale_g,
                                                                            const float* __restrict__ dscale_img, int ipg) {
  __shared__ float coef[64];
  if (threadIdx.x < P.iters) {
    const int it = threadIdx.x;
    float d = dscale_g ? dscale_g[it] : 0.f;
    if (dscale_img)
      for (int j = 0; j < ipg; ++j) d += dscale_img[it * ipg + j];
    const float sg = P.sigma[it * P.it_sigma + ti];
    coef[it] = -d / (sg * sg);
  }
  __syncthreads();
  const SnTensor& T = P.t[ti];
  const int K = T.cin * T.taps;
  const long long total = (long long)T.co * K;
  for (int item = P.itemC[ti] + blockIdx.x; item < P.itemC[ti + 1]; item += gridDim.x) {
    const long long i0 = (long long)(item - P.itemC[ti]) * kSnChunk;
    const long long i1 = i0 + kSnChunk < total ? i0 + kSnChunk : total;
    if ((K & 3) == 0) {
      for (long long i = i0 + 4 * threadIdx.x; i < i1; i += 4 * kSnThreads) {
        const int r = (int)(i / K), p = (int)(i - (long long)r * K);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int it = 0; it < P.iters; ++it) {
          const float cu = coef[it] * P.usave[it * P.it_u + T.roff + r];
          const float4 v4 = *reinterpret_cast<const float4*>(P.vphys + it * P.it_v + T.koff + p);
          o.x = fmaf(cu, v4.x, o.x); o.y = fmaf(cu, v4.y, o.y); o.z = fmaf(cu, v4.z, o.z); o.w = fmaf(cu, v4.w, o.w);
        }
        *reinterpret_cast<float4*>(T.out + i) = o;
      }
    } else {
      for (long long i = i0 + threadIdx.x; i < i1; i += kSnThreads) {
        const int r = (int)(i / K), p = (int)(i - (long long)r * K);
        float o = 0.f;
        for (int it = 0; it < P.iters; ++it)
          o = fmaf(coef[it] * P.usave[it * P.it_u + T.roff + r], P.vphys[it * P.it_v + T.koff + p], o);
        T.out[i] = o;
      }
    }
  }
}

static int fill_params(SnParams& P, int n, const void* const* w, void* const* out, const void* const* g,
                       void* const* u, void* const* v, const int* co, const int* cin, const int* taps,
                       const int* cl, float* save, size_t save_floats, float* scratch, size_t scratch_floats,
                       bool backward, int iters = 1, bool sigma_only = false) {
  AG2V_REQUIRE(n >= 1 && n <= kSnMax, "spectral norm: n=%d outside [1,%d]", n, kSnMax);
  AG2V_REQUIRE(iters >= 1 && iters <= 64, "spectral norm: iters=%d outside [1,64]", iters);
  int koff = 0, roff = 0;
  P.itemA[0] = 0; P.itemC[0] = 0;
  for (int i = 0; i < n; ++i) {
    const bool need_w = !(sigma_only && backward), need_out = !(sigma_only && !backward);
    AG2V_REQUIRE((!need_w || w[i]) && (!need_out || out[i]) && co[i] > 0 && cin[i] > 0 && taps[i] > 0,
                 "spectral norm: bad weight %d", i);
    AG2V_REQUIRE((!need_w || ((uintptr_t)w[i] & 15) == 0) && (!need_out || ((uintptr_t)out[i] & 15) == 0),
                 "spectral norm: weight %d not 16-byte aligned", i);
    SnTensor& T = P.t[i];
    T.w = need_w ? (const float*)w[i] : nullptr; T.out = need_out ? (float*)out[i] : nullptr;
    T.g = backward && !sigma_only ? (const float*)g[i] : nullptr;
    if (backward && !sigma_only) AG2V_REQUIRE(g[i] && ((uintptr_t)g[i] & 15) == 0, "spectral norm: gradient %d null or misaligned", i);
    T.u = backward ? nullptr : (float*)u[i];
    T.v = backward ? nullptr : (float*)v[i];
    if (!backward) AG2V_REQUIRE(u[i] && v[i], "spectral norm: u/v of weight %d is null", i);
    T.co = co[i]; T.cin = cin[i]; T.taps = taps[i]; T.cl = cl[i] && taps[i] > 1 && cin[i] > 1;
    T.koff = koff; T.roff = roff;
    const long long K = (long long)cin[i] * taps[i];
    koff += ((int)K + 3) & ~3; roff += co[i];       // packed vectors keep every weight 16-byte aligned
    P.itemA[i + 1] = P.itemA[i] + ceil_div((int)K, 32);
    P.itemC[i + 1] = P.itemC[i] + (int)ceil_div_ll((long long)co[i] * K, kSnChunk);
  }
  for (int i = n; i < kSnMax; ++i) { P.t[i] = P.t[n - 1]; P.itemA[i + 1] = P.itemA[n]; P.itemC[i + 1] = P.itemC[n]; }
  P.n = n; P.rows = roff;
  // saved: sigma[iters][n] | usave[iters][rows] | vphys[iters][sum K], rows padded to multiples of 4 floats
  const int n4 = (n + 3) & ~3, r4 = (roff + 3) & ~3;
  const size_t save_need = (size_t)iters * ((size_t)n4 + r4 + koff);
  AG2V_REQUIRE(save && ((uintptr_t)save & 15) == 0 && save_floats >= save_need,
               "spectral norm: saved buffer too small or misaligned (%zu < %zu)", save_floats, save_need);
  P.sigma = save; P.usave = save + (size_t)iters * n4; P.vphys = save + (size_t)iters * (n4 + r4);
  P.iters = iters; P.sigma_only = sigma_only ? 1 : 0;
  P.it_sigma = n4; P.it_u = r4; P.it_v = koff;
  const size_t need = backward ? (sigma_only ? 0 : (size_t)P.itemC[n] + n) : (size_t)roff + P.itemA[n];
  AG2V_REQUIRE(need == 0 || (scratch && scratch_floats >= need), "spectral norm: scratch too small (%zu < %zu)", scratch_floats, need);
  P.sraw = scratch; P.tsq = scratch + roff; P.partial = scratch;
  return AG2V_OK;
}

}  // namespace ag2v

using namespace ag2v;

extern "C" int ag2v_spectral_norm_sizes(int n, const int* co, const int* cin, const int* taps,
                                        size_t* save_floats, size_t* fwd_scratch_floats, size_t* bwd_scratch_floats) {
  AG2V_REQUIRE(n >= 1 && n <= kSnMax && co && cin && taps, "spectral norm sizes: bad arguments");
  size_t rows = 0, cols = 0, ia = 0, ic = 0;
  for (int i = 0; i < n; ++i) {
    const long long K = (long long)cin[i] * taps[i];
    rows += co[i]; cols += (K + 3) & ~3LL; ia += ceil_div((int)K, 32); ic += (size_t)ceil_div_ll((long long)co[i] * K, kSnChunk);
  }
  if (save_floats) *save_floats = ((n + 3) & ~3) + ((rows + 3) & ~(size_t)3) + cols;
  if (fwd_scratch_floats) *fwd_scratch_floats = rows + ia;
  if (bwd_scratch_floats) *bwd_scratch_floats = ic + n;
  return AG2V_OK;
}

extern "C" int ag2v_spectral_norm_fwd(int n, const void* const* w, void* const* out, void* const* u, void* const* v,
                                      const int* co, const int* cin, const int* taps, const int* channels_last,
                                      float* save, size_t save_floats, float* scratch, size_t scratch_floats,
                                      int power_iteration, float eps, cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  SnParams P;
  if (int rc = fill_params(P, n, w, out, nullptr, u, v, co, cin, taps, channels_last, save, save_floats, scratch,
                           scratch_floats, false)) return rc;
  P.power_iter = power_iteration ? 1 : 0;
  P.eps = eps;
  void* args[] = {&P};
  AG2V_COOP_LAUNCH(specnorm_fwd_kernel, dim3(sm_count()), dim3(kSnThreads), args, 0, stream);
  return AG2V_OK;
}

extern "C" int ag2v_spectral_norm_bwd(int n, const void* const* w, const void* const* grad_out, void* const* grad_w,
                                      const int* co, const int* cin, const int* taps, const int* channels_last,
                                      const float* save, size_t save_floats, float* scratch, size_t scratch_floats,
                                      cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  SnParams P;
  if (int rc = fill_params(P, n, w, grad_w, grad_out, nullptr, nullptr, co, cin, taps, channels_last,
                           const_cast<float*>(save), save_floats, scratch, scratch_floats, true)) return rc;
  P.power_iter = 0;
  P.eps = 0.f;
  void* args[] = {&P};
  AG2V_COOP_LAUNCH(specnorm_bwd_kernel, dim3(sm_count()), dim3(kSnThreads), args, 0, stream);
  return AG2V_OK;
}

// Sigma mode: `iters` successive power iterations (one per frame group of a batched call); writes
// no weight.  save (iters * ag2v_spectral_norm_sizes floats) starts with sigma [iters][(n+3)&~3].
extern "C" int ag2v_spectral_norm_sigma_fwd(int n, const void* const* w, void* const* u, void* const* v, const int* co,
                                            const int* cin, const int* taps, const int* channels_last, int iters,
                                            float* save, size_t save_floats, float* scratch, size_t scratch_floats,
                                            int power_iteration, float eps, cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  SnParams P;
  if (int rc = fill_params(P, n, w, nullptr, nullptr, u, v, co, cin, taps, channels_last, save, save_floats, scratch,
                           scratch_floats, false, iters, true)) return rc;
  P.power_iter = power_iteration ? 1 : 0;
  P.eps = eps;
  void* args[] = {&P};
  AG2V_COOP_LAUNCH(specnorm_fwd_kernel, dim3(sm_count()), dim3(kSnThreads), args, 0, stream);
  return AG2V_OK;
}

// grad_w[i] = sum_it dsigma[it][i] * u_it v_it^T  (dsigma: device array [iters][n])
extern "C" int ag2v_spectral_norm_sigma_bwd(int n, const float* dsigma, void* const* grad_w, const int* co,
                                            const int* cin, const int* taps, const int* channels_last, int iters,
                                            const float* save, size_t save_floats, cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  AG2V_REQUIRE(dsigma, "spectral norm: dsigma is null");
  SnParams P;
  if (int rc = fill_params(P, n, nullptr, grad_w, nullptr, nullptr, nullptr, co, cin, taps, channels_last,
                           const_cast<float*>(save), save_floats, nullptr, 0, true, iters, true)) return rc;
  P.power_iter = 0;
  P.eps = 0.f;
  const int items = P.itemC[n];
  specnorm_sigma_bwd_kernel<<<items < 4 * sm_count() ? items : 4 * sm_count(), kSnThreads, 0, stream>>>(P, dsigma);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// grad_w = gradient of weight `only` of the list (geometry arrays and `save` describe the WHOLE list of the forward call)
// from the gradients of its scales: dscale_g [iters] (per frame group) and / or dscale_img [iters * images_per_group]
// (per image); either may be null.  One launch per weight, issued by that weight's own autograd node.
extern "C" int ag2v_spectral_norm_scale_bwd_one(int n, int only, const float* dscale_g, const float* dscale_img,
                                                int images_per_group, void* grad_w, const int* co, const int* cin,
                                                const int* taps, const int* channels_last, int iters, const float* save,
                                                size_t save_floats, cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  AG2V_REQUIRE(only >= 0 && only < n && grad_w, "spectral norm: weight %d of %d / null gradient", only, n);
  AG2V_REQUIRE(dscale_g || dscale_img, "spectral norm: no scale gradient");
  AG2V_REQUIRE(!dscale_img || images_per_group >= 1, "spectral norm: images_per_group=%d", images_per_group);
  void* outs[kSnMax];
  for (int i = 0; i < n && i < kSnMax; ++i) outs[i] = grad_w;        // only entry `only` is written
  SnParams P;
  if (int rc = fill_params(P, n, nullptr, outs, nullptr, nullptr, nullptr, co, cin, taps, channels_last,
                           const_cast<float*>(save), save_floats, nullptr, 0, true, iters, true)) return rc;
  P.power_iter = 0;
  P.eps = 0.f;
  const int items = P.itemC[only + 1] - P.itemC[only];
  specnorm_scale_bwd_one_kernel<<<items < 4 * sm_count() ? items : 4 * sm_count(), kSnThreads, 0, stream>>>(
      P, only, dscale_g, dscale_img, images_per_group);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}
