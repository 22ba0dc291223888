// K3 — HBM-bound pieces of SPADE: batch-norm statistics, the backward
// element-wise passes, weight (un)packing.  All activations are NHWC fp32.
// Per-channel reductions are two-level and fixed-order (deterministic): per-CTA
// partials in fp32, combined in double.
#include "k3_common.cuh"

namespace ag2v {

constexpr int kElemThreads = 256;

// tcgen05.mma kind::tf32 reads fp32 words and ignores the low 13 mantissa bits
// (truncation, a biased error of ~2^-11 per operand).  Every GEMM operand is
// therefore rounded to nearest TF32 where it is produced.
__device__ __forceinline__ float elem_round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
constexpr int kMaxSplit = 296;       // 2 CTAs per SM

struct ChanGeom { int cx, py, chunks, nsplit; };

static ChanGeom chan_geom(long long P, int C) {
  ChanGeom g;
  int c4 = C / 4, cx = 1;
  while (cx < c4 && cx < kElemThreads) cx <<= 1;
  g.cx = cx; g.py = kElemThreads / cx; g.chunks = ceil_div(c4, cx);
  long long ns = ceil_div_ll(P, (long long)g.py * 8);
  g.nsplit = (int)(ns < 1 ? 1 : (ns > kMaxSplit ? kMaxSplit : ns));
  return g;
}

// MODE 0: s0 = sum x, s1 = sum x^2                      (F.batch_norm statistics)
// MODE 1: SPADE backward, pass 1 (see spade_bwd_pre below); s4 = sum dout*out serves the
//         spectral norm of the convolution that consumed `out` (d sigma, see spade.py)
// MODE 3: gradient of a scaled convolution output: dxhat = tf32(dout * scale[group]), s0 = sum dout
struct ChanArgs {
  const float* x; long long P; int C;
  float* part;                      // [nsplit][NS][C]
  // MODE 1
  const float* dout; const float* out; const float* gamma; const float* mean; const float* rstd;
  float* dgb; float* dxhat; float slope; int act; int round_ops;
  int chan_gamma;                   // MODE 1: gamma is a per-channel scale w[c] (affine batch norm) instead of per-pixel 1+gamma
  int up_h, up_w;                   // MODE 1, > 0: x is [.., up_h/2, up_w/2, C], read through a nearest 2x up-sampling
};
// Groups: blockIdx.y selects one of G independent pixel ranges of P rows each (the frames of a
// clip batched into one call keep per-call batch statistics); every per-group array
// (x .. dxhat, mean/rstd, part, sums) is laid out group-major.

template <int MODE>
__global__ void __launch_bounds__(kElemThreads) chan_partial_kernel(ChanArgs a, ChanGeom gm) {
  constexpr int NS = MODE == 0 ? 2 : (MODE == 1 ? 5 : 1);
  __shared__ float4 red[kElemThreads];
  const int tx = threadIdx.x % gm.cx, ty = threadIdx.x / gm.cx;
  const int C = a.C, c4n = C >> 2;
  {                                                     // this group's slices
    const size_t go = (size_t)blockIdx.y * a.P * C;
    if (MODE != 3) a.x += (MODE == 1 && a.up_w > 0) ? go / 4 : go;
    a.part += (size_t)blockIdx.y * gm.nsplit * NS * C;
    if (MODE == 3) { a.dout += go; a.dxhat += go; }
    if (MODE == 1) {
      a.dout += go; a.out += go; a.dxhat += go;
      if (!a.chan_gamma) a.gamma += go;
      if (a.dgb) a.dgb += 2 * go;
      a.mean += (size_t)blockIdx.y * C; a.rstd += (size_t)blockIdx.y * C;
    }
  }
  for (int chunk = 0; chunk < gm.chunks; ++chunk) {
    const int c4 = chunk * gm.cx + tx;
    const bool live = c4 < c4n;
    const int c = c4 << 2;
    float4 s[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) s[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), rs = mu;
    if (MODE == 1 && live) { mu = *reinterpret_cast<const float4*>(a.mean + c); rs = *reinterpret_cast<const float4*>(a.rstd + c); }
    if (live) {
      for (long long p = (long long)blockIdx.x * gm.py + ty; p < a.P; p += (long long)gm.nsplit * gm.py) {
        const size_t off = (size_t)p * C + c;
        if (MODE == 3) {
          const float4 dv = *reinterpret_cast<const float4*>(a.dout + off);
          const float sc = a.gamma ? a.gamma[blockIdx.y] : 1.f;
          *reinterpret_cast<float4*>(a.dxhat + off) = make_float4(elem_round_tf32(dv.x * sc), elem_round_tf32(dv.y * sc),
                                                                  elem_round_tf32(dv.z * sc), elem_round_tf32(dv.w * sc));
          s[0].x += dv.x; s[0].y += dv.y; s[0].z += dv.z; s[0].w += dv.w;
          continue;
        }
        size_t xoff = off;
        if (MODE == 1 && a.up_w > 0) {                    // pixel p of this group -> its low-resolution parent
          const long long per = (long long)a.up_h * a.up_w;
          const long long b = p / per;
          const int rem = (int)(p - b * per);
          const int y = rem / a.up_w, xx = rem - y * a.up_w;
          xoff = (size_t)((b * (a.up_h >> 1) + (y >> 1)) * (a.up_w >> 1) + (xx >> 1)) * C + c;
        }
        const float4 xv = *reinterpret_cast<const float4*>(a.x + xoff);
        if (MODE == 0) {
          s[0].x += xv.x; s[0].y += xv.y; s[0].z += xv.z; s[0].w += xv.w;
          s[1].x = fmaf(xv.x, xv.x, s[1].x); s[1].y = fmaf(xv.y, xv.y, s[1].y);
          s[1].z = fmaf(xv.z, xv.z, s[1].z); s[1].w = fmaf(xv.w, xv.w, s[1].w);
        } else {
          const float4 dv = *reinterpret_cast<const float4*>(a.dout + off);
          const float4 ov = *reinterpret_cast<const float4*>(a.out + off);
          float4 gv = *reinterpret_cast<const float4*>(a.gamma + (a.chan_gamma ? (size_t)c : off));
          if (!a.chan_gamma) { gv.x += 1.f; gv.y += 1.f; gv.z += 1.f; gv.w += 1.f; }
          float4 g, xh, dxh;
          const float sl = a.slope;
          g.x = (a.act && !(ov.x > 0.f)) ? dv.x * sl : dv.x;
          g.y = (a.act && !(ov.y > 0.f)) ? dv.y * sl : dv.y;
          g.z = (a.act && !(ov.z > 0.f)) ? dv.z * sl : dv.z;
          g.w = (a.act && !(ov.w > 0.f)) ? dv.w * sl : dv.w;
          xh.x = (xv.x - mu.x) * rs.x; xh.y = (xv.y - mu.y) * rs.y;
          xh.z = (xv.z - mu.z) * rs.z; xh.w = (xv.w - mu.w) * rs.w;
          const float4 dgam = make_float4(g.x * xh.x, g.y * xh.y, g.z * xh.z, g.w * xh.w);
          dxh.x = g.x * gv.x; dxh.y = g.y * gv.y;
          dxh.z = g.z * gv.z; dxh.w = g.w * gv.w;
          float* row = a.dgb + (size_t)p * 2 * C;
          // the GEMM operand copies are rounded to TF32; the sums below use the fp32 values
          if (a.dgb == nullptr) {
          } else if (a.round_ops) {
            *reinterpret_cast<float4*>(row + gb8_col(c, 0)) = make_float4(elem_round_tf32(dgam.x), elem_round_tf32(dgam.y), elem_round_tf32(dgam.z), elem_round_tf32(dgam.w));
            *reinterpret_cast<float4*>(row + gb8_col(c, 1)) = make_float4(elem_round_tf32(g.x), elem_round_tf32(g.y), elem_round_tf32(g.z), elem_round_tf32(g.w));
          } else {
            *reinterpret_cast<float4*>(row + gb8_col(c, 0)) = dgam;
            *reinterpret_cast<float4*>(row + gb8_col(c, 1)) = g;
          }
          *reinterpret_cast<float4*>(a.dxhat + off) = dxh;
          s[0].x += g.x; s[0].y += g.y; s[0].z += g.z; s[0].w += g.w;
          s[1].x += dgam.x; s[1].y += dgam.y; s[1].z += dgam.z; s[1].w += dgam.w;
          s[2].x += dxh.x; s[2].y += dxh.y; s[2].z += dxh.z; s[2].w += dxh.w;
          s[3].x = fmaf(dxh.x, xh.x, s[3].x); s[3].y = fmaf(dxh.y, xh.y, s[3].y);
          s[3].z = fmaf(dxh.z, xh.z, s[3].z); s[3].w = fmaf(dxh.w, xh.w, s[3].w);
          s[NS - 1].x = fmaf(dv.x, ov.x, s[NS - 1].x); s[NS - 1].y = fmaf(dv.y, ov.y, s[NS - 1].y);
          s[NS - 1].z = fmaf(dv.z, ov.z, s[NS - 1].z); s[NS - 1].w = fmaf(dv.w, ov.w, s[NS - 1].w);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      __syncthreads();
      red[threadIdx.x] = s[i];
      __syncthreads();
      if (ty == 0 && live) {
        float4 t = red[tx];
        for (int r = 1; r < gm.py; ++r) {
          float4 u = red[r * gm.cx + tx];
          t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
        }
        *reinterpret_cast<float4*>(a.part + ((size_t)blockIdx.x * NS + i) * C + c) = t;
      }
    }
  }
}

// sums[i][c] (double) = sum over splits of part[split][i][c].  Block = 32 columns x 8
// split lanes (coalesced 128-byte rows); fixed summation order => deterministic.
__global__ void __launch_bounds__(256) chan_reduce_kernel(const float* __restrict__ part, int nsplit, int NSC,
                                                          double* __restrict__ sums) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  part += (size_t)blockIdx.y * nsplit * NSC;
  sums += (size_t)blockIdx.y * NSC;
  double acc = 0.0;
  if (col < NSC) {
    int s = ty;
    for (; s + 24 < nsplit; s += 32) {
      const float a = part[(size_t)s * NSC + col], b = part[(size_t)(s + 8) * NSC + col];
      const float c = part[(size_t)(s + 16) * NSC + col], d = part[(size_t)(s + 24) * NSC + col];
      acc += (double)a; acc += (double)b; acc += (double)c; acc += (double)d;
    }
    for (; s < nsplit; s += 8) acc += (double)part[(size_t)s * NSC + col];
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < NSC) {
    double t = red[0][tx];
#pragma unroll
    for (int r = 1; r < 8; ++r) t += red[r][tx];
    sums[col] = t;
  }
}

// F.batch_norm(training=True, momentum, eps): biased variance for normalisation,
// unbiased for the running estimate (sync_batchnorm/batchnorm.py:63-68,128-145).
// With G groups the running estimates are updated G times in group order, exactly as G
// successive calls of the layer would.
// in_scale (optional, [G]): the layer's input is s_g * x but the statistics were taken on x (a
// spectrally normalised convolution evaluated on weight_orig, 1/sigma_g not yet applied).
// BN(s x; eps) == BN(x; eps / s^2), so only eps and the running estimates change and the
// multiplication by s never touches the activation.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, double unbias_count, int C, int G,
                                   float eps, float momentum, const float* __restrict__ in_scale,
                                   float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int g = 0; g < G; ++g) {
    const double sc = in_scale ? (double)in_scale[g] : 1.0;
    const double m = sums[(size_t)g * 2 * C + c] / count;
    double var = sums[(size_t)g * 2 * C + C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[(size_t)g * C + c] = (float)m;
    rstd[(size_t)g * C + c] = (float)(1.0 / sqrt(var + (double)eps / (sc * sc)));
    if (running_mean != nullptr) {
      const double unbiased = unbias_count > 1.0 ? var * unbias_count / (unbias_count - 1.0) : var;
      running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * sc * m);
      running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * sc * sc * unbiased);
    }
  }
}

// eval mode: statistics are the running estimates
// (with in_scale: (s x - rm) / sqrt(rv + eps) == (x - rm / s) * (s / sqrt(rv + eps)), per group)
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                     int C, int G, float eps, const float* __restrict__ in_scale,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int g = 0; g < G; ++g) {
    const float sc = in_scale ? in_scale[g] : 1.f;
    mean[(size_t)g * C + c] = running_mean[c] / sc;
    rstd[(size_t)g * C + c] = sc / sqrtf(running_var[c] + eps);
  }
}

// pass 2 of the SPADE backward: batch-norm input gradient, in place on dxhat.
//   training: dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat));  eval: dx = rstd * dxhat
__global__ void __launch_bounds__(kElemThreads)
spade_bwd_dx_kernel(const float* __restrict__ x, float* __restrict__ dxhat, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const double* __restrict__ sums, double count, int training,
                    long long P, int C) {
  const long long n4 = P * (C >> 2);
  x += (size_t)blockIdx.y * P * C; dxhat += (size_t)blockIdx.y * P * C;
  mean += (size_t)blockIdx.y * C; rstd += (size_t)blockIdx.y * C; sums += (size_t)blockIdx.y * 5 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (C >> 2)) << 2;
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    float4 d = reinterpret_cast<float4*>(dxhat)[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    if (training) {
      const float m1x = (float)(sums[2 * C + c] / count), m2x = (float)(sums[3 * C + c] / count);
      const float m1y = (float)(sums[2 * C + c + 1] / count), m2y = (float)(sums[3 * C + c + 1] / count);
      const float m1z = (float)(sums[2 * C + c + 2] / count), m2z = (float)(sums[3 * C + c + 2] / count);
      const float m1w = (float)(sums[2 * C + c + 3] / count), m2w = (float)(sums[3 * C + c + 3] / count);
      d.x = rs.x * (d.x - m1x - (xv.x - mu.x) * rs.x * m2x);
      d.y = rs.y * (d.y - m1y - (xv.y - mu.y) * rs.y * m2y);
      d.z = rs.z * (d.z - m1z - (xv.z - mu.z) * rs.z * m2z);
      d.w = rs.w * (d.w - m1w - (xv.w - mu.w) * rs.w * m2w);
    } else {
      d.x *= rs.x; d.y *= rs.y; d.z *= rs.z; d.w *= rs.w;
    }
    reinterpret_cast<float4*>(dxhat)[i] = d;
  }
}

// The same pass when x was up-sampled 2x (nearest) on the fly: dxhat is full resolution
// [.., H, W, C]; the gradient of the low-resolution x [.., H/2, W/2, C] is the sum over its four
// children, written to dx_low.  m1 / m2 are means over the FULL-resolution elements (count).
__global__ void __launch_bounds__(kElemThreads)
spade_bwd_dx_up_kernel(const float* __restrict__ x, const float* __restrict__ dxhat, const float* __restrict__ mean,
                       const float* __restrict__ rstd, const double* __restrict__ sums, double count, int training,
                       long long P_low, int C, int H, int W, float* __restrict__ dx_low) {
  const long long n4 = P_low * (C >> 2);
  const int h = H >> 1, w = W >> 1;
  x += (size_t)blockIdx.y * P_low * C; dx_low += (size_t)blockIdx.y * P_low * C;
  dxhat += (size_t)blockIdx.y * P_low * 4 * C;
  mean += (size_t)blockIdx.y * C; rstd += (size_t)blockIdx.y * C; sums += (size_t)blockIdx.y * 5 * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (C >> 2)) << 2;
    const long long pl = i / (C >> 2);
    const long long b = pl / ((long long)h * w);
    const int rem = (int)(pl - b * h * w);
    const int yl = rem / w, xl = rem - yl * w;
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const size_t pf = (size_t)((b * H + 2 * yl + (k >> 1)) * W + 2 * xl + (k & 1));
      const float4 t = *reinterpret_cast<const float4*>(dxhat + pf * C + c);
      d.x += t.x; d.y += t.y; d.z += t.z; d.w += t.w;
    }
    if (training) {     // sum over the 4 children of rstd * (dxhat - m1 - xhat * m2)
      const float m1x = (float)(sums[2 * C + c] / count), m2x = (float)(sums[3 * C + c] / count);
      const float m1y = (float)(sums[2 * C + c + 1] / count), m2y = (float)(sums[3 * C + c + 1] / count);
      const float m1z = (float)(sums[2 * C + c + 2] / count), m2z = (float)(sums[3 * C + c + 2] / count);
      const float m1w = (float)(sums[2 * C + c + 3] / count), m2w = (float)(sums[3 * C + c + 3] / count);
      d.x = rs.x * (d.x - 4.f * (m1x + (xv.x - mu.x) * rs.x * m2x));
      d.y = rs.y * (d.y - 4.f * (m1y + (xv.y - mu.y) * rs.y * m2y));
      d.z = rs.z * (d.z - 4.f * (m1z + (xv.z - mu.z) * rs.z * m2z));
      d.w = rs.w * (d.w - 4.f * (m1w + (xv.w - mu.w) * rs.w * m2w));
    } else {
      d.x *= rs.x; d.y *= rs.y; d.z *= rs.z; d.w *= rs.w;
    }
    reinterpret_cast<float4*>(dx_low)[i] = d;
  }
}

// Pack OIHW 3x3 weights for the implicit GEMMs.
//   forward : dst[tap][n][ci]  = W(n)[co(n)][ci][tap]          rows = output columns
//   dgrad   : dst[tap][ci][k]  = W(k)[co(k)][ci][8 - tap]      (transposed + flipped)
// With two sources (gamma, beta) the combined channel index follows gb8_col.
__global__ void pack_w3x3_kernel(const float* __restrict__ wa, const float* __restrict__ wb,
                                 const float* __restrict__ ba, const float* __restrict__ bb, int Co, int Ci,
                                 int dgrad, int round_ops, float* __restrict__ dst, float* __restrict__ bias_dst) {
  const int Ntot = wb ? 2 * Co : Co;
  const long long total = 9LL * Ntot * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int tap, n, ci;
    if (!dgrad) { ci = (int)(i % Ci); n = (int)((i / Ci) % Ntot); tap = (int)(i / ((long long)Ci * Ntot)); }
    else { n = (int)(i % Ntot); ci = (int)((i / Ntot) % Ci); tap = (int)(i / ((long long)Ci * Ntot)); }
    int co = n, is_b = 0;
    if (wb) { is_b = (n >> 3) & 1; co = ((n >> 4) << 3) + (n & 7); }
    const float* src = is_b ? wb : wa;
    const int t = dgrad ? 8 - tap : tap;
    const float w = src[((size_t)co * Ci + ci) * 9 + t];
    dst[i] = round_ops ? elem_round_tf32(w) : w;                     // tcgen05 kind::tf32 truncates: round here
  }
  if (bias_dst != nullptr && blockIdx.x == 0) {
    for (int n = threadIdx.x; n < Ntot; n += blockDim.x) {
      int co = n, is_b = 0;
      if (wb) { is_b = (n >> 3) & 1; co = ((n >> 4) << 3) + (n & 7); }
      const float* b = is_b ? bb : ba;
      bias_dst[n] = b ? b[co] : 0.f;
    }
  }
}

// Sum the split-K partials of a weight gradient and scatter them back to OIHW.
//   part[split][tap][n][ci]  ->  dW(n)[co(n)][ci][tap]
__global__ void unpack_dw_kernel(const float* __restrict__ part, int nsplit, int Co, int Ci, int two,
                                 float* __restrict__ dwa, float* __restrict__ dwb) {
  const int Ntot = two ? 2 * Co : Co;
  const long long per = 9LL * Ntot * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Ci); const int n = (int)((i / Ci) % Ntot); const int tap = (int)(i / ((long long)Ci * Ntot));
    float acc = 0.f;
    for (int s = 0; s < nsplit; ++s) acc += part[(size_t)s * per + i];
    int co = n, is_b = 0;
    if (two) { is_b = (n >> 3) & 1; co = ((n >> 4) << 3) + (n & 7); }
    float* dst = is_b ? dwb : dwa;
    dst[((size_t)co * Ci + ci) * 9 + tap] = acc;
  }
}

// Channels-last weights [Co][3][3][Ci] (the convolutions of a channels_last module).  With a second
// source (gamma, beta) the combined output index n follows gb8_col, as in pack_w3x3_kernel.
//   forward : dst[tap][n][ci] = W(n)[co(n)][tap][ci]                (row copies)
__global__ void pack_w3x3_cl_fwd_kernel(const float* __restrict__ wa, const float* __restrict__ wb,
                                        const float* __restrict__ ba, const float* __restrict__ bb, int Co, int Ci,
                                        int round_ops, float* __restrict__ dst, float* __restrict__ bias_dst) {
  const int Ntot = wb ? 2 * Co : Co;
  const int Ci4 = Ci >> 2;
  const long long total = 9LL * Ntot * Ci4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % Ci4); const int n = (int)((i / Ci4) % Ntot); const int tap = (int)(i / ((long long)Ci4 * Ntot));
    int co = n, is_b = 0;
    if (wb) { is_b = (n >> 3) & 1; co = ((n >> 4) << 3) + (n & 7); }
    float4 v = *reinterpret_cast<const float4*>((is_b ? wb : wa) + ((size_t)co * 9 + tap) * Ci + 4 * c4);
    if (round_ops) v = make_float4(elem_round_tf32(v.x), elem_round_tf32(v.y), elem_round_tf32(v.z), elem_round_tf32(v.w));
    *reinterpret_cast<float4*>(dst + ((size_t)tap * Ntot + n) * Ci + 4 * c4) = v;
  }
  if (bias_dst != nullptr && blockIdx.x == 0) {
    for (int n = threadIdx.x; n < Ntot; n += blockDim.x) {
      int co = n, is_b = 0;
      if (wb) { is_b = (n >> 3) & 1; co = ((n >> 4) << 3) + (n & 7); }
      const float* bsrc = is_b ? bb : ba;
      bias_dst[n] = bsrc ? bsrc[co] : 0.f;
    }
  }
}
//   dgrad   : dst[tap][ci][k] = W(k)[co(k)][8 - tap][ci]             (32x32 tile transposes through smem)
__global__ void __launch_bounds__(256) pack_w3x3_cl_dgrad_kernel(const float* __restrict__ wa, const float* __restrict__ wb,
                                                                 int Co, int Ci, int round_ops, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int Ntot = wb ? 2 * Co : Co;
  const int tap = blockIdx.z, k0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, ci = ci0 + tx;
    int co = k, is_b = 0;
    if (wb) { is_b = (k >> 3) & 1; co = ((k >> 4) << 3) + (k & 7); }
    tile[r][tx] = (k < Ntot && ci < Ci) ? (is_b ? wb : wa)[((size_t)co * 9 + (8 - tap)) * Ci + ci] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int ci = ci0 + r, k = k0 + tx;
    if (ci < Ci && k < Ntot) {
      const float v = tile[tx][r];
      dst[((size_t)tap * Ci + ci) * Ntot + k] = round_ops ? elem_round_tf32(v) : v;
    }
  }
}
// part[split][tap][n][ci] -> dW(n)[co(n)][tap][ci] (channels-last weight gradients)
__global__ void unpack_dw_cl_kernel(const float* __restrict__ part, int nsplit, int Co, int Ci, int two,
                                    float* __restrict__ dwa, float* __restrict__ dwb) {
  const int Ntot = two ? 2 * Co : Co;
  const int Ci4 = Ci >> 2;
  const long long per4 = 9LL * Ntot * Ci4, per = 9LL * Ntot * Ci;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per4; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % Ci4); const int n = (int)((i / Ci4) % Ntot); const int tap = (int)(i / ((long long)Ci4 * Ntot));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < nsplit; ++s) {
      const float4 t = *reinterpret_cast<const float4*>(part + (size_t)s * per + 4 * i);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    int co = n, is_b = 0;
    if (two) { is_b = (n >> 3) & 1; co = ((n >> 4) << 3) + (n & 7); }
    *reinterpret_cast<float4*>((is_b ? dwb : dwa) + ((size_t)co * 9 + tap) * Ci + 4 * c4) = acc;
  }
}

// gb8-ordered per-column sums (double) -> the two bias gradients
__global__ void unpack_db_kernel(const double* __restrict__ s_beta, const double* __restrict__ s_gamma, int C,
                                 float* __restrict__ db_gamma, float* __restrict__ db_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  db_gamma[c] = (float)s_gamma[c];
  db_beta[c] = (float)s_beta[c];
}

__global__ void round_tf32_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_float4(elem_round_tf32(v.x), elem_round_tf32(v.y), elem_round_tf32(v.z), elem_round_tf32(v.w));
  }
}

// dst[i] = sum over groups of src[g * gstride + i]
__global__ void double_to_float_kernel(const double* __restrict__ src, int n, int groups, long long gstride,
                                       float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc = 0.0;
  for (int g = 0; g < groups; ++g) acc += src[(size_t)g * gstride + i];
  dst[i] = (float)acc;
}

// Affine batch norm + LeakyReLU (the conv -> SyncBN -> LeakyReLU(0.2) stages of the flow network and
// of conv_dim_in, normalization.py:16-50 / flows_generator.py):  y = act((x - mean_g) * rstd_g * w + b)
__global__ void __launch_bounds__(kElemThreads)
bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ weight, const float* __restrict__ bias, long long P, int C, float slope,
                  float* __restrict__ y) {
  const long long n4 = P * (C >> 2);
  x += (size_t)blockIdx.y * P * C; y += (size_t)blockIdx.y * P * C;
  mean += (size_t)blockIdx.y * C; rstd += (size_t)blockIdx.y * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % (C >> 2)) << 2;
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 w = *reinterpret_cast<const float4*>(weight + c);
    const float4 b = *reinterpret_cast<const float4*>(bias + c);
    float4 o;
    o.x = (xv.x - mu.x) * rs.x * w.x + b.x; o.y = (xv.y - mu.y) * rs.y * w.y + b.y;
    o.z = (xv.z - mu.z) * rs.z * w.z + b.z; o.w = (xv.w - mu.w) * rs.w * w.w + b.w;
    if (slope != 1.f) {
      o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
      o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
    }
    reinterpret_cast<float4*>(y)[i] = o;
  }
}

}  // namespace ag2v

using namespace ag2v;

static int chan_check(long long P, int C) {
  AG2V_REQUIRE(P > 0 && C > 0 && C % 4 == 0, "per-channel reduction: need P > 0 and C %% 4 == 0 (P=%lld C=%d)", P, C);
  return AG2V_OK;
}

// Floats needed for the partial buffer of a per-channel reduction with NS sums.
extern "C" size_t ag2v_chan_partial_floats(long long P, int C, int NS) {
  if (P <= 0 || C <= 0) return 0;
  ChanGeom g = chan_geom(P, C);
  return (size_t)g.nsplit * NS * C;
}

// sums[0..C) = sum x, sums[C..2C) = sum x^2 over the P rows of x [P, C] (doubles).
// x is [groups][P][C]; partial holds groups * ag2v_chan_partial_floats(P, C, 2) floats, sums [groups][2C].
extern "C" int ag2v_bn_stats(const float* x, long long P, int C, int groups, float* partial, double* sums,
                             cudaStream_t stream) {
  int rc = chan_check(P, C);
  if (rc) return rc;
  AG2V_REQUIRE(x && partial && sums && groups >= 1 && groups <= 65535, "bn_stats: null pointer or bad group count");
  ChanGeom g = chan_geom(P, C);
  ChanArgs a{};
  a.x = x; a.P = P; a.C = C; a.part = partial;
  chan_partial_kernel<0><<<dim3(g.nsplit, groups), kElemThreads, 0, stream>>>(a, g);
  AG2V_LAUNCH_CHECK();
  chan_reduce_kernel<<<dim3(ceil_div(2 * C, 32), groups), 256, 0, stream>>>(partial, g.nsplit, 2 * C, sums);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// mean/rstd from (possibly all-reduced) sums; updates running stats when given.
// sums [groups][2C] -> mean / rstd [groups][C]; count = elements per channel in ONE group.
extern "C" int ag2v_bn_finalize(const double* sums, double count, double unbias_count, int C, int groups, float eps,
                                float momentum, const float* in_scale, float* running_mean, float* running_var,
                                float* mean, float* rstd, cudaStream_t stream) {
  AG2V_REQUIRE(sums && mean && rstd && C > 0 && count > 0 && groups >= 1, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<ceil_div(C, 256), 256, 0, stream>>>(sums, count, unbias_count > 0 ? unbias_count : count, C, groups,
                                                          eps, momentum, in_scale, running_mean, running_var, mean, rstd);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

extern "C" int ag2v_bn_eval_stats(const float* running_mean, const float* running_var, int C, int groups, float eps,
                                  const float* in_scale, float* mean, float* rstd, cudaStream_t stream) {
  AG2V_REQUIRE(running_mean && running_var && mean && rstd && C > 0 && groups >= 1, "bn_eval_stats: bad arguments");
  bn_eval_stats_kernel<<<ceil_div(C, 256), 256, 0, stream>>>(running_mean, running_var, C, groups, eps, in_scale, mean, rstd);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// SPADE backward, element-wise pass 1.  Inputs [P, C] NHWC: dout, out (post
// activation), x, gamma (saved by the forward), mean/rstd [C].  Writes dgb [P, 2C]
// (gb8 columns: d gamma = g*xhat, d beta = g), dxhat [P, C] and the four
// per-channel sums (doubles, [4][C]): sum g, sum g*xhat, sum dxhat, sum dxhat*xhat.
extern "C" int ag2v_spade_bwd_pre(const float* dout, const float* out, const float* x, const float* gamma,
                                  const float* mean, const float* rstd, long long P, int C, int groups, int act,
                                  float slope, int round_ops, int chan_gamma, int up_h, int up_w, float* dgb,
                                  float* dxhat, float* partial, double* sums, cudaStream_t stream) {
  int rc = chan_check(P, C);
  if (rc) return rc;
  AG2V_REQUIRE(dgb == nullptr || C % 8 == 0, "spade_bwd_pre: C %% 8 == 0 required (C=%d)", C);
  AG2V_REQUIRE(dout && out && x && gamma && mean && rstd && dxhat && partial && sums && (dgb || chan_gamma),
               "spade_bwd_pre: null pointer");
  AG2V_REQUIRE(groups >= 1 && groups <= 65535, "spade_bwd_pre: bad group count %d", groups);
  ChanGeom g = chan_geom(P, C);
  ChanArgs a{};
  a.x = x; a.P = P; a.C = C; a.part = partial; a.dout = dout; a.out = out; a.gamma = gamma;
  a.mean = mean; a.rstd = rstd; a.dgb = dgb; a.dxhat = dxhat; a.slope = slope; a.act = act; a.round_ops = round_ops;
  a.chan_gamma = chan_gamma;
  AG2V_REQUIRE((up_h == 0 && up_w == 0) || (up_h > 0 && up_w > 0 && up_h % 2 == 0 && up_w % 2 == 0 && P % ((long long)up_h * up_w) == 0),
               "spade_bwd_pre: bad up-sampling geometry %dx%d for %lld pixels", up_h, up_w, P);
  a.up_h = up_h; a.up_w = up_w;
  chan_partial_kernel<1><<<dim3(g.nsplit, groups), kElemThreads, 0, stream>>>(a, g);
  AG2V_LAUNCH_CHECK();
  chan_reduce_kernel<<<dim3(ceil_div(5 * C, 32), groups), 256, 0, stream>>>(partial, g.nsplit, 5 * C, sums);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// SPADE backward pass 2: dxhat -> dx in place (sums as produced by spade_bwd_pre,
// all-reduced across ranks by the caller under SyncBN; count = global element count).
extern "C" int ag2v_spade_bwd_dx(const float* x, float* dxhat, const float* mean, const float* rstd,
                                 const double* sums, double count, int training, long long P, int C, int groups,
                                 int up_h, int up_w, float* dx_low, cudaStream_t stream) {
  int rc = chan_check(P, C);
  if (rc) return rc;
  AG2V_REQUIRE(x && dxhat && mean && rstd && sums && groups >= 1 && groups <= 65535, "spade_bwd_dx: bad arguments");
  if (up_w > 0) {      // x is low resolution: P full-resolution pixels per group, P / 4 parents
    AG2V_REQUIRE(dx_low && up_h > 0 && up_h % 2 == 0 && up_w % 2 == 0 && P % ((long long)up_h * up_w) == 0,
                 "spade_bwd_dx: bad up-sampling geometry");
    const long long n4l = (P / 4) * (C / 4);
    const long long capl = 4 * 148 * 8 / groups + 1;
    int bl = (int)(ceil_div_ll(n4l, kElemThreads) > capl ? capl : ceil_div_ll(n4l, kElemThreads));
    if (bl < 1) bl = 1;
    spade_bwd_dx_up_kernel<<<dim3(bl, groups), kElemThreads, 0, stream>>>(x, dxhat, mean, rstd, sums, count, training, P / 4,
                                                                         C, up_h, up_w, dx_low);
    AG2V_LAUNCH_CHECK();
    return AG2V_OK;
  }
  long long n4 = P * (C / 4);
  const long long cap = 4 * 148 * 8 / groups + 1;
  int blocks = (int)(ceil_div_ll(n4, kElemThreads * 4) > cap ? cap : ceil_div_ll(n4, kElemThreads * 4));
  if (blocks < 1) blocks = 1;
  spade_bwd_dx_kernel<<<dim3(blocks, groups), kElemThreads, 0, stream>>>(x, dxhat, mean, rstd, sums, count, training, P, C);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// wa/wb: OIHW [Co, Ci, 3, 3] (wb null for a single conv); see pack_w3x3_kernel.
extern "C" int ag2v_pack_w3x3(const float* wa, const float* wb, const float* ba, const float* bb, int Co, int Ci,
                              int dgrad, int round_ops, float* dst, float* bias_dst, cudaStream_t stream) {
  AG2V_REQUIRE(wa && dst && Co > 0 && Ci > 0, "pack_w3x3: bad arguments");
  AG2V_REQUIRE(!wb || Co % 8 == 0, "pack_w3x3: gamma/beta packing needs Co %% 8 == 0");
  long long total = 9LL * (wb ? 2 * Co : Co) * Ci;
  int blocks = (int)(ceil_div_ll(total, 256) > 2368 ? 2368 : ceil_div_ll(total, 256));
  pack_w3x3_kernel<<<blocks, 256, 0, stream>>>(wa, wb, ba, bb, Co, Ci, dgrad, round_ops, dst, bias_dst);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

extern "C" int ag2v_unpack_dw3x3(const float* part, int nsplit, int Co, int Ci, int two, float* dwa, float* dwb,
                                 cudaStream_t stream) {
  AG2V_REQUIRE(part && dwa && nsplit > 0 && Co > 0 && Ci > 0 && (!two || dwb), "unpack_dw3x3: bad arguments");
  long long total = 9LL * (two ? 2 * Co : Co) * Ci;
  int blocks = (int)(ceil_div_ll(total, 256) > 2368 ? 2368 : ceil_div_ll(total, 256));
  unpack_dw_kernel<<<blocks, 256, 0, stream>>>(part, nsplit, Co, Ci, two, dwa, dwb);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// dst = round-to-nearest-TF32(src), same layout, n % 4 == 0 (src == dst allowed)
extern "C" int ag2v_round_tf32(const float* src, float* dst, long long n, cudaStream_t stream) {
  AG2V_REQUIRE(src && dst && n > 0 && n % 4 == 0, "round_tf32: bad arguments");
  AG2V_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, "round_tf32: 16-byte alignment required");
  long long n4 = n / 4;
  int blocks = (int)(ceil_div_ll(n4, 256 * 4) > 148 * 16 ? 148 * 16 : ceil_div_ll(n4, 256 * 4));
  round_tf32_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst), n4);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// Gradient of y = conv(x, W) * scale[group] + bias w.r.t. the convolution output, as the operand of
// the input / weight gradient GEMMs: dys = tf32(dy * scale[group]) ([groups][P][C] NHWC, scale may
// be NULL), sums[g][0..C) = sum dy (the bias gradient).  partial: groups * chan_partial_floats(P, C, 1).
extern "C" int ag2v_scaled_grad_pre(const float* dy, const float* scale, long long P, int C, int groups, float* dys,
                                    float* partial, double* sums, cudaStream_t stream) {
  int rc = chan_check(P, C);
  if (rc) return rc;
  AG2V_REQUIRE(dy && dys && partial && sums && groups >= 1 && groups <= 65535, "scaled_grad_pre: bad arguments");
  ChanGeom g = chan_geom(P, C);
  ChanArgs a{};
  a.P = P; a.C = C; a.part = partial; a.dout = dy; a.dxhat = dys; a.gamma = scale;
  chan_partial_kernel<3><<<dim3(g.nsplit, groups), kElemThreads, 0, stream>>>(a, g);
  AG2V_LAUNCH_CHECK();
  chan_reduce_kernel<<<dim3(ceil_div(C, 32), groups), 256, 0, stream>>>(partial, g.nsplit, C, sums);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// Channels-last 3x3 weights [Co][3][3][Ci] -> [9][N][Ci] (dgrad = 0) or [9][Ci][N] flipped (dgrad = 1);
// wb != NULL: (wa, wb) = (mlp_gamma, mlp_beta) interleaved in gb8 order, N = 2 Co; bias_dst (dgrad = 0,
// optional) receives the biases in the same column order.
extern "C" int ag2v_pack_w3x3_cl(const float* wa, const float* wb, const float* ba, const float* bb, int Co, int Ci,
                                 int dgrad, int round_ops, float* dst, float* bias_dst, cudaStream_t stream) {
  AG2V_REQUIRE(wa && dst && Co > 0 && Ci > 0 && Ci % 4 == 0, "pack_w3x3_cl: bad arguments (Ci %% 4 == 0 required)");
  AG2V_REQUIRE(!wb || Co % 8 == 0, "pack_w3x3_cl: gamma/beta packing needs Co %% 8 == 0");
  const int Ntot = wb ? 2 * Co : Co;
  if (!dgrad) {
    const long long total = 9LL * Ntot * (Ci / 4);
    const int blocks = (int)(ceil_div_ll(total, 256) > 2368 ? 2368 : ceil_div_ll(total, 256));
    pack_w3x3_cl_fwd_kernel<<<blocks, 256, 0, stream>>>(wa, wb, ba, bb, Co, Ci, round_ops, dst, bias_dst);
  } else {
    pack_w3x3_cl_dgrad_kernel<<<dim3(ceil_div(Ci, 32), ceil_div(Ntot, 32), 9), 256, 0, stream>>>(wa, wb, Co, Ci, round_ops, dst);
  }
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// split-K partials [nsplit][9][N][Ci] -> channels-last weight gradient(s) [Co][3][3][Ci]
extern "C" int ag2v_unpack_dw3x3_cl(const float* part, int nsplit, int Co, int Ci, int two, float* dwa, float* dwb,
                                    cudaStream_t stream) {
  AG2V_REQUIRE(part && dwa && nsplit > 0 && Co > 0 && Ci > 0 && Ci % 4 == 0 && (!two || dwb), "unpack_dw3x3_cl: bad arguments");
  const long long total = 9LL * (two ? 2 * Co : Co) * (Ci / 4);
  const int blocks = (int)(ceil_div_ll(total, 256) > 2368 ? 2368 : ceil_div_ll(total, 256));
  unpack_dw_cl_kernel<<<blocks, 256, 0, stream>>>(part, nsplit, Co, Ci, two, dwa, dwb);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// dst[i] = (float) sum_g src[g * group_stride + i], i < n
extern "C" int ag2v_double_to_float(const double* src, int n, int groups, long long group_stride, float* dst,
                                    cudaStream_t stream) {
  AG2V_REQUIRE(src && dst && n > 0 && groups >= 1, "double_to_float: bad arguments");
  double_to_float_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(src, n, groups, group_stride, dst);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// y = act((x - mean_g) * rstd_g * weight + bias) on [groups][P][C] (NHWC), slope 1 = no activation
extern "C" int ag2v_bn_act_fwd(const float* x, const float* mean, const float* rstd, const float* weight,
                               const float* bias, long long P, int C, int groups, float slope, float* y,
                               cudaStream_t stream) {
  int rc = chan_check(P, C);
  if (rc) return rc;
  AG2V_REQUIRE(x && mean && rstd && weight && bias && y && groups >= 1 && groups <= 65535, "bn_act_fwd: bad arguments");
  long long n4 = P * (C / 4);
  const long long cap = 4 * 148 * 8 / groups + 1;
  int blocks = (int)(ceil_div_ll(n4, kElemThreads * 4) > cap ? cap : ceil_div_ll(n4, kElemThreads * 4));
  if (blocks < 1) blocks = 1;
  bn_act_fwd_kernel<<<dim3(blocks, groups), kElemThreads, 0, stream>>>(x, mean, rstd, weight, bias, P, C, slope, y);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}
