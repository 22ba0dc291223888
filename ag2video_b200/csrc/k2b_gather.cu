// K2b — the two true bilinear gathers of the path:
//   masks_to_layout (models/layout.py:66-95 + _pool_mask_samples :164-202)
//   crop_bbox       (models/bilinear.py:102-131 + tensor_linspace :192-221)
// Index arithmetic replays the reference op by op (no FMA contraction) so the tap
// indices floor(ix), floor(iy) are bit-exact; taps are accumulated nw, ne, sw, se.
#include "common.cuh"

namespace ag2v {

struct Taps { int x0, y0; float nw, ne, sw, se; bool in_x0, in_x1, in_y0, in_y1; };

// grid_sample(align_corners=True, zeros padding) taps for normalised coords (gx, gy)
__device__ __forceinline__ Taps make_taps(float gx, float gy, int Wsrc, int Hsrc) {
  Taps t;
  const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(gx, 1.f), 2.f), (float)(Wsrc - 1));
  const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(gy, 1.f), 2.f), (float)(Hsrc - 1));
  const float fx = floorf(ix), fy = floorf(iy);
  const float wx1 = __fsub_rn(ix, fx), wx0 = __fsub_rn(__fadd_rn(fx, 1.f), ix);
  const float wy1 = __fsub_rn(iy, fy), wy0 = __fsub_rn(__fadd_rn(fy, 1.f), iy);
  t.nw = __fmul_rn(wx0, wy0); t.ne = __fmul_rn(wx1, wy0);
  t.sw = __fmul_rn(wx0, wy1); t.se = __fmul_rn(wx1, wy1);
  t.in_x0 = fx >= 0.f && fx <= (float)(Wsrc - 1);
  t.in_x1 = fx >= -1.f && fx <= (float)(Wsrc - 2);
  t.in_y0 = fy >= 0.f && fy <= (float)(Hsrc - 1);
  t.in_y1 = fy >= -1.f && fy <= (float)(Hsrc - 2);
  // NaN / inf coordinates fail every test above; the integer values are then unused
  t.x0 = (t.in_x0 || t.in_x1) ? (int)fx : 0;
  t.y0 = (t.in_y0 || t.in_y1) ? (int)fy : 0;
  return t;
}

__device__ __forceinline__ float sample4(const float* __restrict__ src, int Wsrc, const Taps& t) {
  float acc = 0.f;
  if (t.in_y0 && t.in_x0) acc = __fadd_rn(acc, __fmul_rn(src[t.y0 * Wsrc + t.x0], t.nw));
  if (t.in_y0 && t.in_x1) acc = __fadd_rn(acc, __fmul_rn(src[t.y0 * Wsrc + t.x0 + 1], t.ne));
  if (t.in_y1 && t.in_x0) acc = __fadd_rn(acc, __fmul_rn(src[(t.y0 + 1) * Wsrc + t.x0], t.sw));
  if (t.in_y1 && t.in_x1) acc = __fadd_rn(acc, __fmul_rn(src[(t.y0 + 1) * Wsrc + t.x0 + 1], t.se));
  return acc;
}

// S[o, y, x] = grid_sample(mask_o)[y, x] with the grid of _boxes_to_grid (layout.py:98-130)
__global__ void mask_sample_kernel(const float* __restrict__ boxes, const float* __restrict__ masks,
                                   const float* __restrict__ lin_x, const float* __restrict__ lin_y, int O, int M,
                                   int H, int W, float* __restrict__ S) {
  const long long total = (long long)O * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), o = (int)(i / ((long long)W * H));
    const float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)o * 4);
    const float gx = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(lin_x[x], bx.x), bx.z), 2.f), 1.f);
    const float gy = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(lin_y[y], bx.y), bx.w), 2.f), 1.f);
    const Taps t = make_taps(gx, gy, M, M);
    S[i] = sample4(masks + (size_t)o * M * M, M, t);
  }
}

// train mode: out[d, y, x] = sum_o v[o, d] * S[o, y, x]   (object order)
__global__ void __launch_bounds__(256)
dense_compose_kernel(const float* __restrict__ vecs, const float* __restrict__ S, int O, int D, int HW,
                     float* __restrict__ out) {
  const int d = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int o = 0; o < O; ++o) acc = fmaf(vecs[(size_t)o * D + d], S[(size_t)o * HW + i], acc);
    out[(size_t)d * HW + i] = acc;
  }
}

// dvecs[o, d] = sum_i dout[d, i] * S[o, i]; one block per (o, d), fixed-order reduction
__global__ void __launch_bounds__(256)
dense_compose_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ S, int O, int D, int HW,
                         float* __restrict__ dvecs) {
  __shared__ float red[8];
  const int o = blockIdx.x / D, d = blockIdx.x - o * D;
  float acc = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) acc = fmaf(dout[(size_t)d * HW + i], S[(size_t)o * HW + i], acc);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    dvecs[blockIdx.x] = s;
  }
}

// ---- gradients of masks_to_layout with respect to the masks and the boxes (autograd of layout.py:66-95) ----
// G[o, i] = sum_d v[o, d] * dout[d, i]: the gradient arriving at the sampled plane S[o]
__global__ void __launch_bounds__(256)
mask_plane_grad_kernel(const float* __restrict__ dout, const float* __restrict__ vecs, int O, int D, int HW,
                       float* __restrict__ G) {
  const int o = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(dout[(size_t)d * HW + i], vecs[(size_t)o * D + d], acc);
    G[(size_t)o * HW + i] = acc;
  }
}

// One block per object: scatters G through the four taps into dmasks (atomics: several pixels share a tap, the
// summation order is not fixed, as in torch's grid_sampler backward) and reduces the grid gradient to dboxes.
// grid gradient as grid_sampler_2d_backward: gix = sum over taps of (+-) value * y-weight, scaled by (M - 1) / 2.
__global__ void __launch_bounds__(256)
mask_sample_bwd_kernel(const float* __restrict__ G, const float* __restrict__ boxes, const float* __restrict__ masks,
                       const float* __restrict__ lin_x, const float* __restrict__ lin_y, int M, int H, int W,
                       float* __restrict__ dmasks, float* __restrict__ dboxes) {
  __shared__ float red[8][4];
  const int o = blockIdx.x;
  const float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)o * 4);
  const float* mk = masks + (size_t)o * M * M;
  float* dm = dmasks ? dmasks + (size_t)o * M * M : nullptr;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float g = G[(size_t)o * H * W + i];
    const float gxn = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(lin_x[x], bx.x), bx.z), 2.f), 1.f);
    const float gyn = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(lin_y[y], bx.y), bx.w), 2.f), 1.f);
    const Taps t = make_taps(gxn, gyn, M, M);
    const bool a = t.in_y0 && t.in_x0, b = t.in_y0 && t.in_x1, c = t.in_y1 && t.in_x0, d = t.in_y1 && t.in_x1;
    if (!(a || b || c || d)) continue;
    if (dm) {
      if (a) atomicAdd(dm + t.y0 * M + t.x0, g * t.nw);
      if (b) atomicAdd(dm + t.y0 * M + t.x0 + 1, g * t.ne);
      if (c) atomicAdd(dm + (t.y0 + 1) * M + t.x0, g * t.sw);
      if (d) atomicAdd(dm + (t.y0 + 1) * M + t.x0 + 1, g * t.se);
    }
    if (dboxes) {
      const float nw = a ? mk[t.y0 * M + t.x0] : 0.f, ne = b ? mk[t.y0 * M + t.x0 + 1] : 0.f;
      const float sw = c ? mk[(t.y0 + 1) * M + t.x0] : 0.f, se = d ? mk[(t.y0 + 1) * M + t.x0 + 1] : 0.f;
      // the separable weights back out of the products: nw = wx0 * wy0, ne = wx1 * wy0, sw = wx0 * wy1 (wx0 + wx1 = 1)
      const float wy0 = t.nw + t.ne, wy1 = t.sw + t.se, wx0 = t.nw + t.sw, wx1 = t.ne + t.se;
      const float gix = g * ((ne - nw) * wy0 + (se - sw) * wy1) * (float)(M - 1);   // dL / d(normalised X) = gix * (M-1)/2 * 2
      const float giy = g * ((sw - nw) * wx0 + (se - ne) * wx1) * (float)(M - 1);
      acc[0] -= gix / bx.z;
      acc[2] -= gix * (lin_x[x] - bx.x) / (bx.z * bx.z);
      acc[1] -= giy / bx.w;
      acc[3] -= giy * (lin_y[y] - bx.y) / (bx.w * bx.w);
    }
  }
  if (!dboxes) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 4; ++k) red[threadIdx.x >> 5][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    dboxes[(size_t)o * 4 + threadIdx.x] = s;
  }
}

// test mode (layout.py:185-197): objects in ascending order of their sampled mass
// paint the pixels where their clean mask > 0.5 and nothing has painted before.
__global__ void mask_order_kernel(const float* __restrict__ vecs, const float* __restrict__ S, int O, int D, int HW,
                                  int* __restrict__ order) {
  extern __shared__ double mass[];
  for (int o = threadIdx.x >> 5; o < O; o += blockDim.x >> 5) {
    const int lane = threadIdx.x & 31;
    double sv = 0.0, ss = 0.0;
    for (int d = lane; d < D; d += 32) sv += (double)vecs[(size_t)o * D + d];
    for (int i = lane; i < HW; i += 32) ss += (double)S[(size_t)o * HW + i];
    for (int k = 16; k > 0; k >>= 1) { sv += __shfl_xor_sync(0xffffffffu, sv, k); ss += __shfl_xor_sync(0xffffffffu, ss, k); }
    if (lane == 0) mass[o] = sv * ss;
  }
  __syncthreads();
  if (threadIdx.x == 0) {           // stable insertion sort, O is tiny
    for (int o = 0; o < O; ++o) order[o] = o;
    for (int i = 1; i < O; ++i) {
      const int k = order[i]; int j = i - 1;
      while (j >= 0 && mass[order[j]] > mass[k]) { order[j + 1] = order[j]; --j; }
      order[j + 1] = k;
    }
  }
}

__global__ void __launch_bounds__(256)
mask_paint_kernel(const float* __restrict__ vecs, const float* __restrict__ S, const int* __restrict__ order, int O,
                  int D, int HW, float* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    int who = -1;
    for (int k = 0; k < O && who < 0; ++k) { const int o = order[k]; if (S[(size_t)o * HW + i] > 0.5f) who = o; }
    const float s = who >= 0 ? S[(size_t)who * HW + i] : 0.f;
    for (int d = 0; d < D; ++d) out[(size_t)d * HW + i] = who >= 0 ? vecs[(size_t)who * D + d] * s : 0.f;
  }
}

// crop_bbox: crops[n, c, i, j] = bilinear(feats[frame[n]], X_j, Y_i),  box xywh -> points -> [-1,1];
// X_j = start_w[j] * x0 + end_w[j] * x1 as in tensor_linspace (bilinear.py:212-220)
__global__ void crop_fwd_kernel(const float* __restrict__ feats, const int* __restrict__ frame,
                                const float* __restrict__ boxes, const float* __restrict__ ws_x,
                                const float* __restrict__ we_x, const float* __restrict__ ws_y,
                                const float* __restrict__ we_y, int n_crops, int C, int H, int W, int HH, int WW,
                                float* __restrict__ out) {
  const long long total = (long long)n_crops * HH * WW;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(q % WW), i = (int)((q / WW) % HH), n = (int)(q / ((long long)WW * HH));
    const float* box = boxes + (size_t)n * 4;
    float gx, gy;
    {
      const float x0 = __fsub_rn(__fmul_rn(2.f, box[0]), 1.f), y0 = __fsub_rn(__fmul_rn(2.f, box[1]), 1.f);
      const float x1 = __fsub_rn(__fmul_rn(2.f, __fadd_rn(box[0], box[2])), 1.f);
      const float y1 = __fsub_rn(__fmul_rn(2.f, __fadd_rn(box[1], box[3])), 1.f);
      gx = __fadd_rn(__fmul_rn(ws_x[j], x0), __fmul_rn(we_x[j], x1));
      gy = __fadd_rn(__fmul_rn(ws_y[i], y0), __fmul_rn(we_y[i], y1));
    }
    const Taps t = make_taps(gx, gy, W, H);
    const float* src = feats + (size_t)frame[n] * C * H * W;
    for (int c = 0; c < C; ++c)
      out[(((size_t)n * C + c) * HH + i) * WW + j] = sample4(src + (size_t)c * H * W, W, t);
  }
}

// input gradient of crop_bbox; crops overlap, so contributions are added atomically
// (same as torch's grid_sampler backward; the path has no live caller in training).
__global__ void crop_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ frame,
                                const float* __restrict__ boxes, const float* __restrict__ ws_x,
                                const float* __restrict__ we_x, const float* __restrict__ ws_y,
                                const float* __restrict__ we_y, int n_crops, int C, int H, int W, int HH, int WW,
                                float* __restrict__ dfeats) {
  const long long total = (long long)n_crops * HH * WW;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(q % WW), i = (int)((q / WW) % HH), n = (int)(q / ((long long)WW * HH));
    const float* box = boxes + (size_t)n * 4;
    const float x0 = __fsub_rn(__fmul_rn(2.f, box[0]), 1.f), y0 = __fsub_rn(__fmul_rn(2.f, box[1]), 1.f);
    const float x1 = __fsub_rn(__fmul_rn(2.f, __fadd_rn(box[0], box[2])), 1.f);
    const float y1 = __fsub_rn(__fmul_rn(2.f, __fadd_rn(box[1], box[3])), 1.f);
    const float gx = __fadd_rn(__fmul_rn(ws_x[j], x0), __fmul_rn(we_x[j], x1));
    const float gy = __fadd_rn(__fmul_rn(ws_y[i], y0), __fmul_rn(we_y[i], y1));
    const Taps t = make_taps(gx, gy, W, H);
    float* dst = dfeats + (size_t)frame[n] * C * H * W;
    for (int c = 0; c < C; ++c) {
      const float g = dout[(((size_t)n * C + c) * HH + i) * WW + j];
      float* pl = dst + (size_t)c * H * W;
      if (t.in_y0 && t.in_x0) atomicAdd(pl + t.y0 * W + t.x0, g * t.nw);
      if (t.in_y0 && t.in_x1) atomicAdd(pl + t.y0 * W + t.x0 + 1, g * t.ne);
      if (t.in_y1 && t.in_x0) atomicAdd(pl + (t.y0 + 1) * W + t.x0, g * t.sw);
      if (t.in_y1 && t.in_x1) atomicAdd(pl + (t.y0 + 1) * W + t.x0 + 1, g * t.se);
    }
  }
}

// backend='jj' of crop_bbox (bilinear.py:134-189 bilinear_sample): coordinates stay in [0, 1] and are scaled by the
// source size, taps are CLAMPED to the image (no zero padding), weights are the distances to the clamped taps, and the
// four products are summed in the reference's order w1*v1 + w2*v2 + w3*v3 + w4*v4 (v2 is the tap BELOW, v3 to the right).
struct TapsJJ { int x0, x1, y0, y1; float w1, w2, w3, w4; };

__device__ __forceinline__ TapsJJ make_taps_jj(float X, float Y, int Wsrc, int Hsrc) {
  TapsJJ t;
  X = __fmul_rn(X, (float)Wsrc); Y = __fmul_rn(Y, (float)Hsrc);
  const float x0 = fminf(fmaxf(floorf(X), 0.f), (float)(Wsrc - 1)), x1 = fminf(fmaxf(__fadd_rn(x0, 1.f), 0.f), (float)(Wsrc - 1));
  const float y0 = fminf(fmaxf(floorf(Y), 0.f), (float)(Hsrc - 1)), y1 = fminf(fmaxf(__fadd_rn(y0, 1.f), 0.f), (float)(Hsrc - 1));
  t.w1 = __fmul_rn(__fsub_rn(x1, X), __fsub_rn(y1, Y)); t.w2 = __fmul_rn(__fsub_rn(x1, X), __fsub_rn(Y, y0));
  t.w3 = __fmul_rn(__fsub_rn(X, x0), __fsub_rn(y1, Y)); t.w4 = __fmul_rn(__fsub_rn(X, x0), __fsub_rn(Y, y0));
  // NaN coordinates: fmaxf / fminf return the non-NaN operand, so the taps stay inside the image (the weights are NaN,
  // as in the reference)
  t.x0 = (int)x0; t.x1 = (int)x1; t.y0 = (int)y0; t.y1 = (int)y1;
  return t;
}

__device__ __forceinline__ void crop_coords_jj(const float* box, float wsx, float wex, float wsy, float wey, float* X, float* Y) {
  const float x1 = __fadd_rn(box[0], box[2]), y1 = __fadd_rn(box[1], box[3]);       // xywh_to_points, no [-1, 1] remap
  *X = __fadd_rn(__fmul_rn(wsx, box[0]), __fmul_rn(wex, x1));
  *Y = __fadd_rn(__fmul_rn(wsy, box[1]), __fmul_rn(wey, y1));
}

__global__ void crop_jj_fwd_kernel(const float* __restrict__ feats, const int* __restrict__ frame,
                                   const float* __restrict__ boxes, const float* __restrict__ ws_x,
                                   const float* __restrict__ we_x, const float* __restrict__ ws_y,
                                   const float* __restrict__ we_y, int n_crops, int C, int H, int W, int HH, int WW,
                                   float* __restrict__ out) {
  const long long total = (long long)n_crops * HH * WW;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(q % WW), i = (int)((q / WW) % HH), n = (int)(q / ((long long)WW * HH));
    float X, Y;
    crop_coords_jj(boxes + (size_t)n * 4, ws_x[j], we_x[j], ws_y[i], we_y[i], &X, &Y);
    const TapsJJ t = make_taps_jj(X, Y, W, H);
    const float* src = feats + (size_t)frame[n] * C * H * W;
    for (int c = 0; c < C; ++c) {
      const float* pl = src + (size_t)c * H * W;
      float acc = __fmul_rn(t.w1, pl[t.y0 * W + t.x0]);
      acc = __fadd_rn(acc, __fmul_rn(t.w2, pl[t.y1 * W + t.x0]));
      acc = __fadd_rn(acc, __fmul_rn(t.w3, pl[t.y0 * W + t.x1]));
      acc = __fadd_rn(acc, __fmul_rn(t.w4, pl[t.y1 * W + t.x1]));
      out[(((size_t)n * C + c) * HH + i) * WW + j] = acc;
    }
  }
}

__global__ void crop_jj_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ frame,
                                   const float* __restrict__ boxes, const float* __restrict__ ws_x,
                                   const float* __restrict__ we_x, const float* __restrict__ ws_y,
                                   const float* __restrict__ we_y, int n_crops, int C, int H, int W, int HH, int WW,
                                   float* __restrict__ dfeats) {
  const long long total = (long long)n_crops * HH * WW;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(q % WW), i = (int)((q / WW) % HH), n = (int)(q / ((long long)WW * HH));
    float X, Y;
    crop_coords_jj(boxes + (size_t)n * 4, ws_x[j], we_x[j], ws_y[i], we_y[i], &X, &Y);
    const TapsJJ t = make_taps_jj(X, Y, W, H);
    float* dst = dfeats + (size_t)frame[n] * C * H * W;
    for (int c = 0; c < C; ++c) {
      const float g = dout[(((size_t)n * C + c) * HH + i) * WW + j];
      float* pl = dst + (size_t)c * H * W;
      atomicAdd(pl + t.y0 * W + t.x0, g * t.w1);
      atomicAdd(pl + t.y1 * W + t.x0, g * t.w2);
      atomicAdd(pl + t.y0 * W + t.x1, g * t.w3);
      atomicAdd(pl + t.y1 * W + t.x1, g * t.w4);
    }
  }
}

static int blocks_for(long long n, int per = 256, int cap = 148 * 16) {
  long long b = ceil_div_ll(n, per);
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace ag2v

using namespace ag2v;

// masks_to_layout for one (clip, frame): vecs [O,D], boxes [O,4] xywh, masks [O,M,M] fp32,
// S = workspace [O,H,W] (kept for the backward), out [D,H,W].  test_mode != 0 composites
// (order = workspace of O ints).  No zero-box filter (layout.py:66-95).
extern "C" int ag2v_masks_to_layout_fwd(const float* vecs, const float* boxes, const float* masks,
                                        const float* lin_x, const float* lin_y, int O, int D, int M, int H, int W,
                                        int test_mode, float* S, int* order, float* out, cudaStream_t stream) {
  AG2V_REQUIRE(O >= 0 && D > 0 && M > 1 && H > 0 && W > 0, "masks_to_layout: bad sizes");
  AG2V_REQUIRE(out && lin_x && lin_y, "masks_to_layout: null pointer");
  if (O == 0) { AG2V_CUDA(cudaMemsetAsync(out, 0, (size_t)D * H * W * sizeof(float), stream)); return AG2V_OK; }
  AG2V_REQUIRE(vecs && boxes && masks && S, "masks_to_layout: null pointer");
  const int HW = H * W;
  mask_sample_kernel<<<blocks_for((long long)O * HW), 256, 0, stream>>>(boxes, masks, lin_x, lin_y, O, M, H, W, S);
  AG2V_LAUNCH_CHECK();
  if (!test_mode) {
    dim3 grid(blocks_for(HW, 256, 64), D);
    dense_compose_kernel<<<grid, 256, 0, stream>>>(vecs, S, O, D, HW, out);
    AG2V_LAUNCH_CHECK();
  } else {
    AG2V_REQUIRE(order, "masks_to_layout: test_mode needs the order workspace");
    mask_order_kernel<<<1, 256, O * sizeof(double), stream>>>(vecs, S, O, D, HW, order);
    AG2V_LAUNCH_CHECK();
    mask_paint_kernel<<<blocks_for(HW), 256, 0, stream>>>(vecs, S, order, O, D, HW, out);
    AG2V_LAUNCH_CHECK();
  }
  return AG2V_OK;
}

extern "C" int ag2v_masks_to_layout_bwd(const float* dout, const float* S, int O, int D, int H, int W, float* dvecs,
                                        cudaStream_t stream) {
  AG2V_REQUIRE(O >= 0 && D > 0 && H > 0 && W > 0, "masks_to_layout_bwd: bad sizes");
  if (O == 0) return AG2V_OK;
  AG2V_REQUIRE(dout && S && dvecs, "masks_to_layout_bwd: null pointer");
  dense_compose_bwd_kernel<<<O * D, 256, 0, stream>>>(dout, S, O, D, H * W, dvecs);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// crop_bbox over a flat list of crops: feats [NF,C,H,W] (NCHW), frame[n] = source image of crop n,
// boxes [n,4] xywh, w_start/w_end = torch.linspace(1,0,steps) / (0,1,steps) for WW and HH.
extern "C" int ag2v_crop_bbox_fwd(const float* feats, const int* frame, const float* boxes, const float* ws_x,
                                  const float* we_x, const float* ws_y, const float* we_y, int n_crops, int C,
                                  int H, int W, int HH, int WW, float* out, cudaStream_t stream) {
  AG2V_REQUIRE(n_crops >= 0 && C > 0 && H > 1 && W > 1 && HH > 0 && WW > 0, "crop_bbox: bad sizes");
  if (n_crops == 0) return AG2V_OK;
  AG2V_REQUIRE(feats && frame && boxes && ws_x && we_x && ws_y && we_y && out, "crop_bbox: null pointer");
  crop_fwd_kernel<<<blocks_for((long long)n_crops * HH * WW), 256, 0, stream>>>(feats, frame, boxes, ws_x, we_x, ws_y,
                                                                              we_y, n_crops, C, H, W, HH, WW, out);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// dfeats must be zero-initialised by the caller
extern "C" int ag2v_crop_bbox_bwd(const float* dout, const int* frame, const float* boxes, const float* ws_x,
                                  const float* we_x, const float* ws_y, const float* we_y, int n_crops, int C,
                                  int H, int W, int HH, int WW, float* dfeats, cudaStream_t stream) {
  AG2V_REQUIRE(n_crops >= 0 && C > 0 && H > 1 && W > 1 && HH > 0 && WW > 0, "crop_bbox_bwd: bad sizes");
  if (n_crops == 0) return AG2V_OK;
  AG2V_REQUIRE(dout && frame && boxes && ws_x && we_x && ws_y && we_y && dfeats, "crop_bbox_bwd: null pointer");
  crop_bwd_kernel<<<blocks_for((long long)n_crops * HH * WW), 256, 0, stream>>>(dout, frame, boxes, ws_x, we_x, ws_y,
                                                                              we_y, n_crops, C, H, W, HH, WW, dfeats);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// Gradients of masks_to_layout (train mode) with respect to the masks [O,M,M] and the xywh boxes [O,4]; either output
// may be null.  G = workspace [O,H,W]; dmasks must be zero-initialised by the caller.
extern "C" int ag2v_masks_to_layout_bwd_inputs(const float* dout, const float* vecs, const float* boxes, const float* masks,
                                               const float* lin_x, const float* lin_y, int O, int D, int M, int H, int W,
                                               float* G, float* dmasks, float* dboxes, cudaStream_t stream) {
  AG2V_REQUIRE(O >= 0 && D > 0 && M > 1 && H > 0 && W > 0, "masks_to_layout_bwd_inputs: bad sizes");
  if (O == 0 || (!dmasks && !dboxes)) return AG2V_OK;
  AG2V_REQUIRE(dout && vecs && boxes && masks && lin_x && lin_y && G, "masks_to_layout_bwd_inputs: null pointer");
  dim3 grid(blocks_for(H * W, 256, 64), O);
  mask_plane_grad_kernel<<<grid, 256, 0, stream>>>(dout, vecs, O, D, H * W, G);
  AG2V_LAUNCH_CHECK();
  mask_sample_bwd_kernel<<<O, 256, 0, stream>>>(G, boxes, masks, lin_x, lin_y, M, H, W, dmasks, dboxes);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// crop_bbox(backend='jj') (bilinear.py:127-128, :134-189): same arguments as ag2v_crop_bbox_fwd / _bwd
extern "C" int ag2v_crop_bbox_jj_fwd(const float* feats, const int* frame, const float* boxes, const float* ws_x,
                                     const float* we_x, const float* ws_y, const float* we_y, int n_crops, int C,
                                     int H, int W, int HH, int WW, float* out, cudaStream_t stream) {
  AG2V_REQUIRE(n_crops >= 0 && C > 0 && H > 0 && W > 0 && HH > 0 && WW > 0, "crop_bbox_jj: bad sizes");
  if (n_crops == 0) return AG2V_OK;
  AG2V_REQUIRE(feats && frame && boxes && ws_x && we_x && ws_y && we_y && out, "crop_bbox_jj: null pointer");
  crop_jj_fwd_kernel<<<blocks_for((long long)n_crops * HH * WW), 256, 0, stream>>>(feats, frame, boxes, ws_x, we_x, ws_y,
                                                                                 we_y, n_crops, C, H, W, HH, WW, out);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

extern "C" int ag2v_crop_bbox_jj_bwd(const float* dout, const int* frame, const float* boxes, const float* ws_x,
                                     const float* we_x, const float* ws_y, const float* we_y, int n_crops, int C,
                                     int H, int W, int HH, int WW, float* dfeats, cudaStream_t stream) {
  AG2V_REQUIRE(n_crops >= 0 && C > 0 && H > 0 && W > 0 && HH > 0 && WW > 0, "crop_bbox_jj_bwd: bad sizes");
  if (n_crops == 0) return AG2V_OK;
  AG2V_REQUIRE(dout && frame && boxes && ws_x && we_x && ws_y && we_y && dfeats, "crop_bbox_jj_bwd: null pointer");
  crop_jj_bwd_kernel<<<blocks_for((long long)n_crops * HH * WW), 256, 0, stream>>>(dout, frame, boxes, ws_x, we_x, ws_y,
                                                                                 we_y, n_crops, C, H, W, HH, WW, dfeats);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}
