// K2 — layout composition (reference: models/layout.py:28-63 boxes_to_layout,
// :98-130 _boxes_to_grid, :205-237 _pool_samples).
//
// boxes_to_layout bilinearly samples a CONSTANT 8x8 image per object, so per
// object the sampled plane is separable: m_o(y,x) = wy_o(y) * wx_o(x), where
// w(t) is the in-bounds bilinear weight mass at source coordinate t (zeros
// padding, align_corners=True).  out[n,d,y,x] = sum_o v[n,o,d] * m_o(y,x), objects
// in index order.  The fp32 coordinate arithmetic replays the reference op by op
// (sub, IEEE div, *2, -1, +1, /2, *7, floor) so the pixel support is bit-exact;
// the grid [O,H,W,2] and the sampled [O,D,H,W] tensors are never materialised.
//
// HBM traffic: forward writes 4*N*D*H*W bytes once (write-bound); backward reads
// only each object's support window of dout.
#include "common.cuh"

namespace ag2v {

__device__ __forceinline__ float axis_weight(float lin, float p0, float extent) {
  // _boxes_to_grid (layout.py:119-128) then grid_sample's unnormalise for an
  // 8-wide source with align_corners=True.  No FMA contraction: every op rounds.
  float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(lin, p0), extent), 2.f), 1.f);
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), 7.f);
  float fl = floorf(ix);
  float w_hi = __fsub_rn(ix, fl);                      // weight of tap fl+1
  float w_lo = __fsub_rn(__fadd_rn(fl, 1.f), ix);      // weight of tap fl
  float w = 0.f;
  if (fl >= 0.f && fl <= 7.f) w = w_lo;
  if (fl >= -1.f && fl <= 6.f) w = __fadd_rn(w, w_hi);
  return w;                                            // NaN / inf coordinates contribute 0
}

// One block per (n, o): separable weight tables + support window.
// range[n*O+o] = (xlo, xhi, ylo, yhi), empty window for illegal objects.
__global__ void layout_tables_kernel(const float* __restrict__ boxes, const uint8_t* __restrict__ valid,
                                     const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                                     int O, int H, int W, int avg, float* __restrict__ wx,
                                     float* __restrict__ wy, int4* __restrict__ range,
                                     float* __restrict__ scale) {
  const int no = blockIdx.x;
  const int n = no / O, o = no - n * O;
  __shared__ int s_lo[2], s_hi[2];
  const float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)no * 4);
  // layout.py:40-42 drops all-zero boxes; callers' object masks arrive in `valid`.
  bool legal = (bx.x != 0.f) || (bx.y != 0.f) || (bx.z != 0.f) || (bx.w != 0.f);
  if (valid != nullptr) legal = legal && (valid[no] != 0);
  if (threadIdx.x < 2) { s_lo[threadIdx.x] = 1 << 30; s_hi[threadIdx.x] = -1; }
  __syncthreads();
  int lo = 1 << 30, hi = -1;
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    float w = legal ? axis_weight(lin_x[j], bx.x, bx.z) : 0.f;
    wx[(size_t)no * W + j] = w;
    if (w != 0.f) { lo = min(lo, j); hi = max(hi, j); }
  }
  atomicMin(&s_lo[0], lo); atomicMax(&s_hi[0], hi);
  lo = 1 << 30; hi = -1;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    float w = legal ? axis_weight(lin_y[i], bx.y, bx.w) : 0.f;
    wy[(size_t)no * H + i] = w;
    if (w != 0.f) { lo = min(lo, i); hi = max(hi, i); }
  }
  atomicMin(&s_lo[1], lo); atomicMax(&s_hi[1], hi);
  __syncthreads();
  if (threadIdx.x == 0) {
    int4 r;
    r.x = s_hi[0] < 0 ? 0 : s_lo[0]; r.y = s_hi[0] + 1;
    r.z = s_hi[1] < 0 ? 0 : s_lo[1]; r.w = s_hi[1] + 1;
    if (s_hi[0] < 0 || s_hi[1] < 0) { r.x = r.y = r.z = r.w = 0; }
    range[no] = r;
    if (o == 0) {
      // _pool_samples 'avg' (layout.py:225-233): divide by the legal-object count (>= 1)
      float s = 1.f;
      if (avg) {
        int cnt = 0;
        for (int k = 0; k < O; ++k) {
          const float* b = boxes + (size_t)(n * O + k) * 4;
          bool lg = (b[0] != 0.f) || (b[1] != 0.f) || (b[2] != 0.f) || (b[3] != 0.f);
          if (valid != nullptr) lg = lg && (valid[n * O + k] != 0);
          cnt += lg ? 1 : 0;
        }
        s = 1.f / (float)max(cnt, 1);
      }
      scale[n] = s;
    }
  }
}

constexpr int kDC = 16;   // channels per CTA
constexpr int kRB = 32;   // rows per CTA

// grid (ceil(H/kRB), ceil(D/kDC), N), 256 threads.  Each warp owns rows; for a row
// the active objects (wy != 0) are found once, then kDC channel rows of W floats are
// written with 128-bit stores.  smem: wx[O][W] | wy[O][kRB] | v[O][kDC].
template <bool VEC4>
__global__ void __launch_bounds__(256)
layout_fwd_kernel(const float* __restrict__ vecs, const float* __restrict__ wx_g,
                  const float* __restrict__ wy_g, const float* __restrict__ scale, int O, int D, int H,
                  int W, size_t out_nstride, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* wx_s = smem;
  float* wy_s = wx_s + (size_t)O * W;
  float* v_s = wy_s + O * kRB;
  const int n = blockIdx.z, d0 = blockIdx.y * kDC, y0 = blockIdx.x * kRB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float sc = scale[n];
  for (int i = tid; i < O * W; i += blockDim.x) wx_s[i] = wx_g[(size_t)n * O * W + i];
  for (int i = tid; i < O * kRB; i += blockDim.x) {
    int o = i / kRB, r = i - o * kRB;
    wy_s[i] = (y0 + r < H) ? wy_g[((size_t)n * O + o) * H + y0 + r] : 0.f;
  }
  for (int i = tid; i < O * kDC; i += blockDim.x) {
    int o = i / kDC, c = i - o * kDC;
    v_s[i] = (d0 + c < D) ? vecs[((size_t)n * O + o) * D + d0 + c] * sc : 0.f;
  }
  __syncthreads();
  const int rows = min(kRB, H - y0), chans = min(kDC, D - d0);
  for (int r = warp; r < rows; r += 8) {
    // active objects of this row, up to 64 (two ballots), ascending object order
    unsigned m0 = __ballot_sync(0xffffffffu, lane < O && wy_s[lane * kRB + r] != 0.f);
    unsigned m1 = (O > 32) ? __ballot_sync(0xffffffffu, lane + 32 < O && wy_s[(lane + 32) * kRB + r] != 0.f) : 0u;
    // Fast path (the common case: at most 4 objects touch a row, W <= 256): the row's x-weights
    // and y-weights stay in registers across the channel loop, so each output float4 costs one
    // broadcast LDS per active object instead of three loads.
    const int n_act = __popc(m0) + __popc(m1);
    if (VEC4 && n_act <= 4 && W <= 256) {
      float4 wxa[4][2];
      float wya[4];
      int oa[4];
      {
        unsigned mm0 = m0, mm1 = m1;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          int o = -1;
          if (mm0) { o = __ffs(mm0) - 1; mm0 &= mm0 - 1; }
          else if (mm1) { o = __ffs(mm1) - 1 + 32; mm1 &= mm1 - 1; }
          oa[a] = o < 0 ? 0 : o;
          wya[a] = o < 0 ? 0.f : wy_s[o * kRB + r];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int x4 = lane + 32 * k;
            wxa[a][k] = (o >= 0 && x4 < (W >> 2)) ? reinterpret_cast<const float4*>(wx_s + (size_t)o * W)[x4]
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      for (int c = 0; c < chans; ++c) {
        float* dst = out + (size_t)n * out_nstride + (((size_t)d0 + c) * H + y0 + r) * W;
        float coef[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) coef[a] = v_s[oa[a] * kDC + c] * wya[a];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int x4 = lane + 32 * k;
          if (x4 < (W >> 2)) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int a = 0; a < 4; ++a) {          // object order; inactive slots have coef == 0 and w == 0
              acc.x = fmaf(coef[a], wxa[a][k].x, acc.x); acc.y = fmaf(coef[a], wxa[a][k].y, acc.y);
              acc.z = fmaf(coef[a], wxa[a][k].z, acc.z); acc.w = fmaf(coef[a], wxa[a][k].w, acc.w);
            }
            reinterpret_cast<float4*>(dst)[x4] = acc;
          }
        }
      }
      continue;
    }
    for (int c = 0; c < chans; ++c) {
      float* dst = out + (size_t)n * out_nstride + (((size_t)d0 + c) * H + y0 + r) * W;
      if (VEC4) {
        for (int x4 = lane; x4 < (W >> 2); x4 += 32) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int half = 0; half < 2; ++half) {
            unsigned m = half ? m1 : m0;
            while (m) {
              int o = __ffs(m) - 1 + half * 32; m &= m - 1;
              float coef = v_s[o * kDC + c] * wy_s[o * kRB + r];
              float4 w = reinterpret_cast<const float4*>(wx_s + (size_t)o * W)[x4];
              acc.x = fmaf(coef, w.x, acc.x); acc.y = fmaf(coef, w.y, acc.y);
              acc.z = fmaf(coef, w.z, acc.z); acc.w = fmaf(coef, w.w, acc.w);
            }
          }
          reinterpret_cast<float4*>(dst)[x4] = acc;
        }
      } else {
        for (int x = lane; x < W; x += 32) {
          float acc = 0.f;
          for (int half = 0; half < 2; ++half) {
            unsigned m = half ? m1 : m0;
            while (m) {
              int o = __ffs(m) - 1 + half * 32; m &= m - 1;
              acc = fmaf(v_s[o * kDC + c] * wy_s[o * kRB + r], wx_s[(size_t)o * W + x], acc);
            }
          }
          dst[x] = acc;
        }
      }
    }
  }
}

// dvecs[n,o,d] = scale[n] * sum_{y,x in window(o)} dout[n,d,y,x] * wy_o(y) * wx_o(x).
// One warp per (n, o, d); fixed summation order => deterministic.
__global__ void __launch_bounds__(256)
layout_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ wx_g,
                  const float* __restrict__ wy_g, const int4* __restrict__ range,
                  const float* __restrict__ scale, int N, int O, int D, int H, int W,
                  size_t dout_nstride, float* __restrict__ dvecs) {
  const int lane = threadIdx.x & 31;
  const long long task = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= (long long)N * O * D) return;
  const int d = (int)(task % D);
  const int no = (int)(task / D);
  const int n = no / O;
  const int4 rg = range[no];
  float acc = 0.f;
  if (rg.y > rg.x && rg.w > rg.z) {
    const float* wx = wx_g + (size_t)no * W;
    const float* wy = wy_g + (size_t)no * H;
    const float* src = dout + (size_t)n * dout_nstride + (size_t)d * H * W;
    const int xa = rg.x & ~31;
    for (int x = xa + lane; x < rg.y; x += 32) {
      const float wxx = (x >= rg.x) ? wx[x] : 0.f;
      float col = 0.f;
      int y = rg.z;
      for (; y + 4 <= rg.w; y += 4) {
        float a0 = src[(size_t)y * W + x], a1 = src[(size_t)(y + 1) * W + x];
        float a2 = src[(size_t)(y + 2) * W + x], a3 = src[(size_t)(y + 3) * W + x];
        col = fmaf(a0, wy[y], col); col = fmaf(a1, wy[y + 1], col);
        col = fmaf(a2, wy[y + 2], col); col = fmaf(a3, wy[y + 3], col);
      }
      for (; y < rg.w; ++y) col = fmaf(src[(size_t)y * W + x], wy[y], col);
      acc = fmaf(col, wxx, acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) dvecs[task] = acc * scale[n];
}

// ---- gradient with respect to the boxes (autograd of layout.py:98-130 through grid_sample's grid gradient) ----
// w(ix) is piecewise linear in the source coordinate ix = 7 * (lin - p0) / extent: slope +1 while only the upper tap
// is inside the 8-wide source (floor(ix) == -1), -1 while only the lower tap is (floor(ix) == 7), 0 elsewhere; torch
// computes the same from the tap values of the constant image.  d ix / d p0 = -7 / extent, d ix / d extent =
// -7 (lin - p0) / extent^2 (the chain of sub, div, *2 - 1 and the (g + 1) / 2 * 7 unnormalisation).
__device__ __forceinline__ float axis_slope(float lin, float p0, float extent) {
  float g = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(lin, p0), extent), 2.f), 1.f);
  float ix = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.f), 2.f), 7.f);
  float fl = floorf(ix);
  float s = 0.f;
  if (fl >= -1.f && fl <= 6.f) s += 1.f;
  if (fl >= 0.f && fl <= 7.f) s -= 1.f;
  return s;
}

// One block per (n, o).  Only the border pixels of the object's window (slope != 0 along x or y) contribute; for each
// of them G = scale * sum_d dout[n,d,y,x] * v[n,o,d].  Thread-strided pixels and a fixed reduction tree: deterministic.
__global__ void __launch_bounds__(256)
layout_dboxes_kernel(const float* __restrict__ dout, size_t dout_nstride, const float* __restrict__ vecs,
                     const float* __restrict__ boxes, const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                     const float* __restrict__ wx_g, const float* __restrict__ wy_g, const int4* __restrict__ range,
                     const float* __restrict__ scale, int O, int D, int H, int W, float* __restrict__ dboxes) {
  extern __shared__ float v_s[];
  __shared__ float red[8][4];
  const int no = blockIdx.x, n = no / O;
  const int4 rg = range[no];
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (rg.y > rg.x && rg.w > rg.z) {                              // illegal objects have an empty window: zero gradient
    for (int d = threadIdx.x; d < D; d += blockDim.x) v_s[d] = vecs[(size_t)no * D + d];
    __syncthreads();
    const float4 bx = *reinterpret_cast<const float4*>(boxes + (size_t)no * 4);
    const float sc = scale[n];
    const float* wx = wx_g + (size_t)no * W;
    const float* wy = wy_g + (size_t)no * H;
    // one pixel further than the support on each side: w can be exactly 0 where the slope is not
    const int xa = max(rg.x - 1, 0), xb = min(rg.y + 1, W), ya = max(rg.z - 1, 0), yb = min(rg.w + 1, H);
    const int cols = xb - xa, total = cols * (yb - ya);
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int y = ya + i / cols, x = xa + i % cols;
      const float sx = axis_slope(lin_x[x], bx.x, bx.z), sy = axis_slope(lin_y[y], bx.y, bx.w);
      if (sx == 0.f && sy == 0.f) continue;
      const float* src = dout + (size_t)n * dout_nstride + (size_t)y * W + x;
      float g = 0.f;
      for (int d = 0; d < D; ++d) g = fmaf(src[(size_t)d * H * W], v_s[d], g);
      g *= sc;
      const float gx = 7.f * g * wy[y] * sx, gy = 7.f * g * wx[x] * sy;      // dL / d(normalised X), dL / d(normalised Y)
      acc[0] -= gx / bx.z;
      acc[2] -= gx * (lin_x[x] - bx.x) / (bx.z * bx.z);
      acc[1] -= gy / bx.w;
      acc[3] -= gy * (lin_y[y] - bx.y) / (bx.w * bx.w);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = warp_sum(acc[k]);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 4; ++k) red[threadIdx.x >> 5][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
    dboxes[(size_t)no * 4 + threadIdx.x] = s;
  }
}

struct LayoutWs {
  float* wx; float* wy; int4* range; float* scale;
};

static size_t layout_ws_bytes(int N, int O, int H, int W) {
  size_t b = (size_t)N * O * (W + H) * sizeof(float);
  b = (b + 15) & ~(size_t)15;
  b += (size_t)N * O * sizeof(int4);
  b += (size_t)N * sizeof(float);
  return (b + 255) & ~(size_t)255;
}

static LayoutWs layout_ws_carve(void* ws, int N, int O, int H, int W) {
  LayoutWs r;
  char* p = (char*)ws;
  r.wx = (float*)p; p += (size_t)N * O * W * sizeof(float);
  r.wy = (float*)p; p += (size_t)N * O * H * sizeof(float);
  p = (char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  r.range = (int4*)p; p += (size_t)N * O * sizeof(int4);
  r.scale = (float*)p;
  return r;
}

}  // namespace ag2v

using namespace ag2v;

extern "C" size_t ag2v_boxes_to_layout_workspace_bytes(int N, int O, int H, int W) {
  return layout_ws_bytes(N, O, H, W);
}

static int layout_check(int N, int O, int D, int H, int W) {
  AG2V_REQUIRE(N >= 0 && O >= 0 && D >= 0 && H > 0 && W > 0, "boxes_to_layout: bad sizes N=%d O=%d D=%d H=%d W=%d", N, O, D, H, W);
  AG2V_REQUIRE(O <= 64, "boxes_to_layout: at most 64 objects per frame are supported (got %d)", O);
  AG2V_REQUIRE((size_t)O * W * 4 + (size_t)O * (kRB + kDC) * 4 <= 200 * 1024, "boxes_to_layout: O*W too large for shared memory");
  return AG2V_OK;
}

// Only the per-object separable weight tables + support windows (used by the fused
// layout->conv kernels of k4_layoutconv.cu); workspace as for ag2v_boxes_to_layout_fwd.
extern "C" int ag2v_boxes_to_layout_tables(const float* boxes, const uint8_t* valid, const float* lin_x,
                                           const float* lin_y, int N, int O, int H, int W, void* workspace,
                                           cudaStream_t stream) {
  int rc = layout_check(N, O, 1, H, W);
  if (rc) return rc;
  if (N == 0 || O == 0) return AG2V_OK;
  AG2V_REQUIRE(boxes && lin_x && lin_y && workspace, "boxes_to_layout_tables: null pointer");
  LayoutWs ws = layout_ws_carve(workspace, N, O, H, W);
  layout_tables_kernel<<<N * O, 256, 0, stream>>>(boxes, valid, lin_x, lin_y, O, H, W, 0, ws.wx, ws.wy, ws.range, ws.scale);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// out[n] starts at out + n * out_nstride floats and holds [D,H,W]: out_nstride = D*H*W is the plain
// [N,D,H,W] tensor, a larger stride writes the layout straight into a channel slice of a wider
// NCHW buffer (the discriminator's cat[img, seg], discriminator.py:338-342).
static int layout_fwd_impl(const float* vecs, const float* boxes, const uint8_t* valid,
                           const float* lin_x, const float* lin_y, int N, int O, int D,
                           int H, int W, int avg, void* workspace, float* out, size_t out_nstride,
                           cudaStream_t stream) {
  int rc = layout_check(N, O, D, H, W);
  if (rc) return rc;
  if (N == 0 || D == 0) return AG2V_OK;
  AG2V_REQUIRE(out && workspace && lin_x && lin_y, "boxes_to_layout_fwd: null pointer");
  AG2V_REQUIRE(out_nstride >= (size_t)D * H * W, "boxes_to_layout_fwd: batch stride %zu smaller than D*H*W", out_nstride);
  if (O == 0) {
    AG2V_CUDA(cudaMemset2DAsync(out, out_nstride * sizeof(float), 0, (size_t)D * H * W * sizeof(float), N, stream));
    return AG2V_OK;
  }
  AG2V_REQUIRE(vecs && boxes, "boxes_to_layout_fwd: null pointer");
  LayoutWs ws = layout_ws_carve(workspace, N, O, H, W);
  layout_tables_kernel<<<N * O, 256, 0, stream>>>(boxes, valid, lin_x, lin_y, O, H, W, avg, ws.wx, ws.wy,
                                                  ws.range, ws.scale);
  AG2V_LAUNCH_CHECK();
  size_t smem = ((size_t)O * W + (size_t)O * (kRB + kDC)) * sizeof(float);
  dim3 grid(ceil_div(H, kRB), ceil_div(D, kDC), N);
  bool vec4 = (W % 4 == 0) && (((uintptr_t)out & 15) == 0) && (out_nstride % 4 == 0);
  if (vec4) {
    AG2V_CUDA(cudaFuncSetAttribute(layout_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_fwd_kernel<true><<<grid, 256, smem, stream>>>(vecs, ws.wx, ws.wy, ws.scale, O, D, H, W, out_nstride, out);
  } else {
    AG2V_CUDA(cudaFuncSetAttribute(layout_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_fwd_kernel<false><<<grid, 256, smem, stream>>>(vecs, ws.wx, ws.wy, ws.scale, O, D, H, W, out_nstride, out);
  }
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

extern "C" int ag2v_boxes_to_layout_fwd(const float* vecs, const float* boxes, const uint8_t* valid,
                                        const float* lin_x, const float* lin_y, int N, int O, int D,
                                        int H, int W, int avg, void* workspace, float* out,
                                        cudaStream_t stream) {
  return layout_fwd_impl(vecs, boxes, valid, lin_x, lin_y, N, O, D, H, W, avg, workspace, out, (size_t)D * H * W, stream);
}

extern "C" int ag2v_boxes_to_layout_fwd_strided(const float* vecs, const float* boxes, const uint8_t* valid,
                                                const float* lin_x, const float* lin_y, int N, int O, int D,
                                                int H, int W, int avg, void* workspace, float* out,
                                                long long out_batch_stride, cudaStream_t stream) {
  AG2V_REQUIRE(out_batch_stride >= 0, "boxes_to_layout_fwd_strided: negative batch stride");
  return layout_fwd_impl(vecs, boxes, valid, lin_x, lin_y, N, O, D, H, W, avg, workspace, out, (size_t)out_batch_stride, stream);
}

// `workspace` must be the buffer a forward call with the same boxes filled
// (the tables are reused); pass recompute=1 to rebuild them from boxes.
static int layout_bwd_impl(const float* dout, size_t dout_nstride, const float* boxes, const uint8_t* valid,
                           const float* lin_x, const float* lin_y, int N, int O, int D,
                           int H, int W, int avg, int recompute, void* workspace,
                           float* dvecs, cudaStream_t stream) {
  int rc = layout_check(N, O, D, H, W);
  if (rc) return rc;
  if (N == 0 || D == 0 || O == 0) return AG2V_OK;
  AG2V_REQUIRE(dout && workspace && dvecs, "boxes_to_layout_bwd: null pointer");
  LayoutWs ws = layout_ws_carve(workspace, N, O, H, W);
  if (recompute) {
    AG2V_REQUIRE(boxes && lin_x && lin_y, "boxes_to_layout_bwd: null pointer");
    layout_tables_kernel<<<N * O, 256, 0, stream>>>(boxes, valid, lin_x, lin_y, O, H, W, avg, ws.wx,
                                                    ws.wy, ws.range, ws.scale);
    AG2V_LAUNCH_CHECK();
  }
  long long tasks = (long long)N * O * D;
  layout_bwd_kernel<<<(unsigned)ceil_div_ll(tasks, 8), 256, 0, stream>>>(dout, ws.wx, ws.wy, ws.range,
                                                                         ws.scale, N, O, D, H, W, dout_nstride, dvecs);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

extern "C" int ag2v_boxes_to_layout_bwd(const float* dout, const float* boxes, const uint8_t* valid,
                                        const float* lin_x, const float* lin_y, int N, int O, int D,
                                        int H, int W, int avg, int recompute, void* workspace,
                                        float* dvecs, cudaStream_t stream) {
  return layout_bwd_impl(dout, (size_t)D * H * W, boxes, valid, lin_x, lin_y, N, O, D, H, W, avg, recompute, workspace, dvecs, stream);
}

extern "C" int ag2v_boxes_to_layout_bwd_strided(const float* dout, long long dout_batch_stride, const float* boxes,
                                                const uint8_t* valid, const float* lin_x, const float* lin_y, int N,
                                                int O, int D, int H, int W, int avg, int recompute, void* workspace,
                                                float* dvecs, cudaStream_t stream) {
  AG2V_REQUIRE(dout_batch_stride >= (long long)D * H * W, "boxes_to_layout_bwd_strided: batch stride smaller than D*H*W");
  return layout_bwd_impl(dout, (size_t)dout_batch_stride, boxes, valid, lin_x, lin_y, N, O, D, H, W, avg, recompute, workspace, dvecs, stream);
}

// dboxes [N,O,4] = gradient of sum(dout * boxes_to_layout(vecs, boxes)) with respect to the xywh boxes, what autograd
// gives the reference through grid_sample (layout.py:55-57).  `workspace` is the buffer the forward call filled.
// dout[n] starts at dout + n * dout_batch_stride floats.  Objects the forward dropped get zero.
extern "C" int ag2v_boxes_to_layout_dboxes(const float* dout, long long dout_batch_stride, const float* vecs,
                                           const float* boxes, const float* lin_x, const float* lin_y, int N, int O,
                                           int D, int H, int W, void* workspace, float* dboxes, cudaStream_t stream) {
  int rc = layout_check(N, O, D, H, W);
  if (rc) return rc;
  if (N == 0 || O == 0) return AG2V_OK;
  AG2V_REQUIRE(dout && vecs && boxes && lin_x && lin_y && workspace && dboxes, "boxes_to_layout_dboxes: null pointer");
  AG2V_REQUIRE(dout_batch_stride >= (long long)D * H * W, "boxes_to_layout_dboxes: batch stride smaller than D*H*W");
  AG2V_REQUIRE(D * sizeof(float) <= 48 * 1024, "boxes_to_layout_dboxes: D=%d too large", D);
  LayoutWs ws = layout_ws_carve(workspace, N, O, H, W);
  layout_dboxes_kernel<<<N * O, 256, D * sizeof(float), stream>>>(dout, (size_t)dout_batch_stride, vecs, boxes, lin_x, lin_y,
                                                                   ws.wx, ws.wy, ws.range, ws.scale, O, D, H, W, dboxes);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}
