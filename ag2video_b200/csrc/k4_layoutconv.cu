// K4 — layout composition fused into its consumer convolution.
//
// The reference materialises the layout  seg[c, p] = sum_o v[o, c] * m_o(p)  (1.34 GB per
// batch at 256^2) and then runs a dense 3x3 convolution 1027 -> 512 over it
// (generator.py:29-33,82-83; also flows_generator.py:32, 1027 -> 32).  Because the layout is
// rank-1 per object, the convolution collapses:
//
//   conv(seg)[co, p] = sum_o sum_k U[o, k, co] * m_o(p + k),   U[o, k, :] = W[:, :, k] v[o, :]
//
// with m_o(y, x) = wy_o(y) * wx_o(x) the separable weights of K2 (zero outside the image:
// that is the convolution's zero padding).  U is a tiny GEMM (done by the host with torch);
// this file holds the two memory-bound kernels around it:
//
//   forward : out[n, p, :] += sum over the objects that touch p and the 9 taps  (NHWC, in place on
//             a base that already holds the convolution of the 3 image channels; pixels that no
//             object touches are not even read)
//   backward: dU[n, o, k, :] = sum_p dout[n, p, :] * m_o(p + k)  over the object's dilated window.
//
// ~500x fewer FLOPs than the dense convolution and the 1024-channel layout never exists.
#include "common.cuh"

namespace ag2v {

constexpr int LC_TX = 16, LC_TY = 8;          // pixel tile of one CTA (128 pixels)
constexpr int LC_THREADS = 256;
constexpr int LC_MAXACT = 6;                  // objects staged in shared memory at once

// tables as produced by layout_tables_kernel (k2_layout.cu): wx [N,S,W], wy [N,S,H], range [N,S]
struct LcTables { const float* wx; const float* wy; const int4* range; };

// CPL = output channels per lane (Co = 32 * CPL): 16 for the 512-channel conv, 1 for the 32-channel one
template <int CPL>
__global__ void __launch_bounds__(LC_THREADS)
layout_conv_fwd_kernel(const float* __restrict__ U, LcTables tb, int S, int H, int W, float* __restrict__ out) {
  constexpr int Co = 32 * CPL;
  extern __shared__ __align__(16) float lc_smem[];           // [LC_MAXACT][9][Co] U | [LC_MAXACT][3 + LC_TY-1+...] weights
  float* u_s = lc_smem;
  float* wy_s = u_s + LC_MAXACT * 9 * Co;                   // [LC_MAXACT][LC_TY + 2]
  float* wx_s = wy_s + LC_MAXACT * (LC_TY + 2);             // [LC_MAXACT][LC_TX + 2]
  __shared__ int act[64];
  __shared__ int n_act;
  const int n = blockIdx.z, x0 = blockIdx.x * LC_TX, y0 = blockIdx.y * LC_TY;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) n_act = 0;
  __syncthreads();
  // objects whose support, dilated by the 3x3 taps, meets this tile (ascending object order)
  if (warp == 0) {
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      bool hit = false;
      if (s < S) {
        const int4 r = tb.range[(size_t)n * S + s];
        hit = r.y > r.x && r.w > r.z && r.x - 1 < x0 + LC_TX && r.y + 1 > x0 && r.z - 1 < y0 + LC_TY && r.w + 1 > y0;
      }
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (hit) act[n_act + __popc(m & ((1u << lane) - 1u))] = s;
      __syncwarp();
      if (lane == 0) n_act += __popc(m);
      __syncwarp();
    }
  }
  __syncthreads();
  const int total = n_act;
  if (total == 0) return;                                    // untouched tile: the base stays as it is
  for (int g0 = 0; g0 < total; g0 += LC_MAXACT) {
    const int ng = min(LC_MAXACT, total - g0);
    __syncthreads();
    for (int i = tid; i < ng * 9 * Co; i += LC_THREADS) {
      const int a = i / (9 * Co);
      u_s[i] = U[((size_t)n * S + act[g0 + a]) * 9 * Co + (i - a * 9 * Co)];
    }
    for (int i = tid; i < ng * (LC_TY + 2); i += LC_THREADS) {
      const int a = i / (LC_TY + 2), yy = y0 - 1 + (i - a * (LC_TY + 2));
      wy_s[i] = (yy >= 0 && yy < H) ? tb.wy[((size_t)n * S + act[g0 + a]) * H + yy] : 0.f;
    }
    for (int i = tid; i < ng * (LC_TX + 2); i += LC_THREADS) {
      const int a = i / (LC_TX + 2), xx = x0 - 1 + (i - a * (LC_TX + 2));
      wx_s[i] = (xx >= 0 && xx < W) ? tb.wx[((size_t)n * S + act[g0 + a]) * W + xx] : 0.f;
    }
    __syncthreads();
    // one warp per pixel, lanes over output channels
    for (int px = warp; px < LC_TX * LC_TY; px += LC_THREADS / 32) {
      const int ty = px / LC_TX, tx = px - ty * LC_TX;
      const int y = y0 + ty, x = x0 + tx;
      if (y >= H || x >= W) continue;
      float acc[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
      bool any = false;
      for (int a = 0; a < ng; ++a) {
        const float* wya = wy_s + a * (LC_TY + 2) + ty;       // wy(y-1), wy(y), wy(y+1)
        const float* wxa = wx_s + a * (LC_TX + 2) + tx;
        const float wy3[3] = {wya[0], wya[1], wya[2]};
        const float wx3[3] = {wxa[0], wxa[1], wxa[2]};
        if ((wy3[0] == 0.f && wy3[1] == 0.f && wy3[2] == 0.f) || (wx3[0] == 0.f && wx3[1] == 0.f && wx3[2] == 0.f)) continue;
        any = true;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float c = wy3[k / 3] * wx3[k % 3];
          if (c != 0.f) {
            const float* u = u_s + (a * 9 + k) * Co;
            if (CPL == 16) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 uv = *reinterpret_cast<const float4*>(u + 4 * lane + 128 * j);
                acc[4 * j] = fmaf(c, uv.x, acc[4 * j]); acc[4 * j + 1] = fmaf(c, uv.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(c, uv.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(c, uv.w, acc[4 * j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < CPL; ++j) acc[j] = fmaf(c, u[lane + 32 * j], acc[j]);
            }
          }
        }
      }
      if (!any) continue;
      float* dst = out + (((size_t)n * H + y) * W + x) * Co;
      if (CPL == 16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4* q = reinterpret_cast<float4*>(dst + 4 * lane + 128 * j);
          float4 o = *q;
          o.x += acc[4 * j]; o.y += acc[4 * j + 1]; o.z += acc[4 * j + 2]; o.w += acc[4 * j + 3];
          *q = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < CPL; ++j) dst[lane + 32 * j] += acc[j];
      }
    }
  }
}

// dU partials: one CTA per (n, s, chunk of LC_BR rows of the object's dilated window).
// Each warp walks pixels of the chunk; a lane keeps 9 taps x CPL channels of accumulators.
constexpr int LC_BR = 8;

template <int CPL>
__global__ void __launch_bounds__(LC_THREADS)
layout_conv_bwd_kernel(const float* __restrict__ dout, LcTables tb, int S, int H, int W, int max_chunks,
                       float* __restrict__ part /*[N*S][max_chunks][9][Co]*/) {
  constexpr int Co = 32 * CPL;
  extern __shared__ __align__(16) float lc_smem[];           // [9][Co] cross-warp accumulation
  const int ns = blockIdx.x, chunk = blockIdx.y;
  const int n = ns / S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int4 r = tb.range[ns];
  float* dst = part + ((size_t)ns * max_chunks + chunk) * 9 * Co;
  const bool empty = !(r.y > r.x && r.w > r.z);
  const int ylo = max(r.z - 1, 0), yhi = min(r.w + 1, H), xlo = max(r.x - 1, 0), xhi = min(r.y + 1, W);
  const int yc0 = ylo + chunk * LC_BR, yc1 = min(yc0 + LC_BR, yhi);
  if (empty || yc0 >= yhi) {                                 // nothing to add: partial = 0
    for (int i = tid; i < 9 * Co; i += LC_THREADS) dst[i] = 0.f;
    return;
  }
  const float* wy = tb.wy + (size_t)ns * H;
  const float* wx = tb.wx + (size_t)ns * W;
  float acc[9][CPL];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[k][j] = 0.f;
  const int wpix = xhi - xlo, npix = (yc1 - yc0) * wpix;
  for (int i = warp; i < npix; i += LC_THREADS / 32) {
    const int y = yc0 + i / wpix, x = xlo + i % wpix;
    float wy3[3], wx3[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const int yy = y + d - 1, xx = x + d - 1;
      wy3[d] = (yy >= 0 && yy < H) ? wy[yy] : 0.f;
      wx3[d] = (xx >= 0 && xx < W) ? wx[xx] : 0.f;
    }
    const float* src = dout + (((size_t)n * H + y) * W + x) * Co;
    float g[CPL];
    if (CPL == 16) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(src + 4 * lane + 128 * j);
        g[4 * j] = v.x; g[4 * j + 1] = v.y; g[4 * j + 2] = v.z; g[4 * j + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CPL; ++j) g[j] = src[lane + 32 * j];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float c = wy3[k / 3] * wx3[k % 3];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[k][j] = fmaf(c, g[j], acc[k][j]);
    }
  }
  // cross-warp sum in warp order (deterministic)
  for (int w = 0; w < LC_THREADS / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        if (CPL == 16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float4* q = reinterpret_cast<float4*>(lc_smem + k * Co + 4 * lane + 128 * j);
            float4 o = (w == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : *q;
            o.x += acc[k][4 * j]; o.y += acc[k][4 * j + 1]; o.z += acc[k][4 * j + 2]; o.w += acc[k][4 * j + 3];
            *q = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            float* q = lc_smem + k * Co + lane + 32 * j;
            *q = (w == 0 ? 0.f : *q) + acc[k][j];
          }
        }
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < 9 * Co; i += LC_THREADS) dst[i] = lc_smem[i];
}

// dU[ns][k][co] = sum over chunks (in chunk order) of the partials
__global__ void layout_conv_reduce_kernel(const float* __restrict__ part, int max_chunks, int per, long long total,
                                          float* __restrict__ dU) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long ns = i / per; const int e = (int)(i - ns * per);
    float acc = 0.f;
    for (int c = 0; c < max_chunks; ++c) acc += part[((size_t)ns * max_chunks + c) * per + e];
    dU[i] = acc;
  }
}

static size_t lc_tables_bytes(int N, int S, int H, int W) {
  size_t b = (size_t)N * S * (W + H) * sizeof(float);
  b = (b + 15) & ~(size_t)15;
  return b + (size_t)N * S * sizeof(int4) + (size_t)N * sizeof(float) + 256;
}

}  // namespace ag2v

using namespace ag2v;

// the table workspace has the layout of ag2v_boxes_to_layout_workspace_bytes(N, S, H, W)
static LcTables lc_carve(const void* ws, int N, int S, int H, int W) {
  LcTables t;
  const char* p = (const char*)ws;
  t.wx = (const float*)p; p += (size_t)N * S * W * sizeof(float);
  t.wy = (const float*)p; p += (size_t)N * S * H * sizeof(float);
  p = (const char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  t.range = (const int4*)p;
  return t;
}

extern "C" int ag2v_boxes_to_layout_tables(const float* boxes, const uint8_t* valid, const float* lin_x,
                                           const float* lin_y, int N, int O, int H, int W, void* workspace,
                                           cudaStream_t stream);

static int lc_rows_chunks(int H) { return ceil_div(H + 2, LC_BR); }

// Floats of scratch for the backward partials.
extern "C" size_t ag2v_layout_conv_bwd_workspace_floats(int N, int S, int Co, int H) {
  return (size_t)N * S * lc_rows_chunks(H) * 9 * Co;
}

// out [N,H,W,Co] (NHWC) += sum_s sum_k U[n,s,k,:] * m_s(p + k).  `tables` = workspace filled by
// ag2v_boxes_to_layout_tables for boxes [N,S,4] (S = frame slots x objects).  Co in {32, 512}.
extern "C" int ag2v_layout_conv_fwd(const float* U, const void* tables, int N, int S, int Co, int H, int W,
                                    float* out, cudaStream_t stream) {
  AG2V_REQUIRE(U && tables && out && N > 0 && S > 0 && S <= 64 && H > 0 && W > 0, "layout_conv_fwd: bad arguments");
  AG2V_REQUIRE(Co == 512 || Co == 32, "layout_conv_fwd: Co must be 512 or 32 (got %d)", Co);
  LcTables tb = lc_carve(tables, N, S, H, W);
  dim3 grid(ceil_div(W, LC_TX), ceil_div(H, LC_TY), N);
  const size_t smem = ((size_t)LC_MAXACT * 9 * Co + LC_MAXACT * (LC_TY + 2 + LC_TX + 2)) * sizeof(float);
  if (Co == 512) {
    AG2V_CUDA(cudaFuncSetAttribute(layout_conv_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layout_conv_fwd_kernel<16><<<grid, LC_THREADS, smem, stream>>>(U, tb, S, H, W, out);
  } else {
    layout_conv_fwd_kernel<1><<<grid, LC_THREADS, smem, stream>>>(U, tb, S, H, W, out);
  }
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// dU [N,S,9,Co] = sum_p dout[n,p,:] * m_s(p + k); `part` = ag2v_layout_conv_bwd_workspace_floats scratch
extern "C" int ag2v_layout_conv_bwd(const float* dout, const void* tables, int N, int S, int Co, int H, int W,
                                    float* part, float* dU, cudaStream_t stream) {
  AG2V_REQUIRE(dout && tables && part && dU && N > 0 && S > 0 && H > 0 && W > 0, "layout_conv_bwd: bad arguments");
  AG2V_REQUIRE(Co == 512 || Co == 32, "layout_conv_bwd: Co must be 512 or 32 (got %d)", Co);
  LcTables tb = lc_carve(tables, N, S, H, W);
  const int chunks = lc_rows_chunks(H);
  dim3 grid(N * S, chunks);
  const size_t smem = (size_t)9 * Co * sizeof(float);
  if (Co == 512) layout_conv_bwd_kernel<16><<<grid, LC_THREADS, smem, stream>>>(dout, tb, S, H, W, chunks, part);
  else layout_conv_bwd_kernel<1><<<grid, LC_THREADS, smem, stream>>>(dout, tb, S, H, W, chunks, part);
  AG2V_LAUNCH_CHECK();
  const long long total = (long long)N * S * 9 * Co;
  layout_conv_reduce_kernel<<<(unsigned)(ceil_div_ll(total, 256) > 4096 ? 4096 : ceil_div_ll(total, 256)), 256, 0, stream>>>(
      part, chunks, 9 * Co, total, dU);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}
