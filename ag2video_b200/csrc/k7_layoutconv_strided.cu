// K7 — the discriminator's first PatchGAN convolution evaluated through the rank-1 structure of the
// layout (same idea as K4, csrc/k4_layoutconv.cu).
//
// The reference builds seg[c,p] = sum_o v[o,c] m_o(p) with 256 channels per (clip, frame), concatenates
// it with the image (259 channels, discriminator.py:317-342) and runs a dense 4x4 stride-2 convolution
// 259 -> 64 over it at every scale (:326-372; the second scale sees the 3x3/stride-2 average pool of
// the same tensor, :271, :350).  With m_o(y,x) = wy_o(y) wx_o(x):
//
//   conv(seg)[co, py, px] = sum_o sum_{ky,kx} U[o, ky*K+kx, co] * wy_o(S*py + ky - P) * wx_o(S*px + kx - P)
//   U[o, k, :] = W[:, 3:, k] v[o, :]                       (a tiny GEMM, done by the host)
//
// and the average pool commutes with the sum over objects and keeps the mask separable
// (count_include_pad=False divides by cnt_y(py)*cnt_x(px)):  avgpool(m_o) = wy'_o (x) wx'_o with
// w'(p) = sum_{k<3, 0<=2p+k-1<n} w(2p+k-1) / cnt(p).  So neither the 256-channel layout, nor the
// 259-channel concatenation, nor its pooled copy, nor the dense convolution over them exist:
//
//   layout_tables_pool_kernel : tables (H,W) -> tables (H2,W2) of the pooled masks
//   layout_sconv_fwd_kernel   : out[n,py,px,:] += sum over the objects that touch the pixel  (NHWC, in
//                               place on a base holding the convolution of the 3 image channels + bias)
//   layout_sconv_bwd_kernel   : dU[n,o,k,:] = sum_p dout[n,p,:] * m_o(S*p + k - P)  over the object's window
#include "common.cuh"

namespace ag2v {

constexpr int SC_TX = 16, SC_TY = 8;          // output-pixel tile of one CTA
constexpr int SC_THREADS = 256;
constexpr int SC_MAXACT = 6;                  // objects staged in shared memory at once
constexpr int SC_BR = 8;                      // output rows per backward chunk

struct ScTables { const float* wx; const float* wy; const int4* range; };

// One block per (n, o): 3x3 / stride 2 / pad 1 average pool (count_include_pad = False) of the separable
// weights, and the support window of the result.
__global__ void layout_tables_pool_kernel(const float* __restrict__ wx, const float* __restrict__ wy, int H, int W,
                                          int H2, int W2, float* __restrict__ wx2, float* __restrict__ wy2,
                                          int4* __restrict__ range2) {
  const int no = blockIdx.x;
  __shared__ int s_lo[2], s_hi[2];
  if (threadIdx.x < 2) { s_lo[threadIdx.x] = 1 << 30; s_hi[threadIdx.x] = -1; }
  __syncthreads();
  for (int axis = 0; axis < 2; ++axis) {
    const float* src = axis == 0 ? wx + (size_t)no * W : wy + (size_t)no * H;
    float* dst = axis == 0 ? wx2 + (size_t)no * W2 : wy2 + (size_t)no * H2;
    const int n_in = axis == 0 ? W : H, n_out = axis == 0 ? W2 : H2;
    int lo = 1 << 30, hi = -1;
    for (int p = threadIdx.x; p < n_out; p += blockDim.x) {
      float s = 0.f;
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int q = 2 * p + k - 1;
        if (q >= 0 && q < n_in) { s += src[q]; ++cnt; }
      }
      const float w = cnt > 0 ? s / (float)cnt : 0.f;
      dst[p] = w;
      if (w != 0.f) { lo = min(lo, p); hi = max(hi, p); }
    }
    atomicMin(&s_lo[axis], lo); atomicMax(&s_hi[axis], hi);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int4 r;
    r.x = s_lo[0]; r.y = s_hi[0] + 1; r.z = s_lo[1]; r.w = s_hi[1] + 1;
    if (s_hi[0] < 0 || s_hi[1] < 0) { r.x = r.y = r.z = r.w = 0; }
    range2[no] = r;
  }
}

// K x K taps, stride S, padding P; CPL = output channels per lane (Co = 32 * CPL).
template <int K, int S, int P, int CPL>
__global__ void __launch_bounds__(SC_THREADS)
layout_sconv_fwd_kernel(const float* __restrict__ U, ScTables tb, int NS, int H, int W, int Ho, int Wo,
                        float* __restrict__ out) {
  constexpr int Co = 32 * CPL;
  constexpr int KK = K * K;
  constexpr int NY = S * (SC_TY - 1) + K, NX = S * (SC_TX - 1) + K;     // input rows / columns a tile reads
  extern __shared__ __align__(16) float sc_smem[];
  float* u_s = sc_smem;                              // [SC_MAXACT][KK][Co]
  float* wy_s = u_s + SC_MAXACT * KK * Co;           // [SC_MAXACT][NY]
  float* wx_s = wy_s + SC_MAXACT * NY;               // [SC_MAXACT][NX]
  __shared__ int act[64];
  __shared__ int n_act;
  const int n = blockIdx.z, x0 = blockIdx.x * SC_TX, y0 = blockIdx.y * SC_TY;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ix0 = S * x0 - P, iy0 = S * y0 - P;      // first input column / row of the tile
  if (tid == 0) n_act = 0;
  __syncthreads();
  if (warp == 0) {                                   // objects whose support meets the tile's input window, ascending
    for (int s0 = 0; s0 < NS; s0 += 32) {
      const int s = s0 + lane;
      bool hit = false;
      if (s < NS) {
        const int4 r = tb.range[(size_t)n * NS + s];
        hit = r.y > r.x && r.w > r.z && r.x < ix0 + NX && r.y > ix0 && r.z < iy0 + NY && r.w > iy0;
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
  if (total == 0) return;                            // untouched tile: the base stays as it is
  for (int g0 = 0; g0 < total; g0 += SC_MAXACT) {
    const int ng = min(SC_MAXACT, total - g0);
    __syncthreads();
    for (int i = tid; i < ng * KK * Co; i += SC_THREADS) {
      const int a = i / (KK * Co);
      u_s[i] = U[((size_t)n * NS + act[g0 + a]) * KK * Co + (i - a * KK * Co)];
    }
    for (int i = tid; i < ng * NY; i += SC_THREADS) {
      const int a = i / NY, yy = iy0 + (i - a * NY);
      wy_s[i] = (yy >= 0 && yy < H) ? tb.wy[((size_t)n * NS + act[g0 + a]) * H + yy] : 0.f;
    }
    for (int i = tid; i < ng * NX; i += SC_THREADS) {
      const int a = i / NX, xx = ix0 + (i - a * NX);
      wx_s[i] = (xx >= 0 && xx < W) ? tb.wx[((size_t)n * NS + act[g0 + a]) * W + xx] : 0.f;
    }
    __syncthreads();
    for (int px = warp; px < SC_TX * SC_TY; px += SC_THREADS / 32) {       // one warp per output pixel
      const int ty = px / SC_TX, tx = px - ty * SC_TX;
      const int y = y0 + ty, x = x0 + tx;
      if (y >= Ho || x >= Wo) continue;
      float acc[CPL];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[j] = 0.f;
      bool any = false;
      for (int a = 0; a < ng; ++a) {
        const float* wya = wy_s + a * NY + S * ty;
        const float* wxa = wx_s + a * NX + S * tx;
        float wyk[K], wxk[K];
        bool zy = true, zx = true;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          wyk[k] = wya[k]; wxk[k] = wxa[k];
          zy = zy && (wyk[k] == 0.f); zx = zx && (wxk[k] == 0.f);
        }
        if (zy || zx) continue;
        any = true;
#pragma unroll
        for (int k = 0; k < KK; ++k) {
          const float c = wyk[k / K] * wxk[k % K];
          if (c != 0.f) {
            const float* u = u_s + (a * KK + k) * Co;
#pragma unroll
            for (int j = 0; j < CPL; ++j) acc[j] = fmaf(c, u[lane + 32 * j], acc[j]);
          }
        }
      }
      if (!any) continue;
      float* dst = out + (((size_t)n * Ho + y) * Wo + x) * Co;
#pragma unroll
      for (int j = 0; j < CPL; ++j) dst[lane + 32 * j] += acc[j];
    }
  }
}

// dU partials: one CTA per (n*NS + s, chunk of SC_BR output rows of the object's window); a lane keeps
// K*K taps x CPL channels of accumulators; cross-warp sum in warp order (deterministic).
template <int K, int S, int P, int CPL>
__global__ void __launch_bounds__(SC_THREADS)
layout_sconv_bwd_kernel(const float* __restrict__ dout, ScTables tb, int NS, int H, int W, int Ho, int Wo,
                        int max_chunks, float* __restrict__ part /*[N*NS][max_chunks][KK][Co]*/) {
  constexpr int Co = 32 * CPL;
  constexpr int KK = K * K;
  extern __shared__ __align__(16) float sc_smem[];           // [KK][Co]
  const int ns = blockIdx.x, chunk = blockIdx.y;
  const int n = ns / NS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int4 r = tb.range[ns];
  float* dst = part + ((size_t)ns * max_chunks + chunk) * KK * Co;
  const bool empty = !(r.y > r.x && r.w > r.z);
  // output pixels with at least one tap inside the support: S*p + k - P in [lo, hi) for some k in [0, K)
  const int ylo = empty ? 0 : max((r.z + P - (K - 1) + S - 1) / S, 0), yhi = empty ? 0 : min((r.w - 1 + P) / S + 1, Ho);
  const int xlo = empty ? 0 : max((r.x + P - (K - 1) + S - 1) / S, 0), xhi = empty ? 0 : min((r.y - 1 + P) / S + 1, Wo);
  const int yc0 = ylo + chunk * SC_BR, yc1 = min(yc0 + SC_BR, yhi);
  if (empty || yc0 >= yhi || xhi <= xlo) {
    for (int i = tid; i < KK * Co; i += SC_THREADS) dst[i] = 0.f;
    return;
  }
  const float* wy = tb.wy + (size_t)ns * H;
  const float* wx = tb.wx + (size_t)ns * W;
  float acc[KK][CPL];
#pragma unroll
  for (int k = 0; k < KK; ++k)
#pragma unroll
    for (int j = 0; j < CPL; ++j) acc[k][j] = 0.f;
  const int wpix = xhi - xlo, npix = (yc1 - yc0) * wpix;
  for (int i = warp; i < npix; i += SC_THREADS / 32) {
    const int y = yc0 + i / wpix, x = xlo + i % wpix;
    float wyk[K], wxk[K];
#pragma unroll
    for (int d = 0; d < K; ++d) {
      const int yy = S * y + d - P, xx = S * x + d - P;
      wyk[d] = (yy >= 0 && yy < H) ? wy[yy] : 0.f;
      wxk[d] = (xx >= 0 && xx < W) ? wx[xx] : 0.f;
    }
    const float* src = dout + (((size_t)n * Ho + y) * Wo + x) * Co;
    float g[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) g[j] = src[lane + 32 * j];
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      const float c = wyk[k / K] * wxk[k % K];
#pragma unroll
      for (int j = 0; j < CPL; ++j) acc[k][j] = fmaf(c, g[j], acc[k][j]);
    }
  }
  for (int w = 0; w < SC_THREADS / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int k = 0; k < KK; ++k)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          float* q = sc_smem + k * Co + lane + 32 * j;
          *q = (w == 0 ? 0.f : *q) + acc[k][j];
        }
    }
    __syncthreads();
  }
  for (int i = tid; i < KK * Co; i += SC_THREADS) dst[i] = sc_smem[i];
}

__global__ void layout_sconv_reduce_kernel(const float* __restrict__ part, int max_chunks, int per, long long total,
                                           float* __restrict__ dU) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long ns = i / per; const int e = (int)(i - ns * per);
    float acc = 0.f;
    for (int c = 0; c < max_chunks; ++c) acc += part[((size_t)ns * max_chunks + c) * per + e];
    dU[i] = acc;
  }
}

struct ScWs { float* wx; float* wy; int4* range; float* scale; };

// same carving as layout_ws_carve (k2_layout.cu) / ag2v_boxes_to_layout_workspace_bytes
static ScWs sc_carve(void* ws, int N, int O, int H, int W) {
  ScWs r;
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

static int sc_out(int n, int K, int S, int P) { return (n + 2 * P - K) / S + 1; }
static int sc_chunks(int Ho) { return ceil_div(Ho, SC_BR); }

// tables_out (workspace of ag2v_boxes_to_layout_workspace_bytes(N,O,H2,W2), H2 = (H-1)/2+1) = the
// separable tables of avg_pool2d(mask, 3, stride 2, padding 1, count_include_pad=False).
extern "C" int ag2v_layout_tables_avgpool(const void* tables, int N, int O, int H, int W, void* tables_out,
                                          cudaStream_t stream) {
  AG2V_REQUIRE(tables && tables_out && N > 0 && O > 0 && H > 0 && W > 0, "layout_tables_avgpool: bad arguments");
  const int H2 = (H - 1) / 2 + 1, W2 = (W - 1) / 2 + 1;
  ScWs a = sc_carve(const_cast<void*>(tables), N, O, H, W), b = sc_carve(tables_out, N, O, H2, W2);
  layout_tables_pool_kernel<<<N * O, 128, 0, stream>>>(a.wx, a.wy, H, W, H2, W2, b.wx, b.wy, b.range);
  AG2V_LAUNCH_CHECK();
  AG2V_CUDA(cudaMemcpyAsync(b.scale, a.scale, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  return AG2V_OK;
}

extern "C" size_t ag2v_layout_sconv_bwd_workspace_floats(int N, int S, int Co, int Ho, int KK) {
  return (size_t)N * S * sc_chunks(Ho) * KK * Co;
}

// out [N,Ho,Wo,Co] (NHWC) += sum_s sum_k U[n,s,k,:] * m_s(stride*p + k - pad); tables for boxes [N,S,4] at
// the convolution's input resolution (H,W).  Built: kernel 4, stride 2, pad 2, Co = 64 (the PatchGAN stem).
extern "C" int ag2v_layout_sconv_fwd(const float* U, const void* tables, int N, int S, int Co, int H, int W,
                                     int kernel, int stride, int pad, float* out, cudaStream_t stream) {
  AG2V_REQUIRE(U && tables && out && N > 0 && S > 0 && S <= 64 && H > 0 && W > 0, "layout_sconv_fwd: bad arguments");
  AG2V_REQUIRE(kernel == 4 && stride == 2 && pad == 2 && Co == 64,
               "layout_sconv_fwd: built for kernel 4 / stride 2 / pad 2 / Co 64 (got %d/%d/%d/%d)", kernel, stride, pad, Co);
  ScWs w = sc_carve(const_cast<void*>(tables), N, S, H, W);
  ScTables tb{w.wx, w.wy, w.range};
  const int Ho = sc_out(H, 4, 2, 2), Wo = sc_out(W, 4, 2, 2);
  dim3 grid(ceil_div(Wo, SC_TX), ceil_div(Ho, SC_TY), N);
  const size_t smem = ((size_t)SC_MAXACT * 16 * Co + SC_MAXACT * (2 * (SC_TY - 1) + 4 + 2 * (SC_TX - 1) + 4)) * sizeof(float);
  layout_sconv_fwd_kernel<4, 2, 2, 2><<<grid, SC_THREADS, smem, stream>>>(U, tb, S, H, W, Ho, Wo, out);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// dU [N,S,K*K,Co] = sum_p dout[n,p,:] * m_s(stride*p + k - pad); part = ag2v_layout_sconv_bwd_workspace_floats scratch
extern "C" int ag2v_layout_sconv_bwd(const float* dout, const void* tables, int N, int S, int Co, int H, int W,
                                     int kernel, int stride, int pad, float* part, float* dU, cudaStream_t stream) {
  AG2V_REQUIRE(dout && tables && part && dU && N > 0 && S > 0 && H > 0 && W > 0, "layout_sconv_bwd: bad arguments");
  AG2V_REQUIRE(kernel == 4 && stride == 2 && pad == 2 && Co == 64,
               "layout_sconv_bwd: built for kernel 4 / stride 2 / pad 2 / Co 64 (got %d/%d/%d/%d)", kernel, stride, pad, Co);
  ScWs w = sc_carve(const_cast<void*>(tables), N, S, H, W);
  ScTables tb{w.wx, w.wy, w.range};
  const int Ho = sc_out(H, 4, 2, 2), Wo = sc_out(W, 4, 2, 2);
  const int chunks = sc_chunks(Ho);
  dim3 grid(N * S, chunks);
  layout_sconv_bwd_kernel<4, 2, 2, 2><<<grid, SC_THREADS, 16 * Co * sizeof(float), stream>>>(dout, tb, S, H, W, Ho, Wo, chunks, part);
  AG2V_LAUNCH_CHECK();
  const long long total = (long long)N * S * 16 * Co;
  layout_sconv_reduce_kernel<<<(unsigned)(ceil_div_ll(total, 256) > 4096 ? 4096 : ceil_div_ll(total, 256)), 256, 0, stream>>>(
      part, chunks, 16 * Co, total, dU);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}
