// K1 — one action-graph convolution layer, forward and backward, each as ONE
// cooperative kernel (reference: models/graph_models/graph.py:41-107).
//
//   gather [obj[s] | pred | obj[o]]  ->  net1 (Linear-ReLU-Linear-ReLU)  ->
//   masked scatter-average to nodes (subjects first, then objects, edge order;
//   deterministic, no atomics)  ->  net2 (Linear-ReLU-Linear-ReLU)
//
// The layer has M = B*E <= ~128 edge rows and R = B*O <= ~32 node rows against
// 4.5-6 MB of fp32 weights: it is weight-streaming / latency bound, not a
// tensor-throughput problem.  The whole grid (one CTA per SM) walks the stages
// with grid-wide barriers in between; every stage is a skinny GEMM whose weight
// matrix is streamed exactly once from HBM/L2, split over the CTAs by output
// columns.  Products run on mma.sync m16n8k8 with the 3xTF32 split (hi/lo) so
// results are fp32-accurate (~1e-6), which the box recurrence needs.
#include <cooperative_groups.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cg = cooperative_groups;

namespace ag2v {

struct GcnDims {
  int B, O, E, Din, Dp, H, Dout, Dpo;
  __host__ __device__ int M() const { return B * E; }
  __host__ __device__ int R() const { return B * O; }
  __host__ __device__ int K1() const { return 2 * Din + Dp; }
  __host__ __device__ int N2() const { return 2 * H + Dpo; }
};

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kMChunk = 128;          // rows held in registers per pass
constexpr int kMT = kMChunk / 16;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  float rem = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// c += a * b with fp32-class accuracy: small cross terms first.
__device__ __forceinline__ void mma_3x(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}

// ---------------------------------------------------------------------------
// CTA-level skinny product:  C[m, n] = sum_r A(m, r) * B(r, n),  m < M, all n.
// CTAs take 8-column tiles of n.  The activation block A[:, k-chunk] is staged in
// shared memory with bulk async copies (cp.async.bulk + mbarrier: one instruction per
// row segment, no register staging), the CTA's weight slab is prefetched into
// registers while those copies fly, then the 8 warps split the reduction in steps of 8;
// their partials are summed in warp order through shared memory (deterministic).
//   A::issue(m, k0, kn, dst, bar)  bulk-copies row m, columns [k0, k0+kn) to smem (kBulk)
//   A(m, r) -> (A(m,r), A(m,r+1))  element access for the non-bulk (masked) operands
//   B2(r, n) -> (B(r,n), B(r+1,n)),  r even.
// ---------------------------------------------------------------------------
constexpr int kKc = 1024;             // reduction columns staged per pass
constexpr int kXsFloats = 36 * 1024;  // 144 KB activation staging buffer
constexpr int kWsFloats = 8 * (kKc + 8);   // the CTA's 8-column weight slab for one pass (33 KB)

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct StageCtx { float* red; float* xs; float* ws; uint32_t bar; uint32_t phase; };

// compact inner loop (kept small on purpose: a fully unrolled version is ~0.5 MB of SASS and
// the kernel becomes instruction-fetch bound): MT m16 tiles of accumulators in registers.
template <int MT>
__device__ __forceinline__ void skinny_core(const float* xs, const float* ws, int ld, int mrows, int ksteps,
                                            float (&acc)[kMT][4]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int ks = warp; ks < ksteps; ks += kWarps) {
    const float2 wv = *reinterpret_cast<const float2*>(ws + g * ld + (ks << 3) + 2 * t);
    const float b[2] = {wv.x, wv.y};
    const float* col = xs + (ks << 3) + 2 * t;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int r_lo = mt * 16 + g, r_hi = r_lo + 8;
      const float2 p = (r_lo < mrows) ? *reinterpret_cast<const float2*>(col + r_lo * ld) : make_float2(0.f, 0.f);
      const float2 q = (r_hi < mrows) ? *reinterpret_cast<const float2*>(col + r_hi * ld) : make_float2(0.f, 0.f);
      const float a[4] = {p.x, q.x, p.y, q.y};
      mma_3x(acc[mt], a, b);
    }
  }
}

template <class A2, class B2, class Epi>
__device__ void cta_skinny_gemm(int M, int N, int Kred, A2 a2, B2 b2, Epi epi, StageCtx& cx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int ntiles = N >> 3;
  float* red = cx.red;
  float* xs = cx.xs;
  float* ws = cx.ws;
  const uint32_t xs_u32 = smem_u32(xs), ws_u32 = smem_u32(ws);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n0 = tile << 3;
    for (int mbase = 0; mbase < M; mbase += kMChunk) {
      const int mrows = min(kMChunk, M - mbase);
      const int mt_n = (mrows + 15) >> 4;
      int kc_max = (kXsFloats / (mt_n * 16) - 8) & ~7;
      if (kc_max > kKc) kc_max = kKc;
      float acc[kMT][4];
#pragma unroll
      for (int i = 0; i < kMT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
      for (int k0 = 0; k0 < Kred; k0 += kc_max) {
        const int kn = min(kc_max, Kred - k0);
        const int ld = kn + 8;                          // +8 floats: conflict-free 64-bit fragment reads
        // ---- stage A[mbase : mbase+mrows, k0 : k0+kn] and the weight slab B[k0 : k0+kn, n0 : n0+8] ----
        const uint32_t tx_bytes = (A2::kBulk ? (uint32_t)(mrows * kn * 4) : 0u) + (B2::kBulk ? (uint32_t)(8 * kn * 4) : 0u);
        if (tx_bytes) {
          if (threadIdx.x == 0) mbar_expect_tx(cx.bar, tx_bytes);
          __syncthreads();
        }
        if (A2::kBulk) {
          if ((int)threadIdx.x < mrows) a2.issue(mbase + threadIdx.x, k0, kn, xs_u32 + (uint32_t)(threadIdx.x * ld) * 4, cx.bar);
        } else {
          const int half = kn >> 1;
          for (int i = threadIdx.x; i < mrows * half; i += kThreads) {
            const int m = i / half, c2 = (i - m * half) * 2;
            *reinterpret_cast<float2*>(xs + m * ld + c2) = a2(mbase + m, k0 + c2);
          }
        }
        if (B2::kBulk) {
          if (threadIdx.x >= 128 && threadIdx.x < 136) b2.issue(n0 + threadIdx.x - 128, k0, kn, ws_u32 + (uint32_t)((threadIdx.x - 128) * ld) * 4, cx.bar);
        } else {
          for (int r = threadIdx.x; r < kn; r += kThreads) b2.store_row(k0 + r, n0, ws + r, ld);
        }
        if (tx_bytes) { mbar_wait(cx.bar, cx.phase); cx.phase ^= 1u; }
        if (!A2::kBulk || !B2::kBulk) __syncthreads();
        const int ksteps = kn >> 3;
        if (mt_n <= 2) skinny_core<2>(xs, ws, ld, mrows, ksteps, acc);
        else if (mt_n <= 5) skinny_core<5>(xs, ws, ld, mrows, ksteps, acc);
        else skinny_core<8>(xs, ws, ld, mrows, ksteps, acc);
        __syncthreads();                                 // everyone is done with xs / ws before the next pass lands
      }
#pragma unroll
      for (int mt = 0; mt < kMT; ++mt) {
        if (mt < mt_n) {
          float* dst = red + ((size_t)warp * kMChunk + mt * 16 + g) * 8 + 2 * t;
          *reinterpret_cast<float2*>(dst) = make_float2(acc[mt][0], acc[mt][1]);
          *reinterpret_cast<float2*>(dst + 64) = make_float2(acc[mt][2], acc[mt][3]);
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < mrows * 8; i += kThreads) {
        float sum = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < kWarps; ++w8) sum += red[(size_t)w8 * kMChunk * 8 + i];
        epi(mbase + (i >> 3), n0 + (i & 7), sum);
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------
// Warp-level weight gradient:  dW[n, k] = sum_{m<M} Z(m, n) * X(m, k)  and
// db[n] = sum_m Z(m, n).  Task = (16 rows of n) x (64 columns of k); tasks are
// spread over all warps of the grid.  Z(m, n), X(m, k) are scalar accessors.
// ---------------------------------------------------------------------------
template <class Z, class X>
__device__ void warp_wgrad(int M, int N, int K, Z z, X x, float* __restrict__ dW,
                           float* __restrict__ db, int busy_tiles, int total_warps, int gwarp) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int ntiles = (N + 15) >> 4, kchunks = (K + 63) >> 6;
  const int tasks = ntiles * kchunks;
  // CTAs [0, busy_tiles) are working on the skinny GEMM of the same phase: hand the
  // first tasks to the warps of the idle CTAs
  const int task_offset = (busy_tiles >= (int)gridDim.x) ? 0 : busy_tiles * kWarps;
  for (int task = (gwarp + total_warps - task_offset) % total_warps; task < tasks; task += total_warps) {
    const int nt = task / kchunks, kc = task - nt * kchunks;
    const int n_lo = nt * 16 + g, n_hi = n_lo + 8;
    const int k0 = kc << 6;
    const int kt_n = min(8, (K - k0) >> 3);
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    float s_lo = 0.f, s_hi = 0.f;
    for (int m0 = 0; m0 < M; m0 += 8) {
      const int m = m0 + 2 * t;
      float2 p = make_float2(0.f, 0.f), q = make_float2(0.f, 0.f);
      const bool r0 = m < M, r1 = m + 1 < M;
      if (n_lo < N) { if (r0) p.x = z(m, n_lo); if (r1) p.y = z(m + 1, n_lo); }
      if (n_hi < N) { if (r0) q.x = z(m, n_hi); if (r1) q.y = z(m + 1, n_hi); }
      const float a[4] = {p.x, q.x, p.y, q.y};
      s_lo += p.x + p.y; s_hi += q.x + q.y;
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
        if (kt < kt_n) {
          float2 xb = make_float2(0.f, 0.f);
          if (r0) xb.x = x(m, k0 + kt * 8 + g);
          if (r1) xb.y = x(m + 1, k0 + kt * 8 + g);
          const float b[2] = {xb.x, xb.y};
          mma_3x(acc[kt], a, b);
        }
      }
    }
#pragma unroll
    for (int kt = 0; kt < 8; ++kt) {
      if (kt < kt_n) {
        const int k = k0 + kt * 8 + 2 * t;
        if (n_lo < N) *reinterpret_cast<float2*>(dW + (size_t)n_lo * K + k) = make_float2(acc[kt][0], acc[kt][1]);
        if (n_hi < N) *reinterpret_cast<float2*>(dW + (size_t)n_hi * K + k) = make_float2(acc[kt][2], acc[kt][3]);
      }
    }
    if (kc == 0 && db != nullptr) {
      s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 1); s_lo += __shfl_xor_sync(0xffffffffu, s_lo, 2);
      s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 1); s_hi += __shfl_xor_sync(0xffffffffu, s_hi, 2);
      if (t == 0) { if (n_lo < N) db[n_lo] = s_lo; if (n_hi < N) db[n_hi] = s_hi; }
    }
  }
}

// ---------------------------------------------------------------------------
// Deterministic incident-edge sums (the scatter_add of graph.py:89-91 and its
// transpose).  One warp builds, in its private shared list, the rows of the edges
// touching `node` — subject hits in edge order, then object hits in edge order (bit
// 30 marks an object hit) — and then adds the source rows in exactly that order,
// four independent loads in flight at a time.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int build_incident(int node, int b, int E, const int* s_idx, const int* o_idx,
                                              const uint8_t* __restrict__ ind, int* lst) {
  const int lane = threadIdx.x & 31;
  int n = 0;
  for (int pass = 0; pass < 2; ++pass) {
    const int* idx = pass ? o_idx : s_idx;
    for (int e0 = 0; e0 < E; e0 += 32) {
      const int e = e0 + lane, m = b * E + e;
      const bool hit = e < E && (ind == nullptr || ind[m] != 0) && idx[m] == node;
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      if (hit) lst[n + __popc(mask & ((1u << lane) - 1u))] = pass ? (m | (1 << 30)) : m;
      n += __popc(mask);
    }
  }
  __syncwarp();
  return n;
}

// acc[c..c+3] = sum over the list of src[row, (object ? off_o : off_s) + c..c+3]
__device__ __forceinline__ float4 incident_sum4(const float* src, int ld, int off_s, int off_o, int c,
                                                const int* lst, int n) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = 0; i < n; i += 4) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i + j < n) {
        const int ent = lst[i + j];
        const int m = ent & ~(1 << 30);
        v[j] = *reinterpret_cast<const float4*>(src + (size_t)m * ld + ((ent >> 30) ? off_o : off_s) + c);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
  }
  return acc;
}

struct GcnFwd {
  GcnDims d;
  const float *obj, *pred; const long long* edges; const uint8_t* ind;
  const float *W1a, *b1a, *W1b, *b1b, *W2a, *b2a, *W2b, *b2b;
  float *new_obj, *new_p;
  float *h1, *h2, *pooled, *g1, *cnt;      // saved for backward
};

struct GcnBwd {
  GcnDims d;
  const float *obj, *pred; const long long* edges; const uint8_t* ind;
  const float *W1a, *W1b, *W2a, *W2b;
  const float *new_obj, *h1, *h2, *pooled, *g1, *cnt;
  const float *d_new_obj, *d_new_p;        // d_new_p may be null (return_new_p_vecs=False)
  float *dz4, *dpooled, *dz2, *dz1, *dT;   // workspace
  float *dobj, *dpred;
  float *dW1a, *db1a, *dW1b, *db1b, *dW2a, *db2a, *dW2b, *db2b;
};

__device__ __forceinline__ int clamp_idx(long long v, int O) {
  return v < 0 ? 0 : (v >= O ? O - 1 : (int)v);
}

// [obj[b, s_e] | pred[b, e] | obj[b, o_e]] row m = (b, e), columns r, r+1 (graph.py:67-70)
struct GatherT {
  static constexpr bool kBulk = true;
  const float *obj, *pred; const int *s_idx, *o_idx; int Din, Dp;
  // columns [k0, k0+kn) of the virtual row [obj[s] | pred | obj[o]]: up to three contiguous pieces
  __device__ __forceinline__ void issue(int m, int k0, int kn, uint32_t dst, uint32_t bar) const {
    const int k1 = k0 + kn;
    const int segb[4] = {0, Din, Din + Dp, 2 * Din + Dp};
    const float* srcs[3] = {obj + (size_t)s_idx[m] * Din, pred + (size_t)m * Dp, obj + (size_t)o_idx[m] * Din};
#pragma unroll
    for (int sgm = 0; sgm < 3; ++sgm) {
      const int lo = max(k0, segb[sgm]), hi = min(k1, segb[sgm + 1]);
      if (hi > lo) bulk_g2s(dst + (uint32_t)(lo - k0) * 4, srcs[sgm] + (lo - segb[sgm]), (uint32_t)(hi - lo) * 4, bar);
    }
  }
  __device__ __forceinline__ float2 operator()(int m, int r) const {
    const float* src;
    if (r < Din) src = obj + (size_t)s_idx[m] * Din + r;
    else if (r < Din + Dp) src = pred + (size_t)m * Dp + (r - Din);
    else src = obj + (size_t)o_idx[m] * Din + (r - Din - Dp);
    return *reinterpret_cast<const float2*>(src);
  }
};
struct RowMajor2 {      // X[m, r..r+1], contiguous
  static constexpr bool kBulk = true;
  const float* x; int ld;
  __device__ __forceinline__ void issue(int m, int k0, int kn, uint32_t dst, uint32_t bar) const {
    bulk_g2s(dst, x + (size_t)m * ld + k0, (uint32_t)kn * 4, bar);
  }
  __device__ __forceinline__ float2 operator()(int m, int r) const {
    return *reinterpret_cast<const float2*>(x + (size_t)m * ld + r);
  }
};
struct RowMajorMasked2 {  // X[m, r] where gate[m, r] > 0, else 0
  static constexpr bool kBulk = false;
  const float *x, *gate; int ld;
  __device__ __forceinline__ void issue(int, int, int, uint32_t, uint32_t) const {}
  __device__ __forceinline__ float2 operator()(int m, int r) const {
    float2 v = *reinterpret_cast<const float2*>(x + (size_t)m * ld + r);
    float2 q = *reinterpret_cast<const float2*>(gate + (size_t)m * ld + r);
    return make_float2(q.x > 0.f ? v.x : 0.f, q.y > 0.f ? v.y : 0.f);
  }
};
struct WeightNT2 {      // B(r, n) = W[n, r]  (forward: reduce along W's contiguous dim)
  static constexpr bool kBulk = true;
  const float* w; int ld;
  // row n of the slab = W[n, k0 : k0+kn], contiguous
  __device__ __forceinline__ void issue(int n, int k0, int kn, uint32_t dst, uint32_t bar) const {
    bulk_g2s(dst, w + (size_t)n * ld + k0, (uint32_t)kn * 4, bar);
  }
  __device__ __forceinline__ void store_row(int, int, float*, int) const {}
  __device__ __forceinline__ float2 operator()(int r, int n) const {
    return *reinterpret_cast<const float2*>(w + (size_t)n * ld + r);
  }
};
struct WeightNN2 {      // B(r, n) = W[r, n]  (input gradient: reduce along W's rows)
  static constexpr bool kBulk = false;
  const float* w; int ld;
  __device__ __forceinline__ void issue(int, int, int, uint32_t, uint32_t) const {}
  // W[r, n0 : n0+8] (32 contiguous bytes) transposed into the [8][ld] slab at column `dst - slab`
  __device__ __forceinline__ void store_row(int r, int n0, float* dst, int ldw) const {
    const float4 lo = *reinterpret_cast<const float4*>(w + (size_t)r * ld + n0);
    const float4 hi = *reinterpret_cast<const float4*>(w + (size_t)r * ld + n0 + 4);
    dst[0] = lo.x; dst[ldw] = lo.y; dst[2 * ldw] = lo.z; dst[3 * ldw] = lo.w;
    dst[4 * ldw] = hi.x; dst[5 * ldw] = hi.y; dst[6 * ldw] = hi.z; dst[7 * ldw] = hi.w;
  }
  __device__ __forceinline__ float2 operator()(int r, int n) const {
    return make_float2(w[(size_t)r * ld + n], w[(size_t)(r + 1) * ld + n]);
  }
};
struct Elem {            // X[m, c] for the weight-gradient operands
  const float* x; int ld;
  __device__ __forceinline__ float operator()(int m, int c) const { return x[(size_t)m * ld + c]; }
};
struct ElemMasked {
  const float *x, *gate; int ld;
  __device__ __forceinline__ float operator()(int m, int c) const {
    size_t i = (size_t)m * ld + c;
    return gate[i] > 0.f ? x[i] : 0.f;
  }
};
struct ElemGatherT {
  const float *obj, *pred; const int *s_idx, *o_idx; int Din, Dp;
  __device__ __forceinline__ float operator()(int m, int c) const {
    if (c < Din) return obj[(size_t)s_idx[m] * Din + c];
    if (c < Din + Dp) return pred[(size_t)m * Dp + (c - Din)];
    return obj[(size_t)o_idx[m] * Din + (c - Din - Dp)];
  }
};

extern __shared__ __align__(16) unsigned char gcn_smem[];

__device__ __forceinline__ void load_indices(const GcnDims& d, const long long* edges, int* s_idx, int* o_idx) {
  for (int m = threadIdx.x; m < d.M(); m += kThreads) {
    const int b = m / d.E;
    s_idx[m] = b * d.O + clamp_idx(edges[(size_t)m * 2 + 0], d.O);
    o_idx[m] = b * d.O + clamp_idx(edges[(size_t)m * 2 + 1], d.O);
  }
  // one padded row so (m + 1) operand pairs never index out of bounds
  if (threadIdx.x == 0) { s_idx[d.M()] = 0; o_idx[d.M()] = 0; }
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads, 1) gcn_fwd_kernel(GcnFwd p) {
  cg::grid_group grid = cg::this_grid();
  const GcnDims d = p.d;
  const int M = d.M(), R = d.R(), K1 = d.K1(), N2 = d.N2(), H = d.H;
  float* red = reinterpret_cast<float*>(gcn_smem);
  float* xs = red + kWarps * kMChunk * 8;
  float* ws = xs + kXsFloats;
  uint64_t* barp = reinterpret_cast<uint64_t*>(ws + kWsFloats);
  int* s_idx = reinterpret_cast<int*>(barp + 2);
  int* o_idx = s_idx + M + 1;
  StageCtx cx{red, xs, ws, smem_u32(barp), 0u};
  if (threadIdx.x == 0) { mbar_init(cx.bar, 1); fence_barrier_init(); }
  load_indices(d, p.edges, s_idx, o_idx);

  {  // stage 1: h1 = relu(T W1a^T + b1a)                       graph.py:71 (net1[0:2])
    GatherT a{p.obj, p.pred, s_idx, o_idx, d.Din, d.Dp};
    WeightNT2 b{p.W1a, K1};
    const float* bias = p.b1a; float* out = p.h1;
    cta_skinny_gemm(M, H, K1, a, b, [=](int m, int n, float v) { out[(size_t)m * H + n] = fmaxf(v + bias[n], 0.f); }, cx);
  }
  grid.sync();
  {  // stage 2: h2 = relu(h1 W1b^T + b1b); the middle slice is new_p   graph.py:71-75
    RowMajor2 a{p.h1, H};
    WeightNT2 b{p.W1b, H};
    const float* bias = p.b1b; float* out = p.h2; float* np = p.new_p; const int Dpo = d.Dpo;
    cta_skinny_gemm(M, N2, H, a, b, [=](int m, int n, float v) {
      v = fmaxf(v + bias[n], 0.f);
      out[(size_t)m * N2 + n] = v;
      if (n >= H && n < H + Dpo) np[(size_t)m * Dpo + (n - H)] = v;
    }, cx);
  }
  grid.sync();
  {  // stage 3: masked average pooling, subjects then objects, edge order   graph.py:79-100
    int* lst = reinterpret_cast<int*>(red) + (threadIdx.x >> 5) * 1024;
    const int groups = (H + 127) >> 7, lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * kWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kWarps;
    for (int task = gwarp; task < R * groups; task += nwarps) {
      const int node = task / groups, c = (task - node * groups) * 128 + lane * 4;
      const int n = build_incident(node, node / d.O, d.E, s_idx, o_idx, p.ind, lst);
      if (c < H) {
        float4 acc = incident_sum4(p.h2, N2, 0, H + d.Dpo, c, lst, n);
        if (n > 0) { const float cnt = (float)n; acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt; }
        *reinterpret_cast<float4*>(p.pooled + (size_t)node * H + c) = acc;
      }
      if (c == 0) p.cnt[node] = (float)n;
      __syncwarp();
    }
  }
  grid.sync();
  {  // stage 4: g1 = relu(pooled W2a^T + b2a)                   graph.py:103 (net2[0:2])
    RowMajor2 a{p.pooled, H};
    WeightNT2 b{p.W2a, H};
    const float* bias = p.b2a; float* out = p.g1;
    cta_skinny_gemm(R, H, H, a, b, [=](int m, int n, float v) { out[(size_t)m * H + n] = fmaxf(v + bias[n], 0.f); }, cx);
  }
  grid.sync();
  {  // stage 5: new_obj = relu(g1 W2b^T + b2b)
    RowMajor2 a{p.g1, H};
    WeightNT2 b{p.W2b, H};
    const float* bias = p.b2b; float* out = p.new_obj; const int Dout = d.Dout;
    cta_skinny_gemm(R, Dout, H, a, b, [=](int m, int n, float v) { out[(size_t)m * Dout + n] = fmaxf(v + bias[n], 0.f); }, cx);
  }
}

__global__ void __launch_bounds__(kThreads, 1) gcn_bwd_kernel(GcnBwd p) {
  cg::grid_group grid = cg::this_grid();
  const GcnDims d = p.d;
  const int M = d.M(), R = d.R(), K1 = d.K1(), N2 = d.N2(), H = d.H, Dout = d.Dout, Dpo = d.Dpo;
  float* red = reinterpret_cast<float*>(gcn_smem);
  float* xs = red + kWarps * kMChunk * 8;
  float* ws = xs + kXsFloats;
  uint64_t* barp = reinterpret_cast<uint64_t*>(ws + kWsFloats);
  int* s_idx = reinterpret_cast<int*>(barp + 2);
  int* o_idx = s_idx + M + 1;
  StageCtx cx{red, xs, ws, smem_u32(barp), 0u};
  if (threadIdx.x == 0) { mbar_init(cx.bar, 1); fence_barrier_init(); }
  load_indices(d, p.edges, s_idx, o_idx);
  const int total_warps = gridDim.x * kWarps;
  const int gwarp = blockIdx.x * kWarps + (threadIdx.x >> 5);

  {  // A: through net2's second Linear.  dz5 = d_new_obj * [new_obj > 0]
    RowMajorMasked2 a{p.d_new_obj, p.new_obj, Dout};
    WeightNN2 b{p.W2b, H};
    const float* gate = p.g1; float* out = p.dz4;
    cta_skinny_gemm(R, H, Dout, a, b, [=](int m, int n, float v) {
      size_t i = (size_t)m * H + n; out[i] = gate[i] > 0.f ? v : 0.f; }, cx);
    ElemMasked z{p.d_new_obj, p.new_obj, Dout};
    Elem x{p.g1, H};
    warp_wgrad(R, Dout, H, z, x, p.dW2b, p.db2b, H >> 3, total_warps, gwarp);
  }
  grid.sync();
  {  // B: through net2's first Linear
    RowMajor2 a{p.dz4, H};
    WeightNN2 b{p.W2a, H};
    float* out = p.dpooled;
    cta_skinny_gemm(R, H, H, a, b, [=](int m, int n, float v) { out[(size_t)m * H + n] = v; }, cx);
    Elem z{p.dz4, H};
    Elem x{p.pooled, H};
    warp_wgrad(R, H, H, z, x, p.dW2a, p.db2a, H >> 3, total_warps, gwarp);
  }
  grid.sync();
  {  // C: transpose of the pooling, fold in d_new_p, gate by relu of h2
    const int total = M * N2;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < total; i += gridDim.x * kThreads) {
      const int m = i / N2, c = i - m * N2;
      float v = 0.f;
      if (c < H) {
        if (p.ind[m]) { const int node = s_idx[m]; v = p.dpooled[(size_t)node * H + c] / fmaxf(p.cnt[node], 1.f); }
      } else if (c < H + Dpo) {
        if (p.d_new_p != nullptr) v = p.d_new_p[(size_t)m * Dpo + (c - H)];
      } else {
        if (p.ind[m]) { const int node = o_idx[m]; v = p.dpooled[(size_t)node * H + (c - H - Dpo)] / fmaxf(p.cnt[node], 1.f); }
      }
      p.dz2[i] = p.h2[i] > 0.f ? v : 0.f;
    }
  }
  grid.sync();
  {  // D: through net1's second Linear
    RowMajor2 a{p.dz2, N2};
    WeightNN2 b{p.W1b, H};
    const float* gate = p.h1; float* out = p.dz1;
    cta_skinny_gemm(M, H, N2, a, b, [=](int m, int n, float v) {
      size_t i = (size_t)m * H + n; out[i] = gate[i] > 0.f ? v : 0.f; }, cx);
    Elem z{p.dz2, N2};
    Elem x{p.h1, H};
    warp_wgrad(M, N2, H, z, x, p.dW1b, p.db1b, H >> 3, total_warps, gwarp);
  }
  grid.sync();
  {  // E: through net1's first Linear; dT holds gradients of the gathered triple rows
    RowMajor2 a{p.dz1, H};
    WeightNN2 b{p.W1a, K1};
    float* out = p.dT;
    cta_skinny_gemm(M, K1, H, a, b, [=](int m, int n, float v) { out[(size_t)m * K1 + n] = v; }, cx);
    Elem z{p.dz1, H};
    ElemGatherT x{p.obj, p.pred, s_idx, o_idx, d.Din, d.Dp};
    warp_wgrad(M, H, K1, z, x, p.dW1a, p.db1a, K1 >> 3, total_warps, gwarp);
  }
  grid.sync();
  {  // F: transpose of the gather — every edge contributes, masked or not (graph.py:67-71)
    const int Din = d.Din, Dp = d.Dp;
    __syncthreads();
    int* lst = reinterpret_cast<int*>(red) + (threadIdx.x >> 5) * 1024;
    const int groups = (Din + 127) >> 7, lane = threadIdx.x & 31;
    for (int task = gwarp; task < R * groups; task += total_warps) {
      const int node = task / groups, c = (task - node * groups) * 128 + lane * 4;
      const int n = build_incident(node, node / d.O, d.E, s_idx, o_idx, nullptr, lst);
      if (c < Din) *reinterpret_cast<float4*>(p.dobj + (size_t)node * Din + c) = incident_sum4(p.dT, K1, 0, Din + Dp, c, lst, n);
      __syncwarp();
    }
    const int total_p = M * Dp;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < total_p; i += gridDim.x * kThreads) {
      const int m = i / Dp, c = i - m * Dp;
      p.dpred[i] = p.dT[(size_t)m * K1 + Din + c];
    }
  }
}

static int gcn_check(const GcnDims& d) {
  AG2V_REQUIRE(d.B > 0 && d.O > 0 && d.E > 0, "gcn_layer: empty graph B=%d O=%d E=%d", d.B, d.O, d.E);
  AG2V_REQUIRE(d.Din % 8 == 0 && d.Dp % 8 == 0 && d.H % 8 == 0 && d.Dout % 8 == 0 && d.Dpo % 8 == 0,
               "gcn_layer: feature sizes must be multiples of 8 (Din=%d Dp=%d H=%d Dout=%d Dpo=%d)", d.Din, d.Dp, d.H, d.Dout, d.Dpo);
  AG2V_REQUIRE(d.M() <= 1024 && d.E <= 512, "gcn_layer: at most 1024 edge rows per call and 512 edges per clip (got %d, %d)", d.M(), d.E);
  AG2V_REQUIRE(d.Din % 4 == 0 && d.H % 4 == 0, "gcn_layer: Din and H must be multiples of 4");
  return AG2V_OK;
}

static size_t gcn_smem_bytes(const GcnDims& d) {
  return (size_t)kWarps * kMChunk * 8 * sizeof(float) + (size_t)(kXsFloats + kWsFloats) * sizeof(float) + 16 +
         2 * (size_t)(d.M() + 1) * sizeof(int);
}

}  // namespace ag2v

using namespace ag2v;

// Floats the caller must provide for tensors saved between forward and backward.
extern "C" size_t ag2v_gcn_layer_saved_floats(int B, int O, int E, int H, int Dpo) {
  GcnDims d{B, O, E, 0, 0, H, 0, Dpo};
  return (size_t)d.M() * H + (size_t)d.M() * d.N2() + 2 * (size_t)d.R() * H + (size_t)d.R();
}

extern "C" size_t ag2v_gcn_layer_bwd_workspace_floats(int B, int O, int E, int Din, int Dp, int H, int Dpo) {
  GcnDims d{B, O, E, Din, Dp, H, 0, Dpo};
  return 2 * (size_t)d.R() * H + (size_t)d.M() * d.N2() + (size_t)d.M() * H + (size_t)d.M() * d.K1();
}

extern "C" int ag2v_gcn_layer_fwd(const float* obj, const float* pred, const long long* edges,
                                  const uint8_t* ind, const float* W1a, const float* b1a,
                                  const float* W1b, const float* b1b, const float* W2a,
                                  const float* b2a, const float* W2b, const float* b2b, int B, int O,
                                  int E, int Din, int Dp, int H, int Dout, int Dpo, float* new_obj,
                                  float* new_p, float* saved, cudaStream_t stream) {
  GcnDims d{B, O, E, Din, Dp, H, Dout, Dpo};
  int rc = gcn_check(d);
  if (rc) return rc;
  AG2V_REQUIRE(obj && pred && edges && ind && W1a && b1a && W1b && b1b && W2a && b2a && W2b && b2b && new_obj && new_p && saved,
               "gcn_layer_fwd: null pointer");
  GcnFwd p;
  p.d = d; p.obj = obj; p.pred = pred; p.edges = edges; p.ind = ind;
  p.W1a = W1a; p.b1a = b1a; p.W1b = W1b; p.b1b = b1b; p.W2a = W2a; p.b2a = b2a; p.W2b = W2b; p.b2b = b2b;
  p.new_obj = new_obj; p.new_p = new_p;
  float* s = saved;
  p.h1 = s; s += (size_t)d.M() * H;
  p.h2 = s; s += (size_t)d.M() * d.N2();
  p.pooled = s; s += (size_t)d.R() * H;
  p.g1 = s; s += (size_t)d.R() * H;
  p.cnt = s;
  size_t smem = gcn_smem_bytes(d);
  AG2V_CUDA(cudaFuncSetAttribute(gcn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {&p};
  AG2V_COOP_LAUNCH(gcn_fwd_kernel, dim3(sm_count()), dim3(kThreads), args, smem, stream);
  return AG2V_OK;
}

extern "C" int ag2v_gcn_layer_bwd(const float* obj, const float* pred, const long long* edges,
                                  const uint8_t* ind, const float* W1a, const float* W1b,
                                  const float* W2a, const float* W2b, const float* new_obj,
                                  const float* saved, const float* d_new_obj, const float* d_new_p,
                                  int B, int O, int E, int Din, int Dp, int H, int Dout, int Dpo,
                                  float* workspace, float* dobj, float* dpred, float* dW1a, float* db1a,
                                  float* dW1b, float* db1b, float* dW2a, float* db2a, float* dW2b,
                                  float* db2b, cudaStream_t stream) {
  GcnDims d{B, O, E, Din, Dp, H, Dout, Dpo};
  int rc = gcn_check(d);
  if (rc) return rc;
  AG2V_REQUIRE(obj && pred && edges && ind && W1a && W1b && W2a && W2b && new_obj && saved && d_new_obj && workspace &&
               dobj && dpred && dW1a && db1a && dW1b && db1b && dW2a && db2a && dW2b && db2b, "gcn_layer_bwd: null pointer");
  GcnBwd p;
  p.d = d; p.obj = obj; p.pred = pred; p.edges = edges; p.ind = ind;
  p.W1a = W1a; p.W1b = W1b; p.W2a = W2a; p.W2b = W2b; p.new_obj = new_obj;
  const float* s = saved;
  p.h1 = s; s += (size_t)d.M() * H;
  p.h2 = s; s += (size_t)d.M() * d.N2();
  p.pooled = s; s += (size_t)d.R() * H;
  p.g1 = s; s += (size_t)d.R() * H;
  p.cnt = s;
  p.d_new_obj = d_new_obj; p.d_new_p = d_new_p;
  float* w = workspace;
  p.dz4 = w; w += (size_t)d.R() * H;
  p.dpooled = w; w += (size_t)d.R() * H;
  p.dz2 = w; w += (size_t)d.M() * d.N2();
  p.dz1 = w; w += (size_t)d.M() * H;
  p.dT = w;
  p.dobj = dobj; p.dpred = dpred;
  p.dW1a = dW1a; p.db1a = db1a; p.dW1b = dW1b; p.db1b = db1b;
  p.dW2a = dW2a; p.db2a = db2a; p.dW2b = dW2b; p.db2b = db2b;
  size_t smem = gcn_smem_bytes(d);
  AG2V_CUDA(cudaFuncSetAttribute(gcn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {&p};
  AG2V_COOP_LAUNCH(gcn_bwd_kernel, dim3(sm_count()), dim3(kThreads), args, smem, stream);
  return AG2V_OK;
}
