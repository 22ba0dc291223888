// K1r — the Acts2LayoutModel recurrence as ONE persistent kernel per direction (SURVEY.md section 8, row f2;
// reference: models/graph_models/model.py:126-169 with GraphTripleConv graph.py:41-107 inside).
//
//   for t = 1 .. T-1:   x   = obj_vecs_net([emb | boxes[t-1]])                 (2 Linear + ReLU, no bias)
//                       x,p = gconv_l(x, p, edges_t, ind_t)   l = 0 .. NL-1    (4 Linear + masked average pooling each)
//                       boxes[t] = boxes[t-1] + box_net(x)                     (Linear-ReLU-Linear)
//
// The chain over t is sequential (boxes[t] feeds step t+1), every step is 2 + 4 NL + 2 skinny GEMMs with
// M = E <= 16 edge rows / O <= 16 node rows against 18.4 MB of fp32 weights: pure latency.  Design:
//   * one THREAD-BLOCK CLUSTER (16 CTAs, non-portable size; fewer for narrow test widths) per chain = (model,
//     clip); independent chains (clips, the two generator-side models) run as independent clusters of one launch.
//   * every GEMM stage splits its OUTPUT columns over the CTAs of the cluster; a CTA's weight slab for every stage
//     is pre-packed (ag2v_recur_pack) into the order the FFMA register tiles read it so one bulk copy (cp.async.bulk +
//     mbarrier) brings a chunk that the warps read with conflict-free LDS.128.
//   * a dedicated producer warp streams the slabs of ALL stages of ALL timesteps through a shared-memory ring,
//     running ahead of the consumers: weights do not depend on the recurrence, so the L2 -> SM stream never waits
//     for a stage boundary.
//   * 8 consumer warps split the reduction dimension; products are exact fp32 FFMA register tiles (the box
//     recurrence amplifies rounding; and with 16 rows legacy mma.sync is latency-bound on sm_100, see run_stage),
//     partials are summed in warp order (deterministic).
//   * stages are separated by a cluster-scope mbarrier handshake (remote arrive on every peer, ~0.3 us) instead of a
//     grid-wide barrier; activations travel through L2 (they are saved for the backward anyway) and come back
//     as bulk copies.
//   * column ownership of the "split" GEMMs is permuted so that pooling (fwd) and both scatter transposes (bwd)
//     happen inside the epilogue of the CTA that owns the columns: no extra stage, no atomics, fixed edge order.
// Backward = the reverse chain for the DATA gradients only; every Linear's gated output gradient Z is stored and
// all weight gradients come from ONE grouped GEMM over all (chain, t) rows afterwards (ag2v_recur_wgrad).
#include <cooperative_groups.h>
#include "common.cuh"
#include "tc_ptx.cuh"

namespace cg = cooperative_groups;

namespace ag2v {
namespace recur {

constexpr int kMaxLayers = 4;
constexpr int kMaxStages = 2 + 4 * kMaxLayers + 2;
constexpr int kConsumers = 256;                   // 8 MMA warps
constexpr int kThreads = kConsumers + 32;         // + 1 producer warp
constexpr int kRing = 4;
constexpr int kChunkFloats = 6144;                // 24 KB ring slots (a 72-column stage uses 18 KB: 8 k-steps)
constexpr int kMaxNcta = 72;                      // columns of one stage owned by one CTA
constexpr int kMaxK = 1152;
constexpr int kLda = kMaxK + 8;
constexpr int kRows = 16;                         // one m16 tile: E <= 16 edges, O <= 16 nodes per chain
constexpr int kMaxModels = 4;

struct Dims {
  int O, E, T;                 // nodes, edges per timestep, timesteps (frames)
  int Kx, De, Dp, H, Dout, Dpo, NL;
  int CS;                      // CTAs per cluster
  __host__ __device__ int Din(int l) const { return l == 0 ? De : Dout; }
  __host__ __device__ int Dpl(int l) const { return l == 0 ? Dp : Dpo; }
  __host__ __device__ int K1(int l) const { return 2 * Din(l) + Dpl(l); }
  __host__ __device__ int N2() const { return 2 * H + Dpo; }
  __host__ __device__ int nstages() const { return 2 + 4 * NL + 1; }
};

// ---- packed model: [fwd slabs | bwd slabs | small] -----------------------------------------------------------------
struct Stage {                 // one GEMM stage: C[m, n] = sum_k A[m, k] * B(k, n)
  int N, K;                    // logical sizes
  int nseg, base[3], w[3];     // column ownership: CTA j owns columns base[s] + j*w[s] .. + w[s] of every segment s
  int ncta;                    // sum of w[s]
  int kc, nchunk;              // k-steps (of 8) per ring chunk, chunks per stage
  long long off;               // float offset of CTA 0's slab inside the packed model
  int src, ld, trans;          // source parameter index, its leading dimension, B(k,n) = trans ? W[k*ld+n] : W[n*ld+k]
};

struct Small {                 // float offsets of the small tensors inside the packed model
  long long w0box;             // [De][4]  columns Kx..Kx+3 of obj_vecs_net[0].weight
  long long bias[kMaxLayers][4];   // b1a [H], b1b [N2], b2a [H], b2b [Dout]
  long long bb1, wb2, bb2;     // box_net: [H], [4][H], [4]
  long long end;
};

struct Plan {
  Dims d;
  int nf, nb;
  Stage f[kMaxStages], b[kMaxStages];
  Small sm;
  long long total;
};

static void set_cols(Stage& s, int CS, int N0, int N1 = 0, int N2 = 0) {
  // N1 == 0: one segment of N0 columns; else three segments [N0 | N1 | N2] (subject | predicate | object parts)
  s.nseg = N1 ? 3 : 1;
  s.base[0] = 0; s.w[0] = N0 / CS;
  s.base[1] = N0; s.w[1] = N1 / CS;
  s.base[2] = N0 + N1; s.w[2] = N2 / CS;
  s.N = N0 + N1 + N2;
  s.ncta = s.w[0] + s.w[1] + s.w[2];
}

static void finish_stage(Stage& s, long long& off) {
  const int k8 = s.K / 8;
  int kc = (kChunkFloats / (s.ncta * 8)) & ~7;
  if (kc < 8) kc = 8;
  if (kc > k8) kc = k8;
  s.kc = kc;
  s.nchunk = (k8 + kc - 1) / kc;
  s.off = off;
  off += (long long)s.N * s.K;
}

// parameter order of the host pointer array: [0] ovn0.W [De][Kx+4], [1] ovn1.W [De][De], per layer
// [2+8l ..]: W1a b1a W1b b1b W2a b2a W2b b2b, then box_net: Wb1 [H][Dout], bb1, Wb2 [4][H], bb2
static int param_index_box(const Dims& d) { return 2 + 8 * d.NL; }

static int make_plan(const Dims& d, Plan& p) {
  p.d = d;
  const int CS = d.CS;
  long long off = 0;
  int n = 0;
  auto add = [&](Stage* arr, int& cnt, int src, int ld, int trans, int K, int N0, int N1, int N2) {
    Stage& s = arr[cnt++];
    s.src = src; s.ld = ld; s.trans = trans; s.K = K;
    set_cols(s, CS, N0, N1, N2);
    finish_stage(s, off);
  };
  // forward, in execution order
  add(p.f, n, 0, d.Kx + 4, 0, d.Kx, d.De, 0, 0);
  add(p.f, n, 1, d.De, 0, d.De, d.De, 0, 0);
  for (int l = 0; l < d.NL; ++l) {
    const int b = 2 + 8 * l;
    add(p.f, n, b + 0, d.K1(l), 0, d.K1(l), d.H, 0, 0);
    add(p.f, n, b + 2, d.H, 0, d.H, d.H, d.Dpo, d.H);
    add(p.f, n, b + 4, d.H, 0, d.H, d.H, 0, 0);
    add(p.f, n, b + 6, d.H, 0, d.H, d.Dout, 0, 0);
  }
  add(p.f, n, param_index_box(d), d.Dout, 0, d.Dout, d.H, 0, 0);
  p.nf = n;
  // backward (transposed products), in execution order
  n = 0;
  add(p.b, n, param_index_box(d), d.Dout, 1, d.H, d.Dout, 0, 0);
  for (int l = d.NL - 1; l >= 0; --l) {
    const int b = 2 + 8 * l;
    add(p.b, n, b + 6, d.H, 1, d.Dout, d.H, 0, 0);
    add(p.b, n, b + 4, d.H, 1, d.H, d.H, 0, 0);
    add(p.b, n, b + 2, d.H, 1, d.N2(), d.H, 0, 0);
    add(p.b, n, b + 0, d.K1(l), 1, d.H, d.Din(l), d.Dpl(l), d.Din(l));
  }
  add(p.b, n, 1, d.De, 1, d.De, d.De, 0, 0);
  add(p.b, n, 0, d.Kx + 4, 1, d.De, d.Kx, 0, 0);
  p.nb = n;
  Small& s = p.sm;
  s.w0box = off; off += (long long)d.De * 4;
  for (int l = 0; l < d.NL; ++l) {
    s.bias[l][0] = off; off += d.H;
    s.bias[l][1] = off; off += d.N2();
    s.bias[l][2] = off; off += d.H;
    s.bias[l][3] = off; off += d.Dout;
  }
  s.bb1 = off; off += d.H;
  s.wb2 = off; off += 4LL * d.H;
  s.bb2 = off; off += 8;
  s.end = off;
  p.total = (off + 7) & ~7LL;
  return AG2V_OK;
}

static bool plan_ok(const Plan& p) {
  auto ok = [&](const Stage& s) {
    if (s.K % 8 || s.K > kMaxK || s.ncta > kMaxNcta || s.ncta % 8) return false;
    for (int i = 0; i < s.nseg; ++i)
      if (s.w[i] % 8 || s.w[i] * p.d.CS != (i == 0 ? (s.nseg == 1 ? s.N : s.base[1]) : (i == 1 ? s.base[2] - s.base[1] : s.N - s.base[2])))
        return false;
    return s.kc * s.ncta * 8 <= kChunkFloats;
  };
  for (int i = 0; i < p.nf; ++i) if (!ok(p.f[i])) return false;
  for (int i = 0; i < p.nb; ++i) if (!ok(p.b[i])) return false;
  return true;
}

// ---- activations saved by the forward / gradients stored by the backward: structure of arrays, each tensor
//      [NC * (T-1)][rows][cols] so that the grouped weight-gradient GEMM sees plain row-major matrices -----------------
struct Saved {
  long long u0, x0, hb, cnt;                                         // [O][De], [O][De], [O][H], [16]
  long long part;                                                    // scratch [NC][2][16 ranks][16][4]: box partial sums
  long long p0;                                                      // [NC][O][De]: emb W0[:, :Kx]^T, the time-invariant part of obj_vecs_net[0]
  long long trow[kMaxLayers], h1[kMaxLayers], h2[kMaxLayers], pooled[kMaxLayers], g1[kMaxLayers], nobj[kMaxLayers];
  long long total;
};
struct Zbuf {
  long long zu0, zx0, zb1, zb2;                                      // [O][De], [O][De], [O][H], [O][4]
  long long part, ssum;                                              // scratch [NC][2][16][16][4]; [NC][O][De] = sum_t zu0
  long long z1[kMaxLayers], z2[kMaxLayers], z3[kMaxLayers], z4[kMaxLayers];
  long long total;
};

__host__ __device__ inline Saved saved_layout(const Dims& d, int NC) {
  Saved s;
  const long long ct = (long long)NC * (d.T - 1);
  long long off = 0;
  auto take = [&](long long rows, long long cols) { long long o = off; off += ct * rows * cols; return o; };
  s.u0 = take(d.O, d.De); s.x0 = take(d.O, d.De); s.hb = take(d.O, d.H); s.cnt = take(1, 16);
  s.part = off; off += (long long)NC * 2 * 16 * (kRows * 4);
  s.p0 = off; off += (long long)NC * d.O * d.De;
  for (int l = 0; l < d.NL; ++l) {
    s.trow[l] = take(d.E, d.K1(l)); s.h1[l] = take(d.E, d.H); s.h2[l] = take(d.E, d.N2());
    s.pooled[l] = take(d.O, d.H); s.g1[l] = take(d.O, d.H); s.nobj[l] = take(d.O, d.Dout);
  }
  s.total = off;
  return s;
}
__host__ __device__ inline Zbuf z_layout(const Dims& d, int NC) {
  Zbuf z;
  const long long ct = (long long)NC * (d.T - 1);
  long long off = 0;
  auto take = [&](long long rows, long long cols) { long long o = off; off += ct * rows * cols; return o; };
  z.zu0 = take(d.O, d.De); z.zx0 = take(d.O, d.De); z.zb1 = take(d.O, d.H); z.zb2 = take(d.O, 8);
  for (int l = 0; l < d.NL; ++l) {
    z.z1[l] = take(d.E, d.H); z.z2[l] = take(d.E, d.N2()); z.z3[l] = take(d.O, d.H); z.z4[l] = take(d.O, d.Dout);
  }
  z.part = off; off += (long long)NC * 2 * 16 * (kRows * 4);
  z.ssum = off; off += (long long)NC * d.O * d.De;
  z.total = off;
  return z;
}

// ---- device-side stage table (passed by value) ----------------------------------------------------------------------
struct StageDev { int K8, ncta, kc, nchunk, nseg, base[3], w[3]; long long off; };
struct Table { int n; StageDev s[kMaxStages]; };

static Table to_table(const Stage* st, int n) {
  Table t;
  t.n = n;
  for (int i = 0; i < n; ++i) {
    StageDev& d = t.s[i];
    d.K8 = st[i].K / 8; d.ncta = st[i].ncta; d.kc = st[i].kc; d.nchunk = st[i].nchunk; d.nseg = st[i].nseg; d.off = st[i].off;
    for (int k = 0; k < 3; ++k) { d.base[k] = st[i].base[k]; d.w[k] = st[i].w[k]; }
  }
  return t;
}

// global column of local column nl of CTA `rank`
__device__ __forceinline__ int col_of(const StageDev& s, int rank, int nl) {
  if (nl < s.w[0]) return s.base[0] + rank * s.w[0] + nl;
  nl -= s.w[0];
  if (nl < s.w[1]) return s.base[1] + rank * s.w[1] + nl;
  nl -= s.w[1];
  return s.base[2] + rank * s.w[2] + nl;
}

// ---- pack kernel: parameters -> per-CTA slabs in mma.sync B-fragment order ------------------------------------------
struct PackJob { const float* W; int ld, trans, N, K, ncta, CS, nseg, base[3], w[3]; long long off; };
struct PackArgs { int njobs; int mma; PackJob job[2 * kMaxStages]; };

__global__ void __launch_bounds__(256) recur_pack_kernel(PackArgs a, float* __restrict__ dst) {
  // slab of CTA `rank`: [k4 = K/4][j = ncta/8][ng = 8][4]  ->  B(k = 4*k4 + i, local column nl = ng + 8*j):
  // for one (k4, j) the 8 column groups are 128 contiguous bytes = one conflict-free LDS.128 per warp
  const PackJob& j = a.job[blockIdx.y];
  const long long total = (long long)j.N * j.K;
  const int slab = j.ncta * j.K, tn = j.ncta >> 3;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int rank = (int)(e / slab);
    int r = (int)(e - (long long)rank * slab);
    int nl, k;
    if (a.mma) {       // mma.sync m16n8k8 B fragments: [k8][j][lane = 4g + t][2] -> B(k = 8*k8 + 2t + i, nl = 8j + g)
      const int i = r & 1; r >>= 1;
      const int lane = r & 31; r >>= 5;
      const int jj = r % tn, k8 = r / tn;
      nl = 8 * jj + (lane >> 2);
      k = k8 * 8 + 2 * (lane & 3) + i;
    } else {
      const int i = r & 3; r >>= 2;
      const int ng = r & 7; r >>= 3;
      const int jj = r % tn, k4 = r / tn;
      nl = ng + 8 * jj;
      k = k4 * 4 + i;
    }
    int col;
    if (nl < j.w[0]) col = j.base[0] + rank * j.w[0] + nl;
    else if (nl < j.w[0] + j.w[1]) col = j.base[1] + rank * j.w[1] + (nl - j.w[0]);
    else col = j.base[2] + rank * j.w[2] + (nl - j.w[0] - j.w[1]);
    dst[j.off + e] = j.trans ? j.W[(size_t)k * j.ld + col] : j.W[(size_t)col * j.ld + k];
  }
}

struct SmallJob { const float* src; long long off; int rows, cols, ld, col0; };
struct SmallArgs { int njobs; SmallJob job[4 * kMaxLayers + 4]; };
__global__ void recur_pack_small_kernel(SmallArgs a, float* __restrict__ dst) {
  const SmallJob& j = a.job[blockIdx.x];
  for (int e = threadIdx.x; e < j.rows * j.cols; e += blockDim.x) {
    const int r = e / j.cols, c = e - r * j.cols;
    dst[j.off + e] = j.src[(size_t)r * j.ld + j.col0 + c];
  }
}

// ---- device helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float rem = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(rem));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kConsumers) : "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }

// shared-memory image of one CTA
struct Smem {
  float ring[kRing][kChunkFloats];
  float abuf[kRows * kLda];
  float red[8 * kRows * kMaxNcta];
  float tile[kRows * kMaxNcta];
  float acc0[kRows * kMaxNcta];        // cross-timestep accumulators owned by this CTA (d emb columns)
  float acc1[4 * kMaxNcta];            // (d Wb2 columns)
  float acc2[4 * kMaxNcta];            // (d W0 box part)
  float keep[kRows * kMaxNcta];        // d pred columns handed from layer l+1 to layer l (backward)
  float box[kRows * 4], gb[kRows * 4], tmp4[kRows * 4];
  float cnt[kRows];
  int s_idx[kRows], o_idx[kRows], ind[kRows];
  int nxt_s[kRows], nxt_o[kRows], nxt_i[kRows];   // the next timestep's edges, fetched one timestep ahead
  int lst[kRows][2 * kRows], nlst[kRows];        // live incident edges per node: subjects first, then objects (bit 30)
  int lsta[kRows][2 * kRows], nlsta[kRows];      // all incident edges (transpose of the gather)
  unsigned long long full[kRing], empty[kRing], abar, sbar[2];
};

extern __shared__ __align__(128) unsigned char recur_smem_raw[];
// Every function takes the shared image from the symbol itself, never through a pointer that crossed a call: a
// pointer handed to a non-inlined function is generic, and its loads become LD.E (long scoreboard, ~10x the latency
// of LDS; ncu source view of round 2) instead of LDS.
#define RECUR_SMEM() (reinterpret_cast<Smem*>(recur_smem_raw))

struct Pipe { int slot; uint32_t phase; };        // ring position shared by construction between producer and consumers

struct Ctx {
  Smem* sm;
  const float* model;
  int rank, CS;
  Pipe pipe;
  uint32_t aphase, sphase[2];
  int scount;
  unsigned long long* prof;     // optional timeline of CTA 0 (tools/recur_timeline.py), else null
  int mma;                      // product core: 0 = fp32 FFMA register tiles, 1 = 3xTF32 mma.sync fragments
  int a_pending;                // an operand load is in flight: the stage waits for it after its epilogue prefetch
};

__device__ __forceinline__ void stamp(Ctx& cx, int what) {
  if (cx.prof != nullptr && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    const unsigned long long n = cx.prof[0];
    if (n < 4000) { cx.prof[1 + 2 * n] = (unsigned long long)what; cx.prof[2 + 2 * n] = t; cx.prof[0] = n + 1; }
  }
}

// the producer warp: stream every chunk of every stage of every timestep, in consumer order
__device__ __noinline__ void producer_loop(const float* model, const Table& tab, int rank, int steps, bool once_first) {
  // the time-invariant stage (obj_vecs_net[0] on the embedding) runs once: before the loop in the forward
  // (first entry of the table), after it in the backward (last entry)
  if ((threadIdx.x & 31) != 0) return;
  Smem* sm = RECUR_SMEM();
  int slot = 0; uint32_t phase = 0;
  const int lo = once_first ? 1 : 0, hi = once_first ? tab.n : tab.n - 1;
  for (int t = -1; t <= steps; ++t) {
    int s0, s1;
    if (t < 0) { if (!once_first) continue; s0 = 0; s1 = 1; }
    else if (t == steps) { if (once_first) continue; s0 = tab.n - 1; s1 = tab.n; }
    else { s0 = lo; s1 = hi; }
    for (int s = s0; s < s1; ++s) {
      const StageDev& st = tab.s[s];
      const float* slab = model + st.off + (long long)rank * st.ncta * st.K8 * 8;
      for (int c = 0; c < st.nchunk; ++c) {
        const int kcount = min(st.kc, st.K8 - c * st.kc);
        const uint32_t bytes = (uint32_t)(kcount * st.ncta * 8) * 4u;
        mbar_wait(smem_u32(&sm->empty[slot]), phase ^ 1u);
        mbar_expect_tx(smem_u32(&sm->full[slot]), bytes);
        bulk_g2s(smem_u32(sm->ring[slot]), slab + (long long)c * st.kc * st.ncta * 8, bytes, smem_u32(&sm->full[slot]));
        if (++slot == kRing) { slot = 0; phase ^= 1u; }
      }
    }
  }
}

// A operand: rows [0, M) of a row-major global matrix -> abuf (bulk copies by warp 0, everyone waits)
struct RowSrc { const float* p0; const float* p1; const float* p2; };      // up to three pieces per row
__device__ __forceinline__ void a_begin(Ctx& cx, uint32_t bytes) {
  if (threadIdx.x == 0) mbar_expect_tx(smem_u32(&RECUR_SMEM()->abar), bytes);
  __syncwarp();
}
__device__ __forceinline__ void a_wait(Ctx& cx) {
  mbar_wait(smem_u32(&RECUR_SMEM()->abar), cx.aphase);
  cx.aphase ^= 1u;
}
__device__ __noinline__ void load_rows(Ctx& cx, const float* src, int ld, int M, int K) {
  // issue only: the consumer of the operand (run_stage, or a caller that edits it in place) calls a_finish()
  stamp(cx, 1);
  if (threadIdx.x < 32) {
    a_begin(cx, (uint32_t)(M * K) * 4u);
    for (int m = threadIdx.x; m < M; m += 32)
      bulk_g2s(smem_u32(RECUR_SMEM()->abuf + m * kLda), src + (size_t)m * ld, (uint32_t)K * 4u, smem_u32(&RECUR_SMEM()->abar));
  }
  cx.a_pending = 1;
}
__device__ __forceinline__ void a_finish(Ctx& cx) {
  if (cx.a_pending) { a_wait(cx); cx.a_pending = 0; }
}

// Epilogue of a stage, applied to every element of the CTA's [16 x ncta] tile after the 8 k-slices are summed:
//   v = sum (+ bias[n]) (+ sum_j box[m][j] * w4[n][j]) (+ ext[m*ld+n]);  relu;  gate[m*ld+n] > 0 ? v : 0;
//   rows m < rows are stored to out / out2 (global, leading dimension ld); the tile keeps v for CTA-local post phases.
struct Epi {
  const float* bias; const float* w4; const float* ext; const float* gate;
  float* out; float* out2;
  int ld, rows, relu;
};
__device__ __forceinline__ Epi epi_make(const float* bias, const float* gate, float* out, int ld, int rows, int relu) {
  Epi e; e.bias = bias; e.w4 = nullptr; e.ext = nullptr; e.gate = gate; e.out = out; e.out2 = nullptr; e.ld = ld; e.rows = rows; e.relu = relu;
  return e;
}

// One GEMM stage: tile[m][nl] = epilogue(sum_k abuf[m][k] * B(k, col(nl))).
// fp32 FFMA register tiles (exact fp32 products): a warp owns one k-slice (k4 = warp, warp + 8, ...), a lane owns
// rows {mg, mg+4, mg+8, mg+12} x columns {ng + 8j}; operands come as LDS.128 along k (A: 4 rows, conflict-free with
// the row stride kLda = 8 mod 32; B: packed, 128 B per (k4, j)).  The function exists ONCE per column-tile count
// (not once per call site): the stages of a timestep run one after the other, each only a few iterations long, and
// with an inlined copy per stage the kernel was instruction-fetch bound (tools/recur_timeline.py).  The epilogue's
// global operands (bias, gate, ...) are fetched BEFORE the product so their L2 latency hides behind it.
template <int TN>
__device__ __noinline__ void stage_tn(Ctx& cx, const StageDev& st, const Epi& ep) {
  Smem* sm = RECUR_SMEM();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mg = lane >> 3, ng = lane & 7;
  constexpr int kElems = (kRows * TN * 8 + kConsumers - 1) / kConsumers;      // tile elements per thread in the epilogue
  const int ncta = TN * 8, total = kRows * ncta;
  float pb[kElems], pe[kElems], pg[kElems];
  int pidx[kElems];
#pragma unroll
  for (int q = 0; q < kElems; ++q) {
    const int i = threadIdx.x + q * kConsumers;
    pb[q] = 0.f; pe[q] = 0.f; pg[q] = 1.f; pidx[q] = -1;
    if (i < total) {
      const int m = i / ncta, nl = i - m * ncta;
      const int n = col_of(st, cx.rank, nl);
      if (ep.bias) pb[q] = ep.bias[n];
      if (ep.w4) {
        const float4 w = *reinterpret_cast<const float4*>(ep.w4 + (size_t)n * 4);
        const float* bx = sm->box + m * 4;
        float v = pb[q];
        v += bx[0] * w.x; v += bx[1] * w.y; v += bx[2] * w.z; v += bx[3] * w.w;
        pb[q] = v;
      }
      if (m < ep.rows) {
        pidx[q] = m * ep.ld + n;
        if (ep.ext) pe[q] = ep.ext[pidx[q]];
        if (ep.gate) pg[q] = ep.gate[pidx[q]];
      }
    }
  }
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  a_finish(cx);
  stamp(cx, 2);
  const float* arow = sm->abuf + mg * kLda;
  for (int c = 0; c < st.nchunk; ++c) {
    const int kcount = min(st.kc, st.K8 - c * st.kc);
    const int slot = cx.pipe.slot;
    mbar_wait(smem_u32(&sm->full[slot]), cx.pipe.phase);
    const float* ring = sm->ring[slot];
    const int nk4 = kcount * 2;
#pragma unroll 2
    for (int q = warp; q < nk4; q += 8) {
      const int k = (c * st.kc * 2 + q) << 2;
      float4 av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const float4*>(arow + (4 * i) * kLda + k);
      const float* bp = ring + ((size_t)q * TN * 8 + ng) * 4;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(bp + j * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i][j] = fmaf(av[i].x, b.x, acc[i][j]);
          acc[i][j] = fmaf(av[i].y, b.y, acc[i][j]);
          acc[i][j] = fmaf(av[i].z, b.z, acc[i][j]);
          acc[i][j] = fmaf(av[i].w, b.w, acc[i][j]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&sm->empty[slot]));
    if (++cx.pipe.slot == kRing) { cx.pipe.slot = 0; cx.pipe.phase ^= 1u; }
  }
  stamp(cx, 3);
  float* red = sm->red + (size_t)warp * total;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
#pragma unroll
    for (int i = 0; i < 4; ++i) red[(4 * i + mg) * ncta + ng + 8 * j] = acc[i][j];
  }
  consumer_sync();
#pragma unroll
  for (int q = 0; q < kElems; ++q) {
    const int i = threadIdx.x + q * kConsumers;
    if (i < total) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += sm->red[(size_t)w * total + i];
      v += pb[q];
      v += pe[q];
      if (ep.relu) v = fmaxf(v, 0.f);
      v = pg[q] > 0.f ? v : 0.f;
      if (pidx[q] >= 0) {
        if (ep.out) ep.out[pidx[q]] = v;
        if (ep.out2) ep.out2[pidx[q]] = v;
      }
      sm->tile[i] = v;
    }
  }
  consumer_sync();
}

// The same stage on legacy tensor cores: 3xTF32 mma.sync m16n8k8 (hi/lo split, fp32-class accuracy).  A fragment
// load feeds 8x more multiply-adds than an FFMA operand load, so this core is not bound by the shared-memory
// pipe; the two small cross terms and the main term go to separate accumulators (no chain of three dependent MMAs).
template <int TN>
__device__ __noinline__ void stage_tn_mma(Ctx& cx, const StageDev& st, const Epi& ep) {
  Smem* sm = RECUR_SMEM();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int kElems = (kRows * TN * 8 + kConsumers - 1) / kConsumers;
  const int ncta = TN * 8, total = kRows * ncta;
  float pb[kElems], pe[kElems], pg[kElems];
  int pidx[kElems];
#pragma unroll
  for (int q = 0; q < kElems; ++q) {
    const int i = threadIdx.x + q * kConsumers;
    pb[q] = 0.f; pe[q] = 0.f; pg[q] = 1.f; pidx[q] = -1;
    if (i < total) {
      const int m = i / ncta, nl = i - m * ncta;
      const int n = col_of(st, cx.rank, nl);
      if (ep.bias) pb[q] = ep.bias[n];
      if (m < ep.rows) {
        pidx[q] = m * ep.ld + n;
        if (ep.ext) pe[q] = ep.ext[pidx[q]];
        if (ep.gate) pg[q] = ep.gate[pidx[q]];
      }
    }
  }
  float acs[TN][4], ach[TN][4];
#pragma unroll
  for (int j = 0; j < TN; ++j) { acs[j][0] = acs[j][1] = acs[j][2] = acs[j][3] = 0.f; ach[j][0] = ach[j][1] = ach[j][2] = ach[j][3] = 0.f; }
  a_finish(cx);
  stamp(cx, 2);
  for (int c = 0; c < st.nchunk; ++c) {
    const int kcount = min(st.kc, st.K8 - c * st.kc);
    const int slot = cx.pipe.slot;
    mbar_wait(smem_u32(&sm->full[slot]), cx.pipe.phase);
    const float* ring = sm->ring[slot];
#pragma unroll 2
    for (int kl = warp; kl < kcount; kl += 8) {
      const int k0 = (c * st.kc + kl) << 3;
      const float2 lo = *reinterpret_cast<const float2*>(sm->abuf + g * kLda + k0 + 2 * t);
      const float2 hi = *reinterpret_cast<const float2*>(sm->abuf + (g + 8) * kLda + k0 + 2 * t);
      uint32_t ah[4], al[4];
      split_tf32(lo.x, ah[0], al[0]); split_tf32(hi.x, ah[1], al[1]);
      split_tf32(lo.y, ah[2], al[2]); split_tf32(hi.y, ah[3], al[3]);
      const float* bp = ring + ((size_t)kl * TN * 32 + lane) * 2;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const float2 b = *reinterpret_cast<const float2*>(bp + j * 64);
        uint32_t bh[2], bl[2];
        split_tf32(b.x, bh[0], bl[0]); split_tf32(b.y, bh[1], bl[1]);
        mma_tf32(acs[j], al, bh);
        mma_tf32(ach[j], ah, bh);
        mma_tf32(acs[j], ah, bl);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&sm->empty[slot]));
    if (++cx.pipe.slot == kRing) { cx.pipe.slot = 0; cx.pipe.phase ^= 1u; }
  }
  stamp(cx, 3);
  float* red = sm->red + (size_t)warp * total;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    *reinterpret_cast<float2*>(red + g * ncta + j * 8 + 2 * t) = make_float2(acs[j][0] + ach[j][0], acs[j][1] + ach[j][1]);
    *reinterpret_cast<float2*>(red + (g + 8) * ncta + j * 8 + 2 * t) = make_float2(acs[j][2] + ach[j][2], acs[j][3] + ach[j][3]);
  }
  consumer_sync();
#pragma unroll
  for (int q = 0; q < kElems; ++q) {
    const int i = threadIdx.x + q * kConsumers;
    if (i < total) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += sm->red[(size_t)w * total + i];
      v += pb[q];
      v += pe[q];
      if (ep.relu) v = fmaxf(v, 0.f);
      v = pg[q] > 0.f ? v : 0.f;
      if (pidx[q] >= 0) {
        if (ep.out) ep.out[pidx[q]] = v;
        if (ep.out2) ep.out2[pidx[q]] = v;
      }
      sm->tile[i] = v;
    }
  }
  consumer_sync();
}

__device__ __noinline__ void run_stage(Ctx& cx, const StageDev& st, const Epi& ep) {
  if (cx.mma) {
    switch (st.ncta >> 3) {
      case 1: stage_tn_mma<1>(cx, st, ep); break;
      case 2: stage_tn_mma<2>(cx, st, ep); break;
      case 3: stage_tn_mma<3>(cx, st, ep); break;
      case 4: stage_tn_mma<4>(cx, st, ep); break;
      case 5: stage_tn_mma<5>(cx, st, ep); break;
      case 6: stage_tn_mma<6>(cx, st, ep); break;
      case 7: stage_tn_mma<7>(cx, st, ep); break;
      case 8: stage_tn_mma<8>(cx, st, ep); break;
      default: stage_tn_mma<9>(cx, st, ep); break;
    }
    return;
  }
  switch (st.ncta >> 3) {
    case 1: stage_tn<1>(cx, st, ep); break;
    case 2: stage_tn<2>(cx, st, ep); break;
    case 3: stage_tn<3>(cx, st, ep); break;
    case 4: stage_tn<4>(cx, st, ep); break;
    case 5: stage_tn<5>(cx, st, ep); break;
    case 6: stage_tn<6>(cx, st, ep); break;
    case 7: stage_tn<7>(cx, st, ep); break;
    case 8: stage_tn<8>(cx, st, ep); break;
    default: stage_tn<9>(cx, st, ep); break;
  }
}

// end of a stage: everything this CTA wrote to global memory becomes visible to the cluster, and vice versa
__device__ __noinline__ void stage_sync(Ctx& cx) {
  // release: every consumer's global writes happen before the CTA barrier, then CS threads arrive (release.cluster) on
  // the peers' barriers.  acquire: only warp 0 waits at cluster scope - it alone issues the bulk copies that read what
  // the peers wrote (everything else a CTA reads from peers goes through ld.global.cg) - and the second CTA barrier
  // releases the other warps.  An acquire.cluster by all 256 threads costs a CCTL.IVALL each (6.5% of all stall
  // samples in the first ncu capture).
  Smem* sm = RECUR_SMEM();
  stamp(cx, 4);
  consumer_sync();
  const int which = cx.scount & 1;
  const uint32_t bar = smem_u32(&sm->sbar[which]);
  if (threadIdx.x < cx.CS) mbar_arrive_remote(bar, threadIdx.x);
  if (threadIdx.x < 32) {
    mbar_wait_cluster(bar, cx.sphase[which]);
    fence_proxy_async();
  }
  cx.sphase[which] ^= 1u;
  ++cx.scount;
  consumer_sync();
  stamp(cx, 5);
}

__device__ void init_cta(int CS) {
  Smem* sm = RECUR_SMEM();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(smem_u32(&sm->full[i]), 1); mbar_init(smem_u32(&sm->empty[i]), 8); }
    mbar_init(smem_u32(&sm->abar), 1);
    mbar_init(smem_u32(&sm->sbar[0]), CS);
    mbar_init(smem_u32(&sm->sbar[1]), CS);
    fence_barrier_init();
  }
  __syncthreads();
  cg::this_cluster().sync();          // every CTA's barriers exist before anyone arrives remotely
}

// indices, live / all incident-edge lists and counts of one timestep (graph.py:79-100).  `cur` = this timestep's
// edges (used when nothing was prefetched), `nxt` = the timestep the chain visits next (or null): its edges are
// fetched by the last consumer warp while the stages of this timestep run.
__device__ __noinline__ void load_graph(const long long* cur_e, const unsigned char* cur_i, bool have,
                                        const long long* nxt_e, const unsigned char* nxt_i, int O, int E) {
  Smem* sm = RECUR_SMEM();
  if (threadIdx.x < kRows) {
    const int e = threadIdx.x;
    int s = 0, o = 0, live = 0;
    if (have) { s = sm->nxt_s[e]; o = sm->nxt_o[e]; live = sm->nxt_i[e]; }
    else if (e < E) {
      const long long a = cur_e[2 * e], b = cur_e[2 * e + 1];
      s = a < 0 ? 0 : (a >= O ? O - 1 : (int)a);
      o = b < 0 ? 0 : (b >= O ? O - 1 : (int)b);
      live = cur_i[e] != 0;
    }
    sm->s_idx[e] = s; sm->o_idx[e] = o; sm->ind[e] = live;
  }
  consumer_sync();
  if (threadIdx.x < kRows) {
    const int node = threadIdx.x;
    int n = 0, na = 0;
    for (int pass = 0; pass < 2; ++pass) {
      const int* idx = pass ? sm->o_idx : sm->s_idx;
      for (int e = 0; e < E; ++e) {
        if (idx[e] == node) {
          const int ent = pass ? (e | (1 << 30)) : e;
          sm->lsta[node][na++] = ent;
          if (sm->ind[e]) sm->lst[node][n++] = ent;
        }
      }
    }
    sm->nlst[node] = n; sm->nlsta[node] = na;
    sm->cnt[node] = (float)n;
  } else if (nxt_e != nullptr && threadIdx.x >= kConsumers - 32 && threadIdx.x < kConsumers - 32 + kRows) {
    const int e = threadIdx.x - (kConsumers - 32);
    int s = 0, o = 0, live = 0;
    if (e < E) {
      const long long a = nxt_e[2 * e], b = nxt_e[2 * e + 1];
      s = a < 0 ? 0 : (a >= O ? O - 1 : (int)a);
      o = b < 0 ? 0 : (b >= O ? O - 1 : (int)b);
      live = nxt_i[e] != 0;
    }
    sm->nxt_s[e] = s; sm->nxt_o[e] = o; sm->nxt_i[e] = live;
  }
  consumer_sync();
}

// ===================================================================================================================
// forward
// ===================================================================================================================
struct FwdArgs {
  Dims d;
  Table tab;
  Small small;
  int NC, chains_per_model;
  const float* model[kMaxModels];
  const float *emb, *box0, *pred; const long long* edges; const unsigned char* ind;
  float *objv, *boxes, *saved;
  unsigned long long* prof;
  int mma;
};

// per-CTA partial sums of a 4-wide product, exchanged through global memory and added in rank order
__device__ __forceinline__ float sum_parts(const float* part, int CS, int i) {
  float v = 0.f;
  for (int r = 0; r < CS; ++r) v += __ldcg(part + (size_t)r * (kRows * 4) + i);
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) recur_fwd_kernel(const __grid_constant__ FwdArgs a) {
  Smem* sm = reinterpret_cast<Smem*>(recur_smem_raw);
  const Dims& d = a.d;
  const int CS = d.CS, rank = (int)cluster_rank(), chain = blockIdx.x / CS;
  const int clip = chain % a.chains_per_model;        // box0 / edges / ind are data: shared by the models of a launch
  const float* model = a.model[chain / a.chains_per_model];
  init_cta(CS);
  const int steps = d.T - 1;
  if (threadIdx.x >= kConsumers) { producer_loop(model, a.tab, rank, steps, true); return; }

  Ctx cx{sm, model, rank, CS, {0, 0u}, 0u, {0u, 0u}, 0, (blockIdx.x == 0) ? a.prof : nullptr, a.mma, 0};
  const Saved sv = saved_layout(d, a.NC);
  const int O = d.O, E = d.E, H = d.H, De = d.De, N2 = d.N2(), Dout = d.Dout;
  for (int i = threadIdx.x; i < O * 4; i += kConsumers) sm->box[i] = a.box0[(size_t)clip * O * 4 + i];
  // frame 0: boxes[0] = box0, objv[0] = 0 (model.py:124-125)
  if (rank == 0) {
    for (int i = threadIdx.x; i < O * 4; i += kConsumers) a.boxes[((size_t)chain * d.T) * O * 4 + i] = a.box0[(size_t)clip * O * 4 + i];
    for (int i = threadIdx.x; i < O * Dout; i += kConsumers) a.objv[((size_t)chain * d.T) * O * Dout + i] = 0.f;
  }
  consumer_sync();
  const long long* edges0 = a.edges + (size_t)clip * d.T * E * 2;
  const unsigned char* ind0 = a.ind + (size_t)clip * d.T * E;
  const int wDe = De / CS;
  float* p0 = a.saved + sv.p0 + (size_t)chain * O * De;
  {  // obj_vecs_net[0] = [emb | box] W0^T: the embedding part does not depend on t - ONE product per chain
    load_rows(cx, a.emb + (size_t)chain * O * d.Kx, d.Kx, O, d.Kx);
    run_stage(cx, a.tab.s[0], epi_make(nullptr, nullptr, p0, De, O, 0));
    stage_sync(cx);
  }

  for (int t = 1; t < d.T; ++t) {
    const long long ct = (long long)chain * steps + (t - 1);
    load_graph(edges0 + (size_t)t * E * 2, ind0 + (size_t)t * E, t > 1,
               t + 1 < d.T ? edges0 + (size_t)(t + 1) * E * 2 : nullptr, ind0 + (size_t)(t + 1 < d.T ? t + 1 : t) * E, O, E);
    if (rank == 0 && threadIdx.x < kRows) a.saved[sv.cnt + ct * 16 + threadIdx.x] = sm->cnt[threadIdx.x];
    int s = 1;
    {  // u0 = relu(P + box W0[:, Kx:]^T) (model.py:136-137): 4 FMAs per element, every CTA builds the whole operand
       // for itself and keeps its own columns for the backward
      load_rows(cx, p0, De, O, De);
      a_finish(cx);
      const float* wbox = model + a.small.w0box;
      float* u0 = a.saved + sv.u0 + ct * O * De;
      for (int i = threadIdx.x; i < O * De; i += kConsumers) {
        const int m = i / De, k = i - m * De;
        const float4 w = *reinterpret_cast<const float4*>(wbox + (size_t)k * 4);
        const float* bx = sm->box + m * 4;
        float v = sm->abuf[m * kLda + k];
        v += bx[0] * w.x; v += bx[1] * w.y; v += bx[2] * w.z; v += bx[3] * w.w;
        v = fmaxf(v, 0.f);
        sm->abuf[m * kLda + k] = v;
        if (k / wDe == rank) u0[i] = v;
      }
      consumer_sync();
    }
    {  // obj_vecs_net[2]
      const StageDev& st = a.tab.s[s++];
      run_stage(cx, st, epi_make(nullptr, nullptr, a.saved + sv.x0 + ct * O * De, De, O, 1));
      stage_sync(cx);
    }
    const float* obj_in = a.saved + sv.x0 + ct * O * De;
    const float* pred_in = a.pred + ((size_t)chain * d.T + t) * E * d.Dp;
    int pred_ld = d.Dp;
    for (int l = 0; l < d.NL; ++l) {
      const int Din = d.Din(l), Dpl = d.Dpl(l), K1 = d.K1(l);
      {  // net1[0] on the gathered triples [obj[s] | pred | obj[o]] (graph.py:67-71)
        const StageDev& st = a.tab.s[s++];
        stamp(cx, 1);
        if (threadIdx.x < 32) {
          a_begin(cx, (uint32_t)(E * K1) * 4u);
          for (int e = threadIdx.x; e < E; e += 32) {
            const uint32_t dst = smem_u32(sm->abuf + e * kLda), bar = smem_u32(&sm->abar);
            bulk_g2s(dst, obj_in + (size_t)sm->s_idx[e] * Din, (uint32_t)Din * 4u, bar);
            bulk_g2s(dst + (uint32_t)Din * 4u, pred_in + (size_t)e * pred_ld, (uint32_t)Dpl * 4u, bar);
            bulk_g2s(dst + (uint32_t)(Din + Dpl) * 4u, obj_in + (size_t)sm->o_idx[e] * Din, (uint32_t)Din * 4u, bar);
          }
        }
        a_wait(cx);
        const bool keeper = rank == (l % CS) && threadIdx.x < 32;
        if (keeper) {                                      // keep the gathered rows: X operand of dW1a
          float* trow = a.saved + sv.trow[l] + ct * E * K1;
          for (int e = threadIdx.x; e < E; e += 32) bulk_s2g(trow + (size_t)e * K1, smem_u32(sm->abuf + e * kLda), (uint32_t)K1 * 4u);
          bulk_commit();
        }
        run_stage(cx, st, epi_make(model + a.small.bias[l][0], nullptr, a.saved + sv.h1[l] + ct * E * H, H, E, 1));
        if (keeper) bulk_wait_read();
        stage_sync(cx);
      }
      {  // net1[2]; the CTA owns matching subject / predicate / object columns, so it pools its own columns
        const StageDev& st = a.tab.s[s++];
        load_rows(cx, a.saved + sv.h1[l] + ct * E * H, H, E, H);
        run_stage(cx, st, epi_make(model + a.small.bias[l][1], nullptr, a.saved + sv.h2[l] + ct * E * N2, N2, E, 1));
        const int ws = st.w[0], wp = st.w[1];
        float* pooled = a.saved + sv.pooled[l] + ct * O * H;
        for (int i = threadIdx.x; i < O * ws; i += kConsumers) {      // graph.py:89-99, subjects then objects, edge order
          const int node = i / ws, c = i - node * ws;
          const int n = sm->nlst[node];
          float acc = 0.f;
          for (int q = 0; q < n; ++q) {
            const int ent = sm->lst[node][q], e = ent & 0xffff;
            acc += sm->tile[e * st.ncta + ((ent >> 30) ? ws + wp + c : c)];
          }
          if (n > 0) acc /= (float)n;
          pooled[(size_t)node * H + rank * ws + c] = acc;
        }
        stage_sync(cx);
      }
      {  // net2[0]
        const StageDev& st = a.tab.s[s++];
        load_rows(cx, a.saved + sv.pooled[l] + ct * O * H, H, O, H);
        run_stage(cx, st, epi_make(model + a.small.bias[l][2], nullptr, a.saved + sv.g1[l] + ct * O * H, H, O, 1));
        stage_sync(cx);
      }
      {  // net2[2]: the layer's new object vectors; the last layer's are the frame's output (model.py:166)
        const StageDev& st = a.tab.s[s++];
        load_rows(cx, a.saved + sv.g1[l] + ct * O * H, H, O, H);
        Epi ep = epi_make(model + a.small.bias[l][3], nullptr, a.saved + sv.nobj[l] + ct * O * Dout, Dout, O, 1);
        if (l == d.NL - 1) ep.out2 = a.objv + ((size_t)chain * d.T + t) * O * Dout;
        run_stage(cx, st, ep);
        stage_sync(cx);
      }
      obj_in = a.saved + sv.nobj[l] + ct * O * Dout;
      pred_in = a.saved + sv.h2[l] + ct * E * N2 + H;          // new_p = the middle slice of h2 (graph.py:74)
      pred_ld = N2;
    }
    {  // box_net (model.py:168): the hidden columns owned here give a partial of the 4 box outputs; the partials of
       // the cluster are exchanged through L2 and added in rank order by every CTA for itself
      const StageDev& st = a.tab.s[s++];
      load_rows(cx, obj_in, Dout, O, Dout);
      run_stage(cx, st, epi_make(model + a.small.bb1, nullptr, a.saved + sv.hb + ct * O * H, H, O, 1));
      const float* wb2 = model + a.small.wb2;
      float* part = a.saved + sv.part + ((size_t)chain * 2 + (t & 1)) * 16 * (kRows * 4);
      for (int i = threadIdx.x; i < O * 4; i += kConsumers) {
        const int m = i >> 2, j = i & 3;
        float acc = 0.f;
        for (int c = 0; c < st.ncta; ++c) acc = fmaf(sm->tile[m * st.ncta + c], wb2[(size_t)j * H + rank * st.ncta + c], acc);
        part[(size_t)rank * (kRows * 4) + i] = acc;
      }
      stage_sync(cx);
      const float* bb2 = model + a.small.bb2;
      for (int i = threadIdx.x; i < O * 4; i += kConsumers) {
        const float v = sm->box[i] + (sum_parts(part, CS, i) + bb2[i & 3]);
        sm->box[i] = v;
        if (rank == 0) a.boxes[((size_t)chain * d.T + t) * O * 4 + i] = v;
      }
      consumer_sync();
    }
  }
}

// ===================================================================================================================
// backward: the reverse chain for the data gradients; every Linear's gated output gradient goes to the Z buffer
// ===================================================================================================================
struct BwdArgs {
  Dims d;
  Table tab;
  Small small;
  int NC, chain0;                             // chains of the forward launch; first chain of this model
  const float* model;
  const float *saved, *boxes;                 // forward activations (all NC chains); this model's boxes [nc][T][O][4]
  const float *d_objv, *d_boxes;              // incoming gradients [nc][T][O][Dout] (may be null), [nc][T][O][4] (may be null)
  const long long* edges; const unsigned char* ind;
  float* z;                                   // Z buffer (z_layout)
  float *d_emb, *d_box0, *d_pred;             // [nc][O][Kx], [nc][O][4], [nc][T][E][Dp]
  float *dw0box, *dwb2, *dbb2;                // per-chain partial sums [nc][De][4], [nc][4][H], [nc][4]
  int mma;
};

__global__ void __launch_bounds__(kThreads, 1) recur_bwd_kernel(const __grid_constant__ BwdArgs a) {
  Smem* sm = reinterpret_cast<Smem*>(recur_smem_raw);
  const Dims& d = a.d;
  const int CS = d.CS, rank = (int)cluster_rank(), chain = blockIdx.x / CS;      // model-local chain (clip) index
  const float* model = a.model;
  init_cta(CS);
  const int steps = d.T - 1;
  if (threadIdx.x >= kConsumers) { producer_loop(model, a.tab, rank, steps, false); return; }

  Ctx cx{sm, model, rank, CS, {0, 0u}, 0u, {0u, 0u}, 0, nullptr, a.mma, 0};
  const Saved sv = saved_layout(d, a.NC);
  const Zbuf zz = z_layout(d, a.NC);
  const int O = d.O, E = d.E, H = d.H, De = d.De, N2 = d.N2(), Dout = d.Dout, Dpo = d.Dpo;
  const int wH = H / CS, wDe = De / CS;
  for (int i = threadIdx.x; i < kRows * kMaxNcta; i += kConsumers) sm->acc0[i] = 0.f;
  for (int i = threadIdx.x; i < 4 * kMaxNcta; i += kConsumers) { sm->acc1[i] = 0.f; sm->acc2[i] = 0.f; }
  for (int i = threadIdx.x; i < kRows * 4; i += kConsumers) sm->box[i] = 0.f;           // carry: d loss / d boxes[t] from step t+1
  float dbb2_acc = 0.f;                                                                  // thread j < 4 of every CTA
  // frame 0 receives no predicate gradient (the recurrence starts at t = 1)
  if (rank == 0) for (int i = threadIdx.x; i < E * d.Dp; i += kConsumers) a.d_pred[((size_t)chain * d.T) * E * d.Dp + i] = 0.f;
  consumer_sync();
  const long long* edges0 = a.edges + (size_t)chain * d.T * E * 2;
  const unsigned char* ind0 = a.ind + (size_t)chain * d.T * E;

  for (int t = d.T - 1; t >= 1; --t) {
    const long long ct = (long long)(a.chain0 + chain) * steps + (t - 1);      // row block inside the saved / Z buffers
    load_graph(edges0 + (size_t)t * E * 2, ind0 + (size_t)t * E, t < d.T - 1,
               t > 1 ? edges0 + (size_t)(t - 1) * E * 2 : nullptr, ind0 + (size_t)(t > 1 ? t - 1 : t) * E, O, E);
    for (int i = threadIdx.x; i < O * 4; i += kConsumers)
      sm->gb[i] = sm->box[i] + (a.d_boxes ? a.d_boxes[((size_t)chain * d.T + t) * O * 4 + i] : 0.f);
    int s = 0;
    {  // box_net[2]^T for every CTA: zb1 = (gB Wb2) * [hb > 0], built in place of hb; d Wb2 columns owned by this CTA
      load_rows(cx, a.saved + sv.hb + ct * O * H, H, O, H);
      a_finish(cx);
      const float* wb2 = model + a.small.wb2;
      for (int i = threadIdx.x; i < 4 * wH; i += kConsumers) {      // acc1[j][c] += sum_m gB[m][j] * hb[m][rank*wH + c]
        const int j = i / wH, c = i - j * wH;
        float acc = sm->acc1[i];
        for (int m = 0; m < O; ++m) acc = fmaf(sm->gb[m * 4 + j], sm->abuf[m * kLda + rank * wH + c], acc);
        sm->acc1[i] = acc;
      }
      if (threadIdx.x < 4) { float acc = dbb2_acc; for (int m = 0; m < O; ++m) acc += sm->gb[m * 4 + threadIdx.x]; dbb2_acc = acc; }
      consumer_sync();
      float* zb1 = a.z + zz.zb1 + ct * O * H;
      for (int i = threadIdx.x; i < O * H; i += kConsumers) {
        const int m = i / H, k = i - m * H;
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) v = fmaf(sm->gb[m * 4 + j], wb2[(size_t)j * H + k], v);
        v = sm->abuf[m * kLda + k] > 0.f ? v : 0.f;
        sm->abuf[m * kLda + k] = v;
        if (rank == 0) zb1[i] = v;
      }
      consumer_sync();
    }
    {  // box_net[0]^T + the frame's own output gradient, gated by the last layer's ReLU -> z4[NL-1]
      const StageDev& st = a.tab.s[s++];
      Epi ep = epi_make(nullptr, a.saved + sv.nobj[d.NL - 1] + ct * O * Dout, a.z + zz.z4[d.NL - 1] + ct * O * Dout, Dout, O, 0);
      ep.ext = a.d_objv ? a.d_objv + ((size_t)chain * d.T + t) * O * Dout : nullptr;
      run_stage(cx, st, ep);
      stage_sync(cx);
    }
    for (int l = d.NL - 1; l >= 0; --l) {
      const int Din = d.Din(l), Dpl = d.Dpl(l);
      {  // net2[2]^T
        const StageDev& st = a.tab.s[s++];
        load_rows(cx, a.z + zz.z4[l] + ct * O * Dout, Dout, O, Dout);
        run_stage(cx, st, epi_make(nullptr, a.saved + sv.g1[l] + ct * O * H, a.z + zz.z3[l] + ct * O * H, H, O, 0));
        stage_sync(cx);
      }
      {  // net2[0]^T -> d pooled (columns owned here); transpose of the pooling + ReLU gate of h2 -> z2 slices
        const StageDev& st = a.tab.s[s++];
        load_rows(cx, a.z + zz.z3[l] + ct * O * H, H, O, H);
        run_stage(cx, st, epi_make(nullptr, nullptr, nullptr, H, O, 0));
        const int ws = st.ncta, wp = Dpo / CS;
        const float* h2 = a.saved + sv.h2[l] + ct * E * N2;
        float* z2 = a.z + zz.z2[l] + ct * E * N2;
        for (int i = threadIdx.x; i < E * (2 * ws + wp); i += kConsumers) {
          const int e = i / (2 * ws + wp), q = i - e * (2 * ws + wp);
          int n; float v;
          if (q < ws) {
            const int node = sm->s_idx[e];
            n = rank * ws + q;
            v = sm->ind[e] ? sm->tile[node * st.ncta + q] / fmaxf(sm->cnt[node], 1.f) : 0.f;
          } else if (q < ws + wp) {
            n = H + rank * wp + (q - ws);
            v = (l == d.NL - 1) ? 0.f : sm->keep[e * kMaxNcta + (q - ws)];
          } else {
            const int node = sm->o_idx[e], c = q - ws - wp;
            n = H + Dpo + rank * ws + c;
            v = sm->ind[e] ? sm->tile[node * st.ncta + c] / fmaxf(sm->cnt[node], 1.f) : 0.f;
          }
          const size_t idx = (size_t)e * N2 + n;
          z2[idx] = h2[idx] > 0.f ? v : 0.f;
        }
        stage_sync(cx);
      }
      {  // net1[2]^T
        const StageDev& st = a.tab.s[s++];
        load_rows(cx, a.z + zz.z2[l] + ct * E * N2, N2, E, N2);
        run_stage(cx, st, epi_make(nullptr, a.saved + sv.h1[l] + ct * E * H, a.z + zz.z1[l] + ct * E * H, H, E, 0));
        stage_sync(cx);
      }
      {  // net1[0]^T -> d triples (matching subject / predicate / object columns owned here); transpose of the gather
        const StageDev& st = a.tab.s[s++];
        load_rows(cx, a.z + zz.z1[l] + ct * E * H, H, E, H);
        run_stage(cx, st, epi_make(nullptr, nullptr, nullptr, H, E, 0));
        const int wd = st.w[0], wpd = st.w[1];
        const float* gate = (l > 0) ? a.saved + sv.nobj[l - 1] + ct * O * Dout : a.saved + sv.x0 + ct * O * De;
        float* out = (l > 0) ? a.z + zz.z4[l - 1] + ct * O * Dout : a.z + zz.zx0 + ct * O * De;
        for (int i = threadIdx.x; i < O * wd; i += kConsumers) {
          const int node = i / wd, c = i - node * wd;
          const int n = sm->nlsta[node];
          float acc = 0.f;
          for (int q = 0; q < n; ++q) {
            const int ent = sm->lsta[node][q], e = ent & 0xffff;
            acc += sm->tile[e * st.ncta + ((ent >> 30) ? wd + wpd + c : c)];
          }
          const size_t idx = (size_t)node * Din + rank * wd + c;
          out[idx] = gate[idx] > 0.f ? acc : 0.f;
        }
        float* dp = a.d_pred + ((size_t)chain * d.T + t) * E * d.Dp;
        for (int i = threadIdx.x; i < E * wpd; i += kConsumers) {
          const int e = i / wpd, q = i - e * wpd;
          const float v = sm->tile[e * st.ncta + wd + q];
          if (l > 0) sm->keep[e * kMaxNcta + q] = v;
          else dp[(size_t)e * Dpl + rank * wpd + q] = v;
        }
        stage_sync(cx);
      }
    }
    {  // obj_vecs_net[2]^T -> zu0 columns owned here; obj_vecs_net[0]^T is linear in zu0, so its embedding part waits
       // for the sum over t (one product per chain, below) and only the 4 box columns are needed now: per-CTA
       // partials over the owned columns, exchanged through L2 and added in rank order -> the carry into step t-1
      const StageDev& st = a.tab.s[s++];
      load_rows(cx, a.z + zz.zx0 + ct * O * De, De, O, De);
      run_stage(cx, st, epi_make(nullptr, a.saved + sv.u0 + ct * O * De, a.z + zz.zu0 + ct * O * De, De, O, 0));
      for (int i = threadIdx.x; i < O * st.ncta; i += kConsumers) sm->acc0[i] += sm->tile[i];
      const float* wbox = model + a.small.w0box;
      float* part = a.z + zz.part + ((size_t)(a.chain0 + chain) * 2 + (t & 1)) * 16 * (kRows * 4);
      for (int i = threadIdx.x; i < O * 4; i += kConsumers) {
        const int m = i >> 2, j = i & 3;
        float acc = 0.f;
        for (int c = 0; c < st.ncta; ++c) acc = fmaf(sm->tile[m * st.ncta + c], wbox[(size_t)(rank * st.ncta + c) * 4 + j], acc);
        part[(size_t)rank * (kRows * 4) + i] = acc;
      }
      // d W0[:, Kx:Kx+4] rows owned here: acc2[c][j] += sum_m zu0[m][rank*wDe + c] * boxes[t-1][m][j]
      const float* bprev = a.boxes + ((size_t)chain * d.T + (t - 1)) * O * 4;
      for (int i = threadIdx.x; i < 4 * wDe; i += kConsumers) {
        const int c = i >> 2, j = i & 3;
        float acc = sm->acc2[i];
        for (int m = 0; m < O; ++m) acc = fmaf(sm->tile[m * st.ncta + c], bprev[m * 4 + j], acc);
        sm->acc2[i] = acc;
      }
      stage_sync(cx);
      for (int i = threadIdx.x; i < O * 4; i += kConsumers) sm->box[i] = sm->gb[i] + sum_parts(part, CS, i);
      consumer_sync();
    }
  }
  {  // d emb = (sum_t zu0) W0[:, :Kx]: the owned columns of the sum go to L2, then one product per chain
    float* ssum = a.z + zz.ssum + (size_t)(a.chain0 + chain) * O * De;
    for (int i = threadIdx.x; i < O * wDe; i += kConsumers) {
      const int m = i / wDe, c = i - m * wDe;
      ssum[(size_t)m * De + rank * wDe + c] = sm->acc0[i];
    }
    stage_sync(cx);
    load_rows(cx, ssum, De, O, De);
    run_stage(cx, a.tab.s[a.tab.n - 1], epi_make(nullptr, nullptr, a.d_emb + (size_t)chain * O * d.Kx, d.Kx, O, 0));
  }
  // results that accumulate over the chain
  for (int i = threadIdx.x; i < 4 * wDe; i += kConsumers) a.dw0box[((size_t)chain * De + rank * wDe) * 4 + i] = sm->acc2[i];
  for (int i = threadIdx.x; i < 4 * wH; i += kConsumers) {
    const int j = i / wH, c = i - j * wH;
    a.dwb2[((size_t)chain * 4 + j) * H + rank * wH + c] = sm->acc1[i];
  }
  if (rank == 0) {
    if (threadIdx.x < 4) a.dbb2[(size_t)chain * 4 + threadIdx.x] = dbb2_acc;
    for (int i = threadIdx.x; i < O * 4; i += kConsumers) a.d_box0[(size_t)chain * O * 4 + i] = sm->box[i];
  }
}

// ===================================================================================================================
// grouped weight gradient: dW[n][k] = sum_r Z[r][n] * X[r][k], db[n] = sum_r Z[r][n] over all (chain, t, row) rows.
// CTA = 64 x 64 output tile of one matrix; rows stream through shared memory in chunks of 32 (cp.async, 2 stages);
// 4 warps x (16 n x 64 k) each, 3xTF32 mma.sync with the reduction index = row (both operands are read "transposed"
// from their row-major tiles; row stride 72 floats keeps the fragment loads conflict-free).  Fixed order: deterministic.
// ===================================================================================================================
struct WJob {
  const float* Z; const float* X;     // Z [R][N] row-major; X [Rx][K] row-major
  float* dW; float* db;               // dW [N][ldw] (+ column offset already applied), db [N] or null
  int N, K, ldw;
  int R;                              // rows
  int xdiv, xrows;                    // X row of Z row r: (r / (xdiv * xrows)) * xrows + r % xrows   (xdiv = T-1 for the
                                      // time-invariant embedding, else 1)
  int tile0, tiles_k;                 // first CTA tile of this job, tiles along k
};
struct WArgs { int njobs; WJob job[4 * kMaxLayers + 4]; };

constexpr int kWT = 64, kWR = 32, kWLd = 72;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(128) recur_wgrad_kernel(const __grid_constant__ WArgs a) {
  __shared__ __align__(16) float zs[2][kWR * kWLd];
  __shared__ __align__(16) float xs[2][kWR * kWLd];
  int ji = 0;
  while (ji + 1 < a.njobs && (int)blockIdx.x >= a.job[ji + 1].tile0) ++ji;
  const WJob& j = a.job[ji];
  const int tile = blockIdx.x - j.tile0;
  const int n0 = (tile / j.tiles_k) * kWT, k0 = (tile % j.tiles_k) * kWT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  const int nchunks = (j.R + kWR - 1) / kWR;
  auto issue = [&](int c, int buf) {
    // 32 rows x 64 floats per operand = 512 x 16-byte pieces each, 128 threads -> 4 + 4 per thread
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int piece = threadIdx.x + q * 128, r = piece >> 4, col = (piece & 15) << 2;
      const int row = c * kWR + r;
      const bool rv = row < j.R;
      const bool zv = rv && (n0 + col) < j.N;
      cp_async16(smem_u32(&zs[buf][r * kWLd + col]), j.Z + (zv ? (size_t)row * j.N + n0 + col : 0), zv);
      const bool xv = rv && (k0 + col) < j.K;
      const int xr = rv ? (row / (j.xdiv * j.xrows)) * j.xrows + row % j.xrows : 0;
      cp_async16(smem_u32(&xs[buf][r * kWLd + col]), j.X + (xv ? (size_t)xr * j.K + k0 + col : 0), xv);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  issue(0, 0);
  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) { issue(c + 1, (c + 1) & 1); asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    const float* zb = zs[c & 1];
    const float* xb = xs[c & 1];
#pragma unroll
    for (int rs = 0; rs < kWR / 8; ++rs) {
      // A(m = n index, k = row): a0 = Z[r = rs*8+t][n = 16*warp+g], a1 = n+8, a2 = r+4, a3 = (r+4, n+8)
      const float* zr = zb + (rs * 8 + t) * kWLd + warp * 16 + g;
      const float a4[4] = {zr[0], zr[8], zr[4 * kWLd], zr[4 * kWLd + 8]};
      uint32_t ah[4], al[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_tf32(a4[i], ah[i], al[i]);
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
        // B(k = row, n = k index): b0 = X[r = rs*8+t][k = 8*kt+g], b1 = X[r+4][..]
        const float* xr = xb + (rs * 8 + t) * kWLd + kt * 8 + g;
        uint32_t bh[2], bl[2];
        split_tf32(xr[0], bh[0], bl[0]); split_tf32(xr[4 * kWLd], bh[1], bl[1]);
        mma_tf32(acc[kt], al, bh);
        mma_tf32(acc[kt], ah, bl);
        mma_tf32(acc[kt], ah, bh);
      }
    }
    __syncthreads();
  }
  const int n_lo = n0 + warp * 16 + g, n_hi = n_lo + 8;
#pragma unroll
  for (int kt = 0; kt < 8; ++kt) {
    const int k = k0 + kt * 8 + 2 * t;
    if (k < j.K) {
      if (n_lo < j.N) *reinterpret_cast<float2*>(j.dW + (size_t)n_lo * j.ldw + k) = make_float2(acc[kt][0], acc[kt][1]);
      if (n_hi < j.N) *reinterpret_cast<float2*>(j.dW + (size_t)n_hi * j.ldw + k) = make_float2(acc[kt][2], acc[kt][3]);
    }
  }
}

// bias gradients: db[n] = sum_r Z[r][n], one warp per 32 columns, rows in order (deterministic)
struct BJob { const float* Z; float* db; int N, R; int blk0; };
struct BArgs { int njobs; BJob job[4 * kMaxLayers + 4]; };
__global__ void __launch_bounds__(256) recur_bgrad_kernel(const __grid_constant__ BArgs a) {
  int ji = 0;
  while (ji + 1 < a.njobs && (int)blockIdx.x >= a.job[ji + 1].blk0) ++ji;
  const BJob& j = a.job[ji];
  __shared__ float part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = (blockIdx.x - j.blk0) * 32 + lane;
  float acc = 0.f;
  if (n < j.N) for (int r = warp; r < j.R; r += 8) acc += j.Z[(size_t)r * j.N + n];
  part[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && n < j.N) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][lane];
    j.db[n] = s;
  }
}

}  // namespace recur
}  // namespace ag2v

using namespace ag2v;
using namespace ag2v::recur;

static int g_recur_core = 0;
// Product core of the recurrence kernels: 0 = fp32 FFMA register tiles (default), 1 = 3xTF32 mma.sync.  Process-wide;
// the packed weights are laid out for the core that is active when ag2v_recur_pack runs, so switch before packing.
extern "C" int ag2v_recur_set_core(int core) { g_recur_core = core ? 1 : 0; return 0; }
extern "C" int ag2v_recur_get_core(void) { return g_recur_core; }

static_assert(sizeof(Smem) <= 227 * 1024, "recurrence kernel: shared-memory image exceeds 227 KB");

static int dims_check(const Dims& d) {
  AG2V_REQUIRE(d.O >= 1 && d.O <= kRows && d.E >= 1 && d.E <= kRows, "recur: at most %d nodes and %d edges per clip and timestep (got O=%d, E=%d)",
               kRows, kRows, d.O, d.E);
  AG2V_REQUIRE(d.T >= 2, "recur: at least 2 timesteps (got %d)", d.T);
  AG2V_REQUIRE(d.NL >= 1 && d.NL <= kMaxLayers, "recur: 1..%d graph layers (got %d)", kMaxLayers, d.NL);
  AG2V_REQUIRE(d.Kx % 8 == 0 && d.De % 8 == 0 && d.Dp % 8 == 0 && d.H % 8 == 0 && d.Dout % 8 == 0 && d.Dpo % 8 == 0,
               "recur: feature widths must be multiples of 8");
  return AG2V_OK;
}

static Dims mk_dims(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS) {
  Dims d; d.O = O; d.E = E; d.T = T; d.Kx = Kx; d.De = De; d.Dp = Dp; d.H = H; d.Dout = Dout; d.Dpo = Dpo; d.NL = NL; d.CS = CS;
  return d;
}

// Largest cluster size (16, 8, 4, 2, 1) whose column split fits the kernel, 0 if none does.
extern "C" int ag2v_recur_cluster_size(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL) {
  Dims d = mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, 1);
  if (d.O < 1 || d.O > kRows || d.E < 1 || d.E > kRows || d.T < 2 || d.NL < 1 || d.NL > kMaxLayers) return 0;
  if (Kx % 8 || De % 8 || Dp % 8 || H % 8 || Dout % 8 || Dpo % 8) return 0;
  for (int cs = 16; cs >= 1; cs >>= 1) {
    d.CS = cs;
    Plan p;
    make_plan(d, p);
    if (plan_ok(p)) return cs;
  }
  return 0;
}

// 1 if a cluster of CS CTAs fits these widths (every CTA's column share a multiple of 8 and at most 72 wide).
extern "C" int ag2v_recur_cluster_fits(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS) {
  if (CS < 1 || CS > 16 || (CS & (CS - 1))) return 0;
  Dims d = mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, CS);
  if (d.O < 1 || d.O > kRows || d.E < 1 || d.E > kRows || d.T < 2 || d.NL < 1 || d.NL > kMaxLayers) return 0;
  if (Kx % 8 || De % 8 || Dp % 8 || H % 8 || Dout % 8 || Dpo % 8) return 0;
  Plan p;
  make_plan(d, p);
  return plan_ok(p) ? 1 : 0;
}

extern "C" size_t ag2v_recur_pack_floats(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS) {
  Plan p;
  make_plan(mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, CS), p);
  return (size_t)p.total;
}
extern "C" int ag2v_recur_num_params(int NL) { return 2 + 8 * NL + 4; }
extern "C" size_t ag2v_recur_saved_floats(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int NC) {
  return (size_t)saved_layout(mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, 1), NC).total;
}
extern "C" size_t ag2v_recur_z_floats(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int NC) {
  return (size_t)z_layout(mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, 1), NC).total;
}

// params: host array of device pointers in the order documented at param_index_box().
extern "C" int ag2v_recur_pack(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS,
                               const float* const* params, float* pack, cudaStream_t stream) {
  Dims d = mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, CS);
  if (int rc = dims_check(d)) return rc;
  AG2V_REQUIRE(params && pack, "recur_pack: null pointer");
  Plan p;
  make_plan(d, p);
  AG2V_REQUIRE(plan_ok(p), "recur_pack: cluster size %d does not fit these widths", CS);
  PackArgs pa;
  pa.njobs = 0;
  pa.mma = g_recur_core;
  auto push = [&](const Stage& s) {
    PackJob& j = pa.job[pa.njobs++];
    j.W = params[s.src]; j.ld = s.ld; j.trans = s.trans; j.N = s.N; j.K = s.K; j.ncta = s.ncta; j.CS = CS; j.nseg = s.nseg; j.off = s.off;
    for (int k = 0; k < 3; ++k) { j.base[k] = s.base[k]; j.w[k] = s.w[k]; }
  };
  for (int i = 0; i < p.nf; ++i) push(p.f[i]);
  for (int i = 0; i < p.nb; ++i) push(p.b[i]);
  for (int i = 0; i < pa.njobs; ++i) AG2V_REQUIRE(pa.job[i].W, "recur_pack: parameter %d is null", i);
  recur_pack_kernel<<<dim3(64, pa.njobs), 256, 0, stream>>>(pa, pack);
  AG2V_LAUNCH_CHECK();
  SmallArgs sa;
  sa.njobs = 0;
  auto small = [&](const float* src, long long off, int rows, int cols, int ld, int col0) {
    SmallJob& j = sa.job[sa.njobs++];
    j.src = src; j.off = off; j.rows = rows; j.cols = cols; j.ld = ld; j.col0 = col0;
  };
  small(params[0], p.sm.w0box, De, 4, Kx + 4, Kx);
  for (int l = 0; l < NL; ++l) {
    const int b = 2 + 8 * l;
    small(params[b + 1], p.sm.bias[l][0], 1, H, H, 0);
    small(params[b + 3], p.sm.bias[l][1], 1, d.N2(), d.N2(), 0);
    small(params[b + 5], p.sm.bias[l][2], 1, H, H, 0);
    small(params[b + 7], p.sm.bias[l][3], 1, Dout, Dout, 0);
  }
  const int bx = param_index_box(d);
  small(params[bx + 1], p.sm.bb1, 1, H, H, 0);
  small(params[bx + 2], p.sm.wb2, 4, H, H, 0);
  small(params[bx + 3], p.sm.bb2, 1, 4, 4, 0);
  for (int i = 0; i < sa.njobs; ++i) AG2V_REQUIRE(sa.job[i].src, "recur_pack: small parameter %d is null", i);
  recur_pack_small_kernel<<<sa.njobs, 256, 0, stream>>>(sa, pack);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

template <class Args>
static int launch_cluster(void (*kernel)(const Args), const Args& a, int NC, int CS, cudaStream_t stream) {
  const size_t smem = sizeof(Smem);
  AG2V_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CS > 8) AG2V_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NC * CS);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ::ag2v::count_launch();
  AG2V_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
  return AG2V_OK;
}

// How many clusters of `CS` CTAs of the recurrence kernel the device can hold at once (0 = cannot launch).
extern "C" int ag2v_recur_max_active_clusters(int CS) {
  const size_t smem = sizeof(Smem);
  if (cudaFuncSetAttribute((const void*)recur_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  if (CS > 8 && cudaFuncSetAttribute((const void*)recur_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, (const void*)recur_fwd_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// Forward of the whole recurrence for NC chains (n_models weight sets, NC / n_models chains each, model-major).
//   emb [NC][O][Kx], pred [NC][T][E][Dp] (model-major); box0 [NC/n_models][O][4], edges [NC/n_models][T][E][2] int64,
//   ind [NC/n_models][T][E] uint8 are data shared by the models
//   objv [NC][T][O][Dout], boxes [NC][T][O][4], saved: ag2v_recur_saved_floats floats (kept for the backward)
static unsigned long long* g_recur_prof = nullptr;
// Debug: device buffer of 1 + 2 * 4000 uint64 that CTA 0 of the next forward launches fills with (event, globaltimer)
// pairs (1 = A-operand load issued, 2 = operand in shared memory, 3 = MMA loop done, 4 = epilogue done, 5 = cluster
// handshake done); NULL switches it off.  tools/recur_timeline.py prints the breakdown.
extern "C" int ag2v_recur_set_profile(unsigned long long* buf) { g_recur_prof = buf; return 0; }

extern "C" int ag2v_recur_fwd(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS, int NC,
                              int n_models, const float* const* packs, const float* emb, const float* box0,
                              const float* pred, const long long* edges, const unsigned char* ind, float* objv,
                              float* boxes, float* saved, cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  Dims d = mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, CS);
  if (int rc = dims_check(d)) return rc;
  AG2V_REQUIRE(NC >= 1 && n_models >= 1 && n_models <= kMaxModels && NC % n_models == 0, "recur_fwd: bad chain / model counts %d / %d", NC, n_models);
  AG2V_REQUIRE(packs && emb && box0 && pred && edges && ind && objv && boxes && saved, "recur_fwd: null pointer");
  Plan p;
  make_plan(d, p);
  AG2V_REQUIRE(plan_ok(p), "recur_fwd: cluster size %d does not fit these widths", CS);
  FwdArgs a;
  a.d = d; a.tab = to_table(p.f, p.nf); a.small = p.sm; a.NC = NC; a.chains_per_model = NC / n_models;
  for (int i = 0; i < kMaxModels; ++i) a.model[i] = i < n_models ? packs[i] : nullptr;
  for (int i = 0; i < n_models; ++i) AG2V_REQUIRE(a.model[i], "recur_fwd: packed model %d is null", i);
  a.emb = emb; a.box0 = box0; a.pred = pred; a.edges = edges; a.ind = ind; a.objv = objv; a.boxes = boxes; a.saved = saved;
  a.prof = g_recur_prof;
  a.mma = g_recur_core;
  return launch_cluster(recur_fwd_kernel, a, NC, CS, stream);
}

// Backward chain of ONE model: data gradients + the Z buffer for ag2v_recur_wgrad.  The model's chains are
// [chain0, chain0 + nchains) of a forward launch over NC chains (saved / z are that launch's buffers); every other
// array is model-local: boxes [nchains][T][O][4] (forward output), d_objv / d_boxes (may be null = zero),
// d_emb [nchains][O][Kx], d_box0 [nchains][O][4], d_pred [nchains][T][E][Dp]; per-chain partial sums dw0box
// [nchains][De][4], dwb2 [nchains][4][H], dbb2 [nchains][4] (the caller adds them over the chains).
extern "C" int ag2v_recur_bwd(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int CS, int NC,
                              int chain0, int nchains, const float* pack, const float* saved, const float* boxes,
                              const float* d_objv, const float* d_boxes, const long long* edges,
                              const unsigned char* ind, float* z, float* d_emb, float* d_box0, float* d_pred,
                              float* dw0box, float* dwb2, float* dbb2, cudaStream_t stream) {
  if (int rc = check_arch()) return rc;
  Dims d = mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, CS);
  if (int rc = dims_check(d)) return rc;
  AG2V_REQUIRE(NC >= 1 && chain0 >= 0 && nchains >= 1 && chain0 + nchains <= NC, "recur_bwd: bad chain range %d + %d of %d", chain0, nchains, NC);
  AG2V_REQUIRE(pack && saved && boxes && edges && ind && z && d_emb && d_box0 && d_pred && dw0box && dwb2 && dbb2, "recur_bwd: null pointer");
  Plan p;
  make_plan(d, p);
  AG2V_REQUIRE(plan_ok(p), "recur_bwd: cluster size %d does not fit these widths", CS);
  BwdArgs a;
  a.d = d; a.tab = to_table(p.b, p.nb); a.small = p.sm; a.NC = NC; a.chain0 = chain0; a.model = pack;
  a.saved = saved; a.boxes = boxes; a.d_objv = d_objv; a.d_boxes = d_boxes; a.edges = edges; a.ind = ind; a.z = z;
  a.d_emb = d_emb; a.d_box0 = d_box0; a.d_pred = d_pred; a.dw0box = dw0box; a.dwb2 = dwb2; a.dbb2 = dbb2;
  a.mma = g_recur_core;
  return launch_cluster(recur_bwd_kernel, a, nchains, CS, stream);
}

// Weight and bias gradients of ONE model from the chains [chain0, chain0 + nchains) of a forward / backward pair that
// ran NC chains: grads = host array of device pointers in the parameter order (entries may be null to skip);
// obj_vecs_net[0].weight receives only its first Kx columns here (the 4 box columns come from dw0box), and
// box_net[2] comes from dwb2 / dbb2.
extern "C" int ag2v_recur_wgrad(int O, int E, int T, int Kx, int De, int Dp, int H, int Dout, int Dpo, int NL, int NC,
                                int chain0, int nchains, const float* saved, const float* z, const float* emb,
                                float* const* grads, cudaStream_t stream) {   // emb: model-local
  if (int rc = check_arch()) return rc;
  Dims d = mk_dims(O, E, T, Kx, De, Dp, H, Dout, Dpo, NL, 1);
  if (int rc = dims_check(d)) return rc;
  AG2V_REQUIRE(saved && z && emb && grads && chain0 >= 0 && nchains >= 1 && chain0 + nchains <= NC, "recur_wgrad: bad arguments");
  const Saved sv = saved_layout(d, NC);
  const Zbuf zz = z_layout(d, NC);
  const long long ct0 = (long long)chain0 * (T - 1), ctn = (long long)nchains * (T - 1);
  WArgs wa; wa.njobs = 0;
  BArgs ba; ba.njobs = 0;
  int tiles = 0, blks = 0;
  auto job2 = [&](const float* Z, long long R, int rows, int N, const float* X, int K, int xdiv, float* dW, int ldw, float* db) {
    if (dW) {
      WJob& j = wa.job[wa.njobs++];
      j.Z = Z; j.X = X; j.dW = dW; j.db = nullptr; j.N = N; j.K = K; j.ldw = ldw; j.R = (int)R;
      j.xdiv = xdiv; j.xrows = rows; j.tile0 = tiles; j.tiles_k = (K + kWT - 1) / kWT;
      tiles += ((N + kWT - 1) / kWT) * j.tiles_k;
    }
    if (db) {
      BJob& b = ba.job[ba.njobs++];
      b.Z = Z; b.db = db; b.N = N; b.R = (int)R; b.blk0 = blks;
      blks += (N + 31) / 32;
    }
  };
  auto job = [&](long long zoff, int rows, int N, const float* X, long long xoff, int K, int xdiv, float* dW, int ldw, float* db) {
    job2(z + zoff + ct0 * rows * N, ctn * rows, rows, N, X + xoff, K, xdiv, dW, ldw, db);
  };
  // obj_vecs_net[0].weight[:, :Kx] = (sum_t zu0)^T emb: the backward leaves the per-chain sum in the Z buffer; emb is
  // the model's own [nchains][O][Kx]
  job2(z + zz.ssum + (long long)chain0 * O * De, (long long)nchains * O, O, De, emb, Kx, 1, grads[0], Kx + 4, nullptr);
  job(zz.zx0, O, De, saved, sv.u0 + ct0 * O * De, De, 1, grads[1], De, nullptr);
  for (int l = 0; l < NL; ++l) {
    const int b = 2 + 8 * l, K1 = d.K1(l), N2 = d.N2();
    job(zz.z1[l], E, H, saved, sv.trow[l] + ct0 * E * K1, K1, 1, grads[b + 0], K1, grads[b + 1]);
    job(zz.z2[l], E, N2, saved, sv.h1[l] + ct0 * E * H, H, 1, grads[b + 2], H, grads[b + 3]);
    job(zz.z3[l], O, H, saved, sv.pooled[l] + ct0 * O * H, H, 1, grads[b + 4], H, grads[b + 5]);
    job(zz.z4[l], O, Dout, saved, sv.g1[l] + ct0 * O * H, H, 1, grads[b + 6], H, grads[b + 7]);
  }
  const int bx = param_index_box(d);
  job(zz.zb1, O, H, saved, sv.nobj[NL - 1] + ct0 * O * Dout, Dout, 1, grads[bx + 0], Dout, grads[bx + 1]);
  if (tiles > 0) {
    recur_wgrad_kernel<<<tiles, 128, 0, stream>>>(wa);
    AG2V_LAUNCH_CHECK();
  }
  if (blks > 0) {
    recur_bgrad_kernel<<<blks, 256, 0, stream>>>(ba);
    AG2V_LAUNCH_CHECK();
  }
  return AG2V_OK;
}
