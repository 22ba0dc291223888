// K8 - the SyncBN statistics exchange over NVLink peer memory (reference: the all-reduce of the per-channel sums in
// sync_batchnorm/batchnorm.py:74-83 forward and :105-145 backward, there a master/slave pipe between device threads).
//
// One process per GPU.  Every rank owns one window of device memory that all ranks of the node map (CUDA IPC):
//     data  [kSlots][world][cap] doubles     slot = call number % kSlots, one row per SOURCE rank
//     flags [kSlots][kMaxChunks][kMaxWorld]  u32, the call number of the last complete row chunk
//     calls, tickets                         u32, the number of exchanges this rank has finished (device-side, so that
//                                            a launch captured in a CUDA graph numbers itself correctly on every replay)
// An exchange is ONE kernel per rank: every CTA owns a chunk of the vector, stores its chunk into the row `rank` of
// EVERY rank's window (16-byte stores over NVLink / NVSwitch), publishes the call number with a system-scope release
// store per destination, spins (acquire) until all `world` rows of its own window carry that number, and sums the rows
// in rank order - the same order on every rank, so all ranks hold bit-identical totals, and the same totals run to run.
// No host synchronisation, no NCCL launch: the latency is one NVLink store round (a few microseconds), which is what
// the 2C..5C doubles of a batch-norm layer need; NCCL's ring / tree set-up costs several times that per call.
//
// A rank can be at most one call ahead of any other (it needs everyone's row of call k before it leaves call k), so
// kSlots >= 2 makes reuse safe; flags only ever grow, nothing is reset.  A peer that never arrives (a crashed rank)
// trips a 20 s device-side deadline and traps instead of hanging the GPU.
#include "common.cuh"
#include <string.h>

namespace ag2v {
namespace peer {

constexpr int kSlots = 4;
constexpr int kMaxWorld = 16;
constexpr int kChunk = 512;            // doubles per CTA: 256 threads x one 16-byte store per destination
constexpr int kMaxChunks = 64;         // cap = 32768 doubles per call
constexpr int kThreads = 256;

__host__ __device__ inline size_t data_doubles(int world, int cap) { return (size_t)kSlots * world * cap; }
__host__ __device__ inline size_t window_bytes(int world, int cap) {
  return data_doubles(world, cap) * sizeof(double) + (size_t)kSlots * kMaxChunks * kMaxWorld * sizeof(unsigned) + 64;
}
__host__ __device__ inline size_t counter_offset(int world, int cap) {
  return data_doubles(world, cap) * sizeof(double) + (size_t)kSlots * kMaxChunks * kMaxWorld * sizeof(unsigned);
}

struct Args {
  void* win[kMaxWorld];                // window base of every rank, as mapped in THIS process
  double* vec;                         // in: this rank's sums; out: the totals (in place)
  int n, cap, rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(kThreads) exchange_kernel(const Args a) {
  const int chunk = blockIdx.x, tid = threadIdx.x;
  // call number: 1 + the exchanges already finished on this window set.  Launches on one window set are stream-ordered,
  // every CTA of this launch reads the same value, and the last CTA to leave bumps it for the next launch.
  unsigned* counter = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(a.win[a.rank]) + counter_offset(a.world, a.cap));
  const unsigned seq = *reinterpret_cast<volatile unsigned*>(counter) + 1u;
  const int slot = seq % kSlots;
  const int i0 = chunk * kChunk + 2 * tid;                       // this thread's pair of doubles
  const size_t row = ((size_t)slot * a.world + a.rank) * a.cap;  // where this rank's row lives in every window
  double2 mine = make_double2(0.0, 0.0);
  if (i0 < a.n) {
    mine.x = a.vec[i0];
    if (i0 + 1 < a.n) mine.y = a.vec[i0 + 1];
    for (int p = 0; p < a.world; ++p) {
      const int dst = (a.rank + p) % a.world;                    // start with itself, spread the NVLink traffic
      double* d = reinterpret_cast<double*>(a.win[dst]) + row + i0;
      *reinterpret_cast<double2*>(d) = mine;                     // cap is even and rows are 16-byte aligned
    }
  }
  __threadfence_system();
  __syncthreads();
  const size_t flag_off = data_doubles(a.world, a.cap) * sizeof(double);
  if (tid < a.world) {
    unsigned* f = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(a.win[tid]) + flag_off) +
                  ((size_t)slot * kMaxChunks + chunk) * kMaxWorld + a.rank;
    st_release_sys(f, seq);
    // now wait for row `tid` of the own window
    const unsigned* g = reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(a.win[a.rank]) + flag_off) +
                        ((size_t)slot * kMaxChunks + chunk) * kMaxWorld + tid;
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(g) != seq) {
      if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 20000000000ull) {
        printf("ag2v peer exchange: rank %d waited 20 s for rank %d (call %u, chunk %d)\n", a.rank, tid, seq, chunk);
        __trap();
      }
    }
  }
  __syncthreads();
  if (i0 < a.n) {
    const double* base = reinterpret_cast<const double*>(a.win[a.rank]) + (size_t)slot * a.world * a.cap + i0;
    double2 s = make_double2(0.0, 0.0);
    for (int r = 0; r < a.world; ++r) {                          // rank order: identical totals on every rank
      const double2 v = __ldcg(reinterpret_cast<const double2*>(base + (size_t)r * a.cap));
      s.x += v.x; s.y += v.y;
    }
    a.vec[i0] = s.x;
    if (i0 + 1 < a.n) a.vec[i0 + 1] = s.y;
  }
  __syncthreads();                       // every thread of this CTA has read `counter`
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(counter + 1, 1u) == gridDim.x - 1) {           // last CTA of this launch
      counter[1] = 0;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(counter) = seq;
    }
  }
}

}  // namespace peer
}  // namespace ag2v

using namespace ag2v;
using namespace ag2v::peer;

// Window size in bytes for `world` ranks exchanging up to `cap` doubles per call (cap even, <= 32768).
extern "C" size_t ag2v_peer_window_bytes(int world, int cap) {
  if (world < 1 || world > kMaxWorld || cap < 2 || (cap & 1) || cap > kChunk * kMaxChunks) return 0;
  return window_bytes(world, cap);
}

// cudaMalloc'ed, zeroed window (plain cudaMalloc: CUDA IPC cannot export memory of a caching / virtual-memory pool).
extern "C" int ag2v_peer_window_alloc(int world, int cap, void** window) {
  AG2V_REQUIRE(window, "peer_window_alloc: null pointer");
  const size_t bytes = ag2v_peer_window_bytes(world, cap);
  AG2V_REQUIRE(bytes > 0, "peer_window_alloc: world=%d (1..%d), cap=%d (even, 2..%d)", world, kMaxWorld, cap, kChunk * kMaxChunks);
  AG2V_CUDA(cudaMalloc(window, bytes));
  AG2V_CUDA(cudaMemset(*window, 0, bytes));
  AG2V_CUDA(cudaDeviceSynchronize());
  return AG2V_OK;
}

extern "C" int ag2v_peer_window_free(void* window) {
  if (window) AG2V_CUDA(cudaFree(window));
  return AG2V_OK;
}

// 64-byte CUDA IPC handle of a window, to be sent to the other ranks (any host channel), and its import there.
extern "C" int ag2v_peer_window_export(void* window, unsigned char* handle64) {
  AG2V_REQUIRE(window && handle64, "peer_window_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  AG2V_CUDA(cudaIpcGetMemHandle(&h, window));
  memcpy(handle64, &h, 64);
  return AG2V_OK;
}

extern "C" int ag2v_peer_window_import(const unsigned char* handle64, void** window) {
  AG2V_REQUIRE(window && handle64, "peer_window_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  AG2V_CUDA(cudaIpcOpenMemHandle(window, h, cudaIpcMemLazyEnablePeerAccess));
  return AG2V_OK;
}

extern "C" int ag2v_peer_window_close(void* window) {
  if (window) AG2V_CUDA(cudaIpcCloseMemHandle(window));
  return AG2V_OK;
}

// In-place sum of vec[0..n) (doubles) over the `world` ranks of a node.  windows[r] = rank r's window as mapped in this
// process (windows[rank] = the own allocation).  All ranks must issue the same calls on a window set in the same order,
// and the calls on one window set must be ordered on the device (one stream, or event dependencies): the call number
// is counted in the window itself, so the launch can be captured in a CUDA graph and replayed.  Asynchronous on
// `stream`; the result is bit-identical on all ranks.
extern "C" int ag2v_peer_allreduce_f64(double* vec, int n, void* const* windows, int rank, int world, int cap,
                                       cudaStream_t stream) {
  AG2V_REQUIRE(vec && windows, "peer_allreduce: null pointer");
  AG2V_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "peer_allreduce: rank %d of %d", rank, world);
  AG2V_REQUIRE(ag2v_peer_window_bytes(world, cap) > 0, "peer_allreduce: bad cap %d", cap);
  AG2V_REQUIRE(n >= 0 && n <= cap, "peer_allreduce: n=%d exceeds the window capacity %d", n, cap);
  AG2V_REQUIRE(((uintptr_t)vec & 7) == 0, "peer_allreduce: vec must be 8-byte aligned");
  if (n == 0 || world == 1) return AG2V_OK;
  Args a{};
  for (int r = 0; r < world; ++r) {
    AG2V_REQUIRE(windows[r], "peer_allreduce: window of rank %d is null", r);
    a.win[r] = windows[r];
  }
  a.vec = vec; a.n = n; a.cap = cap; a.rank = rank; a.world = world;
  exchange_kernel<<<ceil_div(n, kChunk), kThreads, 0, stream>>>(a);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// =====================================================================================================================
// K8b - gradient all-reduce(avg) over the same kind of CUDA-IPC windows with the COPY ENGINES doing the transport.
// NCCL moves the 530 MB of gradients of a step through SMs (16-32 CTAs per collective); next to a compute stream made
// of persistent kernels that hold every SM those CTAs are placed late and the collectives occupy 5.4 ms of GPU time per
// step for 1.3 ms of work.  Here the bytes move by cudaMemcpyAsync between mapped windows (no SM),
// and SMs only run two flag kernels of one CTA and one HBM-bound reduction kernel per bucket:
//     sync A  -> every rank's bucket is complete in its window
//     pull    -> rank r copies slice r of the bucket from every peer into its staging area          (copy engines)
//     reduce  -> slice r = (own + staged) / world, in rank order                                    (one kernel)
//     sync B  -> every rank's slice is reduced
//     pull    -> rank r copies the reduced slices of the other ranks into its bucket                (copy engines)
//     post C  -> "I no longer read your window for this bucket" (waited for before the bucket is refilled)
// Flags carry the bucket's own exchange count (kept on the device), so everything replays inside a CUDA graph.
namespace ag2v {
namespace peer {

constexpr int kCeMaxBuckets = 64;

// flags: [bucket][phase 0..2][source rank] u32, then one exchange counter per bucket (how often the bucket was refilled)
__host__ __device__ inline size_t ce_flag_index(int bucket, int phase, int src) { return ((size_t)bucket * 3 + phase) * kMaxWorld + src; }
constexpr size_t kCeCounterBase = (size_t)kCeMaxBuckets * 3 * kMaxWorld;
constexpr size_t kCeFlagWords = kCeCounterBase + kCeMaxBuckets;

struct CeSyncArgs {
  unsigned* flags[kMaxWorld];          // flag block of every rank's window, as mapped here
  int bucket, phase, rank, world;
  int bump, post, wait, back;          // bump: this is the refill of the bucket (counter += 1 first); wait for counter - back
};

__global__ void ce_sync_kernel(const CeSyncArgs a) {
  __shared__ unsigned seq_s;
  const int tid = threadIdx.x;
  if (tid == 0) {
    volatile unsigned* c = a.flags[a.rank] + kCeCounterBase + a.bucket;
    unsigned v = *c;
    if (a.bump) { v += 1u; *c = v; }
    seq_s = v;
  }
  __syncthreads();
  const unsigned seq = seq_s;
  if (tid < a.world) {
    if (a.post) {
      __threadfence_system();          // everything this stream did before (kernels, copies) is ordered before the flag
      st_release_sys(a.flags[tid] + ce_flag_index(a.bucket, a.phase, a.rank), seq);
    }
    if (a.wait) {
      const unsigned want = seq - (unsigned)a.back;
      const unsigned* g = a.flags[a.rank] + ce_flag_index(a.bucket, a.phase, tid);
      const unsigned long long t0 = globaltimer_ns();
      unsigned spins = 0;
      while ((int)(ld_acquire_sys(g) - want) < 0) {
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 20000000000ull) {
          printf("ag2v gradient exchange: rank %d waited 20 s for rank %d (bucket %d phase %d exchange %u)\n", a.rank, tid, a.bucket, a.phase, seq);
          __trap();
        }
      }
    }
  }
}

// own[i] = (own[i] + sum_k staged[k][i]) * scale, parts = world - 1 staged copies of n floats each (n % 4 == 0)
__global__ void __launch_bounds__(256) ce_reduce_kernel(float* __restrict__ own, const float* __restrict__ staged, int parts,
                                                        long long n, float scale) {
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 s = reinterpret_cast<const float4*>(own)[i];
    for (int k = 0; k < parts; ++k) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(staged + (size_t)k * n) + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
    reinterpret_cast<float4*>(own)[i] = s;
  }
}

}  // namespace peer
}  // namespace ag2v

// Plain device memory that CUDA IPC can export (cudaMalloc, zeroed); export / import / close / free as for the windows.
extern "C" int ag2v_peer_alloc(size_t bytes, void** ptr) {
  AG2V_REQUIRE(ptr && bytes > 0, "peer_alloc: bad arguments");
  AG2V_CUDA(cudaMalloc(ptr, bytes));
  AG2V_CUDA(cudaMemset(*ptr, 0, bytes));
  AG2V_CUDA(cudaDeviceSynchronize());
  return AG2V_OK;
}

extern "C" size_t ag2v_ce_flag_bytes(void) { return ag2v::peer::kCeFlagWords * sizeof(unsigned); }

// Asynchronous copy between two mapped windows (or inside one): the copy engines move the bytes.
extern "C" int ag2v_peer_memcpy(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
  AG2V_REQUIRE(dst && src, "peer_memcpy: null pointer");
  if (bytes) AG2V_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream));
  return AG2V_OK;
}

// One CTA.  Every bucket counts its exchanges in the own flag block (bump = 1 on the refill that starts an exchange);
// post = publish "phase of this exchange reached" into every rank's flag block; wait = spin until every rank has
// published it for exchange number (own count - back).  flags[r] = rank r's flag block as mapped in this process.
extern "C" int ag2v_ce_sync(void* const* flags, int rank, int world, int bucket, int phase, int bump, int post, int wait,
                            int back, cudaStream_t stream) {
  using namespace ag2v::peer;
  AG2V_REQUIRE(flags && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "ce_sync: rank %d of %d", rank, world);
  AG2V_REQUIRE(bucket >= 0 && bucket < kCeMaxBuckets && phase >= 0 && phase < 3, "ce_sync: bucket %d phase %d", bucket, phase);
  CeSyncArgs a{};
  for (int r = 0; r < world; ++r) {
    AG2V_REQUIRE(flags[r], "ce_sync: flag block of rank %d is null", r);
    a.flags[r] = (unsigned*)flags[r];
  }
  a.bucket = bucket; a.phase = phase; a.rank = rank; a.world = world; a.bump = bump; a.post = post; a.wait = wait; a.back = back;
  ce_sync_kernel<<<1, 32, 0, stream>>>(a);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}

// own[0..n) = (own + the `parts` staged copies) * scale; n a multiple of 4, pointers 16-byte aligned.
extern "C" int ag2v_ce_reduce(float* own, const float* staged, int parts, long long n, float scale, int ctas,
                              cudaStream_t stream) {
  AG2V_REQUIRE(own && (staged || parts == 0) && n >= 0 && (n & 3) == 0, "ce_reduce: bad arguments (n=%lld)", n);
  AG2V_REQUIRE((((uintptr_t)own | (uintptr_t)staged) & 15) == 0, "ce_reduce: pointers must be 16-byte aligned");
  if (n == 0) return AG2V_OK;
  long long want = ag2v::ceil_div_ll(n >> 2, 256);
  if (ctas < 1) ctas = 1;
  ag2v::peer::ce_reduce_kernel<<<(unsigned)(want < ctas ? want : ctas), 256, 0, stream>>>(own, staged, parts, n, scale);
  AG2V_LAUNCH_CHECK();
  return AG2V_OK;
}
