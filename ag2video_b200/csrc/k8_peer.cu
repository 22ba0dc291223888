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
