// Data-parallel gradient exchange fused with the optimizer: one kernel per step and per GPU that
//   1. waits until every rank's local gradient is complete (flag exchange over NVLink),
//   2. reduces ITS slice of the flat gradient by loading the slice from every rank's buffer (peer loads, fixed
//      rank order, so every run sums in the same order),
//   3. applies keras.optimizers.Nadam (model.py:152) to that slice -- the m / v moments of a slice live only on its
//      owner -- and stores the updated weights into every rank's parameter buffer (peer stores),
//   4. waits until every rank has finished reading and writing before it lets the next step start.
// It replaces "NCCL all-reduce of the flat gradient, then dj_nadam_step on the whole buffer" (SURVEY.md 8e); the
// reference itself is single-device (train.py:29 model.fit).
//
// The buffers are plain cudaMalloc allocations shared between the one-process-per-GPU ranks with CUDA IPC handles
// (dj_peer_alloc / dj_peer_open); NVSwitch gives every pair of GPUs the full link bandwidth, so the all-to-all
// access pattern of a one-kernel reduce-scatter + all-gather is the natural one here.
#include "dj_common.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int PEER_MAX = 16;
constexpr int FLAG_A = 0;            // [PEER_MAX] "gradient of rank j ready", written by rank j
constexpr int FLAG_B = PEER_MAX;     // [PEER_MAX] "rank j finished its reads and writes"
constexpr int FLAG_DONE = 2 * PEER_MAX;      // local: CTAs of this launch that finished their slice
constexpr int FLAG_STATUS = 2 * PEER_MAX + 1;  // local: 0 ok, 1 = a wait timed out (results invalid)

struct PeerPtrs {
  float* p[PEER_MAX];
  const float* g[PEER_MAX];
  uint32_t* flags[PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;   // L2-only: the line's home is the owning GPU's L2, nothing of it may linger in this SM's L1
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// Spin until the local flag word reaches `epoch` (epochs only grow; compared as a signed difference so the counter
// may wrap).  Bounded: after timeout_ns the launch gives up and raises the status word instead of hanging.
// Returns false on a timeout.
__device__ __forceinline__ bool wait_flag(uint32_t* flag, uint32_t epoch, uint32_t* status,
                                          unsigned long long timeout_ns) {
  const unsigned long long t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    __nanosleep(64);
    if (globaltimer_ns() - t0 > timeout_ns) {
      atomicExch(status, 1u);
      return false;
    }
  }
  return true;
}

// epoch_dev / sc_dev (nullable): the epoch and the ten Nadam scalars of dj_nadam_step_dev read from device memory, so
// that a captured CUDA graph of the training step replays with the step's own values
__global__ void __launch_bounds__(256) peer_nadam_kernel(PeerPtrs pp, int rank, int world, float* __restrict__ m,
                                                         float* __restrict__ v, int64_t lo, int64_t hi,
                                                         uint32_t epoch, unsigned long long timeout_ns, float gscale,
                                                         float lr, float beta1, float beta2, float eps, float mu_t, float mu_t1,
                                                         float inv_1m_ms_new, float inv_1m_ms_next,
                                                         float inv_bias2, const uint32_t* __restrict__ epoch_dev,
                                                         const float* __restrict__ sc_dev) {
  if (sc_dev != nullptr) {
    epoch = *epoch_dev;
    gscale = sc_dev[0]; lr = sc_dev[1]; beta1 = sc_dev[2]; beta2 = sc_dev[3]; eps = sc_dev[4]; mu_t = sc_dev[5];
    mu_t1 = sc_dev[6]; inv_1m_ms_new = sc_dev[7]; inv_1m_ms_next = sc_dev[8]; inv_bias2 = sc_dev[9];
  }
  uint32_t* my = pp.flags[rank];
  const int tid = threadIdx.x;
  // ---- 1. every rank's gradient is complete.  This kernel is stream-ordered behind the local backward pass, so
  // its start already means "my gradient is in memory"; tell everybody, then wait for everybody.
  if (blockIdx.x == 0 && tid < world) st_release_sys(pp.flags[tid] + FLAG_A + rank, epoch);
  if (tid < world) wait_flag(my + FLAG_A + tid, epoch, my + FLAG_STATUS, timeout_ns);
  __syncthreads();
  // A timeout is NOT destructive: if any wait of this launch -- or of an earlier one: the status word is sticky --
  // gave up, the gradient of some rank is missing, so this CTA touches neither the moments nor anybody's weights.
  // (A rank that never arrives makes every CTA of every rank time out alike; the host reads the status word once per
  // epoch, before the checkpoint callback, and aborts: PeerNadam.raise_if_timed_out.)
  const bool healthy = (ld_acquire_sys(my + FLAG_STATUS) == 0u);

  // ---- 2.+3. reduce my slice over the ranks, Nadam, broadcast the new weights
  const float* gl = pp.g[rank];
  float* pl = pp.p[rank];
  for (int64_t i = lo + 4 * ((int64_t)blockIdx.x * blockDim.x + tid); healthy && i < hi;
       i += 4 * (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 x = (r == rank) ? *reinterpret_cast<const float4*>(gl + i) : ld_peer_f4(pp.g[r] + i);
      s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
    }
    float4 pv = *reinterpret_cast<const float4*>(pl + i);
    float4 mv = *reinterpret_cast<const float4*>(m + i);
    float4 vv = *reinterpret_cast<const float4*>(v + i);
    pv.x = dj_nadam_one(pv.x, __fmul_rn(s.x, gscale), mv.x, vv.x, lr, beta1, beta2, eps, mu_t, mu_t1, inv_1m_ms_new, inv_1m_ms_next, inv_bias2);
    pv.y = dj_nadam_one(pv.y, __fmul_rn(s.y, gscale), mv.y, vv.y, lr, beta1, beta2, eps, mu_t, mu_t1, inv_1m_ms_new, inv_1m_ms_next, inv_bias2);
    pv.z = dj_nadam_one(pv.z, __fmul_rn(s.z, gscale), mv.z, vv.z, lr, beta1, beta2, eps, mu_t, mu_t1, inv_1m_ms_new, inv_1m_ms_next, inv_bias2);
    pv.w = dj_nadam_one(pv.w, __fmul_rn(s.w, gscale), mv.w, vv.w, lr, beta1, beta2, eps, mu_t, mu_t1, inv_1m_ms_new, inv_1m_ms_next, inv_bias2);
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(pp.p[r] + i) = pv;
  }

  // ---- 4. my peer stores are visible system-wide before the last CTA of this launch says so
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (tid == 0) {
    const uint32_t done = atomicAdd(my + FLAG_DONE, 1u);
    last = (done == gridDim.x - 1);
    if (last) my[FLAG_DONE] = 0;     // nobody else touches it until the next launch
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  if (tid < world) {
    st_release_sys(pp.flags[tid] + FLAG_B + rank, epoch);
    // my weights are complete, and nobody still reads my gradient, once every rank has said the same
    wait_flag(my + FLAG_B + tid, epoch, my + FLAG_STATUS, timeout_ns);
  }
}

}  // namespace

extern "C" int64_t dj_peer_flag_words(void) { return 2 * PEER_MAX + 2; }

extern "C" int dj_peer_alloc(int64_t bytes, void** ptr, unsigned char* handle64) {
  DJ_CHECK_ARG(bytes > 0 && ptr && handle64, "dj_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  void* p = nullptr;
  DJ_CUDA(cudaMalloc(&p, (size_t)bytes));
  DJ_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    dj_set_error("dj_peer_alloc: cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    return (int)e;
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

extern "C" int dj_peer_open(const unsigned char* handle64, void** ptr) {
  DJ_CHECK_ARG(handle64 && ptr, "dj_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  DJ_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int dj_peer_close(void* ptr) {
  DJ_CHECK_ARG(ptr, "dj_peer_close: NULL pointer");
  DJ_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int dj_peer_free(void* ptr) {
  DJ_CHECK_ARG(ptr, "dj_peer_free: NULL pointer");
  DJ_CUDA(cudaFree(ptr));
  return 0;
}

static int peer_launch(float* const* peer_params, const float* const* peer_grads, uint32_t* const* peer_flags, int rank,
                       int world, float* m, float* v, int64_t n, uint32_t epoch, float gscale, float lr, float beta1,
                       float beta2, float eps, float mu_t, float mu_t1, float m_sched_new, float m_sched_next, float bias2,
                       const uint32_t* epoch_dev, const float* sc_dev, void* stream);

extern "C" int dj_nadam_allreduce_peer(float* const* peer_params, const float* const* peer_grads,
                                       uint32_t* const* peer_flags, int rank, int world, float* m, float* v,
                                       int64_t n, uint32_t epoch, float gscale, float lr, float beta1, float beta2,
                                       float eps, float mu_t, float mu_t1, float m_sched_new, float m_sched_next,
                                       float bias2, void* stream) {
  DJ_CHECK_ARG(epoch != 0, "dj_nadam_allreduce_peer: epoch counts from 1");
  DJ_CHECK_ARG(m_sched_new < 1.f && m_sched_next < 1.f && bias2 > 0.f, "dj_nadam_allreduce_peer: bad schedule scalars");
  return peer_launch(peer_params, peer_grads, peer_flags, rank, world, m, v, n, epoch, gscale, lr, beta1, beta2, eps, mu_t,
                     mu_t1, m_sched_new, m_sched_next, bias2, nullptr, nullptr, stream);
}

extern "C" int dj_nadam_allreduce_peer_dev(float* const* peer_params, const float* const* peer_grads,
                                           uint32_t* const* peer_flags, int rank, int world, float* m, float* v,
                                           int64_t n, const uint32_t* epoch_dev, const float* sc, void* stream) {
  DJ_CHECK_ARG(epoch_dev && sc, "dj_nadam_allreduce_peer_dev: NULL device parameter block");
  return peer_launch(peer_params, peer_grads, peer_flags, rank, world, m, v, n, 1u, 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f,
                     0.f, 1.f, epoch_dev, sc, stream);
}

static int peer_launch(float* const* peer_params, const float* const* peer_grads, uint32_t* const* peer_flags, int rank,
                       int world, float* m, float* v, int64_t n, uint32_t epoch, float gscale, float lr, float beta1,
                       float beta2, float eps, float mu_t, float mu_t1, float m_sched_new, float m_sched_next, float bias2,
                       const uint32_t* epoch_dev, const float* sc_dev, void* stream) {
  DJ_CHECK_ARG(peer_params && peer_grads && peer_flags && m && v, "dj_nadam_allreduce_peer: NULL pointer");
  DJ_CHECK_ARG(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world,
               "dj_nadam_allreduce_peer: rank %d / world %d (at most %d ranks)", rank, world, PEER_MAX);
  DJ_CHECK_ARG(n > 0 && n % 4 == 0, "dj_nadam_allreduce_peer: n must be a positive multiple of 4");
  PeerPtrs pp;
  for (int r = 0; r < world; ++r) {
    DJ_CHECK_ARG(peer_params[r] && peer_grads[r] && peer_flags[r], "dj_nadam_allreduce_peer: rank %d pointers", r);
    pp.p[r] = peer_params[r]; pp.g[r] = peer_grads[r]; pp.flags[r] = peer_flags[r];
  }
  for (int r = world; r < PEER_MAX; ++r) { pp.p[r] = nullptr; pp.g[r] = nullptr; pp.flags[r] = nullptr; }
  // slice of this rank: whole float4s, the last rank takes what is left
  int64_t chunk = ((n / 4 + world - 1) / world) * 4;
  int64_t lo = (int64_t)rank * chunk, hi = lo + chunk;
  if (lo > n) lo = n;
  if (hi > n) hi = n;
  int64_t blocks = (hi - lo + 1023) / 1024;
  const int maxb = dj_num_sms();        // every CTA polls the ready flags: keep the whole launch resident
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  // bound on each of the two waits: 20 s unless DJ_PEER_TIMEOUT_MS says otherwise
  static unsigned long long timeout_ns = 0;
  if (timeout_ns == 0) {
    const char* e = getenv("DJ_PEER_TIMEOUT_MS");
    const long ms = e ? atol(e) : 0;
    timeout_ns = (unsigned long long)(ms > 0 ? ms : 20000) * 1000000ull;
  }
  peer_nadam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      pp, rank, world, m, v, lo, hi, epoch, timeout_ns, gscale, lr, beta1, beta2, eps, mu_t, mu_t1, 1.f / (1.f - m_sched_new),
      1.f / (1.f - m_sched_next), 1.f / bias2, epoch_dev, sc_dev);
  DJ_LAUNCH_CHECK();
  return 0;
}
