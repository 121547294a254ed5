// Shared device/host helpers for the DeepJ sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/deepj_b200.h"

// ---- error plumbing (C-ABI: every export returns int, never throws) ---------
void dj_set_error(const char* fmt, ...);
#define DJ_CHECK_ARG(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) { dj_set_error(__VA_ARGS__); return -1; }        \
  } while (0)
#define DJ_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      dj_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (int)e__;                                                                  \
    }                                                                                   \
  } while (0)
#define DJ_LAUNCH_CHECK() DJ_CUDA(cudaGetLastError())

// deterministic-reduction plumbing (api.cu): *out = the workspace registered for a stream (nullptr: none; registered
// but smaller than need_floats: error), and  out[r*ldo + c] += sum over p < P, in order, of part[p*pstride + r*ldp + c]
int dj_reduce_workspace(void* stream, int64_t need_floats, float** out);
int dj_ordered_reduce(const float* part, int P, int64_t pstride, int64_t rows, int cols, int64_t ldp, float* out,
                      int64_t ldo, void* stream);

static inline int dj_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

#ifdef __CUDACC__
// Blackwell packed fp32 FMA (two independent IEEE fp32 FMAs per issue slot, SASS
// FFMA2): the recurrent h.U product is CUDA-core work (fp32 recurrence, see
// DESIGN.md) and plain 3-register FFMA issues at half rate on sm_100.
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
#endif

// ---- keras.optimizers.Nadam.get_updates (model.py:152), one element ------------
// Shared by dj_nadam_step and dj_nadam_allreduce_peer; the roundings are spelled out so both kernels produce the
// same bits whatever the compiler would contract.
#ifdef __CUDACC__
__device__ __forceinline__ float dj_nadam_one(float p, float gi, float& m, float& v, float lr, float beta1,
                                              float beta2, float eps, float mu_t, float mu_t1, float inv_1m_ms_new,
                                              float inv_1m_ms_next, float inv_bias2) {
  const float g_prime = __fmul_rn(gi, inv_1m_ms_new);
  const float mt = __fmaf_rn(beta1, m, __fmul_rn(1.f - beta1, gi));
  const float m_prime = __fmul_rn(mt, inv_1m_ms_next);
  const float vt = __fmaf_rn(beta2, v, __fmul_rn(__fmul_rn(1.f - beta2, gi), gi));
  const float v_prime = __fmul_rn(vt, inv_bias2);
  const float m_bar = __fmaf_rn(mu_t1, m_prime, __fmul_rn(1.f - mu_t, g_prime));
  m = mt;
  v = vt;
  return __fsub_rn(p, __fdiv_rn(__fmul_rn(lr, m_bar), __fadd_rn(__fsqrt_rn(v_prime), eps)));
}
#endif

// ---- counter-based dropout masks --------------------------------------------
// A mask bit is a pure function of (seed, site, element index), so the backward
// kernels regenerate it instead of reading a stored mask.  One lowbias32 round of
// (element group + site key) per 32-bit word -- the hash is made for counters, and
// the glue kernels are bound by these integer instructions (round 1: two rounds,
// 17 instructions per word); for rates whose threshold is a whole number of
// 1/256ths one word serves 4 consecutive elements (one byte each), otherwise one
// word per element.  tests/ re-implement the same function in numpy.
__host__ __device__ __forceinline__ uint32_t dj_mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t dj_site_key(uint64_t seed, int site) {
  uint32_t k = dj_mix32((uint32_t)seed ^ 0x9E3779B9U * (uint32_t)(site + 1));
  return dj_mix32(k + (uint32_t)(seed >> 32));
}
// word for element-group q of a site
__host__ __device__ __forceinline__ uint32_t dj_mask_word(uint32_t key, uint32_t q) {
  return dj_mix32(q + key);
}
// the per-step site key may live in device memory (graph replay: include/deepj_b200.h); call once per kernel
__device__ __forceinline__ void dj_resolve(dj_dropout& d) {
  if (d.key_ptr != nullptr) d.key = __ldg(d.key_ptr);
}
// single element keep decision (element index e within the site's padded layout)
__device__ __forceinline__ bool dj_keep(const dj_dropout& d, uint32_t e) {
  if (d.mode == 1) {
    uint32_t w = dj_mask_word(d.key, e >> 2);
    return ((w >> (8 * (e & 3))) & 0xffU) >= (d.thr >> 24);
  }
  return dj_mask_word(d.key, e) >= d.thr;
}
// multiplier (0 or scale) for one element; identity when the site is off
__device__ __forceinline__ float dj_dropmul(const dj_dropout& d, uint32_t e) {
  if (d.mode == 0) return 1.0f;
  return dj_keep(d, e) ? d.scale : 0.0f;
}
// 4 consecutive elements starting at e (e % 4 == 0)
__device__ __forceinline__ void dj_dropmul4(const dj_dropout& d, uint32_t e, float m[4]) {
  if (d.mode == 0) { m[0] = m[1] = m[2] = m[3] = 1.0f; return; }
  if (d.mode == 1) {
    uint32_t w = dj_mask_word(d.key, e >> 2), t = d.thr >> 24;
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = (((w >> (8 * j)) & 0xffU) >= t) ? d.scale : 0.0f;
    return;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) m[j] = (dj_mask_word(d.key, e + j) >= d.thr) ? d.scale : 0.0f;
}

// ---- activations (Keras-2 semantics) ------------------------------------------
__device__ __forceinline__ float dj_hard_sigmoid(float x) {
  return fminf(fmaxf(0.2f * x + 0.5f, 0.0f), 1.0f);
}
__device__ __forceinline__ float dj_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dj_gate_act(float x, int hard) {
  return hard ? dj_hard_sigmoid(x) : dj_sigmoid(x);
}
// derivative of the gate activation expressed through the stored activation a
__device__ __forceinline__ float dj_gate_dact(float a, int hard) {
  return hard ? ((a > 0.0f && a < 1.0f) ? 0.2f : 0.0f) : a * (1.0f - a);
}

__device__ __forceinline__ float dj_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> __device__ __forceinline__ T dj_from_float(float v);
template <> __device__ __forceinline__ float dj_from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 dj_from_float<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}
