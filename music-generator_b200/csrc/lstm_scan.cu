// Persistent LSTM recurrence for the DeepJ time axis and note axis.
//
// Replaces the sequential part of keras.layers.LSTM at model.py:84 and
// model.py:120 (the tf.while_loop of h.U matmuls + gate nonlinearities), forward
// and backward.  Design (B200-first, not the grid-barrier version of SURVEY H4):
// the independent sequences are tiled over thread-block CLUSTERS; inside a
// cluster the recurrent matrix U [units, 4*units] is split by hidden unit over
// the C CTAs and stays resident in shared memory (fp32) for the whole launch;
// h_t is exchanged through distributed shared memory with ONE cluster barrier
// per step; the cell state never leaves registers.  No grid-wide barrier exists
// because sequences in different clusters never interact.
//
// Row addressing: all activations are [M, width] with the canonical row order
// row = (b*T + t)*48 + n; a (sequence, step) pair maps to a row through ScanMap.
// Gate columns use the library-internal GATE-INTERLEAVED order col = 4*unit + gate
// (gate 0..3 = i,f,c,o), so the four gates of a cell are one 16-byte vector; the
// host converts from/to the Keras [.., 4U] block order at the API boundary.
#include <cooperative_groups.h>

#include "dj_common.cuh"

namespace cg = cooperative_groups;

#ifdef DJ_TRACE
// debug build only (tools/scan_trace.py fp32): clock64() stamps of thread 0 of block 0
__device__ long long* g_dj_ftrace = nullptr;
extern "C" int dj_debug_ftrace_set(void* buf) { return (int)cudaMemcpyToSymbol(g_dj_ftrace, &buf, sizeof(buf)); }
#define DJ_FTR(t, k)                                                                   \
  do {                                                                                 \
    if (g_dj_ftrace != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && (t) < 512)    \
      g_dj_ftrace[(t) * 8 + (k)] = clock64();                                          \
  } while (0)
#else
#define DJ_FTR(t, k) do {} while (0)
#endif

namespace {

struct ScanMap {
  int seq_inner;
  int64_t outer_stride, inner_stride, step_stride;
};
__device__ __forceinline__ int64_t scan_row0(const ScanMap& m, int seq) {
  return (int64_t)(seq / m.seq_inner) * m.outer_stride + (int64_t)(seq % m.seq_inner) * m.inner_stride;
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p));
}

constexpr int UC = 32;   // hidden units owned by one CTA (=> 128 gate columns)

// ---- DSMEM all-gather without a cluster barrier -----------------------------------------------
// h_t goes to every CTA of the cluster as st.async stores that complete on the DESTINATION's
// mbarrier (complete_tx bytes), so the consumer waits on a local mbarrier and nobody executes a
// barrier.cluster (whose release is a gpu-scope membar that also drains the gate stores to HBM).
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_async_f4(uint32_t raddr, float a, float b, float c, float d, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];\n" ::"r"(raddr),
               "f"(a), "f"(b), "f"(c), "f"(d), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void sbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) {   // ~2 s: a protocol bug must not hang the GPU
      printf("deepj lstm_scan: h all-gather wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void store_dz4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_dz4(__nv_bfloat16* p, const float v[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

template <int U, int BS>
struct FwdSmem {
  static constexpr int HSTR = U * 4 + 4;            // floats per 4-sequence group (+4: bank skew)
  static constexpr int HBUF = (BS / 4) * HSTR;      // one h buffer
  static constexpr int US = U * UC * 4;             // resident U slice
  static constexpr size_t BYTES = sizeof(float) * (size_t)(US + 2 * HBUF) + 16;   // + two mbarriers
};

// ---------------------------------------------------------------------------
// forward, many sequences
//   thread (sg, ug): sequences 4*sg..4*sg+3 of the tile, hidden units ug and
//   ug+16 of this CTA's slice, all four gates -> 32 accumulators, c in registers
//   (16 FFMA2 per 3 shared-memory loads: the best FMA : load ratio).  BS = 32 gives
//   4 warps, one per scheduler.
// ---------------------------------------------------------------------------
template <int U, int C, int BS>
__global__ void __launch_bounds__((BS / 4) * 16, 1)
scan_fwd_kernel(float* __restrict__ Z, float* __restrict__ Hout, float* __restrict__ Cout,
                __nv_bfloat16* __restrict__ Hbf, const float* __restrict__ Uw, int S, int steps,
                ScanMap map, int hard) {
  static_assert(U / C == UC, "each CTA owns 32 hidden units");
  constexpr int NT = (BS / 4) * 16;
  using SM = FwdSmem<U, BS>;
  extern __shared__ __align__(16) float smem[];
  float* Us = smem;               // [U][32 units][4 gates]
  float* hbuf = smem + SM::US;    // [2][BS/4][U*4+4]

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / C, ncl = gridDim.x / C;
  const int tid = threadIdx.x, sg = tid >> 4, ug = tid & 15;
  const int unit0 = rank * UC;

  for (int idx = tid; idx < SM::US; idx += NT)   // [k][32 units][4 gates] is a contiguous 512 B run per k
    Us[idx] = Uw[(size_t)(idx >> 7) * 4 * U + 4 * unit0 + (idx & 127)];
  const uint32_t bar0 = sm_u32(hbuf + 2 * SM::HBUF);   // bar0 + 8*b: "buffer b holds the whole h of a step"
  if (tid == 0) {
    sbar_init(bar0, 1); sbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  uint32_t rdst[C], rbar[C];
#pragma unroll
  for (int r = 0; r < C; ++r) { rdst[r] = mapa_u32(sm_u32(hbuf), r); rbar[r] = mapa_u32(bar0, r); }
  constexpr uint32_t STEP_BYTES = (uint32_t)C * NT * 32;   // every thread of the cluster sends 2 x 16 B to each CTA
  uint32_t ph[2] = {0u, 0u};
  cluster.sync();   // every CTA of the cluster is resident, its barriers initialised, before any remote store

  const int ntiles = (S + BS - 1) / BS;
  for (int tile = cid; tile < ntiles; tile += ncl) {
    for (int idx = tid; idx < SM::HBUF; idx += NT) hbuf[idx] = 0.f;   // h_{-1} = 0 (buffer 0; nothing is in flight)
    float c[4][2];
    bool ok[4];
    int64_t row0[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int seq = tile * BS + sg * 4 + s;
      ok[s] = seq < S;
      row0[s] = scan_row0(map, ok[s] ? seq : 0);
      c[s][0] = c[s][1] = 0.f;
    }
    float4 zn[4][2];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
        zn[s][hf] = ok[s] ? *reinterpret_cast<const float4*>(Z + row0[s] * (4 * U) + 4 * (unit0 + hf * 16 + ug))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    for (int t = 0; t < steps; ++t) {
      const int cur = t & 1, nxt = cur ^ 1;
      if (tid == 0) sbar_expect_tx(bar0 + 8 * nxt, STEP_BYTES);   // h_t will arrive in buffer nxt
      uint64_t acc2[4][2][2];   // [seq][unit half][gate pair (i,f) / (g,o)]
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          acc2[s][hf][0] = pack2(zn[s][hf].x, zn[s][hf].y);
          acc2[s][hf][1] = pack2(zn[s][hf].z, zn[s][hf].w);
        }
      if (t + 1 < steps) {   // register prefetch of the next step's pre-activations
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int64_t r = row0[s] + (int64_t)(t + 1) * map.step_stride;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
            zn[s][hf] = ok[s] ? *reinterpret_cast<const float4*>(Z + r * (4 * U) + 4 * (unit0 + hf * 16 + ug))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      DJ_FTR(t, 0);
      if (t > 0) { sbar_wait(bar0 + 8 * cur, ph[cur]); ph[cur] ^= 1u; }   // all of h_{t-1} has landed
      DJ_FTR(t, 1);
      const float* hb = hbuf + cur * SM::HBUF + sg * SM::HSTR;
      const float* ua = Us + ug * 4;
#pragma unroll 4
      for (int k = 0; k < U; ++k) {
        const float4 hv = *reinterpret_cast<const float4*>(hb + k * 4);
        const ulonglong2 u0 = *reinterpret_cast<const ulonglong2*>(ua + k * (UC * 4));
        const ulonglong2 u1 = *reinterpret_cast<const ulonglong2*>(ua + k * (UC * 4) + 64);
        const uint64_t hs[4] = {pack2(hv.x, hv.x), pack2(hv.y, hv.y), pack2(hv.z, hv.z), pack2(hv.w, hv.w)};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          acc2[s][0][0] = ffma2(hs[s], u0.x, acc2[s][0][0]);
          acc2[s][0][1] = ffma2(hs[s], u0.y, acc2[s][0][1]);
          acc2[s][1][0] = ffma2(hs[s], u1.x, acc2[s][1][0]);
          acc2[s][1][1] = ffma2(hs[s], u1.y, acc2[s][1][1]);
        }
      }
      DJ_FTR(t, 2);
      float acc[4][2][4];
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          unpack2(acc2[s][hf][0], acc[s][hf][0], acc[s][hf][1]);
          unpack2(acc2[s][hf][1], acc[s][hf][2], acc[s][hf][3]);
        }
      float hnew[2][4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int64_t r = row0[s] + (int64_t)t * map.step_stride;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const float gi = dj_gate_act(acc[s][hf][0], hard);
          const float gf = dj_gate_act(acc[s][hf][1], hard);
          const float gg = tanhf(acc[s][hf][2]);
          const float go = dj_gate_act(acc[s][hf][3], hard);
          const float cn = fmaf(gf, c[s][hf], gi * gg);
          const float hn = go * tanhf(cn);
          c[s][hf] = cn;
          hnew[hf][s] = hn;
          if (ok[s]) {
            const int col = unit0 + hf * 16 + ug;
            *reinterpret_cast<float4*>(Z + r * (4 * U) + 4 * col) = make_float4(gi, gf, gg, go);
            Hout[r * U + col] = hn;
            if (Cout != nullptr) Cout[r * U + col] = cn;
            if (Hbf != nullptr) {   // h_{t} is the "previous h" of step t+1; step 0 sees zeros
              if (t + 1 < steps) Hbf[(r + map.step_stride) * U + col] = __float2bfloat16_rn(hn);
              if (t == 0) Hbf[r * U + col] = __float2bfloat16_rn(0.f);
            }
          }
        }
      }
      // all-gather of h_t into every CTA's next buffer (DSMEM, 16 lanes x 16 B contiguous).  Two buffers
      // are enough without a barrier: a CTA that is already sending h_{t+1} into buffer `cur` has received
      // every h_t, so every peer has finished reading h_{t-1} from it.
      DJ_FTR(t, 3);
      const uint32_t off = (uint32_t)(nxt * SM::HBUF + sg * SM::HSTR + (unit0 + ug) * 4) * 4u;
#pragma unroll
      for (int r = 0; r < C; ++r) {
        st_async_f4(rdst[r] + off, hnew[0][0], hnew[0][1], hnew[0][2], hnew[0][3], rbar[r] + 8 * nxt);
        st_async_f4(rdst[r] + off + 256, hnew[1][0], hnew[1][1], hnew[1][2], hnew[1][3], rbar[r] + 8 * nxt);
      }
    }
    // drain the last step's (unused) h so that no store of this tile is in flight anywhere in the cluster
    const int last = steps & 1;
    sbar_wait(bar0 + 8 * last, ph[last]); ph[last] ^= 1u;
    cluster.sync();
  }
}

// ---------------------------------------------------------------------------
// forward, few sequences
//   thread (sg, ug) = sequences 4*sg..4*sg+3 of the tile, hidden unit ug of this CTA's slice, four
//   gates -> 16 accumulators (8 packed FFMA2 per k), c in registers.  Every accumulator is one FMA
//   chain over k in order, exactly as in scan_fwd_kernel, so the result does not depend on which of
//   the two kernels ran.  Used when all 16-sequence tiles fit one round of clusters: a generation
//   window (48 sequences) runs on 3 clusters x 8 CTAs x 4 warps instead of one cluster.
// ---------------------------------------------------------------------------
template <int U, int C, int SBS>
__global__ void __launch_bounds__((SBS / 4) * UC, 1)
scan_fwd_u1_kernel(float* __restrict__ Z, float* __restrict__ Hout, float* __restrict__ Cout,
                      __nv_bfloat16* __restrict__ Hbf, const float* __restrict__ Uw, int S, int steps,
                      ScanMap map, int hard) {
  static_assert(U / C == UC, "each CTA owns 32 hidden units");
  constexpr int NT = (SBS / 4) * UC;
  using SM = FwdSmem<U, SBS>;
  extern __shared__ __align__(16) float smem[];
  float* Us = smem;               // [U][32 units][4 gates]
  float* hbuf = smem + SM::US;    // [2][SBS/4][U*4+4]

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / C, ncl = gridDim.x / C;
  const int tid = threadIdx.x, sg = tid >> 5, ug = tid & 31;
  const int unit0 = rank * UC;
  const int col = unit0 + ug;

  for (int idx = tid; idx < SM::US; idx += NT)
    Us[idx] = Uw[(size_t)(idx >> 7) * 4 * U + 4 * unit0 + (idx & 127)];
  const uint32_t bar0 = sm_u32(hbuf + 2 * SM::HBUF);   // bar0 + 8*b: "buffer b holds the whole h of a step"
  if (tid == 0) {
    sbar_init(bar0, 1); sbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  uint32_t rdst[C], rbar[C];
#pragma unroll
  for (int r = 0; r < C; ++r) { rdst[r] = mapa_u32(sm_u32(hbuf), r); rbar[r] = mapa_u32(bar0, r); }
  constexpr uint32_t STEP_BYTES = (uint32_t)C * NT * 16;   // every thread of the cluster sends 16 B to each CTA
  uint32_t ph[2] = {0u, 0u};
  cluster.sync();   // every CTA of the cluster is resident, its barriers initialised, before any remote store

  const int ntiles = (S + SBS - 1) / SBS;
  for (int tile = cid; tile < ntiles; tile += ncl) {
    for (int idx = tid; idx < SM::HBUF; idx += NT) hbuf[idx] = 0.f;   // h_{-1} = 0 (buffer 0; nothing is in flight)
    float c[4];
    bool ok[4];
    int64_t row0[4];
    float4 zn[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int seq = tile * SBS + sg * 4 + s;
      ok[s] = seq < S;
      row0[s] = scan_row0(map, ok[s] ? seq : 0);
      c[s] = 0.f;
      zn[s] = ok[s] ? *reinterpret_cast<const float4*>(Z + row0[s] * (4 * U) + 4 * col) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    for (int t = 0; t < steps; ++t) {
      const int cur = t & 1, nxt = cur ^ 1;
      if (tid == 0) sbar_expect_tx(bar0 + 8 * nxt, STEP_BYTES);   // h_t will arrive in buffer nxt
      uint64_t acc2[4][2];   // [seq][gate pair (i,f) / (g,o)]
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        acc2[s][0] = pack2(zn[s].x, zn[s].y);
        acc2[s][1] = pack2(zn[s].z, zn[s].w);
      }
      if (t + 1 < steps) {   // register prefetch of the next step's pre-activations
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int64_t r = row0[s] + (int64_t)(t + 1) * map.step_stride;
          zn[s] = ok[s] ? *reinterpret_cast<const float4*>(Z + r * (4 * U) + 4 * col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      DJ_FTR(t, 0);
      if (t > 0) { sbar_wait(bar0 + 8 * cur, ph[cur]); ph[cur] ^= 1u; }   // all of h_{t-1} has landed
      DJ_FTR(t, 1);
      const float* hb = hbuf + cur * SM::HBUF + sg * SM::HSTR;
      const float* ua = Us + ug * 4;
#pragma unroll 8
      for (int k = 0; k < U; ++k) {
        const float4 hv = *reinterpret_cast<const float4*>(hb + k * 4);
        const ulonglong2 u0 = *reinterpret_cast<const ulonglong2*>(ua + k * (UC * 4));
        const uint64_t hs[4] = {pack2(hv.x, hv.x), pack2(hv.y, hv.y), pack2(hv.z, hv.z), pack2(hv.w, hv.w)};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          acc2[s][0] = ffma2(hs[s], u0.x, acc2[s][0]);
          acc2[s][1] = ffma2(hs[s], u0.y, acc2[s][1]);
        }
      }
      DJ_FTR(t, 2);
      float hnew[4];
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        float a0, a1, a2, a3;
        unpack2(acc2[s][0], a0, a1);
        unpack2(acc2[s][1], a2, a3);
        const int64_t r = row0[s] + (int64_t)t * map.step_stride;
        const float gi = dj_gate_act(a0, hard);
        const float gf = dj_gate_act(a1, hard);
        const float gg = tanhf(a2);
        const float go = dj_gate_act(a3, hard);
        const float cn = fmaf(gf, c[s], gi * gg);
        const float hn = go * tanhf(cn);
        c[s] = cn;
        hnew[s] = hn;
        if (ok[s]) {
          *reinterpret_cast<float4*>(Z + r * (4 * U) + 4 * col) = make_float4(gi, gf, gg, go);
          Hout[r * U + col] = hn;
          if (Cout != nullptr) Cout[r * U + col] = cn;
          if (Hbf != nullptr) {
            if (t + 1 < steps) Hbf[(r + map.step_stride) * U + col] = __float2bfloat16_rn(hn);
            if (t == 0) Hbf[r * U + col] = __float2bfloat16_rn(0.f);
          }
        }
      }
      // all-gather of h_t into every CTA's next buffer (DSMEM, 32 lanes x 16 B contiguous); see
      // scan_fwd_kernel for why two buffers need no barrier
      DJ_FTR(t, 3);
      const uint32_t off = (uint32_t)(nxt * SM::HBUF + sg * SM::HSTR + col * 4) * 4u;
#pragma unroll
      for (int r = 0; r < C; ++r)
        st_async_f4(rdst[r] + off, hnew[0], hnew[1], hnew[2], hnew[3], rbar[r] + 8 * nxt);
    }
    const int last = steps & 1;
    sbar_wait(bar0 + 8 * last, ph[last]); ph[last] ^= 1u;
    cluster.sync();
  }
}

// ---------------------------------------------------------------------------
// backward (reverse scan).  Per step:
//   D  dh_rec = sum over the C source CTAs of the partial dz.U^T slots
//   A  gate derivatives for this CTA's (sequence, unit) pairs -> dz (global + smem)
//   B  P[s,k] = sum_j dz[s,j] * U[k,col(j)]  over this CTA's 128 gate columns, all k
//   C  reduce-scatter: P[:, units of CTA d] -> slot[my rank] of CTA d (DSMEM)
// ---------------------------------------------------------------------------
template <int U, int C, int BS>
struct BwdSmem {
  static constexpr int UT = 128 * U;                     // U^T slice [j][k]
  static constexpr int DZS = BS + 4;                     // dz row stride (bank skew)
  static constexpr int DZ = 128 * DZS;
  static constexpr int SLOT = (BS / 4) * UC * 4;         // one source's contribution
  static constexpr size_t BYTES = sizeof(float) * (size_t)(UT + DZ + C * SLOT);
};

template <int U, int C, int BS, typename TZ>
__global__ void __launch_bounds__((BS / 4) * 16, 1)
scan_bwd_kernel(const float* __restrict__ G, const float* __restrict__ Cst, const float* __restrict__ dY,
                int64_t ldY, dj_dropout d_y, const float* __restrict__ Uw, TZ* __restrict__ dZ,
                float* __restrict__ db, int S, int steps, ScanMap map, int hard, float* __restrict__ db_part) {
  dj_resolve(d_y);
  static_assert(U / C == UC, "each CTA owns 32 hidden units");
  constexpr int NT = (BS / 4) * 16;
  constexpr int KQ = U / 64;        // k quads per thread in phase B (4 consecutive k each)
  using SM = BwdSmem<U, C, BS>;
  extern __shared__ __align__(16) float smem[];
  float* UsT = smem;                // [128 j = g*32+u][U k]
  float* dzb = smem + SM::UT;       // [128 j][BS+4]
  float* slots = dzb + SM::DZ;      // [C src][BS/4][32 units][4 seqs]

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / C, ncl = gridDim.x / C;
  const int tid = threadIdx.x, sg = tid >> 4, ug = tid & 15;
  const int unit0 = rank * UC;

  for (int idx = tid; idx < SM::UT; idx += NT) {
    const int k = idx % U, j = idx / U;   // j = gate*32 + local unit (smem order); global order is 4*unit + gate
    UsT[idx] = Uw[(size_t)k * 4 * U + 4 * (unit0 + (j & 31)) + (j >> 5)];
  }
  float* rslots[C];
#pragma unroll
  for (int r = 0; r < C; ++r) rslots[r] = cluster.map_shared_rank(slots, r);
  float dbacc[2][4];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf)
#pragma unroll
    for (int g = 0; g < 4; ++g) dbacc[hf][g] = 0.f;
  cluster.sync();

  const int ntiles = (S + BS - 1) / BS;
  for (int tile = cid; tile < ntiles; tile += ncl) {
    bool ok[4];
    int64_t row0[4];
    float dcn[4][2];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int seq = tile * BS + sg * 4 + s;
      ok[s] = seq < S;
      row0[s] = scan_row0(map, ok[s] ? seq : 0);
      dcn[s][0] = dcn[s][1] = 0.f;
    }
    for (int t = steps - 1; t >= 0; --t) {
      // ---- D: recurrent gradient arriving from step t+1
      float dh[4][2];
#pragma unroll
      for (int s = 0; s < 4; ++s) dh[s][0] = dh[s][1] = 0.f;
      if (t != steps - 1) {
#pragma unroll
        for (int r = 0; r < C; ++r) {
          const float* sl = slots + r * SM::SLOT + (sg * UC + ug) * 4;
          const float4 a = *reinterpret_cast<const float4*>(sl);
          const float4 b = *reinterpret_cast<const float4*>(sl + 64);
          dh[0][0] += a.x; dh[1][0] += a.y; dh[2][0] += a.z; dh[3][0] += a.w;
          dh[0][1] += b.x; dh[1][1] += b.y; dh[2][1] += b.z; dh[3][1] += b.w;
        }
      }
      cluster_arrive();   // my reads of `slots` are done; matched by cluster_wait before phase C
      // ---- A: elementwise gate backward
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int64_t r = row0[s] + (int64_t)t * map.step_stride;
        const int64_t rp = r - map.step_stride;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int ul = hf * 16 + ug, col = unit0 + ul;
          float dz[4] = {0.f, 0.f, 0.f, 0.f};
          if (ok[s]) {
            const float4 g4 = *reinterpret_cast<const float4*>(G + r * (4 * U) + 4 * col);
            const float gi = g4.x, gf = g4.y, gg = g4.z, go = g4.w;
            const float ct = Cst[r * U + col];
            const float cp = (t > 0) ? Cst[rp * U + col] : 0.f;
            const float dy = dY[r * ldY + col] * dj_dropmul(d_y, (uint32_t)(r * U + col));
            const float dht = dy + dh[s][hf];
            const float tc = tanhf(ct);
            const float d_o = dht * tc;
            const float dc = fmaf(dht * go, 1.f - tc * tc, dcn[s][hf]);
            dcn[s][hf] = dc * gf;
            dz[0] = dc * gg * dj_gate_dact(gi, hard);
            dz[1] = dc * cp * dj_gate_dact(gf, hard);
            dz[2] = dc * gi * (1.f - gg * gg);
            dz[3] = d_o * dj_gate_dact(go, hard);
            store_dz4(dZ + r * (4 * U) + 4 * col, dz);
            if (t > 0) {   // warm L2 for the next (earlier) step while phase B runs
              const int64_t r2 = rp - map.step_stride;
              prefetch_l2(G + rp * (4 * U) + 4 * col);
              prefetch_l2(dY + rp * ldY + col);
              if (t > 1) prefetch_l2(Cst + r2 * U + col);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            dbacc[hf][g] += dz[g];
            dzb[(g * 32 + ul) * SM::DZS + sg * 4 + s] = dz[g];
          }
        }
      }
      __syncthreads();
      // ---- B: partial dh_{t-1} over this CTA's 128 gate columns
      uint64_t P2[4][KQ][2];
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int q = 0; q < KQ; ++q) P2[s][q][0] = P2[s][q][1] = 0ull;
      if (t > 0) {
#pragma unroll 2
        for (int j = 0; j < 128; ++j) {
          const float4 dv = *reinterpret_cast<const float4*>(dzb + j * SM::DZS + sg * 4);
          const uint64_t d2[4] = {pack2(dv.x, dv.x), pack2(dv.y, dv.y), pack2(dv.z, dv.z), pack2(dv.w, dv.w)};
#pragma unroll
          for (int q = 0; q < KQ; ++q) {
            const ulonglong2 uv = *reinterpret_cast<const ulonglong2*>(UsT + j * U + q * 64 + ug * 4);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
              P2[s][q][0] = ffma2(d2[s], uv.x, P2[s][q][0]);
              P2[s][q][1] = ffma2(d2[s], uv.y, P2[s][q][1]);
            }
          }
        }
      }
      float P[4][KQ][4];
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int q = 0; q < KQ; ++q) {
          unpack2(P2[s][q][0], P[s][q][0], P[s][q][1]);
          unpack2(P2[s][q][1], P[s][q][2], P[s][q][3]);
        }
      cluster_wait();     // every CTA has consumed last step's slots
      // ---- C: reduce-scatter.  k = q*64 + ug*4 + i  ->  owner CTA 2q + (ug>=8), unit (ug*4+i)%32
      if (t > 0) {
#pragma unroll
        for (int q = 0; q < KQ; ++q) {
          float* dst = rslots[2 * q + (ug >> 3)] + rank * SM::SLOT + (sg * UC + (ug & 7) * 4) * 4;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(dst + i * 4) = make_float4(P[0][q][i], P[1][q][i], P[2][q][i], P[3][q][i]);
        }
      }
      cluster.sync();     // slots complete and dzb free for the next step
    }
  }
  // bias gradient: reduce the per-thread sums over the sequence groups, one atomic per column
#pragma unroll
  for (int hf = 0; hf < 2; ++hf)
#pragma unroll
    for (int g = 0; g < 4; ++g) dzb[(g * 32 + hf * 16 + ug) * SM::DZS + sg] = dbacc[hf][g];
  __syncthreads();
  if (tid < 128) {
    float s = 0.f;
    for (int q = 0; q < BS / 4; ++q) s += dzb[tid * SM::DZS + q];
    const int col = 4 * (unit0 + (tid & 31)) + (tid >> 5);
    if (db_part != nullptr) db_part[(size_t)cid * (4 * U) + col] = s;   // deterministic mode: one partial per cluster
    else atomicAdd(db + col, s);
  }
}

// Clusters of C CTAs that can be resident at once.  This is NOT num_sms / C: a cluster lives inside one GPC, and on
// B200 only 15 clusters of 8 one-CTA-per-SM blocks fit (ncu launch__cluster_max_active), so a grid of 16 such
// clusters runs as two waves.  Asked from the driver with the real launch configuration.
template <typename K>
int max_active_clusters(K kernel, int C, int nthreads, size_t smem) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C);
  cfg.blockDim = dim3(nthreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, (const void*)kernel, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = dj_num_sms() / C > 1 ? dj_num_sms() / C - 1 : 1;
  }
  return n;
}

template <typename K>
int prepare_cluster_kernel(K kernel, int C, size_t smem) {
  DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (C > 8) DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  return 0;
}

// Persistent clusters: `ntiles` tiles are walked by as many clusters as can be resident, balanced so that every
// cluster does the same number of rounds.
template <typename K>
int launch_cluster(K kernel, int C, int nthreads, size_t smem, int ntiles, cudaStream_t st, void** args,
                   int* ncl_out = nullptr) {
  int rc = prepare_cluster_kernel(kernel, C, smem);
  if (rc) return rc;
  int ncl = max_active_clusters(kernel, C, nthreads, smem);
  if (ncl > ntiles) ncl = ntiles;
  const int rounds = (ntiles + ncl - 1) / ncl;
  ncl = (ntiles + rounds - 1) / rounds;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ncl * C);
  cfg.blockDim = dim3(nthreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  DJ_CUDA(cudaLaunchKernelExC(&cfg, (const void*)kernel, args));
  if (ncl_out != nullptr) *ncl_out = ncl;
  return 0;
}

}  // namespace

extern "C" int dj_lstm_scan_fwd(float* Z, float* h_out, float* c_out, void* h_prev_bf16, const float* Uw, int S,
                                int steps, int units, int seq_inner, int64_t seq_outer_stride,
                                int64_t seq_inner_stride, int64_t step_stride, int hard, void* stream) {
  DJ_CHECK_ARG(Z && h_out && Uw, "dj_lstm_scan_fwd: NULL pointer");
  DJ_CHECK_ARG(S > 0 && steps > 0 && seq_inner > 0, "dj_lstm_scan_fwd: bad sizes");
  ScanMap map{seq_inner, seq_outer_stride, seq_inner_stride, step_stride};
  __nv_bfloat16* hb = (__nv_bfloat16*)h_prev_bf16;
  void* args[] = {&Z, &h_out, &c_out, &hb, (void*)&Uw, &S, &steps, &map, &hard};
  cudaStream_t st = (cudaStream_t)stream;
  if (units == 256) {
    constexpr int C = 8;
    // few sequences (a generation window is 48): 16-sequence tiles spread the step over more clusters
    if (prepare_cluster_kernel(scan_fwd_u1_kernel<256, C, 16>, C, FwdSmem<256, 16>::BYTES) == 0 &&
        (S + 15) / 16 <= max_active_clusters(scan_fwd_u1_kernel<256, C, 16>, C, 4 * UC, FwdSmem<256, 16>::BYTES))
      return launch_cluster(scan_fwd_u1_kernel<256, C, 16>, C, 4 * UC, FwdSmem<256, 16>::BYTES, (S + 15) / 16, st, args);
    constexpr int BS = 32;
    return launch_cluster(scan_fwd_kernel<256, C, BS>, C, (BS / 4) * 16, FwdSmem<256, BS>::BYTES, (S + BS - 1) / BS, st,
                          args);
  } else if (units == 128) {
    constexpr int C = 4;
    if (prepare_cluster_kernel(scan_fwd_u1_kernel<128, C, 16>, C, FwdSmem<128, 16>::BYTES) == 0 &&
        (S + 15) / 16 <= max_active_clusters(scan_fwd_u1_kernel<128, C, 16>, C, 4 * UC, FwdSmem<128, 16>::BYTES))
      return launch_cluster(scan_fwd_u1_kernel<128, C, 16>, C, 4 * UC, FwdSmem<128, 16>::BYTES, (S + 15) / 16, st, args);
    constexpr int BS = 32;
    return launch_cluster(scan_fwd_kernel<128, C, BS>, C, (BS / 4) * 16, FwdSmem<128, BS>::BYTES, (S + BS - 1) / BS, st,
                          args);
  }
  DJ_CHECK_ARG(false, "dj_lstm_scan_fwd: units=%d unsupported (128 or 256; U must fit cluster shared memory in fp32)", units);
  return -1;
}

extern "C" int dj_lstm_scan_bwd(const float* gates, const float* c, const float* dY, int64_t ldY,
                                dj_dropout d_y, const float* Uw, void* dZ, int dz_dtype, float* db, int S,
                                int steps, int units, int seq_inner, int64_t seq_outer_stride,
                                int64_t seq_inner_stride, int64_t step_stride, int hard, void* stream) {
  DJ_CHECK_ARG(gates && c && dY && Uw && dZ && db, "dj_lstm_scan_bwd: NULL pointer");
  DJ_CHECK_ARG(S > 0 && steps > 0 && seq_inner > 0 && ldY >= units, "dj_lstm_scan_bwd: bad sizes");
  DJ_CHECK_ARG(dz_dtype == DJ_F32 || dz_dtype == DJ_BF16, "dj_lstm_scan_bwd: unknown dz dtype %d", dz_dtype);
  ScanMap map{seq_inner, seq_outer_stride, seq_inner_stride, step_stride};
  cudaStream_t st = (cudaStream_t)stream;
  const int BSu = units == 256 ? 48 : 64, ntiles = (S + BSu - 1) / BSu;   // tiles; launch_cluster picks the resident cluster count
  // deterministic mode: one partial bias gradient per cluster, added in cluster order afterwards
  float* db_part = nullptr;
  if (dj_reduce_workspace(stream, (int64_t)ntiles * 4 * units, &db_part)) return -1;
  void* args[] = {(void*)&gates, (void*)&c, (void*)&dY, &ldY, &d_y, (void*)&Uw, &dZ, &db, &S, &steps, &map, &hard, &db_part};
  int ncl = 0, rc = -1;
  if (units == 256) {
    constexpr int C = 8, BS = 48;
    if (dz_dtype == DJ_F32)
      rc = launch_cluster(scan_bwd_kernel<256, C, BS, float>, C, (BS / 4) * 16, BwdSmem<256, C, BS>::BYTES, ntiles, st, args, &ncl);
    else
      rc = launch_cluster(scan_bwd_kernel<256, C, BS, __nv_bfloat16>, C, (BS / 4) * 16, BwdSmem<256, C, BS>::BYTES, ntiles, st, args, &ncl);
  } else if (units == 128) {
    constexpr int C = 4, BS = 64;
    if (dz_dtype == DJ_F32)
      rc = launch_cluster(scan_bwd_kernel<128, C, BS, float>, C, (BS / 4) * 16, BwdSmem<128, C, BS>::BYTES, ntiles, st, args, &ncl);
    else
      rc = launch_cluster(scan_bwd_kernel<128, C, BS, __nv_bfloat16>, C, (BS / 4) * 16, BwdSmem<128, C, BS>::BYTES, ntiles, st, args, &ncl);
  }
  if (units == 256 || units == 128) {
    if (rc == 0 && db_part != nullptr)
      rc = dj_ordered_reduce(db_part, ncl, (int64_t)4 * units, 1, 4 * units, 4 * units, db, 4 * units, stream);
    return rc;
  }
  DJ_CHECK_ARG(false, "dj_lstm_scan_bwd: units=%d unsupported (128 or 256)", units);
  return -1;
}
