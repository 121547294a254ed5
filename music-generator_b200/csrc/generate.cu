// Persistent note-by-note sampler for generation.
// Replaces the inner loop of generate.py:112-121 (48 x note_model.predict + the
// NumPy MusicGeneration.choose / end_time, generate.py:47-79): one launch walks
// all 48 notes of a timestep carrying the note-axis LSTM state (mathematically
// identical to the reference's 48 full re-evaluations because the note LSTM is
// causal in n), applies the temperature transform in float32 as
// NumPy does, compares float64 uniforms with `<=`, and draws the replay uniform
// only when the note is played.
//
// A cluster of 4 CTAs serves one group of sequences.  CTA r keeps, in fp32 and
// resident in shared memory for the whole launch, the 128 gate-interleaved columns
// (32 hidden units) of U0, W1 and U1 that it owns (3 x 64 KB).  Per note: every CTA
// computes its column slice of the layer-0 gates, updates its 32 cells and pushes
// h0 to all CTAs through distributed shared memory; cluster barrier; same for layer
// 1; cluster barrier; CTA 0 evaluates the heads, samples, and pushes the event row
// (the next note's `chosen` input) to all CTAs; cluster barrier.  fp32 FMA arithmetic
// throughout: the sampled events have to equal the fp32 oracle's bit for bit.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "dj_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int N_ = DJ_NUM_NOTES;
constexpr int MAXG = 8;       // sequences one cluster can walk (reference stream order needs them in one place)
constexpr int UN = 128;       // note-axis units
constexpr int G4 = 4 * UN;    // gate columns
constexpr int CL = 4;         // cluster size
constexpr int CW = G4 / CL;   // 128 columns (= 32 units) per CTA
constexpr int KQ = 4;         // k-split of every mat-vec across thread groups
constexpr int NTHR = CW * KQ; // 512 threads

struct GenSmem {
  float U0[UN * CW], W1[UN * CW], U1[UN * CW];   // [k][local column]
  float part[KQ][CW];                            // k-split partial sums
  float h0[2][MAXG][UN], h1[2][MAXG][UN];        // full hidden states (all-gathered), double-buffered by note
                                                 // parity: a fast peer may already push h_n while this CTA still reads h_{n-1}
  float c0[MAXG][32], c1[MAXG][32];              // cell states of the 32 local units
  float prev[MAXG][4];                           // chosen_{n-1} (play, replay, volume)
  float sp1[MAXG][UN];                           // tanh(Dense(style)) added to h0 before layer 1 (constant over notes)
  float wh[3][UN];                               // head weights: play, replay, volume
  float bh[4];
  double temp[MAXG];
  float head[3];
  int played_any[MAXG];
  double margin[MAXG];
  long long cursor;
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NTHR, 1) gen_sample_kernel(
    const float* __restrict__ zpre, const float* __restrict__ W0c, const float* __restrict__ U0,
    const float* __restrict__ W1, const float* __restrict__ U1, const float* __restrict__ b1,
    const float* __restrict__ sp1, const float* __restrict__ Wn, const float* __restrict__ bn,
    const float* __restrict__ Wv, const float* __restrict__ bv, int gcount, const double* __restrict__ uniforms,
    int64_t* ucursor, int stream_mode, double* temperature, int32_t* silent_time, double default_temp, int hard,
    float* __restrict__ events, float* __restrict__ probs_out, double* __restrict__ margin_out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  GenSmem& S = *reinterpret_cast<GenSmem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, c = tid & (CW - 1), kq = tid >> 7;
  const int g_base = (blockIdx.x / CL) * gcount;

  for (int i = tid; i < UN * CW; i += NTHR) {          // resident weight slices
    const int k = i / CW, cc = i % CW;
    S.U0[i] = U0[k * G4 + rank * CW + cc];
    S.W1[i] = W1[k * G4 + rank * CW + cc];
    S.U1[i] = U1[k * G4 + rank * CW + cc];
  }
  for (int i = tid; i < 2 * MAXG * UN; i += NTHR) { (&S.h0[0][0][0])[i] = 0.f; (&S.h1[0][0][0])[i] = 0.f; }
  for (int i = tid; i < MAXG * 32; i += NTHR) { (&S.c0[0][0])[i] = 0.f; (&S.c1[0][0])[i] = 0.f; }
  if (tid < MAXG) {
    S.prev[tid][0] = S.prev[tid][1] = S.prev[tid][2] = S.prev[tid][3] = 0.f;
    S.played_any[tid] = 0;
    S.margin[tid] = 1e300;
  }
  if (tid == 0) S.cursor = (stream_mode == 0 && ucursor != nullptr) ? *ucursor : 0;
  // everything the per-note critical path would otherwise fetch from global memory
  for (int i = tid; i < gcount * UN; i += NTHR) S.sp1[i / UN][i % UN] = sp1[(int64_t)(g_base + i / UN) * UN + i % UN];
  for (int i = tid; i < UN; i += NTHR) { S.wh[0][i] = Wn[i * 2]; S.wh[1][i] = Wn[i * 2 + 1]; S.wh[2][i] = Wv[i]; }
  if (tid == 0) { S.bh[0] = bn[0]; S.bh[1] = bn[1]; S.bh[2] = bv[0]; }
  if (tid < gcount) S.temp[tid] = temperature[g_base + tid];
  float w0c[3][4], bias1[4];                            // tid < 32: the four gate columns of local unit `tid`
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int gc = rank * CW + 4 * (tid & 31) + q;
    w0c[0][q] = W0c[gc]; w0c[1][q] = W0c[G4 + gc]; w0c[2][q] = W0c[2 * G4 + gc];
    bias1[q] = b1[gc];
  }
  // x.W0 pre-activations of the next (sequence, note) are fetched one unit of work ahead
  auto load_zpre = [&](int gl, int n) {
    return (tid < 32 && n < N_) ? *reinterpret_cast<const float4*>(zpre + ((int64_t)(g_base + gl) * N_ + n) * G4 +
                                                                     rank * CW + 4 * tid)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 zp_next = load_zpre(0, 0);
  GenSmem* peer[CL];
#pragma unroll
  for (int r = 0; r < CL; ++r) peer[r] = cluster.map_shared_rank(&S, r);
  cluster.sync();

  // this thread's k range of every mat-vec
  const int k0 = kq * (UN / KQ);
  auto matvec = [&](const float* __restrict__ Wm, const float* __restrict__ x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < UN / KQ; k += 4) {
      a0 = fmaf(x[k0 + k], Wm[(k0 + k) * CW + c], a0);
      a1 = fmaf(x[k0 + k + 1], Wm[(k0 + k + 1) * CW + c], a1);
      a2 = fmaf(x[k0 + k + 2], Wm[(k0 + k + 2) * CW + c], a2);
      a3 = fmaf(x[k0 + k + 3], Wm[(k0 + k + 3) * CW + c], a3);
    }
    return (a0 + a1) + (a2 + a3);
  };

  for (int n = 0; n < N_; ++n) {
    const int cur = n & 1, nxt = cur ^ 1;
    // the first uniform this note is certain to consume (the rest of its cache line is then in L1)
    double upre0 = 0.0;
    if (rank == 0 && tid == 0)
      upre0 = (stream_mode == 0) ? uniforms[S.cursor] : uniforms[((int64_t)g_base * N_ + n) * 2];
    // ---- note layer 0 for every sequence of the group: z = zpre + chosen_{n-1}.W0[Ut:Ut+3] + h0.U0
    for (int gl = 0; gl < gcount; ++gl) {
      const float4 zp4 = zp_next;
      zp_next = (gl + 1 < gcount) ? load_zpre(gl + 1, n) : load_zpre(0, n + 1);
      S.part[kq][c] = matvec(S.U0, S.h0[cur][gl]);
      __syncthreads();
      if (tid < 32) {   // one thread per local hidden unit: its four gate columns are adjacent
        const float zp[4] = {zp4.x, zp4.y, zp4.z, zp4.w};
        float z[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = 4 * tid + q;
          float acc = zp[q];
          acc = fmaf(S.prev[gl][0], w0c[0][q], acc);
          acc = fmaf(S.prev[gl][1], w0c[1][q], acc);
          acc = fmaf(S.prev[gl][2], w0c[2][q], acc);
          z[q] = acc + ((S.part[0][cc] + S.part[1][cc]) + (S.part[2][cc] + S.part[3][cc]));
        }
        const float gi = dj_gate_act(z[0], hard), gf = dj_gate_act(z[1], hard);
        const float gg = tanhf(z[2]), go = dj_gate_act(z[3], hard);
        const float cn = fmaf(gf, S.c0[gl][tid], gi * gg);
        S.c0[gl][tid] = cn;
        // x1 = h0 + tanh(Dense(style)) is what layer 1 consumes (model.py:113-117 at inference); h0 itself
        // feeds the layer-0 recurrence: store both views
        const float hn = go * tanhf(cn);
#pragma unroll
        for (int r = 0; r < CL; ++r) peer[r]->h0[nxt][gl][rank * 32 + tid] = hn;
      }
      __syncthreads();
    }
    cluster.sync();
    // ---- note layer 1
    for (int gl = 0; gl < gcount; ++gl) {
      // x1 = h0 + sp1 (built on the fly from the all-gathered h0)
      float a = 0.f;
      {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
        for (int k = 0; k < UN / KQ; k += 2) {
          a0 = fmaf(S.h0[nxt][gl][k0 + k] + S.sp1[gl][k0 + k], S.W1[(k0 + k) * CW + c], a0);
          a1 = fmaf(S.h0[nxt][gl][k0 + k + 1] + S.sp1[gl][k0 + k + 1], S.W1[(k0 + k + 1) * CW + c], a1);
        }
        a = a0 + a1;
      }
      S.part[kq][c] = a + matvec(S.U1, S.h1[cur][gl]);
      __syncthreads();
      if (tid < 32) {
        float z[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = 4 * tid + q;
          z[q] = bias1[q] + ((S.part[0][cc] + S.part[1][cc]) + (S.part[2][cc] + S.part[3][cc]));
        }
        const float gi = dj_gate_act(z[0], hard), gf = dj_gate_act(z[1], hard);
        const float gg = tanhf(z[2]), go = dj_gate_act(z[3], hard);
        const float cn = fmaf(gf, S.c1[gl][tid], gi * gg);
        S.c1[gl][tid] = cn;
        const float hn = go * tanhf(cn);
#pragma unroll
        for (int r = 0; r < CL; ++r) peer[r]->h1[nxt][gl][rank * 32 + tid] = hn;
      }
      __syncthreads();
    }
    cluster.sync();
    // ---- heads + sampling on CTA 0 (model.py:94-95, generate.py:47-58), sequences in order
    if (rank == 0) {
      for (int gl = 0; gl < gcount; ++gl) {
        const int g = g_base + gl;
        if (tid < 96) {
          const int o = tid >> 5, lane = tid & 31;
          float s = 0.f;
          for (int k = lane; k < UN; k += 32) s = fmaf(S.h1[nxt][gl][k], S.wh[o][k], s);
          s = dj_warp_sum(s);
          if (lane == 0) S.head[o] = s + S.bh[o];
        }
        __syncthreads();
        if (tid == 0) {
          float p0 = dj_sigmoid(S.head[0]), p1 = dj_sigmoid(S.head[1]);
          const float vol = S.head[2];
          if (probs_out != nullptr) {
            float* po = probs_out + ((int64_t)g * N_ + n) * 3;
            po[0] = p0; po[1] = p1; po[2] = vol;
          }
          const double temp = S.temp[gl];
          if (temp != 1.0) {   // generate.py:81-91, float32 arithmetic like NumPy on a float32 array
            const float tf = (float)temp;
            const float xa = -logf(1.0f / p0 - 1.0f), xb = -logf(1.0f / p1 - 1.0f);
            p0 = 1.0f / (1.0f + expf(-xa / tf));
            p1 = 1.0f / (1.0f + expf(-xb / tf));
          }
          double u1, u2;
          const double* ui = uniforms + ((int64_t)g * N_ + n) * 2;
          if (stream_mode == 0) { u1 = (gl == 0) ? upre0 : uniforms[S.cursor]; S.cursor++; } else u1 = (gl == 0) ? upre0 : ui[0];
          float e0 = 0.f, e1 = 0.f, e2 = 0.f;
          double mg = fabs(u1 - (double)p0);
          if (u1 <= (double)p0) {   // generate.py:52
            e0 = 1.f; e2 = vol;
            if (stream_mode == 0) u2 = uniforms[S.cursor++]; else u2 = ui[1];
            mg = fmin(mg, fabs(u2 - (double)p1));
            if (u2 <= (double)p1) e1 = 1.f;   // generate.py:57
            S.played_any[gl] = 1;
          }
          S.margin[gl] = fmin(S.margin[gl], mg);
          float* ev = events + ((int64_t)g * N_ + n) * 3;
          ev[0] = e0; ev[1] = e1; ev[2] = e2;
#pragma unroll
          for (int r = 0; r < CL; ++r) {
            peer[r]->prev[gl][0] = e0; peer[r]->prev[gl][1] = e1; peer[r]->prev[gl][2] = e2;
          }
        }
        __syncthreads();
      }
    }
    cluster.sync();
  }
  // ---- end_time (generate.py:60-79): silence raises the temperature
  if (rank == 0 && tid < gcount) {
    const int g = g_base + tid;
    // np.count_nonzero(next_note) == 0  <=>  nothing played (a played note sets channel 0 to 1)
    if (!S.played_any[tid]) {
      const int st = silent_time[g] + 1;
      silent_time[g] = st;
      if (st >= DJ_BEAT) temperature[g] = temperature[g] + 0.1;
    } else {
      silent_time[g] = 0;
      temperature[g] = default_temp;
    }
    if (margin_out != nullptr) margin_out[g] = fmin(margin_out[g], S.margin[tid]);
  }
  if (rank == 0 && tid == 0 && stream_mode == 0 && ucursor != nullptr) *ucursor = S.cursor;
  cluster.sync();
}

// ---------------------------------------------------------------------------
// One sequence per cluster (the indexed stream mode, and reference order with a single sequence): the same
// arithmetic with the weights in REGISTERS and no barrier inside the note loop.  The kernel above spends 60 % of a
// note in barriers (ncu: stall_barrier + stall_membar; three barrier.cluster round trips with their membar and L1
// flush, six block barriers), and a first barrier-free version was bound by the shared-memory port instead (three
// 128 x 128 mat-vecs per note = 1 500 warp-wide LDS per CTA and note).  Here
//   * the CTA's 3 x 128 x 128 weight slice lives in the register file: 16 warps, lane = (column c8 of the warp's 8
//     gate columns, k-quarter kq), 32 weights of each of U0, W1, U1 per lane; a mat-vec is 32 FMAs on x read as 8
//     conflict-free LDS.128 plus two shuffles over the four k-quarters;
//   * the four gates of a unit sit in four lane groups of the same warp and are exchanged by shuffle;
//   * h0 and h1 travel as st.async stores that complete on the DESTINATION's mbarrier, double-buffered by note parity;
//     a consumer waits on a local mbarrier: two exchanges per note and nothing else;
//   * the heads and the sampling decision are evaluated REDUNDANTLY by every warp of the cluster from the same
//     all-gathered h1 (same instructions on the same bits: the same decision everywhere), so the sampled event -- the
//     next note's `chosen` input -- never has to be sent anywhere; one lane of the cluster writes the outputs;
//   * U1.h1 and the next note's U0.h0 are computed while the exchanges are in flight: the per-note critical path is
//     cell 0 -> push -> W1.h0 -> cell 1 -> push -> heads + sample.
// W1.(h0 + sp) is evaluated as W1.h0 + (W1.sp + b1) with the bracket precomputed per launch.
// ---------------------------------------------------------------------------
struct Gen1Smem {
  float h0[2][UN], h1[2][UN];                    // all-gathered hidden states, by note parity
  float wh[3][UN];
  float bh[4];
  float sp[UN];
  unsigned long long bar_h0[2], bar_h1[2];
};

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t g_mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void g_st_async_f32(uint32_t raddr, float v, uint32_t rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n" ::"r"(raddr),
               "r"(__float_as_uint(v)), "r"(rbar)
               : "memory");
}
__device__ __forceinline__ void g_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void g_bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_bar_wait(uint32_t bar, uint32_t parity) {
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
      printf("deepj gen_sample1: exchange wait timed out (block %d thread %d bar 0x%x)\n", blockIdx.x, threadIdx.x, bar);
      __trap();
    }
  }
}

constexpr int NTHR1 = 512;    // 16 warps x (8 or 4) gate columns

// CLS = cluster size: 4 (32 units per CTA, 96 weight registers per lane: the register file is full and ~35 of them
// spill to local memory) or 8 (16 units per CTA, 48 weight registers; used when all clusters still fit the GPU at once)
template <int CLS>
__global__ void __launch_bounds__(NTHR1, 1) gen_sample1_kernel(
    const float* __restrict__ zpre, const float* __restrict__ W0c, const float* __restrict__ U0,
    const float* __restrict__ W1, const float* __restrict__ U1, const float* __restrict__ b1,
    const float* __restrict__ sp1, const float* __restrict__ Wn, const float* __restrict__ bn,
    const float* __restrict__ Wv, const float* __restrict__ bv, const double* __restrict__ uniforms,
    int64_t* ucursor, int stream_mode, double* temperature, int32_t* silent_time, double default_temp, int hard,
    float* __restrict__ events, float* __restrict__ probs_out, double* __restrict__ margin_out) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Gen1Smem& S = *reinterpret_cast<Gen1Smem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int CWc = G4 / CLS;                        // gate columns per CTA (128 / 64)
  constexpr int UPCc = CWc / 4;                        // hidden units per CTA (32 / 16)
  constexpr int CPW = CWc / 16;                        // columns per warp (8 / 4)
  constexpr int KQN = 32 / CPW;                        // lanes sharing a column = k-slices (4 / 8) = CLS
  constexpr int KPL = UN / KQN;                        // weights of one matrix per lane (32 / 16)
  constexpr int NJ = KPL / 4;                          // float4 chunks of x per lane (8 / 4)
  static_assert(KQN == CLS, "one pushing lane per destination CTA");
  const int g = blockIdx.x / CLS;                      // the sequence of this cluster
  const int kq = lane % KQN, cw = lane / KQN;          // k-slice; column inside the warp's CPW
  const int col = CPW * warp + cw;                     // local gate column (= 4*unit + gate)
  const int gate = cw & 3, ub16 = (cw >> 2) * 4 * KQN; // first lane of this lane's unit

  // the lane's KPL weights of each matrix: k = 4*KQN*j + 4*kq + i  (j < NJ, i < 4), column `col`
  float wU0[KPL], wW1[KPL], wU1[KPL];
  {
    const size_t cg_ = (size_t)rank * CWc + col;
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const size_t src = (size_t)(4 * KQN * j + 4 * kq + i) * G4 + cg_;
        wU0[4 * j + i] = U0[src]; wW1[4 * j + i] = W1[src]; wU1[4 * j + i] = U1[src];
      }
  }
  for (int i = tid; i < 2 * UN; i += NTHR1) { (&S.h0[0][0])[i] = 0.f; (&S.h1[0][0])[i] = 0.f; }
  for (int i = tid; i < UN; i += NTHR1) {
    S.wh[0][i] = Wn[i * 2]; S.wh[1][i] = Wn[i * 2 + 1]; S.wh[2][i] = Wv[i];
    S.sp[i] = sp1[(int64_t)g * UN + i];
  }
  if (tid == 0) {
    S.bh[0] = bn[0]; S.bh[1] = bn[1]; S.bh[2] = bv[0];
#pragma unroll
    for (int b = 0; b < 2; ++b) { g_bar_init(g_smem_u32(&S.bar_h0[b]), 1); g_bar_init(g_smem_u32(&S.bar_h1[b]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  // one column of W.x: KPL FMAs per lane, then the k-slices are added by shuffle (every lane gets the sum)
  auto matvec = [&](const float (&w)[KPL], const float* __restrict__ x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 x4 = *reinterpret_cast<const float4*>(x + 4 * KQN * j + 4 * kq);
      a0 = fmaf(x4.x, w[4 * j], a0); a1 = fmaf(x4.y, w[4 * j + 1], a1);
      a2 = fmaf(x4.z, w[4 * j + 2], a2); a3 = fmaf(x4.w, w[4 * j + 3], a3);
    }
    float s = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 1; o < KQN; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
  };
  // one LSTM cell from the four gate pre-activations of a unit (lanes ub16 + KQN*gate; every lane of the unit computes it)
  auto cell = [&](float z, float& c) {
    const float zi = __shfl_sync(0xffffffffu, z, ub16), zf = __shfl_sync(0xffffffffu, z, ub16 + KQN);
    const float zg = __shfl_sync(0xffffffffu, z, ub16 + 2 * KQN), zo = __shfl_sync(0xffffffffu, z, ub16 + 3 * KQN);
    const float gi = dj_gate_act(zi, hard), gf = dj_gate_act(zf, hard);
    const float gg = tanhf(zg), go = dj_gate_act(zo, hard);
    c = fmaf(gf, c, gi * gg);
    return go * tanhf(c);
  };
  const int gc = rank * CWc + col;
  const float w0c0 = W0c[gc], w0c1 = W0c[G4 + gc], w0c2 = W0c[2 * G4 + gc];
  const float c1const = b1[gc] + matvec(wW1, S.sp);    // W1.(h0 + sp) + b1 = W1.h0 + (W1.sp + b1)
  const uint32_t sbase = g_smem_u32(&S);
  const uint32_t rdst = g_mapa(sbase, (uint32_t)kq);   // lane kq of a unit's gate-0 group pushes the unit's h to CTA kq
  const bool pusher = (gate == 0);
  const uint32_t off_h0 = (uint32_t)((const uint8_t*)&S.h0[0][0] - (const uint8_t*)&S),
                 off_h1 = (uint32_t)((const uint8_t*)&S.h1[0][0] - (const uint8_t*)&S),
                 off_bh0 = (uint32_t)((const uint8_t*)&S.bar_h0[0] - (const uint8_t*)&S),
                 off_bh1 = (uint32_t)((const uint8_t*)&S.bar_h1[0] - (const uint8_t*)&S);
  cluster.sync();   // every CTA's barriers and buffers are initialised before anybody pushes into them

  const bool writer = (rank == 0 && tid == 0);         // the one lane of the cluster that owns the outputs
  float c0 = 0.f, c1 = 0.f, a0 = 0.f, a1 = 0.f;        // cell states; U0.h0 and U1.h1 of the coming note (h = 0 at n = 0)
  float e0 = 0.f, e1 = 0.f, e2 = 0.f;                  // chosen_{n-1}: the event sampled for the previous note
  int64_t cursor = (stream_mode == 0 && ucursor != nullptr) ? *ucursor : 0;
  const double temp = temperature[g];
  int played_any = 0;
  double margin = 1e300;
  const float* zrow = zpre + (int64_t)g * N_ * G4 + gc;
  float zp = zrow[0];
  const uint32_t hslot = (uint32_t)((rank * UPCc + (col >> 2)) * 4);   // this unit inside an all-gathered h vector
  for (int n = 0; n < N_; ++n) {
    const int cur = n & 1, nxt = cur ^ 1;
    const uint32_t par = (uint32_t)(n >> 1) & 1u;
    if (tid == 0) {   // arm this note's two exchanges (their previous phases were consumed two notes ago)
      g_bar_expect_tx(sbase + off_bh0 + 8 * nxt, UN * 4);
      g_bar_expect_tx(sbase + off_bh1 + 8 * nxt, UN * 4);
    }
    const float zp_next = (n + 1 < N_) ? zrow[(int64_t)(n + 1) * G4] : 0.f;
    // the first uniform this note is certain to consume
    const double u1 = (stream_mode == 0) ? uniforms[cursor] : uniforms[((int64_t)g * N_ + n) * 2];
    // ---- layer 0: z = zpre + chosen_{n-1}.W0[Ut:Ut+3] + U0.h0
    float z = zp;
    z = fmaf(e0, w0c0, z);
    z = fmaf(e1, w0c1, z);
    z = fmaf(e2, w0c2, z);
    z += a0;
    const float h0n = cell(z, c0);
    if (pusher) g_st_async_f32(rdst + off_h0 + (uint32_t)(nxt * UN * 4) + hslot, h0n, rdst + off_bh0 + 8 * nxt);
    zp = zp_next;
    if (n > 0) a1 = matvec(wU1, S.h1[cur]);            // recurrent term of layer 1, while h0 is in flight
    // ---- layer 1: z = (W1.sp + b1) + W1.h0 + U1.h1
    g_bar_wait(sbase + off_bh0 + 8 * nxt, par);
    const float z1 = c1const + matvec(wW1, S.h0[nxt]) + a1;
    const float h1n = cell(z1, c1);
    if (pusher) g_st_async_f32(rdst + off_h1 + (uint32_t)(nxt * UN * 4) + hslot, h1n, rdst + off_bh1 + 8 * nxt);
    a0 = matvec(wU0, S.h0[nxt]);                       // recurrent term of layer 0 of the NEXT note, while h1 is in flight
    // ---- heads + sampling (model.py:94-95, generate.py:47-58), by every warp alike
    g_bar_wait(sbase + off_bh1 + 8 * nxt, par);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = lane; k < UN; k += 32) {
      const float hv = S.h1[nxt][k];
      s0 = fmaf(hv, S.wh[0][k], s0); s1 = fmaf(hv, S.wh[1][k], s1); s2 = fmaf(hv, S.wh[2][k], s2);
    }
    s0 = dj_warp_sum(s0); s1 = dj_warp_sum(s1); s2 = dj_warp_sum(s2);
    float p0 = dj_sigmoid(s0 + S.bh[0]), p1 = dj_sigmoid(s1 + S.bh[1]);
    const float vol = s2 + S.bh[2];
    if (writer && probs_out != nullptr) {
      float* po = probs_out + ((int64_t)g * N_ + n) * 3;
      po[0] = p0; po[1] = p1; po[2] = vol;
    }
    if (temp != 1.0) {   // generate.py:81-91, float32 arithmetic like NumPy on a float32 array
      const float tf = (float)temp;
      const float xa = -logf(1.0f / p0 - 1.0f), xb = -logf(1.0f / p1 - 1.0f);
      p0 = 1.0f / (1.0f + expf(-xa / tf));
      p1 = 1.0f / (1.0f + expf(-xb / tf));
    }
    if (stream_mode == 0) cursor++;
    e0 = 0.f; e1 = 0.f; e2 = 0.f;
    double mg = fabs(u1 - (double)p0);
    if (u1 <= (double)p0) {   // generate.py:52 (warp-uniform: every lane holds the same values)
      e0 = 1.f; e2 = vol;
      const double u2 = (stream_mode == 0) ? uniforms[cursor++] : uniforms[((int64_t)g * N_ + n) * 2 + 1];
      mg = fmin(mg, fabs(u2 - (double)p1));
      if (u2 <= (double)p1) e1 = 1.f;   // generate.py:57
      played_any = 1;
    }
    margin = fmin(margin, mg);
    if (writer) {
      float* ev = events + ((int64_t)g * N_ + n) * 3;
      ev[0] = e0; ev[1] = e1; ev[2] = e2;
    }
  }
  // ---- end_time (generate.py:60-79): silence raises the temperature
  if (writer) {
    if (!played_any) {   // np.count_nonzero(next_note) == 0  <=>  nothing played
      const int st = silent_time[g] + 1;
      silent_time[g] = st;
      if (st >= DJ_BEAT) temperature[g] = temperature[g] + 0.1;
    } else {
      silent_time[g] = 0;
      temperature[g] = default_temp;
    }
    if (margin_out != nullptr) margin_out[g] = fmin(margin_out[g], margin);
    if (stream_mode == 0 && ucursor != nullptr) *ucursor = cursor;
  }
  cluster.sync();   // nobody exits while a peer may still push into it (every push of the last note has been waited for)
}

}  // namespace

extern "C" int dj_gen_sample(const float* zpre, const float* W0c, const float* U0, const float* W1,
                             const float* U1, const float* b1, const float* sp1, const float* Wn,
                             const float* bn, const float* Wv, const float* bv, int units, int G,
                             const double* uniforms, int64_t* ucursor, int stream_mode, double* temperature,
                             int32_t* silent_time, double default_temp, int hard, float* events,
                             float* probs_out, double* margin_out, void* stream) {
  DJ_CHECK_ARG(zpre && W0c && U0 && W1 && U1 && b1 && sp1 && Wn && bn && Wv && bv, "dj_gen_sample: NULL weight");
  DJ_CHECK_ARG(uniforms && temperature && silent_time && events, "dj_gen_sample: NULL state/output");
  DJ_CHECK_ARG(G > 0, "dj_gen_sample: G must be positive");
  DJ_CHECK_ARG(units == UN, "dj_gen_sample: units=%d unsupported (128)", units);
  DJ_CHECK_ARG(stream_mode == 0 || stream_mode == 1, "dj_gen_sample: stream_mode must be 0 or 1");
  int nclusters = G, gcount = 1;
  if (stream_mode == 0) {
    DJ_CHECK_ARG(G <= MAXG, "dj_gen_sample: reference stream order supports at most %d sequences", MAXG);
    DJ_CHECK_ARG(ucursor != nullptr, "dj_gen_sample: ucursor required in reference stream mode");
    nclusters = 1; gcount = G;
  }
  static int fast1 = -1;      // DJ_GEN_SAMPLE1=0: the barrier-synchronised kernel also for one sequence per cluster
  if (fast1 < 0) { const char* e = getenv("DJ_GEN_SAMPLE1"); fast1 = (e && atoi(e) == 0) ? 0 : 1; }
  if (gcount == 1 && fast1) {
    // clusters of 8 halve the weights per lane (no register spills) as long as every cluster is resident at once
    const int cls = (nclusters * 8 <= 120) ? 8 : 4;
    auto kernel = cls == 8 ? gen_sample1_kernel<8> : gen_sample1_kernel<4>;
    DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Gen1Smem)));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(nclusters * cls);
    cfg.blockDim = dim3(NTHR1);
    cfg.dynamicSmemBytes = sizeof(Gen1Smem);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cls; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    DJ_CUDA(cudaLaunchKernelEx(&cfg, kernel, zpre, W0c, U0, W1, U1, b1, sp1, Wn, bn, Wv, bv, uniforms, ucursor, stream_mode,
                               temperature, silent_time, default_temp, hard, events, probs_out, margin_out));
    return 0;
  }
  DJ_CUDA(cudaFuncSetAttribute((const void*)gen_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(GenSmem)));
  gen_sample_kernel<<<nclusters * CL, NTHR, sizeof(GenSmem), (cudaStream_t)stream>>>(
      zpre, W0c, U0, W1, U1, b1, sp1, Wn, bn, Wv, bv, gcount, uniforms, ucursor, stream_mode, temperature,
      silent_time, default_temp, hard, events, probs_out, margin_out);
  DJ_LAUNCH_CHECK();
  return 0;
}
