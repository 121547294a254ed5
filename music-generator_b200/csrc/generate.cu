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
  DJ_CUDA(cudaFuncSetAttribute((const void*)gen_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(GenSmem)));
  gen_sample_kernel<<<nclusters * CL, NTHR, sizeof(GenSmem), (cudaStream_t)stream>>>(
      zpre, W0c, U0, W1, U1, b1, sp1, Wn, bn, Wv, bv, gcount, uniforms, ucursor, stream_mode, temperature,
      silent_time, default_temp, hard, events, probs_out, margin_out);
  DJ_LAUNCH_CHECK();
  return 0;
}
