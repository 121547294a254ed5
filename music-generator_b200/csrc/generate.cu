// Persistent note-by-note sampler for generation.
// Replaces the inner loop of generate.py:112-121 (48 x note_model.predict + the
// NumPy MusicGeneration.choose / end_time, generate.py:47-79): one launch walks
// all 48 notes of a timestep carrying the note-axis LSTM state (mathematically
// identical to the reference's 48 full re-evaluations because the note LSTM is
// causal in n), applies the temperature transform in float32 as
// NumPy does, compares float64 uniforms with `<=`, and draws the replay uniform
// only when the note is played.
#include "dj_common.cuh"

namespace {

constexpr int N_ = DJ_NUM_NOTES;
constexpr int MAXG = 16;   // sequences one CTA can walk serially (reference stream order)

template <int UN>
__global__ void __launch_bounds__(4 * UN) gen_sample_kernel(
    const float* __restrict__ zpre, const float* __restrict__ W0c, const float* __restrict__ U0,
    const float* __restrict__ W1, const float* __restrict__ U1, const float* __restrict__ b1,
    const float* __restrict__ sp1, const float* __restrict__ Wn, const float* __restrict__ bn,
    const float* __restrict__ Wv, const float* __restrict__ bv, int gcount, const double* __restrict__ uniforms,
    int64_t* ucursor, int stream_mode, double* temperature, int32_t* silent_time, double default_temp, int hard,
    float* __restrict__ events, float* __restrict__ probs_out, double* __restrict__ margin_out) {
  constexpr int G4 = 4 * UN;
  __shared__ float zbuf[G4];
  __shared__ float h0[MAXG][UN], h1[MAXG][UN], c0[MAXG][UN], c1[MAXG][UN];
  __shared__ float x1[UN];
  __shared__ float prev[MAXG][3];
  __shared__ float head[3];
  __shared__ int played_any[MAXG];
  __shared__ double margin[MAXG];
  __shared__ long long cursor;

  const int j = threadIdx.x;
  const int g_base = blockIdx.x * gcount;
  for (int i = j; i < MAXG * UN; i += G4) {
    (&h0[0][0])[i] = 0.f; (&h1[0][0])[i] = 0.f; (&c0[0][0])[i] = 0.f; (&c1[0][0])[i] = 0.f;
  }
  if (j < MAXG) {
    prev[j][0] = prev[j][1] = prev[j][2] = 0.f;
    played_any[j] = 0;
    margin[j] = 1e300;
  }
  if (j == 0) cursor = (stream_mode == 0 && ucursor != nullptr) ? *ucursor : 0;
  const float w0c0 = W0c[j], w0c1 = W0c[G4 + j], w0c2 = W0c[2 * G4 + j];
  const float bias1 = b1[j];
  __syncthreads();

  for (int n = 0; n < N_; ++n) {
    for (int gl = 0; gl < gcount; ++gl) {
      const int g = g_base + gl;
      // ---- note layer 0: z = zpre + chosen_{n-1}.W0[Ut:Ut+3] + h0.U0
      {
        float acc = zpre[((int64_t)g * N_ + n) * G4 + j];
        acc = fmaf(prev[gl][0], w0c0, acc);
        acc = fmaf(prev[gl][1], w0c1, acc);
        acc = fmaf(prev[gl][2], w0c2, acc);
        const float* hp = h0[gl];
#pragma unroll 16
        for (int k = 0; k < UN; ++k) acc = fmaf(hp[k], U0[k * G4 + j], acc);
        zbuf[j] = acc;
      }
      __syncthreads();
      if (j < UN) {
        // gate columns are interleaved: col = 4*unit + gate (i,f,c,o)
        const float gi = dj_gate_act(zbuf[4 * j], hard), gf = dj_gate_act(zbuf[4 * j + 1], hard);
        const float gg = tanhf(zbuf[4 * j + 2]), go = dj_gate_act(zbuf[4 * j + 3], hard);
        const float cn = fmaf(gf, c0[gl][j], gi * gg);
        c0[gl][j] = cn;
        const float hn = go * tanhf(cn);
        h0[gl][j] = hn;
        x1[j] = hn + sp1[(int64_t)g * UN + j];   // model.py:113-117 at inference: x + tanh(Dense(style))
      }
      __syncthreads();
      // ---- note layer 1
      {
        float acc = bias1;
        const float* hp = h1[gl];
#pragma unroll 16
        for (int k = 0; k < UN; ++k) acc = fmaf(x1[k], W1[k * G4 + j], acc);
#pragma unroll 16
        for (int k = 0; k < UN; ++k) acc = fmaf(hp[k], U1[k * G4 + j], acc);
        zbuf[j] = acc;
      }
      __syncthreads();
      if (j < UN) {
        const float gi = dj_gate_act(zbuf[4 * j], hard), gf = dj_gate_act(zbuf[4 * j + 1], hard);
        const float gg = tanhf(zbuf[4 * j + 2]), go = dj_gate_act(zbuf[4 * j + 3], hard);
        const float cn = fmaf(gf, c1[gl][j], gi * gg);
        c1[gl][j] = cn;
        h1[gl][j] = go * tanhf(cn);
      }
      __syncthreads();
      // ---- heads (model.py:94-95): warps 0..2 each reduce one output
      if (j < 96) {
        const int o = j >> 5, lane = j & 31;
        float s = 0.f;
        for (int k = lane; k < UN; k += 32) {
          const float w = (o < 2) ? Wn[k * 2 + o] : Wv[k];
          s = fmaf(h1[gl][k], w, s);
        }
        s = dj_warp_sum(s);
        if (lane == 0) head[o] = s + ((o < 2) ? bn[o] : bv[0]);
      }
      __syncthreads();
      if (j == 0) {
        float p0 = dj_sigmoid(head[0]), p1 = dj_sigmoid(head[1]);
        const float vol = head[2];
        if (probs_out != nullptr) {
          float* po = probs_out + ((int64_t)g * N_ + n) * 3;
          po[0] = p0; po[1] = p1; po[2] = vol;
        }
        const double temp = temperature[g];
        if (temp != 1.0) {   // generate.py:81-91, float32 arithmetic like NumPy on a float32 array
          const float tf = (float)temp;
          const float xa = -logf(1.0f / p0 - 1.0f), xb = -logf(1.0f / p1 - 1.0f);
          p0 = 1.0f / (1.0f + expf(-xa / tf));
          p1 = 1.0f / (1.0f + expf(-xb / tf));
        }
        double u1, u2;
        const double* ui = uniforms + ((int64_t)g * N_ + n) * 2;
        if (stream_mode == 0) u1 = uniforms[cursor++]; else u1 = ui[0];
        float e0 = 0.f, e1 = 0.f, e2 = 0.f;
        double mg = fabs(u1 - (double)p0);
        if (u1 <= (double)p0) {   // generate.py:52
          e0 = 1.f; e2 = vol;
          if (stream_mode == 0) u2 = uniforms[cursor++]; else u2 = ui[1];
          mg = fmin(mg, fabs(u2 - (double)p1));
          if (u2 <= (double)p1) e1 = 1.f;   // generate.py:57
          played_any[gl] = 1;
        }
        margin[gl] = fmin(margin[gl], mg);
        prev[gl][0] = e0; prev[gl][1] = e1; prev[gl][2] = e2;
        float* ev = events + ((int64_t)g * N_ + n) * 3;
        ev[0] = e0; ev[1] = e1; ev[2] = e2;
      }
      __syncthreads();
    }
  }
  // ---- end_time (generate.py:60-79): silence raises the temperature
  if (j < gcount) {
    const int g = g_base + j;
    // np.count_nonzero(next_note) == 0  <=>  nothing played (a played note sets channel 0 to 1)
    if (!played_any[j]) {
      const int st = silent_time[g] + 1;
      silent_time[g] = st;
      if (st >= DJ_BEAT) temperature[g] = temperature[g] + 0.1;
    } else {
      silent_time[g] = 0;
      temperature[g] = default_temp;
    }
    if (margin_out != nullptr) margin_out[g] = fmin(margin_out[g], margin[j]);
  }
  if (j == 0 && stream_mode == 0 && ucursor != nullptr) *ucursor = cursor;
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
  DJ_CHECK_ARG(units == 128, "dj_gen_sample: units=%d unsupported (128)", units);
  DJ_CHECK_ARG(stream_mode == 0 || stream_mode == 1, "dj_gen_sample: stream_mode must be 0 or 1");
  int grid = G, gcount = 1;
  if (stream_mode == 0) {
    DJ_CHECK_ARG(G <= MAXG, "dj_gen_sample: reference stream order supports at most %d sequences", MAXG);
    DJ_CHECK_ARG(ucursor != nullptr, "dj_gen_sample: ucursor required in reference stream mode");
    grid = 1; gcount = G;
  }
  gen_sample_kernel<128><<<grid, 512, 0, (cudaStream_t)stream>>>(
      zpre, W0c, U0, W1, U1, b1, sp1, Wn, bn, Wv, bv, gcount, uniforms, ucursor, stream_mode, temperature,
      silent_time, default_temp, hard, events, probs_out, margin_out);
  DJ_LAUNCH_CHECK();
  return 0;
}
