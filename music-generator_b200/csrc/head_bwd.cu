// Heads + loss (forward and its own backward), style/conv parameter gradients,
// fused Nadam.  Replaces model.py:14-20 (primary_loss), model.py:94-95,125
// (note_dense / volume_dense / Concatenate), the tf.gradients of the front end
// (model.py:56-58,77-82,113-117,141-142) and keras Nadam (model.py:152).
#include "dj_common.cuh"

namespace {

constexpr int N_ = DJ_NUM_NOTES, NU_ = DJ_NOTE_UNITS, CK_ = DJ_CONV_K, OU_ = DJ_OCTAVE_UNITS;
constexpr int PADL_ = (CK_ - 1) / 2;
constexpr int HEAD_BLOCKS = 512;

__device__ __forceinline__ float bce_term(float t, float o, float& dLdo) {
  // keras.losses.binary_crossentropy on the TF backend: clip to [1e-7, 1-1e-7],
  // logit, sigmoid-CE.  Gradient is zero where the clip is active.
  const float eps = 1e-7f;
  const float oc = fminf(fmaxf(o, eps), 1.0f - eps);
  dLdo = (o > eps && o < 1.0f - eps) ? (-t / oc + (1.0f - t) / (1.0f - oc)) : 0.0f;
  return -(t * logf(oc) + (1.0f - t) * log1pf(-oc));
}

template <int V>   // V = units / 32 values per lane
__global__ void __launch_bounds__(256) head_loss_kernel(
    const float* __restrict__ h, dj_dropout d_h, const float* __restrict__ Wn, const float* __restrict__ bn,
    const float* __restrict__ Wv, const float* __restrict__ bv, const float* __restrict__ y,
    float* __restrict__ probs, float* __restrict__ dX, float* __restrict__ partials, int64_t M, float inv_M) {
  dj_resolve(d_h);
  constexpr int UNITS = V * 32;
  constexpr int PSZ = 3 * UNITS + 4;
  __shared__ float red[8][PSZ];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float wn0[V], wn1[V], wv[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int u = lane * V + i;
    wn0[i] = Wn[u * 2]; wn1[i] = Wn[u * 2 + 1]; wv[i] = Wv[u];
  }
  const float b0 = bn[0], b1 = bn[1], b2 = bv[0];
  float g0[V], g1[V], g2[V], gb0 = 0.f, gb1 = 0.f, gb2 = 0.f, lsum = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) g0[i] = g1[i] = g2[i] = 0.f;

  // RB rows per warp iteration.  All their loads are issued before any is used; the 3 dot products of the RB rows
  // are reduced with one butterfly that leaves row r's sums in lane group r (G = 32/RB lanes), so the scalar part
  // (sigmoids, logs, loss derivatives) is computed G times per row instead of 32 times, and the reduction costs
  // ~3.4 shuffles per row instead of 15.
  constexpr int RB = (V <= 4) ? 8 : 4;
  constexpr int G = 32 / RB;
  const unsigned FULL = 0xffffffffu;
  const int grp = lane / G;
  for (int64_t row0 = ((int64_t)blockIdx.x * 8 + warp) * RB; row0 < M; row0 += (int64_t)gridDim.x * 8 * RB) {
    float x[RB][V];
    const int64_t myrow = row0 + grp;                      // the row whose scalars this lane owns
    const bool myvalid = myrow < M;
    float y0 = 0.f, y1 = 0.f, y2 = 0.f;
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int64_t row = (row0 + rb < M) ? row0 + rb : M - 1;
#pragma unroll
      for (int i4 = 0; i4 < V; i4 += 4) {
        const float4 hv = __ldg(reinterpret_cast<const float4*>(h + row * UNITS + lane * V + i4));
        x[rb][i4] = hv.x; x[rb][i4 + 1] = hv.y; x[rb][i4 + 2] = hv.z; x[rb][i4 + 3] = hv.w;
      }
    }
    if (y != nullptr && myvalid) { y0 = __ldg(y + myrow * 3); y1 = __ldg(y + myrow * 3 + 1); y2 = __ldg(y + myrow * 3 + 2); }
    float s[RB][3];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int64_t row = (row0 + rb < M) ? row0 + rb : M - 1;
#pragma unroll
      for (int i4 = 0; i4 < V; i4 += 4) {
        float m[4];
        dj_dropmul4(d_h, (uint32_t)(row * UNITS + lane * V + i4), m);
        x[rb][i4] *= m[0]; x[rb][i4 + 1] *= m[1]; x[rb][i4 + 2] *= m[2]; x[rb][i4 + 3] *= m[3];
      }
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s0 = fmaf(x[rb][i], wn0[i], s0); s1 = fmaf(x[rb][i], wn1[i], s1); s2 = fmaf(x[rb][i], wv[i], s2);
      }
      s[rb][0] = s0; s[rb][1] = s1; s[rb][2] = s2;
    }
    // butterfly: each step halves the rows a lane carries (upper lanes keep the upper half)
#pragma unroll
    for (int half = RB / 2, off = 16; half >= 1; half >>= 1, off >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int j = 0; j < half; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float keep = upper ? s[j + half][k] : s[j][k];
          const float send = upper ? s[j][k] : s[j + half][k];
          s[j][k] = keep + __shfl_xor_sync(FULL, send, off);
        }
    }
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1)
#pragma unroll
      for (int k = 0; k < 3; ++k) s[0][k] += __shfl_xor_sync(FULL, s[0][k], off);
    // scalars of row `myrow` (identical in the G lanes of the group)
    const float p0 = dj_sigmoid(s[0][0] + b0), p1 = dj_sigmoid(s[0][1] + b1), vol = s[0][2] + b2;
    if (myvalid && (lane % G) == 0) { probs[myrow * 3] = p0; probs[myrow * 3 + 1] = p1; probs[myrow * 3 + 2] = vol; }
    if (y != nullptr) {
      float da0 = 0.f, da1 = 0.f, da2 = 0.f;
      if (myvalid) {
        float dl0, dlq;
        const float l0 = bce_term(y0, p0, dl0);
        const float q = y0 * p1 + (1.f - y0) * y1;
        const float l1 = bce_term(y1, q, dlq);
        const float d = y2 - (y0 * vol + (1.f - y0) * y2);
        da0 = dl0 * p0 * (1.f - p0) * inv_M;
        da1 = dlq * y0 * p1 * (1.f - p1) * inv_M;
        da2 = -2.f * d * y0 * inv_M;
        if ((lane % G) == 0) { lsum += l0 + l1 + d * d; gb0 += da0; gb1 += da1; gb2 += da2; }
      }
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        const float a0 = __shfl_sync(FULL, da0, rb * G), a1 = __shfl_sync(FULL, da1, rb * G),
                    a2 = __shfl_sync(FULL, da2, rb * G);
        if (row0 + rb >= M) break;                       // warp-uniform; a* of invalid rows are zero anyway
#pragma unroll
        for (int i = 0; i < V; ++i) {
          g0[i] = fmaf(x[rb][i], a0, g0[i]); g1[i] = fmaf(x[rb][i], a1, g1[i]); g2[i] = fmaf(x[rb][i], a2, g2[i]);
        }
        if (dX != nullptr) {
#pragma unroll
          for (int i4 = 0; i4 < V; i4 += 4) {
            float4 o;
            o.x = a0 * wn0[i4] + a1 * wn1[i4] + a2 * wv[i4];
            o.y = a0 * wn0[i4 + 1] + a1 * wn1[i4 + 1] + a2 * wv[i4 + 1];
            o.z = a0 * wn0[i4 + 2] + a1 * wn1[i4 + 2] + a2 * wv[i4 + 2];
            o.w = a0 * wn0[i4 + 3] + a1 * wn1[i4 + 3] + a2 * wv[i4 + 3];
            *reinterpret_cast<float4*>(dX + (row0 + rb) * UNITS + lane * V + i4) = o;
          }
        }
      }
    }
  }
  if (y == nullptr || partials == nullptr) return;
  // the loss and bias-gradient terms were accumulated by the leader lane of each row group
  lsum = dj_warp_sum(lsum); gb0 = dj_warp_sum(gb0); gb1 = dj_warp_sum(gb1); gb2 = dj_warp_sum(gb2);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int u = lane * V + i;
    red[warp][u] = g0[i]; red[warp][UNITS + u] = g1[i]; red[warp][2 * UNITS + u] = g2[i];
  }
  if (lane == 0) {
    red[warp][3 * UNITS] = gb0; red[warp][3 * UNITS + 1] = gb1; red[warp][3 * UNITS + 2] = gb2;
    red[warp][3 * UNITS + 3] = lsum * inv_M;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < PSZ; e += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][e];
    partials[(int64_t)blockIdx.x * PSZ + e] = s;
  }
}

__global__ void head_finalize_kernel(const float* __restrict__ partials, int units, float* loss_out,
                                     float* dWn, float* dbn, float* dWv, float* dbv) {
  const int PSZ = 3 * units + 4;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= PSZ) return;
  float s = 0.f;
  for (int b = 0; b < HEAD_BLOCKS; ++b) s += partials[(int64_t)b * PSZ + e];
  if (e < units) dWn[e * 2] = s;
  else if (e < 2 * units) dWn[(e - units) * 2 + 1] = s;
  else if (e < 3 * units) dWv[e - 2 * units] = s;
  else if (e < 3 * units + 2) dbn[e - 3 * units] = s;
  else if (e == 3 * units + 2) dbv[0] = s;
  else loss_out[0] = s;
}

__global__ void __launch_bounds__(128) style_bwd_reduce_kernel(const float* __restrict__ dA, int64_t ldA, int F,
                                                               const float* __restrict__ sp, dj_dropout d_sp,
                                                               float* __restrict__ ds) {
  dj_resolve(d_sp);
  const int ld4 = (F + 3) & ~3;
  const int64_t bt = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += 128) {
    float acc = 0.f;
    for (int n = 0; n < N_; ++n) {
      const int64_t row = bt * N_ + n;
      acc = fmaf(dA[row * ldA + f], dj_dropmul(d_sp, (uint32_t)(row * ld4 + f)), acc);
    }
    const float s = sp[bt * F + f];
    ds[bt * F + f] = acc * (1.f - s * s);
  }
}

// grid (column blocks of 32, row slices): each block sums its slice, one atomic per column per block
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t ldx, int64_t R, int C,
                                                     float* __restrict__ out, float* __restrict__ part) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5, c = blockIdx.x * 32 + cl;
  const int64_t per = (R + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * per, r1 = (r0 + per < R) ? r0 + per : R;
  float s = 0.f;
  if (c < C)
    for (int64_t r = r0 + rl; r < r1; r += 8) s += X[r * ldx + c];
  red[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cl];
    if (part != nullptr) part[(int64_t)blockIdx.y * C + c] = t;      // deterministic mode: added in slice order later
    else atomicAdd(out + c, t);
  }
}

// conv weight gradient: recompute tanh(conv), form dpre, accumulate x^T.dpre per block.
// Register tiles of 4 output channels keep the shared-memory loads per FMA low (the kernel is FMA work fed from
// shared memory): a thread is (og = 4 adjacent channels, r = one of 16 groups).
//   recompute: r = 3 adjacent notes          -> per tap 1 LDS.128 of weights + <= 3 LDS of x for 12 FMAs
//   x^T.dpre : r = (q: 18 of the 72 taps, ng: 12 of the 48 notes) -> per note 1 LDS.128 of dpre + 18 LDS of x for 72 FMAs
// The 72 partial sums per thread live in registers over all (b,t) of the block; at the end the four note groups are
// folded in shared memory so that a block issues one atomic per weight.
__global__ void __launch_bounds__(256, 2) conv_bwd_kernel(
    const float* __restrict__ notes_in, int64_t notes_bstride, int B, int T, const float* __restrict__ Wc,
    const float* __restrict__ bc, dj_dropout d_notes, dj_dropout d_conv, const float* __restrict__ dA0,
    int64_t ldA, float* __restrict__ dWc, float* __restrict__ dbc, float* __restrict__ part) {
  dj_resolve(d_notes); dj_resolve(d_conv);
  constexpr int KC = CK_ * NU_;          // 72 (tap, channel-in) pairs
  constexpr int KPT = KC / 4;            // 18 of them per q
  constexpr int NPT = N_ / 4;            // 12 notes per ng
  static_assert(OU_ == 64 && N_ == 48 && KC == 72, "conv_bwd_kernel tiling");
  __shared__ __align__(16) float Wc_s[KC * OU_];     // reused as the fold buffer at the end
  __shared__ float bc_s[OU_];
  __shared__ float xs[(N_ + CK_ - 1) * NU_];
  __shared__ __align__(16) float dpre[N_][OU_];
  const int tid = threadIdx.x, BT = B * T;
  const int og = tid & 15, r = tid >> 4, q = r & 3, ng = r >> 2;
  for (int i = tid; i < KC * OU_; i += 256) Wc_s[i] = Wc[i];
  if (tid < OU_) bc_s[tid] = bc[tid];
  for (int i = tid; i < (N_ + CK_ - 1) * NU_; i += 256) xs[i] = 0.f;
  // packed fp32 FMAs (FFMA2): one instruction updates two adjacent channels; each lane is an IEEE fma
  uint64_t gw2[KPT][2];
  float gb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < KPT; ++i) gw2[i][0] = gw2[i][1] = pack2(0.f, 0.f);
  __syncthreads();
  for (int bt = blockIdx.x; bt < BT; bt += gridDim.x) {
    const int b = bt / T, t = bt % T;
    if (tid < N_ * NU_) {
      const int n = tid / NU_, c = tid % NU_;
      xs[(n + PADL_) * NU_ + c] = notes_in[(int64_t)b * notes_bstride + (int64_t)t * (N_ * NU_) + tid] *
                                  dj_dropmul(d_notes, (uint32_t)(bt * N_ + n) * 4u + c);
    }
    __syncthreads();
    {
      // notes 3r .. 3r+2, channels 4og .. 4og+3; same summation order per output as the forward kernel
      uint64_t acc2[3][2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc2[j][0] = pack2(bc_s[4 * og], bc_s[4 * og + 1]);
        acc2[j][1] = pack2(bc_s[4 * og + 2], bc_s[4 * og + 3]);
      }
      const float* xw = xs + 3 * r * NU_;
      // upstream gradient of these 12 outputs: issued before the FMA loop so the loads are in flight under it
      float up[3][4];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int64_t row = (int64_t)bt * N_ + 3 * r + j;
#pragma unroll
        for (int c = 0; c < 4; ++c) up[j][c] = dA0[row * ldA + 14 + 4 * og + c];
      }
#pragma unroll 8
      for (int kc = 0; kc < KC; ++kc) {
        const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(Wc_s + kc * OU_ + 4 * og);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float x = xw[j * NU_ + kc];
          const uint64_t xx = pack2(x, x);
          acc2[j][0] = ffma2(xx, w.x, acc2[j][0]);
          acc2[j][1] = ffma2(xx, w.y, acc2[j][1]);
        }
      }
      float acc[3][4];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        unpack2(acc2[j][0], acc[j][0], acc[j][1]);
        unpack2(acc2[j][1], acc[j][2], acc[j][3]);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int n = 3 * r + j;
        const int64_t row = (int64_t)bt * N_ + n;
        float4 d;
        float* dv = reinterpret_cast<float*>(&d);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int o = 4 * og + c;
          const float a = tanhf(acc[j][c]);
          dv[c] = up[j][c] * dj_dropmul(d_conv, (uint32_t)(row * OU_ + o)) * (1.f - a * a);
        }
        *reinterpret_cast<float4*>(&dpre[n][4 * og]) = d;
      }
    }
    __syncthreads();
#pragma unroll 2
    for (int nn = 0; nn < NPT; ++nn) {
      const int n = ng * NPT + nn;
      const ulonglong2 d2 = *reinterpret_cast<const ulonglong2*>(&dpre[n][4 * og]);
      if (q == 0) {
        float dx, dy, dz, dw;
        unpack2(d2.x, dx, dy); unpack2(d2.y, dz, dw);
        gb[0] += dx; gb[1] += dy; gb[2] += dz; gb[3] += dw;
      }
      const float* xr = xs + n * NU_ + q * KPT;
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        const float x = xr[i];
        const uint64_t xx = pack2(x, x);
        gw2[i][0] = ffma2(xx, d2.x, gw2[i][0]);
        gw2[i][1] = ffma2(xx, d2.y, gw2[i][1]);
      }
    }
    __syncthreads();
  }
  // fold the four note groups (one after the other: within a round every (tap, channel) has one owner)
  float* red = Wc_s;
  for (int i = tid; i < KC * OU_; i += 256) red[i] = 0.f;
  if (tid < OU_) bc_s[tid] = 0.f;
  __syncthreads();
  for (int g = 0; g < 4; ++g) {
    if (ng == g) {
#pragma unroll
      for (int i = 0; i < KPT; ++i) {
        float4* p = reinterpret_cast<float4*>(red + (q * KPT + i) * OU_ + 4 * og);
        float4 v = *p;
        float g0, g1, g2, g3;
        unpack2(gw2[i][0], g0, g1); unpack2(gw2[i][1], g2, g3);
        v.x += g0; v.y += g1; v.z += g2; v.w += g3;
        *p = v;
      }
      if (q == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) bc_s[4 * og + c] += gb[c];
      }
    }
    __syncthreads();
  }
  if (part != nullptr) {   // deterministic mode: this block's partial [72*64 weights | 64 biases], added in block order later
    float* dst = part + (size_t)blockIdx.x * (KC * OU_ + OU_);
    for (int i = tid; i < KC * OU_; i += 256) dst[i] = red[i];
    if (tid < OU_) dst[KC * OU_ + tid] = bc_s[tid];
    return;
  }
  for (int i = tid; i < KC * OU_; i += 256) atomicAdd(dWc + i, red[i]);
  if (tid < OU_) atomicAdd(dbc + tid, bc_s[tid]);
}

__global__ void __launch_bounds__(256) nadam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                    float gscale, float lr, float beta1, float beta2, float eps,
                                                    float mu_t, float mu_t1, float inv_1m_ms_new,
                                                    float inv_1m_ms_next, float inv_bias2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float mi = m[i], vi = v[i];
    p[i] = dj_nadam_one(p[i], __fmul_rn(g[i], gscale), mi, vi, lr, beta1, beta2, eps, mu_t, mu_t1, inv_1m_ms_new,
                        inv_1m_ms_next, inv_bias2);
    m[i] = mi;
    v[i] = vi;
  }
}

// the ten per-step scalars from device memory (a replayed CUDA graph gets new ones every step)
__global__ void __launch_bounds__(256) nadam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const float* __restrict__ sc) {
  const float gscale = sc[0], lr = sc[1], beta1 = sc[2], beta2 = sc[3], eps = sc[4], mu_t = sc[5], mu_t1 = sc[6],
              i1 = sc[7], i2 = sc[8], ib = sc[9];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float mi = m[i], vi = v[i];
    p[i] = dj_nadam_one(p[i], __fmul_rn(g[i], gscale), mi, vi, lr, beta1, beta2, eps, mu_t, mu_t1, i1, i2, ib);
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace

extern "C" int dj_nadam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* sc, void* stream) {
  DJ_CHECK_ARG(p && g && m && v && sc && n > 0, "dj_nadam_step_dev: bad arguments");
  int64_t blocks = (n + 255) / 256;
  if (blocks > dj_num_sms() * 8) blocks = dj_num_sms() * 8;
  nadam_dev_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sc);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int64_t dj_head_partials_size(int units) { return (int64_t)HEAD_BLOCKS * (3 * units + 4); }

extern "C" int dj_head_loss(const float* h, int units, dj_dropout d_h, const float* Wn, const float* bn,
                            const float* Wv, const float* bv, const float* y_true, float* probs, float* dX,
                            float* partials, int64_t M, void* stream) {
  DJ_CHECK_ARG(h && Wn && bn && Wv && bv && probs, "dj_head_loss: NULL pointer");
  DJ_CHECK_ARG(M > 0 && M * units < (int64_t)4294967296LL, "dj_head_loss: bad M");
  DJ_CHECK_ARG(y_true == nullptr || partials != nullptr, "dj_head_loss: partials required with y_true");
  const float inv_M = 1.0f / (float)M;
  cudaStream_t st = (cudaStream_t)stream;
  if (units == 128)
    head_loss_kernel<4><<<HEAD_BLOCKS, 256, 0, st>>>(h, d_h, Wn, bn, Wv, bv, y_true, probs, dX, partials, M, inv_M);
  else if (units == 256)
    head_loss_kernel<8><<<HEAD_BLOCKS, 256, 0, st>>>(h, d_h, Wn, bn, Wv, bv, y_true, probs, dX, partials, M, inv_M);
  else DJ_CHECK_ARG(false, "dj_head_loss: units=%d unsupported (128 or 256)", units);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_head_finalize(const float* partials, int units, float* loss_out, float* dWn, float* dbn,
                                float* dWv, float* dbv, void* stream) {
  DJ_CHECK_ARG(partials && loss_out && dWn && dbn && dWv && dbv, "dj_head_finalize: NULL pointer");
  const int psz = 3 * units + 4;
  head_finalize_kernel<<<(psz + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partials, units, loss_out, dWn, dbn, dWv, dbv);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_style_bwd_reduce(const float* dA, int64_t ldA, int F, const float* sp, dj_dropout d_sp,
                                   int BT, float* ds, void* stream) {
  DJ_CHECK_ARG(dA && sp && ds && F > 0 && BT > 0 && ldA >= F, "dj_style_bwd_reduce: bad arguments");
  style_bwd_reduce_kernel<<<BT, 128, 0, (cudaStream_t)stream>>>(dA, ldA, F, sp, d_sp, ds);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_colsum(const float* X, int64_t ldx, int64_t R, int C, float* out, int accumulate, void* stream) {
  DJ_CHECK_ARG(X && out && R > 0 && C > 0 && ldx >= C, "dj_colsum: bad arguments");
  if (!accumulate) DJ_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)C, (cudaStream_t)stream));
  int slices = (int)((R + 127) / 128);
  if (slices > 64) slices = 64;
  if (slices < 1) slices = 1;
  float* part = nullptr;
  if (slices > 1 && dj_reduce_workspace(stream, (int64_t)slices * C, &part)) return -1;
  colsum_kernel<<<dim3((C + 31) / 32, slices), 256, 0, (cudaStream_t)stream>>>(X, ldx, R, C, out, part);
  DJ_LAUNCH_CHECK();
  if (part != nullptr) return dj_ordered_reduce(part, slices, C, 1, C, C, out, C, stream);
  return 0;
}

extern "C" int dj_conv_bwd(const float* notes_in, int64_t notes_bstride, int B, int T, const float* Wc,
                           const float* bc, dj_dropout d_notes, dj_dropout d_conv, const float* dA0, int64_t ldA,
                           float* dWc, float* dbc, void* stream) {
  DJ_CHECK_ARG(notes_in && Wc && bc && dA0 && dWc && dbc, "dj_conv_bwd: NULL pointer");
  DJ_CHECK_ARG(B > 0 && T > 0 && ldA >= DJ_FEAT0, "dj_conv_bwd: bad sizes");
  int grid = dj_num_sms() * 2;   // 31 KB smem, 256 threads x <=128 registers: 2 resident blocks overlap the barriers
  if (grid > B * T) grid = B * T;
  constexpr int PSZ = CK_ * NU_ * OU_ + OU_;
  float* part = nullptr;
  if (dj_reduce_workspace(stream, (int64_t)grid * PSZ, &part)) return -1;
  conv_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(notes_in, notes_bstride, B, T, Wc, bc, d_notes, d_conv,
                                                          dA0, ldA, dWc, dbc, part);
  DJ_LAUNCH_CHECK();
  if (part != nullptr) {
    int rc = dj_ordered_reduce(part, grid, PSZ, 1, CK_ * NU_ * OU_, PSZ, dWc, CK_ * NU_ * OU_, stream);
    if (rc) return rc;
    return dj_ordered_reduce(part + CK_ * NU_ * OU_, grid, PSZ, 1, OU_, PSZ, dbc, OU_, stream);
  }
  return 0;
}

extern "C" int dj_nadam_step(float* p, const float* g, float* m, float* v, int64_t n, float gscale, float lr,
                             float beta1, float beta2, float eps, float mu_t, float mu_t1, float m_sched_new,
                             float m_sched_next, float bias2, void* stream) {
  DJ_CHECK_ARG(p && g && m && v && n > 0, "dj_nadam_step: bad arguments");
  DJ_CHECK_ARG(m_sched_new < 1.f && m_sched_next < 1.f && bias2 > 0.f, "dj_nadam_step: bad schedule scalars");
  int64_t blocks = (n + 255) / 256;
  const int maxb = dj_num_sms() * 8;
  if (blocks > maxb) blocks = maxb;
  nadam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, gscale, lr, beta1, beta2, eps, mu_t,
                                                              mu_t1, 1.f / (1.f - m_sched_new),
                                                              1.f / (1.f - m_sched_next), 1.f / bias2);
  DJ_LAUNCH_CHECK();
  return 0;
}
