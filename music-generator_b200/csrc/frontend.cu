// Note-feature front end and inter-layer glue (forward).
// Replaces model.py:22-49,56-82,101-117,136-142 of the reference (Keras Lambda /
// Conv1D / Dense / Dropout / Concatenate / Permute layers) with fused kernels
// that write the LSTM gate-GEMM A operands directly in canonical row order.
#include <cuda_fp16.h>

#include "dj_common.cuh"

namespace {

constexpr int N_ = DJ_NUM_NOTES;      // 48
constexpr int NU_ = DJ_NOTE_UNITS;    // 3
constexpr int CK_ = DJ_CONV_K;        // 24
constexpr int OU_ = DJ_OCTAVE_UNITS;  // 64
constexpr int F0_ = DJ_FEAT0;         // 94
constexpr int F0P_ = 96;              // roundup4(94)
constexpr int PADL_ = (CK_ - 1) / 2;  // TF 'SAME', even kernel: 11 left / 12 right

__device__ __forceinline__ void store4(float* p, const float v[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float v[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ void store4(__half* p, const float v[4]) {
  const __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void store4_lo(__half* p, const float v[4]) {   // half hi + lo: ~22 mantissa bits
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = v[j] - __half2float(__float2half_rn(v[j]));
  store4(p, r);
}
// residual of the bf16 rounding, itself rounded to bf16: hi + lo carries 16 mantissa bits, which is what the
// 3-pass split gate GEMM (dj_gate_gemm_16) multiplies
__device__ __forceinline__ void store4_lo(__nv_bfloat16* p, const float v[4]) {
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = v[j] - __bfloat162float(__float2bfloat16_rn(v[j]));
  store4(p, r);
}
__device__ __forceinline__ void store4_lo(float*, const float[4]) {}   // fp32 operands have no residual

// ---------------------------------------------------------------------------
// style embedding + the four tanh style projections; 8 (b,t) rows per block
// ---------------------------------------------------------------------------
struct StyleProj {
  const float* W[4];
  const float* b[4];
  float* out[4];
  int F[4];
};

__global__ void __launch_bounds__(256) style_fwd_kernel(const float* __restrict__ style_in,
                                                        int64_t bstride, int64_t tstride, int ns,
                                                        int B, int T, const float* __restrict__ Ws,
                                                        const float* __restrict__ bs, int n_proj,
                                                        StyleProj pr, float* __restrict__ emb) {
  constexpr int R = 8, SU = DJ_STYLE_UNITS;
  __shared__ float sin_[R][32];
  __shared__ float embs[R][SU];
  const int BT = B * T, row0 = blockIdx.x * R, tid = threadIdx.x;
  for (int i = tid; i < R * 32; i += 256) {
    int r = i / 32, s = i % 32, row = row0 + r;
    float v = 0.f;
    if (row < BT && s < ns) v = style_in[(int64_t)(row / T) * bstride + (int64_t)(row % T) * tstride + s];
    sin_[r][s] = v;
  }
  __syncthreads();
  for (int i = tid; i < R * SU; i += 256) {
    int r = i / SU, o = i % SU;
    float acc = bs[o];
    for (int s = 0; s < ns; ++s) acc = fmaf(sin_[r][s], Ws[s * SU + o], acc);
    embs[r][o] = acc;
    if (row0 + r < BT) emb[(int64_t)(row0 + r) * SU + o] = acc;
  }
  __syncthreads();
  for (int l = 0; l < n_proj; ++l) {
    const int F = pr.F[l];
    const float* __restrict__ W = pr.W[l];
    for (int j = tid; j < F; j += 256) {
      float acc[R];
      const float bj = pr.b[l][j];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = bj;
      for (int k = 0; k < SU; ++k) {
        const float w = W[k * F + j];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(embs[r][k], w, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (row0 + r < BT) pr.out[l][(int64_t)(row0 + r) * F + j] = tanhf(acc[r]);
    }
  }
}

// ---------------------------------------------------------------------------
// front end: one (b,t) per loop iteration of a persistent block
// ---------------------------------------------------------------------------
template <typename TA>
__global__ void __launch_bounds__(256) frontend_fwd_kernel(
    const float* __restrict__ notes_in, int64_t notes_bstride, const float* __restrict__ beat_in,
    int64_t beat_bstride, int B, int T, const float* __restrict__ Wc, const float* __restrict__ bc,
    const float* __restrict__ sp0, dj_dropout d_notes, dj_dropout d_beat, dj_dropout d_conv,
    dj_dropout d_sp, TA* __restrict__ A0, TA* __restrict__ A0lo, int ldA) {
  dj_resolve(d_notes); dj_resolve(d_beat); dj_resolve(d_conv); dj_resolve(d_sp);
  __shared__ __align__(16) float Wc_s[CK_ * NU_ * OU_];
  __shared__ float bc_s[OU_];
  __shared__ float xs[(N_ + CK_ - 1) * NU_];
  __shared__ __align__(16) float tile[N_][F0P_];
  __shared__ float bins_s[N_];
  __shared__ float beat_s[DJ_BEAT];
  const int tid = threadIdx.x, BT = B * T;
  for (int i = tid; i < CK_ * NU_ * OU_; i += 256) Wc_s[i] = Wc[i];
  if (tid < OU_) bc_s[tid] = bc[tid];
  for (int i = tid; i < (N_ + CK_ - 1) * NU_; i += 256) xs[i] = 0.f;   // halo stays zero
  __syncthreads();

  for (int bt = blockIdx.x; bt < BT; bt += gridDim.x) {
    const int b = bt / T, t = bt % T;
    const float* nrow = notes_in + (int64_t)b * notes_bstride + (int64_t)t * (N_ * NU_);
    if (tid < N_ * NU_) {
      const int n = tid / NU_, c = tid % NU_;
      xs[(n + PADL_) * NU_ + c] = nrow[tid] * dj_dropmul(d_notes, (uint32_t)(bt * N_ + n) * 4u + c);
    } else if (tid < N_ * NU_ + DJ_BEAT) {
      const int j = tid - N_ * NU_;
      beat_s[j] = beat_in[(int64_t)b * beat_bstride + (int64_t)t * DJ_BEAT + j] *
                  dj_dropmul(d_beat, (uint32_t)bt * DJ_BEAT + j);
    } else if (tid < N_ * NU_ + DJ_BEAT + N_) {
      // pitch_bins_f (model.py:43-49): the raw reshape of the [48,B,T] tensor
      const int n = tid - N_ * NU_ - DJ_BEAT;
      const int64_t flat = (int64_t)bt * N_ + n;
      const int p = (int)(flat / BT), r = (int)(flat % BT);
      const int pc = p % DJ_OCTAVE;
      const float* src = notes_in + (int64_t)(r / T) * notes_bstride + (int64_t)(r % T) * (N_ * NU_);
      float s = 0.f;
#pragma unroll
      for (int o = 0; o < N_ / DJ_OCTAVE; ++o) {
        const int nn = o * DJ_OCTAVE + pc;
        s += src[nn * NU_] * dj_dropmul(d_notes, (uint32_t)(r * N_ + nn) * 4u);
      }
      bins_s[n] = s;
    }
    __syncthreads();
    {  // octave convolution: thread = (4 adjacent out channels, 3 adjacent notes): per tap one LDS.128 of weights
       // and three of x feed 12 FMAs; every output sums bias + taps in ascending (k, c) order
      const int og = tid & 15, r = tid >> 4;
      // packed fp32 FMAs (FFMA2): one instruction updates two adjacent channels; each lane is an IEEE fma
      uint64_t acc2[3][2];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        acc2[j][0] = pack2(bc_s[4 * og], bc_s[4 * og + 1]);
        acc2[j][1] = pack2(bc_s[4 * og + 2], bc_s[4 * og + 3]);
      }
      const float* xw = xs + 3 * r * NU_;
#pragma unroll 8
      for (int kc = 0; kc < CK_ * NU_; ++kc) {
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
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int o = 4 * og + c;
          tile[n][14 + o] = tanhf(acc[j][c]) * dj_dropmul(d_conv, (uint32_t)(bt * N_ + n) * OU_ + o);
        }
      }
    }
    for (int i = tid; i < N_ * 32; i += 256) {   // the 30 non-conv features + 2 pad columns
      const int n = i / 32, q = i % 32;
      const int col = q < 14 ? q : 78 + (q - 14);
      float v;
      if (col == 0) v = (float)n / (float)N_;
      else if (col <= DJ_OCTAVE) v = (n % DJ_OCTAVE == col - 1) ? 1.f : 0.f;
      else if (col == 13) v = bins_s[n];
      else if (col < F0_) v = beat_s[col - 78];
      else v = 0.f;
      tile[n][col] = v;
    }
    __syncthreads();
    for (int i = tid; i < N_ * (F0P_ / 4); i += 256) {   // + style term, coalesced write-out
      const int n = i / (F0P_ / 4), f4 = (i % (F0P_ / 4)) * 4;
      const uint32_t row = (uint32_t)bt * N_ + n;
      float m[4], v[4];
      dj_dropmul4(d_sp, row * F0P_ + f4, m);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int f = f4 + j;
        v[j] = tile[n][f];
        if (f < F0_) v[j] = fmaf(sp0[(int64_t)bt * F0_ + f], m[j], v[j]);
      }
      store4(A0 + (int64_t)row * ldA + f4, v);
      if (A0lo != nullptr) store4_lo(A0lo + (int64_t)row * ldA + f4, v);
    }
    for (int i = tid; i < N_ * ((ldA - F0P_) / 4); i += 256) {   // zero any extra padding
      const int w = (ldA - F0P_) / 4, n = i / w, f4 = F0P_ + (i % w) * 4;
      const float z[4] = {0.f, 0.f, 0.f, 0.f};
      store4(A0 + ((int64_t)bt * N_ + n) * ldA + f4, z);
      if (A0lo != nullptr) store4(A0lo + ((int64_t)bt * N_ + n) * ldA + f4, z);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// A operand of LSTM layers 1.. : drop(h_prev) (+ shifted chosen) + drop(style)
// ---------------------------------------------------------------------------
// One warp per row (grid-stride over rows): the row's (b, t, n) split and base pointers are computed once
// with 32-bit arithmetic, then each lane converts 8 consecutive features per trip (two 16-byte loads of h,
// one 16-byte bf16 store).
template <typename TA>
__global__ void __launch_bounds__(256) layer_input_kernel(
    const float* __restrict__ h_prev, int Uprev, int64_t h_row0, int64_t h_b_rows, dj_dropout d_h,
    const float* __restrict__ sp, int F, dj_dropout d_sp, const float* __restrict__ chosen_in,
    int64_t chosen_bstride, dj_dropout d_chosen, int B, int T, TA* __restrict__ A, TA* __restrict__ Alo, int ldA) {
  dj_resolve(d_h); dj_resolve(d_sp); dj_resolve(d_chosen);
  const int ld4 = (F + 3) & ~3;
  const int lane = threadIdx.x & 31;
  const uint32_t rows = (uint32_t)B * (uint32_t)T * N_;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += nwarps) {
    const uint32_t n = row % N_, bt = row / N_;
    const uint32_t b = bt / (uint32_t)T, t = bt - b * (uint32_t)T;
    const float* hrow = h_prev + (h_row0 + (int64_t)b * h_b_rows + (int64_t)t * N_ + n) * Uprev;
    const float* sprow = sp + (int64_t)bt * F;
    // shift_chosen (model.py:101): previous NOTE of the same timestep, 3 channels
    const float* cr = (chosen_in != nullptr && n > 0)
                          ? chosen_in + (int64_t)b * chosen_bstride + ((int64_t)t * N_ + (n - 1)) * NU_ : nullptr;
    for (int f8 = lane * 8; f8 < ldA; f8 += 256) {
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int f4 = f8 + 4 * hh;
        float* vv = v + 4 * hh;
        if (f4 < Uprev) {
          const float4 hv = *reinterpret_cast<const float4*>(hrow + f4);
          float m[4];
          dj_dropmul4(d_h, row * (uint32_t)Uprev + f4, m);
          vv[0] = hv.x * m[0]; vv[1] = hv.y * m[1]; vv[2] = hv.z * m[2]; vv[3] = hv.w * m[3];
        } else if (f4 < F && cr != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = f4 + j - Uprev;
            if (c < NU_) vv[j] = cr[c] * dj_dropmul(d_chosen, (row - 1) * 4u + c);
          }
        }
        if (f4 < F) {
          float m[4];
          dj_dropmul4(d_sp, row * (uint32_t)ld4 + f4, m);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (f4 + j < F) vv[j] = fmaf(sprow[f4 + j], m[j], vv[j]);
        }
      }
      TA* out = A + (int64_t)row * ldA + f8;
      store4(out, v);
      store4(out + 4, v + 4);
      if (Alo != nullptr) {
        TA* olo = Alo + (int64_t)row * ldA + f8;
        store4_lo(olo, v);
        store4_lo(olo + 4, v + 4);
      }
    }
  }
}

// Fast path of the same function for the shapes the engine launches: both dropout sites in byte mode (training) or
// both off (generation), 16-bit operands, rows laid out b-major without gaps (h row = h_row0 + row, chosen row =
// row), Uprev a multiple of 8.  One thread per 8-feature chunk over a flat (row, chunk) index -- no idle lanes
// whatever ldA is --, 16-byte loads of h and of the style projection, 16-byte stores of A and A_lo, the byte test
// of a mask word done on the word in place.  Same arithmetic, element by element, as layer_input_kernel.
__device__ __forceinline__ uint4 pack8(const float v[8], __nv_bfloat16) {
  uint4 u;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}
__device__ __forceinline__ uint4 pack8(const float v[8], __half) {
  uint4 u;
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  return u;
}
__device__ __forceinline__ void unpack8(const uint4& u, float r[8], __nv_bfloat16) {   // bf16 -> fp32 is a shift
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    r[2 * j] = __uint_as_float(w[j] << 16);
    r[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000U);
  }
}
__device__ __forceinline__ void unpack8(const uint4& u, float r[8], __half) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    r[2 * j] = f.x; r[2 * j + 1] = f.y;
  }
}
// multipliers of the 4 elements of one byte-mode mask word: byte j >= t ? scale : 0 (compared in place)
__device__ __forceinline__ void mask_mul4(uint32_t w, uint32_t t, float scale, float m[4]) {
  m[0] = ((w & 0xffU) >= t) ? scale : 0.f;
  m[1] = ((w & 0xff00U) >= (t << 8)) ? scale : 0.f;
  m[2] = ((w & 0xff0000U) >= (t << 16)) ? scale : 0.f;
  m[3] = (w >= (t << 24)) ? scale : 0.f;
}

template <typename TA, bool CHOSEN, bool DROP>
__global__ void __launch_bounds__(256) layer_input_fast_kernel(
    const float* __restrict__ h_prev, uint32_t Uprev, const float* __restrict__ sp, uint32_t F, dj_dropout d_h,
    dj_dropout d_sp, const float* __restrict__ chosen_in, dj_dropout d_chosen, uint32_t rows, uint32_t ld8,
    uint32_t ld8_magic, TA* __restrict__ A, TA* __restrict__ Alo) {
  dj_resolve(d_h); dj_resolve(d_sp); dj_resolve(d_chosen);
  const uint32_t ld4 = (F + 3) & ~3u, total = rows * ld8;
  const uint32_t th = d_h.thr >> 24, ts = d_sp.thr >> 24;
  for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const uint32_t row = __umulhi(i, ld8_magic);        // i / ld8 (exact: host checks total * (magic * ld8 - 2^32) < 2^32)
    const uint32_t f8 = (i - row * ld8) * 8;
    const uint32_t bt = row / N_, n = row - bt * N_;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (f8 < Uprev) {
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(h_prev + (size_t)row * Uprev + f8));
      const float4 h1 = __ldg(reinterpret_cast<const float4*>(h_prev + (size_t)row * Uprev + f8 + 4));
      v[0] = h0.x; v[1] = h0.y; v[2] = h0.z; v[3] = h0.w; v[4] = h1.x; v[5] = h1.y; v[6] = h1.z; v[7] = h1.w;
      if (DROP) {
        const uint32_t q = (row * Uprev + f8) >> 2;
        float m[8];
        mask_mul4(dj_mask_word(d_h.key, q), th, d_h.scale, m);
        mask_mul4(dj_mask_word(d_h.key, q + 1), th, d_h.scale, m + 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= m[j];
      }
    } else if (CHOSEN && f8 == Uprev && n > 0) {
      // shift_chosen (model.py:101): previous NOTE of the same timestep, 3 channels
      const float* cr = chosen_in + (size_t)(row - 1) * NU_;
#pragma unroll
      for (int c = 0; c < NU_; ++c) v[c] = __ldg(cr + c) * dj_dropmul(d_chosen, (row - 1) * 4u + c);
    }
    if (f8 < F) {
      float s[8], m[8];
      if (!CHOSEN) {      // F == Uprev: whole, 16-byte aligned chunks
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(sp + (size_t)bt * F + f8));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(sp + (size_t)bt * F + f8 + 4));
        s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
      } else if (f8 + 8 <= F) {   // F = Uprev + 3: the rows of sp are not 16-byte aligned, but 32 of the 36 chunks are whole
        const float* sr = sp + (size_t)bt * F + f8;
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = __ldg(sr + j);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = (f8 + j < F) ? __ldg(sp + (size_t)bt * F + f8 + j) : 0.f;
      }
      if (DROP) {
        const uint32_t q = (row * ld4 + f8) >> 2;
        mask_mul4(dj_mask_word(d_sp.key, q), ts, d_sp.scale, m);
        mask_mul4(dj_mask_word(d_sp.key, q + 1), ts, d_sp.scale, m + 4);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = 1.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(s[j], m[j], v[j]);      // s is zero beyond F
    }
    const uint4 hi = pack8(v, TA());
    *reinterpret_cast<uint4*>(A + (size_t)row * (ld8 * 8) + f8) = hi;
    if (Alo != nullptr) {   // residual against the rounded values just packed (no second rounding of v)
      float r[8];
      unpack8(hi, r, TA());
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = v[j] - r[j];
      *reinterpret_cast<uint4*>(Alo + (size_t)row * (ld8 * 8) + f8) = pack8(r, TA());
    }
  }
}

__global__ void mask_materialize_kernel(dj_dropout d, int64_t rows, int F, float* __restrict__ out) {
  dj_resolve(d);
  const int ld4 = (F + 3) & ~3;
  const int64_t total = rows * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / F;
    const int f = (int)(i % F);
    out[i] = (d.mode == 0 || dj_keep(d, (uint32_t)(row * ld4 + f))) ? 1.f : 0.f;
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ in, int rows, int cols,
                                 __nv_bfloat16* __restrict__ out, int ldo, int transpose, int orows) {
  // out has `orows` rows of ldo; element (r_o, c_o) comes from in[r_o, c_o] or in[c_o, r_o]
  const int64_t total = (int64_t)orows * ldo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int ro = (int)(i / ldo), co = (int)(i % ldo);
    float v = 0.f;
    if (!transpose) { if (co < cols) v = in[(int64_t)ro * cols + co]; }
    else { if (co < rows) v = in[(int64_t)co * cols + ro]; }
    out[i] = __float2bfloat16_rn(v);
  }
}

struct CastBatch {
  const float* in[16];
  uint16_t* out[16];
  uint16_t* out_lo[16];      // nullable: 16-bit residual v - hi (same format)
  int rows[16], cols[16], ldo[16], transpose[16], fmt[16];
  float scale[16];           // the value is multiplied by this (a power of two) before the split
};
__device__ __forceinline__ uint16_t to16(float v, int fmt, float& back) {
  if (fmt == DJ_F16) { const __half h = __float2half_rn(v); back = __half2float(h); return __half_as_ushort(h); }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  back = __bfloat162float(h);
  return __bfloat16_as_ushort(h);
}
// all 16-bit operand copies of a step in one launch: blockIdx.y selects the tensor
__global__ void cast_bf16_multi_kernel(CastBatch b) {
  const int e = blockIdx.y;
  const float* __restrict__ in = b.in[e];
  uint16_t* __restrict__ out = b.out[e];
  uint16_t* __restrict__ out_lo = b.out_lo[e];
  const int rows = b.rows[e], cols = b.cols[e], ldo = b.ldo[e], tr = b.transpose[e], fmt = b.fmt[e];
  const int orows = tr ? cols : rows;
  const int64_t total = (int64_t)orows * ldo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ro = (int)(i / ldo), co = (int)(i % ldo);
    float v = 0.f;
    if (!tr) { if (co < cols) v = in[(int64_t)ro * cols + co]; }
    else { if (co < rows) v = in[(int64_t)co * cols + ro]; }
    float back, back2;
    v *= b.scale[e];
    out[i] = to16(v, fmt, back);
    if (out_lo != nullptr) out_lo[i] = to16(v - back, fmt, back2);
  }
}

// 8 elements (one 16-byte vector) per thread and trip
__global__ void half_to_bf16_kernel(uint4* __restrict__ buf, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 q = buf[i];
    uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[j]);
      const __nv_bfloat162 b = __floats2bfloat162_rn(__low2float(h), __high2float(h));
      w[j] = *reinterpret_cast<const uint32_t*>(&b);
    }
    buf[i] = q;
  }
}

inline int grid_for(int64_t total, int threads, int max_blocks) {
  int64_t b = (total + threads - 1) / threads;
  if (b < 1) b = 1;
  return (int)(b > max_blocks ? max_blocks : b);
}

}  // namespace

extern "C" int dj_make_dropout(uint64_t seed, int site, float rate, dj_dropout* out) {
  DJ_CHECK_ARG(out != nullptr, "dj_make_dropout: out is NULL");
  DJ_CHECK_ARG(rate >= 0.f && rate < 1.f, "dj_make_dropout: rate %f outside [0,1)", rate);
  out->key_ptr = nullptr;
  if (rate == 0.f) { out->key = 0; out->thr = 0; out->scale = 1.f; out->mode = 0; return 0; }
  const double thr = (double)rate * 4294967296.0;
  out->key = dj_site_key(seed, site);
  out->thr = (uint32_t)thr;
  out->scale = 1.0f / (1.0f - rate);
  const double t256 = (double)rate * 256.0;
  out->mode = (t256 == (double)(int)t256) ? 1 : 2;
  return 0;
}

extern "C" uint32_t dj_dropout_site_key(uint64_t seed, int site) { return dj_site_key(seed, site); }

extern "C" int dj_dropout_mask_materialize(dj_dropout d, int64_t rows, int F, float* out, void* stream) {
  DJ_CHECK_ARG(out && rows > 0 && F > 0, "dj_dropout_mask_materialize: bad arguments");
  DJ_CHECK_ARG(rows * ((F + 3) & ~3) < (int64_t)4294967296LL, "dj_dropout_mask_materialize: site too large");
  mask_materialize_kernel<<<grid_for(rows * F, 256, dj_num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(d, rows, F, out);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_style_fwd(const float* style_in, int64_t style_bstride, int64_t style_tstride,
                            int num_styles, int B, int T, const float* Ws, const float* bs, int n_proj,
                            const float* const* Wsd, const float* const* bsd, const int* F, float* emb,
                            float* const* sp, void* stream) {
  DJ_CHECK_ARG(style_in && Ws && bs && emb, "dj_style_fwd: NULL pointer");
  DJ_CHECK_ARG(num_styles > 0 && num_styles <= 32, "dj_style_fwd: num_styles %d not in 1..32", num_styles);
  DJ_CHECK_ARG(B > 0 && T > 0 && n_proj >= 0 && n_proj <= 4, "dj_style_fwd: bad sizes");
  StyleProj pr{};
  for (int l = 0; l < n_proj; ++l) {
    DJ_CHECK_ARG(Wsd[l] && bsd[l] && sp[l] && F[l] > 0, "dj_style_fwd: projection %d incomplete", l);
    pr.W[l] = Wsd[l]; pr.b[l] = bsd[l]; pr.out[l] = sp[l]; pr.F[l] = F[l];
  }
  const int BT = B * T;
  style_fwd_kernel<<<(BT + 7) / 8, 256, 0, (cudaStream_t)stream>>>(style_in, style_bstride, style_tstride,
                                                                  num_styles, B, T, Ws, bs, n_proj, pr, emb);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_frontend_fwd(const float* notes_in, int64_t notes_bstride, const float* beat_in,
                               int64_t beat_bstride, int B, int T, const float* Wc, const float* bc,
                               const float* sp0, dj_dropout d_notes, dj_dropout d_beat, dj_dropout d_conv,
                               dj_dropout d_sp, void* A0, void* A0_lo, int ldA, int a_dtype, void* stream) {
  DJ_CHECK_ARG(notes_in && beat_in && Wc && bc && sp0 && A0, "dj_frontend_fwd: NULL pointer");
  DJ_CHECK_ARG(A0_lo == nullptr || a_dtype == DJ_BF16 || a_dtype == DJ_F16, "dj_frontend_fwd: the residual operand A0_lo exists for 16-bit operands only");
  DJ_CHECK_ARG(B > 0 && T > 0, "dj_frontend_fwd: bad B/T");
  DJ_CHECK_ARG(ldA >= F0P_ && ldA % 8 == 0, "dj_frontend_fwd: ldA %d must be >=96 and a multiple of 8", ldA);
  DJ_CHECK_ARG((int64_t)B * T * N_ * F0P_ < (int64_t)4294967296LL, "dj_frontend_fwd: batch too large");
  const int grid = grid_for((int64_t)B * T, 1, dj_num_sms() * 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (a_dtype == DJ_F32)
    frontend_fwd_kernel<float><<<grid, 256, 0, st>>>(notes_in, notes_bstride, beat_in, beat_bstride, B, T, Wc,
                                                     bc, sp0, d_notes, d_beat, d_conv, d_sp, (float*)A0, (float*)nullptr, ldA);
  else if (a_dtype == DJ_BF16)
    frontend_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(notes_in, notes_bstride, beat_in, beat_bstride, B,
                                                             T, Wc, bc, sp0, d_notes, d_beat, d_conv, d_sp,
                                                             (__nv_bfloat16*)A0, (__nv_bfloat16*)A0_lo, ldA);
  else if (a_dtype == DJ_F16)
    frontend_fwd_kernel<__half><<<grid, 256, 0, st>>>(notes_in, notes_bstride, beat_in, beat_bstride, B, T, Wc, bc, sp0,
                                                      d_notes, d_beat, d_conv, d_sp, (__half*)A0, (__half*)A0_lo, ldA);
  else DJ_CHECK_ARG(false, "dj_frontend_fwd: unknown dtype %d", a_dtype);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_layer_input(const float* h_prev, int Uprev, int64_t h_row0, int64_t h_b_rows,
                              dj_dropout d_h, const float* sp, int F, dj_dropout d_sp,
                              const float* chosen_in, int64_t chosen_bstride, dj_dropout d_chosen, int B,
                              int T, void* A, void* A_lo, int ldA, int a_dtype, void* stream) {
  DJ_CHECK_ARG(h_prev && sp && A, "dj_layer_input: NULL pointer");
  DJ_CHECK_ARG(A_lo == nullptr || a_dtype == DJ_BF16 || a_dtype == DJ_F16, "dj_layer_input: the residual operand A_lo exists for 16-bit operands only");
  DJ_CHECK_ARG(Uprev > 0 && Uprev % 4 == 0 && F >= Uprev, "dj_layer_input: bad Uprev %d / F %d", Uprev, F);
  DJ_CHECK_ARG(F == Uprev || (chosen_in && F == Uprev + NU_), "dj_layer_input: F must be Uprev or Uprev+3 with chosen");
  DJ_CHECK_ARG(ldA >= ((F + 3) & ~3) && ldA % 8 == 0, "dj_layer_input: ldA %d too small or not a multiple of 8", ldA);
  DJ_CHECK_ARG((int64_t)B * T * N_ * ((F + 3) & ~3) < (int64_t)4294967296LL, "dj_layer_input: batch too large");
  cudaStream_t st = (cudaStream_t)stream;
  {
    // fast path (what the engine launches): see layer_input_fast_kernel
    const bool drop = d_h.mode == 1 && d_sp.mode == 1, nodrop = d_h.mode == 0 && d_sp.mode == 0;
    const uint32_t rows = (uint32_t)B * T * N_, ld8 = (uint32_t)ldA / 8;
    const uint64_t magic = ((1ull << 32) + ld8 - 1) / ld8, total8 = (uint64_t)rows * ld8;
    const bool exact = total8 < (1ull << 32) && total8 * (magic * ld8 - (1ull << 32)) < (1ull << 32) && magic < (1ull << 32);
    const bool dense = h_b_rows == (int64_t)T * N_ && (chosen_in == nullptr || chosen_bstride == (int64_t)T * N_ * NU_);
    const bool aligned = Uprev % 8 == 0 && ((uintptr_t)h_prev % 16 == 0) && ((uintptr_t)sp % 16 == 0) &&
                         ((uintptr_t)A % 16 == 0) && ((uintptr_t)A_lo % 16 == 0);
    if ((drop || nodrop) && exact && dense && aligned && (a_dtype == DJ_BF16 || a_dtype == DJ_F16) &&
        (chosen_in != nullptr || F == Uprev)) {
      const float* hp = h_prev + h_row0 * Uprev;
      const int grid = grid_for((int64_t)total8, 256, dj_num_sms() * 8);
#define DJ_LI_FAST(TA, CH, DR)                                                                                      \
  layer_input_fast_kernel<TA, CH, DR><<<grid, 256, 0, st>>>(hp, (uint32_t)Uprev, sp, (uint32_t)F, d_h, d_sp, chosen_in, \
                                                            d_chosen, rows, ld8, (uint32_t)magic, (TA*)A, (TA*)A_lo)
      const bool ch = chosen_in != nullptr;
      if (a_dtype == DJ_BF16) {
        if (ch && drop) DJ_LI_FAST(__nv_bfloat16, true, true);
        else if (ch) DJ_LI_FAST(__nv_bfloat16, true, false);
        else if (drop) DJ_LI_FAST(__nv_bfloat16, false, true);
        else DJ_LI_FAST(__nv_bfloat16, false, false);
      } else {
        if (ch && drop) DJ_LI_FAST(__half, true, true);
        else if (ch) DJ_LI_FAST(__half, true, false);
        else if (drop) DJ_LI_FAST(__half, false, true);
        else DJ_LI_FAST(__half, false, false);
      }
#undef DJ_LI_FAST
      DJ_LAUNCH_CHECK();
      return 0;
    }
  }
  const int64_t total = (int64_t)B * T * N_ * 32;   // one warp per row
  const int grid = grid_for(total, 256, dj_num_sms() * 16);
  if (a_dtype == DJ_F32)
    layer_input_kernel<float><<<grid, 256, 0, st>>>(h_prev, Uprev, h_row0, h_b_rows, d_h, sp, F, d_sp,
                                                    chosen_in, chosen_bstride, d_chosen, B, T, (float*)A, (float*)nullptr, ldA);
  else if (a_dtype == DJ_BF16)
    layer_input_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(h_prev, Uprev, h_row0, h_b_rows, d_h, sp, F, d_sp,
                                                            chosen_in, chosen_bstride, d_chosen, B, T,
                                                            (__nv_bfloat16*)A, (__nv_bfloat16*)A_lo, ldA);
  else if (a_dtype == DJ_F16)
    layer_input_kernel<__half><<<grid, 256, 0, st>>>(h_prev, Uprev, h_row0, h_b_rows, d_h, sp, F, d_sp, chosen_in,
                                                     chosen_bstride, d_chosen, B, T, (__half*)A, (__half*)A_lo, ldA);
  else DJ_CHECK_ARG(false, "dj_layer_input: unknown dtype %d", a_dtype);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_cast_bf16(const float* in, int rows, int cols, void* out, int ldo, int transpose,
                            void* stream) {
  DJ_CHECK_ARG(in && out && rows > 0 && cols > 0, "dj_cast_bf16: bad arguments");
  const int orows = transpose ? cols : rows;
  DJ_CHECK_ARG(ldo >= (transpose ? rows : cols), "dj_cast_bf16: ldo %d too small", ldo);
  cast_bf16_kernel<<<grid_for((int64_t)orows * ldo, 256, dj_num_sms() * 8), 256, 0, (cudaStream_t)stream>>>(
      in, rows, cols, (__nv_bfloat16*)out, ldo, transpose, orows);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_half_to_bf16_inplace(void* buf, int64_t n, void* stream) {
  DJ_CHECK_ARG(buf && n > 0 && n % 8 == 0 && ((uintptr_t)buf % 16) == 0, "dj_half_to_bf16_inplace: needs a 16-byte aligned buffer of 8k elements");
  half_to_bf16_kernel<<<grid_for(n / 8, 256, dj_num_sms() * 8), 256, 0, (cudaStream_t)stream>>>((uint4*)buf, n / 8);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_cast16_multi(int n, const float* const* in, const int* rows, const int* cols, void* const* out,
                               void* const* out_lo, const int* ldo, const int* transpose, const int* fmt,
                               const float* scale, void* stream) {
  DJ_CHECK_ARG(n > 0 && n <= 16 && in && rows && cols && out && ldo && transpose && fmt, "dj_cast16_multi: bad arguments");
  CastBatch b{};
  for (int i = 0; i < n; ++i) {
    DJ_CHECK_ARG(in[i] && out[i] && rows[i] > 0 && cols[i] > 0 && ldo[i] >= (transpose[i] ? rows[i] : cols[i]) &&
                     (fmt[i] == DJ_BF16 || fmt[i] == DJ_F16),
                 "dj_cast16_multi: entry %d invalid", i);
    b.in[i] = in[i]; b.out[i] = (uint16_t*)out[i]; b.out_lo[i] = out_lo ? (uint16_t*)out_lo[i] : nullptr;
    b.rows[i] = rows[i]; b.cols[i] = cols[i]; b.ldo[i] = ldo[i]; b.transpose[i] = transpose[i]; b.fmt[i] = fmt[i];
    b.scale[i] = scale ? scale[i] : 1.0f;
  }
  cast_bf16_multi_kernel<<<dim3(64, n), 256, 0, (cudaStream_t)stream>>>(b);
  DJ_LAUNCH_CHECK();
  return 0;
}
