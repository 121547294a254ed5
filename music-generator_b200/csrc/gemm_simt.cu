// Generic fp32 CUDA-core GEMM with arbitrary operand strides.
// Used where fp32-grade results are required (generation: bit-exact sampling
// needs |p_gpu - p_oracle| ~ 1e-6), for the small style/conv
// weight gradients, and as the cross-check of the tcgen05 kernels.
#include "dj_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename TA, typename TB, bool A_KCONTIG, bool B_NCONTIG>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(
    const TA* __restrict__ A, int64_t a_sm, int64_t a_sk, const TB* __restrict__ B, int64_t b_sk,
    int64_t b_sn, float* __restrict__ C, int64_t ldc, const float* __restrict__ bias, int M, int N, int K,
    int accumulate, int64_t a_shift, int64_t a_period, int k_per_split, int use_atomic, float* __restrict__ part) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * NT;
      int m, k;
      if (A_KCONTIG) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) {
        if (a_period > 0) {
          if ((gk % a_period) >= a_shift) v = ldf(A + (int64_t)gm * a_sm + (int64_t)(gk - a_shift) * a_sk);
        } else {
          v = ldf(A + (int64_t)gm * a_sm + (int64_t)gk * a_sk);
        }
      }
      ra[i] = v;
      int n, kb;
      if (B_NCONTIG) { n = idx % BN; kb = idx / BN; } else { kb = idx % BK; n = idx / BK; }
      const int gn = n0 + n, gkb = k0 + kb;
      rb[i] = (gn < N && gkb < kend) ? ldf(B + (int64_t)gkb * b_sk + (int64_t)gn * b_sn) : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * NT;
      int m, k;
      if (A_KCONTIG) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      As[buf][k][m] = ra[i];
      int n, kb;
      if (B_NCONTIG) { n = idx % BN; kb = idx / BN; } else { kb = idx % BK; n = idx / BK; }
      Bs[buf][kb][n] = rb[i];
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = (k0 + BK) < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[gn];
      float* dst = C + (int64_t)gm * ldc + gn;
      if (part != nullptr) part[((int64_t)blockIdx.z * M + gm) * N + gn] = v;   // deterministic split-K: added in order later
      else if (use_atomic) atomicAdd(dst, v);
      else *dst = accumulate ? (*dst + v) : v;
    }
  }
}

template <typename TA, typename TB>
int launch_simt(const void* A, int64_t a_sm, int64_t a_sk, const void* B, int64_t b_sk, int64_t b_sn, float* C,
                int64_t ldc, const float* bias, int M, int N, int K, int accumulate, int64_t a_shift,
                int64_t a_period, cudaStream_t st) {
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int splits = 1;
  const int target = dj_num_sms() * 4;
  if (tiles < target && K >= 4096) {
    splits = (target + tiles - 1) / tiles;
    const int maxs = K / 512;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  int kps = (K + splits - 1) / splits;
  kps = (kps + BK - 1) / BK * BK;
  splits = (K + kps - 1) / kps;
  const int use_atomic = splits > 1;
  if (use_atomic && !accumulate) {
    if (ldc == N) DJ_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, st));
    else DJ_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, M, st));
  }
  float* part = nullptr;
  if (use_atomic && dj_reduce_workspace((void*)st, (int64_t)splits * M * N, &part)) return -1;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
  const bool akc = (a_sk == 1), bnc = (b_sn == 1);
  const TA* Ap = (const TA*)A;
  const TB* Bp = (const TB*)B;
#define LAUNCH(AK, BNc)                                                                                      \
  gemm_simt_kernel<TA, TB, AK, BNc><<<grid, NT, 0, st>>>(Ap, a_sm, a_sk, Bp, b_sk, b_sn, C, ldc, bias, M, N, \
                                                         K, accumulate, a_shift, a_period, kps, use_atomic, part)
  if (akc && bnc) LAUNCH(true, true);
  else if (akc && !bnc) LAUNCH(true, false);
  else if (!akc && bnc) LAUNCH(false, true);
  else LAUNCH(false, false);
#undef LAUNCH
  DJ_LAUNCH_CHECK();
  if (part != nullptr) return dj_ordered_reduce(part, splits, (int64_t)M * N, M, N, N, C, ldc, (void*)st);
  return 0;
}

}  // namespace

extern "C" int dj_gemm_simt(const void* A, int a_dtype, int64_t a_sm, int64_t a_sk, const void* B, int b_dtype,
                            int64_t b_sk, int64_t b_sn, float* C, int64_t ldc, const float* bias, int M, int N,
                            int K, int accumulate, int64_t a_shift, int64_t a_period, void* stream) {
  DJ_CHECK_ARG(A && B && C, "dj_gemm_simt: NULL pointer");
  DJ_CHECK_ARG(M > 0 && N > 0 && K > 0 && ldc >= N, "dj_gemm_simt: bad shape M=%d N=%d K=%d ldc=%lld", M, N, K,
               (long long)ldc);
  DJ_CHECK_ARG(a_period == 0 || (a_shift >= 0 && a_shift < a_period), "dj_gemm_simt: bad shift/period");
  cudaStream_t st = (cudaStream_t)stream;
  if (a_dtype == DJ_F32 && b_dtype == DJ_F32)
    return launch_simt<float, float>(A, a_sm, a_sk, B, b_sk, b_sn, C, ldc, bias, M, N, K, accumulate, a_shift, a_period, st);
  if (a_dtype == DJ_BF16 && b_dtype == DJ_BF16)
    return launch_simt<__nv_bfloat16, __nv_bfloat16>(A, a_sm, a_sk, B, b_sk, b_sn, C, ldc, bias, M, N, K, accumulate, a_shift, a_period, st);
  if (a_dtype == DJ_F32 && b_dtype == DJ_BF16)
    return launch_simt<float, __nv_bfloat16>(A, a_sm, a_sk, B, b_sk, b_sn, C, ldc, bias, M, N, K, accumulate, a_shift, a_period, st);
  if (a_dtype == DJ_BF16 && b_dtype == DJ_F32)
    return launch_simt<__nv_bfloat16, float>(A, a_sm, a_sk, B, b_sk, b_sn, C, ldc, bias, M, N, K, accumulate, a_shift, a_period, st);
  DJ_CHECK_ARG(false, "dj_gemm_simt: unknown dtypes %d/%d", a_dtype, b_dtype);
  return -1;
}
