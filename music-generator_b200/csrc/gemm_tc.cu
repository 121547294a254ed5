// tcgen05 / TMA / TMEM GEMMs for the LSTM gate projections (sm_100a only).
//
//   dj_gate_gemm_bf16 : C[M,N] = A[M,K] . Bt[N,K]^T + bias     (x.W+b of keras
//       LSTM, model.py:84,120, and its data gradient dX = dZ.W^T)
//   dj_wgrad_gemm_bf16: C[Ka,Nb] += A[M,Ka]^T . B[M,Nb]         (dW = X^T.dZ and
//       dU = H_{t-1}^T.dZ: contraction over the M = B*T*48 rows)
//
// Structure of both: persistent CTAs, warp 0 = TMA producer (one lane), warp 1 =
// MMA issuer (one lane, tcgen05.mma cta_group::1, 128xN tile, fp32 accumulators
// in TMEM), warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> regs ->
// swizzled smem -> TMA store / fp32 reduction).  smem stages are 128B-swizzled
// tiles written by TMA and read by the tensor core through shared-memory matrix
// descriptors; full/empty mbarriers pipeline TMA against MMA, tmem_full/empty
// pipeline MMA against the epilogue (two accumulator stages).
#include <stdlib.h>

#include "dj_tc.cuh"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int ACC_STAGES = 2;
constexpr int A_BYTES = BM * BK * 2;
constexpr int EPI_BOXES = 4;                     // 32x32 fp32 staging boxes per epilogue warp (TMA stores in flight)
constexpr int EPI_WARP_BYTES = EPI_BOXES * 32 * 128;
constexpr int NUM_THREADS = 256;

constexpr int MAX_BIAS_N = 2048;                  // bias is staged in shared memory once per CTA
// BN = 128: four stages; BN = 256 (one 128x256x16 MMA feeds the tensor core twice as long per byte of A read from
// shared memory: the tile of the 3-pass split product, which is bound by the MMA loop, not by writing C): three
// stages of 48 KB and all 512 tensor-memory columns for the two accumulator stages
template <int BN> struct GemmSmem {
  static constexpr int STAGES = BN == 256 ? 3 : 4;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = STAGES * A_BYTES;
  static constexpr int EPI_OFF = B_OFF + STAGES * B_BYTES;
  static constexpr int BIAS_OFF = EPI_OFF + 4 * EPI_WARP_BYTES;
  static constexpr int BAR_OFF = BIAS_OFF + MAX_BIAS_N * 4;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;   // + alignment slack
};

// ---------------------------------------------------------------------------
// C = A . Bt^T + bias, both operands K-major
// ---------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gate_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                 const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, int M, int N, int K,
                 int npass, uint32_t idesc, float out_scale) {
  using SM = GemmSmem<BN>;
  constexpr int STAGES = SM::STAGES, B_BYTES = SM::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + SM::BAR_OFF;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + ACC_STAGES + a); };
  uint32_t* tmem_slot = (uint32_t*)(smem + SM::BAR_OFF + 8 * (2 * STAGES + 2 * ACC_STAGES));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_n = (N + BN - 1) / BN, num_m = (M + BM - 1) / BM;
  const int num_tiles = num_m * num_n, num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC);
    if (npass > 1) { prefetch_tmap(&tmAlo); prefetch_tmap(&tmBlo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < ACC_STAGES; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), ACC_STAGES * BN);
  float* bias_s = reinterpret_cast<float*>(smem + SM::BIAS_OFF);
  if (bias != nullptr)   // one global read of the bias per CTA; the epilogue then reads it from shared memory
    for (int i = threadIdx.x; i < ((N + 127) / 128) * 128; i += NUM_THREADS) bias_s[i] = (i < N) ? bias[i] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer =====
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
        // split-operand passes (fp32-grade product from 16-bit operands): hi.hi, then lo.hi, then hi.lo
        for (int pass = 0; pass < npass; ++pass) {
          const CUtensorMap* mA = (pass == 1) ? &tmAlo : &tmA;
          const CUtensorMap* mB = (pass == 2) ? &tmBlo : &tmB;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), A_BYTES + B_BYTES);
            tma_load_2d(sbase + SM::A_OFF + stage * A_BYTES, mA, full_bar(stage), kb * BK, m0);
            tma_load_2d(sbase + SM::B_OFF + stage * B_BYTES, mB, full_bar(stage), kb * BK, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ===== MMA issuer =====
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t aphase = 0;
      const int num_it = npass * num_kb;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_it; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(sbase + SM::A_OFF + stage * A_BYTES, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sbase + SM::B_OFF + stage * B_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)   // UMMA_K = 16 bf16 = 32 B inside the 128 B swizzle row
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          umma_commit(empty_bar(stage));   // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == ACC_STAGES) { acc = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {   // ===== epilogue =====
    const int e = warp - 4;
    uint8_t* ebuf = smem + SM::EPI_OFF + e * EPI_WARP_BYTES;
    const uint32_t ebuf_s = smem_u32(ebuf);
    int acc = 0; uint32_t aphase = 0; int nbuf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM, n0 = (tile % num_n) * BN;
      mbar_wait(tfull_bar(acc), aphase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * e) << 16) + (uint32_t)(acc * BN + c * 32), v);
        if (lane == 0) tma_store_wait_read<EPI_BOXES - 1>();   // the store that last used this staging box is done
        __syncwarp();
        const int ncol = n0 + c * 32;
        uint8_t* box = ebuf + nbuf * 4096;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o;
          o.x = __uint_as_float(v[4 * q + 0]); o.y = __uint_as_float(v[4 * q + 1]);
          o.z = __uint_as_float(v[4 * q + 2]); o.w = __uint_as_float(v[4 * q + 3]);
          o.x *= out_scale; o.y *= out_scale; o.z *= out_scale; o.w *= out_scale;   // exact for the power-of-two scales used
          if (bias != nullptr) {
            const float4 bv = *reinterpret_cast<const float4*>(bias_s + ncol + 4 * q);   // broadcast smem read
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          }
          // 128B swizzle: 16-byte chunk q of row `lane` lives at chunk q ^ (lane & 7)
          *reinterpret_cast<float4*>(box + lane * 128 + ((q ^ (lane & 7)) << 4)) = o;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, ebuf_s + nbuf * 4096, ncol, m0 + 32 * e);
          tma_store_commit();
        }
        nbuf = (nbuf + 1) % EPI_BOXES;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == ACC_STAGES) { acc = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, ACC_STAGES * BN);
}

// ---------------------------------------------------------------------------
// C[Ka,Nb] += A[M,Ka]^T . B[M,Nb]   (weight gradients; both operands MN-major)
//
// The contraction runs over the M activation rows, so both operands sit in
// memory with the NON-contracted index contiguous.  TMA boxes of 64 columns x
// 64 rows with the 128B swizzle land in smem exactly in UMMA's MN-major SW128
// canonical layout (8 rows x 128 B atoms; SBO = 1024 B between 8-row groups along
// K, LBO = one whole box between 64-column groups).  One CTA owns one 128x256
// output tile and a slice of M; fp32 partial sums leave TMEM through vectorised
// global reductions (red.global.add.v4.f32).
// ---------------------------------------------------------------------------
constexpr int WG_BM = 128, WG_BN = 256, WG_BK = 64, WG_STAGES = 4;
constexpr int WG_BOX_BYTES = WG_BK * 128;                 // one 64-column box: 8 KB
constexpr int WG_A_BYTES = (WG_BM / 64) * WG_BOX_BYTES;   // 16 KB
constexpr int WG_B_BYTES = (WG_BN / 64) * WG_BOX_BYTES;   // 32 KB
struct WgradSmem {
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = WG_STAGES * WG_A_BYTES;
  static constexpr int BAR_OFF = B_OFF + WG_STAGES * WG_B_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  float* __restrict__ C, int64_t ldc, int Ka, int Nb, int64_t M, int num_tiles, int kb_per_split,
                  uint32_t idesc, float* __restrict__ part) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + WgradSmem::BAR_OFF;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (WG_STAGES + s); };
  const uint32_t done_bar = bars + 8u * (2 * WG_STAGES);
  uint32_t* tmem_slot = (uint32_t*)(smem + WgradSmem::BAR_OFF + 8 * (2 * WG_STAGES + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (Nb + WG_BN - 1) / WG_BN;
  const int tile = blockIdx.x % num_tiles, split = blockIdx.x / num_tiles;
  const int ka0 = (tile / tiles_n) * WG_BM, nb0 = (tile % tiles_n) * WG_BN;
  const int total_kb = (int)((M + WG_BK - 1) / WG_BK);
  const int kb_begin = split * kb_per_split;
  const int kb_end = min(total_kb, kb_begin + kb_per_split);
  const int nkb = max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_slot), WG_BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer: 2 boxes of A, 4 boxes of B per stage =====
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), WG_A_BYTES + WG_B_BYTES);
        const int r0 = kb * WG_BK;
#pragma unroll
        for (int j = 0; j < WG_BM / 64; ++j)
          tma_load_2d(sbase + WgradSmem::A_OFF + stage * WG_A_BYTES + j * WG_BOX_BYTES, &tmA, full_bar(stage), ka0 + 64 * j, r0);
#pragma unroll
        for (int j = 0; j < WG_BN / 64; ++j)
          tma_load_2d(sbase + WgradSmem::B_OFF + stage * WG_B_BYTES + j * WG_BOX_BYTES, &tmB, full_bar(stage), nb0 + 64 * j, r0);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {   // ===== MMA issuer =====
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t a0 = sbase + WgradSmem::A_OFF + stage * WG_A_BYTES;
        const uint32_t b0 = sbase + WgradSmem::B_OFF + stage * WG_B_BYTES;
#pragma unroll
        for (int k = 0; k < WG_BK / 16; ++k) {   // 16 contraction rows = 2048 B further down the box
          const uint64_t adesc = make_smem_desc(a0 + k * 2048, WG_BOX_BYTES, 1024);
          const uint64_t bdesc = make_smem_desc(b0 + k * 2048, WG_BOX_BYTES, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (i | k) != 0);
        }
        umma_commit(empty_bar(stage));
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else if (warp >= 4 && nkb > 0) {   // ===== epilogue: TMEM -> fp32 global reductions =====
    const int e = warp - 4;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int row = ka0 + 32 * e + lane;
#pragma unroll 1
    for (int c = 0; c < WG_BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(32 * e) << 16) + (uint32_t)(c * 32), v);
      if (row < Ka && part != nullptr) {
        // deterministic mode: this split's partial tile, added later in split order (dj_ordered_reduce)
        float* dst = part + ((int64_t)split * Ka + row) * Nb + nb0 + c * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = nb0 + c * 32 + 4 * q;
          if (n + 3 < Nb) {
            *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                  __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          } else {
            for (int j = 0; j < 4; ++j)
              if (n + j < Nb) dst[4 * q + j] = __uint_as_float(v[4 * q + j]);
          }
        }
      } else if (row < Ka) {
        float* dst = C + (int64_t)row * ldc + nb0 + c * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int n = nb0 + c * 32 + 4 * q;
          if (n + 3 < Nb) {
            red_add_v4(dst + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
          } else {
            for (int j = 0; j < 4; ++j)
              if (n + j < Nb) atomicAdd(dst + 4 * q + j, __uint_as_float(v[4 * q + j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, WG_BN);
}

}  // namespace

// kind::f16 operand formats in the instruction descriptor: bits 7-9 (A) / 10-12 (B): 0 = f16, 1 = bf16.
// make_idesc() sets both to bf16; this clears the bit of every operand that is IEEE half.
static inline bool fmt16_ok(int f) { return f == DJ_BF16 || f == DJ_F16; }
static inline uint32_t idesc_with_formats(uint32_t idesc, int a_fmt, int b_fmt) {
  if (a_fmt == DJ_F16) idesc &= ~(1u << 7);
  if (b_fmt == DJ_F16) idesc &= ~(1u << 10);
  return idesc;
}
static inline CUtensorMapDataType tmap_dtype(int fmt) {
  return fmt == DJ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

template <int BN>
static int launch_gate_gemm(const CUtensorMap& tmA, const CUtensorMap& tmAlo, const void* Bt, const void* Bt_lo, int b_fmt,
                            int64_t ldb, const CUtensorMap& tmC, const float* bias, int M, int N, int K, int npass,
                            uint32_t idesc_fmt_mask, float out_scale, cudaStream_t st) {
  CUtensorMap tmB, tmBlo;
  int rc;
  if ((rc = make_map_2d(&tmB, tmap_dtype(b_fmt), 2, Bt, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN))) return rc;
  tmBlo = tmB;
  if (npass == 3 && (rc = make_map_2d(&tmBlo, tmap_dtype(b_fmt), 2, Bt_lo, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN)))
    return rc;
  DJ_CUDA(cudaFuncSetAttribute((const void*)gate_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               GemmSmem<BN>::TOTAL));
  const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  int grid = dj_num_sms();
  if (grid > tiles) grid = tiles;
  const uint32_t idesc = make_idesc(BM, BN, 0, 0) & idesc_fmt_mask;
  gate_gemm_kernel<BN><<<grid, NUM_THREADS, GemmSmem<BN>::TOTAL, st>>>(tmA, tmAlo, tmB, tmBlo, tmC, bias, M, N, K, npass, idesc,
                                                                      out_scale);
  DJ_LAUNCH_CHECK();
  return 0;
}

extern "C" int dj_gate_gemm_16s(const void* A, const void* A_lo, int a_fmt, int64_t lda, const void* Bt,
                                const void* Bt_lo, int b_fmt, int64_t ldb, float* C, int64_t ldc, const float* bias,
                                float out_scale, int M, int N, int K, void* stream) {
  DJ_CHECK_ARG(A && Bt && C, "dj_gate_gemm_16s: NULL pointer");
  DJ_CHECK_ARG((A_lo == nullptr) == (Bt_lo == nullptr), "dj_gate_gemm_16s: A_lo and Bt_lo come together (3-pass split product)");
  DJ_CHECK_ARG(fmt16_ok(a_fmt) && a_fmt == b_fmt,
               "dj_gate_gemm_16s: operand formats must be DJ_BF16 or DJ_F16 and equal (kind::f16 cannot mix half with bf16)");
  DJ_CHECK_ARG(M > 0 && N > 0 && K > 0, "dj_gate_gemm_16s: bad shape");
  DJ_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0 && lda >= K && ldb >= K && ldc >= N,
               "dj_gate_gemm_16s: leading dimensions must be 16-byte multiples and cover K/N (lda=%lld ldb=%lld ldc=%lld)",
               (long long)lda, (long long)ldb, (long long)ldc);
  DJ_CHECK_ARG(bias == nullptr || N <= MAX_BIAS_N, "dj_gate_gemm_16s: N=%d exceeds the bias staging capacity %d", N, MAX_BIAS_N);
  DJ_CHECK_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)Bt % 16) == 0 && ((uintptr_t)C % 16) == 0 &&
                   ((uintptr_t)A_lo % 16) == 0 && ((uintptr_t)Bt_lo % 16) == 0 &&
                   (bias == nullptr || ((uintptr_t)bias % 16) == 0),
               "dj_gate_gemm_16s: pointers must be 16-byte aligned");
  const int npass = A_lo ? 3 : 1;
  CUtensorMap tmA, tmAlo, tmC;
  int rc;
  if ((rc = make_map_2d(&tmA, tmap_dtype(a_fmt), 2, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM))) return rc;
  tmAlo = tmA;
  if (npass == 3 && (rc = make_map_2d(&tmAlo, tmap_dtype(a_fmt), 2, A_lo, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)))
    return rc;
  if ((rc = make_map_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32, 32))) return rc;
  const uint32_t fmt_mask = idesc_with_formats(0xFFFFFFFFu, a_fmt, b_fmt);
  // the split product is bound by its MMA loop: 128x256 tiles when N allows (DJ_GEMM_BN=128 forces the small tile)
  static int bn_env = -1;
  if (bn_env < 0) { const char* e = getenv("DJ_GEMM_BN"); bn_env = e ? atoi(e) : 0; }
  const bool wide = (npass == 3 && N % 256 == 0 && bn_env != 128) || bn_env == 256;
  if (wide)
    return launch_gate_gemm<256>(tmA, tmAlo, Bt, Bt_lo, b_fmt, ldb, tmC, bias, M, N, K, npass, fmt_mask, out_scale,
                                 (cudaStream_t)stream);
  return launch_gate_gemm<128>(tmA, tmAlo, Bt, Bt_lo, b_fmt, ldb, tmC, bias, M, N, K, npass, fmt_mask, out_scale,
                               (cudaStream_t)stream);
}

extern "C" int dj_gate_gemm_16(const void* A, const void* A_lo, int a_fmt, int64_t lda, const void* Bt,
                               const void* Bt_lo, int b_fmt, int64_t ldb, float* C, int64_t ldc, const float* bias,
                               int M, int N, int K, void* stream) {
  return dj_gate_gemm_16s(A, A_lo, a_fmt, lda, Bt, Bt_lo, b_fmt, ldb, C, ldc, bias, 1.0f, M, N, K, stream);
}

extern "C" int dj_gate_gemm_bf16(const void* A, int64_t lda, const void* Bt, int64_t ldb, float* C, int64_t ldc,
                                 const float* bias, int M, int N, int K, void* stream) {
  return dj_gate_gemm_16(A, nullptr, DJ_BF16, lda, Bt, nullptr, DJ_BF16, ldb, C, ldc, bias, M, N, K, stream);
}

extern "C" int dj_wgrad_gemm_16(const void* A, int a_fmt, int64_t lda, const void* B, int b_fmt, int64_t ldb, float* C,
                                int64_t ldc, int Ka, int Nb, int64_t M, void* stream) {
  DJ_CHECK_ARG(A && B && C, "dj_wgrad_gemm_16: NULL pointer");
  DJ_CHECK_ARG(fmt16_ok(a_fmt) && a_fmt == b_fmt,
               "dj_wgrad_gemm_16: operand formats must be DJ_BF16 or DJ_F16 and equal (kind::f16 cannot mix half with bf16)");
  DJ_CHECK_ARG(Ka > 0 && Nb > 0 && M > 0, "dj_wgrad_gemm_16: bad shape");
  DJ_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0 && lda >= Ka && ldb >= Nb && ldc >= Nb,
               "dj_wgrad_gemm_16: leading dimensions must be 16-byte multiples and cover the widths");
  DJ_CHECK_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0 && ((uintptr_t)C % 16) == 0,
               "dj_wgrad_gemm_16: pointers must be 16-byte aligned");
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_map_2d(&tmA, tmap_dtype(a_fmt), 2, A, (uint64_t)Ka, (uint64_t)M, (uint64_t)lda, 64, WG_BK))) return rc;
  if ((rc = make_map_2d(&tmB, tmap_dtype(b_fmt), 2, B, (uint64_t)Nb, (uint64_t)M, (uint64_t)ldb, 64, WG_BK))) return rc;
  DJ_CUDA(cudaFuncSetAttribute((const void*)wgrad_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgradSmem::TOTAL));
  const int tiles = ((Ka + WG_BM - 1) / WG_BM) * ((Nb + WG_BN - 1) / WG_BN);
  const int total_kb = (int)((M + WG_BK - 1) / WG_BK);
  int splits = dj_num_sms() / tiles;
  if (splits < 1) splits = 1;
  if (splits > total_kb) splits = total_kb;
  const int kbps = (total_kb + splits - 1) / splits;
  splits = (total_kb + kbps - 1) / kbps;
  const uint32_t idesc = idesc_with_formats(make_idesc(WG_BM, WG_BN, 1, 1), a_fmt, b_fmt);
  float* part = nullptr;
  if (dj_reduce_workspace(stream, (int64_t)splits * Ka * Nb, &part)) return -1;
  DJ_CHECK_ARG(part == nullptr || Nb % 4 == 0, "dj_wgrad_gemm_16: deterministic mode needs Nb %% 4 == 0 (got %d)", Nb);
  wgrad_gemm_kernel<<<tiles * splits, NUM_THREADS, WgradSmem::TOTAL, (cudaStream_t)stream>>>(tmA, tmB, C, ldc, Ka, Nb, M,
                                                                                             tiles, kbps, idesc, part);
  DJ_LAUNCH_CHECK();
  if (part != nullptr) return dj_ordered_reduce(part, splits, (int64_t)Ka * Nb, Ka, Nb, Nb, C, ldc, stream);
  return 0;
}

extern "C" int dj_wgrad_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                                  int Ka, int Nb, int64_t M, void* stream) {
  return dj_wgrad_gemm_16(A, DJ_BF16, lda, B, DJ_BF16, ldb, C, ldc, Ka, Nb, M, stream);
}
