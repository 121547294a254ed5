// LSTM recurrence on the 5th-generation tensor cores (training path).
//
// Same job as lstm_scan.cu (the sequential half of keras.layers.LSTM,
// model.py:84,120) but h_{t-1}.U runs as tcgen05.mma with bf16 operands and fp32
// accumulators in TMEM; the gate nonlinearities / cell update run in the
// TMEM-load epilogue with warp shuffles, as registers-only work.
//
// Per cluster (C = U/32 CTAs) one tile of BS independent sequences:
//   CTA r owns hidden units [32r, 32r+32) = 128 gate-interleaved columns.
//   A operand  = its slice of U^T  [128 gate cols x U]  bf16, K-major SW128, loaded
//                once by TMA and resident in shared memory for the whole launch.
//   B operand  = the tile's h_{t-1} [BS x U] bf16, K-major SW128.
//   D (TMEM)   = [128 lanes x BS columns] fp32:  lane = gate column, column = sequence.
// Step t:  (1) MMA warp waits for h_{t-1} to land, issues U/16 MMAs, commits;
//          (2) 8 epilogue warps: tcgen05.ld + x.W pre-activations (fp32, global),
//              4x4 shuffle transpose so each lane holds i,f,g,o of one cell,
//              gates -> c,h; store gates (in place, for backward), h, c, and h_t as
//              bf16 into the `hprev` buffer at the NEXT step's row;
//          (3) one cluster barrier;
//          (4) every CTA TMA-loads 1/C of the tile's h_t rows from `hprev` (L2) and
//              MULTICASTS it to all C CTAs -> B operand of step t+1.
// `hprev` is needed anyway as the A operand of dU = H_{t-1}^T.dZ, so the all-gather
// costs no extra HBM traffic and avoids the ~20 B/clk DSMEM store path.
#include <cooperative_groups.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "dj_tc.cuh"

namespace cg = cooperative_groups;

#ifdef DJ_TRACE
// debug build only (tools/scan_trace.py): per-step clock64() stamps of one CTA's issuer and epilogue lanes
__device__ long long* g_dj_trace = nullptr;
extern "C" int dj_debug_trace_set(void* buf) {
  return (int)cudaMemcpyToSymbol(g_dj_trace, &buf, sizeof(buf));
}
#define DJ_TR(t, k)                                                                                   \
  do {                                                                                                \
    if (g_dj_trace != nullptr && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2 + 1))                \
      g_dj_trace[(((blockIdx.x == 0) ? 0 : 1) * 512 + (t)) * 16 + (k)] = clock64();                   \
  } while (0)
#define DJ_TRV(t, k, v)                                                                               \
  do {                                                                                                \
    if (g_dj_trace != nullptr && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2 + 1))                \
      g_dj_trace[(((blockIdx.x == 0) ? 0 : 1) * 512 + (t)) * 16 + (k)] = (v);                         \
  } while (0)
#define DJ_CLK() clock64()
#else
#define DJ_TR(t, k) do {} while (0)
#define DJ_TRV(t, k, v) do {} while (0)
#define DJ_CLK() 0ll
#endif

namespace {

struct TcMap {
  int seq_inner;                       // 48 (time axis: seq = b*48+n) or 1 (note axis)
  int64_t outer_stride, inner_stride, step_stride;
  // TMA coordinates of a half-tile slice of hprev: c1 = t*step1 + hh*off1, c2 = tile*base2 + hh*off2
  int step1, off1, base2, off2;
  int64_t seq_stride;                  // row distance of consecutive sequences inside a 16-chunk
};

constexpr int TC_EPI_WARPS = 8;       // two per TMEM lane quarter
constexpr int TC_THREADS = 32 * (2 + TC_EPI_WARPS);   // warp 0 and warp 9 each drive one half-tile's chain

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// INFER: the B operand is h as IEEE half hi + lo (two tiles)
template <int U, int BS, bool INFER = false>
struct TcFwdSmem {
  static constexpr int A_BYTES = 128 * U * 2;
  static constexpr int H_BYTES = BS * U * 2 * (INFER ? 2 : 1);
  static constexpr int A_OFF = 0, H_OFF = A_BYTES, BAR_OFF = A_BYTES + H_BYTES;
  static constexpr int TOTAL = BAR_OFF + 96;          // no static smem: two CTAs must fit one SM
};

// single-instruction MUFU forms (ftz: no denormal range fix-up around ex2, no Newton step after rcp)
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}
// tanh(x) = 1 - 2 / (e^{2x} + 1): exact limits at +-inf (ex2 -> inf / 0), 5 instructions
__device__ __forceinline__ float fast_tanh(float x) {
  return fmaf(-2.0f, rcp_ftz(ex2_ftz(x * 2.885390081777927f) + 1.0f), 1.0f);
}
__device__ __forceinline__ float hard_sig_sat(float x) { return __saturatef(fmaf(x, 0.2f, 0.5f)); }
template <bool HARD> __device__ __forceinline__ float gate_act_fast(float x) {
  if constexpr (HARD) return hard_sig_sat(x);
  else return rcp_ftz(1.0f + ex2_ftz(x * -1.4426950408889634f));
}
// The activated gates are saved for the reverse scan as IEEE half (8 bytes per cell instead of 16: the forward
// epilogue is bound by the SM's store path and the reverse scan by its loads).  The reverse scan only needs them to
// ~1e-3, EXCEPT for the hard-sigmoid derivative, which is an indicator of 0 < a < 1: a value strictly inside the
// interval must not round onto its end, so it is stored as the nearest half strictly inside.
template <bool HARD> __device__ __forceinline__ uint32_t gate_half(float a) {
  uint32_t u = __half_as_ushort(__float2half_rn(a));
  if (HARD) {
    if (a < 1.0f && u == 0x3C00u) u = 0x3BFFu;     // largest half below 1
    if (a > 0.0f && u == 0u) u = 1u;               // smallest positive half
  }
  return u;
}
template <bool HARD> __device__ __forceinline__ uint2 pack_gates16(float gi, float gf, float gg, float go) {
  uint2 r;
  r.x = gate_half<HARD>(gi) | (gate_half<HARD>(gf) << 16);
  r.y = (uint32_t)__half_as_ushort(__float2half_rn(gg)) | (gate_half<HARD>(go) << 16);
  return r;
}
__device__ __forceinline__ float4 unpack_gates16(uint2 r) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ int64_t tc_row0(const TcMap& m, int seq) {
  return (int64_t)(seq / m.seq_inner) * m.outer_stride + (int64_t)(seq % m.seq_inner) * m.inner_stride;
}
// Forward recurrence.  h_t leaves the epilogue as bf16 into `hprev` (global, the row of
// step t+1); after ONE cluster barrier per step every CTA TMA-loads 1/C of the tile's
// rows back from L2 and multicasts them into all C CTAs' B-operand tiles.  (A variant
// that pushed 16-byte chunks through distributed shared memory instead measured 25 %
// slower per step on B200: DSMEM stores are the slow path.)  Two CTAs of different
// clusters share an SM so one tile's barrier / TMA / MMA latency chain is covered by the
// other tile's epilogue; the next pre-activations are always in flight in registers.
// TIME: sequences (b, n) stepping t (sequence stride 1 row, step stride 48 rows); else the note axis:
// sequences (b, t) stepping n (sequence stride 48 rows, step stride 1).  Both strides and the gate
// activation are compile-time so every row offset inside a chunk is an immediate.
//
// Warp roles: warps 0 and 9 issue (TMA, MMA) and publish, one half-tile each; warps 1..8 are the epilogue,
// two per TMEM lane quarter, each taking 8 of the 16 sequences of every chunk.
//
// NS = 2 software-pipelines the tile as two half-tiles that alternate through the same epilogue warps:
// while they do the gate math of half B, half A's publish -> all-gather -> MMA chain is in flight.
// Per half-step the hand-offs are mbarriers, not the hardware cluster barrier:
//   epilogue warps  --done[hf] (count 8, release.cta)-->  issuer
//   issuer: ONE gpu-scope release for the whole CTA (the pattern of a grid barrier: CTA-level sync, then
//           one thread's fence publishes everybody's stores), then a remote arrive on pub[hf] of all C CTAs
//   pub[hf] complete (count C) -> this CTA's multicast TMA of its slice of h_t -> bar_h[hf] -> MMA -> bar_acc[hf]
// so the epilogue warps never execute a gpu-scope membar or a cluster barrier.
// ATM: the resident A operand (this CTA's 128 rows of U^T) lives in TENSOR MEMORY instead of shared memory
// (tcgen05.mma with A from TMEM): the per-step MMAs then read only the small B operand from shared memory.  With A
// in shared memory the 16 MMAs of a half-step are bound by re-reading the 64 KB of weights (~80 cycles per MMA
// against a 24-cycle floor).  Needs 2 x (BS + U/2 rounded up) <= 512 TMEM columns for two CTAs per SM: U <= 256.
// INFER (generation: the sampled events must equal the fp32 model's, so the recurrence has to be fp32-grade):
// h_{t-1} enters as IEEE half hi + lo and U as half hi + lo (both ~22 mantissa bits; U pre-scaled by a power of
// two, undone by acc_scale, so its residual stays out of half's subnormal range), three MMA passes per step on the
// same accumulator: U_hi.h_hi + U_lo.h_hi + U_hi.h_lo.  Nothing is saved for a backward pass: no gate / c stores.
template <int U, int BS, bool TIME, bool HARD, int NB, int NS, bool ATM, bool INFER = false>
__global__ void __launch_bounds__(TC_THREADS, 2)
scan_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmH,
                   const __grid_constant__ CUtensorMap tmHlo, const float* __restrict__ Z, uint2* __restrict__ G16,
                   float* __restrict__ Hout, float* __restrict__ Cout, uint16_t* __restrict__ Hprev,
                   uint16_t* __restrict__ Hprev_lo,
                   const uint32_t* __restrict__ Ut_words, int steps, TcMap map, int f16, int has_lo, float acc_scale) {
  static_assert(!INFER || ATM, "inference mode keeps the A operand in tensor memory");
  constexpr uint32_t SSTR = TIME ? 1u : 48u, TSTR = TIME ? 48u : 1u;
  constexpr int C = U / 32;            // cluster size
  constexpr int KA = U / 64;           // 64-wide K atoms
  constexpr int RH = BS / 2;           // rows per multicast slice
  constexpr int NCH = BS / 16;         // 16-sequence chunks per tile
  constexpr int HB = BS / NS;          // sequences per half-tile (= MMA N)
  constexpr int NCHH = NCH / NS;       // chunks per half-tile
  static_assert(NS == 1 || NS == 2, "one tile or two half-tiles");
  static_assert(NS == 1 || (NCH % 2 == 0), "half-tiles are whole 16-sequence chunks");
  static_assert(C == 2 * KA && RH % 8 == 0, "slices = K atoms x 2 row halves");
  constexpr uint32_t D_COLS = BS <= 32 ? 32 : BS <= 64 ? 64 : BS <= 128 ? 128 : 256;   // accumulators: columns [0, BS)
  constexpr uint32_t A_COL0 = D_COLS;                                                   // A operand: columns [A_COL0, +U/2)
  constexpr uint32_t NEED = ATM ? D_COLS + U / 2 : D_COLS;
  constexpr uint32_t TMEM_COLS = NEED <= 32 ? 32 : NEED <= 64 ? 64 : NEED <= 128 ? 128 : NEED <= 256 ? 256 : 512;
  // U = 512 needs 176 KB of shared memory, so only one CTA is resident per SM and it may take all 512 columns
  static_assert(BS % 16 == 0 && BS <= 256 && (!ATM || TMEM_COLS <= 256 || U == 512), "tile shape / two CTAs share 512 TMEM columns");
  // h and U operands: bf16 or IEEE half (kind::f16 format bits 7 / 10 of the instruction descriptor: 0 = f16)
  const uint32_t idesc = f16 ? (make_idesc(128, HB, 0, 0) & ~((1u << 7) | (1u << 10))) : make_idesc(128, HB, 0, 0);
  // has_lo (ATM only): shared memory holds the 16-bit residual U - hi(U) of this CTA's slice; a second MMA pass adds
  // h.U_lo so the recurrent weights enter the product with ~22 mantissa bits
  const bool smem_a = !ATM || has_lo;
  using SM = TcFwdSmem<U, BS, INFER>;
  constexpr int HLO_OFF = BS * U * 2;     // INFER: the lo tiles follow the hi tiles inside the H region

  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  // barriers: a | acc[2] | h[2] | done[2] | pub[2] | tmem slot
  const uint32_t bar_a = sbase + SM::BAR_OFF, bar_acc0 = bar_a + 8, bar_h0 = bar_a + 24, bar_done0 = bar_a + 40,
                 bar_pub0 = bar_a + 56;
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(smem + SM::BAR_OFF + 72);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tile = blockIdx.x / C;     // one tile of BS sequences per cluster
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      if (sbase & 1023u) { printf("deepj scan_tc_fwd: dynamic smem not 1024-aligned\n"); __trap(); }
      prefetch_tmap(&tmU); prefetch_tmap(&tmH);
      if (INFER) prefetch_tmap(&tmHlo);
      mbar_init(bar_a, 1);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        mbar_init(bar_acc0 + 8 * hf, 1); mbar_init(bar_h0 + 8 * hf, 1);
        mbar_init(bar_done0 + 8 * hf, TC_EPI_WARPS); mbar_init(bar_pub0 + 8 * hf, C);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(sbase + SM::BAR_OFF + 72, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  if (ATM && warp >= 1 && warp <= 4) {
    // lane = one row of this CTA's slice of U^T (256 B .. 1 KB contiguous in global memory): 16 packed bf16 pairs
    // per tcgen05.st, straight into the TMEM lanes of this warp's quarter
    const uint32_t* src = Ut_words + (size_t)(128 * rank + 32 * (warp & 3) + lane) * (U / 2);
#pragma unroll 1
    for (int c = 0; c < U / 2; c += 16) {
      uint32_t w[16];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 q4 = *reinterpret_cast<const uint4*>(src + c + 4 * j);
        w[4 * j] = q4.x; w[4 * j + 1] = q4.y; w[4 * j + 2] = q4.z; w[4 * j + 3] = q4.w;
      }
      tmem_st16(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + A_COL0 + (uint32_t)c, w);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  cluster.sync();   // peers' barriers are initialised before any multicast / remote arrive can target them (and A is in TMEM)
  tc_fence_after();

  if (warp == 0 || warp == 1 + TC_EPI_WARPS) {
    // ================= issuer + publisher of half-tile hf =================
    // the whole warp walks the loop converged; single-thread work sits under elect_one()
    const int hf = (warp == 0) ? 0 : 1;
    if (smem_a && warp == 0 && elect_one()) {   // resident shared-memory A operand (U^T, or its residual): rows [128*rank, +128)
      mbar_expect_tx(bar_a, SM::A_BYTES);
#pragma unroll
      for (int ka = 0; ka < KA; ++ka)
        tma_load_2d(sbase + SM::A_OFF + ka * 16384, &tmU, bar_a, ka * 64, 128 * rank);
    }
    __syncwarp();
    const int my_ka = rank >> 1, my_hh = rank & 1;       // the slice of h_t this CTA multicasts
    // time-axis tiles smaller than a batch element (BS < 48): batch element and first row of this tile
    constexpr int TPB = (TIME && BS < 48) ? 48 / BS : 1;
    const int tile_b = tile / TPB, tile_r = (tile % TPB) * BS;
    const uint32_t bar_done = bar_done0 + 8 * hf, bar_pub = bar_pub0 + 8 * hf, bar_h = bar_h0 + 8 * hf,
                   bar_acc = bar_acc0 + 8 * hf;
    if (hf < NS) {
      for (int t = 0; t + 1 < steps; ++t) {               // the last step publishes nothing
        const uint32_t par = (uint32_t)t & 1u;
        mbar_wait(bar_done, par);                         // this CTA's epilogue warps stored this half of h_t
        DJ_TR(t, 4 * hf + 0);
        if (lane < C) mbar_arrive_remote(bar_pub, (uint32_t)lane);   // release.cluster, cumulative
        __syncwarp();
        // every CTA of the cluster has published.  The producers' release put h_t in L2 and the only consumer is the
        // TMA below, which reads L2: a CTA-scope wait is enough (a cluster-scope acquire adds an L1 invalidate)
        mbar_wait(bar_pub, par);
        DJ_TR(t, 4 * hf + 1);
        if (elect_one()) {
          mbar_expect_tx(bar_h, HB * U * 2 * (INFER ? 2 : 1));
          if (NS == 1 || my_hh == hf) {
            tma_load_3d_mc(sbase + SM::H_OFF + my_ka * (BS * 128) + my_hh * (RH * 128), &tmH, bar_h, my_ka * 64,
                           (t + 1) * map.step1 + tile_r + my_hh * map.off1, tile_b * map.base2 + my_hh * map.off2,
                           (uint16_t)((1u << C) - 1u));
            if constexpr (INFER)
              tma_load_3d_mc(sbase + SM::H_OFF + HLO_OFF + my_ka * (BS * 128) + my_hh * (RH * 128), &tmHlo, bar_h,
                             my_ka * 64, (t + 1) * map.step1 + tile_r + my_hh * map.off1,
                             tile_b * map.base2 + my_hh * map.off2, (uint16_t)((1u << C) - 1u));
          }
        }
        __syncwarp();
        if (smem_a && t == 0) mbar_wait(bar_a, 0);
        mbar_wait(bar_h, par);                            // all C slices of this half landed
        DJ_TR(t, 4 * hf + 2);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc0 = make_smem_desc(sbase + SM::A_OFF, 16, 1024);
          const uint64_t bdesc0 = make_smem_desc(sbase + SM::H_OFF + hf * (HB * 128), 16, 1024);
#pragma unroll
          for (int ka = 0; ka < KA; ++ka)
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // descriptor start addresses are in 16-byte units
              const uint64_t bdesc = bdesc0 + (uint64_t)((ka * (BS * 128) + k * 32) >> 4);
              if constexpr (ATM)
                umma_bf16_ts(tmem_base + (uint32_t)(hf * HB), tmem_base + A_COL0 + (uint32_t)(ka * 32 + k * 8), bdesc,
                             idesc, (ka | k) != 0);
              else
                umma_bf16(tmem_base + (uint32_t)(hf * HB), adesc0 + (uint64_t)((ka * 16384 + k * 32) >> 4), bdesc,
                          idesc, (ka | k) != 0);
            }
          if (ATM && has_lo) {
#pragma unroll
            for (int ka = 0; ka < KA; ++ka)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + (uint32_t)(hf * HB), adesc0 + (uint64_t)((ka * 16384 + k * 32) >> 4),
                          bdesc0 + (uint64_t)((ka * (BS * 128) + k * 32) >> 4), idesc, 1);
          }
          if constexpr (INFER) {   // U_hi . h_lo
#pragma unroll
            for (int ka = 0; ka < KA; ++ka)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ts(tmem_base + (uint32_t)(hf * HB), tmem_base + A_COL0 + (uint32_t)(ka * 32 + k * 8),
                             bdesc0 + (uint64_t)((HLO_OFF + ka * (BS * 128) + k * 32) >> 4), idesc, 1);
          }
          umma_commit(bar_acc);
        }
        __syncwarp();
        DJ_TR(t, 4 * hf + 3);
      }
    }
  } else {
    // ================= epilogue warps =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int w2 = (warp - 1) >> 2;         // which 8 of a chunk's 16 sequences
    const int up = lane >> 2;               // unit inside the warp (0..7)
    const int g = lane & 3;                 // gate held before the transpose / sequence slot after it
    const uint32_t col = 32 * rank + 8 * q + up;          // global hidden unit
    const uint32_t zc = 128 * rank + 32 * q + lane;       // gate-interleaved column this lane reads
    float cst[NCH][2];
    uint32_t rowb[NCH];                     // row of this warp's first sequence of each chunk at step 0
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      cst[i][0] = cst[i][1] = 0.f;
      // a 16-chunk never straddles a batch element
      rowb[i] = (uint32_t)tc_row0(map, tile * BS + i * 16) + (uint32_t)(8 * w2) * SSTR;
    }
    // x.W pre-activations are prefetched NB chunks ahead into registers (NB divides NCH, so after
    // unrolling every buffer index is static)
    static_assert(NCH % NB == 0, "prefetch ring must divide the chunk count");
    float zreg[NB][8];
    auto load_z = [&](int ci, uint32_t t) {
      const float* zp = Z + (size_t)(rowb[ci] + t * TSTR) * (4 * U) + zc;
#pragma unroll
      for (int j = 0; j < 8; ++j) zreg[ci % NB][j] = zp[(size_t)j * SSTR * (4 * U)];
    };
#pragma unroll
    for (int ci = 0; ci < NB; ++ci) load_z(ci, 0);
    const bool tr = (warp == 1 && lane == 0);
    for (int t = 0; t < steps; ++t) {
      const bool not_last = (t + 1 < steps);
      const uint32_t par = (uint32_t)(t + 1) & 1u;        // accumulators of step t were produced in issuer round t-1
#pragma unroll
      for (int hf = 0; hf < NS; ++hf) {
        if (t > 0) {
          mbar_wait(bar_acc0 + 8 * hf, par);
          tc_fence_after();
        }
        if (tr) DJ_TR(t, 8 + 4 * hf);
        if (not_last && lane < 8) {   // warm L2 with the next step's x.W rows (this warp's 128-byte segments)
#pragma unroll
          for (int ci = hf * NCHH; ci < (hf + 1) * NCHH; ++ci)
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(Z + (size_t)(rowb[ci] + lane * SSTR + (t + 1) * TSTR) * (4 * U) +
                                                                128 * rank + 32 * q));
        }
#pragma unroll
        for (int ci = hf * NCHH; ci < (hf + 1) * NCHH; ++ci) {
          float v[8];
          if (t > 0) {
            uint32_t acc[8];
            tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(ci * 16 + 8 * w2), acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(acc[j]), acc_scale, zreg[ci % NB][j]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = zreg[ci % NB][j];
          }
          // refill the buffer just consumed: NB chunks ahead, wrapping into the next step
          if (ci + NB < NCH) load_z(ci + NB, t);
          else if (not_last) load_z(ci + NB - NCH, t + 1);
          // this lane's cell after the transpose: sequence 16*ci + 8*w2 + 4*blk + g, unit `col`
          const uint32_t row_c = rowb[ci] + (uint32_t)g * SSTR + (uint32_t)t * TSTR;
          const size_t o1 = (size_t)row_c * U + col;
          uint2* const g16 = G16 + o1;                       // the four activated gates of this cell as IEEE half
          float* const hp = Hout + o1;
          float* const cp = Cout + o1;                       // only dereferenced when Cout != nullptr
          uint16_t* const hb = Hprev + o1;
          uint16_t* const hbl = Hprev_lo + o1;                // INFER only
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            // 4x4 transpose across the 4 lanes of a unit: lane g ends with i,f,g,o of sequence 4*blk+g
            const float a0 = v[4 * blk], a1 = v[4 * blk + 1], a2 = v[4 * blk + 2], a3 = v[4 * blk + 3];
            const bool odd = g & 1, hi = g & 2;
            const float x1 = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
            const float x2 = __shfl_xor_sync(0xffffffffu, odd ? a2 : a3, 1);
            const float b0 = odd ? x1 : a0, b1 = odd ? a1 : x1, b2 = odd ? x2 : a2, b3 = odd ? a3 : x2;
            const float y0 = __shfl_xor_sync(0xffffffffu, hi ? b0 : b2, 2);
            const float y1 = __shfl_xor_sync(0xffffffffu, hi ? b1 : b3, 2);
            const float zi = hi ? y0 : b0, zf = hi ? y1 : b1, zg_ = hi ? b2 : y0, zo = hi ? b3 : y1;

            const float gi = gate_act_fast<HARD>(zi), gf = gate_act_fast<HARD>(zf);
            const float gg = fast_tanh(zg_), go = gate_act_fast<HARD>(zo);
            const float cn = fmaf(gf, cst[ci][blk], gi * gg);
            const float hn = go * fast_tanh(cn);
            cst[ci][blk] = cn;

            constexpr size_t RO = (size_t)4 * SSTR;          // rows between consecutive blocks
            // h_t (bf16) at the NEXT step's row: the A operand of dU = H_{t-1}^T.dZ
            if constexpr (INFER) {
              if (not_last) {
                const __half hh = __float2half_rn(hn);
                hb[(blk * RO + TSTR) * U] = __half_as_ushort(hh);
                hbl[(blk * RO + TSTR) * U] = __half_as_ushort(__float2half_rn(hn - __half2float(hh)));
              }
            } else {
              if (not_last)
                hb[(blk * RO + TSTR) * U] = f16 ? __half_as_ushort(__float2half_rn(hn)) : __bfloat16_as_ushort(__float2bfloat16_rn(hn));
              if (t == 0) hb[blk * RO * U] = 0;
            }
#ifndef DJ_EXP
#define DJ_EXP 0   // timing experiments only: bit0/1/2 drop the gate / h / c stores
#endif
            if (!INFER && !(DJ_EXP & 1)) g16[blk * RO * U] = pack_gates16<HARD>(gi, gf, gg, go);
            if (!(DJ_EXP & 2)) hp[blk * RO * U] = hn;
            if (!INFER && !(DJ_EXP & 4) && Cout != nullptr) cp[blk * RO * U] = cn;
          }
        }
        if (tr) DJ_TR(t, 9 + 4 * hf);
        if (not_last) {
          tc_fence_before();
          fence_proxy_async_all();   // generic-proxy global stores of h_t -> later async-proxy (TMA) reads
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_done0 + 8 * hf);   // release.cta: hands this warp's stores to the issuer
        }
        if (tr) DJ_TR(t, 10 + 4 * hf);
      }
    }
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

inline bool fwd_atm_enabled() {   // DJ_FWD_ATM=0 keeps the A operand in shared memory (experiments)
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("DJ_FWD_ATM");
    env = (e && atoi(e) == 0) ? 0 : 1;
  }
  return env != 0;
}

template <int U, int BS, bool TIME, bool HARD, int NB, int NS, bool INFER = false>
int launch_tc_fwd_inst(const void* Ut_bf, const void* Ut_lo, int f16, const float* Z, void* gates16, float* h_out,
                       float* c_out, void* hprev, int S, int steps, const TcMap& map_in, cudaStream_t st,
                       void* hprev_lo = nullptr, float acc_scale = 1.0f) {
  constexpr int C = U / 32;
  using SM = TcFwdSmem<U, BS, INFER>;
  DJ_CHECK_ARG(S % BS == 0, "dj_lstm_scan_tc_fwd: the number of sequences (%d) must be a multiple of %d", S, BS);
  TcMap map = map_in;
  CUtensorMap tmU, tmH, tmHlo;
  int rc;
  // U <= 256: A operand in tensor memory unless DJ_FWD_ATM=0; U = 512: only when the residual pass needs the shared-memory slot
  const bool atm = (U <= 256) ? fwd_atm_enabled() : (Ut_lo != nullptr);
  DJ_CHECK_ARG(Ut_lo == nullptr || atm, "dj_lstm_scan_tc_fwd: the residual pass (Ut_lo) needs the tensor-memory A operand (DJ_FWD_ATM=0 is set)");
  // the shared-memory A slot holds U^T (A not in tensor memory) or the residual U^T - hi(U^T)
  if ((rc = make_map_2d(&tmU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (atm && Ut_lo) ? Ut_lo : Ut_bf, (uint64_t)U, (uint64_t)4 * U,
                        (uint64_t)U, 64, 128)))
    return rc;
  constexpr int RH = BS / 2;
  if (TIME) {   // hprev viewed as [b][t*48+n][U]
    const uint64_t rows_per_b = (uint64_t)map.outer_stride, B = (uint64_t)(S / 48);
    const uint64_t dims[3] = {(uint64_t)U, rows_per_b, B}, str[2] = {(uint64_t)U, rows_per_b * U};
    const uint32_t box[3] = {64, (uint32_t)(RH <= 48 ? RH : 48), (uint32_t)(RH <= 48 ? 1 : RH / 48)};
    if ((rc = make_map(&tmH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hprev, 3, dims, str, box))) return rc;
    tmHlo = tmH;
    if (INFER && (rc = make_map(&tmHlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hprev_lo, 3, dims, str, box))) return rc;
    map.step1 = 48; map.off1 = (RH % 48); map.base2 = BS >= 48 ? BS / 48 : 1; map.off2 = RH / 48;
    map.seq_stride = map.inner_stride;
  } else {           // hprev viewed as [seq][n][U]
    const uint64_t dims[3] = {(uint64_t)U, 48, (uint64_t)S}, str[2] = {(uint64_t)U, (uint64_t)48 * U};
    const uint32_t box[3] = {64, 1, (uint32_t)RH};
    if ((rc = make_map(&tmH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hprev, 3, dims, str, box))) return rc;
    tmHlo = tmH;
    if (INFER && (rc = make_map(&tmHlo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hprev_lo, 3, dims, str, box))) return rc;
    map.step1 = 1; map.off1 = 0; map.base2 = BS; map.off2 = RH;
    map.seq_stride = map.outer_stride;
  }
  // Shipped instances: A operand in tensor memory (U <= 256 always; U = 512 when the residual pass needs the
  // shared-memory slot, else A in shared memory).  The other placement is an experiment (-DDJ_EXPERIMENTS).
  constexpr bool SHIP_SMEM_A = (U > 256) && !INFER;
  auto kernel = scan_tc_fwd_kernel<U, BS, TIME, HARD, NB, NS, !SHIP_SMEM_A, INFER>;
  if constexpr (INFER) {
    DJ_CHECK_ARG(atm && Ut_lo && hprev_lo, "dj_lstm_scan_tc_infer: needs Ut_lo, h_lo and the tensor-memory A operand");
  } else if constexpr (SHIP_SMEM_A) {
    if (atm) kernel = scan_tc_fwd_kernel<U, BS, TIME, HARD, NB, NS, true, false>;
  } else {
#ifdef DJ_EXPERIMENTS
    if (!atm) kernel = scan_tc_fwd_kernel<U, BS, TIME, HARD, NB, NS, false, false>;
#else
    DJ_CHECK_ARG(atm, "dj_lstm_scan_tc_fwd: DJ_FWD_ATM=0 needs a build with -DDJ_EXPERIMENTS");
#endif
  }
  DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
  if (C > 8) DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((S / BS) * C);   // one cluster per tile; the hardware keeps 2 CTAs resident per SM
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SM::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (getenv("DJ_DEBUG_OCC")) {   // how many clusters the driver can keep resident vs how many tiles there are
    int n = 0;
    cudaOccupancyMaxActiveClusters(&n, (const void*)kernel, &cfg);
    fprintf(stderr, "scan_tc_fwd<%d,%d>: %d clusters of %d launched, %d can be resident\n", U, BS, S / BS, C, n);
  }
  uint16_t* hp = (uint16_t*)hprev;
  const uint32_t* utw = (const uint32_t*)Ut_bf;
  const int has_lo = Ut_lo != nullptr ? 1 : 0;
  uint16_t* hpl = (uint16_t*)hprev_lo;
  uint2* g16 = (uint2*)gates16;
  DJ_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmU, tmH, tmHlo, Z, g16, h_out, c_out, hp, hpl, utw, steps, map, f16, has_lo, acc_scale));
  return 0;
}

// prefetch depth of the x.W pre-activations, in 16-sequence chunks (DJ_FWD_NB overrides for experiments)
[[maybe_unused]] inline int fwd_prefetch_depth(int nch, int dflt) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("DJ_FWD_NB");
    env = e ? atoi(e) : 0;
  }
  const int nb = env > 0 ? env : dflt;
  return (nb >= 1 && nb <= 2 && nch % nb == 0) ? nb : dflt;
}

// DJ_FWD_NS=1 forces the unsplit tile (experiments)
[[maybe_unused]] inline bool fwd_split_enabled() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("DJ_FWD_NS");
    env = (e && atoi(e) == 1) ? 0 : 1;
  }
  return env != 0;
}

template <int U, int BS, bool TIME>
int launch_tc_fwd(const void* Ut_bf, const void* Ut_lo, int f16, const float* Z, void* gates16, float* h_out, float* c_out,
                  void* hprev, int S, int steps, const TcMap& map, int /*axis_time*/, int hard, cudaStream_t st) {
  constexpr int NCH = BS / 16;
  constexpr bool CAN_SPLIT = (NCH % 2 == 0);
  constexpr int NB_DEF = (NCH % 2 == 0) ? 2 : 1;
#ifndef DJ_EXPERIMENTS
  // shipped: two half-tiles when the tile is an even number of chunks, prefetch depth 2 then (else 1)
  return hard ? launch_tc_fwd_inst<U, BS, TIME, true, NB_DEF, CAN_SPLIT ? 2 : 1>(Ut_bf, Ut_lo, f16, Z, gates16, h_out, c_out, hprev, S, steps, map, st)
              : launch_tc_fwd_inst<U, BS, TIME, false, NB_DEF, CAN_SPLIT ? 2 : 1>(Ut_bf, Ut_lo, f16, Z, gates16, h_out, c_out, hprev, S, steps, map, st);
#else
  const int nb = fwd_prefetch_depth(NCH, NB_DEF);
  const bool split = CAN_SPLIT && fwd_split_enabled();
#define DJ_FWD_CASE(NBV, NSV)                                                                                       \
  if constexpr (NCH % NBV == 0 && (NSV == 1 || CAN_SPLIT)) {                                                        \
    if (nb == NBV && split == (NSV == 2))                                                                           \
      return hard ? launch_tc_fwd_inst<U, BS, TIME, true, NBV, NSV>(Ut_bf, Ut_lo, f16, Z, gates16, h_out, c_out, hprev, S, steps, map, st) \
                  : launch_tc_fwd_inst<U, BS, TIME, false, NBV, NSV>(Ut_bf, Ut_lo, f16, Z, gates16, h_out, c_out, hprev, S, steps, map, st); \
  }
  DJ_FWD_CASE(1, 1)
  DJ_FWD_CASE(1, 2)
  DJ_FWD_CASE(2, 1)
  DJ_FWD_CASE(2, 2)
#undef DJ_FWD_CASE
  DJ_CHECK_ARG(false, "dj_lstm_scan_tc_fwd: no kernel instance for prefetch depth %d", nb);
  return -1;
#endif
}

// ---------------------------------------------------------------------------
// backward (reverse scan) on tcgen05.
//   CTA r of a C = U/64 cluster owns hidden units [64r, 64r+64): it does the gate
//   derivatives of those cells and produces dh_{t-1} for them:
//       D[k, s] = sum_j U[k, j] * dz_t[s, j]          (M = 64 units, N = BS, K = 4U)
//   A operand = 64 rows of U [U, 4U] bf16 (natural, gate-interleaved), resident.
//   B operand = the whole dz_t tile [BS x 4U] bf16, all-gathered every step by TMA
//   multicast from the dZ buffer in L2 (dZ must be written anyway for the weight
//   gradients).  M = 64 accumulators sit in lanes 0..15 of each TMEM quarter.
// ---------------------------------------------------------------------------
// UPC = hidden units owned by one CTA: 64 (M = 64 fully used) or 32 (U = 512: only 32 rows of U fit; the
// M = 64 MMA then reads 32 further rows of the next K atom, whose accumulator rows are never read).
// SHARED: the two half-tiles take turns in ONE operand staging buffer (HB rows), so a tile can be twice as large
template <int U, int BS, int UPC, bool SHARED = false>
struct TcBwdSmem {
  static constexpr int A_BYTES = UPC * 4 * U * 2;
  static constexpr int B_BYTES = (SHARED ? BS / 2 : BS) * 4 * U * 2;
  static constexpr int A_OFF = 0, B_OFF = A_BYTES, BAR_OFF = A_BYTES + B_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128 + 1024;
};

constexpr int TCB_THREADS = 32 * 10;   // warps 0 and 9: issuers of the two half-tiles; warps 1..8: epilogue

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t v[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}

// NS = 2 software-pipelines the tile as two half-tiles (HB = BS/2 sequences, MMA N = HB) that alternate through
// the same epilogue warps: while they compute the gate derivatives of half B, half A's publish -> all-gather ->
// 64-MMA chain is in flight.  Hand-offs are mbarriers: epilogue --done[hf]--> issuer --TMA multicast of this CTA's own
// dz columns--> bar_z[hf] of every CTA --MMA--> bar_acc[hf]; the MMA commit also multicasts to free[hf] of every CTA
// ("my operand tile may be overwritten").  A CTA's dz is only ever read from memory by that CTA's own TMA, so the
// chain contains no gpu-scope membar and no cluster barrier.
//
// SHARED (time axis at large batch: one wave instead of two).  Shared memory holds the resident U slice (128 KB)
// and ONE staging buffer of 48 sequences x 1024 gate columns (96 KB); with two buffers a cluster could own only 48
// sequences and 64 sequences per GPU needed 64 clusters where 33-37 are resident.  Here a cluster owns 96 sequences
// = two half-tiles of 48 (one batch element each) that ALTERNATE in the single staging buffer: the multicast of half
// B's dz waits until every CTA's MMAs over half A's dz have retired (free[A], signalled by the multicast
// tcgen05.commit) and vice versa.  The MMA time per half is set by re-reading the resident U slice, not by N, so
// 48-wide halves cost about what the 24-wide halves of the two-wave variant cost: per step and SM the tensor pipe
// does half the work of before.  The epilogue then owns 24 cells per lane: only ONE half's gate / c / dY operands
// are held in registers at a time (loaded from L2 right after the other half's epilogue; the lines are pulled into
// L2 by prefetches issued a whole half-step earlier).
template <int U, int BS, int UPC, bool AXIS_TIME, int NS, bool SHARED = false>
__global__ void __launch_bounds__(TCB_THREADS, (U == 128 && BS <= 32) ? 2 : 1)
scan_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmZ,
                   const uint2* __restrict__ G16, const float* __restrict__ Cst, const float* __restrict__ dY,
                   uint32_t ldY, dj_dropout d_y, __nv_bfloat16* __restrict__ dZ, float* __restrict__ db,
                   int steps, TcMap map, int hard, float* __restrict__ db_part) {
  dj_resolve(d_y);
  constexpr int C = U / UPC;            // cluster size
  constexpr int KA = 4 * U / 64;        // K atoms of the contraction (gate columns)
  constexpr int ATOM = UPC * 128;       // bytes of one resident K atom of U (UPC rows x 128 B)
  static_assert(UPC == 64 || UPC == 32, "units per CTA");
  constexpr int KPC = KA / C;           // atoms each CTA multicasts per step
  constexpr int HB = BS / NS;           // sequences per half-tile
  constexpr int WC = HB / 2;            // accumulator columns (sequences) per epilogue warp and half
  constexpr int CPL = HB / 4;           // cells per lane and half
  constexpr int HALF_BYTES = KA * HB * 128;   // one half of the B operand: [KA atoms][HB rows][128 B]
  constexpr uint32_t TMEM_COLS = BS <= 32 ? 32 : BS <= 64 ? 64 : BS <= 128 ? 128 : 256;
  static_assert(NS == 1 || NS == 2, "one tile or two half-tiles");
  static_assert(HB % 8 == 0 && WC % 4 == 0, "half-tiles are whole 8-row swizzle groups; columns load in 4-column pieces");
  static_assert(!SHARED || (NS == 2 && AXIS_TIME && HB == 48), "shared staging: two half-tiles of one batch element each");
  using SM = TcBwdSmem<U, BS, UPC, SHARED>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  // barriers: a | z[2] | acc[2] | done[2] | free[2] | tmem slot
  constexpr uint32_t BAR_Z = SM::BAR_OFF + 8, BAR_ACC = SM::BAR_OFF + 24, BAR_DONE = SM::BAR_OFF + 40,
                     BAR_FREE = SM::BAR_OFF + 56;
  const uint32_t bar_a = sbase + SM::BAR_OFF;
  uint32_t* tmem_slot = (uint32_t*)(smem + SM::BAR_OFF + 72);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tile = blockIdx.x / C;
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmU); prefetch_tmap(&tmZ);
      mbar_init(bar_a, 1);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        mbar_init(sbase + BAR_Z + 8 * hf, 1); mbar_init(sbase + BAR_ACC + 8 * hf, 1);
        mbar_init(sbase + BAR_DONE + 8 * hf, 8); mbar_init(sbase + BAR_FREE + 8 * hf, C);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster.sync();

  if (warp == 0 || warp == 9) {
    // ================= issuer + publisher of half-tile hf: dh_t = U . dz_{t+1}^T =================
    // the whole warp walks the loop converged; single-thread work sits under elect_one()
    const int hf = (warp == 0) ? 0 : 1;
    if (warp == 0 && elect_one()) {   // resident A operand: rows [UPC*rank, +UPC) of U, all 4U columns
      mbar_expect_tx(bar_a, SM::A_BYTES);
      for (int ja = 0; ja < KA; ++ja)
        tma_load_2d(sbase + SM::A_OFF + ja * ATOM, &tmU, bar_a, ja * 64, UPC * rank);
    }
    __syncwarp();
    const uint32_t bar_z = sbase + BAR_Z + 8 * hf, bar_acc = sbase + BAR_ACC + 8 * hf,
                   bar_done = sbase + BAR_DONE + 8 * hf, bar_free = sbase + BAR_FREE + 8 * hf;
    // the barrier that says "the staging I am about to overwrite is no longer being read": my own half's MMAs of the
    // previous round, or (shared staging) the OTHER half's MMAs, which were the last to read it
    const uint32_t bar_free_wait = SHARED ? sbase + BAR_FREE + 8 * (hf ^ 1) : bar_free;
    const uint32_t b_half = sbase + SM::B_OFF + (SHARED ? 0 : hf * HALF_BYTES);
    // TMA coordinates of this half-tile: rows (time axis: within the batch element) / sequences (note axis)
    int c_row0, c_outer;
    if constexpr (!AXIS_TIME) { c_row0 = tile * BS + hf * HB; c_outer = 0; }
    else if constexpr (BS <= 48) { c_row0 = (tile % (48 / BS)) * BS + hf * HB; c_outer = tile / (48 / BS); }
    else { c_row0 = 0; c_outer = tile * (BS / 48) + hf * (HB / 48); }       // a half-tile = whole batch elements
    if (hf < NS) {
      uint32_t par = 0;
      for (int t = steps - 1; t > 0; --t, par ^= 1u) {   // round: dz_t in, dh of step t-1 out
        mbar_wait(bar_done, par);                         // this CTA's epilogue warps stored (and proxy-fenced) their dz_t
        DJ_TR(t, 4 * hf + 0);
        // Nobody else reads this CTA's dz_t from memory: THIS CTA's TMA multicasts it into every CTA's operand tile,
        // so no gpu-scope release and no cluster-wide "published" round trip is needed -- only the guarantee that
        // every peer's MMAs of the previous round have finished reading the tile we are about to overwrite.
        // (orders TMA writes after MMA reads: no data to acquire)
        if constexpr (SHARED) {
          if (hf == 1) mbar_wait(bar_free_wait, par);                        // half A's MMAs of THIS round
          else if (t != steps - 1) mbar_wait(bar_free_wait, par ^ 1u);       // half B's MMAs of the previous round
        } else {
          if (t != steps - 1) mbar_wait(bar_free_wait, par ^ 1u);
        }
        DJ_TR(t, 4 * hf + 1);
        if (elect_one()) {
          mbar_expect_tx(bar_z, HALF_BYTES);
          // one 4-D box {64 columns, HB rows, KPC atoms, 1}: lands as KPC consecutive [HB x 128 B] swizzled atoms
          tma_load_4d_mc(b_half + rank * KPC * (HB * 128), &tmZ, bar_z, 0, AXIS_TIME ? t * 48 + c_row0 : c_row0,
                         rank * KPC, AXIS_TIME ? c_outer : t, (uint16_t)((1u << C) - 1u));
        }
        __syncwarp();
        if (t == steps - 1) mbar_wait(bar_a, 0);
        mbar_wait(bar_z, par);
        DJ_TR(t, 4 * hf + 2);
        tc_fence_after();
        if (elect_one()) {
          // both operands bf16: dz needs bf16's exponent range and kind::f16 cannot mix half with bf16
          constexpr uint32_t idesc = make_idesc(64, HB, 0, 0);
          const uint64_t adesc0 = make_smem_desc(sbase + SM::A_OFF, 16, 1024);
          const uint64_t bdesc0 = make_smem_desc(b_half, 16, 1024);
#pragma unroll
          for (int ja = 0; ja < KA; ++ja)
#pragma unroll
            for (int k = 0; k < 4; ++k)   // descriptor start addresses are in 16-byte units
              umma_bf16(tmem_base + (uint32_t)(hf * HB), adesc0 + (uint64_t)((ja * ATOM + k * 32) >> 4),
                        bdesc0 + (uint64_t)((ja * (HB * 128) + k * 32) >> 4), idesc, (ja | k) != 0);
          umma_commit(bar_acc);                                   // -> this CTA's epilogue
          umma_commit_mc(bar_free, (uint16_t)((1u << C) - 1u));   // -> every CTA: my operand tile may be rewritten
        }
        __syncwarp();
        DJ_TR(t, 4 * hf + 3);
      }
      // drain the peers' last multicast arrives before this CTA can exit
      if (steps > 1) mbar_wait(bar_free_wait, par ^ 1u);
    }
  } else {
    // ================= epilogue: gate derivatives =================
    // TMEM quarter q holds units 16q..16q+15 in its lanes 0..15 (M = 64 layout); the two warps of a
    // quarter split a half-tile's sequences, and lanes 16..31 take half of their warp's sequences by shuffle.
    const int q = warp & 3;
    const int w2 = (warp - 1) >> 2;
    const int sh = lane >> 4;
    const uint32_t col = UPC * rank + 16 * q + (lane & 15);       // global hidden unit
    const bool active = (16 * q < UPC);                            // UPC = 32: only TMEM quarters 0,1 hold real rows
    // the row strides of the canonical layout are fixed per axis (the ABI accepts only these two maps), so every
    // per-cell offset below is an immediate off ONE base address per array, half-tile and step
    constexpr uint32_t sstr = AXIS_TIME ? 1u : 48u, tstr = AXIS_TIME ? 48u : 1u;
    const uint32_t ystep = sstr * ldY;                             // dY floats between consecutive sequences
    uint32_t row00[NS];                                            // row of this lane's first sequence of each half, step 0
#pragma unroll
    for (int hf = 0; hf < NS; ++hf)   // a half-tile never straddles a batch element
      row00[hf] = (uint32_t)tc_row0(map, tile * BS + hf * HB) + (uint32_t)(w2 * WC + sh * CPL) * sstr;
    float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int LSN = SHARED ? 1 : NS;                           // operand sets held in registers at a time
    float dcn[NS][CPL], ct[NS][CPL], cpv[LSN][CPL], dyv[LSN][CPL];
    uint2 gv[LSN][CPL];                                            // four gates of a cell as IEEE half
    // Pure loads only: anything computed on a just-loaded value would serialise the loads (each use
    // waits for its own DRAM round trip).  Masks and the t==0 special case are applied at use time.
    auto issue_loads = [&](int hf, int t) {   // everything of (half hf, step t) that does not depend on the recurrence
      const uint32_t tp = (t > 0) ? (uint32_t)(t - 1) : 0u;      // clamped: row of c_{t-1} (ignored at t == 0)
      const int ls = SHARED ? 0 : hf;
      if (!active) return;
      const uint32_t r0 = row00[hf] + (uint32_t)t * tstr;
      const uint2* gp = G16 + (size_t)r0 * U + col;
      const float* cp = Cst + (size_t)(row00[hf] + tp * tstr) * U + col;
      const float* yp = dY + (size_t)r0 * ldY + col;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        gv[ls][j] = __ldg(gp + (size_t)j * (sstr * U));
        cpv[ls][j] = __ldg(cp + (size_t)j * (sstr * U));
        dyv[ls][j] = __ldg(yp + (size_t)j * ystep);
      }
    };
    // pull the lines issue_loads(hf, t) will read into L2: per cell the 16 lanes of a unit group read 256 B of gates
    // (lanes 0 and 8 of the group prefetch a 128-byte line each) and 64 B each of c and dY (lane 0 of the group)
    auto prefetch_l2 = [&](int hf, int t) {
      const uint32_t tp = (t > 0) ? (uint32_t)(t - 1) : 0u;
      if (!active) return;
      const uint32_t r0 = row00[hf] + (uint32_t)t * tstr;
      const uint2* gp = G16 + (size_t)r0 * U + col;
      const float* cp = Cst + (size_t)(row00[hf] + tp * tstr) * U + col;
      const float* yp = dY + (size_t)r0 * ldY + col;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        if ((lane & 15) == 0) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(gp + (size_t)j * (sstr * U)));
        if ((lane & 15) == 0) {
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(cp + (size_t)j * (sstr * U)));
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(yp + (size_t)j * ystep));
        }
      }
    };
#pragma unroll
    for (int hf = 0; hf < NS; ++hf) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        dcn[hf][j] = 0.f;
        ct[hf][j] = active ? Cst[(size_t)(row00[hf] + (uint32_t)(steps - 1) * tstr) * U + col + (size_t)j * (sstr * U)] : 0.f;
      }
    }
#pragma unroll
    for (int ls = 0; ls < LSN; ++ls) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) { gv[ls][j] = make_uint2(0u, 0u); cpv[ls][j] = 0.f; dyv[ls][j] = 0.f; }
      issue_loads(ls, steps - 1);
    }
    if constexpr (SHARED) prefetch_l2(1, steps - 1);
    const bool tr = (warp == 1 && lane == 0);
    uint32_t par = 0;
    for (int t = steps - 1; t >= 0; --t) {
#pragma unroll
      for (int hf = 0; hf < NS; ++hf) {
        float dh[CPL];
        if (t != steps - 1) {
          mbar_wait(sbase + BAR_ACC + 8 * hf, par ^ 1u);   // produced by issuer round t+1
          tc_fence_after();
          if (tr) DJ_TR(t, 8 + 3 * hf);
          uint32_t acc[WC];
#pragma unroll
          for (int p4 = 0; p4 < WC / 4; ++p4)
            tmem_ld4(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(hf * HB + w2 * WC + p4 * 4), acc + p4 * 4);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const float up = __shfl_sync(0xffffffffu, __uint_as_float(acc[CPL + j]), lane & 15);
            dh[j] = sh ? up : __uint_as_float(acc[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) dh[j] = 0.f;
        }
        const uint32_t r0 = row00[hf] + (uint32_t)t * tstr;
        const uint32_t e0 = r0 * U + col;               // dropout element index of the first cell (32-bit by contract)
        __nv_bfloat16* const zp = dZ + (size_t)r0 * (4 * U) + 4 * col;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          if (!active) break;
          constexpr int LS_ = SHARED ? 0 : -1;
          const int ls = (LS_ == 0) ? 0 : hf;
          const float4 g4 = unpack_gates16(gv[ls][j]);
          const float cprev = (t > 0) ? cpv[ls][j] : 0.f;
          const float dht = fmaf(dyv[ls][j], dj_dropmul(d_y, e0 + (uint32_t)j * (sstr * U)), dh[j]);
          const float tc = fast_tanh(ct[hf][j]);
          const float d_o = dht * tc;
          const float dc = fmaf(dht * g4.w, 1.f - tc * tc, dcn[hf][j]);
          dcn[hf][j] = dc * g4.y;
          const float dz0 = dc * g4.z * dj_gate_dact(g4.x, hard);
          const float dz1 = dc * cprev * dj_gate_dact(g4.y, hard);
          const float dz2 = dc * g4.x * (1.f - g4.z * g4.z);
          const float dz3 = d_o * dj_gate_dact(g4.w, hard);
          __nv_bfloat162 lo = __floats2bfloat162_rn(dz0, dz1), hi = __floats2bfloat162_rn(dz2, dz3);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(zp + (size_t)j * (sstr * 4 * U)) = pk;
          dbacc[0] += dz0; dbacc[1] += dz1; dbacc[2] += dz2; dbacc[3] += dz3;
          ct[hf][j] = cprev;                // c_{t-1} is the cell state of the next (earlier) step
        }
        if (tr) DJ_TR(t, 9 + 3 * hf);
        if (t > 0) {
          tc_fence_before();
          fence_proxy_async_all();          // dz_t (generic-proxy global stores) -> later TMA (async proxy) reads
          __syncwarp();
          if (lane == 0) mbar_arrive(sbase + BAR_DONE + 8 * hf);   // release.cta: hands this warp's stores to the issuer
          if constexpr (!SHARED) issue_loads(hf, t - 1);   // land during this half's publish / TMA / MMA chain
        }
        if constexpr (SHARED) {
          // one operand set in registers: load what the NEXT epilogue in program order needs (its lines were
          // prefetched into L2 one epilogue ago), and prefetch for the one after it
          if (hf == 0) { issue_loads(1, t); if (t > 0) prefetch_l2(0, t - 1); }
          else if (t > 0) { issue_loads(0, t - 1); prefetch_l2(1, t - 1); }
        }
        if (tr) DJ_TR(t, 10 + 3 * hf);
      }
      par ^= 1u;
    }
    if (active && db_part != nullptr) {
      // deterministic mode: one partial per (tile, sequence group of the CTA), added later in index order; the four
      // sequence groups (w2, sh) of a CTA and the CTAs of a cluster together cover every column exactly once
      float* dst = db_part + ((size_t)(tile * 4 + w2 * 2 + sh) * (4 * U)) + 4 * col;
      *reinterpret_cast<float4*>(dst) = make_float4(dbacc[0], dbacc[1], dbacc[2], dbacc[3]);
    } else if (active) {
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) atomicAdd(db + 4 * col + gq, dbacc[gq]);
    }
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

#ifdef DJ_EXPERIMENTS
// ---------------------------------------------------------------------------
// reverse scan on CTA PAIRS (tcgen05.mma.cta_group::2).
// The reverse scan above is bound by every SM ingesting the whole dz half-tile each step (2 KB per sequence), so a
// larger tile does not help it.  Here the C = U/64 CTAs of a cluster form C/2 pairs; a pair runs ONE M = 128 MMA
// (each CTA its own 64 rows of U), and the B operand -- the dz half-tile -- is split between the two CTAs of a pair
// by sequence: CTA 2p stages sequences [0, NH) of the half-tile, CTA 2p+1 sequences [NH, 2 NH).  Each CTA therefore
// stages and ingests HALF of what it did, so a tile twice as large (96 time-axis sequences = two batch elements,
// 64 note-axis sequences) costs the same shared memory and the same ingest per step, and a layer that needed two
// waves of clusters runs as one.  Per half-step:
//   epilogue --done[hf]--> issuer (every CTA): wait free[hf] (all pairs have read the previous round's tiles), two
//   multicast TMAs of this CTA's own dz columns (sequences [0,NH) -> the even CTAs, [NH,2NH) -> the odd CTAs)
//   --> z[hf] of the destination pair's LEADER (cta_group::2 TMA: loads landing in the odd CTA signal the even one, so
//   nothing is forwarded); the leader issues the pair's 4U/16 MMAs and commits to acc[hf] of its pair and to free[hf]
//   of every CTA.
// Accumulator (M = 128 pair layout, dj_tc.cuh): unit u of this CTA in lane u (sequences of the leader's half) and
// lane 64 + u (the peer's half), columns [hf*NH, +NH).  The 8 epilogue warps: quarter q = lanes [32q, +32) -> units
// 32*(q&1) + lane, sequence side q>>1; the two warps of a quarter split the NH columns.  Every lane owns CPL = NH/2
// cells per half-tile; one operand set (gates, c, dY of the NEXT epilogue in program order) is held in registers.
// ---------------------------------------------------------------------------
template <int U, int BS>
struct TcBwdPairSmem {
  static constexpr int A_BYTES = 64 * 4 * U * 2;
  static constexpr int B_BYTES = (BS / 2) * 4 * U * 2;      // two half-tiles x NH = BS/4 sequences x 4U columns
  static constexpr int A_OFF = 0, B_OFF = A_BYTES, BAR_OFF = A_BYTES + B_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128 + 1024;
};

// DIRECT: operand loads signal the leader's barriers (cta_group::2 TMA); else the odd CTA waits for its own half and
// forwards one remote arrive to its leader (measured faster for the single-pair note-axis clusters).
template <int U, int BS, bool AXIS_TIME, int LSN, bool DIRECT>
__global__ void __launch_bounds__(TCB_THREADS, (U == 128) ? 2 : 1)
scan_tc_bwd_pair_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmZ,
                        const uint2* __restrict__ G16, const float* __restrict__ Cst, const float* __restrict__ dY,
                        uint32_t ldY, dj_dropout d_y, __nv_bfloat16* __restrict__ dZ, float* __restrict__ db,
                        int steps, TcMap map, int hard, float* __restrict__ db_part) {
  dj_resolve(d_y);
  constexpr int C = U / 64;             // cluster size (CTAs), C/2 pairs
  constexpr int KA = 4 * U / 64;        // K atoms of the contraction (gate columns)
  constexpr int ATOM = 64 * 128;        // bytes of one resident K atom of U (64 rows x 128 B)
  constexpr int KPC = KA / C;           // atoms each CTA multicasts per step
  constexpr int HB = BS / 2;            // sequences per half-tile (= MMA N)
  constexpr int NH = HB / 2;            // sequences of a half-tile staged in one CTA of a pair
  constexpr int CPL = NH / 2;           // cells per lane and half-tile
  constexpr int HALF_BYTES = KA * NH * 128;   // one CTA's share of a half-tile's B operand: [KA atoms][NH rows][128 B]
  constexpr uint32_t TMEM_COLS = 2 * NH <= 32 ? 32 : 64;
  // row strides of the canonical layout (row = (b*T + t)*48 + n) are fixed per axis, so every per-cell offset below
  // is an immediate: time axis = sequences (b, n) one row apart, steps 48 rows apart; note axis the other way round
  constexpr uint32_t SSTR = AXIS_TIME ? 1u : 48u, TSTR = AXIS_TIME ? 48u : 1u;
  static_assert(C == 2 || C == 4, "one or two pairs per cluster");
  static_assert(NH % 8 == 0 && CPL % 4 == 0 && 2 * NH <= 64, "whole swizzle groups; columns load in 4-column pieces");
  static_assert(LSN == 1 || LSN == 2, "operand sets held in registers");
  using SM = TcBwdPairSmem<U, BS>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  // barriers: a | z[2] | acc[2] | done[2] | free[2] | peer[2] | tmem slot
  constexpr uint32_t BAR_Z = SM::BAR_OFF + 8, BAR_ACC = SM::BAR_OFF + 24, BAR_DONE = SM::BAR_OFF + 40,
                     BAR_FREE = SM::BAR_OFF + 56, BAR_PEER = SM::BAR_OFF + 72;
  const uint32_t bar_a = sbase + SM::BAR_OFF;
  uint32_t* tmem_slot = (uint32_t*)(smem + SM::BAR_OFF + 88);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const bool leader = (rank & 1) == 0;
  const int tile = blockIdx.x / C;
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmU); prefetch_tmap(&tmZ);
      mbar_init(bar_a, 1);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        mbar_init(sbase + BAR_Z + 8 * hf, 1); mbar_init(sbase + BAR_ACC + 8 * hf, 1);
        mbar_init(sbase + BAR_DONE + 8 * hf, 8); mbar_init(sbase + BAR_FREE + 8 * hf, C / 2);
        mbar_init(sbase + BAR_PEER + 8 * hf, 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc2(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster.sync();

  if (warp == 0 || warp == 9) {
    // ================= issuer of half-tile hf (every CTA multicasts; the leader of a pair also multiplies) =================
    // Every operand load is a cta_group::2 TMA whose mbarrier operand names the LEADER's barrier (dj_tc.cuh), so the
    // leader's a / z[hf] barriers count the bytes landing in both CTAs of the pair and the odd CTA never waits for,
    // or forwards, anything.
    const int hf = (warp == 0) ? 0 : 1;
    if (warp == 0 && elect_one()) {   // resident A operand: rows [64*rank, +64) of U, all 4U columns
      if (DIRECT) {
        if (leader) mbar_expect_tx(bar_a, 2 * SM::A_BYTES);
        for (int ja = 0; ja < KA; ++ja)
          tma_load_2d_pair(sbase + SM::A_OFF + ja * ATOM, &tmU, bar_a, ja * 64, 64 * rank);
      } else {
        mbar_expect_tx(bar_a, SM::A_BYTES);
        for (int ja = 0; ja < KA; ++ja)
          tma_load_2d(sbase + SM::A_OFF + ja * ATOM, &tmU, bar_a, ja * 64, 64 * rank);
      }
    }
    __syncwarp();
    const uint32_t bar_z = sbase + BAR_Z + 8 * hf, bar_acc = sbase + BAR_ACC + 8 * hf,
                   bar_done = sbase + BAR_DONE + 8 * hf, bar_free = sbase + BAR_FREE + 8 * hf,
                   bar_peer = sbase + BAR_PEER + 8 * hf;
    const uint32_t b_half = sbase + SM::B_OFF + hf * HALF_BYTES;
    constexpr uint16_t MASK_ALL = (uint16_t)((1u << C) - 1u), MASK_EVEN = (uint16_t)(0x5555u & MASK_ALL),
                       MASK_ODD = (uint16_t)(0xAAAAu & MASK_ALL);
    const uint16_t mask_pair = (uint16_t)(3u << (rank & ~1));
    // TMA coordinates of this half-tile: rows (time axis: one batch element per half-tile) / sequences (note axis)
    int c_row0, c_outer;
    if constexpr (AXIS_TIME) { c_row0 = 0; c_outer = tile * 2 + hf; }
    else { c_row0 = tile * BS + hf * HB; c_outer = 0; }
    uint32_t par = 0;
    for (int t = steps - 1; t > 0; --t, par ^= 1u) {   // round: dz_t in, dh of step t-1 out
      mbar_wait(bar_done, par);                         // this CTA's epilogue warps stored (and proxy-fenced) their dz_t
      DJ_TR(t, 4 * hf + 0);
      if (t != steps - 1) mbar_wait(bar_free, par ^ 1u);   // every pair's MMAs of the previous round have read their tiles
      DJ_TR(t, 4 * hf + 1);
      if (elect_one()) {
        // this CTA's KPC atoms of dz_t, as {64 columns, NH rows, KPC atoms} boxes: sequences [0, NH) of the half-tile
        // to the even CTAs, [NH, 2 NH) to the odd CTAs
        const int r1 = AXIS_TIME ? t * 48 + c_row0 : c_row0, r3 = AXIS_TIME ? c_outer : t;
        if (DIRECT) {
          if (leader) mbar_expect_tx(bar_z, 2 * HALF_BYTES);
          tma_load_4d_mc_pair(b_half + rank * KPC * (NH * 128), &tmZ, bar_z, 0, r1, rank * KPC, r3, MASK_EVEN);
          tma_load_4d_mc_pair(b_half + rank * KPC * (NH * 128), &tmZ, bar_z, 0, r1 + NH, rank * KPC, r3, MASK_ODD);
        } else {
          mbar_expect_tx(bar_z, HALF_BYTES);
          tma_load_4d_mc(b_half + rank * KPC * (NH * 128), &tmZ, bar_z, 0, r1, rank * KPC, r3, MASK_EVEN);
          tma_load_4d_mc(b_half + rank * KPC * (NH * 128), &tmZ, bar_z, 0, r1 + NH, rank * KPC, r3, MASK_ODD);
        }
      }
      __syncwarp();
      if (!DIRECT && !leader) {       // forwarded signalling: the odd CTA tells its leader that its half landed
        if (t == steps - 1) mbar_wait(bar_a, 0);
        mbar_wait(bar_z, par);
        if (elect_one()) mbar_arrive_remote(bar_peer, (uint32_t)(rank & ~1));   // release.cluster
        __syncwarp();
      }
      if (leader) {
        if (t == steps - 1) mbar_wait(bar_a, 0);
        mbar_wait(bar_z, par);                          // all C CTAs' columns of the pair's sequences landed
        if (!DIRECT) mbar_wait_cluster(bar_peer, par);
        DJ_TR(t, 4 * hf + 2);
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = make_idesc(128, HB, 0, 0);
          const uint64_t adesc0 = make_smem_desc(sbase + SM::A_OFF, 16, 1024);
          const uint64_t bdesc0 = make_smem_desc(b_half, 16, 1024);
#pragma unroll
          for (int ja = 0; ja < KA; ++ja)
#pragma unroll
            for (int k = 0; k < 4; ++k)   // descriptor start addresses are in 16-byte units
              umma2_bf16(tmem_base + (uint32_t)(hf * NH), adesc0 + (uint64_t)((ja * ATOM + k * 32) >> 4),
                         bdesc0 + (uint64_t)((ja * (NH * 128) + k * 32) >> 4), idesc, (ja | k) != 0);
          umma2_commit_mc(bar_acc, mask_pair);          // -> the epilogues of both CTAs of the pair
          umma2_commit_mc(bar_free, MASK_ALL);          // -> every CTA: this pair's operand tiles may be rewritten
        }
        __syncwarp();
      }
      DJ_TR(t, 4 * hf + 3);
    }
    // drain the last multicast arrives on this CTA's barriers before it can exit
    if (steps > 1) mbar_wait(bar_free, par ^ 1u);
  } else {
    // ================= epilogue: gate derivatives =================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int w2 = (warp - 1) >> 2;                  // which half of the NH columns
    const int side = q >> 1;                         // sequences staged in the leader (0) or in the peer (1)
    const uint32_t col = 64 * rank + 32 * (q & 1) + lane;          // global hidden unit
    uint32_t row00[2];                               // row of this lane's first sequence of each half-tile, step 0
#pragma unroll
    for (int hf = 0; hf < 2; ++hf)
      row00[hf] = (uint32_t)tc_row0(map, tile * BS + hf * HB) + (uint32_t)(side * NH + w2 * CPL) * SSTR;
    const uint32_t ystep = SSTR * ldY;               // dY floats between consecutive sequences
    float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
    float dcn[2][CPL], ct[2][CPL], cpv[LSN][CPL], dyv[LSN][CPL];
    uint2 gv[LSN][CPL];                              // four gates of a cell as IEEE half
    auto issue_loads = [&](int hf, int t) {          // everything of (half hf, step t) that does not depend on the recurrence
      const int ls = (LSN == 1) ? 0 : hf;
      const uint32_t r0 = row00[hf] + (uint32_t)t * TSTR;
      const uint32_t rp = (t > 0) ? r0 - TSTR : r0;  // row of c_{t-1} (clamped; ignored at t == 0)
      const uint2* gp = G16 + (size_t)r0 * U + col;
      const float* cp = Cst + (size_t)rp * U + col;
      const float* yp = dY + (size_t)r0 * ldY + col;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        gv[ls][j] = __ldg(gp + (size_t)j * (SSTR * U));
        cpv[ls][j] = __ldg(cp + (size_t)j * (SSTR * U));
        dyv[ls][j] = __ldg(yp + (size_t)j * ystep);
      }
    };
    // pull the lines issue_loads(hf, t) will read into L2: a warp reads 256 B of gates and 128 B each of c and dY per cell
    auto prefetch_l2 = [&](int hf, int t) {
      const uint32_t r0 = row00[hf] + (uint32_t)t * TSTR;
      const uint32_t rp = (t > 0) ? r0 - TSTR : r0;
      const uint2* gp = G16 + (size_t)r0 * U + col;
      const float* cp = Cst + (size_t)rp * U + col;
      const float* yp = dY + (size_t)r0 * ldY + col;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        if ((lane & 15) == 0) asm volatile("prefetch.global.L2 [%0];\n" ::"l"(gp + (size_t)j * (SSTR * U)));
        if (lane == 0) {
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(cp + (size_t)j * (SSTR * U)));
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(yp + (size_t)j * ystep));
        }
      }
    };
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float* cp = Cst + (size_t)(row00[hf] + (uint32_t)(steps - 1) * TSTR) * U + col;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        dcn[hf][j] = 0.f;
        ct[hf][j] = cp[(size_t)j * (SSTR * U)];
      }
    }
#pragma unroll
    for (int ls = 0; ls < LSN; ++ls) issue_loads(ls, steps - 1);
    if constexpr (LSN == 1) prefetch_l2(1, steps - 1);
    uint32_t par = 0;
    const bool tr = (warp == 1 && lane == 0);
    for (int t = steps - 1; t >= 0; --t) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float dh[CPL];
        if (t != steps - 1) {
          mbar_wait(sbase + BAR_ACC + 8 * hf, par ^ 1u);   // produced by issuer round t+1
          tc_fence_after();
          if (tr) DJ_TR(t, 8 + 3 * hf);
          uint32_t acc[CPL];
#pragma unroll
          for (int p4 = 0; p4 < CPL / 4; ++p4)
            tmem_ld4(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(hf * NH + w2 * CPL + p4 * 4), acc + p4 * 4);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CPL; ++j) dh[j] = __uint_as_float(acc[j]);
        } else {
#pragma unroll
          for (int j = 0; j < CPL; ++j) dh[j] = 0.f;
        }
        const int ls = (LSN == 1) ? 0 : hf;
        const uint32_t r0 = row00[hf] + (uint32_t)t * TSTR;
        const uint32_t e0 = r0 * U + col;               // dropout element index of the first cell (32-bit by contract)
        __nv_bfloat16* const zp = dZ + (size_t)r0 * (4 * U) + 4 * col;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const float4 g4 = unpack_gates16(gv[ls][j]);
          const float cprev = (t > 0) ? cpv[ls][j] : 0.f;
          const float dht = fmaf(dyv[ls][j], dj_dropmul(d_y, e0 + (uint32_t)j * (SSTR * U)), dh[j]);
          const float tc = fast_tanh(ct[hf][j]);
          const float d_o = dht * tc;
          const float dc = fmaf(dht * g4.w, 1.f - tc * tc, dcn[hf][j]);
          dcn[hf][j] = dc * g4.y;
          const float dz0 = dc * g4.z * dj_gate_dact(g4.x, hard);
          const float dz1 = dc * cprev * dj_gate_dact(g4.y, hard);
          const float dz2 = dc * g4.x * (1.f - g4.z * g4.z);
          const float dz3 = d_o * dj_gate_dact(g4.w, hard);
          __nv_bfloat162 lo = __floats2bfloat162_rn(dz0, dz1), hi = __floats2bfloat162_rn(dz2, dz3);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&lo);
          pk.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(zp + (size_t)j * (SSTR * 4 * U)) = pk;
          dbacc[0] += dz0; dbacc[1] += dz1; dbacc[2] += dz2; dbacc[3] += dz3;
          ct[hf][j] = cprev;                // c_{t-1} is the cell state of the next (earlier) step
        }
        if (tr) DJ_TR(t, 9 + 3 * hf);
        if (t > 0) {
          tc_fence_before();
          fence_proxy_async_all();          // dz_t (generic-proxy global stores) -> later TMA (async proxy) reads
          __syncwarp();
          if (lane == 0) mbar_arrive(sbase + BAR_DONE + 8 * hf);   // release.cta: hands this warp's stores to the issuer
          if constexpr (LSN == 2) issue_loads(hf, t - 1);   // land during this half's publish / TMA / MMA chain
        }
        if (tr) DJ_TR(t, 10 + 3 * hf);
        if constexpr (LSN == 1) {
          // one operand set in registers: load what the NEXT epilogue in program order needs (its lines were
          // prefetched into L2 one epilogue ago), and prefetch for the one after it
          if (hf == 0) { issue_loads(1, t); if (t > 0) prefetch_l2(0, t - 1); }
          else if (t > 0) { issue_loads(0, t - 1); prefetch_l2(1, t - 1); }
        }
      }
      par ^= 1u;
    }
    if (db_part != nullptr) {
      // deterministic mode: one partial per (tile, sequence group of the CTA), added later in index order; the four
      // sequence groups (side, w2) of a CTA and the CTAs of a cluster together cover every column exactly once
      float* dst = db_part + ((size_t)(tile * 4 + side * 2 + w2) * (4 * U)) + 4 * col;
      *reinterpret_cast<float4*>(dst) = make_float4(dbacc[0], dbacc[1], dbacc[2], dbacc[3]);
    } else {
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) atomicAdd(db + 4 * col + gq, dbacc[gq]);
    }
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc2(tmem_base, TMEM_COLS);
}

#endif  // DJ_EXPERIMENTS (pair reverse scan)

[[maybe_unused]] inline bool bwd_split_enabled() {   // DJ_BWD_NS=1 forces the unsplit tile (experiments)
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("DJ_BWD_NS");
    env = (e && atoi(e) == 1) ? 0 : 1;
  }
  return env != 0;
}

template <int U, int BS, int UPC, bool AXIS_TIME, int NS, bool SHARED = false>
int launch_tc_bwd_inst(const void* Un_bf, const void* gates, const float* c, const float* dY, int64_t ldY, dj_dropout d_y,
                       void* dZ, float* db, int S, int steps, const TcMap& map_in, int hard, cudaStream_t st) {
  constexpr int C = U / UPC;
  constexpr int HB = BS / NS;
  using SM = TcBwdSmem<U, BS, UPC, SHARED>;
  DJ_CHECK_ARG(S % BS == 0, "dj_lstm_scan_tc_bwd: the number of sequences (%d) must be a multiple of %d", S, BS);
  TcMap map = map_in;
  CUtensorMap tmU, tmZ;
  int rc;
  if ((rc = make_map_2d(&tmU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Un_bf, (uint64_t)4 * U, (uint64_t)U, (uint64_t)4 * U, 64, UPC)))
    return rc;
  constexpr int KA = 4 * U / 64, KPC = KA / C;
  if (AXIS_TIME) {   // dZ viewed as [b][t*48+n][atom][64]; a tile is BS consecutive notes of one batch element
    static_assert(!AXIS_TIME || 48 % BS == 0 || (BS % 48 == 0 && HB % 48 == 0),
                  "time-axis tiles divide a batch element, or their halves are whole batch elements");
    const uint64_t rows_per_b = (uint64_t)map.outer_stride, B = (uint64_t)(S / 48);
    const uint64_t dims[4] = {64, rows_per_b, (uint64_t)KA, B}, str[3] = {(uint64_t)4 * U, 64, rows_per_b * 4 * U};
    const uint32_t box[4] = {64, (uint32_t)HB, (uint32_t)KPC, 1};
    if ((rc = make_map(&tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dZ, 4, dims, str, box))) return rc;
    map.step1 = 48; map.off1 = BS; map.base2 = 1; map.off2 = BS <= 48 ? 48 / BS : 1;
    map.seq_stride = map.inner_stride;
  } else {           // dZ viewed as [n][atom][seq][64] (strides: seq 48*4U, atom 64, n 4U)
    const uint64_t dims[4] = {64, (uint64_t)S, (uint64_t)KA, 48}, str[3] = {(uint64_t)48 * 4 * U, 64, (uint64_t)4 * U};
    const uint32_t box[4] = {64, (uint32_t)HB, (uint32_t)KPC, 1};
    if ((rc = make_map(&tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dZ, 4, dims, str, box))) return rc;
    map.step1 = 1; map.off1 = 0; map.base2 = BS; map.off2 = 1;
    map.seq_stride = map.outer_stride;
  }
  auto kernel = scan_tc_bwd_kernel<U, BS, UPC, AXIS_TIME, NS, SHARED>;
  DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
  if (C > 8) DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((S / BS) * C);   // one cluster per tile
  cfg.blockDim = dim3(TCB_THREADS);
  cfg.dynamicSmemBytes = SM::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (getenv("DJ_DEBUG_OCC")) {
    int n = 0;
    cudaOccupancyMaxActiveClusters(&n, (const void*)kernel, &cfg);
    fprintf(stderr, "scan_tc_bwd<%d,%d>: %d clusters of %d launched, %d can be resident\n", U, BS, S / BS, C, n);
  }
  __nv_bfloat16* dzp = (__nv_bfloat16*)dZ;
  const uint32_t ldy32 = (uint32_t)ldY;
  const uint2* g16 = (const uint2*)gates;
  const int tiles = S / BS;
  float* db_part = nullptr;
  if (dj_reduce_workspace((void*)st, (int64_t)tiles * 4 * 4 * U, &db_part)) return -1;
  DJ_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmU, tmZ, g16, c, dY, ldy32, d_y, dzp, db, steps, map, hard, db_part));
  if (db_part != nullptr) return dj_ordered_reduce(db_part, tiles * 4, (int64_t)4 * U, 1, 4 * U, 4 * U, db, 4 * U, (void*)st);
  return 0;
}

#ifdef DJ_EXPERIMENTS
inline bool bwd_pair_enabled() {   // DJ_BWD_PAIR=1: the note-axis reverse scan on CTA pairs
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("DJ_BWD_PAIR");
    env = (e && atoi(e) == 1) ? 1 : 0;
  }
  return env != 0;
}

template <int U, int BS, bool AXIS_TIME, int LSN, bool DIRECT>
int launch_tc_bwd_pair(const void* Un_bf, const void* gates, const float* c, const float* dY, int64_t ldY, dj_dropout d_y,
                       void* dZ, float* db, int S, int steps, const TcMap& map_in, int hard, cudaStream_t st) {
  constexpr int C = U / 64, KA = 4 * U / 64, KPC = KA / C, NH = BS / 4;
  using SM = TcBwdPairSmem<U, BS>;
  DJ_CHECK_ARG(S % BS == 0, "dj_lstm_scan_tc_bwd: the number of sequences (%d) must be a multiple of %d", S, BS);
  static_assert(!AXIS_TIME || BS == 96, "time-axis pair tiles are two batch elements");
  TcMap map = map_in;
  CUtensorMap tmU, tmZ;
  int rc;
  if ((rc = make_map_2d(&tmU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Un_bf, (uint64_t)4 * U, (uint64_t)U, (uint64_t)4 * U, 64, 64)))
    return rc;
  if (AXIS_TIME) {   // dZ viewed as [b][t*48+n][atom][64]; a half-tile is one batch element
    const uint64_t rows_per_b = (uint64_t)map.outer_stride, B = (uint64_t)(S / 48);
    const uint64_t dims[4] = {64, rows_per_b, (uint64_t)KA, B}, str[3] = {(uint64_t)4 * U, 64, rows_per_b * 4 * U};
    const uint32_t box[4] = {64, (uint32_t)NH, (uint32_t)KPC, 1};
    if ((rc = make_map(&tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dZ, 4, dims, str, box))) return rc;
    map.seq_stride = map.inner_stride;
  } else {           // dZ viewed as [n][atom][seq][64] (strides: seq 48*4U, atom 64, n 4U)
    const uint64_t dims[4] = {64, (uint64_t)S, (uint64_t)KA, 48}, str[3] = {(uint64_t)48 * 4 * U, 64, (uint64_t)4 * U};
    const uint32_t box[4] = {64, (uint32_t)NH, (uint32_t)KPC, 1};
    if ((rc = make_map(&tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dZ, 4, dims, str, box))) return rc;
    map.seq_stride = map.outer_stride;
  }
  auto kernel = scan_tc_bwd_pair_kernel<U, BS, AXIS_TIME, LSN, DIRECT>;
  DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((S / BS) * C);   // one cluster per tile
  cfg.blockDim = dim3(TCB_THREADS);
  cfg.dynamicSmemBytes = SM::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (getenv("DJ_DEBUG_OCC")) {
    int n = 0;
    cudaOccupancyMaxActiveClusters(&n, (const void*)kernel, &cfg);
    fprintf(stderr, "scan_tc_bwd_pair<%d,%d>: %d clusters of %d launched, %d can be resident\n", U, BS, S / BS, C, n);
  }
  __nv_bfloat16* dzp = (__nv_bfloat16*)dZ;
  const uint32_t ldy32 = (uint32_t)ldY;
  const uint2* g16 = (const uint2*)gates;
  const int tiles = S / BS;
  float* db_part = nullptr;
  if (dj_reduce_workspace((void*)st, (int64_t)tiles * 4 * 4 * U, &db_part)) return -1;
  DJ_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmU, tmZ, g16, c, dY, ldy32, d_y, dzp, db, steps, map, hard, db_part));
  if (db_part != nullptr) return dj_ordered_reduce(db_part, tiles * 4, (int64_t)4 * U, 1, 4 * U, 4 * U, db, 4 * U, (void*)st);
  return 0;
}

#endif  // DJ_EXPERIMENTS

// ---------------------------------------------------------------------------
// Generation, a handful of sequences: BOTH time-axis layers in one launch, as a wavefront.
// The window recompute of generate.py:106-109 is 2 layers x 128 sequential steps; for one to a few sequences every
// step is a pure latency chain (publish -> multicast -> MMA -> epilogue, ~2.2 us), so the layers are overlapped instead
// of run back to back: clusters [0, NT) run layer 0 exactly as the inference scan does and, after every step, bump a
// per-tile counter in global memory (gpu-scope release); clusters [NT, 2 NT) run layer 1 one step behind.  Layer 1's
// input projection is folded into its recurrent MMA: with no dropout at inference its input is h0_t + sp (sp = the
// style projection, constant over the window), so
//     z1_t = h0_t.W1 + (sp.W1 + b1) + h1_{t-1}.U1
// -- the bracket is a per-sequence constant computed once (c1), and h0_t.W1 is three more MMA passes on the SAME
// accumulator (W1^T hi in tensor memory next to U1^T hi, W1^T lo in shared memory next to U1^T lo) whose B operand is
// layer 0's own exchange buffer (h0_t as half hi + lo, TMA-loaded once the counter says all eight CTAs of the layer-0
// tile have published step t).  The W pass of step t does not depend on h1_{t-1}: it is issued while the epilogue of
// step t-1 is still running, into the other of two accumulator column sets, so the sequential chain stays one
// recurrent step per timestep and the window costs 129 step latencies instead of 256.
// 16-sequence tiles, 8 CTAs per cluster (32 hidden units = 128 gate rows each), one CTA per SM.
// ---------------------------------------------------------------------------
struct TcGen2Smem {
  static constexpr int A_BYTES = 128 * 256 * 2;                 // U^T lo slice: [4 atoms][128 rows][128 B]
  static constexpr int T_BYTES = 16 * 256 * 2;                  // one [16 x 256] half tile
  static constexpr int A_OFF = 0, W_OFF = A_BYTES, H_OFF = 2 * A_BYTES, HLO_OFF = H_OFF + T_BYTES,
                       X_OFF = HLO_OFF + T_BYTES;               // X: [2 buffers][hi | lo]
  static constexpr int BAR_OFF = X_OFF + 4 * T_BYTES;
  static constexpr int TOTAL = BAR_OFF + 128;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// ROLES = 2: layer 0 | layer 1 with the fused input pass (above).  ROLES = 3 (one sequence: 9 clusters): the input
// pass doubles layer 1's tensor-pipe time per step (96 small MMAs, ~2 us), which then bounds the step, so it moves to
// a THIRD set of clusters: "P" multiplies h0_t by W1 as soon as layer 0 publishes it and writes z1in_t = h0_t.W1 + c1
// (fp32 rows of Z1) behind a second counter; layer 1 is then the plain inference scan that waits for its Z rows.
template <bool HARD, int ROLES>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_gen2_kernel(const __grid_constant__ CUtensorMap tmU0lo, const __grid_constant__ CUtensorMap tmU1lo,
                    const __grid_constant__ CUtensorMap tmW1lo, const __grid_constant__ CUtensorMap tmH0,
                    const __grid_constant__ CUtensorMap tmH0lo, const __grid_constant__ CUtensorMap tmH1,
                    const __grid_constant__ CUtensorMap tmH1lo, const __grid_constant__ CUtensorMap tmX,
                    const __grid_constant__ CUtensorMap tmXlo, const float* __restrict__ Z0, float* __restrict__ Z1,
                    const float* __restrict__ C1, float* __restrict__ H1out, uint16_t* __restrict__ H0hi,
                    uint16_t* __restrict__ H0lo, uint16_t* __restrict__ H1hi, uint16_t* __restrict__ H1lo,
                    const uint32_t* __restrict__ Ut0_words, const uint32_t* __restrict__ Ut1_words,
                    const uint32_t* __restrict__ Wt1_words, uint32_t* __restrict__ flags, int ntiles, int steps,
                    float acc_scale) {
  constexpr int U = 256, C = 8, KA = 4, BS = 16, RH = 8;
  constexpr uint32_t AU_COL0 = 32, AW_COL0 = 32 + U / 2, TMEM_COLS = 512;
  using SM = TcGen2Smem;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  // barriers: a | w | acc[2] | h | x[2] | done | pub | tmem slot
  const uint32_t bar_a = sbase + SM::BAR_OFF, bar_w = bar_a + 8, bar_acc0 = bar_a + 16, bar_h = bar_a + 32,
                 bar_x0 = bar_a + 40, bar_done = bar_a + 56, bar_pub = bar_a + 64;
  volatile uint32_t* tmem_slot_p = (volatile uint32_t*)(smem + SM::BAR_OFF + 72);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cl = blockIdx.x / C;
  const bool isL0 = cl < ntiles;
  const bool isP = (ROLES == 3) && !isL0 && cl < 2 * ntiles;     // input projection of layer 1 (ROLES = 3)
  const bool isL1 = !isL0 && !isP;
  const bool fusedL1 = isL1 && ROLES == 2;      // layer 1 multiplies h0_t by W1 itself
  const bool plainL1 = isL1 && ROLES == 3;      // layer 1 reads z1in_t from Z1
  const int tile = cl % ntiles;                 // 16 sequences: pitches [16*(tile%3), +16) of batch element tile/3
  const int tile_b = tile / 3, tile_r = (tile % 3) * BS;
  const int warp = uniform_warp_id(), lane = threadIdx.x & 31;
  uint32_t* const flags0 = flags + tile;            // steps of layer 0 published, x8 CTAs
  uint32_t* const flags1 = flags + ntiles + tile;   // rows of Z1 published by P, x8 CTAs
  uint16_t* Hhi = isL1 ? H1hi : H0hi;
  uint16_t* Hlo = isL1 ? H1lo : H0lo;
  const CUtensorMap* tmH = isL1 ? &tmH1 : &tmH0;
  const CUtensorMap* tmHl = isL1 ? &tmH1lo : &tmH0lo;
  // the "recurrent" operand slots (tensor-memory columns AU, shared-memory slot A) hold U^T of the layer -- or, in
  // a P cluster, W1^T; the second pair (AW, slot W) holds W1^T in a fused layer-1 cluster
  const uint32_t* a_words = isP ? Wt1_words : isL1 ? Ut1_words : Ut0_words;
  const CUtensorMap* tmAlo = isP ? &tmW1lo : isL1 ? &tmU1lo : &tmU0lo;

  if (warp == 0) {
    if (lane == 0) {
      if (sbase & 1023u) { printf("deepj scan_tc_gen2: dynamic smem not 1024-aligned\n"); __trap(); }
      prefetch_tmap(tmAlo); prefetch_tmap(&tmW1lo); prefetch_tmap(tmH); prefetch_tmap(tmHl);
      prefetch_tmap(&tmX); prefetch_tmap(&tmXlo);
      mbar_init(bar_a, 1); mbar_init(bar_w, 1);
      mbar_init(bar_acc0, 1); mbar_init(bar_acc0 + 8, 1);
      mbar_init(bar_h, 1); mbar_init(bar_x0, 1); mbar_init(bar_x0 + 8, 1);
      mbar_init(bar_done, TC_EPI_WARPS); mbar_init(bar_pub, C);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(sbase + SM::BAR_OFF + 72, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_p;
  if (warp >= 1 && warp <= 4) {
    // resident A operands in tensor memory: lane = one row of this CTA's 128-row slice
    const uint32_t row = (uint32_t)(128 * rank + 32 * (warp & 3) + lane);
    const uint32_t* srcs[2] = {a_words + (size_t)row * (U / 2), Wt1_words + (size_t)row * (U / 2)};
    const uint32_t col0[2] = {AU_COL0, AW_COL0};
    for (int m = 0; m < (fusedL1 ? 2 : 1); ++m) {
#pragma unroll 1
      for (int c = 0; c < U / 2; c += 16) {
        uint32_t w[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 q4 = *reinterpret_cast<const uint4*>(srcs[m] + c + 4 * j);
          w[4 * j] = q4.x; w[4 * j + 1] = q4.y; w[4 * j + 2] = q4.z; w[4 * j + 3] = q4.w;
        }
        tmem_st16(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + col0[m] + (uint32_t)c, w);
      }
    }
    tmem_st_wait();
    tc_fence_before();
  }
  cluster.sync();   // peers' barriers are initialised before any multicast / remote arrive can target them (and A is in TMEM)
  tc_fence_after();

  if (warp == 0) {
    // ================= issuer =================
    if (elect_one()) {   // shared-memory A operands: the half-precision residuals of the tensor-memory ones, rows [128*rank, +128)
      mbar_expect_tx(bar_a, SM::A_BYTES);
#pragma unroll
      for (int ka = 0; ka < KA; ++ka)
        tma_load_2d(sbase + SM::A_OFF + ka * 16384, tmAlo, bar_a, ka * 64, 128 * rank);
      if (fusedL1) {
        mbar_expect_tx(bar_w, SM::A_BYTES);
#pragma unroll
        for (int ka = 0; ka < KA; ++ka)
          tma_load_2d(sbase + SM::W_OFF + ka * 16384, &tmW1lo, bar_w, ka * 64, 128 * rank);
      }
    }
    __syncwarp();
    const int my_ka = rank >> 1, my_hh = rank & 1;       // the slice of h_t this CTA multicasts
    // kind::f16 with IEEE-half operands (format bits 7 / 10 of the instruction descriptor: 0 = f16)
    const uint32_t idesc = make_idesc(128, BS, 0, 0) & ~((1u << 7) | (1u << 10));
    // three passes of one 256-deep product A.b on accumulator `d`: A_hi.b_hi + A_lo.b_hi + A_hi.b_lo
    auto three_pass = [&](uint32_t d, uint32_t a_tmem, uint32_t a_smem, uint32_t b_hi, uint32_t b_lo, bool first_zero) {
      const uint64_t adesc0 = make_smem_desc(a_smem, 16, 1024);
      const uint64_t bh0 = make_smem_desc(b_hi, 16, 1024), bl0 = make_smem_desc(b_lo, 16, 1024);
#pragma unroll
      for (int ka = 0; ka < KA; ++ka)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(d, a_tmem + (uint32_t)(ka * 32 + k * 8), bh0 + (uint64_t)((ka * (BS * 128) + k * 32) >> 4), idesc,
                       (first_zero && (ka | k) == 0) ? 0u : 1u);
#pragma unroll
      for (int ka = 0; ka < KA; ++ka)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(d, adesc0 + (uint64_t)((ka * 16384 + k * 32) >> 4), bh0 + (uint64_t)((ka * (BS * 128) + k * 32) >> 4),
                    idesc, 1);
#pragma unroll
      for (int ka = 0; ka < KA; ++ka)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(d, a_tmem + (uint32_t)(ka * 32 + k * 8), bl0 + (uint64_t)((ka * (BS * 128) + k * 32) >> 4), idesc, 1);
    };
    auto wait_counter = [&](const uint32_t* ctr, uint32_t want, int t) {
      if (lane == 0) {
        const long long t0 = clock64();
        while (ld_acquire_gpu(ctr) < want) {
          if (clock64() - t0 > 4000000000LL) { printf("deepj scan_tc_gen2: cluster %d waited in vain for step %d\n", cl, t); __trap(); }
        }
      }
      __syncwarp();
    };
    // Round t produces the accumulator of step t: [fused layer 1 / P: h0_t.W1] + [layers: h_{t-1}.U].  (Layer 0,
    // t = 0: nothing to multiply; the empty commit keeps the phases of the accumulator barriers the same for all roles.)
    for (int t = 0; t < steps; ++t) {
      const uint32_t d = tmem_base + (uint32_t)((t & 1) * BS);
      if (fusedL1 || isP) {
        // ---- input pass: wait until all eight CTAs of the layer-0 tile have published h0_t, fetch it, multiply
        wait_counter(flags0, (uint32_t)(C * (t + 1)), t);
        const uint32_t xb = sbase + SM::X_OFF + (uint32_t)((t & 1) * 2 * SM::T_BYTES), bar_x = bar_x0 + 8 * (t & 1);
        if (elect_one()) {
          mbar_expect_tx(bar_x, 2 * SM::T_BYTES);
#pragma unroll
          for (int ka = 0; ka < KA; ++ka) {
            tma_load_3d(xb + ka * (BS * 128), &tmX, bar_x, ka * 64, (t + 1) * 48 + tile_r, tile_b);
            tma_load_3d(xb + SM::T_BYTES + ka * (BS * 128), &tmXlo, bar_x, ka * 64, (t + 1) * 48 + tile_r, tile_b);
          }
        }
        __syncwarp();
        if (t == 0) mbar_wait(isP ? bar_a : bar_w, 0);
        mbar_wait(bar_x, (uint32_t)(t >> 1) & 1u);
        tc_fence_after();
        if (elect_one())
          three_pass(d, tmem_base + (isP ? AU_COL0 : AW_COL0), sbase + (isP ? SM::A_OFF : SM::W_OFF), xb, xb + SM::T_BYTES, true);
        __syncwarp();
      }
      if (isP) {
        // the epilogue of step t-1 must have read its accumulator set before round t+1 reuses it, and it cannot be
        // further than t-1 (it needs the commit below): waiting here keeps the parity waits one phase apart
        if (t > 0) mbar_wait(bar_done, (uint32_t)(t - 1) & 1u);
      } else if (t > 0) {
        // ---- recurrent pass: h_{t-1} published by every CTA of this cluster, all-gathered by multicast
        const uint32_t par = (uint32_t)(t - 1) & 1u;
        mbar_wait(bar_done, par);                         // this CTA's epilogue warps stored their part of h_{t-1}
        if (lane < C) mbar_arrive_remote(bar_pub, (uint32_t)lane);  // release.cluster, cumulative
        __syncwarp();
        mbar_wait(bar_pub, par);
        if (elect_one()) {
          mbar_expect_tx(bar_h, 2 * SM::T_BYTES);
          tma_load_3d_mc(sbase + SM::H_OFF + my_ka * (BS * 128) + my_hh * (RH * 128), tmH, bar_h, my_ka * 64,
                         t * 48 + tile_r + my_hh * RH, tile_b, (uint16_t)((1u << C) - 1u));
          tma_load_3d_mc(sbase + SM::HLO_OFF + my_ka * (BS * 128) + my_hh * (RH * 128), tmHl, bar_h, my_ka * 64,
                         t * 48 + tile_r + my_hh * RH, tile_b, (uint16_t)((1u << C) - 1u));
        }
        __syncwarp();
        if (t == 1) mbar_wait(bar_a, 0);
        mbar_wait(bar_h, par);
        tc_fence_after();
        if (elect_one()) three_pass(d, tmem_base + AU_COL0, sbase + SM::A_OFF, sbase + SM::H_OFF, sbase + SM::HLO_OFF, !fusedL1);
        __syncwarp();
      }
      if (elect_one()) umma_commit(bar_acc0 + 8 * (t & 1));
      __syncwarp();
    }
  } else if (warp == 1 + TC_EPI_WARPS) {
    // ================= publisher (layer 0 and P): one counter bump per step for the next stage =================
    // A gpu-scope release costs ~1 000 cycles; in the issuer warp it would sit on the step's chain, so this otherwise
    // idle warp watches the same `done` barrier (the epilogue's stores of step t are complete and CTA-visible) and
    // publishes them: the release is cumulative over what the acquire of the barrier wait made visible to this thread.
    if (isL0 || isP) {
      for (int t = 0; t < steps; ++t) {
        mbar_wait(bar_done, (uint32_t)t & 1u);
        if (lane == 0) red_release_gpu_add(isL0 ? flags0 : flags1, 1u);
        __syncwarp();
      }
    }
  } else if (warp <= TC_EPI_WARPS) {
    // ================= epilogue warps (as the inference scan: one 16-sequence chunk) =================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int w2 = (warp - 1) >> 2;         // which 8 of the 16 sequences
    const int up = lane >> 2;               // unit inside the warp (0..7)
    const int g = lane & 3;                 // gate held before the transpose / sequence slot after it
    const uint32_t col = 32 * rank + 8 * q + up;          // global hidden unit
    const uint32_t zc = 128 * rank + 32 * q + lane;       // gate-interleaved column this lane reads
    const uint32_t rowb = (uint32_t)(tile_b * steps * 48 + tile_r + 8 * w2);   // row of this warp's first sequence, step 0
    const float c1 = (isL0 || plainL1) ? 0.f : C1[(size_t)tile_b * (4 * U) + zc];
    if (isP) {
      // ---- P: z1in_t = h0_t.W1 + c1 -> rows t of Z1 (lane = gate column, 8 sequences: eight 128-byte row segments per warp)
      for (int t = 0; t < steps; ++t) {
        mbar_wait(bar_acc0 + 8 * (t & 1), (uint32_t)(t >> 1) & 1u);
        tc_fence_after();
        uint32_t acc[8];
        tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((t & 1) * BS + 8 * w2), acc);
        tmem_ld_wait();
        float* zp = Z1 + (size_t)(rowb + (uint32_t)t * 48u) * (4 * U) + zc;
#pragma unroll
        for (int j = 0; j < 8; ++j) __stcg(zp + (size_t)j * (4 * U), fmaf(__uint_as_float(acc[j]), acc_scale, c1));
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_done);   // release.cta: the issuer's gpu-scope release publishes these stores
      }
    } else {
      float cst[2] = {0.f, 0.f};
      float zreg[8];
      auto load_z = [&](int t) {      // pre-activations of step t that do not depend on this layer's recurrence
        if (isL0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) zreg[j] = Z0[(size_t)(rowb + j + (uint32_t)t * 48u) * (4 * U) + zc];
        } else if (plainL1) {
          if (lane == 0) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(flags1) < (uint32_t)(C * (t + 1))) {
              if (clock64() - t0 > 4000000000LL) { printf("deepj scan_tc_gen2: layer 1 waited in vain for Z1 rows of step %d\n", t); __trap(); }
            }
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) zreg[j] = __ldcg(Z1 + (size_t)(rowb + j + (uint32_t)t * 48u) * (4 * U) + zc);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) zreg[j] = c1;
        }
      };
      load_z(0);
      for (int t = 0; t < steps; ++t) {
        const bool not_last = (t + 1 < steps);
        float v[8];
        if (fusedL1 || t > 0) {
          mbar_wait(bar_acc0 + 8 * (t & 1), (uint32_t)(t >> 1) & 1u);
          tc_fence_after();
          uint32_t acc[8];
          tmem_ld8(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((t & 1) * BS + 8 * w2), acc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = fmaf(__uint_as_float(acc[j]), acc_scale, zreg[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = zreg[j];
        }
        // this lane's cell after the transpose: sequence 8*w2 + 4*blk + g, unit `col`
        const size_t o1 = (size_t)(rowb + (uint32_t)g + (uint32_t)t * 48u) * U + col;
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          // 4x4 transpose across the 4 lanes of a unit: lane g ends with i,f,g,o of sequence 4*blk+g
          const float a0 = v[4 * blk], a1 = v[4 * blk + 1], a2 = v[4 * blk + 2], a3 = v[4 * blk + 3];
          const bool odd = g & 1, hi = g & 2;
          const float x1 = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
          const float x2 = __shfl_xor_sync(0xffffffffu, odd ? a2 : a3, 1);
          const float b0 = odd ? x1 : a0, b1 = odd ? a1 : x1, b2 = odd ? x2 : a2, b3 = odd ? a3 : x2;
          const float y0 = __shfl_xor_sync(0xffffffffu, hi ? b0 : b2, 2);
          const float y1 = __shfl_xor_sync(0xffffffffu, hi ? b1 : b3, 2);
          const float zi = hi ? y0 : b0, zf = hi ? y1 : b1, zg_ = hi ? b2 : y0, zo = hi ? b3 : y1;
          const float gi = gate_act_fast<HARD>(zi), gf = gate_act_fast<HARD>(zf);
          const float gg = fast_tanh(zg_), go = gate_act_fast<HARD>(zo);
          const float cn = fmaf(gf, cst[blk], gi * gg);
          const float hn = go * fast_tanh(cn);
          cst[blk] = cn;
          // h_t as half hi + lo at the NEXT step's row of this layer's exchange buffer (layer 0: also after its last
          // step, for its consumer; the buffers have one spare timestep of rows)
          if (isL0 || not_last) {
            const __half hh = __float2half_rn(hn);
            Hhi[o1 + (size_t)(blk * 4 + 48) * U] = __half_as_ushort(hh);
            Hlo[o1 + (size_t)(blk * 4 + 48) * U] = __half_as_ushort(__float2half_rn(hn - __half2float(hh)));
          }
          if (isL1 && !not_last) H1out[o1 + (size_t)(blk * 4) * U] = hn;     // time_out of the window's last step
        }
        if (isL0 || not_last) {
          tc_fence_before();
          fence_proxy_async_all();   // generic-proxy global stores of h_t -> later async-proxy (TMA) reads
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_done);   // release.cta: hands this warp's stores to the issuer
        }
        if (not_last) load_z(t + 1);              // lands during this step's publish / all-gather / MMA chain
      }
    }
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int U, int BS, int UPC, bool AXIS_TIME>
int launch_tc_bwd(const void* Un_bf, const void* gates, const float* c, const float* dY, int64_t ldY, dj_dropout d_y,
                  void* dZ, float* db, int S, int steps, const TcMap& map, int hard, cudaStream_t st) {
#ifdef DJ_EXPERIMENTS
  if (!bwd_split_enabled())
    return launch_tc_bwd_inst<U, BS, UPC, AXIS_TIME, 1>(Un_bf, gates, c, dY, ldY, d_y, dZ, db, S, steps, map, hard, st);
#endif
  return launch_tc_bwd_inst<U, BS, UPC, AXIS_TIME, 2>(Un_bf, gates, c, dY, ldY, d_y, dZ, db, S, steps, map, hard, st);
}

}  // namespace

extern "C" int dj_lstm_scan_tc_fwd(const float* Z, void* gates16, float* h_out, float* c_out, void* h_prev_bf16,
                                   const void* Ut_bf16, const void* Ut_lo, int fmt, int S, int steps, int units, int seq_inner,
                                   int64_t seq_outer_stride, int64_t seq_inner_stride, int64_t step_stride, int hard,
                                   void* stream) {
  DJ_CHECK_ARG(Z && gates16 && h_out && h_prev_bf16 && Ut_bf16, "dj_lstm_scan_tc_fwd: NULL pointer");
  DJ_CHECK_ARG(fmt == DJ_BF16 || fmt == DJ_F16, "dj_lstm_scan_tc_fwd: operand format must be DJ_BF16 or DJ_F16");
  const int f16 = (fmt == DJ_F16) ? 1 : 0;
  DJ_CHECK_ARG(S > 0 && steps > 0, "dj_lstm_scan_tc_fwd: bad sizes");
  TcMap map{seq_inner, seq_outer_stride, seq_inner_stride, step_stride, 0, 0, 0, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  const bool time_map = (seq_inner == 48 && seq_inner_stride == 1 && step_stride == 48 && S % 48 == 0);
  const bool note_map = (seq_inner == 1 && seq_outer_stride == 48 && step_stride == 1 && steps <= 48);
  DJ_CHECK_ARG(time_map || note_map, "dj_lstm_scan_tc_fwd: only the time-axis (seq=(b,n)) and note-axis (seq=(b,t)) maps are supported");
  if (time_map && units == 256) {
    // 2 CTAs/SM x 16 resident 8-CTA clusters = 32 tile slots: beyond that, double the tile (two batch
    // elements per cluster) so the whole layer still runs as one wave of step chains
    if (S % 96 == 0 && S / 48 > 32)
      return launch_tc_fwd<256, 96, true>(Ut_bf16, Ut_lo, f16, Z, gates16, h_out, c_out, h_prev_bf16, S, steps, map, 1, hard, st);
    return launch_tc_fwd<256, 48, true>(Ut_bf16, Ut_lo, f16, Z, gates16, h_out, c_out, h_prev_bf16, S, steps, map, 1, hard, st);
  } else if (time_map && units == 512) {   // scaled model (BASELINE configs[4]): 16-CTA clusters
    return launch_tc_fwd<512, 48, true>(Ut_bf16, Ut_lo, f16, Z, gates16, h_out, c_out, h_prev_bf16, S, steps, map, 1, hard, st);
  } else if (note_map && units == 128) {
    if (S % 128 == 0 && S / 64 > 74)
      return launch_tc_fwd<128, 128, false>(Ut_bf16, Ut_lo, f16, Z, gates16, h_out, c_out, h_prev_bf16, S, steps, map, 0, hard, st);
    return launch_tc_fwd<128, 64, false>(Ut_bf16, Ut_lo, f16, Z, gates16, h_out, c_out, h_prev_bf16, S, steps, map, 0, hard, st);
  } else if (note_map && units == 256) {   // scaled model, note axis
    return launch_tc_fwd<256, 64, false>(Ut_bf16, Ut_lo, f16, Z, gates16, h_out, c_out, h_prev_bf16, S, steps, map, 0, hard, st);
  }
  DJ_CHECK_ARG(false, "dj_lstm_scan_tc_fwd: units=%d unsupported on this axis (time: 256/512, note: 128/256)", units);
  return -1;
}

extern "C" int dj_lstm_scan_tc_infer(const float* Z, float* h_out, void* h_hi, void* h_lo, const void* Ut_hi,
                                     const void* Ut_lo, float acc_scale, int S, int steps, int units, int seq_inner,
                                     int64_t seq_outer_stride, int64_t seq_inner_stride, int64_t step_stride, int hard,
                                     void* stream) {
  DJ_CHECK_ARG(Z && h_out && h_hi && h_lo && Ut_hi && Ut_lo, "dj_lstm_scan_tc_infer: NULL pointer");
  DJ_CHECK_ARG(S > 0 && steps > 0 && acc_scale > 0.f, "dj_lstm_scan_tc_infer: bad sizes");
  TcMap map{seq_inner, seq_outer_stride, seq_inner_stride, step_stride, 0, 0, 0, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  const bool time_map = (seq_inner == 48 && seq_inner_stride == 1 && step_stride == 48 && S % 48 == 0);
  DJ_CHECK_ARG(time_map && units == 256, "dj_lstm_scan_tc_infer: the time-axis map with 256 units is the supported shape");
  if (S % 96 == 0 && S / 48 > 32) {
    if (hard) return launch_tc_fwd_inst<256, 96, true, true, 2, 2, true>(Ut_hi, Ut_lo, 1, Z, nullptr, h_out, nullptr, h_hi, S, steps, map, st, h_lo, acc_scale);
    return launch_tc_fwd_inst<256, 96, true, false, 2, 2, true>(Ut_hi, Ut_lo, 1, Z, nullptr, h_out, nullptr, h_hi, S, steps, map, st, h_lo, acc_scale);
  }
  if (S / 48 <= 4) {
    // a handful of sequences (generate.py's default: 1 to 3): the recurrence is one latency chain per cluster, so the
    // 48 pitches of a sequence go to three clusters of 16-sequence tiles -- a third of the epilogue per step
    if (hard) return launch_tc_fwd_inst<256, 16, true, true, 1, 1, true>(Ut_hi, Ut_lo, 1, Z, nullptr, h_out, nullptr, h_hi, S, steps, map, st, h_lo, acc_scale);
    return launch_tc_fwd_inst<256, 16, true, false, 1, 1, true>(Ut_hi, Ut_lo, 1, Z, nullptr, h_out, nullptr, h_hi, S, steps, map, st, h_lo, acc_scale);
  }
  if (hard) return launch_tc_fwd_inst<256, 48, true, true, 1, 1, true>(Ut_hi, Ut_lo, 1, Z, nullptr, h_out, nullptr, h_hi, S, steps, map, st, h_lo, acc_scale);
  return launch_tc_fwd_inst<256, 48, true, false, 1, 1, true>(Ut_hi, Ut_lo, 1, Z, nullptr, h_out, nullptr, h_hi, S, steps, map, st, h_lo, acc_scale);
}

extern "C" int dj_lstm_scan_tc_gen2(const float* Z0, float* Z1, const float* c1, float* h1_out, void* h0_hi, void* h0_lo, void* h1_hi,
                                    void* h1_lo, const void* Ut0_hi, const void* Ut0_lo, const void* Ut1_hi,
                                    const void* Ut1_lo, const void* Wt1_hi, const void* Wt1_lo, float acc_scale,
                                    uint32_t* flags, int S, int steps, int hard, void* stream) {
  DJ_CHECK_ARG(Z0 && Z1 && c1 && h1_out && h0_hi && h0_lo && h1_hi && h1_lo && Ut0_hi && Ut0_lo && Ut1_hi && Ut1_lo && Wt1_hi &&
                   Wt1_lo && flags, "dj_lstm_scan_tc_gen2: NULL pointer");
  DJ_CHECK_ARG(S > 0 && S % 48 == 0 && steps > 0 && acc_scale > 0.f, "dj_lstm_scan_tc_gen2: bad sizes");
  constexpr int U = 256, C = 8;
  const int ntiles = S / 16;
  // every cluster of both layers must be resident at once (layer 1 spins on layer 0's counters): B200 keeps 15
  // clusters of 8 one-CTA-per-SM blocks resident
  DJ_CHECK_ARG(2 * ntiles <= 14, "dj_lstm_scan_tc_gen2: at most 2 sequences x 48 pitches (got %d rows); use dj_lstm_scan_tc_infer", S);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap tmU0lo, tmU1lo, tmW1lo, tmH0, tmH0lo, tmH1, tmH1lo, tmX, tmXlo;
  int rc;
  const CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;   // any 16-bit type: TMA only moves the bytes
  if ((rc = make_map_2d(&tmU0lo, dt, 2, Ut0_lo, U, 4 * U, U, 64, 128))) return rc;
  if ((rc = make_map_2d(&tmU1lo, dt, 2, Ut1_lo, U, 4 * U, U, 64, 128))) return rc;
  if ((rc = make_map_2d(&tmW1lo, dt, 2, Wt1_lo, U, 4 * U, U, 64, 128))) return rc;
  const uint64_t rows_per_b = (uint64_t)steps * 48, B = (uint64_t)(S / 48);
  {   // own-layer exchange buffers viewed as [b][t*48+n][U]: 8-row slices
    const uint64_t dims[3] = {(uint64_t)U, rows_per_b, B}, str[2] = {(uint64_t)U, rows_per_b * U};
    const uint32_t box[3] = {64, 8, 1};
    if ((rc = make_map(&tmH0, dt, 2, h0_hi, 3, dims, str, box))) return rc;
    if ((rc = make_map(&tmH0lo, dt, 2, h0_lo, 3, dims, str, box))) return rc;
    if ((rc = make_map(&tmH1, dt, 2, h1_hi, 3, dims, str, box))) return rc;
    if ((rc = make_map(&tmH1lo, dt, 2, h1_lo, 3, dims, str, box))) return rc;
  }
  {   // layer 0's buffers as layer 1 reads them: whole 16-row tiles, one extra timestep of rows per batch element (h0 of
      // the last step sits at row T; the views of consecutive batch elements overlap by that timestep)
    const uint64_t dims[3] = {(uint64_t)U, rows_per_b + 48, B}, str[2] = {(uint64_t)U, rows_per_b * U};
    const uint32_t box[3] = {64, 16, 1};
    if ((rc = make_map(&tmX, dt, 2, h0_hi, 3, dims, str, box))) return rc;
    if ((rc = make_map(&tmXlo, dt, 2, h0_lo, 3, dims, str, box))) return rc;
  }
  DJ_CUDA(cudaMemsetAsync(flags, 0, sizeof(uint32_t) * 2 * (size_t)ntiles, st));
  // one sequence: three roles (layer 0 | input projection of layer 1 | layer 1) on 9 clusters; two: two roles on 12
  const int roles = (3 * ntiles <= 14 && !getenv("DJ_GEN2_ROLES2")) ? 3 : 2;
  auto kernel = roles == 3 ? (hard ? scan_tc_gen2_kernel<true, 3> : scan_tc_gen2_kernel<false, 3>)
                           : (hard ? scan_tc_gen2_kernel<true, 2> : scan_tc_gen2_kernel<false, 2>);
  DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcGen2Smem::TOTAL));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(roles * ntiles * C);   // clusters [0, ntiles): layer 0; then (3 roles) P; the last ntiles: layer 1
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = TcGen2Smem::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  DJ_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmU0lo, tmU1lo, tmW1lo, tmH0, tmH0lo, tmH1, tmH1lo, tmX, tmXlo, Z0, Z1, c1, h1_out,
                             (uint16_t*)h0_hi, (uint16_t*)h0_lo, (uint16_t*)h1_hi, (uint16_t*)h1_lo,
                             (const uint32_t*)Ut0_hi, (const uint32_t*)Ut1_hi, (const uint32_t*)Wt1_hi, flags, ntiles, steps,
                             acc_scale));
  return 0;
}

extern "C" int dj_lstm_scan_tc_bwd(const void* gates, const float* c, const float* dY, int64_t ldY, dj_dropout d_y,
                                   const void* Un_bf16, void* dZ_bf16, float* db, int S, int steps, int units,
                                   int seq_inner, int64_t seq_outer_stride, int64_t seq_inner_stride,
                                   int64_t step_stride, int hard, void* stream) {
  DJ_CHECK_ARG(gates && c && dY && Un_bf16 && dZ_bf16 && db, "dj_lstm_scan_tc_bwd: NULL pointer");
  DJ_CHECK_ARG(S > 0 && steps > 0 && ldY >= units, "dj_lstm_scan_tc_bwd: bad sizes");
  TcMap map{seq_inner, seq_outer_stride, seq_inner_stride, step_stride, 0, 0, 0, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  const bool time_map = (seq_inner == 48 && seq_inner_stride == 1 && step_stride == 48 && S % 48 == 0);
  const bool note_map = (seq_inner == 1 && seq_outer_stride == 48 && step_stride == 1 && steps <= 48);
  DJ_CHECK_ARG(time_map || note_map, "dj_lstm_scan_tc_bwd: only the time-axis (seq=(b,n)) and note-axis (seq=(b,t)) maps are supported");
  if (time_map && units == 256) {
    // 64 units per CTA, clusters of 4: 33-37 clusters are resident, so 64 sequences per GPU (64 tiles of 48) run as
    // two waves.  A one-wave variant (96 sequences per cluster through one shared staging buffer) exists:
#ifdef DJ_EXPERIMENTS
    // (measured in round 2: correct but SLOWER than two waves, 1.24 ms against 0.88 ms at 64 sequences per GPU -- the
    // step is bound by how fast an SM ingests the multicast dz all-gather, ~30 B/clk, and that volume is per
    // sequence; see DESIGN.md section 4.  Kept as an experiment: DJ_BWD_SHARED=1 in a -DDJ_EXPERIMENTS build.)
    static int one_wave = -1;
    if (one_wave < 0) { const char* e = getenv("DJ_BWD_SHARED"); one_wave = (e && atoi(e) == 1) ? 1 : 0; }
    if (one_wave && S % 96 == 0 && S / 48 > 33)
      return launch_tc_bwd_inst<256, 96, 64, true, 2, true>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, hard, st);
#endif
#ifdef DJ_EXPERIMENTS
    // CTA pairs (96 sequences per cluster at the staging and ingest cost of 48, one wave): correct, and measured
    // SLOWER than two waves at 64 sequences per GPU (0.85 against 0.71 ms).  DJ_BWD_PAIR_TIME=1.
    static int pair_time = -1;
    if (pair_time < 0) { const char* e = getenv("DJ_BWD_PAIR_TIME"); pair_time = (e && atoi(e) == 1) ? 1 : 0; }
    if (pair_time && S % 96 == 0)
      return launch_tc_bwd_pair<256, 96, true, 1, true>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, hard, st);
#endif
    return launch_tc_bwd<256, 48, 64, true>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, hard, st);
  }
  if (time_map && units == 512)   // scaled model: 32 units per CTA, 16-CTA clusters, a third of a batch element per tile
    return launch_tc_bwd<512, 16, 32, true>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, hard, st);
  if (note_map && units == 128) {
#ifdef DJ_EXPERIMENTS
    // one pair per cluster, 64 sequences: 0.30 against 0.33 ms alone, no difference inside the step (DJ_BWD_PAIR=1)
    if (bwd_pair_enabled() && S % 64 == 0)
      return launch_tc_bwd_pair<128, 64, false, 1, false>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, hard, st);
#endif
    return launch_tc_bwd<128, 32, 64, false>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, hard, st);
  }
  if (note_map && units == 256)   // scaled model, note axis
    return launch_tc_bwd<256, 32, 64, false>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, hard, st);
  DJ_CHECK_ARG(false, "dj_lstm_scan_tc_bwd: units=%d unsupported on this axis (time: 256/512, note: 128/256)", units);
  return -1;
}
