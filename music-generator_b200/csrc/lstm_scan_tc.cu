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

#include "dj_tc.cuh"

namespace cg = cooperative_groups;

namespace {

struct TcMap {
  int seq_inner;                       // 48 (time axis: seq = b*48+n) or 1 (note axis)
  int64_t outer_stride, inner_stride, step_stride;
  // TMA coordinates of a half-tile slice of hprev: c1 = t*step1 + hh*off1, c2 = tile*base2 + hh*off2
  int step1, off1, base2, off2;
  int64_t seq_stride;                  // row distance of consecutive sequences inside a 16-chunk
};

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 32 * (1 + TC_EPI_WARPS);

template <int U, int BS>
struct TcFwdSmem {
  static constexpr int A_BYTES = 128 * U * 2;
  static constexpr int H_BYTES = BS * U * 2;
  static constexpr int A_OFF = 0, H_OFF = A_BYTES, BAR_OFF = A_BYTES + H_BYTES;
  static constexpr int TOTAL = BAR_OFF + 64 + 1024;
};

__device__ __forceinline__ float fast_tanh(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float gate_act_fast(float x, int hard) {
  return hard ? dj_hard_sigmoid(x) : __fdividef(1.0f, 1.0f + __expf(-x));
}
__device__ __forceinline__ int64_t tc_row0(const TcMap& m, int seq) {
  return (int64_t)(seq / m.seq_inner) * m.outer_stride + (int64_t)(seq % m.seq_inner) * m.inner_stride;
}

template <int U, int BS>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmH,
                   float* __restrict__ Z, float* __restrict__ Hout, float* __restrict__ Cout,
                   __nv_bfloat16* __restrict__ Hprev, int S, int steps, TcMap map, int hard) {
  constexpr int C = U / 32;            // cluster size
  constexpr int KA = U / 64;           // 64-wide K atoms
  constexpr int RH = BS / 2;           // rows per multicast slice
  constexpr int NCH = BS / 16;         // 16-sequence chunks per tile
  constexpr int MAXCH = (NCH + 1) / 2; // chunks per epilogue warp
  constexpr uint32_t TMEM_COLS = BS <= 32 ? 32 : BS <= 64 ? 64 : BS <= 128 ? 128 : 256;
  static_assert(C == 2 * KA, "slices = K atoms x 2 row halves");
  static_assert(BS % 16 == 0 && BS <= 256 && RH % 8 == 0, "tile shape");
  using SM = TcFwdSmem<U, BS>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_a = sbase + SM::BAR_OFF, bar_h = bar_a + 8, bar_acc = bar_a + 16;
  uint32_t* tmem_slot = (uint32_t*)(smem + SM::BAR_OFF + 24);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / C, ncl = gridDim.x / C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmU); prefetch_tmap(&tmH);
      mbar_init(bar_a, 1); mbar_init(bar_h, 1); mbar_init(bar_acc, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster.sync();   // peers' barriers are initialised before any multicast can target them

  if (warp == 0 && lane == 0) {   // resident A operand: rows [128*rank, +128) of U^T
    mbar_expect_tx(bar_a, SM::A_BYTES);
#pragma unroll
    for (int ka = 0; ka < KA; ++ka)
      tma_load_2d(sbase + SM::A_OFF + ka * 16384, &tmU, bar_a, ka * 64, 128 * rank);
  }

  const int ntiles = (S + BS - 1) / BS;
  uint32_t h_phase = 0, acc_phase = 0;
  bool a_ready = false;

  // epilogue thread coordinates
  const int ew = warp - 1;
  const int q = warp & 3;                 // TMEM lane quarter this warp may access
  const int cp = (ew >= 4) ? 1 : 0;       // chunk parity handled by this warp
  const int ul = 8 * q + (lane >> 2);     // local hidden unit
  const int g = lane & 3;                 // gate held before the transpose / sequence slot after it
  const int col = 32 * rank + ul;         // global hidden unit
  const int zc = 128 * rank + 32 * q + lane;   // gate-interleaved column this lane reads

  for (int tile = cid; tile < ntiles; tile += ncl) {
    float cst[MAXCH][4];
    int64_t rowbase[MAXCH];   // row of the first sequence of each of this warp's chunks at step 0
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) cst[i][j] = 0.f;
      const int seq0 = tile * BS + (cp + 2 * i) * 16;   // a 16-chunk never straddles a batch element (48 = 3*16)
      rowbase[i] = tc_row0(map, seq0 < S ? seq0 : 0);
    }

    for (int t = 0; t < steps; ++t) {
      if (warp == 0) {
        // ================= MMA issuer =================
        if (lane == 0 && t > 0) {
          if (!a_ready) { mbar_wait(bar_a, 0); a_ready = true; }
          mbar_wait(bar_h, h_phase);
          tc_fence_after();
          constexpr uint32_t idesc = make_idesc(128, BS, 0, 0);
#pragma unroll
          for (int ka = 0; ka < KA; ++ka)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = make_smem_desc(sbase + SM::A_OFF + ka * 16384 + k * 32, 16, 1024);
              const uint64_t bdesc = make_smem_desc(sbase + SM::H_OFF + ka * (BS * 128) + k * 32, 16, 1024);
              umma_bf16(tmem_base, adesc, bdesc, idesc, (ka | k) != 0);
            }
          umma_commit(bar_acc);
        }
        if (t > 0) h_phase ^= 1;
        __syncwarp();
      } else {
        // ================= epilogue =================
        float zreg[16];
        auto load_z = [&](int ci) {
          const int seq0 = tile * BS + (cp + 2 * ci) * 16;
          const float* zp = Z + (rowbase[ci] + (int64_t)t * map.step_stride) * (4 * U) + zc;
#pragma unroll
          for (int j = 0; j < 16; ++j) zreg[j] = (seq0 + j < S) ? zp[(int64_t)j * map.seq_stride * (4 * U)] : 0.f;
        };
        load_z(0);                        // overlaps the MMA / TMA latency
        if (t > 0) {
          mbar_wait(bar_acc, acc_phase);
          tc_fence_after();
        }
#pragma unroll
        for (int ci = 0; ci < MAXCH; ++ci) {
          const int ch = cp + 2 * ci;
          if (ch < NCH) {
            float v[16];
            if (t > 0) {
              uint32_t acc[16];
              tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(ch * 16), acc);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]) + zreg[j];
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = zreg[j];
            }
            if (ch + 2 < NCH) load_z(ci + 1);   // next chunk's pre-activations in flight during the math
#pragma unroll
            for (int blk = 0; blk < 4; ++blk) {
              // 4x4 transpose across the 4 lanes of a unit: lane g ends with i,f,g,o of sequence 4*blk+g
              const float a0 = v[4 * blk], a1 = v[4 * blk + 1], a2 = v[4 * blk + 2], a3 = v[4 * blk + 3];
              const bool odd = g & 1, hi = g & 2;
              const float x1 = __shfl_xor_sync(0xffffffffu, odd ? a0 : a1, 1);
              const float x2 = __shfl_xor_sync(0xffffffffu, odd ? a2 : a3, 1);
              const float b0 = odd ? x1 : a0, b1 = odd ? a1 : x1, b2 = odd ? x2 : a2, b3 = odd ? a3 : x2;
              const float y0 = __shfl_xor_sync(0xffffffffu, hi ? b0 : b2, 2);
              const float y1 = __shfl_xor_sync(0xffffffffu, hi ? b1 : b3, 2);
              const float zi = hi ? y0 : b0, zf = hi ? y1 : b1, zg = hi ? b2 : y0, zo = hi ? b3 : y1;

              const int seq = tile * BS + ch * 16 + 4 * blk + g;
              const bool ok = seq < S;
              const int64_t row = rowbase[ci] + (int64_t)(4 * blk + g) * map.seq_stride + (int64_t)t * map.step_stride;
              const float gi = gate_act_fast(zi, hard), gf = gate_act_fast(zf, hard);
              const float gg = fast_tanh(zg), go = gate_act_fast(zo, hard);
              const float cn = fmaf(gf, cst[ci][blk], gi * gg);
              const float hn = go * fast_tanh(cn);
              cst[ci][blk] = cn;
              if (ok) {
                *reinterpret_cast<float4*>(Z + row * (4 * U) + 4 * col) = make_float4(gi, gf, gg, go);
                Hout[row * U + col] = hn;
                if (Cout != nullptr) Cout[row * U + col] = cn;
              }
              // gather the 8 units of this warp (same sequence) into one 16-byte bf16 chunk
              uint32_t p = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hn));
              const uint32_t y = __shfl_xor_sync(0xffffffffu, p, 4);
              p = (lane & 4) ? (y | (p << 16)) : (p | (y << 16));
              const uint32_t r8 = __shfl_xor_sync(0xffffffffu, p, 8);
              const uint32_t lo = (lane & 8) ? r8 : p, hi2 = (lane & 8) ? p : r8;
              const uint32_t rl = __shfl_xor_sync(0xffffffffu, lo, 16), rh = __shfl_xor_sync(0xffffffffu, hi2, 16);
              if (ok && lane < 4) {
                const uint4 chunk = (lane & 16) ? make_uint4(rl, rh, lo, hi2) : make_uint4(lo, hi2, rl, rh);
                __nv_bfloat16* hp = Hprev + row * U + 32 * rank + 8 * q;
                if (t + 1 < steps) *reinterpret_cast<uint4*>(hp + map.step_stride * U) = chunk;
                if (t == 0) *reinterpret_cast<uint4*>(hp) = make_uint4(0u, 0u, 0u, 0u);
              }
            }
          }
        }
        if (t > 0) acc_phase ^= 1;
        tc_fence_before();
        fence_proxy_async_all();   // generic-proxy global writes of h_t -> visible to the TMA (async proxy) reads
      }
      // ---- all CTAs of the cluster have published their slice of h_t
      cluster.sync();
      if (warp == 0 && lane == 0 && t + 1 < steps) {
        fence_proxy_async_all();
        mbar_expect_tx(bar_h, SM::H_BYTES);
        const int ka = rank >> 1, hh = rank & 1;
        tma_load_3d_mc(sbase + SM::H_OFF + ka * (BS * 128) + hh * (RH * 128), &tmH, bar_h, ka * 64,
                       (t + 1) * map.step1 + hh * map.off1, tile * map.base2 + hh * map.off2,
                       (uint16_t)((1u << C) - 1u));
      }
    }
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int U, int BS>
int launch_tc_fwd(const void* Ut_bf, float* Z, float* h_out, float* c_out, void* hprev, int S, int steps,
                  const TcMap& map_in, int axis_time, int hard, cudaStream_t st) {
  constexpr int C = U / 32;
  using SM = TcFwdSmem<U, BS>;
  TcMap map = map_in;
  CUtensorMap tmU, tmH;
  int rc;
  if ((rc = make_map_2d(&tmU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Ut_bf, (uint64_t)U, (uint64_t)4 * U, (uint64_t)U, 64, 128)))
    return rc;
  constexpr int RH = BS / 2;
  if (axis_time) {
    // hprev viewed as [b][t*48+n][U]
    const uint64_t rows_per_b = (uint64_t)map.outer_stride;            // T*48
    const uint64_t B = (uint64_t)(S / 48);
    const uint64_t dims[3] = {(uint64_t)U, rows_per_b, B}, str[2] = {(uint64_t)U, rows_per_b * U};
    const uint32_t box[3] = {64, (uint32_t)(RH <= 48 ? RH : 48), (uint32_t)(RH <= 48 ? 1 : RH / 48)};
    if ((rc = make_map(&tmH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hprev, 3, dims, str, box))) return rc;
    map.step1 = 48; map.off1 = (RH % 48); map.base2 = BS / 48; map.off2 = RH / 48;
    map.seq_stride = map.inner_stride;
  } else {
    // hprev viewed as [seq][n][U]
    const uint64_t dims[3] = {(uint64_t)U, 48, (uint64_t)S}, str[2] = {(uint64_t)U, (uint64_t)48 * U};
    const uint32_t box[3] = {64, 1, (uint32_t)RH};
    if ((rc = make_map(&tmH, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hprev, 3, dims, str, box))) return rc;
    map.step1 = 1; map.off1 = 0; map.base2 = BS; map.off2 = RH;
    map.seq_stride = map.outer_stride;
  }
  auto kernel = scan_tc_fwd_kernel<U, BS>;
  DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
  const int ntiles = (S + BS - 1) / BS;
  int ncl = dj_num_sms() / C;
  if (ncl > 16 && C == 8) ncl = 16;          // at most two 8-CTA clusters fit a GPC
  if (ncl > ntiles) ncl = ntiles;
  const int rounds = (ntiles + ncl - 1) / ncl;
  ncl = (ntiles + rounds - 1) / rounds;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ncl * C);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SM::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  __nv_bfloat16* hp = (__nv_bfloat16*)hprev;
  DJ_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmU, tmH, Z, h_out, c_out, hp, S, steps, map, hard));
  return 0;
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
template <int U, int BS>
struct TcBwdSmem {
  static constexpr int A_BYTES = 64 * 4 * U * 2;
  static constexpr int B_BYTES = BS * 4 * U * 2;
  static constexpr int A_OFF = 0, B_OFF = A_BYTES, BAR_OFF = A_BYTES + B_BYTES;
  static constexpr int TOTAL = BAR_OFF + 64 + 1024;
};

template <int U, int BS>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmZ,
                   const float* __restrict__ G, const float* __restrict__ Cst, const float* __restrict__ dY,
                   int64_t ldY, dj_dropout d_y, __nv_bfloat16* __restrict__ dZ, float* __restrict__ db, int S,
                   int steps, TcMap map, int hard) {
  constexpr int C = U / 64;             // cluster size
  constexpr int KA = 4 * U / 64;        // K atoms of the contraction (gate columns)
  constexpr int KPC = KA / C;           // atoms each CTA multicasts per step
  constexpr int NCH = BS / 16;
  constexpr int MAXCH = (NCH + 1) / 2;
  constexpr uint32_t TMEM_COLS = BS <= 32 ? 32 : BS <= 64 ? 64 : BS <= 128 ? 128 : 256;
  using SM = TcBwdSmem<U, BS>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_a = sbase + SM::BAR_OFF, bar_z = bar_a + 8, bar_acc = bar_a + 16;
  uint32_t* tmem_slot = (uint32_t*)(smem + SM::BAR_OFF + 24);

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / C, ncl = gridDim.x / C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmU); prefetch_tmap(&tmZ);
      mbar_init(bar_a, 1); mbar_init(bar_z, 1); mbar_init(bar_acc, 1);
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

  if (warp == 0 && lane == 0) {   // resident A operand: rows [64*rank, +64) of U, all 4U columns
    mbar_expect_tx(bar_a, SM::A_BYTES);
    for (int ja = 0; ja < KA; ++ja)
      tma_load_2d(sbase + SM::A_OFF + ja * 8192, &tmU, bar_a, ja * 64, 64 * rank);
  }

  const int ntiles = (S + BS - 1) / BS;
  uint32_t z_phase = 0, acc_phase = 0;
  bool a_ready = false;

  // epilogue thread coordinates: TMEM quarter q holds units 16q..16q+15 in its lanes 0..15;
  // lanes 16..31 take over half of each chunk's sequences by shuffle.
  const int ew = warp - 1;
  const int q = warp & 3;
  const int cp = (ew >= 4) ? 1 : 0;
  const int ul = 16 * q + (lane & 15);
  const int sh = lane >> 4;               // which 8 sequences of a 16-chunk this lane handles
  const int col = 64 * rank + ul;         // global hidden unit
  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};

  for (int tile = cid; tile < ntiles; tile += ncl) {
    float dcn[MAXCH][8];
    int64_t rowbase[MAXCH];
#pragma unroll
    for (int i = 0; i < MAXCH; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dcn[i][j] = 0.f;
      const int seq0 = tile * BS + (cp + 2 * i) * 16;
      rowbase[i] = tc_row0(map, seq0 < S ? seq0 : 0);
    }

    for (int t = steps - 1; t >= 0; --t) {
      const bool have_rec = (t != steps - 1);
      if (warp == 0) {
        // ================= MMA issuer: dh_t = U . dz_{t+1}^T =================
        if (lane == 0 && have_rec) {
          if (!a_ready) { mbar_wait(bar_a, 0); a_ready = true; }
          mbar_wait(bar_z, z_phase);
          tc_fence_after();
          constexpr uint32_t idesc = make_idesc(64, BS, 0, 0);
#pragma unroll 4
          for (int ja = 0; ja < KA; ++ja)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adesc = make_smem_desc(sbase + SM::A_OFF + ja * 8192 + k * 32, 16, 1024);
              const uint64_t bdesc = make_smem_desc(sbase + SM::B_OFF + ja * (BS * 128) + k * 32, 16, 1024);
              umma_bf16(tmem_base, adesc, bdesc, idesc, (ja | k) != 0);
            }
          umma_commit(bar_acc);
        }
        if (have_rec) z_phase ^= 1;
        __syncwarp();
      } else {
        // ================= epilogue: gate derivatives =================
        if (have_rec) {
          mbar_wait(bar_acc, acc_phase);
          tc_fence_after();
        }
#pragma unroll
        for (int ci = 0; ci < MAXCH; ++ci) {
          const int ch = cp + 2 * ci;
          if (ch < NCH) {
            float dh[8];
            if (have_rec) {
              uint32_t acc[16];
              tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(ch * 16), acc);
              // lanes 0..15 own the data; lanes 16..31 fetch sequences 8..15 of the chunk from lane-16
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float up = __shfl_sync(0xffffffffu, __uint_as_float(acc[8 + j]), lane & 15);
                dh[j] = sh ? up : __uint_as_float(acc[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) dh[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int sl = 8 * sh + j;                       // sequence inside the chunk
              const int seq = tile * BS + ch * 16 + sl;
              if (seq < S) {
                const int64_t r = rowbase[ci] + (int64_t)sl * map.seq_stride + (int64_t)t * map.step_stride;
                const int64_t rp = r - map.step_stride;
                const float4 g4 = *reinterpret_cast<const float4*>(G + r * (4 * U) + 4 * col);
                const float ct = Cst[r * U + col];
                const float cprev = (t > 0) ? Cst[rp * U + col] : 0.f;
                const float dy = dY[r * ldY + col] * dj_dropmul(d_y, (uint32_t)(r * U + col));
                const float dht = dy + dh[j];
                const float tc = fast_tanh(ct);
                const float d_o = dht * tc;
                const float dc = fmaf(dht * g4.w, 1.f - tc * tc, dcn[ci][j]);
                dcn[ci][j] = dc * g4.y;
                float dz[4];
                dz[0] = dc * g4.z * dj_gate_dact(g4.x, hard);
                dz[1] = dc * cprev * dj_gate_dact(g4.y, hard);
                dz[2] = dc * g4.x * (1.f - g4.z * g4.z);
                dz[3] = d_o * dj_gate_dact(g4.w, hard);
                __nv_bfloat162 lo = __floats2bfloat162_rn(dz[0], dz[1]), hi = __floats2bfloat162_rn(dz[2], dz[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(dZ + r * (4 * U) + 4 * col) = pk;
                dbacc[0] += dz[0]; dbacc[1] += dz[1]; dbacc[2] += dz[2]; dbacc[3] += dz[3];
              }
            }
          }
        }
        if (have_rec) acc_phase ^= 1;
        tc_fence_before();
        fence_proxy_async_all();
      }
      cluster.sync();
      if (warp == 0 && lane == 0 && t > 0) {   // all-gather dz_t: this CTA multicasts KPC of the KA column atoms
        fence_proxy_async_all();
        mbar_expect_tx(bar_z, SM::B_BYTES);
        for (int i = 0; i < KPC; ++i) {
          const int ja = rank * KPC + i;
          tma_load_3d_mc(sbase + SM::B_OFF + ja * (BS * 128), &tmZ, bar_z, ja * 64, t * map.step1, tile * map.base2,
                         (uint16_t)((1u << C) - 1u));
        }
      }
    }
  }
  if (warp > 0) {
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) atomicAdd(db + 4 * col + gq, dbacc[gq]);
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int U, int BS>
int launch_tc_bwd(const void* Un_bf, const float* gates, const float* c, const float* dY, int64_t ldY, dj_dropout d_y,
                  void* dZ, float* db, int S, int steps, const TcMap& map_in, int axis_time, int hard, cudaStream_t st) {
  constexpr int C = U / 64;
  using SM = TcBwdSmem<U, BS>;
  TcMap map = map_in;
  CUtensorMap tmU, tmZ;
  int rc;
  if ((rc = make_map_2d(&tmU, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, Un_bf, (uint64_t)4 * U, (uint64_t)U, (uint64_t)4 * U, 64, 64)))
    return rc;
  if (axis_time) {
    const uint64_t rows_per_b = (uint64_t)map.outer_stride, B = (uint64_t)(S / 48);
    const uint64_t dims[3] = {(uint64_t)4 * U, rows_per_b, B}, str[2] = {(uint64_t)4 * U, rows_per_b * 4 * U};
    const uint32_t box[3] = {64, 48, (uint32_t)(BS / 48)};
    if ((rc = make_map(&tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dZ, 3, dims, str, box))) return rc;
    map.step1 = 48; map.off1 = 0; map.base2 = BS / 48; map.off2 = 0;
    map.seq_stride = map.inner_stride;
  } else {
    const uint64_t dims[3] = {(uint64_t)4 * U, 48, (uint64_t)S}, str[2] = {(uint64_t)4 * U, (uint64_t)48 * 4 * U};
    const uint32_t box[3] = {64, 1, (uint32_t)BS};
    if ((rc = make_map(&tmZ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dZ, 3, dims, str, box))) return rc;
    map.step1 = 1; map.off1 = 0; map.base2 = BS; map.off2 = 0;
    map.seq_stride = map.outer_stride;
  }
  auto kernel = scan_tc_bwd_kernel<U, BS>;
  DJ_CUDA(cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL));
  const int ntiles = (S + BS - 1) / BS;
  int ncl = dj_num_sms() / C;
  if (ncl > ntiles) ncl = ntiles;
  const int rounds = (ntiles + ncl - 1) / ncl;
  ncl = (ntiles + rounds - 1) / rounds;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ncl * C);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = SM::TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  __nv_bfloat16* dzp = (__nv_bfloat16*)dZ;
  DJ_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmU, tmZ, gates, c, dY, ldY, d_y, dzp, db, S, steps, map, hard));
  return 0;
}

}  // namespace

extern "C" int dj_lstm_scan_tc_fwd(float* Z, float* h_out, float* c_out, void* h_prev_bf16, const void* Ut_bf16, int S,
                                   int steps, int units, int seq_inner, int64_t seq_outer_stride,
                                   int64_t seq_inner_stride, int64_t step_stride, int hard, void* stream) {
  DJ_CHECK_ARG(Z && h_out && h_prev_bf16 && Ut_bf16, "dj_lstm_scan_tc_fwd: NULL pointer");
  DJ_CHECK_ARG(S > 0 && steps > 0, "dj_lstm_scan_tc_fwd: bad sizes");
  TcMap map{seq_inner, seq_outer_stride, seq_inner_stride, step_stride, 0, 0, 0, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = dj_num_sms();
  if (units == 256) {
    // time axis: sequences (b, n), rows of one b are contiguous at each step
    DJ_CHECK_ARG(seq_inner == 48 && seq_inner_stride == 1 && step_stride == 48 && S % 48 == 0,
                 "dj_lstm_scan_tc_fwd: units=256 expects the time-axis map (seq=(b,n))");
    const int max_cl = sms / 8 > 16 ? 16 : sms / 8;
    if (S <= 48 * max_cl) return launch_tc_fwd<256, 48>(Ut_bf16, Z, h_out, c_out, h_prev_bf16, S, steps, map, 1, hard, st);
    if (S <= 96 * max_cl) return launch_tc_fwd<256, 96>(Ut_bf16, Z, h_out, c_out, h_prev_bf16, S, steps, map, 1, hard, st);
    return launch_tc_fwd<256, 192>(Ut_bf16, Z, h_out, c_out, h_prev_bf16, S, steps, map, 1, hard, st);
  } else if (units == 128) {
    DJ_CHECK_ARG(seq_inner == 1 && seq_outer_stride == 48 && step_stride == 1 && steps <= 48,
                 "dj_lstm_scan_tc_fwd: units=128 expects the note-axis map (seq=(b,t))");
    const int max_cl = sms / 4;
    if (S <= 64 * max_cl) return launch_tc_fwd<128, 64>(Ut_bf16, Z, h_out, c_out, h_prev_bf16, S, steps, map, 0, hard, st);
    if (S <= 128 * max_cl) return launch_tc_fwd<128, 128>(Ut_bf16, Z, h_out, c_out, h_prev_bf16, S, steps, map, 0, hard, st);
    return launch_tc_fwd<128, 256>(Ut_bf16, Z, h_out, c_out, h_prev_bf16, S, steps, map, 0, hard, st);
  }
  DJ_CHECK_ARG(false, "dj_lstm_scan_tc_fwd: units=%d unsupported (128 or 256)", units);
  return -1;
}

extern "C" int dj_lstm_scan_tc_bwd(const float* gates, const float* c, const float* dY, int64_t ldY, dj_dropout d_y,
                                   const void* Un_bf16, void* dZ_bf16, float* db, int S, int steps, int units,
                                   int seq_inner, int64_t seq_outer_stride, int64_t seq_inner_stride,
                                   int64_t step_stride, int hard, void* stream) {
  DJ_CHECK_ARG(gates && c && dY && Un_bf16 && dZ_bf16 && db, "dj_lstm_scan_tc_bwd: NULL pointer");
  DJ_CHECK_ARG(S > 0 && steps > 0 && ldY >= units, "dj_lstm_scan_tc_bwd: bad sizes");
  TcMap map{seq_inner, seq_outer_stride, seq_inner_stride, step_stride, 0, 0, 0, 0, 0};
  cudaStream_t st = (cudaStream_t)stream;
  if (units == 256) {
    DJ_CHECK_ARG(seq_inner == 48 && seq_inner_stride == 1 && step_stride == 48 && S % 48 == 0,
                 "dj_lstm_scan_tc_bwd: units=256 expects the time-axis map (seq=(b,n))");
    return launch_tc_bwd<256, 48>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, 1, hard, st);
  } else if (units == 128) {
    DJ_CHECK_ARG(seq_inner == 1 && seq_outer_stride == 48 && step_stride == 1 && steps <= 48,
                 "dj_lstm_scan_tc_bwd: units=128 expects the note-axis map (seq=(b,t))");
    const int max_cl = dj_num_sms() / 2;
    if (S <= 32 * max_cl) return launch_tc_bwd<128, 32>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, 0, hard, st);
    if (S <= 64 * max_cl) return launch_tc_bwd<128, 64>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, 0, hard, st);
    return launch_tc_bwd<128, 128>(Un_bf16, gates, c, dY, ldY, d_y, dZ_bf16, db, S, steps, map, 0, hard, st);
  }
  DJ_CHECK_ARG(false, "dj_lstm_scan_tc_bwd: units=%d unsupported (128 or 256)", units);
  return -1;
}
