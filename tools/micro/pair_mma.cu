// Micro-test of the data flow a CTA-pair reverse scan step would use (lead 1):
//   cluster of 4 CTAs = 2 pairs; CTA c owns 64 rows of A (= U rows) and produces one 64-column K atom of B (= dz);
//   every CTA multicasts its atom -- sequences [0, NH) to the even CTAs, [NH, 2 NH) to the odd CTAs --, the odd CTA of
//   a pair forwards "my half landed" to its leader, the leader issues tcgen05.mma.cta_group::2 (M = 128, N = 2 NH)
//   and commits to the accumulator barrier of its pair and to a count-2 "free" barrier of all four CTAs.
// Prints where every accumulator element landed (lane, column) against the expected layout
//   unit = 128*pair + 64*(c & 1) + (lane & 63),  sequence = (lane >> 6) * NH + column.
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/micro/pair_mma.bin tools/micro/pair_mma.cu && tools/micro/pair_mma.bin
#include <cooperative_groups.h>
#include <cstdarg>
#include <cstdlib>
#include <vector>

#include "../../music-generator_b200/csrc/dj_tc.cuh"

namespace cg = cooperative_groups;
static char g_err[512];
void dj_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int NH = 24;            // sequences per CTA of a pair (MMA N = 48)
constexpr int KTOT = 256;         // 4 atoms of 64
constexpr int A_BYTES = 4 * 64 * 128, B_BYTES = 4 * NH * 128;

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// cta_group::2 forms: the mbarrier operand may name the barrier of the destination CTA's PEER (here: bit 24 of the
// shared::cluster address cleared = the even CTA of the pair), so loads landing in an odd CTA signal its leader
__device__ __forceinline__ void tma_load_2d_mc2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(dst),
      "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
      "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// out[c][lane][col] (col < 32), rounds = how many MMA rounds to run (tests the free barrier's phase too)
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int rounds, int direct) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t a_off = sbase, b_off = sbase + A_BYTES, bars = b_off + B_BYTES;
  const uint32_t bar_a = bars, bar_z = bars + 8, bar_peer = bars + 16, bar_acc = bars + 24, bar_free = bars + 32;
  uint32_t* tmem_slot = (uint32_t*)(smem + A_BYTES + B_BYTES + 64);
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = (c & 1) == 0;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar_a, 1); mbar_init(bar_z, 1); mbar_init(bar_peer, 1); mbar_init(bar_acc, 1); mbar_init(bar_free, 2);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc2(smem_u32(tmem_slot), 32);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster.sync();
  if (warp == 0 && direct) {
    // direct mode: every load that lands in an odd CTA signals the barrier of its LEADER (cta_group::2 TMA); the odd
    // CTA neither waits for its operands nor forwards anything
    if (c == 0 && lane == 0) printf("bar_z address in CTA 0: 0x%x\n", bar_z);
    if (c == 1 && lane == 0) printf("bar_z address in CTA 1: 0x%x\n", bar_z);
    if (elect_one()) {
      if (leader) mbar_expect_tx(bar_a, 2 * A_BYTES);
      for (int ka = 0; ka < 4; ++ka) tma_load_2d_2(a_off + ka * (64 * 128), &tmA, bar_a, ka * 64, 64 * c);
    }
    __syncwarp();
    for (int r = 0; r < rounds; ++r) {
      const uint32_t par = (uint32_t)r & 1u;
      if (r > 0) mbar_wait(bar_free, par ^ 1u);
      if (elect_one()) {
        if (leader) mbar_expect_tx(bar_z, 2 * B_BYTES);
        tma_load_2d_mc2(b_off + c * (NH * 128), &tmB, bar_z, c * 64, 0, (uint16_t)0b0101);
        tma_load_2d_mc2(b_off + c * (NH * 128), &tmB, bar_z, c * 64, NH, (uint16_t)0b1010);
      }
      __syncwarp();
      if (leader) {
        if (r == 0) mbar_wait(bar_a, 0);
        mbar_wait(bar_z, par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc = make_idesc(128, 2 * NH, 0, 0);
          const uint64_t adesc0 = make_smem_desc(a_off, 16, 1024), bdesc0 = make_smem_desc(b_off, 16, 1024);
          for (int ka = 0; ka < 4; ++ka)
            for (int k = 0; k < 4; ++k)
              umma2_bf16(tmem_base, adesc0 + (uint64_t)((ka * (64 * 128) + k * 32) >> 4),
                         bdesc0 + (uint64_t)((ka * (NH * 128) + k * 32) >> 4), idesc, (ka | k) != 0);
          umma2_commit_mc(bar_acc, (uint16_t)(0b11 << (c & ~1)));
          umma2_commit_mc(bar_free, (uint16_t)0b1111);
        }
        __syncwarp();
      }
      mbar_wait(bar_acc, par);
    }
    mbar_wait(bar_free, (uint32_t)(rounds - 1) & 1u);
  } else
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bar_a, A_BYTES);
      for (int ka = 0; ka < 4; ++ka) tma_load_2d(a_off + ka * (64 * 128), &tmA, bar_a, ka * 64, 64 * c);
    }
    __syncwarp();
    for (int r = 0; r < rounds; ++r) {
      const uint32_t par = (uint32_t)r & 1u;
      if (r > 0) mbar_wait(bar_free, par ^ 1u);          // both pairs finished reading round r-1's tiles
      if (elect_one()) {
        mbar_expect_tx(bar_z, B_BYTES);
        // my K atom `c`: sequences [0, NH) -> even CTAs, [NH, 2 NH) -> odd CTAs (same destination offset in each)
        tma_load_2d_mc(b_off + c * (NH * 128), &tmB, bar_z, c * 64, 0, (uint16_t)0b0101);
        tma_load_2d_mc(b_off + c * (NH * 128), &tmB, bar_z, c * 64, NH, (uint16_t)0b1010);
      }
      __syncwarp();
      if (r == 0) mbar_wait(bar_a, 0);
      mbar_wait(bar_z, par);
      if (!leader) {
        // A's loads of round 0 are covered too: this arrive follows the wait on bar_a in program order
        if (elect_one()) mbar_arrive_remote(bar_peer, (uint32_t)(c & ~1));
        __syncwarp();
      } else {
        mbar_wait_cluster(bar_peer, par);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc = make_idesc(128, 2 * NH, 0, 0);
          const uint64_t adesc0 = make_smem_desc(a_off, 16, 1024), bdesc0 = make_smem_desc(b_off, 16, 1024);
          for (int ka = 0; ka < 4; ++ka)
            for (int k = 0; k < 4; ++k)
              umma2_bf16(tmem_base, adesc0 + (uint64_t)((ka * (64 * 128) + k * 32) >> 4),
                         bdesc0 + (uint64_t)((ka * (NH * 128) + k * 32) >> 4), idesc, (ka | k) != 0);
          umma2_commit_mc(bar_acc, (uint16_t)(0b11 << (c & ~1)));
          umma2_commit_mc(bar_free, (uint16_t)0b1111);
        }
        __syncwarp();
      }
      mbar_wait(bar_acc, par);      // every phase of both barriers is observed in order (a parity wait sees one phase back)
    }
    mbar_wait(bar_free, (uint32_t)(rounds - 1) & 1u);
  }
  __syncthreads();
  tc_fence_after();
  for (int c8 = 0; c8 < 32; c8 += 8) {
    uint32_t v[8];
    tmem_ld8(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)c8, v);
    for (int j = 0; j < 8; ++j) out[((size_t)c * 128 + threadIdx.x) * 32 + c8 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  cluster.sync();
  if (warp == 0) tmem_dealloc2(tmem_base, 32);
}

static int run() {
  const int M = 256, N = 2 * NH;
  std::vector<__nv_bfloat16> A((size_t)M * KTOT), B((size_t)N * KTOT);
  std::vector<float> Af(A.size()), Bf(B.size());
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { Af[i] = (float)(rand() % 9 - 4); A[i] = __float2bfloat16(Af[i]); }
  for (size_t i = 0; i < B.size(); ++i) { Bf[i] = (float)(rand() % 7 - 3); B[i] = __float2bfloat16(Bf[i]); }
  std::vector<float> D((size_t)M * N);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < KTOT; ++k) s += Af[(size_t)m * KTOT + k] * Bf[(size_t)n * KTOT + k];
      D[(size_t)m * N + n] = s;
    }
  __nv_bfloat16 *dA, *dB;
  float* dout;
  DJ_CUDA(cudaMalloc(&dA, A.size() * 2)); DJ_CUDA(cudaMalloc(&dB, B.size() * 2));
  DJ_CUDA(cudaMalloc(&dout, 4 * 128 * 32 * 4));
  DJ_CUDA(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  DJ_CUDA(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = make_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, KTOT, M, KTOT, 64, 64))) return rc;
  if ((rc = make_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, KTOT, N, KTOT, 64, NH))) return rc;
  const int smem = A_BYTES + B_BYTES + 128 + 1024;
  DJ_CUDA(cudaFuncSetAttribute((const void*)pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int variant = 0; variant < 4; ++variant) {
    const int rounds = 1 + 2 * (variant & 1), direct = variant >> 1;
    printf("--- %s signalling, %d round(s)\n", direct ? "direct (cta_group::2 TMA -> leader barrier)" : "forwarded", rounds);
    DJ_CUDA(cudaMemset(dout, 0xff, 4 * 128 * 32 * 4));
    pair_kernel<<<4, 128, smem>>>(tmA, tmB, dout, rounds, direct);
    DJ_CUDA(cudaGetLastError());
    DJ_CUDA(cudaDeviceSynchronize());
    std::vector<float> out(4 * 128 * 32);
    DJ_CUDA(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int c = 0; c < 4; ++c)
      for (int lane = 0; lane < 128; ++lane)
        for (int col = 0; col < NH; ++col) {
          const int m = 128 * (c >> 1) + 64 * (c & 1) + (lane & 63), n = (lane >> 6) * NH + col;
          const float got = out[((size_t)c * 128 + lane) * 32 + col], want = D[(size_t)m * N + n];
          if (got != want && bad++ < 6) printf("  rounds %d: cta %d lane %d col %d: got %g want %g\n", rounds, c, lane, col, got, want);
        }
    printf("rounds %d: %s (%d mismatches of %d)\n", rounds, bad ? "LAYOUT MISMATCH" : "pair MMA layout as expected", bad, 4 * 128 * NH);
    if (bad) {   // where DID the values go?  report, for CTA 0, which (m, n) each of a few positions holds
      for (int lane : {0, 1, 16, 32, 63, 64, 65, 96, 127})
        for (int col : {0, 1, 23}) {
          const float got = out[((size_t)0 * 128 + lane) * 32 + col];
          int hits = 0, hm = -1, hn = -1;
          for (int m = 0; m < M; ++m)
            for (int n = 0; n < N; ++n)
              if (D[(size_t)m * N + n] == got) { if (!hits) { hm = m; hn = n; } ++hits; }
          printf("  cta 0 lane %3d col %2d = %g -> first match (m %d, n %d), %d candidates\n", lane, col, got, hm, hn, hits);
        }
    }
  }
  return 0;
}

int main() {
  int rc = run();
  if (rc) printf("error %d: %s\n", rc, g_err);
  return rc;
}
