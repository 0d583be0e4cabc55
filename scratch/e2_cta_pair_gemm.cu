// Experiment E2 (NOT part of the library; drafted at the end of round 1 without a GPU at hand, compile-checked only — to be run and debugged in round 2):
// the mechanics of a CTA PAIR (cta_group::2) tcgen05 GEMM, the step DESIGN.md section 9 ranks first. D[256 x 256] = A[256 x K] * B[256 x K]^T, bf16 -> fp32.
//
//   cluster of 2 CTAs (same TPC). CTA r holds rows [128 r, 128 r + 128) of A and rows [128 r, 128 r + 128) of B (= HALF of the N dimension) in ITS shared
//   memory, at the same offsets in both CTAs; the leader (rank 0) issues ONE tcgen05.mma.cta_group::2 with M = 256, N = 256 per K = 16: each SM multiplies
//   its 128 A rows by all 256 B rows (its own half + the peer's half, fetched over the pair's operand path) into ITS tensor memory (128 lanes x 256 columns).
//   Per SM and k-block that is 16 KB (A) + 16 KB (B half) through the shared-memory port instead of 16 + 32 KB: the measured single-CTA ceiling
//   (0.67-0.75 of peak at N_TILE = 256) should move to ~1.0.
//
// What this file is meant to pin down on the GPU, in this order (each has an `EXPECT` line in main):
//   1. tcgen05.alloc.cta_group::2 executed by one warp of BOTH CTAs; the two base addresses are equal.
//   2. TMA loads issued by each CTA into its own shared memory but signalling the LEADER's mbarrier (`.cta_group::2` form, barrier address with the
//      peer bit cleared); the leader expects the bytes of both CTAs.
//   3. tcgen05.mma.cta_group::2 with an M = 256 instruction descriptor, issued by the leader only.
//   4. tcgen05.commit.cta_group::2 ... multicast::cluster with mask 0b11: the `done` barrier of BOTH CTAs flips.
//   5. each CTA reads its own 128 accumulator rows.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o e2_cta_pair_gemm e2_cta_pair_gemm.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int M_TOTAL = 256, N_TOTAL = 256, K_TOTAL = 256, BLOCK_K = 64, KBLOCKS = K_TOTAL / BLOCK_K;
constexpr int A_BYTES = 128 * BLOCK_K * 2, BH_BYTES = 128 * BLOCK_K * 2;   // per CTA and k-block: 128 A rows, 128 B rows (half of N)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t out;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_addr), "r"(rank));
  return out;
}
// 2-SM TMA load: data lands in the issuing CTA's shared memory, the transaction bytes are credited to the mbarrier at `bar_cluster_addr`
// (a shared::cluster address: the leader's barrier for both CTAs).
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
e2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out, uint32_t* info) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + KBLOCKS * A_BYTES, bars = sb + KBLOCKS * BH_BYTES;
  const uint32_t full = bars, done = bars + 8, slot = bars + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (threadIdx.x == 0) {
    mbar_init(full, 1);   // the leader's `full` collects the bytes of both CTAs (one arrive.expect_tx by the leader)
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();   // both CTAs' barriers exist before anybody signals the peer's
  if (warp == 0) {  // (1) pair allocation, one warp of each CTA
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) info[rank] = tmem;

  if (warp == 0 && elect_one()) {
    // (2) every CTA loads ITS halves of A and B for all k-blocks; bytes are credited to the leader's `full`
    const uint32_t leader_full = map_to_rank(full, 0);
    if (rank == 0) mbar_expect_tx(full, 2u * KBLOCKS * (A_BYTES + BH_BYTES));
    for (int kb = 0; kb < KBLOCKS; ++kb) {
      tma_load_2d_2sm(sa + kb * A_BYTES, &map_a, leader_full, kb * BLOCK_K, (int)rank * 128);
      tma_load_2d_2sm(sb + kb * BH_BYTES, &map_b, leader_full, kb * BLOCK_K, (int)rank * 128);
    }
  }
  if (warp == 1 && rank == 0 && elect_one()) {
    // (3) the leader issues the pair MMAs: descriptors are shared-memory OFFSETS valid in both CTAs
    mbar_wait(full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = make_idesc(M_TOTAL, N_TOTAL);
    for (int kb = 0; kb < KBLOCKS; ++kb)
      for (int k = 0; k < BLOCK_K / 16; ++k) {
        const uint64_t ad = make_desc(sa + kb * A_BYTES + k * 32, 1024), bd = make_desc(sb + kb * BH_BYTES + k * 32, 1024);
        const uint32_t acc = (kb | k) != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
      }
    // (4) completion to the `done` barrier of BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(done), "h"((uint16_t)0b11) : "memory");
  }
  // (5) every CTA reads its own 128 rows
  mbar_wait(done, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  float* o = out + ((size_t)rank * 128 + warp * 32 + lane) * N_TOTAL;
  for (int c0 = 0; c0 < N_TOTAL; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) o[c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();   // nobody frees tensor memory the peer's MMA may still write
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  std::vector<__nv_bfloat16> ha((size_t)M_TOTAL * K_TOTAL), hb((size_t)N_TOTAL * K_TOTAL);
  std::vector<float> fa(ha.size()), fb(hb.size());
  srand(2);
  for (size_t i = 0; i < ha.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db; float* dout; uint32_t* dinfo;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, (size_t)M_TOTAL * N_TOTAL * 4); cudaMalloc(&dinfo, 8);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, (size_t)M_TOTAL * N_TOTAL * 4);
  CUtensorMap ma, mb; cuuint32_t es[2] = {1, 1};
  { cuuint64_t d[2] = {(cuuint64_t)K_TOTAL, (cuuint64_t)M_TOTAL}, st[1] = {(cuuint64_t)K_TOTAL * 2}; cuuint32_t b[2] = {(cuuint32_t)BLOCK_K, 128};
    enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  { cuuint64_t d[2] = {(cuuint64_t)K_TOTAL, (cuuint64_t)N_TOTAL}, st[1] = {(cuuint64_t)K_TOTAL * 2}; cuuint32_t b[2] = {(cuuint32_t)BLOCK_K, 128};
    enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  const size_t smem = 1024 + (size_t)KBLOCKS * (A_BYTES + BH_BYTES) + 64;
  cudaFuncSetAttribute(e2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  e2_kernel<<<2, 128, smem>>>(ma, mb, dout, dinfo);   // one cluster of two CTAs (__cluster_dims__)
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  uint32_t hinfo[2] = {0, 0};
  cudaMemcpy(hinfo, dinfo, 8, cudaMemcpyDeviceToHost);
  printf("EXPECT equal tensor-memory bases in both CTAs: %u %u\n", hinfo[0], hinfo[1]);
  std::vector<float> ho((size_t)M_TOTAL * N_TOTAL);
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  for (int half = 0; half < 2; ++half) {
    double worst = 0;
    for (int m = half * 128; m < half * 128 + 128; ++m)
      for (int n = 0; n < N_TOTAL; ++n) {
        double ref = 0;
        for (int k = 0; k < K_TOTAL; ++k) ref += (double)fa[(size_t)m * K_TOTAL + k] * fb[(size_t)n * K_TOTAL + k];
        worst = fmax(worst, fabs(ref - ho[(size_t)m * N_TOTAL + n]));
      }
    printf("EXPECT 0: rows of CTA %d max_abs_err=%g\n", half, worst);
  }
  return 0;
}
