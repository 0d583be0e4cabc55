// Experiment E1: can a K-major SWIZZLE_128B UMMA operand start at a 128-byte row offset inside a TMA-written tile (a pixel-shifted window of a halo
// tile), and what must the descriptor's base_offset field hold? One CTA: TMA loads A_all[144 rows][64 bf16] and B[64][64]; for each shift s in 0..15
// and each base_offset candidate it computes D = A_all[s .. s+127] * B^T with four K=16 MMAs and compares with the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o e1_shift_test e1_shift_test.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }

constexpr int ROWS = 144;

// out[variant][s][128][64] fp32; variant 0: base_offset = 0, variant 1: base_offset = s & 7
__global__ void __launch_bounds__(128, 1) e1_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out, int sbo) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + ROWS * 128, bars = sb + 64 * 128;
  const uint32_t full = bars, done = bars + 8, slot = bars + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(full, 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_expect_tx(full, ROWS * 128 + 64 * 128);
    tma_load_2d(sa, &map_a, full, 0, 0);
    tma_load_2d(sb, &map_b, full, 0, 0);
  }
  mbar_wait(full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant)
    for (int s = 0; s < 16; ++s) {
      if (threadIdx.x == 0) {
        const uint32_t bo = variant ? (uint32_t)(s & 7) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_desc(sa + s * 128 + k * 32, (uint32_t)sbo, bo), bd = make_desc(sb + k * 32, 1024, 0);
          const uint32_t acc = k != 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad), "l"(bd), "r"(make_idesc(128, 64)), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
      }
      mbar_wait(done, phase);
      phase ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
      float* o = out + (((size_t)variant * 16 + s) * 128 + warp * 32 + lane) * 64;
      for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) o[c0 + j] = __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  std::vector<__nv_bfloat16> ha(ROWS * 64), hb(64 * 64);
  std::vector<float> fa(ROWS * 64), fb(64 * 64);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(v); fa[i] = __bfloat162float(ha[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { float v = (rand() % 13 - 6) / 4.f; hb[i] = __float2bfloat16(v); fb[i] = __bfloat162float(hb[i]); }
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 2 * 16 * 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb; cuuint32_t es[2] = {1, 1};
  { cuuint64_t d[2] = {64, ROWS}, st[1] = {128}; cuuint32_t b[2] = {64, ROWS};
    enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  { cuuint64_t d[2] = {64, 64}, st[1] = {128}; cuuint32_t b[2] = {64, 64};
    enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  const size_t smem = 1024 + ROWS * 128 + 64 * 128 + 64;
  cudaFuncSetAttribute(e1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> ho(2 * 16 * 128 * 64);
  // SBO = 1024: rows m, m+8, ... of the window are 1024 B apart (a plain row-shifted window of a contiguous tile)
  e1_kernel<<<1, 128, smem>>>(ma, mb, dout, 1024);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  for (int variant = 0; variant < 2; ++variant)
    for (int s = 0; s < 16; ++s) {
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)fa[(m + s) * 64 + k] * fb[n * 64 + k];
          worst = fmax(worst, fabs(ref - ho[(((size_t)variant * 16 + s) * 128 + m) * 64 + n]));
        }
      printf("base_offset=%s shift=%2d max_abs_err=%g\n", variant ? "s&7" : "0  ", s, worst);
    }
  return 0;
}
