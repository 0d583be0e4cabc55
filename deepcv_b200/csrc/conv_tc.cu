// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
// GEMM view of y[n][p][q][k] = sum_{r,s,c} x[n][p+r-pad][q+s-pad][c] * w[k][r][s][c]   (stride 1, dilation 1, NHWC / KRSC):
//     M = output pixels (128 per tile),  N = output channels (N_TILE = 64 / 128 / 256 per tile),  K = R*S*C walked as (r, s, 64-channel block).
//   A operand: for one (r, s, channel block) the 128 x 64 im2col slab is ONE 4-D TMA box {64 ch, tw, th, tn} of x whose start
//     coordinate is shifted by (s - pad, r - pad); TMA zero-fills whatever falls outside the image (the convolution padding) or beyond the
//     batch. The box lands in shared memory as 128 rows of 128 B, hardware-swizzled (SWIZZLE_128B) = the K-major UMMA operand layout.
//     (tw, th, tn) with tw*th*tn = 128 is picked per layer to minimise the padded overhang of the P x Q plane.
//   B operand: weights viewed as the 2-D matrix [K_out][R*S*C]; a 2-D TMA box {64, N_TILE} at column (r*S+s)*C + c0 is the K-major B slab.
//   D: 128 lanes x N_TILE fp32 columns of tensor memory, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles (192 threads, persistent CTAs, one per SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane; also owns the TMEM
// allocation), warps 2..5 = epilogue (each owns the TMEM lane quarter warp%4: tcgen05.ld -> bias + activation -> bf16 -> 16-byte global stores).
// Pipelines: shared-memory ring full/empty mbarriers (TMA <-> MMA, released by tcgen05.commit), TMEM full/empty mbarriers (MMA <-> epilogue).
//
// The same kernel computes the data gradient (x := dy, w := transposed + flipped weights, pad := R-1-pad). The weight gradient kernel below
// contracts over pixels instead: dw[k][r][s][c] = sum_pix dy[pix][k] * x[pix + (r,s)][c], with both operands MN-major (channels contiguous).
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace dcv {
namespace tc {

constexpr int BLOCK_M = 128;   // pixels per tile = UMMA M = TMEM lanes
constexpr int BLOCK_K = 64;    // channels per k-block: 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
// One lane of a converged warp. With `if (lane == 0)` ptxas cannot prove that a single thread executes the warp-level tcgen05 / TMA instructions and wraps
// each of them in an ELECT + BRA.U.ANY loop (4 extra instructions per MMA: the issue loop, not the tensor pipe, paced the 64-channel layers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// Same, for the many-thread waiters (epilogue / producer warps): back off between polls so that spinning warps do not eat the issue slots of the
// warps that have work (ncu on the gather kernel: 20 % of all warp instructions were try_wait polls).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t done;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    __nanosleep(40);
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned, size a multiple of 16), completion on an mbarrier like the tensor variants
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
constexpr int kZeroRowBytes = 8192;
__device__ __align__(16) unsigned char g_zero_row[kZeroRowBytes];   // source of the rows above / below the image (vertical padding)
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// Shared-memory matrix descriptor (SWIZZLE_128B). K-major operands: 8-row groups 1024 B apart (SBO), LBO unused. MN-major operands: rows
// are K indices (128 B each = 64 MN elements), SBO = 1024 B between 8-row K groups, LBO = byte distance between 64-element MN atoms.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: fp32 accumulate, bf16 A and B, M x N tile, operand majors (0 = K-major, 1 = MN-major).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                 "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
                 "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Column sums of a warp's 32 x 32 block held one row per lane, by recursive halving: after the step with distance d a lane keeps the half of its current
// columns selected by (lane & d) and adds its partner's copy of them — 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 5 x 32. On return lane l holds the sum
// of column l over the 32 rows (v is destroyed).
__device__ __forceinline__ float warp_colsum32(float* v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const bool up = (lane & d) != 0;
#pragma unroll
    for (int i = 0; i < d; ++i) {
      const float send = up ? v[i] : v[i + d];
      const float recv = __shfl_xor_sync(0xffffffffu, send, d);
      v[i] = (up ? v[i + d] : v[i]) + recv;
    }
  }
  return v[0];
}

// Epilogue of 32 accumulator columns of one pixel: bias (vector loads, hoisted out of the element loop) + activation resolved at compile
// time + packed bf16 conversion + two 32-byte stores. STATS: the BatchNorm batch statistics ride along (north star: "BatchNorm statistics ... fused into
// the conv epilogue"): sum and sum of squares of the STORED (bf16-rounded) values of each of the 32 channels over the warp's 32 pixels (rows that fall
// outside the tensor contribute zero), added to the caller's running totals (s1, s2) of channel `lane` of this 32-channel chunk — registers kept across
// all the tiles of the CTA and flushed with one global atomic each at the end (a first version added them to shared-memory totals per tile: fp32 shared
// atomics are compare-and-swap loops and the four epilogue warps hit the same addresses — the epilogue became the critical path, +46 us on a 77 us
// kernel). Every lane of the warp must call it (shuffles); only `valid` rows store.
// STATS = 1: per-tile column sums (62 shuffles per chunk) into run[0], run[1] (lane = channel). STATS = 2: the thread's own 32 + 32 partial sums
// run[0..31], run[32..63] (its pixel row's values), summed over the warp ONCE at the end of the kernel — for N_TILE = 64, where a tile's MMAs are too
// short to hide a per-tile butterfly (measured: the 64 -> 64 layer at 56x56 went from 77 to 123 us with it).
template <int ACT, int STATS>
__device__ __forceinline__ void epilogue_store32(const uint32_t* v, const float* __restrict__ bias32, float slope, __nv_bfloat16* dst, bool valid, float* run) {
  if (!STATS && !valid) return;
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
  if (bias32) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias32 + j));
      f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (ACT == DCV_ACT_RELU) f[j] = fmaxf(f[j], 0.f);
    else if (ACT == DCV_ACT_LEAKY_RELU) f[j] = f[j] > 0.f ? f[j] : f[j] * slope;
    else if (ACT == DCV_ACT_SIGMOID) f[j] = 1.f / (1.f + expf(-f[j]));
  }
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) { const uint4 q = vec_pack<__nv_bfloat16>(f + 8 * j); pk[4 * j] = q.x; pk[4 * j + 1] = q.y; pk[4 * j + 2] = q.z; pk[4 * j + 3] = q.w; }
  if (STATS == 1) {
    float sq[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      f[2 * j] = valid ? __uint_as_float(pk[j] << 16) : 0.f;
      f[2 * j + 1] = valid ? __uint_as_float(pk[j] & 0xffff0000u) : 0.f;
      sq[2 * j] = f[2 * j] * f[2 * j]; sq[2 * j + 1] = f[2 * j + 1] * f[2 * j + 1];
    }
    run[0] += warp_colsum32(f);
    run[1] += warp_colsum32(sq);
    if (!valid) return;
  } else if (STATS == 2) {
    if (!valid) return;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float lo = __uint_as_float(pk[j] << 16), hi = __uint_as_float(pk[j] & 0xffff0000u);
      run[2 * j] += lo; run[2 * j + 1] += hi;
      run[32 + 2 * j] = fmaf(lo, lo, run[32 + 2 * j]); run[32 + 2 * j + 1] = fmaf(hi, hi, run[32 + 2 * j + 1]);
    }
  }
  if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
    // a lane owns 64 contiguous bytes of its pixel's row, the lanes of a warp are >= 128 bytes apart: two 256-bit stores (sm_100 STG.256) instead of four
    // 128-bit ones halve the store requests (the uint8 preprocess kernel went from 62 % to 88 % of HBM bandwidth with the same change)
#pragma unroll
    for (int w = 0; w < 16; w += 8)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(reinterpret_cast<uint32_t*>(dst) + w), "r"(pk[w]), "r"(pk[w + 1]), "r"(pk[w + 2]), "r"(pk[w + 3]),
                   "r"(pk[w + 4]), "r"(pk[w + 5]), "r"(pk[w + 6]), "r"(pk[w + 7]) : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  }
}

template <int STATS>
__device__ __forceinline__ void epilogue_store32_dyn(int act, const uint32_t* v, const float* bias32, float slope, __nv_bfloat16* dst, bool valid, float* run) {
  switch (act) {   // warp-uniform
    case DCV_ACT_RELU: epilogue_store32<DCV_ACT_RELU, STATS>(v, bias32, slope, dst, valid, run); break;
    case DCV_ACT_LEAKY_RELU: epilogue_store32<DCV_ACT_LEAKY_RELU, STATS>(v, bias32, slope, dst, valid, run); break;
    case DCV_ACT_SIGMOID: epilogue_store32<DCV_ACT_SIGMOID, STATS>(v, bias32, slope, dst, valid, run); break;
    default: epilogue_store32<DCV_ACT_NONE, STATS>(v, bias32, slope, dst, valid, run); break;
  }
}

// Running channel totals of one epilogue warp over every tile the CTA has finished for output-channel tile nt; `flush` adds them to stats[k][2] (one
// global atomic per value) when nt changes and at the end. N_TILE >= 128: lane l holds {sum, sum of squares} of channel 32 * chunk + l (per-tile
// butterfly). N_TILE = 64: every thread holds the 2 x 32 partial sums of its own pixel row per chunk (128 registers) and the butterfly runs in flush.
constexpr int kMaxStatChannels = 512;
template <int N_TILE, int MODE_ = 1> struct RunStats {
  static constexpr int MODE = MODE_, PER = MODE == 2 ? 64 : 2;
  float v[N_TILE / 32][PER];
  int nt;
  __device__ __forceinline__ void reset(int nt_) {
    nt = nt_;
#pragma unroll
    for (int i = 0; i < N_TILE / 32; ++i)
#pragma unroll
      for (int j = 0; j < PER; ++j) v[i][j] = 0.f;
  }
  __device__ __forceinline__ void flush(float* stats) {
    if (nt < 0) return;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < N_TILE / 32; ++i) {
      const float s1 = MODE == 2 ? warp_colsum32(v[i]) : v[i][0];
      const float s2 = MODE == 2 ? warp_colsum32(v[i] + 32) : v[i][1];
      atomicAdd(stats + 2 * (nt * N_TILE + 32 * i + lane), s1);
      atomicAdd(stats + 2 * (nt * N_TILE + 32 * i + lane) + 1, s2);
    }
  }
};

struct FwdParams {
  int n, h, w, c, k, r, s, pad_h, pad_w, p, q;
  int tw, th, tn, tiles_w, tiles_h, tiles_n;   // pixel tile geometry
  int n_tiles_k;                               // K_out / N_TILE
  int total_tiles;
  int act; float slope;
  const float* bias;
  __nv_bfloat16* y;
  float* stats;   // non-null: per-channel {sum, sum of squares} of y are added to stats[k][2] by the epilogue (k <= kMaxStatChannels)
};

// A pipeline stage holds G consecutive k-blocks (G A slabs + G B slabs) behind ONE full/empty barrier pair: with 64 output channels a k-block is only
// 4 x 32 MMA cycles, and the single-thread producer / issuer loops (mbarrier try_wait ~90 cycles each) would otherwise dominate.
// RES = 1 (64 -> 64 channel 3x3 layers: all R*S*C/64 <= 9 weight slabs = 72 KB): the weights are loaded ONCE per CTA into a resident region and the
// ring only carries the A slabs. These layers are bound by L2 -> SM traffic (24 KB per 1 MFLOP k-block = 44 FLOP/B against a ~12 TB/s TMA ceiling);
// dropping the 8 KB weight slab per k-block raises that to 66 FLOP/B.
constexpr int kMaxResidentKb = 9;
template <int N_TILE, int RES = 0>
struct FwdSmem {
  static constexpr int G = N_TILE == 64 ? 3 : (N_TILE == 128 ? 2 : 1);
  static constexpr int kStages = N_TILE == 256 ? 4 : 3;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2, B_BYTES = N_TILE * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = RES ? G * A_BYTES : G * (A_BYTES + B_BYTES);
  static constexpr int RES_BYTES = RES ? kMaxResidentKb * B_BYTES : 0;
  static constexpr size_t kBytes = 1024 /* alignment slack */ + (size_t)RES_BYTES + (size_t)kStages * STAGE_BYTES + 256;
};

template <int N_TILE, int RES>
__global__ void __launch_bounds__(kThreads, 1) conv_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const FwdParams prm) {
  using S = FwdSmem<N_TILE, RES>;
  constexpr int kStages = S::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t res_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // resident weight slabs (RES), then the ring
  const uint32_t base = res_base + S::RES_BYTES;
  constexpr int G = S::G;
  const uint32_t bars = base + kStages * S::STAGE_BYTES;        // 8-byte mbarriers
  auto slab_a = [&](int stage, int j) { return base + stage * S::STAGE_BYTES + j * S::A_BYTES; };
  auto slab_b = [&](int stage, int j) { return base + stage * S::STAGE_BYTES + G * S::A_BYTES + j * S::B_BYTES; };
  auto slab_res = [&](int kb) { return res_base + kb * S::B_BYTES; };
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (kStages + i); };
  auto tfull = [&](int i) { return bars + 8u * (2 * kStages + i); };
  auto tempty = [&](int i) { return bars + 8u * (2 * kStages + 2 + i); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  const uint32_t bres = bars + 8u * (2 * kStages + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 4); }
    mbar_init(bres, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * N_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int cblocks = prm.c / BLOCK_K;
  const int num_kb = prm.r * prm.s * cblocks;

  if (warp == 0) {
    // ===== TMA producer (one lane) =====
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      if (RES) {   // all weight slabs, once (RES implies a single output-channel tile)
        mbar_expect_tx(bres, (uint32_t)num_kb * S::B_BYTES);
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(slab_res(kb), &map_w, bres, kb * BLOCK_K, 0);
      }
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
        const int nt = tile % prm.n_tiles_k;
        int pt = tile / prm.n_tiles_k;
        const int pw = pt % prm.tiles_w; pt /= prm.tiles_w;
        const int ph = pt % prm.tiles_h; const int pn = pt / prm.tiles_h;
        const int q0 = pw * prm.tw, p0 = ph * prm.th, n0 = pn * prm.tn;
        // k-block cursor (r, s, channel block), channel block fastest. Every CTA starts at a different k-block (the accumulation order is free): all
        // 148 CTAs would otherwise request the same weight slab from the same L2 slice at the same time.
        const int kstart = (int)(((unsigned)blockIdx.x * 2654435761u) % (unsigned)num_kb);
        int cb = kstart % cblocks, ss = (kstart / cblocks) % prm.s, rr = kstart / (cblocks * prm.s);
        for (int kb0 = 0; kb0 < num_kb; kb0 += G) {
          const int cnt = min(G, num_kb - kb0);
          mbar_wait(empty(stage), phase ^ 1u);
          mbar_expect_tx(full(stage), (uint32_t)cnt * (RES ? S::A_BYTES : S::A_BYTES + S::B_BYTES));
          for (int j = 0; j < cnt; ++j) {
            tma_load_4d(slab_a(stage, j), &map_x, full(stage), cb * BLOCK_K, q0 + ss - prm.pad_w, p0 + rr - prm.pad_h, n0);
            if (!RES) tma_load_2d(slab_b(stage, j), &map_w, full(stage), (rr * prm.s + ss) * prm.c + cb * BLOCK_K, nt * N_TILE);
            if (++cb == cblocks) { cb = 0; if (++ss == prm.s) { ss = 0; if (++rr == prm.r) rr = 0; } }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane) =====
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, N_TILE, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      // the producer starts every tile at k-block kstart (see there); with resident weights the slab index must follow it
      const int kstart = (int)(((unsigned)blockIdx.x * 2654435761u) % (unsigned)num_kb);
      if (RES) mbar_wait(bres, 0);
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * N_TILE);
        int kb = kstart;
        for (int kb0 = 0; kb0 < num_kb; kb0 += G) {
          const int cnt = min(G, num_kb - kb0);
          mbar_wait(full(stage), phase);
          tc_fence_after();
          for (int j = 0; j < cnt; ++j) {
            const uint64_t adesc = make_desc(slab_a(stage, j), 0, 1024);
            const uint64_t bdesc = make_desc(RES ? slab_res(kb) : slab_b(stage, j), 0, 1024);
            if (++kb == num_kb) kb = 0;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16(d_tmem, adesc + (uint64_t)(k * UMMA_K * 2 / 16), bdesc + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (kb0 | j | k) != 0);
          }
          umma_commit(empty(stage));          // frees the stage once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull(as));               // accumulator complete -> epilogue
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps: TMEM -> registers -> bias + activation -> bf16 -> global =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int wl = row % prm.tw, hl = (row / prm.tw) % prm.th, nl = row / (prm.tw * prm.th);
    int as = 0; uint32_t aphase = 0;
    const bool stats = prm.stats != nullptr;
    RunStats<N_TILE> run; run.reset(-1);
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
      const int nt = tile % prm.n_tiles_k;
      if (stats && nt != run.nt) { run.flush(prm.stats); run.reset(nt); }
      int pt = tile / prm.n_tiles_k;
      const int pw = pt % prm.tiles_w; pt /= prm.tiles_w;
      const int ph = pt % prm.tiles_h; const int pn = pt / prm.tiles_h;
      const int q = pw * prm.tw + wl, p = ph * prm.th + hl, n = pn * prm.tn + nl;
      const bool valid = q < prm.q && p < prm.p && n < prm.n;
      __nv_bfloat16* dst = prm.y + (((size_t)n * prm.p + p) * prm.q + q) * prm.k + (size_t)nt * N_TILE;
      mbar_wait(tfull(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(as * N_TILE) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int ci = 0; ci < N_TILE / 32; ++ci) {
        const int c0 = 32 * ci;
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        if (stats) epilogue_store32_dyn<RunStats<N_TILE>::MODE>(prm.act, v, prm.bias ? prm.bias + nt * N_TILE + c0 : nullptr, prm.slope, dst + c0, valid, run.v[ci]);
        else epilogue_store32_dyn<0>(prm.act, v, prm.bias ? prm.bias + nt * N_TILE + c0 : nullptr, prm.slope, dst + c0, valid, nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (stats) run.flush(prm.stats);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * N_TILE) : "memory");
  }
}

// ---- forward / data gradient, halo variant: 64-channel-class 3x3 layers on large maps (the 64 -> 64 layers at 56x56) ------------------------------
// Measured ceiling of the per-tap kernel above at N_TILE = 64: 35-38 % of the tensor peak, and it is the SHARED-MEMORY PORT (128 B/clk/SM), not L2: a
// k-block writes 16 KB (A) + 8 KB (B) through TMA and the four MMAs read the same 24 KB back, 48 KB of port traffic per 128 MMA cycles = 33 %
// (N_TILE 128: 50 %, 256: 67 % — the three measured plateaus). Keeping the weights resident moved it to 40 %. Here the A operand is written ONCE per tile:
//   * pixel tile = 8 wide x th <= 16 tall of one image; ONE TMA box {64 ch, 8 + S - 1, th + R - 1} brings the tile with its halo (TMA zero fill =
//     padding), 180 rows x 128 B for a 3x3 filter instead of nine 16 KB slabs;
//   * the operand of tap (r, s) is the same buffer read through a K-major SWIZZLE_128B descriptor that starts (r * (8 + S - 1) + s) rows = 128-byte
//     steps into it, 8-row groups (8 + S - 1) * 128 B apart. Experiment scratch/e1_shift_test.cu (B200): the tcgen05 swizzle is a function of the
//     absolute shared-memory address, so any 128-byte row offset reads TMA-written data correctly with base_offset = 0 (max error 0 for shifts 0..15);
//   * all R*S*C/64 weight slabs stay resident (<= 144 KB).
// Port traffic per tile: 9 x 24 KB of MMA reads + 23 KB of writes instead of + 144 KB.
struct FwdHaloParams {
  int n, c, k, r, s, pad_h, pad_w, p, q;
  int th, tiles_w, tiles_h, total_tiles, stages;
  int act; float slope;
  const float* bias;
  __nv_bfloat16* y;
  float* stats;   // as FwdParams::stats
};

constexpr int HALO_TW = 8, HALO_STAGE_BYTES = 23 * 1024;   // (16 + 2) x (8 + 2) rows x 128 B = 23040, rounded up to the 1024-byte swizzle period

template <int N_TILE>
__global__ void __launch_bounds__(kThreads, 1) conv_fwd_tc_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const FwdHaloParams prm) {
  constexpr int B_BYTES = N_TILE * BLOCK_K * 2, kMaxStages = 8;
  extern __shared__ uint8_t smem_raw[];
  const int cblocks = prm.c / BLOCK_K, num_kb = prm.r * prm.s * cblocks, stages = prm.stages;
  const uint32_t res_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t ring = res_base + (uint32_t)num_kb * B_BYTES;
  const uint32_t bars = ring + (uint32_t)stages * HALO_STAGE_BYTES;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (kMaxStages + i); };
  auto tfull = [&](int i) { return bars + 8u * (2 * kMaxStages + i); };
  auto tempty = [&](int i) { return bars + 8u * (2 * kMaxStages + 2 + i); };
  const uint32_t bres = bars + 8u * (2 * kMaxStages + 4), tmem_slot = bars + 8u * (2 * kMaxStages + 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int box_w = HALO_TW + prm.s - 1, box_h = prm.th + prm.r - 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 4); }
    mbar_init(bres, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * N_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto decode = [&](int tile, int& q0, int& p0, int& img) {
    q0 = (tile % prm.tiles_w) * HALO_TW; int t = tile / prm.tiles_w;
    p0 = (t % prm.tiles_h) * prm.th; img = t / prm.tiles_h;
  };

  if (warp == 0) {
    if (elect_one()) {   // ===== producer: the weights once, then one halo box per (tile, channel block)
      mbar_expect_tx(bres, (uint32_t)num_kb * B_BYTES);
      for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(res_base + kb * B_BYTES, &map_w, bres, kb * BLOCK_K, 0);
      int stage = 0; uint32_t phase = 0;
      const uint32_t box_bytes = (uint32_t)box_w * box_h * 128u;
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
        int q0, p0, img; decode(tile, q0, p0, img);
        for (int cb = 0; cb < cblocks; ++cb) {
          mbar_wait(empty(stage), phase ^ 1u);
          mbar_expect_tx(full(stage), box_bytes);
          tma_load_4d(ring + stage * HALO_STAGE_BYTES, &map_x, full(stage), cb * BLOCK_K, q0 - prm.pad_w, p0 - prm.pad_h, img);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // ===== MMA issuer
      constexpr uint32_t idesc = make_idesc(BLOCK_M, N_TILE, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      const uint32_t sbo = (uint32_t)box_w * 128u;
      mbar_wait(bres, 0);
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * N_TILE);
        uint32_t accumulate = 0;
        for (int cb = 0; cb < cblocks; ++cb) {
          mbar_wait(full(stage), phase);
          tc_fence_after();
          const uint32_t abuf = ring + stage * HALO_STAGE_BYTES;
          for (int rr = 0; rr < prm.r; ++rr)
            for (int ss = 0; ss < prm.s; ++ss) {
              const uint64_t adesc = make_desc(abuf + (uint32_t)(rr * box_w + ss) * 128u, 0, sbo);
              const uint64_t bdesc = make_desc(res_base + (uint32_t)((rr * prm.s + ss) * cblocks + cb) * B_BYTES, 0, 1024);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, accumulate);
                accumulate = 1;
              }
            }
          umma_commit(empty(stage));
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull(as));
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===== epilogue warps: accumulator row = tile pixel (y = row / 8, x = row % 8)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int yl = row >> 3, xl = row & 7;
    int as = 0; uint32_t aphase = 0;
    const bool stats = prm.stats != nullptr;
    typedef RunStats<N_TILE, N_TILE == 64 ? 2 : 1> RS;   // 64 channels: a tile's MMAs are too short to hide a per-tile butterfly
    RS run; run.reset(0);
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
      int q0, p0, img; decode(tile, q0, p0, img);
      const int q = q0 + xl, p = p0 + yl;
      const bool valid = yl < prm.th && q < prm.q && p < prm.p;
      __nv_bfloat16* dst = prm.y + (((size_t)img * prm.p + p) * prm.q + q) * prm.k;
      mbar_wait(tfull(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(as * N_TILE) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int ci = 0; ci < N_TILE / 32; ++ci) {
        const int c0 = 32 * ci;
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        if (stats) epilogue_store32_dyn<RS::MODE>(prm.act, v, prm.bias ? prm.bias + c0 : nullptr, prm.slope, dst + c0, valid, run.v[ci]);
        else epilogue_store32_dyn<0>(prm.act, v, prm.bias ? prm.bias + c0 : nullptr, prm.slope, dst + c0, valid, nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (stats) run.flush(prm.stats);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * N_TILE) : "memory");
  }
}

// ---- forward, gather variant: convolutions TMA cannot address (few input channels, strides: the 3 -> 64, 7x7 / stride-2 stem) ---------------------
// The explicit im2col route materialises col[n][p][q][kpad] (1.2 GB for the stem at batch 256: 338 us to write it + a GEMM that is HBM-bound reading it
// back = the most expensive layer of the ImageNet-shaped step). Here eight PRODUCER warps build the 128 x kpad im2col tile of one output-row segment
// directly in shared memory, in the K-major SWIZZLE_128B layout TMA would have produced (16-byte chunk j of row m lives at chunk j ^ (m & 7) of its
// 128-byte line): the R input rows the segment reads are staged by ONE bulk copy each (cp.async.bulk, rows outside the image come from a zero row) into
// a triple-buffered area with zeroed margins. (cp.async staging by the producers themselves was 5x slower: the fence.proxy.async they need after
// writing the tile compiles to MEMBAR.ALL.CTA, which waits for their in-flight prefetch of the NEXT tile — a full L2 / DRAM round trip per tile.)
// K order ("row padded"): column kk = r * RP + x, x < RP = S*C rounded up to 8; for x >= S*C the WEIGHT is zero, so the tile may hold whatever follows in
// the staged row there. A 16-byte chunk is then always 8 CONTIGUOUS staged elements at a 2-byte-aligned address: five aligned 32-bit shared loads and a
// funnel shift, no per-element gather, no masks. Lanes of a warp take 32 consecutive pixels of one chunk column (12 bytes apart: conflict-free).
// Weights ([K][kpad] in the same K order) are resident. MMA / TMEM / epilogue as above.
struct GatherParams {
  dcv_conv_shape s;
  int kpad, kblocks, rowlen, lpad, rows_bytes, rp;   // rp: elements per filter row in the K order (S*C rounded up to 8)
  int tiles_q, total_tiles;
  int act; float slope;
  const float* bias;
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  float* stats;   // as FwdParams::stats
};
constexpr int GA_THREADS = 448, GA_STAGES = 2, GA_FWD_STAGES = 3, GA_PRODUCERS = 256, GA_ROW_BUFS = 3;   // warps: 0 weights (TMA), 1 MMA, 2..5 epilogue, 6..13 producers

// Producer warp `pw` of GA_PRODUCERS / 32 builds its share of the (chunk column, 32-row band) units of the tile; lane = row within the band. The
// unit list of a warp is the same for every tile: it is decoded ONCE into registers (ncu on the version that decoded it per unit: three runtime integer
// divisions = ~75 instructions per unit, 63 % of the kernel's 355 M warp instructions, the kernel issue-bound at 5300 cycles per 128-pixel tile).
// (Earlier versions: one thread per row walking an offset table with data-dependent branches; one thread per chunk column with the 8 offsets in
// registers — 8 two-byte loads per chunk with 4-way bank conflicts.)
constexpr int GA_MAX_UNITS = (256 / 8) * (BLOCK_M / 32) / (GA_PRODUCERS / 32);   // 16
struct GatherUnits {
  int n;
  uint32_t dst[GA_MAX_UNITS];   // byte offset of the unit's 16-byte chunk of row `lane` inside the stage, before the swizzle XOR
  int src[GA_MAX_UNITS];        // element offset of the chunk's first element relative to the pixel's window start; -1: zero chunk (K padding)
  int band[GA_MAX_UNITS];
};

__device__ __forceinline__ void gather_units_setup(const GatherParams& prm, int pw, GatherUnits& g) {
  const int ncols = prm.kpad / 8, cpr = prm.rp / 8, real_cols = prm.s.r * cpr, units = ncols * (BLOCK_M / 32);
  g.n = 0;
#pragma unroll
  for (int i = 0; i < GA_MAX_UNITS; ++i) {
    const int u = pw + i * (GA_PRODUCERS / 32);
    g.dst[i] = 0u; g.src[i] = -1; g.band[i] = 0;
    if (u < units) {
      const int cc = u % ncols, band = u / ncols;   // consecutive warps take consecutive chunk columns of one band
      g.n = i + 1;
      g.band[i] = band;
      g.dst[i] = (uint32_t)(cc >> 3) * (BLOCK_M * 128) + (uint32_t)(band * 32) * 128u + (uint32_t)(cc & 7) * 16u;
      if (cc < real_cols) { const int rr = cc / cpr; g.src[i] = rr * prm.rowlen + (cc - rr * cpr) * 8; }
    }
  }
}

__device__ __forceinline__ void gather_build_tile(const GatherParams& prm, const GatherUnits& g, const __nv_bfloat16* rows, uint32_t a_stage, int q0, int lane) {
  const dcv_conv_shape& s = prm.s;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(rows);
  const uint32_t lane_dst = a_stage + (uint32_t)lane * 128u, sw = (uint32_t)(lane & 7) * 16u;   // band * 32 is a multiple of 8: (m & 7) == (lane & 7)
  const int step = s.stride_w * s.c, base0 = prm.lpad - s.pad_w * s.c;
#pragma unroll
  for (int i = 0; i < GA_MAX_UNITS; ++i) {
    if (i < g.n) {
      uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
      if (g.src[i] >= 0) {
        // overhang pixels (q >= Q) read the first pixel's window: finite data, and their accumulator rows are never stored
        const int q = q0 + g.band[i] * 32 + lane;
        const int idx = base0 + (q < s.q ? q : 0) * step + g.src[i];   // first element of the chunk (2-byte units)
        const uint32_t* wp = words + (idx >> 1);
        const uint32_t a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3], a4 = wp[4];
        const uint32_t sh = (uint32_t)(idx & 1) * 16u;
        w0 = __funnelshift_r(a0, a1, sh); w1 = __funnelshift_r(a1, a2, sh); w2 = __funnelshift_r(a2, a3, sh); w3 = __funnelshift_r(a3, a4, sh);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"((lane_dst + g.dst[i]) ^ sw), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
    }
  }
}

template <int N_TILE>
__global__ void __launch_bounds__(GA_THREADS, 1) conv_fwd_tc_gather_kernel(const __grid_constant__ CUtensorMap map_w, const GatherParams prm) {
  constexpr int B_BYTES = N_TILE * BLOCK_K * 2, A_SLAB = BLOCK_M * 128;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t res_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = res_base + (uint32_t)prm.kblocks * B_BYTES;
  const uint32_t a_stage_bytes = (uint32_t)prm.kblocks * A_SLAB;
  const uint32_t rows_base = a_base + GA_FWD_STAGES * a_stage_bytes;
  const uint32_t bars = rows_base + (uint32_t)GA_ROW_BUFS * (uint32_t)prm.rows_bytes;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (GA_FWD_STAGES + i); };
  auto tfull = [&](int i) { return bars + 8u * (2 * GA_FWD_STAGES + i); };
  auto tempty = [&](int i) { return bars + 8u * (2 * GA_FWD_STAGES + 2 + i); };
  const uint32_t bres = bars + 8u * (2 * GA_FWD_STAGES + 4), tmem_slot = bars + 8u * (2 * GA_FWD_STAGES + 5);
  auto rfull = [&](int i) { return bars + 8u * (2 * GA_FWD_STAGES + 6 + i); };    // staged rows of buffer i landed (TMA)
  auto rempty = [&](int i) { return bars + 8u * (2 * GA_FWD_STAGES + 6 + GA_ROW_BUFS + i); };   // all producer warps are done reading buffer i
  uint8_t* gen_base = smem_raw + (res_base - smem_u32(smem_raw));   // generic-address view of the same buffer
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const dcv_conv_shape& s = prm.s;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GA_FWD_STAGES; ++i) { mbar_init(full(i), GA_PRODUCERS / 32); mbar_init(empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 4); }
    mbar_init(bres, 1);
    for (int i = 0; i < GA_ROW_BUFS; ++i) { mbar_init(rfull(i), 1); mbar_init(rempty(i), GA_PRODUCERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * N_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the staging buffers start as zeros: the bulk copies only ever write [lpad, lpad + W*C) of each row, the margins are the convolution padding
  for (uint32_t i = threadIdx.x; i < (uint32_t)GA_ROW_BUFS * (uint32_t)prm.rows_bytes / 16u; i += GA_THREADS)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(rows_base + i * 16u), "r"(0) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto decode = [&](int tile, int& q0, int& op, int& img) {
    q0 = (tile % prm.tiles_q) * BLOCK_M; const int t = tile / prm.tiles_q;
    op = t % s.p; img = t / s.p;
  };

  if (warp == 0) {
    if (elect_one()) {   // ===== weights once, then the input rows of every tile
      mbar_expect_tx(bres, (uint32_t)prm.kblocks * B_BYTES);
      for (int kb = 0; kb < prm.kblocks; ++kb) tma_load_2d(res_base + kb * B_BYTES, &map_w, bres, kb * BLOCK_K, 0);
      int buf = 0; uint32_t rphase = 0;
      const uint32_t row_bytes = (uint32_t)(s.w * s.c) * 2u;
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
        int q0, op, img; decode(tile, q0, op, img);
        mbar_wait(rempty(buf), rphase ^ 1u);
        mbar_expect_tx(rfull(buf), (uint32_t)s.r * row_bytes);
        const int iy0 = op * s.stride_h - s.pad_h;
        for (int r = 0; r < s.r; ++r) {   // ONE bulk copy per input row (28 small tensor boxes per tile made the single issuing thread the bottleneck)
          const int iy = iy0 + r * s.dil_h;
          const void* src = (iy >= 0 && iy < s.h) ? (const void*)(prm.x + ((size_t)img * s.h + iy) * (size_t)(s.w * s.c)) : (const void*)g_zero_row;
          bulk_load_1d(rows_base + (uint32_t)buf * (uint32_t)prm.rows_bytes + (uint32_t)(r * prm.rowlen + prm.lpad) * 2u, src, row_bytes, rfull(buf));
        }
        if (++buf == GA_ROW_BUFS) { buf = 0; rphase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // ===== MMA issuer
      constexpr uint32_t idesc = make_idesc(BLOCK_M, N_TILE, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      mbar_wait(bres, 0);
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty(as), aphase ^ 1u);
        mbar_wait(full(stage), phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * N_TILE);
        for (int kb = 0; kb < prm.kblocks; ++kb) {
          const uint64_t adesc = make_desc(a_base + stage * a_stage_bytes + kb * A_SLAB, 0, 1024);
          const uint64_t bdesc = make_desc(res_base + kb * B_BYTES, 0, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        umma_commit(empty(stage));
        umma_commit(tfull(as));
        if (++stage == GA_FWD_STAGES) { stage = 0; phase ^= 1u; }
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue warps 2..5 (TMEM lane quarter = warp % 4): accumulator row = pixel q0 + row of output row (img, op)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    int as = 0; uint32_t aphase = 0;
    const bool stats = prm.stats != nullptr;
    RunStats<N_TILE> run; run.reset(0);
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
      int q0, op, img; decode(tile, q0, op, img);
      const int q = q0 + row;
      const bool valid = q < s.q;
      __nv_bfloat16* dst = prm.y + (((size_t)img * s.p + op) * s.q + q) * s.k;
      mbar_wait_relaxed(tfull(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(as * N_TILE) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int ci = 0; ci < N_TILE / 32; ++ci) {
        const int c0 = 32 * ci;
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        if (stats) epilogue_store32_dyn<RunStats<N_TILE>::MODE>(prm.act, v, prm.bias ? prm.bias + c0 : nullptr, prm.slope, dst + c0, valid, run.v[ci]);
        else epilogue_store32_dyn<0>(prm.act, v, prm.bias ? prm.bias + c0 : nullptr, prm.slope, dst + c0, valid, nullptr);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (stats) run.flush(prm.stats);
  } else {
    // ===== producer warps 6..13
    const int pw = warp - (GA_THREADS - GA_PRODUCERS) / 32;
    GatherUnits units;
    gather_units_setup(prm, pw, units);
    int stage = 0; uint32_t phase = 0;
    int buf = 0; uint32_t rphase = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
      int q0, op, img; decode(tile, q0, op, img);
      mbar_wait_relaxed(rfull(buf), rphase);
      mbar_wait_relaxed(empty(stage), phase ^ 1u);
      const __nv_bfloat16* rows = reinterpret_cast<const __nv_bfloat16*>(gen_base + (rows_base - res_base) + (size_t)buf * prm.rows_bytes);
      gather_build_tile(prm, units, rows, a_base + stage * a_stage_bytes, q0, lane);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) { mbar_arrive(full(stage)); mbar_arrive(rempty(buf)); }
      if (++stage == GA_FWD_STAGES) { stage = 0; phase ^= 1u; }
      if (++buf == GA_ROW_BUFS) { buf = 0; rphase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * N_TILE) : "memory");
  }
}

// ---- weight gradient, gather variant ---------------------------------------------------------------------------------------------------------
// dw_col[k][kk] = sum over output pixels of dy[pix][k] * col[pix][kk] with col built on the fly: the tile the producers write for the forward kernel
// ([128 pixels][kpad], 128-byte rows, one 16 KB slab per 64 columns) is, read MN-major, exactly the B operand of the weight gradient (pixels = K rows,
// the kpad / 64 slabs = MN atoms one LBO apart). A = the dy tile of the same 128 pixels through TMA (a box along q; pixels beyond the row are zero filled
// and so switch off whatever the producers left in those rows). One N = kpad MMA per 16 pixels; each CTA accumulates its share of the tiles in TMEM and
// adds it into dw_col with vector reductions at the end.
struct GatherWgradParams {
  GatherParams g;
  float* dw_col;
};
constexpr int GW_NA = 2;   // dy tile slots (2 x 16 KB atoms each)

__global__ void __launch_bounds__(GA_THREADS, 1) conv_wgrad_tc_gather_kernel(const __grid_constant__ CUtensorMap map_dy, const GatherWgradParams wp) {
  const GatherParams& prm = wp.g;
  constexpr int A_SLAB = BLOCK_M * 128, DY_BYTES = 2 * A_SLAB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t dy_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = dy_base + GW_NA * DY_BYTES;
  const uint32_t b_stage_bytes = (uint32_t)prm.kblocks * A_SLAB;
  const uint32_t rows_base = b_base + GA_STAGES * b_stage_bytes;
  const uint32_t bars = rows_base + (uint32_t)GA_ROW_BUFS * (uint32_t)prm.rows_bytes;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (GA_STAGES + i); };
  auto afull = [&](int i) { return bars + 8u * (2 * GA_STAGES + i); };
  auto aempty = [&](int i) { return bars + 8u * (2 * GA_STAGES + GW_NA + i); };
  auto rfull = [&](int i) { return bars + 8u * (2 * GA_STAGES + 2 * GW_NA + i); };
  auto rempty = [&](int i) { return bars + 8u * (2 * GA_STAGES + 2 * GW_NA + GA_ROW_BUFS + i); };
  const uint32_t done = bars + 8u * (2 * GA_STAGES + 2 * GW_NA + 2 * GA_ROW_BUFS), tmem_slot = done + 8u;
  uint8_t* gen_base = smem_raw + (dy_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const dcv_conv_shape& s = prm.s;
  const bool second_atom = s.k > 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < GA_STAGES; ++i) { mbar_init(full(i), GA_PRODUCERS / 32); mbar_init(empty(i), 1); }
    for (int i = 0; i < GW_NA; ++i) { mbar_init(afull(i), 1); mbar_init(aempty(i), 1); }
    for (int i = 0; i < GA_ROW_BUFS; ++i) { mbar_init(rfull(i), 1); mbar_init(rempty(i), GA_PRODUCERS / 32); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_dy)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // zero the staging rows (margins = padding) and, for 64 output channels, the unused second atom of every dy slot
  for (uint32_t i = threadIdx.x; i < (uint32_t)GA_ROW_BUFS * (uint32_t)prm.rows_bytes / 16u; i += GA_THREADS)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(rows_base + i * 16u), "r"(0) : "memory");
  if (!second_atom)
    for (uint32_t i = threadIdx.x; i < (uint32_t)GW_NA * (A_SLAB / 16); i += GA_THREADS)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dy_base + (i / (A_SLAB / 16)) * DY_BYTES + A_SLAB + (i % (A_SLAB / 16)) * 16u), "r"(0) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  auto decode = [&](int tile, int& q0, int& op, int& img) {
    q0 = (tile % prm.tiles_q) * BLOCK_M; const int t = tile / prm.tiles_q;
    op = t % s.p; img = t / s.p;
  };
  const int my_tiles = (int)blockIdx.x < prm.total_tiles ? (prm.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    if (elect_one()) {   // ===== TMA: per tile the dy slab(s) and the R input rows
      int buf = 0; uint32_t rphase = 0;
      int as = 0; uint32_t aph = 0;
      const uint32_t row_bytes = (uint32_t)(s.w * s.c) * 2u;
      for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
        int q0, op, img; decode(tile, q0, op, img);
        mbar_wait(aempty(as), aph ^ 1u);
        mbar_expect_tx(afull(as), second_atom ? DY_BYTES : A_SLAB);
        tma_load_4d(dy_base + as * DY_BYTES, &map_dy, afull(as), 0, q0, op, img);
        if (second_atom) tma_load_4d(dy_base + as * DY_BYTES + A_SLAB, &map_dy, afull(as), 64, q0, op, img);
        if (++as == GW_NA) { as = 0; aph ^= 1u; }
        mbar_wait(rempty(buf), rphase ^ 1u);
        mbar_expect_tx(rfull(buf), (uint32_t)s.r * row_bytes);
        const int iy0 = op * s.stride_h - s.pad_h;
        for (int r = 0; r < s.r; ++r) {
          const int iy = iy0 + r * s.dil_h;
          const void* src = (iy >= 0 && iy < s.h) ? (const void*)(prm.x + ((size_t)img * s.h + iy) * (size_t)(s.w * s.c)) : (const void*)g_zero_row;
          bulk_load_1d(rows_base + (uint32_t)buf * (uint32_t)prm.rows_bytes + (uint32_t)(r * prm.rowlen + prm.lpad) * 2u, src, row_bytes, rfull(buf));
        }
        if (++buf == GA_ROW_BUFS) { buf = 0; rphase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // ===== MMA issuer: D[128 k][kpad] += dy^T * col over this CTA's tiles
      const uint32_t idesc = make_idesc(128, prm.kpad, 1, 1);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(afull(as), aph);
        mbar_wait(full(stage), phase);
        tc_fence_after();
        const uint32_t sa = dy_base + as * DY_BYTES, sb = b_base + stage * b_stage_bytes;
#pragma unroll
        for (int ks = 0; ks < BLOCK_M / UMMA_K; ++ks) {
          const uint64_t adesc = make_desc(sa + ks * UMMA_K * 128, A_SLAB, 1024);
          const uint64_t bdesc = make_desc(sb + ks * UMMA_K * 128, A_SLAB, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (i | ks) != 0);
        }
        umma_commit(empty(stage));
        umma_commit(aempty(as));
        if (++stage == GA_STAGES) { stage = 0; phase ^= 1u; }
        if (++as == GW_NA) { as = 0; aph ^= 1u; }
      }
      umma_commit(done);
    }
  } else if (warp < 6) {
    if (my_tiles > 0) {   // ===== epilogue, once: TMEM lane = output channel, column = im2col column
      const int quarter = warp & 3;
      const int krow = quarter * 32 + lane;
      mbar_wait_relaxed(done, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < prm.kpad; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        if (krow < s.k) {
          float* dst = wp.dw_col + (size_t)krow * prm.kpad + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                         "f"(__uint_as_float(v[j + 3])) : "memory");
        }
      }
    }
  } else {
    // ===== producer warps: the im2col tile of every tile, as in the forward kernel
    const int pw = warp - (GA_THREADS - GA_PRODUCERS) / 32;
    GatherUnits units;
    gather_units_setup(prm, pw, units);
    int stage = 0; uint32_t phase = 0;
    int buf = 0; uint32_t rphase = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
      int q0, op, img; decode(tile, q0, op, img);
      mbar_wait_relaxed(rfull(buf), rphase);
      mbar_wait_relaxed(empty(stage), phase ^ 1u);
      const __nv_bfloat16* rows = reinterpret_cast<const __nv_bfloat16*>(gen_base + (rows_base - dy_base) + (size_t)buf * prm.rows_bytes);
      gather_build_tile(prm, units, rows, b_base + stage * b_stage_bytes, q0, lane);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(full(stage)); mbar_arrive(rempty(buf)); }
      if (++stage == GA_STAGES) { stage = 0; phase ^= 1u; }
      if (++buf == GA_ROW_BUFS) { buf = 0; rphase ^= 1u; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// ---- stride-2 few-channel convolutions ("pixel pair" variant: the 3 -> 64, 7x7 / stride-2 stem) ----------------------------------------------------
// The gather kernels above spend their time in the producers (ncu, stem at batch 256: 415 us, tensor pipe 10 %, 23 warp instructions per 16-byte chunk of the
// im2col tile, 96 chunks per pixel band; every role waits on them). With stride 2 the im2col matrix has a structure the tensor core can address by itself:
// stage the R input rows of an output row TRANSPOSED, one 128-byte line per PIXEL PAIR:   T[L][r][8 elements] = row r, pixels 2P and 2P + 1 (2C <= 8
// elements; the rest of the 16-byte chunk is whatever follows in the row and meets zero weights), P = q0 - hp + L.   Output pixel q reads the pixel pairs
// q - hp .. q - hp + 3 (8 pixels >= S + (pad & 1) taps), i.e. FOUR CONSECUTIVE LINES starting at line q - q0:
//   * forward: the A operand of K block pp (64 elements = line m + pp of row m) is the K-major SWIZZLE_128B tile that starts at line pp — overlapping
//     tiles, the swizzle is a function of the shared-memory address only, so every view reads the same bytes. K = 4 x 64 = 256; 16 MMAs per 128 pixels.
//   * weight gradient: read MN-major, the same lines are the B operand with N = 256 (four 64-element atoms ONE LINE apart: LBO = 128 bytes).
// The producers only move 16-byte chunks (4 shared loads + 1 store per chunk, R x 131 chunks per tile): 14x fewer instructions than the gather, no K-sized
// tile in shared memory (17 KB per stage instead of 48 KB). K order of the weights: column pp * 64 + r * 8 + px * C + c = w[k][r][2 pp + px - (pad & 1)][c].
struct PairParams {
  dcv_conv_shape s;
  int rowlen, lpad, rows_bytes;   // staged input rows, in elements: [lpad zeros][W * C][zeros up to rowlen]
  int hp;                         // output pixel q reads the pixel pairs q - hp .. q - hp + 3
  int tiles_q, total_tiles;
  int act; float slope;
  const float* bias;
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
};
constexpr int PR_LINES = BLOCK_M + 3, PR_STAGE_BYTES = 18 * 1024, PR_K = 256, PR_KBLOCKS = PR_K / BLOCK_K;
// Four producer GROUPS of two warps: group g builds the tiles i = g (mod 4) of the CTA in line stage g from row buffer g — four tiles are under construction
// at once. (All eight warps on one tile ran at the latency of wait -> 4 units -> fence -> arrive per tile: 0.96 us per tile with the stores switched off.)
constexpr int PR_GROUPS = 4, PR_FWD_STAGES = PR_GROUPS, PR_WG_STAGES = PR_GROUPS, PR_ROW_BUFS = 2 * PR_GROUPS, PR_PRODUCERS = PR_GROUPS * 64;   // two row buffers per group: the rows of its next tile land while it builds the current one
constexpr int PR_FWD_THREADS = 64 + 256 + PR_PRODUCERS;   // warps: 0 TMA, 1 MMA, 2..9 epilogue, 10..17 producers
constexpr int PR_WG_THREADS = 64 + 128 + PR_PRODUCERS;                                          // warps: 0 TMA, 1 MMA, 2..5 epilogue, 6..13 producers

// The tiles of a CTA are a CONTIGUOUS range of (image, output row, row segment) triples walked by increments: the first versions decoded `tile` with two
// runtime divisions per tile in every one of the 18 warps (ncu: 15 % of the kernel's instructions, and on the epilogue's critical path).
struct PairTiles {
  int left, tq, op, img;
  __device__ __forceinline__ PairTiles(const PairParams& prm) {
    const long long t = prm.total_tiles, g = gridDim.x, b = blockIdx.x;
    const int first = (int)(b * t / g);
    left = (int)((b + 1) * t / g) - first;
    tq = first % prm.tiles_q; const int u = first / prm.tiles_q;
    op = u % prm.s.p; img = u / prm.s.p;
  }
  __device__ __forceinline__ bool valid() const { return left > 0; }
  __device__ __forceinline__ void next(const PairParams& prm) {
    --left;
    if (++tq == prm.tiles_q) { tq = 0; if (++op == prm.s.p) { op = 0; ++img; } }
  }
};

// epilogue_store32 with the 32 bias values held in registers (loaded once per kernel: an epilogue warp of the pixel-pair kernel always serves the same columns)
template <int ACT>
__device__ __forceinline__ void pairs_store32(const uint32_t* v, const float* bias_regs, float slope, __nv_bfloat16* dst) {
  float f[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    f[j] = __uint_as_float(v[j]) + bias_regs[j];
    if (ACT == DCV_ACT_RELU) f[j] = fmaxf(f[j], 0.f);
    else if (ACT == DCV_ACT_LEAKY_RELU) f[j] = f[j] > 0.f ? f[j] : f[j] * slope;
    else if (ACT == DCV_ACT_SIGMOID) f[j] = 1.f / (1.f + expf(-f[j]));
  }
  uint32_t pk[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) { const uint4 q = vec_pack<__nv_bfloat16>(f + 8 * j); pk[4 * j] = q.x; pk[4 * j + 1] = q.y; pk[4 * j + 2] = q.z; pk[4 * j + 3] = q.w; }
  if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
#pragma unroll
    for (int w = 0; w < 16; w += 8)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(reinterpret_cast<uint32_t*>(dst) + w), "r"(pk[w]), "r"(pk[w + 1]), "r"(pk[w + 2]), "r"(pk[w + 3]),
                   "r"(pk[w + 4]), "r"(pk[w + 5]), "r"(pk[w + 6]), "r"(pk[w + 7]) : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
  }
}
__device__ __forceinline__ void pairs_store32_dyn(int act, const uint32_t* v, const float* bias_regs, float slope, __nv_bfloat16* dst) {
  switch (act) {   // warp-uniform
    case DCV_ACT_RELU: pairs_store32<DCV_ACT_RELU>(v, bias_regs, slope, dst); break;
    case DCV_ACT_LEAKY_RELU: pairs_store32<DCV_ACT_LEAKY_RELU>(v, bias_regs, slope, dst); break;
    case DCV_ACT_SIGMOID: pairs_store32<DCV_ACT_SIGMOID>(v, bias_regs, slope, dst); break;
    default: pairs_store32<DCV_ACT_NONE>(v, bias_regs, slope, dst); break;
  }
}

// One of the two warps of a producer group builds its share of a tile: filter rows sub, sub + 2, ...; per row the 16-byte chunks of all the lines, the
// shared loads of a whole row issued before its stores (the stores are volatile asm: the compiler does not move loads across them).
__device__ __forceinline__ void pairs_build_tile(const PairParams& prm, const uint32_t* row_words, uint32_t t_stage, int q0, int sub, int lane) {
  const dcv_conv_shape& s = prm.s;
  constexpr int NLB = (PR_LINES + 31) / 32;
  const int e_lane = prm.lpad + 2 * (q0 - prm.hp + lane) * s.c, e_step = 64 * s.c;   // first element of the pixel pair of line `lane` (even); 32 lines further
  for (int r = sub; r < s.r; r += 2) {
    uint32_t w[NLB][4];
#pragma unroll
    for (int b = 0; b < NLB; ++b) {
      const int e0 = e_lane + b * e_step;
      w[b][0] = w[b][1] = w[b][2] = w[b][3] = 0u;
      if (b * 32 + lane < PR_LINES && e0 + 8 <= prm.rowlen) {   // beyond the staged row: only lines of overhang pixels (never stored / zero dy) — keep them finite
        const uint32_t* wp = row_words + ((r * prm.rowlen + e0) >> 1);
        w[b][0] = wp[0]; w[b][1] = wp[1]; w[b][2] = wp[2]; w[b][3] = wp[3];
      }
    }
#pragma unroll
    for (int b = 0; b < NLB; ++b)
      if (b * 32 + lane < PR_LINES)   // 32 b + lane and lane agree modulo 8
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_stage + (uint32_t)(b * 32 + lane) * 128u + ((uint32_t)(r ^ (lane & 7)) << 4)), "r"(w[b][0]), "r"(w[b][1]), "r"(w[b][2]),
                     "r"(w[b][3]) : "memory");
  }
}

template <int N_TILE>
__global__ void __launch_bounds__(PR_FWD_THREADS, 1) conv_fwd_tc_pairs_kernel(const __grid_constant__ CUtensorMap map_w, const PairParams prm) {
  constexpr int B_BYTES = N_TILE * BLOCK_K * 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t res_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t t_base = res_base + PR_KBLOCKS * B_BYTES;
  const uint32_t rows_base = t_base + PR_FWD_STAGES * PR_STAGE_BYTES;
  const uint32_t bars = rows_base + (uint32_t)PR_ROW_BUFS * (uint32_t)prm.rows_bytes;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (PR_FWD_STAGES + i); };
  constexpr int NACC = 512 / N_TILE >= 4 ? 4 : 2;   // TMEM accumulator buffers: four tiles of output may be waiting for their (HBM-bound) stores
  auto tfull = [&](int i) { return bars + 8u * (2 * PR_FWD_STAGES + i); };
  auto tempty = [&](int i) { return bars + 8u * (2 * PR_FWD_STAGES + NACC + i); };
  const uint32_t bres = bars + 8u * (2 * PR_FWD_STAGES + 2 * NACC), tmem_slot = bars + 8u * (2 * PR_FWD_STAGES + 2 * NACC + 1);
  auto rfull = [&](int i) { return bars + 8u * (2 * PR_FWD_STAGES + 2 * NACC + 2 + i); };
  auto rempty = [&](int i) { return bars + 8u * (2 * PR_FWD_STAGES + 2 * NACC + 2 + PR_ROW_BUFS + i); };
  uint8_t* gen_base = smem_raw + (res_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const dcv_conv_shape& s = prm.s;

  if (threadIdx.x == 0) {
    for (int i = 0; i < PR_FWD_STAGES; ++i) { mbar_init(full(i), 2); mbar_init(empty(i), 1); }
    for (int i = 0; i < NACC; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 8); }
    mbar_init(bres, 1);
    for (int i = 0; i < PR_ROW_BUFS; ++i) { mbar_init(rfull(i), 1); mbar_init(rempty(i), 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(NACC * N_TILE) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the line stages (chunks of filter rows >= R stay zero) and the row buffers (their margins are the horizontal padding) start as zeros
  for (uint32_t i = threadIdx.x; i < (PR_FWD_STAGES * PR_STAGE_BYTES + (uint32_t)PR_ROW_BUFS * (uint32_t)prm.rows_bytes) / 16u; i += PR_FWD_THREADS)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(t_base + i * 16u), "r"(0) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (elect_one()) {   // ===== weights once, then the R input rows of every tile (one bulk copy each; rows outside the image come from a zero row)
      mbar_expect_tx(bres, PR_KBLOCKS * B_BYTES);
      for (int kb = 0; kb < PR_KBLOCKS; ++kb) tma_load_2d(res_base + kb * B_BYTES, &map_w, bres, kb * BLOCK_K, 0);
      int buf = 0; uint32_t rphase = 0;
      const uint32_t row_bytes = (uint32_t)(s.w * s.c) * 2u;
      for (PairTiles t(prm); t.valid(); t.next(prm)) {
        const int op = t.op, img = t.img;
        mbar_wait(rempty(buf), rphase ^ 1u);
        mbar_expect_tx(rfull(buf), (uint32_t)s.r * row_bytes);
        const int iy0 = op * s.stride_h - s.pad_h;
        for (int r = 0; r < s.r; ++r) {
          const int iy = iy0 + r * s.dil_h;
          const void* src = (iy >= 0 && iy < s.h) ? (const void*)(prm.x + ((size_t)img * s.h + iy) * (size_t)(s.w * s.c)) : (const void*)g_zero_row;
          bulk_load_1d(rows_base + (uint32_t)buf * (uint32_t)prm.rows_bytes + (uint32_t)(r * prm.rowlen + prm.lpad) * 2u, src, row_bytes, rfull(buf));
        }
        if (++buf == PR_ROW_BUFS) { buf = 0; rphase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // ===== MMA issuer: K block pp of pixel row m is line m + pp
      constexpr uint32_t idesc = make_idesc(BLOCK_M, N_TILE, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aphase = 0;
      mbar_wait(bres, 0);
      for (PairTiles t(prm); t.valid(); t.next(prm)) {
        mbar_wait(tempty(as), aphase ^ 1u);
        mbar_wait(full(stage), phase);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * N_TILE);
#pragma unroll
        for (int kb = 0; kb < PR_KBLOCKS; ++kb) {
          const uint64_t adesc = make_desc(t_base + stage * PR_STAGE_BYTES + kb * 128, 0, 1024);
          const uint64_t bdesc = make_desc(res_base + kb * B_BYTES, 0, 1024);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        umma_commit(empty(stage));
        umma_commit(tfull(as));
        if (++stage == PR_FWD_STAGES) { stage = 0; phase ^= 1u; }
        if (++as == NACC) { as = 0; aphase ^= 1u; }
      }
    }
  } else if (warp < 10) {
    // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    int as = 0; uint32_t aphase = 0;
    float bias_regs[N_TILE / 64][32];
#pragma unroll
    for (int ci = 0; ci < N_TILE / 64; ++ci)
#pragma unroll
      for (int j = 0; j < 32; ++j) bias_regs[ci][j] = prm.bias ? __ldg(prm.bias + half * (N_TILE / 2) + 32 * ci + j) : 0.f;
    for (PairTiles t(prm); t.valid(); t.next(prm)) {
      const int q = t.tq * BLOCK_M + row;
      const bool valid = q < s.q;
      __nv_bfloat16* dst = prm.y + (((size_t)t.img * s.p + t.op) * s.q + q) * s.k;
      mbar_wait_relaxed(tfull(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(as * N_TILE) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
      for (int ci = 0; ci < N_TILE / 64; ++ci) {
        const int c0 = half * (N_TILE / 2) + 32 * ci;
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        if (valid) pairs_store32_dyn(prm.act, v, bias_regs[ci], prm.slope, dst + c0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(as));
      if (++as == NACC) { as = 0; aphase ^= 1u; }
    }
  } else {
    // ===== producer warps 10..17: transpose the staged rows into pixel-pair lines
    const int g = (warp - 10) >> 1, sub = warp & 1;
    int i = 0, own = 0;
    for (PairTiles t(prm); t.valid(); t.next(prm), ++i) {
      if ((i & (PR_GROUPS - 1)) != g) continue;
      const int buf = g + PR_GROUPS * (own & 1);   // = i mod PR_ROW_BUFS, the loader's order
      const uint32_t phase = (uint32_t)(own & 1), rphase = (uint32_t)((own >> 1) & 1);
      const uint32_t* rows = reinterpret_cast<const uint32_t*>(gen_base + (rows_base - res_base) + (size_t)buf * prm.rows_bytes);
      mbar_wait_relaxed(rfull(buf), rphase);
      mbar_wait_relaxed(empty(g), phase ^ 1u);
      pairs_build_tile(prm, rows, t_base + g * PR_STAGE_BYTES, t.tq * BLOCK_M, sub, lane);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(full(g)); mbar_arrive(rempty(buf)); }
      ++own;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NACC * N_TILE) : "memory");
  }
}

// Weight gradient on the same lines, transposed: D^T[256 weight columns][K output channels] += lines (A, MN-major: M atoms of 64 columns ONE LINE apart,
// two M = 128 halves) x dy (B, MN-major through TMA: N = K channels, pixels beyond the row zero filled) over the CTA's tiles — N = K, not 256: a quarter of
// the tensor time of the D[128 channels][256] form (whose M was half zero padding for K = 64), and a 16 KB dy slot per tile (K = 64), so six tiles of dy
// are in flight: with two 32 KB slots the kernel ran at DRAM latency (ncu: 255 us, tensor pipe 44 %, every role waiting on the dy box).
struct PairWgradParams {
  PairParams g;
  float* dw_col;
  int dy_slots;
};

__global__ void __launch_bounds__(PR_WG_THREADS, 1) conv_wgrad_tc_pairs_kernel(const __grid_constant__ CUtensorMap map_dy, const PairWgradParams wp) {
  const PairParams& prm = wp.g;
  constexpr int A_SLAB = BLOCK_M * 128, MAX_SLOTS = 6;
  extern __shared__ uint8_t smem_raw[];
  const dcv_conv_shape& s = prm.s;
  const int atoms = s.k / 64, slots = wp.dy_slots;
  const uint32_t dy_bytes = (uint32_t)atoms * A_SLAB;
  const uint32_t dy_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t t_base = dy_base + (uint32_t)slots * dy_bytes;
  const uint32_t rows_base = t_base + PR_WG_STAGES * PR_STAGE_BYTES;
  const uint32_t bars = rows_base + (uint32_t)PR_ROW_BUFS * (uint32_t)prm.rows_bytes;
  auto full = [&](int i) { return bars + 8u * i; };
  auto empty = [&](int i) { return bars + 8u * (PR_WG_STAGES + i); };
  auto afull = [&](int i) { return bars + 8u * (2 * PR_WG_STAGES + i); };
  auto aempty = [&](int i) { return bars + 8u * (2 * PR_WG_STAGES + MAX_SLOTS + i); };
  auto rfull = [&](int i) { return bars + 8u * (2 * PR_WG_STAGES + 2 * MAX_SLOTS + i); };
  auto rempty = [&](int i) { return bars + 8u * (2 * PR_WG_STAGES + 2 * MAX_SLOTS + PR_ROW_BUFS + i); };
  const uint32_t done = bars + 8u * (2 * PR_WG_STAGES + 2 * MAX_SLOTS + 2 * PR_ROW_BUFS), tmem_slot = done + 8u;
  uint8_t* gen_base = smem_raw + (dy_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < PR_WG_STAGES; ++i) { mbar_init(full(i), 2); mbar_init(empty(i), 1); }
    for (int i = 0; i < MAX_SLOTS; ++i) { mbar_init(afull(i), 1); mbar_init(aempty(i), 1); }
    for (int i = 0; i < PR_ROW_BUFS; ++i) { mbar_init(rfull(i), 1); mbar_init(rempty(i), 2); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_dy)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x; i < (PR_WG_STAGES * PR_STAGE_BYTES + (uint32_t)PR_ROW_BUFS * (uint32_t)prm.rows_bytes) / 16u; i += PR_WG_THREADS)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(t_base + i * 16u), "r"(0) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const bool any_tiles = PairTiles(prm).valid();

  if (warp == 0) {
    if (elect_one()) {   // ===== TMA: per tile the dy box(es) (pixels beyond the row are zero filled) and the R input rows
      int buf = 0; uint32_t rphase = 0;
      int as = 0; uint32_t aph = 0;
      const uint32_t row_bytes = (uint32_t)(s.w * s.c) * 2u;
      for (PairTiles t(prm); t.valid(); t.next(prm)) {
        const int q0 = t.tq * BLOCK_M;
        mbar_wait(aempty(as), aph ^ 1u);
        mbar_expect_tx(afull(as), dy_bytes);
        for (int a = 0; a < atoms; ++a) tma_load_4d(dy_base + as * dy_bytes + a * A_SLAB, &map_dy, afull(as), 64 * a, q0, t.op, t.img);
        if (++as == slots) { as = 0; aph ^= 1u; }
        mbar_wait(rempty(buf), rphase ^ 1u);
        mbar_expect_tx(rfull(buf), (uint32_t)s.r * row_bytes);
        const int iy0 = t.op * s.stride_h - s.pad_h;
        for (int r = 0; r < s.r; ++r) {
          const int iy = iy0 + r * s.dil_h;
          const void* src = (iy >= 0 && iy < s.h) ? (const void*)(prm.x + ((size_t)t.img * s.h + iy) * (size_t)(s.w * s.c)) : (const void*)g_zero_row;
          bulk_load_1d(rows_base + (uint32_t)buf * (uint32_t)prm.rows_bytes + (uint32_t)(r * prm.rowlen + prm.lpad) * 2u, src, row_bytes, rfull(buf));
        }
        if (++buf == PR_ROW_BUFS) { buf = 0; rphase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // ===== MMA issuer: per 16 pixels one M = 128 x N = K MMA for each half of the 256 weight columns
      const uint32_t idesc = make_idesc(128, s.k, 1, 1);
      int stage = 0; uint32_t phase = 0;
      int as = 0; uint32_t aph = 0;
      bool first = true;
      for (PairTiles t(prm); t.valid(); t.next(prm)) {
        mbar_wait(afull(as), aph);
        mbar_wait(full(stage), phase);
        tc_fence_after();
        const uint32_t sdy = dy_base + as * dy_bytes, sl = t_base + stage * PR_STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < BLOCK_M / UMMA_K; ++ks) {
          const uint64_t bdesc = make_desc(sdy + ks * UMMA_K * 128, A_SLAB, 1024);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t adesc = make_desc(sl + h * 256 + ks * UMMA_K * 128, 128, 1024);
            umma_bf16(tmem_base + (uint32_t)(h * s.k), adesc, bdesc, idesc, !(first && ks == 0));
          }
        }
        first = false;
        umma_commit(empty(stage));
        umma_commit(aempty(as));
        if (++stage == PR_WG_STAGES) { stage = 0; phase ^= 1u; }
        if (++as == slots) { as = 0; aph ^= 1u; }
      }
      umma_commit(done);
    }
  } else if (warp < 6) {
    if (any_tiles) {   // ===== epilogue, once: TMEM lane = weight column within the half, TMEM column = output channel
      const int quarter = warp & 3;
      mbar_wait_relaxed(done, 0);
      tc_fence_after();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int col = h * 128 + quarter * 32 + lane;
        const uint32_t taddr = tmem_base + (uint32_t)(h * s.k) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < s.k; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          float* dst = wp.dw_col + (size_t)c0 * PR_K + col;   // the 32 lanes of a warp: 32 consecutive floats
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + (size_t)j * PR_K, __uint_as_float(v[j]));
        }
      }
    }
  } else {
    const int g = (warp - 6) >> 1, sub = warp & 1;   // ===== producer warps 6..13, as in the forward kernel
    int i = 0, own = 0;
    for (PairTiles t(prm); t.valid(); t.next(prm), ++i) {
      if ((i & (PR_GROUPS - 1)) != g) continue;
      const int buf = g + PR_GROUPS * (own & 1);   // = i mod PR_ROW_BUFS, the loader's order
      const uint32_t phase = (uint32_t)(own & 1), rphase = (uint32_t)((own >> 1) & 1);
      const uint32_t* rows = reinterpret_cast<const uint32_t*>(gen_base + (rows_base - dy_base) + (size_t)buf * prm.rows_bytes);
      mbar_wait_relaxed(rfull(buf), rphase);
      mbar_wait_relaxed(empty(g), phase ^ 1u);
      pairs_build_tile(prm, rows, t_base + g * PR_STAGE_BYTES, t.tq * BLOCK_M, sub, lane);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(full(g)); mbar_arrive(rempty(buf)); }
      ++own;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 tensor map, SWIZZLE_128B, zero fill outside the tensor. dims / box innermost first; strides in bytes for dims 1..rank-1.
static int make_map(CUtensorMap* map, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn fn = encode_tiled();
  DCV_REQUIRE(fn, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DCV_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// (tw, th, tn), powers of two with product 128, minimising the padded overhang of (q, p, n).
static void pick_pixel_tile(int q, int p, int n, int* tw, int* th, int* tn) {
  double best = 1e30;
  for (int a = 1; a <= 128; a *= 2)
    for (int b = 1; a * b <= 128; b *= 2) {
      const int c = 128 / (a * b);
      if (a > 256 || b > 256 || c > 256) continue;
      const double padded = (double)((q + a - 1) / a * a) * ((p + b - 1) / b * b) * ((n + c - 1) / c * c);
      // prefer wider rows (longer contiguous global segments per TMA box) on ties
      const double cost = padded / ((double)q * p * n) - 1e-6 * a;
      if (cost < best) { best = cost; *tw = a; *th = b; *tn = c; }
    }
}

static int num_sms() {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = kNumSMs; }
  return sms;
}

template <int N_TILE, int RES = 0>
static int launch_fwd(const CUtensorMap& mx, const CUtensorMap& mw, const FwdParams& prm, cudaStream_t st) {
  auto kern = conv_fwd_tc_kernel<N_TILE, RES>;
  const size_t smem = FwdSmem<N_TILE, RES>::kBytes;
  static bool configured = false;
  if (!configured) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = true; }
  const int grid = prm.total_tiles < num_sms() ? prm.total_tiles : num_sms();
  kern<<<grid, kThreads, smem, st>>>(mx, mw, prm);
  DCV_LAUNCH_CHECK("conv_fwd_tc_kernel");
  return 0;
}


// Halo variant: filters wider than one tap, a single output-channel tile of 64 or 128 whose weight slabs all stay resident (<= 144 KB), and a map that
// 8 x th tiles cover with <= 15 % overhang (56 x 56: 7 x 4 tiles of 8 x 14, 12.5 %; the 28 / 14 / 7 pixel maps lose more than they gain).
static bool fwd_halo_applicable(const dcv_conv_shape* s) {
  static const bool disabled = getenv("DCV_TC_NO_HALO") != nullptr;
  if (disabled || (s->k != 64 && s->k != 128) || s->r * s->s < 2 || s->r > 3 || s->s > 3) return false;
  const int num_kb = s->r * s->s * (s->c / BLOCK_K);
  if ((size_t)num_kb * s->k * BLOCK_K * 2 > 144 * 1024) return false;
  const int tiles_h = (s->p + 15) / 16, th = (s->p + tiles_h - 1) / tiles_h, tiles_w = (s->q + HALO_TW - 1) / HALO_TW;
  const double cover = (double)s->p * s->q / ((double)tiles_h * 16 * tiles_w * HALO_TW);
  (void)th;
  return cover >= 0.85;
}

template <int N_TILE>
static int launch_fwd_halo(const CUtensorMap& mx, const CUtensorMap& mw, const FwdHaloParams& prm, size_t smem, cudaStream_t st) {
  auto kern = conv_fwd_tc_halo_kernel<N_TILE>;
  static size_t configured = 0;
  if (configured < smem) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
  const int grid = prm.total_tiles < num_sms() ? prm.total_tiles : num_sms();
  kern<<<grid, kThreads, smem, st>>>(mx, mw, prm);
  DCV_LAUNCH_CHECK("conv_fwd_tc_halo_kernel");
  return 0;
}

static int conv_fwd_tc_halo(const dcv_conv_shape* s, const void* x, const void* w, const float* bias, void* y, int act, float slope, float* stats, cudaStream_t st) {
  FwdHaloParams prm{};
  prm.n = s->n; prm.c = s->c; prm.k = s->k; prm.r = s->r; prm.s = s->s; prm.pad_h = s->pad_h; prm.pad_w = s->pad_w; prm.p = s->p; prm.q = s->q;
  prm.tiles_h = (s->p + 15) / 16;
  prm.th = (s->p + prm.tiles_h - 1) / prm.tiles_h;
  prm.tiles_w = (s->q + HALO_TW - 1) / HALO_TW;
  const long long tiles = (long long)s->n * prm.tiles_h * prm.tiles_w;
  DCV_REQUIRE(tiles < (1ll << 31), "conv2d_fwd (tcgen05): too many tiles");
  prm.total_tiles = (int)tiles;
  prm.act = act; prm.slope = slope; prm.bias = bias; prm.y = reinterpret_cast<__nv_bfloat16*>(y); prm.stats = stats;
  const size_t res_bytes = (size_t)s->r * s->s * (s->c / BLOCK_K) * s->k * BLOCK_K * 2;
  int stages = (int)((227 * 1024 - 1024 - 256 - res_bytes) / HALO_STAGE_BYTES);
  if (stages > 8) stages = 8;
  DCV_REQUIRE(stages >= 2, "conv2d_fwd (tcgen05 halo): %zu bytes of weights leave no room for the activation ring", res_bytes);
  prm.stages = stages;
  const size_t smem = 1024 + res_bytes + (size_t)stages * HALO_STAGE_BYTES + 256;
  CUtensorMap mx, mw;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)s->c, (cuuint64_t)s->w, (cuuint64_t)s->h, (cuuint64_t)s->n};
    const cuuint64_t strides[3] = {(cuuint64_t)s->c * 2, (cuuint64_t)s->w * s->c * 2, (cuuint64_t)s->h * s->w * s->c * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(HALO_TW + s->s - 1), (cuuint32_t)(prm.th + s->r - 1), 1u};
    if (make_map(&mx, x, 4, dims, strides, box)) return 1;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)s->r * s->s * s->c, (cuuint64_t)s->k};
    const cuuint64_t strides[1] = {(cuuint64_t)s->r * s->s * s->c * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)s->k};
    if (make_map(&mw, w, 2, dims, strides, box)) return 1;
  }
  return s->k == 128 ? launch_fwd_halo<128>(mx, mw, prm, smem, st) : launch_fwd_halo<64>(mx, mw, prm, smem, st);
}

// Geometry of the gather kernels' staging area; false when the layer does not fit (falls back to the explicit im2col route).
static bool gather_geometry(const dcv_conv_shape* s, const void* x, int kpad, GatherParams* prm, size_t* smem, int b_bytes) {
  const int wc = s->w * s->c, rp = (s->s * s->c + 7) / 8 * 8;
  if (wc * 2 > kZeroRowBytes || s->dil_w != 1 || kpad % BLOCK_K != 0 || kpad < s->r * rp || kpad > 256 || wc % 8 != 0 || reinterpret_cast<uintptr_t>(x) % 16 != 0) return false;
  prm->rp = rp;
  int lpad = s->pad_w * s->c;
  lpad = (lpad + 63) / 64 * 64;   // every staged row (and every box inside it) starts on a 128-byte boundary
  // the last tile's overhang pixels are not gathered, so the farthest element read belongs to pixel q - 1
  int need = lpad + ((s->q - 1) * s->stride_w - s->pad_w) * s->c + rp + 2;   // a chunk reads 5 aligned words = up to 10 elements from its first one
  if (need < lpad + wc) need = lpad + wc;
  const int rowlen = (need + 63) / 64 * 64;
  prm->s = *s; prm->kpad = kpad; prm->kblocks = kpad / BLOCK_K; prm->rowlen = rowlen; prm->lpad = lpad;
  prm->rows_bytes = s->r * rowlen * 2;   // a multiple of 128
  prm->tiles_q = (s->q + BLOCK_M - 1) / BLOCK_M;
  const long long tiles = (long long)s->n * s->p * prm->tiles_q;
  if (tiles >= (1ll << 31)) return false;
  prm->total_tiles = (int)tiles;
  if (GA_PRODUCERS % (kpad / 8) != 0 && GA_PRODUCERS / (kpad / 8) < 1) return false;
  *smem = 1024 + (size_t)prm->kblocks * b_bytes + (size_t)GA_FWD_STAGES * prm->kblocks * BLOCK_M * 128 + (size_t)GA_ROW_BUFS * prm->rows_bytes + 256;
  return *smem <= 227 * 1024;
}

}  // namespace tc

bool conv_fwd_tc_gather_supported(const dcv_conv_shape* s, const void* x, int kpad, int dtype) {
  if (!s || dtype != DCV_BF16 || (s->k != 64 && s->k != 128) || tc::encode_tiled() == nullptr) return false;
  if ((long long)s->n * s->p * s->q < 128) return false;
  tc::GatherParams prm{}; size_t smem = 0;
  return tc::gather_geometry(s, x, kpad, &prm, &smem, s->k * tc::BLOCK_K * 2);
}

// w_col: [K][kpad] bf16, columns (r, s, c) of the [K][R][S][C] weights followed by zeros.
int conv_fwd_tc_gather(const dcv_conv_shape* s, const void* x, const void* w_col, int kpad, const float* bias, void* y, float* stats_nc, int act, float slope, int stats_flags, cudaStream_t st) {
  using namespace tc;
  DCV_REQUIRE(x && w_col && y, "conv2d_fwd_gather: null pointer");
  GatherParams prm{}; size_t smem = 0;
  DCV_REQUIRE(gather_geometry(s, x, kpad, &prm, &smem, s->k * BLOCK_K * 2), "conv2d_fwd_gather: shape not supported (see dcv_conv2d_gather_supported)");
  prm.act = act; prm.slope = slope; prm.bias = bias; prm.x = reinterpret_cast<const __nv_bfloat16*>(x); prm.y = reinterpret_cast<__nv_bfloat16*>(y);
  const bool fused_stats = stats_nc && (stats_flags & DCV_STATS_CHANNEL_TOTALS) && (stats_flags & DCV_STATS_IN_EPILOGUE);   // per-channel totals in the epilogue, credited to image 0 of stats_nc[n][k][2] (rows of the other images stay zero)
  prm.stats = fused_stats ? stats_nc : nullptr;
  CUtensorMap mw;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)s->k};
    const cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)s->k};
    if (make_map(&mw, w_col, 2, dims, strides, box)) return 1;
  }
  const int grid = prm.total_tiles < num_sms() ? prm.total_tiles : num_sms();
  if (s->k == 128) {
    auto kern = conv_fwd_tc_gather_kernel<128>;
    static size_t configured = 0;
    if (configured < smem) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
    kern<<<grid, GA_THREADS, smem, st>>>(mw, prm);
  } else {
    auto kern = conv_fwd_tc_gather_kernel<64>;
    static size_t configured = 0;
    if (configured < smem) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
    kern<<<grid, GA_THREADS, smem, st>>>(mw, prm);
  }
  DCV_LAUNCH_CHECK("conv_fwd_tc_gather_kernel");
  if (stats_nc && !fused_stats) return dcv_norm_stats(y, stats_nc, s->n, s->p * s->q, s->k, DCV_BF16, DCV_ACC_PREZEROED | (stats_flags & DCV_STATS_CHANNEL_TOTALS), st);
  return 0;
}

// dw_col: [K][kpad] fp32 in the gather K order (overwritten).
int conv_wgrad_tc_gather(const dcv_conv_shape* s, const void* x, const void* dy, float* dw_col, int kpad, bool prezeroed, cudaStream_t st) {
  using namespace tc;
  DCV_REQUIRE(x && dy && dw_col, "conv2d_wgrad_gather: null pointer");
  DCV_REQUIRE(reinterpret_cast<uintptr_t>(dy) % 16 == 0 && reinterpret_cast<uintptr_t>(dw_col) % 16 == 0, "conv2d_wgrad_gather: pointers must be 16-byte aligned");
  GatherWgradParams wp{}; size_t smem_fwd = 0;
  DCV_REQUIRE(gather_geometry(s, x, kpad, &wp.g, &smem_fwd, s->k * BLOCK_K * 2), "conv2d_wgrad_gather: shape not supported (see dcv_conv2d_gather_supported)");
  wp.g.x = reinterpret_cast<const __nv_bfloat16*>(x);
  wp.dw_col = dw_col;
  const size_t smem = 1024 + (size_t)GW_NA * 2 * BLOCK_M * 128 + (size_t)GA_STAGES * wp.g.kblocks * BLOCK_M * 128 + (size_t)GA_ROW_BUFS * wp.g.rows_bytes + 256;
  DCV_REQUIRE(smem <= 227 * 1024, "conv2d_wgrad_gather: %zu bytes of shared memory needed", smem);
  zero_accumulator(dw_col, (size_t)s->k * kpad * sizeof(float), st, prezeroed);
  CUtensorMap mdy;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)s->k, (cuuint64_t)s->q, (cuuint64_t)s->p, (cuuint64_t)s->n};
    const cuuint64_t strides[3] = {(cuuint64_t)s->k * 2, (cuuint64_t)s->q * s->k * 2, (cuuint64_t)s->p * s->q * s->k * 2};
    const cuuint32_t box[4] = {64u, (cuuint32_t)BLOCK_M, 1u, 1u};
    if (make_map(&mdy, dy, 4, dims, strides, box)) return 1;
  }
  auto kern = conv_wgrad_tc_gather_kernel;
  static size_t configured = 0;
  if (configured < smem) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
  const int grid = wp.g.total_tiles < num_sms() ? wp.g.total_tiles : num_sms();
  kern<<<grid, GA_THREADS, smem, st>>>(mdy, wp);
  DCV_LAUNCH_CHECK("conv_wgrad_tc_gather_kernel");
  return 0;
}

namespace tc {
static int pairs_dy_slots(int k) { return k > 64 ? 2 : 4; }   // 64 KB of dy boxes in flight
// Geometry of the pixel-pair kernels; false when the layer is not of that form (stride_w = 2, at most 4 input channels, window of 8 pixels, R <= 8).
static bool pairs_geometry(const dcv_conv_shape* s, const void* x, PairParams* prm, size_t* smem_fwd, size_t* smem_wgrad) {
  if (s->stride_w != 2 || s->dil_w != 1 || s->c < 1 || s->c > 4 || s->r < 1 || s->r > 8 || s->pad_w < 0 || s->pad_h < 0 || (s->k != 64 && s->k != 128)) return false;
  const int e = s->pad_w & 1, wc = s->w * s->c;
  if (s->s + e > 8 || wc % 8 != 0 || wc * 2 > kZeroRowBytes || reinterpret_cast<uintptr_t>(x) % 16 != 0) return false;
  prm->s = *s;
  prm->hp = (s->pad_w + e) / 2;
  prm->lpad = (2 * prm->hp * s->c + 7) / 8 * 8;
  int need = prm->lpad + 2 * (s->q + 2 - prm->hp) * s->c + 8;   // the last chunk a stored pixel reads (line q - 1 + 3) lies inside the row
  if (need < prm->lpad + wc) need = prm->lpad + wc;
  prm->rowlen = (need + 7) / 8 * 8;
  prm->rows_bytes = s->r * prm->rowlen * 2;
  prm->tiles_q = (s->q + BLOCK_M - 1) / BLOCK_M;
  const long long tiles = (long long)s->n * s->p * prm->tiles_q;
  if (tiles >= (1ll << 31) || (long long)s->n * s->p * s->q < 128) return false;
  prm->total_tiles = (int)tiles;
  *smem_fwd = 1024 + (size_t)PR_KBLOCKS * s->k * BLOCK_K * 2 + (size_t)PR_FWD_STAGES * PR_STAGE_BYTES + (size_t)PR_ROW_BUFS * prm->rows_bytes + 512;
  *smem_wgrad = 1024 + (size_t)pairs_dy_slots(s->k) * (s->k / 64) * BLOCK_M * 128 + (size_t)PR_WG_STAGES * PR_STAGE_BYTES + (size_t)PR_ROW_BUFS * prm->rows_bytes + 512;
  return *smem_fwd <= 227 * 1024 && *smem_wgrad <= 227 * 1024;
}
}  // namespace tc

bool conv_tc_pairs_supported(const dcv_conv_shape* s, const void* x, int dtype) {
  if (!s || dtype != DCV_BF16 || tc::encode_tiled() == nullptr) return false;
  tc::PairParams prm{}; size_t a = 0, b = 0;
  return tc::pairs_geometry(s, x, &prm, &a, &b);
}

// w_col: [K][256] bf16 in the pixel-pair K order (dcv_pairs_pack_weight).
int conv_fwd_tc_pairs(const dcv_conv_shape* s, const void* x, const void* w_col, const float* bias, void* y, float* stats_nc, int act, float slope, int stats_flags, cudaStream_t st) {
  using namespace tc;
  DCV_REQUIRE(x && w_col && y, "conv2d_fwd_pairs: null pointer");
  PairParams prm{}; size_t smem = 0, smem_w = 0;
  DCV_REQUIRE(pairs_geometry(s, x, &prm, &smem, &smem_w), "conv2d_fwd_pairs: shape not supported (see dcv_conv2d_pairs_supported)");
  prm.act = act; prm.slope = slope; prm.bias = bias; prm.x = reinterpret_cast<const __nv_bfloat16*>(x); prm.y = reinterpret_cast<__nv_bfloat16*>(y);
  CUtensorMap mw;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)PR_K, (cuuint64_t)s->k};
    const cuuint64_t strides[1] = {(cuuint64_t)PR_K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)s->k};
    if (make_map(&mw, w_col, 2, dims, strides, box)) return 1;
  }
  const int grid = prm.total_tiles < num_sms() ? prm.total_tiles : num_sms();
  if (s->k == 128) {
    auto kern = conv_fwd_tc_pairs_kernel<128>;
    static size_t configured = 0;
    if (configured < smem) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
    kern<<<grid, PR_FWD_THREADS, smem, st>>>(mw, prm);
  } else {
    auto kern = conv_fwd_tc_pairs_kernel<64>;
    static size_t configured = 0;
    if (configured < smem) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
    kern<<<grid, PR_FWD_THREADS, smem, st>>>(mw, prm);
  }
  DCV_LAUNCH_CHECK("conv_fwd_tc_pairs_kernel");
  if (stats_nc) return dcv_norm_stats(y, stats_nc, s->n, s->p * s->q, s->k, DCV_BF16, DCV_ACC_PREZEROED | (stats_flags & DCV_STATS_CHANNEL_TOTALS), st);
  return 0;
}

// dw_col: [K][256] fp32 in the pixel-pair K order (overwritten).
int conv_wgrad_tc_pairs(const dcv_conv_shape* s, const void* x, const void* dy, float* dw_col, bool prezeroed, cudaStream_t st) {
  using namespace tc;
  DCV_REQUIRE(x && dy && dw_col, "conv2d_wgrad_pairs: null pointer");
  DCV_REQUIRE(reinterpret_cast<uintptr_t>(dy) % 16 == 0 && reinterpret_cast<uintptr_t>(dw_col) % 16 == 0, "conv2d_wgrad_pairs: pointers must be 16-byte aligned");
  PairWgradParams wp{}; size_t smem_f = 0, smem = 0;
  DCV_REQUIRE(pairs_geometry(s, x, &wp.g, &smem_f, &smem), "conv2d_wgrad_pairs: shape not supported (see dcv_conv2d_pairs_supported)");
  wp.g.x = reinterpret_cast<const __nv_bfloat16*>(x);
  wp.dw_col = dw_col;
  wp.dy_slots = pairs_dy_slots(s->k);
  zero_accumulator(dw_col, (size_t)s->k * PR_K * sizeof(float), st, prezeroed);
  CUtensorMap mdy;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)s->k, (cuuint64_t)s->q, (cuuint64_t)s->p, (cuuint64_t)s->n};
    const cuuint64_t strides[3] = {(cuuint64_t)s->k * 2, (cuuint64_t)s->q * s->k * 2, (cuuint64_t)s->p * s->q * s->k * 2};
    const cuuint32_t box[4] = {64u, (cuuint32_t)BLOCK_M, 1u, 1u};
    if (make_map(&mdy, dy, 4, dims, strides, box)) return 1;
  }
  auto kern = conv_wgrad_tc_pairs_kernel;
  static size_t configured = 0;
  if (configured < smem) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
  const int grid = wp.g.total_tiles < num_sms() ? wp.g.total_tiles : num_sms();
  kern<<<grid, PR_WG_THREADS, smem, st>>>(mdy, wp);
  DCV_LAUNCH_CHECK("conv_wgrad_tc_pairs_kernel");
  return 0;
}

bool conv_tc_fwd_supported(const dcv_conv_shape* s, int dtype) {
  if (!s || dtype != DCV_BF16) return false;
  if (s->stride_h != 1 || s->stride_w != 1 || s->dil_h != 1 || s->dil_w != 1) return false;
  if (s->c % 64 != 0 || s->k % 64 != 0 || s->pad_h < 0 || s->pad_w < 0) return false;
  if ((long long)s->n * s->p * s->q < 128) return false;
  return tc::encode_tiled() != nullptr;
}

int conv_fwd_tc(const dcv_conv_shape* s, const void* x, const void* w, const float* bias, void* y, float* stats_nc, int act, float slope, int stats_flags, cudaStream_t st) {
  using namespace tc;
  DCV_REQUIRE(x && w && y, "conv2d_fwd (tcgen05): null pointer");
  DCV_REQUIRE((reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(w) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0), "conv2d_fwd (tcgen05): pointers must be 16-byte aligned");
  // BatchNorm-only blocks need per-CHANNEL totals only: the epilogue produces them (credited to image 0 of stats_nc[n][k][2]; the rows of the other
  // images stay zero). A GroupNorm / InstanceNorm needs per-(image, channel) sums: pixel tiles span images, so those come from the statistics kernel.
  const bool fused_stats = stats_nc && (stats_flags & DCV_STATS_CHANNEL_TOTALS) && (stats_flags & DCV_STATS_IN_EPILOGUE) && s->k <= kMaxStatChannels;
  if (fwd_halo_applicable(s)) {
    if (conv_fwd_tc_halo(s, x, w, bias, y, act, slope, fused_stats ? stats_nc : nullptr, st)) return 1;
    if (stats_nc && !fused_stats) return dcv_norm_stats(y, stats_nc, s->n, s->p * s->q, s->k, DCV_BF16, DCV_ACC_PREZEROED | (stats_flags & DCV_STATS_CHANNEL_TOTALS), st);
    return 0;
  }
  FwdParams prm{};
  prm.n = s->n; prm.h = s->h; prm.w = s->w; prm.c = s->c; prm.k = s->k; prm.r = s->r; prm.s = s->s; prm.pad_h = s->pad_h; prm.pad_w = s->pad_w; prm.p = s->p; prm.q = s->q;
  pick_pixel_tile(s->q, s->p, s->n, &prm.tw, &prm.th, &prm.tn);
  prm.tiles_w = (s->q + prm.tw - 1) / prm.tw; prm.tiles_h = (s->p + prm.th - 1) / prm.th; prm.tiles_n = (s->n + prm.tn - 1) / prm.tn;
  const int n_tile = s->k % 256 == 0 ? 256 : (s->k % 128 == 0 ? 128 : 64);
  prm.n_tiles_k = s->k / n_tile;
  const long long tiles = (long long)prm.tiles_w * prm.tiles_h * prm.tiles_n * prm.n_tiles_k;
  DCV_REQUIRE(tiles < (1ll << 31), "conv2d_fwd (tcgen05): too many tiles");
  prm.total_tiles = (int)tiles;
  prm.act = act; prm.slope = slope; prm.bias = bias; prm.y = reinterpret_cast<__nv_bfloat16*>(y);
  prm.stats = fused_stats ? stats_nc : nullptr;

  CUtensorMap mx, mw;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)s->c, (cuuint64_t)s->w, (cuuint64_t)s->h, (cuuint64_t)s->n};
    const cuuint64_t strides[3] = {(cuuint64_t)s->c * 2, (cuuint64_t)s->w * s->c * 2, (cuuint64_t)s->h * s->w * s->c * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)prm.tw, (cuuint32_t)prm.th, (cuuint32_t)prm.tn};
    if (make_map(&mx, x, 4, dims, strides, box)) return 1;
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)s->r * s->s * s->c, (cuuint64_t)s->k};
    const cuuint64_t strides[1] = {(cuuint64_t)s->r * s->s * s->c * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)n_tile};
    if (make_map(&mw, w, 2, dims, strides, box)) return 1;
  }
  int rc;
  if (n_tile == 256) rc = launch_fwd<256>(mx, mw, prm, st);
  else if (n_tile == 128) rc = launch_fwd<128>(mx, mw, prm, st);
  else if (prm.n_tiles_k == 1 && s->r * s->s * (s->c / BLOCK_K) <= kMaxResidentKb && getenv("DCV_TC_NO_RESIDENT") == nullptr) rc = launch_fwd<64, 1>(mx, mw, prm, st);
  else rc = launch_fwd<64>(mx, mw, prm, st);
  if (rc) return rc;
  if (stats_nc && !fused_stats) return dcv_norm_stats(y, stats_nc, s->n, s->p * s->q, s->k, DCV_BF16, DCV_ACC_PREZEROED | (stats_flags & DCV_STATS_CHANNEL_TOTALS), st);
  return 0;
}

// ---- weight gradient ---------------------------------------------------------------------------------------------------------------------
// dw[k][r][s][c] = sum over output pixels of dy[pix][k] * x[pix + (r - pad, s - pad)][c]     (stride 1, dilation 1)
// GEMM per filter tap: M = 128 output channels k, N = 64 input channels c, K = pixels (128 per step). Both operands are MN-major: a TMA box
// {64 channels, tw, th, tn} lands as 128 pixel rows x 128 B, i.e. K (pixel) rows of one 64-element MN atom; the k side uses two atoms (LBO apart).
// A CTA owns (128-k tile, filter row r, 64-c block, pixel split): per pixel tile it loads the dy slab once and the S column-shifted x slabs of
// its filter row, and accumulates S tap tiles [128 x 64] in tensor memory over all its pixel tiles; the epilogue adds them into dw with fp32 atomics
// (dw zeroed first; the splits of the pixel range combine there too).
namespace tc {

struct WgradParams {
  int n, h, w, c, k, r, s, pad_h, pad_w, p, q;
  int tw, th, tn, tiles_w, tiles_h, tiles_n, pixel_tiles;
  int k_tiles, c_tiles, splits;
  int chan_taps;   // 1: the S_TAPS atoms of the B operand are consecutive 64-channel blocks of x (1x1 filters: c >= 128), not column taps
  int row_pairs, r_units;   // row_pairs: layers with at most 64 output channels — a unit owns filter rows (rr, rr + 1), see the kernel; r_units = units along r
  float* dw;
};

constexpr int WG_SLAB = BLOCK_M * 128;   // one TMA box: 128 pixel rows x 128 B = 16 KB
constexpr int WG_MAX_S = 3;

constexpr int WG_NA = 3, WG_NB = 2;   // dy-tile slots (2 slabs = 32 KB each) and x slots (S_TAPS column-shifted slabs = up to 48 KB each): 192 KB in flight

template <int S_TAPS>
__global__ void __launch_bounds__(kThreads, 1) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgradParams prm) {
  constexpr int TMEM_COLS = S_TAPS * 64 <= 64 ? 64 : (S_TAPS * 64 <= 128 ? 128 : 256);
  constexpr int A_BYTES = 2 * WG_SLAB;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr int B_BYTES = S_TAPS * WG_SLAB;
  const uint32_t smem_a = base, smem_b = base + WG_NA * A_BYTES;
  const uint32_t bars = smem_b + WG_NB * B_BYTES;
  auto afull = [&](int i) { return bars + 8u * i; };
  auto aempty = [&](int i) { return bars + 8u * (WG_NA + i); };
  auto bfull = [&](int i) { return bars + 8u * (2 * WG_NA + i); };
  auto bempty = [&](int i) { return bars + 8u * (2 * WG_NA + WG_NB + i); };
  const uint32_t done = bars + 8u * (2 * WG_NA + 2 * WG_NB);
  const uint32_t tmem_slot = done + 8u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work unit
  int u = blockIdx.x;
  const int split = u % prm.splits; u /= prm.splits;
  const int ct = u % prm.c_tiles; u /= prm.c_tiles;
  const int rr = (u % prm.r_units) * (prm.row_pairs ? 2 : 1); const int kt = u / prm.r_units;
  const int my_tiles = split < prm.pixel_tiles ? (prm.pixel_tiles - split + prm.splits - 1) / prm.splits : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_NA; ++i) { mbar_init(afull(i), 1); mbar_init(aempty(i), 1); }
    for (int i = 0; i < WG_NB; ++i) { mbar_init(bfull(i), 1); mbar_init(bempty(i), 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Output channels 64..127 of this k tile: when the layer has only 64 of them the second MN atom of the dy operand is all zeros. It is zeroed here once
  // instead of being zero-filled by a second TMA box per pixel tile (16 KB of shared-memory writes per tile for nothing).
  // Row pairs (at most 64 output channels): the second atom is not wasted on zeros but holds the SAME dy tile one image row up (a second TMA box at p0 - 1,
  // rows outside the image zero-filled): accumulator rows 64..127 then are sum_p dy[p - 1][k] * x[p + rr - pad][c] = the gradient of filter row rr + 1.
  // One unit does two filter rows with a full M = 128 (the pixel-tile grid covers one extra row so that the shifted copy reaches the last dy row).
  const bool second_atom = prm.k > kt * 128 + 64;
  const bool pair = prm.row_pairs && !second_atom && rr + 1 < prm.r;
  if (!second_atom && !pair) {
    for (int i = threadIdx.x; i < WG_NA * (WG_SLAB / 16); i += kThreads) {
      const int slot = i / (WG_SLAB / 16), off = i - slot * (WG_SLAB / 16);
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_a + slot * A_BYTES + WG_SLAB + off * 16), "r"(0) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor-core (async proxy) reads
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (elect_one()) {   // ===== producer: per pixel tile one dy tile (A ring) then its S column-shifted x tiles (B ring)
      int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
      // CTAs of different units walk the same pixel tiles: start each unit at a different tile so that concurrently running CTAs do not all pull the
      // same lines out of one L2 slice at the same moment (measured 3x slower on the 7x7 maps when they did)
      const int rot = my_tiles > 0 ? (int)(((unsigned)(blockIdx.x / prm.splits) * 2654435761u) % (unsigned)my_tiles) : 0;
      for (int i = 0; i < my_tiles; ++i) {
        int it = i + rot; if (it >= my_tiles) it -= my_tiles;
        int pt = split + it * prm.splits;
        const int pw = pt % prm.tiles_w; pt /= prm.tiles_w;
        const int ph = pt % prm.tiles_h; const int pn = pt / prm.tiles_h;
        const int q0 = pw * prm.tw, p0 = ph * prm.th, n0 = pn * prm.tn;
        mbar_wait(aempty(as), aph ^ 1u);
        mbar_expect_tx(afull(as), (second_atom || pair) ? A_BYTES : WG_SLAB);
        tma_load_4d(smem_a + as * A_BYTES, &map_dy, afull(as), kt * 128, q0, p0, n0);
        if (second_atom) tma_load_4d(smem_a + as * A_BYTES + WG_SLAB, &map_dy, afull(as), kt * 128 + 64, q0, p0, n0);
        else if (pair) tma_load_4d(smem_a + as * A_BYTES + WG_SLAB, &map_dy, afull(as), kt * 128, q0, p0 - 1, n0);
        if (++as == WG_NA) { as = 0; aph ^= 1u; }
        mbar_wait(bempty(bs), bph ^ 1u);
        mbar_expect_tx(bfull(bs), B_BYTES);
#pragma unroll
        for (int ss = 0; ss < S_TAPS; ++ss) {
          if (prm.chan_taps) tma_load_4d(smem_b + bs * B_BYTES + ss * WG_SLAB, &map_x, bfull(bs), (ct * S_TAPS + ss) * 64, q0 - prm.pad_w, p0 + rr - prm.pad_h, n0);
          else tma_load_4d(smem_b + bs * B_BYTES + ss * WG_SLAB, &map_x, bfull(bs), ct * 64, q0 + ss - prm.pad_w, p0 + rr - prm.pad_h, n0);
        }
        if (++bs == WG_NB) { bs = 0; bph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // ===== MMA issuer
      // ONE MMA covers all S_TAPS taps: N = S_TAPS x 64, the taps' x slabs being the 64-element MN atoms of B, one slab (LBO) apart. Issuing the taps
      // separately re-read the dy operand from shared memory once per tap (6 KB per 32 MMA cycles: shared-memory-read bound at 2/3 of the tensor rate).
      constexpr uint32_t idesc = make_idesc(128, S_TAPS * 64, 1, 1);
      int as = 0, bs = 0; uint32_t aph = 0, bph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        mbar_wait(afull(as), aph);
        mbar_wait(bfull(bs), bph);
        tc_fence_after();
        const uint32_t sa = smem_a + as * A_BYTES, sb = smem_b + bs * B_BYTES;
#pragma unroll
        for (int ks = 0; ks < BLOCK_M / UMMA_K; ++ks) {
          // MN-major: 16 pixel rows per MMA = 2048 B; SBO = 1024 B between 8-row groups; LBO = one slab between 64-channel atoms (of k for A, of the taps for B)
          const uint64_t adesc = make_desc(sa + ks * UMMA_K * 128, WG_SLAB, 1024);
          const uint64_t bdesc = make_desc(sb + ks * UMMA_K * 128, WG_SLAB, 1024);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (i | ks) != 0);
        }
        umma_commit(bempty(bs));
        umma_commit(aempty(as));
        if (++bs == WG_NB) { bs = 0; bph ^= 1u; }
        if (++as == WG_NA) { as = 0; aph ^= 1u; }
      }
      umma_commit(done);
    }
  } else if (my_tiles > 0) {
    const int quarter = warp & 3;
    const int arow = quarter * 32 + lane;                              // accumulator row
    const int krow = kt * 128 + (pair ? (arow & 63) : arow);           // output channel; a row pair's rows 64..127 are the channels again, for filter row rr + 1
    const int frow = rr + (pair ? (arow >> 6) : 0);
    mbar_wait(done, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
    for (int ss = 0; ss < S_TAPS; ++ss) {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + ss * 64 + c0, v);
        if (krow < prm.k && (pair || arow < 64 || second_atom)) {
          float* dst = prm.chan_taps ? prm.dw + (((size_t)krow * prm.r + frow) * prm.s) * prm.c + (ct * S_TAPS + ss) * 64 + c0
                                     : prm.dw + (((size_t)krow * prm.r + frow) * prm.s + ss) * prm.c + ct * 64 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)   // 16-byte vector reductions: 4x fewer L2 atomic operations
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                         "f"(__uint_as_float(v[j + 3])) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int S_TAPS>
static int launch_wgrad(const CUtensorMap& mdy, const CUtensorMap& mx, const WgradParams& prm, int grid, cudaStream_t st) {
  auto kern = conv_wgrad_tc_kernel<S_TAPS>;
  const size_t smem = 1024 + (size_t)WG_NA * 2 * WG_SLAB + (size_t)WG_NB * S_TAPS * WG_SLAB + 256;
  static bool configured = false;
  if (!configured) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = true; }
  kern<<<grid, kThreads, smem, st>>>(mdy, mx, prm);
  DCV_LAUNCH_CHECK("conv_wgrad_tc_kernel");
  return 0;
}

}  // namespace tc

bool conv_tc_wgrad_supported(const dcv_conv_shape* s, int dtype) {
  if (!conv_tc_fwd_supported(s, dtype)) return false;
  return s->s <= tc::WG_MAX_S;
}

size_t conv_wgrad_tc_workspace(const dcv_conv_shape*) { return 0; }

int conv_wgrad_tc(const dcv_conv_shape* s, const void* x, const void* dy, float* dw, void* /*workspace*/, bool prezeroed, cudaStream_t st) {
  using namespace tc;
  DCV_REQUIRE(x && dy && dw, "conv2d_wgrad (tcgen05): null pointer");
  DCV_REQUIRE((reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(dy) % 16 == 0), "conv2d_wgrad (tcgen05): pointers must be 16-byte aligned");
  WgradParams prm{};
  prm.n = s->n; prm.h = s->h; prm.w = s->w; prm.c = s->c; prm.k = s->k; prm.r = s->r; prm.s = s->s; prm.pad_h = s->pad_h; prm.pad_w = s->pad_w; prm.p = s->p; prm.q = s->q;
  // at most 64 output channels (M = 128 would be half zero-fill): a unit takes two filter rows, the second through a dy tile shifted by one image row;
  // the pixel tiles then cover p = 0 .. P (one extra row: the shifted copy of the last dy row)
  prm.row_pairs = (s->k <= 64 && s->r >= 2 && s->s > 1 && getenv("DCV_WGRAD_NO_ROW_PAIRS") == nullptr) ? 1 : 0;
  const int p_cover = s->p + prm.row_pairs;
  pick_pixel_tile(s->q, p_cover, s->n, &prm.tw, &prm.th, &prm.tn);
  prm.tiles_w = (s->q + prm.tw - 1) / prm.tw; prm.tiles_h = (p_cover + prm.th - 1) / prm.th; prm.tiles_n = (s->n + prm.tn - 1) / prm.tn;
  const long long ptiles = (long long)prm.tiles_w * prm.tiles_h * prm.tiles_n;
  DCV_REQUIRE(ptiles < (1ll << 30), "conv2d_wgrad (tcgen05): too many pixel tiles");
  prm.pixel_tiles = (int)ptiles;
  prm.k_tiles = (s->k + 127) / 128; prm.c_tiles = s->c / 64;
  // 1x1-wide filters (the im2col GEMM of the stem: c = kpad = 192): group 3 or 2 channel blocks into one MMA instead of column taps
  int taps = s->s;
  if (s->s == 1 && prm.c_tiles >= 2) {
    taps = prm.c_tiles % 3 == 0 ? 3 : (prm.c_tiles % 2 == 0 ? 2 : 1);
    if (taps > 1) { prm.chan_taps = 1; prm.c_tiles /= taps; }
  }
  if (prm.chan_taps) prm.row_pairs = 0;
  prm.r_units = prm.row_pairs ? (s->r + 1) / 2 : s->r;
  const int units = prm.k_tiles * prm.r_units * prm.c_tiles;
  // One CTA per SM fits (192 KB of shared memory): pick the pixel split that minimises (waves of CTAs) x (pixel tiles per CTA + fixed cost) — e.g.
  // 3 units x 49 splits = 147 CTAs in one wave, never 297 CTAs in two waves plus a one-CTA tail.
  int splits = 1;
  long long best_cost = -1;
  const int max_splits = prm.pixel_tiles < 4 * num_sms() ? prm.pixel_tiles : 4 * num_sms();
  for (int sp = 1; sp <= max_splits; ++sp) {
    const long long waves = ((long long)units * sp + num_sms() - 1) / num_sms();
    const long long cost = waves * ((prm.pixel_tiles + sp - 1) / sp + 5);   // + 5: a CTA's prologue + atomic epilogue cost about 5 pixel tiles (measured, 6.8 us vs 1.3 us)
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = sp; }
  }
  if (const char* e = getenv("DCV_WGRAD_SPLITS")) { const int v = atoi(e); if (v >= 1 && v <= prm.pixel_tiles) splits = v; }   // tuning aid
  prm.splits = splits;
  prm.dw = dw;
  zero_accumulator(dw, (size_t)s->k * s->r * s->s * s->c * sizeof(float), st, prezeroed);

  CUtensorMap mdy, mx;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)s->k, (cuuint64_t)s->q, (cuuint64_t)s->p, (cuuint64_t)s->n};
    const cuuint64_t strides[3] = {(cuuint64_t)s->k * 2, (cuuint64_t)s->q * s->k * 2, (cuuint64_t)s->p * s->q * s->k * 2};
    const cuuint32_t box[4] = {64u, (cuuint32_t)prm.tw, (cuuint32_t)prm.th, (cuuint32_t)prm.tn};
    if (make_map(&mdy, dy, 4, dims, strides, box)) return 1;
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)s->c, (cuuint64_t)s->w, (cuuint64_t)s->h, (cuuint64_t)s->n};
    const cuuint64_t strides[3] = {(cuuint64_t)s->c * 2, (cuuint64_t)s->w * s->c * 2, (cuuint64_t)s->h * s->w * s->c * 2};
    const cuuint32_t box[4] = {64u, (cuuint32_t)prm.tw, (cuuint32_t)prm.th, (cuuint32_t)prm.tn};
    if (make_map(&mx, x, 4, dims, strides, box)) return 1;
  }
  const int grid = units * splits;
  if (taps == 1) return launch_wgrad<1>(mdy, mx, prm, grid, st);
  if (taps == 2) return launch_wgrad<2>(mdy, mx, prm, grid, st);
  return launch_wgrad<3>(mdy, mx, prm, grid, st);
}

}  // namespace dcv
