// Shared helpers for the deepcv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include "../../include/deepcv_b200.h"

namespace dcv {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define DCV_REQUIRE(cond, ...)                       \
  do {                                               \
    if (!(cond)) {                                   \
      ::dcv::set_error(__VA_ARGS__);                 \
      return 1;                                      \
    }                                                \
  } while (0)

// Checks the launch (not the execution: nothing here synchronises) and counts it.
#define DCV_LAUNCH_CHECK(name)                                                            \
  do {                                                                                    \
    cudaError_t e__ = cudaPeekAtLastError();                                              \
    ::dcv::g_launches.fetch_add(1, std::memory_order_relaxed);                            \
    if (e__ != cudaSuccess) {                                                             \
      ::dcv::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));           \
      (void)cudaGetLastError();                                                           \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)

constexpr int kNumSMs = 148;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Activation on the pre-activation value, and its derivative expressed through the activation OUTPUT y.
__device__ __forceinline__ float act_apply(float v, int act, float slope) {
  switch (act) {
    case DCV_ACT_RELU: return v > 0.f ? v : 0.f;
    case DCV_ACT_LEAKY_RELU: return v > 0.f ? v : v * slope;
    case DCV_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}
__device__ __forceinline__ float act_grad_from_output(float y, int act, float slope) {
  switch (act) {
    case DCV_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case DCV_ACT_LEAKY_RELU: return y > 0.f ? 1.f : slope;
    case DCV_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 16-byte vector of T (4 floats or 8 bf16) for 128-bit global accesses.
template <typename T> struct Vec16 { static constexpr int kElems = 16 / sizeof(T); uint4 raw; };
template <typename T> __device__ __forceinline__ void vec_unpack(const uint4& raw, float* out);
template <> __device__ __forceinline__ void vec_unpack<float>(const uint4& raw, float* out) {
  out[0] = __uint_as_float(raw.x); out[1] = __uint_as_float(raw.y); out[2] = __uint_as_float(raw.z); out[3] = __uint_as_float(raw.w);
}
template <> __device__ __forceinline__ void vec_unpack<__nv_bfloat16>(const uint4& raw, float* out) {
  const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = __uint_as_float(u[i] << 16);
    out[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}
template <typename T> __device__ __forceinline__ uint4 vec_pack(const float* in);
template <> __device__ __forceinline__ uint4 vec_pack<float>(const float* in) {
  return make_uint4(__float_as_uint(in[0]), __float_as_uint(in[1]), __float_as_uint(in[2]), __float_as_uint(in[3]));
}
template <> __device__ __forceinline__ uint4 vec_pack<__nv_bfloat16>(const float* in) {
  uint32_t u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 p = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
    u[i] = *reinterpret_cast<uint32_t*>(&p);
  }
  return make_uint4(u[0], u[1], u[2], u[3]);
}

struct FastDiv {  // n / d for 0 <= n < 2^31, d >= 1
  uint32_t mul, shr, d;
  __host__ FastDiv() : mul(0), shr(0), d(1) {}
  __host__ explicit FastDiv(uint32_t d_) : d(d_) {
    if (d_ == 1) { mul = 0; shr = 0; return; }
    uint32_t l = 0;
    while ((1u << l) < d_) ++l;
    uint64_t m = ((uint64_t(1) << (31 + l)) + d_ - 1) / d_;  // ceil(2^(31+l) / d)
    mul = (uint32_t)m; shr = l;                               // valid for n < 2^31
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (uint32_t)(((uint64_t)n * mul) >> 31) >> shr; }
};


// ---- programmatic dependent launch ----------------------------------------------------------------------------------------------------------------
// `launch_pdl` launches with cudaLaunchAttributeProgrammaticStreamSerialization: the kernel's CTAs may be scheduled while the previous kernel of the stream
// is still draining. Such a kernel calls `pdl_wait()` before its first access to global data of its predecessor — it returns once the predecessor has
// completed and flushed — and `pdl_trigger()` right after, which lets its own successor begin to launch. What overlaps is the launch ramp and a prologue
// that touches no fresh data (zeroed tiles, weight fragments). Used by the few-channel kernels (conv_small.cu: a chain of 18 latency-bound kernels of
// 10-20 us per CIFAR step, -13 us per step, also inside the captured graph). Measured and NOT adopted library-wide (round 2): with every kernel of the
// ImageNet-shaped step launched this way the step got 1.3 % slower and the overlapped host -> device prefetch collapsed (7.14 -> 8.56 ms end to end).
// DCV_NO_PDL=1: plain launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool use_pdl() { static const bool on = getenv("DCV_NO_PDL") == nullptr; return on; }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = use_pdl() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
// Accumulator buffers (per-(n,c) statistics, gradient sums filled by atomics) are normally zeroed by the launcher that fills them: one memset node
// per kernel. A caller that zeroes ALL of them itself, once per step (deepcv_b200/ops.py: AccumulatorArena), says so per call (`acc_prezeroed` argument of
// the accumulating entry points) and turns those ~22 memsets of a CIFAR step into one. Per call, not per process: two models / streams / threads never
// see each other's setting.
inline void zero_accumulator(void* p, size_t bytes, cudaStream_t st, bool prezeroed) { if (p && bytes && !prezeroed) cudaMemsetAsync(p, 0, bytes, st); }
inline int grid_for(size_t work_items, int block, int max_blocks = kNumSMs * 16) {
  size_t g = (work_items + block - 1) / block;
  if (g < 1) g = 1;
  if (g > (size_t)max_blocks) g = max_blocks;
  return (int)g;
}

// Dispatch on DCV_F32 / DCV_BF16.
#define DCV_DISPATCH_DTYPE(dtype, T, ...)                                 \
  do {                                                                    \
    if ((dtype) == DCV_F32) { using T = float; __VA_ARGS__; }             \
    else if ((dtype) == DCV_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { ::dcv::set_error("unsupported dtype %d", (int)(dtype)); return 1; } \
  } while (0)

}  // namespace dcv
