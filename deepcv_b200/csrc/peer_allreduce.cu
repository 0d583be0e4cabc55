// One-shot gradient all-reduce over NVLink peer memory for the small buckets of the default net (reference: DistributedDataParallel's gradient
// averaging, /root/reference/src/deepcv/meta/ignite_training.py:373-390). The 68 KB of gradients of the CIFAR-10 `image_classifier` are pure latency for a
// collective library (~20 us exposed per step, measured). Here every rank owns a RECEIVE area in symmetric memory (mapped into every peer's address
// space; PyTorch's symmetric-memory allocator is only the plumbing) and ONE kernel per bucket does, per 4096-float slice (= one CTA):
//   1. PUSH the own slice into every peer's receive area (posted remote stores over NVLink / NVSwitch: no round trip), then publish "slice of epoch e
//      delivered" with a release store (system scope) into the peer's flag word;
//   2. wait until every peer's flag shows epoch e, then add the W slices — the own one from the local gradient buffer, the others from the local receive
//      area — in rank order 0..W-1 (the same order on every rank: bit-identical results, replicas cannot drift) and overwrite the own slice in place.
// One one-way trip + local work; no second barrier is needed because nobody reads the gradient buffer remotely. The receive area is double-buffered by
// epoch parity: a rank can be at most one all-reduce of the same bucket ahead of a peer (it cannot finish epoch e + 1 before the peer has pushed e + 1,
// i.e. finished summing e), so parity e + 2 = parity e is never overwritten while still being read. (A first pull-based version — ready barrier, remote
// loads, done barrier — measured 12-20 us per call on 2 x B200, no better than NCCL: two system-scope round trips and serialised remote loads.)
// Flags are monotonically increasing epochs (one counter per bucket slot, advanced by the last CTA to finish): nothing is ever reset, the kernel is
// CUDA-graph replayable. CTAs spin on flags, so a launch must be co-resident: at most DCV_PEER_MAX_CTAS (<< 148) CTAs, i.e. buckets up to 512 KB.
#include "common.cuh"

namespace dcv {
namespace peer {

constexpr int kThreads = 256, kVec = 4 /* float4 per thread */, kCtaFloats = kThreads * kVec * 4;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) { uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
// received data was written by a peer into this GPU's memory: read it past L1 (the same addresses held the epoch e - 2 data)
__device__ __forceinline__ float4 ld_cg_v4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// flag word of (slot, cta) written BY rank `src` — lives in the flag area of the rank that waits on it
__device__ __forceinline__ size_t flag_index(int slot, int cta, int src) { return ((size_t)slot * DCV_PEER_MAX_CTAS + cta) * DCV_PEER_MAX_WORLD + src; }

// recv[r]: rank r's receive area, float [2 parities][world sources][recv_stride]; element i of the flat buffer at index i (i < recv_stride)
__global__ void __launch_bounds__(kThreads) peer_allreduce_kernel(float* __restrict__ grads, float* const* __restrict__ recv, uint32_t* const* __restrict__ flags, int rank, int world,
                                                                  size_t recv_stride, size_t offset, size_t count, int slot, uint32_t* __restrict__ state) {
  const uint32_t epoch = state[2 * slot] + 1u;   // advanced only after every CTA of this launch has finished (below)
  const int tid = threadIdx.x;
  const size_t lo = offset + (size_t)blockIdx.x * kCtaFloats, hi = min(offset + count, lo + (size_t)kCtaFloats);
  const size_t par = (size_t)(epoch & 1u) * world * recv_stride;
  // ---- 1. own slice -> registers -> every peer's receive area
  float4 mine[kVec];
  float mine_tail = 0.f;
  const size_t tail0 = lo + ((hi - lo) / 4) * 4;   // a ragged end (count not a multiple of 4): scalar, one element per thread
#pragma unroll
  for (int u = 0; u < kVec; ++u) {
    const size_t i = lo + ((size_t)u * kThreads + tid) * 4;
    mine[u] = i + 4 <= hi ? *reinterpret_cast<const float4*>(grads + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (tail0 + tid < hi) mine_tail = grads[tail0 + tid];
  for (int p = 0; p < world; ++p) {
    if (p == rank) continue;
    float* dst = recv[p] + par + (size_t)rank * recv_stride;
#pragma unroll
    for (int u = 0; u < kVec; ++u) {
      const size_t i = lo + ((size_t)u * kThreads + tid) * 4;
      if (i + 4 <= hi) *reinterpret_cast<float4*>(dst + i) = mine[u];
    }
    if (tail0 + tid < hi) dst[tail0 + tid] = mine_tail;
  }
  __syncthreads();   // all of this CTA's pushes are ordered before the signal (cumulativity through the barrier)
  if (tid < world && tid != rank) {
    __threadfence_system();
    st_release_sys(flags[tid] + flag_index(slot, blockIdx.x, rank), epoch);
    const uint32_t* from = flags[rank] + flag_index(slot, blockIdx.x, tid);
    while ((int32_t)(ld_acquire_sys(from) - epoch) < 0) { }
  }
  __syncthreads();
  // ---- 2. sum in rank order, in place
  const float* in = recv[rank] + par;
  float4 got[DCV_PEER_MAX_WORLD][kVec];
  float got_tail[DCV_PEER_MAX_WORLD];
#pragma unroll
  for (int r = 0; r < DCV_PEER_MAX_WORLD; ++r) {
    if (r < world && r != rank) {   // all loads in flight together
#pragma unroll
      for (int u = 0; u < kVec; ++u) {
        const size_t i = lo + ((size_t)u * kThreads + tid) * 4;
        if (i + 4 <= hi) got[r][u] = ld_cg_v4(in + (size_t)r * recv_stride + i);
      }
      if (tail0 + tid < hi) got_tail[r] = __ldcg(in + (size_t)r * recv_stride + tail0 + tid);
    }
  }
#pragma unroll
  for (int u = 0; u < kVec; ++u) {
    const size_t i = lo + ((size_t)u * kThreads + tid) * 4;
    if (i + 4 <= hi) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < DCV_PEER_MAX_WORLD; ++r) {
        if (r < world) {
          const float4 v = r == rank ? mine[u] : got[r][u];
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      *reinterpret_cast<float4*>(grads + i) = acc;
    }
  }
  if (tail0 + tid < hi) {
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < DCV_PEER_MAX_WORLD; ++r) if (r < world) acc += r == rank ? mine_tail : got_tail[r];
    grads[tail0 + tid] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(&state[2 * slot + 1], 1u);
    if (done == gridDim.x - 1) { state[2 * slot + 1] = 0u; __threadfence(); state[2 * slot] = epoch; }
  }
}

}  // namespace peer
}  // namespace dcv

extern "C" {

size_t dcv_peer_flag_words(void) { return (size_t)DCV_PEER_MAX_SLOTS * DCV_PEER_MAX_CTAS * DCV_PEER_MAX_WORLD; }
size_t dcv_peer_max_floats(void) { return (size_t)DCV_PEER_MAX_CTAS * dcv::peer::kCtaFloats; }

int dcv_peer_allreduce_sum(float* grads, float* const* peer_recv_dev, uint32_t* const* peer_flags_dev, int rank, int world, size_t recv_stride, size_t offset, size_t count, int slot,
                           uint32_t* state_dev, void* stream) {
  using namespace dcv; using namespace dcv::peer;
  DCV_REQUIRE(grads && peer_recv_dev && peer_flags_dev && state_dev, "peer_allreduce_sum: null pointer");
  DCV_REQUIRE(world >= 1 && world <= DCV_PEER_MAX_WORLD && rank >= 0 && rank < world, "peer_allreduce_sum: rank %d / world %d (at most %d ranks)", rank, world, DCV_PEER_MAX_WORLD);
  DCV_REQUIRE(slot >= 0 && slot < DCV_PEER_MAX_SLOTS, "peer_allreduce_sum: slot %d outside [0, %d)", slot, DCV_PEER_MAX_SLOTS);
  DCV_REQUIRE(offset % 4 == 0 && recv_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(grads) & 15) == 0, "peer_allreduce_sum: offset / stride must be multiples of 4 floats, buffers 16-byte aligned");
  DCV_REQUIRE(offset + count <= recv_stride, "peer_allreduce_sum: slice [%zu, %zu) outside the receive area of %zu floats per source", offset, offset + count, recv_stride);
  DCV_REQUIRE(count <= dcv_peer_max_floats(), "peer_allreduce_sum: %zu floats exceed the one-shot limit of %zu (use the collective library for large buckets)", count, dcv_peer_max_floats());
  if (count == 0) return 0;
  const int grid = (int)((count + kCtaFloats - 1) / kCtaFloats);
  peer_allreduce_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(grads, peer_recv_dev, peer_flags_dev, rank, world, recv_stride, offset, count, slot, state_dev);
  DCV_LAUNCH_CHECK("peer_allreduce_kernel");
  return 0;
}

}  // extern "C"
