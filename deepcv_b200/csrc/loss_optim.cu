// The two ops either side of forward/backward on the training step: softmax
// cross-entropy (value + gradient in one pass) and AdamW over a flat parameter buffer.
#include "common.cuh"

namespace dcv {

// one warp per row
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, float* __restrict__ loss,
                                  float* __restrict__ dlogits, int m, int n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= m) return;
  const float* row = logits + (size_t)warp * n;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, row[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f;
  for (int j = lane; j < n; j += 32) se += expf(row[j] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  const int64_t t = target[warp];
  const float inv_m = 1.f / (float)m;
  if (dlogits)
    for (int j = lane; j < n; j += 32) dlogits[(size_t)warp * n + j] = (expf(row[j] - lse) - (j == t ? 1.f : 0.f)) * inv_m;
  if (lane == 0) atomicAdd(loss, (lse - row[t]) * inv_m);
}

__global__ void scale_dev_kernel(const float* __restrict__ src, const float* __restrict__ scale_dev, float* __restrict__ dst, size_t count) {
  const float sc = *scale_dev;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i] * sc;
}

__global__ void counter_add_kernel(int32_t* counter, int32_t delta) { *counter += delta; }

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t count,
                             const float* __restrict__ lr_dev, float b1, float b2, float eps, float wd, float gscale, const int32_t* __restrict__ step_dev) {
  const float lr = *lr_dev;
  const float t = (float)(*step_dev);
  const float bc1 = 1.f - powf(b1, t), bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const float grad = g[i] * gscale;
    float param = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * grad;
    const float vi = b2 * v[i] + (1.f - b2) * grad * grad;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    param -= step_size * (mi / denom);
    p[i] = param;
  }
}

}  // namespace dcv

extern "C" {

int dcv_softmax_ce(const float* logits, const int64_t* target, float* loss, float* dlogits, int m, int n, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(logits && target && loss && m > 0 && n > 0, "softmax_ce: bad arguments");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(loss, 0, sizeof(float), st);
  softmax_ce_kernel<<<(m * 32 + 255) / 256, 256, 0, st>>>(logits, target, loss, dlogits, m, n);
  DCV_LAUNCH_CHECK("softmax_ce_kernel");
  return 0;
}

int dcv_scale_by_device_scalar(const float* src, const float* scale_dev, float* dst, size_t count, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(src && scale_dev && dst, "scale_by_device_scalar: null pointer");
  if (count == 0) return 0;
  scale_dev_kernel<<<grid_for(count, 256), 256, 0, as_stream(stream)>>>(src, scale_dev, dst, count);
  DCV_LAUNCH_CHECK("scale_dev_kernel");
  return 0;
}

int dcv_counter_add(int32_t* counter_dev, int32_t delta, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(counter_dev, "counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, as_stream(stream)>>>(counter_dev, delta);
  DCV_LAUNCH_CHECK("counter_add_kernel");
  return 0;
}

int dcv_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t count, const float* lr_dev,
                   float beta1, float beta2, float eps, float weight_decay, float grad_scale, const int32_t* step_dev, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(param && grad && exp_avg && exp_avg_sq && lr_dev && step_dev, "adamw_flat: null pointer");
  if (count == 0) return 0;
  adamw_kernel<<<grid_for(count, 256, kNumSMs * 8), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, count, lr_dev, beta1, beta2, eps, weight_decay, grad_scale, step_dev);
  DCV_LAUNCH_CHECK("adamw_kernel");
  return 0;
}

}  // extern "C"
