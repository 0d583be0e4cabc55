// The two ops either side of forward/backward on the training step: softmax
// cross-entropy (value + gradient in one pass) and AdamW over a flat parameter buffer.
#include "common.cuh"

namespace dcv {

constexpr int64_t kIgnoreIndex = -100;   // torch.nn.CrossEntropyLoss default `ignore_index`

// one warp per row. Rows whose target is `ignore_index` contribute nothing and the mean runs over the other rows (torch semantics); any other target
// outside [0, n) poisons the loss with NaN instead of reading out of bounds (torch device-asserts there).
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, float* __restrict__ loss,
                                  float* __restrict__ dlogits, int m, int n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= m) return;
  int valid = 0;
  for (int i = lane; i < m; i += 32) valid += target[i] != kIgnoreIndex;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
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
  const bool ignored = t == kIgnoreIndex, bad = !ignored && (t < 0 || t >= n);
  const float inv_m = valid > 0 ? 1.f / (float)valid : 0.f;
  if (dlogits)
    for (int j = lane; j < n; j += 32) dlogits[(size_t)warp * n + j] = (ignored || bad) ? 0.f : (expf(row[j] - lse) - (j == t ? 1.f : 0.f)) * inv_m;
  if (lane == 0) {
    if (bad) atomicAdd(loss, __int_as_float(0x7fc00000));
    else if (!ignored) atomicAdd(loss, (lse - row[t]) * inv_m);
  }
}

// Evaluation metrics, accumulated over batches: acc[0] += sum of per-row cross entropies, acc[1] += rows whose argmax is the target, acc[2] += rows counted.
__global__ void classification_metrics_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target, float* __restrict__ acc, int m, int n) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= m) return;
  const int64_t t = target[warp];
  if (t == kIgnoreIndex) return;
  const float* row = logits + (size_t)warp * n;
  float mx = -INFINITY; int arg = 0x7fffffff;
  for (int j = lane; j < n; j += 32) { const float v = row[j]; if (v > mx) { mx = v; arg = j; } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {   // max with the lowest index on ties (torch.argmax)
    const float om = __shfl_xor_sync(0xffffffffu, mx, o); const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float se = 0.f;
  for (int j = lane; j < n; j += 32) se += expf(row[j] - mx);
  se = warp_sum(se);
  if (lane == 0) {
    const bool bad = t < 0 || t >= n;
    atomicAdd(acc, bad ? __int_as_float(0x7fc00000) : (mx + logf(se) - row[t]));
    if (!bad && arg == (int)t) atomicAdd(acc + 1, 1.f);
    atomicAdd(acc + 2, 1.f);
  }
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

int dcv_classification_metrics(const float* logits, const int64_t* target, float* acc3, int m, int n, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(logits && target && acc3 && m > 0 && n > 0, "classification_metrics: bad arguments");
  classification_metrics_kernel<<<(m * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(logits, target, acc3, m, n);
  DCV_LAUNCH_CHECK("classification_metrics_kernel");
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
