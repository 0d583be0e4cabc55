// Public convolution entry points: pick the algorithm, forward to the direct (CUDA-core) or tcgen05 (tensor-core) kernels.
#include "common.cuh"

namespace dcv {
int conv_fwd_direct(const dcv_conv_shape*, const void*, const void*, const float*, void*, float*, int, float, int, cudaStream_t);
int conv_dgrad_direct(const dcv_conv_shape*, const void*, const void*, void*, int, cudaStream_t);
int conv_wgrad_direct(const dcv_conv_shape*, const void*, const void*, float*, int, bool prezeroed, cudaStream_t);
// conv_tc.cu
bool conv_tc_fwd_supported(const dcv_conv_shape*, int dtype);
bool conv_tc_wgrad_supported(const dcv_conv_shape*, int dtype);
int conv_fwd_tc(const dcv_conv_shape*, const void*, const void*, const float*, void*, float*, int, float, int stats_flags, cudaStream_t);
int conv_wgrad_tc(const dcv_conv_shape*, const void*, const void*, float*, void*, bool prezeroed, cudaStream_t);
bool conv_fwd_tc_gather_supported(const dcv_conv_shape*, const void* x, int kpad, int dtype);
int conv_fwd_tc_gather(const dcv_conv_shape*, const void* x, const void* w_col, int kpad, const float* bias, void* y, float* stats_nc, int act, float slope, int stats_flags, cudaStream_t);
int conv_wgrad_tc_gather(const dcv_conv_shape*, const void* x, const void* dy, float* dw_col, int kpad, bool prezeroed, cudaStream_t);
bool conv_tc_pairs_supported(const dcv_conv_shape*, const void* x, int dtype);
int conv_fwd_tc_pairs(const dcv_conv_shape*, const void* x, const void* w_col, const float* bias, void* y, float* stats_nc, int act, float slope, int stats_flags, cudaStream_t);
int conv_wgrad_tc_pairs(const dcv_conv_shape*, const void* x, const void* dy, float* dw_col, bool prezeroed, cudaStream_t);
size_t conv_wgrad_tc_workspace(const dcv_conv_shape*);

static dcv_conv_shape dgrad_as_fwd(const dcv_conv_shape& s) {
  dcv_conv_shape t{};
  t.n = s.n; t.h = s.p; t.w = s.q; t.c = s.k; t.k = s.c; t.r = s.r; t.s = s.s;
  t.stride_h = t.stride_w = 1; t.pad_h = s.dil_h * (s.r - 1) - s.pad_h; t.pad_w = s.dil_w * (s.s - 1) - s.pad_w;
  t.dil_h = s.dil_h; t.dil_w = s.dil_w; t.p = s.h; t.q = s.w;
  return t;
}
}  // namespace dcv

extern "C" {

int dcv_conv2d_tc_supported(const dcv_conv_shape* shape, int dtype, int op) {
  using namespace dcv;
  if (!shape) return 0;
  if (op == 0) return conv_tc_fwd_supported(shape, dtype) ? 1 : 0;
  if (op == 1) {
    if (shape->stride_h != 1 || shape->stride_w != 1) return 0;
    const dcv_conv_shape t = dgrad_as_fwd(*shape);
    return (t.pad_h >= 0 && t.pad_w >= 0 && conv_tc_fwd_supported(&t, dtype)) ? 1 : 0;
  }
  return conv_tc_wgrad_supported(shape, dtype) ? 1 : 0;
}

int dcv_conv2d_fwd(const dcv_conv_shape* shape, const void* x, const void* w, const float* bias, void* y, float* stats_nc,
                   int act, float slope, int dtype, int algo, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(shape, "conv2d_fwd: null shape");
  cudaStream_t st = as_stream(stream);
  const bool tc_ok = conv_tc_fwd_supported(shape, dtype);
  zero_accumulator(stats_nc, (size_t)shape->n * shape->k * 2 * sizeof(float), st, (acc_prezeroed & DCV_ACC_PREZEROED) != 0);
  DCV_REQUIRE(algo != DCV_ALGO_TCGEN05 || tc_ok, "conv2d_fwd: tcgen05 algorithm does not support this shape/dtype (needs bf16, c %% 64 == 0, k %% 16 == 0, stride 1, dilation 1)");
  if (algo == DCV_ALGO_TCGEN05 || (algo == DCV_ALGO_AUTO && tc_ok)) return conv_fwd_tc(shape, x, w, bias, y, stats_nc, act, slope, acc_prezeroed, st);
  return conv_fwd_direct(shape, x, w, bias, y, stats_nc, act, slope, dtype, st);
}

int dcv_conv2d_gather_supported(const dcv_conv_shape* shape, const void* x, int kpad, int dtype) {
  return dcv::conv_fwd_tc_gather_supported(shape, x, kpad, dtype) ? 1 : 0;
}

int dcv_conv2d_fwd_gather(const dcv_conv_shape* shape, const void* x, const void* w_col, int kpad, const float* bias, void* y, float* stats_nc,
                          int act, float slope, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(shape, "conv2d_fwd_gather: null shape");
  cudaStream_t st = as_stream(stream);
  zero_accumulator(stats_nc, (size_t)shape->n * shape->k * 2 * sizeof(float), st, (acc_prezeroed & DCV_ACC_PREZEROED) != 0);
  return conv_fwd_tc_gather(shape, x, w_col, kpad, bias, y, stats_nc, act, slope, acc_prezeroed, st);
}

int dcv_conv2d_wgrad_gather(const dcv_conv_shape* shape, const void* x, const void* dy, float* dw_col, int kpad, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(shape, "conv2d_wgrad_gather: null shape");
  return conv_wgrad_tc_gather(shape, x, dy, dw_col, kpad, acc_prezeroed != 0, as_stream(stream));
}

int dcv_conv2d_pairs_supported(const dcv_conv_shape* shape, const void* x, int dtype) {
  return dcv::conv_tc_pairs_supported(shape, x, dtype) ? 1 : 0;
}

int dcv_conv2d_fwd_pairs(const dcv_conv_shape* shape, const void* x, const void* w_col, const float* bias, void* y, float* stats_nc, int act, float slope, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(shape, "conv2d_fwd_pairs: null shape");
  cudaStream_t st = as_stream(stream);
  zero_accumulator(stats_nc, (size_t)shape->n * shape->k * 2 * sizeof(float), st, (acc_prezeroed & DCV_ACC_PREZEROED) != 0);
  return conv_fwd_tc_pairs(shape, x, w_col, bias, y, stats_nc, act, slope, acc_prezeroed, st);
}

int dcv_conv2d_wgrad_pairs(const dcv_conv_shape* shape, const void* x, const void* dy, float* dw_col, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(shape, "conv2d_wgrad_pairs: null shape");
  return conv_wgrad_tc_pairs(shape, x, dy, dw_col, acc_prezeroed != 0, as_stream(stream));
}

int dcv_conv2d_dgrad(const dcv_conv_shape* shape, const void* dy, const void* w, const void* wt, void* dx, int dtype, int algo, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(shape, "conv2d_dgrad: null shape");
  cudaStream_t st = as_stream(stream);
  bool tc_ok = false;
  dcv_conv_shape t{};
  if (shape->stride_h == 1 && shape->stride_w == 1 && wt) {
    t = dgrad_as_fwd(*shape);
    tc_ok = t.pad_h >= 0 && t.pad_w >= 0 && conv_tc_fwd_supported(&t, dtype);
  }
  DCV_REQUIRE(algo != DCV_ALGO_TCGEN05 || tc_ok, "conv2d_dgrad: tcgen05 algorithm does not support this shape/dtype (or `wt` is NULL)");
  if (algo == DCV_ALGO_TCGEN05 || (algo == DCV_ALGO_AUTO && tc_ok)) return conv_fwd_tc(&t, dy, wt, nullptr, dx, nullptr, DCV_ACT_NONE, 0.f, 0, st);
  return conv_dgrad_direct(shape, dy, w, dx, dtype, st);
}

size_t dcv_conv2d_wgrad_workspace(const dcv_conv_shape* shape, int dtype, int algo) {
  using namespace dcv;
  if (!shape) return 0;
  const bool tc_ok = conv_tc_wgrad_supported(shape, dtype);
  if (algo == DCV_ALGO_TCGEN05 || (algo == DCV_ALGO_AUTO && tc_ok)) return tc_ok ? conv_wgrad_tc_workspace(shape) : 0;
  return 0;
}

int dcv_conv2d_wgrad(const dcv_conv_shape* shape, const void* x, const void* dy, float* dw, void* workspace, int dtype, int algo, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(shape, "conv2d_wgrad: null shape");
  cudaStream_t st = as_stream(stream);
  const bool tc_ok = conv_tc_wgrad_supported(shape, dtype);
  DCV_REQUIRE(algo != DCV_ALGO_TCGEN05 || tc_ok, "conv2d_wgrad: tcgen05 algorithm does not support this shape/dtype");
  if (algo == DCV_ALGO_TCGEN05 || (algo == DCV_ALGO_AUTO && tc_ok)) return conv_wgrad_tc(shape, x, dy, dw, workspace, acc_prezeroed != 0, st);
  return conv_wgrad_direct(shape, x, dy, dw, dtype, acc_prezeroed != 0, st);
}

}  // extern "C"
