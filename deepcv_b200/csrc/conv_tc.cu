// tcgen05 / TMEM / TMA implicit-GEMM convolution (placeholder until the kernels land: reports "unsupported").
#include "common.cuh"

namespace dcv {
bool conv_tc_fwd_supported(const dcv_conv_shape*, int) { return false; }
bool conv_tc_wgrad_supported(const dcv_conv_shape*, int) { return false; }
int conv_fwd_tc(const dcv_conv_shape*, const void*, const void*, const float*, void*, float*, int, float, cudaStream_t) { set_error("tcgen05 convolution not built"); return 1; }
int conv_wgrad_tc(const dcv_conv_shape*, const void*, const void*, float*, void*, cudaStream_t) { set_error("tcgen05 convolution not built"); return 1; }
size_t conv_wgrad_tc_workspace(const dcv_conv_shape*) { return 0; }
}  // namespace dcv
