// Fully connected head: y = act(x @ W^T + b) and its adjoints. The FC layers on the path are small (1280x10 on the
// default net, 768x1000 on the ResNet-style one: < 0.2 % of the step's FLOPs), so one generic strided, shared-memory
// tiled fp32-accumulate GEMM serves forward, dx and dW.
#include "common.cuh"

namespace dcv {

__device__ __forceinline__ float ld_any(const void* p, int dtype, size_t i) {
  return dtype == DCV_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_any(void* p, int dtype, size_t i, float v) {
  if (dtype == DCV_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

// C[i][j] = epilogue( sum_l A(i,l) * B(l,j) ),  A(i,l) = A[i*sai + l*sal],  B(l,j) = B[l*sbl + j*sbj],  C row-major [M][N]
struct GemmArgs {
  const void* A; const void* B; void* C; const float* bias;
  int M, N, K; long long sai, sal, sbl, sbj;
  int a_dtype, b_dtype, c_dtype, act; float slope;
};

__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs g) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float sA[BK][BM + 4], sB[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    for (int i = tid; i < BM * BK; i += 256) {
      // walk the contiguous dimension of each operand with consecutive threads
      int mi, ki;
      if (g.sal == 1) { ki = i % BK; mi = i / BK; } else { mi = i % BM; ki = i / BM; }
      const int m = m0 + mi, k = k0 + ki;
      sA[ki][mi] = (m < g.M && k < g.K) ? ld_any(g.A, g.a_dtype, (size_t)((long long)m * g.sai + (long long)k * g.sal)) : 0.f;
    }
    for (int i = tid; i < BN * BK; i += 256) {
      int ni, ki;
      if (g.sbl == 1) { ki = i % BK; ni = i / BK; } else { ni = i % BN; ki = i / BN; }
      const int n = n0 + ni, k = k0 + ki;
      sB[ki][ni] = (n < g.N && k < g.K) ? ld_any(g.B, g.b_dtype, (size_t)((long long)k * g.sbl + (long long)n * g.sbj)) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < g.M && n < g.N) {
        float v = acc[i][j] + (g.bias ? g.bias[n] : 0.f);
        st_any(g.C, g.c_dtype, (size_t)m * g.N + n, act_apply(v, g.act, g.slope));
      }
    }
}

static int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  dim3 grid((g.N + 63) / 64, (g.M + 63) / 64);
  gemm_kernel<<<grid, 256, 0, st>>>(g);
  DCV_LAUNCH_CHECK("gemm_kernel");
  return 0;
}

// dpre[m][n] = act'(y) * dy; db[n] += sum_m dpre (db zeroed by the wrapper)
__global__ void linear_dpre_kernel(const void* __restrict__ y, const void* __restrict__ dy, float* __restrict__ dpre, float* __restrict__ db,
                                   int m, int n, int act, float slope, int y_dtype) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n) return;
  float s = 0.f;
  for (int r = blockIdx.y; r < m; r += gridDim.y) {
    const size_t i = (size_t)r * n + col;
    const float g = ld_any(dy, y_dtype, i) * act_grad_from_output(ld_any(y, y_dtype, i), act, slope);
    dpre[i] = g;
    s += g;
  }
  if (db) atomicAdd(db + col, s);
}

}  // namespace dcv

extern "C" {

int dcv_linear_fwd(const void* x, const float* w, const float* bias, void* y, int m, int n, int k, int act, float slope,
                   int x_dtype, int y_dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(x && w && y && m > 0 && n > 0 && k > 0, "linear_fwd: bad arguments");
  GemmArgs g{x, w, y, bias, m, n, k, k, 1, 1, k, x_dtype, DCV_F32, y_dtype, act, slope};
  return launch_gemm(g, as_stream(stream));
}

int dcv_linear_bwd(const void* x, const float* w, const void* y, const void* dy, void* dx, float* dw, float* db, float* dpre_ws,
                   int m, int n, int k, int act, float slope, int x_dtype, int y_dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(x && w && y && dy && dpre_ws && m > 0 && n > 0 && k > 0, "linear_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  if (db) cudaMemsetAsync(db, 0, (size_t)n * sizeof(float), st);
  int gy = (m + 31) / 32; if (gy > 64) gy = 64;
  linear_dpre_kernel<<<dim3((n + 127) / 128, gy), 128, 0, st>>>(y, dy, dpre_ws, db, m, n, act, slope, y_dtype);
  DCV_LAUNCH_CHECK("linear_dpre_kernel");
  if (dx) {  // dx[m][k] = dpre[m][n] @ w[n][k]
    GemmArgs g{dpre_ws, w, dx, nullptr, m, k, n, n, 1, k, 1, DCV_F32, DCV_F32, x_dtype, DCV_ACT_NONE, 0.f};
    if (launch_gemm(g, st)) return 2;
  }
  if (dw) {  // dw[n][k] = dpre^T[n][m] @ x[m][k]
    GemmArgs g{dpre_ws, x, dw, nullptr, n, k, m, 1, n, k, 1, DCV_F32, x_dtype, DCV_F32, DCV_ACT_NONE, 0.f};
    if (launch_gemm(g, st)) return 2;
  }
  return 0;
}

}  // extern "C"
