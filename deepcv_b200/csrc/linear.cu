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

// TM x TM register block per thread, 16 x 16 threads: 64 x 64 tiles (TM = 4) for large problems, 32 x 32 (TM = 2) when the large tile would leave
// most SMs without a CTA (the 256 x 1000 x 768 classifier head: 64 CTAs on 148 SMs, 159 us). The next k-slice is fetched into registers while the
// current one is multiplied out of shared memory: the loop used to expose one global-load round trip per 16 columns (85 us for 0.4 GFLOP).
template <int TM>
__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs g) {
  constexpr int BM = 16 * TM, BN = 16 * TM, BK = 32, PER = BM * BK / 256;
  __shared__ float sA[BK][BM + 4], sB[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  float acc[TM][TM] = {};
  float ra[PER], rb[PER];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int i = tid + e * 256;
      // walk the contiguous dimension of each operand with consecutive threads
      int mi, ki;
      if (g.sal == 1) { ki = i % BK; mi = i / BK; } else { mi = i % BM; ki = i / BM; }
      const int m = m0 + mi, k = k0 + ki;
      ra[e] = (m < g.M && k < g.K) ? ld_any(g.A, g.a_dtype, (size_t)((long long)m * g.sai + (long long)k * g.sal)) : 0.f;
      int ni, kj;
      if (g.sbl == 1) { kj = i % BK; ni = i / BK; } else { ni = i % BN; kj = i / BN; }
      const int n = n0 + ni, kk = k0 + kj;
      rb[e] = (n < g.N && kk < g.K) ? ld_any(g.B, g.b_dtype, (size_t)((long long)kk * g.sbl + (long long)n * g.sbj)) : 0.f;
    }
  };
  auto stage = [&]() {
#pragma unroll
    for (int e = 0; e < PER; ++e) {
      const int i = tid + e * 256;
      int mi, ki;
      if (g.sal == 1) { ki = i % BK; mi = i / BK; } else { mi = i % BM; ki = i / BM; }
      sA[ki][mi] = ra[e];
      int ni, kj;
      if (g.sbl == 1) { kj = i % BK; ni = i / BK; } else { ni = i % BN; kj = i / BN; }
      sB[kj][ni] = rb[e];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    stage();
    __syncthreads();
    if (k0 + BK < g.K) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) { a[i] = sA[kk][ty * TM + i]; b[i] = sB[kk][tx * TM + i]; }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) {
      const int m = m0 + ty * TM + i, n = n0 + tx * TM + j;
      if (m < g.M && n < g.N) {
        float v = acc[i][j] + (g.bias ? g.bias[n] : 0.f);
        st_any(g.C, g.c_dtype, (size_t)m * g.N + n, act_apply(v, g.act, g.slope));
      }
    }
}

static int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  const long long big = (long long)((g.N + 63) / 64) * ((g.M + 63) / 64);
  if (big >= 2 * kNumSMs) {
    dim3 grid((g.N + 63) / 64, (g.M + 63) / 64);
    gemm_kernel<4><<<grid, 256, 0, st>>>(g);
  } else {
    dim3 grid((g.N + 31) / 32, (g.M + 31) / 32);
    gemm_kernel<2><<<grid, 256, 0, st>>>(g);
  }
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


// ---- "skinny" head: n <= 32 output features (the 1280 -> 10 classifier of the default net). The generic 64x64 tile above would launch a
// handful of CTAs; here a warp owns a row of x and keeps all n accumulators in registers (forward / dx), and the weight gradient is a
// split-M reduction with one thread per input feature.
template <typename TX> __device__ __forceinline__ float ldx(const TX* p, size_t i) { return to_f<TX>(p[i]); }

template <int NB, typename TX>
__global__ void __launch_bounds__(256) linear_skinny_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                                                                int m, int n, int k, int act, float slope) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= m) return;
  float acc[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) acc[j] = 0.f;
  const TX* xr = x + (size_t)row * k;
  int kk = lane;
  for (; kk + 96 < k; kk += 128) {   // 4 independent k positions per lane and iteration: loads of all four are in flight together
    float xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xv[u] = ldx<TX>(xr, kk + 32 * u);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (j < n) {
        float wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = __ldg(w + (size_t)j * k + kk + 32 * u);
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[j] = fmaf(xv[u], wv[u], acc[j]);
      }
    }
  }
  for (; kk < k; kk += 32) {
    const float xv = ldx<TX>(xr, kk);
#pragma unroll
    for (int j = 0; j < NB; ++j) if (j < n) acc[j] = fmaf(xv, w[(size_t)j * k + kk], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) acc[j] = warp_sum(acc[j]);
  if (lane < n) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) if (j == lane) v = acc[j];
    v += bias ? bias[lane] : 0.f;
    y[(size_t)row * n + lane] = act_apply(v, act, slope);
  }
}

// dpre[row][j] = act'(y)*dy (written to dpre_ws) and dx[row][:] = dpre[row][:] @ w
template <int NB, typename TX>
__global__ void __launch_bounds__(256) linear_skinny_dx_kernel(const float* __restrict__ w, const float* __restrict__ y, const float* __restrict__ dy, TX* __restrict__ dx,
                                                               float* __restrict__ dpre_ws, int m, int n, int k, int act, float slope) {
  const int lane = threadIdx.x & 31, row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= m) return;
  float mine = 0.f;
  if (lane < n) {
    const size_t i = (size_t)row * n + lane;
    mine = dy[i] * act_grad_from_output(y[i], act, slope);
    dpre_ws[i] = mine;
  }
  float d[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) d[j] = __shfl_sync(0xffffffffu, mine, j);
  if (!dx) return;
  TX* dxr = dx + (size_t)row * k;
  int kk = lane;
  for (; kk + 96 < k; kk += 128) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (j < n) {
        float wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = __ldg(w + (size_t)j * k + kk + 32 * u);
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = fmaf(d[j], wv[u], v[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) dxr[kk + 32 * u] = from_f<TX>(v[u]);
  }
  for (; kk < k; kk += 32) {
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < NB; ++j) if (j < n) v = fmaf(d[j], w[(size_t)j * k + kk], v);
    dxr[kk] = from_f<TX>(v);
  }
}

// dw[j][kk] += sum over a chunk of rows of dpre[r][j] * x[r][kk]; db[j] += sum dpre[r][j]   (dw, db zeroed by the wrapper)
template <int NB, typename TX>
__global__ void __launch_bounds__(256) linear_skinny_dw_kernel(const TX* __restrict__ x, const float* __restrict__ dpre, float* __restrict__ dw, float* __restrict__ db,
                                                               int m, int n, int k, int rows_per_cta) {
  __shared__ float s_d[64 * NB];
  const int kk = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(r0 + rows_per_cta, m);
  for (int i = threadIdx.x; i < (r1 - r0) * n; i += blockDim.x) s_d[(i / n) * NB + i % n] = dpre[(size_t)r0 * n + i];
  __syncthreads();
  float acc[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) acc[j] = 0.f;
  if (kk < k) {
    for (int r = r0; r < r1; ++r) {
      const float xv = ldx<TX>(x, (size_t)r * k + kk);
#pragma unroll
      for (int j = 0; j < NB; ++j) if (j < n) acc[j] = fmaf(s_d[(r - r0) * NB + j], xv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) if (j < n) atomicAdd(dw + (size_t)j * k + kk, acc[j]);
  }
  if (db && blockIdx.x == 0 && threadIdx.x < n) {
    float sd = 0.f;
    for (int r = r0; r < r1; ++r) sd += s_d[(r - r0) * NB + threadIdx.x];
    atomicAdd(db + threadIdx.x, sd);
  }
}

template <int NB, typename TX>
static int skinny_fwd(const void* x, const float* w, const float* bias, float* y, int m, int n, int k, int act, float slope, cudaStream_t st) {
  linear_skinny_fwd_kernel<NB, TX><<<(m + 3) / 4, 128, 0, st>>>((const TX*)x, w, bias, y, m, n, k, act, slope);
  DCV_LAUNCH_CHECK("linear_skinny_fwd_kernel");
  return 0;
}

template <int NB, typename TX>
static int skinny_bwd(const void* x, const float* w, const float* y, const float* dy, void* dx, float* dw, float* db, float* dpre, int m, int n, int k, int act, float slope, bool prezeroed, cudaStream_t st) {
  linear_skinny_dx_kernel<NB, TX><<<(m + 3) / 4, 128, 0, st>>>(w, y, dy, (TX*)dx, dpre, m, n, k, act, slope);
  DCV_LAUNCH_CHECK("linear_skinny_dx_kernel");
  if (dw) {
    zero_accumulator(dw, (size_t)n * k * sizeof(float), st, prezeroed);
    const int rows = 32;
    linear_skinny_dw_kernel<NB, TX><<<dim3((k + 255) / 256, (m + rows - 1) / rows), 256, 0, st>>>((const TX*)x, dpre, dw, db, m, n, k, rows);
    DCV_LAUNCH_CHECK("linear_skinny_dw_kernel");
  }
  return 0;
}

}  // namespace dcv

extern "C" {

int dcv_linear_fwd(const void* x, const float* w, const float* bias, void* y, int m, int n, int k, int act, float slope,
                   int x_dtype, int y_dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(x && w && y && m > 0 && n > 0 && k > 0, "linear_fwd: bad arguments");
  if (n <= 32 && y_dtype == DCV_F32) {
    cudaStream_t st = as_stream(stream);
    DCV_DISPATCH_DTYPE(x_dtype, TX, {
      if (n <= 8) return skinny_fwd<8, TX>(x, w, bias, (float*)y, m, n, k, act, slope, st);
      if (n <= 16) return skinny_fwd<16, TX>(x, w, bias, (float*)y, m, n, k, act, slope, st);
      return skinny_fwd<32, TX>(x, w, bias, (float*)y, m, n, k, act, slope, st);
    });
  }
  GemmArgs g{x, w, y, bias, m, n, k, k, 1, 1, k, x_dtype, DCV_F32, y_dtype, act, slope};
  return launch_gemm(g, as_stream(stream));
}

int dcv_linear_bwd(const void* x, const float* w, const void* y, const void* dy, void* dx, float* dw, float* db, float* dpre_ws,
                   int m, int n, int k, int act, float slope, int x_dtype, int y_dtype, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(x && w && y && dy && dpre_ws && m > 0 && n > 0 && k > 0, "linear_bwd: bad arguments");
  cudaStream_t st = as_stream(stream);
  zero_accumulator(db, (size_t)n * sizeof(float), st, acc_prezeroed != 0);
  if (n <= 32 && y_dtype == DCV_F32) {
    DCV_DISPATCH_DTYPE(x_dtype, TX, {
      if (n <= 8) return skinny_bwd<8, TX>(x, w, (const float*)y, (const float*)dy, dx, dw, db, dpre_ws, m, n, k, act, slope, acc_prezeroed != 0, st);
      if (n <= 16) return skinny_bwd<16, TX>(x, w, (const float*)y, (const float*)dy, dx, dw, db, dpre_ws, m, n, k, act, slope, acc_prezeroed != 0, st);
      return skinny_bwd<32, TX>(x, w, (const float*)y, (const float*)dy, dx, dw, db, dpre_ws, m, n, k, act, slope, acc_prezeroed != 0, st);
    });
  }
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
