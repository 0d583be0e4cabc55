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


// ---- "skinny" head: n <= 32 output features (the 1280 -> 10 classifier of the default net) --------------------------------------------------------
// 13 MFLOP: pure latency. One CTA per row of x for forward / dx (512 CTAs of 4 warps: every SM holds several rows, all loads of a thread are
// independent), 4 rows per CTA for the weight gradient; the row(s) of x are staged in shared memory with 16-byte loads first.
// `torch.nn.Flatten` fused in (x_c > 0): x is then the NHWC image tensor [m][hw][x_c] itself and feature f = c * hw + p (the (C, H, W) order Flatten
// gives the logical N x C x H x W tensor, reference conf/base/parameters.yml:87) is element p * x_c + c of its row — the kernels walk f (contiguous in
// the weight matrix) and read / write x through that index map in shared memory: no transposed copy of the activations in either direction.
template <typename TX> __device__ __forceinline__ float ldx(const TX* p, size_t i) { return to_f<TX>(p[i]); }
constexpr int kSkinnyThreads = 128;

struct FlatMap {   // feature index (weight order) -> element of the row of x
  int x_c, hw;
  __device__ __forceinline__ int operator()(int f) const {
    if (x_c <= 0) return f;
    const int c = f / hw, p = f - c * hw;
    return p * x_c + c;
  }
};

// Row `row` of x (k elements) -> shared memory as fp32, in x's own element order. 16-byte loads when the row is aligned.
template <typename TX> __device__ __forceinline__ void stage_row(const TX* __restrict__ xr, float* xs, int k, int tid, int nthr) {
  constexpr int V = 16 / sizeof(TX);
  if ((reinterpret_cast<uintptr_t>(xr) & 15) == 0 && k % V == 0) {
    for (int v = tid; v < k / V; v += nthr) {
      float f[V];
      vec_unpack<TX>(*reinterpret_cast<const uint4*>(xr + (size_t)v * V), f);
#pragma unroll
      for (int i = 0; i < V; ++i) xs[v * V + i] = f[i];
    }
  } else {
    for (int i = tid; i < k; i += nthr) xs[i] = ldx<TX>(xr, i);
  }
}

template <int NB, typename TX>
__global__ void __launch_bounds__(kSkinnyThreads) linear_skinny_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                                                                           int m, int n, int k, int act, float slope, FlatMap map) {
  extern __shared__ float sh[];   // [k] row of x, then [warps][NB] partial sums
  float* xs = sh;
  float* part = sh + k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, row = blockIdx.x;
  stage_row<TX>(x + (size_t)row * k, xs, k, tid, kSkinnyThreads);
  __syncthreads();
  float acc[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) acc[j] = 0.f;
  for (int f0 = tid; f0 < k; f0 += 4 * kSkinnyThreads) {   // 4 independent feature positions per thread and iteration
    float xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int f = f0 + u * kSkinnyThreads; xv[u] = f < k ? xs[map(f)] : 0.f; }
    // ALL of the iteration's weight loads first (4 * n independent L2 requests in flight per thread), then the FMAs: issued per output feature they
    // were 30 dependent L2 round trips per thread (ncu: 78 % of the stall samples on the first FMA of each feature)
    float wv[NB][4];
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int f = f0 + u * kSkinnyThreads; wv[j][u] = (j < n && f < k) ? __ldg(w + (size_t)j * k + f) : 0.f; }
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[j] = fmaf(xv[u], wv[j][u], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    acc[j] = warp_sum(acc[j]);
    if (lane == 0) part[warp * NB + j] = acc[j];
  }
  __syncthreads();
  if (tid < n) {
    float v = bias ? bias[tid] : 0.f;
    for (int wq = 0; wq < kSkinnyThreads / 32; ++wq) v += part[wq * NB + tid];
    y[(size_t)row * n + tid] = act_apply(v, act, slope);
  }
}

// dpre[row][j] = act'(y)*dy (written to dpre_ws) and dx[row][:] = dpre[row][:] @ w
template <int NB, typename TX>
__global__ void __launch_bounds__(kSkinnyThreads) linear_skinny_dx_kernel(const float* __restrict__ w, const float* __restrict__ y, const float* __restrict__ dy, TX* __restrict__ dx,
                                                                          float* __restrict__ dpre_ws, int m, int n, int k, int act, float slope, FlatMap map) {
  extern __shared__ float sh[];   // [k] row of dx in x's element order, then [NB] dpre of the row
  float* dxs = sh;
  float* d = sh + k;
  const int tid = threadIdx.x, row = blockIdx.x;
  if (tid < NB) {
    float mine = 0.f;
    if (tid < n) {
      const size_t i = (size_t)row * n + tid;
      mine = dy[i] * act_grad_from_output(y[i], act, slope);
      dpre_ws[i] = mine;
    }
    d[tid] = mine;
  }
  __syncthreads();
  if (!dx) return;
  float dj[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) dj[j] = d[j];
  for (int f0 = tid; f0 < k; f0 += 4 * kSkinnyThreads) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    float wv[NB][4];   // all loads of the iteration in flight together (see the forward kernel)
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int f = f0 + u * kSkinnyThreads; wv[j][u] = (j < n && f < k) ? __ldg(w + (size_t)j * k + f) : 0.f; }
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = fmaf(dj[j], wv[j][u], v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int f = f0 + u * kSkinnyThreads; if (f < k) dxs[map(f)] = v[u]; }
  }
  __syncthreads();
  TX* dxr = dx + (size_t)row * k;
  constexpr int V = 16 / sizeof(TX);
  if ((reinterpret_cast<uintptr_t>(dxr) & 15) == 0 && k % V == 0) {
    for (int v = tid; v < k / V; v += kSkinnyThreads) *reinterpret_cast<uint4*>(dxr + (size_t)v * V) = vec_pack<TX>(dxs + v * V);
  } else {
    for (int i = tid; i < k; i += kSkinnyThreads) dxr[i] = from_f<TX>(dxs[i]);
  }
}

// dw[j][f] += sum over the CTA's rows of dpre[r][j] * x[r][f]; db[j] += sum dpre[r][j]   (dw, db zeroed by the wrapper / the caller)
constexpr int kDwRows = 4, kDwThreads = 256, kDwPer = 5;   // features per thread: covers k <= 1280 in one CTA column, more through blockIdx.x
template <int NB, typename TX>
__global__ void __launch_bounds__(kDwThreads) linear_skinny_dw_kernel(const TX* __restrict__ x, const float* __restrict__ dpre, float* __restrict__ dw, float* __restrict__ db,
                                                                      int m, int n, int k, FlatMap map) {
  extern __shared__ float sh[];   // [kDwRows][k] rows of x, then [kDwRows][NB] dpre
  float* xs = sh;
  float* s_d = sh + (size_t)kDwRows * k;
  const int tid = threadIdx.x;
  const int r0 = blockIdx.y * kDwRows, rows = min(kDwRows, m - r0);
  for (int r = 0; r < rows; ++r) stage_row<TX>(x + (size_t)(r0 + r) * k, xs + (size_t)r * k, k, tid, kDwThreads);
  for (int i = tid; i < kDwRows * NB; i += kDwThreads) { const int r = i / NB, j = i - r * NB; s_d[i] = (r < rows && j < n) ? dpre[(size_t)(r0 + r) * n + j] : 0.f; }
  __syncthreads();
  const int fbase = blockIdx.x * kDwThreads * kDwPer + tid;
  float acc[kDwPer][NB];
#pragma unroll
  for (int i = 0; i < kDwPer; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) acc[i][j] = 0.f;
  int e[kDwPer];
#pragma unroll
  for (int i = 0; i < kDwPer; ++i) { const int f = fbase + i * kDwThreads; e[i] = f < k ? map(f) : 0; }
  for (int r = 0; r < rows; ++r) {
    float xv[kDwPer];
#pragma unroll
    for (int i = 0; i < kDwPer; ++i) xv[i] = xs[(size_t)r * k + e[i]];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const float dv = s_d[r * NB + j];
#pragma unroll
      for (int i = 0; i < kDwPer; ++i) acc[i][j] = fmaf(dv, xv[i], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < kDwPer; ++i) {
    const int f = fbase + i * kDwThreads;
    if (f < k) {
#pragma unroll
      for (int j = 0; j < NB; ++j) if (j < n) atomicAdd(dw + (size_t)j * k + f, acc[i][j]);
    }
  }
  if (db && blockIdx.x == 0 && tid < n) {
    float sd = 0.f;
    for (int r = 0; r < rows; ++r) sd += s_d[r * NB + tid];
    atomicAdd(db + tid, sd);
  }
}

static bool skinny_fits(int n, int k) { return n <= 32 && (size_t)kDwRows * k * 4 + 1024 <= 160 * 1024; }

template <int NB, typename TX>
static int skinny_fwd(const void* x, const float* w, const float* bias, float* y, int m, int n, int k, int act, float slope, FlatMap map, cudaStream_t st) {
  const size_t smem = ((size_t)k + (kSkinnyThreads / 32) * NB) * 4;
  auto kern = linear_skinny_fwd_kernel<NB, TX>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<m, kSkinnyThreads, smem, st>>>((const TX*)x, w, bias, y, m, n, k, act, slope, map);
  DCV_LAUNCH_CHECK("linear_skinny_fwd_kernel");
  return 0;
}

template <int NB, typename TX>
static int skinny_bwd(const void* x, const float* w, const float* y, const float* dy, void* dx, float* dw, float* db, float* dpre, int m, int n, int k, int act, float slope,
                      bool prezeroed, FlatMap map, cudaStream_t st) {
  {
    const size_t smem = ((size_t)k + NB) * 4;
    auto kern = linear_skinny_dx_kernel<NB, TX>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<m, kSkinnyThreads, smem, st>>>(w, y, dy, (TX*)dx, dpre, m, n, k, act, slope, map);
    DCV_LAUNCH_CHECK("linear_skinny_dx_kernel");
  }
  if (dw || db) {
    zero_accumulator(dw, (size_t)n * k * sizeof(float), st, prezeroed);
    const size_t smem = ((size_t)kDwRows * k + kDwRows * NB) * 4;
    auto kern = linear_skinny_dw_kernel<NB, TX>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (dw) {
      kern<<<dim3((k + kDwThreads * kDwPer - 1) / (kDwThreads * kDwPer), (m + kDwRows - 1) / kDwRows), kDwThreads, smem, st>>>((const TX*)x, dpre, dw, db, m, n, k, map);
      DCV_LAUNCH_CHECK("linear_skinny_dw_kernel");
    } else {   // frozen weight, trainable bias: the bias gradient alone
      int gy = (m + 31) / 32; if (gy > 64) gy = 64;
      linear_dpre_kernel<<<dim3((n + 127) / 128, gy), 128, 0, st>>>(y, dy, dpre, db, m, n, act, slope, DCV_F32);
      DCV_LAUNCH_CHECK("linear_dpre_kernel");
    }
  }
  return 0;
}

}  // namespace dcv

extern "C" {

int dcv_linear_flatten_fused(int n, int k) { return dcv::skinny_fits(n, k) ? 1 : 0; }

int dcv_linear_fwd(const void* x, const float* w, const float* bias, void* y, int m, int n, int k, int act, float slope,
                   int x_dtype, int y_dtype, int x_nhwc_channels, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(x && w && y && m > 0 && n > 0 && k > 0, "linear_fwd: bad arguments");
  DCV_REQUIRE(x_nhwc_channels == 0 || (x_nhwc_channels > 0 && k % x_nhwc_channels == 0 && skinny_fits(n, k) && y_dtype == DCV_F32),
              "linear_fwd: fused Flatten (x_nhwc_channels = %d) needs n <= 32 output features and k a multiple of the channel count (see dcv_linear_flatten_fused)", x_nhwc_channels);
  const FlatMap map{x_nhwc_channels, x_nhwc_channels > 0 ? k / x_nhwc_channels : 0};
  if (skinny_fits(n, k) && y_dtype == DCV_F32) {
    cudaStream_t st = as_stream(stream);
    DCV_DISPATCH_DTYPE(x_dtype, TX, {
      if (n <= 8) return skinny_fwd<8, TX>(x, w, bias, (float*)y, m, n, k, act, slope, map, st);
      if (n <= 16) return skinny_fwd<16, TX>(x, w, bias, (float*)y, m, n, k, act, slope, map, st);
      return skinny_fwd<32, TX>(x, w, bias, (float*)y, m, n, k, act, slope, map, st);
    });
  }
  GemmArgs g{x, w, y, bias, m, n, k, k, 1, 1, k, x_dtype, DCV_F32, y_dtype, act, slope};
  return launch_gemm(g, as_stream(stream));
}

int dcv_linear_bwd(const void* x, const float* w, const void* y, const void* dy, void* dx, float* dw, float* db, float* dpre_ws,
                   int m, int n, int k, int act, float slope, int x_dtype, int y_dtype, int acc_prezeroed, int x_nhwc_channels, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(x && w && y && dy && dpre_ws && m > 0 && n > 0 && k > 0, "linear_bwd: bad arguments");
  DCV_REQUIRE(x_nhwc_channels == 0 || (x_nhwc_channels > 0 && k % x_nhwc_channels == 0 && skinny_fits(n, k) && y_dtype == DCV_F32),
              "linear_bwd: fused Flatten (x_nhwc_channels = %d) needs n <= 32 output features and k a multiple of the channel count", x_nhwc_channels);
  const FlatMap map{x_nhwc_channels, x_nhwc_channels > 0 ? k / x_nhwc_channels : 0};
  cudaStream_t st = as_stream(stream);
  zero_accumulator(db, (size_t)n * sizeof(float), st, acc_prezeroed != 0);
  if (skinny_fits(n, k) && y_dtype == DCV_F32) {
    DCV_DISPATCH_DTYPE(x_dtype, TX, {
      if (n <= 8) return skinny_bwd<8, TX>(x, w, (const float*)y, (const float*)dy, dx, dw, db, dpre_ws, m, n, k, act, slope, acc_prezeroed != 0, map, st);
      if (n <= 16) return skinny_bwd<16, TX>(x, w, (const float*)y, (const float*)dy, dx, dw, db, dpre_ws, m, n, k, act, slope, acc_prezeroed != 0, map, st);
      return skinny_bwd<32, TX>(x, w, (const float*)y, (const float*)dy, dx, dw, db, dpre_ws, m, n, k, act, slope, acc_prezeroed != 0, map, st);
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
