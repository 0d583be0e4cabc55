// BatchNorm / GroupNorm after the activation, restated as ONE affine per (image, channel).
//
// Forward:   y = act(conv(x))  --(sum y, sum y^2 per (n,c))-->  finalize  -->  z = A[n][c]*y + B[n][c]
// Backward:  (sum dz, sum dz*y per (n,c))  -->  finalize  -->  dy = act'(y) * (P[n][c]*dz + Q[n][c]*y + R[n][c])
//
// Why this is exact: BatchNorm is a per-channel affine u = alpha_c*y + beta_c whose coefficients depend only on
// per-channel sums; GroupNorm applied to u is an affine per (n, group) whose mean / variance over u are closed-form in
// the per-(n,c) sums of y and y^2. The same holds for the adjoint: every reduction BatchNorm-backward and
// GroupNorm-backward need is bilinear in (dz, y) per (n,c). So the tensor is read once for statistics (or not at all
// when the convolution epilogue produced them) and once for the apply, instead of the >= 4 passes per block of the
// module-by-module reference path (meta/nn.py:553). The finalize kernels do the small per-(n,c) algebra in fp64.
//
// HBM roofline: apply fwd = 2*E*s bytes, stats = E*s, bwd reduce = 2*E*s, bwd apply = 3*E*s  (E elements, s bytes each).
#include "common.cuh"

namespace dcv {

// Work decomposition shared by the streaming kernels. A tensor is n images of hw pixels of c channels (NHWC). It is walked as "vector rows": one row
// = `span` contiguous elements = `cv` 16-byte vectors (scalars when the channel count / alignment does not allow vectors). Normally a row is one pixel
// (span = c); when c is smaller than a vector and divides it (the 4-channel CIFAR layers in bf16) a row packs `pack` pixels (span = c * pack = one
// vector) and element e of the vector belongs to channel e % c. The n * hwv rows are cut into equal contiguous ranges, one per CTA, with the CTA count
// fitted to ONE resident wave (ncu on the first version: 768 CTAs on 740 slots = 1.04 waves, SMs idle 40 % of the kernel); a range may cross image
// boundaries, the per-(image, channel) state is flushed / reloaded there. A thread keeps ONE column for its whole life.
struct NcGeom {
  int n, c;
  int hwv;             // vector rows per image
  int span;            // elements per vector row (c * pack)
  int pack;            // pixels per vector row
  int cv;              // vectors per row (span / VE)
  int cols_per_block;  // <= 256
  int rows;            // threads per column
  uint32_t total_rows, rows_per_cta;
};

static int streaming_ctas_per_sm(const void* kernel) {
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0) != cudaSuccess || occ < 1) occ = 1;
  return occ > 4 ? 4 : occ;   // 4 x 256 threads x 4 loads in flight per SM already covers the HBM latency-bandwidth product
}

// VE = elements per vector (1 = scalar fallback). ctas_per_sm: resident CTAs of the kernel about to be launched.
template <int VE>
static NcGeom make_geom(int n, int hw, int c, int ctas_per_sm, dim3* grid, int* block, int min_rows_per_thread = 2) {
  NcGeom g;
  g.n = n; g.c = c; g.pack = 1;
  if (VE > 1 && c < VE && VE % c == 0 && hw % (VE / c) == 0) g.pack = VE / c;
  g.hwv = hw / g.pack; g.span = c * g.pack; g.cv = g.span / VE;
  g.cols_per_block = g.cv < 256 ? g.cv : 256;
  g.rows = 256 / g.cols_per_block;
  *block = g.cols_per_block * g.rows;
  const int col_blocks = (g.cv + g.cols_per_block - 1) / g.cols_per_block;
  g.total_rows = (uint32_t)n * (uint32_t)g.hwv;
  // one wave at most, and at least `min_rows_per_thread` rows per thread (small tensors: fewer CTAs; the kernel that ends in same-address atomics
  // asks for 8). A range of several images is cut at image boundaries, so that no CTA pays an extra flush for a partial image.
  long long ctas = (long long)kNumSMs * ctas_per_sm / col_blocks;
  const long long by_work = ((long long)g.total_rows + (long long)g.rows * min_rows_per_thread - 1) / ((long long)g.rows * min_rows_per_thread);
  if (ctas > by_work) ctas = by_work;
  if (ctas < 1) ctas = 1;
  uint32_t per = (uint32_t)(((long long)g.total_rows + ctas - 1) / ctas);
  if (per >= (uint32_t)g.hwv) per = per / g.hwv * g.hwv;
  else per = (per + g.rows - 1) / g.rows * g.rows;
  g.rows_per_cta = per;
  *grid = dim3((g.total_rows + per - 1) / per, 1, col_blocks);
  return g;
}

template <typename T, int VE> __device__ __forceinline__ void load_vec(const T* p, float* out) {
  if constexpr (VE == 1) out[0] = to_f<T>(*p);
  else vec_unpack<T>(*reinterpret_cast<const uint4*>(p), out);
}
template <typename T, int VE> __device__ __forceinline__ void store_vec(T* p, const float* in) {
  if constexpr (VE == 1) *p = from_f<T>(in[0]);
  else *reinterpret_cast<uint4*>(p) = vec_pack<T>(in);
}

// Calls f(img, p0, p1) for every image the CTA's row range touches ([p0, p1) = vector rows inside that image). CTA-uniform.
template <typename F>
__device__ __forceinline__ void for_each_segment(const NcGeom& g, F&& f) {
  uint32_t r0 = blockIdx.x * g.rows_per_cta;
  const uint32_t r1 = min(r0 + g.rows_per_cta, g.total_rows);
  while (r0 < r1) {
    const uint32_t img = r0 / (uint32_t)g.hwv;
    const int p0 = (int)(r0 - img * (uint32_t)g.hwv);
    const int p1 = (int)min((uint32_t)g.hwv, (uint32_t)p0 + (r1 - r0));
    f((int)img, p0, p1);
    r0 += (uint32_t)(p1 - p0);
  }
}

// channel of element e of column colg
__device__ __forceinline__ int chan_of(const NcGeom& g, int colg, int ve, int e) { return g.pack > 1 ? (e & (g.c - 1)) : colg * ve + e; }

// Sums `NV` per-thread values per channel over the CTA (over the `rows` threads that share a column and, in packed mode, over the elements of a
// vector that belong to the same channel), then ONE atomicAdd per (channel, v) into dst[channel * dst_stride + v]. Two stages so that all threads
// take part: the outputs x `parts` row slices are spread over the threads, then the slices are folded.
// (Earlier versions: the row-0 threads summing all rows serially = 512 dependent shared loads; one atomic per slice = 256 atomics per CTA onto 4
// addresses for the 4-channel layers, 105 us of L2 atomic serialisation.)
template <int VE, int NV>
__device__ __forceinline__ void column_reduce_atomic(float (*acc)[VE], const NcGeom& g, int col, int row, float* dst, int dst_stride) {
  constexpr int PER = VE * NV;
  static_assert(PER <= 16, "column_reduce_atomic: too many values per thread");
  constexpr int PST = PER + 1;           // odd per-thread stride: the 16 stores of a warp's 32 threads would otherwise hit 2 banks (16-way conflict)
  __shared__ float red[256 * PST];  // [row][col][NV*VE] flattened
  __shared__ float red2[256];
  const int t = row * g.cols_per_block + col;
  __syncthreads();   // previous use of red / red2
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < VE; ++e) red[t * PST + v * VE + e] = acc[v][e];
  __syncthreads();
  const int epc = g.pack > 1 ? g.c : VE;                  // distinct channels per column
  const int outs = g.cols_per_block * NV * epc, nthr = blockDim.x;
  const int parts = outs < nthr ? nthr / outs : 1;
  for (int o = threadIdx.x; o < outs * parts; o += nthr) {
    const int part = o / outs, oo = o - part * outs;
    const int cl = oo / (NV * epc), k = oo - cl * (NV * epc), v = k / epc, e0 = k - v * epc;
    float s = 0.f;
    for (int r = part; r < g.rows; r += parts)
      for (int e = e0; e < VE; e += epc) s += red[(r * g.cols_per_block + cl) * PST + v * VE + e];
    if (parts > 1) red2[o] = s;
    else {
      const int colg = blockIdx.z * g.cols_per_block + cl;
      if (colg < g.cv) atomicAdd(dst + (size_t)chan_of(g, colg, VE, e0) * dst_stride + v, s);
    }
  }
  if (parts > 1) {
    __syncthreads();
    if ((int)threadIdx.x < outs) {
      const int oo = threadIdx.x;
      const int cl = oo / (NV * epc), k = oo - cl * (NV * epc), v = k / epc, e0 = k - v * epc;
      float s = 0.f;
      for (int q = 0; q < parts; ++q) s += red2[q * outs + oo];
      const int colg = blockIdx.z * g.cols_per_block + cl;
      if (colg < g.cv) atomicAdd(dst + (size_t)chan_of(g, colg, VE, e0) * dst_stride + v, s);
    }
  }
}

// Raw (not yet unpacked) 16-byte vector or scalar element: lets the streaming loops below issue UNR independent loads before touching any of them.
template <typename T, int VE> struct Raw { uint4 v; };
template <typename T> struct Raw<T, 1> { T v; };
template <typename T, int VE> __device__ __forceinline__ Raw<T, VE> load_raw(const T* p) {
  Raw<T, VE> r;
  if constexpr (VE == 1) r.v = *p; else r.v = *reinterpret_cast<const uint4*>(p);
  return r;
}
template <typename T, int VE> __device__ __forceinline__ Raw<T, VE> zero_raw() {
  Raw<T, VE> r;
  if constexpr (VE == 1) r.v = from_f<T>(0.f); else r.v = make_uint4(0u, 0u, 0u, 0u);
  return r;
}
template <typename T, int VE> __device__ __forceinline__ void unpack_raw(const Raw<T, VE>& r, float* out) {
  if constexpr (VE == 1) out[0] = to_f<T>(r.v); else vec_unpack<T>(r.v, out);
}
#ifndef DCV_UNR
#define DCV_UNR 4
#endif
constexpr int UNR = DCV_UNR;   // rows in flight per thread: the bf16 kernels were latency-bound at ~45 % of HBM with one load per iteration

// ---- statistics: stats[n][c][2] += {sum y, sum y^2}
constexpr int SUNR = UNR;   // (8 rows in flight measured no better than 4, and slower once the tail loads were predicated)
template <typename T, int VE>
__global__ void __launch_bounds__(256) stats_kernel(const T* __restrict__ y, float* __restrict__ stats, const NcGeom g) {
  const int col = threadIdx.x % g.cols_per_block, row = threadIdx.x / g.cols_per_block;
  const int colg = blockIdx.z * g.cols_per_block + col;
  for_each_segment(g, [&](int img, int p0, int p1) {
    float acc[2][VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[0][e] = acc[1][e] = 0.f;
    if (colg < g.cv) {
      const T* base = y + ((size_t)img * g.hwv) * g.span + (size_t)colg * VE;
      // SUNR rows in flight per thread; the loads of a short range / the tail are predicated, not serialised (a thread with 6 rows used to issue 6
      // dependent loads: 15 us for the 12.8 MB 7x7 tensors)
      for (int p = p0 + row; p < p1; p += SUNR * g.rows) {
        Raw<T, VE> r[SUNR];
#pragma unroll
        for (int u = 0; u < SUNR; ++u) r[u] = (p + u * g.rows < p1) ? load_raw<T, VE>(base + (size_t)(p + u * g.rows) * g.span) : zero_raw<T, VE>();
#pragma unroll
        for (int u = 0; u < SUNR; ++u) {
          float v[VE];
          unpack_raw<T, VE>(r[u], v);
#pragma unroll
          for (int e = 0; e < VE; ++e) { acc[0][e] += v[e]; acc[1][e] = fmaf(v[e], v[e], acc[1][e]); }
        }
      }
    }
    column_reduce_atomic<VE, 2>(acc, g, col, row, stats + (size_t)img * g.c * 2, 2);
  });
}

// Forward counterpart of BnFold: a BatchNorm-only block whose batch is ONE image — the consumer of the raw output (apply / apply + residual sum / apply + pool)
// computes alpha = gamma * rstd, beta = bias - mean * alpha from the channel totals {sum y, sum y^2} itself; the CTAs of the first row range also write what
// the backward needs (`saved`: mean, rstd | alpha, beta) and update the running statistics. Same arithmetic as fwd_finalize_kernel, minus its launch.
struct BnFwdFold {
  const float* stats;   // [c][2]
  const float* gamma; const float* beta;
  float* run_mean; float* run_var; long long* nbt;
  float* saved;
  double m, eps, momentum;
  int c, training;
};

__device__ __forceinline__ void bn_fwd_coeffs(const BnFwdFold& f, int ch, bool owner, float& A, float& B) {
  const double gamma = f.gamma ? (double)__ldg(f.gamma + ch) : 1.0, bias = f.beta ? (double)__ldg(f.beta + ch) : 0.0;
  const bool run = f.run_mean && f.run_var;
  double mean, var;
  if (f.training) {
    mean = (double)__ldg(f.stats + 2 * ch) / f.m;
    var = (double)__ldg(f.stats + 2 * ch + 1) / f.m - mean * mean;
    if (var < 0.0) var = 0.0;
    if (owner && run) {
      const double unbiased = f.m > 1.0 ? var * f.m / (f.m - 1.0) : var;
      f.run_mean[ch] = (float)((1.0 - f.momentum) * (double)f.run_mean[ch] + f.momentum * mean);
      f.run_var[ch] = (float)((1.0 - f.momentum) * (double)f.run_var[ch] + f.momentum * unbiased);
    }
  } else {
    mean = (double)f.run_mean[ch]; var = (double)f.run_var[ch];
  }
  // only the owner (one thread per channel) pays for the fp64 reciprocal square root; everybody else: fp32 rsqrt + one Newton step on the fp64 variance
  // (relative error ~1e-7 on alpha / beta: below the bf16 / fp32 rounding of the tensor they are applied to)
  if (owner) {
    const double rstd = rsqrt(var + f.eps), alpha = gamma * rstd, beta = bias - mean * alpha;
    f.saved[2 * ch] = (float)mean; f.saved[2 * ch + 1] = (float)rstd;
    f.saved[2 * f.c + 2 * ch] = (float)alpha; f.saved[2 * f.c + 2 * ch + 1] = (float)beta;
    A = (float)alpha; B = (float)beta;
  } else {
    const float v = (float)(var + f.eps);
    float r = rsqrtf(v);
    r = r * (1.5f - 0.5f * v * r * r);
    const float alpha = (float)gamma * r;
    A = alpha; B = (float)bias - (float)mean * alpha;
  }
}

// ---- forward apply: z = A*y + B
template <typename T, int VE, bool ADD = false, bool FOLD = false>
__global__ void __launch_bounds__(256) apply_fwd_kernel(const T* __restrict__ y, const float* __restrict__ ab, T* __restrict__ z, const NcGeom g, const T* __restrict__ other = nullptr,
                                                        const BnFwdFold fold = BnFwdFold()) {
  const int col = threadIdx.x % g.cols_per_block, row = threadIdx.x / g.cols_per_block;
  const int colg = blockIdx.z * g.cols_per_block + col;
  if constexpr (FOLD) {
    if (blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0 && fold.training && fold.nbt) *fold.nbt += 1;
  }
  if (colg >= g.cv) return;
  for_each_segment(g, [&](int img, int p0, int p1) {
    float A[VE], B[VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      if constexpr (FOLD) {
        bn_fwd_coeffs(fold, chan_of(g, colg, VE, e), blockIdx.x == 0 && row == 0 && (g.pack <= 1 || e < g.c), A[e], B[e]);
      } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(ab) + (size_t)img * g.c + chan_of(g, colg, VE, e));
        A[e] = t.x; B[e] = t.y;
      }
    }
    const size_t base = ((size_t)img * g.hwv) * g.span + (size_t)colg * VE;
    for (int p = p0 + row; p < p1; p += UNR * g.rows) {
      Raw<T, VE> r[UNR], o[ADD ? UNR : 1];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        r[u] = (p + u * g.rows < p1) ? load_raw<T, VE>(y + base + (size_t)(p + u * g.rows) * g.span) : zero_raw<T, VE>();
        if constexpr (ADD) o[u] = (p + u * g.rows < p1) ? load_raw<T, VE>(other + base + (size_t)(p + u * g.rows) * g.span) : zero_raw<T, VE>();
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        float v[VE];
        unpack_raw<T, VE>(r[u], v);
#pragma unroll
        for (int e = 0; e < VE; ++e) v[e] = fmaf(A[e], v[e], B[e]);
        if constexpr (ADD) {   // the residual link that follows the block (sum with a tensor of the same shape), in the same pass
          float w[VE];
          unpack_raw<T, VE>(o[u], w);
#pragma unroll
          for (int e = 0; e < VE; ++e) v[e] += w[e];
        }
        if (p + u * g.rows < p1) store_vec<T, VE>(z + base + (size_t)(p + u * g.rows) * g.span, v);
      }
    }
  });
}

// ---- forward apply fused with the 2x2 / stride-2 average pooling that follows the block: zp = A * avgpool(y) + B (= avgpool(A*y + B)); the normalised
// full-resolution tensor is never written. One thread per (pooled pixel, channel vector).
template <typename T, int VE, bool FOLD = false>
__global__ void __launch_bounds__(256) apply_pool_fwd_kernel(const T* __restrict__ y, const float* __restrict__ ab, T* __restrict__ zp, int h, int w, int c, uint32_t total,
                                                             const FastDiv div_cv, const FastDiv div_q, const FastDiv div_p, const BnFwdFold fold = BnFwdFold()) {
  const int cv = c / VE, q = w >> 1;
  extern __shared__ float2 sh_ab[];   // FOLD: the c coefficient pairs, computed by every CTA from the channel totals
  if constexpr (FOLD) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && fold.training && fold.nbt) *fold.nbt += 1;
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
      float A, B;
      bn_fwd_coeffs(fold, ch, blockIdx.x == 0, A, B);
      sh_ab[ch] = make_float2(A, B);
    }
    __syncthreads();
  }
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t t = div_cv.div(idx);
    const int cc = (int)(idx - t * cv);
    const uint32_t row = div_q.div(t);             // row = img * p + oy
    const int ox = (int)(t - row * q);
    const uint32_t img = div_p.div(row);
    const T* src = y + (((size_t)row * 2) * w + (size_t)ox * 2) * c + (size_t)cc * VE;   // (img*p + oy) * 2 == img*h + 2*oy
    float v0[VE], v1[VE], v2[VE], v3[VE], out[VE];
    load_vec<T, VE>(src, v0); load_vec<T, VE>(src + c, v1); load_vec<T, VE>(src + (size_t)w * c, v2); load_vec<T, VE>(src + (size_t)(w + 1) * c, v3);
    const float2* abp = FOLD ? sh_ab + (size_t)cc * VE : reinterpret_cast<const float2*>(ab) + (size_t)img * c + (size_t)cc * VE;
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      const float2 k = FOLD ? abp[e] : __ldg(abp + e);
      out[e] = fmaf(k.x, ((v0[e] + v1[e]) + (v2[e] + v3[e])) * 0.25f, k.y);
    }
    store_vec<T, VE>(zp + (size_t)idx * VE, out);
  }
}

// ---- backward reduce: s[n][c][2] += {sum dz, sum dz*y}
// (Tried and dropped: three more activation-mask sums here so that the finalize kernel could give the convolution-bias gradient in closed form and the
// apply pass would need no reduction. Measured on B200: the reduce went from 42.7 to 60.5 us on the 64-channel 56x56 tensor — 40 accumulators per thread,
// ALU-bound — while the apply only gained 65.2 -> 62.6 us.)
constexpr int kBwdSums = 2;
// POOL (both backward kernels): `dz` is the gradient of the 2x2 / stride-2 average pooling that consumed the normalised tensor, at POOLED resolution; the
// gradient of pixel p is a quarter of the pooled pixel's (exact in bf16: the value the stand-alone pooling backward would have stored), read in place —
// the full-resolution dz is never written. p counts pixels over whole rows of width w (h even: image boundaries fall on even rows).
__device__ __forceinline__ uint32_t pooled_pixel(uint32_t p, const FastDiv& div_w, uint32_t w) {
  const uint32_t r = div_w.div(p), x = p - r * w;
  return (r >> 1) * (w >> 1) + (x >> 1);
}

template <typename T, int VE, bool POOL = false>
__global__ void __launch_bounds__(256) bwd_reduce_kernel(const T* __restrict__ dz, const T* __restrict__ y, float* __restrict__ s, const NcGeom g, const FastDiv div_w = FastDiv(), uint32_t w = 0) {
  const int col = threadIdx.x % g.cols_per_block, row = threadIdx.x / g.cols_per_block;
  const int colg = blockIdx.z * g.cols_per_block + col;
  for_each_segment(g, [&](int img, int p0, int p1) {
    float acc[2][VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[0][e] = acc[1][e] = 0.f;
    if (colg < g.cv) {
      const size_t base = ((size_t)img * g.hwv) * g.span + (size_t)colg * VE, pbase = ((size_t)img * (g.hwv >> 2)) * g.span + (size_t)colg * VE;
      for (int p = p0 + row; p < p1; p += UNR * g.rows) {
        Raw<T, VE> ra[UNR], rb[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const bool on = p + u * g.rows < p1;
          if constexpr (POOL) ra[u] = on ? load_raw<T, VE>(dz + pbase + (size_t)pooled_pixel((uint32_t)(p + u * g.rows), div_w, w) * g.span) : zero_raw<T, VE>();
          else ra[u] = on ? load_raw<T, VE>(dz + base + (size_t)(p + u * g.rows) * g.span) : zero_raw<T, VE>();
          rb[u] = on ? load_raw<T, VE>(y + base + (size_t)(p + u * g.rows) * g.span) : zero_raw<T, VE>();
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          float a[VE], b[VE];
          unpack_raw<T, VE>(ra[u], a); unpack_raw<T, VE>(rb[u], b);
          if constexpr (POOL) {
#pragma unroll
            for (int e = 0; e < VE; ++e) a[e] *= 0.25f;
          }
#pragma unroll
          for (int e = 0; e < VE; ++e) { acc[0][e] += a[e]; acc[1][e] = fmaf(a[e], b[e], acc[1][e]); }
        }
      }
    }
    column_reduce_atomic<VE, 2>(acc, g, col, row, s + (size_t)img * g.c * kBwdSums, kBwdSums);
  });
}

// ---- backward apply: dy = act'(y) * (P*dz + Q*y + R); dbias[c] += sum dy. The activation is a template parameter: no per-element switch.
template <int ACT> __device__ __forceinline__ float act_grad_t(float y, float slope) {
  if (ACT == DCV_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (ACT == DCV_ACT_LEAKY_RELU) return y > 0.f ? 1.f : slope;
  if (ACT == DCV_ACT_SIGMOID) return y * (1.f - y);
  return 1.f;
}

// FOLD (BatchNorm-only block whose batch is handed over as ONE image): the coefficients P, Q, R come out of the prologue — a few fp64 operations per thread
// on the channel totals {sum dz, sum dz*y} (`fold.s`) and the forward's saved mean / rstd / alpha — and the CTAs of the first row range write the
// BatchNorm parameter gradients: the bwd_finalize launch in between (6-8 us of latency per layer) is gone. Same arithmetic as bwd_finalize_kernel.
struct BnFold {
  const float* s;       // [c][2] channel totals of the backward reduce
  const float* saved;   // forward finalize: [0,2c) mean,rstd | [2c,4c) alpha,beta
  float* d_w; float* d_b;
  double inv_m;         // 1 / (N*H*W)
  int c, training;
};

template <typename T, int VE, int ACT, bool POOL = false, bool FOLD = false>
__global__ void __launch_bounds__(256) bwd_apply_kernel(const T* __restrict__ dz, const T* __restrict__ y, const float* __restrict__ pqr,
                                                        T* __restrict__ dy, float* __restrict__ dbias, float slope, const NcGeom g, const FastDiv div_w = FastDiv(), uint32_t w = 0,
                                                        const BnFold fold = BnFold()) {
  const int col = threadIdx.x % g.cols_per_block, row = threadIdx.x / g.cols_per_block;
  const int colg = blockIdx.z * g.cols_per_block + col;
  float acc[1][VE];   // bias-gradient partial sums, kept over all the CTA's images
#pragma unroll
  for (int e = 0; e < VE; ++e) acc[0][e] = 0.f;
  if (colg < g.cv) {
    for_each_segment(g, [&](int img, int p0, int p1) {
      float P[VE], Q[VE], R[VE];
      if constexpr (FOLD) {
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const int ch = chan_of(g, colg, VE, e);
          const double mu = (double)__ldg(fold.saved + 2 * ch), rc = (double)__ldg(fold.saved + 2 * ch + 1), al = (double)__ldg(fold.saved + 2 * fold.c + 2 * ch);
          const double u1 = (double)__ldg(fold.s + kBwdSums * ch), u2 = rc * ((double)__ldg(fold.s + kBwdSums * ch + 1) - mu * u1);
          P[e] = (float)al;
          Q[e] = fold.training ? (float)(-al * rc * u2 * fold.inv_m) : 0.f;
          R[e] = fold.training ? (float)(al * (-u1 * fold.inv_m + mu * rc * u2 * fold.inv_m)) : 0.f;
          if (blockIdx.x == 0 && row == 0 && (g.pack <= 1 || e < g.c)) {
            if (fold.d_w) fold.d_w[ch] = (float)u2;
            if (fold.d_b) fold.d_b[ch] = (float)u1;
          }
        }
      } else if (pqr) {
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          const float* t = pqr + ((size_t)img * g.c + chan_of(g, colg, VE, e)) * 3;
          P[e] = __ldg(t); Q[e] = __ldg(t + 1); R[e] = __ldg(t + 2);
        }
      } else {
#pragma unroll
        for (int e = 0; e < VE; ++e) { P[e] = 1.f; Q[e] = 0.f; R[e] = 0.f; }
      }
      const size_t base = ((size_t)img * g.hwv) * g.span + (size_t)colg * VE, pbase = ((size_t)img * (g.hwv >> 2)) * g.span + (size_t)colg * VE;
      auto one = [&](float* a, const float* b, size_t off) {
#pragma unroll
        for (int e = 0; e < VE; ++e) {
          if constexpr (POOL) a[e] *= 0.25f;
          const float pre = fmaf(P[e], a[e], fmaf(Q[e], b[e], R[e]));
          a[e] = pre * act_grad_t<ACT>(b[e], slope);
          acc[0][e] += a[e];
        }
        store_vec<T, VE>(dy + off, a);
      };
      for (int p = p0 + row; p < p1; p += UNR * g.rows) {
        Raw<T, VE> ra[UNR], rb[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const bool on = p + u * g.rows < p1;
          if constexpr (POOL) ra[u] = on ? load_raw<T, VE>(dz + pbase + (size_t)pooled_pixel((uint32_t)(p + u * g.rows), div_w, w) * g.span) : zero_raw<T, VE>();
          else ra[u] = on ? load_raw<T, VE>(dz + base + (size_t)(p + u * g.rows) * g.span) : zero_raw<T, VE>();
          rb[u] = on ? load_raw<T, VE>(y + base + (size_t)(p + u * g.rows) * g.span) : zero_raw<T, VE>();
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          if (p + u * g.rows < p1) {
            float a[VE], b[VE];
            unpack_raw<T, VE>(ra[u], a); unpack_raw<T, VE>(rb[u], b);
            one(a, b, base + (size_t)(p + u * g.rows) * g.span);
          }
        }
      }
    });
  }
  if (dbias) column_reduce_atomic<VE, 1>(acc, g, col, row, dbias, 1);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- finalize kernels: one CTA, fp64 algebra on the tiny per-(n,c) arrays -------------------------------------------
// `saved` layout (floats): [0,2c) BN mean,rstd | [2c,4c) alpha,beta | [4c, 4c+2nG) GN mean,rstd | [.., +2nG) scratch A,B per
// (n,g) | [.., +2c) scratch U1,U2 per channel.
__device__ __forceinline__ size_t off_alpha(int c) { return 2 * (size_t)c; }
__device__ __forceinline__ size_t off_gn(int c) { return 4 * (size_t)c; }
__device__ __forceinline__ size_t off_ab(int c, int n, int G) { return 4 * (size_t)c + 2 * (size_t)n * G; }
__device__ __forceinline__ size_t off_u(int c, int n, int G) { return 4 * (size_t)c + 4 * (size_t)n * G; }

__global__ void __launch_bounds__(1024) fwd_finalize_kernel(const dcv_norm_params prm, const float* __restrict__ stats, float* __restrict__ ab, float* __restrict__ saved, const int cpb) {
  const int n = prm.n, c = prm.c, G = prm.use_gn ? prm.gn_groups : 1;
  const double hw = (double)prm.hw;
  const int tid = threadIdx.x, nt = blockDim.x;
  // this CTA owns channels [ch_lo, ch_hi): whole GroupNorm groups (cpb is a multiple of the group size), so every phase below is CTA-local
  const int ch_lo = blockIdx.x * cpb, ch_hi = min(c, ch_lo + cpb), cl = ch_hi - ch_lo;
  // 1. BatchNorm per channel: `wpc` warps per channel (all warps busy even with 4 channels), lanes stride over the images, fp64 warp reduction,
  //    partials combined through shared memory by one thread per channel
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  __shared__ double part[32][2];
  constexpr int kShCh = 64;
  __shared__ float sh_ab[kShCh][2];   // alpha, beta of this CTA's channels (when it has at most kShCh): phases 2 / 3 read them here, not from global
  const int wpc = max(1, nwarps / max(cl, 1)), cpp = nwarps / wpc;   // warps per channel, channels per pass
  for (int c0 = 0; c0 < cl; c0 += cpp) {
    const int chl = c0 + warp / wpc, sub = warp % wpc;
    const bool on = warp / wpc < cpp && chl < cl && prm.use_bn && prm.bn_training;
    if (on) {
      const int ch = ch_lo + chl;
      double s1 = 0.0, s2 = 0.0;
      for (int i = sub * 32 + lane; i < n; i += wpc * 32) { s1 += (double)stats[((size_t)i * c + ch) * 2]; s2 += (double)stats[((size_t)i * c + ch) * 2 + 1]; }
      s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
      if (lane == 0) { part[warp][0] = s1; part[warp][1] = s2; }
    }
    __syncthreads();
    if (warp < cpp && c0 + warp < cl) {   // warp w finishes channel c0 + w: its lanes fold the wpc partials
      const int ch = ch_lo + c0 + warp;
      double alpha = 1.0, beta = 0.0, mean = 0.0, rstd = 1.0;
      // parameters are fetched before the reduction so that their latency overlaps it
      const double gamma = (prm.use_bn && prm.bn_weight) ? (double)prm.bn_weight[ch] : 1.0;
      const double bias = (prm.use_bn && prm.bn_bias) ? (double)prm.bn_bias[ch] : 0.0;
      const bool run = prm.use_bn && prm.bn_running_mean && prm.bn_running_var;
      const double rm = run ? (double)prm.bn_running_mean[ch] : 0.0, rv = run ? (double)prm.bn_running_var[ch] : 1.0;
      if (prm.use_bn) {
        double var;
        if (prm.bn_training) {
          double s1 = lane < wpc ? part[warp * wpc + lane][0] : 0.0, s2 = lane < wpc ? part[warp * wpc + lane][1] : 0.0;
          s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
          const double m = (double)n * hw;
          mean = s1 / m;
          var = s2 / m - mean * mean;
          if (var < 0.0) var = 0.0;
          if (run && lane == 0) {
            double mom = (double)prm.bn_momentum;
            if (mom < 0.0) mom = 1.0 / (double)((prm.bn_num_batches_tracked ? *prm.bn_num_batches_tracked : 0) + 1);
            const double unbiased = m > 1.0 ? var * m / (m - 1.0) : var;
            prm.bn_running_mean[ch] = (float)((1.0 - mom) * rm + mom * mean);
            prm.bn_running_var[ch] = (float)((1.0 - mom) * rv + mom * unbiased);
          }
        } else {
          mean = rm;
          var = rv;
        }
        rstd = rsqrt(var + (double)prm.bn_eps);
        alpha = gamma * rstd;
        beta = bias - mean * alpha;
      }
      if (lane == 0) {
        saved[2 * ch] = (float)mean; saved[2 * ch + 1] = (float)rstd;
        saved[off_alpha(c) + 2 * ch] = (float)alpha; saved[off_alpha(c) + 2 * ch + 1] = (float)beta;
        if (cl <= kShCh) { sh_ab[c0 + warp][0] = (float)alpha; sh_ab[c0 + warp][1] = (float)beta; }   // same float rounding as the copy the backward reads
      }
    }
    __syncthreads();
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0 && prm.use_bn && prm.bn_training && prm.bn_num_batches_tracked) *prm.bn_num_batches_tracked += 1;
  const bool ab_in_smem = cl <= kShCh;
  auto alpha_of = [&](int ch) { return ab_in_smem ? (double)sh_ab[ch - ch_lo][0] : (double)saved[off_alpha(c) + 2 * ch]; };
  auto beta_of = [&](int ch) { return ab_in_smem ? (double)sh_ab[ch - ch_lo][1] : (double)saved[off_alpha(c) + 2 * ch + 1]; };
  const int cg = c / G;
  if (prm.use_gn && cg <= 16) {
    // 2 + 3 merged (small groups: the CIFAR layers have 1 or 4 channels per group): one thread per (image, group) computes the group statistics
    // and writes the combined affine of its channels — no second pass over global memory, no barrier in between
    const double mg = (double)cg * hw;
    const int g_lo = ch_lo / cg, gl = cl / cg;
    for (int j = tid; j < n * gl; j += nt) {
      const int img = j / gl, grp = g_lo + (j - img * gl);
      const size_t i = (size_t)img * G + grp;
      double su = 0.0, suu = 0.0;
      for (int k = 0; k < cg; ++k) {
        const int ch = grp * cg + k;
        const double al = alpha_of(ch), be = beta_of(ch);
        const double sy = (double)stats[((size_t)img * c + ch) * 2], syy = (double)stats[((size_t)img * c + ch) * 2 + 1];
        su += al * sy + hw * be;
        suu += al * al * syy + 2.0 * al * be * sy + hw * be * be;
      }
      const double meand = su / mg;
      double var = suu / mg - meand * meand;
      if (var < 0.0) var = 0.0;
      const float mean_f = (float)meand, rstd_f = (float)rsqrt(var + (double)prm.gn_eps);
      saved[off_gn(c) + 2 * (size_t)i] = mean_f;
      saved[off_gn(c) + 2 * (size_t)i + 1] = rstd_f;
      for (int k = 0; k < cg; ++k) {
        const int ch = grp * cg + k;
        const double a2 = (prm.gn_weight ? (double)prm.gn_weight[ch] : 1.0) * (double)rstd_f;
        const double b2 = (prm.gn_bias ? (double)prm.gn_bias[ch] : 0.0) - (double)mean_f * a2;
        const size_t o = (size_t)img * c + ch;
        ab[2 * o] = (float)(a2 * alpha_of(ch)); ab[2 * o + 1] = (float)(a2 * beta_of(ch) + b2);
      }
    }
    return;
  }
  // 2. GroupNorm per (n, group) on u = alpha*y + beta
  if (prm.use_gn) {
    const double mg = (double)cg * hw;
    const int g_lo = ch_lo / cg, gl = cl / cg;
    for (int j = tid; j < n * gl; j += nt) {
      const int img = j / gl, grp = g_lo + (j - img * gl);
      const size_t i = (size_t)img * G + grp;
      double su = 0.0, suu = 0.0;
      for (int k = 0; k < cg; ++k) {
        const int ch = grp * cg + k;
        const double al = alpha_of(ch), be = beta_of(ch);
        const double sy = (double)stats[((size_t)img * c + ch) * 2], syy = (double)stats[((size_t)img * c + ch) * 2 + 1];
        su += al * sy + hw * be;
        suu += al * al * syy + 2.0 * al * be * sy + hw * be * be;
      }
      const double mean = su / mg;
      double var = suu / mg - mean * mean;
      if (var < 0.0) var = 0.0;
      saved[off_gn(c) + 2 * (size_t)i] = (float)mean;
      saved[off_gn(c) + 2 * (size_t)i + 1] = (float)rsqrt(var + (double)prm.gn_eps);
    }
    __syncthreads();
  }
  // 3. combined affine per (n, c)
  for (int j = tid; j < n * cl; j += nt) {
    const int img = j / cl, ch = ch_lo + (j - img * cl);
    const size_t i = (size_t)img * c + ch;
    double A = alpha_of(ch), B = beta_of(ch);
    if (prm.use_gn) {
      const size_t gi = (size_t)img * G + ch / cg;
      const double mean = (double)saved[off_gn(c) + 2 * gi], rstd = (double)saved[off_gn(c) + 2 * gi + 1];
      const double a2 = (prm.gn_weight ? (double)prm.gn_weight[ch] : 1.0) * rstd;
      const double b2 = (prm.gn_bias ? (double)prm.gn_bias[ch] : 0.0) - mean * a2;
      B = a2 * B + b2;
      A = a2 * A;
    }
    ab[2 * (size_t)i] = (float)A; ab[2 * (size_t)i + 1] = (float)B;
  }
}

// D1, D2, D3 of du = D1*dz + D2*y + D3 (GroupNorm adjoint w.r.t. its input u, expressed on (dz, y)).
__device__ __forceinline__ void gn_adjoint_coeffs(const dcv_norm_params& prm, const float* saved, int img, int ch, int G, double& D1, double& D2, double& D3) {
  if (!prm.use_gn) { D1 = 1.0; D2 = 0.0; D3 = 0.0; return; }
  const int c = prm.c, cg = c / G;
  const size_t gi = (size_t)img * G + ch / cg;
  const double mean_u = (double)saved[off_gn(c) + 2 * gi], r = (double)saved[off_gn(c) + 2 * gi + 1];
  const double Ag = (double)saved[off_ab(c, prm.n, G) + 2 * gi], Bg = (double)saved[off_ab(c, prm.n, G) + 2 * gi + 1];
  const double al = (double)saved[off_alpha(c) + 2 * ch], be = (double)saved[off_alpha(c) + 2 * ch + 1];
  const double gam = prm.gn_weight ? (double)prm.gn_weight[ch] : 1.0;
  D1 = r * gam;
  D2 = -r * Bg * r * al;
  D3 = -r * Ag - r * Bg * r * (be - mean_u);
}

__global__ void __launch_bounds__(1024) bwd_finalize_kernel(const dcv_norm_params prm, const float* __restrict__ stats, const float* __restrict__ s,
                                                            float* __restrict__ saved, float* __restrict__ pqr, float* __restrict__ d_bn_w, float* __restrict__ d_bn_b,
                                                            float* __restrict__ d_gn_w, float* __restrict__ d_gn_b, const int cpb) {
  const int n = prm.n, c = prm.c, G = prm.use_gn ? prm.gn_groups : 1, cg = c / G;
  const double hw = (double)prm.hw;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ch_lo = blockIdx.x * cpb, ch_hi = min(c, ch_lo + cpb), cl = ch_hi - ch_lo;   // whole groups per CTA (see fwd_finalize_kernel)
  // 1. GroupNorm group sums A_g = mean_g(gamma*dz), B_g = mean_g(gamma*dz*u_hat)
  if (prm.use_gn) {
    const double mg = (double)cg * hw;
    const int g_lo = ch_lo / cg, gl = cl / cg;
    for (int j = tid; j < n * gl; j += nt) {
      const int img = j / gl, grp = g_lo + (j - img * gl);
      const size_t i = (size_t)img * G + grp;
      const double mean_u = (double)saved[off_gn(c) + 2 * (size_t)i], r = (double)saved[off_gn(c) + 2 * (size_t)i + 1];
      double a = 0.0, b = 0.0;
      for (int k = 0; k < cg; ++k) {
        const int ch = grp * cg + k;
        const double gam = prm.gn_weight ? (double)prm.gn_weight[ch] : 1.0;
        const double al = (double)saved[off_alpha(c) + 2 * ch], be = (double)saved[off_alpha(c) + 2 * ch + 1];
        const double s1 = (double)s[((size_t)img * c + ch) * kBwdSums], s2 = (double)s[((size_t)img * c + ch) * kBwdSums + 1];
        a += gam * s1;
        b += gam * r * (al * s2 + (be - mean_u) * s1);
      }
      saved[off_ab(c, n, G) + 2 * (size_t)i] = (float)(a / mg);
      saved[off_ab(c, n, G) + 2 * (size_t)i + 1] = (float)(b / mg);
    }
    __syncthreads();
  }
  // 2. per-channel sums over images: BatchNorm adjoint sums U1 = sum du, U2 = sum du*y_hat; parameter gradients.
  //    `wpc` warps per channel, lanes stride over the images, fp64 warp reductions, partials combined through shared memory.
  const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
  __shared__ double part[32][4];
  const int wpc = max(1, nwarps / max(cl, 1)), cpp = nwarps / wpc;
  for (int c0 = 0; c0 < cl; c0 += cpp) {
    const int chl = c0 + warp / wpc, sub = warp % wpc;
    if (warp / wpc < cpp && chl < cl) {
      const int ch = ch_lo + chl;
      const double al = (double)saved[off_alpha(c) + 2 * ch], be = (double)saved[off_alpha(c) + 2 * ch + 1];
      double u1 = 0.0, u2raw = 0.0, dgw = 0.0, dgb = 0.0;
      for (int img = sub * 32 + lane; img < n; img += wpc * 32) {
        const size_t i = (size_t)img * c + ch;
        const double sy = (double)stats[2 * i], syy = (double)stats[2 * i + 1], s1 = (double)s[kBwdSums * i], s2 = (double)s[kBwdSums * i + 1];
        double D1, D2, D3;
        gn_adjoint_coeffs(prm, saved, img, ch, G, D1, D2, D3);
        u1 += D1 * s1 + D2 * sy + D3 * hw;
        u2raw += D1 * s2 + D2 * syy + D3 * sy;
        if (prm.use_gn) {
          const size_t gi = (size_t)img * G + ch / cg;
          const double mean_u = (double)saved[off_gn(c) + 2 * gi], r = (double)saved[off_gn(c) + 2 * gi + 1];
          dgw += r * (al * s2 + (be - mean_u) * s1);
          dgb += s1;
        }
      }
      u1 = warp_sum_d(u1); u2raw = warp_sum_d(u2raw);
      if (prm.use_gn) { dgw = warp_sum_d(dgw); dgb = warp_sum_d(dgb); }
      if (lane == 0) { part[warp][0] = u1; part[warp][1] = u2raw; part[warp][2] = dgw; part[warp][3] = dgb; }
    }
    __syncthreads();
    if (warp < cpp && c0 + warp < cl) {   // warp w finishes channel c0 + w: its lanes fold the wpc partials
      const int ch = ch_lo + c0 + warp;
      const double mu = (double)saved[2 * ch], rc = (double)saved[2 * ch + 1];
      const bool has = lane < wpc;
      double u1 = has ? part[warp * wpc + lane][0] : 0.0, u2raw = has ? part[warp * wpc + lane][1] : 0.0;
      double dgw = has ? part[warp * wpc + lane][2] : 0.0, dgb = has ? part[warp * wpc + lane][3] : 0.0;
      u1 = warp_sum_d(u1); u2raw = warp_sum_d(u2raw);
      if (prm.use_gn) { dgw = warp_sum_d(dgw); dgb = warp_sum_d(dgb); }
      if (lane == 0) {
        const double u2 = rc * (u2raw - mu * u1);
        saved[off_u(c, n, G) + 2 * ch] = (float)u1;
        saved[off_u(c, n, G) + 2 * ch + 1] = (float)u2;
        if (prm.use_bn) { if (d_bn_w) d_bn_w[ch] = (float)u2; if (d_bn_b) d_bn_b[ch] = (float)u1; }
        if (prm.use_gn) { if (d_gn_w) d_gn_w[ch] = (float)dgw; if (d_gn_b) d_gn_b[ch] = (float)dgb; }
      }
    }
    __syncthreads();
  }
  // 3. P, Q, R per (n, c)
  const double m = (double)n * hw;
  for (int j = tid; j < n * cl; j += nt) {
    const int img = j / cl, ch = ch_lo + (j - img * cl);
    const size_t i = (size_t)img * c + ch;
    double D1, D2, D3;
    gn_adjoint_coeffs(prm, saved, img, ch, G, D1, D2, D3);
    double P = D1, Q = D2, R = D3;
    if (prm.use_bn) {
      const double al = (double)saved[off_alpha(c) + 2 * ch];  // gamma * rstd
      if (prm.bn_training) {
        const double mu = (double)saved[2 * ch], rc = (double)saved[2 * ch + 1];
        const double u1 = (double)saved[off_u(c, n, G) + 2 * ch], u2 = (double)saved[off_u(c, n, G) + 2 * ch + 1];
        P = al * D1;
        Q = al * (D2 - rc * u2 / m);
        R = al * (D3 - u1 / m + mu * rc * u2 / m);
      } else {
        P = al * D1; Q = al * D2; R = al * D3;
      }
    }
    pqr[3 * (size_t)i] = (float)P; pqr[3 * (size_t)i + 1] = (float)Q; pqr[3 * (size_t)i + 2] = (float)R;
  }
}

// 16-byte vectors are usable when a pixel is a whole number of vectors, or a vector a whole number of pixels (and of an image's pixels)
static bool vec_ok(const void* a, const void* b, const void* d, int n, int hw, int c, int ve) {
  auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % 16 == 0); };
  (void)n;
  const bool shape_ok = c % ve == 0 || (c < ve && ve % c == 0 && hw % (ve / c) == 0);
  return shape_ok && al(a) && al(b) && al(d);
}

static int check_nc(const char* name, int n, int hw, int c) {
  DCV_REQUIRE(n > 0 && hw > 0 && c > 0, "%s: bad shape n=%d hw=%d c=%d", name, n, hw, c);
  DCV_REQUIRE((long long)n * hw < (1ll << 31), "%s: n*hw=%lld exceeds the 32-bit row index range", name, (long long)n * hw);
  return 0;
}

}  // namespace dcv

extern "C" {

size_t dcv_norm_saved_floats(int n, int c, int groups) {
  const size_t G = groups > 0 ? groups : 1;
  return 4 * (size_t)c + 4 * (size_t)n * G + 2 * (size_t)c;
}

int dcv_norm_stats(const void* y, float* stats_nc, int n, int hw, int c, int dtype, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(y && stats_nc, "norm_stats: null pointer");
  if (check_nc("norm_stats", n, hw, c)) return 1;
  cudaStream_t st = as_stream(stream);
  zero_accumulator(stats_nc, (size_t)n * c * 2 * sizeof(float), st, (acc_prezeroed & DCV_ACC_PREZEROED) != 0);
  // DCV_STATS_CHANNEL_TOTALS (a BatchNorm-only block): the batch is walked as ONE image of n * hw pixels — one CTA reduction per CTA instead of one per
  // (CTA, image) — and the totals are credited to image 0. (7 x 7 maps: a CTA's range used to cross dozens of images, each with its own flush.)
  if ((acc_prezeroed & DCV_STATS_CHANNEL_TOTALS) && (long long)n * hw < (1ll << 31)) { hw *= n; n = 1; }
  dim3 grid; int block;
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    if (vec_ok(y, nullptr, nullptr, n, hw, c, VE)) { static const int occ = streaming_ctas_per_sm((const void*)stats_kernel<T, VE>); NcGeom g = make_geom<VE>(n, hw, c, occ, &grid, &block); stats_kernel<T, VE><<<grid, block, 0, st>>>((const T*)y, stats_nc, g); }
    else { static const int occ = streaming_ctas_per_sm((const void*)stats_kernel<T, 1>); NcGeom g = make_geom<1>(n, hw, c, occ, &grid, &block); stats_kernel<T, 1><<<grid, block, 0, st>>>((const T*)y, stats_nc, g); }
  });
  DCV_LAUNCH_CHECK("stats_kernel");
  return 0;
}

// Channels per finalize CTA: whole GroupNorm groups, about 512 (image, channel) items per CTA (a 1024-thread CTA then makes one pass over them; ncu on
// the former single-CTA launch for the CIFAR layers: 13 us of serial latency), at most one CTA per SM.
static void finalize_grid(const dcv_norm_params* prm, int* cpb, int* blocks) {
  const int cg = prm->use_gn ? prm->c / prm->gn_groups : 1;
  int per = 512 / (prm->n > 0 ? prm->n : 1);
  if (per < 1) per = 1;
  if (per > 32) per = 32;   // one warp per channel and pass: a handful of images (or a whole batch handed over as ONE image) must not end up in a single CTA
  per = (per + cg - 1) / cg * cg;
  while ((prm->c + per - 1) / per > dcv::kNumSMs) per += cg;
  *cpb = per;
  *blocks = (prm->c + per - 1) / per;
  // momentum = None (cumulative moving average): every channel's update reads num_batches_tracked and block 0 increments it — one CTA, no race
  if (prm->use_bn && prm->bn_training && prm->bn_momentum < 0.f && prm->bn_num_batches_tracked) { *cpb = prm->c; *blocks = 1; }
}

static int check_norm_params(const dcv_norm_params* prm, const char* name) {
  using namespace dcv;
  DCV_REQUIRE(prm, "%s: null params", name);
  if (check_nc(name, prm->n, prm->hw, prm->c)) return 1;
  DCV_REQUIRE(!prm->use_gn || (prm->gn_groups > 0 && prm->c % prm->gn_groups == 0), "%s: num_channels=%d not divisible by num_groups=%d", name, prm->c, prm->gn_groups);
  DCV_REQUIRE(!prm->use_bn || prm->bn_training || (prm->bn_running_mean && prm->bn_running_var), "%s: eval-mode BatchNorm needs running statistics", name);
  return 0;
}

int dcv_norm_fwd_finalize(const dcv_norm_params* prm, const float* stats_nc, float* ab_nc, float* saved, void* stream) {
  using namespace dcv;
  if (check_norm_params(prm, "norm_fwd_finalize")) return 1;
  DCV_REQUIRE(stats_nc && ab_nc && saved, "norm_fwd_finalize: null pointer");
  int cpb, blocks;
  finalize_grid(prm, &cpb, &blocks);
  fwd_finalize_kernel<<<blocks, 1024, 0, as_stream(stream)>>>(*prm, stats_nc, ab_nc, saved, cpb);
  DCV_LAUNCH_CHECK("fwd_finalize_kernel");
  return 0;
}

int dcv_bn_apply_fold_fwd(const void* y, const void* other, int pool, void* out, const float* stats_c, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, long long* num_batches_tracked, float eps, float momentum, int bn_training, float* saved, int rows, int w, int c, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(y && out && stats_c && saved, "bn_apply_fold_fwd: null pointer");
  DCV_REQUIRE(rows > 0 && w > 0 && c > 0 && (long long)rows * w < (1ll << 31) && momentum >= 0.f, "bn_apply_fold_fwd: bad arguments (momentum = None is served by dcv_norm_fwd_finalize)");
  DCV_REQUIRE(bn_training || (running_mean && running_var), "bn_apply_fold_fwd: eval-mode BatchNorm needs running statistics");
  DCV_REQUIRE(!(pool && other), "bn_apply_fold_fwd: pooling and a residual operand are exclusive");
  cudaStream_t st = as_stream(stream);
  const int hw = rows * w;
  BnFwdFold fold{stats_c, gamma, beta, running_mean, running_var, num_batches_tracked, saved, (double)hw, (double)eps, (double)momentum, c, bn_training};
  dim3 grid; int block;
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    DCV_REQUIRE(c % VE == 0 && vec_ok(y, out, other, 1, hw, c, VE), "bn_apply_fold_fwd: channels must be whole 16-byte vectors at 16-byte aligned pointers");
    if (pool) {
      DCV_REQUIRE(rows % 2 == 0 && w % 2 == 0 && c <= 4096, "bn_apply_fold_fwd: the 2x2 / stride-2 windows must tile the input (and c <= 4096)");
      const int p = rows / 2, q = w / 2;
      const uint32_t total = (uint32_t)((size_t)p * q * (c / VE));
      apply_pool_fwd_kernel<T, VE, true><<<grid_for(total, 256), 256, (size_t)c * sizeof(float2), st>>>((const T*)y, nullptr, (T*)out, rows, w, c, total, FastDiv(c / VE), FastDiv(q), FastDiv(p), fold);
    } else if (other) {
      static const int occ = streaming_ctas_per_sm((const void*)apply_fwd_kernel<T, VE, true, true>);
      NcGeom g = make_geom<VE>(1, hw, c, occ, &grid, &block);
      apply_fwd_kernel<T, VE, true, true><<<grid, block, 0, st>>>((const T*)y, nullptr, (T*)out, g, (const T*)other, fold);
    } else {
      static const int occ = streaming_ctas_per_sm((const void*)apply_fwd_kernel<T, VE, false, true>);
      NcGeom g = make_geom<VE>(1, hw, c, occ, &grid, &block);
      apply_fwd_kernel<T, VE, false, true><<<grid, block, 0, st>>>((const T*)y, nullptr, (T*)out, g, nullptr, fold);
    }
  });
  DCV_LAUNCH_CHECK("bn_apply_fold_fwd");
  return 0;
}

int dcv_norm_apply_add_fwd(const void* y, const float* ab_nc, const void* other, void* z, int n, int hw, int c, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(y && ab_nc && other && z, "norm_apply_add_fwd: null pointer");
  if (check_nc("norm_apply_add_fwd", n, hw, c)) return 1;
  cudaStream_t st = as_stream(stream);
  dim3 grid; int block;
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    if (vec_ok(y, z, other, n, hw, c, VE)) { static const int occ = streaming_ctas_per_sm((const void*)apply_fwd_kernel<T, VE, true>); NcGeom g = make_geom<VE>(n, hw, c, occ, &grid, &block); apply_fwd_kernel<T, VE, true><<<grid, block, 0, st>>>((const T*)y, ab_nc, (T*)z, g, (const T*)other); }
    else { static const int occ = streaming_ctas_per_sm((const void*)apply_fwd_kernel<T, 1, true>); NcGeom g = make_geom<1>(n, hw, c, occ, &grid, &block); apply_fwd_kernel<T, 1, true><<<grid, block, 0, st>>>((const T*)y, ab_nc, (T*)z, g, (const T*)other); }
  });
  DCV_LAUNCH_CHECK("apply_fwd_kernel(add)");
  return 0;
}

int dcv_norm_apply_pool_fwd(const void* y, const float* ab_nc, void* zp, int n, int h, int w, int c, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(y && ab_nc && zp, "norm_apply_pool_fwd: null pointer");
  DCV_REQUIRE(n > 0 && c > 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0, "norm_apply_pool_fwd: the 2x2 / stride-2 windows must tile the %dx%d input", h, w);
  DCV_REQUIRE((size_t)n * (h / 2) * (w / 2) * c < (1ull << 31), "norm_apply_pool_fwd: tensor too large");
  cudaStream_t st = as_stream(stream);
  const int p = h / 2, q = w / 2;
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    if (c % VE == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0 && reinterpret_cast<uintptr_t>(zp) % 16 == 0) {
      const uint32_t total = (uint32_t)((size_t)n * p * q * (c / VE));
      apply_pool_fwd_kernel<T, VE><<<grid_for(total, 256), 256, 0, st>>>((const T*)y, ab_nc, (T*)zp, h, w, c, total, FastDiv(c / VE), FastDiv(q), FastDiv(p));
    } else {
      const uint32_t total = (uint32_t)((size_t)n * p * q * c);
      apply_pool_fwd_kernel<T, 1><<<grid_for(total, 256), 256, 0, st>>>((const T*)y, ab_nc, (T*)zp, h, w, c, total, FastDiv(c), FastDiv(q), FastDiv(p));
    }
  });
  DCV_LAUNCH_CHECK("apply_pool_fwd_kernel");
  return 0;
}

int dcv_norm_apply_fwd(const void* y, const float* ab_nc, void* z, int n, int hw, int c, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(y && ab_nc && z, "norm_apply_fwd: null pointer");
  if (check_nc("norm_apply_fwd", n, hw, c)) return 1;
  cudaStream_t st = as_stream(stream);
  dim3 grid; int block;
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    if (vec_ok(y, z, nullptr, n, hw, c, VE)) { static const int occ = streaming_ctas_per_sm((const void*)apply_fwd_kernel<T, VE>); NcGeom g = make_geom<VE>(n, hw, c, occ, &grid, &block); apply_fwd_kernel<T, VE><<<grid, block, 0, st>>>((const T*)y, ab_nc, (T*)z, g); }
    else { static const int occ = streaming_ctas_per_sm((const void*)apply_fwd_kernel<T, 1>); NcGeom g = make_geom<1>(n, hw, c, occ, &grid, &block); apply_fwd_kernel<T, 1><<<grid, block, 0, st>>>((const T*)y, ab_nc, (T*)z, g); }
  });
  DCV_LAUNCH_CHECK("apply_fwd_kernel");
  return 0;
}

int dcv_norm_bwd_reduce(const void* dz, const void* y, float* s_nc, int n, int hw, int c, int dtype, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(dz && y && s_nc, "norm_bwd_reduce: null pointer");
  if (check_nc("norm_bwd_reduce", n, hw, c)) return 1;
  cudaStream_t st = as_stream(stream);
  zero_accumulator(s_nc, (size_t)n * c * kBwdSums * sizeof(float), st, (acc_prezeroed & DCV_ACC_PREZEROED) != 0);
  if ((acc_prezeroed & DCV_STATS_CHANNEL_TOTALS) && (long long)n * hw < (1ll << 31)) { hw *= n; n = 1; }   // as dcv_norm_stats: totals credited to image 0
  dim3 grid; int block;
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    if (vec_ok(dz, y, nullptr, n, hw, c, VE)) { static const int occ = streaming_ctas_per_sm((const void*)bwd_reduce_kernel<T, VE>); NcGeom g = make_geom<VE>(n, hw, c, occ, &grid, &block); bwd_reduce_kernel<T, VE><<<grid, block, 0, st>>>((const T*)dz, (const T*)y, s_nc, g); }
    else { static const int occ = streaming_ctas_per_sm((const void*)bwd_reduce_kernel<T, 1>); NcGeom g = make_geom<1>(n, hw, c, occ, &grid, &block); bwd_reduce_kernel<T, 1><<<grid, block, 0, st>>>((const T*)dz, (const T*)y, s_nc, g); }
  });
  DCV_LAUNCH_CHECK("bwd_reduce_kernel");
  return 0;
}

int dcv_norm_bwd_pooled_supported(int n, int h, int w, int c, int dtype) {
  const int ve = dtype == DCV_BF16 ? 8 : 4;
  return (dtype == DCV_BF16 || dtype == DCV_F32) && n > 0 && c > 0 && c % ve == 0 && h >= 2 && w >= 2 && h % 2 == 0 && w % 2 == 0 && (long long)n * h * w < (1ll << 31);
}

int dcv_norm_bwd_reduce_pooled(const void* dzp, const void* y, float* s_nc, int n, int h, int w, int c, int dtype, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(dzp && y && s_nc, "norm_bwd_reduce_pooled: null pointer");
  DCV_REQUIRE(dcv_norm_bwd_pooled_supported(n, h, w, c, dtype), "norm_bwd_reduce_pooled: shape not supported (see dcv_norm_bwd_pooled_supported)");
  cudaStream_t st = as_stream(stream);
  zero_accumulator(s_nc, (size_t)n * c * kBwdSums * sizeof(float), st, (acc_prezeroed & DCV_ACC_PREZEROED) != 0);
  int hw = h * w;
  if (acc_prezeroed & DCV_STATS_CHANNEL_TOTALS) { hw *= n; n = 1; }
  dim3 grid; int block;
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    DCV_REQUIRE(vec_ok(dzp, y, nullptr, n, hw, c, VE), "norm_bwd_reduce_pooled: pointers must be 16-byte aligned");
    static const int occ = streaming_ctas_per_sm((const void*)bwd_reduce_kernel<T, VE, true>);
    NcGeom g = make_geom<VE>(n, hw, c, occ, &grid, &block);
    bwd_reduce_kernel<T, VE, true><<<grid, block, 0, st>>>((const T*)dzp, (const T*)y, s_nc, g, FastDiv((uint32_t)w), (uint32_t)w);
  });
  DCV_LAUNCH_CHECK("bwd_reduce_kernel(pooled)");
  return 0;
}

int dcv_norm_bwd_finalize(const dcv_norm_params* prm, const float* stats_nc, const float* s_nc, float* saved, float* pqr_nc,
                          float* d_bn_weight, float* d_bn_bias, float* d_gn_weight, float* d_gn_bias, void* stream) {
  using namespace dcv;
  if (check_norm_params(prm, "norm_bwd_finalize")) return 1;
  DCV_REQUIRE(stats_nc && s_nc && saved && pqr_nc, "norm_bwd_finalize: null pointer");
  int cpb, blocks;
  finalize_grid(prm, &cpb, &blocks);
  bwd_finalize_kernel<<<blocks, 1024, 0, as_stream(stream)>>>(*prm, stats_nc, s_nc, saved, pqr_nc, d_bn_weight, d_bn_bias, d_gn_weight, d_gn_bias, cpb);
  DCV_LAUNCH_CHECK("bwd_finalize_kernel");
  return 0;
}

int dcv_act_norm_bwd_apply(const void* dz, const void* y, const float* pqr_nc, void* dy, float* dbias_c, int act, float slope,
                           int n, int hw, int c, int dtype, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(dz && y && dy, "act_norm_bwd_apply: null pointer");
  if (check_nc("act_norm_bwd_apply", n, hw, c)) return 1;
  cudaStream_t st = as_stream(stream);
  zero_accumulator(dbias_c, (size_t)c * sizeof(float), st, acc_prezeroed != 0);
  dim3 grid; int block;
#define DCV_BWD_APPLY(ACT_)                                                                                                                                   \
  DCV_DISPATCH_DTYPE(dtype, T, {                                                                                                                              \
    constexpr int VE = 16 / sizeof(T);                                                                                                                        \
    if (vec_ok(dz, y, dy, n, hw, c, VE)) { static const int occ = streaming_ctas_per_sm((const void*)bwd_apply_kernel<T, VE, ACT_>); NcGeom g = make_geom<VE>(n, hw, c, occ, &grid, &block, dbias_c ? 8 : 2); bwd_apply_kernel<T, VE, ACT_><<<grid, block, 0, st>>>((const T*)dz, (const T*)y, pqr_nc, (T*)dy, dbias_c, slope, g); } \
    else { static const int occ = streaming_ctas_per_sm((const void*)bwd_apply_kernel<T, 1, ACT_>); NcGeom g = make_geom<1>(n, hw, c, occ, &grid, &block, dbias_c ? 8 : 2); bwd_apply_kernel<T, 1, ACT_><<<grid, block, 0, st>>>((const T*)dz, (const T*)y, pqr_nc, (T*)dy, dbias_c, slope, g); }                      \
  })
  switch (act) {
    case DCV_ACT_RELU: DCV_BWD_APPLY(DCV_ACT_RELU); break;
    case DCV_ACT_LEAKY_RELU: DCV_BWD_APPLY(DCV_ACT_LEAKY_RELU); break;
    case DCV_ACT_SIGMOID: DCV_BWD_APPLY(DCV_ACT_SIGMOID); break;
    default: DCV_BWD_APPLY(DCV_ACT_NONE); break;
  }
#undef DCV_BWD_APPLY
  DCV_LAUNCH_CHECK("bwd_apply_kernel");
  return 0;
}

int dcv_act_bn_bwd_apply_fold(const void* dz, int pooled, const void* y, const float* s_c, const float* saved, int bn_training, float* d_bn_weight, float* d_bn_bias,
                              void* dy, float* dbias_c, int act, float slope, int rows, int w, int c, int dtype, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(dz && y && dy && s_c && saved, "act_bn_bwd_apply_fold: null pointer");
  DCV_REQUIRE(rows > 0 && w > 0 && c > 0 && (long long)rows * w < (1ll << 31), "act_bn_bwd_apply_fold: bad shape");
  DCV_REQUIRE(!pooled || dcv_norm_bwd_pooled_supported(1, rows, w, c, dtype), "act_bn_bwd_apply_fold: pooled gradient needs even rows / width and whole 16-byte channel vectors");
  cudaStream_t st = as_stream(stream);
  zero_accumulator(dbias_c, (size_t)c * sizeof(float), st, acc_prezeroed != 0);
  const int hw = rows * w;
  BnFold fold{s_c, saved, d_bn_weight, d_bn_bias, 1.0 / (double)hw, c, bn_training};
  dim3 grid; int block;
#define DCV_FOLD_LAUNCH(ACT_, POOL_)                                                                                                                          \
  DCV_DISPATCH_DTYPE(dtype, T, {                                                                                                                              \
    constexpr int VE = 16 / sizeof(T);                                                                                                                        \
    DCV_REQUIRE(c % VE == 0 && vec_ok(dz, y, dy, 1, hw, c, VE), "act_bn_bwd_apply_fold: channels must be whole 16-byte vectors at 16-byte aligned pointers");  \
    static const int occ = streaming_ctas_per_sm((const void*)bwd_apply_kernel<T, VE, ACT_, POOL_, true>);                                                  \
    NcGeom g = make_geom<VE>(1, hw, c, occ, &grid, &block, dbias_c ? 8 : 2);                                                                                  \
    bwd_apply_kernel<T, VE, ACT_, POOL_, true><<<grid, block, 0, st>>>((const T*)dz, (const T*)y, nullptr, (T*)dy, dbias_c, slope, g, FastDiv((uint32_t)w), (uint32_t)w, fold); \
  })
#define DCV_FOLD_ACT(ACT_) do { if (pooled) { DCV_FOLD_LAUNCH(ACT_, true); } else { DCV_FOLD_LAUNCH(ACT_, false); } } while (0)
  switch (act) {
    case DCV_ACT_RELU: DCV_FOLD_ACT(DCV_ACT_RELU); break;
    case DCV_ACT_LEAKY_RELU: DCV_FOLD_ACT(DCV_ACT_LEAKY_RELU); break;
    case DCV_ACT_SIGMOID: DCV_FOLD_ACT(DCV_ACT_SIGMOID); break;
    default: DCV_FOLD_ACT(DCV_ACT_NONE); break;
  }
#undef DCV_FOLD_ACT
#undef DCV_FOLD_LAUNCH
  DCV_LAUNCH_CHECK("bwd_apply_kernel(fold)");
  return 0;
}

int dcv_act_norm_bwd_apply_pooled(const void* dzp, const void* y, const float* pqr_nc, void* dy, float* dbias_c, int act, float slope,
                                  int n, int h, int w, int c, int dtype, int acc_prezeroed, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(dzp && y && dy, "act_norm_bwd_apply_pooled: null pointer");
  DCV_REQUIRE(dcv_norm_bwd_pooled_supported(n, h, w, c, dtype), "act_norm_bwd_apply_pooled: shape not supported (see dcv_norm_bwd_pooled_supported)");
  cudaStream_t st = as_stream(stream);
  zero_accumulator(dbias_c, (size_t)c * sizeof(float), st, acc_prezeroed != 0);
  const int hw = h * w;
  dim3 grid; int block;
#define DCV_BWD_APPLY_POOLED(ACT_)                                                                                                                            \
  DCV_DISPATCH_DTYPE(dtype, T, {                                                                                                                              \
    constexpr int VE = 16 / sizeof(T);                                                                                                                        \
    DCV_REQUIRE(vec_ok(dzp, y, dy, n, hw, c, VE), "act_norm_bwd_apply_pooled: pointers must be 16-byte aligned");                                             \
    static const int occ = streaming_ctas_per_sm((const void*)bwd_apply_kernel<T, VE, ACT_, true>);                                                          \
    NcGeom g = make_geom<VE>(n, hw, c, occ, &grid, &block, dbias_c ? 8 : 2);                                                                                  \
    bwd_apply_kernel<T, VE, ACT_, true><<<grid, block, 0, st>>>((const T*)dzp, (const T*)y, pqr_nc, (T*)dy, dbias_c, slope, g, FastDiv((uint32_t)w), (uint32_t)w); \
  })
  switch (act) {
    case DCV_ACT_RELU: DCV_BWD_APPLY_POOLED(DCV_ACT_RELU); break;
    case DCV_ACT_LEAKY_RELU: DCV_BWD_APPLY_POOLED(DCV_ACT_LEAKY_RELU); break;
    case DCV_ACT_SIGMOID: DCV_BWD_APPLY_POOLED(DCV_ACT_SIGMOID); break;
    default: DCV_BWD_APPLY_POOLED(DCV_ACT_NONE); break;
  }
#undef DCV_BWD_APPLY_POOLED
  DCV_LAUNCH_CHECK("bwd_apply_kernel(pooled)");
  return 0;
}

}  // extern "C"
