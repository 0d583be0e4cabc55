// Layout / dtype plumbing: NCHW <-> NHWC transposes (tiled through shared memory so both sides stay coalesced),
// flat casts, and convolution-weight packing for the tensor-core / data-gradient operands.
#include "common.cuh"

namespace dcv {

// src viewed as [batch][rows][cols] -> dst [batch][cols][rows]
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) transpose_kernel(const TS* __restrict__ src, TD* __restrict__ dst, int rows, int cols, int tiles_r, int tiles_c) {
  __shared__ float tile[32][33];
  const int tiles_per_batch = tiles_r * tiles_c;
  const int b = blockIdx.x / tiles_per_batch;
  const int t = blockIdx.x - b * tiles_per_batch;
  const int tr = t / tiles_c, tc = t - tr * tiles_c;
  const size_t base = (size_t)b * rows * cols;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = tr * 32 + ty + 8 * k, c = tc * 32 + tx;
    if (r < rows && c < cols) tile[ty + 8 * k][tx] = to_f<TS>(src[base + (size_t)r * cols + c]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = tc * 32 + ty + 8 * k, r = tr * 32 + tx;
    if (r < rows && c < cols) dst[base + (size_t)c * rows + r] = from_f<TD>(tile[tx][ty + 8 * k]);
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, TD* __restrict__ dst, size_t count) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) dst[i] = from_f<TD>(to_f<TS>(src[i]));
}

template <typename TD>
__global__ void pack_weight_kernel(const float* __restrict__ w, TD* __restrict__ dst, int k, int r, int s, int c, int transpose_flip) {
  const size_t total = (size_t)k * r * s * c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t t = i;
    const int ci = t % c; t /= c;
    const int si = t % s; t /= s;
    const int ri = t % r; const int ki = t / r;
    const size_t o = transpose_flip ? ((((size_t)ci * r + (r - 1 - ri)) * s + (s - 1 - si)) * k + ki) : i;
    dst[o] = from_f<TD>(w[i]);
  }
}

// All data-gradient weight operands of a step in ONE launch (17 pack launches of ~8.5 us each on the ImageNet-shaped step otherwise): entry e is a
// [K][R][S][C] fp32 weight inside the flat parameter buffer and its [C][R-1-r][S-1-s][K] copy inside one destination buffer. Work unit = (filter tap,
// 8-channel block, 32-filter block): a lane reads 8 contiguous floats of its filter (32 bytes) and writes one element of each of 8 channel rows, the
// warp's stores forming 64-byte runs.
template <typename TD>
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const float* __restrict__ flat, TD* __restrict__ dst, const dcv_pack_entry* __restrict__ entries, int n_entries,
                                                                   long long total_units) {
  __shared__ dcv_pack_entry sh[64];
  for (int i = threadIdx.x; i < n_entries; i += blockDim.x) sh[i] = entries[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long u = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); u < total_units; u += warps) {
    int e = 0;
    while (e + 1 < n_entries && u >= sh[e + 1].unit0) ++e;   // <= 64 entries: a short scan (warp-uniform)
    const dcv_pack_entry& en = sh[e];
    long long t = u - en.unit0;
    const int kb_n = (en.k + 31) / 32, cb_n = (en.c + 7) / 8;
    const int kb = (int)(t % kb_n); t /= kb_n;
    const int cb = (int)(t % cb_n); t /= cb_n;
    const int si = (int)(t % en.s), ri = (int)(t / en.s);
    const int ki = kb * 32 + lane, c0 = cb * 8;
    if (ki >= en.k) continue;
    const float* src = flat + en.src_off + (((size_t)ki * en.r + ri) * en.s + si) * en.c + c0;
    float v[8];
    if (c0 + 8 <= en.c && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = c0 + i < en.c ? __ldg(src + i) : 0.f;
    }
    TD* out = dst + en.dst_off + (((size_t)c0 * en.r + (en.r - 1 - ri)) * en.s + (en.s - 1 - si)) * en.k + ki;
    const size_t cstride = (size_t)en.r * en.s * en.k;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (c0 + i < en.c) out[i * cstride] = from_f<TD>(v[i]);
  }
}

// col[n][p][q][kpad]: entries (r, s, c) of the receptive field of output pixel (p, q) in the order of the [K][R][S][C] weight layout, zero
// padded from R*S*C to kpad. One thread per 16-byte output vector: the (cached, redundant) gathers are scalar, the 1.2 GB-class store is coalesced.
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ x, T* __restrict__ col, const dcv_conv_shape s, const int kpad, const FastDiv div_vpp, const FastDiv div_q, const FastDiv div_p,
                              const FastDiv div_sc, const FastDiv div_c) {
  constexpr int VE = 16 / sizeof(T);
  const int vec_per_pix = kpad / VE, rsc = s.r * s.s * s.c, sc = s.s * s.c;
  const uint32_t total = (uint32_t)((size_t)s.n * s.p * s.q * vec_per_pix);
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    uint32_t t = div_vpp.div(idx);
    const int v = (int)(idx - t * vec_per_pix);
    uint32_t t2 = div_q.div(t);
    const int oq = (int)(t - t2 * s.q);
    const int img = (int)div_p.div(t2), op = (int)(t2 - (uint32_t)img * s.p);
    const int iy0 = op * s.stride_h - s.pad_h, ix0 = oq * s.stride_w - s.pad_w;
    const T* xin = x + (size_t)img * s.h * s.w * s.c;
    // (r, s, c) of the first element by division, of the following ones by carry
    int kk = v * VE;
    int rr = (int)div_sc.div((uint32_t)kk), rem = kk - rr * sc, ss = (int)div_c.div((uint32_t)rem), cc = rem - ss * s.c;
    float vals[VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) {
      float val = 0.f;
      if (kk + e < rsc) {
        const int iy = iy0 + rr * s.dil_h, ix = ix0 + ss * s.dil_w;
        if (iy >= 0 && iy < s.h && ix >= 0 && ix < s.w) val = to_f<T>(__ldg(xin + ((size_t)iy * s.w + ix) * s.c + cc));
      }
      vals[e] = val;
      if (++cc == s.c) { cc = 0; if (++ss == s.s) { ss = 0; ++rr; } }
    }
    __stcs(reinterpret_cast<uint4*>(col + (size_t)idx * VE), vec_pack<T>(vals));
  }
}

// Row-staged variant (the 7x7 / stride-2 stem): a CTA owns ONE output row (image, p). The R input rows that row reads are copied once into shared
// memory with coalesced 16-byte loads — zero-filled left / right margins and zero rows stand in for the convolution padding, so the gather has no
// bounds checks — and every 16-byte output vector is then assembled from 2-byte shared-memory reads. The first version gathered straight from global
// memory (8 scalar loads with 4 bounds checks each per vector): 1.05 ms for the 1.2 GB stem buffer = 19 % of HBM bandwidth, LSU-bound.
template <typename T>
__global__ void __launch_bounds__(256) im2col_rows_kernel(const T* __restrict__ x, T* __restrict__ col, const dcv_conv_shape s, const int kpad, const int rowlen, const int lpad,
                                                          const int vec_stage, const int tab_offset_bytes, const FastDiv div_vpp, const FastDiv div_sc, const FastDiv div_c, const FastDiv div_p) {
  constexpr int VE = 16 / sizeof(T);
  extern __shared__ __align__(16) unsigned char im2col_smem[];
  T* rows = reinterpret_cast<T*>(im2col_smem);   // [R][rowlen]
  const int img = (int)div_p.div(blockIdx.x), op = (int)blockIdx.x - img * s.p;
  const int wc = s.w * s.c, iy0 = op * s.stride_h - s.pad_h;
  const T* ximg = x + (size_t)img * s.h * wc;
  if (vec_stage) {
    const int gpr = rowlen / VE;   // 16-byte groups per staged row; groups are entirely inside or outside [lpad, lpad + wc)
    for (int i = threadIdx.x; i < s.r * gpr; i += 256) {
      const int r = i / gpr, o = (i - r * gpr) * VE, e = o - lpad;
      const int iy = iy0 + r * s.dil_h;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (iy >= 0 && iy < s.h && e >= 0 && e < wc) v = __ldg(reinterpret_cast<const uint4*>(ximg + (size_t)iy * wc + e));
      *reinterpret_cast<uint4*>(rows + r * rowlen + o) = v;
    }
  } else {
    for (int i = threadIdx.x; i < s.r * rowlen; i += 256) {
      const int r = i / rowlen, e = (i - r * rowlen) - lpad;
      const int iy = iy0 + r * s.dil_h;
      T v = from_f<T>(0.f);
      if (iy >= 0 && iy < s.h && e >= 0 && e < wc) v = ximg[(size_t)iy * wc + e];
      rows[i] = v;
    }
  }
  // offset table: tab[kk] = position of column kk = (r, s, c) relative to the first staged element of an output pixel's window, -1 for the zero padding
  int* tab = reinterpret_cast<int*>(im2col_smem + tab_offset_bytes);
  {
    const int rsc = s.r * s.s * s.c, sc = s.s * s.c;
    for (int kk = threadIdx.x; kk < kpad; kk += 256) {
      int o = -1;
      if (kk < rsc) { const int rr = (int)div_sc.div((uint32_t)kk), rem = kk - rr * sc, ss = (int)div_c.div((uint32_t)rem), cc = rem - ss * s.c; o = rr * rowlen + ss * s.dil_w * s.c + cc; }
      tab[kk] = o;
    }
    if (threadIdx.x < VE) rows[s.r * rowlen + threadIdx.x] = from_f<T>(0.f);   // the zero every padding column reads
  }
  __syncthreads();
  const int vpp = kpad / VE;
  const int total = s.q * vpp, zero_at = s.r * rowlen;
  T* out = col + (size_t)blockIdx.x * s.q * kpad;
  for (int idx = threadIdx.x; idx < total; idx += 256) {
    const int q = (int)div_vpp.div((uint32_t)idx), v = idx - q * vpp;
    const int base = lpad + (q * s.stride_w - s.pad_w) * s.c;
    int o[VE];
#pragma unroll
    for (int e = 0; e < VE; e += 4) { const int4 t = *reinterpret_cast<const int4*>(tab + v * VE + e); o[e] = t.x; o[e + 1] = t.y; o[e + 2] = t.z; o[e + 3] = t.w; }
    __align__(16) T vals[VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) vals[e] = rows[o[e] >= 0 ? base + o[e] : zero_at];
    __stcs(reinterpret_cast<uint4*>(out + (size_t)idx * VE), *reinterpret_cast<const uint4*>(vals));
  }
}

template <typename TS, typename TD>
static int launch_transpose(const void* src, void* dst, int batch, int rows, int cols, cudaStream_t st) {
  const int tiles_r = (rows + 31) / 32, tiles_c = (cols + 31) / 32;
  const size_t blocks = (size_t)batch * tiles_r * tiles_c;
  DCV_REQUIRE(blocks < (1u << 31), "transpose: too many tiles");
  transpose_kernel<TS, TD><<<(unsigned)blocks, 256, 0, st>>>((const TS*)src, (TD*)dst, rows, cols, tiles_r, tiles_c);
  DCV_LAUNCH_CHECK("transpose_kernel");
  return 0;
}

static int transpose_dispatch(const void* src, int sd, void* dst, int dd, int batch, int rows, int cols, cudaStream_t st) {
  DCV_REQUIRE(src && dst && batch > 0 && rows > 0 && cols > 0, "transpose: bad arguments");
  if (sd == DCV_F32 && dd == DCV_F32) return launch_transpose<float, float>(src, dst, batch, rows, cols, st);
  if (sd == DCV_F32 && dd == DCV_BF16) return launch_transpose<float, __nv_bfloat16>(src, dst, batch, rows, cols, st);
  if (sd == DCV_BF16 && dd == DCV_F32) return launch_transpose<__nv_bfloat16, float>(src, dst, batch, rows, cols, st);
  if (sd == DCV_BF16 && dd == DCV_BF16) return launch_transpose<__nv_bfloat16, __nv_bfloat16>(src, dst, batch, rows, cols, st);
  DCV_REQUIRE(false, "transpose: unsupported dtypes %d -> %d", sd, dd);
  return 1;
}

// Weight operand of the gather convolution kernels (conv_tc.cu): [K][kpad] with column r * rp + x = w[k][r][x] for x < sc = S*C, zero elsewhere
// ("row padded" K order: every 16-byte chunk of an im2col row is 8 contiguous input elements); and the way back for the weight gradient.
template <typename T>
__global__ void gather_pack_weight_kernel(const T* __restrict__ w, T* __restrict__ w_col, int k, int r, int sc, int rp, int kpad) {
  const size_t total = (size_t)k * kpad;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % kpad), kk = (int)(i / kpad), rr = j / rp, x = j - rr * rp;
    w_col[i] = (rr < r && x < sc) ? w[((size_t)kk * r + rr) * sc + x] : from_f<T>(0.f);
  }
}
__global__ void gather_unpack_wgrad_kernel(const float* __restrict__ dw_col, float* __restrict__ dw, int k, int r, int sc, int rp, int kpad) {
  const size_t total = (size_t)k * r * sc;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % sc); const size_t t = i / sc; const int rr = (int)(t % r), kk = (int)(t / r);
    dw[i] = dw_col[(size_t)kk * kpad + rr * rp + x];
  }
}

// Weight operand of the pixel-pair convolution kernels (conv_tc.cu): [K][256], column pp * 64 + r * 8 + px * C + c = w[k][r][2 pp + px - e][c] with e = pad_w & 1
// (output pixel q reads the pixel pairs q - hp .. q - hp + 3 of every filter row, one 128-byte line per pair), zero elsewhere; and the way back.
constexpr int kPairsK = 256;
template <typename T>
__global__ void pairs_pack_weight_kernel(const T* __restrict__ w, T* __restrict__ w_col, int k, int r, int s, int c, int e) {
  const int total = k * kPairsK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kk = i & (kPairsK - 1), kr = i / kPairsK, pp = kk >> 6, rr = (kk >> 3) & 7, el = kk & 7, px = el / c, ch = el - px * c, tap = 2 * pp + px - e;
    w_col[i] = (rr < r && px < 2 && tap >= 0 && tap < s) ? w[(((size_t)kr * r + rr) * s + tap) * c + ch] : from_f<T>(0.f);
  }
}
__global__ void pairs_unpack_wgrad_kernel(const float* __restrict__ dw_col, float* __restrict__ dw, int k, int r, int s, int c, int e) {
  const int total = k * r * s * c;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t = i;
    const int ch = t % c; t /= c;
    const int tap = t % s; t /= s;
    const int rr = t % r, kr = t / r, wp = tap + e;
    dw[i] = dw_col[(size_t)kr * kPairsK + (wp >> 1) * 64 + rr * 8 + (wp & 1) * c + ch];
  }
}

// Batch assembly for the device-resident input pipeline: dst row j = src row idx[j]; rows are `row_bytes` long (a multiple of 16, both bases 16-byte
// aligned). One warp streams a row with 16-byte accesses; an index outside [0, n_src) traps nothing and copies nothing but flags the batch (err != 0).
__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int64_t* __restrict__ idx, uint4* __restrict__ dst, int n, long long n_src, uint32_t vec_per_row, int* __restrict__ err) {
  const int warps_per_block = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = blockIdx.x * warps_per_block + warp; j < n; j += gridDim.x * warps_per_block) {
    const long long i = idx[j];
    if (i < 0 || i >= n_src) { if (lane == 0 && err) *err = 1; continue; }
    const uint4* s = src + (size_t)i * vec_per_row;
    uint4* d = dst + (size_t)j * vec_per_row;
    for (uint32_t v = lane; v < vec_per_row; v += 128) {   // 4 independent 16-byte loads in flight per lane
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) if (v + u * 32 < vec_per_row) r[u] = __ldg(s + v + u * 32);
#pragma unroll
      for (int u = 0; u < 4; ++u) if (v + u * 32 < vec_per_row) d[v + u * 32] = r[u];
    }
  }
}

// rows that are not 16-byte multiples (labels: one int64 per row): one thread per 4-byte word
__global__ void gather_rows_words_kernel(const uint32_t* __restrict__ src, const int64_t* __restrict__ idx, uint32_t* __restrict__ dst, int n, long long n_src, uint32_t words_per_row, int* __restrict__ err) {
  const size_t total = (size_t)n * words_per_row;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(t / words_per_row); const uint32_t w = (uint32_t)(t - (size_t)j * words_per_row);
    const long long i = idx[j];
    if (i < 0 || i >= n_src) { if (err) *err = 1; continue; }
    dst[t] = src[(size_t)i * words_per_row + w];
  }
}

}  // namespace dcv

extern "C" {

int dcv_nchw_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype, int n, int c, int h, int w, void* stream) {
  return dcv::transpose_dispatch(src, src_dtype, dst, dst_dtype, n, c, h * w, dcv::as_stream(stream));
}

int dcv_nhwc_to_nchw(const void* src, int src_dtype, void* dst, int dst_dtype, int n, int c, int h, int w, void* stream) {
  return dcv::transpose_dispatch(src, src_dtype, dst, dst_dtype, n, h * w, c, dcv::as_stream(stream));
}

int dcv_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t count, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(src && dst, "cast: null pointer");
  if (count == 0) return 0;
  cudaStream_t st = as_stream(stream);
  const int g = grid_for(count, 256);
  if (src_dtype == DCV_F32 && dst_dtype == DCV_BF16) cast_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, count);
  else if (src_dtype == DCV_BF16 && dst_dtype == DCV_F32) cast_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, count);
  else if (src_dtype == DCV_F32 && dst_dtype == DCV_F32) cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, count);
  else if (src_dtype == DCV_BF16 && dst_dtype == DCV_BF16) cast_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, count);
  else DCV_REQUIRE(false, "cast: unsupported dtypes %d -> %d", src_dtype, dst_dtype);
  DCV_LAUNCH_CHECK("cast_kernel");
  return 0;
}

int dcv_im2col(const dcv_conv_shape* shape, const void* x, void* col, int kpad, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(shape && x && col, "im2col: null pointer");
  const int ve = dtype == DCV_BF16 ? 8 : 4;
  DCV_REQUIRE(kpad >= shape->r * shape->s * shape->c && kpad % ve == 0 && reinterpret_cast<uintptr_t>(col) % 16 == 0, "im2col: kpad=%d must cover r*s*c=%d and be a multiple of %d", kpad,
              shape->r * shape->s * shape->c, ve);
  const size_t total = (size_t)shape->n * shape->p * shape->q * (kpad / ve);
  DCV_REQUIRE(total < (1ull << 31), "im2col: %zu output vectors exceed the 32-bit index range", total);
  const FastDiv d_vpp(kpad / ve), d_q(shape->q), d_p(shape->p), d_sc(shape->s * shape->c), d_c(shape->c);
  {
    // row-staged kernel when the R input rows (plus zero margins) fit in shared memory
    const int wc = shape->w * shape->c;
    int lpad = shape->pad_w * shape->c;
    lpad = (lpad + ve - 1) / ve * ve;
    // largest staged index read: lpad + ((q-1)*stride - pad)*c + (s-1)*dil*c + c - 1
    int need = lpad + ((shape->q - 1) * shape->stride_w - shape->pad_w + (shape->s - 1) * shape->dil_w + 1) * shape->c;
    if (need < lpad + wc) need = lpad + wc;
    const int rowlen = (need + ve - 1) / ve * ve;
    const size_t esize = dtype == DCV_BF16 ? 2 : 4;
    const size_t tab_off = ((size_t)(shape->r * rowlen + ve) * esize + 15) / 16 * 16;   // staged rows + one vector of zeros, then the offset table
    const size_t smem = tab_off + (size_t)kpad * sizeof(int);
    const long long ctas = (long long)shape->n * shape->p;
    if (smem <= 96 * 1024 && ctas < (1ll << 31)) {
      const int vec_stage = (wc % ve == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0) ? 1 : 0;
      DCV_DISPATCH_DTYPE(dtype, T, {
        auto kern = im2col_rows_kernel<T>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<(unsigned)ctas, 256, smem, as_stream(stream)>>>((const T*)x, (T*)col, *shape, kpad, rowlen, lpad, vec_stage, (int)tab_off, d_vpp, d_sc, d_c, d_p);
      });
      DCV_LAUNCH_CHECK("im2col_rows_kernel");
      return 0;
    }
  }
  DCV_DISPATCH_DTYPE(dtype, T, (im2col_kernel<T><<<grid_for(total, 256, kNumSMs * 32), 256, 0, as_stream(stream)>>>((const T*)x, (T*)col, *shape, kpad, d_vpp, d_q, d_p, d_sc, d_c)));
  DCV_LAUNCH_CHECK("im2col_kernel");
  return 0;
}

int dcv_gather_pack_weight(const void* w_krsc, void* w_col, int k, int r, int sc, int kpad, int dtype, void* stream) {
  using namespace dcv;
  const int rp = (sc + 7) / 8 * 8;
  DCV_REQUIRE(w_krsc && w_col && k > 0 && r > 0 && sc > 0 && kpad >= r * rp, "gather_pack_weight: bad arguments (kpad=%d must cover %d filter rows of %d)", kpad, r, rp);
  DCV_DISPATCH_DTYPE(dtype, T, (gather_pack_weight_kernel<T><<<grid_for((size_t)k * kpad, 256), 256, 0, as_stream(stream)>>>((const T*)w_krsc, (T*)w_col, k, r, sc, rp, kpad)));
  DCV_LAUNCH_CHECK("gather_pack_weight_kernel");
  return 0;
}

int dcv_gather_unpack_wgrad(const float* dw_col, float* dw_krsc, int k, int r, int sc, int kpad, void* stream) {
  using namespace dcv;
  const int rp = (sc + 7) / 8 * 8;
  DCV_REQUIRE(dw_col && dw_krsc && k > 0 && r > 0 && sc > 0 && kpad >= r * rp, "gather_unpack_wgrad: bad arguments");
  gather_unpack_wgrad_kernel<<<grid_for((size_t)k * r * sc, 256), 256, 0, as_stream(stream)>>>(dw_col, dw_krsc, k, r, sc, rp, kpad);
  DCV_LAUNCH_CHECK("gather_unpack_wgrad_kernel");
  return 0;
}

int dcv_pairs_pack_weight(const void* w_krsc, void* w_col, const dcv_conv_shape* shape, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(w_krsc && w_col && shape && shape->c >= 1 && shape->c <= 4 && shape->r <= 8 && shape->s + (shape->pad_w & 1) <= 8, "pairs_pack_weight: bad arguments");
  DCV_DISPATCH_DTYPE(dtype, T, (pairs_pack_weight_kernel<T><<<grid_for((size_t)shape->k * kPairsK, 256), 256, 0, as_stream(stream)>>>((const T*)w_krsc, (T*)w_col, shape->k, shape->r, shape->s, shape->c, shape->pad_w & 1)));
  DCV_LAUNCH_CHECK("pairs_pack_weight_kernel");
  return 0;
}

int dcv_pairs_unpack_wgrad(const float* dw_col, float* dw_krsc, const dcv_conv_shape* shape, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(dw_col && dw_krsc && shape && shape->c >= 1 && shape->c <= 4 && shape->r <= 8 && shape->s + (shape->pad_w & 1) <= 8, "pairs_unpack_wgrad: bad arguments");
  pairs_unpack_wgrad_kernel<<<grid_for((size_t)shape->k * shape->r * shape->s * shape->c, 256), 256, 0, as_stream(stream)>>>(dw_col, dw_krsc, shape->k, shape->r, shape->s, shape->c, shape->pad_w & 1);
  DCV_LAUNCH_CHECK("pairs_unpack_wgrad_kernel");
  return 0;
}

int dcv_gather_rows(const void* src, const int64_t* idx, void* dst, int n, long long n_src, size_t row_bytes, int* err_flag, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(src && idx && dst && n > 0 && n_src > 0, "gather_rows: bad arguments");
  if (row_bytes % 16 != 0 || reinterpret_cast<uintptr_t>(src) % 16 != 0 || reinterpret_cast<uintptr_t>(dst) % 16 != 0) {
    DCV_REQUIRE(row_bytes % 4 == 0 && reinterpret_cast<uintptr_t>(src) % 4 == 0 && reinterpret_cast<uintptr_t>(dst) % 4 == 0, "gather_rows: rows must be multiples of 4 bytes at 4-byte aligned bases");
    gather_rows_words_kernel<<<grid_for((size_t)n * (row_bytes / 4), 256), 256, 0, as_stream(stream)>>>((const uint32_t*)src, idx, (uint32_t*)dst, n, n_src, (uint32_t)(row_bytes / 4), err_flag);
    DCV_LAUNCH_CHECK("gather_rows_words_kernel");
    return 0;
  }
  const int wpb = 8;
  int grid = (n + wpb - 1) / wpb; if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  gather_rows_kernel<<<grid, wpb * 32, 0, as_stream(stream)>>>((const uint4*)src, idx, (uint4*)dst, n, n_src, (uint32_t)(row_bytes / 16), err_flag);
  DCV_LAUNCH_CHECK("gather_rows_kernel");
  return 0;
}

int dcv_fill_zero(void* dst, size_t bytes, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(dst || bytes == 0, "fill_zero: null pointer");
  if (bytes && cudaMemsetAsync(dst, 0, bytes, as_stream(stream)) != cudaSuccess) { set_error("fill_zero: cudaMemsetAsync failed"); (void)cudaGetLastError(); return 2; }
  return 0;
}

int dcv_pack_conv_weight(const float* w_krsc, void* dst, int dst_dtype, int k, int r, int s, int c, int transpose_flip, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(w_krsc && dst && k > 0 && r > 0 && s > 0 && c > 0, "pack_conv_weight: bad arguments");
  const size_t total = (size_t)k * r * s * c;
  DCV_DISPATCH_DTYPE(dst_dtype, T, (pack_weight_kernel<T><<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(w_krsc, (T*)dst, k, r, s, c, transpose_flip)));
  DCV_LAUNCH_CHECK("pack_weight_kernel");
  return 0;
}

int dcv_pack_conv_weights_batched(const float* flat_params, void* dst, int dst_dtype, const dcv_pack_entry* entries_dev, int n_entries, long long total_units, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(flat_params && dst && entries_dev && n_entries > 0 && n_entries <= 64 && total_units > 0, "pack_conv_weights_batched: bad arguments (at most 64 entries)");
  const int grid = grid_for((size_t)total_units * 32, 256, kNumSMs * 8);
  DCV_DISPATCH_DTYPE(dst_dtype, T, (pack_weights_batched_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(flat_params, (T*)dst, entries_dev, n_entries, total_units)));
  DCV_LAUNCH_CHECK("pack_weights_batched_kernel");
  return 0;
}

}  // extern "C"
