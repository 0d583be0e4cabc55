// Average pooling, link reductions (sum / mean / concat-by-slice) and bilinear rescaling of referenced tensors.
// All HBM-bound streaming kernels on NHWC tensors; 16-byte vectors over the channel dimension whenever it allows.
#include "common.cuh"

namespace dcv {

template <typename T, int VE>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c, int p, int q, int kh, int kw, int sh, int sw) {
  const int cv = c / VE;
  const size_t total = (size_t)n * p * q * cv;
  const float inv = 1.f / (float)(kh * kw);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    const int cc = t % cv; t /= cv;
    const int ox = t % q; t /= q;
    const int oy = t % p; const int img = t / p;
    float acc[VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[e] = 0.f;
    for (int r = 0; r < kh; ++r)
      for (int s = 0; s < kw; ++s) {
        const T* src = x + (((size_t)img * h + oy * sh + r) * w + ox * sw + s) * c + (size_t)cc * VE;
        float v[VE];
        if constexpr (VE == 1) v[0] = to_f<T>(*src); else vec_unpack<T>(*reinterpret_cast<const uint4*>(src), v);
#pragma unroll
        for (int e = 0; e < VE; ++e) acc[e] += v[e];
      }
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[e] *= inv;
    T* dst = y + idx * VE;
    if constexpr (VE == 1) *dst = from_f<T>(acc[0]); else *reinterpret_cast<uint4*>(dst) = vec_pack<T>(acc);
  }
}

template <typename T, int VE>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int n, int h, int w, int c, int p, int q, int kh, int kw, int sh, int sw) {
  const int cv = c / VE;
  const size_t total = (size_t)n * h * w * cv;
  const float inv = 1.f / (float)(kh * kw);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    const int cc = t % cv; t /= cv;
    const int ix = t % w; t /= w;
    const int iy = t % h; const int img = t / h;
    float acc[VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[e] = 0.f;
    const int oy_lo = iy - kh + 1 > 0 ? (iy - kh + 1 + sh - 1) / sh : 0, oy_hi = min(iy / sh, p - 1);
    const int ox_lo = ix - kw + 1 > 0 ? (ix - kw + 1 + sw - 1) / sw : 0, ox_hi = min(ix / sw, q - 1);
    for (int oy = oy_lo; oy <= oy_hi; ++oy)
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        const T* src = dy + (((size_t)img * p + oy) * q + ox) * c + (size_t)cc * VE;
        float v[VE];
        if constexpr (VE == 1) v[0] = to_f<T>(*src); else vec_unpack<T>(*reinterpret_cast<const uint4*>(src), v);
#pragma unroll
        for (int e = 0; e < VE; ++e) acc[e] += v[e];
      }
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[e] *= inv;
    T* dst = dx + idx * VE;
    if constexpr (VE == 1) *dst = from_f<T>(acc[0]); else *reinterpret_cast<uint4*>(dst) = vec_pack<T>(acc);
  }
}

// Non-overlapping windows that tile the input exactly (kernel == stride, h % kh == 0, w % kw == 0: the 2x2/2 pooling of the specs): one thread per
// OUTPUT-gradient vector reads it once and writes its kh x kw input positions — no per-input-pixel window search, dy is read once instead of kh*kw times.
template <typename T, int VE>
__global__ void avgpool_bwd_tiled_kernel(const T* __restrict__ dy, T* __restrict__ dx, int w, int c, int kh, int kw, const FastDiv div_cv, const FastDiv div_q, uint32_t total, int q) {
  const int cv = c / VE;
  const float inv = 1.f / (float)(kh * kw);
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t t = div_cv.div(idx);
    const int cc = (int)(idx - t * cv);
    const uint32_t row = div_q.div(t);           // row = img * p + oy
    const int ox = (int)(t - row * q);
    float v[VE];
    const T* src = dy + (size_t)idx * VE;
    if constexpr (VE == 1) v[0] = to_f<T>(*src); else vec_unpack<T>(*reinterpret_cast<const uint4*>(src), v);
#pragma unroll
    for (int e = 0; e < VE; ++e) v[e] *= inv;
    T* dst = dx + (((size_t)row * kh) * w + (size_t)ox * kw) * c + (size_t)cc * VE;   // (img*p + oy) * kh == img*h + oy*kh because h == p*kh
    for (int r = 0; r < kh; ++r)
      for (int s2 = 0; s2 < kw; ++s2) {
        T* d = dst + ((size_t)r * w + s2) * c;
        if constexpr (VE == 1) *d = from_f<T>(v[0]); else *reinterpret_cast<uint4*>(d) = vec_pack<T>(v);
      }
  }
}

template <typename T, int VE>
__global__ void axpby_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, float alpha, float beta, size_t nvec) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    float va[VE], vb[VE];
    if constexpr (VE == 1) va[0] = to_f<T>(a[i]); else vec_unpack<T>(reinterpret_cast<const uint4*>(a)[i], va);
    if (b) {
      if constexpr (VE == 1) vb[0] = to_f<T>(b[i]); else vec_unpack<T>(reinterpret_cast<const uint4*>(b)[i], vb);
#pragma unroll
      for (int e = 0; e < VE; ++e) va[e] = alpha * va[e] + beta * vb[e];
    } else {
#pragma unroll
      for (int e = 0; e < VE; ++e) va[e] = alpha * va[e];
    }
    if constexpr (VE == 1) out[i] = from_f<T>(va[0]); else reinterpret_cast<uint4*>(out)[i] = vec_pack<T>(va);
  }
}

// dst[pix][dst_off + j] = src[pix][src_off + j], j < cn  (element granularity E bytes: 16 when everything is 16B-aligned)
template <typename V>
__global__ void copy_channels_kernel(const V* __restrict__ src, V* __restrict__ dst, size_t pixels, int c_src, int src_off, int c_dst, int dst_off, int cn) {
  const size_t total = pixels * cn;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i / cn;
    const int j = (int)(i - pix * cn);
    dst[pix * c_dst + dst_off + j] = src[pix * c_src + src_off + j];
  }
}

static bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

// Dense link in ONE launch (`dense_link` = concat on the channel dim of the previous output and the `_from` tensors, a `_from`
// tensor twice the spatial size being the 2x2 average — what the bilinear rescale with align_corners=False is at an exact 2x reduction). One thread per
// (output pixel, channel granule of VE elements); the granule divides every source's channel count, so a thread never straddles two sources.
struct LinkSources {
  const void* ptr[DCV_LINK_MAX_SOURCES];
  int units[DCV_LINK_MAX_SOURCES];   // channels / VE
  int pool[DCV_LINK_MAX_SOURCES];    // 1 | 2
  int count, units_total;
};

template <typename T, int VE>
struct Granule {
  static __device__ __forceinline__ void load(const T* p, float (&v)[VE]) {
    if constexpr (VE * sizeof(T) == 16) vec_unpack<T>(*reinterpret_cast<const uint4*>(p), v);
    else {
#pragma unroll
      for (int e = 0; e < VE; ++e) v[e] = to_f<T>(p[e]);
    }
  }
  static __device__ __forceinline__ void store(T* p, const float (&v)[VE]) {
    if constexpr (VE * sizeof(T) == 16) *reinterpret_cast<uint4*>(p) = vec_pack<T>(v);
    else {
#pragma unroll
      for (int e = 0; e < VE; ++e) p[e] = from_f<T>(v[e]);
    }
  }
};

template <typename T, int VE, bool BWD>
__global__ void link_concat_kernel(const LinkSources src, T* __restrict__ cat, int h, int w, uint32_t total, const FastDiv div_units, const FastDiv div_w) {
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t pix = div_units.div(idx);
    int u = (int)(idx - pix * src.units_total), s = 0;
#pragma unroll
    for (int i = 0; i < DCV_LINK_MAX_SOURCES - 1; ++i)
      if (i + 1 < src.count && s == i && u >= src.units[i]) { u -= src.units[i]; s = i + 1; }
    T* t = (T*)src.ptr[s];
    if (BWD && t == nullptr) continue;
    const int c = src.units[s] * VE;
    T* cat_p = cat + (size_t)idx * VE;
    float v[VE];
    if (src.pool[s] == 1) {
      T* p = t + (size_t)pix * c + (size_t)u * VE;
      if constexpr (BWD) { Granule<T, VE>::load(cat_p, v); Granule<T, VE>::store(p, v); }
      else { Granule<T, VE>::load(p, v); Granule<T, VE>::store(cat_p, v); }
    } else {
      const uint32_t row = div_w.div(pix);       // row = img * h + y
      const int x = (int)(pix - row * w);
      T* p = t + (((size_t)row * 2) * (2 * w) + (size_t)x * 2) * c + (size_t)u * VE;   // (img*h + y)*2 == img*2h + 2y
      if constexpr (BWD) {
        Granule<T, VE>::load(cat_p, v);
#pragma unroll
        for (int e = 0; e < VE; ++e) v[e] *= 0.25f;
#pragma unroll
        for (int q = 0; q < 4; ++q) Granule<T, VE>::store(p + ((size_t)(q >> 1) * (2 * w) + (q & 1)) * c, v);
      } else {
        float v1[VE], v2[VE], v3[VE], acc[VE];   // (top pair) + (bottom pair), each pair left + right: the order torch's bilinear kernel adds them in
        Granule<T, VE>::load(p, v);
        Granule<T, VE>::load(p + c, v1);
        Granule<T, VE>::load(p + (size_t)(2 * w) * c, v2);
        Granule<T, VE>::load(p + (size_t)(2 * w + 1) * c, v3);
#pragma unroll
        for (int e = 0; e < VE; ++e) acc[e] = ((v[e] + v1[e]) + (v2[e] + v3[e])) * 0.25f;
        Granule<T, VE>::store(cat_p, acc);
      }
    }
  }
}

template <typename T, int VE, bool BWD>
static int launch_link(const dcv_link_source* sources, int count, int ctot, void* cat, int n, int h, int w, cudaStream_t st) {
  LinkSources ls{};
  ls.count = count; ls.units_total = ctot / VE;
  for (int i = 0; i < count; ++i) { ls.ptr[i] = sources[i].ptr; ls.units[i] = sources[i].channels / VE; ls.pool[i] = sources[i].pool; }
  const uint32_t total = (uint32_t)((size_t)n * h * w * ls.units_total);
  link_concat_kernel<T, VE, BWD><<<grid_for(total, 256), 256, 0, st>>>(ls, (T*)cat, h, w, total, FastDiv(ls.units_total), FastDiv(w));
  DCV_LAUNCH_CHECK("link_concat_kernel");
  return 0;
}

template <bool BWD>
static int link_concat(const char* name, const dcv_link_source* sources, int count, void* cat, int n, int h, int w, int dtype, cudaStream_t st) {
  DCV_REQUIRE(sources && cat && count > 0 && count <= DCV_LINK_MAX_SOURCES && n > 0 && h > 0 && w > 0, "%s: bad arguments", name);
  DCV_REQUIRE(dtype == DCV_F32 || dtype == DCV_BF16, "%s: unsupported dtype %d", name, dtype);
  const int es = dtype == DCV_BF16 ? 2 : 4;
  int ctot = 0, ve = 16 / es;
  bool al16 = aligned16(cat);
  for (int i = 0; i < count; ++i) {
    DCV_REQUIRE(sources[i].channels > 0 && (sources[i].pool == 1 || sources[i].pool == 2) && (BWD || sources[i].ptr), "%s: bad source %d", name, i);
    ctot += sources[i].channels;
    while (sources[i].channels % ve) ve >>= 1;
    al16 = al16 && (!sources[i].ptr || aligned16(sources[i].ptr));
  }
  DCV_REQUIRE((size_t)n * h * w * ctot < (1ull << 31), "%s: tensor too large", name);
  if (ve * es == 16 && !al16) ve >>= 1;
#define DCV_LINK_CASE(T, V) if (ve == V) return launch_link<T, V, BWD>(sources, count, ctot, cat, n, h, w, st)
  if (dtype == DCV_BF16) { DCV_LINK_CASE(__nv_bfloat16, 8); DCV_LINK_CASE(__nv_bfloat16, 4); DCV_LINK_CASE(__nv_bfloat16, 2); DCV_LINK_CASE(__nv_bfloat16, 1); }
  else { DCV_LINK_CASE(float, 4); DCV_LINK_CASE(float, 2); DCV_LINK_CASE(float, 1); }
#undef DCV_LINK_CASE
  return 1;
}

// torch upsample_bilinear2d source index / weights (aten/native/UpSample.h: area_pixel_compute_source_index)
__device__ __forceinline__ void bilinear_coord(int dst, float scale, int in_size, int align_corners, int& i0, int& i1, float& l0, float& l1) {
  float src = align_corners ? scale * (float)dst : scale * ((float)dst + 0.5f) - 0.5f;
  if (!align_corners && src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}
__host__ __device__ inline float bilinear_scale(int in_size, int out_size, int align_corners) {
  if (align_corners) return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
  return (float)in_size / (float)out_size;
}

template <typename T>
__global__ void bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int n, int h, int w, int c, int oh, int ow, int align, float sh, float sw) {
  const size_t total = (size_t)n * oh * ow * c;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    const int ch = t % c; t /= c;
    const int ox = t % ow; t /= ow;
    const int oy = t % oh; const int img = t / oh;
    int y0, y1, x0, x1; float hl0, hl1, wl0, wl1;
    bilinear_coord(oy, sh, h, align, y0, y1, hl0, hl1);
    bilinear_coord(ox, sw, w, align, x0, x1, wl0, wl1);
    const T* base = x + (size_t)img * h * w * c + ch;
    const float v00 = to_f<T>(base[((size_t)y0 * w + x0) * c]), v01 = to_f<T>(base[((size_t)y0 * w + x1) * c]);
    const float v10 = to_f<T>(base[((size_t)y1 * w + x0) * c]), v11 = to_f<T>(base[((size_t)y1 * w + x1) * c]);
    y[idx] = from_f<T>(hl0 * (wl0 * v00 + wl1 * v01) + hl1 * (wl0 * v10 + wl1 * v11));
  }
}

template <typename T>
__global__ void bilinear_bwd_kernel(const T* __restrict__ dy, float* __restrict__ dx, int n, int h, int w, int c, int oh, int ow, int align, float sh, float sw) {
  const size_t total = (size_t)n * oh * ow * c;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    const int ch = t % c; t /= c;
    const int ox = t % ow; t /= ow;
    const int oy = t % oh; const int img = t / oh;
    int y0, y1, x0, x1; float hl0, hl1, wl0, wl1;
    bilinear_coord(oy, sh, h, align, y0, y1, hl0, hl1);
    bilinear_coord(ox, sw, w, align, x0, x1, wl0, wl1);
    const float g = to_f<T>(dy[idx]);
    float* base = dx + (size_t)img * h * w * c + ch;
    atomicAdd(base + ((size_t)y0 * w + x0) * c, hl0 * wl0 * g);
    atomicAdd(base + ((size_t)y0 * w + x1) * c, hl0 * wl1 * g);
    atomicAdd(base + ((size_t)y1 * w + x0) * c, hl1 * wl0 * g);
    atomicAdd(base + ((size_t)y1 * w + x1) * c, hl1 * wl1 * g);
  }
}

static int pool_check(const char* name, int n, int h, int w, int c, int kh, int kw, int sh, int sw, int* p, int* q) {
  DCV_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0, "%s: bad arguments", name);
  DCV_REQUIRE(h >= kh && w >= kw, "%s: kernel %dx%d larger than input %dx%d", name, kh, kw, h, w);
  *p = (h - kh) / sh + 1; *q = (w - kw) / sw + 1;
  return 0;
}

template <typename V>
static int launch_copy(const void* src, void* dst, size_t pixels, int c_src, int src_off, int c_dst, int dst_off, int cn, cudaStream_t st) {
  copy_channels_kernel<V><<<grid_for(pixels * cn, 256), 256, 0, st>>>((const V*)src, (V*)dst, pixels, c_src, src_off, c_dst, dst_off, cn);
  DCV_LAUNCH_CHECK("copy_channels_kernel");
  return 0;
}

static int copy_channels(const void* src, void* dst, size_t pixels, int c_src, int src_off, int c_dst, int dst_off, int cn, int dtype, cudaStream_t st) {
  DCV_REQUIRE(src && dst && pixels > 0 && cn > 0 && src_off >= 0 && dst_off >= 0 && src_off + cn <= c_src && dst_off + cn <= c_dst, "copy_channels: bad arguments");
  DCV_REQUIRE(dtype == DCV_F32 || dtype == DCV_BF16, "copy_channels: unsupported dtype %d", dtype);
  const int es = dtype == DCV_BF16 ? 2 : 4, ve = 16 / es;
  if (c_src % ve == 0 && c_dst % ve == 0 && src_off % ve == 0 && dst_off % ve == 0 && cn % ve == 0 && aligned16(src) && aligned16(dst))
    return launch_copy<uint4>(src, dst, pixels, c_src / ve, src_off / ve, c_dst / ve, dst_off / ve, cn / ve, st);
  if (es == 2) return launch_copy<uint16_t>(src, dst, pixels, c_src, src_off, c_dst, dst_off, cn, st);
  return launch_copy<uint32_t>(src, dst, pixels, c_src, src_off, c_dst, dst_off, cn, st);
}

}  // namespace dcv

extern "C" {

int dcv_avgpool2d_fwd(const void* x, void* y, int n, int h, int w, int c, int kh, int kw, int sh, int sw, int dtype, void* stream) {
  using namespace dcv;
  int p, q;
  DCV_REQUIRE(x && y, "avgpool2d_fwd: null pointer");
  if (pool_check("avgpool2d_fwd", n, h, w, c, kh, kw, sh, sw, &p, &q)) return 1;
  cudaStream_t st = as_stream(stream);
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    if (c % VE == 0 && aligned16(x) && aligned16(y)) avgpool_fwd_kernel<T, VE><<<grid_for((size_t)n * p * q * (c / VE), 256), 256, 0, st>>>((const T*)x, (T*)y, n, h, w, c, p, q, kh, kw, sh, sw);
    else avgpool_fwd_kernel<T, 1><<<grid_for((size_t)n * p * q * c, 256), 256, 0, st>>>((const T*)x, (T*)y, n, h, w, c, p, q, kh, kw, sh, sw);
  });
  DCV_LAUNCH_CHECK("avgpool_fwd_kernel");
  return 0;
}

int dcv_avgpool2d_bwd(const void* dy, void* dx, int n, int h, int w, int c, int kh, int kw, int sh, int sw, int dtype, void* stream) {
  using namespace dcv;
  int p, q;
  DCV_REQUIRE(dy && dx, "avgpool2d_bwd: null pointer");
  if (pool_check("avgpool2d_bwd", n, h, w, c, kh, kw, sh, sw, &p, &q)) return 1;
  cudaStream_t st = as_stream(stream);
  if (kh == sh && kw == sw && h % kh == 0 && w % kw == 0 && (size_t)n * p * q * c < (1ull << 31)) {
    DCV_DISPATCH_DTYPE(dtype, T, {
      constexpr int VE = 16 / sizeof(T);
      if (c % VE == 0 && aligned16(dy) && aligned16(dx)) {
        const uint32_t total = (uint32_t)((size_t)n * p * q * (c / VE));
        avgpool_bwd_tiled_kernel<T, VE><<<grid_for(total, 256), 256, 0, st>>>((const T*)dy, (T*)dx, w, c, kh, kw, FastDiv(c / VE), FastDiv(q), total, q);
      } else {
        const uint32_t total = (uint32_t)((size_t)n * p * q * c);
        avgpool_bwd_tiled_kernel<T, 1><<<grid_for(total, 256), 256, 0, st>>>((const T*)dy, (T*)dx, w, c, kh, kw, FastDiv(c), FastDiv(q), total, q);
      }
    });
    DCV_LAUNCH_CHECK("avgpool_bwd_tiled_kernel");
    return 0;
  }
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    if (c % VE == 0 && aligned16(dy) && aligned16(dx)) avgpool_bwd_kernel<T, VE><<<grid_for((size_t)n * h * w * (c / VE), 256), 256, 0, st>>>((const T*)dy, (T*)dx, n, h, w, c, p, q, kh, kw, sh, sw);
    else avgpool_bwd_kernel<T, 1><<<grid_for((size_t)n * h * w * c, 256), 256, 0, st>>>((const T*)dy, (T*)dx, n, h, w, c, p, q, kh, kw, sh, sw);
  });
  DCV_LAUNCH_CHECK("avgpool_bwd_kernel");
  return 0;
}

int dcv_axpby(const void* a, const void* b, void* out, float alpha, float beta, size_t count, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(a && out, "axpby: null pointer");
  if (count == 0) return 0;
  cudaStream_t st = as_stream(stream);
  DCV_DISPATCH_DTYPE(dtype, T, {
    constexpr int VE = 16 / sizeof(T);
    if (count % VE == 0 && aligned16(a) && aligned16(out) && (!b || aligned16(b))) axpby_kernel<T, VE><<<grid_for(count / VE, 256), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, alpha, beta, count / VE);
    else axpby_kernel<T, 1><<<grid_for(count, 256), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, alpha, beta, count);
  });
  DCV_LAUNCH_CHECK("axpby_kernel");
  return 0;
}

int dcv_copy_channels_in(const void* src, void* dst, size_t pixels, int c_src, int c_dst, int c_off, int dtype, void* stream) {
  return dcv::copy_channels(src, dst, pixels, c_src, 0, c_dst, c_off, c_src, dtype, dcv::as_stream(stream));
}

int dcv_copy_channels_out(const void* src, void* dst, size_t pixels, int c_src, int c_off, int c_dst, int dtype, void* stream) {
  return dcv::copy_channels(src, dst, pixels, c_src, c_off, c_dst, 0, c_dst, dtype, dcv::as_stream(stream));
}

int dcv_link_concat_fwd(const dcv_link_source* sources, int count, void* out, int n, int h, int w, int dtype, void* stream) {
  return dcv::link_concat<false>("link_concat_fwd", sources, count, out, n, h, w, dtype, dcv::as_stream(stream));
}

int dcv_link_concat_bwd(const void* dout, const dcv_link_source* grads, int count, int n, int h, int w, int dtype, void* stream) {
  return dcv::link_concat<true>("link_concat_bwd", grads, count, const_cast<void*>(dout), n, h, w, dtype, dcv::as_stream(stream));
}

int dcv_bilinear_fwd(const void* x, void* y, int n, int h, int w, int c, int oh, int ow, int align_corners, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "bilinear_fwd: bad arguments");
  const float sh = bilinear_scale(h, oh, align_corners), sw = bilinear_scale(w, ow, align_corners);
  DCV_DISPATCH_DTYPE(dtype, T, (bilinear_fwd_kernel<T><<<grid_for((size_t)n * oh * ow * c, 256), 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, n, h, w, c, oh, ow, align_corners, sh, sw)));
  DCV_LAUNCH_CHECK("bilinear_fwd_kernel");
  return 0;
}

int dcv_bilinear_bwd(const void* dy, float* dx_f32, int n, int h, int w, int c, int oh, int ow, int align_corners, int dtype, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(dy && dx_f32 && n > 0 && h > 0 && w > 0 && c > 0 && oh > 0 && ow > 0, "bilinear_bwd: bad arguments");
  const float sh = bilinear_scale(h, oh, align_corners), sw = bilinear_scale(w, ow, align_corners);
  cudaMemsetAsync(dx_f32, 0, (size_t)n * h * w * c * sizeof(float), as_stream(stream));
  DCV_DISPATCH_DTYPE(dtype, T, (bilinear_bwd_kernel<T><<<grid_for((size_t)n * oh * ow * c, 256), 256, 0, as_stream(stream)>>>((const T*)dy, dx_f32, n, h, w, c, oh, ow, align_corners, sh, sw)));
  DCV_LAUNCH_CHECK("bilinear_bwd_kernel");
  return 0;
}

}  // extern "C"
