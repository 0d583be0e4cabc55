// Direct (CUDA-core, fp32-accumulate) convolution kernels: forward, data gradient and weight gradient.
//
// These serve (a) the fp32 parity mode (1e-4 against the CPU oracle, which tensor cores at bf16/tf32 cannot give) and
// (b) layers whose channel counts are too small for an MMA tile to pay (the default CIFAR net: 3/4/16 channels,
// K = 36..144, arithmetic intensity 43-72 FLOP/B => HBM/latency bound, not tensor bound).
// Large-channel bf16 layers go to the tcgen05 implicit-GEMM kernels in conv_tc.cu.
//
// Forward tiling: a CTA owns a TW x TH tile of output pixels of one image and KC output channels; input-channel chunks
// of CC are staged in shared memory as [cc][iy][ix] planes (lanes walk ix => conflict-free) next to the matching
// [r][s][cc][KC] weight slab (read as broadcast float4). Each thread accumulates PX pixels x KC channels in registers.
// The epilogue fuses bias + activation + the per-(image, channel) sum / sum-of-squares that BatchNorm / GroupNorm need.
#include "common.cuh"
#include <stdlib.h>

namespace dcv {

struct DirectConvArgs {
  dcv_conv_shape s;
  const void* x; const void* w; const float* bias; void* y; float* stats;
  int act; float slope;
  int transposed;      // weights are read as W'[k'][r'][s'][c'] = w[c'][R-1-r'][S-1-s'][k'] (data-gradient operand)
  int cc;              // input channels per shared-memory chunk
  int in_th, in_tw, in_pitch;
  int tiles_x;
  int vec4;            // input channels are a multiple of 4 and 4-channel vector loads are aligned
};

template <typename T> __device__ __forceinline__ void load4(const T* p, float* v);
template <> __device__ __forceinline__ void load4<float>(const float* p, float* v) { const float4 f = *reinterpret_cast<const float4*>(p); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u); v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
}

template <int BX, int BY, int PX> struct Tile { static constexpr int TW = BX * PX, TH = BY, THREADS = BX * BY; };

template <typename T, int KC, int BX, int BY, int PX>
__global__ void __launch_bounds__(BX * BY) conv_fwd_direct_kernel(const DirectConvArgs a) {
  extern __shared__ __align__(16) float smem[];
  const dcv_conv_shape& s = a.s;
  constexpr int TW = BX * PX, TH = BY, NT = BX * BY;
  float* s_in = smem;                                            // [cc][in_th][in_pitch]
  float* s_w = smem + (size_t)a.cc * a.in_th * a.in_pitch;       // [r][s][cc][KC]
  const int tid = threadIdx.x, tx = tid % BX, ty = tid / BX;
  const int tile = blockIdx.x, tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
  const int img = blockIdx.y, k0 = blockIdx.z * KC;
  const int oy0 = tile_y * TH, ox0 = tile_x * TW;
  const int iy0 = oy0 * s.stride_h - s.pad_h, ix0 = ox0 * s.stride_w - s.pad_w;
  const T* xin = reinterpret_cast<const T*>(a.x) + (size_t)img * s.h * s.w * s.c;
  const T* wgt = reinterpret_cast<const T*>(a.w);

  float acc[PX][KC];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[p][k] = 0.f;

  for (int c0 = 0; c0 < s.c; c0 += a.cc) {
    const int ccn = min(a.cc, s.c - c0);
    __syncthreads();
    // stage input planes (zero outside the image: convolution padding)
    const int plane = a.in_th * a.in_tw;
    for (int i = tid; i < plane * ccn; i += NT) {
      // channel fastest in the global read (NHWC), plane-major in shared memory
      const int cc = i % ccn, pix = i / ccn;
      const int iy = pix / a.in_tw, ix = pix - iy * a.in_tw;
      const int gy = iy0 + iy, gx = ix0 + ix;
      float v = 0.f;
      if (gy >= 0 && gy < s.h && gx >= 0 && gx < s.w) v = to_f<T>(xin[((size_t)gy * s.w + gx) * s.c + c0 + cc]);
      s_in[(cc * a.in_th + iy) * a.in_pitch + ix] = v;
    }
    // stage weights [r][s][cc][KC]
    const int wcount = s.r * s.s * ccn * KC;
    for (int i = tid; i < wcount; i += NT) {
      const int kk = i % KC;
      int t = i / KC;
      const int cc = t % ccn; t /= ccn;
      const int ss = t % s.s, rr = t / s.s;
      float v = 0.f;
      if (k0 + kk < s.k) {
        size_t src;
        if (!a.transposed) src = (((size_t)(k0 + kk) * s.r + rr) * s.s + ss) * s.c + c0 + cc;
        else src = (((size_t)(c0 + cc) * s.r + (s.r - 1 - rr)) * s.s + (s.s - 1 - ss)) * s.k + k0 + kk;  // w is [C'=s.c][R][S][K'=s.k]
        v = to_f<T>(wgt[src]);
      }
      s_w[((rr * s.s + ss) * a.cc + cc) * KC + kk] = v;
    }
    __syncthreads();
    for (int cc = 0; cc < ccn; ++cc) {
      const float* pin = s_in + (size_t)cc * a.in_th * a.in_pitch + (ty * s.stride_h) * a.in_pitch + tx * s.stride_w;
      for (int rr = 0; rr < s.r; ++rr) {
        const float* prow = pin + rr * s.dil_h * a.in_pitch;
        for (int ss = 0; ss < s.s; ++ss) {
          const float* wp = s_w + ((rr * s.s + ss) * a.cc + cc) * KC;
          float wv[KC];
          if constexpr (KC % 4 == 0) {
#pragma unroll
            for (int k = 0; k < KC; k += 4) {
              const float4 f = *reinterpret_cast<const float4*>(wp + k);
              wv[k] = f.x; wv[k + 1] = f.y; wv[k + 2] = f.z; wv[k + 3] = f.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < KC; ++k) wv[k] = wp[k];
          }
#pragma unroll
          for (int p = 0; p < PX; ++p) {
            const float v = prow[(p * BX) * s.stride_w + ss * s.dil_w];
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[p][k] = fmaf(v, wv[k], acc[p][k]);
          }
        }
      }
    }
  }

  // ---- epilogue: bias + activation, store NHWC, fused statistics
  T* yout = reinterpret_cast<T*>(a.y) + (size_t)img * s.p * s.q * s.k;
  float ssum[KC], ssq[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) ssum[k] = ssq[k] = 0.f;
  const int oy = oy0 + ty;
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int ox = ox0 + p * BX + tx;
    if (oy < s.p && ox < s.q) {
      T* dst = yout + ((size_t)oy * s.q + ox) * s.k + k0;
      float out[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        float v = acc[p][k] + ((a.bias && k0 + k < s.k) ? a.bias[k0 + k] : 0.f);
        v = act_apply(v, a.act, a.slope);
        v = to_f<T>(from_f<T>(v));   // statistics are taken on the value as stored
        out[k] = v;
        if (k0 + k < s.k) { ssum[k] += v; ssq[k] = fmaf(v, v, ssq[k]); }
      }
      constexpr int VE = 16 / sizeof(T);
      if (KC % VE == 0 && k0 + KC <= s.k && (s.k % VE == 0) && (reinterpret_cast<uintptr_t>(a.y) % 16 == 0)) {
#pragma unroll
        for (int k = 0; k < KC; k += VE) *reinterpret_cast<uint4*>(dst + k) = vec_pack<T>(out + k);
      } else {
#pragma unroll
        for (int k = 0; k < KC; ++k) if (k0 + k < s.k) dst[k] = from_f<T>(out[k]);
      }
    }
  }
  if (a.stats) {
    __shared__ float s_red[NT / 32][KC * 2];
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float v1 = warp_sum(ssum[k]), v2 = warp_sum(ssq[k]);
      if (lane == 0) { s_red[warp][2 * k] = v1; s_red[warp][2 * k + 1] = v2; }
    }
    __syncthreads();
    if (tid < KC * 2 && k0 + tid / 2 < s.k) {
      float v = 0.f;
#pragma unroll
      for (int wv = 0; wv < NT / 32; ++wv) v += s_red[wv][tid];
      atomicAdd(a.stats + ((size_t)img * s.k + k0) * 2 + tid, v);
    }
  }
}


// Stride-1, dilation-1 specialisation with a compile-time filter width S: a thread owns PX *consecutive* output pixels of one row and KC output
// channels. Per (input channel, filter row) it loads its PX+S-1 input values once (128-bit shared loads, conflict-free when pitch/4 is odd) and
// slides the S taps over them in registers: S*PX*KC FMAs for (PX+S-1)/4 + S*KC/4 shared loads, i.e. FMA-bound instead of LDS-bound.
template <typename T, int KC, int BX, int BY, int PX, int S>
__global__ void __launch_bounds__(BX * BY) conv_fwd_direct_s1_kernel(const DirectConvArgs a) {
  extern __shared__ __align__(16) float smem[];
  const dcv_conv_shape& s = a.s;
  constexpr int TW = BX * PX, TH = BY, NT = BX * BY, NIN = PX + S - 1, NV = (NIN + 3) / 4;
  float* s_in = smem;                                                  // [cc][in_th][in_pitch] (+4 floats of slack)
  float* s_w = smem + (size_t)a.cc * a.in_th * a.in_pitch + 4;         // [r][S][cc][KC]
  const int tid = threadIdx.x, tx = tid % BX, ty = tid / BX;
  const int tile = blockIdx.x, tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
  const int img = blockIdx.y, k0 = blockIdx.z * KC;
  const int oy0 = tile_y * TH, ox0 = tile_x * TW;
  const int iy0 = oy0 - s.pad_h, ix0 = ox0 - s.pad_w;
  const T* xin = reinterpret_cast<const T*>(a.x) + (size_t)img * s.h * s.w * s.c;
  const T* wgt = reinterpret_cast<const T*>(a.w);

  float acc[PX][KC];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[p][k] = 0.f;

  for (int c0 = 0; c0 < s.c; c0 += a.cc) {
    const int ccn = min(a.cc, s.c - c0);
    __syncthreads();
    const int plane = a.in_th * a.in_tw;
    if (a.vec4) {
      // 4 channels of a pixel per load (8 B of bf16 / 16 B of fp32): one index computation per pixel instead of per element
      const int groups = ccn >> 2;
      for (int i = tid; i < plane * groups; i += NT) {
        const int g4 = i % groups, pix = i / groups;
        const int iy = pix / a.in_tw, ix = pix - iy * a.in_tw;
        const int gy = iy0 + iy, gx = ix0 + ix;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gy >= 0 && gy < s.h && gx >= 0 && gx < s.w) load4<T>(xin + ((size_t)gy * s.w + gx) * s.c + c0 + 4 * g4, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) s_in[((4 * g4 + e) * a.in_th + iy) * a.in_pitch + ix] = v[e];
      }
    } else {
      for (int i = tid; i < plane * ccn; i += NT) {
        const int cc = i % ccn, pix = i / ccn;
        const int iy = pix / a.in_tw, ix = pix - iy * a.in_tw;
        const int gy = iy0 + iy, gx = ix0 + ix;
        float v = 0.f;
        if (gy >= 0 && gy < s.h && gx >= 0 && gx < s.w) v = to_f<T>(xin[((size_t)gy * s.w + gx) * s.c + c0 + cc]);
        s_in[(cc * a.in_th + iy) * a.in_pitch + ix] = v;
      }
    }
    const int wcount = s.r * S * ccn * KC;
    for (int i = tid; i < wcount; i += NT) {
      const int kk = i % KC;
      int t = i / KC;
      const int cc = t % ccn; t /= ccn;
      const int ss = t % S, rr = t / S;
      float v = 0.f;
      if (k0 + kk < s.k) {
        size_t src;
        if (!a.transposed) src = (((size_t)(k0 + kk) * s.r + rr) * S + ss) * s.c + c0 + cc;
        else src = (((size_t)(c0 + cc) * s.r + (s.r - 1 - rr)) * S + (S - 1 - ss)) * s.k + k0 + kk;
        v = to_f<T>(wgt[src]);
      }
      s_w[((rr * S + ss) * a.cc + cc) * KC + kk] = v;
    }
    __syncthreads();
    for (int cc = 0; cc < ccn; ++cc) {
      const float* pin = s_in + ((size_t)cc * a.in_th + ty) * a.in_pitch + tx * PX;
      for (int rr = 0; rr < s.r; ++rr) {
        float in[NV * 4];
        const float4* prow = reinterpret_cast<const float4*>(pin + rr * a.in_pitch);
#pragma unroll
        for (int v = 0; v < NV; ++v) { const float4 f = prow[v]; in[4 * v] = f.x; in[4 * v + 1] = f.y; in[4 * v + 2] = f.z; in[4 * v + 3] = f.w; }
#pragma unroll
        for (int ss = 0; ss < S; ++ss) {
          const float* wp = s_w + ((rr * S + ss) * a.cc + cc) * KC;
          float wv[KC];
#pragma unroll
          for (int k = 0; k < KC; k += 4) { const float4 f = *reinterpret_cast<const float4*>(wp + k); wv[k] = f.x; wv[k + 1] = f.y; wv[k + 2] = f.z; wv[k + 3] = f.w; }
#pragma unroll
          for (int p = 0; p < PX; ++p)
#pragma unroll
            for (int k = 0; k < KC; ++k) acc[p][k] = fmaf(in[p + ss], wv[k], acc[p][k]);
        }
      }
    }
  }

  T* yout = reinterpret_cast<T*>(a.y) + (size_t)img * s.p * s.q * s.k;
  float ssum[KC], ssq[KC];
#pragma unroll
  for (int k = 0; k < KC; ++k) ssum[k] = ssq[k] = 0.f;
  const int oy = oy0 + ty;
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int ox = ox0 + tx * PX + p;
    if (oy < s.p && ox < s.q) {
      T* dst = yout + ((size_t)oy * s.q + ox) * s.k + k0;
      float out[KC];
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        float v = acc[p][k] + ((a.bias && k0 + k < s.k) ? a.bias[k0 + k] : 0.f);
        v = act_apply(v, a.act, a.slope);
        v = to_f<T>(from_f<T>(v));
        out[k] = v;
        if (k0 + k < s.k) { ssum[k] += v; ssq[k] = fmaf(v, v, ssq[k]); }
      }
      constexpr int VE = 16 / sizeof(T);
      if (KC % VE == 0 && k0 + KC <= s.k && (s.k % VE == 0) && (reinterpret_cast<uintptr_t>(a.y) % 16 == 0)) {
#pragma unroll
        for (int k = 0; k < KC; k += VE) *reinterpret_cast<uint4*>(dst + k) = vec_pack<T>(out + k);
      } else if (sizeof(T) == 2 && KC % 4 == 0 && k0 + KC <= s.k && (s.k % 4 == 0) && (reinterpret_cast<uintptr_t>(a.y) % 8 == 0)) {
#pragma unroll
        for (int k = 0; k < KC; k += 4) {   // 4 bf16 = 8 bytes
          __nv_bfloat162 lo = __floats2bfloat162_rn(out[k], out[k + 1]), hi = __floats2bfloat162_rn(out[k + 2], out[k + 3]);
          *reinterpret_cast<uint2*>(dst + k) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
      } else {
#pragma unroll
        for (int k = 0; k < KC; ++k) if (k0 + k < s.k) dst[k] = from_f<T>(out[k]);
      }
    }
  }
  if (a.stats) {
    __shared__ float s_red[(NT + 31) / 32][KC * 2];
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const float v1 = warp_sum(ssum[k]), v2 = warp_sum(ssq[k]);
      if (lane == 0) { s_red[warp][2 * k] = v1; s_red[warp][2 * k + 1] = v2; }
    }
    __syncthreads();
    if (tid < KC * 2 && k0 + tid / 2 < s.k) {
      float v = 0.f;
#pragma unroll
      for (int wv = 0; wv < (NT + 31) / 32; ++wv) v += s_red[wv][tid];
      atomicAdd(a.stats + ((size_t)img * s.k + k0) * 2 + tid, v);
    }
  }
}

// Data gradient for strided convolutions (never the hot case: only a strided layer that is not first needs it).
template <typename T>
__global__ void conv_dgrad_generic_kernel(const dcv_conv_shape s, const T* __restrict__ dy, const T* __restrict__ w, T* __restrict__ dx) {
  const size_t total = (size_t)s.n * s.h * s.w * s.c;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    size_t t = idx;
    const int ci = t % s.c; t /= s.c;
    const int ix = t % s.w; t /= s.w;
    const int iy = t % s.h; const int img = t / s.h;
    float acc = 0.f;
    for (int rr = 0; rr < s.r; ++rr) {
      const int ny = iy + s.pad_h - rr * s.dil_h;
      if (ny < 0 || ny % s.stride_h) continue;
      const int oy = ny / s.stride_h;
      if (oy >= s.p) continue;
      for (int ss = 0; ss < s.s; ++ss) {
        const int nx = ix + s.pad_w - ss * s.dil_w;
        if (nx < 0 || nx % s.stride_w) continue;
        const int ox = nx / s.stride_w;
        if (ox >= s.q) continue;
        const T* dyp = dy + (((size_t)img * s.p + oy) * s.q + ox) * s.k;
        const T* wp = w + ((size_t)rr * s.s + ss) * s.c + ci;
        for (int k = 0; k < s.k; ++k) acc = fmaf(to_f<T>(dyp[k]), to_f<T>(wp[(size_t)k * s.r * s.s * s.c]), acc);
      }
    }
    dx[idx] = from_f<T>(acc);
  }
}

// ---- weight gradient --------------------------------------------------------------------------------------------------
// A CTA owns one (16 k x 16 c x tap-chunk) slab of dw and one spatial tile position; it walks images n = blockIdx.y,
// +gridDim.y, ... staging dy[pix][16] and x[ipix][16] tiles (channel-innermost => float4 reads). A thread accumulates 4k x 4c
// register blocks for one filter tap over its share of the tile's pixels. Partials are combined through shared memory and
// leave the CTA as one atomicAdd per dw element (dw is zeroed by the host wrapper).
struct WgradArgs {
  dcv_conv_shape s;
  const void* x; const void* dy; float* dw;
  int in_th, in_tw, tiles_x;
  int kslabs, cslabs, tap_chunk;
  int vec_dy, vec_x;   // channel counts are multiples of 4 and the tensors are aligned for 4-channel vector loads
};

template <typename T, int TW, int TH>
__global__ void __launch_bounds__(256) conv_wgrad_direct_kernel(const WgradArgs a) {
  constexpr int KC = 16, CC = 16, NT = 256, MAXPASS = 4, NPIX = TW * TH;
  extern __shared__ __align__(16) float smem[];
  const dcv_conv_shape& s = a.s;
  float* s_dy = smem;                 // [NPIX][KC]
  float* s_x = smem + NPIX * KC;      // [in_th*in_tw][CC]
  const int tid = threadIdx.x;
  const int tile = blockIdx.x, tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
  int z = blockIdx.z;
  const int k0 = (z % a.kslabs) * KC; z /= a.kslabs;
  const int c0 = (z % a.cslabs) * CC; z /= a.cslabs;
  const int tap0 = z * a.tap_chunk;
  const int ntaps = min(a.tap_chunk, s.r * s.s - tap0);
  const int kq_n = (min(KC, s.k - k0) + 3) / 4, cq_n = (min(CC, s.c - c0) + 3) / 4;
  const int groups = kq_n * cq_n * ntaps;
  const int parts = groups >= NT ? 1 : NT / groups;
  const int oy0 = tile_y * TH, ox0 = tile_x * TW;
  const int iy0 = oy0 * s.stride_h - s.pad_h, ix0 = ox0 * s.stride_w - s.pad_w;
  const int in_pix = a.in_th * a.in_tw;

  // per-pass register block description
  int g_kq[MAXPASS], g_cq[MAXPASS], g_off[MAXPASS];
  bool g_on[MAXPASS];
  const int part = groups >= NT ? 0 : tid / groups;
#pragma unroll
  for (int ps = 0; ps < MAXPASS; ++ps) {
    const int g = (groups >= NT ? tid : tid % groups) + ps * NT;
    g_on[ps] = (g < groups) && (part < parts) && (groups >= NT || ps == 0);
    const int gg = g_on[ps] ? g : 0;
    const int tap = tap0 + gg / (kq_n * cq_n), rem = gg % (kq_n * cq_n);
    g_kq[ps] = rem / cq_n; g_cq[ps] = rem % cq_n;
    const int rr = tap / s.s, ss = tap - rr * s.s;
    g_off[ps] = (rr * s.dil_h) * a.in_tw + ss * s.dil_w;
  }
  float acc[MAXPASS][16];
#pragma unroll
  for (int ps = 0; ps < MAXPASS; ++ps)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[ps][i] = 0.f;

  const T* xall = reinterpret_cast<const T*>(a.x);
  const T* dyall = reinterpret_cast<const T*>(a.dy);
  for (int img = blockIdx.y; img < s.n; img += gridDim.y) {
    __syncthreads();
    const T* dyp = dyall + (size_t)img * s.p * s.q * s.k;
    for (int i = tid; i < NPIX * KC; i += NT) {
      const int kk = i % KC, pix = i / KC;
      const int oy = oy0 + pix / TW, ox = ox0 + pix % TW;
      float v = 0.f;
      if (oy < s.p && ox < s.q && k0 + kk < s.k) v = to_f<T>(dyp[((size_t)oy * s.q + ox) * s.k + k0 + kk]);
      s_dy[i] = v;
    }
    const T* xp = xall + (size_t)img * s.h * s.w * s.c;
    for (int i = tid; i < in_pix * CC; i += NT) {
      const int cc = i % CC, pix = i / CC;
      const int gy = iy0 + pix / a.in_tw, gx = ix0 + pix % a.in_tw;
      float v = 0.f;
      if (gy >= 0 && gy < s.h && gx >= 0 && gx < s.w && c0 + cc < s.c) v = to_f<T>(xp[((size_t)gy * s.w + gx) * s.c + c0 + cc]);
      s_x[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ps = 0; ps < MAXPASS; ++ps) {
      if (!g_on[ps]) continue;
      for (int pix = part; pix < NPIX; pix += parts) {
        const int ty = pix / TW, tx = pix % TW;
        const float4 d = *reinterpret_cast<const float4*>(s_dy + pix * KC + g_kq[ps] * 4);
        const float4 xv = *reinterpret_cast<const float4*>(s_x + ((ty * s.stride_h) * a.in_tw + tx * s.stride_w + g_off[ps]) * CC + g_cq[ps] * 4);
        const float dd[4] = {d.x, d.y, d.z, d.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[ps][i * 4 + j] = fmaf(dd[i], xx[j], acc[ps][i * 4 + j]);
      }
    }
  }
  // combine pixel partitions, then one atomic per element
  __syncthreads();
  float* s_red = smem;   // [NT][16] per pass, reused
#pragma unroll
  for (int ps = 0; ps < MAXPASS; ++ps) {
    if (ps > 0 && groups < NT) break;
    if (ps * NT >= groups) break;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) s_red[tid * 16 + i] = g_on[ps] ? acc[ps][i] : 0.f;
    __syncthreads();
    if (g_on[ps] && part == 0) {
      const int g = (groups >= NT ? tid : tid % groups) + ps * NT;
      const int tap = tap0 + g / (kq_n * cq_n);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int kk = k0 + g_kq[ps] * 4 + i, cc = c0 + g_cq[ps] * 4 + j;
          if (kk < s.k && cc < s.c) {
            float v = 0.f;
            if (groups >= NT) v = s_red[tid * 16 + i * 4 + j];
            else for (int q = 0; q < parts; ++q) v += s_red[(q * groups + (tid % groups)) * 16 + i * 4 + j];
            atomicAdd(a.dw + ((size_t)kk * s.r * s.s + tap) * s.c + cc, v);
          }
        }
    }
  }
}


// Stride-1 / dilation-1 weight gradient with a compile-time filter width S. A thread owns one filter row r, a 4k x 4c register block and ALL S
// column taps of that row (S*16 accumulators); it walks output rows y (strided over the threads that share the unit) and slides along x keeping the
// S input columns of row y+r in registers: per pixel 2 shared 128-bit loads (dy, newest input column) feed 16*S FMAs.
template <typename T, int S, int TW, int TH, int KC, int CC>
__global__ void __launch_bounds__(256, 2) conv_wgrad_direct_s1_kernel(const WgradArgs a) {   // <= 128 registers: 2 CTAs of 256 threads or 3 of 160 per SM
  constexpr int NPIX = TW * TH, TWP = TW + 1;   // TWP: odd row pitch (in pixels) of the staged dy tile => conflict-free 128-bit loads across rows   // KC / CC: channel slab widths (4, 8 or 16) = shared-memory pixel strides
  extern __shared__ __align__(16) float smem[];
  const dcv_conv_shape& s = a.s;
  float* s_dy = smem;                 // [NPIX][KC]
  float* s_x = smem + TH * TWP * KC;  // [in_th][in_tw (odd)][CC]
  const int tid = threadIdx.x, NT = blockDim.x;   // NT = units * L rounded up to whole warps (<= 256): no idle warps holding registers
  const int tile = blockIdx.x, tile_y = tile / a.tiles_x, tile_x = tile - tile_y * a.tiles_x;
  int z = blockIdx.z;
  const int k0 = (z % a.kslabs) * KC; z /= a.kslabs;
  const int c0 = z * CC;
  const int kq_n = (min(KC, s.k - k0) + 3) / 4, cq_n = (min(CC, s.c - c0) + 3) / 4;
  const int units = s.r * kq_n * cq_n;
  // L consecutive lanes (a power of two, <= 32) share a unit and split the tile into (row, x-segment) pieces: their partial sums are combined
  // with warp shuffles at the end, no shared-memory reduction
  int L = 32;
  while (L > 1 && units * L > NT) L >>= 1;
  const int xseg = max(1, min(L / TH, TW / S)), seglen = (TW + xseg - 1) / xseg, nseg = TH * xseg;
  const int u = tid / L, part = tid % L;
  const bool active = u < units;
  const bool warp_has_work = ((tid & ~31) / L) < units;   // warp-uniform
  const int rr = u / (kq_n * cq_n), kq = (u / cq_n) % kq_n, cq = u % cq_n;
  const int oy0 = tile_y * TH, ox0 = tile_x * TW;
  const int iy0 = oy0 - s.pad_h, ix0 = ox0 - s.pad_w;
  const int in_pix = a.in_th * a.in_tw;

  float acc[S][16];
#pragma unroll
  for (int ss = 0; ss < S; ++ss)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[ss][i] = 0.f;

  const T* xall = reinterpret_cast<const T*>(a.x);
  const T* dyall = reinterpret_cast<const T*>(a.dy);
  for (int img = blockIdx.y; img < s.n; img += gridDim.y) {
    __syncthreads();
    const T* dyp = dyall + (size_t)img * s.p * s.q * s.k;
    if (a.vec_dy) {   // 4 channels per load
      for (int i = tid; i < NPIX * (KC / 4); i += NT) {
        const int g4 = i % (KC / 4), pix = i / (KC / 4);
        const int oy = oy0 + pix / TW, ox = ox0 + pix % TW;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (oy < s.p && ox < s.q && k0 + 4 * g4 < s.k) load4<T>(dyp + ((size_t)oy * s.q + ox) * s.k + k0 + 4 * g4, v);
        *reinterpret_cast<float4*>(s_dy + ((pix / TW) * TWP + pix % TW) * KC + 4 * g4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
      for (int i = tid; i < NPIX * KC; i += NT) {
        const int kk = i % KC, pix = i / KC;
        const int oy = oy0 + pix / TW, ox = ox0 + pix % TW;
        float v = 0.f;
        if (oy < s.p && ox < s.q && k0 + kk < s.k) v = to_f<T>(dyp[((size_t)oy * s.q + ox) * s.k + k0 + kk]);
        s_dy[((pix / TW) * TWP + pix % TW) * KC + kk] = v;
      }
    }
    const T* xp = xall + (size_t)img * s.h * s.w * s.c;
    if (a.vec_x) {
      for (int i = tid; i < in_pix * (CC / 4); i += NT) {
        const int g4 = i % (CC / 4), pix = i / (CC / 4);
        const int gy = iy0 + pix / a.in_tw, gx = ix0 + pix % a.in_tw;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (gy >= 0 && gy < s.h && gx >= 0 && gx < s.w && c0 + 4 * g4 < s.c) load4<T>(xp + ((size_t)gy * s.w + gx) * s.c + c0 + 4 * g4, v);
        *reinterpret_cast<float4*>(s_x + pix * CC + 4 * g4) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
      for (int i = tid; i < in_pix * CC; i += NT) {
        const int cc = i % CC, pix = i / CC;
        const int gy = iy0 + pix / a.in_tw, gx = ix0 + pix % a.in_tw;
        float v = 0.f;
        if (gy >= 0 && gy < s.h && gx >= 0 && gx < s.w && c0 + cc < s.c) v = to_f<T>(xp[((size_t)gy * s.w + gx) * s.c + c0 + cc]);
        s_x[i] = v;
      }
    }
    __syncthreads();
    if (active) {
      for (int seg = part; seg < nseg; seg += L) {
        const int xs = seg / TH, y = seg - xs * TH;   // consecutive lanes walk rows: with odd row pitches their 16-byte loads hit distinct banks
        const int xbeg = xs * seglen, xend = min(TW, xbeg + seglen);
        const float* xrow = s_x + (size_t)((y + rr) * a.in_tw) * CC + cq * 4;
        const float* drow = s_dy + (size_t)(y * TWP) * KC + kq * 4;
        float4 win[S];
#pragma unroll
        for (int j = 0; j < S - 1; ++j) win[j] = *reinterpret_cast<const float4*>(xrow + (xbeg + j) * CC);
        for (int x0 = xbeg; x0 < xend; x0 += S) {
#pragma unroll
          for (int j = 0; j < S; ++j) {
            const int x = x0 + j;
            if (x < xend) {
              win[(j + S - 1) % S] = *reinterpret_cast<const float4*>(xrow + (x + S - 1) * CC);
              const float4 d = *reinterpret_cast<const float4*>(drow + x * KC);
              const float dd[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
              for (int ss = 0; ss < S; ++ss) {
                const float4 w4 = win[(j + ss) % S];
                const float xx[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) acc[ss][i * 4 + jj] = fmaf(dd[i], xx[jj], acc[ss][i * 4 + jj]);
              }
            }
          }
        }
      }
    }
  }
  // Combine the L lanes of each unit. Halving butterfly over the 16 (k, c) register-block entries: at lane bit b a lane keeps one half of its live
  // entries and receives the partner's sums of that half, so after log2(min(L, 16)) steps every lane owns 16 / min(L, 16) entries, fully summed over
  // those lanes (15 shuffles per tap instead of 16 per step), then (L == 32) one plain exchange. The atomics are spread over the lanes as well.
  // (First version: every lane reduced all 16 * S values with log2(L) shuffle steps each and lane 0 issued all the atomics: ~35 % of the kernel's
  // instructions on the 4-channel 5x5 layers.)
  int cur_shift = 0;   // entries still live per tap = 16 >> cur_shift
  if (warp_has_work) {
#pragma unroll
    for (int st = 0; st < 4; ++st) {
      const int b = 1 << st;
      if (b < L) {
        const bool upper = (part & b) != 0;
        const int half = 8 >> st;
#pragma unroll
        for (int ss = 0; ss < S; ++ss)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < half) {
              const float send = upper ? acc[ss][j] : acc[ss][j + half];
              const float keep = upper ? acc[ss][j + half] : acc[ss][j];
              acc[ss][j] = keep + __shfl_xor_sync(0xffffffffu, send, b);
            }
        cur_shift = st + 1;
      }
    }
    if (L == 32) {
#pragma unroll
      for (int ss = 0; ss < S; ++ss) acc[ss][0] += __shfl_xor_sync(0xffffffffu, acc[ss][0], 16);
    }
  }
  if (active && part < 16) {
    const int live = 16 >> cur_shift;
    // entry index of acc[.][j]: the halving steps consumed the index bits from the top (bit 3 at lane bit 0, bit 2 at lane bit 1, ...)
    int ibase = 0;
    for (int st = 0; st < cur_shift; ++st) ibase += ((part >> st) & 1) * (8 >> st);
#pragma unroll
    for (int ss = 0; ss < S; ++ss)
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < live) {
          const int e = ibase + j, i = e >> 2, jj = e & 3;
          const int kk = k0 + kq * 4 + i, cc = c0 + cq * 4 + jj;
          if (kk < s.k && cc < s.c) atomicAdd(a.dw + (((size_t)kk * s.r + rr) * S + ss) * s.c + cc, acc[ss][j]);
        }
  }
}

// ---- host-side launchers ------------------------------------------------------------------------------------------------
static int check_shape(const dcv_conv_shape* s, const char* name) {
  DCV_REQUIRE(s, "%s: null shape", name);
  DCV_REQUIRE(s->n > 0 && s->h > 0 && s->w > 0 && s->c > 0 && s->k > 0 && s->r > 0 && s->s > 0, "%s: non-positive dimension", name);
  DCV_REQUIRE(s->stride_h > 0 && s->stride_w > 0 && s->dil_h > 0 && s->dil_w > 0 && s->pad_h >= 0 && s->pad_w >= 0, "%s: bad stride/dilation/padding", name);
  const int p = (s->h + 2 * s->pad_h - s->dil_h * (s->r - 1) - 1) / s->stride_h + 1;
  const int q = (s->w + 2 * s->pad_w - s->dil_w * (s->s - 1) - 1) / s->stride_w + 1;
  DCV_REQUIRE(p == s->p && q == s->q && p > 0 && q > 0, "%s: output size %dx%d inconsistent with geometry (expected %dx%d)", name, s->p, s->q, p, q);
  DCV_REQUIRE(s->n < 65536, "%s: batch %d exceeds grid limit", name, s->n);
  return 0;
}

template <typename T, int KC, int BX, int BY, int PX>
static int launch_fwd_cfg(DirectConvArgs a, cudaStream_t st) {
  const dcv_conv_shape& s = a.s;
  constexpr int TW = BX * PX, TH = BY;
  a.in_th = (TH - 1) * s.stride_h + (s.r - 1) * s.dil_h + 1;
  a.in_tw = (TW - 1) * s.stride_w + (s.s - 1) * s.dil_w + 1;
  a.in_pitch = a.in_tw;
  while (a.in_pitch % 32 != 8) ++a.in_pitch;
  int cc = s.c < 16 ? s.c : 16;
  auto bytes = [&](int c) { return ((size_t)c * a.in_th * a.in_pitch + (size_t)s.r * s.s * c * KC) * sizeof(float); };
  while (cc > 1 && bytes(cc) > 160 * 1024) --cc;
  DCV_REQUIRE(bytes(cc) <= 220 * 1024, "conv direct: kernel %dx%d (dilation %dx%d) too large for the shared-memory tile", s.r, s.s, s.dil_h, s.dil_w);
  a.cc = cc;
  a.tiles_x = (s.q + TW - 1) / TW;
  const int tiles = a.tiles_x * ((s.p + TH - 1) / TH);
  dim3 grid(tiles, s.n, (s.k + KC - 1) / KC);
  auto kern = conv_fwd_direct_kernel<T, KC, BX, BY, PX>;
  const size_t smem = bytes(cc);
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, BX * BY, smem, st>>>(a);
  DCV_LAUNCH_CHECK("conv_fwd_direct_kernel");
  return 0;
}


template <typename T, int KC, int BX, int BY, int PX, int S>
static int launch_fwd_s1_cfg(DirectConvArgs a, cudaStream_t st) {
  const dcv_conv_shape& s = a.s;
  constexpr int TW = BX * PX, TH = BY;
  a.in_th = TH + s.r - 1;
  a.in_tw = TW + S - 1;
  a.in_pitch = (a.in_tw + 3) / 4 * 4;
  if ((a.in_pitch / 4) % 2 == 0) a.in_pitch += 4;          // pitch/4 odd: the 128-bit row loads of a quarter warp hit distinct banks
  int cc = s.c < 16 ? s.c : 16;
  auto bytes = [&](int c) { return ((size_t)c * a.in_th * a.in_pitch + 4 + (size_t)s.r * S * c * KC) * sizeof(float); };
  while (cc > 1 && bytes(cc) > 96 * 1024) --cc;
  a.cc = cc;
  a.vec4 = (s.c % 4 == 0 && cc % 4 == 0 && reinterpret_cast<uintptr_t>(a.x) % (4 * sizeof(T)) == 0) ? 1 : 0;
  a.tiles_x = (s.q + TW - 1) / TW;
  const int tiles = a.tiles_x * ((s.p + TH - 1) / TH);
  dim3 grid(tiles, s.n, (s.k + KC - 1) / KC);
  auto kern = conv_fwd_direct_s1_kernel<T, KC, BX, BY, PX, S>;
  const size_t smem = bytes(cc);
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, BX * BY, smem, st>>>(a);
  DCV_LAUNCH_CHECK("conv_fwd_direct_s1_kernel");
  return 0;
}

// Returns -1 when the specialised kernel does not apply (caller falls back to the generic one).
template <typename T, int KC>
static int launch_fwd_s1(const DirectConvArgs& a, cudaStream_t st) {
  const dcv_conv_shape& s = a.s;
  if (s.stride_h != 1 || s.stride_w != 1 || s.dil_h != 1 || s.dil_w != 1 || s.r > 7) return -1;
  if (s.q >= 24) {
    // 128-thread CTAs, 4 pixels per thread: measured 15-30 % faster than 64 threads x 8 pixels on the 32x32 CIFAR layers (more warps per SM to hide the
    // shared-memory latency; ncu: 19 % of peak warps active with the 64-thread tiles)
    if (s.s == 3) return launch_fwd_s1_cfg<T, KC, 8, 16, 4, 3>(a, st);
    if (s.s == 5) return launch_fwd_s1_cfg<T, KC, 8, 16, 4, 5>(a, st);
  } else if (s.q >= 12) {
    if (s.s == 3) return launch_fwd_s1_cfg<T, KC, 4, 16, 4, 3>(a, st);
    if (s.s == 5) return launch_fwd_s1_cfg<T, KC, 4, 16, 4, 5>(a, st);
  }
  return -1;
}

template <typename T, int KC>
static int launch_fwd_kc(const DirectConvArgs& a, cudaStream_t st) {
  if (KC % 4 == 0) { const int rc = launch_fwd_s1<T, KC>(a, st); if (rc >= 0) return rc; }
  if (a.s.q >= 24) return launch_fwd_cfg<T, KC, 8, 16, 4>(a, st);
  if (a.s.q >= 12) return launch_fwd_cfg<T, KC, 8, 16, 2>(a, st);
  return launch_fwd_cfg<T, KC, 8, 8, 1>(a, st);
}

template <typename T>
static int launch_fwd(const DirectConvArgs& a, cudaStream_t st) {
  if (a.s.k <= 4) return launch_fwd_kc<T, 4>(a, st);
  if (a.s.k <= 8) return launch_fwd_kc<T, 8>(a, st);
  // Small maps (the 16x16 CIFAR layers: one 64-thread tile per image) leave most of the machine idle with 16-channel slabs: 512 CTAs x 2 warps = 7 warps
  // per SM. 8-channel slabs double the CTAs for the price of staging the input tile twice.
  static const bool kc16_only = getenv("DCV_FWD_KC16") != nullptr;
  const long long tiles16 = (long long)((a.s.q + 15) / 16) * ((a.s.p + 15) / 16) * a.s.n * ((a.s.k + 15) / 16);
  if (!kc16_only && a.s.q < 24 && tiles16 * 64 < (long long)kNumSMs * 1024) return launch_fwd_kc<T, 8>(a, st);
  return launch_fwd_kc<T, 16>(a, st);
}

int conv_fwd_direct(const dcv_conv_shape* shape, const void* x, const void* w, const float* bias, void* y, float* stats, int act, float slope, int dtype, cudaStream_t st) {
  if (check_shape(shape, "conv2d_fwd")) return 1;
  DCV_REQUIRE(x && w && y, "conv2d_fwd: null pointer");
  DirectConvArgs a{};
  a.s = *shape; a.x = x; a.w = w; a.bias = bias; a.y = y; a.stats = stats; a.act = act; a.slope = slope; a.transposed = 0;
  DCV_DISPATCH_DTYPE(dtype, T, return launch_fwd<T>(a, st));
  return 0;
}

int conv_dgrad_direct(const dcv_conv_shape* shape, const void* dy, const void* w, void* dx, int dtype, cudaStream_t st) {
  if (check_shape(shape, "conv2d_dgrad")) return 1;
  DCV_REQUIRE(dy && w && dx, "conv2d_dgrad: null pointer");
  const dcv_conv_shape& s = *shape;
  const int tp_h = s.dil_h * (s.r - 1) - s.pad_h, tp_w = s.dil_w * (s.s - 1) - s.pad_w;
  if (s.stride_h == 1 && s.stride_w == 1 && tp_h >= 0 && tp_w >= 0) {
    // dx = correlate(dy, flipped/transposed w) with padding dil*(R-1)-pad: the forward kernel on swapped roles
    DirectConvArgs a{};
    a.s.n = s.n; a.s.h = s.p; a.s.w = s.q; a.s.c = s.k; a.s.k = s.c; a.s.r = s.r; a.s.s = s.s;
    a.s.stride_h = a.s.stride_w = 1; a.s.pad_h = tp_h; a.s.pad_w = tp_w; a.s.dil_h = s.dil_h; a.s.dil_w = s.dil_w;
    a.s.p = s.h; a.s.q = s.w;
    a.x = dy; a.w = w; a.bias = nullptr; a.y = dx; a.stats = nullptr; a.act = DCV_ACT_NONE; a.slope = 0.f; a.transposed = 1;
    DCV_DISPATCH_DTYPE(dtype, T, return launch_fwd<T>(a, st));
    return 0;
  }
  const size_t total = (size_t)s.n * s.h * s.w * s.c;
  DCV_DISPATCH_DTYPE(dtype, T, (conv_dgrad_generic_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(s, (const T*)dy, (const T*)w, (T*)dx)));
  DCV_LAUNCH_CHECK("conv_dgrad_generic_kernel");
  return 0;
}

template <typename T, int TW, int TH>
static int launch_wgrad_cfg(WgradArgs a, cudaStream_t st) {
  const dcv_conv_shape& s = a.s;
  a.in_th = (TH - 1) * s.stride_h + (s.r - 1) * s.dil_h + 1;
  a.in_tw = (TW - 1) * s.stride_w + (s.s - 1) * s.dil_w + 1;
  a.tiles_x = (s.q + TW - 1) / TW;
  const int tiles = a.tiles_x * ((s.p + TH - 1) / TH);
  a.kslabs = (s.k + 15) / 16; a.cslabs = (s.c + 15) / 16;
  const int taps = s.r * s.s;
  a.tap_chunk = taps < 64 ? taps : 64;   // <= 64 taps * 16 register blocks = 1024 = 4 passes of 256 threads
  const int tap_chunks = (taps + a.tap_chunk - 1) / a.tap_chunk;
  const long long slabs = (long long)a.kslabs * a.cslabs * tap_chunks;
  DCV_REQUIRE(slabs < 65536, "conv2d_wgrad (direct): %lld channel slabs exceed the grid limit; use the tcgen05 algorithm", slabs);
  size_t smem = ((size_t)TW * TH * 16 + (size_t)a.in_th * a.in_tw * 16) * sizeof(float);
  if (smem < 256 * 16 * sizeof(float)) smem = 256 * 16 * sizeof(float);
  DCV_REQUIRE(smem <= 220 * 1024, "conv2d_wgrad (direct): kernel %dx%d too large for the shared-memory tile", s.r, s.s);
  long long gy = (long long)kNumSMs * 4 / ((long long)tiles * slabs) + 1;
  // at most 8 images per CTA: bounds the length of each sequential fp32 accumulation chain (the per-CTA partials are then combined by
  // atomics, i.e. blocked summation), unless that would take more than ~32M atomics
  const long long elems = (long long)s.k * s.r * s.s * s.c;
  long long gy_acc = (s.n + 7) / 8;
  while (gy_acc > gy && elems * gy_acc * tiles > (32ll << 20)) gy_acc /= 2;
  if (gy_acc > gy) gy = gy_acc;
  if (gy > s.n) gy = s.n;
  if (gy > 65535) gy = 65535;
  dim3 grid(tiles, (unsigned)gy, (unsigned)slabs);
  auto kern = conv_wgrad_direct_kernel<T, TW, TH>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  kern<<<grid, 256, smem, st>>>(a);
  DCV_LAUNCH_CHECK("conv_wgrad_direct_kernel");
  return 0;
}


template <typename T, int S, int TW, int TH, int KC, int CC>
static int launch_wgrad_s1_cfg(WgradArgs a, cudaStream_t st) {
  const dcv_conv_shape& s = a.s;
  a.in_th = TH + s.r - 1;
  a.in_tw = (TW + S - 1) | 1;        // odd row pitch (see the kernel)
  a.tiles_x = (s.q + TW - 1) / TW;
  const int tiles = a.tiles_x * ((s.p + TH - 1) / TH);
  a.kslabs = (s.k + KC - 1) / KC; a.cslabs = (s.c + CC - 1) / CC;
  const long long slabs = (long long)a.kslabs * a.cslabs;
  DCV_REQUIRE(slabs < 65536, "conv2d_wgrad (direct): %lld channel slabs exceed the grid limit; use the tcgen05 algorithm", slabs);
  size_t smem = ((size_t)(TW + 1) * TH * KC + (size_t)a.in_th * a.in_tw * CC) * sizeof(float);
  auto kern = conv_wgrad_direct_s1_kernel<T, S, TW, TH, KC, CC>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  // Exactly one resident wave of CTAs when the batch allows it (ncu on the 4->4 5x5 layer: 128 registers => 2 CTAs per SM = 296 slots, and the former
  // "4 per SM" guess launched 594 CTAs = 2.01 waves, the last two CTAs costing a third of the kernel); at most 8 images per CTA otherwise.
  // threads = (filter rows x 4k x 4c register blocks) x L lanes each, whole warps; same rule as in the kernel
  const int units_full = s.r * ((KC < s.k ? KC : s.k) + 3) / 4 * (((CC < s.c ? CC : s.c) + 3) / 4);
  int lanes = 32;
  while (lanes > 1 && units_full * lanes > 256) lanes >>= 1;
  int nthreads = (units_full * lanes + 31) / 32 * 32;
  if (nthreads > 256) nthreads = 256;
  if (nthreads < 64) nthreads = 64;
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nthreads, smem) != cudaSuccess || occ < 1) occ = 1;
  long long gy = (long long)kNumSMs * occ / ((long long)tiles * slabs);
  if (gy < 1) gy = 1;
  const long long elems = (long long)s.k * s.r * s.s * s.c;
  long long gy_acc = (s.n + 7) / 8;
  while (gy_acc > gy && elems * gy_acc * tiles > (32ll << 20)) gy_acc /= 2;
  if (gy_acc > gy) gy = gy_acc / gy * gy;   // whole waves
  if (gy > s.n) gy = s.n;
  if (gy > 65535) gy = 65535;
  a.vec_dy = (s.k % 4 == 0 && reinterpret_cast<uintptr_t>(a.dy) % (4 * sizeof(T)) == 0) ? 1 : 0;
  a.vec_x = (s.c % 4 == 0 && reinterpret_cast<uintptr_t>(a.x) % (4 * sizeof(T)) == 0) ? 1 : 0;
  dim3 grid(tiles, (unsigned)gy, (unsigned)slabs);
  kern<<<grid, nthreads, smem, st>>>(a);
  DCV_LAUNCH_CHECK("conv_wgrad_direct_s1_kernel");
  return 0;
}

template <typename T, int S, int TW, int TH>
static int launch_wgrad_s1_slabs(const WgradArgs& a, cudaStream_t st) {
  const dcv_conv_shape& s = a.s;
  const bool k4 = s.k <= 4, c4 = s.c <= 4;
  if (k4 && c4) return launch_wgrad_s1_cfg<T, S, TW, TH, 4, 4>(a, st);
  if (k4) return launch_wgrad_s1_cfg<T, S, TW, TH, 4, 16>(a, st);
  if (c4) return launch_wgrad_s1_cfg<T, S, TW, TH, 16, 4>(a, st);
  return launch_wgrad_s1_cfg<T, S, TW, TH, 16, 16>(a, st);
}

template <typename T>
static int launch_wgrad_s1(const WgradArgs& a, cudaStream_t st) {   // -1: not applicable
  const dcv_conv_shape& s = a.s;
  if (s.stride_h != 1 || s.stride_w != 1 || s.dil_h != 1 || s.dil_w != 1 || s.r > 7) return -1;
  if (s.q >= 24) {
    if (s.s == 3) return launch_wgrad_s1_slabs<T, 3, 32, 16>(a, st);
    if (s.s == 5) return launch_wgrad_s1_slabs<T, 5, 32, 16>(a, st);
  } else if (s.q >= 12) {
    if (s.s == 3) return launch_wgrad_s1_slabs<T, 3, 16, 16>(a, st);
    if (s.s == 5) return launch_wgrad_s1_slabs<T, 5, 16, 16>(a, st);
  }
  return -1;
}

int conv_wgrad_direct(const dcv_conv_shape* shape, const void* x, const void* dy, float* dw, int dtype, bool prezeroed, cudaStream_t st) {
  if (check_shape(shape, "conv2d_wgrad")) return 1;
  DCV_REQUIRE(x && dy && dw, "conv2d_wgrad: null pointer");
  const dcv_conv_shape& s = *shape;
  zero_accumulator(dw, (size_t)s.k * s.r * s.s * s.c * sizeof(float), st, prezeroed);
  WgradArgs a{};
  a.s = s; a.x = x; a.dy = dy; a.dw = dw;
  // the 7x7/stride-2 class of layers needs the small tile to fit the input halo in shared memory
  const bool big_halo = ((31 * s.stride_w + (s.s - 1) * s.dil_w + 1) * (15 * s.stride_h + (s.r - 1) * s.dil_h + 1)) > 2200;
  DCV_DISPATCH_DTYPE(dtype, T, {
    { const int rc = launch_wgrad_s1<T>(a, st); if (rc >= 0) return rc; }
    if (s.q >= 24 && !big_halo) return launch_wgrad_cfg<T, 32, 16>(a, st);
    if (s.q >= 12) return launch_wgrad_cfg<T, 16, 16>(a, st);
    return launch_wgrad_cfg<T, 8, 8>(a, st);
  });
  return 0;
}

}  // namespace dcv
