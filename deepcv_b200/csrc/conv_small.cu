// Few-channel convolution blocks (the default CIFAR-10 `image_classifier`: 3/4/16 channels, 5x5 and 3x3 filters, 32x32 and 16x16 maps) as three
// fused kernels per block — forward, weight gradient, data gradient — with the BatchNorm / GroupNorm that FOLLOWS the block's activation never run
// as passes of its own (reference block order: meta/nn.py:553 `Conv2d -> act -> BatchNorm2d -> GroupNorm`, spec conf/base/parameters.yml:8-19).
//
// Why not the tcgen05 kernels: K = R*S*C is 36..144 and N = 4..16 output channels; an M = 128 UMMA tile needs a software im2col tile in the swizzled
// operand layout (2.3 us per 128-pixel tile measured for the gather kernel) for 0.1 us of MMA. Here warp-level `mma.sync.m16n8k16` (bf16 -> fp32)
// takes its operands with `ldmatrix` STRAIGHT from the NHWC image tile in shared memory (one tile layout for all three kernels, no im2col, no
// transposed copies):
//   * 16-channel tensors: an 8x8 ldmatrix block is 8 pixels x 8 channels (16 bytes of a pixel); forward / data gradient read it as the A operand
//     (M = pixels, K = (tap, channel)), the weight gradient reads the same block TRANSPOSED (`ldmatrix.trans`: M = (tap, channel), K = pixels).
//   * 4-channel tensors (8 bytes per pixel): a 16-byte ldmatrix row is a PIXEL PAIR (2j, 2j+1), so the GEMM runs over pixel pairs and the pair's
//     parity moves into N: out[2j + par][k] = sum_{s', c} z[2j + s'][c] * w[s' - par][c][k] with s' = 0..KS (a (KS+1)-wide window, zero weights where
//     s' - par falls outside the filter), N = (par, k). Every MMA column is useful (N = 8 for 4 output channels instead of 4 + 4 padding), M halves:
//     256 MMAs per 32x32 image of a 5x5 4->4 layer instead of 448, each fed by ONE ldmatrix.x4 instead of four 32-bit loads.
//     The weight gradient uses the same identity transposed: D[(s', c)][(par, k)] = sum_j z[2j + s'][c] * dy[2j + par][k], dw[s] = D[s][par 0] + D[s + 1][par 1].
// The step is bound by latency and instruction issue, not math or HBM (arithmetic intensity 43-72 FLOP/B; 8 MB per launch is
// 1.3 us of HBM time), so what matters is the number of launches, passes, instructions and dependent round trips:
//
//   forward   reads the producer's RAW output y_prev and applies its pending normalisation z = A[n][c]*y + B[n][c] while staging the tile
//             ("normalise on load": z is never written), convolves, adds bias, activates, stores y (bf16) and accumulates the statistics of y:
//             per-(image, channel) sums (plain stores: a CTA owns whole images) and sharded per-channel batch sums (atomics). The raw image is
//             fetched into registers BEFORE the coefficient prologue, so its L2 round trip overlaps the prologue's.
//   A, B      are NOT produced by a finalize kernel: the FIRST consumer CTA of an image derives them from the raw sums in a ~100-flop fp64 prologue
//             (BatchNorm from the batch sums, GroupNorm from the image's sums) and caches the image's eight coefficients in `coef_nc`; the backward
//             kernels load them (one round trip) instead of repeating the division / rsqrt chains. The CTA of image 0 updates the running statistics.
//   backward  the consumer's data-gradient kernel writes dz (gradient w.r.t. the producer's normalised output) and, in its epilogue, the sums
//             s[n][c] = {sum dz, sum dz*y}, the image's GroupNorm adjoint coefficients (`d_nc`) and its contribution to the BatchNorm adjoint sums
//             (sharded atomics). The producer's weight- and data-gradient kernels then derive P, Q, R of dy = act'(y)*(P*dz + Q*y + R) per image in
//             their prologue and apply it WHILE LOADING their operand: no reduce / finalize / apply passes, dy is never written.
// Per block: 3 launches instead of 9 (conv, finalize, apply, reduce, finalize, apply, wgrad, dgrad + weight packing), and three activation-sized
// tensors (z, dy, the transposed weights) never touch HBM.
#include <stdlib.h>
#include "common.cuh"

namespace dcv {
namespace sc {

constexpr int kThreads = 128;   // forward / data gradient / pooling: 4 warps per image, >= 4 CTAs resident per SM => a batch of 512 images is ONE wave
constexpr int kWarps = kThreads / 32;
constexpr int kMinCtas = 4;
constexpr int kWgThreads = 256; // weight gradient: 8 warps, 2 CTAs per SM, persistent over images (fewer global atomics on dw)
constexpr int kWgWarps = kWgThreads / 32;
constexpr int kShards = 16;     // per-channel batch sums are spread over this many accumulators (index image % kShards): 32 atomics per address at batch 512
constexpr int kMaxC = 32;       // channels of a normalised tensor on this path

typedef __nv_bfloat16 bf16;

// Normalisation algebra of ONE image held in shared memory (lane = channel). z = A*y + B; BatchNorm u = al*y + be (mean mu, rstd rc);
// GroupNorm of u: mean gmean, rstd gr of the channel's group; backward: du = D1*dz + D2*y + D3, dy_pre = P*dz + Q*y + R.
struct Coef {
  float A[kMaxC], B[kMaxC], al[kMaxC], be[kMaxC], mu[kMaxC], rc[kMaxC], gmean[kMaxC], gr[kMaxC];
  float D1[kMaxC], D2[kMaxC], D3[kMaxC], P[kMaxC], Q[kMaxC], R[kMaxC];
};

// ---- forward coefficients of image `img` from the raw sums. Called by ONE whole warp (lane = channel, c <= 32).
__device__ __noinline__ void norm_forward_coeffs(const dcv_sc_norm& nd, int img, Coef& cf, bool update_running) {
  const int lane = threadIdx.x & 31, c = nd.c;
  double al = 1.0, be = 0.0, mean = 0.0, rstd = 1.0;
  if (lane < c && nd.use_bn) {
    double var;
    const bool run = nd.bn_running_mean && nd.bn_running_var;
    if (nd.bn_training) {
      double s1 = 0.0, s2 = 0.0;
      for (int sh = 0; sh < kShards; ++sh) { s1 += (double)nd.bn_sums[(sh * c + lane) * 2]; s2 += (double)nd.bn_sums[(sh * c + lane) * 2 + 1]; }
      const double m = (double)nd.n * (double)nd.hw;
      mean = s1 / m;
      var = s2 / m - mean * mean;
      if (var < 0.0) var = 0.0;
      if (update_running && run) {
        double mom = (double)nd.bn_momentum;
        if (mom < 0.0) mom = 1.0 / (double)((nd.bn_num_batches_tracked ? *nd.bn_num_batches_tracked : 0) + 1);
        const double unbiased = m > 1.0 ? var * m / (m - 1.0) : var;
        nd.bn_running_mean[lane] = (float)((1.0 - mom) * (double)nd.bn_running_mean[lane] + mom * mean);
        nd.bn_running_var[lane] = (float)((1.0 - mom) * (double)nd.bn_running_var[lane] + mom * unbiased);
      }
    } else {
      mean = (double)nd.bn_running_mean[lane];
      var = (double)nd.bn_running_var[lane];
    }
    rstd = rsqrt(var + (double)nd.bn_eps);
    al = (nd.bn_weight ? (double)nd.bn_weight[lane] : 1.0) * rstd;
    be = (nd.bn_bias ? (double)nd.bn_bias[lane] : 0.0) - mean * al;
  }
  if (lane < c) { cf.al[lane] = (float)al; cf.be[lane] = (float)be; cf.mu[lane] = (float)mean; cf.rc[lane] = (float)rstd; }
  __syncwarp();
  if (update_running && lane == 0 && nd.use_bn && nd.bn_training && nd.bn_num_batches_tracked) *nd.bn_num_batches_tracked += 1;   // every lane has read it above
  if (lane < c) {
    double A = (double)cf.al[lane], B = (double)cf.be[lane];   // the float-rounded values: the backward kernels recompute exactly these
    float gmean = 0.f, gr = 1.f;
    if (nd.use_gn) {
      const int cg = c / nd.gn_groups, g0 = lane / cg * cg;
      const double hw = (double)nd.hw;
      double su = 0.0, suu = 0.0;
      for (int k = 0; k < cg; ++k) {
        const int ch = g0 + k;
        const double a_ = (double)cf.al[ch], b_ = (double)cf.be[ch];
        const double sy = (double)nd.stats_nc[((size_t)img * c + ch) * 2], syy = (double)nd.stats_nc[((size_t)img * c + ch) * 2 + 1];
        su += a_ * sy + hw * b_;
        suu += a_ * a_ * syy + 2.0 * a_ * b_ * sy + hw * b_ * b_;
      }
      const double mg = (double)cg * hw, mean_u = su / mg;
      double var = suu / mg - mean_u * mean_u;
      if (var < 0.0) var = 0.0;
      gmean = (float)mean_u; gr = (float)rsqrt(var + (double)nd.gn_eps);
      const double a2 = (nd.gn_weight ? (double)nd.gn_weight[lane] : 1.0) * (double)gr;
      const double b2 = (nd.gn_bias ? (double)nd.gn_bias[lane] : 0.0) - (double)gmean * a2;
      B = a2 * B + b2;
      A = a2 * A;
    }
    cf.gmean[lane] = gmean; cf.gr[lane] = gr; cf.A[lane] = (float)A; cf.B[lane] = (float)B;
    if (nd.coef_nc) {   // the backward kernels load exactly these floats
      float4* o = reinterpret_cast<float4*>(nd.coef_nc + ((size_t)img * c + lane) * 8);
      o[0] = make_float4(cf.A[lane], cf.B[lane], cf.al[lane], cf.be[lane]);
      o[1] = make_float4(cf.mu[lane], cf.rc[lane], gmean, gr);
    }
  }
  __syncwarp();
}

// ---- the cached forward coefficients of image `img` (backward kernels). One warp.
__device__ __forceinline__ void load_coeffs(const dcv_sc_norm& nd, int img, Coef& cf) {
  const int lane = threadIdx.x & 31;
  if (lane < nd.c) {
    const float4* o = reinterpret_cast<const float4*>(nd.coef_nc + ((size_t)img * nd.c + lane) * 8);
    const float4 a = o[0], b = o[1];
    cf.A[lane] = a.x; cf.B[lane] = a.y; cf.al[lane] = a.z; cf.be[lane] = a.w; cf.mu[lane] = b.x; cf.rc[lane] = b.y; cf.gmean[lane] = b.z; cf.gr[lane] = b.w;
  }
  __syncwarp();
}

// ---- GroupNorm adjoint of image `img`: D1, D2, D3 from the image's sums s[c][2] = {sum dz, sum dz*y} (any memory). One warp, after norm_forward_coeffs.
__device__ __noinline__ void norm_backward_D(const dcv_sc_norm& nd, Coef& cf, const float* s) {
  const int lane = threadIdx.x & 31, c = nd.c;
  if (lane < c) {
    double D1 = 1.0, D2 = 0.0, D3 = 0.0;
    if (nd.use_gn) {
      const int cg = c / nd.gn_groups, g0 = lane / cg * cg;
      const double r = (double)cf.gr[lane], mean_u = (double)cf.gmean[lane], mg = (double)cg * (double)nd.hw;
      double a = 0.0, b = 0.0;
      for (int k = 0; k < cg; ++k) {
        const int ch = g0 + k;
        const double gam = nd.gn_weight ? (double)nd.gn_weight[ch] : 1.0;
        const double s1 = (double)s[2 * ch], s2 = (double)s[2 * ch + 1];
        a += gam * s1;
        b += gam * r * ((double)cf.al[ch] * s2 + ((double)cf.be[ch] - mean_u) * s1);
      }
      const double Ag = a / mg, Bg = b / mg, gam = nd.gn_weight ? (double)nd.gn_weight[lane] : 1.0;
      D1 = r * gam;
      D2 = -r * Bg * r * (double)cf.al[lane];
      D3 = -r * Ag - r * Bg * r * ((double)cf.be[lane] - mean_u);
    }
    cf.D1[lane] = (float)D1; cf.D2[lane] = (float)D2; cf.D3[lane] = (float)D3;
  }
  __syncwarp();
}

// ---- what the kernel that COMPLETES the sums s of image `img` does with them (one warp; s in shared memory): stores them for the producer's backward
// kernels and adds the image's terms of the BatchNorm adjoint sums / GroupNorm parameter gradients to the sharded accumulators u_sums[shard][c][4] =
// {U1 = sum du, U2raw = sum du*y, d gn_weight, d gn_bias}.
__device__ __noinline__ void norm_backward_image_sums(const dcv_sc_norm& nd, int img, Coef& cf, const float* s) {
  load_coeffs(nd, img, cf);
  norm_backward_D(nd, cf, s);
  const int lane = threadIdx.x & 31, c = nd.c;
  if (lane < c) {
    *reinterpret_cast<float4*>(nd.d_nc + ((size_t)img * c + lane) * 4) = make_float4(cf.D1[lane], cf.D2[lane], cf.D3[lane], 0.f);   // for the producer's backward kernels
    const double s1 = (double)s[2 * lane], s2 = (double)s[2 * lane + 1], hw = (double)nd.hw;
    const double sy = (double)nd.stats_nc[((size_t)img * c + lane) * 2], syy = (double)nd.stats_nc[((size_t)img * c + lane) * 2 + 1];
    const double D1 = (double)cf.D1[lane], D2 = (double)cf.D2[lane], D3 = (double)cf.D3[lane];
    if (nd.s_nc) { nd.s_nc[((size_t)img * c + lane) * 2] = (float)s1; nd.s_nc[((size_t)img * c + lane) * 2 + 1] = (float)s2; }   // diagnostic only: nothing reads it
    float* u = nd.u_sums + ((size_t)(img % kShards) * c + lane) * 4;
    atomicAdd(u, (float)(D1 * s1 + D2 * sy + D3 * hw));
    atomicAdd(u + 1, (float)(D1 * s2 + D2 * syy + D3 * sy));
    if (nd.use_gn) {
      atomicAdd(u + 2, (float)((double)cf.gr[lane] * ((double)cf.al[lane] * s2 + ((double)cf.be[lane] - (double)cf.gmean[lane]) * s1)));
      atomicAdd(u + 3, (float)s1);
    }
  }
  __syncwarp();
}

// ---- P, Q, R of image `img` (one warp): cached forward coefficients and D, BatchNorm adjoint from the complete batch sums.
// `param_grads`: this warp also writes the normalisation parameter gradients (one CTA of one kernel per block does).
__device__ __noinline__ void norm_backward_pqr(const dcv_sc_norm& nd, int img, Coef& cf, bool param_grads, float* d_bn_w, float* d_bn_b, float* d_gn_w, float* d_gn_b) {
  load_coeffs(nd, img, cf);
  const int lane = threadIdx.x & 31, c = nd.c;
  if (lane < c) {
    const float4 d = *reinterpret_cast<const float4*>(nd.d_nc + ((size_t)img * c + lane) * 4);
    cf.D1[lane] = d.x; cf.D2[lane] = d.y; cf.D3[lane] = d.z;
    double U1 = 0.0, U2raw = 0.0, dgw = 0.0, dgb = 0.0;
    for (int sh = 0; sh < kShards; ++sh) {
      const float* u = nd.u_sums + ((size_t)sh * c + lane) * 4;
      U1 += (double)u[0]; U2raw += (double)u[1]; dgw += (double)u[2]; dgb += (double)u[3];
    }
    const double D1 = (double)cf.D1[lane], D2 = (double)cf.D2[lane], D3 = (double)cf.D3[lane];
    double P = D1, Q = D2, R = D3, u2 = 0.0;
    if (nd.use_bn) {
      const double al = (double)cf.al[lane];
      if (nd.bn_training) {
        const double mu = (double)cf.mu[lane], rc = (double)cf.rc[lane], m = (double)nd.n * (double)nd.hw;
        u2 = rc * (U2raw - mu * U1);
        P = al * D1;
        Q = al * (D2 - rc * u2 / m);
        R = al * (D3 - U1 / m + mu * rc * u2 / m);
      } else {
        u2 = (double)cf.rc[lane] * (U2raw - (double)cf.mu[lane] * U1);
        P = al * D1; Q = al * D2; R = al * D3;
      }
    }
    cf.P[lane] = (float)P; cf.Q[lane] = (float)Q; cf.R[lane] = (float)R;
    if (param_grads) {
      if (nd.use_bn) { if (d_bn_w) d_bn_w[lane] = (float)u2; if (d_bn_b) d_bn_b[lane] = (float)U1; }
      if (nd.use_gn) { if (d_gn_w) d_gn_w[lane] = (float)dgw; if (d_gn_b) d_gn_b[lane] = (float)dgb; }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == DCV_ACT_RELU) return fmaxf(v, 0.f);
  if (act == DCV_ACT_LEAKY_RELU) return v > 0.f ? v : v * slope;
  if (act == DCV_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}
__device__ __forceinline__ float act_bwd(float y, int act, float slope) {
  if (act == DCV_ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == DCV_ACT_LEAKY_RELU) return y > 0.f ? 1.f : slope;
  if (act == DCV_ACT_SIGMOID) return y * (1.f - y);
  return 1.f;
}

__device__ __forceinline__ void mma16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// ldmatrix: lane l supplies the 16-byte row (l & 7) of 8x8 matrix (l >> 3); register i of thread (g = lane / 4, t = lane % 4) is elements (2t, 2t + 1) of
// row g of matrix i — or, with .trans, rows (2t, 2t + 1) of column g.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// (pdl_wait / pdl_trigger: common.cuh. Here the wait comes AFTER everything that does not depend on the predecessor: zeroing the tile, building the
// weight fragments — the weights were written several kernels earlier.)
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ float round_bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---- operand staging -------------------------------------------------------------------------------------------------------------------------
// Source element of image-local index (pixel, channel): either the plain tensor, the producer's raw output with its pending affine (A, B), or the
// pre-activation gradient dy = act'(y) * (P*dz + Q*y + R) assembled from dz and y. `fetch` only issues the global loads (their results are first
// touched by `decode`), so a kernel can put a prologue and a barrier between the two.
struct SrcPlain {
  const bf16* x; const Coef* cf;   // cf == nullptr: plain
  struct Raw { uint4 a; };
  __device__ __forceinline__ Raw fetch(size_t e0) const { Raw r; r.a = *reinterpret_cast<const uint4*>(x + e0); return r; }
  __device__ __forceinline__ void decode(const Raw& r, int c0, int cmask, float* v) const {   // channel of element i = (c0 + i) & cmask
    vec_unpack<bf16>(r.a, v);
    if (cf) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int ch = (c0 + i) & cmask; v[i] = fmaf(cf->A[ch], v[i], cf->B[ch]); }
    }
  }
  __device__ __forceinline__ float load1(size_t e, int ch) const {
    const float v = __bfloat162float(x[e]);
    return cf ? fmaf(cf->A[ch], v, cf->B[ch]) : v;
  }
};
struct SrcDy {
  const bf16* dz; const bf16* y; const Coef* cf;   // cf == nullptr: P = 1, Q = R = 0
  int act; float slope;
  struct Raw { uint4 a, b; };
  __device__ __forceinline__ Raw fetch(size_t e0) const {
    Raw r; r.a = *reinterpret_cast<const uint4*>(dz + e0); r.b = *reinterpret_cast<const uint4*>(y + e0); return r;
  }
  __device__ __forceinline__ void decode(const Raw& r, int c0, int cmask, float* v) const {
    float b[8];
    vec_unpack<bf16>(r.a, v);
    vec_unpack<bf16>(r.b, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = (c0 + i) & cmask;
      const float pre = cf ? fmaf(cf->P[ch], v[i], fmaf(cf->Q[ch], b[i], cf->R[ch])) : v[i];
      v[i] = pre * act_bwd(b[i], act, slope);
    }
  }
  __device__ __forceinline__ float load1(size_t e, int ch) const {
    const float a = __bfloat162float(dz[e]), b = __bfloat162float(y[e]);
    const float pre = cf ? fmaf(cf->P[ch], a, fmaf(cf->Q[ch], b, cf->R[ch])) : a;
    return pre * act_bwd(b, act, slope);
  }
};

// NHWC tile with halo: tile[(y + PAD) * Wp + x + PAD][CI] = src(y, x, 0..c_src) (channels >= c_src zero). The halo and the padding channels were zeroed
// once and are never written. Three routes: whole 16-byte vectors when the tensor has exactly CI channels (two-phase: `fetch` the first PF vectors of
// every thread, [the caller's prologue + barrier], `store`); 16-byte loads + 2-byte tile stores when c_src < CI but the image is a whole number of
// aligned vectors (the 3-channel network input); scalar otherwise.
constexpr int kPF = 4;   // vectors of a thread in flight across the prologue
template <int CI, int NTHR, typename Src, int PF = kPF> struct Stager {
  typename Src::Raw raw[PF];
  bool vec;
  __device__ __forceinline__ void fetch(const Src& src, size_t e_img, int HW, int c_src, int tid) {
    vec = (c_src == CI) && (e_img % 8) == 0;
    if (vec) {
      const int nvec = HW * CI / 8;
#pragma unroll
      for (int i = 0; i < PF; ++i) { const int v = tid + i * NTHR; if (v < nvec) raw[i] = src.fetch(e_img + (size_t)v * 8); }
    }
  }
  float* vsum = nullptr;   // when set: 8 running sums of the (bf16-rounded) staged values per vector slot (the bias gradient; vector route only)
  __device__ __forceinline__ void put(bf16* tile, const Src& src, const typename Src::Raw& r, int v, int W, int wlog, int Wp, int PAD) const {
    const int e0 = v * 8, pix = e0 / CI, c0 = e0 % CI, y = pix >> wlog, x = pix & (W - 1);
    float f[8];
    src.decode(r, c0, CI - 1, f);
    if (vsum) {
#pragma unroll
      for (int i = 0; i < 8; ++i) vsum[i] += round_bf(f[i]);
    }
    bf16* dst = tile + ((size_t)(y + PAD) * Wp + x + PAD) * CI + c0;
    if (CI == 4) {   // two pixels of 8 bytes each (x is even, so both are in the same row)
      *reinterpret_cast<uint2*>(dst) = make_uint2(pack2(f[0], f[1]), pack2(f[2], f[3]));
      *reinterpret_cast<uint2*>(dst + 4) = make_uint2(pack2(f[4], f[5]), pack2(f[6], f[7]));
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
    }
  }
  __device__ __forceinline__ void store(bf16* tile, const Src& src, size_t e_img, int H, int W, int Wp, int PAD, int c_src, int tid) {
    const int wlog = 31 - __clz(W), HW = H * W;
    if (vec) {
      const int nvec = HW * CI / 8;
#pragma unroll
      for (int i = 0; i < PF; ++i) { const int v = tid + i * NTHR; if (v < nvec) put(tile, src, raw[i], v, W, wlog, Wp, PAD); }
      for (int v = tid + PF * NTHR; v < nvec; v += NTHR) put(tile, src, src.fetch(e_img + (size_t)v * 8), v, W, wlog, Wp, PAD);
    } else if (src.cf == nullptr && c_src < CI && (HW * c_src) % 8 == 0 && ((e_img * 2) % 16) == 0 && sizeof(typename Src::Raw) == sizeof(uint4)) {
      const int nvec = HW * c_src / 8;
      for (int v = tid; v < nvec; v += NTHR) {
        const typename Src::Raw r = src.fetch(e_img + (size_t)v * 8);
        float f[8];
        vec_unpack<bf16>(r.a, f);
        int pix = (v * 8) / c_src, c = (v * 8) - pix * c_src;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int y = pix >> wlog, x = pix & (W - 1);
          tile[((size_t)(y + PAD) * Wp + x + PAD) * CI + c] = __float2bfloat16_rn(f[i]);
          if (++c == c_src) { c = 0; ++pix; }
        }
      }
    } else {
      for (int pix = tid; pix < HW; pix += NTHR) {
        const int y = pix >> wlog, x = pix & (W - 1);
        bf16* dst = tile + ((size_t)(y + PAD) * Wp + x + PAD) * CI;
#pragma unroll
        for (int c = 0; c < CI; ++c) dst[c] = __float2bfloat16_rn(c < c_src ? src.load1(e_img + (size_t)pix * c_src + c, c) : 0.f);
      }
    }
  }
};

// ---- GEMM geometry shared by the three kernels ------------------------------------------------------------------------------------------------------
// CI: channels of the tile (4: pixel-pair mode, 16: plain mode); NO: channels on the other side (4 or 16); KS: filter size.
// A "chunk" is one 16-byte ldmatrix row's worth of the (tap, channel) index: pair mode: two adjacent taps (s0, s0 + 1), s0 even, of filter row r, 4 channels
// each; plain mode: half (8 channels) of one tap.
template <int CI, int NO, int KS> struct Geo {
  static constexpr bool PAIR = (CI == 4);
  static constexpr int CPR = (KS + 1) / 2;                              // chunks per filter row (pair mode)
  static constexpr int NCH = PAIR ? KS * CPR : KS * KS * 2;             // chunks
  static constexpr int KSTEPS = (NCH + 1) / 2;                          // forward / data gradient: k-steps of 16 = 2 chunks
  static constexpr int N = PAIR ? 2 * NO : (NO < 8 ? 8 : NO);           // GEMM columns: (parity, channel) or channel
  static constexpr int NT = N / 8;
  static constexpr int PIXB = CI * 2;                                   // bytes per tile pixel
  // byte offset of chunk `ch` relative to the row address of the pixel (pair)
  __device__ static __forceinline__ int chunk_off(int ch, int Wp) {
    if (ch >= NCH) return 0;   // padding chunk: any valid address (its weights / results are zero / ignored)
    if (PAIR) return ((ch / CPR) * Wp + 2 * (ch % CPR)) * PIXB;
    const int tap = ch >> 1;
    return ((tap / KS) * Wp + tap % KS) * PIXB + (ch & 1) * 16;
  }
  // (filter row, filter column, channel) of element e (0..7) of chunk ch for GEMM column n; returns false when the product term does not exist
  __device__ static __forceinline__ bool decode(int ch, int e, int n, int& r, int& s, int& c, int& o) {
    if (ch >= NCH) return false;
    if (PAIR) {
      r = ch / CPR; c = e & 3; o = n % NO;
      s = 2 * (ch % CPR) + (e >> 2) - n / NO;
      return s >= 0 && s < KS;
    }
    const int tap = ch >> 1;
    r = tap / KS; s = tap % KS; c = (ch & 1) * 8 + e; o = n;
    return n < NO;
  }
};

// ---- implicit-GEMM core shared by forward and data gradient ------------------------------------------------------------------------------------
// One image: M = pixel pairs (pair mode) or pixels, 16 consecutive per m-tile; A by ldmatrix.x4 from the tile; `breg`: the B fragments (weights),
// resident in registers for the whole kernel. `pre(y, x, o, slot)` may start global loads whose results the epilogue needs (issued before the tile's
// MMAs); `epi(y, x, o, v0, v1, slot)` receives output channels o, o + 1 of pixel (y, x).
template <int CI, int NO, int KS> struct Core {
  typedef Geo<CI, NO, KS> G;
  static constexpr int KSTEPS = G::KSTEPS, NT = G::NT;
  uint32_t breg[KSTEPS][NT][2];
  int aoff[KSTEPS];

  // wsel(r, s, c, o) -> weight of tile channel c for output channel o at filter position (r, s), as bf16 bits (0 when out of range)
  template <typename WSel>
  __device__ __forceinline__ void setup(int Wp, WSel wsel) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int j = 0; j < KSTEPS; ++j) {
      aoff[j] = G::chunk_off(2 * j + (lane >> 4), Wp);
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          int r, s, c, o;
          const bool on = G::decode(2 * j + h2, 2 * t, nt * 8 + g, r, s, c, o);
          const uint32_t lo = on ? wsel(r, s, c, o) : 0u, hi = on ? wsel(r, s, c + 1, o) : 0u;
          breg[j][nt][h2] = lo | (hi << 16);
        }
    }
  }

  template <typename Pre, typename Epi>
  __device__ __forceinline__ void run(const bf16* tile, int H, int W, int Wp, Pre pre, Epi epi) const { run_w(tile, H, W, Wp, pre, epi, threadIdx.x >> 5); }
  // `warp`: index of this warp among the kWarps warps that share the image (a CTA may hold several such groups)
  template <typename Pre, typename Epi>
  __device__ __forceinline__ void run_w(const bf16* tile, int H, int W, int Wp, Pre pre, Epi epi, const int warp) const {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const uint32_t tile_u = smem_u32(tile);
    const int wlog = 31 - __clz(W);
    const int mtiles = G::PAIR ? H * W / 32 : H * W / 16;
    const int arow = ((lane >> 3) & 1) * 8 + (lane & 7);   // this lane's ldmatrix row of the m-tile
    for (int mt = warp; mt < mtiles; mt += kWarps) {
      int ya, xa, y0, x0, y1, x1;   // (ya, xa): this lane's A row; (y0, x0) / (y1, x1): the pixels (pair mode: even pixel of the pair) of accumulator rows g / g + 8
      if (G::PAIR) {
        const int qa = mt * 16 + arow, q0 = mt * 16 + g, q1 = q0 + 8, hl = wlog - 1, hm = (W >> 1) - 1;
        ya = qa >> hl; xa = (qa & hm) * 2; y0 = q0 >> hl; x0 = (q0 & hm) * 2; y1 = q1 >> hl; x1 = (q1 & hm) * 2;
      } else {
        const int pa = mt * 16 + arow, p0 = mt * 16 + g, p1 = p0 + 8;
        ya = pa >> wlog; xa = pa & (W - 1); y0 = p0 >> wlog; x0 = p0 & (W - 1); y1 = p1 >> wlog; x1 = p1 & (W - 1);
      }
      const uint32_t abase = tile_u + (uint32_t)((ya * Wp + xa) * G::PIXB);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + 2 * t, par = G::PAIR ? n / NO : 0, o = G::PAIR ? n % NO : n;
        pre(y0, x0 + par, o, 2 * nt); pre(y1, x1 + par, o, 2 * nt + 1);
      }
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int j = 0; j < KSTEPS; ++j) {
        uint32_t a[4];
        ldsm_x4(abase + aoff[j], a);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma16816(acc[nt], a, breg[j][nt][0], breg[j][nt][1]);
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + 2 * t, par = G::PAIR ? n / NO : 0, o = G::PAIR ? n % NO : n;
        epi(y0, x0 + par, o, acc[nt][0], acc[nt][1], 2 * nt);
        epi(y1, x1 + par, o, acc[nt][2], acc[nt][3], 2 * nt + 1);
      }
    }
  }
  // output channel of accumulator columns (nt, 2t), for per-thread column bookkeeping
  __device__ static __forceinline__ int out_channel(int nt, int t) { const int n = nt * 8 + 2 * t; return G::PAIR ? n % NO : n; }
};

// Per-column partial sums of a thread -> sum over the 8 lanes that share `t` (same columns) -> the warp's own slot part[warp][column][which] (a plain
// store: fp32 shared-memory atomics are compare-and-swap spin loops, ATOMS.CAST.SPIN — 30 % of the weight-gradient kernel before they were removed).
// `fold_cols` then adds the warps' slots of the GEMM columns that belong to output channel o (pair mode: both parities).
constexpr int kMaxN = 32;   // GEMM columns
__device__ __forceinline__ void reduce_cols_to_slot(float v, float* slot) {
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  if ((threadIdx.x & 31) < 4) *slot = v;
}
template <typename G> __device__ __forceinline__ float fold_cols(const float (*part)[kMaxN][2], int o, int which) {
  float v = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) {
    if (G::PAIR) {
#pragma unroll
      for (int n = 0; n < G::N; n += G::N / 2) v += part[w][n + o][which];
    } else {
      v += part[w][o][which];
    }
  }
  return v;
}

// ---- forward ----------------------------------------------------------------------------------------------------------------------------------
struct FwdArgs {
  int n, h, w, c_src, k_out, act, update_running; float slope;
  const bf16* x; const bf16* wgt; const float* bias; bf16* y;
  dcv_sc_norm xn, yn;
};

template <int CI, int NO, int KS>
__global__ void __launch_bounds__(kThreads, kMinCtas) sc_fwd_kernel(const FwdArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* tile = reinterpret_cast<bf16*>(smem_raw);
  typedef Core<CI, NO, KS> C;
  constexpr int NT = C::NT, PAD = KS / 2;
  __shared__ Coef cfx;
  __shared__ float sh_part[kWarps][kMaxN][2];
  const int H = a.h, W = a.w, Hp = H + KS - 1, Wp = W + KS - 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, t = lane & 3;
  SrcPlain src{a.x, a.xn.enabled ? &cfx : nullptr};
  Stager<CI, kThreads, SrcPlain> stager;
  for (int i = tid; i < Hp * Wp * CI / 8; i += kThreads) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0u, 0u, 0u, 0u);
  C core;
  const unsigned short* wb = reinterpret_cast<const unsigned short*>(a.wgt);
  core.setup(Wp, [&](int r, int s, int c, int o) -> uint32_t {
    return (c < a.c_src && o < a.k_out) ? (uint32_t)wb[((size_t)(o * KS + r) * KS + s) * a.c_src + c] : 0u;
  });
  float bias_r[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) { const int o = C::out_channel(nt, t) + e; bias_r[nt][e] = (a.bias && o < a.k_out) ? a.bias[o] : 0.f; }
  pdl_wait();      // the predecessor's output (x, its sums) may be read from here on
  pdl_trigger();
  stager.fetch(src, (size_t)blockIdx.x * H * W * a.c_src, H * W, a.c_src, tid);   // the first image's L2 round trip overlaps the coefficient prologue

  for (int img = blockIdx.x; img < a.n; img += gridDim.x) {
    const size_t e_img = (size_t)img * H * W * a.c_src;
    if (img != (int)blockIdx.x) { __syncthreads(); stager.fetch(src, e_img, H * W, a.c_src, tid); }   // the previous image's tile has been consumed
    if (tid < 32 && a.xn.enabled) norm_forward_coeffs(a.xn, img, cfx, a.update_running != 0 && img == 0);
    __syncthreads();   // coefficients ready, halo zeroed
    stager.store(tile, src, e_img, H, W, Wp, PAD, a.c_src, tid);
    __syncthreads();
    float s1[NT][2], s2[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;
    bf16* yimg = a.y + (size_t)img * H * W * a.k_out;
    core.run(tile, H, W, Wp, [](int, int, int, int) {}, [&](int y, int x, int o, float v0, float v1, int slot) {
      const int nt = slot >> 1;
      v0 = round_bf(act_fwd(v0 + bias_r[nt][0], a.act, a.slope));
      v1 = round_bf(act_fwd(v1 + bias_r[nt][1], a.act, a.slope));
      if (o < a.k_out) *reinterpret_cast<uint32_t*>(yimg + ((size_t)y * W + x) * a.k_out + o) = pack2(v0, v1);
      s1[nt][0] += v0; s1[nt][1] += v1; s2[nt][0] = fmaf(v0, v0, s2[nt][0]); s2[nt][1] = fmaf(v1, v1, s2[nt][1]);   // statistics of the STORED values
    });
    if (a.yn.enabled) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          reduce_cols_to_slot(s1[nt][e], &sh_part[warp][nt * 8 + 2 * t + e][0]);
          reduce_cols_to_slot(s2[nt][e], &sh_part[warp][nt * 8 + 2 * t + e][1]);
        }
      __syncthreads();
      if (tid < a.k_out) {   // this CTA owns the whole image: plain stores per (image, channel); the batch sums are sharded atomics
        const float v1 = fold_cols<typename C::G>(sh_part, tid, 0), v2 = fold_cols<typename C::G>(sh_part, tid, 1);
        a.yn.stats_nc[((size_t)img * a.k_out + tid) * 2] = v1;
        a.yn.stats_nc[((size_t)img * a.k_out + tid) * 2 + 1] = v2;
        if (a.yn.use_bn && a.yn.bn_training) {
          atomicAdd(a.yn.bn_sums + ((size_t)(img % kShards) * a.k_out + tid) * 2, v1);
          atomicAdd(a.yn.bn_sums + ((size_t)(img % kShards) * a.k_out + tid) * 2 + 1, v2);
        }
      }
    }
  }
}

// ---- data gradient ------------------------------------------------------------------------------------------------------------------------------
// dx[n][y][x][c] = sum_{k,r,s} dy[n][y + PAD - r][x + PAD - s][k] * w[k][r][s][c]: the forward core over the dy tile (KI = channels of dy) with the weights
// read transposed and flipped. dy is assembled while staging (SrcDy). When the layer's input is a producer's raw output with a pending normalisation
// (xn.enabled), dx is the gradient w.r.t. the NORMALISED input and the epilogue accumulates the producer's backward sums against its raw output x_raw.
struct DgradArgs {
  int n, h, w, c_in, k_out, act; float slope;
  const bf16* dz; const bf16* y; const bf16* wgt; bf16* dx; const bf16* x_raw;
  dcv_sc_norm yn, xn;
};

template <int KI, int NO, int KS>
__global__ void __launch_bounds__(kThreads, kMinCtas) sc_dgrad_kernel(const DgradArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* tile = reinterpret_cast<bf16*>(smem_raw);
  typedef Core<KI, NO, KS> C;
  constexpr int NT = C::NT, PAD = KS / 2;
  __shared__ Coef cfy, cfx;
  __shared__ float sh_part[kWarps][kMaxN][2];
  __shared__ float sh_s[kMaxC][2];
  const int H = a.h, W = a.w, Hp = H + KS - 1, Wp = W + KS - 1, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, t = lane & 3;
  SrcDy src{a.dz, a.y, a.yn.enabled ? &cfy : nullptr, a.act, a.slope};
  Stager<KI, kThreads, SrcDy> stager;
  for (int i = tid; i < Hp * Wp * KI / 8; i += kThreads) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0u, 0u, 0u, 0u);
  C core;
  const unsigned short* wb = reinterpret_cast<const unsigned short*>(a.wgt);
  core.setup(Wp, [&](int r, int s, int k, int o) -> uint32_t {   // tile channel k = output channel of the layer, GEMM column o = input channel of the layer
    return (k < a.k_out && o < a.c_in) ? (uint32_t)wb[((size_t)(k * KS + (KS - 1 - r)) * KS + (KS - 1 - s)) * a.c_in + o] : 0u;
  });
  pdl_wait();
  pdl_trigger();
  stager.fetch(src, (size_t)blockIdx.x * H * W * a.k_out, H * W, a.k_out, tid);
  for (int img = blockIdx.x; img < a.n; img += gridDim.x) {
    const size_t e_img = (size_t)img * H * W * a.k_out;
    if (img != (int)blockIdx.x) { __syncthreads(); stager.fetch(src, e_img, H * W, a.k_out, tid); }
    if (tid < 32 && a.yn.enabled) norm_backward_pqr(a.yn, img, cfy, false, nullptr, nullptr, nullptr, nullptr);
    __syncthreads();
    stager.store(tile, src, e_img, H, W, Wp, PAD, a.k_out, tid);
    __syncthreads();
    float s1[NT][2], s2[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;
    const size_t img0 = (size_t)img * H * W * a.c_in;
    uint32_t yraw[2 * NT];   // the producer's raw output at this thread's output elements: loaded BEFORE the tile's MMAs (an L2 round trip), used after them
    core.run(tile, H, W, Wp, [&](int y, int x, int o, int slot) {
      if (a.xn.enabled && o < a.c_in) yraw[slot] = *reinterpret_cast<const uint32_t*>(a.x_raw + img0 + ((size_t)y * W + x) * a.c_in + o);   // c_in is even here (check_norm)
    }, [&](int y, int x, int o, float v0, float v1, int slot) {
      if (o < a.c_in) {
        const size_t e = img0 + ((size_t)y * W + x) * a.c_in + o;
        v0 = round_bf(v0); v1 = round_bf(v1);
        if ((a.c_in & 1) == 0) *reinterpret_cast<uint32_t*>(a.dx + e) = pack2(v0, v1);
        else { a.dx[e] = __float2bfloat16_rn(v0); if (o + 1 < a.c_in) a.dx[e + 1] = __float2bfloat16_rn(v1); }   // odd channel counts (a 3-channel input image)
        if (a.xn.enabled) {
          const uint32_t yr = yraw[slot];
          const int nt = slot >> 1;
          s1[nt][0] += v0; s1[nt][1] += v1; s2[nt][0] = fmaf(v0, bf_lo(yr), s2[nt][0]); s2[nt][1] = fmaf(v1, bf_hi(yr), s2[nt][1]);
        }
      }
    });
    if (a.xn.enabled) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          reduce_cols_to_slot(s1[nt][e], &sh_part[warp][nt * 8 + 2 * t + e][0]);
          reduce_cols_to_slot(s2[nt][e], &sh_part[warp][nt * 8 + 2 * t + e][1]);
        }
      __syncthreads();
      if (tid < 32) {
        if (lane < a.c_in) { sh_s[lane][0] = fold_cols<typename C::G>(sh_part, lane, 0); sh_s[lane][1] = fold_cols<typename C::G>(sh_part, lane, 1); }
        __syncwarp();
        norm_backward_image_sums(a.xn, img, cfx, &sh_s[0][0]);
      }
    }
  }
}

// ---- weight gradient ----------------------------------------------------------------------------------------------------------------------------
// dw[k][r][s][c] = sum over (image, pixel) of dy[pix][k] * z[pix + (r, s) - PAD][c] as the GEMM D[(tap, c)][k] += Zt[(tap, c)][pix] * dy[pix][k]: M = chunks
// of 8 (tap, channel) rows (two per m-tile), K = pixels (pair mode: pixel pairs), N = output channels (pair mode: (parity, channel)). Both operands come
// TRANSPOSED out of NHWC tiles with ldmatrix.trans: z is the same halo tile the forward kernel builds (normalised while staging), dy is assembled from dz
// and y while staging (and summed into the bias gradient).
// A CTA is TWO groups of 4 warps; each group works on its own image with its own tiles and its own named barrier (the kernel is a chain of dependent
// round trips per image — fetch, coefficients, stage, MMA — so two images in flight per CTA halve the critical path without doubling the global atomics
// the way twice as many CTAs would). Inside a group the warps split K; every warp keeps its share of D in registers across all its images. At the end
// the 8 warps store their D tiles to shared memory (plain stores, the tiles' space is free by then) and each dw element adds up its <= 2 x 8 sources and
// issues one global atomic.
struct WgradArgs {
  int n, h, w, c_src, k_out, act; float slope;
  const bf16* x; const bf16* dz; const bf16* y;
  float* dw; float* dbias; float* d_bn_w; float* d_bn_b; float* d_gn_w; float* d_gn_b;
  dcv_sc_norm xn, yn;
};

constexpr int kGrp = 128, kGrpWarps = kGrp / 32;   // threads / warps of one image group of the weight-gradient CTA
__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, %1;" :: "r"(group + 1), "n"(kGrp) : "memory"); }

template <int CI, int NO, int KS> struct WgradLayout {
  typedef Geo<CI, NO, KS> G;
  static constexpr int MT = (G::NCH + 1) / 2, NT = G::NT, KOP = G::PAIR ? NO : (NO < 8 ? 8 : NO);   // KOP: channel stride of the dy tile
  __host__ __device__ static size_t group_bytes(int h, int w) { return (size_t)(h + KS - 1) * (w + KS - 1) * CI * 2 + (size_t)h * w * KOP * 2; }
  __host__ __device__ static size_t reduce_bytes() { return (size_t)kWgWarps * MT * 16 * (NT * 8) * 4; }
  __host__ __device__ static size_t smem_bytes(int h, int w) { const size_t a = 2 * group_bytes(h, w), b = reduce_bytes(); return a > b ? a : b; }
};

template <int CI, int NO, int KS>
__global__ void __launch_bounds__(kWgThreads, 2) sc_wgrad_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef Geo<CI, NO, KS> G;
  typedef WgradLayout<CI, NO, KS> L;
  constexpr int PAD = KS / 2, MT = L::MT, NT = L::NT, KOP = L::KOP, ND = NT * 8;
  constexpr int PF = MT * NT * 4 > 40 ? 2 : 4;   // vectors of a thread in flight across the prologue (z: PF, dy: 2 * PF registers x 4): fewer when the accumulators are many
  const int H = a.h, W = a.w, Hp = H + KS - 1, Wp = W + KS - 1, HW = H * W, wlog = 31 - __clz(W);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int group = warp / kGrpWarps, gtid = tid - group * kGrp, gwarp = warp - group * kGrpWarps;
  bf16* tile = reinterpret_cast<bf16*>(smem_raw + (size_t)group * L::group_bytes(H, W));   // [Hp][Wp][CI]
  bf16* dyt = tile + (size_t)Hp * Wp * CI;                                                  // [HW][KOP]
  __shared__ Coef cfx[2], cfy[2];
  __shared__ float sh_db[kMaxC];
  for (int i = gtid; i < Hp * Wp * CI / 8; i += kGrp) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = gtid; i < HW * KOP / 8; i += kGrp) reinterpret_cast<uint4*>(dyt)[i] = make_uint4(0u, 0u, 0u, 0u);   // padding channels stay zero
  if (tid < kMaxC) sh_db[tid] = 0.f;
  // per-lane ldmatrix geometry. A (x4.trans): matrix mi = lane / 8: chunk (mi & 1) of the m-tile, k half (mi >> 1); B: see below.
  const int mi = lane >> 3, li = lane & 7;
  int aoff[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) aoff[mt] = G::chunk_off(2 * mt + (mi & 1), Wp);
  const uint32_t tile_u = smem_u32(tile), dyt_u = smem_u32(dyt);
  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  float db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) db[i] = 0.f;
  SrcPlain sx{a.x, a.xn.enabled ? &cfx[group] : nullptr};
  SrcDy sd{a.dz, a.y, a.yn.enabled ? &cfy[group] : nullptr, a.act, a.slope};
  Stager<CI, kGrp, SrcPlain, PF> stx;
  const int kv = a.k_out;   // channels of dy
  const int nks = G::PAIR ? HW / 32 : HW / 16;
  pdl_wait();
  pdl_trigger();

  for (int img = blockIdx.x * 2 + group; img < a.n; img += gridDim.x * 2) {
    const size_t e_img = (size_t)img * HW * a.c_src, k_img = (size_t)img * HW * kv;
    group_sync(group);   // the group's previous image consumed / initial zeroing done
    stx.fetch(sx, e_img, HW, a.c_src, gtid);
    const bool dvec = (kv == 4 || kv % 8 == 0) && (k_img % 8) == 0;
    const int ndv = HW * kv / 8;
    SrcDy::Raw draw[PF];
    if (dvec) {
#pragma unroll
      for (int i = 0; i < PF; ++i) { const int v = gtid + i * kGrp; if (v < ndv) draw[i] = sd.fetch(k_img + (size_t)v * 8); }
    }
    if (gwarp == 0 && a.xn.enabled) load_coeffs(a.xn, img, cfx[group]);
    if (gwarp == 1 && a.yn.enabled) norm_backward_pqr(a.yn, img, cfy[group], img == 0, a.d_bn_w, a.d_bn_b, a.d_gn_w, a.d_gn_b);
    group_sync(group);
    // ---- stage z (halo tile) and dy
    stx.store(tile, sx, e_img, H, W, Wp, PAD, a.c_src, gtid);
    auto put_dy = [&](const SrcDy::Raw& r, int v) {
      const int e0 = v * 8, pix = e0 / kv, c0 = e0 % kv;
      float f[8];
      sd.decode(r, c0, kv - 1, f);   // kv is a power of two here (4, 8, 16, 32)
#pragma unroll
      for (int i = 0; i < 8; ++i) db[i] += round_bf(f[i]);
      if (kv == KOP) {
        *reinterpret_cast<uint4*>(dyt + e0) = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
      } else if (kv == 4) {   // two pixels of 4 channels into 8-channel rows
        *reinterpret_cast<uint2*>(dyt + (size_t)pix * KOP) = make_uint2(pack2(f[0], f[1]), pack2(f[2], f[3]));
        *reinterpret_cast<uint2*>(dyt + (size_t)(pix + 1) * KOP) = make_uint2(pack2(f[4], f[5]), pack2(f[6], f[7]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) dyt[(size_t)pix * KOP + c0 + i] = __float2bfloat16_rn(f[i]);
      }
    };
    if (dvec) {
#pragma unroll
      for (int i = 0; i < PF; ++i) { const int v = gtid + i * kGrp; if (v < ndv) put_dy(draw[i], v); }
      for (int v = gtid + PF * kGrp; v < ndv; v += kGrp) put_dy(sd.fetch(k_img + (size_t)v * 8), v);
    } else {
      for (int pix = gtid; pix < HW; pix += kGrp)
        for (int c = 0; c < kv; ++c) {
          const float v = round_bf(sd.load1(k_img + (size_t)pix * kv + c, c));
          dyt[(size_t)pix * KOP + c] = __float2bfloat16_rn(v);
          atomicAdd(&sh_db[c], v);
        }
    }
    group_sync(group);
    // ---- MMA over the image's k-steps (16 pixel pairs / pixels each), split over the group's warps
    for (int ks = gwarp; ks < nks; ks += kGrpWarps) {
      // A rows: the k index (pixel pair / pixel) ks * 16 + (mi >> 1) * 8 + li
      const int ka = ks * 16 + (mi >> 1) * 8 + li;
      int ya, xa;
      if (G::PAIR) { ya = ka >> (wlog - 1); xa = (ka & ((W >> 1) - 1)) * 2; } else { ya = ka >> wlog; xa = ka & (W - 1); }
      const uint32_t abase = tile_u + (uint32_t)((ya * Wp + xa) * G::PIXB);
      // B fragments of this k-step: rows = k index ks * 16 + (mi & 1) * 8 + li
      const int kb = ks * 16 + (mi & 1) * 8 + li;
      uint32_t b[NT][2];
      if (NT == 1) {             // pair mode, NO = 4: a row = pixel pair = (parity, 4 channels); plain mode: 8 (padded) channels of a pixel; x2: matrix = k half
        ldsm_x2_t(dyt_u + (uint32_t)(G::PAIR ? kb * 16 : kb * KOP * 2), b[0]);
      } else {                   // x4: matrices (n-tile 2h + (mi >> 1), k half mi & 1); pair mode: n-tile nt = (parity nt >> 1, channel half nt & 1)
#pragma unroll
        for (int h = 0; h < NT / 2; ++h) {
          const int nt = 2 * h + (mi >> 1);
          uint32_t r4[4];
          ldsm_x4_t(dyt_u + (uint32_t)(G::PAIR ? ((2 * kb + (nt >> 1)) * NO + (nt & 1) * 8) * 2 : (kb * KOP + nt * 8) * 2), r4);
          b[2 * h][0] = r4[0]; b[2 * h][1] = r4[1]; b[2 * h + 1][0] = r4[2]; b[2 * h + 1][1] = r4[3];
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t af[4];
        ldsm_x4_t(abase + aoff[mt], af);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma16816(acc[mt][nt], af, b[nt][0], b[nt][1]);
      }
    }
  }
  // ---- bias gradient: vector slot i of a thread is always the same channel: (c0 + i) & (kv - 1) with c0 = (gtid * 8) % kv (kGrp * 8 is a multiple of kv).
  // Lanes whose (lane * 8) % kv agree hold the same channels: butterfly over them, then one shared-memory atomic per (warp, channel).
  if ((kv == 4 || kv % 8 == 0)) {
    if (kv == 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { db[i] += db[i + 4]; db[i + 4] = 0.f; }
    }
    const int stride = kv >= 8 ? kv / 8 : 1, nval = kv == 4 ? 4 : 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nval) {
        float v = db[i];
        for (int o = 16; o >= stride; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < stride) atomicAdd(&sh_db[(lane * 8 + i) & (kv - 1)], v);
      }
    }
  }
  // ---- the 8 warps' D tiles -> shared memory (each element has one owner: plain stores), then one thread per dw element adds its sources up
  __syncthreads();   // every warp is done with the image tiles: their space is reused
  float* red = reinterpret_cast<float*>(smem_raw);   // [kWgWarps][MT * 16][ND]
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float* d0 = red + ((size_t)warp * MT * 16 + mt * 16 + g) * ND + nt * 8 + 2 * t;
      *reinterpret_cast<float2*>(d0) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
      *reinterpret_cast<float2*>(d0 + 8 * ND) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
    }
  __syncthreads();
  const int total = a.k_out * KS * KS * CI;   // compile-time divisors (see sc_bwd_kernel); tile channels beyond c_src are skipped
  for (int j = tid; j < total; j += kWgThreads) {
    const int c = j % CI, tap = (j / CI) % (KS * KS), o = j / (CI * KS * KS), r = tap / KS, s_ = tap % KS;
    if (c >= a.c_src) continue;
    const int i = (o * KS * KS + tap) * a.c_src + c;
    float v = 0.f;
    if (G::PAIR) {   // dw[s] = D[(r, s' = s, c)][(par 0, o)] + D[(r, s' = s + 1, c)][(par 1, o)]
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int sp = s_ + par, m = (r * G::CPR + (sp >> 1)) * 8 + (sp & 1) * 4 + c, n = par * NO + o;
        for (int w = 0; w < kWgWarps; ++w) v += red[((size_t)w * MT * 16 + m) * ND + n];
      }
    } else {
      const int m = tap * 16 + c;   // chunk 2 * tap + (c >> 3), element c & 7
      for (int w = 0; w < kWgWarps; ++w) v += red[((size_t)w * MT * 16 + m) * ND + o];
    }
    if (v != 0.f) atomicAdd(a.dw + i, v);
  }
  if (a.dbias && tid < a.k_out) atomicAdd(a.dbias + tid, sh_db[tid]);
}

// ---- backward of a block in ONE launch: data gradient + weight gradient ---------------------------------------------------------------------------
// Both gradients of a layer read the same dy = act'(y)*(P*dz + Q*y + R); as two kernels it was fetched, decoded and staged twice and the launch ramp /
// coefficient prologue paid twice (per layer: 23-35 us + 15-20 us). Here a group of 4 warps (two groups = two images per CTA, as in the weight-gradient
// kernel) stages the dy HALO tile once — the data gradient's A operand (ldmatrix) and, through per-lane row addresses into the same padded tile, the
// weight gradient's B operand (ldmatrix.trans) — plus the normalised input tile z, then runs the data-gradient GEMM (dx, the producer's backward sums)
// and the weight-gradient GEMM back to back. One image per group and launch (no persistent loop: the weight-gradient accumulators are only live after
// the data-gradient phase, so both phases fit the 128-register budget). Served: 4->4 5x5, 4->16 3x3, 16->16 3x3 (the layers of the default net with a
// data gradient); everything else runs the two kernels above.
struct BwdArgs {
  int n, h, w, c_in, k_out, act; float slope;
  const bf16* x; const bf16* dz; const bf16* y; const bf16* wgt; bf16* dx;
  float* dw; float* dbias; float* d_bn_w; float* d_bn_b; float* d_gn_w; float* d_gn_b;
  dcv_sc_norm xn, yn;
};

template <int CI, int KI, int KS> struct BwdLayout {
  typedef Geo<CI, KI, KS> WG;   // weight gradient: rows from the z tile (CI channels), columns = output channels (KI)
  static constexpr int MT = (WG::NCH + 1) / 2, NT = WG::NT, ND = NT * 8;
  __host__ __device__ static size_t group_bytes(int h, int w) { return (size_t)(h + KS - 1) * (w + KS - 1) * (CI + KI) * 2; }
  __host__ __device__ static size_t reduce_bytes() { return (size_t)kWgWarps * MT * 16 * ND * 4; }
  __host__ __device__ static size_t smem_bytes(int h, int w) { const size_t a = 2 * group_bytes(h, w), b = reduce_bytes(); return a > b ? a : b; }
};

template <int CI, int KI, int KS>
__global__ void __launch_bounds__(kWgThreads, 2) sc_bwd_kernel(const BwdArgs a) {
  static_assert((CI == 4 && KI == 4 && KS == 5) || (CI == 4 && KI == 16 && KS == 3) || (CI == 16 && KI == 16 && KS == 3), "combination not served");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef Core<KI, CI, KS> DC;      // data gradient: tile = dy (KI channels), GEMM columns = input channels
  typedef BwdLayout<CI, KI, KS> L;
  typedef typename L::WG WG;
  constexpr int PAD = KS / 2, DNT = DC::NT, MT = L::MT, NT = L::NT, ND = L::ND;
  const int H = a.h, W = a.w, Hp = H + KS - 1, Wp = W + KS - 1, HW = H * W, wlog = 31 - __clz(W);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int group = warp / kGrpWarps, gtid = tid - group * kGrp, gwarp = warp - group * kGrpWarps;
  bf16* ztile = reinterpret_cast<bf16*>(smem_raw + (size_t)group * L::group_bytes(H, W));   // [Hp][Wp][CI]
  bf16* dyt = ztile + (size_t)Hp * Wp * CI;                                                  // [Hp][Wp][KI]  (halo tile)
  __shared__ Coef cfx[2], cfy[2], cfp[2];
  __shared__ float sh_part[2][kWarps][kMaxN][2];
  __shared__ float sh_s[2][kMaxC][2];
  __shared__ float sh_db[kMaxC];
  for (int i = gtid; i < Hp * Wp * (CI + KI) / 8; i += kGrp) reinterpret_cast<uint4*>(ztile)[i] = make_uint4(0u, 0u, 0u, 0u);   // both tiles (contiguous)
  if (tid < kMaxC) sh_db[tid] = 0.f;
  DC core;
  const unsigned short* wb = reinterpret_cast<const unsigned short*>(a.wgt);
  core.setup(Wp, [&](int r, int s, int k, int o) -> uint32_t {
    return (k < a.k_out && o < a.c_in) ? (uint32_t)wb[((size_t)(k * KS + (KS - 1 - r)) * KS + (KS - 1 - s)) * a.c_in + o] : 0u;
  });
  const int mi = lane >> 3, li = lane & 7;
  int aoff[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) aoff[mt] = WG::chunk_off(2 * mt + (mi & 1), Wp);
  const uint32_t ztile_u = smem_u32(ztile), dyt_u = smem_u32(dyt);
  pdl_wait();
  pdl_trigger();
  const int img = blockIdx.x * 2 + group;
  const bool active = img < a.n;
  float db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) db[i] = 0.f;
  SrcPlain sx{a.x, a.xn.enabled ? &cfx[group] : nullptr};
  SrcDy sd{a.dz, a.y, a.yn.enabled ? &cfy[group] : nullptr, a.act, a.slope};
  if (active) {
    const size_t e_img = (size_t)img * HW * a.c_in, k_img = (size_t)img * HW * a.k_out;
    Stager<CI, kGrp, SrcPlain> stx;
    Stager<KI, kGrp, SrcDy, 2> std_;
    std_.vsum = db;
    stx.fetch(sx, e_img, HW, a.c_in, gtid);
    std_.fetch(sd, k_img, HW, a.k_out, gtid);
    if (gwarp == 0 && a.xn.enabled) load_coeffs(a.xn, img, cfx[group]);
    if (gwarp == 1 && a.yn.enabled) norm_backward_pqr(a.yn, img, cfy[group], img == 0, a.d_bn_w, a.d_bn_b, a.d_gn_w, a.d_gn_b);
    group_sync(group);   // coefficients ready, tiles zeroed
    stx.store(ztile, sx, e_img, H, W, Wp, PAD, a.c_in, gtid);
    std_.store(dyt, sd, k_img, H, W, Wp, PAD, a.k_out, gtid);
    group_sync(group);
    // ---- data gradient (+ the producer's backward sums)
    float s1[DNT][2], s2[DNT][2];
#pragma unroll
    for (int nt = 0; nt < DNT; ++nt) s1[nt][0] = s1[nt][1] = s2[nt][0] = s2[nt][1] = 0.f;
    uint32_t yraw[2 * DNT];
    core.run_w(dyt, H, W, Wp, [&](int y, int x, int o, int slot) {
      if (a.xn.enabled && o < a.c_in) yraw[slot] = *reinterpret_cast<const uint32_t*>(a.x + e_img + ((size_t)y * W + x) * a.c_in + o);
    }, [&](int y, int x, int o, float v0, float v1, int slot) {
      if (o < a.c_in) {
        const size_t e = e_img + ((size_t)y * W + x) * a.c_in + o;
        v0 = round_bf(v0); v1 = round_bf(v1);
        *reinterpret_cast<uint32_t*>(a.dx + e) = pack2(v0, v1);   // c_in is even on this path (host check)
        if (a.xn.enabled) {
          const uint32_t yr = yraw[slot];
          const int nt = slot >> 1;
          s1[nt][0] += v0; s1[nt][1] += v1; s2[nt][0] = fmaf(v0, bf_lo(yr), s2[nt][0]); s2[nt][1] = fmaf(v1, bf_hi(yr), s2[nt][1]);
        }
      }
    }, gwarp);
    if (a.xn.enabled) {
#pragma unroll
      for (int nt = 0; nt < DNT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          reduce_cols_to_slot(s1[nt][e], &sh_part[group][gwarp][nt * 8 + 2 * t + e][0]);
          reduce_cols_to_slot(s2[nt][e], &sh_part[group][gwarp][nt * 8 + 2 * t + e][1]);
        }
      group_sync(group);
      if (gwarp == 0) {
        if (lane < a.c_in) { sh_s[group][lane][0] = fold_cols<typename DC::G>(sh_part[group], lane, 0); sh_s[group][lane][1] = fold_cols<typename DC::G>(sh_part[group], lane, 1); }
        __syncwarp();
        norm_backward_image_sums(a.xn, img, cfp[group], &sh_s[group][0][0]);
      }
    }
  }
  // ---- weight gradient: D[(tap, c)][k] += z^T dy over the image's k-steps, split over the group's warps
  float acc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
  if (active) {
    const int nks = WG::PAIR ? HW / 32 : HW / 16;
    for (int ks = gwarp; ks < nks; ks += kGrpWarps) {
      const int ka = ks * 16 + (mi >> 1) * 8 + li, kb = ks * 16 + (mi & 1) * 8 + li;
      int ya, xa, yb, xb;
      if (WG::PAIR) { ya = ka >> (wlog - 1); xa = (ka & ((W >> 1) - 1)) * 2; yb = kb >> (wlog - 1); xb = (kb & ((W >> 1) - 1)) * 2; }
      else { ya = ka >> wlog; xa = ka & (W - 1); yb = kb >> wlog; xb = kb & (W - 1); }
      const uint32_t abase = ztile_u + (uint32_t)((ya * Wp + xa) * WG::PIXB);
      const uint32_t bpix = (uint32_t)((yb + PAD) * Wp + xb + PAD);   // the dy tile has a halo: pixel (y, x) sits at (y + PAD, x + PAD)
      uint32_t b[NT][2];
      if (NT == 1) {             // 4 -> 4 channels, 5x5: a row = pixel pair = (parity, 4 channels); PAD = 2 keeps it 16-byte aligned
        ldsm_x2_t(dyt_u + bpix * (KI * 2), b[0]);
      } else {
#pragma unroll
        for (int h = 0; h < NT / 2; ++h) {
          const int nt = 2 * h + (mi >> 1);
          uint32_t r4[4];
          ldsm_x4_t(dyt_u + (WG::PAIR ? ((bpix + (nt >> 1)) * KI + (nt & 1) * 8) * 2 : (bpix * KI + nt * 8) * 2), r4);
          b[2 * h][0] = r4[0]; b[2 * h][1] = r4[1]; b[2 * h + 1][0] = r4[2]; b[2 * h + 1][1] = r4[3];
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        uint32_t af[4];
        ldsm_x4_t(abase + aoff[mt], af);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma16816(acc[mt][nt], af, b[nt][0], b[nt][1]);
      }
    }
    // bias gradient: slot i of a thread is channel (c0 + i) & (kv - 1), c0 = (gtid * 8) % kv
    const int kv = a.k_out;
    if (kv == 4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { db[i] += db[i + 4]; db[i + 4] = 0.f; }
    }
    const int stride = kv >= 8 ? kv / 8 : 1, nval = kv == 4 ? 4 : 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nval) {
        float v = db[i];
        for (int o = 16; o >= stride; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < stride) atomicAdd(&sh_db[(lane * 8 + i) & (kv - 1)], v);
      }
    }
  }
  // ---- the 8 warps' D tiles -> shared memory (plain stores; the tiles' space is free), one thread per dw element adds its sources, one global atomic each
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem_raw);   // [kWgWarps][MT * 16][ND]
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float* d0 = red + ((size_t)warp * MT * 16 + mt * 16 + g) * ND + nt * 8 + 2 * t;
      *reinterpret_cast<float2*>(d0) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
      *reinterpret_cast<float2*>(d0 + 8 * ND) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
    }
  __syncthreads();
  // walk the (output channel, tap, TILE channel) index space: every divisor is a compile-time constant (dividing by the runtime channel count cost three
  // integer divisions per element: 11.6 % of this kernel's instructions in the ncu source view); tile channels beyond c_in are skipped
  const int total = a.k_out * KS * KS * CI;
  for (int j = tid; j < total; j += kWgThreads) {
    const int c = j % CI, tap = (j / CI) % (KS * KS), o = j / (CI * KS * KS), r = tap / KS, s_ = tap % KS;
    if (c >= a.c_in) continue;
    const int i = (o * KS * KS + tap) * a.c_in + c;
    float v = 0.f;
    if (WG::PAIR) {
#pragma unroll
      for (int par = 0; par < 2; ++par) {
        const int sp = s_ + par, m = (r * WG::CPR + (sp >> 1)) * 8 + (sp & 1) * 4 + c, n = par * KI + o;
        for (int w = 0; w < kWgWarps; ++w) v += red[((size_t)w * MT * 16 + m) * ND + n];
      }
    } else {
      const int m = tap * 16 + c;
      for (int w = 0; w < kWgWarps; ++w) v += red[((size_t)w * MT * 16 + m) * ND + o];
    }
    if (v != 0.f) atomicAdd(a.dw + i, v);
  }
  if (a.dbias && tid < a.k_out) atomicAdd(a.dbias + tid, sh_db[tid]);
}

// ---- normalise (+ average-pool) a raw output into a plain tensor: the consumer of a pending normalisation that is not one of the kernels above ----------
// zp[n][oy][ox][c] = A[n][c] * mean_{pool x pool}(y) + B[n][c]   (pooling is linear and the affine is per (image, channel): pool(z) = A*pool(y) + B).
struct PoolArgs { int n, h, w, c, pool, update_running; const bf16* y; bf16* z; const bf16* dzp; bf16* dz; dcv_sc_norm nd; };

__global__ void __launch_bounds__(kThreads, kMinCtas) sc_affine_pool_fwd_kernel(const PoolArgs a) {
  __shared__ Coef cf;
  const int tid = threadIdx.x, c2 = a.c / 2, oh = a.h / a.pool, ow = a.w / a.pool;
  const float inv = 1.f / (float)(a.pool * a.pool);
  pdl_wait();
  pdl_trigger();
  for (int img = blockIdx.x; img < a.n; img += gridDim.x) {
    __syncthreads();
    if (tid < 32) norm_forward_coeffs(a.nd, img, cf, a.update_running != 0 && img == 0);
    __syncthreads();
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.y + (size_t)img * a.h * a.w * a.c);
    uint32_t* dst = reinterpret_cast<uint32_t*>(a.z + (size_t)img * oh * ow * a.c);
    for (int i = tid; i < oh * ow * c2; i += kThreads) {
      const int cp = i % c2, opix = i / c2, oy = opix / ow, ox = opix - oy * ow;
      float v0 = 0.f, v1 = 0.f;
      for (int r = 0; r < a.pool; ++r)
        for (int s = 0; s < a.pool; ++s) {
          const uint32_t u = src[((size_t)(oy * a.pool + r) * a.w + ox * a.pool + s) * c2 + cp];
          v0 += bf_lo(u); v1 += bf_hi(u);
        }
      dst[i] = pack2(fmaf(cf.A[2 * cp], v0 * inv, cf.B[2 * cp]), fmaf(cf.A[2 * cp + 1], v1 * inv, cf.B[2 * cp + 1]));
    }
  }
}

// backward: dz = dzp / pool^2 spread over the window (written, bf16), s[n][c] = {sum dz, sum dz*y}, then the image's adjoint sums.
__global__ void __launch_bounds__(kThreads, kMinCtas) sc_affine_pool_bwd_kernel(const PoolArgs a) {
  __shared__ Coef cf;
  __shared__ float sh_s[kMaxC][2];
  __shared__ float sh_pw[kWarps][kMaxC][2];   // per-warp slots (plain stores instead of shared-memory atomics)
  bool slots = false;
  const int tid = threadIdx.x, c2 = a.c / 2, oh = a.h / a.pool, ow = a.w / a.pool;
  const float inv = 1.f / (float)(a.pool * a.pool);
  const bool fixed_cp = (kThreads % c2) == 0;   // a thread then always works on the same channel pair: sums stay in registers
  pdl_wait();
  pdl_trigger();
  for (int img = blockIdx.x; img < a.n; img += gridDim.x) {
    __syncthreads();
    if (tid < kMaxC * 2) (&sh_s[0][0])[tid] = 0.f;
    __syncthreads();
    const uint32_t* ysrc = reinterpret_cast<const uint32_t*>(a.y + (size_t)img * a.h * a.w * a.c);
    const uint32_t* gsrc = reinterpret_cast<const uint32_t*>(a.dzp + (size_t)img * oh * ow * a.c);
    uint32_t* dst = reinterpret_cast<uint32_t*>(a.dz + (size_t)img * a.h * a.w * a.c);
    float t1a = 0.f, t1b = 0.f, t2a = 0.f, t2b = 0.f;
    for (int i = tid; i < oh * ow * c2; i += kThreads) {
      const int cp = i % c2, opix = i / c2, oy = opix / ow, ox = opix - oy * ow;
      const uint32_t gu = gsrc[i];
      const float g0 = round_bf(bf_lo(gu) * inv), g1 = round_bf(bf_hi(gu) * inv);
      const uint32_t gp = pack2(g0, g1);
      float l1a = 0.f, l1b = 0.f, l2a = 0.f, l2b = 0.f;
      for (int r = 0; r < a.pool; ++r)
        for (int s = 0; s < a.pool; ++s) {
          const size_t e = ((size_t)(oy * a.pool + r) * a.w + ox * a.pool + s) * c2 + cp;
          const uint32_t yu = ysrc[e];
          dst[e] = gp;
          l1a += g0; l1b += g1; l2a = fmaf(g0, bf_lo(yu), l2a); l2b = fmaf(g1, bf_hi(yu), l2b);
        }
      if (fixed_cp) { t1a += l1a; t1b += l1b; t2a += l2a; t2b += l2b; }
      else { atomicAdd(&sh_s[2 * cp][0], l1a); atomicAdd(&sh_s[2 * cp + 1][0], l1b); atomicAdd(&sh_s[2 * cp][1], l2a); atomicAdd(&sh_s[2 * cp + 1][1], l2b); }
    }
    if (fixed_cp) {
      const int cp = tid % c2;
      if ((c2 & (c2 - 1)) == 0 && c2 <= 32) {   // lanes with equal lane % c2 hold the same channel pair: butterfly, then one atomic per (warp, value)
        for (int o = 16; o >= c2; o >>= 1) {
          t1a += __shfl_xor_sync(0xffffffffu, t1a, o); t1b += __shfl_xor_sync(0xffffffffu, t1b, o);
          t2a += __shfl_xor_sync(0xffffffffu, t2a, o); t2b += __shfl_xor_sync(0xffffffffu, t2b, o);
        }
        if ((tid & 31) < c2) { float (*pw)[2] = sh_pw[tid >> 5]; pw[2 * cp][0] = t1a; pw[2 * cp + 1][0] = t1b; pw[2 * cp][1] = t2a; pw[2 * cp + 1][1] = t2b; }
        slots = true;
      } else {
        atomicAdd(&sh_s[2 * cp][0], t1a); atomicAdd(&sh_s[2 * cp + 1][0], t1b); atomicAdd(&sh_s[2 * cp][1], t2a); atomicAdd(&sh_s[2 * cp + 1][1], t2b);
      }
    }
    __syncthreads();
    if (tid < 32) {
      if (slots && tid < a.c) {
        float v0 = 0.f, v1 = 0.f;
        for (int w = 0; w < kWarps; ++w) { v0 += sh_pw[w][tid][0]; v1 += sh_pw[w][tid][1]; }
        sh_s[tid][0] = v0; sh_s[tid][1] = v1;
      }
      __syncwarp();
      norm_backward_image_sums(a.nd, img, cf, &sh_s[0][0]);
    }
  }
}

// ---- host side --------------------------------------------------------------------------------------------------------------------------------
static int pick_ci(int c) { return c <= 4 ? 4 : (c == 16 ? 16 : 0); }

static size_t wgrad_smem(const dcv_conv_shape* s) {
  const int ci = pick_ci(s->c), no = pick_ci(s->k);
  if (s->r == 5) return WgradLayout<4, 4, 5>::smem_bytes(s->h, s->w);
  if (ci == 4 && no == 4) return WgradLayout<4, 4, 3>::smem_bytes(s->h, s->w);
  if (ci == 4) return WgradLayout<4, 16, 3>::smem_bytes(s->h, s->w);
  if (no == 4) return WgradLayout<16, 4, 3>::smem_bytes(s->h, s->w);
  return WgradLayout<16, 16, 3>::smem_bytes(s->h, s->w);
}

static bool shape_ok(const dcv_conv_shape* s, int dtype) {
  if (!s || dtype != DCV_BF16) return false;
  if (s->stride_h != 1 || s->stride_w != 1 || s->dil_h != 1 || s->dil_w != 1 || s->r != s->s || (s->r != 3 && s->r != 5)) return false;
  if (s->pad_h != s->r / 2 || s->pad_w != s->r / 2 || s->p != s->h || s->q != s->w) return false;
  if ((s->w != 16 && s->w != 32 && s->w != 64) || s->h > 64 || (s->h * s->w) % 32 != 0) return false;   // pixel indices split with shifts; 16 pixel pairs per m-tile
  const int ci = pick_ci(s->c), ki = pick_ci(s->k);
  if (!ci || !ki || s->k % 2 != 0 || s->c < 1) return false;
  if (s->r == 5 && (ci != 4 || ki != 4)) return false;   // 5x5 over 16 channels: 25 k-steps of resident weight fragments do not fit the register file
  if (wgrad_smem(s) > 110 * 1024) return false;          // two weight-gradient CTAs per SM up to ~105 KB, one above
  return true;
}

static int num_ctas(int n) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = kNumSMs; }
  return n < 4 * sms ? n : 4 * sms;
}

template <typename K> static int set_smem(K kern, size_t bytes) {
  if (bytes > 32 * 1024) {   // dynamic + static (coefficient structs, slots: up to ~8 KB) must stay under the 48 KB default
    DCV_REQUIRE(bytes <= 200 * 1024, "sc: %zu bytes of shared memory needed", bytes);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  }
  return 0;
}

static dcv_sc_norm norm_or_off(const dcv_sc_norm* nd) { dcv_sc_norm z{}; return nd ? *nd : z; }

static int check_norm(const dcv_sc_norm* nd, int n, int c, int hw, const char* what, bool backward) {
  if (!nd || !nd->enabled) return 0;
  DCV_REQUIRE(nd->n == n && nd->c == c && nd->hw == hw, "%s: normalisation descriptor is for [%d][%d][%d], tensor is [%d][%d][%d]", what, nd->n, nd->hw, nd->c, n, hw, c);
  DCV_REQUIRE(c <= kMaxC && c % 2 == 0, "%s: %d channels (this path serves even channel counts up to %d)", what, c, kMaxC);
  DCV_REQUIRE(nd->stats_nc && (!nd->use_bn || !nd->bn_training || nd->bn_sums), "%s: missing statistics buffers", what);
  DCV_REQUIRE(!nd->use_gn || (nd->gn_groups > 0 && c % nd->gn_groups == 0), "%s: num_channels=%d not divisible by num_groups=%d", what, c, nd->gn_groups);
  DCV_REQUIRE(!nd->use_bn || nd->bn_training || (nd->bn_running_mean && nd->bn_running_var), "%s: eval-mode BatchNorm needs running statistics", what);
  DCV_REQUIRE(!backward || (nd->u_sums && nd->coef_nc && nd->d_nc), "%s: missing backward buffers (u_sums, coef_nc, d_nc)", what);
  return 0;
}

}  // namespace sc
}  // namespace dcv

extern "C" {

int dcv_sc_conv_supported(const dcv_conv_shape* shape, int dtype) { return dcv::sc::shape_ok(shape, dtype) ? 1 : 0; }

size_t dcv_sc_norm_floats(int n, int c, int which) {
  /* which: 0 stats_nc, 1 bn_sums, 2 s_nc, 3 u_sums, 4 coef_nc, 5 d_nc */
  switch (which) {
    case 0: case 2: return (size_t)n * c * 2;
    case 1: return (size_t)dcv::sc::kShards * c * 2;
    case 3: return (size_t)dcv::sc::kShards * c * 4;
    case 4: return (size_t)n * c * 8;
    default: return (size_t)n * c * 4;
  }
}

#define SC_DISPATCH(CI_, NT_, KS_, KERN, ARGS, SMEM)                                                                \
  do {                                                                                                              \
    auto kern = KERN<CI_, NT_, KS_>;                                                                                \
    if (set_smem(kern, SMEM)) return 1;                                                                             \
    launch_pdl(kern, num_ctas((ARGS).n), kThreads, SMEM, st, ARGS);                                                 \
  } while (0)

int dcv_sc_conv_fwd(const dcv_conv_shape* s, const void* x, const dcv_sc_norm* x_norm, int update_running, const void* w, const float* bias, int act, float slope,
                    void* y, const dcv_sc_norm* y_norm, void* stream) {
  using namespace dcv; using namespace dcv::sc;
  DCV_REQUIRE(shape_ok(s, DCV_BF16), "sc_conv_fwd: shape not served by the few-channel kernels (see dcv_sc_conv_supported)");
  DCV_REQUIRE(x && w && y, "sc_conv_fwd: null pointer");
  if (check_norm(x_norm, s->n, s->c, s->h * s->w, "sc_conv_fwd (input)", false) || check_norm(y_norm, s->n, s->k, s->h * s->w, "sc_conv_fwd (output)", false)) return 1;
  cudaStream_t st = as_stream(stream);
  FwdArgs a{};
  a.n = s->n; a.h = s->h; a.w = s->w; a.c_src = s->c; a.k_out = s->k; a.act = act; a.slope = slope; a.update_running = update_running;
  a.x = (const bf16*)x; a.wgt = (const bf16*)w; a.bias = bias; a.y = (bf16*)y; a.xn = norm_or_off(x_norm); a.yn = norm_or_off(y_norm);
  const int ci = pick_ci(s->c), no = pick_ci(s->k);
  const size_t smem = (size_t)(s->h + s->r - 1) * (s->w + s->r - 1) * ci * 2;
  if (s->r == 5) SC_DISPATCH(4, 4, 5, sc_fwd_kernel, a, smem);
  else if (ci == 4 && no == 4) SC_DISPATCH(4, 4, 3, sc_fwd_kernel, a, smem);
  else if (ci == 4) SC_DISPATCH(4, 16, 3, sc_fwd_kernel, a, smem);
  else if (no == 4) SC_DISPATCH(16, 4, 3, sc_fwd_kernel, a, smem);
  else SC_DISPATCH(16, 16, 3, sc_fwd_kernel, a, smem);
  DCV_LAUNCH_CHECK("sc_fwd_kernel");
  return 0;
}

int dcv_sc_conv_dgrad(const dcv_conv_shape* s, const void* dz, const void* y, const dcv_sc_norm* y_norm, int act, float slope, const void* w, void* dx,
                      const void* x_raw, const dcv_sc_norm* x_norm, void* stream) {
  using namespace dcv; using namespace dcv::sc;
  DCV_REQUIRE(shape_ok(s, DCV_BF16), "sc_conv_dgrad: shape not served by the few-channel kernels");
  DCV_REQUIRE(dz && y && w && dx, "sc_conv_dgrad: null pointer");
  if (check_norm(y_norm, s->n, s->k, s->h * s->w, "sc_conv_dgrad (output)", true) || check_norm(x_norm, s->n, s->c, s->h * s->w, "sc_conv_dgrad (input)", true)) return 1;
  DCV_REQUIRE(!(x_norm && x_norm->enabled) || x_raw, "sc_conv_dgrad: the producer's raw output is needed for its backward sums");
  cudaStream_t st = as_stream(stream);
  DgradArgs a{};
  a.n = s->n; a.h = s->h; a.w = s->w; a.c_in = s->c; a.k_out = s->k; a.act = act; a.slope = slope;
  a.dz = (const bf16*)dz; a.y = (const bf16*)y; a.wgt = (const bf16*)w; a.dx = (bf16*)dx; a.x_raw = (const bf16*)x_raw; a.yn = norm_or_off(y_norm); a.xn = norm_or_off(x_norm);
  const int ki = pick_ci(s->k), no = pick_ci(s->c);
  const size_t smem = (size_t)(s->h + s->r - 1) * (s->w + s->r - 1) * ki * 2;
  if (s->r == 5) SC_DISPATCH(4, 4, 5, sc_dgrad_kernel, a, smem);
  else if (ki == 4 && no == 4) SC_DISPATCH(4, 4, 3, sc_dgrad_kernel, a, smem);
  else if (ki == 4) SC_DISPATCH(4, 16, 3, sc_dgrad_kernel, a, smem);
  else if (no == 4) SC_DISPATCH(16, 4, 3, sc_dgrad_kernel, a, smem);
  else SC_DISPATCH(16, 16, 3, sc_dgrad_kernel, a, smem);
  DCV_LAUNCH_CHECK("sc_dgrad_kernel");
  return 0;
}

int dcv_sc_conv_wgrad(const dcv_conv_shape* s, const void* x, const dcv_sc_norm* x_norm, const void* dz, const void* y, const dcv_sc_norm* y_norm, int act, float slope,
                      float* dw, float* dbias, float* d_bn_w, float* d_bn_b, float* d_gn_w, float* d_gn_b, void* stream) {
  using namespace dcv; using namespace dcv::sc;
  DCV_REQUIRE(shape_ok(s, DCV_BF16), "sc_conv_wgrad: shape not served by the few-channel kernels");
  DCV_REQUIRE(x && dz && y && dw, "sc_conv_wgrad: null pointer");
  if (check_norm(x_norm, s->n, s->c, s->h * s->w, "sc_conv_wgrad (input)", false) || check_norm(y_norm, s->n, s->k, s->h * s->w, "sc_conv_wgrad (output)", true)) return 1;
  cudaStream_t st = as_stream(stream);
  WgradArgs a{};
  a.n = s->n; a.h = s->h; a.w = s->w; a.c_src = s->c; a.k_out = s->k; a.act = act; a.slope = slope;
  a.x = (const bf16*)x; a.dz = (const bf16*)dz; a.y = (const bf16*)y; a.dw = dw; a.dbias = dbias; a.d_bn_w = d_bn_w; a.d_bn_b = d_bn_b; a.d_gn_w = d_gn_w; a.d_gn_b = d_gn_b;
  a.xn = norm_or_off(x_norm); a.yn = norm_or_off(y_norm);
  const int ci = pick_ci(s->c), no = pick_ci(s->k);
  const size_t smem = wgrad_smem(s);
  // persistent over images, two at a time: fewer CTAs = fewer atomics on dw
  auto grid_of = [&](int n) { const int g = num_ctas((n + 1) / 2); return g < 2 * kNumSMs ? g : 2 * kNumSMs; };   // two images in flight per CTA
#define SC_WGRAD(CI_, NT_, KS_)                                                         \
  do {                                                                                  \
    auto kern = sc_wgrad_kernel<CI_, NT_, KS_>;                                         \
    if (set_smem(kern, smem)) return 1;                                                 \
    launch_pdl(kern, grid_of(a.n), kWgThreads, smem, st, a);                              \
  } while (0)
  if (s->r == 5) SC_WGRAD(4, 4, 5);
  else if (ci == 4 && no == 4) SC_WGRAD(4, 4, 3);
  else if (ci == 4) SC_WGRAD(4, 16, 3);
  else if (no == 4) SC_WGRAD(16, 4, 3);
  else SC_WGRAD(16, 16, 3);
#undef SC_WGRAD
  DCV_LAUNCH_CHECK("sc_wgrad_kernel");
  return 0;
}

static int bwd_combo(const dcv_conv_shape* s) {   // 1: <4,4,5>  2: <4,16,3>  3: <16,16,3>  0: not served by the fused backward kernel
  using namespace dcv::sc;
  if (!shape_ok(s, DCV_BF16) || s->c % 2 != 0) return 0;
  const int ci = pick_ci(s->c), ki = pick_ci(s->k);
  if (s->k != ki) return 0;   // dy is staged with whole 16-byte vectors
  int combo = 0;
  size_t smem = 0;
  if (ci == 4 && ki == 4 && s->r == 5) { combo = 1; smem = BwdLayout<4, 4, 5>::smem_bytes(s->h, s->w); }
  else if (ci == 4 && ki == 16 && s->r == 3) { combo = 2; smem = BwdLayout<4, 16, 3>::smem_bytes(s->h, s->w); }
  else if (ci == 16 && ki == 16 && s->r == 3) { combo = 3; smem = BwdLayout<16, 16, 3>::smem_bytes(s->h, s->w); }
  return (combo && smem <= 100 * 1024) ? combo : 0;
}

int dcv_sc_conv_bwd_supported(const dcv_conv_shape* shape, int dtype) { return (shape && dtype == DCV_BF16) ? (bwd_combo(shape) != 0) : 0; }

int dcv_sc_conv_bwd(const dcv_conv_shape* s, const void* x, const dcv_sc_norm* x_norm, const void* dz, const void* y, const dcv_sc_norm* y_norm, int act, float slope, const void* w,
                    void* dx, float* dw, float* dbias, float* d_bn_w, float* d_bn_b, float* d_gn_w, float* d_gn_b, void* stream) {
  using namespace dcv; using namespace dcv::sc;
  const int combo = s ? bwd_combo(s) : 0;
  DCV_REQUIRE(combo != 0, "sc_conv_bwd: shape not served by the fused backward kernel (see dcv_sc_conv_bwd_supported)");
  DCV_REQUIRE(x && dz && y && w && dx && dw, "sc_conv_bwd: null pointer");
  if (check_norm(x_norm, s->n, s->c, s->h * s->w, "sc_conv_bwd (input)", true) || check_norm(y_norm, s->n, s->k, s->h * s->w, "sc_conv_bwd (output)", true)) return 1;
  cudaStream_t st = as_stream(stream);
  BwdArgs a{};
  a.n = s->n; a.h = s->h; a.w = s->w; a.c_in = s->c; a.k_out = s->k; a.act = act; a.slope = slope;
  a.x = (const bf16*)x; a.dz = (const bf16*)dz; a.y = (const bf16*)y; a.wgt = (const bf16*)w; a.dx = (bf16*)dx;
  a.dw = dw; a.dbias = dbias; a.d_bn_w = d_bn_w; a.d_bn_b = d_bn_b; a.d_gn_w = d_gn_w; a.d_gn_b = d_gn_b;
  a.xn = norm_or_off(x_norm); a.yn = norm_or_off(y_norm);
  const int grid = (s->n + 1) / 2;
#define SC_BWD(CI_, KI_, KS_)                                                              \
  do {                                                                                     \
    auto kern = sc_bwd_kernel<CI_, KI_, KS_>;                                              \
    const size_t smem = BwdLayout<CI_, KI_, KS_>::smem_bytes(s->h, s->w);                  \
    if (set_smem(kern, smem)) return 1;                                                    \
    launch_pdl(kern, grid, kWgThreads, smem, st, a);                                       \
  } while (0)
  if (combo == 1) SC_BWD(4, 4, 5);
  else if (combo == 2) SC_BWD(4, 16, 3);
  else SC_BWD(16, 16, 3);
#undef SC_BWD
  DCV_LAUNCH_CHECK("sc_bwd_kernel");
  return 0;
}

int dcv_sc_affine_pool_fwd(const void* y, const dcv_sc_norm* norm, int update_running, void* z, int n, int h, int w, int c, int pool, void* stream) {
  using namespace dcv; using namespace dcv::sc;
  DCV_REQUIRE(y && z && norm && norm->enabled && n > 0 && pool >= 1 && h % pool == 0 && w % pool == 0, "sc_affine_pool_fwd: bad arguments");
  if (check_norm(norm, n, c, h * w, "sc_affine_pool_fwd", false)) return 1;
  PoolArgs a{}; a.n = n; a.h = h; a.w = w; a.c = c; a.pool = pool; a.update_running = update_running; a.y = (const bf16*)y; a.z = (bf16*)z; a.nd = *norm;
  launch_pdl(sc_affine_pool_fwd_kernel, num_ctas(n), kThreads, 0, as_stream(stream), a);
  DCV_LAUNCH_CHECK("sc_affine_pool_fwd_kernel");
  return 0;
}

int dcv_sc_affine_pool_bwd(const void* dzp, const void* y, const dcv_sc_norm* norm, void* dz, int n, int h, int w, int c, int pool, void* stream) {
  using namespace dcv; using namespace dcv::sc;
  DCV_REQUIRE(dzp && y && dz && norm && norm->enabled && n > 0 && pool >= 1 && h % pool == 0 && w % pool == 0, "sc_affine_pool_bwd: bad arguments");
  if (check_norm(norm, n, c, h * w, "sc_affine_pool_bwd", true)) return 1;
  PoolArgs a{}; a.n = n; a.h = h; a.w = w; a.c = c; a.pool = pool; a.y = (const bf16*)y; a.dzp = (const bf16*)dzp; a.dz = (bf16*)dz; a.nd = *norm;
  launch_pdl(sc_affine_pool_bwd_kernel, num_ctas(n), kThreads, 0, as_stream(stream), a);
  DCV_LAUNCH_CHECK("sc_affine_pool_bwd_kernel");
  return 0;
}

}  // extern "C"
