// uint8 -> float normalise + horizontal flip + pad-crop in ONE HBM-bound kernel.
//
// Geometry (bit-exact index selection): for output pixel (i, j) of image n
//   jj = flip[n] ? out_w-1-j : j;   r = top[n] + i - pad;   col = left[n] + jj - pad
//   value = inside(r, col) ? src[n][r][col][ch] : uint8 0,   then  ((value / 255) - mean[ch]) / std[ch]  in fp32.
// The 256-entry-per-channel table of that last expression is computed once per CTA with IEEE fp32 division, in the same
// operation order as torchvision's ToTensor -> Normalize, so the fp32 results are bit-identical to the CPU path and the
// bf16 results are the round-to-nearest-even of them. For large batches the table is replicated once per shared-memory
// bank (lane l only ever touches bank l), which makes the data-dependent lookups conflict-free.
//
// Data movement per CTA iteration ("row group" = RG consecutive output rows of one image):
//   1. the cropped, zero-padded source bytes of the group are staged in shared memory with coalesced 32-bit global loads
//      (two aligned words funnel-shifted into each shared word, so arbitrary crop offsets stay coalesced);
//   2. every thread then emits whole 16-byte output vectors (8 bf16 / 4 fp32) with fully coalesced stores.
// Algorithmic bytes per output pixel: c (uint8 in) + c_out * sizeof(out).
#include "common.cuh"

namespace dcv {

struct PreprocessArgs {
  const uint8_t* src; void* dst;
  int n, h, w, c, out_h, out_w, pad, c_out, nchw_out, rg;   // rg: output rows per group
  const float* mean; const float* stdv; const uint8_t* flip; const int32_t* crop_yx;
  FastDiv div_cout, div_vec_per_row, div_groups_per_img;
  int vec_per_row;      // 16-byte vectors per output row (NHWC: out_w*c_out/VE; NCHW: out_w/VE)
  int row_bytes_smem;   // padded bytes per staged row
  size_t src_total_bytes;
  int wide_store;       // dst is 32-byte aligned: fp32 rows may be written with 256-bit stores
};

template <typename T, int R>
__global__ void __launch_bounds__(256) preprocess_kernel(const PreprocessArgs a) {
  constexpr int VE = 16 / sizeof(T);
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* lut = reinterpret_cast<float*>(smem_raw);                       // [c][256][R]
  uint8_t* rows = smem_raw + (size_t)a.c * 256 * R * sizeof(float);      // [rg][row_bytes_smem]
  const int tid = threadIdx.x, lane = tid & 31;

  for (int e = tid; e < a.c * 256; e += blockDim.x) {
    const int ch = e >> 8, v = e & 255;
    const float val = ((float)v / 255.0f - a.mean[ch]) / a.stdv[ch];   // ToTensor then Normalize, same op order
#pragma unroll
    for (int r = 0; r < R; ++r) lut[e * R + r] = val;
  }

  const int groups_per_img = (a.out_h + a.rg - 1) / a.rg;
  const int total_groups = a.n * groups_per_img;
  const int in_row_bytes = a.w * a.c;
  const int crop_row_bytes = a.out_w * a.c;
  const int words_per_row = (crop_row_bytes + 3) >> 2;
  const uintptr_t src_base = reinterpret_cast<uintptr_t>(a.src);

  for (int g = blockIdx.x; g < total_groups; g += gridDim.x) {
    const int img = a.div_groups_per_img.div(g);
    const int row0 = (g - img * groups_per_img) * a.rg;
    const int nrows = min(a.rg, a.out_h - row0);
    const int top = a.crop_yx ? a.crop_yx[2 * img] : a.pad;
    const int left = a.crop_yx ? a.crop_yx[2 * img + 1] : a.pad;
    const bool flip = a.flip ? (a.flip[img] != 0) : false;
    const int shift = (left - a.pad) * a.c;   // cropped byte b of a row comes from source-row byte shift + b

    __syncthreads();  // previous group's readers are done with `rows` (also orders the LUT fill on the first pass)
    // ---- stage cropped source bytes: one 32-bit shared word per thread-iteration
    for (int wi = tid; wi < nrows * words_per_row; wi += blockDim.x) {
      const int rr = wi / words_per_row, wcol = wi - rr * words_per_row;
      const int r_src = top + row0 + rr - a.pad;
      uint32_t out_word = 0;
      if (r_src >= 0 && r_src < a.h) {
        const int b0 = wcol * 4;                       // first cropped byte of this word
        const long long o = (long long)shift + b0;     // source-row byte offset of it (may be negative)
        if (o + 3 >= 0 && o < in_row_bytes) {
          const size_t row_off = ((size_t)img * a.h + r_src) * (size_t)in_row_bytes;
          const long long abs0 = (long long)row_off + o;                 // absolute byte offset in src (>= -3)
          const uintptr_t addr = src_base + abs0;                        // wraps harmlessly when abs0 < 0: guarded below
          const uintptr_t aligned = addr & ~uintptr_t(3);
          const int mis = (int)(addr & 3);
          const long long abs_aligned = abs0 - mis;
          uint32_t lo = 0, hi = 0;
          if (abs_aligned + 3 >= 0 && abs_aligned < (long long)a.src_total_bytes) lo = *reinterpret_cast<const uint32_t*>(aligned);
          if (mis && abs_aligned + 4 < (long long)a.src_total_bytes) hi = *reinterpret_cast<const uint32_t*>(aligned + 4);
          uint32_t word = __funnelshift_r(lo, hi, mis * 8);
          // zero the bytes that fall outside [0, in_row_bytes) of this source row (left / right padding)
          uint32_t mask = 0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const long long ok = o + k;
            if (ok >= 0 && ok < in_row_bytes && b0 + k < crop_row_bytes) mask |= 0xffu << (8 * k);
          }
          out_word = word & mask;
        }
      }
      *reinterpret_cast<uint32_t*>(rows + (size_t)rr * a.row_bytes_smem + wcol * 4) = out_word;
    }
    __syncthreads();

    // ---- emit 16-byte output vectors
    const int total_vec = nrows * a.vec_per_row * (a.nchw_out ? a.c : 1);
    for (int v = tid; v < total_vec; v += blockDim.x) {
      float vals[VE];
      size_t dst_elem;
      if (!a.nchw_out) {
        const int rr = a.div_vec_per_row.div(v);
        const int e0 = (v - rr * a.vec_per_row) * VE;     // first element (pixel*c_out + ch) of this vector in its row
        const uint8_t* rb = rows + (size_t)rr * a.row_bytes_smem;
#pragma unroll
        for (int k = 0; k < VE; ++k) {
          const int e = e0 + k;
          const int j = a.div_cout.div(e), ch = e - j * a.c_out;
          float val = 0.f;
          if (ch < a.c) {
            const int jj = flip ? (a.out_w - 1 - j) : j;
            const int byte = rb[jj * a.c + ch];
            val = lut[(ch * 256 + byte) * R + (R == 1 ? 0 : lane)];
          }
          vals[k] = val;
        }
        dst_elem = (((size_t)img * a.out_h + row0 + rr) * a.out_w) * a.c_out + e0;
      } else {
        const int per_ch = nrows * a.vec_per_row;
        const int ch = v / per_ch, rem = v - ch * per_ch;
        const int rr = a.div_vec_per_row.div(rem);
        const int j0 = (rem - rr * a.vec_per_row) * VE;
        const uint8_t* rb = rows + (size_t)rr * a.row_bytes_smem;
#pragma unroll
        for (int k = 0; k < VE; ++k) {
          const int j = j0 + k;
          const int jj = flip ? (a.out_w - 1 - j) : j;
          const int byte = rb[jj * a.c + ch];
          vals[k] = lut[(ch * 256 + byte) * R + (R == 1 ? 0 : lane)];
        }
        dst_elem = (((size_t)img * a.c + ch) * a.out_h + row0 + rr) * a.out_w + j0;
      }
      *reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.dst) + dst_elem) = vec_pack<T>(vals);
    }
  }
}


// Streaming path for NHWC output with c_out == c and out_w % 8 == 0 (every shape of the benchmark configs): no shared-memory staging, no block
// barriers. A thread produces 8 consecutive output pixels: their 8 source pixels are contiguous in the source row (reversed when flipped), so the 8*C
// source bytes are fetched as 2C+1 aligned 32-bit words (consecutive lanes read consecutive 8*C-byte windows => coalesced) and funnel-shifted to the
// byte offset the crop asks for. The channel of every element is a compile-time constant; values come from the 256-entry-per-channel table
// ((v/255 - mean)/std evaluated with IEEE division in torchvision's operation order, R bank-spread replicas). Windows that touch the zero padding
// take a per-pixel path (only at the crop border).
template <typename T, int C, int R>
__global__ void __launch_bounds__(512) preprocess_stream_kernel(const PreprocessArgs a, const size_t total_items) {
  constexpr int VE = 16 / sizeof(T), NE = 8 * C, NW = 2 * C, NL = (8 * C + 30) / 16;   // NL: 128-bit loads covering any alignment of the window
  extern __shared__ __align__(16) float lut[];   // [C][256][R]
  const int tid = threadIdx.x, lane = tid & 31;
  for (int e = tid; e < C * 256; e += blockDim.x) {
    const int ch = e >> 8, v = e & 255;
    const float val = ((float)v / 255.0f - a.mean[ch]) / a.stdv[ch];   // ToTensor then Normalize, same op order
#pragma unroll
    for (int r = 0; r < R; ++r) lut[e * R + r] = val;
  }
  __syncthreads();
  const int gpr = a.out_w >> 3;
  const size_t in_row_bytes = (size_t)a.w * C;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + tid; it < total_items; it += (size_t)gridDim.x * blockDim.x) {
    const uint32_t row_lin = (uint32_t)(it / gpr);
    const int gx = (int)(it - (size_t)row_lin * gpr);
    const int img = a.div_vec_per_row.div(row_lin), i = row_lin - img * a.out_h;        // div_vec_per_row holds out_h here
    const int top = a.crop_yx ? a.crop_yx[2 * img] : a.pad;
    const int left = a.crop_yx ? a.crop_yx[2 * img + 1] : a.pad;
    const bool flip = a.flip ? (a.flip[img] != 0) : false;
    const int r_src = top + i - a.pad;
    const int j0 = gx << 3;
    const int px0 = (flip ? (a.out_w - 8 - j0) : j0) + left - a.pad;                     // first (lowest-address) source pixel of the window
    uint32_t w[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = 0u;
    if (r_src >= 0 && r_src < a.h) {
      const uint8_t* row = a.src + ((size_t)img * a.h + r_src) * in_row_bytes;
      const size_t win_off = ((size_t)img * a.h + r_src) * in_row_bytes + (size_t)(px0 > 0 ? px0 : 0) * C;
      if (px0 >= 0 && px0 + 8 <= a.w && (win_off & ~size_t(15)) + 16 * NL <= a.src_total_bytes) {
        // NL aligned 128-bit loads cover the 8*C-byte window at any byte alignment (the ncu profile of the 32-bit version showed the L1TEX pipe at
        // 90 %: 6.6 sectors fetched per useful sector); the window is then funnel-shifted out of them. `mw` takes at most two values in a warp.
        const uintptr_t addr = reinterpret_cast<uintptr_t>(a.src) + win_off;
        const uint4* al = reinterpret_cast<const uint4*>(addr & ~uintptr_t(15));
        const int mw = (int)(addr & 15) >> 2, mis = (int)(addr & 3) * 8;
        uint32_t raw[4 * NL + 1];
#pragma unroll
        for (int k = 0; k < NL; ++k) { const uint4 v4 = __ldg(al + k); raw[4 * k] = v4.x; raw[4 * k + 1] = v4.y; raw[4 * k + 2] = v4.z; raw[4 * k + 3] = v4.w; }
        raw[4 * NL] = 0u;
        switch (mw) {
          case 0:
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = __funnelshift_r(raw[k], raw[k + 1], mis);
            break;
          case 1:
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = __funnelshift_r(raw[k + 1], raw[k + 2], mis);
            break;
          case 2:
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = __funnelshift_r(raw[k + 2], raw[k + 3], mis);
            break;
          default:
#pragma unroll
            for (int k = 0; k < NW; ++k) w[k] = __funnelshift_r(raw[k + 3], raw[k + 4], mis);
            break;
        }
      } else {
#pragma unroll
        for (int pxl = 0; pxl < 8; ++pxl) {
          const int px = px0 + pxl;
          if (px >= 0 && px < a.w) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
              const int kb = pxl * C + ch;
              w[kb >> 2] |= (uint32_t)row[(size_t)px * C + ch] << (8 * (kb & 3));
            }
          }
        }
      }
    }
    float vals[NE];
    if (flip) {
#pragma unroll
      for (int pxl = 0; pxl < 8; ++pxl)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          const int kb = (7 - pxl) * C + ch;
          vals[pxl * C + ch] = lut[(ch * 256 + ((w[kb >> 2] >> (8 * (kb & 3))) & 0xffu)) * R + (R == 1 ? 0 : (lane & (R - 1)))];
        }
    } else {
#pragma unroll
      for (int pxl = 0; pxl < 8; ++pxl)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          const int kb = pxl * C + ch;
          vals[pxl * C + ch] = lut[(ch * 256 + ((w[kb >> 2] >> (8 * (kb & 3))) & 0xffu)) * R + (R == 1 ? 0 : (lane & (R - 1)))];
        }
    }
    T* dst = reinterpret_cast<T*>(a.dst) + ((size_t)row_lin * a.out_w + j0) * C;
    if constexpr (sizeof(T) == 4) {
      if (a.wide_store) {   // fp32: 8*C floats per thread = C 256-bit streaming stores (sm_100 STG.256): half the store requests of the 128-bit version
#pragma unroll
        for (int v = 0; v < NE / 8; ++v)
          asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + v * 8), "f"(vals[v * 8]), "f"(vals[v * 8 + 1]), "f"(vals[v * 8 + 2]), "f"(vals[v * 8 + 3]),
                       "f"(vals[v * 8 + 4]), "f"(vals[v * 8 + 5]), "f"(vals[v * 8 + 6]), "f"(vals[v * 8 + 7]) : "memory");
        continue;
      }
    }
    // (bf16 rows are 16*C bytes per thread: mixing one 256-bit and one 128-bit store by item parity measured slower — 68.6 % vs 73.2 % at 224^2)
#pragma unroll
    for (int v = 0; v < NE / VE; ++v) __stcs(reinterpret_cast<uint4*>(dst + v * VE), vec_pack<T>(vals + v * VE));   // streaming store: written once, read by the next kernel from L2/HBM
  }
}

template <typename T, int C>
static int launch_preprocess_fast(PreprocessArgs a, cudaStream_t st) {
  // Table replicas: lane l reads replica l % R, i.e. bank (byte % (32/R)) * R + l % R. ncu on R = 8: 7.6 extra shared wavefronts per lookup and the
  // L1TEX pipe at 90 %; R = 16 leaves a 2-way conflict only between lanes l and l+16 that look up bytes of the same parity (48 KB for 3 channels).
  constexpr int R = 16;
  a.div_vec_per_row = FastDiv(a.out_h);
  const size_t total_items = (size_t)a.n * a.out_h * (a.out_w >> 3);
  DCV_REQUIRE((size_t)a.n * a.out_h < (1u << 31), "preprocess_u8: too many rows");
  // 512-thread CTAs: one 48 KB table serves twice the threads, so 3 CTAs = 1536 threads fit per SM instead of 4 x 256 = 1024 (each thread has only two
  // 16-byte loads in flight: the kernel was latency-bound at 60-75 % of HBM with 32 KB in flight per SM)
  constexpr int kThreads = 512;
  size_t blocks = (total_items + kThreads - 1) / kThreads;
  const size_t smem = (size_t)C * 256 * R * sizeof(float);
  auto kern = preprocess_stream_kernel<T, C, R>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem) != cudaSuccess || occ < 1) occ = 1;
  const size_t max_grid = (size_t)kNumSMs * occ;              // persistent CTAs, one resident wave: the table is filled once per CTA
  if (blocks > max_grid) blocks = max_grid;
  kern<<<(unsigned)blocks, kThreads, smem, st>>>(a, total_items);
  DCV_LAUNCH_CHECK("preprocess_stream_kernel");
  return 0;
}

// Scalar variant for shapes whose rows are not a whole number of 16-byte vectors.
template <typename T>
__global__ void preprocess_scalar_kernel(const PreprocessArgs a) {
  const size_t total = (size_t)a.n * a.out_h * a.out_w * a.c_out;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int ch, j, i, img;
    size_t t = idx;
    if (!a.nchw_out) { ch = t % a.c_out; t /= a.c_out; j = t % a.out_w; t /= a.out_w; i = t % a.out_h; img = t / a.out_h; }
    else { j = t % a.out_w; t /= a.out_w; i = t % a.out_h; t /= a.out_h; ch = t % a.c_out; img = t / a.c_out; }
    float val = 0.f;
    if (ch < a.c) {
      const int top = a.crop_yx ? a.crop_yx[2 * img] : a.pad, left = a.crop_yx ? a.crop_yx[2 * img + 1] : a.pad;
      const bool flip = a.flip ? (a.flip[img] != 0) : false;
      const int jj = flip ? (a.out_w - 1 - j) : j;
      const int r = top + i - a.pad, col = left + jj - a.pad;
      const int byte = (r >= 0 && r < a.h && col >= 0 && col < a.w) ? a.src[(((size_t)img * a.h + r) * a.w + col) * a.c + ch] : 0;
      val = ((float)byte / 255.0f - a.mean[ch]) / a.stdv[ch];
    }
    reinterpret_cast<T*>(a.dst)[idx] = from_f<T>(val);
  }
}

}  // namespace dcv

extern "C" int dcv_preprocess_u8(const uint8_t* src, void* dst, int n, int h, int w, int c, int out_h, int out_w, int pad,
                                 const float* mean, const float* std, const uint8_t* flip, const int32_t* crop_yx,
                                 int out_dtype, int c_out, int nchw_out, void* stream) {
  using namespace dcv;
  DCV_REQUIRE(src && dst && mean && std, "preprocess_u8: null pointer");
  DCV_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c <= 4 && out_h > 0 && out_w > 0 && pad >= 0, "preprocess_u8: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  DCV_REQUIRE(c_out >= c && (!nchw_out || c_out == c), "preprocess_u8: c_out=%d incompatible with c=%d (nchw_out=%d)", c_out, c, nchw_out);
  DCV_REQUIRE(out_h <= h + 2 * pad && out_w <= w + 2 * pad, "preprocess_u8: crop %dx%d larger than padded source %dx%d", out_h, out_w, h + 2 * pad, w + 2 * pad);
  DCV_REQUIRE((size_t)out_w * c_out < (1u << 30), "preprocess_u8: row too long");
  PreprocessArgs a;
  a.src = src; a.dst = dst; a.n = n; a.h = h; a.w = w; a.c = c; a.out_h = out_h; a.out_w = out_w; a.pad = pad; a.c_out = c_out; a.nchw_out = nchw_out;
  a.mean = mean; a.stdv = std; a.flip = flip; a.crop_yx = crop_yx;
  a.src_total_bytes = (size_t)n * h * w * c;
  a.wide_store = reinterpret_cast<uintptr_t>(dst) % 32 == 0 ? 1 : 0;
  const int esize = out_dtype == DCV_BF16 ? 2 : 4;
  const int ve = 16 / esize;
  const int row_elems = nchw_out ? out_w : out_w * c_out;
  const bool vec_ok = (row_elems % ve == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 4 == 0);
  cudaStream_t st = as_stream(stream);
  if (vec_ok && !nchw_out && c_out == c && out_w % 8 == 0 && out_w >= 8) {
    if (out_dtype == DCV_BF16) {
      if (c == 1) return launch_preprocess_fast<__nv_bfloat16, 1>(a, st);
      if (c == 2) return launch_preprocess_fast<__nv_bfloat16, 2>(a, st);
      if (c == 3) return launch_preprocess_fast<__nv_bfloat16, 3>(a, st);
      return launch_preprocess_fast<__nv_bfloat16, 4>(a, st);
    }
    if (out_dtype == DCV_F32) {
      if (c == 1) return launch_preprocess_fast<float, 1>(a, st);
      if (c == 2) return launch_preprocess_fast<float, 2>(a, st);
      if (c == 3) return launch_preprocess_fast<float, 3>(a, st);
      return launch_preprocess_fast<float, 4>(a, st);
    }
  }
  if (!vec_ok) {
    const size_t total = (size_t)n * out_h * out_w * c_out;
    DCV_DISPATCH_DTYPE(out_dtype, T, (preprocess_scalar_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(a)));
    DCV_LAUNCH_CHECK("preprocess_scalar_kernel");
    return 0;
  }
  a.vec_per_row = row_elems / ve;
  // rows per group: enough vectors to keep 256 threads busy for a few iterations, bounded by 16 KB of staged bytes
  a.row_bytes_smem = ((out_w * c + 3) / 4) * 4 + 16;
  int rg = (1024 + a.vec_per_row * (nchw_out ? c : 1) - 1) / (a.vec_per_row * (nchw_out ? c : 1));
  rg = rg < 1 ? 1 : rg;
  if (rg > out_h) rg = out_h;
  while (rg > 1 && (size_t)rg * a.row_bytes_smem > 16384) --rg;
  a.rg = rg;
  a.div_cout = FastDiv(c_out);
  a.div_vec_per_row = FastDiv(a.vec_per_row);
  const int groups_per_img = (out_h + rg - 1) / rg;
  a.div_groups_per_img = FastDiv(groups_per_img);
  const size_t total_groups = (size_t)n * groups_per_img;
  DCV_REQUIRE(total_groups < (1u << 31), "preprocess_u8: too many row groups");
  // Bank-replicated table once the batch is big enough to amortise filling it (>= ~8 groups per CTA).
  const bool replicate = total_groups >= (size_t)kNumSMs * 2 * 8 && (size_t)out_h * out_w >= 64 * 64;
  const int R = replicate ? 32 : 1;
  const size_t smem = (size_t)c * 256 * R * sizeof(float) + (size_t)rg * a.row_bytes_smem;
  const int grid = (int)(total_groups < (size_t)kNumSMs * (replicate ? 2 : 8) ? total_groups : (size_t)kNumSMs * (replicate ? 2 : 8));
  if (out_dtype == DCV_F32) {
    if (replicate) {
      cudaFuncSetAttribute(preprocess_kernel<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      preprocess_kernel<float, 32><<<grid, 256, smem, st>>>(a);
    } else preprocess_kernel<float, 1><<<grid, 256, smem, st>>>(a);
  } else if (out_dtype == DCV_BF16) {
    if (replicate) {
      cudaFuncSetAttribute(preprocess_kernel<__nv_bfloat16, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      preprocess_kernel<__nv_bfloat16, 32><<<grid, 256, smem, st>>>(a);
    } else preprocess_kernel<__nv_bfloat16, 1><<<grid, 256, smem, st>>>(a);
  } else {
    DCV_REQUIRE(false, "preprocess_u8: unsupported out_dtype %d", out_dtype);
  }
  DCV_LAUNCH_CHECK("preprocess_kernel");
  return 0;
}
