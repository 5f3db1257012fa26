// K1 and the other HBM-bound kernels of the path: uint8 crop + bilinear resize + ImageNet normalise,
// NCHW-fp32 -> stem-layout repack (Seam A), 3x3/2 max-pool.
//
// Stem input layout ("NHWC4p"): bf16 [N][224][232][4]; pixel column wp = w + 4 (4 zero pixels on each side so the
// 7-tap window of the stem conv never leaves the row), channel 3 is zero padding.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "ptx_sm100.cuh"

namespace phdfxk {

constexpr int kImg = 224;        // network input side
constexpr int kStemWPad = 232;   // padded row length in pixels
constexpr int kStemLeftPad = 4;

// Mirrors src/dataset.py:141-152 (_crop_and_resize_video_uint8) + :242-245/:429 (Normalize):
//   crop [top:top+h, left:left+w] -> F.resize(224, bilinear, antialias=False) evaluated in fp32 and ROUNDED
//   half-to-even back to uint8 (torchvision transforms/_functional_tensor.py:462-472,536-540) -> /255 ->
//   (x - mean) / std  (fp32) -> bf16.
// ATen's bilinear (align_corners=False), in the exact fp32 operation order of the kernel the reference's
// DataLoader workers run (1 thread -> channels-last kernel; pinned bit-for-bit by oracle/preprocess_ref.py):
//   src = max(0, fma(scale, dst+0.5, -0.5)), scale = in/out;  wij = lam_h_i * lam_w_j;
//   value = fma(w11,v11, fma(w10,v10, fma(w00,v00, w01*v01))).
// One WARP per (frame, output row): both source rows of the bilinear stencil are the same for the whole output row, so
// the warp stages them in its own slice of shared memory with coalesced 16-byte loads (no block-wide barrier) and
// every lane then produces 8 of the row's 232 output pixels from there.
// The /255 and (x - mean)/std steps depend only on (channel, uint8 value): each block builds the 3 x 256 table of
// final bf16 values once, with exactly the reference's fp32 operations, and pixels look their result up.
// Dynamic shared memory: 1536 (table) + kK1Warps * 2 * row_cap bytes, row_cap = round_up(3*W + 32, 16).
constexpr int kK1Warps = 8;
constexpr int kK1LutBytes = 3 * 256 * 2;

// ---- colour jitter (the reference's `cjitter` variant, src/dataset.py:188-198: torchvision v2 ColorJitter on the
// resized [0,1] clip, one parameter draw per clip, before Normalize).  Parameters arrive per FRAME as 12 floats:
//   [0..3] the four ops in application order (0 brightness, 1 contrast, 2 saturation, 3 hue; transforms/v2/_color.py
//   :156-173), [4] brightness factor, [5] contrast factor, [6] float32(1.0 - contrast) (formed in double, as
//   _blend's `alpha`, functional/_color.py:92-97), [7] saturation factor, [8] float32(1.0 - saturation), [9] hue.
// Every op is the fp32 arithmetic of torchvision transforms/v2/functional/_color.py (:31-48 grey, :112-123 brightness,
// :149-165 saturation, :188-205 contrast, :300-395 hue); oracle/preprocess_ref.py restates the same and is pinned
// against torchvision to 1e-6.  Contrast blends with the frame's mean grey level of the image AS IT IS when the op
// runs, so the kernel runs twice: KIND_JITTER_SUMS applies the ops that precede contrast and writes one fp32 sum of
// grey per output row; KIND_JITTER adds the 224 row sums (in double, fixed order) and finishes.
constexpr int kJitterFloats = 12;
enum K1Kind : int { KIND_PLAIN = 0, KIND_JITTER_SUMS = 1, KIND_JITTER = 2 };

__device__ __forceinline__ float jit_clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
// a / b given rb = RN(1 / b): Markstein's correction sequence (q0 = a*rb; e = a - b*q0 exactly, by FMA; q = q0 + e*rb).
// With a correctly rounded reciprocal this IS the correctly rounded quotient (__fdiv_rn) for every operand pair the
// colour ops produce (normal range, |a| <= |b| <= 1 or constant b), at 3 instructions instead of the ~20 of an IEEE
// division — the jitter variant of K1 is bound by its 10 divisions per pixel, not by memory.  Divisors whose
// significand is all ones (where the bound of the theorem is not met) can differ in the last fp32 bit; the bf16
// rounding that follows hides all but ~2^-15 of those.
__device__ __forceinline__ float div_by(float a, float b, float rb) {
  const float q0 = __fmul_rn(a, rb);
  const float e = __fmaf_rn(-b, q0, a);
  return __fmaf_rn(e, rb, q0);
}
__device__ __forceinline__ float jit_gray(const float (&c)[3]) {
  return __fmaf_rn(c[2], 0.114f, __fmaf_rn(c[1], 0.587f, __fmul_rn(c[0], 0.2989f)));
}
__device__ __forceinline__ void jit_blend(float (&c)[3], float other, float ratio, float rest) {
#pragma unroll
  for (int k = 0; k < 3; ++k) c[k] = jit_clamp01(__fmaf_rn(other, rest, __fmul_rn(c[k], ratio)));
}
__device__ __forceinline__ void jit_hue(float (&c)[3], float hue) {
  const float r = c[0], g = c[1], b = c[2];
  const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
  const bool eqc = maxc == minc;
  const float cr = __fsub_rn(maxc, minc);
  const float sdiv = eqc ? 1.0f : maxc;
  const float s = div_by(cr, sdiv, __frcp_rn(sdiv));
  const float div = eqc ? 1.0f : cr;
  const float rdiv = __frcp_rn(div);  // one reciprocal for the three channel ratios
  const float rc = div_by(__fsub_rn(maxc, r), div, rdiv), gc = div_by(__fsub_rn(maxc, g), div, rdiv),
              bc = div_by(__fsub_rn(maxc, b), div, rdiv);
  const bool neq_r = maxc != r, eq_g = maxc == g;
  const float hg = (eq_g && neq_r) ? __fsub_rn(__fadd_rn(rc, 2.0f), bc) : 0.0f;
  const float hr = (!neq_r) ? __fsub_rn(bc, gc) : 0.0f;
  const float hb = (neq_r && !eq_g) ? __fsub_rn(__fadd_rn(gc, 4.0f), rc) : 0.0f;
  float h = __fadd_rn(__fadd_rn(hr, hg), hb);
  h = __fadd_rn(__fmul_rn(h, 1.0f / 6.0f), 1.0f);  // in [5/6, 11/6]: fmod(x, 1) == x - floor(x), exactly
  h = __fsub_rn(h, floorf(h));
  h = __fadd_rn(h, hue);              // h.add_(hue_factor).remainder_(1.0): result takes the divisor's sign
  h = __fsub_rn(h, floorf(h));
  if (h >= 1.0f) h = 0.0f;
  const float v = maxc;
  const float h6 = __fmul_rn(h, 6.0f);
  const float fi = floorf(h6);
  const float f = __fsub_rn(h6, fi);
  int i = static_cast<int>(fi) % 6;
  if (i < 0) i += 6;
  const float sxf = __fmul_rn(s, f);
  const float oms = __fsub_rn(1.0f, s);
  const float q = jit_clamp01(__fmul_rn(__fsub_rn(1.0f, sxf), v));
  const float t = jit_clamp01(__fmul_rn(__fadd_rn(sxf, oms), v));
  const float pp = jit_clamp01(__fmul_rn(oms, v));
  // vpqt -> rgb by sextant (functional/_color.py:359)
  switch (i) {
    case 0: c[0] = v, c[1] = t, c[2] = pp; break;
    case 1: c[0] = q, c[1] = v, c[2] = pp; break;
    case 2: c[0] = pp, c[1] = v, c[2] = t; break;
    case 3: c[0] = pp, c[1] = q, c[2] = v; break;
    case 4: c[0] = t, c[1] = pp, c[2] = v; break;
    default: c[0] = v, c[1] = pp, c[2] = q; break;
  }
}
// the ops of one frame in order; until_contrast: stop in front of the contrast op (first pass)
__device__ __forceinline__ void jit_apply(float (&c)[3], const float* __restrict__ prm, float mean_gray,
                                          bool until_contrast) {
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int op = static_cast<int>(prm[k]);
    if (op == 0) {
#pragma unroll
      for (int j = 0; j < 3; ++j) c[j] = jit_clamp01(__fmul_rn(c[j], prm[4]));
    } else if (op == 1) {
      if (until_contrast) return;
      jit_blend(c, mean_gray, prm[5], prm[6]);
    } else if (op == 2) {
      jit_blend(c, jit_gray(c), prm[7], prm[8]);
    } else {
      jit_hue(c, prm[9]);
    }
  }
}

// ---- per-row pieces of K1, shared by preprocess_u8_kernel and by the stem kernel's converter warps (stem_pool_sm100.cuh,
// FUSE_K1: the same rows are produced straight into the stem's shared-memory ring, no NHWC4p tensor in HBM) ----------
struct K1Box {        // per frame
  int top, left, bh, bw;  // crop, clamped to the frame
  float scale_h, scale_w;
};
struct K1Row {        // per output row
  const uint8_t* r0;  // first byte of source row y0 of the crop (bytes [left*3, (left+bw)*3) of the frame row)
  const uint8_t* r1;  // same for row y1
  int bh, bw;         // crop size after clamping to the frame
  float ly0, ly1;     // vertical weights
  bool need1;         // row y1 carries weight (weight-0 taps add exactly +0: identity-size crops skip it)
};
// crop box of frame n, clamped to the frame as the reference's slicing `frames[:, top:top+h, left:left+w]`
// (src/dataset.py:146) cuts it; row staging is sized from W and relies on left + bw <= W
__device__ __forceinline__ K1Box k1_box(int n, int H, int W, const int32_t* __restrict__ boxes) {
  K1Box b;
  b.top = 0, b.left = 0, b.bh = H, b.bw = W;
  if (boxes != nullptr) {
    b.top = boxes[4 * n + 0];
    b.left = boxes[4 * n + 1];
    b.bh = boxes[4 * n + 2];
    b.bw = boxes[4 * n + 3];
    b.top = min(max(b.top, 0), H - 1);
    b.left = min(max(b.left, 0), W - 1);
    b.bh = min(max(b.bh, 1), H - b.top);
    b.bw = min(max(b.bw, 1), W - b.left);
  }
  b.scale_h = __fdiv_rn(static_cast<float>(b.bh), static_cast<float>(kImg));
  b.scale_w = __fdiv_rn(static_cast<float>(b.bw), static_cast<float>(kImg));
  return b;
}
// the vertical stencil of output row y of frame n
__device__ __forceinline__ K1Row k1_row(const uint8_t* __restrict__ frames, const K1Box& b, int n, int y, int H, int W) {
  K1Row g;
  g.bh = b.bh;
  g.bw = b.bw;
  float sy = __fmaf_rn(b.scale_h, static_cast<float>(y) + 0.5f, -0.5f);
  sy = sy < 0.0f ? 0.0f : sy;
  int y0 = static_cast<int>(sy);
  y0 = y0 > g.bh - 1 ? g.bh - 1 : y0;
  const int y1 = y0 + (y0 < g.bh - 1 ? 1 : 0);
  g.ly1 = __fsub_rn(sy, static_cast<float>(y0));
  g.ly1 = fminf(fmaxf(g.ly1, 0.0f), 1.0f);
  g.ly0 = __fsub_rn(1.0f, g.ly1);
  const uint8_t* base = frames + static_cast<size_t>(n) * H * W * 3;
  g.r0 = base + (static_cast<size_t>(b.top + y0) * W + b.left) * 3;
  g.r1 = base + (static_cast<size_t>(b.top + y1) * W + b.left) * 3;
  g.need1 = g.ly1 != 0.0f;
  return g;
}
// the [3][256] table of final bf16 values: /255 and Normalize in the reference's fp32 operations
__device__ __forceinline__ void k1_build_lut(__nv_bfloat16* lut, int tid, int nthreads) {
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int i = tid; i < 3 * 256; i += nthreads) {
    const int c = i >> 8;
    const float x01 = __fdiv_rn(static_cast<float>(i & 255), 255.0f);             // frames.to(float32) / 255.0
    lut[i] = __float2bfloat16_rn(__fdiv_rn(__fsub_rn(x01, mean[c]), stdv[c]));    // Normalize, then bf16
  }
}
// One NHWC4p pixel (4 x bf16: three channels + zero) of output column x in [0, 224) from the staged source rows.
// IDENT: a 224x224 crop — F.resize returns its input untouched (torchvision transforms/functional.py:470-471); the
// bilinear stencil degenerates to weight 1 on v00 and +0 elsewhere, i.e. the same bytes: table lookups only.
// Both variants are branch-free so that the seven pixels a lane produces per row overlap their shared-memory loads.
template <bool IDENT>
__device__ __forceinline__ uint2 k1_plain_pixel(const uint8_t* t0, const uint8_t* t1, const K1Row& g, float scale_w,
                                                int flip_w, const __nv_bfloat16* lut, int x) {
  const uint16_t* l16 = reinterpret_cast<const uint16_t*>(lut);
  uint2 o;
  x = flip_w ? kImg - 1 - x : x;
  if (IDENT) {
    const uint32_t r0v = l16[t0[x * 3 + 0]], r1v = l16[256 + t0[x * 3 + 1]], r2v = l16[512 + t0[x * 3 + 2]];
    o.x = r0v | (r1v << 16);
    o.y = r2v;
    return o;
  }
  float sx = __fmaf_rn(scale_w, static_cast<float>(x) + 0.5f, -0.5f);
  sx = sx < 0.0f ? 0.0f : sx;
  int x0 = static_cast<int>(sx);
  x0 = x0 > g.bw - 1 ? g.bw - 1 : x0;
  const int x1 = x0 + (x0 < g.bw - 1 ? 1 : 0);
  float lx1 = __fsub_rn(sx, static_cast<float>(x0));
  lx1 = fminf(fmaxf(lx1, 0.0f), 1.0f);
  const float lx0 = __fsub_rn(1.0f, lx1);
  const float w00 = __fmul_rn(g.ly0, lx0), w01 = __fmul_rn(g.ly0, lx1);
  const float w10 = __fmul_rn(g.ly1, lx0), w11 = __fmul_rn(g.ly1, lx1);
  uint32_t r[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v00 = static_cast<float>(t0[x0 * 3 + c]);
    const float v01 = static_cast<float>(t0[x1 * 3 + c]);
    const float v10 = static_cast<float>(t1[x0 * 3 + c]);
    const float v11 = static_cast<float>(t1[x1 * 3 + c]);
    const float v = __fmaf_rn(w11, v11, __fmaf_rn(w10, v10, __fmaf_rn(w00, v00, __fmul_rn(w01, v01))));
    const int u8 = static_cast<int>(fminf(fmaxf(rintf(v), 0.0f), 255.0f));  // round half to even, uint8 range
    r[c] = l16[c * 256 + u8];
  }
  o.x = r[0] | (r[1] << 16);
  o.y = r[2];
  return o;
}
// the 224 image pixels of one output row, seven per lane (x = lane + 32 i); dst points at pixel x = 0
template <bool IDENT>
__device__ __forceinline__ void k1_plain_row(const uint8_t* t0, const uint8_t* t1, const K1Row& g, float scale_w,
                                             int flip_w, const __nv_bfloat16* lut, uint2* dst, int lane) {
  uint2 o[kImg / 32];
  if (IDENT) {
    // all 21 source bytes first, then the 21 table lookups: two shared-memory round trips per row instead of fourteen
    const uint16_t* l16 = reinterpret_cast<const uint16_t*>(lut);
    const uint8_t* tp = t0 + (flip_w ? (kImg - 1 - lane) * 3 : lane * 3);
    const int step = flip_w ? -96 : 96;
    uint32_t v[kImg / 32][3];
#pragma unroll
    for (int i = 0; i < kImg / 32; ++i) {
#pragma unroll
      for (int c = 0; c < 3; ++c) v[i][c] = tp[i * step + c];
    }
#pragma unroll
    for (int i = 0; i < kImg / 32; ++i) {
      const uint32_t r0v = l16[v[i][0]], r1v = l16[256 + v[i][1]], r2v = l16[512 + v[i][2]];
      o[i].x = r0v | (r1v << 16);
      o[i].y = r2v;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kImg / 32; ++i) o[i] = k1_plain_pixel<false>(t0, t1, g, scale_w, flip_w, lut, lane + 32 * i);
  }
#pragma unroll
  for (int i = 0; i < kImg / 32; ++i) dst[lane + 32 * i] = o[i];
}

template <int KIND>
__global__ void __launch_bounds__(kK1Warps * 32)
preprocess_u8_kernel(const uint8_t* __restrict__ frames, int n_frames, int H, int W,
                     const int32_t* __restrict__ boxes, int flip_w, __nv_bfloat16* __restrict__ out,
                     const float* __restrict__ jitter, float* __restrict__ row_sums) {
  extern __shared__ __align__(16) uint8_t s_rows[];
  griddep_launch_dependents();
  griddep_wait();  // the arena input buffer may still be read by the previous step's stem kernel
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const float rstd[3] = {__frcp_rn(0.229f), __frcp_rn(0.224f), __frcp_rn(0.225f)};  // jitter kinds: div_by()
  const int row_cap = (3 * W + 32 + 15) & ~15;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  __nv_bfloat16* lut = reinterpret_cast<__nv_bfloat16*>(s_rows);  // [3][256] (KIND_PLAIN)
  float* lut01 = reinterpret_cast<float*>(s_rows);                // [256] u8 / 255 (jitter kinds; same 1536 bytes)
  if (KIND == KIND_PLAIN) {
    k1_build_lut(lut, threadIdx.x, blockDim.x);
  } else {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut01[i] = __fdiv_rn(static_cast<float>(i), 255.0f);
  }
  __syncthreads();
  uint8_t* s0 = s_rows + kK1LutBytes + static_cast<size_t>(warp) * 2 * row_cap;
  uint8_t* s1 = s0 + row_cap;
  const uint8_t* buf_begin = frames;
  const uint8_t* buf_end = frames + static_cast<size_t>(n_frames) * H * W * 3;
  const int total_rows = n_frames * kImg;
  for (int rowi = blockIdx.x * kK1Warps + warp; rowi < total_rows; rowi += gridDim.x * kK1Warps) {
    const int n = rowi / kImg;
    const int y = rowi - n * kImg;
    const K1Box box = k1_box(n, H, W, boxes);
    const K1Row geo = k1_row(frames, box, n, y, H, W);
    const int bh = geo.bh, bw = geo.bw;
    const float ly0 = geo.ly0, ly1 = geo.ly1;
    const uint8_t* r0 = geo.r0;
    const uint8_t* r1 = geo.r1;
    const int off0 = static_cast<int>(reinterpret_cast<uintptr_t>(r0) & 15);
    const int off1 = static_cast<int>(reinterpret_cast<uintptr_t>(r1) & 15);
    const bool need1 = geo.need1;
    __syncwarp();                    // the previous row's readers are done with s0 / s1
    for (int pass = 0; pass < (need1 ? 2 : 1); ++pass) {
      const uint8_t* src = (pass == 0 ? r0 - off0 : r1 - off1);  // 16-byte aligned
      uint8_t* dst = pass == 0 ? s0 : s1;
      const int nvec = ((pass == 0 ? off0 : off1) + bw * 3 + 15) >> 4;
      for (int i = lane; i < nvec; i += 32) {
        const uint8_t* g = src + 16 * i;
        if (g >= buf_begin && g + 16 <= buf_end) {
          reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(g));
        } else {  // first / last vector of the whole buffer: stay inside it
          for (int k = 0; k < 16; ++k) dst[16 * i + k] = (g + k >= buf_begin && g + k < buf_end) ? g[k] : 0;
        }
      }
    }
    __syncwarp();
    const uint8_t* t0 = s0 + off0;
    const uint8_t* t1 = need1 ? s1 + off1 : t0;
    const float scale_w = box.scale_w;
    const bool ident = (bh == kImg) && (bw == kImg) && KIND == KIND_PLAIN;
    if (KIND == KIND_PLAIN) {
      uint2* out_row = reinterpret_cast<uint2*>(out) + static_cast<size_t>(rowi) * kStemWPad;
      if (ident)
        k1_plain_row<true>(t0, t1, geo, scale_w, flip_w, lut, out_row + kStemLeftPad, lane);
      else
        k1_plain_row<false>(t0, t1, geo, scale_w, flip_w, lut, out_row + kStemLeftPad, lane);
      if (lane < 2 * kStemLeftPad)  // the zero pad pixels left and right of the image
        out_row[lane < kStemLeftPad ? lane : kImg + lane] = make_uint2(0u, 0u);
      continue;
    }
    const float* prm = KIND != KIND_PLAIN ? jitter + static_cast<size_t>(n) * kJitterFloats : nullptr;
    float mean_gray = 0.0f;
    if (KIND == KIND_JITTER) {  // sum of the frame's 224 row sums, in double, fixed order
      double acc = 0.0;
      for (int r = lane; r < kImg; r += 32) acc += static_cast<double>(row_sums[static_cast<size_t>(n) * kImg + r]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
      mean_gray = static_cast<float>(acc / static_cast<double>(kImg * kImg));
    }
    float gray_sum = 0.0f;
    for (int wp = lane; wp < kStemWPad; wp += 32) {
      uint2 o = make_uint2(0u, 0u);
      int x = wp - kStemLeftPad;
      if (x >= 0 && x < kImg) {
        if (flip_w) x = kImg - 1 - x;
        float sx = __fmaf_rn(scale_w, static_cast<float>(x) + 0.5f, -0.5f);
        sx = sx < 0.0f ? 0.0f : sx;
        int x0 = static_cast<int>(sx);
        x0 = x0 > bw - 1 ? bw - 1 : x0;
        const int x1 = x0 + (x0 < bw - 1 ? 1 : 0);
        float lx1 = __fsub_rn(sx, static_cast<float>(x0));
        lx1 = fminf(fmaxf(lx1, 0.0f), 1.0f);
        const float lx0 = __fsub_rn(1.0f, lx1);
        const float w00 = __fmul_rn(ly0, lx0), w01 = __fmul_rn(ly0, lx1);
        const float w10 = __fmul_rn(ly1, lx0), w11 = __fmul_rn(ly1, lx1);
        uint32_t r[3] = {0u, 0u, 0u};
        float px[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v00 = static_cast<float>(t0[x0 * 3 + c]);
          const float v01 = static_cast<float>(t0[x1 * 3 + c]);
          const float v10 = static_cast<float>(t1[x0 * 3 + c]);
          const float v11 = static_cast<float>(t1[x1 * 3 + c]);
          const float v = __fmaf_rn(w11, v11, __fmaf_rn(w10, v10, __fmaf_rn(w00, v00, __fmul_rn(w01, v01))));
          const int u8 = static_cast<int>(fminf(fmaxf(rintf(v), 0.0f), 255.0f));  // round half to even, uint8 range
          if (KIND == KIND_PLAIN)
            r[c] = reinterpret_cast<const uint16_t*>(lut)[c * 256 + u8];
          else
            px[c] = lut01[u8];  // frames.to(float32) / 255.0
        }
        if (KIND == KIND_JITTER_SUMS) {
          jit_apply(px, prm, 0.0f, true);
          gray_sum += jit_gray(px);
        } else if (KIND == KIND_JITTER) {
          jit_apply(px, prm, mean_gray, false);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const __nv_bfloat16 hv = __float2bfloat16_rn(div_by(__fsub_rn(px[c], mean[c]), stdv[c], rstd[c]));
            r[c] = *reinterpret_cast<const uint16_t*>(&hv);
          }
        }
        o.x = r[0] | (r[1] << 16);
        o.y = r[2];
      }
      if (KIND != KIND_JITTER_SUMS) reinterpret_cast<uint2*>(out)[static_cast<size_t>(rowi) * kStemWPad + wp] = o;
    }
    if (KIND == KIND_JITTER_SUMS) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) gray_sum += __shfl_xor_sync(0xffffffffu, gray_sum, d);
      if (lane == 0) row_sums[rowi] = gray_sum;
    }
  }
}

// Seam A repack: already-normalised fp32 NCHW [N,3,224,224] (src/preprocess_resnet_features.py:295) -> NHWC4p.
__global__ void nchw_f32_to_stem_kernel(const float* __restrict__ x, int n_frames, __nv_bfloat16* __restrict__ out) {
  griddep_launch_dependents();
  griddep_wait();
  const int total = n_frames * kImg * kStemWPad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int wp = idx % kStemWPad;
    const int y = (idx / kStemWPad) % kImg;
    const int n = idx / (kStemWPad * kImg);
    uint2 o = make_uint2(0u, 0u);
    const int xx = wp - kStemLeftPad;
    if (xx >= 0 && xx < kImg) {
      const size_t plane = static_cast<size_t>(kImg) * kImg;
      const float* px = x + static_cast<size_t>(n) * 3 * plane + static_cast<size_t>(y) * kImg + xx;
      const __nv_bfloat162 a = __floats2bfloat162_rn(__ldg(px), __ldg(px + plane));
      const __nv_bfloat162 b = __floats2bfloat162_rn(__ldg(px + 2 * plane), 0.0f);
      o.x = *reinterpret_cast<const uint32_t*>(&a);
      o.y = *reinterpret_cast<const uint32_t*>(&b);
    }
    reinterpret_cast<uint2*>(out)[idx] = o;
  }
}

// MaxPool2d(3, stride 2, pad 1) over NHWC bf16 (torchvision models/resnet.py:200); 8 channels per thread.
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in, int n_frames, int Hin, int Win, int C,
                                    __nv_bfloat16* __restrict__ out) {
  griddep_launch_dependents();
  griddep_wait();
  const int Ho = (Hin + 2 - 3) / 2 + 1;
  const int Wo = (Win + 2 - 3) / 2 + 1;
  const int cg = C / 8;
  const long long total = static_cast<long long>(n_frames) * Ho * Wo * cg;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % cg);
    const int qo = static_cast<int>((idx / cg) % Wo);
    const int po = static_cast<int>((idx / (static_cast<long long>(cg) * Wo)) % Ho);
    const int n = static_cast<int>(idx / (static_cast<long long>(cg) * Wo * Ho));
    __nv_bfloat162 m[4];
    const __nv_bfloat162 ninf = __floats2bfloat162_rn(-INFINITY, -INFINITY);
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = ninf;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int y = 2 * po - 1 + dy;
      if (y < 0 || y >= Hin) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int x = 2 * qo - 1 + dx;
        if (x < 0 || x >= Win) continue;
        const uint4 v =
            __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(n) * Hin + y) * Win + x) * C) + g);
        const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], pv[j]);
      }
    }
    uint4 o;
    o.x = *reinterpret_cast<uint32_t*>(&m[0]);
    o.y = *reinterpret_cast<uint32_t*>(&m[1]);
    o.z = *reinterpret_cast<uint32_t*>(&m[2]);
    o.w = *reinterpret_cast<uint32_t*>(&m[3]);
    reinterpret_cast<uint4*>(out)[idx] = o;
  }
}

}  // namespace phdfxk
