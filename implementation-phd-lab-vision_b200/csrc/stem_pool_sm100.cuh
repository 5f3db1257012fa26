// K2: stem conv 7x7/2 (3->64) + folded BN + ReLU + MaxPool 3x3/2, one kernel (torchvision models/resnet.py:197-200,
// :268-271: conv1, bn1, relu, maxpool).  Input NHWC4p bf16 [N][224][232][4], output NHWC bf16 [N][56][56][64].
//
// The conv output (1.6 MB/frame) never leaves the SM.  Each CTA walks a band of conv rows of one frame, two conv rows
// (= one pooled row) per step.
//
// Implicit GEMM without any im2col copy: one conv row = one 128x64 MMA tile (112 valid pixels).  For filter row r the
// A operand is the RAW input row 2p+r-3 as it lies in shared memory: output column q reads the 8-pixel x 4-channel
// window starting 16*q bytes into the row, i.e. a K-major no-swizzle operand whose rows OVERLAP (row stride 16 B,
// K-chunk stride LBO = 16 B, 8-row-group stride SBO = 128 B) — the tensor core does the sliding window.  Rows outside
// the image read a zeroed slot.  K = 7 filter rows x 32 (8 px x 4 ch; tap -1 and channel 3 carry zero weights).
//   warp 0    producer: cp.async.bulk of input row PAIRS into a 16-slot ring (2 x 1856 B per slot)
//   warp 1    MMA issuer: a step's even/odd conv rows go to adjacent 64-column halves of one of two TMEM buffers; an input
//             row that feeds both (filter row e of the even, e-2 of the odd conv row) is ONE N=128 tcgen05.mma pair against
//             the stacked filter rows: 8 N=64 + 10 N=128 MMAs (M=128, K=16) per step; weights resident in smem (68 KB)
//   warp 2    DMA: TMA store of pooled rows (56 px x 64 ch = 7 KB each); owns the TMEM allocation
//   warp 4-11 epilogue, two warps per TMEM lane quadrant (32 channels each).  Thread q owns conv column q:
//             vertical 3-max of rows 2i-1, 2i, 2i+1 in REGISTERS (row 2i-1 is carried from the previous step), then the
//             horizontal 3-max (columns 2j-1..2j+1) by warp shuffles — the neighbours of a centre column are one lane
//             away; only lane 0's left neighbour crosses a warp (64 B exchange slot) — and the pooled row goes straight
//             to the store staging.  The conv rows never touch shared memory, which the N=64 MMAs' operand reads
//             keep busy.  No ReLU instruction anywhere: the max starts at 0 and max(0, max(v)) == max-pool(relu(v)).
//
// FUSE_K1 (template): K1 — uint8 crop + bilinear resize + normalise (elementwise_sm100.cuh) — runs inside this kernel.
// Four more warps ("converters") read the uint8 source rows (cp.async, double-buffered per warp), produce the bf16
// NHWC4p rows with K1's own per-pixel function and write them straight into the ring slots the MMA warp reads; the
// 106 MB NHWC4p tensor (write + read per 256 frames) and K1's launch disappear.  The zero pad pixels of a ring row are
// written once, at kernel start.  Registers: 512 threads leave 128 per thread, the epilogue warps need ~170, so the
// warpgroups re-balance with setmaxnreg (producer / MMA / store warps 56, converters 104, epilogue 176).
#pragma once
#include "elementwise_sm100.cuh"
#include "ptx_sm100.cuh"

namespace phdfxk {

constexpr int kSpIn = 224, kSpInPitchPx = 232, kSpOut = 112, kSpPool = 56;
constexpr int kSpRowBytes = kSpInPitchPx * 8;   // 1856
constexpr int kSpRowPitch = 2048;               // smem pitch of one input row
constexpr int kSpPairSlots = 16;                // ring of row pairs (2 steps x 2 rows in flight need 7; the rest is prefetch)
constexpr int kSpBand = 28;                     // conv rows per work item: 4 bands per frame (1 halo row per 28, and 1024 bands = 6.9 waves of 148 at batch 256)
constexpr int kSpBandsPerFrame = kSpOut / kSpBand;  // 4
constexpr int kSpConvRowBytes = kSpOut * 128;   // 14336: one conv row, 112 px x 64 ch bf16
constexpr int kSpPoolRowBytes = 8192;           // staging pitch (56 x 128 B = 7168 used)
// weights: [7 filter rows][4 k-chunks][64 cout][8] bf16, then for e = 2..6 the STACKED pairs
// [4 k-chunks][128 rows = filter row e (even conv row) | filter row e-2 (odd conv row)][8] (see the MMA warp)
constexpr int kSpWeightBase = 7 * 4096;
constexpr int kSpWeightBytes = kSpWeightBase + 5 * 8192;
constexpr int kSpThreads = 384;
constexpr int kSpFuseThreads = 512;   // + one warpgroup of converters
constexpr int kSpConvWarps = 4;
constexpr int kSpEpiThreads = 256;
constexpr int kSpEpiWarps = 8;

struct StemPoolSmem {
  static constexpr int RING = 0;                                           // 8 x 4096
  static constexpr int ZERO = RING + kSpPairSlots * 2 * kSpRowPitch;       // 2048 + 256 zero bytes
  static constexpr int WEIGHTS = ZERO + kSpRowPitch + 1024;                // 1024-aligned
  static constexpr int POOL = WEIGHTS + kSpWeightBytes;                    // 2 x 8192, 1024-aligned
  static constexpr int BIAS = POOL + 2 * kSpPoolRowBytes;                  // 64 floats
  static constexpr int BARS = BIAS + 256;
  static constexpr int XCHG = BARS + 512;                                  // [2 parities][2 halves][4 quadrants][64 B]
  static constexpr int TOTAL = XCHG + 1024;
  // FUSE_K1 only: the [3][256] bf16 table, then per converter warp 2 buffers x 2 source rows of row_cap bytes
  static constexpr int LUT = TOTAL;
  static constexpr int STAGE = LUT + kK1LutBytes;
  static constexpr int fuse_total(int W) { return STAGE + kSpConvWarps * 4 * ((3 * W + 32 + 15) & ~15); }
};
static_assert(StemPoolSmem::WEIGHTS % 1024 == 0 && StemPoolSmem::POOL % 1024 == 0,
              "swizzled regions must be 1024-byte aligned");

struct StemPoolParams {
  const __nv_bfloat16* in;       // NHWC4p
  const __nv_bfloat16* weights;  // pre-laid-out smem image, kSpWeightBytes
  const float* bias;             // [64]
  int n_frames;
  uint32_t* zero_ptr;            // frame progress counters of the pass this launch opens (conv_igemm_sm100.cuh), or
  int zero_words;                // nullptr: zeroed here, behind this launch's full grid dependency
  // FUSE_K1: the call's uint8 frames [n][H][W][3], crop boxes (nullable) and the mirror flag, as preprocess_u8_kernel
  const uint8_t* frames;
  int H, W;
  const int32_t* boxes;
  int flip_w;
  long long* trace;  // debug (PHDFX_STEM_TRACE): converter warp 12 of CTA 0 accumulates clock64 per phase, [8]
};

__device__ __forceinline__ void cp_async_16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// executed by a whole converged warp; one elected lane issues (keeps operands in uniform registers)
__device__ __forceinline__ void bulk_load_elect(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// no-swizzle K-major descriptor: LBO = K-chunk (8 elements) stride, SBO = 8-row-group stride
__device__ __forceinline__ uint64_t make_nosw_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (static_cast<uint64_t>((addr & 0x3FFFFu) >> 4)) | (static_cast<uint64_t>(lbo >> 4) << 16) |
         (static_cast<uint64_t>(sbo >> 4) << 32) | (1ull << 46);
}

// band b of this launch -> (frame, first conv row computed, first conv row whose pooled output is emitted)
struct Band {
  int n, p0, p_first, p_last, j_first, j_last;
  __device__ __forceinline__ explicit Band(int b) {
    n = b / kSpBandsPerFrame;
    p0 = (b - n * kSpBandsPerFrame) * kSpBand;
    p_first = p0 > 0 ? p0 - 1 : 0;  // one halo conv row: pooled row p0/2 needs conv row p0-1
    p_last = p0 + kSpBand - 1;
    j_first = p_first - 2 < 0 ? 0 : p_first - 2;                       // input row pairs p-2 .. p+1 feed conv row p
    j_last = p_last + 1 > kSpOut - 1 ? kSpOut - 1 : p_last + 1;
  }
  __device__ __forceinline__ int pairs() const { return j_last - j_first + 1; }
};

template <bool FUSE_K1>
__global__ void __launch_bounds__(FUSE_K1 ? kSpFuseThreads : kSpThreads, 1)
stem_pool_kernel(const __grid_constant__ CUtensorMap mapO, const StemPoolParams p) {
  constexpr int NTHREADS = FUSE_K1 ? kSpFuseThreads : kSpThreads;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using L = StemPoolSmem;
  uint64_t* pair_full = reinterpret_cast<uint64_t*>(smem + L::BARS);  // [kSpPairSlots]
  uint64_t* pair_empty = pair_full + kSpPairSlots;                   // [8]
  uint64_t* tmem_full = pair_empty + kSpPairSlots;                   // [2]
  uint64_t* tmem_empty = tmem_full + 2;                              // [2]
  uint64_t* pool_full = tmem_empty + 2;                              // [2]
  uint64_t* pool_free = pool_full + 2;                               // [2]
  uint64_t* w_full = pool_free + 2;                                  // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_full + 1);
  float* s_bias = reinterpret_cast<float*>(smem + L::BIAS);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int num_bands = p.n_frames * kSpBandsPerFrame;

  // zero slot, bias
  for (int i = threadIdx.x; i < (kSpRowPitch + 1024) / 16; i += NTHREADS)
    reinterpret_cast<uint4*>(smem + L::ZERO)[i] = make_uint4(0, 0, 0, 0);
  if (FUSE_K1) {
    // ring rows: the converters only ever write the 224 image pixels of a row; its pad pixels stay zero from here on
    for (int i = threadIdx.x; i < kSpPairSlots * 2 * kSpRowPitch / 16; i += NTHREADS)
      reinterpret_cast<uint4*>(smem + L::RING)[i] = make_uint4(0, 0, 0, 0);
    k1_build_lut(reinterpret_cast<__nv_bfloat16*>(smem + L::LUT), threadIdx.x, NTHREADS);
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = __ldg(&p.bias[threadIdx.x]);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kSpPairSlots; ++i) {
      mbar_init(&pair_full[i], FUSE_K1 ? 2 : 1);  // FUSE_K1: one arrive per converted row of the pair
      mbar_init(&pair_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kSpEpiWarps);
      mbar_init(&pool_full[i], kSpEpiWarps);
      mbar_init(&pool_free[i], 1);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  if (warp == 2 && lane == 0) tma_prefetch_desc(&mapO);
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 256);
    tmem_relinquish();
  }
  fence_proxy_async_smem();  // zero slot is read by the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();  // PDL: see conv_igemm_sm100.cuh
  griddep_wait();
  if (p.zero_ptr != nullptr)
    for (int i = blockIdx.x * NTHREADS + threadIdx.x; i < p.zero_words; i += gridDim.x * NTHREADS) p.zero_ptr[i] = 0u;

  // Role dispatch by warpgroup, so that each setmaxnreg is ONE instruction executed by the four warps of its warpgroup
  // and dominates the code whose register budget it sets.
  if (warp < 4) {
  if (FUSE_K1) setmaxnreg_dec<56>();
  if (warp == 0) {
    // ------------------------------------------------------------------ producer (whole warp, uniform flow)
    {
      mbar_arrive_expect_tx_elect(w_full, kSpWeightBytes);
      bulk_load_elect(smem + L::WEIGHTS, p.weights, kSpWeightBytes, w_full);
      int seq = 0;
      for (int b = blockIdx.x; !FUSE_K1 && b < num_bands; b += gridDim.x) {
        const Band band(b);
        const uint8_t* frame = reinterpret_cast<const uint8_t*>(p.in) +
                               static_cast<size_t>(band.n) * kSpIn * kSpRowBytes;
        for (int j = band.j_first; j <= band.j_last; ++j, ++seq) {
          const int slot = seq % kSpPairSlots;
          mbar_wait(&pair_empty[slot], ((seq / kSpPairSlots) & 1) ^ 1);
          uint8_t* dst = smem + L::RING + slot * 2 * kSpRowPitch;
          mbar_arrive_expect_tx_elect(&pair_full[slot], 2 * kSpRowBytes);
          bulk_load_elect(dst, frame + static_cast<size_t>(2 * j) * kSpRowBytes, kSpRowBytes, &pair_full[slot]);
          bulk_load_elect(dst + kSpRowPitch, frame + static_cast<size_t>(2 * j + 1) * kSpRowBytes, kSpRowBytes,
                          &pair_full[slot]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, uniform flow)
    {
      constexpr uint32_t idesc = make_idesc_bf16(128, 64);
      constexpr uint32_t idesc128 = make_idesc_bf16(128, 128);
      const uint32_t ring_addr = smem_u32(smem + L::RING);
      const uint32_t zero_addr = smem_u32(smem + L::ZERO);
      const uint32_t w_addr = smem_u32(smem + L::WEIGHTS);
      mbar_wait(w_full, 0);
      int buf = 0;  // TMEM buffer of the current step: columns [buf*128, +64) even row, [buf*128+64, +64) odd row
      uint32_t buf_phase = 0;
      int seq_base = 0;
      for (int b = blockIdx.x; b < num_bands; b += gridDim.x) {
        const Band band(b);
        int waited = band.j_first;  // first pair of this band not yet known to have landed
        // input row h of this band in the ring (or the zero slot outside the image)
        auto a_of = [&](int h) -> uint32_t {
          if (h < 0 || h >= kSpIn) return zero_addr;
          const int seq = seq_base + ((h >> 1) - band.j_first);
          return ring_addr + (seq % kSpPairSlots) * 2 * kSpRowPitch + (h & 1) * kSpRowPitch;
        };
        // two K = 16 MMAs of one input row: N = 64 against filter row `fr` into d, or N = 128 against the stacked pair
        // `pair` (filter rows pair+2 | pair) into the whole 128-column step accumulator
        auto mma_row = [&](uint32_t d, uint32_t a_row, int fr, int pair, bool first) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t adesc = make_nosw_desc(a_row + k * 32, 16, 128);
            if (pair < 0)
              umma_bf16_elect(d, adesc, make_nosw_desc(w_addr + fr * 4096 + k * 2048, 1024, 128), idesc,
                              (first && k == 0) ? 0u : 1u);
            else
              umma_bf16_elect(d, adesc, make_nosw_desc(w_addr + kSpWeightBase + pair * 8192 + k * 4096, 2048, 128),
                              idesc128, 1u);
          }
        };
        for (int pr = band.p_first; pr <= band.p_last;) {
          // a step = (even row, odd row); the band's halo row (odd) is a step of its own
          const bool both = !(pr & 1);
          const int last = both ? pr + 1 : pr;
          mbar_wait(&tmem_empty[buf], buf_phase ^ 1);
          // input row pairs pr-2 .. last+1 feed this step; earlier ones were already waited for
          const int need = last + 1 > band.j_last ? band.j_last : last + 1;
          for (; waited <= need; ++waited) {
            const int seq = seq_base + (waited - band.j_first);
            mbar_wait(&pair_full[seq % kSpPairSlots], (seq / kSpPairSlots) & 1);
          }
          tc_fence_after();
          const uint32_t d_even = tmem_base + buf * 128, d_odd = d_even + 64;
          if (!both) {
            for (int r = 0; r < 7; ++r) mma_row(d_odd, a_of(2 * pr + r - 3), r, -1, r == 0);
          } else {
            // Input row h = 2*pr - 3 + e (e = 0..8) is filter row e of the even conv row AND filter row e-2 of the odd
            // one.  Rows that feed only one of them go first (N = 64, initialising that half of the accumulator); the
            // five rows that feed both are ONE N = 128 MMA pair each against the stacked filter rows [e | e-2] — 18
            // instead of 28 MMAs per step, and the slow part of an MMA here is fetching A (overlapping 16-byte rows).
            const int h0 = 2 * pr - 3;
            mma_row(d_odd, a_of(h0 + 7), 5, -1, true);
            mma_row(d_odd, a_of(h0 + 8), 6, -1, false);
            mma_row(d_even, a_of(h0 + 0), 0, -1, true);
            mma_row(d_even, a_of(h0 + 1), 1, -1, false);
            for (int e = 2; e < 7; ++e) mma_row(d_even, a_of(h0 + e), 0, e - 2, false);
          }
          umma_commit_elect(&tmem_full[buf]);
          buf ^= 1;
          if (buf == 0) buf_phase ^= 1;
          // pairs below last-1 are not needed by later rows
          for (int j = pr - 2; j <= last - 2; ++j) {
            if (j >= band.j_first) {
              const int seq = seq_base + (j - band.j_first);
              umma_commit_elect(&pair_empty[seq % kSpPairSlots]);
            }
          }
          if (last == band.p_last) {
            for (int j = (last - 1 < band.j_first ? band.j_first : last - 1); j <= band.j_last; ++j) {
              const int seq = seq_base + (j - band.j_first);
              umma_commit_elect(&pair_empty[seq % kSpPairSlots]);
            }
          }
          pr = last + 1;
        }
        seq_base += band.pairs();
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ DMA: pooled rows out
    if (lane == 0) {
      int k = 0;  // pooled rows stored so far
      for (int b = blockIdx.x; b < num_bands; b += gridDim.x) {
        const Band band(b);
        for (int i = band.p0 / 2; i < (band.p0 + kSpBand) / 2; ++i, ++k) {
          const int buf = k & 1;
          mbar_wait(&pool_full[buf], (k >> 1) & 1);
          tma_store_2d(&mapO, smem + L::POOL + buf * kSpPoolRowBytes, 0, (band.n * kSpPool + i) * kSpPool);
          tma_store_commit();
          if (k >= 1) {
            tma_store_wait_read<1>();          // store k-1 has left smem
            mbar_arrive(&pool_free[(k - 1) & 1]);
          }
        }
      }
      tma_store_wait_all<0>();
    }
  }
  } else if (FUSE_K1 && warp >= 12) {
    setmaxnreg_dec<104>();
    // ------------------------------------------------------------------ converters (warps 12..15): K1 into the ring
    // Ring rows in sequence order are (pair seq, parity); warp cw takes pairs with seq % 2 == cw / 2 and parity cw % 2.
    const int cw = warp - 12;
    const int par = cw & 1;
    const int row_cap = (3 * p.W + 32 + 15) & ~15;
    uint8_t* stage = smem + L::STAGE + cw * 4 * row_cap;  // [2 buffers][2 source rows][row_cap]
    const __nv_bfloat16* lut = reinterpret_cast<const __nv_bfloat16*>(smem + L::LUT);
    const uint8_t* buf_begin = p.frames;
    const uint8_t* buf_end = p.frames + static_cast<size_t>(p.n_frames) * p.H * p.W * 3;
    // cursor over this warp's rows: band b, pair j, ring sequence number seq
    int cb = blockIdx.x, cj = 0, cseq = 0;
    bool band_open = false;
    int cj_last = 0, cn = 0;
    auto next_row = [&](int* n, int* h, int* seq) -> bool {
      for (;;) {
        if (!band_open) {
          if (cb >= num_bands) return false;
          const Band band(cb);
          cj = band.j_first;
          cj_last = band.j_last;
          cn = band.n;
          band_open = true;
        }
        if (cj > cj_last) {
          band_open = false;
          cb += gridDim.x;
          continue;
        }
        const int s = cseq++;
        const int j = cj++;
        if ((s & 1) == (cw >> 1)) {
          *n = cn;
          *h = 2 * j + par;
          *seq = s;
          return true;
        }
      }
    };
    // stage the source row(s) of output row h of frame n into buffer `bsel` (asynchronous copies, one commit group)
    int box_n = -1;
    K1Box box;
    auto issue = [&](int n, int h, int bsel) -> K1Row {
      if (n != box_n) {  // per frame: the crop box and its two scale factors (IEEE divisions)
        box = k1_box(n, p.H, p.W, p.boxes);
        box_n = n;
      }
      const K1Row g = k1_row(p.frames, box, n, h, p.H, p.W);
      for (int pass = 0; pass < (g.need1 ? 2 : 1); ++pass) {
        const uint8_t* r = pass == 0 ? g.r0 : g.r1;
        const int off = static_cast<int>(reinterpret_cast<uintptr_t>(r) & 15);
        const uint8_t* src = r - off;
        uint8_t* dst = stage + (bsel * 2 + pass) * row_cap;
        const int nvec = (off + g.bw * 3 + 15) >> 4;
        if (src >= buf_begin && src + 16 * nvec <= buf_end) {
          for (int i = lane; i < nvec; i += 32) cp_async_16(dst + 16 * i, src + 16 * i);
        } else {  // the first / last row of the whole buffer: its first / last vector must stay inside it
          for (int i = lane; i < nvec; i += 32) {
            const uint8_t* gp = src + 16 * i;
            if (gp >= buf_begin && gp + 16 <= buf_end) {
              cp_async_16(dst + 16 * i, gp);
            } else {
              for (int k = 0; k < 16; ++k) dst[16 * i + k] = (gp + k >= buf_begin && gp + k < buf_end) ? gp[k] : 0;
            }
          }
        }
      }
      cp_async_commit();
      return g;
    };
    long long tr[6] = {0, 0, 0, 0, 0, 0};
#ifdef PHDFX_EXPERIMENTAL
    const bool tracing = p.trace != nullptr && blockIdx.x == 0 && cw == 0;
#else
    constexpr bool tracing = false;
#endif
    auto tick = [&](int k, long long& t) {
      if (tracing) {
        const long long now = clock64();
        tr[k] += now - t;
        t = now;
      }
    };
    long long tt = clock64();
    int n0, h0, s0, n1 = 0, h1 = 0, s1 = 0;
    bool have = next_row(&n0, &h0, &s0);
    int bsel = 0;
    K1Row g0, g1;
    if (have) g0 = issue(n0, h0, 0);
    while (have) {
      const bool have_next = next_row(&n1, &h1, &s1);
      if (have_next) {
        __syncwarp();  // the readers of buffer bsel ^ 1 (two rows ago) are done
        g1 = issue(n1, h1, bsel ^ 1);
        tick(0, tt);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      tick(1, tt);
      const int off0 = static_cast<int>(reinterpret_cast<uintptr_t>(g0.r0) & 15);
      const int off1 = static_cast<int>(reinterpret_cast<uintptr_t>(g0.r1) & 15);
      const uint8_t* t0 = stage + (bsel * 2 + 0) * row_cap + off0;
      const uint8_t* t1 = g0.need1 ? stage + (bsel * 2 + 1) * row_cap + off1 : t0;
      // (scale_w of the row being converted: rows of a band share the frame, a band switch recomputes it exactly)
      const float scale_w = __fdiv_rn(static_cast<float>(g0.bw), static_cast<float>(kImg));
      const bool ident = (g0.bh == kImg) && (g0.bw == kImg);
      const int slot = s0 % kSpPairSlots;
      mbar_wait(&pair_empty[slot], ((s0 / kSpPairSlots) & 1) ^ 1);
      tick(2, tt);
      uint2* dst = reinterpret_cast<uint2*>(smem + L::RING + slot * 2 * kSpRowPitch + par * kSpRowPitch) + kStemLeftPad;
      if (ident)
        k1_plain_row<true>(t0, t1, g0, scale_w, p.flip_w, lut, dst, lane);
      else
        k1_plain_row<false>(t0, t1, g0, scale_w, p.flip_w, lut, dst, lane);
      tick(3, tt);
      fence_proxy_async_smem();  // generic-proxy writes -> the tensor core's operand reads (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&pair_full[slot]);
      tick(4, tt);
      if (tracing) ++tr[5];
      have = have_next;
      n0 = n1, h0 = h1, s0 = s1, g0 = g1;
      bsel ^= 1;
    }
    if (tracing && lane == 0)
      for (int k = 0; k < 6; ++k) p.trace[k] = tr[k];
  } else if (warp < 12) {
    if (FUSE_K1) setmaxnreg_inc<176>();
    // ------------------------------------------------------------------ epilogue + pool (warps 4..11)
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;  // 32-channel half of the 64 output channels
    const int q = quad * 32 + lane;    // conv output column owned by this thread (TMEM lane)
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + half * 32;
    const float4* sb4 = reinterpret_cast<const float4*>(s_bias + half * 32);
    int buf = 0;
    uint32_t buf_phase = 0;
    int k = 0;  // pooled rows produced so far
    // conv row (+bias, bf16x2-packed) of this thread's column and channel half
    auto load_row = [&](uint32_t taddr, __nv_bfloat162 (&dst)[16]) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 bb = sb4[c];
        dst[2 * c + 0] = __floats2bfloat162_rn(__uint_as_float(v[4 * c + 0]) + bb.x, __uint_as_float(v[4 * c + 1]) + bb.y);
        dst[2 * c + 1] = __floats2bfloat162_rn(__uint_as_float(v[4 * c + 2]) + bb.z, __uint_as_float(v[4 * c + 3]) + bb.w);
      }
    };
    for (int b = blockIdx.x; b < num_bands; b += gridDim.x) {
      const Band band(b);
      __nv_bfloat162 carry[16];  // conv row 2i-1 of the next pooled row i
      bool have_carry = false;
      if (band.p_first & 1) {
        // halo step: conv row p0-1 only
        mbar_wait(&tmem_full[buf], buf_phase);
        tc_fence_after();
        load_row(t_lane + buf * 128 + 64, carry);
        have_carry = true;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        buf ^= 1;
        if (buf == 0) buf_phase ^= 1;
      }
      for (int i = band.p0 / 2; i < (band.p0 + kSpBand) / 2; ++i, ++k) {
        mbar_wait(&tmem_full[buf], buf_phase);
        tc_fence_after();
        __nv_bfloat162 ev[16], od[16];
        load_row(t_lane + buf * 128, ev);       // conv row 2i
        load_row(t_lane + buf * 128 + 64, od);  // conv row 2i+1
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        buf ^= 1;
        if (buf == 0) buf_phase ^= 1;
        // vertical 3-max in registers
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat162 m = __hmax2(ev[j], od[j]);
          if (have_carry) m = __hmax2(m, carry[j]);
          carry[j] = od[j];
          ev[j] = m;
        }
        have_carry = true;
        // horizontal 3-max (columns 2j-1 .. 2j+1, >= 0) starting from 0: the 0 is both the ReLU and the -inf padding
        // of the reference's MaxPool2d (its input is post-ReLU, resnet.py:270-271).  Thread q holds column q, so the
        // neighbours of an even (= centre) column are one lane away: two warp shuffles per register.  Only lane 0's left
        // neighbour lives in another warp (lane 31 of the previous lane quadrant, same channel half): those 64 bytes
        // go through a tiny double-buffered exchange slot.  Nothing else of the conv row touches shared memory.
        uint32_t* xch = reinterpret_cast<uint32_t*>(smem + L::XCHG) + ((k & 1) * 8 + half * 4 + quad) * 16;
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < 16; ++j) xch[j] = *reinterpret_cast<uint32_t*>(&ev[j]);
        }
        named_barrier_sync(1, kSpEpiThreads);
        const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
        __nv_bfloat162 m[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t self = *reinterpret_cast<uint32_t*>(&ev[j]);
          uint32_t lft = __shfl_up_sync(0xffffffffu, self, 1);
          const uint32_t rgt = __shfl_down_sync(0xffffffffu, self, 1);
          if (lane == 0) lft = quad > 0 ? (xch - 16)[j] : 0u;  // column q-1 of the previous quadrant; none left of q = 0
          m[j] = __hmax2(__hmax2(z, ev[j]), __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&lft),
                                                    *reinterpret_cast<const __nv_bfloat162*>(&rgt)));
        }
        const int pbuf = k & 1;
        if (k >= 2) mbar_wait(&pool_free[pbuf], ((k >> 1) - 1) & 1);
        uint8_t* pool_row = smem + L::POOL + pbuf * kSpPoolRowBytes;
        if (!(q & 1) && q < kSpOut) {  // even columns are the pooling centres: pooled column j = q / 2
          const int j = q >> 1;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&m[4 * c4 + 0]);
            o.y = *reinterpret_cast<uint32_t*>(&m[4 * c4 + 1]);
            o.z = *reinterpret_cast<uint32_t*>(&m[4 * c4 + 2]);
            o.w = *reinterpret_cast<uint32_t*>(&m[4 * c4 + 3]);
            *reinterpret_cast<uint4*>(pool_row + j * 128 + (((half * 4 + c4) ^ (j & 7)) << 4)) = o;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pool_full[pbuf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace phdfxk
