// K2/K3/K4: implicit-GEMM convolution for sm_100a.
//
//   D[M = pixels, N = Cout] = A[M, K] * W[Cout, K]^T        K ordered (r, s, cin)
//
// replaces, per conv, the reference's cuDNN conv + BatchNorm + ReLU (+ residual add) chain
// (torchvision models/resnet.py:143-163 Bottleneck.forward, :268-271 stem; called from
// src/preprocess_resnet_features.py:296 `backbone(x)`).  BN is folded into W and a per-channel fp32 bias
// on the host, so the epilogue is  y = relu?(acc + bias [+ residual]).
//
// Structure: one persistent CTA per SM, 12 warps, warp-specialised.
//   warp 0    TMA producer   A tile via tiled / im2col / stem-window tensor maps, W tile via a 2-D map, both
//                            K-major with hardware swizzle, NSTAGE-deep mbarrier ring
//   warp 1    MMA issuer     one thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered TMEM
//                            accumulator; tcgen05.commit releases smem slots / publishes the accumulator
//   warp 2    epilogue DMA   one thread: TMA-loads the residual tile into the staging buffer ahead of the epilogue
//                            warps and TMA-stores the finished buffer; NB buffers, look-ahead LOOK groups
//                            (the warp also owns the TMEM allocation)
//   warp 4-11 epilogue       two warps per TMEM lane quadrant (= two per SM sub-partition, so one hides the other's
//                            issue latency); per 64-channel group each takes 32 channels: tcgen05.ld -> + bias
//                            (+ residual read from smem) -> bf16 -> ReLU, written IN PLACE into a 128x128B swizzled
//                            staging buffer (conflict-free 16 B accesses)
// All global traffic of the kernel is therefore TMA (full 128 B lines); the only exception is the fused
// global-average-pool mode, which writes 8 KB of fp32 features per frame directly.
#pragma once
#include "ptx_sm100.cuh"

namespace phdfxk {

enum ConvMode : int {
  MODE_TILED = 0,   // 1x1 stride 1: A is the [M, Cin] activation matrix itself
  MODE_IM2COL = 1,  // 3x3 (any stride) and strided 1x1: TMA im2col mode over NHWC
  MODE_STEM = 2,    // 7x7/2 stem: per filter row an 8-pixel x 4-channel window, overlapping-stride 5-D map
  MODE_GAP = 3,     // last 1x1 conv: tile = 2 whole frames (98 rows), epilogue emits the 7x7 mean in fp32
  MODE_HALO = 4,    // 3x3 stride 1 pad 1 on 56x56 (BN=64) / 28x28 (BN=128): the tile is R_t output rows of one frame
                    // in "padded raster" order (row pitch W+2); ONE zero-padded input patch per 64-channel block is
                    // loaded (tiled 4-D TMA, OOB zero fill = the halo) and all nine taps read it through
                    // row-shifted 128B-swizzled descriptors (start = patch + (r*(W+2)+s)*128 B) — 9x less A traffic
                    // than the im2col path; only the weights stream through the stage ring
};

struct ConvParams {
  int M;           // valid GEMM rows in this launch (output pixels)
  int Cout;        // GEMM N
  int num_kb;      // K blocks per tile
  int kb_per_tap;  // Cin / 64 (im2col)
  int S;           // filter width (im2col)
  int P, Q;        // output spatial size
  int stride, pad;
  int m_tiles, n_tiles;
  int relu;
  int has_res;     // residual tile is TMA-loaded through mapR
  int n_frames;
  int halo_rt;     // MODE_HALO: output rows per tile (2 for 56x56, 4 for 28x28)
  int kb_split;    // MODE_TILED with a second source (fused down-sample): K blocks [0, kb_split) come from mapA,
                   // [kb_split, num_kb) from mapA2; = num_kb when there is no second source
  int src2_stride; // second source: 1 = tiled 2-D map over its activation matrix, 2 = im2col map (1x1 stride 2)
  int rev;         // 1: walk the tiles in descending order.  Consecutive layers alternate direction, so a layer starts
                   // on the rows its predecessor wrote last — the part of the activation tensor still resident in L2
  const float* bias;  // [Cout] folded BN bias
  float* feats;       // MODE_GAP: [n_frames, Cout]
  // Frame progress counters (see dep_wait_frames): wait_ctr != nullptr replaces griddepcontrol.wait — the launch starts
  // a tile as soon as the launch that produces its A operand has published the tile's frames; sig_ctr != nullptr makes
  // this launch publish its own output the same way.  Both index by frame of the call.
  unsigned long long* cta_ts;  // debug (PHDFX_CTA_TRACE): 6 words per CTA: %globaltimer at [4] entry, [0] prologue done, [1] first tile's
                               // inputs available, [2] exit; [3] = ns the TMA producer warp spent waiting on counters
  const uint32_t* wait_ctr;
  uint32_t wait_full;  // counter value of a complete input frame: (producer's output rows per frame) x (its Cout / 64)
  uint32_t* sig_ctr;
  // [n_frames] of either array counts CTAs of the producing grid that have published everything (wait_ctas = that
  // grid's size): once a consumer CTA has seen it complete it stops polling — a poll is an L2 round trip per tile
  int ctr_frames;
  uint32_t wait_ctas;
  long long* trace;   // debug (PHDFX_CONV_TRACE): CTA 0 writes clock64() of pipeline events, [tile < 32][32 events]
};

// Experimental code paths of the single-launch kernels are compiled in only with -DPHDFX_EXPERIMENTAL (PHDFX_EXPERIMENTAL=1
// at build()): the per-CTA timeline (PHDFX_CTA_TRACE) and the frame progress counters BETWEEN launches (PHDFX_FLAGS=1).
// Both are dead branches in a normal pass, yet they cost: batch-1 latency under graph replay 325.8 us with both, 320.4
// without the timeline code, 320.6 without the counter code, 316.4 without either (40 launches of one tile each, where
// every instruction in front of the first TMA load is on the critical path).  The multi-phase CTA-pair kernel uses the
// counters in every build.
#ifdef PHDFX_EXPERIMENTAL
#define P_CTA_TS(p) ((p).cta_ts)
#define P_WAIT(p) ((p).wait_ctr)
#define P_SIG(p) ((p).sig_ctr)
#else
#define P_CTA_TS(p) (static_cast<unsigned long long*>(nullptr))
#define P_WAIT(p) (static_cast<const uint32_t*>(nullptr))
#define P_SIG(p) (static_cast<uint32_t*>(nullptr))
#endif

constexpr int kBlockM = 128;
constexpr int kNumThreads = 384;
constexpr int kEpiThreads = 256;   // warps 4..11
constexpr int kEpiWarps = 8;
constexpr int kGapRowsPerFrame = 49;
constexpr int kGapRows = 98;
constexpr int kStemTileQ = 16, kStemTileP = 8, kStemOut = 112, kStemTilesPerFrame = (112 / 16) * (112 / 8);
constexpr int kGroupCols = 64;                        // channels per epilogue group (= one 128 B swizzle row)
constexpr int kStageOutBytes = kBlockM * 128;         // one staging buffer: 128 rows x 128 B

// ---- Frame progress counters: tile-granular dependencies between ADJACENT launches -------------------------------
// Every kernel of the library is a persistent grid of one CTA per SM that triggers its dependents (PDL) at start, so the
// next launch's CTAs take over SMs as this launch's CTAs exit.  With griddepcontrol.wait they then sit idle until the
// WHOLE grid has drained — up to a full tile time when the tile count is not a multiple of the grid (layer4: 98
// pair-tiles on 74 CTA pairs).  Instead, a producer launch adds (rows x 64-channel groups) to a per-frame counter once
// the TMA stores of a tile have completed, and the consumer's TMA producer warp polls the counters of the frames a tile
// reads before its first load.  Frames are the unit because every consumer on this path needs whole frames (3x3) or a
// row range inside one or two frames (1x1).  Only the predecessor's output is guarded this way; what a launch reads from
// further back (residual, down-sample source) was complete before this launch could start: a grid of num_sms CTAs with
// one CTA per SM is fully resident only after every CTA of the grid before it has exited, and the dependent launch
// starts only after every CTA of this grid has started.  The host links two launches only under those conditions
// (api.cu: plan_links) and zeroes the counters at the start of every pass.
// Whole converged warp: lane i owns frame f_lo + i of the tile (a tile touches at most a few frames).
__device__ __forceinline__ void dep_wait_frames(const uint32_t* ctr, uint32_t full, int f_lo, int f_hi) {
  const int f = f_lo + static_cast<int>(threadIdx.x & 31);
  if (f <= f_hi) {
    uint32_t v = 0u;
#ifdef PHDFX_TRAP
    uint32_t spins = 0;
    unsigned long long t0 = 0;
#endif
    while (v < full) {
      v = ld_acquire_gpu(ctr + f);
#ifdef PHDFX_TRAP
      if ((++spins & 0x3FFu) == 0) {
        const unsigned long long now = global_timer_ns();
        if (t0 == 0)
          t0 = now;
        else if (now - t0 > kMbarTrapNs)
          __trap();
      }
#endif
    }
  }
  __syncwarp();
  fence_proxy_async_all();  // the TMA loads that follow (async proxy) observe what the acquire made visible
}
// whole converged warp: has every CTA of the producing grid published all its tiles?
__device__ __forceinline__ bool dep_grid_done(const uint32_t* done_ctr, uint32_t grid_ctas) {
  const bool done = ld_acquire_gpu(done_ctr) >= grid_ctas;
  if (done) fence_proxy_async_all();
  return done;
}
// one thread, after the TMA stores of output rows [r0, r1) have COMPLETED (wait_group without .read): `per_row` = the
// 64-channel groups of the tile
__device__ __forceinline__ void dep_signal_rows(uint32_t* ctr, int r0, int r1, int pq, uint32_t per_row) {
  fence_proxy_async_all();
  int f = r0 / pq;
  int f_end = (f + 1) * pq;
  while (r0 < r1) {
    const int e = r1 < f_end ? r1 : f_end;
    red_release_gpu_add(ctr + f, static_cast<uint32_t>(e - r0) * per_row);
    r0 = e;
    ++f;
    f_end += pq;
  }
}

template <int BN, int MODE>
struct ConvCfg {
  static constexpr int ROWB = (MODE == MODE_STEM) ? 64 : 128;  // bytes per smem operand row (= BLOCK_K bf16)
  static constexpr int BLOCK_K = ROWB / 2;
  static constexpr int A_BYTES = (MODE == MODE_HALO) ? 0 : kBlockM * ROWB;  // HALO: A lives in the patch buffers
  static constexpr int B_BYTES = BN * ROWB;
  // MODE_HALO patch buffers: (R_t+2) x (W+2) positions x 128 B, sized so the furthest shifted 128-row window stays
  // inside: 56x56 (BN=64): 4x58 = 232 loaded, 118+128 = 246 read -> 32 KB; 28x28 (BN=128): 6x30 = 180, 62+128 = 190
  // read -> 24 KB
  static constexpr int HALO_BYTES = (MODE == MODE_HALO) ? (BN == 64 ? 32768 : 24576) : 0;
  static constexpr int HB = (MODE == MODE_HALO) ? (BN == 64 ? 3 : 4) : 0;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int A_TX = ((MODE == MODE_GAP) ? kGapRows : kBlockM) * ROWB;
  // BN = 256 1x1 convs (the expand convs with their residual): a residual TMA load takes ~2 800 cycles from issue to
  // arrival while a group is processed in ~800 (PHDFX_CONV_TRACE), so three loads have to be in flight: 5 staging
  // buffers, paid for with a single-buffered bias (one more named barrier per tile)
  static constexpr bool DEEP = (BN == 256 && MODE == MODE_TILED);
  static constexpr int NB = DEEP ? 5 : (MODE == MODE_GAP || MODE == MODE_HALO) ? 2 : 4;  // epilogue staging buffers
  static constexpr int LOOK = DEEP ? 3 : (NB == 2) ? 1 : 2;    // residual loads run LOOK groups ahead of the stores
  static constexpr int GROUPS = BN / kGroupCols;
  static constexpr int SCRATCH_BYTES = (MODE == MODE_GAP) ? 2 * kBlockM * 33 * 4 : 0;
  static constexpr int TAIL_BYTES = 1024 + (DEEP ? 1 : 2) * BN * 4 + SCRATCH_BYTES;  // barriers + bias buffer(s) + scratch
  static constexpr int SMEM_MAX = 232448;                              // 227 KB
  static constexpr int NSTAGE_RAW =
      (SMEM_MAX - 1024 - TAIL_BYTES - NB * kStageOutBytes - HB * HALO_BYTES) / STAGE_BYTES;
  // MODE_HALO with BN = 64 (layer1: Cin = Cout = 64): the nine 8 KB weight tiles stay RESIDENT in the nine "stages" for
  // the whole kernel (loaded once), so per tile only the input patch moves.
  static constexpr bool RES_B = (MODE == MODE_HALO && BN == 64);
  static constexpr int NSTAGE = RES_B ? 9 : (NSTAGE_RAW > 8 ? 8 : NSTAGE_RAW);
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + HB * HALO_BYTES + NB * kStageOutBytes + TAIL_BYTES + 1024;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {64,128,256}
};

template <int BN, int MODE>
__global__ void __launch_bounds__(kNumThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapR,
                  const __grid_constant__ CUtensorMap mapA2, const ConvParams p) {
  using Cfg = ConvCfg<BN, MODE>;
  constexpr int NSTAGE = Cfg::NSTAGE;
  constexpr int NB = Cfg::NB;
  constexpr int LOOK = Cfg::LOOK;
  constexpr int GROUPS = Cfg::GROUPS;
  constexpr int HBD = Cfg::HB > 0 ? Cfg::HB : 1;  // patch-buffer count (1 keeps dead non-HALO code well-formed)
  static_assert(NSTAGE >= 2, "pipeline needs at least two stages");
  static_assert(Cfg::SMEM_BYTES <= Cfg::SMEM_MAX, "shared memory budget exceeded");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* halo = smem + NSTAGE * Cfg::STAGE_BYTES;       // MODE_HALO: [HB] input patches, 1024-aligned
  uint8_t* stage_out = halo + Cfg::HB * Cfg::HALO_BYTES;  // [NB][128][128 B], 1024-aligned
  uint8_t* tail = stage_out + NB * kStageOutBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);  // [NSTAGE]
  uint64_t* empty_bar = full_bar + NSTAGE;                 // [NSTAGE]
  uint64_t* tmem_full = empty_bar + NSTAGE;                // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]
  uint64_t* res_full = tmem_empty + 2;                     // [NB] staging buffer holds the residual / is free
  uint64_t* out_full = res_full + NB;                      // [NB] epilogue warps are done with the buffer
  uint64_t* halo_full = out_full + NB;                     // [HB]
  uint64_t* halo_empty = halo_full + (Cfg::HB > 0 ? Cfg::HB : 1);  // [HB]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(halo_empty + (Cfg::HB > 0 ? Cfg::HB : 1));
  float* s_bias = reinterpret_cast<float*>(tail + 1024);   // [2][BN]
  float* s_scratch = s_bias + 2 * BN;                      // MODE_GAP: [2 halves][128][33]

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  if (P_CTA_TS(p) != nullptr && threadIdx.x == 0) P_CTA_TS(p)[blockIdx.x * 6 + 4] = global_timer_ns();  // kernel entry
  const int num_tiles = p.m_tiles * p.n_tiles;
  // debug timeline: event e of this CTA's k-th tile (CTA 0 only, first 32 tiles)
  auto mark = [&](int k, int e) {
    if (p.trace != nullptr && blockIdx.x == 0 && k < 32 && lane == 0) p.trace[k * 32 + e] = clock64();
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if (MODE == MODE_TILED && p.kb_split < p.num_kb) tma_prefetch_desc(&mapA2);
  }
  if (warp == 2 && lane == 0) {
    if (MODE != MODE_GAP) tma_prefetch_desc(&mapO);
    if (p.has_res) tma_prefetch_desc(&mapR);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kEpiWarps);  // one arrive per epilogue warp
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&out_full[i], kEpiWarps);
    }
    for (int i = 0; i < Cfg::HB; ++i) {
      mbar_init(&halo_full[i], 1);
      mbar_init(&halo_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) may overlap the previous kernel's
  // tail; nothing below touches activations before the previous grid has completed.
  griddep_launch_dependents();
  if (P_CTA_TS(p) != nullptr && threadIdx.x == 0) P_CTA_TS(p)[blockIdx.x * 6 + 0] = global_timer_ns();
  if (P_WAIT(p) == nullptr) griddep_wait();
  if (P_CTA_TS(p) != nullptr && threadIdx.x == 0 && P_WAIT(p) == nullptr) P_CTA_TS(p)[blockIdx.x * 6 + 1] = global_timer_ns();
  unsigned long long dep_ns = 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {  // whole warp, warp-uniform control flow; one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      int hseq = 0;  // MODE_HALO: patches loaded so far
      if (Cfg::RES_B) {
        for (int tap = 0; tap < 9; ++tap) {  // resident weights: tile `tap` -> stage `tap`, loaded once
          mbar_arrive_expect_tx_elect(&full_bar[tap], Cfg::B_BYTES);
          tma_load_2d_elect(&mapB, &full_bar[tap], smem + tap * Cfg::STAGE_BYTES, tap * 64, 0);
        }
      }
      bool dep_all = false;  // frame progress counters: the producing grid is known to have published everything
      auto frames_of = [&](int mb, int* f_lo, int* f_hi) {
        if (MODE == MODE_GAP) {
          *f_lo = mb * 2;
          *f_hi = mb * 2 + 1 < p.n_frames ? mb * 2 + 1 : p.n_frames - 1;
        } else {
          const int pq = p.P * p.Q;
          const int r1 = (mb + 1) * kBlockM < p.M ? (mb + 1) * kBlockM : p.M;
          *f_lo = (mb * kBlockM) / pq;
          *f_hi = (r1 - 1) / pq;
        }
      };
      for (int lt = blockIdx.x; lt < num_tiles; lt += gridDim.x) {
        const int tile = p.rev ? num_tiles - 1 - lt : lt;
        const int m_blk = tile / p.n_tiles;
        const int n_blk = tile - m_blk * p.n_tiles;
        // per-tile A coordinates
        int cw = 0, ch = 0, cn = 0;
        if (MODE == MODE_IM2COL || (MODE == MODE_TILED && p.src2_stride == 2)) {
          // base pixel of the tile in the (second) source's input coordinates; for the fused stride-2 down-sample
          // pad is 0 and stride 2
          const int m0 = m_blk * kBlockM;
          const int pq = p.P * p.Q;
          cn = m0 / pq;
          const int rem = m0 - cn * pq;
          const int p0 = rem / p.Q;
          const int q0 = rem - p0 * p.Q;
          if (MODE == MODE_IM2COL) {
            cw = q0 * p.stride - p.pad;
            ch = p0 * p.stride - p.pad;
          } else {
            cw = q0 * 2;
            ch = p0 * 2;
          }
        } else if (MODE == MODE_STEM) {
          cn = m_blk / kStemTilesPerFrame;
          const int t = m_blk - cn * kStemTilesPerFrame;
          ch = (t / (kStemOut / kStemTileQ)) * kStemTileP;  // p0
          cw = (t % (kStemOut / kStemTileQ)) * kStemTileQ;  // q0
        } else if (MODE == MODE_HALO) {
          const int tpf = p.P / p.halo_rt;  // tiles per frame
          cn = m_blk / tpf;
          ch = (m_blk - cn * tpf) * p.halo_rt - 1;  // first input row of the patch (-1 = zero halo)
        }
        if (P_WAIT(p) != nullptr) {
          // frames this tile reads from the previous launch (= the frames of its output rows); then read ahead for
          // the next tile of this CTA
          int f_lo, f_hi;
          const unsigned long long tw = P_CTA_TS(p) != nullptr ? global_timer_ns() : 0ull;
          if (!dep_all) dep_all = dep_grid_done(P_WAIT(p) + p.ctr_frames, p.wait_ctas);
          if (!dep_all) {
            frames_of(m_blk, &f_lo, &f_hi);
            dep_wait_frames(P_WAIT(p), p.wait_full, f_lo, f_hi);
          }
          if (P_CTA_TS(p) != nullptr) {
            const unsigned long long now = global_timer_ns();
            dep_ns += now - tw;
            if (lt == static_cast<int>(blockIdx.x) && lane == 0) P_CTA_TS(p)[blockIdx.x * 6 + 1] = now;
          }
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (kb == 0) mark((lt - blockIdx.x) / gridDim.x, 0);
          if (MODE == MODE_HALO) {
            // kb = cb * 9 + tap: one input patch per 64-channel block, then its nine weight tiles
            const int cb = kb / 9;
            const int tap = kb - cb * 9;
            if (tap == 0) {
              const int hb = hseq % HBD;
              mbar_wait(&halo_empty[hb], ((hseq / HBD) & 1) ^ 1);
              mbar_arrive_expect_tx_elect(&halo_full[hb], 128 * (p.Q + 2) * (p.halo_rt + 2));
              tma_load_4d_elect(&mapA, &halo_full[hb], halo + hb * Cfg::HALO_BYTES, cb * 64, -1, ch, cn);
              ++hseq;
            }
            if (Cfg::RES_B) continue;  // weights are resident
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx_elect(&full_bar[stage], Cfg::B_BYTES);
            tma_load_2d_elect(&mapB, &full_bar[stage], smem + stage * Cfg::STAGE_BYTES,
                              (tap * p.kb_per_tap + cb) * 64, n_blk * BN);
            if (++stage == NSTAGE) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          mbar_arrive_expect_tx_elect(&full_bar[stage], Cfg::A_TX + Cfg::B_BYTES);
          if (MODE == MODE_TILED) {
            if (kb < p.kb_split)
              tma_load_2d_elect(&mapA, &full_bar[stage], sA, kb * Cfg::BLOCK_K, m_blk * kBlockM);
            else if (p.src2_stride == 1)
              tma_load_2d_elect(&mapA2, &full_bar[stage], sA, (kb - p.kb_split) * Cfg::BLOCK_K, m_blk * kBlockM);
            else
              tma_load_im2col_4d_elect(&mapA2, &full_bar[stage], sA, (kb - p.kb_split) * Cfg::BLOCK_K, cw, ch, cn, 0,
                                       0);
            tma_load_2d_elect(&mapB, &full_bar[stage], sB, kb * Cfg::BLOCK_K, n_blk * BN);
          } else if (MODE == MODE_IM2COL) {
            const int tap = kb / p.kb_per_tap;
            const int cb = kb - tap * p.kb_per_tap;
            const int r = tap / p.S;
            const int s = tap - r * p.S;
            tma_load_im2col_4d_elect(&mapA, &full_bar[stage], sA, cb * Cfg::BLOCK_K, cw, ch, cn,
                               static_cast<uint16_t>(s), static_cast<uint16_t>(r));
            tma_load_2d_elect(&mapB, &full_bar[stage], sB, kb * Cfg::BLOCK_K, n_blk * BN);
          } else if (MODE == MODE_STEM) {
            const int d = kb - 3;  // input row = 2*p + d
            tma_load_5d_elect(&mapA, &full_bar[stage], sA, 0, cw, d & 1, ch + (d >> 1), cn);
            tma_load_2d_elect(&mapB, &full_bar[stage], sB, 0, kb * BN);
          } else {  // MODE_GAP
            tma_load_3d_elect(&mapA, &full_bar[stage], sA, kb * Cfg::BLOCK_K, 0, m_blk * 2);
            tma_load_2d_elect(&mapB, &full_bar[stage], sB, kb * Cfg::BLOCK_K, n_blk * BN);
          }
          if (++stage == NSTAGE) {
            stage = 0;
            phase ^= 1;
          }
          if (kb == p.num_kb - 1) mark((lt - blockIdx.x) / gridDim.x, 1);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    {  // whole warp, warp-uniform control flow; one elected lane issues
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int hseq = 0;  // MODE_HALO: patches consumed so far
      bool res_b_ready = false;  // RES_B: the nine resident weight tiles have landed (checked during the first tile)
      int tk = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tk) {
        mark(tk, 2);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mark(tk, 3);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        // `ready`: the barrier of the stage about to be consumed is already known to be complete (probed while the
        // previous K block's MMAs were being issued), so the ~90-cycle query is off the issue path
        bool ready = false;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          int tap = 0;
          if (MODE == MODE_HALO) {
            tap = kb % 9;
            if (tap == 0) mbar_wait(&halo_full[hseq % HBD], (hseq / HBD) & 1);
          }
          if (Cfg::RES_B) {
            stage = tap;  // resident weight tile; its barrier completed phase 0 once and never advances
            phase = 0;
            if (!res_b_ready) mbar_wait(&full_bar[stage], 0);
          } else if (!ready) {
            mbar_wait(&full_bar[stage], phase);
          }
          tc_fence_after();
          uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
          if (MODE == MODE_HALO) {
            // tap (r, s): the same patch, shifted by r rows and s pixels of the padded raster
            const int r = tap / 3;
            const int sft = r * (p.Q + 2) + (tap - r * 3);
            a_addr = smem_u32(halo + (hseq % HBD) * Cfg::HALO_BYTES) + sft * 128;
          }
          // probe the next stage before issuing (non-blocking)
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == NSTAGE) {
            nstage = 0;
            nphase ^= 1;
          }
          ready = false;
          if (!Cfg::RES_B && kb + 1 < p.num_kb) ready = mbar_test(&full_bar[nstage], nphase);
          // descriptors once per K block; each K=16 step advances the start address by 32 B (field unit 16 B)
          const uint64_t adesc0 = make_kmajor_desc(a_addr, Cfg::ROWB);
          const uint64_t bdesc0 = make_kmajor_desc(b_addr, Cfg::ROWB);
          if (Cfg::BLOCK_K == 64)
            umma_bf16_x4_elect(d_tmem, adesc0, bdesc0, idesc, kb != 0 ? 1u : 0u);
          else
            umma_bf16_x2_elect(d_tmem, adesc0, bdesc0, idesc, kb != 0 ? 1u : 0u);
          if (!Cfg::RES_B) umma_commit_elect(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (MODE == MODE_HALO && tap == 8) {
            umma_commit_elect(&halo_empty[hseq % HBD]);  // all nine taps of this patch have been issued
            ++hseq;
          }
          if (!Cfg::RES_B) {
            stage = nstage;
            phase = nphase;
          }
        }
        res_b_ready = true;
        mark(tk, 4);
        umma_commit_elect(&tmem_full[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ epilogue DMA (one thread)
    if (lane == 0) {
      const int my_tiles = (num_tiles > static_cast<int>(blockIdx.x))
                               ? (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                     static_cast<int>(gridDim.x)
                               : 0;
      const int J = my_tiles * GROUPS;
      constexpr int kSigLag = 2;
      const bool sig_now = p.num_kb >= 16;
      // publish this CTA's ti-th tile (its stores have completed): rows x GROUPS per frame
      auto signal_tile = [&](int ti) {
        const int lt = blockIdx.x + ti * gridDim.x;
        const int tile = p.rev ? num_tiles - 1 - lt : lt;
        const int m_blk = tile / p.n_tiles;
        const int r1 = (m_blk + 1) * kBlockM < p.M ? (m_blk + 1) * kBlockM : p.M;
        dep_signal_rows(P_SIG(p), m_blk * kBlockM, r1, p.P * p.Q, GROUPS);
      };
      for (int t = 0; t < J + LOOK; ++t) {
        if (t < J) {
          // A(t): make buffer t % NB available to the epilogue warps (with the residual tile in it, if any)
          if (MODE != MODE_GAP && t >= NB) tma_store_wait_read<NB - LOOK - 1>();  // store t-NB has left smem
          const int b = t % NB;
          if (p.has_res) {
            const int lt = blockIdx.x + (t / GROUPS) * gridDim.x;
            const int tile = p.rev ? num_tiles - 1 - lt : lt;
            const int g = t % GROUPS;
            const int m_blk = tile / p.n_tiles;
            const int n_blk = tile - m_blk * p.n_tiles;
            const int row0 = m_blk * ((MODE == MODE_GAP) ? kGapRows : kBlockM);
            mbar_arrive_expect_tx(&res_full[b], kStageOutBytes);
            tma_load_2d(&mapR, &res_full[b], stage_out + b * kStageOutBytes, n_blk * BN + g * kGroupCols, row0);
            if (g < 4) mark(t / GROUPS, 8 + g);
          } else {
            mbar_arrive(&res_full[b]);
          }
        }
        if (t >= LOOK) {
          // B(u): the epilogue warps finished buffer u % NB -> store it
          const int u = t - LOOK;
          const int b = u % NB;
          mbar_wait(&out_full[b], (u / NB) & 1);
          if (MODE != MODE_GAP) {
            const int lt = blockIdx.x + (u / GROUPS) * gridDim.x;
            const int tile = p.rev ? num_tiles - 1 - lt : lt;
            const int g = u % GROUPS;
            const int m_blk = tile / p.n_tiles;
            const int n_blk = tile - m_blk * p.n_tiles;
            const uint8_t* src = stage_out + b * kStageOutBytes;
            if (MODE == MODE_STEM) {
              const int n = m_blk / kStemTilesPerFrame;
              const int tt = m_blk - n * kStemTilesPerFrame;
              const int p0 = (tt / (kStemOut / kStemTileQ)) * kStemTileP;
              const int q0 = (tt % (kStemOut / kStemTileQ)) * kStemTileQ;
              tma_store_4d(&mapO, src, 0, q0, p0, n);
            } else if (MODE == MODE_HALO) {
              const int tpf = p.P / p.halo_rt;
              const int n = m_blk / tpf;
              tma_store_4d(&mapO, src, n_blk * BN + g * kGroupCols, 0, (m_blk - n * tpf) * p.halo_rt, n);
            } else {
              tma_store_2d(&mapO, src, n_blk * BN + g * kGroupCols, m_blk * kBlockM);
            }
            tma_store_commit();
            if (g < 4) mark(u / GROUPS, 12 + g);
            if ((MODE == MODE_TILED || MODE == MODE_IM2COL) && P_SIG(p) != nullptr) {
              // Publish a tile once its stores have been WRITTEN (wait_group without .read).  MMA-bound launches (long K)
              // emit a tile's groups in a burst and then nothing for a long time: wait right away, this thread has nothing
              // else to do.  Epilogue-bound launches emit groups continuously: publish kSigLag groups late, when the wait
              // returns at once, so the residual loads this thread issues for the groups ahead are not held up.
              if (sig_now) {
                if (g == GROUPS - 1) {
                  tma_store_wait_all<0>();
                  signal_tile(u / GROUPS);
                }
              } else if (u >= kSigLag && (u - kSigLag) % GROUPS == GROUPS - 1) {
                tma_store_wait_all<kSigLag>();
                signal_tile((u - kSigLag) / GROUPS);
              }
            }
          }
        }
      }
      // (exiting on wait_group.read — the writes then complete with the grid — was measured: no change at batch 256,
      // and +0.3 us per launch at batch 1, where the next launch's griddepcontrol.wait sits right behind this exit)
      if (MODE != MODE_GAP) tma_store_wait_all<0>();
      if ((MODE == MODE_TILED || MODE == MODE_IM2COL) && P_SIG(p) != nullptr) {
        if (!sig_now)
          for (int w = (J > kSigLag ? J - kSigLag : 0); w < J; ++w)
            if (w % GROUPS == GROUPS - 1) signal_tile(w / GROUPS);
        red_release_gpu_add(P_SIG(p) + p.ctr_frames, 1u);  // this CTA has published everything
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    const int quad = warp & 3;            // TMEM lane quadrant this warp may read
    const int half = (warp - 4) >> 2;     // which 32 channels of every 64-channel group this warp converts
    const int row = quad * 32 + lane;     // row of the 128-row tile owned by this thread
    const int et = threadIdx.x - 128;     // 0..255
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    int jg = 0;  // running group counter (matches the DMA thread's t / u)
    for (int lt = blockIdx.x; lt < num_tiles; lt += gridDim.x, ++it) {
      const int tile = p.rev ? num_tiles - 1 - lt : lt;
      const int m_blk = tile / p.n_tiles;
      const int n_blk = tile - m_blk * p.n_tiles;
      const int n_base = n_blk * BN;
      float* sb = s_bias + (Cfg::DEEP ? 0 : (it & 1) * BN);
      if (Cfg::DEEP && it > 0) named_barrier_sync(2, kEpiThreads);  // everyone is done with the previous tile's bias
      for (int i = et; i < BN; i += kEpiThreads) sb[i] = __ldg(&p.bias[n_base + i]);
      named_barrier_sync(1, kEpiThreads);

      bool row_ok = true;
      if (MODE == MODE_GAP) row_ok = (row < kGapRows) && (m_blk * kGapRows + row < p.M);
      int srow = row;  // row of the staging buffer this thread writes
      if (MODE == MODE_HALO) {
        // tile row = padded-raster position i*(W+2)+j; only j < W, i < R_t are outputs; staging is dense [R_t][W]
        const int wp = p.Q + 2;
        const int i = row / wp;
        const int j = row - i * wp;
        row_ok = (i < p.halo_rt) && (j < p.Q);
        srow = i * p.Q + j;
      }

      if (warp == 4) mark(it, 16);
      mbar_wait(&tmem_full[acc], acc_phase);
      if (warp == 4) mark(it, 17);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;

#pragma unroll 1
      for (int g = 0; g < GROUPS; ++g, ++jg) {
        const int b = jg % NB;
        mbar_wait(&res_full[b], (jg / NB) & 1);
        if (warp == 4 && g < 4) mark(it, 18 + g);
        uint8_t* row_ptr = stage_out + b * kStageOutBytes + srow * 128;
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + g * kGroupCols + half * 32, v);
        tmem_ld_wait();
        const float4* sb4 = reinterpret_cast<const float4*>(sb + g * kGroupCols + half * 32);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          // 16-byte chunk (8 channels) of this thread's row; 128B-swizzle: physical chunk = logical ^ (row & 7)
          uint4* sp = reinterpret_cast<uint4*>(row_ptr + (((half * 4 + c4) ^ (srow & 7)) << 4));
          const float4 b0 = sb4[2 * c4], b1 = sb4[2 * c4 + 1];
          float f[8];
          f[0] = __uint_as_float(v[8 * c4 + 0]) + b0.x;
          f[1] = __uint_as_float(v[8 * c4 + 1]) + b0.y;
          f[2] = __uint_as_float(v[8 * c4 + 2]) + b0.z;
          f[3] = __uint_as_float(v[8 * c4 + 3]) + b0.w;
          f[4] = __uint_as_float(v[8 * c4 + 4]) + b1.x;
          f[5] = __uint_as_float(v[8 * c4 + 5]) + b1.y;
          f[6] = __uint_as_float(v[8 * c4 + 6]) + b1.z;
          f[7] = __uint_as_float(v[8 * c4 + 7]) + b1.w;
          if (p.has_res) {
            const uint4 rv = *sp;
            f[0] += bf16_lo(rv.x);
            f[1] += bf16_hi(rv.x);
            f[2] += bf16_lo(rv.y);
            f[3] += bf16_hi(rv.y);
            f[4] += bf16_lo(rv.z);
            f[5] += bf16_hi(rv.z);
            f[6] += bf16_lo(rv.w);
            f[7] += bf16_hi(rv.w);
          }
          if (MODE == MODE_GAP) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = p.relu ? fmaxf(f[j], 0.0f) : f[j];
              s_scratch[(half * kBlockM + row) * 33 + c4 * 8 + j] = row_ok ? x : 0.0f;
            }
          } else {
            // round once to bf16, ReLU on the packed pairs (max commutes with the rounding)
            __nv_bfloat162 o2[4];
            o2[0] = __floats2bfloat162_rn(f[0], f[1]);
            o2[1] = __floats2bfloat162_rn(f[2], f[3]);
            o2[2] = __floats2bfloat162_rn(f[4], f[5]);
            o2[3] = __floats2bfloat162_rn(f[6], f[7]);
            if (p.relu) {
              const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
              for (int j = 0; j < 4; ++j) o2[j] = __hmax2(o2[j], z);
            }
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&o2[0]);
            o.y = *reinterpret_cast<uint32_t*>(&o2[1]);
            o.z = *reinterpret_cast<uint32_t*>(&o2[2]);
            o.w = *reinterpret_cast<uint32_t*>(&o2[3]);
            if (MODE != MODE_HALO || row_ok) *sp = o;
          }
        }
        if (MODE == MODE_GAP) {
          // deterministic in-CTA mean over the 49 rows of each of the tile's two frames
          named_barrier_sync(2, kEpiThreads);
          if (et < 128) {
            const int hh = et >> 6;          // which 32-channel half
            const int fr = (et >> 5) & 1;    // which of the tile's two frames
            const int col = et & 31;
            const int frame = m_blk * 2 + fr;
            float acc_sum = 0.0f;
            for (int r = 0; r < kGapRowsPerFrame; ++r)
              acc_sum += s_scratch[(hh * kBlockM + fr * kGapRowsPerFrame + r) * 33 + col];
            if (frame < p.n_frames)
              p.feats[static_cast<size_t>(frame) * p.Cout + n_base + g * kGroupCols + hh * 32 + col] =
                  acc_sum * (1.0f / kGapRowsPerFrame);
          }
          named_barrier_sync(3, kEpiThreads);
        }
        // hand the buffer to the DMA thread (generic-proxy writes -> async proxy)
        if (MODE != MODE_GAP) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_full[b]);
        if (warp == 4 && g < 4) mark(it, 22 + g);
      }
      // release the accumulator buffer to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  if (P_CTA_TS(p) != nullptr && threadIdx.x == 0) P_CTA_TS(p)[blockIdx.x * 6 + 3] = dep_ns;
  tc_fence_before();
  __syncthreads();
  if (P_CTA_TS(p) != nullptr && threadIdx.x == 0) P_CTA_TS(p)[blockIdx.x * 6 + 2] = global_timer_ns();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace phdfxk
