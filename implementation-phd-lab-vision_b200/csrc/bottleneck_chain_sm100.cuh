// K3c: "bottleneck chain" — conv2 (3x3) -> conv3 (1x1, + identity or fused down-sample) -> ReLU
//      [-> conv1 of the NEXT block (1x1)] in ONE kernel, for the stride-1 blocks of layer1 (56x56, width 64, 256
//      output channels) and layer2 (28x28, width 128, 512 output channels).
//
// Replaces, per block (torchvision models/resnet.py:150-161 and, in layer1, the following block's :146-148), two or
// three launches of conv_igemm_kernel and their round trips through HBM:
//   * t2 (conv2's output) never leaves the SM: the epilogue warps round it to bf16 and park it in TENSOR MEMORY
//     (tcgen05.st, two channels per 32-bit column); conv3 takes it from there as its A operand (tcgen05.mma with A in
//     TMEM) — no shared-memory traffic at all for t2;
//   * layer1: the block output (256 channels, the largest activation of the network) is stored once (TMA, through
//     rotating staging buffers) and ALSO written back, as bf16, over the accumulator columns it was computed from;
//     from there it is the A operand (K = 256) of the next block's 1x1 reduce conv — that conv no longer re-reads
//     1.6 MB per frame from HBM, and its A operand costs no shared-memory bandwidth either;
//   * the residual tile is TMA-loaded into the staging buffer ahead of the epilogue and added in place (as in
//     conv_igemm_kernel), by a loader warp that runs as far ahead as free staging buffers allow;
//   * the tensor-bound conv2 of tile k+1 runs while the HBM-/epilogue-bound conv3 + residual of tile k drains, which two
//     separate launches can never do.
//
// Tile = RT output rows of one frame in "padded raster" order (row pitch W+2): GEMM row m = i*(W+2) + j, valid when
// i < RT and j < W (RT = 2 on 56x56: 116 of 128 rows carry pixels; RT = 4 on 28x28: 120).  Every GEMM of the chain
// keeps that row order, rows are independent in all of them, and the TMA stores clip the two pad columns — so the
// invalid rows never need masking.  conv2 reads one zero-padded input patch per 64-channel block through row-shifted
// 128B-swizzled descriptors (conv_igemm_sm100.cuh, MODE_HALO).
//
// Warps (13): 0 weight producer (mbarrier ring of [C2 rows][64 K] tiles: conv2 taps, layer2's conv3 tiles, layer1's
// next-conv1 tiles; layer1's conv3 weights are resident), 1 MMA issuer, 2 store DMA (+ TMEM owner), 3 activation
// producer (input patches, down-sample source), 4-11 epilogue, 12 residual loader.  The MMA warp software-pipelines
// across tiles:
//     layer1:  conv3(k) | conv2(k+1) | conv1n(k)                       epilogue:  B(k) | A(k+1) | C(k)
//     layer2:  conv3.lo(k) | conv2(k+1).cb0 | conv3.hi(k) | conv2(k+1).cb1        B.lo(k) | B.hi(k) | A(k+1)
// (A = conv2's, B = conv3's, C = next-conv1's epilogue; layer2's 512 output channels go through the 256-column conv3
// accumulator in two halves), so the long conv2 always overlaps the long residual epilogue.  Tensor memory: conv2
// accumulator C2 columns, t2 (bf16) C2/2, conv3 accumulator 256 (in layer1 the first 16 columns of every 32 are
// re-used for the bf16 block output), next-conv1 accumulator N1; all single-buffered — the data dependencies of the
// chain and the in-order tensor pipe order every reuse.
#pragma once
#include "conv_igemm_sm100.cuh"

namespace phdfxk {

struct ChainParams {
  int n_frames;
  int num_tiles;        // n_frames * W / RT
  int rev;              // walk tiles in descending order (see ConvParams::rev)
  const float* bias2;   // [C2] conv2 folded-BN bias
  const float* bias3;   // [N3] conv3 (+ down-sample) bias
  const float* bias1n;  // [N1] next block's conv1 bias (N1 > 0)
  long long* trace;     // debug (PHDFX_CHAIN_TRACE): CTA 0 writes clock64() of pipeline events, [tile < 32][32 events]
};

constexpr int kChainThreads = 416;  // 13 warps

// W: spatial size (56 | 28); C2: bottleneck width (64 | 128); N3: block output channels (256 | 512);
// HAS_DS: conv3 carries the block's down-sample branch as a second K block; N1: width of the fused next conv1 (0 = none)
template <int W, int C2, int N3, bool HAS_DS, int N1>
struct ChainCfg {
  static constexpr int RT = (W == 56) ? 2 : 4;                     // output rows per tile
  static constexpr int WP = W + 2;                                 // padded-raster row pitch
  static constexpr int TILES_PER_FRAME = W / RT;
  static constexpr int TILE_ROWS = RT * WP;                        // padded-raster rows that hold output pixels
  static constexpr int TILE_BYTES = TILE_ROWS * 128;               // one 64-channel group of a tile = one TMA box
  static constexpr int PATCH_LOAD_BYTES = (RT + 2) * WP * 128;     // (RT+2) x (W+2) positions x 64 channels
  static constexpr int PATCH_BYTES = (PATCH_LOAD_BYTES + 1023) / 1024 * 1024;
  static constexpr int CB = C2 / 64;                               // 64-channel blocks of conv2's input = patches per tile
  static constexpr int NP = (CB == 1) ? 2 : 3;                     // patch slots
  static constexpr int NH = N3 / 256;                              // 256-channel halves of conv3's output
  static constexpr bool W3_RES = (N3 == 256);                      // conv3 weights resident (else streamed via the ring)
  static constexpr int W3_KB = W3_RES ? (HAS_DS ? 2 : 1) : 0;      // resident K blocks (t2 | x for the down-sample)
  static constexpr int W3_BYTES = W3_KB * 256 * 128;
  static constexpr int RING_STAGE = C2 * 128;                      // one [C2 rows][64 K] weight tile
  static constexpr int NB = HAS_DS ? 3 : 5;                        // rotating staging buffers (residual in / tile out)
  static constexpr int X_BYTES = HAS_DS ? kStageOutBytes : 0;
  static constexpr int B_ITEMS = N3 / 64;                          // 64-channel groups of the block output
  static constexpr int C_ITEMS = N1 / 64;                          // 64-channel groups of the next conv1's output
  static constexpr int ITEMS = B_ITEMS + C_ITEMS;                  // staged groups (= TMA stores) per tile
  static constexpr int TAIL_BYTES = 1024 + 3072;                   // barriers + biases (C2 + N3 + N1 floats)
  static constexpr int SMEM_MAX = 232448;
  static constexpr int FIXED_BYTES = W3_BYTES + NP * PATCH_BYTES + X_BYTES + NB * kStageOutBytes + TAIL_BYTES + 1024;
  static constexpr int RING_RAW = (SMEM_MAX - FIXED_BYTES) / RING_STAGE;
  static constexpr int RING_D = RING_RAW > 9 ? 9 : RING_RAW;
  static constexpr int SMEM_BYTES = FIXED_BYTES + RING_D * RING_STAGE;
  static constexpr int TMEM_COLS = 512;
  static constexpr int ACC2_COL = 0, T2_COL = C2, ACC3_COL = (C2 == 64) ? 128 : 256, ACC1_COL = 384;
};

template <int W, int C2, int N3, bool HAS_DS, int N1>
__global__ void __launch_bounds__(kChainThreads, 1)
bottleneck_chain_kernel(const __grid_constant__ CUtensorMap mapH,   // t1 [n][W][W][C2], box {64, W+2, RT+2, 1}
                        const __grid_constant__ CUtensorMap mapX,   // HAS_DS: x [n][W][W][64], box {64, W+2, RT, 1}
                        const __grid_constant__ CUtensorMap mapW2,  // [C2][9*C2], box {64, C2}
                        const __grid_constant__ CUtensorMap mapW3,  // [N3][C2 (+64)], box {64, 256 (resident) | 128}
                        const __grid_constant__ CUtensorMap mapW1,  // N1 > 0: [N1][256], box {64, 64}
                        const __grid_constant__ CUtensorMap mapO,   // out [n][W][W][N3], box {64, W+2, RT, 1}
                        const __grid_constant__ CUtensorMap mapR,   // !HAS_DS: identity residual, same geometry as mapO
                        const __grid_constant__ CUtensorMap mapT,   // N1 > 0: t1' [n][W][W][N1], box {64, W+2, RT, 1}
                        const ChainParams p) {
  using Cfg = ChainCfg<W, C2, N3, HAS_DS, N1>;
  constexpr int D = Cfg::RING_D;
  constexpr int ITEMS = Cfg::ITEMS;
  constexpr int B_ITEMS = Cfg::B_ITEMS;
  constexpr int NB = Cfg::NB;
  constexpr int NP = Cfg::NP;
  constexpr int CB = Cfg::CB;
  constexpr int NH = Cfg::NH;
  constexpr int WP = Cfg::WP;
  constexpr int RT = Cfg::RT;
  static_assert((W == 56 && C2 == 64 && N3 == 256) || (W == 28 && C2 == 128 && N3 == 512), "layer1 / layer2 geometry");
  static_assert(D >= 4, "weight ring too shallow");
  static_assert(Cfg::SMEM_BYTES <= Cfg::SMEM_MAX, "shared memory budget exceeded");
  static_assert(N1 == 0 || N1 == 64 || N1 == 128, "next conv1 width");
  static_assert(!(HAS_DS && N1 > 64), "the down-sample variant is only paired with a 64-wide next conv1");
  static_assert(Cfg::W3_RES || (!HAS_DS && N1 == 0 && CB == NH), "streamed-conv3 variant: identity residual, no conv1n");
  static_assert(Cfg::TILE_ROWS <= kBlockM, "tile does not fit one M = 128 MMA");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w3 = smem;                                  // [W3_KB][256][128 B]
  uint8_t* s_ring = s_w3 + Cfg::W3_BYTES;                // [D][RING_STAGE]
  uint8_t* s_patch = s_ring + D * Cfg::RING_STAGE;       // [NP][PATCH_BYTES]; shifted windows over-read into what follows
  uint8_t* s_x = s_patch + NP * Cfg::PATCH_BYTES;        // HAS_DS: [128][128 B] down-sample source tile
  uint8_t* s_stage = s_x + Cfg::X_BYTES;                 // [NB][128][128 B]
  uint8_t* tail = s_stage + NB * kStageOutBytes;
  uint64_t* ring_full = reinterpret_cast<uint64_t*>(tail);  // [D]
  uint64_t* ring_empty = ring_full + D;                     // [D]
  uint64_t* patch_full = ring_empty + D;                    // [NP]
  uint64_t* patch_empty = patch_full + NP;                  // [NP]
  uint64_t* x_full = patch_empty + NP;                      // [1]
  uint64_t* x_empty = x_full + 1;                           // [1]
  uint64_t* w3_full = x_empty + 1;                          // [1]
  uint64_t* acc2_full = w3_full + 1;                        // [1] MMA -> epilogue
  uint64_t* acc3_full = acc2_full + 1;                      // [1] MMA -> epilogue, NH phases per tile
  uint64_t* acc3_empty = acc3_full + 1;                     // [1] epilogue -> MMA (NH == 2): the low half is drained
  uint64_t* acc1_full = acc3_empty + 1;
  uint64_t* t2_full = acc1_full + 1;                        // [1] epilogue (8 warps) -> MMA: bf16 t2 is in TMEM
  uint64_t* out_full = t2_full + 1;                         // [4] epilogue -> MMA: 64-channel group g of the bf16 block
                                                            // output is in TMEM (K block g of the next conv1)
  uint64_t* res_full = out_full + 4;                        // [NB] loader -> epilogue: buffer free / residual landed
  uint64_t* st_ready = res_full + NB;                       // [NB] epilogue -> DMA: group staged
  uint64_t* st_free = st_ready + NB;                        // [NB] DMA -> loader: the store has left shared memory
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(st_free + NB);
  float* s_b2 = reinterpret_cast<float*>(tail + 1024);      // [C2]
  float* s_b3 = s_b2 + C2;                                  // [N3]
  float* s_b1 = s_b3 + N3;                                  // [N1]

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_tiles;
  const int my_tiles = (num_tiles > static_cast<int>(blockIdx.x))
                           ? (num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                 static_cast<int>(gridDim.x)
                           : 0;
  // k-th tile of this CTA -> (frame, first output row)
  auto tile_of = [&](int k, int& n, int& r0) {
    const int lt = blockIdx.x + k * gridDim.x;
    const int t = p.rev ? num_tiles - 1 - lt : lt;
    n = t / Cfg::TILES_PER_FRAME;
    r0 = (t - n * Cfg::TILES_PER_FRAME) * RT;
  };
  // debug timeline: event e of this CTA's k-th tile (CTA 0 only, first 32 tiles)
  auto mark = [&](int k, int e) {
    if (p.trace != nullptr && blockIdx.x == 0 && k < 32 && lane == 0) p.trace[k * 32 + e] = clock64();
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapW2);
    tma_prefetch_desc(&mapW3);
    if (N1 > 0) tma_prefetch_desc(&mapW1);
  }
  if (warp == 3 && lane == 0) {
    tma_prefetch_desc(&mapH);
    if (HAS_DS) tma_prefetch_desc(&mapX);
  }
  if (warp == 2 && lane == 0) {
    tma_prefetch_desc(&mapO);
    if (N1 > 0) tma_prefetch_desc(&mapT);
  }
  if (warp == 12 && lane == 0 && !HAS_DS) tma_prefetch_desc(&mapR);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < D; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    for (int i = 0; i < NP; ++i) {
      mbar_init(&patch_full[i], 1);
      mbar_init(&patch_empty[i], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(w3_full, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc3_full, 1);
    mbar_init(acc3_empty, kEpiWarps);
    mbar_init(acc1_full, 1);
    mbar_init(t2_full, kEpiWarps);
    for (int i = 0; i < 4; ++i) mbar_init(&out_full[i], kEpiWarps);
    for (int i = 0; i < NB; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&st_ready[i], kEpiWarps);
      mbar_init(&st_free[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 4 && warp < 12) {  // biases are weights, not activations: safe to read before the grid dependency resolves
    const int et = threadIdx.x - 128;
    if (et < C2) s_b2[et] = __ldg(&p.bias2[et]);
    for (int i = et; i < N3; i += kEpiThreads) s_b3[i] = __ldg(&p.bias3[i]);
    if (N1 > 0 && et < N1) s_b1[et] = __ldg(&p.bias1n[et]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();
  griddep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ weight producer
    if (Cfg::W3_RES) {
      mbar_arrive_expect_tx_elect(w3_full, Cfg::W3_BYTES);
      for (int kb = 0; kb < Cfg::W3_KB; ++kb) tma_load_2d_elect(&mapW3, w3_full, s_w3 + kb * 256 * 128, kb * 64, 0);
    }
    int stage = 0;
    uint32_t phase = 0;
    auto push = [&](const CUtensorMap* m, int k0, int row0) {
      mbar_wait(&ring_empty[stage], phase ^ 1);
      mbar_arrive_expect_tx_elect(&ring_full[stage], Cfg::RING_STAGE);
      tma_load_2d_elect(m, &ring_full[stage], s_ring + stage * Cfg::RING_STAGE, k0, row0);
      if (++stage == D) {
        stage = 0;
        phase ^= 1;
      }
    };
    auto push_conv2 = [&](int cb_begin, int cb_end) {  // K order of the packed weights: (tap, cb, channel)
      for (int cb = cb_begin; cb < cb_end; ++cb)
        for (int tap = 0; tap < 9; ++tap) push(&mapW2, (tap * CB + cb) * 64, 0);
    };
    auto push_conv3 = [&](int h) {  // streamed conv3: output quarter (2h + nq) x K block kb
      for (int nq = 0; nq < 2; ++nq)
        for (int kb = 0; kb < CB; ++kb) push(&mapW3, kb * 64, (2 * h + nq) * 128);
    };
    // same order as the MMA warp consumes
    if (my_tiles > 0) push_conv2(0, CB);
    for (int k = 0; k < my_tiles; ++k) {
      const bool more = k + 1 < my_tiles;
      if (!Cfg::W3_RES) push_conv3(0);
      if (NH == 1) {
        if (more) push_conv2(0, CB);
      } else {
        if (more) push_conv2(0, 1);
        push_conv3(1);
        if (more) push_conv2(1, CB);
      }
      if (N1 > 0)
        for (int kb = 0; kb < 4; ++kb)
          for (int nh = 0; nh < N1 / 64; ++nh) push(&mapW1, kb * 64, nh * 64);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ activation producer
    int q = 0;  // running patch number
    for (int k = 0; k < my_tiles; ++k) {
      int n, r0;
      tile_of(k, n, r0);
      for (int cb = 0; cb < CB; ++cb, ++q) {
        const int slot = q % NP;
        mbar_wait(&patch_empty[slot], ((q / NP) & 1) ^ 1);
        if (cb == 0) mark(k, 15);
        mbar_arrive_expect_tx_elect(&patch_full[slot], Cfg::PATCH_LOAD_BYTES);
        tma_load_4d_elect(&mapH, &patch_full[slot], s_patch + slot * Cfg::PATCH_BYTES, cb * 64, -1, r0 - 1, n);
      }
      if (HAS_DS) {
        mbar_wait(x_empty, (k & 1) ^ 1);
        mbar_arrive_expect_tx_elect(x_full, Cfg::TILE_BYTES);
        tma_load_4d_elect(&mapX, x_full, s_x, 0, 0, r0, n);
      }
    }
  } else if (warp == 12) {
    // ------------------------------------------------------------------ residual loader / staging-buffer recycler
    int j = 0;
    for (int k = 0; k < my_tiles; ++k) {
      int n, r0;
      tile_of(k, n, r0);
      for (int it = 0; it < ITEMS; ++it, ++j) {
        const int b = j % NB;
        const int use = j / NB;
        if (use > 0) mbar_wait(&st_free[b], (use - 1) & 1);  // the previous store out of this buffer has been read
        if (!HAS_DS && it < B_ITEMS) {
          if (it == 0) mark(k, 22);
          mbar_arrive_expect_tx_elect(&res_full[b], Cfg::TILE_BYTES);
          tma_load_4d_elect(&mapR, &res_full[b], s_stage + b * kStageOutBytes, it * 64, 0, r0, n);
        } else {
          if (elect_one()) mbar_arrive(&res_full[b]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc2 = make_idesc_bf16(kBlockM, C2);
    constexpr uint32_t idesc3r = make_idesc_bf16(kBlockM, 256);  // resident conv3: one N = 256 MMA per K step
    constexpr uint32_t idesc3s = make_idesc_bf16(kBlockM, 128);  // streamed conv3: N = 128 quarters
    constexpr uint32_t idesc1 = make_idesc_bf16(kBlockM, 64);
    int stage = 0;
    uint32_t phase = 0;
    int q = 0;  // running patch number
    auto ring_next = [&]() {
      if (++stage == D) {
        stage = 0;
        phase ^= 1;
      }
    };
    // conv2 of tile k, 64-channel input blocks [cb_begin, cb_end)
    auto conv2 = [&](int k, int cb_begin, int cb_end) {
      for (int cb = cb_begin; cb < cb_end; ++cb, ++q) {
        const int slot = q % NP;
        mbar_wait(&patch_full[slot], (q / NP) & 1);
        if (cb == 0) mark(k, 0);
        const uint32_t patch = smem_u32(s_patch + slot * Cfg::PATCH_BYTES);
        for (int tap = 0; tap < 9; ++tap) {
          mbar_wait(&ring_full[stage], phase);
          tc_fence_after();
          const int r = tap / 3;
          const uint32_t a_addr = patch + (r * WP + (tap - r * 3)) * 128;  // row-shifted window of the patch
          const uint32_t b_addr = smem_u32(s_ring + stage * Cfg::RING_STAGE);
          umma_bf16_x4_elect(tmem_base + Cfg::ACC2_COL, make_kmajor_desc(a_addr, 128), make_kmajor_desc(b_addr, 128),
                             idesc2, (cb | tap) != 0 ? 1u : 0u);
          umma_commit_elect(&ring_empty[stage]);
          ring_next();
        }
        umma_commit_elect(&patch_empty[slot]);
      }
      if (cb_end == CB) {
        umma_commit_elect(acc2_full);
        mark(k, 1);
      }
    };
    // conv3 of tile k, output half h: A = bf16 t2 parked in TMEM by the epilogue warps
    auto conv3 = [&](int k, int h) {
      if (Cfg::W3_RES) {
        const uint32_t w3_addr = smem_u32(s_w3);
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_bf16_ts_elect(tmem_base + Cfg::ACC3_COL, tmem_base + Cfg::T2_COL + 8 * s,
                             make_kmajor_desc(w3_addr + 32 * s, 128), idesc3r, s != 0 ? 1u : 0u);
        if (HAS_DS) {  // + the down-sample source tile from smem
          mbar_wait(x_full, k & 1);
          tc_fence_after();
          umma_bf16_x4_elect(tmem_base + Cfg::ACC3_COL, make_kmajor_desc(smem_u32(s_x), 128),
                             make_kmajor_desc(w3_addr + 256 * 128, 128), idesc3r, 1u);
          umma_commit_elect(x_empty);
        }
      } else {
        for (int nq = 0; nq < 2; ++nq) {
          for (int kb = 0; kb < CB; ++kb) {
            mbar_wait(&ring_full[stage], phase);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(s_ring + stage * Cfg::RING_STAGE);
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma_bf16_ts_elect(tmem_base + Cfg::ACC3_COL + nq * 128, tmem_base + Cfg::T2_COL + kb * 32 + 8 * s,
                                 make_kmajor_desc(b_addr + 32 * s, 128), idesc3s, (kb | s) != 0 ? 1u : 0u);
            umma_commit_elect(&ring_empty[stage]);
            ring_next();
          }
        }
      }
      umma_commit_elect(acc3_full);
    };
    if (my_tiles > 0) {
      if (Cfg::W3_RES) mbar_wait(w3_full, 0);
      conv2(0, 0, CB);
    }
    for (int k = 0; k < my_tiles; ++k) {
      const bool more = k + 1 < my_tiles;
      mbar_wait(t2_full, k & 1);
      mark(k, 2);
      tc_fence_after();
      conv3(k, 0);
      if (NH == 1) {
        if (more) conv2(k + 1, 0, CB);
      } else {
        if (more) conv2(k + 1, 0, 1);
        mbar_wait(acc3_empty, k & 1);  // the low half has been drained
        tc_fence_after();
        conv3(k, 1);
        if (more) conv2(k + 1, 1, CB);
      }
      if (N1 > 0) {
        // conv1n(k): A = bf16 block output parked over the conv3 accumulator columns (K block kb, K step s at column
        // 64*kb + 32*(s/2) + 8*(s%2)), B streams through the ring as [64 rows][64 K] tiles.  Each K block is issued as
        // soon as the epilogue warps have parked ITS 64 channels, so the conv runs under the residual epilogue of the
        // groups behind it instead of after the whole of it (the epilogue warps used to idle ~1 300 cycles per tile here)
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(&out_full[kb], k & 1);
          if (kb == 0) mark(k, 3);
          tc_fence_after();
          for (int nh = 0; nh < N1 / 64; ++nh) {
            mbar_wait(&ring_full[stage], phase);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(s_ring + stage * Cfg::RING_STAGE);
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma_bf16_ts_elect(tmem_base + Cfg::ACC1_COL + nh * 64,
                                 tmem_base + Cfg::ACC3_COL + 64 * kb + 32 * (s >> 1) + 8 * (s & 1),
                                 make_kmajor_desc(b_addr + 32 * s, 128), idesc1, (kb | s) != 0 ? 1u : 0u);
            umma_commit_elect(&ring_empty[stage]);
            ring_next();
          }
        }
        umma_commit_elect(acc1_full);
        mark(k, 4);
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ store DMA (one thread)
    if (lane == 0) {
      int j = 0;
      for (int k = 0; k < my_tiles; ++k) {
        int n, r0;
        tile_of(k, n, r0);
        for (int it = 0; it < ITEMS; ++it, ++j) {
          const int b = j % NB;
          mbar_wait(&st_ready[b], (j / NB) & 1);
          const uint8_t* src = s_stage + b * kStageOutBytes;
          if (it < B_ITEMS)
            tma_store_4d(&mapO, src, it * 64, 0, r0, n);
          else
            tma_store_4d(&mapT, src, (it - B_ITEMS) * 64, 0, r0, n);
          tma_store_commit();
          if (it < 8) mark(k, 16 + it);
          if (j > 0) {  // one store stays in flight behind the newest; the one before has left shared memory
            tma_store_wait_read<1>();
            mbar_arrive(&st_free[(j - 1) % NB]);
          }
        }
      }
      tma_store_wait_all<0>();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (warps 4..11)
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int sw = row & 7;

    // relu(acc + bias [+ residual]) of 32 accumulator columns, rounded once to bf16: 16 packed words
    auto convert = [&](const uint32_t (&v)[32], const float* bias32, const uint8_t* res_row, uint32_t (&o)[16]) {
      const float4* b4 = reinterpret_cast<const float4*>(bias32);
      const __nv_bfloat162 z = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 b0 = b4[2 * c4], b1 = b4[2 * c4 + 1];
        float f[8];
        f[0] = __uint_as_float(v[8 * c4 + 0]) + b0.x;
        f[1] = __uint_as_float(v[8 * c4 + 1]) + b0.y;
        f[2] = __uint_as_float(v[8 * c4 + 2]) + b0.z;
        f[3] = __uint_as_float(v[8 * c4 + 3]) + b0.w;
        f[4] = __uint_as_float(v[8 * c4 + 4]) + b1.x;
        f[5] = __uint_as_float(v[8 * c4 + 5]) + b1.y;
        f[6] = __uint_as_float(v[8 * c4 + 6]) + b1.z;
        f[7] = __uint_as_float(v[8 * c4 + 7]) + b1.w;
        if (res_row != nullptr) {
          const uint4 rv = *reinterpret_cast<const uint4*>(res_row + (((half * 4 + c4) ^ sw) << 4));
          f[0] += bf16_lo(rv.x);
          f[1] += bf16_hi(rv.x);
          f[2] += bf16_lo(rv.y);
          f[3] += bf16_hi(rv.y);
          f[4] += bf16_lo(rv.z);
          f[5] += bf16_hi(rv.z);
          f[6] += bf16_lo(rv.w);
          f[7] += bf16_hi(rv.w);
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          __nv_bfloat162 h2 = __hmax2(__floats2bfloat162_rn(f[2 * qq], f[2 * qq + 1]), z);
          o[4 * c4 + qq] = *reinterpret_cast<uint32_t*>(&h2);
        }
      }
    };
    // this thread's 64 B of a 128B-swizzled staging row
    auto stage_row = [&](const uint32_t (&o)[16], uint8_t* row_ptr) {
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
        *reinterpret_cast<uint4*>(row_ptr + (((half * 4 + c4) ^ sw) << 4)) =
            make_uint4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
    };

    auto epi_A = [&](int k) {  // conv2: t2 = relu(acc2 + b2) -> bf16 in TMEM (conv3's A operand)
      mbar_wait(acc2_full, k & 1);
      if (warp == 4) mark(k, 5);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < CB; ++g) {
        uint32_t v[32], o[16];
        tmem_ld_32x32b_x32(t_lane + Cfg::ACC2_COL + g * 64 + half * 32, v);
        tmem_ld_wait();
        convert(v, s_b2 + g * 64 + half * 32, nullptr, o);
        tmem_st_32x32b_x16(t_lane + Cfg::T2_COL + g * 32 + half * 16, o);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t2_full);
      if (warp == 4) mark(k, 6);
    };
    // conv3, output half h: out = relu(acc3 + b3 [+ identity]) -> staged for the TMA store (+ bf16 in TMEM)
    auto epi_B = [&](int k, int h) {
      mbar_wait(acc3_full, (k * NH + h) & 1);
      if (warp == 4) mark(k, h == 0 ? 7 : 14);
      tc_fence_after();
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        const int g = h * 4 + gg;          // 64-channel group of the block output
        const int j = k * ITEMS + g;       // staged-group number (matches the loader's and the DMA thread's)
        const int b = j % NB;
        uint32_t v[32], o[16];
        tmem_ld_32x32b_x32(t_lane + Cfg::ACC3_COL + gg * 64 + half * 32, v);
        mbar_wait(&res_full[b], (j / NB) & 1);  // buffer free (and the residual tile in it)
        uint8_t* row_ptr = s_stage + b * kStageOutBytes + row * 128;
        tmem_ld_wait();
        convert(v, s_b3 + g * 64 + half * 32, HAS_DS ? nullptr : row_ptr, o);
        stage_row(o, row_ptr);
        // K block g of the next conv1's A operand: bf16 pairs over the first 16 of the 32 columns just drained
        if (N1 > 0) tmem_st_32x32b_x16(t_lane + Cfg::ACC3_COL + gg * 64 + half * 32, o);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_ready[b]);
        if (N1 > 0) {  // K block gg of the next conv1 may go
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&out_full[gg]);
        }
        if (warp == 4 && h == 0) mark(k, 8 + gg);
      }
      if (NH == 2 && h == 0) {  // the MMA warp may overwrite the accumulator with the high half
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc3_empty);
      }
    };
    auto epi_C = [&](int k) {  // next conv1: t1' = relu(acc1 + b1) -> staged for the TMA store
      mbar_wait(acc1_full, k & 1);
      if (warp == 4) mark(k, 12);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < Cfg::C_ITEMS; ++g) {
        const int j = k * ITEMS + B_ITEMS + g;
        const int b = j % NB;
        uint32_t v[32], o[16];
        tmem_ld_32x32b_x32(t_lane + Cfg::ACC1_COL + g * 64 + half * 32, v);
        mbar_wait(&res_full[b], (j / NB) & 1);
        tmem_ld_wait();
        convert(v, s_b1 + g * 64 + half * 32, nullptr, o);
        stage_row(o, s_stage + b * kStageOutBytes + row * 128);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_ready[b]);
      }
      tc_fence_before();
      if (warp == 4) mark(k, 13);
    };

    if (my_tiles > 0) epi_A(0);
    for (int k = 0; k < my_tiles; ++k) {
      epi_B(k, 0);
      if (NH == 2) epi_B(k, 1);
      if (k + 1 < my_tiles) epi_A(k + 1);
      if (N1 > 0) epi_C(k);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace phdfxk
