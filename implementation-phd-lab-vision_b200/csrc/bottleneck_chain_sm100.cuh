// K3c: layer1 "bottleneck chain" — conv2 (3x3) -> conv3 (1x1, + identity or fused down-sample) -> ReLU
//      [-> conv1 of the NEXT block (1x1)] in ONE kernel, for the 56x56 stage (width 64, 256 output channels).
//
// Replaces, per block of torchvision's layer1 (models/resnet.py:150-161 and the following block's :146-148), three
// launches of conv_igemm_kernel and two round trips through HBM:
//   * t2 (conv2's 64-channel output) never leaves the SM: the epilogue warps round it to bf16 and park it in TENSOR
//     MEMORY (tcgen05.st, two channels per 32-bit column); conv3 takes it from there as its A operand
//     (tcgen05.mma with A in TMEM) — no shared-memory traffic at all for t2;
//   * the block output (256 channels, the largest activation of the network) is stored once (TMA, through rotating
//     staging buffers) and ALSO written back, as bf16, over the accumulator columns it was computed from; from there it
//     is the A operand (K = 256) of the next block's 1x1 reduce conv — that conv no longer re-reads 1.6 MB per frame
//     from HBM, and its A operand costs no shared-memory bandwidth either;
//   * the residual tile is TMA-loaded into the staging buffer ahead of the epilogue and added in place (as in
//     conv_igemm_kernel), by a loader warp that runs as far ahead as free staging buffers allow.
// layer1 is HBM-bound when unfused (conv3 + residual moves 3.6 MB per frame for 51 MMAC); the chain moves, per block,
// t1 in (0.4 MB, halo rows from L2), residual in + out (1.6 MB each) and the next t1 out (0.4 MB).
//
// Tile = 2 output rows of one frame in "padded raster" order (row pitch 58): GEMM row m = i*58 + j, valid when
// i < 2 and j < 56 (116 of 128 rows carry pixels, 112 are valid).  Every GEMM of the chain keeps that row order, rows
// are independent in all of them, and the TMA stores clip the two pad columns — so the invalid rows never need masking.
//
// Warps (13): 0 weight producer (conv2 taps / next-conv1 tiles through an mbarrier ring; conv3's weights resident),
// 1 MMA issuer, 2 store DMA (+ TMEM owner), 3 activation producer (input patches, down-sample source), 4-11 epilogue,
// 12 residual loader.  The MMA warp software-pipelines across tiles:   conv3(k) | conv2(k+1) | conv1n(k)   while the
// epilogue warps run  B(k) | A(k+1) | C(k)  (A = conv2's, B = conv3's, C = next-conv1's epilogue), so the long conv2
// overlaps the long residual epilogue.  Tensor memory: conv2 accumulator 64 columns, t2 (bf16) 32, conv3 accumulator
// 256 (its first 16 columns of every 32 are re-used for the bf16 block output), next-conv1 accumulator N1; all
// single-buffered — the data dependencies of the chain and the in-order tensor pipe order every reuse.
#pragma once
#include "conv_igemm_sm100.cuh"

namespace phdfxk {

struct ChainParams {
  int n_frames;
  int num_tiles;        // n_frames * 28
  int rev;              // walk tiles in descending order (see ConvParams::rev)
  const float* bias2;   // [64]  conv2 folded-BN bias
  const float* bias3;   // [256] conv3 (+ down-sample) bias
  const float* bias1n;  // [N1]  next block's conv1 bias (N1 > 0)
  long long* trace;     // debug (PHDFX_CHAIN_TRACE): CTA 0 writes clock64() of pipeline events, [tile < 32][32 events]
};

constexpr int kChW = 56, kChWP = 58, kChRT = 2, kChTilesPerFrame = kChW / kChRT;
constexpr int kChainThreads = 416;  // 13 warps

template <bool HAS_DS, int N1>
struct ChainCfg {
  static constexpr int TILE_ROWS = kChRT * kChWP;                  // 116 padded-raster rows hold output pixels
  static constexpr int TILE_BYTES = TILE_ROWS * 128;               // one 64-channel group of a tile = one TMA box
  static constexpr int HALO_BYTES = (kChRT + 2) * kChWP * 128;     // 29696 = 29 * 1024: 4 x 58 positions x 64 ch
  static constexpr int W3_KB = HAS_DS ? 2 : 1;                     // K blocks of conv3 (t2 | x for the down-sample)
  static constexpr int W3_BYTES = W3_KB * 256 * 128;
  static constexpr int RING_STAGE = 8192;                          // one [64 rows][64 K] weight tile
  static constexpr int NB = HAS_DS ? 3 : 5;                        // rotating staging buffers (residual in / tile out)
  static constexpr int X_BYTES = HAS_DS ? kStageOutBytes : 0;
  static constexpr int C_ITEMS = N1 / 64;                          // 64-channel groups of the next conv1's output
  static constexpr int ITEMS = 4 + C_ITEMS;                        // staged 64-channel groups (= TMA stores) per tile
  static constexpr int TAIL_BYTES = 1024 + 2048;                   // barriers + biases (64 + 256 + 128 floats)
  static constexpr int SMEM_MAX = 232448;
  static constexpr int FIXED_BYTES = W3_BYTES + 2 * HALO_BYTES + X_BYTES + NB * kStageOutBytes + TAIL_BYTES + 1024;
  static constexpr int RING_RAW = (SMEM_MAX - FIXED_BYTES) / RING_STAGE;
  static constexpr int RING_D = RING_RAW > 9 ? 9 : RING_RAW;
  static constexpr int SMEM_BYTES = FIXED_BYTES + RING_D * RING_STAGE;
  static constexpr int TMEM_COLS = 512;
  static constexpr int ACC2_COL = 0, T2_COL = 64, ACC3_COL = 128, ACC1_COL = 384;
};

template <bool HAS_DS, int N1>
__global__ void __launch_bounds__(kChainThreads, 1)
bottleneck_chain_kernel(const __grid_constant__ CUtensorMap mapH,   // t1 [n][56][56][64], box {64, 58, 4, 1}
                        const __grid_constant__ CUtensorMap mapX,   // HAS_DS: x [n][56][56][64], box {64, 58, 2, 1}
                        const __grid_constant__ CUtensorMap mapW2,  // [64][576], box {64, 64}
                        const __grid_constant__ CUtensorMap mapW3,  // [256][64 | 128], box {64, 256}
                        const __grid_constant__ CUtensorMap mapW1,  // N1 > 0: [N1][256], box {64, 64}
                        const __grid_constant__ CUtensorMap mapO,   // out [n][56][56][256], box {64, 58, 2, 1}
                        const __grid_constant__ CUtensorMap mapR,   // !HAS_DS: identity residual, same geometry as mapO
                        const __grid_constant__ CUtensorMap mapT,   // N1 > 0: t1' [n][56][56][N1], box {64, 58, 2, 1}
                        const ChainParams p) {
  using Cfg = ChainCfg<HAS_DS, N1>;
  constexpr int D = Cfg::RING_D;
  constexpr int ITEMS = Cfg::ITEMS;
  constexpr int NB = Cfg::NB;
  static_assert(D >= 4, "weight ring too shallow");
  static_assert(Cfg::SMEM_BYTES <= Cfg::SMEM_MAX, "shared memory budget exceeded");
  static_assert(N1 == 0 || N1 == 64 || N1 == 128, "next conv1 width");
  static_assert(!(HAS_DS && N1 > 64), "the down-sample variant is only paired with a 64-wide next conv1");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_w3 = smem;                                  // [W3_KB][256][128 B]
  uint8_t* s_ring = s_w3 + Cfg::W3_BYTES;                // [D][RING_STAGE]
  uint8_t* s_halo = s_ring + D * Cfg::RING_STAGE;        // [2][HALO_BYTES]; shifted windows over-read into what follows
  uint8_t* s_x = s_halo + 2 * Cfg::HALO_BYTES;           // HAS_DS: [128][128 B] down-sample source tile
  uint8_t* s_stage = s_x + Cfg::X_BYTES;                 // [NB][128][128 B]
  uint8_t* tail = s_stage + NB * kStageOutBytes;
  uint64_t* ring_full = reinterpret_cast<uint64_t*>(tail);  // [D]
  uint64_t* ring_empty = ring_full + D;                     // [D]
  uint64_t* halo_full = ring_empty + D;                     // [2]
  uint64_t* halo_empty = halo_full + 2;                     // [2]
  uint64_t* x_full = halo_empty + 2;                        // [1]
  uint64_t* x_empty = x_full + 1;                           // [1]
  uint64_t* w3_full = x_empty + 1;                          // [1]
  uint64_t* acc2_full = w3_full + 1;                        // [1] MMA -> epilogue
  uint64_t* acc3_full = acc2_full + 1;
  uint64_t* acc1_full = acc3_full + 1;
  uint64_t* t2_full = acc1_full + 1;                        // [1] epilogue (8 warps) -> MMA: bf16 t2 is in TMEM
  uint64_t* out_full = t2_full + 1;                         // [1] epilogue -> MMA: bf16 block output is in TMEM
  uint64_t* res_full = out_full + 1;                        // [NB] loader -> epilogue: buffer free / residual landed
  uint64_t* st_ready = res_full + NB;                       // [NB] epilogue -> DMA: group staged
  uint64_t* st_free = st_ready + NB;                        // [NB] DMA -> loader: the store has left shared memory
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(st_free + NB);
  float* s_b2 = reinterpret_cast<float*>(tail + 1024);      // [64]
  float* s_b3 = s_b2 + 64;                                  // [256]
  float* s_b1 = s_b3 + 256;                                 // [128]

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
    n = t / kChTilesPerFrame;
    r0 = (t - n * kChTilesPerFrame) * kChRT;
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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&halo_full[i], 1);
      mbar_init(&halo_empty[i], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_empty, 1);
    mbar_init(w3_full, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc3_full, 1);
    mbar_init(acc1_full, 1);
    mbar_init(t2_full, kEpiWarps);
    mbar_init(out_full, kEpiWarps);
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
    if (et < 64) s_b2[et] = __ldg(&p.bias2[et]);
    s_b3[et] = __ldg(&p.bias3[et]);
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
    mbar_arrive_expect_tx_elect(w3_full, Cfg::W3_BYTES);
    for (int kb = 0; kb < Cfg::W3_KB; ++kb) tma_load_2d_elect(&mapW3, w3_full, s_w3 + kb * 256 * 128, kb * 64, 0);
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
    // same order as the MMA warp consumes: conv2(0) | { conv2(k+1) | conv1n(k) }
    if (my_tiles > 0)
      for (int tap = 0; tap < 9; ++tap) push(&mapW2, tap * 64, 0);
    for (int k = 0; k < my_tiles; ++k) {
      if (k + 1 < my_tiles)
        for (int tap = 0; tap < 9; ++tap) push(&mapW2, tap * 64, 0);
      if (N1 > 0)
        for (int kb = 0; kb < 4; ++kb)
          for (int nh = 0; nh < N1 / 64; ++nh) push(&mapW1, kb * 64, nh * 64);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ activation producer
    for (int k = 0; k < my_tiles; ++k) {
      int n, r0;
      tile_of(k, n, r0);
      const int hb = k & 1;
      mbar_wait(&halo_empty[hb], ((k >> 1) & 1) ^ 1);
      mark(k, 15);
      mbar_arrive_expect_tx_elect(&halo_full[hb], Cfg::HALO_BYTES);
      tma_load_4d_elect(&mapH, &halo_full[hb], s_halo + hb * Cfg::HALO_BYTES, 0, -1, r0 - 1, n);
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
        if (!HAS_DS && it < 4) {
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
    constexpr uint32_t idesc2 = make_idesc_bf16(kBlockM, 64);
    constexpr uint32_t idesc3 = make_idesc_bf16(kBlockM, 256);
    int stage = 0;
    uint32_t phase = 0;
    auto ring_next = [&]() {
      if (++stage == D) {
        stage = 0;
        phase ^= 1;
      }
    };
    auto conv2 = [&](int k) {
      const int hb = k & 1;
      mbar_wait(&halo_full[hb], (k >> 1) & 1);
      mark(k, 0);
      const uint32_t patch = smem_u32(s_halo + hb * Cfg::HALO_BYTES);
      for (int tap = 0; tap < 9; ++tap) {
        mbar_wait(&ring_full[stage], phase);
        tc_fence_after();
        const int r = tap / 3;
        const uint32_t a_addr = patch + (r * kChWP + (tap - r * 3)) * 128;  // row-shifted window of the patch
        const uint32_t b_addr = smem_u32(s_ring + stage * Cfg::RING_STAGE);
        umma_bf16_x4_elect(tmem_base + Cfg::ACC2_COL, make_kmajor_desc(a_addr, 128), make_kmajor_desc(b_addr, 128),
                           idesc2, tap != 0 ? 1u : 0u);
        umma_commit_elect(&ring_empty[stage]);
        ring_next();
      }
      umma_commit_elect(&halo_empty[hb]);
      umma_commit_elect(acc2_full);
      mark(k, 1);
    };
    if (my_tiles > 0) {
      mbar_wait(w3_full, 0);
      conv2(0);
    }
    for (int k = 0; k < my_tiles; ++k) {
      // conv3(k): A = bf16 t2 parked in TMEM by the epilogue warps (+ the down-sample source tile from smem),
      // B = resident W3
      mbar_wait(t2_full, k & 1);
      mark(k, 2);
      tc_fence_after();
      const uint32_t w3_addr = smem_u32(s_w3);
#pragma unroll
      for (int s = 0; s < 4; ++s)
        umma_bf16_ts_elect(tmem_base + Cfg::ACC3_COL, tmem_base + Cfg::T2_COL + 8 * s,
                           make_kmajor_desc(w3_addr + 32 * s, 128), idesc3, s != 0 ? 1u : 0u);
      if (HAS_DS) {
        mbar_wait(x_full, k & 1);
        tc_fence_after();
        umma_bf16_x4_elect(tmem_base + Cfg::ACC3_COL, make_kmajor_desc(smem_u32(s_x), 128),
                           make_kmajor_desc(w3_addr + 256 * 128, 128), idesc3, 1u);
        umma_commit_elect(x_empty);
      }
      umma_commit_elect(acc3_full);
      if (k + 1 < my_tiles) conv2(k + 1);
      if (N1 > 0) {
        // conv1n(k): A = bf16 block output parked over the conv3 accumulator columns (K block kb, K step s at column
        // 64*kb + 32*(s/2) + 8*(s%2)), B streams through the ring as [64 rows][64 K] tiles
        mbar_wait(out_full, k & 1);
        mark(k, 3);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) {
          for (int nh = 0; nh < N1 / 64; ++nh) {
            mbar_wait(&ring_full[stage], phase);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(s_ring + stage * Cfg::RING_STAGE);
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma_bf16_ts_elect(tmem_base + Cfg::ACC1_COL + nh * 64,
                                 tmem_base + Cfg::ACC3_COL + 64 * kb + 32 * (s >> 1) + 8 * (s & 1),
                                 make_kmajor_desc(b_addr + 32 * s, 128), idesc2, (kb | s) != 0 ? 1u : 0u);
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
          if (it < 4)
            tma_store_4d(&mapO, src, it * 64, 0, r0, n);
          else
            tma_store_4d(&mapT, src, (it - 4) * 64, 0, r0, n);
          tma_store_commit();
          mark(k, 16 + it);
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
    int j = 0;  // running staged-group counter (matches the loader's and the DMA thread's)

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
        for (int q = 0; q < 4; ++q) {
          __nv_bfloat162 h2 = __hmax2(__floats2bfloat162_rn(f[2 * q], f[2 * q + 1]), z);
          o[4 * c4 + q] = *reinterpret_cast<uint32_t*>(&h2);
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
      uint32_t v[32], o[16];
      tmem_ld_32x32b_x32(t_lane + Cfg::ACC2_COL + half * 32, v);
      tmem_ld_wait();
      convert(v, s_b2 + half * 32, nullptr, o);
      tmem_st_32x32b_x16(t_lane + Cfg::T2_COL + half * 16, o);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t2_full);
      if (warp == 4) mark(k, 6);
    };
    auto epi_B = [&](int k) {  // conv3: out = relu(acc3 + b3 [+ identity]) -> staged for the TMA store (+ bf16 in TMEM)
      mbar_wait(acc3_full, k & 1);
      if (warp == 4) mark(k, 7);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 4; ++g, ++j) {
        const int b = j % NB;
        uint32_t v[32], o[16];
        tmem_ld_32x32b_x32(t_lane + Cfg::ACC3_COL + g * 64 + half * 32, v);
        mbar_wait(&res_full[b], (j / NB) & 1);  // buffer free (and the residual tile in it)
        uint8_t* row_ptr = s_stage + b * kStageOutBytes + row * 128;
        tmem_ld_wait();
        convert(v, s_b3 + g * 64 + half * 32, HAS_DS ? nullptr : row_ptr, o);
        stage_row(o, row_ptr);
        // K block g of the next conv1's A operand: bf16 pairs over the first 16 of the 32 columns just drained
        if (N1 > 0) tmem_st_32x32b_x16(t_lane + Cfg::ACC3_COL + g * 64 + half * 32, o);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&st_ready[b]);
        if (warp == 4) mark(k, 8 + g);
      }
      if (N1 > 0) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(out_full);
      }
    };
    auto epi_C = [&](int k) {  // next conv1: t1' = relu(acc1 + b1) -> staged for the TMA store
      mbar_wait(acc1_full, k & 1);
      if (warp == 4) mark(k, 12);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < Cfg::C_ITEMS; ++g, ++j) {
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
      epi_B(k);
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
