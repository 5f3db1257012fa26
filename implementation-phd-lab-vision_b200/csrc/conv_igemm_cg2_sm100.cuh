// K3 (2-CTA variant): implicit-GEMM conv with tcgen05.mma.cta_group::2 for the K-heavy layers of layer3/layer4
// (1x1 reduce convs, 3x3 convs, strided 1x1 down-samples; Cout tile 256, no residual).
//
// Why: with one CTA per tile (conv_igemm_sm100.cuh) every SM ingests a 128-row activation tile AND the whole 256-row
// weight tile per K block: 48 KB for 512 tensor-core cycles = 96 B/clk, but an SM only receives ~48 B/clk from L2, so
// those layers sit at ~50-60 % tensor utilisation.  Here two SMs of a cluster form one 256 x 256 tile: each SM loads its
// own 128 activation rows and HALF of the weight tile (32 KB per K block = 64 B/clk); the pair's tensor cores read
// both halves (cta_group::2).  Work units per SM stay the same (one 128-row slab each), so wave quantisation does too.
//
// Roles per CTA (12 warps, as in the 1-CTA kernel): warp 0 TMA producer (both CTAs; completions land on the LEADER's
// full barrier), warp 1 MMA issuer (leader CTA only; commits multicast to both CTAs' barriers), warp 2 epilogue DMA +
// TMEM allocation (cta_group::2, both CTAs), warps 4-11 epilogue (each CTA drains its own 128 TMEM lanes; "accumulator
// free" arrives go to the leader's barrier).
#pragma once
#include "conv_igemm_sm100.cuh"

namespace phdfxk {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA loads of a CTA pair: data lands in the executing CTA's smem, the byte count is credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_cg2_elect(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_cg2_elect(const CUtensorMap* m, uint64_t* bar, void* dst, int c,
                                                             int w, int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5, %6}], [%2], {%7, %8};\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n),
      "h"(off_w), "h"(off_h)
      : "memory");
}
// four K=16 MMAs of one K block on the CTA pair (M = 256 over both CTAs), one election (leader CTA only)
__device__ __forceinline__ void umma_bf16_x4_cg2_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                       uint32_t accumulate_first) {
  asm volatile(
      "{\n"
      ".reg .pred p, pe, pt;\n"
      ".reg .b32 alo, ahi, blo, bhi;\n"
      ".reg .b64 da, db;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "setp.eq.b32 pt, 0, 0;\n"
      "mov.b64 {alo, ahi}, %1;\n"
      "mov.b64 {blo, bhi}, %2;\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "add.u32 alo, alo, 2;\n"
      "add.u32 blo, blo, 2;\n"
      "mov.b64 da, {alo, ahi};\n"
      "mov.b64 db, {blo, bhi};\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u32 alo, alo, 2;\n"
      "add.u32 blo, blo, 2;\n"
      "mov.b64 da, {alo, ahi};\n"
      "mov.b64 db, {blo, bhi};\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u32 alo, alo, 2;\n"
      "add.u32 blo, blo, 2;\n"
      "mov.b64 da, {alo, ahi};\n"
      "mov.b64 db, {blo, bhi};\n"
      "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, pt;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
// arrive (once all earlier MMAs of the pair retired) on the barrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_cg2_mc_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// arrive on the LEADER CTA's copy of a barrier
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, 0;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}

struct Cg2Cfg {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = kBlockM * 128;        // this CTA's 128 activation rows x 64 K
  static constexpr int B_BYTES = (BN / 2) * 128;       // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NB = 3;    // output staging buffers (3 x 16 KB leaves room for 5 pipeline stages)
  static constexpr int LOOK = 1;
  static constexpr int GROUPS = BN / kGroupCols;
  static constexpr int TAIL_BYTES = 1024 + 2 * BN * 4;
  static constexpr int NSTAGE = (232448 - 1024 - TAIL_BYTES - NB * kStageOutBytes) / STAGE_BYTES;  // = 5
  static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + NB * kStageOutBytes + TAIL_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BN;
};

// grid = 2 * pairs, cluster (2,1,1).  pair-tile pt -> (m_pair, n_blk); CTA rank r owns output rows of m_blk = 2*m_pair+r.
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
conv_igemm_cg2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                      const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapA2,
                      const ConvParams p) {
  using Cfg = Cg2Cfg;
  constexpr int BN = Cfg::BN;
  constexpr int NSTAGE = Cfg::NSTAGE;
  constexpr int NB = Cfg::NB;
  constexpr int LOOK = Cfg::LOOK;
  constexpr int GROUPS = Cfg::GROUPS;
  static_assert(NSTAGE == 5, "pipeline depth");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + NSTAGE * Cfg::STAGE_BYTES;
  uint8_t* tail = stage_out + NB * kStageOutBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);  // [NSTAGE] leader's copy is the live one
  uint64_t* empty_bar = full_bar + NSTAGE;                 // [NSTAGE] per CTA, fed by the leader's multicast commit
  uint64_t* tmem_full = empty_bar + NSTAGE;                // [2] per CTA, fed by the leader's multicast commit
  uint64_t* tmem_empty = tmem_full + 2;                    // [2] leader's copy: 2 CTAs x 8 epilogue warps
  uint64_t* res_full = tmem_empty + 2;                     // [NB]
  uint64_t* out_full = res_full + NB;                      // [NB]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(out_full + NB);
  float* s_bias = reinterpret_cast<float*>(tail + 1024);   // [2][BN]

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  if (P_CTA_TS(p) != nullptr && threadIdx.x == 0) P_CTA_TS(p)[blockIdx.x * 6 + 4] = global_timer_ns();  // kernel entry
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int num_ptiles = m_pairs * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 2 && lane == 0) tma_prefetch_desc(&mapO);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * kEpiWarps);
    }
    for (int i = 0; i < NB; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&out_full[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  cluster_sync_all();  // both CTAs' barriers are initialised before anything remote touches them
  if (warp == 2) {
    tmem_alloc_cg2(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();
  if (P_CTA_TS(p) != nullptr && threadIdx.x == 0) P_CTA_TS(p)[blockIdx.x * 6 + 0] = global_timer_ns();
  if (P_WAIT(p) == nullptr) griddep_wait();
  if (P_CTA_TS(p) != nullptr && threadIdx.x == 0 && P_WAIT(p) == nullptr) P_CTA_TS(p)[blockIdx.x * 6 + 1] = global_timer_ns();
  unsigned long long dep_ns = 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    bool dep_all = false;
    auto frames_of = [&](int mb, int* f_lo, int* f_hi) {  // empty range for the missing half of an odd last pair
      const int pq = p.P * p.Q;
      const int r1 = (mb + 1) * kBlockM < p.M ? (mb + 1) * kBlockM : p.M;
      *f_lo = (mb * kBlockM) / pq;
      *f_hi = mb * kBlockM < p.M ? (r1 - 1) / pq : *f_lo - 1;
    };
    for (int lp = pair; lp < num_ptiles; lp += num_pairs) {
      const int pt = p.rev ? num_ptiles - 1 - lp : lp;
      const int m_pair = pt / p.n_tiles;
      const int n_blk = pt - m_pair * p.n_tiles;
      const int m_blk = 2 * m_pair + static_cast<int>(rank);
      int cw = 0, ch = 0, cn = 0;
      if (MODE == MODE_IM2COL || p.src2_stride == 2) {
        const int m0 = m_blk * kBlockM;
        const int pq = p.P * p.Q;
        cn = m0 / pq;
        const int rem = m0 - cn * pq;
        const int p0 = rem / p.Q;
        const int q0 = rem - p0 * p.Q;
        if (MODE == MODE_IM2COL) {
          cw = q0 * p.stride - p.pad;
          ch = p0 * p.stride - p.pad;
        } else {  // fused stride-2 down-sample as the second source of a 1x1 conv
          cw = q0 * 2;
          ch = p0 * 2;
        }
      }
      if (P_WAIT(p) != nullptr) {  // frame progress counters: conv_igemm_sm100.cuh
        const unsigned long long tw = P_CTA_TS(p) != nullptr ? global_timer_ns() : 0ull;
        if (!dep_all) dep_all = dep_grid_done(P_WAIT(p) + p.ctr_frames, p.wait_ctas);
        if (!dep_all) {
          int f_lo, f_hi;
          frames_of(m_blk, &f_lo, &f_hi);
          dep_wait_frames(P_WAIT(p), p.wait_full, f_lo, f_hi);
        }
        if (P_CTA_TS(p) != nullptr) {
          const unsigned long long now = global_timer_ns();
          dep_ns += now - tw;
          if (lp == pair && lane == 0) P_CTA_TS(p)[blockIdx.x * 6 + 1] = now;
        }
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        // the leader arms its barrier with the bytes of BOTH CTAs
        if (leader) mbar_arrive_expect_tx_elect(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
        if (MODE == MODE_TILED) {
          if (kb < p.kb_split)
            tma_load_2d_cg2_elect(&mapA, &full_bar[stage], sA, kb * 64, m_blk * kBlockM);
          else if (p.src2_stride == 1)
            tma_load_2d_cg2_elect(&mapA2, &full_bar[stage], sA, (kb - p.kb_split) * 64, m_blk * kBlockM);
          else
            tma_load_im2col_4d_cg2_elect(&mapA2, &full_bar[stage], sA, (kb - p.kb_split) * 64, cw, ch, cn, 0, 0);
        } else {
          const int tap = kb / p.kb_per_tap;
          const int cb = kb - tap * p.kb_per_tap;
          const int r = tap / p.S;
          const int s = tap - r * p.S;
          tma_load_im2col_4d_cg2_elect(&mapA, &full_bar[stage], sA, cb * 64, cw, ch, cn, static_cast<uint16_t>(s),
                                       static_cast<uint16_t>(r));
        }
        tma_load_2d_cg2_elect(&mapB, &full_bar[stage], sB, kb * 64, n_blk * BN + static_cast<int>(rank) * (BN / 2));
        if (++stage == NSTAGE) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int pt = pair; pt < num_ptiles; pt += num_pairs) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        bool ready = false;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (!ready) mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
          int nstage = stage + 1;
          uint32_t nphase = phase;
          if (nstage == NSTAGE) {
            nstage = 0;
            nphase ^= 1;
          }
          ready = (kb + 1 < p.num_kb) ? mbar_test(&full_bar[nstage], nphase) : false;
          umma_bf16_x4_cg2_elect(d_tmem, make_kmajor_desc(a_addr, 128), make_kmajor_desc(b_addr, 128), idesc,
                                 kb != 0 ? 1u : 0u);
          umma_commit_cg2_mc_elect(&empty_bar[stage]);  // frees this stage in both CTAs
          stage = nstage;
          phase = nphase;
        }
        umma_commit_cg2_mc_elect(&tmem_full[acc]);  // accumulator complete, both CTAs' epilogues may read
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ epilogue DMA (one thread per CTA)
    if (lane == 0) {
      const int my_tiles = (num_ptiles > pair) ? (num_ptiles - pair + num_pairs - 1) / num_pairs : 0;
      const int J = my_tiles * GROUPS;
      // publish this CTA's half of the pair tile lp (its stores have completed)
      auto signal_tile = [&](int lp) {
        const int pt = p.rev ? num_ptiles - 1 - lp : lp;
        const int m_blk = 2 * (pt / p.n_tiles) + static_cast<int>(rank);
        if (m_blk * kBlockM >= p.M) return;
        const int r1 = (m_blk + 1) * kBlockM < p.M ? (m_blk + 1) * kBlockM : p.M;
        dep_signal_rows(P_SIG(p), m_blk * kBlockM, r1, p.P * p.Q, GROUPS);
      };
      for (int t = 0; t < J + LOOK; ++t) {
        if (t < J) {
          if (t >= NB) tma_store_wait_read<NB - LOOK - 1>();
          mbar_arrive(&res_full[t % NB]);  // no residual in this variant: the buffer is simply free
        }
        if (t >= LOOK) {
          const int u = t - LOOK;
          const int b = u % NB;
          mbar_wait(&out_full[b], (u / NB) & 1);
          const int lp = pair + (u / GROUPS) * num_pairs;
          const int pt = p.rev ? num_ptiles - 1 - lp : lp;
          const int g = u % GROUPS;
          const int m_pair = pt / p.n_tiles;
          const int n_blk = pt - m_pair * p.n_tiles;
          const int m_blk = 2 * m_pair + static_cast<int>(rank);
          tma_store_2d(&mapO, stage_out + b * kStageOutBytes, n_blk * BN + g * kGroupCols, m_blk * kBlockM);
          tma_store_commit();
          if (P_SIG(p) != nullptr && g == GROUPS - 1) {
            // the pair's next accumulator is many K blocks away: wait until the tile has been WRITTEN and publish it
            tma_store_wait_all<0>();
            signal_tile(lp);
          }
        }
      }
      tma_store_wait_all<0>();
      if (P_SIG(p) != nullptr) red_release_gpu_add(P_SIG(p) + p.ctr_frames, 1u);  // this CTA has published everything
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (warps 4..11, both CTAs)
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quad * 32 + lane;
    const int et = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    int jg = 0;
    for (int lp = pair; lp < num_ptiles; lp += num_pairs, ++it) {
      const int pt = p.rev ? num_ptiles - 1 - lp : lp;
      const int m_pair = pt / p.n_tiles;
      const int n_blk = pt - m_pair * p.n_tiles;
      const int n_base = n_blk * BN;
      float* sb = s_bias + (it & 1) * BN;
      for (int i = et; i < BN; i += kEpiThreads) sb[i] = __ldg(&p.bias[n_base + i]);
      named_barrier_sync(1, kEpiThreads);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int g = 0; g < GROUPS; ++g, ++jg) {
        const int b = jg % NB;
        mbar_wait(&res_full[b], (jg / NB) & 1);
        uint8_t* row_ptr = stage_out + b * kStageOutBytes + row * 128;
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + g * kGroupCols + half * 32, v);
        tmem_ld_wait();
        const float4* sb4 = reinterpret_cast<const float4*>(sb + g * kGroupCols + half * 32);
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          uint4* sp = reinterpret_cast<uint4*>(row_ptr + (((half * 4 + c4) ^ (row & 7)) << 4));
          const float4 b0 = sb4[2 * c4], b1 = sb4[2 * c4 + 1];
          __nv_bfloat162 o2[4];
          o2[0] = __floats2bfloat162_rn(__uint_as_float(v[8 * c4 + 0]) + b0.x, __uint_as_float(v[8 * c4 + 1]) + b0.y);
          o2[1] = __floats2bfloat162_rn(__uint_as_float(v[8 * c4 + 2]) + b0.z, __uint_as_float(v[8 * c4 + 3]) + b0.w);
          o2[2] = __floats2bfloat162_rn(__uint_as_float(v[8 * c4 + 4]) + b1.x, __uint_as_float(v[8 * c4 + 5]) + b1.y);
          o2[3] = __floats2bfloat162_rn(__uint_as_float(v[8 * c4 + 6]) + b1.z, __uint_as_float(v[8 * c4 + 7]) + b1.w);
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
          *sp = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&out_full[b]);
      }
      // this CTA's half of the accumulator is drained: tell the leader's MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
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
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while the other can still signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace phdfxk
