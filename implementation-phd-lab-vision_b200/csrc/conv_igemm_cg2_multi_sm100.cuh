// K3b, multi-phase: up to three CONSECUTIVE CTA-pair convs (conv_igemm_cg2_sm100.cuh) of a bottleneck block in ONE launch
// — conv1 (1x1) -> conv2 (3x3) [-> conv3 + fused down-sample] of layer3 / layer4.
//
// Why: a launch boundary costs more than the pipeline fill.  The per-CTA timeline (tools/trace_ctas.py) shows a CTA of
// the next launch entering ~5 us after its predecessor's CTA exits, and a launch whose tile count is not a multiple of
// the grid leaves SMs idle for up to a whole tile time at its end (layer4: 98 pair-tiles on 74 CTA pairs = 66 %).
// Here the pair-tiles of the phases form ONE list, walked round-robin by the persistent CTA pairs; a tile of phase k > 0
// waits, in the TMA producer warp, for the frames it reads to have been published by phase k-1 (per-frame progress
// counters, conv_igemm_sm100.cuh) — no CTA is relaunched and a pair that runs out of phase-k tiles goes straight on with
// phase k+1.  All CTAs are resident (grid <= SMs) and the list order is a topological order, so the waits cannot
// deadlock.  Every tile is computed exactly as by conv_igemm_cg2_kernel: results are bit-identical.
#pragma once
#include "conv_igemm_cg2_sm100.cuh"

namespace phdfxk {

constexpr int kMaxPhases = 3;

__device__ __forceinline__ void st_release_cta_smem(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta_smem(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

struct Cg2MultiMaps {
  CUtensorMap a[kMaxPhases], b[kMaxPhases], o[kMaxPhases], a2[kMaxPhases];
};
struct Cg2MultiParams {
  int n_phases;
  int begin[kMaxPhases + 1];  // pair-tile list: phase k owns [begin[k], begin[k+1])
  int mode[kMaxPhases];       // MODE_TILED | MODE_IM2COL
  ConvParams ph[kMaxPhases];  // per phase; wait_ctr / sig_ctr / ctr_frames tie phase k to k-1 (phase 0: wait_ctr = null)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
conv_igemm_cg2_multi_kernel(const __grid_constant__ Cg2MultiMaps maps, const __grid_constant__ Cg2MultiParams mp) {
  using Cfg = Cg2Cfg;
  constexpr int BN = Cfg::BN;
  constexpr int NSTAGE = Cfg::NSTAGE;
  constexpr int NB = Cfg::NB;
  constexpr int LOOK = Cfg::LOOK;
  constexpr int GROUPS = Cfg::GROUPS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + NSTAGE * Cfg::STAGE_BYTES;
  uint8_t* tail = stage_out + NB * kStageOutBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + NSTAGE;
  uint64_t* tmem_full = empty_bar + NSTAGE;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;
  uint64_t* out_full = res_full + NB;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(out_full + NB);
  int* dep_cnt = reinterpret_cast<int*>(tail + 512);  // tiles of this CTA whose inputs the scout warp has seen published
  float* s_bias = reinterpret_cast<float*>(tail + 1024);

  const int warp = uniform_warp_idx();
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int total = mp.begin[mp.n_phases];
  // list position -> phase (warp-uniform)
  auto phase_of = [&](int lp) -> int {
    int k = 0;
    while (k + 1 < mp.n_phases && lp >= mp.begin[k + 1]) ++k;
    return k;
  };
  // list position -> pair-tile of its phase
  auto tile_of = [&](int lp, int k) -> int {
    const int i = lp - mp.begin[k];
    return mp.ph[k].rev ? mp.begin[k + 1] - mp.begin[k] - 1 - i : i;
  };

  if (warp == 0 && lane == 0)
    for (int k = 0; k < mp.n_phases; ++k) {
      tma_prefetch_desc(&maps.a[k]);
      tma_prefetch_desc(&maps.b[k]);
    }
  if (warp == 2 && lane == 0)
    for (int k = 0; k < mp.n_phases; ++k) tma_prefetch_desc(&maps.o[k]);
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
    *dep_cnt = 0;
    fence_mbar_init();
  }
  cluster_sync_all();
  if (warp == 2) {
    tmem_alloc_cg2(tmem_ptr, Cfg::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_launch_dependents();
  griddep_wait();  // phase 0 reads the previous launch's output

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    int stage = 0;
    uint32_t phase = 0;
    int ord = 0;
    for (int lp = pair; lp < total; lp += num_pairs, ++ord) {
      const int k = phase_of(lp);
      const ConvParams& p = mp.ph[k];
      const int mode = mp.mode[k];
      const int pt = tile_of(lp, k);
      const int m_pair = pt / p.n_tiles;
      const int n_blk = pt - m_pair * p.n_tiles;
      const int m_blk = 2 * m_pair + static_cast<int>(rank);
      int cw = 0, ch = 0, cn = 0;
      if (mode == MODE_IM2COL || p.src2_stride == 2) {
        const int m0 = m_blk * kBlockM;
        const int pq = p.P * p.Q;
        cn = m0 / pq;
        const int rem = m0 - cn * pq;
        const int p0 = rem / p.Q;
        const int q0 = rem - p0 * p.Q;
        if (mode == MODE_IM2COL) {
          cw = q0 * p.stride - p.pad;
          ch = p0 * p.stride - p.pad;
        } else {
          cw = q0 * 2;
          ch = p0 * 2;
        }
      }
      if (p.wait_ctr != nullptr) {
        // the scout warp (warp 3) polls the progress counters — an L2 round trip or two per tile — ahead of time; here
        // it only takes a look at shared memory
#ifdef PHDFX_TRAP
        unsigned long long t0 = 0;
        uint32_t spins = 0;
#endif
        while (ld_acquire_cta_smem(dep_cnt) <= ord) {
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
        fence_proxy_async_all();
      }
      const CUtensorMap* mA = &maps.a[k];
      const CUtensorMap* mB = &maps.b[k];
      const CUtensorMap* mA2 = &maps.a2[k];
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sB = sA + Cfg::A_BYTES;
        if (leader) mbar_arrive_expect_tx_elect(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
        if (mode == MODE_TILED) {
          if (kb < p.kb_split)
            tma_load_2d_cg2_elect(mA, &full_bar[stage], sA, kb * 64, m_blk * kBlockM);
          else if (p.src2_stride == 1)
            tma_load_2d_cg2_elect(mA2, &full_bar[stage], sA, (kb - p.kb_split) * 64, m_blk * kBlockM);
          else
            tma_load_im2col_4d_cg2_elect(mA2, &full_bar[stage], sA, (kb - p.kb_split) * 64, cw, ch, cn, 0, 0);
        } else {
          const int tap = kb / p.kb_per_tap;
          const int cb = kb - tap * p.kb_per_tap;
          const int r = tap / p.S;
          const int s = tap - r * p.S;
          tma_load_im2col_4d_cg2_elect(mA, &full_bar[stage], sA, cb * 64, cw, ch, cn, static_cast<uint16_t>(s),
                                       static_cast<uint16_t>(r));
        }
        tma_load_2d_cg2_elect(mB, &full_bar[stage], sB, kb * 64, n_blk * BN + static_cast<int>(rank) * (BN / 2));
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
      for (int lp = pair; lp < total; lp += num_pairs) {
        const int num_kb = mp.ph[phase_of(lp)].num_kb;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        bool ready = false;
        for (int kb = 0; kb < num_kb; ++kb) {
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
          ready = (kb + 1 < num_kb) ? mbar_test(&full_bar[nstage], nphase) : false;
          umma_bf16_x4_cg2_elect(d_tmem, make_kmajor_desc(a_addr, 128), make_kmajor_desc(b_addr, 128), idesc,
                                 kb != 0 ? 1u : 0u);
          umma_commit_cg2_mc_elect(&empty_bar[stage]);
          stage = nstage;
          phase = nphase;
        }
        umma_commit_cg2_mc_elect(&tmem_full[acc]);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------ epilogue DMA (one thread per CTA)
    if (lane == 0) {
      const int my_tiles = (total > pair) ? (total - pair + num_pairs - 1) / num_pairs : 0;
      const int J = my_tiles * GROUPS;
      int published = 0;  // phases this CTA has reported complete (every tile of theirs it owns is written)
      // this CTA has nothing (more) to write in phases < upto: tell their consumers
      auto publish_phases = [&](int upto) {
        for (; published < upto; ++published)
          if (mp.ph[published].sig_ctr != nullptr)
            red_release_gpu_add(mp.ph[published].sig_ctr + mp.ph[published].ctr_frames, 1u);
      };
      for (int t = 0; t < J + LOOK; ++t) {
        if (t < J) {
          if (t >= NB) tma_store_wait_read<NB - LOOK - 1>();
          mbar_arrive(&res_full[t % NB]);
        }
        if (t >= LOOK) {
          const int u = t - LOOK;
          const int b = u % NB;
          const int lp = pair + (u / GROUPS) * num_pairs;
          const int k = phase_of(lp);
          const ConvParams& p = mp.ph[k];
          const int g = u % GROUPS;
          if (g == 0) publish_phases(k);  // stores of earlier phases were waited for when their last tile was published
          mbar_wait(&out_full[b], (u / NB) & 1);
          const int pt = tile_of(lp, k);
          const int m_pair = pt / p.n_tiles;
          const int n_blk = pt - m_pair * p.n_tiles;
          const int m_blk = 2 * m_pair + static_cast<int>(rank);
          tma_store_2d(&maps.o[k], stage_out + b * kStageOutBytes, n_blk * BN + g * kGroupCols, m_blk * kBlockM);
          tma_store_commit();
          if (p.sig_ctr != nullptr && g == GROUPS - 1) {
            // the pair's next accumulator is many K blocks away: wait until the tile has been WRITTEN and publish it
            tma_store_wait_all<0>();
            if (m_blk * kBlockM < p.M) {
              const int r1 = (m_blk + 1) * kBlockM < p.M ? (m_blk + 1) * kBlockM : p.M;
              dep_signal_rows(p.sig_ctr, m_blk * kBlockM, r1, p.P * p.Q, GROUPS);
            }
          }
        }
      }
      tma_store_wait_all<0>();
      publish_phases(mp.n_phases);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ dependency scout (both CTAs)
    // Walks this CTA's tile list ahead of the producer and waits, tile by tile, until the frames the tile reads have been
    // published by the phase before; reports how far it got through shared memory.
    int done_phase = -1;  // phases <= done_phase are known to have published everything
    int ord = 0;
    for (int lp = pair; lp < total; lp += num_pairs, ++ord) {
      const int k = phase_of(lp);
      const ConvParams& p = mp.ph[k];
      if (p.wait_ctr != nullptr && done_phase < k - 1) {
        const int m_blk = 2 * (tile_of(lp, k) / p.n_tiles) + static_cast<int>(rank);
        if (m_blk * kBlockM < p.M) {
          const int pq = p.P * p.Q;
          const int r1 = (m_blk + 1) * kBlockM < p.M ? (m_blk + 1) * kBlockM : p.M;
          dep_wait_frames(p.wait_ctr, p.wait_full, (m_blk * kBlockM) / pq, (r1 - 1) / pq);
        }
        if (dep_grid_done(p.wait_ctr + p.ctr_frames, p.wait_ctas)) done_phase = k - 1;
      }
      __syncwarp();
      if (lane == 0) st_release_cta_smem(dep_cnt, ord + 1);
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
    for (int lp = pair; lp < total; lp += num_pairs, ++it) {
      const int k = phase_of(lp);
      const ConvParams& p = mp.ph[k];
      const int pt = tile_of(lp, k);
      const int m_pair = pt / p.n_tiles;
      const int n_blk = pt - m_pair * p.n_tiles;
      const int n_base = n_blk * BN;
      float* sb = s_bias + (it & 1) * BN;
      for (int i = et; i < BN; i += kEpiThreads) sb[i] = __ldg(&p.bias[n_base + i]);
      named_barrier_sync(1, kEpiThreads);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
      const int relu = p.relu;
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
          if (relu) {
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace phdfxk
