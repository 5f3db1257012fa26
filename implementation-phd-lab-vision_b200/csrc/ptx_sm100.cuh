// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (tiled + im2col), tcgen05 (alloc/mma/commit/ld).
// Everything the conv kernels issue to the async proxy or the tensor pipe goes through here.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace phdfxk {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Spin on an mbarrier phase.  Production builds poll without a bound: measured on B200 (profiles/r02/ab_trap_*.txt),
// any extra work on the failed-poll path costs throughput — the waiting warps share issue slots and shared-memory
// bandwidth with the epilogue warps (poll-count trap: +2.1 % on the whole trunk; %globaltimer-based trap: +4.9 %).
// -DPHDFX_TRAP (debug builds, used while developing a kernel) bounds every wait by TIME (%globaltimer, 20 s) and traps,
// so a pipeline bug (lost arrive, wrong tx count) ends the kernel instead of hanging the GPU; a legitimately long wait
// (time-sliced GPU, MPS neighbour, profiler replay) never comes near the bound.
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#ifdef PHDFX_TRAP
constexpr unsigned long long kMbarTrapNs = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s
#endif
#ifndef PHDFX_SPIN_SLEEP_NS
#define PHDFX_SPIN_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
#ifdef PHDFX_TRAP
  uint32_t spins = 0;
  unsigned long long t0 = 0;
#endif
  do {
#ifdef PHDFX_WAIT_HINT_NS
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(PHDFX_WAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
#endif
#if PHDFX_SPIN_SLEEP_NS > 0
    if (!done) __nanosleep(PHDFX_SPIN_SLEEP_NS);
#endif
#ifdef PHDFX_TRAP
    if (!done && (++spins & 0xFFFFu) == 0) {
      const unsigned long long now = global_timer_ns();
      if (t0 == 0)
        t0 = now;
      else if (now - t0 > kMbarTrapNs)
        __trap();
    }
#endif
  } while (!done);
}

// Non-blocking probe: has the phase with this parity completed?  Used to overlap the ~90-cycle barrier query of the
// NEXT pipeline stage with issuing the current stage's MMAs.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
// im2col-mode load over an NHWC tensor: (c, w, h, n) is the base pixel (input coordinates, starting at the
// bounding-box lower corner), (off_w, off_h) the filter-tap offset.
__device__ __forceinline__ void tma_load_im2col_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2], {%7, %8};" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- tile-granular dependencies between launches
// A launch may start consuming its predecessor's output before the predecessor's grid has completed (conv_igemm_sm100.cuh,
// "Frame progress counters"): the producer publishes per-frame progress with a release-add AFTER its TMA stores have
// completed (cp.async.bulk.wait_group without .read) and a proxy fence; the consumer polls with acquire loads and fences
// the async proxy before its TMA loads of those frames.
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> fp32; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once every previously issued tcgen05.mma of this thread has retired
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (lane t <-> TMEM lane base+t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 consecutive 32-bit columns <- 16 registers per thread (lane t <-> TMEM lane base+t).  Used to park a
// bf16 tile (two K elements per column) in TMEM as the A operand of a later tcgen05.mma.
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- warp-uniform issue: every lane of the (converged) warp executes the statement, one elected lane performs it.
// Keeping the surrounding control flow warp-uniform lets ptxas hold descriptors / coordinates in uniform registers
// instead of wrapping each UTCHMMA / UTMALDG in an ELECT + R2UR.BROADCAST loop.
__device__ __forceinline__ void umma_bf16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand in TENSOR MEMORY (128 lanes = GEMM rows, bf16 pairs along the columns: K = 16 is 8 columns), B in smem.
__device__ __forceinline__ void umma_bf16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One K block = four (or two) K=16 MMAs issued back to back under ONE election; each step advances both descriptors'
// start-address fields by 2 (= 32 bytes).  Keeping the whole group in one asm block keeps the per-MMA issue cost to a
// couple of uniform instructions (it is the single MMA-issuing warp that bounds small-N layers).
__device__ __forceinline__ void umma_bf16_x4_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "add.u32 alo, alo, 2;\n"
      "add.u32 blo, blo, 2;\n"
      "mov.b64 da, {alo, ahi};\n"
      "mov.b64 db, {blo, bhi};\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u32 alo, alo, 2;\n"
      "add.u32 blo, blo, 2;\n"
      "mov.b64 da, {alo, ahi};\n"
      "mov.b64 db, {blo, bhi};\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
      "add.u32 alo, alo, 2;\n"
      "add.u32 blo, blo, 2;\n"
      "mov.b64 da, {alo, ahi};\n"
      "mov.b64 db, {blo, bhi};\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_x2_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
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
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "add.u32 alo, alo, 2;\n"
      "add.u32 blo, blo, 2;\n"
      "mov.b64 da, {alo, ahi};\n"
      "mov.b64 db, {blo, bhi};\n"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, pt;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate_first)
      : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_elect(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                                  int c2) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_elect(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                                  int c2, int c3) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_elect(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                                  int c2, int c3, int c4) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_elect(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w,
                                                         int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "{\n"
      ".reg .pred pe;\n"
      "elect.sync _|pe, 0xffffffff;\n"
      "@pe cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2], {%7, %8};\n"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
      "h"(off_h)
      : "memory");
}
// warp index that ptxas can treat as warp-uniform
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor for a K-major operand tile whose rows are `row_bytes` (64 or 128) wide and
// stored densely with the matching TMA swizzle (SWIZZLE_64B / SWIZZLE_128B): 8-row groups are
// 8*row_bytes apart (SBO); LBO is unused for swizzled K-major layouts (canonical value 1).
// Bit layout: [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout type.
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr, uint32_t row_bytes) {
  const uint64_t layout = (row_bytes == 128) ? 2ull : 4ull;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  const uint64_t sbo = (8ull * row_bytes) >> 4;
  return (static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4)) | (1ull << 16) | (sbo << 32) | (1ull << 46) |
         (layout << 61);
}
// Instruction descriptor: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1), both K-major (bits 15,16 = 0),
// N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---------------------------------------------------------------- programmatic dependent launch
// launch_dependents: the next kernel in the stream (launched with programmaticStreamSerialization) may start its
// CTAs on SMs this grid has vacated; wait: block until every prerequisite grid has completed and flushed.
// Both are no-ops when the kernel was launched without the attribute.
__device__ __forceinline__ void griddep_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- misc
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ void named_barrier_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace phdfxk
