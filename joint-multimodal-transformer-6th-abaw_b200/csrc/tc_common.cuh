// Shared tcgen05 / TMA / mbarrier PTX wrappers and tensor-map helpers of the sm_100a tensor-core kernels
// (gemm_tc.cu, attn_tc.cu).
#pragma once
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace jmt {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;           // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kTmemCols = 512;
constexpr uint32_t kSpinLimit = 1u << 24;   // ~1 s of polling, far beyond any legitimate wait

// Division by a runtime constant without the ~150-cycle integer-divide sequence: the single-thread TMA / MMA roles
// decode tile coordinates on their critical path.  q = (n * mul) >> 32 >> shr is exact for 0 <= n < 2^31, d >= 1.
struct FastDiv {
  uint32_t mul, shr, d;
  __host__ void init(uint32_t div) {
    d = div;
    if (div == 1) { mul = 0; shr = 0; return; }
    uint32_t l = 0;
    while ((1ull << l) < div) ++l;                       // ceil(log2(div))
    mul = (uint32_t)((((1ull << l) - div) << 32) / div + 1);
    shr = l - 1;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    if (d == 1) return n;
    const uint32_t hi = __umulhi(n, mul);
    return (hi + ((n - hi) >> 1)) >> shr;
  }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};


// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < kSpinLimit; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();   // a pipeline bug must fail loudly, never hang the GPU
}
// one non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// (pdl_wait / pdl_launch_dependents: common.cuh)
// L2 prefetch of a tensor-map box (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// cta_group::2 load: data lands in this CTA's smem, complete_tx is signalled on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_bar), "r"(bytes) : "memory");
}
// Arrive on a (possibly remote) CTA's mbarrier.  Default semantics (.release.cta): a `.release.cluster` arrive compiles to
// MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR in front of the SYNCS.ARRIVE and measured ~3.5 k cycles per call in the GEMM epilogue.
// The signals sent this way hand over no generic-proxy data (TMEM reads are ordered by tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only cluster barrier (no memory ordering: mbarrier initialisation is published by fence.mbarrier_init.release.cluster,
// nothing else is handed over): avoids the GPU-scope membar + L1 invalidate of the release/acquire form at kernel start and end
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: arrives on `bar` once all prior MMAs of this thread retire.  kCta == 2: the arrive is
// multicast to the same barrier offset in both CTAs of the pair.
template <int kCta>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if constexpr (kCta == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
  }
}
template <int kCta>
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kCta == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// Warp-uniform issue: the WHOLE warp runs the issuer's control flow (so every operand is warp-uniform and lives in uniform
// registers) and one elected lane executes the instruction.  The form above, called under `if (lane == 0)`, makes ptxas wrap
// every tcgen05.mma in an ELECT + 7 x R2UR.BROADCAST + BRA.U.ANY loop (~25 instructions of a single thread per MMA: more
// than the 64-128 cycles the MMA itself takes).  `elected` = elect_one() evaluated once by the caller.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
template <int kCta>
__device__ __forceinline__ void tc_mma_elect(uint32_t elected, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kCta == 1) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(elected) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(elected) : "memory");
  }
}
template <int kCta>
__device__ __forceinline__ void tc_commit_elect(uint32_t elected, uint32_t bar) {
  if constexpr (kCta == 1) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar), "r"(elected) : "memory");
  } else {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %2;\n\t}"
                 ::"r"(bar), "r"(elected), "h"((uint16_t)3) : "memory");
  }
}
// elected-lane forms of the producer's instructions (same reasoning: the whole warp computes the warp-uniform coordinates)
__device__ __forceinline__ void mbar_expect_tx_el(uint32_t el, uint32_t bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
               ::"r"(bar), "r"(bytes), "r"(el) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster_el(uint32_t el, uint32_t cluster_bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;\n\t}"
               ::"r"(cluster_bar), "r"(bytes), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_load_4d_el(uint32_t el, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %7, 0;\n\t"
      "@q cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_load_5d_el(uint32_t el, uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %8, 0;\n\t"
      "@q cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t}"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm_el(uint32_t el, uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %7, 0;\n\t"
      "@q cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n\t}"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm_el(uint32_t el, uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %8, 0;\n\t"
      "@q cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n\t}"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(el) : "memory");
}
// elected-lane TMA stores (epilogue warps: the whole warp computes the warp-uniform coordinates, lane `el` issues, commits and waits)
__device__ __forceinline__ void tma_store_4d_el(uint32_t el, const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
               "@q cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n\t"
               "@q cp.async.bulk.commit_group;\n\t}"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d_el(uint32_t el, const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
               "@q cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n\t"
               "@q cp.async.bulk.commit_group;\n\t}"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(el) : "memory");
}
__device__ __forceinline__ void bulk_wait_read0_el(uint32_t el) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q cp.async.bulk.wait_group.read 0;\n\t}" ::"r"(el) : "memory");
}
__device__ __forceinline__ void bulk_wait_read1_el(uint32_t el) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q cp.async.bulk.wait_group.read 1;\n\t}" ::"r"(el) : "memory");
}
__device__ __forceinline__ void bulk_wait0_el(uint32_t el) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q cp.async.bulk.wait_group 0;\n\t}" ::"r"(el) : "memory");
}
// elected-lane L2 prefetch of a tensor-map box
__device__ __forceinline__ void tma_prefetch_4d_el(uint32_t el, const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
               "@q cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];\n\t}"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(el) : "memory");
}
__device__ __forceinline__ void tma_prefetch_5d_el(uint32_t el, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %6, 0;\n\t"
               "@q cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];\n\t}"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(el) : "memory");
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread = TMEM lane = output row).
// Issue only; tc_wait_ld() must precede the first use of r[] (several loads can be in flight).
__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) { tc_ld32_issue(taddr, r); tc_wait_ld(); }

// UMMA shared-memory matrix descriptor, 128-byte swizzle, Blackwell version bits.
// K-major tile  [rows][64 elem]: 8-row groups 1024 B apart (SBO); LBO unused (encoded 1).
// MN-major tile [chunk][k][64 elem]: 8-k groups 1024 B apart (SBO), 64-wide MN chunks 8192 B apart (LBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;      // SWIZZLE_128B
  return d;
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// L2 promotion of operand loads (experiment knob JMT_TMA_L2PROMO = 0 none, 1 64B, 2 128B, 3 256B; default 256B)
static inline CUtensorMapL2promotion tma_l2_promotion() {
  static const int v = []() { const char* e = getenv("JMT_TMA_L2PROMO"); return e ? atoi(e) : 3; }();
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
       : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

// 4-D bf16 tensor map over (inner, rows, b0, b1) with a {64, box_rows, 1, 1} box, 128B swizzle, zero OOB fill
static inline int make_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t rows, int64_t ld, int64_t nb0, int64_t bs0,
                    int64_t nb1, int64_t bs1, int box_rows, const char* who) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("%s: cuTensorMapEncodeTiled unavailable (no CUDA driver?)", who); return JMT_ERR_CUDA; }
  JMT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: operand pointer must be 16-byte aligned", who);
  JMT_REQUIRE(ld % 8 == 0 && (nb0 == 1 || bs0 % 8 == 0) && (nb1 == 1 || bs1 % 8 == 0),
              "%s: operand ld / batch strides must be multiples of 8 elements (ld=%lld bs0=%lld bs1=%lld)", who,
              (long long)ld, (long long)bs0, (long long)bs1);
  JMT_REQUIRE(box_rows >= 1 && box_rows <= 256, "%s: bad box rows %d", who, box_rows);
  const cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nb0, (cuuint64_t)nb1};
  const cuuint64_t row_bytes = (cuuint64_t)ld * 2;
  const cuuint64_t strides[3] = {row_bytes, nb0 > 1 ? (cuuint64_t)bs0 * 2 : row_bytes, nb1 > 1 ? (cuuint64_t)bs1 * 2 : row_bytes};
  const cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, tma_l2_promotion(),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld ld=%lld nb0=%lld bs0=%lld nb1=%lld bs1=%lld box_rows=%d",
              who, (int)r, (long long)inner, (long long)rows, (long long)ld, (long long)nb0, (long long)bs0,
              (long long)nb1, (long long)bs1, box_rows);
    return JMT_ERR_CUDA;
  }
  return JMT_OK;
}

// 5-D bf16 tensor map for an MN-major operand tile: (64 contiguous elements, k rows, 64-wide column chunks, b0, b1) with box
// {64, box_rows, box_chunks, 1, 1}: ONE TMA instruction lands `box_chunks` chunks of [box_rows x 128 B] back to back in
// shared memory -- exactly the chunked MN-major UMMA layout (LBO = chunk stride) -- instead of one instruction per chunk
// (the single producer thread pays ~150 cycles per TMA instruction).  Columns beyond `inner` read as zero.
static inline int make_map_mn5(CUtensorMap* map, const void* ptr, int64_t inner, int64_t rows, int64_t ld, int64_t nb0, int64_t bs0,
                               int64_t nb1, int64_t bs1, int box_rows, int box_chunks, const char* who) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("%s: cuTensorMapEncodeTiled unavailable (no CUDA driver?)", who); return JMT_ERR_CUDA; }
  JMT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "%s: operand pointer must be 16-byte aligned", who);
  JMT_REQUIRE(ld % 8 == 0 && (nb0 == 1 || bs0 % 8 == 0) && (nb1 == 1 || bs1 % 8 == 0),
              "%s: operand ld / batch strides must be multiples of 8 elements", who);
  JMT_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_chunks >= 1 && box_chunks <= 8, "%s: bad 5-D box", who);
  const int64_t chunks = (inner + 63) / 64;
  // dim0 spans one chunk; a ragged last chunk is expressed by clamping dim0 to what remains when there is only one chunk,
  // otherwise dim0 = 64 and the reads past `inner` in the last chunk are masked by the operand consumer (they multiply
  // zero-filled / never-read accumulator columns): callers keep inner % 64 == 0 or chunks == 1 for exact OOB zero fill.
  const cuuint64_t d0 = chunks == 1 ? (cuuint64_t)inner : 64;
  const cuuint64_t dims[5] = {d0, (cuuint64_t)rows, (cuuint64_t)chunks, (cuuint64_t)nb0, (cuuint64_t)nb1};
  const cuuint64_t row_bytes = (cuuint64_t)ld * 2;
  const cuuint64_t strides[4] = {row_bytes, 128, nb0 > 1 ? (cuuint64_t)bs0 * 2 : row_bytes, nb1 > 1 ? (cuuint64_t)bs1 * 2 : row_bytes};
  const cuuint32_t box[5] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows, (cuuint32_t)box_chunks, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, tma_l2_promotion(),
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled(5-D) failed (%d) inner=%lld rows=%lld ld=%lld box_rows=%d box_chunks=%d", who, (int)r,
              (long long)inner, (long long)rows, (long long)ld, box_rows, box_chunks);
    return JMT_ERR_CUDA;
  }
  return JMT_OK;
}

// D tensor map: (N, M, b0, b1), box {128 bytes of columns, 32 rows}, 128B swizzle (matches the epilogue staging)
static inline int make_map_d(CUtensorMap* map, const void* ptr, int dtype, int64_t inner, int64_t rows, int64_t ld, int64_t nb0,
                      int64_t bs0, int64_t nb1, int64_t bs1, const char* who) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("%s: cuTensorMapEncodeTiled unavailable (no CUDA driver?)", who); return JMT_ERR_CUDA; }
  const cuuint64_t es = dtype == JMT_F32 ? 4 : 2;
  const cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nb0, (cuuint64_t)nb1};
  const cuuint64_t row_bytes = (cuuint64_t)ld * es;
  const cuuint64_t strides[3] = {row_bytes, nb0 > 1 ? (cuuint64_t)bs0 * es : row_bytes, nb1 > 1 ? (cuuint64_t)bs1 * es : row_bytes};
  const cuuint32_t box[4] = {(cuuint32_t)(128 / es), 32, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, dtype == JMT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld ld=%lld nb0=%lld bs0=%lld nb1=%lld bs1=%lld", who,
              (int)r, (long long)inner, (long long)rows, (long long)ld, (long long)nb0, (long long)bs0, (long long)nb1,
              (long long)bs1);
    return JMT_ERR_CUDA;
  }
  return JMT_OK;
}

}  // namespace jmt
