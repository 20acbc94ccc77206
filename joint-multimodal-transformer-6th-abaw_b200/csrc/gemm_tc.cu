// jmt_gemm_bf16: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0 (1 elected thread) : TMA producer   -- cp.async.bulk.tensor.4d -> 128B-swizzled smem ring
//   warp 1 (1 elected thread) : MMA issuer     -- tcgen05.mma.cta_group::1.kind::f16, fp32 accum in TMEM
//   warps 2..9                : epilogue       -- tcgen05.ld 32x32b -> alpha/bias/activation -> global
//
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the mainloop of
// tile i+1.  One descriptor (jmt_gemm_desc) covers every dense contraction of the JMT path: Linear
// fwd/dgrad/wgrad, QK^T / PV and their gradients (batched over (b, head) through 4-D tensor maps,
// K- or MN-major operands selected in the UMMA instruction descriptor) and the dilated causal
// Conv1d of the TCN as an implicit GEMM (taps = extra K blocks with a shifted TMA row coordinate;
// causal zero padding = TMA out-of-bounds fill).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

int jmt_validate_gemm_desc(const jmt_gemm_desc* g, const char* who);

namespace jmt {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;           // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kNumEpiWarps = 8;         // two warps per TMEM lane quarter, interleaved over column chunks
constexpr int kTcThreads = 64 + 32 * kNumEpiWarps;
constexpr int kAStageBytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;       // TMEM columns per accumulator stage
constexpr int kEpiStageBytes = 4096;      // 32 rows x 128 B per epilogue warp
constexpr int kEpiSmemBytes = kNumEpiWarps * kEpiStageBytes + 1024;   // staging tiles + one shared bias tile
constexpr uint32_t kSpinLimit = 1u << 24;   // ~1 s of polling, far beyond any legitimate wait

struct TcParams {
  int M, N, K, block_n;
  int m_tiles, n_tiles, batch_tiles, split_k, total_tiles;
  int kblocks, rb_n, iters_total;
  int nb0;
  int a_major, b_major;
  int a_shift0, a_shift_step, b_shift0, b_shift_step;
  int b_has_b0, b_has_b1;
  int reduce_batch;
  uint32_t idesc;
  int b_stage_bytes, b_tx_bytes, stages;
  // epilogue
  void* d;
  const float* bias;
  int64_t d_ld, d_bs0, d_bs1;
  float alpha, slope;
  int d_dtype, act, store_mode, vec_ok;
  int cluster;        // 1, or 2 = CTA pairs along M sharing the B tile by TMA multicast
  int m_pairs;        // ceil(m_tiles / cluster)
  int tma_store;      // epilogue through swizzled smem + TMA store / reduce-add (needs 16-byte aligned D geometry)
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
__device__ __forceinline__ void tma_load_4d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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

struct TileCoord { int m0, n0, batch, it0, it1; };

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t, int crank) {
  TileCoord c;
  const int nt = t % p.n_tiles; t /= p.n_tiles;
  const int mp = t % p.m_pairs; t /= p.m_pairs;
  c.batch = t % p.batch_tiles;
  const int split = t / p.batch_tiles;
  c.m0 = (mp * p.cluster + crank) * kBlockM;     // an odd tail pair gives rank 1 an all-out-of-range (ghost) tile
  c.n0 = nt * p.block_n;
  c.it0 = (int)(((int64_t)split * p.iters_total) / p.split_k);
  c.it1 = (int)(((int64_t)(split + 1) * p.iters_total) / p.split_k);
  return c;
}

__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_d, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms are 1024 B)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + p.stages * kAStageBytes;
  const uint32_t sD = sB + p.stages * p.b_stage_bytes;       // epilogue warps x 4 KiB staging (1024-aligned)
  const uint32_t sBias = sD + kNumEpiWarps * kEpiStageBytes; // 256 floats: this tile's bias slice
  const uint32_t bars = sBias + 1024;                        // 8-byte aligned
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kMaxStages;
  const uint32_t tfull_bar = bars + 16 * kMaxStages, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_aligned + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = p.cluster == 2 ? (int)cluster_ctarank() : 0;
  const int first_tile = blockIdx.x / p.cluster;          // tile (pair) index owned by this CTA's cluster
  const int tile_stride = gridDim.x / p.cluster;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, p.cluster); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, kNumEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster == 2) cluster_sync_all();       // peer barriers are initialised before any multicast can land
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ================================ TMA producer ================================
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
      int stage = 0; uint32_t phase = 0;
      const int b_chunks = (p.block_n + 63) / 64;
      for (int t = first_tile; t < p.total_tiles; t += tile_stride) {
        const TileCoord c = decode_tile(p, t, crank);
        for (int it = c.it0; it < c.it1; ++it) {
          const int kb = it % p.kblocks;
          const int rb = (it / p.kblocks) % p.rb_n;
          const int tap = it / (p.kblocks * p.rb_n);
          const int bidx = p.reduce_batch ? rb : c.batch;
          const int b0 = bidx % p.nb0, b1 = bidx / p.nb0;
          const int bb0 = p.b_has_b0 ? b0 : 0, bb1 = p.b_has_b1 ? b1 : 0;
          const int ash = p.a_shift0 + tap * p.a_shift_step;
          const int bsh = p.b_shift0 + tap * p.b_shift_step;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          mbar_expect_tx(fb, kAStageBytes + p.b_tx_bytes);
          const uint32_t a_dst = sA + stage * kAStageBytes;
          const uint32_t b_dst = sB + stage * p.b_stage_bytes;
          if (p.a_major == JMT_MAJOR_K) {
            tma_load_4d(a_dst, &tma_a, fb, kb * kBlockK, c.m0 + ash, b0, b1);
          } else {
            tma_load_4d(a_dst, &tma_a, fb, c.m0, kb * kBlockK + ash, b0, b1);
            tma_load_4d(a_dst + 8192, &tma_a, fb, c.m0 + 64, kb * kBlockK + ash, b0, b1);
          }
          if (p.cluster == 2) {
            // each CTA of the pair fetches half of the shared B tile and multicasts it to both
            if (p.b_major == JMT_MAJOR_K) {
              const int half = p.block_n >> 1;
              tma_load_4d_mc(b_dst + crank * half * 128, &tma_b, fb, tap * p.K + kb * kBlockK, c.n0 + crank * half, bb0, bb1, 3);
            } else {
              const int hc = b_chunks >> 1;
              for (int ch = crank * hc; ch < (crank + 1) * hc; ++ch)
                tma_load_4d_mc(b_dst + ch * 8192, &tma_b, fb, c.n0 + ch * 64, kb * kBlockK + bsh, bb0, bb1, 3);
            }
          } else if (p.b_major == JMT_MAJOR_K) {
            tma_load_4d(b_dst, &tma_b, fb, tap * p.K + kb * kBlockK, c.n0, bb0, bb1);
          } else {
            for (int ch = 0; ch < b_chunks; ++ch)
              tma_load_4d(b_dst + ch * 8192, &tma_b, fb, c.n0 + ch * 64, kb * kBlockK + bsh, bb0, bb1);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================================ MMA issuer ================================
      int stage = 0; uint32_t phase = 0;
      int tile_iter = 0;
      const uint32_t a_lbo = p.a_major == JMT_MAJOR_K ? 16u : 8192u;
      const uint32_t b_lbo = p.b_major == JMT_MAJOR_K ? 16u : 8192u;
      const uint32_t a_kstep = p.a_major == JMT_MAJOR_K ? (kUmmaK * 2) >> 4 : (kUmmaK * 128) >> 4;   // desc.lo units (16 B)
      const uint32_t b_kstep = p.b_major == JMT_MAJOR_K ? (kUmmaK * 2) >> 4 : (kUmmaK * 128) >> 4;
      for (int t = first_tile; t < p.total_tiles; t += tile_stride, ++tile_iter) {
        const TileCoord c = decode_tile(p, t, crank);
        const int acc = tile_iter & 1;
        const uint32_t acc_phase = (tile_iter >> 1) & 1;
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        for (int it = c.it0; it < c.it1; ++it) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc(sA + stage * kAStageBytes, a_lbo, 1024);
          const uint64_t b_desc = make_smem_desc(sB + stage * p.b_stage_bytes, b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            tc_mma(d_tmem, a_desc + (uint64_t)(k * a_kstep), b_desc + (uint64_t)(k * b_kstep), p.idesc,
                   (it > c.it0 || k > 0) ? 1u : 0u);
          if (p.cluster == 2) tc_commit_mc(empty_bar + 8 * stage, 3);   // both CTAs' producers write this slot
          else tc_commit(empty_bar + 8 * stage);  // frees the smem slot once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar + 8 * acc);          // accumulator ready for the epilogue
      }
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int ew = warp - 2;                      // epilogue warp index 0..7
    const int half = ew >> 2;                     // which interleaved set of column chunks this warp drains
    const uint32_t stage_smem = sD + ew * kEpiStageBytes;
    float* bias_ptr = reinterpret_cast<float*>(smem_aligned + (sBias - smem_base));
    const int et = threadIdx.x - 64;              // 0..255 among epilogue threads
    const uint32_t row_smem = stage_smem + lane * 128;
    const uint32_t sw = lane & 7;                 // 128B-swizzle phase of this thread's staging row
    int tile_iter = 0;
    for (int t = first_tile; t < p.total_tiles; t += tile_stride, ++tile_iter) {
      const TileCoord c = decode_tile(p, t, crank);
      const int split = (t / (p.n_tiles * p.m_pairs)) / p.batch_tiles;
      const int acc = tile_iter & 1;
      const uint32_t acc_phase = (tile_iter >> 1) & 1;
      const bool add_bias = p.bias != nullptr && split == 0;
      // stage this tile's bias slice in shared memory (overlaps the mainloop); named barrier 1 = epilogue warps
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kNumEpiWarps) : "memory");     // previous tile's readers are done
      if (et < p.block_n) bias_ptr[et] = (add_bias && c.n0 + et < p.N) ? __ldg(p.bias + c.n0 + et) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kNumEpiWarps) : "memory");
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      const int m = c.m0 + q * 32 + lane;
      const int b0 = c.batch % p.nb0, b1 = c.batch / p.nb0;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride);
      if (p.tma_store) {
        const bool warp_rows_valid = c.m0 + q * 32 < p.M;      // warp-uniform
        if (p.d_dtype == JMT_BF16) {
          for (int c0 = half * 64; c0 < p.block_n; c0 += 128) {
            if (c.n0 + c0 >= p.N) break;
            uint32_t pk[32];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (c0 + 32 * h < p.block_n) {
                uint32_t r[32];
                tc_ld32(tbase + c0 + 32 * h, r);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 bv = *reinterpret_cast<const float4*>(bias_ptr + c0 + 32 * h + j);
                  const float x0 = apply_act(fmaf(p.alpha, __uint_as_float(r[j]), bv.x), p.act, p.slope);
                  const float x1 = apply_act(fmaf(p.alpha, __uint_as_float(r[j + 1]), bv.y), p.act, p.slope);
                  const float x2 = apply_act(fmaf(p.alpha, __uint_as_float(r[j + 2]), bv.z), p.act, p.slope);
                  const float x3 = apply_act(fmaf(p.alpha, __uint_as_float(r[j + 3]), bv.w), p.act, p.slope);
                  pk[16 * h + j / 2] = pack_bf16(x0, x1);
                  pk[16 * h + j / 2 + 1] = pack_bf16(x2, x3);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[16 * h + j] = 0u;
              }
            }
            if (lane == 0) bulk_wait_read0();       // previous TMA store has finished reading the staging tile
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              st_shared_v4(row_smem + ((ch ^ sw) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0 && warp_rows_valid) {
              if (p.store_mode == JMT_STORE) tma_store_4d(&tma_d, stage_smem, c.n0 + c0, c.m0 + q * 32, b0, b1);
              else tma_reduce_add_4d(&tma_d, stage_smem, c.n0 + c0, c.m0 + q * 32, b0, b1);
              bulk_commit();
            }
          }
        } else {
          for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
            if (c.n0 + c0 >= p.N) break;
            uint32_t r[32];
            tc_ld32(tbase + c0, r);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = *reinterpret_cast<const float4*>(bias_ptr + c0 + j);
              r[j] = __float_as_uint(apply_act(fmaf(p.alpha, __uint_as_float(r[j]), bv.x), p.act, p.slope));
              r[j + 1] = __float_as_uint(apply_act(fmaf(p.alpha, __uint_as_float(r[j + 1]), bv.y), p.act, p.slope));
              r[j + 2] = __float_as_uint(apply_act(fmaf(p.alpha, __uint_as_float(r[j + 2]), bv.z), p.act, p.slope));
              r[j + 3] = __float_as_uint(apply_act(fmaf(p.alpha, __uint_as_float(r[j + 3]), bv.w), p.act, p.slope));
            }
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              st_shared_v4(row_smem + ((ch ^ sw) << 4), r[4 * ch], r[4 * ch + 1], r[4 * ch + 2], r[4 * ch + 3]);
            fence_async_smem();
            __syncwarp();
            if (lane == 0 && warp_rows_valid) {
              if (p.store_mode == JMT_STORE) tma_store_4d(&tma_d, stage_smem, c.n0 + c0, c.m0 + q * 32, b0, b1);
              else tma_reduce_add_4d(&tma_d, stage_smem, c.n0 + c0, c.m0 + q * 32, b0, b1);
              bulk_commit();
            }
          }
        }
      } else {
        // direct (unaligned D geometry): per-thread row stores / atomics
        const int64_t row_off = (int64_t)b0 * p.d_bs0 + (int64_t)b1 * p.d_bs1 + (int64_t)m * p.d_ld;
        for (int c0 = half * 32; c0 < p.block_n; c0 += 64) {
          const int n = c.n0 + c0;
          if (n >= p.N) break;                      // warp-uniform
          uint32_t r[32];
          tc_ld32(tbase + c0, r);
          if (m >= p.M) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (n + j >= p.N) break;
            const float v = apply_act(fmaf(p.alpha, __uint_as_float(r[j]), bias_ptr[c0 + j]), p.act, p.slope);
            const int64_t idx = row_off + n + j;
            if (p.d_dtype == JMT_F32) {
              float* d = (float*)p.d;
              if (p.store_mode == JMT_STORE) d[idx] = v;
              else if (p.store_mode == JMT_ACCUMULATE) d[idx] += v;
              else atomicAdd(d + idx, v);
            } else {
              __nv_bfloat16* d = (__nv_bfloat16*)p.d;
              d[idx] = __float2bfloat16_rn(p.store_mode == JMT_STORE ? v : __bfloat162float(d[idx]) + v);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
    }
    if (lane == 0) bulk_wait0();                   // all TMA stores of this warp have completed
  }

  tc_fence_before();
  __syncthreads();
  if (p.cluster == 2) cluster_sync_all();       // no CTA exits while its peer may still signal / multicast into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// 4-D bf16 tensor map over (inner, rows, b0, b1) with a {64, box_rows, 1, 1} box, 128B swizzle, zero OOB fill
static int make_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t rows, int64_t ld, int64_t nb0, int64_t bs0,
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
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld ld=%lld nb0=%lld bs0=%lld nb1=%lld bs1=%lld box_rows=%d",
              who, (int)r, (long long)inner, (long long)rows, (long long)ld, (long long)nb0, (long long)bs0,
              (long long)nb1, (long long)bs1, box_rows);
    return JMT_ERR_CUDA;
  }
  return JMT_OK;
}

// D tensor map: (N, M, b0, b1), box {128 bytes of columns, 32 rows}, 128B swizzle (matches the epilogue staging)
static int make_map_d(CUtensorMap* map, const void* ptr, int dtype, int64_t inner, int64_t rows, int64_t ld, int64_t nb0,
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

static int pick_block_n(int N) {
  int best = 32; long best_cost = -1;
  for (int bn = 32; bn <= 256; bn += 32) {
    const int tiles = (N + bn - 1) / bn;
    // padded columns computed + a per-tile overhead of ~24 columns' worth of work
    const long cost = (long)tiles * bn + 24L * tiles;
    if (best_cost < 0 || cost <= best_cost) { best = bn; best_cost = cost; }
  }
  return best;
}

}  // namespace jmt

using namespace jmt;

extern "C" int jmt_gemm_bf16(const jmt_gemm_desc* g, void* stream) {
  int rc = jmt_validate_gemm_desc(g, "jmt_gemm_bf16");
  if (rc != JMT_OK) return rc;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.block_n = pick_block_n(g->N);
  p.m_tiles = (g->M + kBlockM - 1) / kBlockM;
  p.n_tiles = (g->N + p.block_n - 1) / p.block_n;
  const int nb = g->nb0 * g->nb1;
  p.reduce_batch = g->reduce_batch ? 1 : 0;
  p.batch_tiles = p.reduce_batch ? 1 : nb;
  p.kblocks = (g->K + kBlockK - 1) / kBlockK;
  p.rb_n = p.reduce_batch ? nb : 1;
  p.iters_total = g->ntaps * p.rb_n * p.kblocks;
  p.split_k = g->split_k < p.iters_total ? g->split_k : p.iters_total;
  // CTA pairs along M with B multicast when the M tiling wastes < ~10 % on the ghost tile
  const int b_chunks_h = (p.block_n + 63) / 64;
  const bool mc_layout_ok = g->b_major == JMT_MAJOR_K ? true : (b_chunks_h % 2 == 0);
  const char* env_cl = getenv("JMT_GEMM_CLUSTER");
  const bool cl_enabled = env_cl ? atoi(env_cl) != 0 : true;
  p.cluster = (cl_enabled && mc_layout_ok && p.m_tiles >= 2 && (p.m_tiles % 2 == 0 || p.m_tiles >= 9)) ? 2 : 1;
  p.m_pairs = (p.m_tiles + p.cluster - 1) / p.cluster;
  const int64_t total = (int64_t)p.m_pairs * p.n_tiles * p.batch_tiles * p.split_k;
  JMT_REQUIRE(total < (1ll << 31), "jmt_gemm_bf16: too many tiles");
  p.total_tiles = (int)total;
  p.nb0 = g->nb0;
  p.a_major = g->a_major; p.b_major = g->b_major;
  p.a_shift0 = g->a_shift0; p.a_shift_step = g->a_shift_step;
  p.b_shift0 = g->b_shift0; p.b_shift_step = g->b_shift_step;
  p.b_has_b0 = (g->nb0 > 1 && g->b_bs0 != 0) ? 1 : 0;
  p.b_has_b1 = (g->nb1 > 1 && g->b_bs1 != 0) ? 1 : 0;
  JMT_REQUIRE(!(g->ntaps > 1 && g->b_major == JMT_MAJOR_K && g->K % 8 != 0), "jmt_gemm_bf16: taps need K %% 8 == 0");
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)g->a_major << 15) | ((uint32_t)g->b_major << 16) |
            ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
  const int b_chunks = (p.block_n + 63) / 64;
  p.b_stage_bytes = g->b_major == JMT_MAJOR_K ? p.block_n * 128 : b_chunks * 8192;
  p.b_tx_bytes = p.b_stage_bytes;
  const int stage_bytes = kAStageBytes + p.b_stage_bytes;
  const int budget = 227 * 1024 - 1024 /*align*/ - 512 /*barriers*/ - kEpiSmemBytes;
  p.stages = budget / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  JMT_REQUIRE(p.stages >= 2, "jmt_gemm_bf16: shared memory budget");
  p.d = g->d; p.bias = g->bias; p.d_ld = g->d_ld; p.d_bs0 = g->d_bs0; p.d_bs1 = g->d_bs1;
  p.alpha = g->alpha; p.slope = g->slope; p.d_dtype = g->d_dtype; p.act = g->act; p.store_mode = g->store_mode;
  const int64_t es = g->d_dtype == JMT_F32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(g->d) & 15) == 0 && (g->d_ld * es) % 16 == 0 &&
              (g->d_bs0 * es) % 16 == 0 && (g->d_bs1 * es) % 16 == 0) ? 1 : 0;

  CUtensorMap map_a, map_b;
  // A: K-major -> (K, a_rows) box {64,128}; MN-major -> (M, a_rows = k extent) box {64,64}
  if (g->a_major == JMT_MAJOR_K)
    rc = make_map(&map_a, g->a, g->K, g->a_rows, g->a_ld, g->nb0, g->a_bs0, g->nb1, g->a_bs1, kBlockM, "jmt_gemm_bf16(A)");
  else
    rc = make_map(&map_a, g->a, g->M, g->a_rows, g->a_ld, g->nb0, g->a_bs0, g->nb1, g->a_bs1, kBlockK, "jmt_gemm_bf16(A)");
  if (rc != JMT_OK) return rc;
  const int bnb0 = p.b_has_b0 ? g->nb0 : 1, bnb1 = p.b_has_b1 ? g->nb1 : 1;
  if (g->b_major == JMT_MAJOR_K)
    rc = make_map(&map_b, g->b, (int64_t)g->ntaps * g->K, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1,
                  p.block_n / p.cluster /* each CTA of a pair fetches (and multicasts) half of the tile's rows */, "jmt_gemm_bf16(B)");
  else
    rc = make_map(&map_b, g->b, g->N, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1, kBlockK, "jmt_gemm_bf16(B)");
  if (rc != JMT_OK) return rc;

  // D through TMA (store / reduce-add) when its geometry is 16-byte aligned; bf16 read-modify-write
  // accumulation and fp32 atomics both become cp.reduce.async.bulk.tensor .add
  CUtensorMap map_d = map_a;
  p.tma_store = p.vec_ok;
  if (p.tma_store) {
    const int dnb0 = p.reduce_batch ? 1 : g->nb0, dnb1 = p.reduce_batch ? 1 : g->nb1;
    rc = make_map_d(&map_d, g->d, g->d_dtype, g->N, g->M, g->d_ld, dnb0, g->d_bs0, dnb1, g->d_bs1, "jmt_gemm_bf16(D)");
    if (rc != JMT_OK) return rc;
  }
  const int smem = 1024 + p.stages * stage_bytes + kEpiSmemBytes + 512;
  static std::atomic<int> attr_set[64];     // per device (immutable once set)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); set_error("jmt_gemm_bf16: no CUDA device"); return JMT_ERR_CUDA; }
  if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("jmt_gemm_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return JMT_ERR_CUDA; }
    attr_set[dev & 63].store(1, std::memory_order_release);
  }
  const int max_groups = kNumSMs / p.cluster;
  const int groups = p.total_tiles < max_groups ? p.total_tiles : max_groups;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(groups * p.cluster);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel, map_a, map_b, map_d, p);
  if (le != cudaSuccess) {
    set_error("jmt_gemm_bf16: cudaLaunchKernelEx: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JMT_ERR_CUDA;
  }
  return check_launch("gemm_tc_kernel");
}
