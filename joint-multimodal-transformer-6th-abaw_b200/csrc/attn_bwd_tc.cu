// jmt_attn_bwd_dqkv_bf16: the three "small-K" GEMMs of the attention backward -- dQ = dS K, dK = dS^T Q, dV = P^T dO -- of every
// (batch, head) window in ONE launch of a dedicated tcgen05 kernel (sm_100a).
//
// Why not three batched jmt_gemm_bf16 launches (round 1 / early round 2): per window these are 300 x 512 outputs reduced over
// 300.  With the 300-extent on the MMA's M axis a window is three 128-row tiles (22 % of the tensor work is padding), the odd
// tile count forbids CTA pairs (a ghost tile measured slower), and a 128 x 256 x 300 tile stages 225 KB for 2.4 k cycles of MMA:
// the launches sat at 2x their L2->SM ingest bound (357 TFLOP/s inside the step, profiles/gemm_table_r2p.txt).
//
// Here the problem is computed TRANSPOSED, out^T (dh x n) = A^T (dh x k) X(n, k)^T:
//   * M axis = head dimension (256 or 512): whole CTA pairs (cta_group::2, M = 256), no ghost tiles; the pair's two CTAs each
//     stage their own 128 rows of A and HALF of X's columns;
//   * N axis = the whole 300-extent in ONE tile, as two MMAs per K = 16 step that share the A tile: a "big" piece of
//     N1 = 512 - r32(max(Lq, S)) columns (192 for 300) and a "small" rest (108 -> 112);
//   * TMEM: two big accumulator regions (alternating by tile) and one small one, N1 + N1 + rest <= 512 columns.  The epilogue
//     drains the small piece first, so the next tile's MMAs (other big region + the small one) start after a quarter of the
//     drain: epilogue and mainloop overlap although a tile's accumulators are 304 columns wide;
//   * a plain k-block pipeline: stage = A k-block (16 KB) + the CTA's half of both X pieces (<= 24 KB), 4 stages in flight --
//     the first versions kept a tile's A resident and streamed X piece by piece through a small ring: 79 us for the loads alone
//     (too few bytes in flight for the round trip of a slot), MMAs and stores came on top (profiles/dqkv_r2_notes.txt);
//   * bound: DRAM.  One launch at the C2 geometry reads K, Q, dO (236 MB) and dS / P (92 MB) and writes dQ, dK, dV (236 MB):
//     564 MB = 87 us at the measured 6.5 TB/s; measured 95-97 us (the three batched GEMMs it replaces: 137 us);
//   * epilogue: thread = TMEM lane = head-dimension index; it writes 32 columns as a COLUMN of a [32 n][128 m] bf16 staging
//     tile (conflict-free 2-byte stores; the four lane-quarter warps of a column half fill one tile) that one TMA store /
//     reduce-add writes into out[n][m] -- the transposition costs no shuffle and no extra pass (16 epilogue warps with a tile
//     each measured the same: the epilogue is bound by its instruction count, ~5 per element); the per-thread row sum is the bias-gradient column sum of the projection (optional).
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA), warps 4..11 = epilogue; producer / issuer run their loops
// warp-uniformly and one elected lane executes the TMA / MMA instructions (operands in uniform registers: under `if (lane == 0)`
// ptxas wraps every tcgen05.mma in a ~25-instruction ELECT / R2UR loop, which alone made the first version 1.8x slower).
// Replaces the bmm calls torch autograd issues for F.multi_head_attention_forward's two bmm's (SURVEY Q4;
// mm_multi_transformers.py:62,142-167) in the backward pass.
#include "tc_common.cuh"

namespace jmt {

constexpr int kBwThreads = 384;      // warp 0 producer, warp 1 MMA issuer, warps 4..11 epilogue
constexpr int kBwMaxKb = 5;          // k-blocks of 64 per tile: Lq, S <= 320
constexpr int kBwMaxStages = 5;      // 4 by default (5 with single-buffered epilogue staging: JMT_DQKV_STAGES=5)
constexpr int kBwStageSz = 16384 + 16384 + 8192;      // A k-block | X big piece (this CTA's half) | X small piece
constexpr int kBwEpiWarps = 8;
constexpr int kBwStgTile = 8192;                      // [32 n][128 m x 2 B] staging tile, shared by the four warps of a column half
constexpr int kBwCsumMax = 1024;     // heads * dh entries per part

struct BwPart {
  int Nn, Kk, nkb, np, trans, store_mode, last_k16;
  int n0[2], nmma[2], units[2];      // per piece (0 = big, 1 = small): first column, MMA N (multiple of 16), 32-column epilogue units
  uint32_t tx_bytes;                 // bytes one CTA lands per stage
  float alpha;
  float* colsum;
};

struct BwParams {
  BwPart part[3];
  int nparts, mtiles, heads, dh, total_tiles, csum_len;
  int n1;                        // columns of a big accumulator region (regions at TMEM columns 0 and n1; the small one starts at 2 * n1)
  FastDiv fd_mt, fd_parts, fd_heads;
  uint32_t idesc_base;           // everything but N and the B major
  unsigned long long* prof;
  int stages, sbufs, csum_stride; // pipeline stages (4 / 5), staging tiles per column half (2 / 1), floats per part in the column-sum area
  int debug;                     // JMT_DQKV_DEBUG bit mask (diagnostics only): 1 = epilogue without staging / stores, 2 = issuer without MMAs
};

struct BwTile { int mtile, part, head, b; };
__device__ __forceinline__ BwTile bw_decode(const BwParams& p, int t) {
  uint32_t q, mt, pa, b, h;
  p.fd_mt.divmod((uint32_t)t, q, mt);
  p.fd_parts.divmod(q, q, pa);
  p.fd_heads.divmod(q, b, h);
  BwTile r; r.mtile = (int)mt; r.part = (int)pa; r.head = (int)h; r.b = (int)b;
  return r;
}

__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// 32 accumulator columns of this thread's row -> [alpha *] -> row sum (valid columns only) -> bf16 -> a COLUMN of the staging tile
// (row pitch 256 B).  Warp-uniform template switches keep the loop at one FADD, half a pack and one 2-byte store per element.
template <bool SCALE, bool FULL>
__device__ __forceinline__ void stage_unit(const uint32_t (&r)[32], float alpha, int nv, uint32_t col, float& rsum) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 2) {
    float v0 = __uint_as_float(r[j]), v1 = __uint_as_float(r[j + 1]);
    if (SCALE) { v0 *= alpha; v1 *= alpha; }
    if (FULL) { s0 += v0; s1 += v1; }
    else { if (j < nv) s0 += v0; if (j + 1 < nv) s1 += v1; }
    const uint32_t pk = pack_bf16(v0, v1);
    st_shared_u16(col + j * 256, (uint16_t)(pk & 0xFFFFu));
    st_shared_u16(col + (j + 1) * 256, (uint16_t)(pk >> 16));
  }
  rsum += s0 + s1;
}

__global__ void __launch_bounds__(kBwThreads, 1)
attn_bwd_dqkv_kernel(const __grid_constant__ CUtensorMap ma0, const __grid_constant__ CUtensorMap ma1, const __grid_constant__ CUtensorMap ma2,
                     const __grid_constant__ CUtensorMap mx0, const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap mx2,
                     const __grid_constant__ CUtensorMap md0, const __grid_constant__ CUtensorMap md1, const __grid_constant__ CUtensorMap md2,
                     const BwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t sRing = base;                                   // stages x (A | X big | X small)
  const uint32_t sStage = sRing + p.stages * kBwStageSz;         // 2 halves x sbufs x [32 n x 256 B]
  const uint32_t sCsum = sStage + 2 * p.sbufs * kBwStgTile;      // 3 x csum_stride floats
  const uint32_t bars = sCsum + 3 * p.csum_stride * 4;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kBwMaxStages;
  const uint32_t t_full = empty_bar + 8 * kBwMaxStages;          // [2]: tile parity
  const uint32_t t_empty_big = t_full + 16;                      // [2]: big region
  const uint32_t t_empty_small = t_empty_big + 16;
  const uint32_t tmem_slot = t_empty_small + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));
  float* csum_sh = reinterpret_cast<float*>(smem_raw + (sCsum - base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (shfl: lets ptxas treat the warp index as warp-uniform)
  const int crank = (int)cluster_ctarank();
  const int first_tile = blockIdx.x >> 1, tile_stride = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar + 8 * s, 2); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(t_full + 8 * s, 1); mbar_init(t_empty_big + 8 * s, kBwEpiWarps * 2); }
    mbar_init(t_empty_small, kBwEpiWarps * 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * p.csum_stride; i += kBwThreads) csum_sh[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_relaxed();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer (every CTA): A k-block + this CTA's half of both X pieces per stage ================================
    const uint32_t el = elect_one();
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ma0)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mx0)) : "memory");
    if (p.nparts > 1) { asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ma1)) : "memory");
                        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mx1)) : "memory"); }
    if (p.nparts > 2) { asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ma2)) : "memory");
                        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mx2)) : "memory"); }
    int stage = 0; uint32_t phase = 0;
    long long w_e = 0; const long long t0 = p.prof ? clock64() : 0;
    const uint32_t full_leader = mapa_rank(full_bar, 0);
    for (int t = first_tile; t < p.total_tiles; t += tile_stride) {
      const BwTile c = bw_decode(p, t);
      const BwPart& P = p.part[c.part];
      const CUtensorMap* ma = c.part == 0 ? &ma0 : (c.part == 1 ? &ma1 : &ma2);
      const CUtensorMap* mx = c.part == 0 ? &mx0 : (c.part == 1 ? &mx1 : &mx2);
      const int nkb = P.nkb, np = P.np, trans = P.trans;
      const uint32_t tx = P.tx_bytes;
      const int mchunk = (c.mtile * 256 + crank * 128) >> 6;
      // this CTA's share of the two pieces: columns [n0 + crank * nmma / 2, ...), nmma / 2 of them
      const int hb = P.nmma[0] >> 1, hs = np > 1 ? P.nmma[1] >> 1 : 0;
      const int nb0 = P.n0[0] + crank * hb, ns0 = np > 1 ? P.n0[1] + crank * hs : 0;
      // MN-major (trans): 64-column chunks [64 k rows][128 B], 8 KB apart; K-major: 32-row boxes [32 n rows][64 k], 4 KB apart
      const int ldb = trans ? (hb + 63) >> 6 : (hb + 31) >> 5, lds = trans ? (hs + 63) >> 6 : (hs + 31) >> 5;
      for (int kb = 0; kb < nkb; ++kb) {
        { const long long tw = p.prof ? clock64() : 0; mbar_wait(empty_bar + 8 * stage, phase ^ 1u); if (p.prof) w_e += clock64() - tw; }
        const uint32_t fb = full_leader + 8 * stage;
        const uint32_t dst = sRing + stage * kBwStageSz;
        mbar_expect_tx_cluster_el(el, fb, tx);
        tma_load_5d_2sm_el(el, dst, ma, fb, 0, kb * 64, mchunk, c.head, c.b);
        if (trans) {
          for (int j = 0; j < ldb; ++j) tma_load_4d_2sm_el(el, dst + 16384 + j * 8192, mx, fb, nb0 + j * 64, kb * 64, c.head, c.b);
          for (int j = 0; j < lds; ++j) tma_load_4d_2sm_el(el, dst + 32768 + j * 8192, mx, fb, ns0 + j * 64, kb * 64, c.head, c.b);
        } else {
          for (int j = 0; j < ldb; ++j) tma_load_4d_2sm_el(el, dst + 16384 + j * 4096, mx, fb, kb * 64, nb0 + j * 32, c.head, c.b);
          for (int j = 0; j < lds; ++j) tma_load_4d_2sm_el(el, dst + 32768 + j * 4096, mx, fb, kb * 64, ns0 + j * 32, c.head, c.b);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    if (p.prof && lane == 0) { unsigned long long* o = p.prof + blockIdx.x * 16; o[6] = w_e; o[7] = clock64() - t0; }
  } else if (warp == 1) {
    if (crank == 0) {
      // ================================ MMA issuer (leader CTA; whole warp, one elected lane issues) ================================
      const uint32_t el = elect_one();
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;                                // tiles done by this pair
      long long w_x = 0, w_t = 0; const long long t0 = p.prof ? clock64() : 0;
      for (int t = first_tile; t < p.total_tiles; t += tile_stride, ++it) {
        const BwTile c = bw_decode(p, t);
        const BwPart& P = p.part[c.part];
        const uint32_t x_lbo = P.trans ? 8192u : 16u;
        const uint32_t x_kstep = P.trans ? 128u : 2u;
        const int nkb = P.nkb, last_k16 = P.last_k16;
        const bool two = P.np > 1;
        const uint32_t idesc_t = p.idesc_base | ((uint32_t)P.trans << 16);
        const uint32_t idesc_b = idesc_t | ((uint32_t)(P.nmma[0] >> 3) << 17);
        const uint32_t idesc_s = idesc_t | ((uint32_t)(P.nmma[1] >> 3) << 17);
        const uint32_t reg = it & 1u;
        { const long long tw = p.prof ? clock64() : 0;
          mbar_wait(t_empty_big + 8 * reg, ((it >> 1) & 1u) ^ 1u);       // this big region was drained (two tiles ago)
          mbar_wait(t_empty_small, (it & 1u) ^ 1u);                      // the small region was drained (previous tile)
          if (p.prof) w_t += clock64() - tw; }
        tc_fence_after();
        const uint32_t d_big = tmem_base + reg * p.n1, d_small = tmem_base + 2 * p.n1;
        for (int kb = 0; kb < nkb; ++kb) {
          { const long long tw = p.prof ? clock64() : 0; mbar_wait(full_bar + 8 * stage, phase); if (p.prof) w_x += clock64() - tw; }
          tc_fence_after();
          const uint32_t st = sRing + stage * kBwStageSz;
          const uint64_t a_desc = make_smem_desc(st, 8192, 1024);
          const uint64_t b_desc = make_smem_desc(st + 16384, x_lbo, 1024);
          const uint64_t s_desc = make_smem_desc(st + 32768, x_lbo, 1024);
          const int nk = kb == nkb - 1 ? last_k16 : 4;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (k < nk && !(p.debug & 2)) {
              tc_mma_elect<2>(el, d_big, a_desc + (uint64_t)(k * 128), b_desc + (uint64_t)(k * x_kstep), idesc_b, (kb > 0 || k > 0) ? 1u : 0u);
              if (two) tc_mma_elect<2>(el, d_small, a_desc + (uint64_t)(k * 128), s_desc + (uint64_t)(k * x_kstep), idesc_s, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit_elect<2>(el, empty_bar + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        tc_commit_elect<2>(el, t_full + 8 * reg);
      }
      if (p.prof && lane == 0) { unsigned long long* o = p.prof + blockIdx.x * 16; o[1] = w_x; o[2] = w_t; o[3] = clock64() - t0; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (warps 4..11) ================================
    const int q = warp & 3;                          // TMEM lane quarter
    const int ew = warp - 4;
    const int half = ew >> 2;                        // this warp takes the 32-column units u = half, half + 2, ... of a piece
    const uint32_t stg0 = sStage + half * p.sbufs * kBwStgTile;
    int sbuf = 0;
    const uint32_t big_leader = mapa_rank(t_empty_big, 0), small_leader = mapa_rank(t_empty_small, 0);
    const uint32_t el = q == 0 ? elect_one() : 0u;   // one lane of the quarter-0 warp issues / commits / waits for the column half's TMA stores
    uint32_t it = 0;
    long long w_f = 0; const long long t0 = p.prof ? clock64() : 0;
    for (int t = first_tile; t < p.total_tiles; t += tile_stride, ++it) {
      const BwTile c = bw_decode(p, t);
      const BwPart& P = p.part[c.part];
      const CUtensorMap* md = c.part == 0 ? &md0 : (c.part == 1 ? &md1 : &md2);
      const int m_cta = c.mtile * 256 + crank * 128;              // first head-dimension index of this CTA's accumulator rows
      const int m_loc = m_cta + q * 32;                           // ... and of this warp's
      const float alpha = P.alpha;
      const int Nn = P.Nn, np = P.np, store_mode = P.store_mode;
      const uint32_t reg = it & 1u;
      float rsum = 0.f;
      { const long long tw = p.prof ? clock64() : 0; mbar_wait(t_full + 8 * reg, (it >> 1) & 1u); if (p.prof) w_f += clock64() - tw; }
      tc_fence_after();
      // the small piece first: its region is what the next tile's MMAs wait for
      for (int pc = 1; pc >= 0; --pc) {
        const uint32_t rel_bar = pc == 1 ? small_leader : big_leader + 8 * reg;
        const int units = pc < np ? P.units[pc] : 0, n0 = pc < np ? P.n0[pc] : 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (pc == 1 ? 2 * p.n1 : reg * p.n1);
        // this warp's units u = half, half + 2, ...: the TMEM load of the next unit is in flight while the current one is converted,
        // staged and stored (two register buffers); the region is handed back as soon as the warp's last load has landed
        uint32_t ra[32], rb[32];
        auto release = [&]() {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(rel_bar);      // (both CTAs' warps arrive on the leader's barrier)
        };
        auto process = [&](const uint32_t (&r)[32], int u) {
          if (p.debug & 1) { rsum += __uint_as_float(r[0]); return; }
          const int n_base = n0 + u * 32;
          // the four warps of this column half (one per TMEM lane quarter) fill ONE [32 n][128 m] staging tile: 256-byte rows for the
          // TMA store instead of a 64-byte-row tile per warp
          const uint32_t stg = stg0 + sbuf * kBwStgTile;
          if (p.sbufs == 2) bulk_wait_read1_el(el); else bulk_wait_read0_el(el);      // the last store out of this buffer has read it
          asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
          const uint32_t col = stg + q * 64 + lane * 2;
          const int nv = Nn - n_base;                    // valid columns of this unit
          if (nv >= 32) { if (alpha == 1.f) stage_unit<false, true>(r, alpha, nv, col, rsum); else stage_unit<true, true>(r, alpha, nv, col, rsum); }
          else { if (alpha == 1.f) stage_unit<false, false>(r, alpha, nv, col, rsum); else stage_unit<true, false>(r, alpha, nv, col, rsum); }
          fence_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
          if (store_mode == JMT_STORE) tma_store_4d_el(el, md, stg, m_cta, n_base, c.head, c.b);
          else tma_reduce_add_4d_el(el, md, stg, m_cta, n_base, c.head, c.b);
          if (p.sbufs == 2) sbuf ^= 1;
        };
        int u = half;
        if (pc == 1) {
          // small piece (<= 4 units, <= 2 per warp): all TMEM loads first, then the region goes back at once -- the next tile's MMAs
          // wait for it (single small region), so nothing but the load latency may sit in front of the release
          const bool h0 = u < units, h1 = u + 2 < units;
          if (h0) tc_ld32_issue(taddr + u * 32, ra);
          if (h1) tc_ld32_issue(taddr + (u + 2) * 32, rb);
          if (h0) tc_wait_ld();
          release();
          if (h0) process(ra, u);
          if (h1) process(rb, u + 2);
          continue;
        }
        if (u < units) tc_ld32_issue(taddr + u * 32, ra);
        else release();
        while (u < units) {
          tc_wait_ld();
          if (u + 2 < units) tc_ld32_issue(taddr + (u + 2) * 32, rb); else release();
          process(ra, u);
          u += 2;
          if (u >= units) break;
          tc_wait_ld();
          if (u + 2 < units) tc_ld32_issue(taddr + (u + 2) * 32, ra); else release();
          process(rb, u);
          u += 2;
        }
      }
      if (P.colsum != nullptr) atomicAdd(csum_sh + c.part * p.csum_stride + c.head * p.dh + m_loc + lane, rsum);
    }
    bulk_wait0_el(el);
    // flush the bias-gradient column sums: one global atomic per entry and CTA
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kBwEpiWarps) : "memory");
    for (int pa = 0; pa < p.nparts; ++pa) {
      float* dst = p.part[pa].colsum;
      if (dst == nullptr) continue;
      for (int i = threadIdx.x - 128; i < p.csum_len; i += 32 * kBwEpiWarps) {
        const float v = csum_sh[pa * p.csum_stride + i];
        if (v != 0.f) atomicAdd(dst + i, v);
      }
    }
    if (p.prof && ew == 0 && lane == 0) { unsigned long long* o = p.prof + blockIdx.x * 16; o[4] = w_f; o[5] = clock64() - t0; }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_relaxed();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// transposed-output tensor map: out(n, m) of (head, batch) at ptr + b*bs + h*hs + n*ld + m; box {128 m, 32 n}, no swizzle
static int make_map_dt(CUtensorMap* map, const void* ptr, int64_t inner, int64_t rows, int64_t ld, int64_t nb0, int64_t bs0,
                       int64_t nb1, int64_t bs1, const char* who) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("%s: cuTensorMapEncodeTiled unavailable (no CUDA driver?)", who); return JMT_ERR_CUDA; }
  JMT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0 && (nb0 == 1 || bs0 % 8 == 0) && (nb1 == 1 || bs1 % 8 == 0),
              "%s: output geometry must be 16-byte aligned", who);
  const cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nb0, (cuuint64_t)nb1};
  const cuuint64_t row_bytes = (cuuint64_t)ld * 2;
  const cuuint64_t strides[3] = {row_bytes, nb0 > 1 ? (cuuint64_t)bs0 * 2 : row_bytes, nb1 > 1 ? (cuuint64_t)bs1 * 2 : row_bytes};
  const cuuint32_t box[4] = {128, 32, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("%s: cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld ld=%lld", who, (int)r, (long long)inner, (long long)rows, (long long)ld);
    return JMT_ERR_CUDA;
  }
  return JMT_OK;
}

static bool bw_geometry_ok(const jmt_attn_bwd_desc* g) {
  if (!g) return false;
  if (g->dh != 256 && g->dh != 512) return false;
  if (g->Lq < 1 || g->S < 1 || g->Lq > 64 * kBwMaxKb || g->S > 64 * kBwMaxKb) return false;
  if (g->heads < 1 || g->NB < 1 || (int64_t)g->heads * g->dh > kBwCsumMax) return false;
  if (g->x_ld % 8 != 0 || g->x_ld < g->S) return false;
  return true;
}

}  // namespace jmt

using namespace jmt;

static std::atomic<unsigned long long*> g_bw_prof{nullptr};
extern "C" int jmt_attn_bwd_set_profile_buffer(void* dev_buf) {
  g_bw_prof.store((unsigned long long*)dev_buf);
  return JMT_OK;
}

extern "C" int jmt_attn_bwd_dqkv_supported(const jmt_attn_bwd_desc* g) { return bw_geometry_ok(g) ? 1 : 0; }

extern "C" int jmt_attn_bwd_dqkv_bf16(const jmt_attn_bwd_desc* g, void* stream) {
  JMT_REQUIRE(g, "jmt_attn_bwd_dqkv_bf16: null descriptor");
  if (!bw_geometry_ok(g)) {
    set_error("jmt_attn_bwd_dqkv_bf16: unsupported geometry (dh=%d Lq=%d S=%d heads=%d x_ld=%lld): use jmt_gemm_bf16", g->dh, g->Lq, g->S,
              g->heads, (long long)g->x_ld);
    return JMT_ERR_UNSUPPORTED;
  }
  BwParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap ma[3], mx[3], md[3];
  // accumulator geometry common to all parts: big slots of n1 columns at TMEM columns 0 and n1, the small slot behind them
  const int lmax = g->Lq > g->S ? g->Lq : g->S;
  const int rmax = (lmax + 31) / 32 * 32;
  p.n1 = rmax <= 256 ? 256 : 512 - rmax;
  int np_ = 0;
  for (int i = 0; i < 3; ++i) {
    const jmt_attn_bwd_part& s = g->part[i];
    if (s.d == nullptr) continue;
    JMT_REQUIRE(s.a && s.x, "jmt_attn_bwd_dqkv_bf16: part %d: null operand", i);
    JMT_REQUIRE(s.store_mode == JMT_STORE || s.store_mode == JMT_ACCUMULATE, "jmt_attn_bwd_dqkv_bf16: part %d: bad store_mode", i);
    BwPart& P = p.part[np_];
    P.trans = s.x_trans ? 1 : 0;
    P.Nn = P.trans ? g->S : g->Lq;
    P.Kk = P.trans ? g->Lq : g->S;
    P.nkb = (P.Kk + 63) / 64;
    P.last_k16 = (P.Kk - 64 * (P.nkb - 1) + 15) / 16;
    // pieces: one big piece when the extent fits a big slot, else big + small; an MMA's N covers the valid columns rounded up to 16
    P.np = P.Nn <= p.n1 ? 1 : 2;
    for (int pc = 0; pc < P.np; ++pc) {
      P.n0[pc] = pc * p.n1;
      const int cols = (pc == P.np - 1 ? P.Nn : p.n1) - P.n0[pc];
      P.nmma[pc] = (cols + 15) / 16 * 16;
      P.units[pc] = (cols + 31) / 32;
    }
    {
      const int hb = P.nmma[0] / 2, hs = P.np > 1 ? P.nmma[1] / 2 : 0;
      P.tx_bytes = 16384u + (P.trans ? (uint32_t)(((hb + 63) / 64 + (hs + 63) / 64) * 8192) : (uint32_t)(((hb + 31) / 32 + (hs + 31) / 32) * 4096));
    }
    P.store_mode = s.store_mode;
    P.alpha = s.alpha;
    P.colsum = s.colsum;
    int rc = make_map_mn5(&ma[np_], s.a, g->dh, P.Kk, s.a_ld, g->heads, s.a_hs, g->NB, s.a_bs, 64, 2, "jmt_attn_bwd_dqkv_bf16(A)");
    if (rc != JMT_OK) return rc;
    // X: MN-major parts fetch [64 k rows][64 columns] chunks, the K-major part [32 n rows][64 k] boxes
    rc = make_map(&mx[np_], s.x, g->S, g->Lq, g->x_ld, g->heads, (int64_t)g->Lq * g->x_ld, g->NB, (int64_t)g->heads * g->Lq * g->x_ld,
                  P.trans ? 64 : 32, "jmt_attn_bwd_dqkv_bf16(X)");
    if (rc != JMT_OK) return rc;
    rc = make_map_dt(&md[np_], s.d, g->dh, P.Nn, s.d_ld, g->heads, s.d_hs, g->NB, s.d_bs, "jmt_attn_bwd_dqkv_bf16(D)");
    if (rc != JMT_OK) return rc;
    ++np_;
  }
  if (np_ == 0) return JMT_OK;
  for (int i = np_; i < 3; ++i) { ma[i] = ma[0]; mx[i] = mx[0]; md[i] = md[0]; }
  p.nparts = np_;
  p.mtiles = g->dh / 256;
  p.heads = g->heads;
  p.dh = g->dh;
  p.csum_len = g->heads * g->dh;
  const int64_t total = (int64_t)g->NB * g->heads * np_ * p.mtiles;
  JMT_REQUIRE(total < (1ll << 31), "jmt_attn_bwd_dqkv_bf16: too many tiles");
  p.total_tiles = (int)total;
  p.fd_mt.init(p.mtiles); p.fd_parts.init(np_); p.fd_heads.init(g->heads);
  // fp32 accumulate, bf16 A / B, A MN-major, M = 256 (cta_group::2); N and the B major are set per piece
  p.idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(256 >> 4) << 24);
  p.prof = g_bw_prof.load();
  { static const int dbg = []() { const char* e = getenv("JMT_DQKV_DEBUG"); return e ? atoi(e) : 0; }(); p.debug = dbg; }

  // shared memory plan: 4 pipeline stages + two staging tiles per column half.  (JMT_DQKV_STAGES=5: 5 stages + one staging tile when
  // the column-sum area is small -- measured slower, 102 vs 97 us: the single staging tile makes the epilogue the bottleneck.)
  p.csum_stride = (p.csum_len + 63) / 64 * 64;
  {
    static const int st_env = []() { const char* e = getenv("JMT_DQKV_STAGES"); return e ? atoi(e) : 4; }();
    const int budget = 227 * 1024 - 512 - 3 * p.csum_stride * 4;
    p.stages = 4; p.sbufs = 2;
    if (st_env == 5 && 5 * kBwStageSz + 2 * kBwStgTile <= budget) { p.stages = 5; p.sbufs = 1; }
  }
  const int smem = p.stages * kBwStageSz + 2 * p.sbufs * kBwStgTile + 3 * p.csum_stride * 4 + 512;
  static std::atomic<int> attr_set[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); set_error("jmt_attn_bwd_dqkv_bf16: no CUDA device"); return JMT_ERR_CUDA; }
  if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dqkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("jmt_attn_bwd_dqkv_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return JMT_ERR_CUDA; }
    attr_set[dev & 63].store(1, std::memory_order_release);
  }
  const int max_pairs = kNumSMs / 2;
  const int pairs = p.total_tiles < max_pairs ? p.total_tiles : max_pairs;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(pairs * 2);
  cfg.blockDim = dim3(kBwThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = []() { const char* e = getenv("JMT_PDL"); return e ? atoi(e) != 0 : true; }();
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, attn_bwd_dqkv_kernel, ma[0], ma[1], ma[2], mx[0], mx[1], mx[2], md[0], md[1], md[2], p);
  if (le != cudaSuccess) {
    set_error("jmt_attn_bwd_dqkv_bf16: cudaLaunchKernelEx: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JMT_ERR_CUDA;
  }
  return check_launch("attn_bwd_dqkv_kernel");
}
