// jmt_gemm_bf16: persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   warp 0 : TMA producer   -- cp.async.bulk.tensor.4d / 5d -> 128B-swizzled smem ring
//   warp 1 : MMA issuer     -- tcgen05.mma.cta_group::{1,2}.kind::f16, fp32 accum in TMEM
//   warps 2..9 : epilogue   -- tcgen05.ld 32x32b -> alpha/bias/activation -> swizzled smem -> TMA store
// The producer and issuer warps run their loops with ALL lanes (warp-uniform control flow, operands in uniform registers) and ONE
// elected lane executes the TMA / MMA / commit instructions (tc_common.cuh: *_el, tc_mma_elect): under `if (lane == 0)` ptxas wraps
// every such instruction in an ELECT + R2UR.BROADCAST loop of ~25 single-thread instructions -- as long as an N = 256 MMA takes.
//
// Two TMEM accumulator stages (2 x 256 columns) let the epilogue of tile i overlap the mainloop of
// tile i+1; an epilogue warp hands its stage back as soon as its last tcgen05.ld of the tile has landed.  One descriptor (jmt_gemm_desc) covers every dense contraction of the JMT path: Linear
// fwd/dgrad/wgrad, QK^T / PV and their gradients (batched over (b, head) through 4-D tensor maps,
// K- or MN-major operands selected in the UMMA instruction descriptor) and the dilated causal
// Conv1d of the TCN as an implicit GEMM (taps = extra K blocks with a shifted TMA row coordinate;
// causal zero padding = TMA out-of-bounds fill).
//
// kCta == 2 (template): CTA pairs (cluster of 2) run one tcgen05.mma.cta_group::2 with M = 256: each CTA
// owns 128 accumulator rows (its own A tile, possibly from another batch element when B is shared across
// the batch) and stages only HALF of the B tile -- the pair's tensor cores read both halves.  That cuts
// the L2->SMEM operand traffic per FLOP by a third (48 KB -> 32 KB per 128x256x64 block), which is the
// measured limit of the 1-CTA kernel (LTS ~12 TB/s => 1.05 PFLOP/s).
// Wide pair tiles (TcParams::wide, N % 512 == 0, >= 8 k-iterations): the pair computes 256 x 512 with TWO N=256 MMAs per k-step
// that share the A tile (TMEM columns 0-255 and 256-511, i.e. a single accumulator stage): 24 KB per 128x256x64 block and CTA.
// The kernel is bound by what an SM can ingest from L2 (~44 B/clk), so bytes per FLOP decide: K=3072 linear 1150 -> 1437 TFLOP/s.
#include "tc_common.cuh"

int jmt_validate_gemm_desc(const jmt_gemm_desc* g, const char* who);

namespace jmt {

// epilogue warps per CTA (template parameter kEpi): 8 = two per TMEM lane quarter, 16 = four per quarter (interleaved over the
// tile's column chunks).  The epilogue is latency-bound (tcgen05.ld -> math -> staging -> TMA store, one chain per warp), so
// short-K tiles, whose mainloop is shorter than one warp-pair's drain of the accumulator, want the 16-warp kernel.
constexpr int kAStageBytes = kBlockM * kBlockK * 2;   // 16 KiB
constexpr int kMaxStages = 8;
constexpr int kWideMinIters = 8;       // k-iterations per tile from which the 256 x 512 pair tile pays (JMT_GEMM_WIDE=<n> overrides)
constexpr int kAccStride = 256;       // TMEM columns per accumulator stage
constexpr int kEpiStageBytes = 4096;      // 32 rows x 128 B per epilogue warp
constexpr int kBiasFlagBytes = 2048 + 256;   // 512-float bias tile (two 256-float buffers for narrow tiles) + 2 x 16 keep-flag words
constexpr int epi_smem_bytes(int epi_warps) { return epi_warps * kEpiStageBytes + kBiasFlagBytes; }

struct TcParams {
  int M, N, K, block_n;
  int m_tiles, n_tiles, batch_tiles, split_k, total_tiles;
  int kblocks, rb_n, iters_total;
  int nb0;
  FastDiv fd_ntiles, fd_mpairs, fd_bslots, fd_nb0, fd_kblocks, fd_rbn;
  int a_major, b_major;
  int a_shift0, a_shift_step, b_shift0, b_shift_step;
  int b_has_b0, b_has_b1;
  int a_has_b0, a_has_b1;   // 0: A is shared by every entry of that batch dimension (stride 0), its coordinate stays 0
  int a_mn5, b_mn5;    // MN-major operand tile fetched by ONE 5-D TMA instruction (all its 64-wide chunks) instead of one per chunk
  int reduce_batch;
  uint32_t idesc;
  int b_stage_bytes, b_tx_bytes, b_chunks_cta, stages;
  // epilogue
  void* d;
  const float* bias;
  const uint8_t* colmask;     // per (sample, column) keep-mask applied after the activation (nullable)
  float colmask_scale;
  int colmask_period;         // > 0: sample = output row / period (flat layout); 0: sample = batch index
  int colmask_samples;        // number of mask rows (flat layout)
  int zrow_period, zrow_count; // output rows m with (m % period) < count are written as zeros (period 0: off)
  FastDiv fd_zrow, fd_cmask;
  int64_t d_ld, d_bs0, d_bs1;
  float alpha, slope;
  int d_dtype, act, store_mode, vec_ok;
  int epi_warps;      // 8 or 16 (kernel template parameter kEpi)
  int wide;           // CTA pairs only: 256 x 512 tile = two N=256 MMAs per k-step sharing the A tile, ONE accumulator stage
                      // (all 512 TMEM columns).  Per 128x256x64 block a CTA stages 24 KB instead of 32 KB: for long-K tiles
                      // (L2->SM ingest-bound) that outweighs the lost epilogue / mainloop overlap
  int cluster;        // 1, or 2 = CTA pairs issuing cta_group::2 MMAs (each CTA stages half of the B tile)
  int pair_batch;     // cluster 2 only: 0 = the pair covers two consecutive M tiles, 1 = two consecutive batch entries
  int m_pairs;        // number of M tile slots per (n, batch): ceil(m_tiles / 2) when pairing along M, else m_tiles
  int batch_slots;    // batch_tiles, or batch_tiles / 2 when pairing along the batch
  // Tail split (wave quantisation): tiles are dealt round-robin to `units` CTAs / CTA pairs, so 300 tiles on 74 pairs take 5 rounds
  // although they are 4.05 rounds of work.  The tiles of the last, nearly empty round (t >= tail_start in regular numbering) are
  // cut along N into tail_split pieces of tail_bn columns each (own UMMA instruction descriptor, own B tensor map with a smaller
  // box), so that round costs a fraction of a tile instead of a whole one.
  // Extended backward epilogue (bf16 D through TMA, no activation): D_out = value * act'(aux) and / or column sums of D_out
  const __nv_bfloat16* aux;   // same element offsets as D (ld, batch strides); value *= aux > 0 ? 1 : aux_slope   (nullable)
  float aux_slope;
  float* colsum;              // colsum[b0 * colsum_bs0 + n] += sum over the rows this launch writes (fp32 atomics; nullable)
  int64_t colsum_bs0;
  int csum_smem;              // the whole column-sum vector (<= 512 entries) is accumulated in shared memory (the bias-tile area: these
                              // launches have no bias) and flushed with one global atomic per entry and CTA at kernel end, instead
                              // of one per (warp, column, tile) -- 1.5 M -> 75 k atomics per attention-backward launch
  int csum_len;
  int two_phase;      // wide tiles: the two 256-column halves of the accumulator are handed back separately (tempty[0] / tempty[1]):
                      // the epilogue drains half 0 first, and the next tile's MMAs into half 0 run while half 1 is still drained
  int tail_start, tail_split, tail_bn, b_tail_bytes;
  uint32_t idesc_tail;
  FastDiv fd_tsplit;
  unsigned long long* prof;   // optional per-CTA cycle counters (16 per CTA), see jmt_gemm_set_profile_buffer
  int tma_store;      // epilogue through swizzled smem + TMA store / reduce-add (needs 16-byte aligned D geometry)
  int l2_ahead, l2_mstep;   // bres: L2 prefetch distance in tiles of this pair, and the M distance of two consecutive tiles of a pair
  int bres;           // B-stationary pairs (K <= 512, N % 256 == 0, no taps / batches): every CTA pair keeps ITS 256-column slice of B (all
                      // k-blocks, <= 128 KB per CTA) resident in shared memory for the whole launch and streams only A.  The K = 512 linears are
                      // bound by what an SM ingests from L2 (narrow tiles: 32 KB per 128x256x64 block, 60 % of the issuer's time waiting for
                      // operands) or by the un-overlapped epilogue (wide tiles, one accumulator stage); with B resident a block costs 16 KB and the
                      // two 256-column accumulator stages keep the epilogue under the next tile's mainloop
};

template <int ACT>
__device__ __forceinline__ float act_t(float x, float slope) {
  if constexpr (ACT == JMT_ACT_RELU) return fmaxf(x, 0.f);
  else if constexpr (ACT == JMT_ACT_LEAKY_RELU) return x >= 0.f ? x : x * slope;
  else return x;
}
// 32 accumulator columns -> act(alpha * acc + bias), branch-free (the activation is a template parameter so the
// compiler can interleave the 32 independent FFMA / FMNMX / F2FP chains)
// 32 accumulator columns -> act(alpha * acc + bias) [* column scale] -> bf16 -> this thread's staging row
// (16-byte pieces `piece0 .. piece0+3` of the 128-byte row, 128B-swizzled).  Branch-free: the activation is a template
// parameter so the compiler interleaves the independent FFMA / FMNMX / F2FP chains; nothing but r[] stays live.
// keep-flags of this thread's sample: one BIT per column, one 32-bit word per 32-column group (behind the 512-float bias tile):
// x *= flag ? scale : 0
__device__ __forceinline__ void apply_flags4(uint32_t fbits, int j, float mscale, float& x0, float& x1, float& x2, float& x3) {
  x0 = (fbits >> j) & 1u ? x0 * mscale : 0.f;
  x1 = (fbits >> (j + 1)) & 1u ? x1 * mscale : 0.f;
  x2 = (fbits >> (j + 2)) & 1u ? x2 * mscale : 0.f;
  x3 = (fbits >> (j + 3)) & 1u ? x3 * mscale : 0.f;
}

template <int ACT, bool MASK>
__device__ __forceinline__ void epi_math_bf16(const uint32_t (&r)[32], const float* bias, uint32_t fbits, float mscale,
                                              uint32_t keep, float alpha, float slope, uint32_t* pk /*16 packed words*/) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 bv = *reinterpret_cast<const float4*>(bias + j);
    float x0 = act_t<ACT>(fmaf(alpha, __uint_as_float(r[j]), bv.x), slope);
    float x1 = act_t<ACT>(fmaf(alpha, __uint_as_float(r[j + 1]), bv.y), slope);
    float x2 = act_t<ACT>(fmaf(alpha, __uint_as_float(r[j + 2]), bv.z), slope);
    float x3 = act_t<ACT>(fmaf(alpha, __uint_as_float(r[j + 3]), bv.w), slope);
    if constexpr (MASK) apply_flags4(fbits, j, mscale, x0, x1, x2, x3);
    // keep = 0 for rows that must read back as zeros (padding rows of the flat TCN layout), else all ones
    pk[j / 2] = pack_bf16(x0, x1) & keep;
    pk[j / 2 + 1] = pack_bf16(x2, x3) & keep;
  }
}
template <int ACT, bool MASK>
__device__ __forceinline__ void epi_math_f32(uint32_t (&r)[32], const float* bias, uint32_t fbits, float mscale, uint32_t keep,
                                             float alpha, float slope) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 bv = *reinterpret_cast<const float4*>(bias + j);
    float x0 = act_t<ACT>(fmaf(alpha, __uint_as_float(r[j]), bv.x), slope);
    float x1 = act_t<ACT>(fmaf(alpha, __uint_as_float(r[j + 1]), bv.y), slope);
    float x2 = act_t<ACT>(fmaf(alpha, __uint_as_float(r[j + 2]), bv.z), slope);
    float x3 = act_t<ACT>(fmaf(alpha, __uint_as_float(r[j + 3]), bv.w), slope);
    if constexpr (MASK) apply_flags4(fbits, j, mscale, x0, x1, x2, x3);
    r[j] = __float_as_uint(x0) & keep; r[j + 1] = __float_as_uint(x1) & keep; r[j + 2] = __float_as_uint(x2) & keep; r[j + 3] = __float_as_uint(x3) & keep;
  }
}

struct TileCoord { int m0, n0, batch, split, it0, it1, bn, tail; };   // bn: columns of this tile (block_n, or tail_bn for a tail piece)

__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int t, int crank) {
  TileCoord c;
  uint32_t q, nt, mp, bs;
  uint32_t sub = 0;
  c.tail = t >= p.tail_start ? 1 : 0;
  c.bn = c.tail ? p.tail_bn : p.block_n;
  if (c.tail) {                // piece `sub` of regular tile tail_start + (t - tail_start) / tail_split
    uint32_t u;
    p.fd_tsplit.divmod((uint32_t)(t - p.tail_start), u, sub);
    t = p.tail_start + (int)u;
  }
  p.fd_ntiles.divmod((uint32_t)t, q, nt);
  p.fd_mpairs.divmod(q, q, mp);
  uint32_t split;
  p.fd_bslots.divmod(q, split, bs);
  c.split = (int)split;
  if (p.pair_batch) {          // the two CTAs of a pair work on consecutive batch entries (B is not batched)
    c.m0 = (int)mp * kBlockM;
    c.batch = (int)bs * 2 + crank;
  } else {                     // consecutive M tiles; an odd tail pair gives rank 1 an all-out-of-range (ghost) tile
    c.m0 = ((int)mp * p.cluster + crank) * kBlockM;
    c.batch = (int)bs;
  }
  c.n0 = (int)nt * p.block_n + (int)sub * p.tail_bn;
  if (p.split_k == 1) { c.it0 = 0; c.it1 = p.iters_total; }
  else {
    c.it0 = (int)(((int64_t)c.split * p.iters_total) / p.split_k);
    c.it1 = (int)(((int64_t)(c.split + 1) * p.iters_total) / p.split_k);
  }
  return c;
}

struct EpiCtx {
  uint32_t tbase, stage_smem, row_smem, sw;
  uint32_t el;                // 1 in the warp's elected lane (issues / commits / waits for this warp's TMA stores), else 0
  const float* bias;          // this tile's bias slice in shared memory (zeros when there is none)
  const uint32_t* fwords;     // MASK: this thread's sample's keep-flag words (one bit per column of the tile, shared memory)
  uint32_t keep;              // 0 when this thread's output row must be written as zeros, else ~0
#ifdef JMT_EPI_PROF
  long long* prof;
#endif
  uint32_t tempty;            // accumulator-stage 'empty' barrier (leader CTA's, cluster address when kCta == 2), released at the end
  uint32_t tempty_mid;        // two-phase wide tile: barrier of accumulator half 0, released once columns [0, 256) are drained (0: none)
  uint32_t tempty_end2;       // a second barrier to release at the end (tail piece inside a two-phase launch), 0: none
  int m0w, n0, b0, b1, part, parts, lane;   // part / parts: this warp's interleaved share of the tile's column chunks
  int bn;                                   // columns of this tile
  float* csum_sh;                           // shared-memory column-sum accumulators (p.csum_smem)
};

// One epilogue warp's share of a 128 x block_n accumulator tile: its 32 TMEM lanes (rows) x every other
// 64-column (bf16 out) / 32-column (fp32 out) chunk -> alpha / bias / activation -> 128B-swizzled staging tile ->
// TMA store or reduce-add; or per-thread stores when D's geometry is not 16-byte aligned.
#ifdef JMT_EPI_PROF
// diagnostic build only (profiles/tools/epi_prof.sh): cycles of epilogue warp 0 of CTA 0 per phase, accumulated in registers
__device__ unsigned long long g_epi_prof[8];
#define EPI_T(i) do { const long long t_ = clock64(); e.prof[i] += t_ - tprev; tprev = t_; } while (0)
#define EPI_T0() long long tprev = clock64()
#else
#define EPI_T(i) do { } while (0)
#define EPI_T0() do { } while (0)
#endif

template <int kCta>
__device__ __forceinline__ void epi_release(uint32_t tempty, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) {
    if constexpr (kCta == 1) mbar_arrive(tempty);
    else mbar_arrive_cluster(tempty);      // the leader's MMA issuer waits for both CTAs' epilogues
  }
}

// returns true when the accumulator stage has already been released (epi_release) inside
template <int ACT, bool MASK, int kCta>
__device__ __forceinline__ bool epi_tile(const TcParams& p, const CUtensorMap* tma_d, const EpiCtx& e) {
  const int lane = e.lane;
  bool released = false;
  if (p.tma_store) {
    const bool warp_rows_valid = e.m0w < p.M;      // warp-uniform
    if (p.d_dtype == JMT_BF16) {
      // one 64-column chunk (a [32 rows x 128 B] staging tile) per iteration, in two 32-column halves so that only
      // r[32] + pk[16] are live (the 16-warp kernel has 112 registers per thread): the second half's tcgen05.ld is in
      // flight while the first half waits for the previous TMA store and is written to the staging tile
      for (int c0 = e.part * 64; c0 < e.bn; c0 += 64 * e.parts) {
        if (e.n0 + c0 >= p.N) break;
        const bool second = c0 + 32 < e.bn;      // block_n is a multiple of 32
        uint32_t r[32], pk[16];
        EPI_T0();
        tc_ld32_issue(e.tbase + c0, r);
        tc_wait_ld();
        EPI_T(0);
        epi_math_bf16<ACT, MASK>(r, e.bias + c0, MASK ? e.fwords[c0 >> 5] : 0u, p.colmask_scale, e.keep, p.alpha, p.slope, pk);
        if (second) tc_ld32_issue(e.tbase + c0 + 32, r);
        EPI_T(1);
        bulk_wait_read0_el(e.el);       // previous TMA store has finished reading the staging tile
        __syncwarp();
        EPI_T(2);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          st_shared_v4(e.row_smem + (((uint32_t)ch ^ e.sw) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        if (second) tc_wait_ld();
        EPI_T(3);
        // last chunk of this warp's share: every TMEM read of the tile has landed in registers, hand the accumulator stage
        // back to the MMA issuer before the remaining math / staging / store
        const int c_next = c0 + 64 * e.parts;
        if (e.tempty_mid != 0u && c0 < 256 && c_next >= 256) epi_release<kCta>(e.tempty_mid, lane);   // accumulator half 0 drained
        if (c_next >= e.bn || e.n0 + c_next >= p.N) { epi_release<kCta>(e.tempty, lane); released = true; }
        if (second) {
          epi_math_bf16<ACT, MASK>(r, e.bias + c0 + 32, MASK ? e.fwords[(c0 >> 5) + 1] : 0u, p.colmask_scale, e.keep, p.alpha, p.slope, pk);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          st_shared_v4(e.row_smem + (((uint32_t)(ch + 4) ^ e.sw) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
        EPI_T(4);
        fence_async_smem();
        __syncwarp();
        {
          const uint32_t elv = warp_rows_valid ? e.el : 0u;
          if (p.store_mode == JMT_STORE) tma_store_4d_el(elv, tma_d, e.stage_smem, e.n0 + c0, e.m0w, e.b0, e.b1);
          else tma_reduce_add_4d_el(elv, tma_d, e.stage_smem, e.n0 + c0, e.m0w, e.b0, e.b1);
        }
        EPI_T(5);
      }
    } else {
      for (int c0 = e.part * 32; c0 < e.bn; c0 += 32 * e.parts) {
        if (e.n0 + c0 >= p.N) break;
        uint32_t r[32];
        tc_ld32(e.tbase + c0, r);
        const int c_next = c0 + 32 * e.parts;
        if (e.tempty_mid != 0u && c0 < 256 && c_next >= 256) epi_release<kCta>(e.tempty_mid, lane);   // accumulator half 0 drained
        if (c_next >= e.bn || e.n0 + c_next >= p.N) { epi_release<kCta>(e.tempty, lane); released = true; }
        epi_math_f32<ACT, MASK>(r, e.bias + c0, MASK ? e.fwords[c0 >> 5] : 0u, p.colmask_scale, e.keep, p.alpha, p.slope);
        bulk_wait_read0_el(e.el);
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          st_shared_v4(e.row_smem + ((ch ^ e.sw) << 4), r[4 * ch], r[4 * ch + 1], r[4 * ch + 2], r[4 * ch + 3]);
        fence_async_smem();
        __syncwarp();
        {
          const uint32_t elv = warp_rows_valid ? e.el : 0u;
          if (p.store_mode == JMT_STORE) tma_store_4d_el(elv, tma_d, e.stage_smem, e.n0 + c0, e.m0w, e.b0, e.b1);
          else tma_reduce_add_4d_el(elv, tma_d, e.stage_smem, e.n0 + c0, e.m0w, e.b0, e.b1);
        }
      }
    }
  } else {
    // direct (unaligned D geometry): per-thread row stores / atomics
    const int m = e.m0w + lane;
    const int64_t row_off = (int64_t)e.b0 * p.d_bs0 + (int64_t)e.b1 * p.d_bs1 + (int64_t)m * p.d_ld;
    for (int c0 = e.part * 32; c0 < e.bn; c0 += 32 * e.parts) {
      const int n = e.n0 + c0;
      if (n >= p.N) break;                      // warp-uniform
      uint32_t r[32];
      tc_ld32(e.tbase + c0, r);
      if (e.tempty_mid != 0u && c0 < 256 && c0 + 32 * e.parts >= 256) epi_release<kCta>(e.tempty_mid, lane);
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (n + j >= p.N) break;
        float v = act_t<ACT>(fmaf(p.alpha, __uint_as_float(r[j]), e.bias[c0 + j]), p.slope);
        if constexpr (MASK) v = (e.fwords[c0 >> 5] >> j) & 1u ? v * p.colmask_scale : 0.f;
        if (e.keep == 0u) v = 0.f;
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
  return released;
}

// Column sums of a 32 x 32 block held row-wise (thread = row, v[j] = column j): butterfly transpose-reduce, 31 shuffles; on
// return lane l holds the sum of column l in v[0].
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, cnt = 32; off >= 1; off >>= 1, cnt >>= 1) {
    const bool up = (lane & off) != 0;
    const int half = cnt >> 1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (i < half) {
        const float send = up ? v[i] : v[i + half];
        const float keep = up ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
  }
  return v[0];
}

// Extended epilogue of the backward GEMMs (bf16 D through TMA store / reduce-add, no activation):
//   value = alpha * acc (+ bias) [* channel keep-flag scale] [* (aux > 0 ? 1 : aux_slope)]      <- activation gradient folded in
//   colsum[n] += sum_m value(m, n)                                                               <- bias gradient of the producer
// so that neither the act-bwd pass over the gradient nor the column-sum pass runs as a separate kernel (SURVEY 8a kernel column).
// aux is read straight from global memory (64 contiguous bytes per thread and 32-column half).
struct ExtRow {
  const __nv_bfloat16* aux_row;   // aux + this thread's row offset (nullptr: no fold)
  float* csum;                    // colsum + batch offset (nullptr: no column sums); shared-memory accumulators when p.csum_smem
  bool row_valid, zero_row;
};

// 32 accumulator columns [cb, cb + 32) of this thread's row: r -> x (math), fold, pack into pk[16], column sums
template <bool MASK, bool CSUM>
__device__ __forceinline__ void ext_half(const TcParams& p, const EpiCtx& e, const ExtRow& xr, const uint32_t (&r)[32], const uint32_t (&aw)[16],
                                         bool ax_vec, int cb, uint32_t (&pk)[16]) {
  const float* bias = e.bias + cb;
  const bool has_bias = p.bias != nullptr;        // (without a bias the bias-tile area may hold the column-sum accumulators)
  const uint32_t fbits = MASK ? e.fwords[cb >> 5] : 0u;
  if constexpr (!CSUM) {
    // fold only: stream pairs of columns straight from the accumulator registers into the packed output (no x[32] array)
    const bool fold_vec = xr.aux_row != nullptr && xr.row_valid && ax_vec;
    const bool fold_el = xr.aux_row != nullptr && xr.row_valid && !ax_vec;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      float x0 = has_bias ? fmaf(p.alpha, __uint_as_float(r[j]), bias[j]) : p.alpha * __uint_as_float(r[j]);
      float x1 = has_bias ? fmaf(p.alpha, __uint_as_float(r[j + 1]), bias[j + 1]) : p.alpha * __uint_as_float(r[j + 1]);
      if constexpr (MASK) {
        x0 = (fbits >> j) & 1u ? x0 * p.colmask_scale : 0.f;
        x1 = (fbits >> (j + 1)) & 1u ? x1 * p.colmask_scale : 0.f;
      }
      if (fold_vec) {
        const float y0 = __uint_as_float(aw[j >> 1] << 16), y1 = __uint_as_float(aw[j >> 1] & 0xFFFF0000u);
        x0 = y0 > 0.f ? x0 : x0 * p.aux_slope;
        x1 = y1 > 0.f ? x1 : x1 * p.aux_slope;
      } else if (fold_el) {
        if (e.n0 + cb + j < p.N) { const float y = __bfloat162float(xr.aux_row[e.n0 + cb + j]); x0 = y > 0.f ? x0 : x0 * p.aux_slope; }
        if (e.n0 + cb + j + 1 < p.N) { const float y = __bfloat162float(xr.aux_row[e.n0 + cb + j + 1]); x1 = y > 0.f ? x1 : x1 * p.aux_slope; }
      }
      pk[j >> 1] = xr.zero_row ? 0u : pack_bf16(x0, x1);
    }
    return;
  }
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float v = has_bias ? fmaf(p.alpha, __uint_as_float(r[j]), bias[j]) : p.alpha * __uint_as_float(r[j]);
    if constexpr (MASK) v = (fbits >> j) & 1u ? v * p.colmask_scale : 0.f;
    x[j] = v;
  }
  if (xr.aux_row != nullptr && xr.row_valid) {
    if (ax_vec) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float y0 = __uint_as_float(aw[j >> 1] << 16), y1 = __uint_as_float(aw[j >> 1] & 0xFFFF0000u);
        x[j] = y0 > 0.f ? x[j] : x[j] * p.aux_slope;
        x[j + 1] = y1 > 0.f ? x[j + 1] : x[j + 1] * p.aux_slope;
      }
    } else {                                       // ragged N tail / unaligned: element-wise
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (e.n0 + cb + j < p.N) {
          const float y = __bfloat162float(xr.aux_row[e.n0 + cb + j]);
          x[j] = y > 0.f ? x[j] : x[j] * p.aux_slope;
        }
    }
  }
  if (xr.zero_row) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 32; j += 2) pk[j >> 1] = pack_bf16(x[j], x[j + 1]);
  if constexpr (CSUM) {
    if (xr.csum != nullptr) {
      const float cs = warp_colsum32(x, e.lane);
      if (e.n0 + cb + e.lane < p.N && cs != 0.f) atomicAdd(xr.csum + e.n0 + cb + e.lane, cs);
    }
  }
}

// 16 packed bf16 pairs (32 columns) of this thread's aux row, straight from global memory; returns whether the vector path applied
__device__ __forceinline__ bool ext_load_aux(const TcParams& p, const EpiCtx& e, const ExtRow& xr, bool aligned, int cb, uint32_t (&ax)[16]) {
  const bool vec = aligned && xr.row_valid && cb < e.bn && e.n0 + cb + 32 <= p.N;
  if (vec) {
    const uint4* ap = reinterpret_cast<const uint4*>(xr.aux_row + e.n0 + cb);
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) { const uint4 t = __ldg(ap + q4); ax[4 * q4] = t.x; ax[4 * q4 + 1] = t.y; ax[4 * q4 + 2] = t.z; ax[4 * q4 + 3] = t.w; }
  }
  return vec;
}

template <bool MASK, bool CSUM, int kCta>
__device__ __forceinline__ bool epi_tile_ext(const TcParams& p, const CUtensorMap* tma_d, const EpiCtx& e) {
  const int lane = e.lane;
  bool released = false;
  const bool warp_rows_valid = e.m0w < p.M;      // warp-uniform
  const int m = e.m0w + lane;
  ExtRow xr;
  xr.row_valid = m < p.M;
  xr.zero_row = e.keep == 0u || !xr.row_valid;
  const int64_t row_off = (int64_t)e.b0 * p.d_bs0 + (int64_t)e.b1 * p.d_bs1 + (int64_t)m * p.d_ld;
  xr.aux_row = p.aux != nullptr ? p.aux + row_off : nullptr;
  xr.csum = p.colsum == nullptr ? nullptr : (p.csum_smem ? e.csum_sh : p.colsum) + (int64_t)e.b0 * p.colsum_bs0;
  const bool aux_aligned = xr.aux_row != nullptr && ((reinterpret_cast<uintptr_t>(xr.aux_row) | (uintptr_t)(e.n0 * 2)) & 15) == 0;
  // (Pipelining the aux loads one half ahead of their use was tried and measured SLOWER -- FFN2 dgrad 77 -> 94 us under ncu --:
  //  the loads are issued right before the TMEM wait they overlap with.)
  uint32_t ax0[16] = {}, ax1[16] = {};
  bool v0 = false, v1 = false;
  for (int c0 = e.part * 64; c0 < e.bn; c0 += 64 * e.parts) {
    if (e.n0 + c0 >= p.N) break;
    const bool second = c0 + 32 < e.bn;
    const int c_next = c0 + 64 * e.parts;
    uint32_t r[32], pk[16];
    tc_ld32_issue(e.tbase + c0, r);
    v0 = ext_load_aux(p, e, xr, aux_aligned, c0, ax0);
    tc_wait_ld();
    // ---- columns [c0, c0 + 32)
    uint32_t r2[32];
    if (second) {
      tc_ld32_issue(e.tbase + c0 + 32, r2);       // the second half's TMEM load and aux loads are in flight under the math
      v1 = ext_load_aux(p, e, xr, aux_aligned, c0 + 32, ax1);
    }
    ext_half<MASK, CSUM>(p, e, xr, r, ax0, v0, c0, pk);
    bulk_wait_read0_el(e.el);       // previous TMA store has finished reading the staging tile
    __syncwarp();
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      st_shared_v4(e.row_smem + (((uint32_t)ch ^ e.sw) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
    if (second) tc_wait_ld();
    // every TMEM read of this chunk has landed: hand the accumulator (half) back before the remaining math / staging / store
    if (e.tempty_mid != 0u && c0 < 256 && c_next >= 256) epi_release<kCta>(e.tempty_mid, lane);
    if (c_next >= e.bn || e.n0 + c_next >= p.N) { epi_release<kCta>(e.tempty, lane); released = true; }
    // ---- columns [c0 + 32, c0 + 64)
    if (second) {
      ext_half<MASK, CSUM>(p, e, xr, r2, ax1, v1, c0 + 32, pk);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = 0u;
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      st_shared_v4(e.row_smem + (((uint32_t)(ch + 4) ^ e.sw) << 4), pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
    fence_async_smem();
    __syncwarp();
    {
      const uint32_t elv = warp_rows_valid ? e.el : 0u;
      if (p.store_mode == JMT_STORE) tma_store_4d_el(elv, tma_d, e.stage_smem, e.n0 + c0, e.m0w, e.b0, e.b1);
      else tma_reduce_add_4d_el(elv, tma_d, e.stage_smem, e.n0 + c0, e.m0w, e.b0, e.b1);
    }
  }
  return released;
}

// kX3: "bf16x3" split-operand mode (SURVEY 7 hard part 4): every fp32 operand is given as two bf16 matrices of identical
// geometry, x = hi + lo (hi = bf16(x), lo = bf16(x - hi), 16 mantissa bits together).  A stage holds the four tiles
// A_hi | A_lo and B_hi | B_lo, and every k-step issues THREE tcgen05.mma into the same fp32 TMEM accumulator:
// A_hi B_hi + A_hi B_lo + A_lo B_hi  (the dropped lo x lo term is 2^-16 relative) -- the 1e-3 parity gate on tensor cores.
// kExt: the extended backward epilogue (epi_tile_ext): 0 = none, 1 = activation-gradient fold (epi_aux), 2 = fold and / or column
// sums (d_colsum) -- own instantiations so that the common kernels keep their register allocation (122-128 registers, no spills)
template <int kCta, bool kMask, int kEpi, bool kX3, int kExt>
__global__ void __launch_bounds__(64 + 32 * kEpi, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a_hi, const __grid_constant__ CUtensorMap tma_b_hi,
               const __grid_constant__ CUtensorMap tma_d, const __grid_constant__ CUtensorMap tma_a_lo,
               const __grid_constant__ CUtensorMap tma_b_lo, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte aligned carve-up (SWIZZLE_128B atoms are 1024 B); identical offsets in both CTAs of a pair
  // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (SWIZZLE_128B atoms); fail loudly if not
  const uint32_t smem_base = smem_u32(smem_raw);
  if ((smem_base & 1023u) != 0u) __trap();
  constexpr int kOps = kX3 ? 2 : 1;                          // tiles per operand and stage (hi | lo)
  const uint32_t a_stride = kOps * kAStageBytes, b_stride = p.bres ? 0u : kOps * p.b_stage_bytes;
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + p.stages * a_stride;              // ring of B tiles, or (bres) this CTA's resident B slice: kblocks tiles
  const uint32_t sD = sB + (p.bres ? p.kblocks * p.b_stage_bytes : p.stages * b_stride);   // epilogue warps x 4 KiB staging (1024-aligned)
  const uint32_t sBias = sD + kEpi * kEpiStageBytes;         // 256 floats: this tile's bias slice
  const uint32_t bars = sBias + kBiasFlagBytes;              // 8-byte aligned (behind the bias tile and the keep-flag words)
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kMaxStages;
  const uint32_t tfull_bar = bars + 16 * kMaxStages, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  const uint32_t bres_bar = tmem_slot + 8;                   // resident B slice has landed (bres)
  uint8_t* smem_aligned = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_aligned + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (shfl: lets ptxas treat the warp index as warp-uniform)
  const int crank = kCta == 2 ? (int)cluster_ctarank() : 0;
  const int first_tile = blockIdx.x / kCta;          // tile (pair) index owned by this CTA's cluster
  const int tile_stride = gridDim.x / kCta;

  if (threadIdx.x == 0) {
    // full: one arrive(+expect_tx) per CTA of the pair (on the leader's barrier); empty / tfull: one tcgen05.commit
    // (multicast to both CTAs); tempty: every epilogue warp of every CTA of the pair (on the leader's barrier)
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar + 8 * s, kCta); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, kEpi * kCta); }
    mbar_init(bres_bar, kCta);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (kCta == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCta == 2) cluster_sync_relaxed();   // peer barriers are initialised before any remote arrive / complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // Everything above (barrier init, TMEM allocation, cluster handshake) touched no global memory, so under programmatic
  // dependent launch it overlaps the tail of the previous kernel in the stream; from here on its results are needed.
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    {
      // ================================ TMA producer (every CTA; whole warp, one elected lane issues) ================================
      const uint32_t el = elect_one();
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a_hi)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b_hi)) : "memory");
      if constexpr (kX3) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a_lo)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b_lo)) : "memory");
      }
      int stage = 0; uint32_t phase = 0;
      long long pr_wait = 0; const long long pr_t0 = p.prof ? clock64() : 0;
      const int b_chunks = p.b_chunks_cta;                    // 64-wide MN chunks staged by this CTA
      const int n_off = crank * (p.block_n / kCta);           // this CTA's half of the B tile (kCta == 2)
      const uint32_t full_leader = kCta == 2 ? mapa_rank(full_bar, 0) : full_bar;
      if constexpr (kCta == 2 && !kX3) {
        if (p.bres && first_tile < p.total_tiles) {
          // B-stationary: this pair's N slice is the same for every tile it owns (grid = a multiple of n_tiles): load all its k-blocks once
          const TileCoord c0 = decode_tile(p, first_tile, crank);
          const uint32_t fb = mapa_rank(bres_bar, 0);
          mbar_expect_tx_cluster_el(el, fb, (uint32_t)(p.kblocks * p.b_tx_bytes));
          for (int kb = 0; kb < p.kblocks; ++kb) {
            const uint32_t dst = sB + kb * p.b_stage_bytes;
            if (p.b_major == JMT_MAJOR_K) tma_load_4d_2sm_el(el, dst, &tma_b_hi, fb, kb * kBlockK, c0.n0 + n_off, 0, 0);
            else if (p.b_mn5) tma_load_5d_2sm_el(el, dst, &tma_b_hi, fb, 0, kb * kBlockK + p.b_shift0, (c0.n0 + n_off) >> 6, 0, 0);
            else for (int ch = 0; ch < b_chunks; ++ch)
              tma_load_4d_2sm_el(el, dst + ch * 8192, &tma_b_hi, fb, c0.n0 + n_off + ch * 64, kb * kBlockK + p.b_shift0, 0, 0);
          }
        }
      }
      for (int t = first_tile; t < p.total_tiles; t += tile_stride) {
        const TileCoord c = decode_tile(p, t, crank);
        // (tap, rb, kb) of the first iteration; afterwards the counters advance incrementally (no divisions in the loop)
        uint32_t q0, kb_u, tap_u, rb_u;
        p.fd_kblocks.divmod((uint32_t)c.it0, q0, kb_u);
        p.fd_rbn.divmod(q0, tap_u, rb_u);
        int kb = (int)kb_u, rb = (int)rb_u, tap = (int)tap_u;
        uint32_t b1_u, b0_u;
        p.fd_nb0.divmod((uint32_t)(p.reduce_batch ? rb : c.batch), b1_u, b0_u);
        int b0 = (int)b0_u, b1 = (int)b1_u;
        for (int it = c.it0; it < c.it1; ++it) {
          const int bb0 = p.b_has_b0 ? b0 : 0, bb1 = p.b_has_b1 ? b1 : 0;
          const int ab0 = p.a_has_b0 ? b0 : 0, ab1 = p.a_has_b1 ? b1 : 0;
          const int ash = p.a_shift0 + tap * p.a_shift_step;
          const int bsh = p.b_shift0 + tap * p.b_shift_step;
          const long long tw = p.prof ? clock64() : 0;
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          if (p.prof) pr_wait += clock64() - tw;
          const int b_tx = p.bres ? 0 : (c.tail ? p.b_tail_bytes : p.b_tx_bytes);
          if constexpr (kCta == 1) mbar_expect_tx_el(el, full_bar + 8 * stage, kOps * (kAStageBytes + b_tx));
          else mbar_expect_tx_cluster_el(el, full_leader + 8 * stage, kOps * (kAStageBytes + b_tx));
#pragma unroll
          for (int part = 0; part < kOps; ++part) {          // kX3: part 0 = hi tiles, part 1 = lo tiles (same coordinates)
          const CUtensorMap& tma_a = part == 0 ? tma_a_hi : tma_a_lo;
          const CUtensorMap& tma_b = part == 0 ? tma_b_hi : tma_b_lo;
          const uint32_t a_dst = sA + stage * a_stride + part * kAStageBytes;
          const uint32_t b_dst = sB + stage * b_stride + part * p.b_stage_bytes;
          if constexpr (kCta == 1) {
            const uint32_t fb = full_bar + 8 * stage;
            if (p.a_major == JMT_MAJOR_K) {
              tma_load_4d_el(el, a_dst, &tma_a, fb, kb * kBlockK, c.m0 + ash, ab0, ab1);
            } else if (p.a_mn5) {
              tma_load_5d_el(el, a_dst, &tma_a, fb, 0, kb * kBlockK + ash, c.m0 >> 6, ab0, ab1);
            } else {
              tma_load_4d_el(el, a_dst, &tma_a, fb, c.m0, kb * kBlockK + ash, ab0, ab1);
              tma_load_4d_el(el, a_dst + 8192, &tma_a, fb, c.m0 + 64, kb * kBlockK + ash, ab0, ab1);
            }
            if (!kX3 && c.tail) {      // tail piece: the tail map's box holds exactly this piece (tma_b_lo slot, unused without kX3)
              if (p.b_major == JMT_MAJOR_K) tma_load_4d_el(el, b_dst, &tma_b_lo, fb, tap * p.K + kb * kBlockK, c.n0, bb0, bb1);
              else tma_load_5d_el(el, b_dst, &tma_b_lo, fb, 0, kb * kBlockK + bsh, c.n0 >> 6, bb0, bb1);
            } else if (p.b_major == JMT_MAJOR_K) {
              tma_load_4d_el(el, b_dst, &tma_b, fb, tap * p.K + kb * kBlockK, c.n0, bb0, bb1);
            } else if (p.b_mn5) {
              tma_load_5d_el(el, b_dst, &tma_b, fb, 0, kb * kBlockK + bsh, c.n0 >> 6, bb0, bb1);
            } else {
              for (int ch = 0; ch < b_chunks; ++ch)
                tma_load_4d_el(el, b_dst + ch * 8192, &tma_b, fb, c.n0 + ch * 64, kb * kBlockK + bsh, bb0, bb1);
            }
          } else {
            const uint32_t fb = full_leader + 8 * stage;
            if (p.bres && p.l2_ahead > 0 && t + p.l2_ahead * tile_stride < p.total_tiles) {
              // B-stationary: only A is streamed and its ring is short (B takes 128 KB): pull the same k-block of the tile this pair
              // works on `l2_ahead` tiles from now into L2, so the later load is an L2 hit (~1/3 of the DRAM round trip)
              const int m_next = c.m0 + p.l2_ahead * p.l2_mstep;
              if (p.a_major == JMT_MAJOR_K) tma_prefetch_4d_el(el, &tma_a, kb * kBlockK, m_next, 0, 0);
              else if (p.a_mn5) tma_prefetch_5d_el(el, &tma_a, 0, kb * kBlockK, m_next >> 6, 0, 0);
            }
            if (p.a_major == JMT_MAJOR_K) {
              tma_load_4d_2sm_el(el, a_dst, &tma_a, fb, kb * kBlockK, c.m0 + ash, ab0, ab1);
            } else if (p.a_mn5) {
              tma_load_5d_2sm_el(el, a_dst, &tma_a, fb, 0, kb * kBlockK + ash, c.m0 >> 6, ab0, ab1);
            } else {
              tma_load_4d_2sm_el(el, a_dst, &tma_a, fb, c.m0, kb * kBlockK + ash, ab0, ab1);
              tma_load_4d_2sm_el(el, a_dst + 8192, &tma_a, fb, c.m0 + 64, kb * kBlockK + ash, ab0, ab1);
            }
            if (p.bres) {               // (B is resident)
            } else if (!kX3 && c.tail) {      // tail piece: this CTA's half of it through the tail map (tma_b_lo slot, unused without kX3)
              const int n_t = c.n0 + crank * (p.tail_bn / kCta);
              if (p.b_major == JMT_MAJOR_K) tma_load_4d_2sm_el(el, b_dst, &tma_b_lo, fb, tap * p.K + kb * kBlockK, n_t, bb0, bb1);
              else tma_load_5d_2sm_el(el, b_dst, &tma_b_lo, fb, 0, kb * kBlockK + bsh, n_t >> 6, bb0, bb1);
            } else if (p.wide) {          // two 128-column pieces: this CTA's share of the B operand of each of the two MMAs
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int n_h = c.n0 + h * 256 + crank * 128;
                if (p.b_major == JMT_MAJOR_K) tma_load_4d_2sm_el(el, b_dst + h * 16384, &tma_b, fb, tap * p.K + kb * kBlockK, n_h, bb0, bb1);
                else tma_load_5d_2sm_el(el, b_dst + h * 16384, &tma_b, fb, 0, kb * kBlockK + bsh, n_h >> 6, bb0, bb1);
              }
            } else if (p.b_major == JMT_MAJOR_K) {
              tma_load_4d_2sm_el(el, b_dst, &tma_b, fb, tap * p.K + kb * kBlockK, c.n0 + n_off, bb0, bb1);
            } else if (p.b_mn5) {
              tma_load_5d_2sm_el(el, b_dst, &tma_b, fb, 0, kb * kBlockK + bsh, (c.n0 + n_off) >> 6, bb0, bb1);
            } else {
              for (int ch = 0; ch < b_chunks; ++ch)
                tma_load_4d_2sm_el(el, b_dst + ch * 8192, &tma_b, fb, c.n0 + n_off + ch * 64, kb * kBlockK + bsh, bb0, bb1);
            }
          }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          if (++kb == p.kblocks) {
            kb = 0;
            if (++rb == p.rb_n) { rb = 0; ++tap; }
            if (p.reduce_batch) { if (++b0 == p.nb0) { b0 = 0; ++b1; } if (rb == 0) { b0 = 0; b1 = 0; } }
          }
        }
      }
      if (p.prof && lane == 0) { p.prof[blockIdx.x * 16 + 3] = pr_wait; p.prof[blockIdx.x * 16 + 4] = clock64() - pr_t0; }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ================================ MMA issuer (leader CTA of a pair; whole warp, one elected lane issues) ================================
      const uint32_t el = elect_one();
      int stage = 0; uint32_t phase = 0;
      int tile_iter = 0;
      long long mw_full = 0, mw_tempty = 0; const long long mw_t0 = p.prof ? clock64() : 0;
      const uint32_t a_lbo = p.a_major == JMT_MAJOR_K ? 16u : 8192u;
      const uint32_t b_lbo = p.b_major == JMT_MAJOR_K ? 16u : 8192u;
      const uint32_t a_kstep = p.a_major == JMT_MAJOR_K ? (kUmmaK * 2) >> 4 : (kUmmaK * 128) >> 4;   // desc.lo units (16 B)
      const uint32_t b_kstep = p.b_major == JMT_MAJOR_K ? (kUmmaK * 2) >> 4 : (kUmmaK * 128) >> 4;
      if (p.bres && first_tile < p.total_tiles) { mbar_wait(bres_bar, 0); tc_fence_after(); }     // the resident B slice has landed
      for (int t = first_tile; t < p.total_tiles; t += tile_stride, ++tile_iter) {
        const TileCoord c = decode_tile(p, t, crank);
        const int acc = p.wide ? 0 : (tile_iter & 1);
        const uint32_t acc_phase = (p.wide ? tile_iter : (tile_iter >> 1)) & 1;
        const long long tw0 = p.prof ? clock64() : 0;
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        if (p.two_phase && c.tail) mbar_wait(tempty_bar + 8, acc_phase ^ 1);     // a tail piece consumes a phase of both halves
        if (p.prof) mw_tempty += clock64() - tw0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccStride;
        if (!kX3 && p.two_phase && !c.tail) {
          // Two-phase wide tile: half 0 (columns 0..255) is free, half 1 (256..511) may still be drained by the epilogue of
          // the previous tile.  Issue the half-0 MMAs of up to `stages` k-blocks ahead; their half-1 MMAs (and the commits
          // that free the shared-memory slots) follow as soon as tempty[1] flips.
          bool h1_free = false;
          int pend = 0, pend_stage = stage, pend_it = c.it0;
          for (int it = c.it0; it < c.it1; ++it) {
            const long long tw1 = p.prof ? clock64() : 0;
            mbar_wait(full_bar + 8 * stage, phase);
            if (p.prof) mw_full += clock64() - tw1;
            tc_fence_after();
            {
              const uint64_t a_desc = make_smem_desc(sA + stage * a_stride, a_lbo, 1024);
              const uint64_t b_desc = make_smem_desc(sB + stage * b_stride, b_lbo, 1024);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                tc_mma_elect<kCta>(el, d_tmem, a_desc + (uint64_t)(k * a_kstep), b_desc + (uint64_t)(k * b_kstep), p.idesc, (it > c.it0 || k > 0) ? 1u : 0u);
            }
            ++pend;
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            if (!h1_free) {
              if (pend == p.stages || it == c.it1 - 1) {
                const long long tw2 = p.prof ? clock64() : 0;
                mbar_wait(tempty_bar + 8, acc_phase ^ 1);
                if (p.prof) mw_tempty += clock64() - tw2;
                h1_free = true;
              } else {
                h1_free = __shfl_sync(0xffffffffu, mbar_test(tempty_bar + 8, acc_phase ^ 1) ? 1 : 0, 0) != 0;
              }
              if (h1_free) tc_fence_after();
            }
            if (h1_free) {
              for (; pend > 0; --pend, ++pend_it) {
                const uint64_t a_desc = make_smem_desc(sA + pend_stage * a_stride, a_lbo, 1024);
                const uint64_t b_desc = make_smem_desc(sB + pend_stage * b_stride, b_lbo, 1024);
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k)
                  tc_mma_elect<kCta>(el, d_tmem + 256, a_desc + (uint64_t)(k * a_kstep), b_desc + (uint64_t)(1024 + k * b_kstep), p.idesc,
                               (pend_it > c.it0 || k > 0) ? 1u : 0u);
                tc_commit_elect<kCta>(el, empty_bar + 8 * pend_stage);
                if (++pend_stage == p.stages) pend_stage = 0;
              }
            }
          }
          tc_commit_elect<kCta>(el, tfull_bar);
          continue;
        }
        for (int it = c.it0; it < c.it1; ++it) {
          const long long tw1 = p.prof ? clock64() : 0;
          mbar_wait(full_bar + 8 * stage, phase);
          if (p.prof) mw_full += clock64() - tw1;
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc(sA + stage * a_stride, a_lbo, 1024);
          const uint64_t b_desc = make_smem_desc(p.bres ? sB + it * p.b_stage_bytes : sB + stage * b_stride, b_lbo, 1024);   // (bres: it = k-block)
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            tc_mma_elect<kCta>(el, d_tmem, a_desc + (uint64_t)(k * a_kstep), b_desc + (uint64_t)(k * b_kstep), c.tail ? p.idesc_tail : p.idesc,
                         (it > c.it0 || k > 0) ? 1u : 0u);
            if constexpr (kX3) {   // + A_hi B_lo + A_lo B_hi (the lo tiles sit one tile behind the hi tiles; descriptor units of 16 B)
              tc_mma_elect<kCta>(el, d_tmem, a_desc + (uint64_t)(k * a_kstep), b_desc + (uint64_t)((p.b_stage_bytes >> 4) + k * b_kstep), p.idesc, 1u);
              tc_mma_elect<kCta>(el, d_tmem, a_desc + (uint64_t)((kAStageBytes >> 4) + k * a_kstep), b_desc + (uint64_t)(k * b_kstep), p.idesc, 1u);
            }
            if (p.wide && !c.tail)      // columns 256..511 of the tile: same A, second B piece (16 KB further), TMEM columns 256..511
              tc_mma_elect<kCta>(el, d_tmem + 256, a_desc + (uint64_t)(k * a_kstep), b_desc + (uint64_t)(1024 + k * b_kstep), p.idesc,
                           (it > c.it0 || k > 0) ? 1u : 0u);
          }
          tc_commit_elect<kCta>(el, empty_bar + 8 * stage);  // frees the smem slot (in both CTAs) once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit_elect<kCta>(el, tfull_bar + 8 * acc);      // accumulator ready for the epilogue warps (of both CTAs)
      }
      if (p.prof && lane == 0) { p.prof[blockIdx.x * 16 + 0] = mw_full; p.prof[blockIdx.x * 16 + 1] = mw_tempty; p.prof[blockIdx.x * 16 + 2] = clock64() - mw_t0; }
    }
  } else {
    // ================================ epilogue (warps 2 .. 2 + kEpi - 1) ================================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int ew = warp - 2;                      // epilogue warp index 0 .. kEpi - 1
    const int part = ew >> 2;                     // which interleaved set of column chunks this warp drains (of kEpi / 4)
    const uint32_t stage_smem = sD + ew * kEpiStageBytes;
    float* bias_ptr = reinterpret_cast<float*>(smem_aligned + (sBias - smem_base));
    const int et = threadIdx.x - 64;              // index among the epilogue threads
    const uint32_t epi_el = elect_one();          // the lane that issues, commits and waits for this warp's TMA stores
    const uint32_t row_smem = stage_smem + lane * 128;
    const uint32_t sw = lane & 7;                 // 128B-swizzle phase of this thread's staging row
    const uint32_t tempty_leader = kCta == 2 ? mapa_rank(tempty_bar, 0) : tempty_bar;
    long long ep_tfull = 0, ep_bar = 0, ep_rd = 0; const long long ep_t0 = p.prof ? clock64() : 0;
    if ((p.bias == nullptr && !kMask) || p.csum_smem) {   // no bias (or column-sum accumulators): one zero fill for the whole kernel
      for (int e2 = et; e2 < 512; e2 += 32 * kEpi) bias_ptr[e2] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");
    }
#ifdef JMT_EPI_PROF
    long long eprof[6] = {0, 0, 0, 0, 0, 0};
#endif
    int tile_iter = 0;
    for (int t = first_tile; t < p.total_tiles; t += tile_stride, ++tile_iter) {
      const TileCoord c = decode_tile(p, t, crank);
      const int split = c.split;
      const int acc = p.wide ? 0 : (tile_iter & 1);
      const uint32_t acc_phase = (p.wide ? tile_iter : (tile_iter >> 1)) & 1;
      // stage this tile's bias slice in shared memory (overlaps the mainloop); named barrier 1 = epilogue warps
      const long long tb0 = p.prof ? clock64() : 0;
      // kMask == false: the bias tile is double-buffered by tile parity (the keep-flag area is free), so ONE barrier per tile
      // suffices -- buffer (i & 1) was last read for tile i - 2, and every warp finished those reads before it arrived at the
      // barrier of tile i - 1, which this writer has passed
      // (a wide tile's 512 bias values fill both buffers: single buffer, two barriers, like kMask)
      const bool bias_dbuf = !kMask && !p.wide;
      float* bias_tile = bias_ptr + ((bias_dbuf && p.bias != nullptr) ? (tile_iter & 1) * 256 : 0);
      if (p.bias != nullptr || kMask) {
        const bool add_bias = p.bias != nullptr && split == 0;
        if (!bias_dbuf) asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");     // previous tile's readers are done
        // every epilogue thread takes columns et, et + 32 * kEpi, ... of the tile (warp-uniform trip count: block_n % 32 == 0)
        for (int e2 = et; e2 < c.bn; e2 += 32 * kEpi) {
          const bool in_n = c.n0 + e2 < p.N;
          if (!p.csum_smem) bias_tile[e2] = (add_bias && in_n) ? __ldg(p.bias + c.n0 + e2) : 0.f;
          if constexpr (kMask) {
            // keep-flags of the (up to two) samples this tile's rows belong to: sample = batch, or row / period (flat layout);
            // one ballot per 32 columns and sample
            uint32_t* fw = reinterpret_cast<uint32_t*>(bias_ptr + 512);
            const int s0 = p.colmask_period ? (int)p.fd_cmask.div((uint32_t)c.m0) : c.batch;
            const uint32_t w0 = __ballot_sync(0xffffffffu, in_n && p.colmask[(int64_t)s0 * p.N + c.n0 + e2]);
            const uint32_t w1 = __ballot_sync(0xffffffffu, in_n && p.colmask_period && s0 + 1 < p.colmask_samples &&
                                                            p.colmask[(int64_t)(s0 + 1) * p.N + c.n0 + e2]);
            if (lane == 0) { fw[e2 >> 5] = w0; fw[16 + (e2 >> 5)] = w1; }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");
      }
      const long long tb1 = p.prof ? clock64() : 0;
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      if (p.prof) { ep_bar += tb1 - tb0; ep_tfull += clock64() - tb1; }
      tc_fence_after();
      uint32_t b1u, b0u;
      p.fd_nb0.divmod((uint32_t)c.batch, b1u, b0u);
      EpiCtx ec;
      ec.tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kAccStride);
      ec.stage_smem = stage_smem; ec.row_smem = row_smem; ec.sw = sw; ec.bias = bias_tile; ec.el = epi_el;
      {
        const uint32_t m = (uint32_t)(c.m0 + q * 32 + lane);
        ec.keep = ~0u;
        if (p.zrow_period) { uint32_t qq, rr; p.fd_zrow.divmod(m, qq, rr); ec.keep = rr < (uint32_t)p.zrow_count ? 0u : ~0u; }
        const uint32_t* fw = reinterpret_cast<const uint32_t*>(bias_ptr + 512);
        if (kMask && p.colmask_period) fw += (p.fd_cmask.div(m) != p.fd_cmask.div((uint32_t)c.m0)) ? 16 : 0;
        ec.fwords = fw;
      }
      ec.m0w = c.m0 + q * 32; ec.n0 = c.n0; ec.b0 = (int)b0u; ec.b1 = (int)b1u; ec.part = part; ec.parts = kEpi / 4; ec.lane = lane;
      ec.bn = c.bn;
      ec.csum_sh = bias_ptr;
      // (kMask: channel dropout fused after the activation, TCN -- a separate kernel instantiation so that its extra
      //  register pressure never touches the common kernels)
      {
        const uint32_t te = kCta == 1 ? tempty_bar : tempty_leader;
        ec.tempty = te + 8 * acc; ec.tempty_mid = 0u; ec.tempty_end2 = 0u;
        if (p.two_phase) {
          if (c.tail) ec.tempty_end2 = te + 8;                   // a tail piece uses (and hands back) both halves at once
          else { ec.tempty_mid = te; ec.tempty = te + 8; }       // half 0 after columns [0, 256), half 1 at the end
        }
      }
#ifdef JMT_EPI_PROF
      ec.prof = eprof;
#endif
      bool released;
      if constexpr (kExt != 0) {
        released = epi_tile_ext<kMask, kExt == 2, kCta>(p, &tma_d, ec);
      } else {
        if (p.act == JMT_ACT_NONE) released = epi_tile<JMT_ACT_NONE, kMask, kCta>(p, &tma_d, ec);
        else if (p.act == JMT_ACT_RELU) released = epi_tile<JMT_ACT_RELU, kMask, kCta>(p, &tma_d, ec);
        else released = epi_tile<JMT_ACT_LEAKY_RELU, kMask, kCta>(p, &tma_d, ec);
      }
      if (!released) {
        // (a two-phase tile that ended before its half boundary -- columns beyond N -- still owes the half-0 arrive)
        if (ec.tempty_mid != 0u) epi_release<kCta>(ec.tempty_mid, lane);
        epi_release<kCta>(ec.tempty, lane);
      }
      if (ec.tempty_end2 != 0u) epi_release<kCta>(ec.tempty_end2, lane);
    }
    if (kExt == 2 && p.csum_smem) {                     // flush the shared-memory column sums: one global atomic per entry and CTA
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpi) : "memory");
      for (int i = et; i < p.csum_len; i += 32 * kEpi) {
        const float v = bias_ptr[i];
        if (v != 0.f) atomicAdd(p.colsum + i, v);
      }
    }
    bulk_wait0_el(epi_el);                         // all TMA stores of this warp have completed
#ifdef JMT_EPI_PROF
    if (blockIdx.x == 0 && threadIdx.x == 64) for (int i = 0; i < 6; ++i) g_epi_prof[i] = (unsigned long long)eprof[i];
#endif
    if (p.prof && ew == 0 && lane == 0) {
      unsigned long long* o = p.prof + blockIdx.x * 16;
      o[5] = ep_tfull; o[6] = ep_bar; o[7] = ep_rd; o[8] = clock64() - ep_t0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kCta == 2) cluster_sync_relaxed();    // no CTA exits (or frees TMEM) while its peer may still read / signal it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kCta == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ----------------------------------------------------------------------------- host side
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

static std::atomic<unsigned long long*> g_prof_buf{nullptr};
// B-stationary mode of the short-K linears: -1 = environment (JMT_GEMM_BRES, default 1 = automatic), 0 = off, 1 = automatic, 2 = forced
// whenever the geometry allows it (tests: small M)
static std::atomic<int> g_bres_mode{-1};
extern "C" int jmt_gemm_set_bres_mode(int mode) { return g_bres_mode.exchange(mode); }
// Debug aid: when set (device buffer of 16 x 148 uint64), every jmt_gemm_bf16 launch overwrites per-CTA cycle
// counters: [0] MMA wait-full [1] MMA wait-tmem-empty [2] MMA total [3] TMA wait-empty [4] TMA total
// [5] epilogue wait-tmem-full [6] epilogue bias barriers [7] epilogue wait-store-read [8] epilogue total
extern "C" int jmt_gemm_set_profile_buffer(void* dev_buf) {
  g_prof_buf.store((unsigned long long*)dev_buf);
  return JMT_OK;
}

#ifdef JMT_EPI_PROF
extern "C" int jmt_gemm_epi_prof_read(unsigned long long* host8, int reset) {
  cudaMemcpyFromSymbol(host8, jmt::g_epi_prof, 8 * sizeof(unsigned long long));
  if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(jmt::g_epi_prof, z, sizeof(z)); }
  return 0;
}
#endif

// a_lo / b_lo != NULL: the bf16x3 split-operand mode (g->a / g->b are the hi parts; lo parts have identical geometry)
static int gemm_tc_launch(const jmt_gemm_desc* g, const void* a_lo, const void* b_lo, void* stream) {
  int rc = jmt_validate_gemm_desc(g, "jmt_gemm_bf16");
  if (rc != JMT_OK) return rc;
  const bool x3 = a_lo != nullptr;
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
  p.b_has_b0 = (g->nb0 > 1 && g->b_bs0 != 0) ? 1 : 0;
  p.b_has_b1 = (g->nb1 > 1 && g->b_bs1 != 0) ? 1 : 0;
  p.a_has_b0 = (g->nb0 > 1 && g->a_bs0 != 0) ? 1 : 0;
  p.a_has_b1 = (g->nb1 > 1 && g->a_bs1 != 0) ? 1 : 0;
  // CTA pairs (cta_group::2, M = 256): along M when the ghost tile of an odd tail wastes < ~10 %, else across two
  // consecutive batch entries when they share B (implicit-GEMM conv: weights are not batched)
  const char* env_cl = getenv("JMT_GEMM_CLUSTER");
  const bool cl_enabled = env_cl ? atoi(env_cl) != 0 : true;
  const bool pair_m_ok = p.m_tiles >= 2 && (p.m_tiles % 2 == 0 || p.m_tiles >= 9);
  const bool pair_b_ok = !p.reduce_batch && !p.b_has_b0 && !p.b_has_b1 && p.batch_tiles >= 2 && p.batch_tiles % 2 == 0;
  // (tiles with fewer than 4 k-iterations are epilogue-bound: the pair's shared tempty / full handshakes only cost there)
  const bool long_enough = p.iters_total / p.split_k >= 4;
  p.cluster = (cl_enabled && long_enough && (pair_m_ok || pair_b_ok)) ? 2 : 1;
  p.pair_batch = (p.cluster == 2 && !pair_m_ok) ? 1 : 0;
  // wide tiles (256 x 512 per pair) for long-K GEMMs whose N is a multiple of 512; JMT_GEMM_WIDE=0 disables
  {
    static const int wide_env = []() { const char* e = getenv("JMT_GEMM_WIDE"); return e ? atoi(e) : 1; }();
    const int min_iters = wide_env > 1 ? wide_env : kWideMinIters;
    // B-stationary pairs for short-K linears (see TcParams::bres); JMT_GEMM_BRES=0 disables
    static const int bres_env0 = []() { const char* e = getenv("JMT_GEMM_BRES"); return e ? atoi(e) : 1; }();
    const int bres_set = g_bres_mode.load();
    const int bres_env = bres_set >= 0 ? bres_set : bres_env0;
    p.bres = (bres_env != 0 && !x3 && p.cluster == 2 && !p.pair_batch && g->ntaps == 1 && nb == 1 && p.split_k == 1 && g->N % 256 == 0 &&
              p.kblocks <= 8 && g->N / 256 <= kNumSMs / 2 &&
              (bres_env == 2 /* forced: tests */ || (p.kblocks >= 4 && p.m_tiles >= 8 * (kNumSMs / 2) / (g->N / 256)))) ? 1 : 0;
    if (p.bres) { p.block_n = 256; p.n_tiles = g->N / 256; }
    p.wide = (wide_env != 0 && !x3 && !p.bres && p.cluster == 2 && !p.pair_batch && g->N % 512 == 0 &&
              p.iters_total / p.split_k >= min_iters) ? 1 : 0;
    if (p.wide) { p.block_n = 512; p.n_tiles = g->N / 512; }
    // measured (round 2, profiles/gemm_roles_r2h_tp*.txt): the MMA issuer's wait for the accumulator halves (20.9 k -> 11.7 k of 62 k
    // cycles on 76800x512x512) but the launch is not faster (44.7 -> 45.3 us; K = 3072: 336 -> 352 us) -- off unless JMT_GEMM_TWO_PHASE=1
    static const int two_env = []() { const char* e = getenv("JMT_GEMM_TWO_PHASE"); return e ? atoi(e) : 0; }();
    p.two_phase = (p.wide && two_env != 0) ? 1 : 0;
  }
  p.m_pairs = (p.cluster == 2 && !p.pair_batch) ? (p.m_tiles + 1) / 2 : p.m_tiles;
  p.batch_slots = p.pair_batch ? p.batch_tiles / 2 : p.batch_tiles;
  const int64_t total = (int64_t)p.m_pairs * p.n_tiles * p.batch_slots * p.split_k;
  JMT_REQUIRE(total < (1ll << 31), "jmt_gemm_bf16: too many tiles");
  p.total_tiles = (int)total;
  p.nb0 = g->nb0;
  p.fd_ntiles.init(p.n_tiles); p.fd_mpairs.init(p.m_pairs); p.fd_bslots.init(p.batch_slots); p.fd_nb0.init(g->nb0);
  p.fd_kblocks.init(p.kblocks); p.fd_rbn.init(p.rb_n);
  p.a_major = g->a_major; p.b_major = g->b_major;
  p.a_shift0 = g->a_shift0; p.a_shift_step = g->a_shift_step;
  p.b_shift0 = g->b_shift0; p.b_shift_step = g->b_shift_step;
  JMT_REQUIRE(!(g->ntaps > 1 && g->b_major == JMT_MAJOR_K && g->K % 8 != 0), "jmt_gemm_bf16: taps need K %% 8 == 0");
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)g->a_major << 15) | ((uint32_t)g->b_major << 16) |
            ((uint32_t)((p.wide ? 256 : p.block_n) >> 3) << 17) | ((uint32_t)((kBlockM * p.cluster) >> 4) << 24);
  // per CTA: the whole B tile, or its half of it under cta_group::2
  // (wide: two 128-column pieces per CTA, one per MMA, 16 KB each)
  const int b_cols_cta = p.wide ? 128 : p.block_n / p.cluster;
  p.b_chunks_cta = (b_cols_cta + 63) / 64;
  p.b_stage_bytes = (g->b_major == JMT_MAJOR_K ? b_cols_cta * 128 : p.b_chunks_cta * 8192) * (p.wide ? 2 : 1);
  p.b_tx_bytes = p.b_stage_bytes;
  const int stage_bytes = p.bres ? kAStageBytes : (kAStageBytes + p.b_stage_bytes) * (x3 ? 2 : 1);     // x3: hi and lo tile of each operand
  const int bres_bytes = p.bres ? p.kblocks * p.b_stage_bytes : 0;
  // 8 epilogue warps, two per TMEM lane quarter (a 16-warp variant measured 3-10 % slower on every shape in round 1: the
  // epilogue waits on the stage hand-over, it is not short of warps; the instantiation was dropped)
  p.epi_warps = 8;
  const int budget = 227 * 1024 - 256 /*barriers*/ - epi_smem_bytes(p.epi_warps);      // the dynamic window is 1024-aligned (checked in the kernel)
  p.stages = (budget - bres_bytes) / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  JMT_REQUIRE(p.stages >= 2, "jmt_gemm_bf16: shared memory budget");
  p.prof = g_prof_buf.load();
  p.colmask = g->colmask; p.colmask_scale = g->colmask_scale;
  p.colmask_period = g->colmask_row_period; p.colmask_samples = g->colmask_row_period > 0 ? (g->M + g->colmask_row_period - 1) / g->colmask_row_period : 0;
  p.zrow_period = g->zero_row_period; p.zrow_count = g->zero_row_count;
  p.fd_zrow.init(g->zero_row_period > 0 ? g->zero_row_period : 1); p.fd_cmask.init(g->colmask_row_period > 0 ? g->colmask_row_period : 1);
  p.d = g->d; p.bias = g->bias; p.d_ld = g->d_ld; p.d_bs0 = g->d_bs0; p.d_bs1 = g->d_bs1;
  p.alpha = g->alpha; p.slope = g->slope; p.d_dtype = g->d_dtype; p.act = g->act; p.store_mode = g->store_mode;
  p.aux = (const __nv_bfloat16*)g->epi_aux; p.aux_slope = g->aux_slope; p.colsum = g->d_colsum; p.colsum_bs0 = g->colsum_bs0;
  JMT_REQUIRE(!(x3 && (g->epi_aux || g->d_colsum)), "jmt_gemm_bf16x3: epi_aux / d_colsum are not available in the split-operand mode");
  {
    const int64_t span = (int64_t)(g->nb0 - 1) * g->colsum_bs0 + g->N;          // entries of d_colsum this launch can touch
    p.csum_smem = (g->d_colsum != nullptr && g->bias == nullptr && g->colsum_bs0 >= 0 && span <= 512) ? 1 : 0;
    p.csum_len = p.csum_smem ? (int)span : 0;
  }
  const int64_t es = g->d_dtype == JMT_F32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(g->d) & 15) == 0 && (g->d_ld * es) % 16 == 0 &&
              (g->d_bs0 * es) % 16 == 0 && (g->d_bs1 * es) % 16 == 0) ? 1 : 0;

  CUtensorMap map_a, map_b;
  // A: K-major -> (K, a_rows) box {64,128}; MN-major -> (M, a_rows = k extent): both 64-wide chunks of the 128-row tile in one
  // 5-D box when M % 64 == 0, else two 4-D {64,64} boxes
  p.a_mn5 = (g->a_major == JMT_MAJOR_MN && g->M % 64 == 0) ? 1 : 0;
  const int anb0 = p.a_has_b0 ? g->nb0 : 1, anb1 = p.a_has_b1 ? g->nb1 : 1;
  if (g->a_major == JMT_MAJOR_K)
    rc = make_map(&map_a, g->a, g->K, g->a_rows, g->a_ld, anb0, g->a_bs0, anb1, g->a_bs1, kBlockM, "jmt_gemm_bf16(A)");
  else if (p.a_mn5)
    rc = make_map_mn5(&map_a, g->a, g->M, g->a_rows, g->a_ld, anb0, g->a_bs0, anb1, g->a_bs1, kBlockK, 2, "jmt_gemm_bf16(A)");
  else
    rc = make_map(&map_a, g->a, g->M, g->a_rows, g->a_ld, anb0, g->a_bs0, anb1, g->a_bs1, kBlockK, "jmt_gemm_bf16(A)");
  if (rc != JMT_OK) return rc;
  const int bnb0 = p.b_has_b0 ? g->nb0 : 1, bnb1 = p.b_has_b1 ? g->nb1 : 1;
  p.b_mn5 = (g->b_major == JMT_MAJOR_MN && g->N % 64 == 0 && p.block_n % 64 == 0 && b_cols_cta % 64 == 0) ? 1 : 0;
  if (g->b_major == JMT_MAJOR_K)
    rc = make_map(&map_b, g->b, (int64_t)g->ntaps * g->K, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1,
                  b_cols_cta /* each CTA of a pair stages half of the tile's rows */, "jmt_gemm_bf16(B)");
  else if (p.b_mn5)
    rc = make_map_mn5(&map_b, g->b, g->N, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1, kBlockK, p.b_chunks_cta, "jmt_gemm_bf16(B)");
  else
    rc = make_map(&map_b, g->b, g->N, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1, kBlockK, "jmt_gemm_bf16(B)");
  if (rc != JMT_OK) return rc;

  CUtensorMap map_a_lo = map_a, map_b_lo = map_b;
  if (x3) {
    if (g->a_major == JMT_MAJOR_K)
      rc = make_map(&map_a_lo, a_lo, g->K, g->a_rows, g->a_ld, anb0, g->a_bs0, anb1, g->a_bs1, kBlockM, "jmt_gemm_bf16x3(A lo)");
    else if (p.a_mn5)
      rc = make_map_mn5(&map_a_lo, a_lo, g->M, g->a_rows, g->a_ld, anb0, g->a_bs0, anb1, g->a_bs1, kBlockK, 2, "jmt_gemm_bf16x3(A lo)");
    else
      rc = make_map(&map_a_lo, a_lo, g->M, g->a_rows, g->a_ld, anb0, g->a_bs0, anb1, g->a_bs1, kBlockK, "jmt_gemm_bf16x3(A lo)");
    if (rc != JMT_OK) return rc;
    if (g->b_major == JMT_MAJOR_K)
      rc = make_map(&map_b_lo, b_lo, (int64_t)g->ntaps * g->K, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1, b_cols_cta, "jmt_gemm_bf16x3(B lo)");
    else if (p.b_mn5)
      rc = make_map_mn5(&map_b_lo, b_lo, g->N, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1, kBlockK, p.b_chunks_cta, "jmt_gemm_bf16x3(B lo)");
    else
      rc = make_map(&map_b_lo, b_lo, g->N, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1, kBlockK, "jmt_gemm_bf16x3(B lo)");
    if (rc != JMT_OK) return rc;
  }

  // Tail split: cut the tiles of a nearly empty last round along N (see TcParams).  Needs whole tiles per CTA in one pass (no
  // split-K / batch reduction) and, for MN-major B, whole 64-column chunks per CTA.  JMT_GEMM_TAIL=0 disables.
  p.tail_start = p.total_tiles; p.tail_split = 1; p.tail_bn = p.block_n; p.idesc_tail = p.idesc; p.b_tail_bytes = p.b_tx_bytes;
  {
    static const int tail_env = []() { const char* e = getenv("JMT_GEMM_TAIL"); return e ? atoi(e) : 1; }();
    const int units = kNumSMs / p.cluster;
    const int rem = p.total_tiles % units;
    const int gran = g->b_major == JMT_MAJOR_K ? 64 : 64 * p.cluster;
    const bool b_ok = g->b_major == JMT_MAJOR_K || p.b_mn5;
    if (tail_env && !x3 && !p.bres && p.split_k == 1 && !p.reduce_batch && b_ok && p.total_tiles > units && rem > 0 && 2 * rem <= units) {
      int sp = 1;
      while (sp < 8 && rem * sp * 2 <= units && p.block_n % (sp * 2) == 0 && (p.block_n / (sp * 2)) % gran == 0) sp *= 2;
      if (sp > 1 && p.block_n / sp <= 256) {
        p.tail_split = sp;
        p.tail_bn = p.block_n / sp;
        p.tail_start = p.total_tiles - rem;
        p.total_tiles = p.tail_start + rem * sp;
        p.idesc_tail = (p.idesc & ~(0x3Fu << 17)) | ((uint32_t)(p.tail_bn >> 3) << 17);
        const int rows_cta = p.tail_bn / p.cluster;
        p.b_tail_bytes = g->b_major == JMT_MAJOR_K ? rows_cta * 128 : (rows_cta / 64) * 8192;
        if (g->b_major == JMT_MAJOR_K)
          rc = make_map(&map_b_lo, g->b, (int64_t)g->ntaps * g->K, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1, rows_cta, "jmt_gemm_bf16(B tail)");
        else
          rc = make_map_mn5(&map_b_lo, g->b, g->N, g->b_rows, g->b_ld, bnb0, g->b_bs0, bnb1, g->b_bs1, kBlockK, rows_cta / 64, "jmt_gemm_bf16(B tail)");
        if (rc != JMT_OK) return rc;
      }
    }
  }
  p.fd_tsplit.init(p.tail_split);

  // D through TMA (store / reduce-add) when its geometry is 16-byte aligned; bf16 read-modify-write
  // accumulation and fp32 atomics both become cp.reduce.async.bulk.tensor .add
  CUtensorMap map_d = map_a;
  p.tma_store = p.vec_ok;
  JMT_REQUIRE(!((g->epi_aux || g->d_colsum) && !p.tma_store), "jmt_gemm_bf16: epi_aux / d_colsum need a 16-byte aligned D geometry");
  if (p.tma_store) {
    const int dnb0 = p.reduce_batch ? 1 : g->nb0, dnb1 = p.reduce_batch ? 1 : g->nb1;
    rc = make_map_d(&map_d, g->d, g->d_dtype, g->N, g->M, g->d_ld, dnb0, g->d_bs0, dnb1, g->d_bs1, "jmt_gemm_bf16(D)");
    if (rc != JMT_OK) return rc;
  }
  const int smem = p.stages * stage_bytes + bres_bytes + epi_smem_bytes(p.epi_warps) + 256;
  static std::atomic<int> attr_set[64];     // per device (immutable once set)
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); set_error("jmt_gemm_bf16: no CUDA device"); return JMT_ERR_CUDA; }
  if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t e = cudaSuccess;
    auto set_smem = [&e](const void* fn) { if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); };
    set_smem((const void*)gemm_tc_kernel<1, false, 8, false, 0>); set_smem((const void*)gemm_tc_kernel<2, false, 8, false, 0>);
    set_smem((const void*)gemm_tc_kernel<1, true, 8, false, 0>); set_smem((const void*)gemm_tc_kernel<2, true, 8, false, 0>);
    set_smem((const void*)gemm_tc_kernel<1, false, 8, true, 0>); set_smem((const void*)gemm_tc_kernel<2, false, 8, true, 0>);
    set_smem((const void*)gemm_tc_kernel<1, true, 8, true, 0>); set_smem((const void*)gemm_tc_kernel<2, true, 8, true, 0>);
    set_smem((const void*)gemm_tc_kernel<1, false, 8, false, 1>); set_smem((const void*)gemm_tc_kernel<2, false, 8, false, 1>);
    set_smem((const void*)gemm_tc_kernel<1, true, 8, false, 1>); set_smem((const void*)gemm_tc_kernel<2, true, 8, false, 1>);
    set_smem((const void*)gemm_tc_kernel<1, false, 8, false, 2>); set_smem((const void*)gemm_tc_kernel<2, false, 8, false, 2>);
    set_smem((const void*)gemm_tc_kernel<1, true, 8, false, 2>); set_smem((const void*)gemm_tc_kernel<2, true, 8, false, 2>);
    if (e != cudaSuccess) { set_error("jmt_gemm_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return JMT_ERR_CUDA; }
    attr_set[dev & 63].store(1, std::memory_order_release);
  }
  int max_groups = kNumSMs / p.cluster;
  if (p.bres) max_groups -= max_groups % p.n_tiles;       // a pair's tiles t, t + groups, ... all have the same N slice (t % n_tiles)
  const int groups = p.total_tiles < max_groups ? p.total_tiles : max_groups;
  {
    static const int l2_env = []() { const char* e = getenv("JMT_GEMM_L2_AHEAD"); return e ? atoi(e) : 1; }();
    p.l2_ahead = p.bres ? l2_env : 0;
    p.l2_mstep = p.bres ? (groups / p.n_tiles) * kBlockM * p.cluster : 0;      // tiles t and t + groups of a pair are groups / n_tiles pair-rows apart
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(groups * p.cluster);
  cfg.blockDim = dim3(64 + 32 * p.epi_warps);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static const bool pdl = []() { const char* e = getenv("JMT_PDL"); return e ? atoi(e) != 0 : true; }();
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  cudaError_t le;
#define JMT_TC_LAUNCH(CTA, MASK, X3, EXT) le = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<CTA, MASK, 8, X3, EXT>, map_a, map_b, map_d, map_a_lo, map_b_lo, p)
#define JMT_TC_LAUNCH_X(CTA, MASK) do { if (x3) JMT_TC_LAUNCH(CTA, MASK, true, 0); else if (ext == 2) JMT_TC_LAUNCH(CTA, MASK, false, 2); \
                                        else if (ext == 1) JMT_TC_LAUNCH(CTA, MASK, false, 1); else JMT_TC_LAUNCH(CTA, MASK, false, 0); } while (0)
  const int ext = p.colsum != nullptr ? 2 : (p.aux != nullptr ? 1 : 0);
  if (p.colmask) { if (p.cluster == 2) JMT_TC_LAUNCH_X(2, true); else JMT_TC_LAUNCH_X(1, true); }
  else { if (p.cluster == 2) JMT_TC_LAUNCH_X(2, false); else JMT_TC_LAUNCH_X(1, false); }
#undef JMT_TC_LAUNCH_X
#undef JMT_TC_LAUNCH
  if (le != cudaSuccess) {
    set_error("jmt_gemm_bf16: cudaLaunchKernelEx: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JMT_ERR_CUDA;
  }
  return check_launch("gemm_tc_kernel");
}

extern "C" int jmt_gemm_bf16(const jmt_gemm_desc* g, void* stream) { return gemm_tc_launch(g, nullptr, nullptr, stream); }

extern "C" int jmt_gemm_bf16x3(const jmt_gemm_desc* g, const void* a_lo, const void* b_lo, void* stream) {
  JMT_REQUIRE(a_lo && b_lo, "jmt_gemm_bf16x3: the lo parts of both operands are required");
  return gemm_tc_launch(g, a_lo, b_lo, stream);
}
