// jmt_attn_chain_bf16: fused attention core for sm_100a -- two chained tcgen05 GEMMs with the row-wise
// softmax algebra in between, nothing but the bf16 probabilities ever leaves the SM.
//
//   forward  (mode 0):  T1 = Q K^T            -> P  = softmax(scale * T1)            -> O  = P V
//   backward (mode 1):  T1 = dO V^T  (= dP)   -> dS = scale * P o (dP - rowsum(P o dP)) -> dQ (+)= dS K
//
// One CTA owns a 128-row query tile of one (batch, head):
//   warp 0 (1 thread)  TMA producer: K-major A1/B1 k-blocks of GEMM1, then MN-major B2 k-blocks of GEMM2, 128B swizzle
//   warp 1 (1 thread)  MMA issuer  : GEMM1 -> T1 (TMEM columns [0, S_pad)), then GEMM2 with A = X from shared memory
//                                    -> T2 (TMEM columns [0, dh)); T1 is dead by then, so dh = 512 (one head) fits
//   warps 2..9         row warps   : thread = TMEM lane = query row, the two warps of a lane quarter split the key
//                                    columns; max / sum (or the P.dP dot) are in-thread reductions plus one shared-
//                                    memory exchange; X (= P or dS, bf16) is written straight into the K-major
//                                    128B-swizzled A-operand layout of GEMM2 and from there to global by TMA (saved
//                                    P for backward / dS for the dK GEMM); finally the T2 epilogue (bf16, TMA store or
//                                    reduce-add).
// Replaces, per attention call, the QK^T GEMM + softmax kernel + PV GEMM (forward) and the dP GEMM + softmax-backward
// kernel + dQ GEMM (backward) of torch's MHA math path (nn.MultiheadAttention call sites
// mm_multi_transformers.py:62,142-167, mm_transformers.py:76,125-134): the fp32 scores (92 MB per attention at
// B=256, T=300) are never materialised.
#include "tc_common.cuh"

namespace jmt {

constexpr int kAtRowWarps = 8;
constexpr int kAtThreads = 64 + 32 * kAtRowWarps;
constexpr int kAtMaxSlots = 6;

struct AtParams {
  int mode;
  int Lq, S, dh, heads, NB;
  int q_tiles, total_tiles;
  int nk1;              // GEMM1 k-blocks (dh / 64)
  int nsplit1, n1;      // GEMM1: MMAs per k-step and N of each (T1 spans nsplit1 * n1 TMEM columns)
  int nkx;              // 64-key chunks of X = GEMM2 k-blocks
  int nh2, n2;          // GEMM2: N groups and N of each (dh = nh2 * n2)
  int slots, slot_bytes, b1_bytes, b2_bytes, x_bytes;   // b1_bytes: ONE N-group of B1 (n1 rows)
  uint32_t idesc1, idesc2;
  float c_exp;          // scale * log2(e)
  float scale;
  const __nv_bfloat16* p_in;   // mode 1: saved probabilities (NB, heads, Lq, x_ld)
  const __nv_bfloat16* do_in;  // mode 1: dO and the saved forward output O (same geometry): delta = rowsum(dO o O)
  const __nv_bfloat16* o_in;
  int64_t do_ld, do_hs, do_bs;
  const float* delta_in;       // mode 1: precomputed delta (NB, heads, Lq) fp32, or NULL (computed in-kernel from dO, O)
  int delta_pdp;               // mode 1 without delta_in / O: delta = rowsum(P o dP) from the staged P and the dP accumulator (one more
                               // pass over TMEM / shared memory instead of a separate rowdot kernel over dO and O)
  int skip2;                   // mode 1: stop after X = dS (no GEMM2 / dQ): dQ, dK, dV are plain GEMMs on the saved dS
  int64_t x_ld;
  int store_mode;
  int store_x;                 // write X (P / dS) to global memory (0: forward-only callers that never run a backward)
  float* lse_out;              // mode 0, optional: log-sum-exp of scale * scores per query row, (NB, heads, Lq) fp32 (natural log):
                               // what a caller needs to merge attention over key CHUNKS (long S, see jmt_attn_merge)
  FastDiv fd_qt, fd_heads;
  unsigned long long* prof;   // optional per-CTA cycle counters (16 per CTA), jmt_attn_set_profile_buffer
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

struct AtTile { int q0, head, b; };
__device__ __forceinline__ AtTile at_decode(const AtParams& p, int t) {
  uint32_t bh, qt, b, head;
  p.fd_qt.divmod((uint32_t)t, bh, qt);
  p.fd_heads.divmod(bh, b, head);
  AtTile r; r.q0 = (int)qt * kBlockM; r.head = (int)head; r.b = (int)b;
  return r;
}

struct RowCtx {
  uint32_t tb, xrow, sw;
  int half, row, nch;
};

// forward, pass 1 on one 32-key chunk: four independent partial maxima
__device__ __forceinline__ void row_fwd_max(const uint32_t (&r)[32], int nvalid, float (&mx)[4]) {
  if (nvalid >= 32) {                      // whole chunk valid (all but the last chunk of a row): no per-element predicate
#pragma unroll
    for (int j = 0; j < 32; ++j) mx[j & 3] = fmaxf(mx[j & 3], __uint_as_float(r[j]));
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) mx[j & 3] = j < nvalid ? fmaxf(mx[j & 3], __uint_as_float(r[j])) : mx[j & 3];
  }
}
// forward, pass 2 on one 32-key chunk: e = exp2(v*c - rowmax*c) -> bf16 -> swizzled X, four independent partial sums
template <bool FULL>
__device__ __forceinline__ void row_fwd_exp_t(const AtParams& p, const uint32_t (&r)[32], int nvalid, float shift, float (&sum)[4],
                                              uint32_t xbase, uint32_t piece0, uint32_t sw) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float x[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const float y = ex2_approx(fmaf(__uint_as_float(r[i * 8 + h]), p.c_exp, -shift));
      x[h] = (FULL || i * 8 + h < nvalid) ? y : 0.f;
      sum[h & 3] += x[h];
    }
    st_shared_v4(xbase + (((piece0 + (uint32_t)i) ^ sw) << 4), pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]),
                 pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
  }
}
__device__ __forceinline__ void row_fwd_exp(const AtParams& p, const uint32_t (&r)[32], int nvalid, float shift, float (&sum)[4],
                                            uint32_t xbase, uint32_t piece0, uint32_t sw) {
  if (nvalid >= 32) row_fwd_exp_t<true>(p, r, nvalid, shift, sum, xbase, piece0, sw);
  else row_fwd_exp_t<false>(p, r, nvalid, shift, sum, xbase, piece0, sw);
}

// backward, one 32-key chunk: dS = scale * P * (dP - delta) with P read from the X tile it overwrites
template <bool FULL>
__device__ __forceinline__ void row_bwd_chunk_t(const AtParams& p, const uint32_t (&r)[32], int nvalid, float delta,
                                                uint32_t xbase, uint32_t piece0, uint32_t sw) {
  const float ds = p.scale * delta;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t addr = xbase + (((piece0 + (uint32_t)i) ^ sw) << 4);
    const uint4 pv = ld_shared_v4(addr);
    const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
    float x[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const float pf = (h & 1) ? bf16hi(w[h >> 1]) : bf16lo(w[h >> 1]);
      const float y = pf * fmaf(p.scale, __uint_as_float(r[i * 8 + h]), -ds);       // scale * P * (dP - delta)
      x[h] = (FULL || i * 8 + h < nvalid) ? y : 0.f;
    }
    st_shared_v4(addr, pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
  }
}
__device__ __forceinline__ void row_bwd_chunk(const AtParams& p, const uint32_t (&r)[32], int nvalid, float delta,
                                              uint32_t xbase, uint32_t piece0, uint32_t sw) {
  if (nvalid >= 32) row_bwd_chunk_t<true>(p, r, nvalid, delta, xbase, piece0, sw);
  else row_bwd_chunk_t<false>(p, r, nvalid, delta, xbase, piece0, sw);
}

// backward, delta pass on one 32-key chunk: acc += P * dP with P read from the X tile (staged by TMA), four independent partial sums
__device__ __forceinline__ void row_bwd_dot(const uint32_t (&r)[32], int nvalid, uint32_t xbase, uint32_t piece0, uint32_t sw, float (&acc)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 pv = ld_shared_v4(xbase + (((piece0 + (uint32_t)i) ^ sw) << 4));
    const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      const float pf = (h & 1) ? bf16hi(w[h >> 1]) : bf16lo(w[h >> 1]);
      const float dp = i * 8 + h < nvalid ? __uint_as_float(r[i * 8 + h]) : 0.f;
      acc[h & 3] = fmaf(pf, dp, acc[h & 3]);
    }
  }
}

// The row-wise algebra between the two GEMMs for one thread (= one query row, every other 32-key chunk):
//   MODE 0: X = softmax(scale * T1): two streaming passes over T1 (row max, then exp2 + row sum; each with one exchange
//           with the warp that owns the other chunks of the row), then the bf16 values are normalised in shared memory.
//           (A single-pass online-softmax variant was measured slower: its per-chunk max -> exp dependency serialises
//           the two warps an SMSP has.)
//   MODE 1: X = scale * P o (T1 - delta), ONE pass: delta = rowsum(dO o O) (= rowsum(P o dP)) was precomputed into
//           `delta_buf` while the previous tile's GEMM2 ran, P was staged in X by TMA.
// X ends up in shared memory as bf16 in the 128B-swizzled K-major A-operand layout of GEMM2.  The TMEM load of chunk
// i+1 is in flight while chunk i is processed (two register buffers).
template <int MODE>
__device__ __forceinline__ void row_op(const AtParams& p, const RowCtx& rc, float* red, const float* delta_buf, float* lse_row) {
  const int half = rc.half, row = rc.row, nch = rc.nch;
  uint32_t ra[32], rb[32];
  float shift = 0.f;
  if (MODE == 0) {
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (half < nch) tc_ld32_issue(rc.tb + half * 32, ra);
    for (int ch = half; ch < nch; ch += 4) {
      tc_wait_ld();
      if (ch + 2 < nch) tc_ld32_issue(rc.tb + (ch + 2) * 32, rb);
      row_fwd_max(ra, p.S - ch * 32, mx);
      if (ch + 2 >= nch) break;
      tc_wait_ld();
      if (ch + 4 < nch) tc_ld32_issue(rc.tb + (ch + 4) * 32, ra);
      row_fwd_max(rb, p.S - (ch + 2) * 32, mx);
    }
    red[half * 128 + row] = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
    __syncwarp();
    if (half < nch) tc_ld32_issue(rc.tb + half * 32, ra);        // pass 2's first chunk is in flight across the barrier
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kAtRowWarps) : "memory");
    shift = fmaxf(red[row], red[128 + row]) * p.c_exp;
  } else if (p.delta_pdp) {
    // delta_i = sum_j P_ij dP_ij: one pass over this warp's chunks of the dP accumulator and the staged P, then the exchange with
    // the warp that owns the other chunks of the row (same structure as the forward's row-max pass)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (half < nch) tc_ld32_issue(rc.tb + half * 32, ra);
    for (int ch = half; ch < nch; ch += 4) {
      tc_wait_ld();
      if (ch + 2 < nch) tc_ld32_issue(rc.tb + (ch + 2) * 32, rb);
      row_bwd_dot(ra, p.S - ch * 32, rc.xrow + (ch >> 1) * 16384, (uint32_t)((ch & 1) * 4), rc.sw, acc);
      if (ch + 2 >= nch) break;
      tc_wait_ld();
      if (ch + 4 < nch) tc_ld32_issue(rc.tb + (ch + 4) * 32, ra);
      row_bwd_dot(rb, p.S - (ch + 2) * 32, rc.xrow + ((ch + 2) >> 1) * 16384, (uint32_t)((ch & 1) * 4), rc.sw, acc);
    }
    red[half * 128 + row] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    __syncwarp();
    if (half < nch) tc_ld32_issue(rc.tb + half * 32, ra);        // the dS pass's first chunk is in flight across the barrier
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kAtRowWarps) : "memory");
    shift = red[row] + red[128 + row];
  } else {
    shift = delta_buf[row];
    if (half < nch) tc_ld32_issue(rc.tb + half * 32, ra);
  }
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ch = half; ch < nch; ch += 4) {
    tc_wait_ld();
    if (ch + 2 < nch) tc_ld32_issue(rc.tb + (ch + 2) * 32, rb);
    if (MODE == 0) row_fwd_exp(p, ra, p.S - ch * 32, shift, sum, rc.xrow + (ch >> 1) * 16384, (uint32_t)((ch & 1) * 4), rc.sw);
    else row_bwd_chunk(p, ra, p.S - ch * 32, shift, rc.xrow + (ch >> 1) * 16384, (uint32_t)((ch & 1) * 4), rc.sw);
    if (ch + 2 >= nch) break;
    tc_wait_ld();
    if (ch + 4 < nch) tc_ld32_issue(rc.tb + (ch + 4) * 32, ra);
    if (MODE == 0) row_fwd_exp(p, rb, p.S - (ch + 2) * 32, shift, sum, rc.xrow + ((ch + 2) >> 1) * 16384, (uint32_t)((ch & 1) * 4), rc.sw);
    else row_bwd_chunk(p, rb, p.S - (ch + 2) * 32, shift, rc.xrow + ((ch + 2) >> 1) * 16384, (uint32_t)((ch & 1) * 4), rc.sw);
  }
  if (half == 0) {                               // zero the key columns [32 * nch, 64 * nkx) nobody computes
    for (int pc = nch * 4; pc < p.nkx * 8; ++pc)
      st_shared_v4(rc.xrow + (pc >> 3) * 16384 + ((((uint32_t)(pc & 7)) ^ rc.sw) << 4), 0u, 0u, 0u, 0u);
  }
  if (MODE == 0) {
    // ---- pass 3 (shared memory only): normalise this thread's pieces in place by 1 / rowsum
    red[256 + half * 128 + row] = (sum[0] + sum[1]) + (sum[2] + sum[3]);
    __syncwarp();
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kAtRowWarps) : "memory");
    const float total = red[256 + row] + red[256 + 128 + row];
    const float inv = 1.f / total;
    // log-sum-exp of this row's scaled scores: shift is rowmax * scale * log2(e), the exponentials were taken base 2
    if (lse_row != nullptr && half == 0) *lse_row = (shift + log2f(total)) * 0.6931471805599453f;
    for (int ch = half; ch < nch; ch += 2) {
      const uint32_t xbase = rc.xrow + (ch >> 1) * 16384;
      uint4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = ld_shared_v4(xbase + ((((uint32_t)((ch & 1) * 4 + i)) ^ rc.sw) << 4));
#pragma unroll
      for (int i = 0; i < 4; ++i)
        st_shared_v4(xbase + ((((uint32_t)((ch & 1) * 4 + i)) ^ rc.sw) << 4),
                     pack_bf16(bf16lo(v[i].x) * inv, bf16hi(v[i].x) * inv), pack_bf16(bf16lo(v[i].y) * inv, bf16hi(v[i].y) * inv),
                     pack_bf16(bf16lo(v[i].z) * inv, bf16hi(v[i].z) * inv), pack_bf16(bf16lo(v[i].w) * inv, bf16hi(v[i].w) * inv));
    }
  }
}

// delta_i = sum_d dO[i, d] * O[i, d] for the 128 rows of one tile: one warp per row, coalesced 16-byte loads, 4 rows in flight
__device__ __forceinline__ void at_delta(const AtParams& p, const AtTile& c, int ew, int lane, float* out) {
  const int64_t hb = (int64_t)c.b * p.do_bs + (int64_t)c.head * p.do_hs;
  const int npc = p.dh >> 3;
  for (int r0 = ew; r0 < kBlockM; r0 += 4 * kAtRowWarps) {
    float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + u * kAtRowWarps;
      if (c.q0 + r < p.Lq) {
        const int64_t off = hb + (int64_t)(c.q0 + r) * p.do_ld;
        for (int pc = lane; pc < npc; pc += 32) {
          const uint4 x = __ldg(reinterpret_cast<const uint4*>(p.do_in + off + pc * 8));
          const uint4 y = __ldg(reinterpret_cast<const uint4*>(p.o_in + off + pc * 8));
          d[u] = fmaf(bf16lo(x.x), bf16lo(y.x), d[u]); d[u] = fmaf(bf16hi(x.x), bf16hi(y.x), d[u]);
          d[u] = fmaf(bf16lo(x.y), bf16lo(y.y), d[u]); d[u] = fmaf(bf16hi(x.y), bf16hi(y.y), d[u]);
          d[u] = fmaf(bf16lo(x.z), bf16lo(y.z), d[u]); d[u] = fmaf(bf16hi(x.z), bf16hi(y.z), d[u]);
          d[u] = fmaf(bf16lo(x.w), bf16lo(y.w), d[u]); d[u] = fmaf(bf16hi(x.w), bf16hi(y.w), d[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float t = warp_sum(d[u]);
      if (lane == 0) out[r0 + u * kAtRowWarps] = t;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(kAtThreads, 1)
attn_chain_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
                  const __grid_constant__ CUtensorMap map_b2, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_p, const AtParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // The kernel has no static shared memory, so the dynamic window starts 1024-byte aligned (after the driver's
  // reserved 1 KiB); the 2-slot configuration at S = 300, dh = 512 has no room for alignment slack.  Fail loudly if not.
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t sX = base;                                       // nkx chunks of [128 rows x 128 B]
  const uint32_t sStage = sX;                                      // 8 x 4 KiB epilogue staging aliases X (dead after GEMM2)
  const uint32_t sRing = sX + p.x_bytes;
  const uint32_t sRed = sRing + p.slots * p.slot_bytes;            // 2 x [2][128] floats
  const uint32_t bars = sRed + 2048;     // mode 0: [0,256) max / [256,512) sum exchange; mode 1: [0,256) two delta buffers
  const uint32_t full_bar = bars, empty_bar = bars + 8 * kAtMaxSlots;
  const uint32_t t1_full = bars + 16 * kAtMaxSlots, x_ready = t1_full + 8, t2_full = t1_full + 16, t_empty = t1_full + 24;
  const uint32_t p_full = t1_full + 32;
  const uint32_t tmem_slot = t1_full + 40;
  uint8_t* smem_aligned = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_aligned + (tmem_slot - base));
  float* red = reinterpret_cast<float*>(smem_aligned + (sRed - base));     // [0..255] max / dot, [256..511] sum

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // (shfl: lets ptxas treat the warp index as warp-uniform)

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.slots; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    mbar_init(t1_full, 1); mbar_init(t2_full, 1); mbar_init(p_full, 1);
    mbar_init(x_ready, kAtRowWarps); mbar_init(t_empty, kAtRowWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    {
      // ================================ TMA producer (whole warp, one elected lane issues) ================================
      const uint32_t el = elect_one();
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b2)) : "memory");
      int slot = 0; uint32_t phase = 0;
      long long pw = 0; const long long pt0 = p.prof ? clock64() : 0;
      const int b2_chunks = p.n2 >> 6;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const AtTile c = at_decode(p, t);
        for (int kb = 0; kb < p.nk1; ++kb) {                  // GEMM1 operands: A1 [128 x 64] + B1 group 0, then B1 group 1
          for (int s = 0; s < p.nsplit1; ++s) {

            const long long tw = p.prof ? clock64() : 0;
            mbar_wait(empty_bar + 8 * slot, phase ^ 1);
            if (p.prof) pw += clock64() - tw;
            const uint32_t fb = full_bar + 8 * slot;
            const uint32_t dst = sRing + slot * p.slot_bytes;
            if (s == 0) {
              mbar_expect_tx_el(el, fb, 16384 + p.b1_bytes);
              tma_load_4d_el(el, dst, &map_a1, fb, kb * kBlockK, c.q0, c.head, c.b);
              tma_load_4d_el(el, dst + 16384, &map_b1, fb, kb * kBlockK, 0, c.head, c.b);
            } else {
              mbar_expect_tx_el(el, fb, p.b1_bytes);
              tma_load_4d_el(el, dst, &map_b1, fb, kb * kBlockK, p.n1, c.head, c.b);
            }
            if (++slot == p.slots) { slot = 0; phase ^= 1; }
          }
        }
        for (int nh = 0; nh < (p.skip2 ? 0 : p.nh2); ++nh)    // GEMM2 operand: B2 [64 keys x n2] (MN-major chunks)
          for (int kb = 0; kb < p.nkx; ++kb) {
            const long long tw = p.prof ? clock64() : 0;
            mbar_wait(empty_bar + 8 * slot, phase ^ 1);
            if (p.prof) pw += clock64() - tw;
            const uint32_t fb = full_bar + 8 * slot;
            const uint32_t dst = sRing + slot * p.slot_bytes;
            mbar_expect_tx_el(el, fb, p.b2_bytes);
            tma_load_5d_el(el, dst, &map_b2, fb, 0, kb * kBlockK, nh * b2_chunks, c.head, c.b);     // all n2 / 64 chunks in one instruction
            if (++slot == p.slots) { slot = 0; phase ^= 1; }
          }
      }
      if (p.prof && lane == 0) { p.prof[blockIdx.x * 16 + 0] = pw; p.prof[blockIdx.x * 16 + 1] = clock64() - pt0; }
    }
  } else if (warp == 1) {
    {
      // ================================ MMA issuer (whole warp, one elected lane issues) ================================
      const uint32_t el = elect_one();
      int slot = 0; uint32_t phase = 0;
      int tile_iter = 0;
      long long mw1 = 0, mw2 = 0, mwx = 0, mwe = 0; const long long mt0 = p.prof ? clock64() : 0;
#define AT_TIMED_WAIT(acc, bar, ph) { const long long tw_ = p.prof ? clock64() : 0; mbar_wait(bar, ph); if (p.prof) acc += clock64() - tw_; }
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tile_iter) {
        const uint32_t par = tile_iter & 1;
        if (!p.skip2) AT_TIMED_WAIT(mwe, t_empty, par ^ 1);   // previous tile's T2 has been drained
        tc_fence_after();
        for (int kb = 0; kb < p.nk1; ++kb) {                  // GEMM1: T1 = A1 B1^T, one stage per N group of B1
          AT_TIMED_WAIT(mw1, full_bar + 8 * slot, phase);
          tc_fence_after();
          const int slot_a = slot;
          const uint32_t st = sRing + slot * p.slot_bytes;
          const uint64_t a_desc = make_smem_desc(st, 16, 1024);
          const uint64_t b_desc = make_smem_desc(st + 16384, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            tc_mma_elect<1>(el, tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), p.idesc1, (kb > 0 || k > 0) ? 1u : 0u);
          if (++slot == p.slots) { slot = 0; phase ^= 1; }
          if (p.nsplit1 == 2) {
            AT_TIMED_WAIT(mw1, full_bar + 8 * slot, phase);
            tc_fence_after();
            const uint64_t b2_desc = make_smem_desc(sRing + slot * p.slot_bytes, 16, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              tc_mma_elect<1>(el, tmem_base + p.n1, a_desc + (uint64_t)(k * 2), b2_desc + (uint64_t)(k * 2), p.idesc1, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit_elect<1>(el, empty_bar + 8 * slot_a);             // the A tile is shared by both groups: freed only now
            tc_commit_elect<1>(el, empty_bar + 8 * slot);
            if (++slot == p.slots) { slot = 0; phase ^= 1; }
          } else {
            tc_commit_elect<1>(el, empty_bar + 8 * slot_a);
          }
        }
        tc_commit_elect<1>(el, t1_full);
        AT_TIMED_WAIT(mwx, x_ready, par);                     // X (P or dS) is in shared memory, T1 is dead
        tc_fence_after();
        if (p.skip2) continue;                                // T1 has been consumed (x_ready): next tile's GEMM1 may overwrite it
        for (int nh = 0; nh < p.nh2; ++nh)                    // GEMM2: T2[:, nh] = X B2[:, nh]
          for (int kb = 0; kb < p.nkx; ++kb) {
            AT_TIMED_WAIT(mw2, full_bar + 8 * slot, phase);
            tc_fence_after();
            const uint64_t a_desc = make_smem_desc(sX + kb * 16384, 16, 1024);
            const uint64_t b_desc = make_smem_desc(sRing + slot * p.slot_bytes, 8192, 1024);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              tc_mma_elect<1>(el, tmem_base + nh * p.n2, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 128), p.idesc2,
                        (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit_elect<1>(el, empty_bar + 8 * slot);
            if (++slot == p.slots) { slot = 0; phase ^= 1; }
          }
        tc_commit_elect<1>(el, t2_full);
      }
      if (p.prof && lane == 0) {
        unsigned long long* o = p.prof + blockIdx.x * 16;
        o[2] = mw1; o[3] = mw2; o[4] = mwx; o[5] = mwe; o[6] = clock64() - mt0;
      }
    }
  } else {
    // ================================ row warps (2..9) ================================
    const int q = warp & 3;                        // TMEM lane quarter
    const int ew = warp - 2;
    const int half = ew >> 2;                      // which interleaved set of column chunks this warp owns
    const int row = q * 32 + lane;                 // row inside the 128-row query tile
    const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sw = lane & 7;
    const uint32_t xrow = sX + row * 128;          // this row inside X chunk 0 (chunk kc: + kc * 16384)
    const uint32_t stage = sStage + ew * 4096;
    const int nch = (p.S + 31) >> 5;               // 32-column chunks that contain valid keys
    long long rw1 = 0, rrow = 0, rw2 = 0, repi = 0; const long long rt0 = p.prof ? clock64() : 0;
    int tile_iter = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++tile_iter) {
      const AtTile c = at_decode(p, t);
      const uint32_t par = tile_iter & 1;
      const int grow = c.q0 + row;
      if (MODE == 1) {
        // Stage the saved probabilities of this tile into X by TMA while GEMM1 (dP = dO V^T) runs: same 128B-swizzled
        // chunk layout that dS will overwrite in place.  X is free: GEMM2 of the previous tile has completed (its
        // t2_full was consumed) and, after this barrier, every warp's TMA stores out of X / the staging tiles have
        // finished reading it.
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kAtRowWarps) : "memory");
        if (ew == 0 && lane == 0) {
          mbar_expect_tx(p_full, p.nkx * 16384);
          for (int kc = 0; kc < p.nkx; ++kc) tma_load_4d(sX + kc * 16384, &map_p, p_full, kc * 64, c.q0, c.head, c.b);
        }
        // delta of this tile while GEMM1 runs.  (Measured: ~9 k cycles per tile of latency-bound global loads that GEMM1
        // only partly hides; computing it for the next tile under GEMM2 instead put 30 k cycles on the critical path.)
        if (p.delta_pdp) {
          // (delta comes out of the row pass itself)
        } else if (p.delta_in != nullptr) {       // precomputed (jmt_rowdot): one coalesced load per row
          if (ew == 0) {
            const float* dl = p.delta_in + ((int64_t)c.b * p.heads + c.head) * p.Lq;
            for (int r = lane; r < kBlockM; r += 32) red[r] = c.q0 + r < p.Lq ? __ldg(dl + c.q0 + r) : 0.f;
          }
        } else {
          at_delta(p, c, ew, lane, red);
        }
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kAtRowWarps) : "memory");     // delta of every row is visible
      }
      const long long ra = p.prof ? clock64() : 0;
      mbar_wait(t1_full, par);
      if (MODE == 1) mbar_wait(p_full, par);         // the P tile has landed in X
      __syncwarp();
      tc_fence_after();
      const long long rb = p.prof ? clock64() : 0;
      RowCtx rc;
      rc.tb = tb; rc.xrow = xrow; rc.sw = sw; rc.half = half; rc.row = row; rc.nch = nch;
      if (lane == 0) bulk_wait_read0();              // this warp's / the X stores of the previous tile have finished reading X
      float* lse_row = nullptr;
      if (MODE == 0 && p.lse_out != nullptr && grow < p.Lq)
        lse_row = p.lse_out + ((int64_t)c.b * p.heads + c.head) * p.Lq + grow;
      row_op<MODE>(p, rc, red, red, lse_row);
      fence_async_smem();                            // X visible to the tensor core / TMA (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_ready);
      if (p.store_x && ew == 0 && lane == 0) {
        // saved probabilities (forward) / dS (backward): X chunks -> global, clipped to (S, Lq) by the tensor map
        mbar_wait(x_ready, par);
        for (int kc = 0; kc < p.nkx; ++kc) tma_store_4d(&map_x, sX + kc * 16384, kc * 64, c.q0, c.head, c.b);
        bulk_commit();
      }
      if (p.skip2) continue;                         // dS-only mode: no GEMM2, nothing to drain
      // ---- T2 epilogue: O (store) / dQ (reduce-add), bf16
      const long long rcx = p.prof ? clock64() : 0;
      mbar_wait(t2_full, par);                       // GEMM2 has consumed X: its first 32 KiB become the epilogue staging
      if (ew == 0 && lane == 0) bulk_wait_read0();   // ... once the X -> global stores have finished reading it too
      __syncwarp();                                  // (lane 0 of warp 2 rejoins after issuing the X stores)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kAtRowWarps) : "memory");
      tc_fence_after();
      const long long rd = p.prof ? clock64() : 0;
      const bool rows_valid = c.q0 + q * 32 < p.Lq;        // warp-uniform
      {
        uint32_t r0[32], r1[32];
        int buf = 0;
        int c0 = half * 64;
        if (c0 < p.dh) { tc_ld32_issue(tb + c0, r0); tc_ld32_issue(tb + c0 + 32, r1); }
        for (; c0 < p.dh; c0 += 128, buf ^= 1) {
          const uint32_t stg = stage + buf * (kAtRowWarps * 4096);     // two staging tiles per warp: store i overlaps chunk i+1
          const uint32_t stg_row = stg + lane * 128;
          if (lane == 0) bulk_wait_read1();          // the store issued two chunks ago has finished reading this buffer
          __syncwarp();
          tc_wait_ld();
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            pk[j >> 1] = pack_bf16(__uint_as_float(r0[j]), __uint_as_float(r0[j + 1]));
            pk[16 + (j >> 1)] = pack_bf16(__uint_as_float(r1[j]), __uint_as_float(r1[j + 1]));
          }
          if (c0 + 128 < p.dh) { tc_ld32_issue(tb + c0 + 128, r0); tc_ld32_issue(tb + c0 + 160, r1); }   // next chunk in flight
#pragma unroll
          for (int pc = 0; pc < 8; ++pc)
            st_shared_v4(stg_row + ((((uint32_t)pc) ^ sw) << 4), pk[4 * pc], pk[4 * pc + 1], pk[4 * pc + 2], pk[4 * pc + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && rows_valid) {
            if (p.store_mode == JMT_STORE) tma_store_4d(&map_d, stg, c0, c.q0 + q * 32, c.head, c.b);
            else tma_reduce_add_4d(&map_d, stg, c0, c.q0 + q * 32, c.head, c.b);
          }
          if (lane == 0) bulk_commit();              // (possibly empty group: keeps the wait_group.read 1 accounting uniform)
        }
        tc_wait_ld();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty);
      if (p.prof) { const long long re = clock64(); rw1 += rb - ra; rrow += rcx - rb; rw2 += rd - rcx; repi += re - rd; }
    }
    if (p.prof && ew == 0 && lane == 0) {
      unsigned long long* o = p.prof + blockIdx.x * 16;
      o[7] = rw1; o[8] = rrow; o[9] = rw2; o[10] = repi; o[11] = clock64() - rt0;
    }
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// geometry shared by the support query and the launcher
static int at_plan(const jmt_attn_desc* g, AtParams* p, int* smem_bytes) {
  memset(p, 0, sizeof(*p));
  if (g->dh < 64 || g->dh > 512 || (g->dh & (g->dh - 1)) != 0) return 0;      // 64, 128, 256, 512
  if (g->S < 1 || g->S > 512 || g->Lq < 1 || g->heads < 1 || g->NB < 1) return 0;
  p->mode = g->mode; p->Lq = g->Lq; p->S = g->S; p->dh = g->dh; p->heads = g->heads; p->NB = g->NB;
  p->q_tiles = (g->Lq + kBlockM - 1) / kBlockM;
  const int64_t total = (int64_t)p->q_tiles * g->heads * g->NB;
  if (total >= (1ll << 31)) return 0;
  p->total_tiles = (int)total;
  p->nk1 = g->dh / 64;
  const int s16 = (g->S + 15) / 16 * 16;
  if (s16 <= 256) { p->nsplit1 = 1; p->n1 = s16; }
  else { p->nsplit1 = 2; p->n1 = (g->S + 31) / 32 * 16; }
  p->nkx = (g->S + 63) / 64;
  p->n2 = g->dh > 256 ? 256 : g->dh;
  p->nh2 = g->dh / p->n2;
  p->b1_bytes = p->n1 * 128;
  p->b2_bytes = p->n2 * 128;
  int slot = 16384 + p->b1_bytes;
  if (p->b2_bytes > slot) slot = p->b2_bytes;
  p->slot_bytes = (slot + 1023) / 1024 * 1024;
  p->x_bytes = p->nkx * 16384;
  if (p->x_bytes < 2 * kAtRowWarps * 4096) p->x_bytes = 2 * kAtRowWarps * 4096;   // the epilogue staging (2 tiles per warp) aliases X
  const int fixed = p->x_bytes + 2048 + 256;
  const int avail = 227 * 1024 - fixed;
  int slots = avail / p->slot_bytes;
  if (slots < 2) return 0;
  if (slots > kAtMaxSlots) slots = kAtMaxSlots;
  p->slots = slots;
  *smem_bytes = fixed + slots * p->slot_bytes;
  p->idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p->n1 >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
  p->idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(p->n2 >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
  p->scale = g->scale;
  p->c_exp = g->scale * 1.4426950408889634f;
  p->p_in = (const __nv_bfloat16*)g->p_in;
  p->do_in = (const __nv_bfloat16*)g->a1; p->o_in = (const __nv_bfloat16*)g->o_in;
  p->do_ld = g->a1_ld; p->do_hs = g->a1_hs; p->do_bs = g->a1_bs;
  p->delta_in = g->delta_in; p->skip2 = g->d == nullptr ? 1 : 0;
  p->delta_pdp = (g->mode == 1 && g->delta_in == nullptr && g->o_in == nullptr) ? 1 : 0;
  p->x_ld = g->x_ld;
  p->store_mode = g->store_mode;
  p->store_x = g->x != nullptr ? 1 : 0;
  p->lse_out = g->mode == 0 ? g->lse_out : nullptr;
  p->fd_qt.init(p->q_tiles);
  p->fd_heads.init(g->heads);
  return 1;
}

}  // namespace jmt

using namespace jmt;

namespace jmt {
// out[(b*heads + h)*rows + r] = sum_d A[b, h, r, d] * B[b, h, r, d]  (delta_i = dO_i . O_i of the attention backward):
// one warp per (b, h, r), coalesced 16-byte loads
__global__ void __launch_bounds__(256)
rowdot_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, int64_t ld, int64_t hs, int64_t bs,
              int rows, int heads, int dh, int64_t total, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  for (int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); w < total; w += (int64_t)gridDim.x * 8) {
    const int64_t bh = w / rows; const int r = (int)(w - bh * rows);
    const int64_t bi = bh / heads; const int h = (int)(bh - bi * heads);
    const int64_t off = bi * bs + (int64_t)h * hs + (int64_t)r * ld;
    float d = 0.f;
    for (int pc = lane; pc < (dh >> 3); pc += 32) {
      const uint4 x = __ldg(reinterpret_cast<const uint4*>(a + off + pc * 8));
      const uint4 y = __ldg(reinterpret_cast<const uint4*>(b + off + pc * 8));
      d = fmaf(bf16lo(x.x), bf16lo(y.x), d); d = fmaf(bf16hi(x.x), bf16hi(y.x), d);
      d = fmaf(bf16lo(x.y), bf16lo(y.y), d); d = fmaf(bf16hi(x.y), bf16hi(y.y), d);
      d = fmaf(bf16lo(x.z), bf16lo(y.z), d); d = fmaf(bf16hi(x.z), bf16hi(y.z), d);
      d = fmaf(bf16lo(x.w), bf16lo(y.w), d); d = fmaf(bf16hi(x.w), bf16hi(y.w), d);
    }
    d = warp_sum(d);
    if (lane == 0) out[w] = d;
  }
}
}  // namespace jmt

extern "C" int jmt_rowdot_bf16(const void* a, const void* b, int64_t ld, int64_t hs, int64_t bs, int NB, int heads, int rows, int dh,
                               float* out, void* stream) {
  JMT_REQUIRE(a && b && out && NB >= 1 && heads >= 1 && rows >= 1 && dh >= 8 && dh % 8 == 0, "jmt_rowdot_bf16: bad arguments");
  JMT_REQUIRE(ld % 8 == 0 && hs % 8 == 0 && bs % 8 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0,
              "jmt_rowdot_bf16: 16-byte aligned geometry required");
  const int64_t total = (int64_t)NB * heads * rows;
  rowdot_kernel<<<grid_for(total, 8, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, ld, hs, bs,
                                                                                  rows, heads, dh, total, out);
  return check_launch("rowdot_kernel");
}

namespace jmt {
// Merge of attention computed over key CHUNKS (long S): chunk c gave O_c = softmax_c(s) V_c and lse_c = logsumexp_c(s) per query row;
// softmax over all keys = sum_c exp(lse_c - lse) O_c with lse = logsumexp_c(lse_c).  Running form, one warp per (b, head, query row):
//   lse' = logaddexp(lse_acc, lse_c);  acc' = acc * exp(lse_acc - lse') + O_c * exp(lse_c - lse')
// acc is fp32 (rows, E) contiguous with the rows of O; the last chunk writes the bf16 result instead of acc.
__global__ void __launch_bounds__(256)
attn_merge_kernel(const __nv_bfloat16* __restrict__ oc, int64_t ld, int64_t hs, int64_t bs, const float* __restrict__ lse_c,
                  float* __restrict__ acc, float* __restrict__ lse_acc, __nv_bfloat16* __restrict__ out, int first,
                  int rows, int heads, int dh, int64_t total) {
  const int lane = threadIdx.x & 31;
  for (int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); w < total; w += (int64_t)gridDim.x * 8) {
    const int64_t bh = w / rows; const int r = (int)(w - bh * rows);
    const int64_t bi = bh / heads; const int h = (int)(bh - bi * heads);
    const int64_t off = bi * bs + (int64_t)h * hs + (int64_t)r * ld;
    const float lc = lse_c[w];
    float wa = 0.f, wc = 1.f, ln = lc;
    if (!first) {
      const float la = lse_acc[w];
      const float m = fmaxf(la, lc);
      ln = m + __logf(__expf(la - m) + __expf(lc - m));
      wa = __expf(la - ln); wc = __expf(lc - ln);
    }
    for (int pc = lane; pc < (dh >> 3); pc += 32) {
      Vec8<__nv_bfloat16> x; x.load(oc + off + pc * 8);
      Vec8<float> a;
      if (!first) {
        a.load(acc + off + pc * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) a.v[k] = fmaf(a.v[k], wa, x.v[k] * wc);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) a.v[k] = x.v[k];
      }
      if (out != nullptr) {
        Vec8<__nv_bfloat16> o;
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = a.v[k];
        o.store(out + off + pc * 8);
      } else {
        a.store(acc + off + pc * 8);
      }
    }
    if (lane == 0) lse_acc[w] = ln;
  }
}
}  // namespace jmt

extern "C" int jmt_attn_merge(const void* o_chunk, int64_t ld, int64_t hs, int64_t bs, const float* lse_chunk, float* acc,
                              float* lse_acc, void* out, int first, int NB, int heads, int rows, int dh, void* stream) {
  JMT_REQUIRE(o_chunk && lse_chunk && acc && lse_acc && NB >= 1 && heads >= 1 && rows >= 1 && dh >= 8 && dh % 8 == 0,
              "jmt_attn_merge: bad arguments");
  JMT_REQUIRE(ld % 8 == 0 && hs % 8 == 0 && bs % 8 == 0 &&
              ((reinterpret_cast<uintptr_t>(o_chunk) | reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(out)) & 31) == 0,
              "jmt_attn_merge: 32-byte aligned geometry required");
  const int64_t total = (int64_t)NB * heads * rows;
  attn_merge_kernel<<<grid_for(total, 8, kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)o_chunk, ld, hs, bs, lse_chunk, acc, lse_acc, (__nv_bfloat16*)out, first, rows, heads, dh, total);
  return check_launch("attn_merge_kernel");
}

static std::atomic<unsigned long long*> g_attn_prof{nullptr};
extern "C" int jmt_attn_set_profile_buffer(void* dev_buf) {
  g_attn_prof.store((unsigned long long*)dev_buf);
  return JMT_OK;
}

extern "C" int jmt_attn_chain_supported(const jmt_attn_desc* g) {
  if (!g) return 0;
  AtParams p; int smem = 0;
  if (!at_plan(g, &p, &smem)) return 0;
  if (g->x != nullptr && (g->x_ld % 8 != 0 || g->x_ld < g->S)) return 0;
  return 1;
}

extern "C" int jmt_attn_chain_bf16(const jmt_attn_desc* g, void* stream) {
  JMT_REQUIRE(g && g->a1 && g->b1, "jmt_attn_chain_bf16: null pointer");
  JMT_REQUIRE(g->x || (g->mode == 0 && g->d), "jmt_attn_chain_bf16: X may only be omitted by a forward call that produces D");
  JMT_REQUIRE(g->mode == 0 || g->mode == 1, "jmt_attn_chain_bf16: bad mode");
  JMT_REQUIRE(g->d == nullptr || g->b2, "jmt_attn_chain_bf16: D needs B2");
  JMT_REQUIRE(g->mode == 0 || g->p_in, "jmt_attn_chain_bf16: mode 1 needs the saved probabilities");
  JMT_REQUIRE(g->store_mode == JMT_STORE || g->store_mode == JMT_ACCUMULATE, "jmt_attn_chain_bf16: bad store_mode");
  AtParams p; int smem = 0;
  if (!at_plan(g, &p, &smem) || (g->x != nullptr && (g->x_ld % 8 != 0 || g->x_ld < g->S))) {
    set_error("jmt_attn_chain_bf16: unsupported geometry (dh=%d S=%d Lq=%d x_ld=%lld): use the unfused path", g->dh, g->S, g->Lq,
              (long long)g->x_ld);
    return JMT_ERR_UNSUPPORTED;
  }
  p.prof = g_attn_prof.load();
  CUtensorMap ma1, mb1, mb2, mx, md;
  int rc = make_map(&ma1, g->a1, g->dh, g->Lq, g->a1_ld, g->heads, g->a1_hs, g->NB, g->a1_bs, kBlockM, "jmt_attn_chain_bf16(A1)");
  if (rc != JMT_OK) return rc;
  rc = make_map(&mb1, g->b1, g->dh, g->S, g->b1_ld, g->heads, g->b1_hs, g->NB, g->b1_bs, p.n1, "jmt_attn_chain_bf16(B1)");
  if (rc != JMT_OK) return rc;
  mb2 = mb1;                    // (unused in the dS-only mode)
  if (!p.skip2) {
    rc = make_map_mn5(&mb2, g->b2, g->dh, g->S, g->b2_ld, g->heads, g->b2_hs, g->NB, g->b2_bs, kBlockK, p.n2 / 64, "jmt_attn_chain_bf16(B2)");
    if (rc != JMT_OK) return rc;
  }
  mx = ma1;                     // (unused when X is not stored)
  if (g->x != nullptr) {
    rc = make_map(&mx, g->x, g->S, g->Lq, g->x_ld, g->heads, (int64_t)g->Lq * g->x_ld, g->NB, (int64_t)g->heads * g->Lq * g->x_ld, kBlockM,
                  "jmt_attn_chain_bf16(X)");
    if (rc != JMT_OK) return rc;
  }
  md = mx;
  if (!p.skip2) {
    JMT_REQUIRE((reinterpret_cast<uintptr_t>(g->d) & 15) == 0 && g->d_ld % 8 == 0 && g->d_hs % 8 == 0 && g->d_bs % 8 == 0,
                "jmt_attn_chain_bf16: D geometry must be 16-byte aligned");
    rc = make_map_d(&md, g->d, JMT_BF16, g->dh, g->Lq, g->d_ld, g->heads, g->d_hs, g->NB, g->d_bs, "jmt_attn_chain_bf16(D)");
    if (rc != JMT_OK) return rc;
  }
  static std::atomic<int> attr_set[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); set_error("jmt_attn_chain_bf16: no CUDA device"); return JMT_ERR_CUDA; }
  if (!attr_set[dev & 63].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(attn_chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("jmt_attn_chain_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); cudaGetLastError(); return JMT_ERR_CUDA; }
    attr_set[dev & 63].store(1, std::memory_order_release);
  }
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  CUtensorMap mp = mx;          // mode 1: the saved probabilities, same geometry as X
  if (p.mode == 1) {
    rc = make_map(&mp, g->p_in, g->S, g->Lq, g->x_ld, g->heads, (int64_t)g->Lq * g->x_ld, g->NB, (int64_t)g->heads * g->Lq * g->x_ld, kBlockM,
                  "jmt_attn_chain_bf16(P)");
    if (rc != JMT_OK) return rc;
  }
  if (p.mode == 0) attn_chain_kernel<0><<<grid, kAtThreads, smem, (cudaStream_t)stream>>>(ma1, mb1, mb2, mx, md, mp, p);
  else attn_chain_kernel<1><<<grid, kAtThreads, smem, (cudaStream_t)stream>>>(ma1, mb1, mb2, mx, md, mp, p);
  return check_launch("attn_chain_kernel");
}
