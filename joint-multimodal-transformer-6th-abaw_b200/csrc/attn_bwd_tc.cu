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
//     stage their own 128 rows of A and HALF of X (64 of a piece's 128 columns);
//   * N axis = the 300-extent, cut into pieces of 128 columns (last piece narrowed to a multiple of 16); the A tile of an M tile
//     (all of its <= 5 k-blocks, 80 KB) stays resident in shared memory for all pieces, only X is streamed through a ring;
//   * every (M tile, piece) accumulates into one of FOUR 128-column TMEM slots, so the epilogue of piece i overlaps the MMAs of
//     pieces i+1 .. i+3 -- also across tiles;
//   * epilogue: thread = TMEM lane = head-dimension index; it writes its 64 columns as a COLUMN of a [64 n][32 m] bf16 staging
//     tile (conflict-free 2-byte stores) that one TMA store / reduce-add writes into out[n][m] -- the transposition costs no
//     shuffle and no extra pass; the per-thread row sum is the bias-gradient column sum of the projection (optional).
// Roles: warp 0 = A producer (TMA), warp 1 = MMA issuer (leader CTA), warp 2 = X producer (TMA), warps 4..11 = epilogue.
// Replaces the bmm calls torch autograd issues for F.multi_head_attention_forward's two bmm's (SURVEY Q4;
// mm_multi_transformers.py:62,142-167) in the backward pass.
#include "tc_common.cuh"

namespace jmt {

constexpr int kBwThreads = 384;
constexpr int kBwMaxKb = 5;          // k-blocks of 64 per tile: Lq, S <= 320
constexpr int kBwRing = 8;           // X ring slots (8 KB each)
constexpr int kBwAcc = 4;            // TMEM accumulator slots of 128 columns
constexpr int kBwPiece = 128;
constexpr int kBwEpiWarps = 8;
constexpr int kBwABytes = kBwMaxKb * 16384;
constexpr int kBwStageBytes = kBwEpiWarps * 2 * 4096;
constexpr int kBwCsumMax = 1024;     // heads * dh entries per part

struct BwPart {
  int Nn, Kk, nkb, np, trans, store_mode, last_k16, n_last;   // n_last: MMA N of the last piece (multiple of 16)
  float alpha;
  float* colsum;
};

struct BwParams {
  BwPart part[3];
  int nparts, mtiles, heads, dh, total_tiles, csum_len;
  FastDiv fd_mt, fd_parts, fd_heads;
  uint32_t idesc_base;           // everything but N and the B major
  unsigned long long* prof;
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

__global__ void __launch_bounds__(kBwThreads, 1)
attn_bwd_dqkv_kernel(const __grid_constant__ CUtensorMap ma0, const __grid_constant__ CUtensorMap ma1, const __grid_constant__ CUtensorMap ma2,
                     const __grid_constant__ CUtensorMap mx0, const __grid_constant__ CUtensorMap mx1, const __grid_constant__ CUtensorMap mx2,
                     const __grid_constant__ CUtensorMap md0, const __grid_constant__ CUtensorMap md1, const __grid_constant__ CUtensorMap md2,
                     const BwParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if ((base & 1023u) != 0u) __trap();
  const uint32_t sA = base;                              // kBwMaxKb x [2 chunks x 64 k x 128 B]
  const uint32_t sX = sA + kBwABytes;                    // ring of [64 x 128 B] tiles
  const uint32_t sStage = sX + kBwRing * 8192;           // 8 warps x 2 x [64 n x 64 B]
  const uint32_t sCsum = sStage + kBwStageBytes;         // 3 x csum_len floats
  const uint32_t bars = sCsum + 3 * kBwCsumMax * 4;
  const uint32_t a_full = bars, a_empty = bars + 8 * kBwMaxKb;
  const uint32_t x_full = a_empty + 8 * kBwMaxKb, x_empty = x_full + 8 * kBwRing;
  const uint32_t t_full = x_empty + 8 * kBwRing, t_empty = t_full + 8 * kBwAcc;
  const uint32_t tmem_slot = t_empty + 8 * kBwAcc;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - base));
  float* csum_sh = reinterpret_cast<float*>(smem_raw + (sCsum - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const int first_tile = blockIdx.x >> 1, tile_stride = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kBwMaxKb; ++s) { mbar_init(a_full + 8 * s, 2); mbar_init(a_empty + 8 * s, 1); }
    for (int s = 0; s < kBwRing; ++s) { mbar_init(x_full + 8 * s, 2); mbar_init(x_empty + 8 * s, 1); }
    for (int s = 0; s < kBwAcc; ++s) { mbar_init(t_full + 8 * s, 1); mbar_init(t_empty + 8 * s, kBwEpiWarps * 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * kBwCsumMax; i += kBwThreads) csum_sh[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  cluster_sync_relaxed();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ================================ A producer: the (k x 128 m) tile of this CTA, one k-block per slot ================================
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ma0)) : "memory");
      if (p.nparts > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ma1)) : "memory");
      if (p.nparts > 2) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ma2)) : "memory");
      uint32_t aph = 0;                                    // bit kb: phase of slot kb
      const uint32_t full_leader = mapa_rank(a_full, 0);
      for (int t = first_tile; t < p.total_tiles; t += tile_stride) {
        const BwTile c = bw_decode(p, t);
        const BwPart& P = p.part[c.part];
        const CUtensorMap* ma = c.part == 0 ? &ma0 : (c.part == 1 ? &ma1 : &ma2);
        const int mchunk = (c.mtile * 256 + crank * 128) >> 6;
        for (int kb = 0; kb < P.nkb; ++kb) {
          mbar_wait(a_empty + 8 * kb, ((aph >> kb) & 1u) ^ 1u);
          aph ^= 1u << kb;
          mbar_expect_tx_cluster(full_leader + 8 * kb, 16384);
          tma_load_5d_2sm(sA + kb * 16384, ma, full_leader + 8 * kb, 0, kb * 64, mchunk, c.head, c.b);
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ================================ X producer: this CTA's half of every piece, k-block by k-block ================================
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mx0)) : "memory");
      if (p.nparts > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mx1)) : "memory");
      if (p.nparts > 2) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mx2)) : "memory");
      int slot = 0; uint32_t phase = 0;
      const uint32_t full_leader = mapa_rank(x_full, 0);
      for (int t = first_tile; t < p.total_tiles; t += tile_stride) {
        const BwTile c = bw_decode(p, t);
        const BwPart& P = p.part[c.part];
        const CUtensorMap* mx = c.part == 0 ? &mx0 : (c.part == 1 ? &mx1 : &mx2);
        for (int pc = 0; pc < P.np; ++pc) {
          const int npc = pc == P.np - 1 ? P.n_last : kBwPiece;       // MMA N of this piece; this CTA supplies columns [crank * npc / 2, ...)
          const int n0 = pc * kBwPiece + crank * (npc >> 1);
          for (int kb = 0; kb < P.nkb; ++kb) {
            mbar_wait(x_empty + 8 * slot, phase ^ 1u);
            const uint32_t fb = full_leader + 8 * slot;
            mbar_expect_tx_cluster(fb, 8192);
            if (P.trans) tma_load_4d_2sm(sX + slot * 8192, mx, fb, n0, kb * 64, c.head, c.b);     // [64 k rows][64 n]: MN-major
            else tma_load_4d_2sm(sX + slot * 8192, mx, fb, kb * 64, n0, c.head, c.b);              // [64 n rows][64 k]: K-major
            if (++slot == kBwRing) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
      // ================================ MMA issuer (leader CTA; whole warp, one elected lane issues) ================================
      const uint32_t el = elect_one();
      int slot = 0; uint32_t phase = 0, aph = 0;
      uint32_t acc_it = 0;
      long long w_a = 0, w_x = 0, w_t = 0; const long long t0 = p.prof ? clock64() : 0;
      for (int t = first_tile; t < p.total_tiles; t += tile_stride) {
        const BwTile c = bw_decode(p, t);
        const BwPart& P = p.part[c.part];
        const uint32_t x_lbo = P.trans ? 8192u : 16u;
        const uint32_t x_kstep = P.trans ? 128u : 2u;
        const int nkb = P.nkb, np = P.np, last_k16 = P.last_k16, n_last = P.n_last;
        const uint32_t idesc_t = p.idesc_base | ((uint32_t)P.trans << 16);
        for (int pc = 0; pc < np; ++pc, ++acc_it) {
          const uint32_t acc = acc_it & (kBwAcc - 1), acc_par = (acc_it >> 2) & 1u;
          const int npc = pc == np - 1 ? n_last : kBwPiece;
          const uint32_t idesc = idesc_t | ((uint32_t)(npc >> 3) << 17);
          { const long long tw = p.prof ? clock64() : 0; mbar_wait(t_empty + 8 * acc, acc_par ^ 1u); if (p.prof) w_t += clock64() - tw; }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * kBwPiece;
          const bool last_piece = pc == np - 1;
          for (int kb = 0; kb < nkb; ++kb) {
            if (pc == 0) {
              const long long tw = p.prof ? clock64() : 0;
              mbar_wait(a_full + 8 * kb, (aph >> kb) & 1u);
              if (p.prof) w_a += clock64() - tw;
              aph ^= 1u << kb;
            }
            { const long long tw = p.prof ? clock64() : 0; mbar_wait(x_full + 8 * slot, phase); if (p.prof) w_x += clock64() - tw; }
            tc_fence_after();
            const uint64_t a_desc = make_smem_desc(sA + kb * 16384, 8192, 1024);
            const uint64_t b_desc = make_smem_desc(sX + slot * 8192, x_lbo, 1024);
            const int nk = kb == nkb - 1 ? last_k16 : 4;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k < nk) tc_mma_elect<2>(el, d_tmem, a_desc + (uint64_t)(k * 128), b_desc + (uint64_t)(k * x_kstep), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit_elect<2>(el, x_empty + 8 * slot);
            if (last_piece) tc_commit_elect<2>(el, a_empty + 8 * kb);     // last piece: this k-block of A is free for the next tile
            if (++slot == kBwRing) { slot = 0; phase ^= 1u; }
          }
          tc_commit_elect<2>(el, t_full + 8 * acc);
        }
      }
      if (p.prof && lane == 0) { unsigned long long* o = p.prof + blockIdx.x * 16; o[0] = w_a; o[1] = w_x; o[2] = w_t; o[3] = clock64() - t0; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue (warps 4..11) ================================
    const int q = warp & 3;                          // TMEM lane quarter
    const int ew = warp - 4;
    const int half = ew >> 2;                        // which 64 columns of a piece
    const uint32_t stg0 = sStage + ew * 8192;
    const uint32_t tempty_leader = mapa_rank(t_empty, 0);
    uint32_t acc_it = 0;
    int sbuf = 0;
    long long w_f = 0; const long long t0 = p.prof ? clock64() : 0;
    for (int t = first_tile; t < p.total_tiles; t += tile_stride) {
      const BwTile c = bw_decode(p, t);
      const BwPart& P = p.part[c.part];
      const CUtensorMap* md = c.part == 0 ? &md0 : (c.part == 1 ? &md1 : &md2);
      const int m_loc = c.mtile * 256 + crank * 128 + q * 32;     // first head-dimension index of this warp
      float rsum = 0.f;
      for (int pc = 0; pc < P.np; ++pc, ++acc_it) {
        const uint32_t acc = acc_it & (kBwAcc - 1), acc_par = (acc_it >> 2) & 1u;
        const int n_base = pc * kBwPiece + half * 64;
        const bool active = n_base < P.Nn;             // warp-uniform
        { const long long tw = p.prof ? clock64() : 0; mbar_wait(t_full + 8 * acc, acc_par); if (p.prof) w_f += clock64() - tw; }
        tc_fence_after();
        uint32_t r0[32], r1[32];
        if (active) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBwPiece + half * 64;
          tc_ld32_issue(taddr, r0);
          tc_ld32_issue(taddr + 32, r1);
          tc_wait_ld();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + 8 * acc);      // accumulator slot handed back (both CTAs' warps arrive)
        if (!active) continue;
        const uint32_t stg = stg0 + sbuf * 4096;
        if (lane == 0) bulk_wait_read1();              // the store issued two pieces ago has finished reading this buffer
        __syncwarp();
        const uint32_t col = stg + lane * 2;
        const int nv = P.Nn - n_base;                  // valid columns of this warp's 64
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = P.alpha * __uint_as_float(r0[j]);
          if (j < nv) rsum += v;
          st_shared_u16(col + j * 64, __bfloat16_as_ushort(__float2bfloat16_rn(v)));
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = P.alpha * __uint_as_float(r1[j]);
          if (32 + j < nv) rsum += v;
          st_shared_u16(col + (32 + j) * 64, __bfloat16_as_ushort(__float2bfloat16_rn(v)));
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (P.store_mode == JMT_STORE) tma_store_4d(md, stg, m_loc, n_base, c.head, c.b);
          else tma_reduce_add_4d(md, stg, m_loc, n_base, c.head, c.b);
          bulk_commit();
        }
        sbuf ^= 1;
      }
      if (P.colsum != nullptr) atomicAdd(csum_sh + c.part * kBwCsumMax + c.head * p.dh + m_loc + lane, rsum);
    }
    if (lane == 0) bulk_wait0();
    // flush the bias-gradient column sums: one global atomic per entry and CTA
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kBwEpiWarps) : "memory");
    for (int pa = 0; pa < p.nparts; ++pa) {
      float* dst = p.part[pa].colsum;
      if (dst == nullptr) continue;
      for (int i = threadIdx.x - 128; i < p.csum_len; i += 32 * kBwEpiWarps) {
        const float v = csum_sh[pa * kBwCsumMax + i];
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

// transposed-output tensor map: out(n, m) of (head, batch) at ptr + b*bs + h*hs + n*ld + m; box {32 m, 64 n}, no swizzle
static int make_map_dt(CUtensorMap* map, const void* ptr, int64_t inner, int64_t rows, int64_t ld, int64_t nb0, int64_t bs0,
                       int64_t nb1, int64_t bs1, const char* who) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("%s: cuTensorMapEncodeTiled unavailable (no CUDA driver?)", who); return JMT_ERR_CUDA; }
  JMT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0 && (nb0 == 1 || bs0 % 8 == 0) && (nb1 == 1 || bs1 % 8 == 0),
              "%s: output geometry must be 16-byte aligned", who);
  const cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nb0, (cuuint64_t)nb1};
  const cuuint64_t row_bytes = (cuuint64_t)ld * 2;
  const cuuint64_t strides[3] = {row_bytes, nb0 > 1 ? (cuuint64_t)bs0 * 2 : row_bytes, nb1 > 1 ? (cuuint64_t)bs1 * 2 : row_bytes};
  const cuuint32_t box[4] = {32, 64, 1, 1};
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
  static const int narrow_env = []() { const char* e = getenv("JMT_DQKV_NARROW"); return e ? atoi(e) : 2; }();
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
    P.np = (P.Nn + kBwPiece - 1) / kBwPiece;
    P.last_k16 = (P.Kk - 64 * (P.nkb - 1) + 15) / 16;
    P.n_last = kBwPiece;
    {
      const int rem = (P.Nn - kBwPiece * (P.np - 1) + 15) / 16 * 16;      // columns the last piece needs (multiple of 16)
      if (narrow_env >= 2 || (narrow_env == 1 && !P.trans)) P.n_last = rem;
    }
    P.store_mode = s.store_mode;
    P.alpha = s.alpha;
    P.colsum = s.colsum;
    int rc = make_map_mn5(&ma[np_], s.a, g->dh, P.Kk, s.a_ld, g->heads, s.a_hs, g->NB, s.a_bs, 64, 2, "jmt_attn_bwd_dqkv_bf16(A)");
    if (rc != JMT_OK) return rc;
    rc = make_map(&mx[np_], s.x, g->S, g->Lq, g->x_ld, g->heads, (int64_t)g->Lq * g->x_ld, g->NB, (int64_t)g->heads * g->Lq * g->x_ld, 64,
                  "jmt_attn_bwd_dqkv_bf16(X)");
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

  const int smem = kBwABytes + kBwRing * 8192 + kBwStageBytes + 3 * kBwCsumMax * 4 + 512;
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
