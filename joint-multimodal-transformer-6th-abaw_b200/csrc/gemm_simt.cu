// jmt_gemm_f32: fp32 FFMA tiled GEMM implementing the full jmt_gemm_desc semantics (batches, taps,
// row shifts with zero fill, MN-/K-major operands, reduce_batch, split-K).  This is the
// high-precision parity mode of the engine (fp32 operands, fp32 accumulate) and the on-device
// cross-check for the tcgen05 kernel; it is not the throughput path.
#include "common.cuh"

namespace jmt {

constexpr int kTM = 64, kTN = 64, kTK = 16;

struct SimtArgs {
  jmt_gemm_desc g;
  int kblocks;       // ceil(K / kTK)
  int iters_total;   // ntaps * (reduce ? nb : 1) * kblocks
  int nb;            // nb0 * nb1
};

__device__ __forceinline__ float load_a(const SimtArgs& s, const float* a, int m, int k, int shift) {
  const jmt_gemm_desc& g = s.g;
  if (g.a_major == JMT_MAJOR_K) {
    const int r = m + shift;
    return (m < g.M && k < g.K && r >= 0 && r < g.a_rows) ? a[(int64_t)r * g.a_ld + k] : 0.f;
  }
  const int r = k + shift;
  return (m < g.M && k < g.K && r >= 0 && r < g.a_rows) ? a[(int64_t)r * g.a_ld + m] : 0.f;
}
__device__ __forceinline__ float load_b(const SimtArgs& s, const float* b, int n, int k, int tap, int shift) {
  const jmt_gemm_desc& g = s.g;
  if (g.b_major == JMT_MAJOR_K)
    return (n < g.N && n < g.b_rows && k < g.K) ? b[(int64_t)n * g.b_ld + (int64_t)tap * g.K + k] : 0.f;
  const int r = k + shift;
  return (n < g.N && k < g.K && r >= 0 && r < g.b_rows) ? b[(int64_t)r * g.b_ld + n] : 0.f;
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtArgs s) {
  const jmt_gemm_desc& g = s.g;
  __shared__ float As[kTK][kTM + 4];
  __shared__ float Bs[kTK][kTN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  int z = blockIdx.z;
  const int split = z % g.split_k; z /= g.split_k;
  const int batch = g.reduce_batch ? 0 : z;
  const int per = (s.iters_total + g.split_k - 1) / g.split_k;
  const int it0 = split * per;
  const int it1 = min(s.iters_total, it0 + per);
  const int rb_n = g.reduce_batch ? s.nb : 1;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int it = it0; it < it1; ++it) {
    const int kb = it % s.kblocks;
    const int rb = (it / s.kblocks) % rb_n;
    const int tap = it / (s.kblocks * rb_n);
    const int bidx = g.reduce_batch ? rb : batch;
    const int b0 = bidx % g.nb0, b1 = bidx / g.nb0;
    const float* a = (const float*)g.a + (int64_t)b0 * g.a_bs0 + (int64_t)b1 * g.a_bs1;
    const float* b = (const float*)g.b + (int64_t)b0 * g.b_bs0 + (int64_t)b1 * g.b_bs1;
    const int ash = g.a_shift0 + tap * g.a_shift_step;
    const int bsh = g.b_shift0 + tap * g.b_shift_step;
    // cooperative tile load: 256 threads, 16x64 elements each for A and B
    for (int e = threadIdx.x; e < kTK * kTM; e += 256) {
      int kk, mm;
      if (g.a_major == JMT_MAJOR_K) { kk = e % kTK; mm = e / kTK; } else { mm = e % kTM; kk = e / kTM; }
      As[kk][mm] = load_a(s, a, m0 + mm, kb * kTK + kk, ash);
    }
    for (int e = threadIdx.x; e < kTK * kTN; e += 256) {
      int kk, nn;
      if (g.b_major == JMT_MAJOR_K) { kk = e % kTK; nn = e / kTK; } else { nn = e % kTN; kk = e / kTN; }
      Bs[kk][nn] = load_b(s, b, n0 + nn, kb * kTK + kk, tap, bsh);
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  const int b0 = batch % g.nb0, b1 = batch / g.nb0;
  const int64_t doff = (int64_t)b0 * g.d_bs0 + (int64_t)b1 * g.d_bs1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias && split == 0) v += g.bias[n];
      v = apply_act(v, g.act, g.slope);
      if (g.colmask) {
        const int64_t srow = g.colmask_row_period > 0 ? m / g.colmask_row_period : batch;
        v = g.colmask[srow * g.N + n] ? v * g.colmask_scale : 0.f;
      }
      if (g.zero_row_period > 0 && (m % g.zero_row_period) < g.zero_row_count) v = 0.f;
      const int64_t idx = doff + (int64_t)m * g.d_ld + n;
      if (g.d_dtype == JMT_F32) {
        float* d = (float*)g.d;
        if (g.store_mode == JMT_STORE) d[idx] = v;
        else if (g.store_mode == JMT_ACCUMULATE) d[idx] += v;
        else atomicAdd(d + idx, v);
      } else {
        __nv_bfloat16* d = (__nv_bfloat16*)g.d;
        if (g.store_mode == JMT_STORE) d[idx] = __float2bfloat16_rn(v);
        else d[idx] = __float2bfloat16_rn(__bfloat162float(d[idx]) + v);
      }
    }
  }
}

}  // namespace jmt

using namespace jmt;

int jmt_validate_gemm_desc(const jmt_gemm_desc* g, const char* who) {
  JMT_REQUIRE(g && g->a && g->b && g->d, "%s: null pointer", who);
  JMT_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "%s: M,N,K must be positive (%d,%d,%d)", who, g->M, g->N, g->K);
  JMT_REQUIRE(g->nb0 >= 1 && g->nb1 >= 1 && g->ntaps >= 1 && g->split_k >= 1, "%s: nb0,nb1,ntaps,split_k must be >= 1", who);
  JMT_REQUIRE(g->a_major == JMT_MAJOR_K || g->a_major == JMT_MAJOR_MN, "%s: bad a_major", who);
  JMT_REQUIRE(g->b_major == JMT_MAJOR_K || g->b_major == JMT_MAJOR_MN, "%s: bad b_major", who);
  JMT_REQUIRE(g->d_dtype == JMT_F32 || g->d_dtype == JMT_BF16, "%s: bad d_dtype", who);
  JMT_REQUIRE(g->store_mode >= JMT_STORE && g->store_mode <= JMT_ATOMIC_ADD, "%s: bad store_mode", who);
  JMT_REQUIRE(!(g->store_mode == JMT_ATOMIC_ADD && g->d_dtype != JMT_F32), "%s: atomic add needs fp32 D", who);
  JMT_REQUIRE(!(g->split_k > 1 && (g->store_mode != JMT_ATOMIC_ADD || g->act != JMT_ACT_NONE)),
              "%s: split_k > 1 needs JMT_ATOMIC_ADD and no activation", who);
  JMT_REQUIRE(g->act >= JMT_ACT_NONE && g->act <= JMT_ACT_LEAKY_RELU, "%s: bad act", who);
  JMT_REQUIRE(!(g->colmask && (g->reduce_batch || g->split_k > 1)), "%s: colmask needs reduce_batch == 0 and split_k == 1", who);
  JMT_REQUIRE(g->colmask_row_period >= 0 && g->zero_row_period >= 0 && g->zero_row_count >= 0, "%s: negative row period", who);
  JMT_REQUIRE(!(g->colmask_row_period > 0 && g->nb0 * g->nb1 != 1), "%s: colmask_row_period needs nb0 = nb1 = 1", who);
  // the tensor-core epilogue stages the keep-flags of at most TWO samples per 128-row tile
  JMT_REQUIRE(!(g->colmask && g->colmask_row_period > 0 && g->colmask_row_period < 127),
              "%s: colmask_row_period must be >= 127 (a 128-row tile may span at most two samples); apply the mask with "
              "jmt_apply_mask instead", who);
  JMT_REQUIRE(!(g->zero_row_period > 0 && g->reduce_batch), "%s: zero_row_period cannot be combined with reduce_batch", who);
  JMT_REQUIRE(!((g->epi_aux || g->d_colsum) && (g->d_dtype != JMT_BF16 || g->act != JMT_ACT_NONE || g->split_k != 1 || g->reduce_batch)),
              "%s: epi_aux / d_colsum need bf16 D, no activation, split_k == 1 and reduce_batch == 0", who);
  return JMT_OK;
}

extern "C" int jmt_gemm_f32(const jmt_gemm_desc* g, void* stream) {
  int rc = jmt_validate_gemm_desc(g, "jmt_gemm_f32");
  if (rc != JMT_OK) return rc;
  JMT_REQUIRE(!g->epi_aux && !g->d_colsum, "jmt_gemm_f32: epi_aux / d_colsum are extensions of jmt_gemm_bf16 only");
  SimtArgs s;
  s.g = *g;
  s.kblocks = (g->K + kTK - 1) / kTK;
  s.nb = g->nb0 * g->nb1;
  s.iters_total = g->ntaps * (g->reduce_batch ? s.nb : 1) * s.kblocks;
  const int zb = (g->reduce_batch ? 1 : s.nb) * g->split_k;
  JMT_REQUIRE(zb <= 65535, "jmt_gemm_f32: too many batches*splits (%d)", zb);
  dim3 grid((g->N + kTN - 1) / kTN, (g->M + kTM - 1) / kTM, zb);
  gemm_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(s);
  return check_launch("gemm_simt_kernel");
}
