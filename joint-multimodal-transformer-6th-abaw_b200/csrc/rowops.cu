// Memory-bound fused row kernels: L2-normalise, residual+LayerNorm (fwd/bwd), softmax (fwd/bwd).
// One warp per row, 16-byte accesses, warp-shuffle reductions, fp32 statistics.
// Replaces F.normalize (two_transformers.py:118-119), x+attn -> nn.LayerNorm
// (mm_multi_transformers.py:62-69) and torch MHA's softmax.
#include "common.cuh"

namespace jmt {

constexpr int kRowThreads = 256;            // 8 warps = 8 rows per block iteration
constexpr int kWarpsPerBlock = kRowThreads / 32;

// ------------------------------------------------------------------ l2norm
template <typename TI, typename TO>
__global__ void __launch_bounds__(kRowThreads)
l2norm_fwd_kernel(const TI* __restrict__ x, int64_t in_ld, TO* __restrict__ out, int64_t rows, int D, float eps,
                  float* __restrict__ inv_norm, int seq_len, int seq_rows, int row0) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    // seq_len > 0: compact row r = n * seq_len + t lives at physical row n * seq_rows + row0 + t of x (the TCN's flat padded layout)
    const int64_t pr = seq_len > 0 ? (r / seq_len) * seq_rows + row0 + r % seq_len : r;
    const TI* xr = x + pr * in_ld;
    float ss = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      Vec8<TI> v; v.load(xr + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss = fmaf(v.v[i], v.v[i], ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    if (lane == 0 && inv_norm) inv_norm[r] = inv;
    TO* orow = out + r * (int64_t)D;
    for (int c = lane * 8; c < D; c += 256) {
      Vec8<TI> v; v.load(xr + c);
      Vec8<TO> o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = v.v[i] * inv;
      o.store(orow + c);
    }
  }
}

template <typename T, typename TO>
__global__ void __launch_bounds__(kRowThreads)
l2norm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, const float* __restrict__ inv_norm, float eps,
                  TO* __restrict__ dx, int64_t rows, int D, int seq_len, int seq_rows, int row0) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  // seq_len > 0: dx is the flat padded layout (seq_rows physical rows per sequence, the first row0 of them padding): the loop runs over
  // PHYSICAL rows, padding rows are written as zeros (the TCN backward relies on them), the others take compact row n * seq_len + t
  const int64_t prows = seq_len > 0 ? rows / seq_len * seq_rows : rows;
  for (int64_t pr = warp0; pr < prows; pr += nwarps) {
    int64_t r = pr;
    if (seq_len > 0) {
      const int64_t n = pr / seq_rows;
      const int t = (int)(pr - n * seq_rows) - row0;
      if (t < 0) {
        TO* z = dx + pr * (int64_t)D;
        Vec8<TO> o;
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = 0.f;
        for (int c = lane * 8; c < D; c += 256) o.store(z + c);
        continue;
      }
      r = n * seq_len + t;
    }
    const T* dyr = dy + r * (int64_t)D;
    const T* yr = y + r * (int64_t)D;
    const float inv = inv_norm[r];
    float dot = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      Vec8<T> a, b; a.load(dyr + c); b.load(yr + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) dot = fmaf(a.v[i], b.v[i], dot);
    }
    dot = warp_sum(dot);
    // ||x|| clamped by eps: y = x/eps is linear in x, no projection term
    if (inv >= 1.f / eps) dot = 0.f;
    TO* dxr = dx + pr * (int64_t)D;
    for (int c = lane * 8; c < D; c += 256) {
      Vec8<T> a, b; a.load(dyr + c); b.load(yr + c);
      Vec8<TO> o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = inv * (a.v[i] - b.v[i] * dot);
      o.store(dxr + c);
    }
  }
}

// ------------------------------------------------------------------ add + LayerNorm
template <typename T, int NCH>   // D = NCH * 256
__global__ void __launch_bounds__(kRowThreads)
add_layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, T* __restrict__ y, float* __restrict__ mean,
                         float* __restrict__ rstd, int64_t rows) {
  constexpr int D = NCH * 256;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float g[NCH][8], b[NCH][8];
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    Vec8<float> t; t.load(gamma + j * 256 + lane * 8);
    Vec8<float> u; u.load(beta + j * 256 + lane * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[j][i] = t.v[i]; b[j][i] = u.v[i]; }
  }
  for (int64_t r = warp0; r < rows; r += nwarps) {
    float z[NCH][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      Vec8<T> a; a.load(x + r * D + j * 256 + lane * 8);
      if (res) {
        Vec8<T> c; c.load(res + r * D + j * 256 + lane * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) a.v[i] += c.v[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { z[j][i] = a.v[i]; s += a.v[i]; }
    }
    const float mu = warp_sum(s) * (1.f / D);
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = z[j][i] - mu; var = fmaf(d, d, var); }
    const float rs = rsqrtf(warp_sum(var) * (1.f / D) + eps);
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      Vec8<T> o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = (z[j][i] - mu) * rs * g[j][i] + b[j][i];
      o.store(y + r * D + j * 256 + lane * 8);
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(kRowThreads)
add_layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ res,
                         const float* __restrict__ gamma, const float* __restrict__ mean,
                         const float* __restrict__ rstd, T* __restrict__ dz, int dz_accumulate,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dz_colsum,
                         int64_t rows) {
  constexpr int D = NCH * 256;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  float g[NCH][8], dg[NCH][8], db[NCH][8], dzs[NCH][8];
#pragma unroll
  for (int j = 0; j < NCH; ++j) {
    Vec8<float> t; t.load(gamma + j * 256 + lane * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[j][i] = t.v[i]; dg[j][i] = 0.f; db[j][i] = 0.f; dzs[j][i] = 0.f; }
  }
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float mu = mean[r], rs = rstd[r];
    float xh[NCH][8], gy[NCH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      Vec8<T> a; a.load(x + r * D + j * 256 + lane * 8);
      if (res) {
        Vec8<T> c; c.load(res + r * D + j * 256 + lane * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) a.v[i] += c.v[i];
      }
      Vec8<T> d; d.load(dy + r * D + j * 256 + lane * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float h = (a.v[i] - mu) * rs;
        const float gg = d.v[i] * g[j][i];
        xh[j][i] = h; gy[j][i] = gg;
        s1 += gg; s2 = fmaf(gg, h, s2);
        dg[j][i] = fmaf(d.v[i], h, dg[j][i]);
        db[j][i] += d.v[i];
      }
    }
    s1 = warp_sum(s1) * (1.f / D);
    s2 = warp_sum(s2) * (1.f / D);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      Vec8<T> o;
      if (dz_accumulate) o.load(dz + r * D + j * 256 + lane * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float v = rs * (gy[j][i] - s1 - xh[j][i] * s2);
        dzs[j][i] += v;
        o.v[i] = dz_accumulate ? o.v[i] + v : v;
      }
      o.store(dz + r * D + j * 256 + lane * 8);
    }
  }
  // block-level reduction of dgamma / dbeta partials, then one atomic per column per block
  __shared__ float sh[kWarpsPerBlock][D + 8];
  // (third pass: column sums of dz = the bias gradient of the Linear that produced the residual branch)
  for (int pass = 0; pass < (dz_colsum ? 3 : 2); ++pass) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) sh[warp][j * 256 + lane * 8 + i] = pass == 0 ? dg[j][i] : (pass == 1 ? db[j][i] : dzs[j][i]);
    __syncthreads();
    float* outp = pass == 0 ? dgamma : (pass == 1 ? dbeta : dz_colsum);
    for (int c = threadIdx.x; c < D; c += kRowThreads) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarpsPerBlock; ++w) s += sh[w][c];
      atomicAdd(outp + c, s);
    }
  }
}

// ------------------------------------------------------------------ softmax
template <typename TP>
__global__ void __launch_bounds__(kRowThreads)
softmax_fwd_kernel(const float* __restrict__ s, int64_t s_ld, TP* __restrict__ p, int64_t p_ld, int64_t rows,
                   int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float* sr = s + r * s_ld;
    float m = -INFINITY;
    for (int c = lane; c < cols; c += 32) m = fmaxf(m, sr[c]);
    m = warp_max(m);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) sum += __expf(sr[c] - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    TP* pr = p + r * p_ld;
    for (int c = lane; c < (int)p_ld; c += 32)
      pr[c] = from_f32<TP>(c < cols ? __expf(sr[c] - m) * inv : 0.f);
  }
}

template <typename TP, typename TS>
__global__ void __launch_bounds__(kRowThreads)
softmax_bwd_kernel(const TP* __restrict__ p, int64_t p_ld, const float* __restrict__ dp, int64_t dp_ld,
                   TS* __restrict__ ds, int64_t ds_ld, int64_t rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const TP* pr = p + r * p_ld;
    const float* dpr = dp + r * dp_ld;
    float dot = 0.f;
    for (int c = lane; c < cols; c += 32) dot = fmaf(to_f32(pr[c]), dpr[c], dot);
    dot = warp_sum(dot);
    TS* dsr = ds + r * ds_ld;
    for (int c = lane; c < (int)ds_ld; c += 32)
      dsr[c] = from_f32<TS>(c < cols ? to_f32(pr[c]) * (dpr[c] - dot) : 0.f);
  }
}

}  // namespace jmt

using namespace jmt;

static inline int row_grid(int64_t rows) { return grid_for(rows, kWarpsPerBlock, kNumSMs * 8); }

extern "C" int jmt_l2norm_fwd(const void* x, int in_dtype, int64_t in_ld, void* out, int out_dtype, int64_t rows,
                              int D, float eps, float* inv_norm, void* stream) {
  JMT_REQUIRE(x && out && rows >= 0 && D > 0 && D % 8 == 0 && in_ld % 8 == 0, "jmt_l2norm_fwd: D and in_ld must be multiples of 8");
  if (rows == 0) return JMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  JMT_DISPATCH_DTYPE(in_dtype, TI, JMT_DISPATCH_DTYPE(out_dtype, TO,
      (l2norm_fwd_kernel<TI, TO><<<row_grid(rows), kRowThreads, 0, st>>>((const TI*)x, in_ld, (TO*)out, rows, D, eps, inv_norm, 0, 0, 0))));
  return check_launch("l2norm_fwd_kernel");
}

extern "C" int jmt_l2norm_fwd_seq(const void* x, int in_dtype, int64_t in_ld, void* out, int out_dtype, int64_t nseq, int seq_len,
                                  int seq_rows, int row0, int D, float eps, float* inv_norm, void* stream) {
  JMT_REQUIRE(x && out && nseq >= 0 && seq_len > 0 && row0 >= 0 && seq_rows >= row0 + seq_len && D > 0 && D % 8 == 0 && in_ld % 8 == 0,
              "jmt_l2norm_fwd_seq: bad arguments");
  const int64_t rows = nseq * seq_len;
  if (rows == 0) return JMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  JMT_DISPATCH_DTYPE(in_dtype, TI, JMT_DISPATCH_DTYPE(out_dtype, TO,
      (l2norm_fwd_kernel<TI, TO><<<row_grid(rows), kRowThreads, 0, st>>>((const TI*)x, in_ld, (TO*)out, rows, D, eps, inv_norm, seq_len, seq_rows, row0))));
  return check_launch("l2norm_fwd_kernel");
}

extern "C" int jmt_l2norm_bwd(const void* dy, const void* y, int dtype, const float* inv_norm, float eps, void* dx, int dx_dtype,
                              int64_t rows, int D, void* stream) {
  JMT_REQUIRE(dy && y && inv_norm && dx && D > 0 && D % 8 == 0, "jmt_l2norm_bwd: bad arguments");
  if (rows == 0) return JMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  JMT_DISPATCH_DTYPE(dtype, T, JMT_DISPATCH_DTYPE(dx_dtype, TO,
      (l2norm_bwd_kernel<T, TO><<<row_grid(rows), kRowThreads, 0, st>>>((const T*)dy, (const T*)y, inv_norm, eps, (TO*)dx, rows, D, 0, 0, 0))));
  return check_launch("l2norm_bwd_kernel");
}

extern "C" int jmt_l2norm_bwd_seq(const void* dy, const void* y, int dtype, const float* inv_norm, float eps, void* dx, int dx_dtype,
                                  int64_t nseq, int seq_len, int seq_rows, int row0, int D, void* stream) {
  JMT_REQUIRE(dy && y && inv_norm && dx && nseq >= 0 && seq_len > 0 && row0 >= 0 && seq_rows >= row0 + seq_len && D > 0 && D % 8 == 0,
              "jmt_l2norm_bwd_seq: bad arguments");
  JMT_REQUIRE(seq_rows == row0 + seq_len, "jmt_l2norm_bwd_seq: rows behind a sequence are not supported (they would stay unwritten)");
  const int64_t rows = nseq * seq_len;
  if (rows == 0) return JMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  JMT_DISPATCH_DTYPE(dtype, T, JMT_DISPATCH_DTYPE(dx_dtype, TO,
      (l2norm_bwd_kernel<T, TO><<<row_grid(nseq * seq_rows), kRowThreads, 0, st>>>((const T*)dy, (const T*)y, inv_norm, eps, (TO*)dx, rows, D,
                                                                                   seq_len, seq_rows, row0))));
  return check_launch("l2norm_bwd_kernel");
}

template <typename T>
static int launch_ln_fwd(const void* x, const void* res, const float* gamma, const float* beta, float eps, void* y,
                         float* mean, float* rstd, int64_t rows, int D, cudaStream_t st) {
  const int g = row_grid(rows);
#define JMT_LN_CASE(N) case N: add_layernorm_fwd_kernel<T, N><<<g, kRowThreads, 0, st>>>((const T*)x, (const T*)res, gamma, beta, eps, (T*)y, mean, rstd, rows); break;
  switch (D / 256) { JMT_LN_CASE(1) JMT_LN_CASE(2) JMT_LN_CASE(3) JMT_LN_CASE(4)
    default: set_error("add_layernorm: D=%d unsupported (need 256..1024, multiple of 256)", D); return JMT_ERR_UNSUPPORTED; }
#undef JMT_LN_CASE
  return check_launch("add_layernorm_fwd_kernel");
}

extern "C" int jmt_add_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, float eps,
                                     void* y, float* mean, float* rstd, int64_t rows, int D, int dtype, void* stream) {
  JMT_REQUIRE(x && gamma && beta && y && mean && rstd && D % 256 == 0 && D >= 256, "jmt_add_layernorm_fwd: bad arguments (D=%d)", D);
  if (rows == 0) return JMT_OK;
  JMT_DISPATCH_DTYPE(dtype, T, return launch_ln_fwd<T>(x, res, gamma, beta, eps, y, mean, rstd, rows, D, (cudaStream_t)stream));
  return JMT_OK;
}

template <typename T>
static int launch_ln_bwd(const void* dy, const void* x, const void* res, const float* gamma, const float* mean,
                         const float* rstd, void* dz, int acc, float* dgamma, float* dbeta, float* dz_colsum, int64_t rows, int D,
                         cudaStream_t st) {
  const int g = grid_for(rows, kWarpsPerBlock * 16, kNumSMs * 2);   // each warp sweeps many rows: few atomics
#define JMT_LN_CASE(N) case N: add_layernorm_bwd_kernel<T, N><<<g, kRowThreads, 0, st>>>((const T*)dy, (const T*)x, (const T*)res, gamma, mean, rstd, (T*)dz, acc, dgamma, dbeta, dz_colsum, rows); break;
  switch (D / 256) { JMT_LN_CASE(1) JMT_LN_CASE(2) JMT_LN_CASE(3) JMT_LN_CASE(4)
    default: set_error("add_layernorm_bwd: D=%d unsupported", D); return JMT_ERR_UNSUPPORTED; }
#undef JMT_LN_CASE
  return check_launch("add_layernorm_bwd_kernel");
}

extern "C" int jmt_add_layernorm_bwd(const void* dy, const void* x, const void* res, const float* gamma,
                                     const float* mean, const float* rstd, void* dz, int dz_accumulate, float* dgamma,
                                     float* dbeta, float* dz_colsum, int64_t rows, int D, int dtype, void* stream) {
  JMT_REQUIRE(dy && x && gamma && mean && rstd && dz && dgamma && dbeta && D % 256 == 0 && D >= 256,
              "jmt_add_layernorm_bwd: bad arguments (D=%d)", D);
  if (rows == 0) return JMT_OK;
  JMT_DISPATCH_DTYPE(dtype, T, return launch_ln_bwd<T>(dy, x, res, gamma, mean, rstd, dz, dz_accumulate, dgamma, dbeta, dz_colsum, rows, D, (cudaStream_t)stream));
  return JMT_OK;
}

extern "C" int jmt_softmax_fwd(const float* s, int64_t s_ld, void* p, int p_dtype, int64_t p_ld, int64_t rows, int cols,
                               void* stream) {
  JMT_REQUIRE(s && p && cols > 0 && s_ld >= cols && p_ld >= cols, "jmt_softmax_fwd: bad arguments");
  if (rows == 0) return JMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  JMT_DISPATCH_DTYPE(p_dtype, TP, (softmax_fwd_kernel<TP><<<row_grid(rows), kRowThreads, 0, st>>>(s, s_ld, (TP*)p, p_ld, rows, cols)));
  return check_launch("softmax_fwd_kernel");
}

extern "C" int jmt_softmax_bwd(const void* p, int p_dtype, int64_t p_ld, const float* dp, int64_t dp_ld, void* ds,
                               int ds_dtype, int64_t ds_ld, int64_t rows, int cols, void* stream) {
  JMT_REQUIRE(p && dp && ds && cols > 0 && p_ld >= cols && dp_ld >= cols && ds_ld >= cols, "jmt_softmax_bwd: bad arguments");
  if (rows == 0) return JMT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  JMT_DISPATCH_DTYPE(p_dtype, TP, JMT_DISPATCH_DTYPE(ds_dtype, TS,
      (softmax_bwd_kernel<TP, TS><<<row_grid(rows), kRowThreads, 0, st>>>((const TP*)p, p_ld, dp, dp_ld, (TS*)ds, ds_ld, rows, cols))));
  return check_launch("softmax_bwd_kernel");
}
