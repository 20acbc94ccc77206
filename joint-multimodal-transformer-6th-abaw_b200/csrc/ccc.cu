// CCC statistics: single-pass six-sum reduction + device finaliser + closed-form backward.
// Replaces losses/loss.py:18-32, losses/CCCLoss.py:12-43, EvaluationMetrics/cccmetric.py:4-56
// (SURVEY.md 8a rows L1-L3, Appendix A).  HBM/latency-bound: 8 bytes per element read once.
#include "common.cuh"

namespace jmt {

constexpr int kCccThreads = 256;

__global__ void __launch_bounds__(kCccThreads)
ccc_sums_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, int64_t stride,
                int use_ignore, float ignore, double* __restrict__ sums) {
  const int pair = blockIdx.y;
  const float* xp = x + (int64_t)pair * stride;
  const float* yp = y + (int64_t)pair * stride;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  auto take = [&](float xf, float yf) {
    if (use_ignore && yf == ignore) return;
    double xd = (double)xf, yd = (double)yf;
    acc[0] += 1.0; acc[1] += xd; acc[2] += yd;
    acc[3] = fma(xd, yd, acc[3]); acc[4] = fma(xd, xd, acc[4]); acc[5] = fma(yd, yd, acc[5]);
  };
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(yp)) & 15) == 0;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    const float4* y4 = reinterpret_cast<const float4*>(yp);
    for (int64_t i = tid; i < n4; i += nthreads) {
      float4 a = __ldg(x4 + i), b = __ldg(y4 + i);
      take(a.x, b.x); take(a.y, b.y); take(a.z, b.z); take(a.w, b.w);
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) take(xp[i], yp[i]);

  __shared__ double sh[6][kCccThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    double v = warp_sum(acc[k]);
    if (lane == 0) sh[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      double v = lane < kCccThreads / 32 ? sh[k][lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) atomicAdd(&sums[pair * 6 + k], v);
    }
  }
}

__global__ void ccc_finalize_kernel(const double* __restrict__ sums, int npairs, int kind, double n_all,
                                    double eps, float* __restrict__ value, double* __restrict__ coef) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  const double n = sums[p * 6 + 0], sx = sums[p * 6 + 1], sy = sums[p * 6 + 2];
  const double sxy = sums[p * 6 + 3], sxx = sums[p * 6 + 4], syy = sums[p * 6 + 5];
  double val = 0.0, c0 = 0.0, cx = 0.0, cy = 0.0, valid = 0.0;
  if (n > 1.0 || (kind != JMT_CCC_LOSS_MASKED && n > 0.0)) {
    const double mx = sx / n, my = sy / n, d = mx - my;
    const double cxx = sxx - sx * sx / n, cyy = syy - sy * sy / n, cxy = sxy - sx * sy / n;
    valid = 1.0;
    if (kind == JMT_CCC_METRIC) {
      const double D = cxx + cyy + n * d * d;
      val = 2.0 * cxy / D;
      cy = 2.0 / D;
      cx = -4.0 * cxy / (D * D);
      c0 = -cy * my + cx * (d - mx);
    } else if (kind == JMT_CCC_NUMPY) {
      const double Den = (cxx + cyy) / n + d * d + 1e-8;
      const double Num = 2.0 * cxy / (n - 1.0);
      val = Num / Den;
      cy = 2.0 / ((n - 1.0) * Den);
      cx = -Num / (Den * Den) * 2.0 / n;
      c0 = -cy * my - cx * mx + cx * d;
    } else if (kind == JMT_CCC_LOSS_LIVE) {
      const double a = sqrt(cxx), b = sqrt(cyy), u = a * b;
      const double f = u / (u + eps), fp = eps / ((u + eps) * (u + eps));
      const double Num = 2.0 * cxy * f / (n - 1.0);
      const double Den = (cxx + cyy) / (n - 1.0) + d * d;
      const double C = Num / Den;
      const double ay = 2.0 * f / (n - 1.0);
      const double ax = a > 0.0 ? 2.0 * cxy * fp * b / (a * (n - 1.0)) : 0.0;
      const double bx = 2.0 / (n - 1.0), b0 = 2.0 * d / n;
      const double cxC = ax / Den - Num * bx / (Den * Den);
      const double cyC = ay / Den;
      const double c0C = -cxC * mx - cyC * my - Num * b0 / (Den * Den);
      val = 1.0 - C; cx = -cxC; cy = -cyC; c0 = -c0C;
    } else {  // JMT_CCC_LOSS_MASKED
      const double Den = (cxx + cyy) / (n - 1.0) + d * d + 1e-8;
      const double C = 2.0 * cxy / (Den * n_all);
      const double k2 = 2.0 * cxy / (Den * Den * n_all);
      const double cyC = 2.0 / (Den * n_all);
      const double cxC = -k2 * 2.0 / (n - 1.0);
      const double c0C = -cyC * my - cxC * mx - k2 * 2.0 * d / n;
      val = 1.0 - C; cx = -cxC; cy = -cyC; c0 = -c0C;
    }
  }
  if (value) value[p] = (float)val;
  if (coef) { coef[p * 4 + 0] = c0; coef[p * 4 + 1] = cx; coef[p * 4 + 2] = cy; coef[p * 4 + 3] = valid; }
}

__global__ void __launch_bounds__(256)
ccc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, int64_t stride,
               const double* __restrict__ coef, const float* __restrict__ gout, int gout_per_pair,
               int use_ignore, float ignore, float* __restrict__ dx) {
  const int pair = blockIdx.y;
  const double g = (double)gout[gout_per_pair ? pair : 0];
  const double c0 = coef[pair * 4 + 0] * g, cx = coef[pair * 4 + 1] * g, cy = coef[pair * 4 + 2] * g;
  const float* xp = x + (int64_t)pair * stride;
  const float* yp = y + (int64_t)pair * stride;
  float* dp = dx + (int64_t)pair * stride;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float yf = yp[i];
    double v = c0 + cx * (double)xp[i] + cy * (double)yf;
    if (use_ignore && yf == ignore) v = 0.0;
    dp[i] = (float)v;
  }
}

__global__ void label_mask_kernel(const float* __restrict__ y, int64_t n, float ignore, uint8_t* __restrict__ mask) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask[i] = y[i] != ignore ? 1 : 0;
}

__global__ void pad_right_align_kernel(const float* __restrict__ in, int64_t rows, int in_w,
                                       float* __restrict__ out, int out_w) {
  const int64_t total = rows * out_w;
  const int off = out_w - in_w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / out_w;
    const int c = (int)(i - r * out_w);
    out[i] = c >= off ? in[r * in_w + (c - off)] : 0.f;
  }
}

}  // namespace jmt

using namespace jmt;

extern "C" int jmt_ccc_sums(const float* x, const float* y, int64_t n, int npairs, int64_t stride, int use_ignore,
                            float ignore, double* sums, void* stream) {
  JMT_REQUIRE(x && y && sums && n >= 0 && npairs >= 1, "jmt_ccc_sums: bad arguments");
  if (n == 0) return JMT_OK;
  dim3 grid(grid_for(n, kCccThreads * 8, kNumSMs * 2), npairs);
  ccc_sums_kernel<<<grid, kCccThreads, 0, (cudaStream_t)stream>>>(x, y, n, stride, use_ignore, ignore, sums);
  return check_launch("ccc_sums_kernel");
}

extern "C" int jmt_ccc_finalize(const double* sums, int npairs, int kind, double n_all, double eps, float* value,
                                double* coef, void* stream) {
  JMT_REQUIRE(sums && npairs >= 1 && kind >= 0 && kind <= 3, "jmt_ccc_finalize: bad arguments");
  ccc_finalize_kernel<<<(npairs + 31) / 32, 32, 0, (cudaStream_t)stream>>>(sums, npairs, kind, n_all, eps, value, coef);
  return check_launch("ccc_finalize_kernel");
}

extern "C" int jmt_ccc_bwd(const float* x, const float* y, int64_t n, int npairs, int64_t stride, const double* coef,
                           const float* gout, int gout_per_pair, int use_ignore, float ignore, float* dx,
                           void* stream) {
  JMT_REQUIRE(x && y && coef && gout && dx && npairs >= 1, "jmt_ccc_bwd: bad arguments");
  if (n == 0) return JMT_OK;
  dim3 grid(grid_for(n, 256 * 4, kNumSMs * 4), npairs);
  ccc_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, n, stride, coef, gout, gout_per_pair, use_ignore,
                                                        ignore, dx);
  return check_launch("ccc_bwd_kernel");
}

extern "C" int jmt_label_mask(const float* y, int64_t n, float ignore, uint8_t* mask, void* stream) {
  JMT_REQUIRE(y && mask && n >= 0, "jmt_label_mask: bad arguments");
  if (n == 0) return JMT_OK;
  label_mask_kernel<<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(y, n, ignore, mask);
  return check_launch("label_mask_kernel");
}

extern "C" int jmt_pad_right_align(const float* in, int64_t rows, int in_w, float* out, int out_w, void* stream) {
  JMT_REQUIRE(in && out && in_w >= 0 && out_w >= in_w, "jmt_pad_right_align: bad arguments");
  if (rows * out_w == 0) return JMT_OK;
  pad_right_align_kernel<<<grid_for(rows * out_w, 256 * 4), 256, 0, (cudaStream_t)stream>>>(in, rows, in_w, out, out_w);
  return check_launch("pad_right_align_kernel");
}
