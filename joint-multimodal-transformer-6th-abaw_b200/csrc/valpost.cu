// Validation post-processing on the device: the step right after the hot path, which the
// reference runs as a Python quadruple loop over .cpu() copies of every batch (val.py:313-382):
//   scatter each window's per-frame predictions / labels into per-video arrays by (video, frame id - 1),
//   skipping frames whose valence OR arousal label is the -5 sentinel and frame ids beyond the video length
//   (later windows overwrite earlier ones; untouched frames stay (0, 0));
//   per video: clip to [-1, 1], scipy.ndimage.uniform_filter1d(size = 20 valence / 50 arousal, mode='constant');
//   CCC (EvaluationMetrics/cccmetric.py:4-21) over the concatenation of all videos, untouched frames included.
// Here: two tiny kernels per batch (deterministic "last writer wins" through 64-bit stamps) and one finalising
// kernel (clip + box filter + the twelve fp64 CCC sums); nothing leaves the GPU but two floats.
#include "common.cuh"

namespace jmt {

constexpr int kVpThreads = 256;

struct VpBatch {
  const float* v; const float* a; const float* lab_v; const float* lab_a;   // (n) each, element e = b*T + t
  const int32_t* frame_id;   // (n) 1-based frame index inside its video
  const int32_t* video;      // (n) video index
  const int64_t* offsets;    // (videos + 1) prefix sums of the video lengths
  int64_t n;
  float ignore;
  unsigned long long seq;    // batch sequence number: later batches win
};

__device__ __forceinline__ int64_t vp_target(const VpBatch& b, int64_t e) {
  if (b.lab_v[e] == b.ignore || b.lab_a[e] == b.ignore) return -1;          // val.py:335-339 / 348-352
  const int vid = b.video[e];
  const int64_t lo = b.offsets[vid], len = b.offsets[vid + 1] - lo;
  const int64_t f = b.frame_id[e];
  if (f < 1 || f > len) return -1;                                          // val.py:346
  return lo + f - 1;
}

__global__ void __launch_bounds__(kVpThreads)
vp_claim_kernel(VpBatch b, unsigned long long* __restrict__ stamp) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < b.n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = vp_target(b, e);
    if (t >= 0) atomicMax(stamp + t, (b.seq << 32) | (unsigned long long)(e + 1));
  }
}

__global__ void __launch_bounds__(kVpThreads)
vp_write_kernel(VpBatch b, const unsigned long long* __restrict__ stamp, float* __restrict__ pv, float* __restrict__ pa,
                float* __restrict__ lv, float* __restrict__ la) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < b.n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = vp_target(b, e);
    if (t >= 0 && stamp[t] == ((b.seq << 32) | (unsigned long long)(e + 1))) {   // the last element in loop order wins
      pv[t] = b.v[e]; pa[t] = b.a[e]; lv[t] = b.lab_v[e]; la[t] = b.lab_a[e];
    }
  }
}

__device__ __forceinline__ float clip1(float x) { return fminf(fmaxf(x, -1.f), 1.f); }

// out[i] = (1/size) * sum_{j = i - size/2}^{i - size/2 + size - 1} clip(x[j]) over the frames of the same video, zeros outside
// (scipy.ndimage.uniform_filter1d, origin 0, mode='constant', cval 0); sums (2, 6) fp64 += CCC sums vs the labels.
__global__ void __launch_bounds__(kVpThreads)
vp_finalize_kernel(const float* __restrict__ pv, const float* __restrict__ pa, const float* __restrict__ lv,
                   const float* __restrict__ la, const int64_t* __restrict__ offsets, int videos, int size_v, int size_a,
                   float* __restrict__ sm_v, float* __restrict__ sm_a, double* __restrict__ sums) {
  const int64_t total = offsets[videos];
  double acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int lo_v = 0, hi_v = videos;                      // video of frame i: offsets[lo_v] <= i < offsets[lo_v + 1]
    while (hi_v - lo_v > 1) { const int mid = (lo_v + hi_v) >> 1; if (offsets[mid] <= i) lo_v = mid; else hi_v = mid; }
    const int64_t lo = offsets[lo_v], hi = offsets[lo_v + 1];
    double sv = 0.0, sa = 0.0;
    for (int64_t j = max(lo, i - size_v / 2), je = min(hi, i - size_v / 2 + size_v); j < je; ++j) sv += (double)clip1(pv[j]);
    for (int64_t j = max(lo, i - size_a / 2), je = min(hi, i - size_a / 2 + size_a); j < je; ++j) sa += (double)clip1(pa[j]);
    const double xv = sv / size_v, xa = sa / size_a;
    if (sm_v) sm_v[i] = (float)xv;
    if (sm_a) sm_a[i] = (float)xa;
    const double yv = (double)lv[i], ya = (double)la[i];
    acc[0] += 1.0; acc[1] += xv; acc[2] += yv; acc[3] = fma(xv, yv, acc[3]); acc[4] = fma(xv, xv, acc[4]); acc[5] = fma(yv, yv, acc[5]);
    acc[6] += 1.0; acc[7] += xa; acc[8] += ya; acc[9] = fma(xa, ya, acc[9]); acc[10] = fma(xa, xa, acc[10]); acc[11] = fma(ya, ya, acc[11]);
  }
  __shared__ double sh[12][kVpThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) sh[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      double v = lane < kVpThreads / 32 ? sh[k][lane] : 0.0;
      v = warp_sum(v);
      if (lane == 0) atomicAdd(&sums[k], v);
    }
  }
}

}  // namespace jmt

using namespace jmt;

extern "C" int jmt_valpost_scatter(const float* v, const float* a, const float* lab_v, const float* lab_a, const int32_t* frame_id,
                                   const int32_t* video, int64_t n, const int64_t* offsets, float ignore, uint64_t seq,
                                   uint64_t* stamp, float* pred_v, float* pred_a, float* label_v, float* label_a, void* stream) {
  JMT_REQUIRE(v && a && lab_v && lab_a && frame_id && video && offsets && stamp && pred_v && pred_a && label_v && label_a && n >= 0,
              "jmt_valpost_scatter: bad arguments");
  JMT_REQUIRE(seq >= 1 && seq < (1ull << 31) && n < (1ll << 31), "jmt_valpost_scatter: seq must be in [1, 2^31), n < 2^31");
  if (n == 0) return JMT_OK;
  VpBatch b{v, a, lab_v, lab_a, frame_id, video, offsets, n, ignore, (unsigned long long)seq};
  const int g = grid_for(n, kVpThreads);
  vp_claim_kernel<<<g, kVpThreads, 0, (cudaStream_t)stream>>>(b, (unsigned long long*)stamp);
  int rc = check_launch("vp_claim_kernel");
  if (rc != JMT_OK) return rc;
  vp_write_kernel<<<g, kVpThreads, 0, (cudaStream_t)stream>>>(b, (const unsigned long long*)stamp, pred_v, pred_a, label_v, label_a);
  return check_launch("vp_write_kernel");
}

extern "C" int jmt_valpost_finalize(const float* pred_v, const float* pred_a, const float* label_v, const float* label_a,
                                    const int64_t* offsets, int videos, int64_t total_frames, int size_v, int size_a,
                                    float* smooth_v, float* smooth_a, double* sums, void* stream) {
  JMT_REQUIRE(pred_v && pred_a && label_v && label_a && offsets && sums && videos >= 1 && size_v >= 1 && size_a >= 1 && total_frames >= 0,
              "jmt_valpost_finalize: bad arguments");
  if (total_frames == 0) return JMT_OK;
  vp_finalize_kernel<<<grid_for(total_frames, kVpThreads, kNumSMs * 4), kVpThreads, 0, (cudaStream_t)stream>>>(
      pred_v, pred_a, label_v, label_a, offsets, videos, size_v, size_a, smooth_v, smooth_a, sums);
  return check_launch("vp_finalize_kernel");
}
