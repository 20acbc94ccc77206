// Regressor tail: the Linear(128, k) that closes each valence/arousal head
// (two_transformers.py:104-114 `nn.Linear(128, 1)`, :146-149 `nn.Linear(128, 2)`), fused with the
// squeeze / output-layout permutation (TRANSFORMER+FC returns (T, B)) and, in backward,
// with the ReLU(+dropout) mask of the hidden layer.  One warp per row; HBM-bound.
#include "common.cuh"

namespace jmt {

constexpr int kMaxGroups = 4;
constexpr int kHid = 128;

struct TailArgs {
  const void* h[kMaxGroups];     // hidden (M, 128) per group, row pitch h_ld (post ReLU, post dropout)
  const float* w[kMaxGroups];    // (128)
  const float* b[kMaxGroups];    // (1)
  float* out[kMaxGroups];        // out[g][b*sb + t*st]
  const float* dout[kMaxGroups];
  void* dh[kMaxGroups];          // d(pre-activation hidden) (M, 128), row pitch h_ld
  float* dw[kMaxGroups];         // (128) accumulated
  float* db[kMaxGroups];         // (1) accumulated
  float scale[kMaxGroups];       // dropout 1/(1-p) (1 when off)
  int accumulate[kMaxGroups];    // dh[g] += (shared hidden between groups)
  int64_t h_ld;
  int64_t M, T, sb, st;          // row m = b*T + t
  int G;
};

template <typename T>
__global__ void __launch_bounds__(256) regressor_tail_fwd_kernel(TailArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  for (int64_t m = w0; m < a.M; m += (int64_t)gridDim.x * 8) {
    const int64_t bb = m / a.T, tt = m - bb * a.T;
    for (int g = 0; g < a.G; ++g) {
      const T* hr = (const T*)a.h[g] + m * a.h_ld;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < kHid / 32; ++j) s = fmaf(to_f32(hr[lane + 32 * j]), __ldg(a.w[g] + lane + 32 * j), s);
      s = warp_sum(s);
      if (lane == 0) a.out[g][bb * a.sb + tt * a.st] = s + a.b[g][0];
    }
  }
}

// 4 consecutive hidden units of one row (lane l owns columns 4l .. 4l+3): one 8-byte (bf16) / 16-byte (fp32) access
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 r = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 r;
  *reinterpret_cast<__nv_bfloat162*>(&r.x) = __floats2bfloat162_rn(v[0], v[1]);
  *reinterpret_cast<__nv_bfloat162*>(&r.y) = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = r;
}

// kVec: rows are 4-element aligned (pointers and pitch), so lane l owns columns 4l..4l+3 and moves them with one vector
// access; otherwise lane l owns columns l, l+32, l+64, l+96 (scalar accesses).  dw[g][j] is the column col(j) either way.
template <typename T, bool kVec>
__global__ void __launch_bounds__(256) regressor_tail_bwd_kernel(TailArgs a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + warp;
  float dw[kMaxGroups][kHid / 32];
  float db[kMaxGroups];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) { db[g] = 0.f;
#pragma unroll
    for (int j = 0; j < kHid / 32; ++j) dw[g][j] = 0.f; }
  for (int64_t m = w0; m < a.M; m += (int64_t)gridDim.x * 8) {
    const int64_t bb = m / a.T, tt = m - bb * a.T;
#pragma unroll
    for (int g = 0; g < kMaxGroups; ++g) {
      if (g >= a.G) break;
      const float go = a.dout[g][bb * a.sb + tt * a.st];
      const T* hr = (const T*)a.h[g] + m * a.h_ld;
      T* dr = (T*)a.dh[g] + m * a.h_ld;
      if (lane == 0) db[g] += go;
      if constexpr (kVec) {
        float hv[4], v[4], old[4];
        load4(hr + 4 * lane, hv);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(a.w[g]) + lane);
        const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
        if (a.accumulate[g]) load4(dr + 4 * lane, old);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dw[g][j] = fmaf(go, hv[j], dw[g][j]);
          v[j] = hv[j] > 0.f ? go * wr[j] * a.scale[g] : 0.f;
          if (a.accumulate[g]) v[j] += old[j];
        }
        store4(dr + 4 * lane, v);
      } else {
#pragma unroll
        for (int j = 0; j < kHid / 32; ++j) {
          const int c = lane + 32 * j;
          const float hv = to_f32(hr[c]);
          dw[g][j] = fmaf(go, hv, dw[g][j]);
          float v = hv > 0.f ? go * __ldg(a.w[g] + c) * a.scale[g] : 0.f;
          if (a.accumulate[g]) v += to_f32(dr[c]);
          dr[c] = from_f32<T>(v);
        }
      }
    }
  }
  __shared__ float sh[8][kHid + 1];
#pragma unroll
  for (int g = 0; g < kMaxGroups; ++g) {
    if (g >= a.G) break;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kHid / 32; ++j) sh[warp][kVec ? 4 * lane + j : lane + 32 * j] = dw[g][j];
    if (lane == 0) sh[warp][kHid] = db[g];
    __syncthreads();
    if (threadIdx.x <= kHid) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += sh[w][threadIdx.x];
      if (threadIdx.x < kHid) atomicAdd(a.dw[g] + threadIdx.x, s);
      else atomicAdd(a.db[g], s);
    }
  }
}

}  // namespace jmt

using namespace jmt;

static int fill_tail(TailArgs& a, int G, const void* const* h, int64_t h_ld, const float* const* w, const float* const* b,
                     int64_t M, int64_t T, int64_t sb, int64_t st) {
  JMT_REQUIRE(G >= 1 && G <= kMaxGroups && h && w && b && M >= 0 && T >= 1, "regressor tail: bad arguments");
  memset(&a, 0, sizeof(a));
  a.G = G; a.h_ld = h_ld; a.M = M; a.T = T; a.sb = sb; a.st = st;
  for (int g = 0; g < G; ++g) { a.h[g] = h[g]; a.w[g] = w[g]; a.b[g] = b[g]; }
  return JMT_OK;
}

extern "C" int jmt_regressor_tail_fwd(int G, const void* const* h, int64_t h_ld, int dtype, const float* const* w,
                                      const float* const* b, float* const* out, int64_t M, int64_t T, int64_t sb,
                                      int64_t st, void* stream) {
  TailArgs a;
  int rc = fill_tail(a, G, h, h_ld, w, b, M, T, sb, st);
  if (rc != JMT_OK) return rc;
  JMT_REQUIRE(out, "jmt_regressor_tail_fwd: null out");
  for (int g = 0; g < G; ++g) a.out[g] = out[g];
  if (M == 0) return JMT_OK;
  const int grid = grid_for(M, 8, kNumSMs * 8);
  JMT_DISPATCH_DTYPE(dtype, T_, (regressor_tail_fwd_kernel<T_><<<grid, 256, 0, (cudaStream_t)stream>>>(a)));
  return check_launch("regressor_tail_fwd_kernel");
}

extern "C" int jmt_regressor_tail_bwd(int G, const void* const* h, int64_t h_ld, int dtype, const float* const* w,
                                      const float* const* dout, void* const* dh, const int* accumulate,
                                      const float* scale, float* const* dw, float* const* db, int64_t M, int64_t T,
                                      int64_t sb, int64_t st, void* stream) {
  TailArgs a;
  int rc = fill_tail(a, G, h, h_ld, w, w /*unused*/, M, T, sb, st);
  if (rc != JMT_OK) return rc;
  JMT_REQUIRE(dout && dh && dw && db && accumulate && scale, "jmt_regressor_tail_bwd: null argument");
  for (int g = 0; g < G; ++g) {
    a.dout[g] = dout[g]; a.dh[g] = dh[g]; a.dw[g] = dw[g]; a.db[g] = db[g];
    a.accumulate[g] = accumulate[g]; a.scale[g] = scale[g];
  }
  if (M == 0) return JMT_OK;
  // 8 rows per block-iteration; enough blocks (6 per SM) to cover the latency of the strided d(out) gather
  const int grid = grid_for(M, 8 * 8, kNumSMs * 6);
  const size_t esz = dtype == JMT_BF16 ? 2 : 4;
  bool vec = (h_ld % 4) == 0;
  for (int g = 0; g < G; ++g)
    vec = vec && ((reinterpret_cast<uintptr_t>(h[g]) | reinterpret_cast<uintptr_t>(dh[g])) % (4 * esz)) == 0 &&
          (reinterpret_cast<uintptr_t>(w[g]) & 15) == 0;
  if (vec) JMT_DISPATCH_DTYPE(dtype, T_, (regressor_tail_bwd_kernel<T_, true><<<grid, 256, 0, (cudaStream_t)stream>>>(a)));
  else JMT_DISPATCH_DTYPE(dtype, T_, (regressor_tail_bwd_kernel<T_, false><<<grid, 256, 0, (cudaStream_t)stream>>>(a)));
  return check_launch("regressor_tail_bwd_kernel");
}
