// Elementwise / small reduction kernels of the fusion + TCN path (all HBM-bound, vectorised).
#include "common.cuh"

namespace jmt {

constexpr int kEwThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t n8, int64_t n, float slope) {
  pdl_wait();          // launched with the PDL attribute: the predecessor grid is complete from here on
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n8; i += nt) {
    Vec8<T> a, b; a.load(dy + i * 8); b.load(y + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) a.v[k] = b.v[k] > 0.f ? a.v[k] : a.v[k] * slope;
    a.store(dx + i * 8);
  }
  for (int64_t i = n8 * 8 + tid; i < n; i += nt) {
    const float g = to_f32(dy[i]);
    dx[i] = from_f32<T>(to_f32(y[i]) > 0.f ? g : g * slope);
  }
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
add_act_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n8, int64_t n, int act, float slope) {
  pdl_wait();          // launched with the PDL attribute: the predecessor grid is complete from here on
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n8; i += nt) {
    Vec8<T> x, y; x.load(a + i * 8); y.load(b + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) x.v[k] = apply_act(x.v[k] + y.v[k], act, slope);
    x.store(out + i * 8);
  }
  for (int64_t i = n8 * 8 + tid; i < n; i += nt) out[i] = from_f32<T>(apply_act(to_f32(a[i]) + to_f32(b[i]), act, slope));
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
axpy_kernel(const T* __restrict__ x, T* __restrict__ y, float a, int64_t n8, int64_t n) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n8; i += nt) {
    Vec8<T> u, v; u.load(x + i * 8); v.load(y + i * 8);
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = fmaf(a, u.v[k], v.v[k]);
    v.store(y + i * 8);
  }
  for (int64_t i = n8 * 8 + tid; i < n; i += nt) y[i] = from_f32<T>(fmaf(a, to_f32(x[i]), to_f32(y[i])));
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kEwThreads)
copy2d_kernel(const TI* __restrict__ in, int64_t in_ld, TO* __restrict__ out, int64_t out_ld, int64_t rows, int cols, int vec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int c8 = cols / 8;
    for (int64_t i = tid; i < rows * c8; i += nt) {
      const int64_t r = i / c8; const int c = (int)(i - r * c8) * 8;
      Vec8<TI> v; v.load(in + r * in_ld + c);
      Vec8<TO> o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = v.v[k];
      o.store(out + r * out_ld + c);
    }
  } else {
    for (int64_t i = tid; i < rows * cols; i += nt) {
      const int64_t r = i / cols; const int c = (int)(i - r * cols);
      out[r * out_ld + c] = from_f32<TO>(to_f32(in[r * in_ld + c]));
    }
  }
}

// per batch b (blockIdx.y): out[b*out_bs + i] = cast(in[b*in_bs + i]), i < per
template <typename TI, typename TO>
__global__ void __launch_bounds__(kEwThreads)
copy_rows3d_kernel(const TI* __restrict__ in, int64_t in_bs, TO* __restrict__ out, int64_t out_bs, int64_t per, int vec) {
  pdl_wait();          // launched with the PDL attribute: the predecessor grid is complete from here on
  const TI* ib = in + (int64_t)blockIdx.y * in_bs;
  TO* ob = out + (int64_t)blockIdx.y * out_bs;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    for (int64_t i = tid; i < per / 8; i += nt) {
      Vec8<TI> v; v.load(ib + i * 8);
      Vec8<TO> o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = v.v[k];
      o.store(ob + i * 8);
    }
  } else {
    for (int64_t i = tid; i < per; i += nt) ob[i] = from_f32<TO>(to_f32(ib[i]));
  }
}

// (nb, R, C) -> (nb, C, R) through a padded 64x64 shared tile; every thread moves PAIRS of adjacent elements on both
// sides (4-byte accesses for bf16, 8-byte for fp32), so a warp touches 128-256 contiguous bytes per instruction.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
transpose_kernel(const TI* __restrict__ in, TO* __restrict__ out, int R, int C, int64_t in_bs, int64_t out_bs) {
  __shared__ float tile[64][65];
  const int64_t b = blockIdx.z;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const TI* ib = in + b * in_bs;
  TO* ob = out + b * out_bs;
  const bool in_vec = (C % 2 == 0) && (in_bs % 2 == 0) && ((reinterpret_cast<uintptr_t>(in) & (2 * sizeof(TI) - 1)) == 0);
  const bool out_vec = (R % 2 == 0) && (out_bs % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & (2 * sizeof(TO) - 1)) == 0);
  for (int i = ty; i < 64; i += 8) {
    const int r = r0 + i, c = c0 + 2 * tx;
    float v0 = 0.f, v1 = 0.f;
    if (r < R) {
      if (in_vec && c + 1 < C) {
        if constexpr (sizeof(TI) == 2) { const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(ib + (int64_t)r * C + c); v0 = __low2float(t); v1 = __high2float(t); }
        else { const float2 t = *reinterpret_cast<const float2*>(ib + (int64_t)r * C + c); v0 = t.x; v1 = t.y; }
      } else {
        if (c < C) v0 = to_f32(ib[(int64_t)r * C + c]);
        if (c + 1 < C) v1 = to_f32(ib[(int64_t)r * C + c + 1]);
      }
    }
    tile[i][2 * tx] = v0; tile[i][2 * tx + 1] = v1;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i, r = r0 + 2 * tx;
    if (c >= C) continue;
    const float v0 = tile[2 * tx][i], v1 = tile[2 * tx + 1][i];
    if (out_vec && r + 1 < R) {
      if constexpr (sizeof(TO) == 2) *reinterpret_cast<__nv_bfloat162*>(ob + (int64_t)c * R + r) = __floats2bfloat162_rn(v0, v1);
      else *reinterpret_cast<float2*>(ob + (int64_t)c * R + r) = make_float2(v0, v1);
    } else {
      if (r < R) ob[(int64_t)c * R + r] = from_f32<TO>(v0);
      if (r + 1 < R) ob[(int64_t)c * R + r + 1] = from_f32<TO>(v1);
    }
  }
}

// bf16 -> bf16 fast path of the transpose (the (B, 1024, T) -> channels-last input transpose of the TCN, 157 MB in + 157 MB out per step):
// 8-byte global loads (16 lanes per 128-byte row segment), 16-byte global stores (8 lanes per 128-byte output row segment); the 64 x 64
// tile sits UN-transposed in shared memory (pitch 33 words), the transposition is the 2-byte gather of the read-out (2-way bank
// conflicts at most).  Needs C % 4 == 0, R % 8 == 0 and 8- / 16-byte aligned batch strides; the generic kernel above is the fallback.
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int R, int C, int64_t in_bs, int64_t out_bs) {
  __shared__ __align__(16) uint16_t tile[64][66];
  const int64_t b = blockIdx.z;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const __nv_bfloat16* ib = in + b * in_bs;
  __nv_bfloat16* ob = out + b * out_bs;
  const int t = threadIdx.x;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int r = pass * 16 + (t >> 4), c = (t & 15) * 4;
    uint2 v = make_uint2(0u, 0u);
    if (r0 + r < R && c0 + c < C) v = *reinterpret_cast<const uint2*>(ib + (int64_t)(r0 + r) * C + c0 + c);     // (C % 4 == 0: whole or nothing)
    uint32_t* dst = reinterpret_cast<uint32_t*>(&tile[r][c]);
    dst[0] = v.x; dst[1] = v.y;
  }
  __syncthreads();
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int c = pass * 32 + (t >> 3), r8 = (t & 7) * 8;
    if (c0 + c >= C || r0 + r8 >= R) continue;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (uint32_t)tile[r8 + 2 * i][c] | ((uint32_t)tile[r8 + 2 * i + 1][c] << 16);
    *reinterpret_cast<uint4*>(ob + (int64_t)(c0 + c) * R + r0 + r8) = make_uint4(w[0], w[1], w[2], w[3]);          // (R % 8 == 0)
  }
}

// out[c] += sum_r x[r*ld + c].  Block = 32 column-groups (8 columns each, one 16-byte load) x 8 row lanes;
// grid.y slabs of rows; per-block shared reduction then one fp32 atomic per column.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int cols, float* __restrict__ out, int vec) {
  pdl_wait();          // launched with the PDL attribute: the predecessor grid is complete from here on
  __shared__ float sh[8][257];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int c0 = blockIdx.x * 256 + tx * 8;
  if (vec) {
    if (c0 < cols) {
      // four independent 16-byte loads in flight per thread (one load per iteration left HBM at ~45 % of peak)
      const int64_t rs = (int64_t)gridDim.y * 8;
      int64_t r = (int64_t)blockIdx.y * 8 + ty;
      for (; r + 3 * rs < rows; r += 4 * rs) {
        Vec8<T> v0, v1, v2, v3;
        v0.load(x + r * ld + c0); v1.load(x + (r + rs) * ld + c0);
        v2.load(x + (r + 2 * rs) * ld + c0); v3.load(x + (r + 3 * rs) * ld + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += (v0.v[i] + v1.v[i]) + (v2.v[i] + v3.v[i]);
      }
      for (; r < rows; r += rs) {
        Vec8<T> v; v.load(x + r * ld + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += v.v[i];
      }
    }
  } else {
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < rows; r += (int64_t)gridDim.y * 8)
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (c0 + i < cols) acc[i] += to_f32(x[r * ld + c0 + i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

// dx = (mask ? dy * scale : 0) * (y > 0 ? 1 : slope) and colsum[c] += sum_r dx[r, c] in ONE pass over dy / y
// (backward of Linear/Conv -> activation -> channel dropout: activation gradient, dropout replay and the bias
// gradient, which were three kernels = 5 passes over a (rows, C) matrix).  Same block shape as colsum_kernel:
// 32 column groups of 8 x 8 row lanes.  mask is per (row / L, c) (channel dropout of one sample), nullable.
template <typename T>
__global__ void __launch_bounds__(256)
act_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ y, const uint8_t* __restrict__ mask, T* __restrict__ dx,
                     int64_t rows, int cols, int L, float scale, float slope, float* __restrict__ colsum) {
  pdl_wait();          // launched with the PDL attribute: the predecessor grid is complete from here on
  __shared__ float sh[8][257];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int c0 = blockIdx.x * 256 + tx * 8;
  if (c0 < cols) {
    auto finish_row = [&](Vec8<T>& g, const Vec8<T>& a, int64_t r) {
      if (mask) {
        const uint2 mk = *reinterpret_cast<const uint2*>(mask + (r / L) * cols + c0);
        const uint8_t* mb = reinterpret_cast<const uint8_t*>(&mk);
#pragma unroll
        for (int k = 0; k < 8; ++k) g.v[k] = mb[k] ? g.v[k] * scale : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        g.v[k] = a.v[k] > 0.f ? g.v[k] : g.v[k] * slope;
        acc[k] += g.v[k];
      }
      g.store(dx + r * cols + c0);
    };
    // two rows (four independent 16-byte loads) in flight per thread
    const int64_t rs = (int64_t)gridDim.y * 8;
    int64_t r = (int64_t)blockIdx.y * 8 + ty;
    for (; r + rs < rows; r += 2 * rs) {
      Vec8<T> g0, a0, g1, a1;
      g0.load(dy + r * cols + c0); a0.load(y + r * cols + c0);
      g1.load(dy + (r + rs) * cols + c0); a1.load(y + (r + rs) * cols + c0);
      finish_row(g0, a0, r);
      finish_row(g1, a1, r + rs);
    }
    if (r < rows) {
      Vec8<T> g, a; g.load(dy + r * cols + c0); a.load(y + r * cols + c0);
      finish_row(g, a, r);
    }
  }
  if (colsum == nullptr) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    atomicAdd(colsum + c, t);
  }
}

// Backward of out = act(a + b) fused with the backward of a = act2(pre) [* channel-dropout mask] (TemporalBlock: the residual LeakyReLU
// and conv2's LeakyReLU + Dropout2d, temporal_convolutional_model.py:54-57, 35-36):
//   dz  = dy * act'(out)                      -> gradient of b (the residual branch) and of a
//   dz2 = dz * [mask * scale] * act2'(a)      -> gradient of conv2's pre-activation;  colsum += column sums of dz2 (conv2's bias gradient)
// One pass (3 reads, 2 writes) instead of act_bwd (2 + 1) followed by act_bwd_fused (2 + 1).  Same block shape as act_bwd_fused_kernel.
template <typename T>
__global__ void __launch_bounds__(256)
add_act_bwd_fused_kernel(const T* __restrict__ dy, const T* __restrict__ out, const T* __restrict__ a, const uint8_t* __restrict__ mask,
                         T* __restrict__ dz, T* __restrict__ dz2, int64_t rows, int cols, int L, float scale, float slope, float slope2,
                         float* __restrict__ colsum) {
  __shared__ float sh[8][257];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const int c0 = blockIdx.x * 256 + tx * 8;
  if (c0 < cols) {
    const int64_t rs = (int64_t)gridDim.y * 8;
    for (int64_t r = (int64_t)blockIdx.y * 8 + ty; r < rows; r += rs) {
      Vec8<T> g, o, y2;
      g.load(dy + r * cols + c0); o.load(out + r * cols + c0); y2.load(a + r * cols + c0);
#pragma unroll
      for (int k = 0; k < 8; ++k) g.v[k] = o.v[k] > 0.f ? g.v[k] : g.v[k] * slope;
      g.store(dz + r * cols + c0);
      // (dz2 continues from the ROUNDED dz, as the two-kernel path does: bit-identical results)
      Vec8<T> h;
#pragma unroll
      for (int k = 0; k < 8; ++k) h.v[k] = to_f32(from_f32<T>(g.v[k]));
      if (mask) {
        const uint2 mk = *reinterpret_cast<const uint2*>(mask + (r / L) * cols + c0);
        const uint8_t* mb = reinterpret_cast<const uint8_t*>(&mk);
#pragma unroll
        for (int k = 0; k < 8; ++k) h.v[k] = mb[k] ? h.v[k] * scale : 0.f;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        h.v[k] = y2.v[k] > 0.f ? h.v[k] : h.v[k] * slope2;
        acc[k] += h.v[k];
      }
      h.store(dz2 + r * cols + c0);
    }
  }
  if (colsum == nullptr) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[ty][tx * 8 + i] = acc[i];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    atomicAdd(colsum + c, t);
  }
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
apply_mask_kernel(const T* __restrict__ x, const uint8_t* __restrict__ mask, T* __restrict__ out, int64_t total, int L,
                  int C, int per_channel, float scale, int vec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  if (vec) {      // C % 8 == 0: one 16-byte (bf16) / 32-byte (fp32) access and one 8-byte mask load per thread-iteration
    const int c8n = C / 8;
    const int64_t rows = total / C;
    for (int64_t i = tid; i < rows * c8n; i += nt) {
      const int64_t r = i / c8n;
      const int c = (int)(i - r * c8n) * 8;
      const int64_t mi = per_channel ? (r / L) * C + c : r * C + c;
      const uint2 mk = *reinterpret_cast<const uint2*>(mask + mi);
      const uint8_t* mb = reinterpret_cast<const uint8_t*>(&mk);
      Vec8<T> v; v.load(x + r * C + c);
#pragma unroll
      for (int k = 0; k < 8; ++k) v.v[k] = mb[k] ? v.v[k] * scale : 0.f;
      v.store(out + r * C + c);
    }
    return;
  }
  for (int64_t i = tid; i < total; i += nt) {
    int64_t mi = i;
    if (per_channel) { const int64_t b = i / ((int64_t)L * C); const int c = (int)(i % C); mi = b * C + c; }
    out[i] = from_f32<T>(mask[mi] ? to_f32(x[i]) * scale : 0.f);
  }
}

// Philox-4x32-10 (Salmon et al. 2011) counter-based generator
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

// `state` (nullable): device-resident (seed increment, counter offset) added to the host-side values.  A captured CUDA graph
// bakes the host values in; the device counter (advanced once per forward by rng_advance_kernel) is what makes every
// REPLAY draw fresh masks -- the same scheme torch uses for graph-safe RNG.
__global__ void __launch_bounds__(kEwThreads)
dropout_mask_kernel(uint8_t* __restrict__ mask, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* __restrict__ state) {
  if (state) { seed += state[0]; offset += state[1]; }
  const int64_t n4 = (n + 3) / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t c = offset + (uint64_t)i;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t j = i * 4 + k;
      if (j < n) mask[j] = ((float)rr[k] * 2.3283064365386963e-10f) >= p ? 1 : 0;
    }
  }
}

__global__ void rng_advance_kernel(uint64_t* state, uint64_t inc) { state[1] += inc; }

// weight_norm: w = g * v / ||v|| per output channel (legacy torch weight_norm), emitted in the two
// implicit-GEMM layouts.  One block per output channel stages its (cin, k) slice in shared memory so that both the
// read of v (cin-major, tap-minor) and the write of w_fwd (tap-major, cin-minor) are coalesced; the dgrad layout
// (cin, k*cout) is a per-tap transpose of w_fwd done by a tiled kernel (coalesced both ways).
// One launch serves up to kWnMax weight-normed convolutions (all convs of a TemporalConvNet: 8 tiny kernels per pass were
// latency-bound at ~10 us each): blockIdx.y (fwd / bwd) or blockIdx.z / k (layout) selects the conv.
constexpr int kWnMax = 16;
struct WnBatch {
  const float* g[kWnMax]; const float* v[kWnMax];
  void* w_fwd[kWnMax]; void* w_dg[kWnMax]; float* norm[kWnMax];
  const float* dw[kWnMax]; float* dgr[kWnMax]; float* dv[kWnMax];
  int cout[kWnMax], cin[kWnMax];
  int n, k;
};

template <typename TO>
__global__ void __launch_bounds__(256)
weight_norm_fwd_kernel(const WnBatch b) {
  extern __shared__ float wn_sh[];            // inner floats
  const int e = blockIdx.y, co = blockIdx.x;
  const int cout = b.cout[e], cin = b.cin[e], k = b.k;
  if (co >= cout) return;
  const float* __restrict__ g = b.g[e];
  TO* __restrict__ w_fwd = (TO*)b.w_fwd[e];
  float* __restrict__ norm = b.norm[e];
  const int inner = cin * k;
  const float* vr = b.v[e] + (int64_t)co * inner;
  float ss = 0.f;
  for (int i = threadIdx.x; i < inner; i += blockDim.x) { const float x = vr[i]; wn_sh[i] = x; ss = fmaf(x, x, ss); }
  __shared__ float sh[8];
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += sh[i];
  const float nrm = sqrtf(tot);
  if (threadIdx.x == 0 && norm) norm[co] = nrm;
  const float sc = g[co] / nrm;
  for (int o = threadIdx.x; o < inner; o += blockDim.x) {        // o = j * cin + ci  <-  v index ci * k + j
    const int j = o / cin, ci = o - j * cin;
    w_fwd[(int64_t)co * inner + o] = from_f32<TO>(wn_sh[ci * k + j] * sc);
  }
}

// dgrad layout: w_dgrad[ci, j*cout + co] = w_fwd[co, j*cin + ci]: per tap a (cout x cin) -> (cin x cout) transpose with row
// pitches k*cin / k*cout, 32x32 tiles through shared memory
template <typename T>
__global__ void __launch_bounds__(256)
weight_norm_dgrad_layout_kernel(const WnBatch b) {
  __shared__ float tile[32][33];
  const int e = blockIdx.z / b.k, j = blockIdx.z - e * b.k;
  const int cout = b.cout[e], cin = b.cin[e], k = b.k;
  const T* __restrict__ w_fwd = (const T*)b.w_fwd[e];
  T* __restrict__ w_dgrad = (T*)b.w_dg[e];
  if (w_dgrad == nullptr) return;
  const int co0 = blockIdx.y * 32, ci0 = blockIdx.x * 32;
  if (co0 >= cout || ci0 >= cin) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int co = co0 + i, ci = ci0 + tx;
    tile[i][tx] = (co < cout && ci < cin) ? to_f32(w_fwd[(int64_t)co * k * cin + (int64_t)j * cin + ci]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int ci = ci0 + i, co = co0 + tx;
    if (ci < cin && co < cout) w_dgrad[(int64_t)ci * k * cout + (int64_t)j * cout + co] = from_f32<T>(tile[tx][i]);
  }
}

// dg[co] = <dw, v> / ||v||,  dv = g/||v|| * (dw - v * <dw, v> / ||v||^2); dw arrives tap-major (co, j*cin + ci): both rows are
// staged in shared memory so every global access is coalesced
__global__ void __launch_bounds__(256)
weight_norm_bwd_kernel(const WnBatch b) {
  const int e = blockIdx.y;
  const int cout = b.cout[e], cin = b.cin[e], k = b.k;
  if ((int)blockIdx.x >= cout || b.dw[e] == nullptr) return;      // (a conv whose output never received a gradient)
  const float* __restrict__ dw_fwd = b.dw[e];
  const float* __restrict__ g = b.g[e];
  const float* __restrict__ v = b.v[e];
  const float* __restrict__ norm = b.norm[e];
  float* __restrict__ dg = b.dgr[e];
  float* __restrict__ dv = b.dv[e];
  extern __shared__ float wn_sh[];            // [0, inner): v row (ci*k + j order), [inner, 2*inner): dw row (j*cin + ci order)
  const int co = blockIdx.x;
  const int inner = cin * k;
  float* vs = wn_sh;
  float* ds = wn_sh + inner;
  const float* vr = v + (int64_t)co * inner;
  const float* dwr = dw_fwd + (int64_t)co * inner;
  for (int i = threadIdx.x; i < inner; i += blockDim.x) { vs[i] = vr[i]; ds[i] = dwr[i]; }
  __syncthreads();
  float dot = 0.f;
  for (int o = threadIdx.x; o < inner; o += blockDim.x) {        // o = j * cin + ci
    const int j = o / cin, ci = o - j * cin;
    dot = fmaf(ds[o], vs[ci * k + j], dot);
  }
  __shared__ float sh[8];
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dot;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += sh[i];
  const float nrm = norm[co];
  if (threadIdx.x == 0) dg[co] = tot / nrm;
  const float sc = g[co] / nrm, proj = tot / (nrm * nrm);
  for (int i = threadIdx.x; i < inner; i += blockDim.x) {        // i = ci * k + j
    const int ci = i / k, j = i - ci * k;
    dv[(int64_t)co * inner + i] = sc * (ds[j * cin + ci] - vs[i] * proj);
  }
}

// max over time of a channels-last sequence batch (`temporal(features).transpose(1,2)` then
// `torch.max(ft, 1)`, I3DWSDDA.py:44 + tsav.py:216, without materialising the transpose): x rows n*bs_rows + t, t < L.
// out (N, C) in the activation dtype, arg (N, C) int32 = first t attaining the maximum (torch.max tie rule).
template <typename T>
__global__ void __launch_bounds__(256)
time_max_fwd_kernel(const T* __restrict__ x, int64_t batch_stride, int L, int C, int64_t total, T* __restrict__ out, int32_t* __restrict__ arg) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / C; const int c = (int)(i - n * C);
    const T* p = x + n * batch_stride + c;
    float best = to_f32(p[0]); int bi = 0;
    for (int t = 1; t < L; ++t) { const float v = to_f32(p[(int64_t)t * C]); if (v > best) { best = v; bi = t; } }
    out[i] = from_f32<T>(best);
    arg[i] = bi;
  }
}
// dx[n, arg[n,c], c] = dout[n, c] (dx pre-zeroed)
template <typename T>
__global__ void __launch_bounds__(256)
time_max_bwd_kernel(const T* __restrict__ dout, const int32_t* __restrict__ arg, int64_t batch_stride, int C, int64_t total, T* __restrict__ dx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / C; const int c = (int)(i - n * C);
    dx[n * batch_stride + (int64_t)arg[i] * C + c] = dout[i];
  }
}

// ---------------------------------------------------------------- tiny-sequence attention
constexpr int kMaxSmallL = 8;

template <typename T, int L>
__global__ void __launch_bounds__(256)
attn_small_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ probs, int64_t N,
                      int E, int heads, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t total = N * heads;
  const int dh = E / heads;
  const int64_t row = 3 * (int64_t)E;          // elements per (l, n) row of qkv
  for (int64_t w = w0; w < total; w += (int64_t)gridDim.x * 8) {
    const int64_t n = w / heads; const int h = (int)(w - n * heads);
    const T* base = qkv + n * row + (int64_t)h * dh;
    float p[L][L];
    for (int i = 0; i < L; ++i) {
      float m = -INFINITY;
      for (int j = 0; j < L; ++j) {
        const T* q = base + (int64_t)i * N * row;
        const T* k = base + (int64_t)j * N * row + E;
        float s = 0.f;
        for (int d = lane; d < dh; d += 32) s = fmaf(to_f32(q[d]), to_f32(k[d]), s);
        s = warp_sum(s) * scale;
        p[i][j] = s; m = fmaxf(m, s);
      }
      float sum = 0.f;
      for (int j = 0; j < L; ++j) { p[i][j] = __expf(p[i][j] - m); sum += p[i][j]; }
      const float inv = 1.f / sum;
      for (int j = 0; j < L; ++j) {
        p[i][j] *= inv;
        if (lane == 0) probs[(w * L + i) * L + j] = p[i][j];
      }
    }
    for (int i = 0; i < L; ++i) {
      T* o = out + ((int64_t)i * N + n) * E + (int64_t)h * dh;
      for (int d = lane; d < dh; d += 32) {
        float acc = 0.f;
        for (int j = 0; j < L; ++j) acc = fmaf(p[i][j], to_f32(base[(int64_t)j * N * row + 2 * E + d]), acc);
        o[d] = from_f32<T>(acc);
      }
    }
  }
}

template <typename T, int L>
__global__ void __launch_bounds__(256)
attn_small_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ probs,
                      T* __restrict__ dqkv, int64_t N, int E, int heads, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int64_t total = N * heads;
  const int dh = E / heads;
  const int64_t row = 3 * (int64_t)E;
  for (int64_t w = w0; w < total; w += (int64_t)gridDim.x * 8) {
    const int64_t n = w / heads; const int h = (int)(w - n * heads);
    const T* base = qkv + n * row + (int64_t)h * dh;
    T* dbase = dqkv + n * row + (int64_t)h * dh;
    float p[L][L], ds[L][L];
    for (int i = 0; i < L; ++i) {
      const T* go = dout + ((int64_t)i * N + n) * E + (int64_t)h * dh;
      float dot = 0.f;
      for (int j = 0; j < L; ++j) {
        p[i][j] = probs[(w * L + i) * L + j];
        const T* v = base + (int64_t)j * N * row + 2 * E;
        float s = 0.f;
        for (int d = lane; d < dh; d += 32) s = fmaf(to_f32(go[d]), to_f32(v[d]), s);
        s = warp_sum(s);
        ds[i][j] = s; dot = fmaf(p[i][j], s, dot);
      }
      for (int j = 0; j < L; ++j) ds[i][j] = p[i][j] * (ds[i][j] - dot) * scale;
    }
    for (int d = lane; d < dh; d += 32) {
      for (int i = 0; i < L; ++i) {
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int j = 0; j < L; ++j) {
          dq = fmaf(ds[i][j], to_f32(base[(int64_t)j * N * row + E + d]), dq);        // dQ_i = sum_j dS_ij K_j
          dk = fmaf(ds[j][i], to_f32(base[(int64_t)j * N * row + d]), dk);            // dK_i = sum_j dS_ji Q_j
          dv = fmaf(p[j][i], to_f32(dout[((int64_t)j * N + n) * E + (int64_t)h * dh + d]), dv);  // dV_i = sum_j P_ji dO_j
        }
        T* drow = dbase + (int64_t)i * N * row;
        drow[d] = from_f32<T>(dq); drow[E + d] = from_f32<T>(dk); drow[2 * E + d] = from_f32<T>(dv);
      }
    }
  }
}

}  // namespace jmt

using namespace jmt;

extern "C" int jmt_act_bwd(const void* dy, const void* y, void* dx, int64_t n, float slope, int dtype, void* stream) {
  JMT_REQUIRE(dy && y && dx && n >= 0, "jmt_act_bwd: bad arguments");
  if (n == 0) return JMT_OK;
  const bool al = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dx)) & 31) == 0;
  const int64_t n8 = al ? n / 8 : 0;
  JMT_DISPATCH_DTYPE(dtype, T, (launch_pdl(act_bwd_kernel<T>, dim3(grid_for(n, kEwThreads * 16)), dim3(kEwThreads), 0, (cudaStream_t)stream, (const T*)dy, (const T*)y, (T*)dx, n8, n, slope)));
  return check_launch("act_bwd_kernel");
}

extern "C" int jmt_add_act(const void* a, const void* b, void* out, int64_t n, int act, float slope, int dtype, void* stream) {
  JMT_REQUIRE(a && b && out && n >= 0, "jmt_add_act: bad arguments");
  if (n == 0) return JMT_OK;
  const bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 31) == 0;
  const int64_t n8 = al ? n / 8 : 0;
  JMT_DISPATCH_DTYPE(dtype, T, (launch_pdl(add_act_kernel<T>, dim3(grid_for(n, kEwThreads * 16)), dim3(kEwThreads), 0, (cudaStream_t)stream, (const T*)a, (const T*)b, (T*)out, n8, n, act, slope)));
  return check_launch("add_act_kernel");
}

extern "C" int jmt_axpy(const void* x, void* y, float a, int64_t n, int dtype, void* stream) {
  JMT_REQUIRE(x && y && n >= 0, "jmt_axpy: bad arguments");
  if (n == 0) return JMT_OK;
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 31) == 0;
  const int64_t n8 = al ? n / 8 : 0;
  JMT_DISPATCH_DTYPE(dtype, T, (axpy_kernel<T><<<grid_for(n, kEwThreads * 16), kEwThreads, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, a, n8, n)));
  return check_launch("axpy_kernel");
}

extern "C" int jmt_copy2d(const void* in, int in_dtype, int64_t in_ld, void* out, int out_dtype, int64_t out_ld,
                          int64_t rows, int cols, void* stream) {
  JMT_REQUIRE(in && out && rows >= 0 && cols >= 0, "jmt_copy2d: bad arguments");
  if (rows * cols == 0) return JMT_OK;
  const int vec = (cols % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 31) == 0) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(rows * cols, kEwThreads * 16);
  JMT_DISPATCH_DTYPE(in_dtype, TI, JMT_DISPATCH_DTYPE(out_dtype, TO,
      (copy2d_kernel<TI, TO><<<g, kEwThreads, 0, st>>>((const TI*)in, in_ld, (TO*)out, out_ld, rows, cols, vec))));
  return check_launch("copy2d_kernel");
}

extern "C" int jmt_cast(const void* in, int in_dtype, void* out, int out_dtype, int64_t n, void* stream) {
  if (n == 0) return JMT_OK;
  if (n % 8 == 0) return jmt_copy2d(in, in_dtype, 8, out, out_dtype, 8, n / 8, 8, stream);
  JMT_REQUIRE(n < (int64_t)1 << 31, "jmt_cast: n too large for the unaligned path");
  return jmt_copy2d(in, in_dtype, n, out, out_dtype, n, 1, (int)n, stream);
}

// out[i0*os0 + i1*os1 + c] = cast(in[i0*is0 + i1*is1 + c]): a row permutation with cast in ONE launch
// (e.g. (B, T, D) activations -> the (T, B, D) layout MultimodalTransformer_w_JR returns)
template <typename TI, typename TO>
__global__ void __launch_bounds__(kEwThreads)
copy3d_kernel(const TI* __restrict__ in, int64_t is0, int64_t is1, TO* __restrict__ out, int64_t os0, int64_t os1,
              int64_t n0, int64_t n1, int cols, int vec) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  const int cw = vec ? cols / 8 : cols;
  const int64_t total = n0 * n1 * cw;
  for (int64_t i = tid; i < total; i += nt) {
    const int64_t r = i / cw; const int c = (int)(i - r * cw);
    const int64_t i0 = r / n1, i1 = r - i0 * n1;
    const TI* ip = in + i0 * is0 + i1 * is1;
    TO* op = out + i0 * os0 + i1 * os1;
    if (vec) {
      Vec8<TI> v; v.load(ip + c * 8);
      Vec8<TO> o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = v.v[k];
      o.store(op + c * 8);
    } else {
      op[c] = from_f32<TO>(to_f32(ip[c]));
    }
  }
}

extern "C" int jmt_copy3d(const void* in, int in_dtype, int64_t is0, int64_t is1, void* out, int out_dtype, int64_t os0, int64_t os1,
                          int64_t n0, int64_t n1, int cols, void* stream) {
  JMT_REQUIRE(in && out && n0 >= 0 && n1 >= 0 && cols >= 0, "jmt_copy3d: bad arguments");
  if (n0 * n1 * cols == 0) return JMT_OK;
  const int vec = (cols % 8 == 0 && is0 % 8 == 0 && is1 % 8 == 0 && os0 % 8 == 0 && os1 % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 31) == 0) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(n0 * n1 * (vec ? cols / 8 : cols), kEwThreads);
  JMT_DISPATCH_DTYPE(in_dtype, TI, JMT_DISPATCH_DTYPE(out_dtype, TO,
      (copy3d_kernel<TI, TO><<<grid, kEwThreads, 0, st>>>((const TI*)in, is0, is1, (TO*)out, os0, os1, n0, n1, cols, vec))));
  return check_launch("copy3d_kernel");
}

// fp32 -> bf16 casts of up to kCastMax tensors in one launch (the bf16 operand copies of all weight matrices of a module,
// refreshed once per optimizer step: 24 separate ~4 us launches before)
constexpr int kCastMax = 96;
struct CastBatch { const float* src[kCastMax]; __nv_bfloat16* dst[kCastMax]; int64_t n[kCastMax]; int count; };
__global__ void __launch_bounds__(kEwThreads)
cast_multi_kernel(const CastBatch b) {
  const int e = blockIdx.y;
  const float* __restrict__ x = b.src[e];
  __nv_bfloat16* __restrict__ y = b.dst[e];
  const int64_t n = b.n[e], n8 = n / 8;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n8; i += nt) {
    Vec8<float> v; v.load(x + i * 8);
    Vec8<__nv_bfloat16> o;
#pragma unroll
    for (int k = 0; k < 8; ++k) o.v[k] = v.v[k];
    o.store(y + i * 8);
  }
  for (int64_t i = n8 * 8 + tid; i < n; i += nt) y[i] = __float2bfloat16_rn(x[i]);
}

extern "C" int jmt_cast_multi(int count, const float* const* src, void* const* dst, const int64_t* n, void* stream) {
  JMT_REQUIRE(count >= 0 && src && dst && n, "jmt_cast_multi: bad arguments");
  for (int base = 0; base < count; base += kCastMax) {
    CastBatch b;
    b.count = count - base < kCastMax ? count - base : kCastMax;
    for (int e = 0; e < b.count; ++e) {
      JMT_REQUIRE(src[base + e] && dst[base + e] && n[base + e] >= 0, "jmt_cast_multi: bad entry %d", base + e);
      JMT_REQUIRE(((reinterpret_cast<uintptr_t>(src[base + e]) | reinterpret_cast<uintptr_t>(dst[base + e])) & 15) == 0,
                  "jmt_cast_multi: entry %d is not 16-byte aligned", base + e);
      b.src[e] = src[base + e]; b.dst[e] = (__nv_bfloat16*)dst[base + e]; b.n[e] = n[base + e];
    }
    cast_multi_kernel<<<dim3(96, b.count), kEwThreads, 0, (cudaStream_t)stream>>>(b);
    int rc = check_launch("cast_multi_kernel");
    if (rc != JMT_OK) return rc;
  }
  return JMT_OK;
}

// bf16x3 operand split: hi = bf16(x), lo = bf16(x - hi); x ~= hi + lo to 2^-17 relative
__global__ void __launch_bounds__(kEwThreads)
split_bf16x2_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t n) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = n / 8;
  for (int64_t i = tid; i < n8; i += nt) {
    Vec8<float> v; v.load(x + i * 8);
    Vec8<__nv_bfloat16> h, l;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float hf = __bfloat162float(__float2bfloat16_rn(v.v[k]));
      h.v[k] = hf; l.v[k] = v.v[k] - hf;
    }
    h.store(hi + i * 8); l.store(lo + i * 8);
  }
  for (int64_t i = n8 * 8 + tid; i < n; i += nt) {
    const __nv_bfloat16 hb = __float2bfloat16_rn(x[i]);
    hi[i] = hb; lo[i] = __float2bfloat16_rn(x[i] - __bfloat162float(hb));
  }
}

extern "C" int jmt_split_bf16x2(const float* x, void* hi, void* lo, int64_t n, void* stream) {
  if (n == 0) return JMT_OK;
  JMT_REQUIRE(x && hi && lo && n > 0, "jmt_split_bf16x2: bad arguments");
  JMT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0,
              "jmt_split_bf16x2: x, hi and lo must be 16-byte aligned");
  split_bf16x2_kernel<<<grid_for((n + 7) / 8, kEwThreads), kEwThreads, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  return check_launch("split_bf16x2_kernel");
}

extern "C" int jmt_transpose_strided(const void* in, int in_dtype, int64_t in_bs, void* out, int out_dtype, int64_t out_bs, int64_t nb,
                                     int R, int C, void* stream) {
  JMT_REQUIRE(in && out && nb >= 0 && R >= 0 && C >= 0 && nb < 65536, "jmt_transpose: bad arguments");
  if (nb * R * C == 0) return JMT_OK;
  if (in_bs == 0) in_bs = (int64_t)R * C;
  if (out_bs == 0) out_bs = (int64_t)R * C;
  dim3 grid((C + 63) / 64, (R + 63) / 64, (unsigned)nb);
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == JMT_BF16 && out_dtype == JMT_BF16 && C % 4 == 0 && R % 8 == 0 && in_bs % 4 == 0 && out_bs % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(in) & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    transpose_bf16_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, C, in_bs, out_bs);
    return check_launch("transpose_bf16_kernel");
  }
  JMT_DISPATCH_DTYPE(in_dtype, TI, JMT_DISPATCH_DTYPE(out_dtype, TO,
      (transpose_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)in, (TO*)out, R, C, in_bs, out_bs))));
  return check_launch("transpose_kernel");
}

extern "C" int jmt_transpose(const void* in, int in_dtype, void* out, int out_dtype, int64_t nb, int R, int C, void* stream) {
  return jmt_transpose_strided(in, in_dtype, 0, out, out_dtype, 0, nb, R, C, stream);
}

extern "C" int jmt_copy_rows3d(const void* in, int in_dtype, int64_t in_bs, void* out, int out_dtype, int64_t out_bs, int64_t nb,
                               int64_t rows, int cols, void* stream) {
  JMT_REQUIRE(in && out && nb >= 0 && rows >= 0 && cols >= 0 && nb < 65536, "jmt_copy_rows3d: bad arguments");
  if (nb * rows * cols == 0) return JMT_OK;
  const int64_t per = rows * cols;
  const int vec = (per % 8 == 0 && in_bs % 8 == 0 && out_bs % 8 == 0 &&
                   ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 31) == 0) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  int gx = grid_for(per, kEwThreads * 8, kNumSMs * 8);
  dim3 grid(gx, (unsigned)nb);
  JMT_DISPATCH_DTYPE(in_dtype, TI, JMT_DISPATCH_DTYPE(out_dtype, TO,
      (launch_pdl(copy_rows3d_kernel<TI, TO>, grid, dim3(kEwThreads), 0, st, (const TI*)in, in_bs, (TO*)out, out_bs, per, vec))));
  return check_launch("copy_rows3d_kernel");
}

extern "C" int jmt_colsum(const void* x, int dtype, int64_t ld, int64_t rows, int cols, float* out, void* stream) {
  JMT_REQUIRE(x && out && rows >= 0 && cols > 0, "jmt_colsum: bad arguments");
  if (rows == 0) return JMT_OK;
  const int gx = (cols + 255) / 256;
  int gy = (int)((rows + 63) / 64);
  const int cap = (kNumSMs * 4 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  const int vec = (cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
  JMT_DISPATCH_DTYPE(dtype, T, (launch_pdl(colsum_kernel<T>, dim3(gx, gy), dim3(256), 0, (cudaStream_t)stream, (const T*)x, ld, rows, cols, out, vec)));
  return check_launch("colsum_kernel");
}

extern "C" int jmt_act_bwd_fused(const void* dy, const void* y, const uint8_t* mask, void* dx, int64_t rows, int cols, int L,
                                 float scale, float slope, float* colsum, int dtype, void* stream) {
  JMT_REQUIRE(dy && y && dx && rows >= 0 && cols > 0 && cols % 8 == 0 && L >= 1, "jmt_act_bwd_fused: bad arguments (cols %% 8)");
  JMT_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(mask) & 7) == 0, "jmt_act_bwd_fused: pointers must be 16-byte aligned");
  if (rows == 0) return JMT_OK;
  const int gx = (cols + 255) / 256;
  int gy = (int)((rows + 63) / 64);
  const int cap = (kNumSMs * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  JMT_DISPATCH_DTYPE(dtype, T, (act_bwd_fused_kernel<T><<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(
      (const T*)dy, (const T*)y, mask, (T*)dx, rows, cols, L, scale, slope, colsum)));
  return check_launch("act_bwd_fused_kernel");
}

extern "C" int jmt_add_act_bwd_fused(const void* dy, const void* out, const void* a, const uint8_t* mask, void* dz, void* dz2,
                                     int64_t rows, int cols, int L, float scale, float slope, float slope2, float* colsum, int dtype,
                                     void* stream) {
  JMT_REQUIRE(dy && out && a && dz && dz2 && rows >= 0 && cols > 0 && cols % 8 == 0 && L >= 1, "jmt_add_act_bwd_fused: bad arguments (cols %% 8)");
  JMT_REQUIRE(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(a) |
                reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dz2)) & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 7) == 0,
              "jmt_add_act_bwd_fused: pointers must be 16-byte aligned");
  if (rows == 0) return JMT_OK;
  const int gx = (cols + 255) / 256;
  int gy = (int)((rows + 63) / 64);
  const int cap = (kNumSMs * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  JMT_DISPATCH_DTYPE(dtype, T, (add_act_bwd_fused_kernel<T><<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(
      (const T*)dy, (const T*)out, (const T*)a, mask, (T*)dz, (T*)dz2, rows, cols, L, scale, slope, slope2, colsum)));
  return check_launch("add_act_bwd_fused_kernel");
}

extern "C" int jmt_apply_mask(const void* x, const uint8_t* mask, void* out, int64_t nb, int L, int C, int per_channel,
                              float scale, int dtype, void* stream) {
  JMT_REQUIRE(x && mask && out, "jmt_apply_mask: bad arguments");
  const int64_t total = nb * L * C;
  if (total == 0) return JMT_OK;
  const int vec = (C % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 31) == 0 &&
                   (reinterpret_cast<uintptr_t>(mask) & 7) == 0) ? 1 : 0;
  JMT_DISPATCH_DTYPE(dtype, T, (apply_mask_kernel<T><<<grid_for(total / 8 + 1, kEwThreads * 2), kEwThreads, 0, (cudaStream_t)stream>>>((const T*)x, mask, (T*)out, total, L, C, per_channel, scale, vec)));
  return check_launch("apply_mask_kernel");
}

extern "C" int jmt_dropout_mask(uint8_t* mask, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* dev_state,
                                void* stream) {
  JMT_REQUIRE(mask && n >= 0 && p >= 0.f && p < 1.f, "jmt_dropout_mask: bad arguments");
  if (n == 0) return JMT_OK;
  dropout_mask_kernel<<<grid_for((n + 3) / 4, kEwThreads), kEwThreads, 0, (cudaStream_t)stream>>>(mask, n, p, seed, offset, dev_state);
  return check_launch("dropout_mask_kernel");
}

extern "C" int jmt_rng_advance(uint64_t* dev_state, uint64_t inc, void* stream) {
  JMT_REQUIRE(dev_state, "jmt_rng_advance: null state");
  rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dev_state, inc);
  return check_launch("rng_advance_kernel");
}

static int wn_fwd_launch(const jmt::WnBatch& b, int out_dtype, bool any_dgrad, cudaStream_t st) {
  int max_cout = 0, max_cin = 0;
  for (int e = 0; e < b.n; ++e) { max_cout = b.cout[e] > max_cout ? b.cout[e] : max_cout; max_cin = b.cin[e] > max_cin ? b.cin[e] : max_cin; }
  const size_t sh = (size_t)max_cin * b.k * sizeof(float);
  JMT_REQUIRE(sh <= 96 * 1024, "jmt_weight_norm_fwd: cin*k too large for the shared-memory staging (%d x %d)", max_cin, b.k);
  if (sh > 48 * 1024) {
    JMT_DISPATCH_DTYPE(out_dtype, TO, cudaFuncSetAttribute(weight_norm_fwd_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  }
  JMT_DISPATCH_DTYPE(out_dtype, TO, (weight_norm_fwd_kernel<TO><<<dim3(max_cout, b.n), 256, sh, st>>>(b)));
  int rc = check_launch("weight_norm_fwd_kernel");
  if (rc != JMT_OK || !any_dgrad) return rc;
  dim3 grid((max_cin + 31) / 32, (max_cout + 31) / 32, b.k * b.n);
  JMT_DISPATCH_DTYPE(out_dtype, TO, (weight_norm_dgrad_layout_kernel<TO><<<grid, 256, 0, st>>>(b)));
  return check_launch("weight_norm_dgrad_layout_kernel");
}

extern "C" int jmt_weight_norm_fwd(const float* g, const float* v, void* w_fwd, void* w_dgrad, int out_dtype, float* norm,
                                   int cout, int cin, int k, void* stream) {
  JMT_REQUIRE(g && v && w_fwd && cout > 0 && cin > 0 && k > 0, "jmt_weight_norm_fwd: bad arguments");
  jmt::WnBatch b;
  memset(&b, 0, sizeof(b));
  b.n = 1; b.k = k; b.g[0] = g; b.v[0] = v; b.w_fwd[0] = w_fwd; b.w_dg[0] = w_dgrad; b.norm[0] = norm; b.cout[0] = cout; b.cin[0] = cin;
  return wn_fwd_launch(b, out_dtype, w_dgrad != nullptr, (cudaStream_t)stream);
}

extern "C" int jmt_weight_norm_fwd_batched(int n, const float* const* g, const float* const* v, void* const* w_fwd, void* const* w_dgrad,
                                           int out_dtype, float* const* norm, const int* cout, const int* cin, int k, void* stream) {
  JMT_REQUIRE(n >= 1 && n <= jmt::kWnMax && g && v && w_fwd && norm && cout && cin && k > 0, "jmt_weight_norm_fwd_batched: bad arguments");
  jmt::WnBatch b;
  memset(&b, 0, sizeof(b));
  b.n = n; b.k = k;
  bool any_dg = false;
  for (int e = 0; e < n; ++e) {
    JMT_REQUIRE(g[e] && v[e] && w_fwd[e] && cout[e] > 0 && cin[e] > 0, "jmt_weight_norm_fwd_batched: bad entry %d", e);
    b.g[e] = g[e]; b.v[e] = v[e]; b.w_fwd[e] = w_fwd[e]; b.w_dg[e] = w_dgrad ? w_dgrad[e] : nullptr; b.norm[e] = norm[e];
    b.cout[e] = cout[e]; b.cin[e] = cin[e];
    any_dg = any_dg || b.w_dg[e] != nullptr;
  }
  return wn_fwd_launch(b, out_dtype, any_dg, (cudaStream_t)stream);
}

static int wn_bwd_launch(const jmt::WnBatch& b, cudaStream_t st) {
  int max_cout = 0, max_cin = 0;
  for (int e = 0; e < b.n; ++e) { max_cout = b.cout[e] > max_cout ? b.cout[e] : max_cout; max_cin = b.cin[e] > max_cin ? b.cin[e] : max_cin; }
  const size_t sh = 2 * (size_t)max_cin * b.k * sizeof(float);
  JMT_REQUIRE(sh <= 96 * 1024, "jmt_weight_norm_bwd: cin*k too large for the shared-memory staging (%d x %d)", max_cin, b.k);
  if (sh > 48 * 1024) cudaFuncSetAttribute(weight_norm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  weight_norm_bwd_kernel<<<dim3(max_cout, b.n), 256, sh, st>>>(b);
  return check_launch("weight_norm_bwd_kernel");
}

extern "C" int jmt_weight_norm_bwd(const float* dw_fwd, const float* g, const float* v, const float* norm, float* dg,
                                   float* dv, int cout, int cin, int k, void* stream) {
  JMT_REQUIRE(dw_fwd && g && v && norm && dg && dv, "jmt_weight_norm_bwd: bad arguments");
  jmt::WnBatch b;
  memset(&b, 0, sizeof(b));
  b.n = 1; b.k = k; b.dw[0] = dw_fwd; b.g[0] = g; b.v[0] = v; b.norm[0] = const_cast<float*>(norm); b.dgr[0] = dg; b.dv[0] = dv;
  b.cout[0] = cout; b.cin[0] = cin;
  return wn_bwd_launch(b, (cudaStream_t)stream);
}

extern "C" int jmt_weight_norm_bwd_batched(int n, const float* const* dw_fwd, const float* const* g, const float* const* v,
                                           const float* const* norm, float* const* dg, float* const* dv, const int* cout,
                                           const int* cin, int k, void* stream) {
  JMT_REQUIRE(n >= 1 && n <= jmt::kWnMax && dw_fwd && g && v && norm && dg && dv && cout && cin && k > 0, "jmt_weight_norm_bwd_batched: bad arguments");
  jmt::WnBatch b;
  memset(&b, 0, sizeof(b));
  b.n = n; b.k = k;
  for (int e = 0; e < n; ++e) {
    JMT_REQUIRE(g[e] && v[e] && norm[e] && dg[e] && dv[e] && cout[e] > 0 && cin[e] > 0, "jmt_weight_norm_bwd_batched: bad entry %d", e);
    b.dw[e] = dw_fwd[e];      // NULL: that conv received no gradient (its dg / dv stay untouched)
    b.g[e] = g[e]; b.v[e] = v[e]; b.norm[e] = const_cast<float*>(norm[e]); b.dgr[e] = dg[e]; b.dv[e] = dv[e];
    b.cout[e] = cout[e]; b.cin[e] = cin[e];
  }
  return wn_bwd_launch(b, (cudaStream_t)stream);
}

extern "C" int jmt_time_max_fwd(const void* x, int64_t batch_stride, int64_t nb, int L, int C, void* out, int32_t* arg, int dtype, void* stream) {
  JMT_REQUIRE(x && out && arg && nb >= 0 && L >= 1 && C >= 1, "jmt_time_max_fwd: bad arguments");
  if (nb == 0) return JMT_OK;
  const int64_t total = nb * C;
  JMT_DISPATCH_DTYPE(dtype, T, (time_max_fwd_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, batch_stride, L, C, total, (T*)out, arg)));
  return check_launch("time_max_fwd_kernel");
}

extern "C" int jmt_time_max_bwd(const void* dout, const int32_t* arg, int64_t batch_stride, int64_t nb, int C, void* dx, int dtype, void* stream) {
  JMT_REQUIRE(dout && arg && dx && nb >= 0 && C >= 1, "jmt_time_max_bwd: bad arguments");
  if (nb == 0) return JMT_OK;
  const int64_t total = nb * C;
  JMT_DISPATCH_DTYPE(dtype, T, (time_max_bwd_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)dout, arg, batch_stride, C, total, (T*)dx)));
  return check_launch("time_max_bwd_kernel");
}

extern "C" int jmt_attn_small_fwd(const void* qkv, void* out, float* probs, int L, int64_t N, int E, int heads,
                                  float scale, int dtype, void* stream) {
  JMT_REQUIRE(qkv && out && probs && L >= 1 && L <= kMaxSmallL && heads >= 1 && E % heads == 0, "jmt_attn_small_fwd: need 1 <= L <= 8");
  if (N == 0) return JMT_OK;
  const int g = grid_for(N * heads, 8, kNumSMs * 8);
#define JMT_L_CASE(LL) case LL: attn_small_fwd_kernel<T, LL><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)qkv, (T*)out, probs, N, E, heads, scale); break;
  JMT_DISPATCH_DTYPE(dtype, T, switch (L) { JMT_L_CASE(1) JMT_L_CASE(2) JMT_L_CASE(3) JMT_L_CASE(4) JMT_L_CASE(5) JMT_L_CASE(6) JMT_L_CASE(7) JMT_L_CASE(8) });
#undef JMT_L_CASE
  return check_launch("attn_small_fwd_kernel");
}

extern "C" int jmt_attn_small_bwd(const void* qkv, const void* dout, const float* probs, void* dqkv, int L, int64_t N,
                                  int E, int heads, float scale, int dtype, void* stream) {
  JMT_REQUIRE(qkv && dout && probs && dqkv && L >= 1 && L <= kMaxSmallL && heads >= 1 && E % heads == 0, "jmt_attn_small_bwd: need 1 <= L <= 8");
  if (N == 0) return JMT_OK;
  const int g = grid_for(N * heads, 8, kNumSMs * 8);
#define JMT_L_CASE(LL) case LL: attn_small_bwd_kernel<T, LL><<<g, 256, 0, (cudaStream_t)stream>>>((const T*)qkv, (const T*)dout, probs, (T*)dqkv, N, E, heads, scale); break;
  JMT_DISPATCH_DTYPE(dtype, T, switch (L) { JMT_L_CASE(1) JMT_L_CASE(2) JMT_L_CASE(3) JMT_L_CASE(4) JMT_L_CASE(5) JMT_L_CASE(6) JMT_L_CASE(7) JMT_L_CASE(8) });
#undef JMT_L_CASE
  return check_launch("attn_small_bwd_kernel");
}
