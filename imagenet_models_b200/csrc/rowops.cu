// Row / column kernels over NHWC activation matrices [M, C]: LayerNorm, BatchNorm statistics + apply, patch
// gathers, casts.  All are HBM-bound: 128-bit (4-element) accesses, one warp per row for row reductions,
// column-strip CTAs with partial buffers for column reductions (no atomics on the result).
#include "common.cuh"

#define DISPATCH_T(dtype, ...)                         \
  if ((dtype) == GA_BF16) { typedef bf16 T; __VA_ARGS__; } \
  else { typedef float T; __VA_ARGS__; }

static inline int launch_ok(const char* n) { ga_count_launch(); return ga_check_launch(n); }

// ---------------------------------------------------------------------------------------------- LayerNorm rows
// one warp per row; row cached in registers (C <= 32*4*MAXV)
template <typename T, int MAXV>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, T* __restrict__ y,
                                                            float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                                            long long M, int C, long long ldx, long long ldy, float eps) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const T* xr = x + row * ldx;
  float4 v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    v[i] = (c < C) ? ld4(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      float a = v[i].x - mean, bb = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += a * a + bb * bb + cc * cc + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0) { if (mean_o) mean_o[row] = mean; if (rstd_o) rstd_o[row] = rstd; }
  T* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < C) {
      float4 o = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
      if (w) {
        float4 ww = *reinterpret_cast<const float4*>(w + c), bb = *reinterpret_cast<const float4*>(b + c);
        o.x = o.x * ww.x + bb.x; o.y = o.y * ww.y + bb.y; o.z = o.z * ww.z + bb.z; o.w = o.w * ww.w + bb.w;
      }
      st4(yr + c, o);
    }
  }
}

// narrow rows (C <= 128: CSWin stem / stage 1-2, ConvNeXt stem): LPR lanes per row, 32/LPR rows per warp, ROWS rows in flight
// per lane group so a warp keeps ROWS*32 128-bit loads outstanding instead of C/4
template <typename T, int LPR, int ROWS>
__global__ void __launch_bounds__(256) layernorm_fwd_narrow_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ b, T* __restrict__ y,
                                                                   float* __restrict__ mean_o, float* __restrict__ rstd_o,
                                                                   long long M, int C, long long ldx, long long ldy, float eps) {
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
  const int c = sub * 4;
  const bool cok = c < C;
  const float invC = 1.f / (float)C;
  float4 ww = make_float4(1.f, 1.f, 1.f, 1.f), bb = make_float4(0.f, 0.f, 0.f, 0.f);
  if (w && cok) { ww = *reinterpret_cast<const float4*>(w + c); bb = *reinterpret_cast<const float4*>(b + c); }
  // the trip count must be warp-uniform: the shuffles below name all 32 lanes
  for (long long base = gw * G * ROWS; base < M; base += nw * G * ROWS) {
    const long long r0 = base + (long long)grp * ROWS;
    float4 v[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) v[r] = (cok && r0 + r < M) ? ld4(x + (r0 + r) * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float mu[ROWS], rs[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) mu[r] = (v[r].x + v[r].y) + (v[r].z + v[r].w);
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) mu[r] += __shfl_xor_sync(0xffffffffu, mu[r], o);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      mu[r] *= invC;
      const float a = v[r].x - mu[r], b2 = v[r].y - mu[r], c2 = v[r].z - mu[r], d = v[r].w - mu[r];
      rs[r] = cok ? (a * a + b2 * b2) + (c2 * c2 + d * d) : 0.f;
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], o);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = r0 + r;
      if (row >= M) continue;
      const float rstd = rsqrtf(rs[r] * invC + eps);
      if (sub == 0) { if (mean_o) mean_o[row] = mu[r]; if (rstd_o) rstd_o[row] = rstd; }
      if (cok) {
        float4 o = make_float4((v[r].x - mu[r]) * rstd, (v[r].y - mu[r]) * rstd, (v[r].z - mu[r]) * rstd, (v[r].w - mu[r]) * rstd);
        o.x = o.x * ww.x + bb.x; o.y = o.y * ww.y + bb.y; o.z = o.z * ww.z + bb.z; o.w = o.w * ww.w + bb.w;
        st4(y + row * ldy + c, o);
      }
    }
  }
}

extern "C" int ga_layernorm_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd, long long M,
                                int C, long long ldx, long long ldy, float eps, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && y && M >= 0 && C > 0 && (C & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0, GA_ERR_ALIGN,
             "ga_layernorm_fwd: C=%d ldx=%lld ldy=%lld must be multiples of 4", C, ldx, ldy);
  GA_REQUIRE(C <= 2048, GA_ERR_UNSUPPORTED, "ga_layernorm_fwd: C=%d > 2048", C);
  if (M == 0) return GA_OK;
  if (C <= 128) {
    const int lpr = C <= 32 ? 8 : (C <= 64 ? 16 : 32);
    long long blocks = (M + 8LL * (32 / lpr) * 4 - 1) / (8LL * (32 / lpr) * 4);
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    DISPATCH_T(dtype, {
      if (lpr == 8) layernorm_fwd_narrow_kernel<T, 8, 4><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>((const T*)x, w, b, (T*)y, mean, rstd, M, C, ldx, ldy, eps);
      else if (lpr == 16) layernorm_fwd_narrow_kernel<T, 16, 4><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>((const T*)x, w, b, (T*)y, mean, rstd, M, C, ldx, ldy, eps);
      else layernorm_fwd_narrow_kernel<T, 32, 4><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>((const T*)x, w, b, (T*)y, mean, rstd, M, C, ldx, ldy, eps);
    });
    return launch_ok("layernorm_fwd_narrow");
  }
  const unsigned grid = (unsigned)((M + 7) / 8);
  DISPATCH_T(dtype, {
    if (C <= 512) layernorm_fwd_kernel<T, 4><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, w, b, (T*)y, mean, rstd, M, C, ldx, ldy, eps);
    else if (C <= 1024) layernorm_fwd_kernel<T, 8><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, w, b, (T*)y, mean, rstd, M, C, ldx, ldy, eps);
    else layernorm_fwd_kernel<T, 16><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, w, b, (T*)y, mean, rstd, M, C, ldx, ldy, eps);
  });
  return launch_ok("layernorm_fwd");
}

// backward.  xhat either given directly (x_is_hat) or recomputed from x, mean, rstd.
// dx = rstd * (g - mean(g) - xhat*mean(g*xhat)),  g = dy*w ; per-CTA partial dw/db -> partial[blockIdx][2][C]
// HAS_PARAM = false (no affine gradients wanted) drops the per-lane dw / db accumulators: 214 -> ~100 registers at MAXV = 8, so
// two CTAs fit an SM; MAXV <= 4 is capped at 80 registers for three.  The kernel keeps one row per warp in flight, so resident
// warps are what hides the HBM latency (ncu-free reasoning from the launch list: 1.1-1.5 TB/s before).
template <typename T, int MAXV, bool HAS_PARAM>
__global__ void __launch_bounds__(256, MAXV <= 4 ? 3 : (HAS_PARAM ? 1 : 2)) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ w, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, T* __restrict__ dx,
                                                            float* __restrict__ partial, long long M, int C, long long lddy,
                                                            long long ldx, long long lddx, int x_is_hat, int rows_per_cta,
                                                            float* __restrict__ adw, float* __restrict__ adb) {
  extern __shared__ float sacc[];  // [2][C] per CTA when partial != NULL
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (HAS_PARAM) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
  }
  float4 dwv[HAS_PARAM ? MAXV : 1], dbv[HAS_PARAM ? MAXV : 1];
#pragma unroll
  for (int i = 0; i < (HAS_PARAM ? MAXV : 1); ++i) { dwv[i] = make_float4(0, 0, 0, 0); dbv[i] = make_float4(0, 0, 0, 0); }
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long row_end = r0 + rows_per_cta;
  if (row_end > M) row_end = M;
  // prefetch hint for the row after next keeps two rows of loads in flight per warp
  for (long long row = r0 + wid; row < row_end; row += nw) {
    if (row + nw < row_end) {
      const int c0 = lane * 4;
      if (c0 < C) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(dy + (row + nw) * lddy + c0));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(x + (row + nw) * ldx + c0));
      }
    }
    const float mu = x_is_hat ? 0.f : mean[row];
    const float rs = rstd[row];
    float4 g[MAXV], xh[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        float4 d = ld4(dy + row * lddy + c);
        float4 xv = ld4(x + row * ldx + c);
        if (!x_is_hat) { xv.x = (xv.x - mu) * rs; xv.y = (xv.y - mu) * rs; xv.z = (xv.z - mu) * rs; xv.w = (xv.w - mu) * rs; }
        if (HAS_PARAM) {
          dwv[i].x += d.x * xv.x; dwv[i].y += d.y * xv.y; dwv[i].z += d.z * xv.z; dwv[i].w += d.w * xv.w;
          dbv[i].x += d.x; dbv[i].y += d.y; dbv[i].z += d.z; dbv[i].w += d.w;
        }
        if (w) { float4 ww = *reinterpret_cast<const float4*>(w + c); d.x *= ww.x; d.y *= ww.y; d.z *= ww.z; d.w *= ww.w; }
        g[i] = d; xh[i] = xv;
        s1 += d.x + d.y + d.z + d.w;
        s2 += d.x * xv.x + d.y * xv.y + d.z * xv.z + d.w * xv.w;
      } else { g[i] = make_float4(0, 0, 0, 0); xh[i] = g[i]; }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        float4 o = make_float4(rs * (g[i].x - s1 - xh[i].x * s2), rs * (g[i].y - s1 - xh[i].y * s2),
                               rs * (g[i].z - s1 - xh[i].z * s2), rs * (g[i].w - s1 - xh[i].w * s2));
        st4(dx + row * lddx + c, o);
      }
    }
  }
  if (HAS_PARAM) {
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = (i * 32 + lane) * 4;
      if (c < C) {
        atomicAdd(&sacc[c], dwv[i].x); atomicAdd(&sacc[c + 1], dwv[i].y); atomicAdd(&sacc[c + 2], dwv[i].z); atomicAdd(&sacc[c + 3], dwv[i].w);
        atomicAdd(&sacc[C + c], dbv[i].x); atomicAdd(&sacc[C + c + 1], dbv[i].y); atomicAdd(&sacc[C + c + 2], dbv[i].z); atomicAdd(&sacc[C + c + 3], dbv[i].w);
      }
    }
    __syncthreads();
    if (adw || adb) {          // single-kernel mode: this CTA's sums go straight into the (pre-zeroed or accumulating) gradients
      for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float* o = i < C ? (adw ? adw + i : nullptr) : (adb ? adb + (i - C) : nullptr);
        if (o) atomicAdd(o, sacc[i]);
      }
    } else {
      for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) partial[(size_t)blockIdx.x * 2 * C + i] = sacc[i];
    }
  }
}

// narrow rows (C <= 128): LPR lanes per row, ROWS rows in flight per lane group (see layernorm_fwd_narrow_kernel)
template <typename T, int LPR, int ROWS>
__global__ void __launch_bounds__(256) layernorm_bwd_narrow_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                                   const float* __restrict__ w, const float* __restrict__ mean,
                                                                   const float* __restrict__ rstd, T* __restrict__ dx,
                                                                   float* __restrict__ partial, long long M, int C, long long lddy,
                                                                   long long ldx, long long lddx, int x_is_hat, int rows_per_cta,
                                                                   float* __restrict__ adw, float* __restrict__ adb) {
  extern __shared__ float sacc[];  // [2][C]
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (partial) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
  }
  const int c = sub * 4;
  const bool cok = c < C;
  const float invC = 1.f / (float)C;
  float4 ww = make_float4(1.f, 1.f, 1.f, 1.f);
  if (w && cok) ww = *reinterpret_cast<const float4*>(w + c);
  float4 dwv = make_float4(0, 0, 0, 0), dbv = make_float4(0, 0, 0, 0);
  const long long rbeg = (long long)blockIdx.x * rows_per_cta;
  long long rend = rbeg + rows_per_cta;
  if (rend > M) rend = M;
  // the trip count must be warp-uniform: the shuffles below name all 32 lanes
  for (long long base = rbeg + (long long)wid * G * ROWS; base < rend; base += (long long)nw * G * ROWS) {
    const long long r0 = base + (long long)grp * ROWS;
    float4 g[ROWS], xh[ROWS];
    float rs[ROWS], s1[ROWS], s2[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = r0 + r;
      const bool ok = cok && row < rend;
      g[r] = ok ? ld4(dy + row * lddy + c) : make_float4(0, 0, 0, 0);
      xh[r] = ok ? ld4(x + row * ldx + c) : make_float4(0, 0, 0, 0);
      rs[r] = row < rend ? rstd[row] : 0.f;
      if (!x_is_hat && ok) {
        const float mu = mean[row];
        xh[r].x = (xh[r].x - mu) * rs[r]; xh[r].y = (xh[r].y - mu) * rs[r]; xh[r].z = (xh[r].z - mu) * rs[r]; xh[r].w = (xh[r].w - mu) * rs[r];
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      dwv.x += g[r].x * xh[r].x; dwv.y += g[r].y * xh[r].y; dwv.z += g[r].z * xh[r].z; dwv.w += g[r].w * xh[r].w;
      dbv.x += g[r].x; dbv.y += g[r].y; dbv.z += g[r].z; dbv.w += g[r].w;
      g[r].x *= ww.x; g[r].y *= ww.y; g[r].z *= ww.z; g[r].w *= ww.w;
      s1[r] = (g[r].x + g[r].y) + (g[r].z + g[r].w);
      s2[r] = (g[r].x * xh[r].x + g[r].y * xh[r].y) + (g[r].z * xh[r].z + g[r].w * xh[r].w);
    }
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o);
        s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
      }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = r0 + r;
      if (!cok || row >= rend) continue;
      const float m1 = s1[r] * invC, m2 = s2[r] * invC;
      st4(dx + row * lddx + c, make_float4(rs[r] * (g[r].x - m1 - xh[r].x * m2), rs[r] * (g[r].y - m1 - xh[r].y * m2),
                                           rs[r] * (g[r].z - m1 - xh[r].z * m2), rs[r] * (g[r].w - m1 - xh[r].w * m2)));
    }
  }
  if (partial) {
    if (cok) {
      atomicAdd(&sacc[c], dwv.x); atomicAdd(&sacc[c + 1], dwv.y); atomicAdd(&sacc[c + 2], dwv.z); atomicAdd(&sacc[c + 3], dwv.w);
      atomicAdd(&sacc[C + c], dbv.x); atomicAdd(&sacc[C + c + 1], dbv.y); atomicAdd(&sacc[C + c + 2], dbv.z); atomicAdd(&sacc[C + c + 3], dbv.w);
    }
    __syncthreads();
    if (adw || adb) {          // single-kernel mode: this CTA's sums go straight into the (pre-zeroed or accumulating) gradients
      for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
        float* o = i < C ? (adw ? adw + i : nullptr) : (adb ? adb + (i - C) : nullptr);
        if (o) atomicAdd(o, sacc[i]);
      }
    } else {
      for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) partial[(size_t)blockIdx.x * 2 * C + i] = sacc[i];
    }
  }
}

// out0[j] += sum_p partial[p][j] (j < n0), out1[j-n0] += ... (j >= n0)
// block = 32 columns x 8 part-slices (256 threads); launch with grid ((n + 31) / 32)
__global__ void __launch_bounds__(256) reduce_parts2_kernel(const float* __restrict__ partial, int nparts, int n,
                                                            float* __restrict__ out0, int n0, float* __restrict__ out1,
                                                            int accumulate) {
  __shared__ float sh[8][33];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (j < n) {
    int p = slice;
    for (; p + 24 < nparts; p += 32)
      s += (partial[(size_t)p * n + j] + partial[(size_t)(p + 8) * n + j]) + (partial[(size_t)(p + 16) * n + j] + partial[(size_t)(p + 24) * n + j]);
    for (; p < nparts; p += 8) s += partial[(size_t)p * n + j];
  }
  sh[slice][lane] = s;
  __syncthreads();
  if (slice == 0 && j < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sh[k][lane];
    float* o = (j < n0) ? (out0 ? out0 + j : nullptr) : (out1 ? out1 + (j - n0) : nullptr);
    if (o) *o = accumulate ? (*o + t) : t;
  }
}

static inline int row_parts(long long M) {
  long long p = (M + 63) / 64;  // >= 64 rows per CTA ...
  const long long cap = 8LL * 148;
  if (p > cap) p = cap;
  if (p < 148) {                // ... unless that leaves most SMs idle: the class-token rows of the heads ([256, 688]: 4 CTAs,
    p = (M + 7) / 8;            // 36 us for 700 KB) get one row per warp instead
    if (p > 148) p = 148;
  }
  if (p < 1) p = 1;
  return (int)p;
}
extern "C" int ga_layernorm_bwd_parts(long long M, int C) { return row_parts(M); }

extern "C" int ga_layernorm_bwd(const void* dy, const void* x, const float* w, const float* mean, const float* rstd, void* dx,
                                float* dw, float* db, float* partial, long long M, int C, long long lddy, long long ldx,
                                long long lddx, int dtype, ga_stream_t s) {
  GA_REQUIRE(dy && x && rstd && dx && (C & 3) == 0 && (lddy & 3) == 0 && (ldx & 3) == 0 && (lddx & 3) == 0, GA_ERR_ALIGN,
             "ga_layernorm_bwd: bad arguments (C=%d)", C);
  GA_REQUIRE(C <= 2048, GA_ERR_UNSUPPORTED, "ga_layernorm_bwd: C=%d > 2048", C);
  if (M == 0) return GA_OK;
  const bool want_param = (dw || db);
  // partial == NULL: single-kernel mode, every CTA adds its sums to dw / db with fp32 atomics (no second launch; the order
  // of the additions, hence the last bits, varies from run to run).  With the workspace the reduction is a second,
  // deterministic kernel.  Either way dw / db are ACCUMULATED onto.
  const bool atomic_out = want_param && !partial;
  float* const adw = atomic_out ? dw : nullptr;
  float* const adb = atomic_out ? db : nullptr;
  if (atomic_out) partial = dw ? dw : db;          // non-null marker: the kernels test `partial` for "parameter gradients wanted"
  int parts = row_parts(M);
  if (!want_param) {                      // no partial buffers needed: one warp-row per slot, many CTAs in flight
    long long p = (M + 15) / 16;
    if (p > 32LL * 148) p = 32LL * 148;
    parts = (int)(p < 1 ? 1 : p);
  }
  const int rows_per_cta = (int)((M + parts - 1) / parts);
  const size_t smem = want_param ? (size_t)2 * C * sizeof(float) : 0;
  const int x_is_hat = (mean == nullptr);
#define GA_LNB_WIDE(MV)                                                                                                        \
  do {                                                                                                                         \
    if (want_param)                                                                                                            \
      layernorm_bwd_kernel<T, MV, true><<<parts, 256, smem, (cudaStream_t)s>>>((const T*)dy, (const T*)x, w, mean, rstd, (T*)dx, partial, M, C, \
                                                                                lddy, ldx, lddx, x_is_hat, rows_per_cta, adw, adb);           \
    else                                                                                                                       \
      layernorm_bwd_kernel<T, MV, false><<<parts, 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)x, w, mean, rstd, (T*)dx, nullptr, M, C,  \
                                                                              lddy, ldx, lddx, x_is_hat, rows_per_cta, nullptr, nullptr);     \
  } while (0)
#define GA_LNB_NARROW(LPR) layernorm_bwd_narrow_kernel<T, LPR, 4><<<parts, 256, smem, (cudaStream_t)s>>>((const T*)dy, (const T*)x, w, mean, rstd, (T*)dx, want_param ? partial : nullptr, M, C, lddy, ldx, lddx, x_is_hat, rows_per_cta, adw, adb)
  DISPATCH_T(dtype, {
    if (C <= 32) GA_LNB_NARROW(8);
    else if (C <= 64) GA_LNB_NARROW(16);
    else if (C <= 128) GA_LNB_NARROW(32);
    else if (C <= 512) GA_LNB_WIDE(4);
    else if (C <= 1024) GA_LNB_WIDE(8);
    else GA_LNB_WIDE(16);
  });
#undef GA_LNB_NARROW
#undef GA_LNB_WIDE
  int rc = launch_ok("layernorm_bwd");
  if (rc || !want_param || atomic_out) return rc;
  reduce_parts2_kernel<<<(2 * C + 31) / 32, 256, 0, (cudaStream_t)s>>>(partial, parts, 2 * C, dw, C, db, 1);
  return launch_ok("layernorm_bwd_reduce");
}

// LN backward with xhat saved and no parameters: the hot form inside the ConvNeXt block.  VPL float4 vectors per lane
// cover C <= 128*VPL channels; ROWS rows are loaded before any reduction so each warp keeps 2*ROWS*VPL 128-bit loads in flight.
// Optional residual: out = res + LN'(dxhat) in the residual dtype TO (the transformer-block form, where the stream gradient
// bypasses the norm), plus an optional compute-dtype shadow of the sum.
template <typename T, typename TO, int VPL, int ROWS>
__global__ void __launch_bounds__(256) ln_bwd_hat_kernel(const T* __restrict__ dxhat, const T* __restrict__ xhat,
                                                         const float* __restrict__ rstd, TO* __restrict__ dconv,
                                                         const TO* __restrict__ res, T* __restrict__ shadow, long long M, int C) {
  const int lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float invC = 1.f / (float)C;
  for (long long row0 = gw * ROWS; row0 < M; row0 += nwarps * ROWS) {
    float4 g[ROWS][VPL], xh[ROWS][VPL];
    float s1[ROWS], s2[ROWS], rs[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r;
      rs[r] = row < M ? rstd[row] : 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = (v * 32 + lane) * 4;
        if (row < M && c < C) { g[r][v] = ld4(dxhat + row * C + c); xh[r][v] = ld4(xhat + row * C + c); }
        else { g[r][v] = make_float4(0, 0, 0, 0); xh[r][v] = g[r][v]; }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        a += (g[r][v].x + g[r][v].y) + (g[r][v].z + g[r][v].w);
        b += (g[r][v].x * xh[r][v].x + g[r][v].y * xh[r][v].y) + (g[r][v].z * xh[r][v].z + g[r][v].w * xh[r][v].w);
      }
      s1[r] = a; s2[r] = b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        s1[r] += __shfl_xor_sync(0xffffffffu, s1[r], o);
        s2[r] += __shfl_xor_sync(0xffffffffu, s2[r], o);
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const long long row = row0 + r;
      if (row >= M) continue;
      const float m1 = s1[r] * invC, m2 = s2[r] * invC;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        const int c = (v * 32 + lane) * 4;
        if (c < C) {
          float4 o = make_float4(rs[r] * (g[r][v].x - m1 - xh[r][v].x * m2), rs[r] * (g[r][v].y - m1 - xh[r][v].y * m2),
                                 rs[r] * (g[r][v].z - m1 - xh[r][v].z * m2), rs[r] * (g[r][v].w - m1 - xh[r][v].w * m2));
          if (res) {
            const float4 rr = ld4(res + row * C + c);
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          st4(dconv + row * C + c, o);
          if (shadow) st4(shadow + row * C + c, o);
        }
      }
    }
  }
}

extern "C" int ga_ln_bwd_rows(const void* dxhat, const void* xhat, const float* rstd, void* dconv, long long M, int C, int dtype,
                              ga_stream_t s) {
  if (C > 512 || (C & 3))
    return ga_layernorm_bwd(dxhat, xhat, nullptr, nullptr, rstd, dconv, nullptr, nullptr, nullptr, M, C, C, C, C, dtype, s);
  GA_REQUIRE(dxhat && xhat && rstd && dconv, GA_ERR_SHAPE, "ga_ln_bwd_rows: null argument");
  if (M == 0) return GA_OK;
  const int rows = C <= 128 ? 4 : (C <= 256 ? 2 : 1);
  long long blocks = (M + 8LL * rows - 1) / (8LL * rows);
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  cudaStream_t st = (cudaStream_t)s;
  DISPATCH_T(dtype, {
    if (C <= 128) ln_bwd_hat_kernel<T, T, 1, 4><<<(unsigned)blocks, 256, 0, st>>>((const T*)dxhat, (const T*)xhat, rstd, (T*)dconv, nullptr, nullptr, M, C);
    else if (C <= 256) ln_bwd_hat_kernel<T, T, 2, 2><<<(unsigned)blocks, 256, 0, st>>>((const T*)dxhat, (const T*)xhat, rstd, (T*)dconv, nullptr, nullptr, M, C);
    else ln_bwd_hat_kernel<T, T, 4, 1><<<(unsigned)blocks, 256, 0, st>>>((const T*)dxhat, (const T*)xhat, rstd, (T*)dconv, nullptr, nullptr, M, C);
  });
  return launch_ok("ln_bwd_hat");
}

extern "C" int ga_ln_bwd_rows_res(const void* dxhat, const void* xhat, const float* rstd, const void* res, void* out, void* shadow,
                                  long long M, int C, int dtype, int res_dtype, ga_stream_t s) {
  GA_REQUIRE(dxhat && xhat && rstd && res && out, GA_ERR_SHAPE, "ga_ln_bwd_rows_res: null argument");
  GA_REQUIRE(C <= 512 && (C & 3) == 0, GA_ERR_UNSUPPORTED, "ga_ln_bwd_rows_res: C=%d (multiple of 4, <= 512)", C);
  GA_REQUIRE(res_dtype == GA_F32 || res_dtype == dtype, GA_ERR_UNSUPPORTED, "ga_ln_bwd_rows_res: residual is fp32 or the compute dtype");
  if (M == 0) return GA_OK;
  const int rows = C <= 128 ? 4 : (C <= 256 ? 2 : 1);
  long long blocks = (M + 8LL * rows - 1) / (8LL * rows);
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  cudaStream_t st = (cudaStream_t)s;
#define GA_LNR(T, TO)                                                                                                          \
  {                                                                                                                            \
    if (C <= 128) ln_bwd_hat_kernel<T, TO, 1, 4><<<(unsigned)blocks, 256, 0, st>>>((const T*)dxhat, (const T*)xhat, rstd, (TO*)out, (const TO*)res, (T*)shadow, M, C); \
    else if (C <= 256) ln_bwd_hat_kernel<T, TO, 2, 2><<<(unsigned)blocks, 256, 0, st>>>((const T*)dxhat, (const T*)xhat, rstd, (TO*)out, (const TO*)res, (T*)shadow, M, C); \
    else ln_bwd_hat_kernel<T, TO, 4, 1><<<(unsigned)blocks, 256, 0, st>>>((const T*)dxhat, (const T*)xhat, rstd, (TO*)out, (const TO*)res, (T*)shadow, M, C); \
  }
  if (dtype == GA_BF16 && res_dtype == GA_F32) GA_LNR(bf16, float)
  else if (dtype == GA_BF16) GA_LNR(bf16, bf16)
  else GA_LNR(float, float)
#undef GA_LNR
  return launch_ok("ln_bwd_hat_res");
}

// ---------------------------------------------------------------------------------------------- patch gathers
// k == stride patches of an NHWC tensor: out row (b, oy, ox), column ((ky*k + kx)*C + c).  inverse: scatter back.
template <typename T>
__global__ void patchify_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int k, int inverse) {
  const int C4 = C >> 2;
  const long long total = (long long)B * H * W * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    long long p = i / C4;
    const int xx = (int)(p % W); p /= W;
    const int yy = (int)(p % H);
    const int b = (int)(p / H);
    const int oy = yy / k, ky = yy - oy * k, ox = xx / k, kx = xx - ox * k;
    const long long row = ((long long)b * (H / k) + oy) * (W / k) + ox;
    const long long po = row * ((long long)k * k * C) + (long long)(ky * k + kx) * C + c4 * 4;
    const long long xo = i * 4;
    if (!inverse) st4(y + po, ld4(x + xo)); else st4(y + xo, ld4(x + po));
  }
}
// the same gather with 16-byte accesses (bf16, C % 8 == 0): half the threads and half the index arithmetic per byte moved
__global__ void patchify8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W, int C8, int k, int inverse) {
  const long long total = (long long)B * H * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long p = i / C8;
    const int xx = (int)(p % W); p /= W;
    const int yy = (int)(p % H);
    const int b = (int)(p / H);
    const int oy = yy / k, ky = yy - oy * k, ox = xx / k, kx = xx - ox * k;
    const long long row = ((long long)b * (H / k) + oy) * (W / k) + ox;
    const long long po = row * ((long long)k * k * C8) + (long long)(ky * k + kx) * C8 + c8;
    if (!inverse) y[po] = x[i]; else y[i] = x[po];
  }
}
extern "C" int ga_patchify(const void* x, void* y, int B, int H, int W, int C, int k, int inverse, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && y && (C & 3) == 0 && H % k == 0 && W % k == 0, GA_ERR_SHAPE, "ga_patchify: bad shape H=%d W=%d C=%d k=%d", H, W, C, k);
  const long long total = (long long)B * H * W * (C >> 2);
  if (total == 0) return GA_OK;
  if (dtype == GA_BF16 && (C & 7) == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
    const long long t8 = total >> 1;
    const int grid8 = (int)((t8 + 255) / 256 > 148 * 16 ? 148 * 16 : (t8 + 255) / 256);
    patchify8_kernel<<<grid8, 256, 0, (cudaStream_t)s>>>((const uint4*)x, (uint4*)y, B, H, W, C >> 3, k, inverse);
    return launch_ok("patchify8");
  }
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  DISPATCH_T(dtype, { patchify_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, B, H, W, C, k, inverse); });
  return launch_ok("patchify");
}

// stem: fp32 image [B,3,H,W] with arbitrary element strides (NCHW or channels_last storage) -> rows (b, oy, ox),
// columns (ky, kx, c)
template <typename T>
__global__ void stem_patchify_kernel(const float* __restrict__ x, T* __restrict__ y, int B, int H, int W, int k, long long sb,
                                     long long sc, long long sy, long long sx) {
  const int Ho = H / k, Wo = W / k, KK = k * k * 3;
  const long long total = (long long)B * Ho * Wo * KK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % KK);
    long long p = i / KK;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const int c = col % 3, kx = (col / 3) % k, ky = col / (3 * k);
    st_f(y + i, x[b * sb + c * sc + (long long)(oy * k + ky) * sy + (long long)(ox * k + kx) * sx]);
  }
}
// k = 4 on an image whose rows are contiguous (NCHW storage): one thread per output row.  The 3 x 4 source segments of a patch
// are 16 contiguous bytes each and adjacent for adjacent patches (12 coalesced 128-bit loads per thread); the 48 outputs of a
// row are contiguous (bf16: six 128-bit stores).  The element-per-thread kernel above reads 4 bytes per thread from three
// channel planes at once: 0.9 TB/s at 224^2 (profiles/r02_step_profile_torchprofiler.txt), this form streams.
template <typename T>
__global__ void __launch_bounds__(128) stem_patchify4_kernel(const float* __restrict__ x, T* __restrict__ y, int B, int H, int W, long long sb,
                                                             long long sc, long long sy) {
  const int Ho = H >> 2, Wo = W >> 2;
  const long long rows = (long long)B * Ho * Wo;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(r % Wo);
    long long p = r / Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const float* src = x + b * sb + (long long)(oy * 4) * sy + ox * 4;
    float v[48];                                   // column (ky * 4 + kx) * 3 + c
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ky = 0; ky < 4; ++ky) {
        const float4 q = *reinterpret_cast<const float4*>(src + c * sc + ky * sy);
        v[(ky * 4 + 0) * 3 + c] = q.x; v[(ky * 4 + 1) * 3 + c] = q.y; v[(ky * 4 + 2) * 3 + c] = q.z; v[(ky * 4 + 3) * 3 + c] = q.w;
      }
    T* dst = y + r * 48;
#pragma unroll
    for (int j = 0; j < 48; j += 4) st4(dst + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
  }
}

extern "C" int ga_stem_patchify(const float* x, void* y, int B, int H, int W, int k, long long sb, long long sc, long long sy,
                                long long sx, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && y && H % k == 0 && W % k == 0, GA_ERR_SHAPE, "ga_stem_patchify: bad shape");
  const long long total = (long long)B * (H / k) * (W / k) * k * k * 3;
  if (total == 0) return GA_OK;
  if (k == 4 && sx == 1 && ((uintptr_t)x & 15) == 0 && ((sb | sc | sy) & 3) == 0 && ((uintptr_t)y & 15) == 0) {
    const long long rows = total / 48;
    const int grid4 = (int)((rows + 127) / 128 > 148 * 32 ? 148 * 32 : (rows + 127) / 128);
    DISPATCH_T(dtype, { stem_patchify4_kernel<T><<<grid4, 128, 0, (cudaStream_t)s>>>(x, (T*)y, B, H, W, sb, sc, sy); });
    return launch_ok("stem_patchify4");
  }
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  DISPATCH_T(dtype, { stem_patchify_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>(x, (T*)y, B, H, W, k, sb, sc, sy, sx); });
  return launch_ok("stem_patchify");
}

// 3x3, pad 1, stride S im2col (forward) and its adjoint in gather form (inverse): C % 4 == 0.  Output map Ho x Wo with
// Ho = (H - 1) / S + 1 (S = 1: the Bottleneck conv; S = 2: CSWin's Merge_Block and deep stem, ga_cswin.py:256, 472)
template <typename T>
__global__ void im2col3_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, long long ldx,
                               long long ldy, int inverse, int S) {
  const int C4 = C >> 2;
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
  if (S != 1) {
    if (!inverse) {
      const long long total = (long long)B * Ho * Wo * 9 * C4;
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long p = i / C4;
        const int tap = (int)(p % 9); p /= 9;
        const int ox = (int)(p % Wo); p /= Wo;
        const int oy = (int)(p % Ho);
        const int b = (int)(p / Ho);
        const int sy = oy * S + tap / 3 - 1, sx = ox * S + tap % 3 - 1;
        float4 v = make_float4(0, 0, 0, 0);
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = ld4(x + (((long long)b * H + sy) * W + sx) * ldx + c4 * 4);
        st4(y + (((long long)b * Ho + oy) * Wo + ox) * ldy + tap * C + c4 * 4, v);
      }
    } else {
      // x = d(col) [B*Ho*Wo, ldx>=9C]; y = d(image) [B*H*W, ldy]: pixel (yy,xx) gathers tap (ky,kx) from output (oy,ox)
      // whenever oy*S + ky - 1 == yy and ox*S + kx - 1 == xx
      const long long total = (long long)B * H * W * C4;
      for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long p = i / C4;
        const int xx = (int)(p % W); p /= W;
        const int yy = (int)(p % H);
        const int b = (int)(p / H);
        float4 a = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int ty = yy + 1 - tap / 3, tx = xx + 1 - tap % 3;
          if (ty < 0 || tx < 0 || ty % S || tx % S) continue;
          const int oy = ty / S, ox = tx / S;
          if (oy >= Ho || ox >= Wo) continue;
          float4 v = ld4(x + (((long long)b * Ho + oy) * Wo + ox) * ldx + tap * C + c4 * 4);
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        st4(y + (((long long)b * H + yy) * W + xx) * ldy + c4 * 4, a);
      }
    }
    return;
  }
  if (!inverse) {
    const long long total = (long long)B * H * W * 9 * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int c4 = (int)(i % C4);
      long long p = i / C4;
      const int tap = (int)(p % 9); p /= 9;
      const int xx = (int)(p % W); p /= W;
      const int yy = (int)(p % H);
      const int b = (int)(p / H);
      const int sy = yy + tap / 3 - 1, sx = xx + tap % 3 - 1;
      float4 v = make_float4(0, 0, 0, 0);
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = ld4(x + (((long long)b * H + sy) * W + sx) * ldx + c4 * 4);
      st4(y + (((long long)b * H + yy) * W + xx) * ldy + tap * C + c4 * 4, v);
    }
  } else {
    // x = d(col) [B*H*W, ldx>=9C] ; y = d(image) [B*H*W, ldy]: pixel (yy,xx) gathers tap t from output pixel (yy - dy, xx - dx)
    const long long total = (long long)B * H * W * C4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const int c4 = (int)(i % C4);
      long long p = i / C4;
      const int xx = (int)(p % W); p /= W;
      const int yy = (int)(p % H);
      const int b = (int)(p / H);
      float4 a = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int oy = yy - (tap / 3 - 1), ox = xx - (tap % 3 - 1);
        if (oy >= 0 && oy < H && ox >= 0 && ox < W) {
          float4 v = ld4(x + (((long long)b * H + oy) * W + ox) * ldx + tap * C + c4 * 4);
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
      }
      st4(y + (((long long)b * H + yy) * W + xx) * ldy + c4 * 4, a);
    }
  }
}
// 16-byte-chunk forms (C * sizeof(T) and both pitches multiples of 16 bytes): the forward gather is pure data movement
__global__ void __launch_bounds__(256) im2col3_fwd16_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W,
                                                            int CV, long long ldx, long long ldy, int S) {
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
  const long long total = (long long)B * Ho * Wo * 9 * CV;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % CV);
    long long p = i / CV;
    const int tap = (int)(p % 9); p /= 9;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const int sy = oy * S + tap / 3 - 1, sx = ox * S + tap % 3 - 1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = x[(((long long)b * H + sy) * W + sx) * ldx + cv];
    y[(((long long)b * Ho + oy) * Wo + ox) * ldy + tap * CV + cv] = v;
  }
}
__global__ void __launch_bounds__(256) im2col3_inv8_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C,
                                                           long long ldx, long long ldy, int S) {
  const int C8 = C >> 3;
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
  const long long total = (long long)B * H * W * C8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    long long p = i / C8;
    const int xx = (int)(p % W); p /= W;
    const int yy = (int)(p % H);
    const int b = (int)(p / H);
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ty = yy + 1 - tap / 3, tx = xx + 1 - tap % 3;
      if (ty < 0 || tx < 0 || ty % S || tx % S) continue;
      const int oy = ty / S, ox = tx / S;
      if (oy >= Ho || ox >= Wo) continue;
      float v[8];
      ld8_bf16(x + (((long long)b * Ho + oy) * Wo + ox) * ldx + tap * C + c8 * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] += v[e];
    }
    st8_bf16(y + (((long long)b * H + yy) * W + xx) * ldy + c8 * 8, a);
  }
}

extern "C" int ga_im2col3s(const void* x, void* y, int B, int H, int W, int C, int stride, long long ldx, long long ldy, int inverse,
                           int dtype, ga_stream_t s) {
  GA_REQUIRE(x && y && (C & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0, GA_ERR_ALIGN, "ga_im2col3: C/ld must be multiples of 4");
  GA_REQUIRE(stride == 1 || stride == 2, GA_ERR_UNSUPPORTED, "ga_im2col3: stride %d (1 or 2)", stride);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int es = dtype == GA_BF16 ? 2 : 4, per16 = 16 / es;
  const bool v16 = (C % per16 == 0) && (ldx % per16 == 0) && (ldy % per16 == 0) && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  if (v16 && !inverse) {
    const int CV = C / per16;
    const long long total = (long long)B * Ho * Wo * 9 * CV;
    if (total == 0) return GA_OK;
    const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    im2col3_fwd16_kernel<<<grid, 256, 0, (cudaStream_t)s>>>((const uint4*)x, (uint4*)y, B, H, W, CV, ldx / per16, ldy / per16, stride);
    return launch_ok("im2col3_fwd16");
  }
  if (v16 && inverse && dtype == GA_BF16) {
    const long long total = (long long)B * H * W * (C >> 3);
    if (total == 0) return GA_OK;
    const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    im2col3_inv8_kernel<<<grid, 256, 0, (cudaStream_t)s>>>((const bf16*)x, (bf16*)y, B, H, W, C, ldx, ldy, stride);
    return launch_ok("im2col3_inv8");
  }
  const long long total = inverse ? (long long)B * H * W * (C >> 2) : (long long)B * Ho * Wo * (C >> 2) * 9;
  if (total == 0) return GA_OK;
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  DISPATCH_T(dtype, { im2col3_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (T*)y, B, H, W, C, ldx, ldy, inverse, stride); });
  return launch_ok("im2col3");
}
extern "C" int ga_im2col3(const void* x, void* y, int B, int H, int W, int C, long long ldx, long long ldy, int inverse,
                          int dtype, ga_stream_t s) {
  return ga_im2col3s(x, y, B, H, W, C, 1, ldx, ldy, inverse, dtype, s);
}

// CSWin deep-stem first conv (3x3, stride S, pad 1, ga_cswin.py:463): fp32 image [B,3,H,W] with element strides ->
// rows (b, oy, ox) of 32 columns: (ky, kx, c) in the first 27, zeros in the pad (16-byte pitch for the GEMM operand)
template <typename T>
__global__ void stem_im2col3_kernel(const float* __restrict__ x, T* __restrict__ y, int B, int H, int W, int S, long long sb,
                                    long long sc, long long sy, long long sx) {
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
  const long long total = (long long)B * Ho * Wo * 32;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i & 31);
    long long p = i >> 5;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float v = 0.f;
    if (col < 27) {
      const int c = col % 3, tap = col / 3;
      const int yy = oy * S + tap / 3 - 1, xx = ox * S + tap % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = x[b * sb + c * sc + yy * sy + xx * sx];
    }
    st_f(y + i, v);
  }
}
extern "C" int ga_stem_im2col3(const float* x, void* y, int B, int H, int W, int stride, long long sb, long long sc, long long sy,
                               long long sx, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && y && stride >= 1, GA_ERR_SHAPE, "ga_stem_im2col3: bad arguments");
  const long long total = (long long)B * ((H - 1) / stride + 1) * ((W - 1) / stride + 1) * 32;
  if (total == 0) return GA_OK;
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  DISPATCH_T(dtype, { stem_im2col3_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>(x, (T*)y, B, H, W, stride, sb, sc, sy, sx); });
  return launch_ok("stem_im2col3");
}

// ---------------------------------------------------------------------------------------------- column statistics
// CTA (strip of 128 columns as 32 lanes x 4, row slice): 8 warps stride rows; result -> partial[slice][2][C]
// MODE 0: s1 = sum x, s2 = sum x^2
// MODE 1: BN backward: d = dy * (relu ? y>0 : 1); s1 = sum d; s2 = sum d * (x-mean)*invstd
// MODE 2: s1 = sum (x - p), s2 = sum (x - p)^2 with the pivot p = row 0 of x (also written to pivot_out): BatchNorm statistics
//         without the cancellation of E[x^2] - E[x]^2 (a batch whose rows differ by 5 % of their magnitude loses 4e-5 there)
template <typename T, int MODE>
__global__ void __launch_bounds__(256) colstats_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ yact,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                       float* __restrict__ partial, long long M, int C, long long ldx,
                                                       long long lddy, long long ldy, int rows_per_cta, int relu,
                                                       float* __restrict__ pivot_out = nullptr, float* __restrict__ out0 = nullptr,
                                                       float* __restrict__ out1 = nullptr) {
  __shared__ float4 sh[2][8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + lane * 4;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  float4 a = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
  if (c < C) {
    float4 mu = make_float4(0, 0, 0, 0), is = make_float4(1, 1, 1, 1);
    if (MODE == 1) { mu = *reinterpret_cast<const float4*>(mean + c); is = *reinterpret_cast<const float4*>(invstd + c); }
    if (MODE == 2) {
      mu = ld4(x + c);
      if (blockIdx.y == 0 && wid == 0) *reinterpret_cast<float4*>(pivot_out + c) = mu;
    }
    long long rend = r0 + rows_per_cta;
    if (rend > M) rend = M;
    long long r = r0 + wid;
    if (MODE == 0) {
      for (; r + 24 < rend; r += 32) {     // four independent row loads in flight
        float4 v0 = ld4(x + r * ldx + c), v1 = ld4(x + (r + 8) * ldx + c), v2 = ld4(x + (r + 16) * ldx + c), v3 = ld4(x + (r + 24) * ldx + c);
        a.x += (v0.x + v1.x) + (v2.x + v3.x); a.y += (v0.y + v1.y) + (v2.y + v3.y);
        a.z += (v0.z + v1.z) + (v2.z + v3.z); a.w += (v0.w + v1.w) + (v2.w + v3.w);
        q.x += (v0.x * v0.x + v1.x * v1.x) + (v2.x * v2.x + v3.x * v3.x); q.y += (v0.y * v0.y + v1.y * v1.y) + (v2.y * v2.y + v3.y * v3.y);
        q.z += (v0.z * v0.z + v1.z * v1.z) + (v2.z * v2.z + v3.z * v3.z); q.w += (v0.w * v0.w + v1.w * v1.w) + (v2.w * v2.w + v3.w * v3.w);
      }
    }
    for (; r < rend; r += 8) {
      float4 v = ld4(x + r * ldx + c);
      if (MODE == 0) {
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
      } else if (MODE == 2) {
        v.x -= mu.x; v.y -= mu.y; v.z -= mu.z; v.w -= mu.w;
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w;
      } else {
        float4 d = ld4(dy + r * lddy + c);
        if (relu) {
          float4 yy = ld4(yact + r * ldy + c);
          d.x = yy.x > 0.f ? d.x : 0.f; d.y = yy.y > 0.f ? d.y : 0.f; d.z = yy.z > 0.f ? d.z : 0.f; d.w = yy.w > 0.f ? d.w : 0.f;
        }
        a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
        q.x += d.x * (v.x - mu.x) * is.x; q.y += d.y * (v.y - mu.y) * is.y; q.z += d.z * (v.z - mu.z) * is.z; q.w += d.w * (v.w - mu.w) * is.w;
      }
    }
  }
  sh[0][wid][lane] = a; sh[1][wid][lane] = q;
  __syncthreads();
  if (wid < 2 && c < C) {
    float4 t = sh[wid][0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) { float4 u = sh[wid][k][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    if (partial) {
      *reinterpret_cast<float4*>(partial + ((size_t)blockIdx.y * 2 + wid) * C + c) = t;
    } else {                   // single-kernel mode: add to the (pre-zeroed or accumulating) outputs
      float* o = wid == 0 ? out0 : out1;
      if (o) atomicAdd(reinterpret_cast<float4*>(o + c), t);
    }
  }
}

extern "C" int ga_colstats_parts(long long M, int C) {
  const int strips = (C + 127) / 128;
  long long p = (8LL * 148 + strips - 1) / strips;
  const long long maxp = (M + 63) / 64;
  if (p > maxp) p = maxp;
  if (p < 1) p = 1;
  return (int)p;
}

extern "C" int ga_colstats(const void* x, float* sum, float* sumsq, float* partial, long long M, int C, long long ldx,
                           int accumulate, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && (C & 3) == 0 && (ldx & 3) == 0, GA_ERR_ALIGN, "ga_colstats: C=%d ldx=%lld must be multiples of 4", C, ldx);
  GA_REQUIRE(partial || accumulate, GA_ERR_SHAPE, "ga_colstats: without the workspace the sums are added atomically: pass accumulate = 1 and zeroed outputs");
  const int parts = ga_colstats_parts(M, C);
  const int rows_per_cta = (int)((M + parts - 1) / parts);
  dim3 grid((C + 127) / 128, parts);
  DISPATCH_T(dtype, { colstats_kernel<T, 0><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, nullptr, nullptr, nullptr, nullptr, partial, M, C, ldx, 0, 0, rows_per_cta, 0, nullptr, sum, sumsq); });
  int rc = launch_ok("colstats");
  if (rc || !partial) return rc;
  reduce_parts2_kernel<<<(2 * C + 31) / 32, 256, 0, (cudaStream_t)s>>>(partial, parts, 2 * C, sum, C, sumsq, accumulate);
  return launch_ok("colstats_reduce");
}

// sums of (x - pivot) and (x - pivot)^2 with pivot = row 0 (pivot[C] written), for ga_bn_finalize(pivot != NULL)
extern "C" int ga_colstats_shifted(const void* x, float* pivot, float* sum, float* sumsq, float* partial, long long M, int C,
                                   long long ldx, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && pivot && sum && sumsq && M > 0 && (C & 3) == 0 && (ldx & 3) == 0, GA_ERR_ALIGN,
             "ga_colstats_shifted: C=%d ldx=%lld must be multiples of 4", C, ldx);
  const int parts = ga_colstats_parts(M, C);
  const int rows_per_cta = (int)((M + parts - 1) / parts);
  dim3 grid((C + 127) / 128, parts);
  DISPATCH_T(dtype, { colstats_kernel<T, 2><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, nullptr, nullptr, nullptr, nullptr, partial, M, C, ldx, 0, 0, rows_per_cta, 0, pivot, sum, sumsq); });
  int rc = launch_ok("colstats_shifted");
  if (rc || !partial) return rc;          // partial == NULL: sum / sumsq were zero and now hold the atomically added sums
  reduce_parts2_kernel<<<(2 * C + 31) / 32, 256, 0, (cudaStream_t)s>>>(partial, parts, 2 * C, sum, C, sumsq, 0);
  return launch_ok("colstats_reduce");
}

// BatchNorm (training) finalize: from sum/sumsq over M rows -> mean, invstd, scale, shift; running stats (momentum, unbiased var)
__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, const float* __restrict__ pivot,
                                   const float* __restrict__ w,
                                   const float* __restrict__ b, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean_o, float* __restrict__ invstd_o, float* __restrict__ scale_o,
                                   float* __restrict__ shift_o, long long M, int C, float momentum, float eps, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, var;
  if (training) {
    const float m0 = sum[c] / (float)M;                         // mean of (x - pivot) when a pivot is given
    var = fmaxf(sumsq[c] / (float)M - m0 * m0, 0.f);
    mean = pivot ? pivot[c] + m0 : m0;
    if (running_mean) {
      const float unb = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const float is = rsqrtf(var + eps);
  const float sc = (w ? w[c] : 1.f) * is;
  if (mean_o) mean_o[c] = mean;
  if (invstd_o) invstd_o[c] = is;
  scale_o[c] = sc;
  shift_o[c] = (b ? b[c] : 0.f) - mean * sc;
}
extern "C" int ga_bn_finalize(const float* sum, const float* sumsq, const float* pivot, const float* w, const float* b, float* running_mean,
                              float* running_var, float* mean, float* invstd, float* scale, float* shift, long long M, int C,
                              float momentum, float eps, int training, ga_stream_t s) {
  GA_REQUIRE(scale && shift && C > 0, GA_ERR_SHAPE, "ga_bn_finalize: bad arguments");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)s>>>(sum, sumsq, pivot, w, b, running_mean, running_var, mean, invstd, scale, shift, M, C, momentum, eps, training);
  return launch_ok("bn_finalize");
}

// y = act( x*scale + shift  (+ x2*scale2 + shift2) )
template <typename T>
__global__ void affine_act_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                  const T* __restrict__ x2, const float* __restrict__ scale2, const float* __restrict__ shift2,
                                  T* __restrict__ y, long long M, int C, long long ldx, long long ldx2, long long ldy, int act) {
  const int C4 = C >> 2;
  const long long total = M * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long r = i / C4;
    float4 v = ld4(x + r * ldx + c);
    float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    v.x = v.x * sc.x + sh.x; v.y = v.y * sc.y + sh.y; v.z = v.z * sc.z + sh.z; v.w = v.w * sc.w + sh.w;
    if (x2) {
      float4 u = ld4(x2 + r * ldx2 + c);
      if (scale2) {
        float4 s2 = *reinterpret_cast<const float4*>(scale2 + c), h2 = *reinterpret_cast<const float4*>(shift2 + c);
        u.x = u.x * s2.x + h2.x; u.y = u.y * s2.y + h2.y; u.z = u.z * s2.z + h2.z; u.w = u.w * s2.w + h2.w;
      }
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (act == GA_ACT_RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    else if (act == GA_ACT_GELU) { v.x = gelu_f(v.x); v.y = gelu_f(v.y); v.z = gelu_f(v.z); v.w = gelu_f(v.w); }
    st4(y + r * ldy + c, v);
  }
}
extern "C" int ga_affine_act(const void* x, const float* scale, const float* shift, const void* x2, const float* scale2,
                             const float* shift2, void* y, long long M, int C, long long ldx, long long ldx2, long long ldy,
                             int act, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && y && scale && shift && (C & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0 && (ldx2 & 3) == 0, GA_ERR_ALIGN,
             "ga_affine_act: C/ld must be multiples of 4 (C=%d)", C);
  const long long total = M * (C >> 2);
  if (total == 0) return GA_OK;
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  DISPATCH_T(dtype, { affine_act_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, scale, shift, (const T*)x2, scale2, shift2, (T*)y, M, C, ldx, ldx2, ldy, act); });
  return launch_ok("affine_act");
}

// BN backward, pass 1: c1[c] = sum d, c2[c] = sum d*xhat  (d = dy masked by relu(y))
extern "C" int ga_bn_bwd_reduce(const void* dy, const void* x, const void* y, const float* mean, const float* invstd, float* c1,
                                float* c2, float* partial, long long M, int C, long long lddy, long long ldx, long long ldy,
                                int relu, int dtype, ga_stream_t s) {
  GA_REQUIRE(dy && x && mean && invstd && (C & 3) == 0 && (ldx & 3) == 0 && (lddy & 3) == 0, GA_ERR_ALIGN,
             "ga_bn_bwd_reduce: bad arguments");
  GA_REQUIRE(!relu || (y && (ldy & 3) == 0), GA_ERR_SHAPE, "ga_bn_bwd_reduce: relu mask needs y");
  const int parts = ga_colstats_parts(M, C);
  const int rows_per_cta = (int)((M + parts - 1) / parts);
  dim3 grid((C + 127) / 128, parts);
  DISPATCH_T(dtype, { colstats_kernel<T, 1><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, (const T*)dy, (const T*)y, mean, invstd, partial, M, C, ldx, lddy, ldy, rows_per_cta, relu, nullptr, c1, c2); });
  int rc = launch_ok("bn_bwd_reduce");
  if (rc || !partial) return rc;          // partial == NULL: c1 / c2 were zero and now hold the atomically added sums
  reduce_parts2_kernel<<<(2 * C + 31) / 32, 256, 0, (cudaStream_t)s>>>(partial, parts, 2 * C, c1, C, c2, 0);
  return launch_ok("bn_bwd_reduce2");
}

// BN backward, pass 2: dx = scale*(d - c1/M - xhat*c2/M)   (training) ; eval: dx = scale*d
template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ scale, const float* __restrict__ c1, const float* __restrict__ c2,
                                    T* __restrict__ dx, long long M, int C, long long lddy, long long ldx, long long ldy,
                                    long long lddx, int relu, float invM) {
  const int C4 = C >> 2;
  const long long total = M * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long r = i / C4;
    float4 d = ld4(dy + r * lddy + c);
    if (relu) {
      float4 yy = ld4(y + r * ldy + c);
      d.x = yy.x > 0.f ? d.x : 0.f; d.y = yy.y > 0.f ? d.y : 0.f; d.z = yy.z > 0.f ? d.z : 0.f; d.w = yy.w > 0.f ? d.w : 0.f;
    }
    float4 sc = *reinterpret_cast<const float4*>(scale + c);
    float4 o;
    if (c1) {
      float4 v = ld4(x + r * ldx + c);
      float4 mu = *reinterpret_cast<const float4*>(mean + c), is = *reinterpret_cast<const float4*>(invstd + c);
      float4 a = *reinterpret_cast<const float4*>(c1 + c), b = *reinterpret_cast<const float4*>(c2 + c);
      o.x = sc.x * (d.x - a.x * invM - (v.x - mu.x) * is.x * b.x * invM);
      o.y = sc.y * (d.y - a.y * invM - (v.y - mu.y) * is.y * b.y * invM);
      o.z = sc.z * (d.z - a.z * invM - (v.z - mu.z) * is.z * b.z * invM);
      o.w = sc.w * (d.w - a.w * invM - (v.w - mu.w) * is.w * b.w * invM);
    } else {
      o = make_float4(sc.x * d.x, sc.y * d.y, sc.z * d.z, sc.w * d.w);
    }
    st4(dx + r * lddx + c, o);
  }
}
extern "C" int ga_bn_bwd_apply(const void* dy, const void* x, const void* y, const float* mean, const float* invstd,
                               const float* scale, const float* c1, const float* c2, void* dx, long long M, int C,
                               long long lddy, long long ldx, long long ldy, long long lddx, int relu, int dtype, ga_stream_t s) {
  GA_REQUIRE(dy && dx && scale && (C & 3) == 0 && (lddy & 3) == 0 && (lddx & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0,
             GA_ERR_ALIGN, "ga_bn_bwd_apply: bad arguments");
  const long long total = M * (C >> 2);
  if (total == 0) return GA_OK;
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  DISPATCH_T(dtype, { bn_bwd_apply_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)x, (const T*)y, mean, invstd, scale, c1, c2, (T*)dx, M, C, lddy, ldx, ldy, lddx, relu, 1.f / (float)M); });
  return launch_ok("bn_bwd_apply");
}

// ---------------------------------------------------------------------------------------------- small element-wise
template <typename TS, typename TD>
__global__ void copy_cols_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long M, int C, long long lds, long long ldd) {
  const long long total = M * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C; const int c = (int)(i - r * C);
    st_f(dst + r * ldd + c, ld_f(src + r * lds + c));
  }
}
template <typename TS, typename TD>
__global__ void copy_cols4_kernel(const TS* __restrict__ src, TD* __restrict__ dst, long long M, int C4, long long lds, long long ldd) {
  const long long total = M * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C4; const int c = (int)(i - r * C4) * 4;
    st4(dst + r * ldd + c, ld4(src + r * lds + c));
  }
}
extern "C" int ga_copy_cols(const void* src, void* dst, long long M, int C, long long lds, long long ldd, int src_dtype,
                            int dst_dtype, ga_stream_t s) {
  const long long total = M * C;
  if (total == 0) return GA_OK;
  if ((C & 3) == 0 && (lds & 3) == 0 && (ldd & 3) == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const long long t4 = total >> 2;
    const int grid4 = (int)((t4 + 255) / 256 > 148 * 16 ? 148 * 16 : (t4 + 255) / 256);
    cudaStream_t st4s = (cudaStream_t)s;
    const int C4 = C >> 2;
    if (src_dtype == GA_F32 && dst_dtype == GA_F32) copy_cols4_kernel<float, float><<<grid4, 256, 0, st4s>>>((const float*)src, (float*)dst, M, C4, lds, ldd);
    else if (src_dtype == GA_F32) copy_cols4_kernel<float, bf16><<<grid4, 256, 0, st4s>>>((const float*)src, (bf16*)dst, M, C4, lds, ldd);
    else if (dst_dtype == GA_F32) copy_cols4_kernel<bf16, float><<<grid4, 256, 0, st4s>>>((const bf16*)src, (float*)dst, M, C4, lds, ldd);
    else copy_cols4_kernel<bf16, bf16><<<grid4, 256, 0, st4s>>>((const bf16*)src, (bf16*)dst, M, C4, lds, ldd);
    return launch_ok("copy_cols4");
  }
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  cudaStream_t st = (cudaStream_t)s;
  if (src_dtype == GA_F32 && dst_dtype == GA_F32) copy_cols_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, M, C, lds, ldd);
  else if (src_dtype == GA_F32) copy_cols_kernel<float, bf16><<<grid, 256, 0, st>>>((const float*)src, (bf16*)dst, M, C, lds, ldd);
  else if (dst_dtype == GA_F32) copy_cols_kernel<bf16, float><<<grid, 256, 0, st>>>((const bf16*)src, (float*)dst, M, C, lds, ldd);
  else copy_cols_kernel<bf16, bf16><<<grid, 256, 0, st>>>((const bf16*)src, (bf16*)dst, M, C, lds, ldd);
  return launch_ok("copy_cols");
}

// dst[r,c] = src[r,c] * rowscale[r] * colscale[c]
template <typename TD>
__global__ void scale_matrix_kernel(const float* __restrict__ src, const float* __restrict__ rs, const float* __restrict__ cs,
                                    TD* __restrict__ dst, int rows, int cols) {
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    float v = src[i];
    if (rs) v *= rs[r];
    if (cs) v *= cs[c];
    st_f(dst + i, v);
  }
}
extern "C" int ga_scale_matrix(const float* src, const float* rowscale, const float* colscale, void* dst, int rows, int cols,
                               int dst_dtype, ga_stream_t s) {
  const long long total = (long long)rows * cols;
  if (total == 0) return GA_OK;
  const int grid = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  if (dst_dtype == GA_BF16) scale_matrix_kernel<bf16><<<grid, 256, 0, (cudaStream_t)s>>>(src, rowscale, colscale, (bf16*)dst, rows, cols);
  else scale_matrix_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>(src, rowscale, colscale, (float*)dst, rows, cols);
  return launch_ok("scale_matrix");
}
// LayerNorm-affine fold of a Linear that follows the norm: Wf[n,k] = W[n,k] * g[k] (cast to the operand dtype) and
// bf[n] = b[n] + sum_k W[n,k] * beta[k].  One warp per output row: W is read once (was: scale_matrix + an M=1 GEMM).
template <typename TD>
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ g, const float* __restrict__ beta,
                                                      const float* __restrict__ b, TD* __restrict__ Wf, float* __restrict__ bf, int N,
                                                      int K, long long ldw) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N) return;
  const float* w = W + (long long)row * K;
  float acc = 0.f;
  if ((K & 3) == 0) {
    for (int k = lane * 4; k < K; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(w + k);
      const float4 gg = *reinterpret_cast<const float4*>(g + k);
      const float4 bb = *reinterpret_cast<const float4*>(beta + k);
      acc = fmaf(v.x, bb.x, fmaf(v.y, bb.y, fmaf(v.z, bb.z, fmaf(v.w, bb.w, acc))));
      st4(Wf + (long long)row * ldw + k, make_float4(v.x * gg.x, v.y * gg.y, v.z * gg.z, v.w * gg.w));
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float v = w[k];
      acc = fmaf(v, beta[k], acc);
      st_f(Wf + (long long)row * ldw + k, v * g[k]);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) bf[row] = acc + (b ? b[row] : 0.f);
}
extern "C" int ga_fold_ln(const float* W, const float* ln_w, const float* ln_b, const float* bias, void* Wf, float* bf, int N, int K,
                          long long ldw, int dst_dtype, ga_stream_t s) {
  GA_REQUIRE(W && ln_w && ln_b && Wf && bf && N > 0 && K > 0 && ldw >= K, GA_ERR_SHAPE, "ga_fold_ln: bad arguments");
  GA_REQUIRE((K & 3) || (ldw & 3) == 0, GA_ERR_ALIGN, "ga_fold_ln: ldw must be a multiple of 4 when K is");
  const int grid = (N + 7) / 8;
  if (dst_dtype == GA_BF16) fold_ln_kernel<bf16><<<grid, 256, 0, (cudaStream_t)s>>>(W, ln_w, ln_b, bias, (bf16*)Wf, bf, N, K, ldw);
  else fold_ln_kernel<float><<<grid, 256, 0, (cudaStream_t)s>>>(W, ln_w, ln_b, bias, (float*)Wf, bf, N, K, ldw);
  return launch_ok("fold_ln");
}
// Operand preparation of EVERY ConvNeXt block of a model in one launch (was: a transpose copy, fold_ln, a cast and a
// scale_matrix per block and step: 72 launches of a few microseconds each).  One warp per unit; a block with C channels
// has 4C units for fc1 rows (W1 diag(ln_w) -> bf16, b1 + W1 ln_b), C units for fc2 rows (bf16 W2 and gamma[c] W2) and 49
// units for the depthwise taps ([C,1,7,7] -> [49,C] fp32).  table: GA_BLOCK_PREP_WORDS int64 words per block (header).
__global__ void __launch_bounds__(256) block_weight_prep_kernel(const long long* __restrict__ table, int nblocks, int total_units) {
  const int lane = threadIdx.x & 31;
  const int unit = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (unit >= total_units) return;
  int b = 0;
  while (b + 1 < nblocks && unit >= (int)table[(long long)(b + 1) * GA_BLOCK_PREP_WORDS + 13]) ++b;
  const long long* t = table + (long long)b * GA_BLOCK_PREP_WORDS;
  const float* dw_w = (const float*)t[0];
  const float* ln_w = (const float*)t[1];
  const float* ln_b = (const float*)t[2];
  const float* w1 = (const float*)t[3];
  const float* b1 = (const float*)t[4];
  const float* w2 = (const float*)t[5];
  const float* gamma = (const float*)t[6];
  float* w49c = (float*)t[7];
  bf16* w1f = (bf16*)t[8];
  float* b1f = (float*)t[9];
  bf16* w2c = (bf16*)t[10];
  bf16* w2s = (bf16*)t[11];
  const int C = (int)t[12], H = 4 * C;
  const int r = unit - (int)t[13];
  if (r < H) {                                   // fc1 row r: K = C (multiple of 8 on this path)
    const float* w = w1 + (long long)r * C;
    float acc = 0.f;
    for (int k = lane * 4; k < C; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(w + k);
      const float4 gg = *reinterpret_cast<const float4*>(ln_w + k);
      const float4 bb = *reinterpret_cast<const float4*>(ln_b + k);
      acc = fmaf(v.x, bb.x, fmaf(v.y, bb.y, fmaf(v.z, bb.z, fmaf(v.w, bb.w, acc))));
      st4(w1f + (long long)r * C + k, make_float4(v.x * gg.x, v.y * gg.y, v.z * gg.z, v.w * gg.w));
    }
    acc = warp_sum(acc);
    if (lane == 0) b1f[r] = acc + b1[r];
  } else if (r < H + C) {                        // fc2 row c: K = 4C
    const int c = r - H;
    const float g = gamma ? gamma[c] : 1.f;
    const float* w = w2 + (long long)c * H;
    for (int k = lane * 4; k < H; k += 128) {
      const float4 v = *reinterpret_cast<const float4*>(w + k);
      st4(w2c + (long long)c * H + k, v);
      st4(w2s + (long long)c * H + k, make_float4(v.x * g, v.y * g, v.z * g, v.w * g));
    }
  } else {                                       // depthwise tap
    const int tap = r - H - C;
    for (int c = lane; c < C; c += 32) w49c[(long long)tap * C + c] = dw_w[(long long)c * 49 + tap];
  }
}
extern "C" int ga_block_weight_prep(const void* table, int nblocks, int total_units, ga_stream_t s) {
  GA_REQUIRE(table && nblocks > 0 && total_units > 0, GA_ERR_SHAPE, "ga_block_weight_prep: bad arguments");
  block_weight_prep_kernel<<<(total_units + 7) / 8, 256, 0, (cudaStream_t)s>>>((const long long*)table, nblocks, total_units);
  return launch_ok("block_weight_prep");
}
extern "C" int ga_cast_bf16(const float* src, void* dst, long long n, ga_stream_t s) {
  if (n == 0) return GA_OK;
  GA_REQUIRE(n < (1LL << 31), GA_ERR_SHAPE, "ga_cast_bf16: too large");
  return ga_scale_matrix(src, nullptr, nullptr, dst, 1, (int)n, GA_BF16, s);
}

// y[r, c] = x[r, c] * rowscale[r / rows_per_scale]   (drop-path on a gradient)
template <typename T>
__global__ void scale_rows_kernel(const T* __restrict__ x, const float* __restrict__ rs, T* __restrict__ y, long long M, int C,
                                  int rows_per_scale) {
  const int C4 = C >> 2;
  const long long total = M * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C4;
    const float sc = rs[r / rows_per_scale];
    float4 v = ld4(x + i * 4);
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    st4(y + i * 4, v);
  }
}
extern "C" int ga_scale_rows(const void* x, const float* rowscale, void* y, long long M, int C, int rows_per_scale, int dtype,
                             ga_stream_t s) {
  GA_REQUIRE(x && y && rowscale && (C & 3) == 0, GA_ERR_ALIGN, "ga_scale_rows: bad arguments");
  const long long total = M * (C >> 2);
  if (total == 0) return GA_OK;
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  DISPATCH_T(dtype, { scale_rows_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, rowscale, (T*)y, M, C, rows_per_scale); });
  return launch_ok("scale_rows");
}

// Linear gradient finalisation with folded scales (see header).  One CTA per output row n (dW, dg) + column pass.
__global__ void __launch_bounds__(256) linear_grad_rows_kernel(const float* __restrict__ G, const float* __restrict__ s,
                                                               const float* __restrict__ W, const float* __restrict__ bias,
                                                               const float* __restrict__ g, const float* __restrict__ w,
                                                               const float* __restrict__ b_in, float* __restrict__ dW,
                                                               float* __restrict__ dbias, float* __restrict__ dg, int N, int K) {
  __shared__ float red[32];
  const int n = blockIdx.x;
  const float gn = g ? g[n] : 1.f;
  const float sn = (b_in && s) ? s[n] : 0.f;
  float dot = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float Gv = G[(size_t)n * K + k];
    const float wk = w ? w[k] : 1.f;
    // the layer saw in' = in*w + b_in, so dOut^T in' = G*w + s (x) b_in
    if (dW) dW[(size_t)n * K + k] += gn * (Gv * wk + (b_in ? sn * b_in[k] : 0.f));
    if (dg) dot += W[(size_t)n * K + k] * wk * Gv;
  }
  if (dg) {
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) dg[n] += dot + (bias ? bias[n] * s[n] : 0.f);
  }
  if (dbias && threadIdx.x == 0) dbias[n] += gn * s[n];
}
// dw[k] += sum_n W[n,k]*g[n]*G[n,k] ; db_in[k] += sum_n W[n,k]*g[n]*s[n]
__global__ void __launch_bounds__(256) linear_grad_cols_kernel(const float* __restrict__ G, const float* __restrict__ s,
                                                               const float* __restrict__ W, const float* __restrict__ g,
                                                               float* __restrict__ dw, float* __restrict__ db_in, int N, int K) {
  const int k = blockIdx.x * 32 + (threadIdx.x & 31);
  const int wid = threadIdx.x >> 5;
  __shared__ float sh[2][8][32];
  float a = 0.f, b = 0.f;
  const int rows_per = (N + gridDim.y - 1) / gridDim.y;
  const int n_begin = blockIdx.y * rows_per;
  const int n_end = min(N, n_begin + rows_per);
  if (k < K) {
    for (int n = n_begin + wid; n < n_end; n += 8) {
      const float wv = W[(size_t)n * K + k] * (g ? g[n] : 1.f);
      a += wv * G[(size_t)n * K + k];
      b += wv * s[n];
    }
  }
  sh[0][wid][threadIdx.x & 31] = a; sh[1][wid][threadIdx.x & 31] = b;
  __syncthreads();
  if (wid == 0 && k < K) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { ta += sh[0][i][threadIdx.x]; tb += sh[1][i][threadIdx.x]; }
    if (dw) atomicAdd(dw + k, ta);
    if (db_in) atomicAdd(db_in + k, tb);
  }
}
extern "C" int ga_linear_grad_finalize(const float* G, const float* s, const float* W, const float* bias, const float* g,
                                       const float* w, const float* b_in, float* dW, float* dbias, float* dg, float* dw,
                                       float* db_in, int N, int K, ga_stream_t st) {
  GA_REQUIRE(G && N > 0 && K > 0, GA_ERR_SHAPE, "ga_linear_grad_finalize: bad arguments");
  GA_REQUIRE((!dg && !dw && !db_in) || W, GA_ERR_SHAPE, "ga_linear_grad_finalize: scale gradients need W");
  GA_REQUIRE((!dbias && !db_in && !(dg && bias)) || s, GA_ERR_SHAPE, "ga_linear_grad_finalize: bias gradients need s");
  if (dW || dg || dbias) {
    linear_grad_rows_kernel<<<N, 256, 0, (cudaStream_t)st>>>(G, s, W, bias, g, w, b_in, dW, dbias, dg, N, K);
    int rc = launch_ok("linear_grad_rows");
    if (rc) return rc;
  }
  if (dw || db_in) {
    const int kb = (K + 31) / 32;
    int ny = (2 * 148 + kb - 1) / kb;
    if (ny > (N + 63) / 64) ny = (N + 63) / 64;
    if (ny < 1) ny = 1;
    linear_grad_cols_kernel<<<dim3(kb, ny), 256, 0, (cudaStream_t)st>>>(G, s, W, g, dw, db_in, N, K);
    return launch_ok("linear_grad_cols");
  }
  return GA_OK;
}

// dz = dy * act'(.)  : GELU'(z) from the saved pre-activation, or the ReLU mask from the saved output
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ zy, T* __restrict__ dz, long long M, int C,
                               long long lddy, long long ldz, long long lddz, int act) {
  const int C4 = C >> 2;
  const long long total = M * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const long long r = i / C4;
    float4 d = ld4(dy + r * lddy + c), z = ld4(zy + r * ldz + c);
    if (act == GA_ACT_GELU) { d.x *= gelu_grad_f(z.x); d.y *= gelu_grad_f(z.y); d.z *= gelu_grad_f(z.z); d.w *= gelu_grad_f(z.w); }
    else { d.x = z.x > 0.f ? d.x : 0.f; d.y = z.y > 0.f ? d.y : 0.f; d.z = z.z > 0.f ? d.z : 0.f; d.w = z.w > 0.f ? d.w : 0.f; }
    st4(dz + r * lddz + c, d);
  }
}
extern "C" int ga_act_bwd(const void* dy, const void* zy, void* dz, long long M, int C, long long lddy, long long ldz,
                          long long lddz, int act, int dtype, ga_stream_t s) {
  GA_REQUIRE(dy && zy && dz && (C & 3) == 0 && (lddy & 3) == 0 && (ldz & 3) == 0 && (lddz & 3) == 0, GA_ERR_ALIGN,
             "ga_act_bwd: C/ld must be multiples of 4 (C=%d)", C);
  const long long total = M * (C >> 2);
  if (total == 0) return GA_OK;
  const int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  DISPATCH_T(dtype, { act_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)dy, (const T*)zy, (T*)dz, M, C, lddy, ldz, lddz, act); });
  return launch_ok("act_bwd");
}


// ---------------------------------------------------------------------------------------------- input batch preparation
// uint8 NCHW -> fp32 NCHW, normalised, optionally mixed with the batch in reverse order (timm Mixup: x.flip(0)); 4 pixels / thread
__global__ void __launch_bounds__(256) prep_batch_kernel(const unsigned char* __restrict__ x, float* __restrict__ y, int B, long long plane,
                                                         int W, float m0, float m1, float m2, float r0, float r1, float r2, int mode,
                                                         float lam, int y0, int y1, int x0, int x1) {
  const long long per_img = 3 * plane;
  const long long total4 = (long long)B * per_img / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 4;
    const int b = (int)(e / per_img);
    const long long r = e - (long long)b * per_img;
    const int c = (int)(r / plane);
    const long long p = r - (long long)c * plane;
    const float mu = c == 0 ? m0 : (c == 1 ? m1 : m2), rs = c == 0 ? r0 : (c == 1 ? r1 : r2);
    const uchar4 a = *reinterpret_cast<const uchar4*>(x + e);
    float v[4] = {(float)a.x, (float)a.y, (float)a.z, (float)a.w};
    if (mode) {
      const uchar4 o = *reinterpret_cast<const uchar4*>(x + (long long)(B - 1 - b) * per_img + r);
      const float w[4] = {(float)o.x, (float)o.y, (float)o.z, (float)o.w};
      const int py = (int)(p / W), px = (int)(p - (long long)py * W);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (mode == 1) v[k] = lam * v[k] + (1.f - lam) * w[k];
        else if (py >= y0 && py < y1 && px + k >= x0 && px + k < x1) v[k] = w[k];
      }
    }
    *reinterpret_cast<float4*>(y + e) = make_float4((v[0] - mu) * rs, (v[1] - mu) * rs, (v[2] - mu) * rs, (v[3] - mu) * rs);
  }
}
extern "C" int ga_prep_batch(const unsigned char* x, float* y, int B, int H, int W, const float* mean3, const float* std3, int mode,
                             float lam, int y0, int y1, int x0, int x1, ga_stream_t s) {
  GA_REQUIRE(x && y && mean3 && std3 && B > 0 && (W & 3) == 0, GA_ERR_SHAPE, "ga_prep_batch: bad arguments (W must be a multiple of 4)");
  const long long plane = (long long)H * W, total4 = (long long)B * 3 * plane / 4;
  long long blocks = (total4 + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  prep_batch_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(x, y, B, plane, W, mean3[0] * 255.f, mean3[1] * 255.f, mean3[2] * 255.f,
                                                                   1.f / (std3[0] * 255.f), 1.f / (std3[1] * 255.f), 1.f / (std3[2] * 255.f),
                                                                   mode, lam, y0, y1, x0, x1);
  return launch_ok("prep_batch");
}
