// Shared device helpers for the ga_sm100 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#define GA_OK 0
#define GA_ERR_SHAPE 1
#define GA_ERR_ALIGN 2
#define GA_ERR_LAUNCH 3
#define GA_ERR_UNSUPPORTED 4

#include "../../include/ga_sm100.h"

typedef __nv_bfloat16 bf16;

void ga_set_error(const char* fmt, ...);
int ga_check_launch(const char* what);
int ga_num_sms();
// per call site, per device "first time" flag: opt-ins such as cudaFuncAttributeMaxDynamicSharedMemorySize are per device,
// so a process that drives several GPUs must repeat them on each
struct GaPerDevice { bool done[64] = {}; };
bool ga_first_on_device(GaPerDevice& s);
void ga_count_launch();
// cached cuTensorMapEncodeTiled: rank<=5, dims/box innermost first, strides in BYTES for dims 1..rank-1
struct CUtensorMap_st;
int ga_tensor_map(CUtensorMap_st* out, int dtype, int rank, const void* ptr, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);

#define GA_REQUIRE(cond, code, ...)            \
  do {                                         \
    if (!(cond)) {                             \
      ga_set_error(__VA_ARGS__);               \
      return (code);                           \
    }                                          \
  } while (0)

// ---- scalar load / store with dtype conversion ------------------------------------------------
template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 4 consecutive elements <-> float4 (16-byte aligned for float, 8-byte aligned for bf16)
template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<bf16>(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  return make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y));
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<bf16>(bf16* p, float4 v) {
  uint2 u; u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = u;
}
// 8 consecutive bf16 (16 B) <-> 8 floats
__device__ __forceinline__ void ld8_bf16(const bf16* p, float* f) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ void st8_bf16(bf16* p, const float* f) {
  uint4 u; u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---- reductions ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result valid in every thread; `red` must hold >= 32 floats; blockDim.x multiple of 32
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (l < nw) ? red[l] : 0.f;
  t = warp_sum(t);
  return t;
}

// ---- activations: nn.GELU() (erf form).  erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, i.e. fp32 round-off).
// The GELU epilogues are issue-bound, so the form below is the shortest instruction sequence found: with
//   h(x) = 0.5 * erfc(|x|/sqrt2) = 0.5 * poly(t) * t * exp(-x^2/2),  t = 1/(1 + p|x|/sqrt2)
//   gelu(x)  = relu(x) - |x| * h            (x>=0: x(1-h);  x<0: x*h)
//   gelu'(x) = Phi(x) + x * phi(x),  Phi = x>=0 ? 1-h : h,  phi = exp(-x^2/2)/sqrt(2 pi)
// = 2 MUFU (rcp, ex2) + 12 FMA-pipe instructions for gelu.  exp(-x^2/2) is shared by h and phi.
__device__ __forceinline__ float gelu_h(float x, float* ex_out) {
  const float u = fabsf(x);
  float t, ex;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.23164189f, u, 1.f)));        // p/sqrt2 = 0.3275911/1.41421356
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"((u * u) * -0.72134752f));           // exp(-x^2/2) = 2^(-x^2 * 0.5*log2 e)
  float poly = fmaf(0.5307027145f, t, -0.7265760135f);                                   // 0.5 * A-S coefficients
  poly = fmaf(poly, t, 0.7107068705f);
  poly = fmaf(poly, t, -0.142248368f);
  poly = fmaf(poly, t, 0.127414796f);
  *ex_out = ex;
  return (poly * t) * ex;
}
__device__ __forceinline__ float gelu_f(float x) {
  float ex;
  const float h = gelu_h(x, &ex);
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}
// ---- gelu and gelu' on two values at once with Blackwell's packed fp32x2 FMA pipe instructions (fma/mul/add.f32x2: two
// results per issue slot; the GELU epilogues are issue-bound).  Bit-identical to the scalar forms above.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) { return ((f32x2_t)__float_as_uint(hi) << 32) | (f32x2_t)__float_as_uint(lo); }
__device__ __forceinline__ float f2_lo(f32x2_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(f32x2_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ f32x2_t f2_splat(float c) { return f2_pack(c, c); }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// y = gelu(x), dy = gelu'(x) for the pair (x0, x1); WANT_D = false skips the derivative.
// h = 1 - Phi(|x|) by Abramowitz-Stegun 26.2.17 (|error| < 7.5e-8), in packed fp32x2 arithmetic where the operands are pairs.
// Sign handling is kept off the 64-bit integer path (every packed AND / XOR is two LOP3): |x| enters through the free source
// modifier of a scalar FFMA, and with the derivative wanted Phi = 0.5 + copysign(0.5 - h, x) serves both outputs
// (y = x Phi, dy = Phi + x phi): 21 instructions per pair instead of 28 in a kernel that is issue-bound.
template <bool WANT_D>
__device__ __forceinline__ void gelu_pair(float x0, float x1, float* y0, float* y1, float* d0, float* d1) {
  const f32x2_t x = f2_pack(x0, x1);
  float t0, t1, e0, e1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(fmaf(0.23164189f, fabsf(x0), 1.f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(fmaf(0.23164189f, fabsf(x1), 1.f)));
  const f32x2_t ein = f2_mul(f2_mul(x, x), f2_splat(-0.72134752f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(f2_lo(ein)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(f2_hi(ein)));
  const f32x2_t t = f2_pack(t0, t1), ex = f2_pack(e0, e1);
  f32x2_t poly = f2_fma(f2_splat(0.5307027145f), t, f2_splat(-0.7265760135f));
  poly = f2_fma(poly, t, f2_splat(0.7107068705f));
  poly = f2_fma(poly, t, f2_splat(-0.142248368f));
  poly = f2_fma(poly, t, f2_splat(0.127414796f));
  const f32x2_t h = f2_mul(f2_mul(poly, t), ex);
  if (WANT_D) {
    const f32x2_t q = f2_fma(h, f2_splat(-1.f), f2_splat(0.5f));                // 0.5 - h  in [0, 0.5)
    const f32x2_t cdf = f2_add(f2_pack(copysignf(f2_lo(q), x0), copysignf(f2_hi(q), x1)), f2_splat(0.5f));
    const f32x2_t yv = f2_mul(x, cdf);
    const f32x2_t dv = f2_fma(f2_mul(x, f2_splat(0.39894228040143268f)), ex, cdf);
    *y0 = f2_lo(yv); *y1 = f2_hi(yv);
    *d0 = f2_lo(dv); *d1 = f2_hi(dv);
  } else {
    const float h0 = f2_lo(h), h1 = f2_hi(h);
    *y0 = fmaf(-fabsf(x0), h0, fmaxf(x0, 0.f));                                 // relu(x) - |x| h
    *y1 = fmaf(-fabsf(x1), h1, fmaxf(x1, 0.f));
  }
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float ex;
  const float h = gelu_h(x, &ex);
  const float cdf = x >= 0.f ? 1.f - h : h;
  return fmaf(x * 0.39894228040143268f, ex, cdf);
}
