// GEMM with fused epilogue for every Linear / 1x1 / k=s conv on the GA / MAP hot path.
//   bf16 operands : tcgen05.mma (cta_group::1, M=128) fed by TMA (SWIZZLE_128B), fp32 accumulators in TMEM,
//                   warp-specialised {TMA producer, MMA issuer, 4 epilogue warps}, smem-staged coalesced epilogue.
//                   Operands may be K-major or MN-major (wgrad / dgrad read activations and weights in place).
//   fp32 operands : register-tiled SIMT kernel (exact fp32 FMA; the 1e-5 parity path and odd-shape fallback).
// Replaces the cuBLAS(Lt)/cuDNN calls behind ga_convnext.py:107-111,127,357,260-282,407,418,163-167,202-205,422,460.
#include "common.cuh"
#include "../../include/ga_sm100.h"
#include <cuda.h>
#include <string.h>
#include <stdlib.h>



// ------------------------------------------------------------------------------------------------ epilogue
struct EpiArgs {
  void* D; long long ldd, d_bs, d_cs;
  int out_f32, accumulate;
  float alpha;
  const float* bias; long long bias_bs;
  int act;
  void* Z;
  const float* colscale; long long colscale_bs;
  const float* rowscale; int rows_per_scale;
  const void* R; long long ldr, r_bs;
  const void* Zin; long long ldz, z_bs; int zmode;
  int z_shadow;  // Z receives a bf16 copy of the final value instead of the pre-activation
  float* colsum;  // x act' epilogue only: colsum[n] += sum_m D[m,n] (the bias gradient of the layer whose dz this GEMM produces)
  const void* ln_xhat; long long ld_xhat; const float* ln_rstd;   // fused LayerNorm backward of the output rows (EPI_LNBWD)
  int M, N;
};

__device__ __forceinline__ float epi_act(float v, int act) {
  if (act == GA_ACT_GELU) return gelu_f(v);
  if (act == GA_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// one element (generic / tail path)
__device__ __forceinline__ void epi_store1(const EpiArgs& e, int b, int m, int n, float acc) {
  float v = acc * e.alpha;
  long long off = (long long)b * e.d_bs + (long long)m * e.ldd + (long long)n * e.d_cs;
  if (e.Zin) {
    long long zo = (long long)b * e.z_bs + (long long)m * e.ldz + n;
    float z = e.out_f32 ? ((const float*)e.Zin)[zo] : __bfloat162float(((const bf16*)e.Zin)[zo]);
    v *= (e.zmode == GA_ACT_MUL) ? z : (e.zmode == GA_ACT_GELU) ? gelu_grad_f(z) : (z > 0.f ? 1.f : 0.f);
  } else {
    if (e.bias) v += e.bias[(long long)b * e.bias_bs + n];
    if (e.Z && e.z_shadow != 1) {
      const float zv = e.z_shadow == 2 ? gelu_grad_f(v) : v;
      if (e.out_f32) ((float*)e.Z)[off] = zv; else ((bf16*)e.Z)[off] = __float2bfloat16_rn(zv);
    }
    v = epi_act(v, e.act);
    if (e.colscale) v *= e.colscale[(long long)b * e.colscale_bs + n];
    if (e.rowscale) v *= e.rowscale[m / e.rows_per_scale];
    if (e.R) {
      long long ro = (long long)b * e.r_bs + (long long)m * e.ldr + n;
      v += e.out_f32 ? ((const float*)e.R)[ro] : __bfloat162float(((const bf16*)e.R)[ro]);
    }
  }
  if (e.Z && e.z_shadow == 1) ((bf16*)e.Z)[off] = __float2bfloat16_rn(v);
  if (e.accumulate) atomicAdd(((float*)e.D) + off, v);
  else if (e.out_f32) ((float*)e.D)[off] = v;
  else ((bf16*)e.D)[off] = __float2bfloat16_rn(v);
}

// four consecutive columns n..n+3 (requires N%4==0, ld%4==0, 16B-aligned bases): coalesced vector path.
// `pre` holds the already-loaded per-element operand: Zin values when e.Zin is set, else the residual R values.
__device__ __forceinline__ void epi_store4p(const EpiArgs& e, int b, int m, int n, float4 acc, float4 pre) {
  float v[4] = {acc.x * e.alpha, acc.y * e.alpha, acc.z * e.alpha, acc.w * e.alpha};
  const long long off = (long long)b * e.d_bs + (long long)m * e.ldd + n;
  if (e.Zin) {
    const float zz[4] = {pre.x, pre.y, pre.z, pre.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      v[i] *= (e.zmode == GA_ACT_MUL) ? zz[i] : (e.zmode == GA_ACT_GELU) ? gelu_grad_f(zz[i]) : (zz[i] > 0.f ? 1.f : 0.f);
  } else {
    if (e.bias) {
      const float4 bb = *reinterpret_cast<const float4*>(e.bias + (long long)b * e.bias_bs + n);
      v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
    }
    if (e.Z && e.z_shadow != 1) {
      const float4 zv = e.z_shadow == 2 ? make_float4(gelu_grad_f(v[0]), gelu_grad_f(v[1]), gelu_grad_f(v[2]), gelu_grad_f(v[3]))
                                        : make_float4(v[0], v[1], v[2], v[3]);
      if (e.out_f32) st4((float*)e.Z + off, zv); else st4((bf16*)e.Z + off, zv);
    }
    if (e.act == GA_ACT_GELU) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = gelu_f(v[i]);
    } else if (e.act == GA_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (e.colscale) {
      const float4 cs = *reinterpret_cast<const float4*>(e.colscale + (long long)b * e.colscale_bs + n);
      v[0] *= cs.x; v[1] *= cs.y; v[2] *= cs.z; v[3] *= cs.w;
    }
    if (e.rowscale) {
      const float rs = e.rowscale[m / e.rows_per_scale];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] *= rs;
    }
    if (e.R) { v[0] += pre.x; v[1] += pre.y; v[2] += pre.z; v[3] += pre.w; }
  }
  if (e.Z && e.z_shadow == 1) st4((bf16*)e.Z + off, make_float4(v[0], v[1], v[2], v[3]));
  if (e.accumulate) {
    float* d = (float*)e.D + off;
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(d + i, v[i]);
  } else {
    const float4 o = make_float4(v[0], v[1], v[2], v[3]);
    if (e.out_f32) st4((float*)e.D + off, o); else st4((bf16*)e.D + off, o);
  }
}

__device__ __forceinline__ void epi_store4(const EpiArgs& e, int b, int m, int n, float4 acc) {
  float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e.Zin) {
    const long long zo = (long long)b * e.z_bs + (long long)m * e.ldz + n;
    pre = e.out_f32 ? ld4((const float*)e.Zin + zo) : ld4((const bf16*)e.Zin + zo);
  } else if (e.R) {
    const long long ro = (long long)b * e.r_bs + (long long)m * e.ldr + n;
    pre = e.out_f32 ? ld4((const float*)e.R + ro) : ld4((const bf16*)e.R + ro);
  }
  epi_store4p(e, b, m, n, acc, pre);
}

// ------------------------------------------------------------------------------------------------ SIMT kernel
// 64x64 tile, BK=16, 256 threads, 4x4 micro-tile, arbitrary element strides, fp32 or bf16 inputs.
template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, long long a_rs, long long a_cs, long long a_bs,
                                                        const T* __restrict__ Bm, long long b_rs, long long b_cs, long long b_bs,
                                                        int K, EpiArgs e) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16x16 threads, each 4x4
  A += (long long)b * a_bs;
  Bm += (long long)b * b_bs;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: choose the fast index along the contiguous dimension
  const bool a_kfast = (a_cs == 1);
  const bool b_kfast = (b_cs == 1);
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int idx = tid + it * 256;  // 0..1023
      int mm, kk;
      if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < e.M && gk < K) v = ld_f(A + (long long)gm * a_rs + (long long)gk * a_cs);
      As[kk][mm] = v;
      int nn;
      if (b_kfast) { kk = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kk = idx >> 6; }
      int gn = n0 + nn; gk = k0 + kk;
      v = 0.f;
      if (gn < e.N && gk < K) v = ld_f(Bm + (long long)gn * b_rs + (long long)gk * b_cs);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a_[4] = {av.x, av.y, av.z, av.w}, b_[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= e.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < e.N) epi_store1(e, b, m, n, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ tcgen05 kernel
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int EPI_N = 64;       // epilogue column chunk staged through smem
constexpr int STAGE_LD = EPI_N + 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread = its own TMEM lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B.  K-major: 8-row groups 1024 B apart (SBO), LBO unused(=1).
// MN-major: 64-element MN chunks `lbo` bytes apart, 8-k groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

struct Params {
  int M, N, K;
  int kb_total, kb_per_split, splits;
  int stages;
  uint32_t idesc;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b;  // descriptor byte offsets (host-selected so they can be probed)
  uint32_t wait_ns;                     // persistent kernel: suspend-time hint of the mbarrier waits (0 = poll + nanosleep)
  int step_n, step_m, step_z;           // persistent kernel: gridDim.x decomposed over (n tiles, m tiles, batch*splits)
};

// grid: (m tiles, n tiles, batch*splits).  192 threads: warp0 TMA, warp1 MMA(+TMEM alloc), warps 2..5 epilogue.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, Params p, EpiArgs e) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B aligned carve-up (SWIZZLE_128B atoms)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  uint8_t* stage_base = smem;
  float* staging = (float*)(smem + (size_t)p.stages * STAGE_BYTES);          // 4 warps x 32 x STAGE_LD floats
  uint64_t* full_bar = (uint64_t*)((uint8_t*)staging + 4 * 32 * STAGE_LD * 4);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full = empty_bar + 8;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int batch = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int kb_begin = split * p.kb_per_split;
  int kb_end = kb_begin + p.kb_per_split;
  if (kb_end > p.kb_total) kb_end = p.kb_total;
  const int nkb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (i / p.stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        uint8_t* sa = stage_base + (size_t)s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        const int k0 = (kb_begin + i) * BK;
        if (!A_MN) {
          tma_load_3d(sa, &tmA, &full_bar[s], k0, m0, batch);
        } else {
          tma_load_3d(sa, &tmA, &full_bar[s], m0, k0, batch);
          tma_load_3d(sa + 64 * BK * 2, &tmA, &full_bar[s], m0 + 64, k0, batch);
        }
        if (!B_MN) {
          tma_load_3d(sb, &tmB, &full_bar[s], k0, n0, batch);
        } else {
#pragma unroll
          for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * 64 * BK * 2, &tmB, &full_bar[s], n0 + 64 * j, k0, batch);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % p.stages;
        const uint32_t ph = (i / p.stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_base + (size_t)s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
        const uint64_t adesc = make_desc(sa, p.lbo_a, p.sbo_a);
        const uint64_t bdesc = make_desc(sb, p.lbo_b, p.sbo_b);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 K-elements: K-major = 32 B inside the swizzle atom; MN-major = 16 rows of 128 B
          const uint64_t ka = A_MN ? (uint64_t)((k * 16 * 128) >> 4) : (uint64_t)((k * 32) >> 4);
          const uint64_t kb = B_MN ? (uint64_t)((k * 16 * 128) >> 4) : (uint64_t)((k * 32) >> 4);
          umma_bf16(tmem_base, adesc + ka, bdesc + kb, p.idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above retire
      }
      umma_commit(tmem_full);        // accumulator complete
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> smem (transpose) -> coalesced global stores
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    float* st = staging + (size_t)(warp - 2) * 32 * STAGE_LD;
    if (nkb > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    const bool vec_ok = ((e.N & 3) == 0) && ((e.ldd & 3) == 0) && (!e.R || (e.ldr & 3) == 0) && (!e.Zin || (e.ldz & 3) == 0);
#pragma unroll 1
    for (int c = 0; c < BN / EPI_N; ++c) {
      if (n0 + c * EPI_N >= e.N) break;
#pragma unroll
      for (int h = 0; h < EPI_N / 32; ++h) {
        uint32_t r[32];
        if (nkb > 0) {
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * EPI_N + h * 32), r);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        float* row = st + lane * STAGE_LD + h * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(row + 4 * i) =
              make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                          __uint_as_float(r[4 * i + 3]));
      }
      __syncwarp();
      const int cl = (lane & 15) * 4;
      const int n = n0 + c * EPI_N + cl;
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int rl = 2 * i + (lane >> 4);
        const int m = m0 + q * 32 + rl;
        if (m < e.M && n < e.N) {
          float4 acc = *reinterpret_cast<const float4*>(st + rl * STAGE_LD + cl);
          if (vec_ok) {
            epi_store4(e, batch, m, n, acc);
          } else {
            float a_[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < e.N) epi_store1(e, batch, m, n + j, a_[j]);
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ------------------------------------------------------------------------------------------------ persistent kernel
// One CTA per SM walks tiles (n fastest, so the CTAs running concurrently share A rows in L2).  576 threads:
//   warp 0  TMA producer   : smem ring of `stages` {A,B} k-blocks that keeps running across tile boundaries
//   warp 1  MMA issuer     : tcgen05.mma into one of TWO TMEM accumulators (2*BN columns), so the MMAs of tile i+1
//                            overlap the epilogue of tile i
//   warps 2..17 epilogue   : warp e owns TMEM lane quadrant (warp id % 4) and a BN/4 column slice; TMEM -> regs -> smem
//                            transpose -> coalesced global stores.  The epilogue is specialised at compile time (EPI_*):
//                            with GELU in it the kernel is issue-bound, so the hot loop carries no runtime flags, column
//                            operands (bias, gamma) are loaded once per chunk and row operands (residual, Zin) are
//                            prefetched four rows ahead.
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Wait used by every role of the persistent kernel: mbarrier.try_wait with a suspend-time hint parks the thread in hardware
// until the phase completes (or the hint expires), so a waiting warp issues ~3 instructions per microsecond instead of per
// 20 ns -- the polling loops were 14 % of all issued instructions of the short-K GEMMs (profiles/r02_ncu_gemm_epilogue.txt).
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  if (hint_ns == 0) {
    while (true) {
      asm volatile(
          "{\n"
          ".reg .pred p;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
          "selp.u32 %0, 1, 0, p;\n"
          "}\n"
          : "=r"(done)
          : "r"(addr), "r"(parity)
          : "memory");
      if (done) break;
      __nanosleep(40);
    }
    return;
  }
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(hint_ns)
        : "memory");
    if (done) break;
  }
}

constexpr int EPI_WARPS = 16;
constexpr int EPI_C = 32;                 // columns per epilogue chunk (one tcgen05.ld.32x32b.x32)
constexpr int ST_LD = EPI_C + 4;          // staging row pitch (floats): conflict-free float4 writes and reads

enum { EPI_GENERIC = 0, EPI_BIAS_GELU_Z = 1, EPI_BIAS_GELU = 2, EPI_RES_F32_SHADOW = 3, EPI_ZIN_GELU = 4, EPI_PLAIN_BF16 = 5,
       EPI_ACCUM = 6, EPI_LNBWD = 7, EPI_PLAIN_F32 = 8 };
constexpr int LNBWD_XCH_BYTES = 2 * 4 * 4 * 32 * 2 * 4;     // [tile parity][row quadrant][column slice][row][2] fp32

// one 32-row x 32-column chunk held in this warp's staging buffer -> global memory
template <int EPI, bool FULL>
__device__ __forceinline__ float4 epilogue_chunk(const EpiArgs& e, const float* st, int lane, int batch, int m_base, int n,
                                                 const uint2* zraw = nullptr) {
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);     // x act' epilogue: sums of this lane's 4 columns over its 8 rows
  // lane -> 4 columns (lane & 7) * 4 of rows (lane >> 3) + 4 * it, it = 0..7
  const int cl = (lane & 7) * 4;
  const int r0 = lane >> 3;
  if (EPI == EPI_GENERIC) {
    const bool vec_ok = ((e.N & 3) == 0) && ((e.ldd & 3) == 0) && (!e.R || (e.ldr & 3) == 0) && (!e.Zin || (e.ldz & 3) == 0) &&
                        (!e.bias || ((e.bias_bs & 3) == 0)) && (!e.colscale || ((e.colscale_bs & 3) == 0));
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
      const int rl = r0 + 4 * it;
      const int m = m_base + rl;
      if (m < e.M && n < e.N) {
        const float4 acc = *reinterpret_cast<const float4*>(st + rl * ST_LD + cl);
        if (vec_ok) {
          epi_store4(e, batch, m, n, acc);
        } else {
          const float a_[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < e.N) epi_store1(e, batch, m, n + j, a_[j]);
        }
      }
    }
    return csum;
  }
  if (n >= e.N) return csum;
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), cs4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if (EPI == EPI_BIAS_GELU_Z || EPI == EPI_BIAS_GELU || EPI == EPI_RES_F32_SHADOW || EPI == EPI_PLAIN_BF16 || EPI == EPI_PLAIN_F32) {
    if (e.bias) bias4 = *reinterpret_cast<const float4*>(e.bias + (long long)batch * e.bias_bs + n);
  }
  if (EPI == EPI_RES_F32_SHADOW) {
    if (e.colscale) cs4 = *reinterpret_cast<const float4*>(e.colscale + (long long)batch * e.colscale_bs + n);
  }
  // Address arithmetic is hoisted: ONE 64-bit element offset per chunk, then a 32-bit row offset (it * 4 * ldd) widened into
  // each base pointer; staging rows sit at compile-time offsets from one shared-space address.  The integer work was half of
  // the instructions this kernel issued (profiles/r02_ncu_gemm_epilogue.txt).
  const int m_first = m_base + r0;
  const long long eoff = (long long)batch * e.d_bs + (long long)m_first * e.ldd + n;
  float* const d32 = (float*)e.D + eoff;
  bf16* const d16 = (bf16*)e.D + eoff;
  bf16* const z16 = (bf16*)e.Z + eoff;
  const float* const r32 = (const float*)e.R + ((long long)batch * e.r_bs + (long long)m_first * e.ldr + n);
  const bf16* const zin16 = (const bf16*)e.Zin + ((long long)batch * e.z_bs + (long long)m_first * e.ldz + n);
  const uint32_t dstep = 4u * (uint32_t)e.ldd, rstep = 4u * (uint32_t)e.ldr, zstep = 4u * (uint32_t)e.ldz;
  const float* const stl = st + (r0 * ST_LD + cl);
  const int z_shadow = e.z_shadow;
  const bool mul_mode = (e.zmode == GA_ACT_MUL);
  const bool has_z = (e.Z != nullptr);
  const float* const rowscale = e.rowscale;
  // per-sample row scale (DropPath): one division per chunk; 32 consecutive rows span at most two samples when a sample has
  // >= 32 rows (it has 49 at the least on this path), so row it's sample is q0 + (r0 + 4 it >= rows_per_scale)
  int rs_q0 = 0, rs_r0 = 0;
  const int rps = e.rows_per_scale;
  if (EPI == EPI_RES_F32_SHADOW && rowscale) { rs_q0 = m_first / rps; rs_r0 = m_first - rs_q0 * rps; }
#pragma unroll
  for (int g4 = 0; g4 < 2; ++g4) {
    float4 pre[4];
    if (EPI == EPI_RES_F32_SHADOW || EPI == EPI_ZIN_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t it = g4 * 4 + j;
        pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (FULL || m_first + 4 * (int)it < e.M) {
          if (EPI == EPI_RES_F32_SHADOW) pre[j] = ld4(r32 + it * rstep);
          else if (zraw) { const uint2 u = zraw[it]; pre[j] = make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y)); }
          else pre[j] = ld4(zin16 + it * zstep);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t it = g4 * 4 + j;
      const int m = m_first + 4 * (int)it;
      if (!FULL && m >= e.M) continue;
      const float4 acc = *reinterpret_cast<const float4*>(stl + it * (4 * ST_LD));
      const uint32_t off = it * dstep;
      float v[4] = {acc.x, acc.y, acc.z, acc.w};
      if (EPI == EPI_BIAS_GELU_Z || EPI == EPI_BIAS_GELU) {
        v[0] += bias4.x; v[1] += bias4.y; v[2] += bias4.z; v[3] += bias4.w;
        if (EPI == EPI_BIAS_GELU_Z) {
          if (z_shadow == 2) {         // save gelu'(z) instead of z: the backward epilogue becomes one multiply
            float gd[4];
            gelu_pair<true>(v[0], v[1], &v[0], &v[1], &gd[0], &gd[1]);
            gelu_pair<true>(v[2], v[3], &v[2], &v[3], &gd[2], &gd[3]);
            st4(z16 + off, make_float4(gd[0], gd[1], gd[2], gd[3]));
          } else {
            st4(z16 + off, make_float4(v[0], v[1], v[2], v[3]));
            gelu_pair<false>(v[0], v[1], &v[0], &v[1], nullptr, nullptr);
            gelu_pair<false>(v[2], v[3], &v[2], &v[3], nullptr, nullptr);
          }
        } else {
          gelu_pair<false>(v[0], v[1], &v[0], &v[1], nullptr, nullptr);
          gelu_pair<false>(v[2], v[3], &v[2], &v[3], nullptr, nullptr);
        }
        st4(d16 + off, make_float4(v[0], v[1], v[2], v[3]));
      } else if (EPI == EPI_RES_F32_SHADOW) {
        float rs = 1.f;
        if (rowscale) {
          const int rr = rs_r0 + 4 * (int)it;
          rs = rowscale[rps >= 32 ? rs_q0 + (rr >= rps ? 1 : 0) : m / rps];
        }
        v[0] = fmaf((v[0] + bias4.x) * cs4.x, rs, pre[j].x); v[1] = fmaf((v[1] + bias4.y) * cs4.y, rs, pre[j].y);
        v[2] = fmaf((v[2] + bias4.z) * cs4.z, rs, pre[j].z); v[3] = fmaf((v[3] + bias4.w) * cs4.w, rs, pre[j].w);
        const float4 o = make_float4(v[0], v[1], v[2], v[3]);
        st4(d32 + off, o);
        if (has_z) st4(z16 + off, o);
      } else if (EPI == EPI_ZIN_GELU) {
        const float zz[4] = {pre[j].x, pre[j].y, pre[j].z, pre[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] *= mul_mode ? zz[i] : gelu_grad_f(zz[i]);
        st4(d16 + off, make_float4(v[0], v[1], v[2], v[3]));
        csum.x += v[0]; csum.y += v[1]; csum.z += v[2]; csum.w += v[3];
      } else if (EPI == EPI_PLAIN_BF16) {
        st4(d16 + off, make_float4(v[0] + bias4.x, v[1] + bias4.y, v[2] + bias4.z, v[3] + bias4.w));
      } else if (EPI == EPI_PLAIN_F32) {
        st4(d32 + off, make_float4(v[0] + bias4.x, v[1] + bias4.y, v[2] + bias4.z, v[3] + bias4.w));
      } else if (EPI == EPI_ACCUM) {
        atomicAdd(reinterpret_cast<float4*>(d32 + off), make_float4(v[0] * e.alpha, v[1] * e.alpha, v[2] * e.alpha, v[3] * e.alpha));
      }
    }
  }
  return csum;
}

// EPI_LNBWD: the GEMM produces dxhat (gradient w.r.t. a LayerNorm's normalised output, affine folded away) and the row's
// LayerNorm backward   dconv = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))   is applied before anything is written,
// so dxhat never reaches HBM (ga_ln_bwd_rows re-read it and xhat: 3 x M x C x 2 bytes).  Needs the whole row in one tile
// (N <= BN = 128).  Within a row quadrant the up-to-four warps that own the row's 32-column slices exchange their two partial
// sums per row through shared memory (double-buffered by tile parity) and meet at a named barrier (ids 1..4, one per quadrant).
__device__ __forceinline__ void epilogue_lnbwd(const EpiArgs& e, const float* st, float* xch, int lane, int q, int slice, int nlive,
                                               int m_base, int n, const uint2* xraw, const float* rsv) {
  const int cl = (lane & 7) * 4, r0 = lane >> 3;
  const float* const stl = st + (r0 * ST_LD + cl);
  float s1[8], s2[8];
  float4 xh[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const float4 g = *reinterpret_cast<const float4*>(stl + it * (4 * ST_LD));
    const uint2 u = xraw[it];
    xh[it] = make_float4(bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y));
    s1[it] = (g.x + g.y) + (g.z + g.w);
    s2[it] = (g.x * xh[it].x + g.y * xh[it].y) + (g.z * xh[it].z + g.w * xh[it].w);
  }
#pragma unroll
  for (int o = 1; o <= 4; o <<= 1) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      s1[it] += __shfl_xor_sync(0xffffffffu, s1[it], o);
      s2[it] += __shfl_xor_sync(0xffffffffu, s2[it], o);
    }
  }
  float2* const mine = reinterpret_cast<float2*>(xch) + ((q * 4 + slice) * 32);
  if ((lane & 7) == 0) {
#pragma unroll
    for (int it = 0; it < 8; ++it) mine[r0 + 4 * it] = make_float2(s1[it], s2[it]);
  }
  asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * nlive) : "memory");
  const float2* const quad = reinterpret_cast<const float2*>(xch) + (q * 4) * 32;
  const float invN = 1.f / (float)e.N;
  const long long eoff = (long long)(m_base + r0) * e.ldd + n;
  bf16* const d16 = (bf16*)e.D + eoff;
  const uint32_t dstep = 4u * (uint32_t)e.ldd;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = r0 + 4 * it;
    float t1 = 0.f, t2 = 0.f;
    for (int sl = 0; sl < nlive; ++sl) { const float2 v = quad[sl * 32 + row]; t1 += v.x; t2 += v.y; }
    if (m_base + row >= e.M || n >= e.N) continue;
    const float m1 = t1 * invN, m2 = t2 * invN, rs = rsv[it];
    const float4 g = *reinterpret_cast<const float4*>(stl + it * (4 * ST_LD));
    st4(d16 + (uint32_t)it * dstep, make_float4(rs * (g.x - m1 - xh[it].x * m2), rs * (g.y - m1 - xh[it].y * m2),
                                                 rs * (g.z - m1 - xh[it].z * m2), rs * (g.w - m1 - xh[it].w * m2)));
  }
}

// Tile walk of the persistent kernel without per-tile divisions: t = blockIdx.x + k * gridDim.x decomposed as (n_t fastest,
// m_t, z) and advanced by the decomposition of gridDim.x with carries.  Four integer divisions per tile and thread (~100
// instructions) were a quarter of the epilogue warps' instruction stream (profiles/r02_ncu_gemm_epilogue.txt).
struct TileIter {
  int n_t, m_t, z;
  __device__ __forceinline__ void init(int t0, int nt, int mt) {
    n_t = t0 % nt; const int r = t0 / nt; m_t = r % mt; z = r / mt;
  }
  // the step (host-decomposed gridDim.x) and the extents stay in the constant bank: three registers per walker
  __device__ __forceinline__ void next(const Params& p, int nt, int mt) {
    n_t += p.step_n;
    int c = (n_t >= nt) ? 1 : 0;
    n_t -= c ? nt : 0;
    m_t += p.step_m + c;
    c = (m_t >= mt) ? 1 : 0;
    m_t -= c ? mt : 0;
    z += p.step_z + c;
  }
};

template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1) gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                          const __grid_constant__ CUtensorMap tmB, Params p,
                                                                          EpiArgs e, int mt, int nt, int total_tiles) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer arithmetic on the shared array (no integer round trip), so the compiler keeps the shared address space
  // and the staging traffic of the epilogue compiles to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int SLICE = BN / 4;                // columns per epilogue warp
  uint8_t* stage_base = smem;
  float* staging = (float*)(smem + (size_t)p.stages * STAGE_BYTES);            // EPI_WARPS x 32 x ST_LD floats
  float* xch_base = (float*)((uint8_t*)staging + EPI_WARPS * 32 * ST_LD * 4);  // EPI_LNBWD only (LNBWD_XCH_BYTES)
  uint64_t* full_bar = (uint64_t*)((uint8_t*)xch_base + (EPI == EPI_LNBWD ? LNBWD_XCH_BYTES : 0));
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full = empty_bar + 8;         // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      TileIter ti;
      ti.init(blockIdx.x, nt, mt);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ti.next(p, nt, mt)) {
        const int n_t = ti.n_t, m_t = ti.m_t, z = ti.z;
        const int batch = (p.splits == 1) ? z : z / p.splits, split = (p.splits == 1) ? 0 : z % p.splits;
        const int kb0 = split * p.kb_per_split;
        int kb1 = kb0 + p.kb_per_split;
        if (kb1 > p.kb_total) kb1 = p.kb_total;
        const int m0 = m_t * BM, n0 = n_t * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait_backoff(&empty_bar[s], ph ^ 1, p.wait_ns);
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          uint8_t* sa = stage_base + (size_t)s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_3d(sa, &tmA, &full_bar[s], k0, m0, batch);
          } else {
            tma_load_3d(sa, &tmA, &full_bar[s], m0, k0, batch);
            tma_load_3d(sa + 64 * BK * 2, &tmA, &full_bar[s], m0 + 64, k0, batch);
          }
          if (!B_MN) {
            tma_load_3d(sb, &tmB, &full_bar[s], k0, n0, batch);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * 64 * BK * 2, &tmB, &full_bar[s], n0 + 64 * j, k0, batch);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, lt = 0;
      TileIter ti;
      ti.init(blockIdx.x, nt, mt);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt, ti.next(p, nt, mt)) {
        const int split = (p.splits == 1) ? 0 : ti.z % p.splits;
        const int kb0 = split * p.kb_per_split;
        int kb1 = kb0 + p.kb_per_split;
        if (kb1 > p.kb_total) kb1 = p.kb_total;
        const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
        mbar_wait_backoff(&tmem_empty[buf], bph ^ 1, p.wait_ns);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait_backoff(&full_bar[s], ph, p.wait_ns);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          const uint64_t adesc = make_desc(sa, p.lbo_a, p.sbo_a);
          const uint64_t bdesc = make_desc(sb, p.lbo_b, p.sbo_b);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t ka = A_MN ? (uint64_t)((k * 16 * 128) >> 4) : (uint64_t)((k * 32) >> 4);
            const uint64_t kb_ = B_MN ? (uint64_t)((k * 16 * 128) >> 4) : (uint64_t)((k * 32) >> 4);
            umma_bf16(tacc, adesc + ka, bdesc + kb_, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;                   // TMEM lane quadrant accessible to this warp (warp id % 4)
    const int slice = ew >> 2;                // which quarter of the BN columns
    float* st = staging + (size_t)ew * 32 * ST_LD;
    uint32_t lt = 0;
    // fused column sums (bias gradient) of the x act' epilogue: each lane keeps the sums of its 4 columns per chunk across
    // the tiles of one column block and flushes (2 shuffles + one 16-byte red per 8 lanes) when the block changes
    float4 csum[SLICE / EPI_C];
    int cs_n0 = -1;
    const bool want_cs = (EPI == EPI_ZIN_GELU) && e.colsum != nullptr;
    auto cs_flush = [&]() {
#pragma unroll
      for (int c = 0; c < SLICE / EPI_C; ++c) {
        float4 v = csum[c];
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
          v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
          v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
        }
        const int ncol = cs_n0 + slice * SLICE + c * EPI_C + (lane & 7) * 4;
        if (lane < 8 && cs_n0 >= 0 && ncol < e.N) atomicAdd(reinterpret_cast<float4*>(e.colsum + ncol), v);
        csum[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
#pragma unroll
    for (int c = 0; c < SLICE / EPI_C; ++c) csum[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    // x act'(z) epilogue: this warp's share of the saved derivative is fetched straight into registers.  With one chunk per
    // warp (BN = 128) the loads run ONE TILE AHEAD, so their HBM latency hides behind a whole tile of work (issued at the top
    // of the same tile they were exposed: long-scoreboard stalls were 3.3 cycles per issued instruction); with two chunks the
    // registers do not allow that and the loads are issued before the wait for this tile's accumulator.
    constexpr int NCH = SLICE / EPI_C;
    constexpr bool IS_Z = (EPI == EPI_ZIN_GELU);
    uint2 zraw[NCH][8], znext[(IS_Z && NCH == 1) ? 8 : 1];
    auto load_z = [&](const TileIter& w, int c, uint2* zr) {   // chunk c of the tile at w
      const int batch = (p.splits == 1) ? w.z : w.z / p.splits;
      const int m_first = w.m_t * BM + q * 32 + (lane >> 3);
      const int ncol = w.n_t * BN + slice * SLICE + c * EPI_C + (lane & 7) * 4;
      const bf16* zp = (const bf16*)e.Zin + ((long long)batch * e.z_bs + (long long)m_first * e.ldz + ncol);
      const uint32_t zstep = 4u * (uint32_t)e.ldz;
      if ((w.m_t * BM + q * 32 + 32 <= e.M) && (ncol < e.N)) {
#pragma unroll
        for (int it = 0; it < 8; ++it) zr[it] = *reinterpret_cast<const uint2*>(zp + (uint32_t)it * zstep);
      } else {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          zr[it] = make_uint2(0u, 0u);
          if (m_first + 4 * it < e.M && ncol < e.N) zr[it] = *reinterpret_cast<const uint2*>(zp + (uint32_t)it * zstep);
        }
      }
    };
    constexpr bool z_ahead = IS_Z;
    TileIter ti;
    ti.init(blockIdx.x, nt, mt);
    if (z_ahead && blockIdx.x < total_tiles) load_z(ti, 0, NCH == 1 ? znext : zraw[0]);
    float* const st_row = st + lane * ST_LD;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
      const int batch = (p.splits == 1) ? ti.z : ti.z / p.splits;
      const int m0 = ti.m_t * BM, n0 = ti.n_t * BN;
      const TileIter tc = ti;                 // this tile; ti moves on to the tile after it (the prefetch target)
      ti.next(p, nt, mt);
      const bool has_next = (t + (int)gridDim.x < total_tiles);
      if (want_cs && n0 != cs_n0) { cs_flush(); cs_n0 = n0; }
      const uint32_t buf = lt & 1, bph = (lt >> 1) & 1;
      // Prefetch schedule of the x act' operand (registers only).  One chunk per warp: the loads run one tile ahead through
      // `znext`.  Two chunks: chunk 1 of this tile is fetched here (chunk 0's work covers it) and chunk 0 of the NEXT tile right
      // after chunk 0 is consumed (chunk 1's work covers it), so no load is exposed and no extra registers are needed.
      if (IS_Z) {
        if (!z_ahead) {
#pragma unroll
          for (int c = 0; c < NCH; ++c) load_z(tc, c, zraw[c]);
        } else if (NCH == 1) {
#pragma unroll
          for (int it = 0; it < 8; ++it) zraw[0][it] = znext[it];
          if (has_next) load_z(ti, 0, znext);
        } else {
#pragma unroll
          for (int c = 1; c < NCH; ++c) load_z(tc, c, zraw[c]);
        }
      }
      // EPI_LNBWD: this warp's xhat values (the layout of the store phase) and the rows' rstd, fetched before the accumulator wait
      uint2 xraw[EPI == EPI_LNBWD ? 8 : 1];
      float rsv[EPI == EPI_LNBWD ? 8 : 1];
      if (EPI == EPI_LNBWD) {
        const int m_first = m0 + q * 32 + (lane >> 3);
        const int ncol = n0 + slice * SLICE + (lane & 7) * 4;
        const bf16* xp = (const bf16*)e.ln_xhat + ((long long)m_first * e.ld_xhat + ncol);
        const uint32_t xstep = 4u * (uint32_t)e.ld_xhat;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const bool ok = (m_first + 4 * it < e.M) && (ncol < e.N);
          xraw[it] = ok ? *reinterpret_cast<const uint2*>(xp + (uint32_t)it * xstep) : make_uint2(0u, 0u);
          rsv[it] = (m_first + 4 * it < e.M) ? e.ln_rstd[m_first + 4 * it] : 0.f;
        }
      }
      mbar_wait_backoff(&tmem_full[buf], bph, p.wait_ns);
      tc_fence_after();
      const uint32_t tacc = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col0 = slice * SLICE + c * EPI_C;
        const bool live = (n0 + col0 < e.N);
        const bool last = (c == NCH - 1);
        if (IS_Z && NCH > 1 && c == 1 && z_ahead && has_next) load_z(ti, 0, zraw[0]);
        if (live) {
          uint32_t r[32];
          tmem_ld32(tacc + (uint32_t)col0, r);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4*>(st_row + 4 * i) =
                make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        }
        if (last) {                           // all TMEM reads of this tile are done: hand the accumulator back
          tc_fence_before();
          if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
        if (!live) continue;
        __syncwarp();
        if (EPI == EPI_LNBWD) {        // one chunk per warp (BN = 128); slices with columns >= N are not live and skip the barrier
          epilogue_lnbwd(e, st, xch_base + (lt & 1) * (LNBWD_XCH_BYTES / 8), lane, q, slice, (e.N + SLICE - 1) / SLICE, m0 + q * 32,
                         n0 + col0 + (lane & 7) * 4, xraw, rsv);
          __syncwarp();
          continue;
        }
        float4 cs;
        if (m0 + q * 32 + 32 <= e.M) cs = epilogue_chunk<EPI, true>(e, st, lane, batch, m0 + q * 32, n0 + col0 + (lane & 7) * 4, EPI == EPI_ZIN_GELU ? zraw[c] : nullptr);
        else cs = epilogue_chunk<EPI, false>(e, st, lane, batch, m0 + q * 32, n0 + col0 + (lane & 7) * 4, EPI == EPI_ZIN_GELU ? zraw[c] : nullptr);
        if (EPI == EPI_ZIN_GELU) { csum[c].x += cs.x; csum[c].y += cs.y; csum[c].z += cs.z; csum[c].w += cs.w; }
        __syncwarp();
      }
    }
    if (want_cs) cs_flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BN);
}

// ---- host: tensor maps (shared cache in runtime.cu) -------------------------------------------------------
// 3-D bf16 tensor map: dims {d0 (contiguous), d1, d2}, strides in elements, box {b0, b1, 1}, SWIZZLE_128B
static int get_map(const void* ptr, long long d0, long long d1, long long d2, long long s1, long long s2, int b0, int b1,
                   CUtensorMap* out) {
  uint64_t dims[3] = {(uint64_t)d0, (uint64_t)d1, (uint64_t)d2};
  uint64_t strides[2] = {(uint64_t)s1 * 2, (uint64_t)(d2 > 1 ? s2 : s1 * d1) * 2};
  uint32_t box[3] = {(uint32_t)b0, (uint32_t)b1, 1};
  return ga_tensor_map(out, GA_BF16, 3, ptr, dims, strides, box, /*swizzle128=*/1);
}

static uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                     // D format f32
  d |= 1u << 7;                     // A bf16
  d |= 1u << 10;                    // B bf16
  d |= (a_mn ? 1u : 0u) << 15;      // A major (0 = K-major)
  d |= (b_mn ? 1u : 0u) << 16;      // B major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

template <int BN, bool A_MN, bool B_MN>
static int launch(const GaGemm* g, const EpiArgs& e, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc;
  // A
  if (!A_MN) rc = get_map(g->A, g->K, g->M, g->batch, g->a_rs, g->a_bs, BK, BM, &ma);
  else rc = get_map(g->A, g->M, g->K, g->batch, g->a_cs, g->a_bs, 64, BK, &ma);
  if (rc) return rc;
  if (!B_MN) rc = get_map(g->B, g->K, g->N, g->batch, g->b_rs, g->b_bs, BK, BN, &mb);
  else rc = get_map(g->B, g->N, g->K, g->batch, g->b_cs, g->b_bs, 64, BK, &mb);
  if (rc) return rc;

  Params p;
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.kb_total = (g->K + BK - 1) / BK;
  int mt = (g->M + BM - 1) / BM, nt = (g->N + BN - 1) / BN;
  int splits = 1;
  if (g->accumulate) {
    splits = g->splits;
    if (splits <= 0) {
      long long tiles = (long long)mt * nt * g->batch;
      int sms = ga_num_sms();
      splits = (int)(sms / tiles);                 // one balanced wave, see launch2
      if (splits == 6) splits = 8;
      int maxs = p.kb_total / 4; if (maxs < 1) maxs = 1;
      if (splits > maxs) splits = maxs;
      if (splits < 1) splits = 1;
    }
    if (splits > p.kb_total) splits = p.kb_total > 0 ? p.kb_total : 1;
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  if (p.kb_per_split < 1) p.kb_per_split = 1;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  if (splits < 1) splits = 1;
  p.splits = splits;
  constexpr int STAGE_BYTES = (BM + BN) * BK * 2;
  static int stages_env = -1;
  if (stages_env < 0) { const char* sv = getenv("GA_GEMM_STAGES"); stages_env = sv ? atoi(sv) : 0; }
  int stages = stages_env > 0 ? stages_env : (BN <= 128 ? 2 : 3);
  if (stages > 8) stages = 8;
  p.stages = stages;
  p.idesc = make_idesc(BM, BN, A_MN, B_MN);
  p.lbo_a = A_MN ? 64 * BK * 2 : 16; p.sbo_a = 1024;
  p.lbo_b = B_MN ? 64 * BK * 2 : 16; p.sbo_b = 1024;
  size_t smem = 1024 + (size_t)stages * STAGE_BYTES + 4 * 32 * STAGE_LD * 4 + 256;
  static GaPerDevice attr_once;
  if (ga_first_on_device(attr_once))
    cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  dim3 grid(mt, nt, g->batch * splits);
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, 192, smem, st>>>(ma, mb, p, e);
  ga_count_launch();
  return ga_check_launch("gemm_tc");
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch2(const GaGemm* g, const EpiArgs& e, cudaStream_t st) {
  CUtensorMap ma, mb;
  int rc;
  if (!A_MN) rc = get_map(g->A, g->K, g->M, g->batch, g->a_rs, g->a_bs, BK, BM, &ma);
  else rc = get_map(g->A, g->M, g->K, g->batch, g->a_cs, g->a_bs, 64, BK, &ma);
  if (rc) return rc;
  if (!B_MN) rc = get_map(g->B, g->K, g->N, g->batch, g->b_rs, g->b_bs, BK, BN, &mb);
  else rc = get_map(g->B, g->N, g->K, g->batch, g->b_cs, g->b_bs, 64, BK, &mb);
  if (rc) return rc;
  Params p;
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.kb_total = (g->K + BK - 1) / BK;
  const int mt = (g->M + BM - 1) / BM, nt = (g->N + BN - 1) / BN;
  const int sms = ga_num_sms();
  int splits = 1;
  if (g->accumulate) {
    splits = g->splits;
    if (splits <= 0) {
      // one balanced wave: tiles * splits <= #SMs.  Measured (scripts/kernel_bench.py, KB_SPLIT_SWEEP=1, B=256 weight-gradient
      // shapes): against the former two-waves rule 56x56 156 -> 123 us (96 % of the HBM peak), 28x28 103 -> 82 / 139 -> 104,
      // 14x14 74 -> 59, 7x7 64 -> 55; more splits only add red.v4 traffic on the [N, K] gradient tile.
      const long long tiles = (long long)mt * nt * g->batch;
      splits = (int)(sms / tiles);
      if (splits == 6) splits = 8;       // 22..24 tiles: the sweep has 6 splits (K blocks of 131) at 104 us and 8 at 82 ([1536,384], K = 50176)
      int maxs = p.kb_total / 4; if (maxs < 1) maxs = 1;
      if (splits > maxs) splits = maxs;
      if (splits < 1) splits = 1;
    }
    if (splits > p.kb_total) splits = p.kb_total;
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  p.splits = splits;
  constexpr int STAGE_BYTES = (BM + BN) * BK * 2;
  constexpr size_t FIXED = 1024 + (size_t)EPI_WARPS * 32 * ST_LD * 4 + 256 + (EPI == EPI_LNBWD ? LNBWD_XCH_BYTES : 0);
  int stages = (int)((225 * 1024 - FIXED) / STAGE_BYTES);
  static int stages_env = -1;
  if (stages_env < 0) { const char* sv = getenv("GA_GEMM_STAGES"); stages_env = sv ? atoi(sv) : 0; }
  if (stages_env > 0 && stages_env < stages) stages = stages_env;
  if (stages > 8) stages = 8;
  p.stages = stages;
  static int wait_env = -1;
  if (wait_env < 0) { const char* sv = getenv("GA_GEMM_WAIT_NS"); wait_env = sv ? atoi(sv) : 0; }
  p.wait_ns = (uint32_t)wait_env;
  p.idesc = make_idesc(BM, BN, A_MN, B_MN);
  p.lbo_a = A_MN ? 64 * BK * 2 : 16; p.sbo_a = 1024;
  p.lbo_b = B_MN ? 64 * BK * 2 : 16; p.sbo_b = 1024;
  const size_t smem = FIXED + (size_t)stages * STAGE_BYTES;
  static GaPerDevice attr_once;
  if (ga_first_on_device(attr_once))
    cudaFuncSetAttribute(gemm_tc2_kernel<BN, A_MN, B_MN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const long long total = (long long)mt * nt * g->batch * splits;
  GA_REQUIRE(total < (1LL << 31), GA_ERR_SHAPE, "ga_gemm: too many tiles");
  const int grid = (int)(total < sms ? total : sms);
  p.step_n = grid % nt; p.step_m = (grid / nt) % mt; p.step_z = grid / (nt * mt);
  gemm_tc2_kernel<BN, A_MN, B_MN, EPI><<<grid, 64 + 32 * EPI_WARPS, smem, st>>>(ma, mb, p, e, mt, nt, (int)total);
  ga_count_launch();
  return ga_check_launch("gemm_tc2");
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------ entry point

static bool tc_eligible(const GaGemm* g, bool* a_mn, bool* b_mn) {
  if (g->in_dtype != GA_BF16) return false;
  if (g->M < 1 || g->N < 1 || g->K < 1) return false;
  auto ok_ptr = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (!ok_ptr(g->A) || !ok_ptr(g->B)) return false;
  if (g->batch > 1 && ((g->a_bs & 7) || (g->b_bs & 7))) return false;
  if (g->a_cs == 1 && (g->a_rs & 7) == 0 && g->a_rs >= g->K) *a_mn = false;
  else if (g->a_rs == 1 && (g->a_cs & 7) == 0 && g->a_cs >= g->M) *a_mn = true;
  else return false;
  if (g->b_cs == 1 && (g->b_rs & 7) == 0 && g->b_rs >= g->K) *b_mn = false;
  else if (g->b_rs == 1 && (g->b_cs & 7) == 0 && g->b_cs >= g->N) *b_mn = true;
  else return false;
  // MN-major operands are loaded in 64-wide chunks: the contiguous extent must cover whole 16-byte rows (checked above)
  return true;
}

extern "C" int ga_gemm(const GaGemm* g, ga_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  GA_REQUIRE(g && g->A && g->B && g->D, GA_ERR_SHAPE, "ga_gemm: null operand");
  GA_REQUIRE(g->M >= 0 && g->N >= 0 && g->K >= 0 && g->batch >= 1, GA_ERR_SHAPE, "ga_gemm: bad sizes M=%d N=%d K=%d batch=%d",
             g->M, g->N, g->K, g->batch);
  GA_REQUIRE(!g->accumulate || g->out_dtype == GA_F32, GA_ERR_UNSUPPORTED, "ga_gemm: accumulate needs fp32 D");
  if (g->M == 0 || g->N == 0) return GA_OK;
  EpiArgs e;
  e.D = g->D; e.ldd = g->ldd; e.d_bs = g->d_bs; e.d_cs = g->d_cs > 0 ? g->d_cs : 1;
  e.out_f32 = (g->out_dtype == GA_F32); e.accumulate = g->accumulate;
  e.alpha = g->alpha;
  e.bias = g->bias; e.bias_bs = g->bias_bs; e.act = g->act; e.Z = g->Z;
  e.colscale = g->colscale; e.colscale_bs = g->colscale_bs;
  e.rowscale = g->rowscale; e.rows_per_scale = g->rows_per_scale > 0 ? g->rows_per_scale : 1;
  e.R = g->R; e.ldr = g->ldr; e.r_bs = g->r_bs;
  e.Zin = g->Zin; e.ldz = g->ldz; e.z_bs = g->z_bs; e.zmode = g->zmode; e.z_shadow = g->z_shadow;
  e.colsum = g->colsum;
  e.ln_xhat = g->ln_xhat; e.ld_xhat = g->ld_xhat; e.ln_rstd = g->ln_rstd;
  e.M = g->M; e.N = g->N;

  bool a_mn = false, b_mn = false;
  GA_REQUIRE(!g->accumulate || (!g->bias && !g->act && !g->R && !g->Zin && !g->Z), GA_ERR_UNSUPPORTED,
             "ga_gemm: accumulate mode takes no bias/act/residual/Zin");
  bool use_tc = (g->backend != GA_BACKEND_SIMT) && e.d_cs == 1 && tc_eligible(g, &a_mn, &b_mn);
  if (g->backend == GA_BACKEND_TCGEN05)
    GA_REQUIRE(use_tc, GA_ERR_ALIGN, "ga_gemm: operands not eligible for the tcgen05 path (bf16, 16B-aligned, unit stride)");
  if (use_tc) {
    if (g->backend_used) *g->backend_used = GA_BACKEND_TCGEN05;
    static int v1_env = -1;
    if (v1_env < 0) { const char* sv = getenv("GA_GEMM_V1"); v1_env = (sv && atoi(sv)) ? 1 : 0; }
    const bool wide = g->N > 32;   // N in (32, 64] also takes the persistent kernel (half of a BN=128 tile idle; these GEMMs are HBM-bound)
    const bool wide256 = (g->N >= 512) || (g->N % 256 == 0);
    // compile-time specialised epilogues for the ConvNeXt-block GEMMs (vector path: widths and pitches multiples of 4)
    const bool v4 = ((g->N & 3) == 0) && ((g->ldd & 3) == 0) && e.d_cs == 1 && g->alpha == 1.0f;
    const bool bf_out = (g->out_dtype == GA_BF16);
    const bool b4 = !g->bias || (((uintptr_t)g->bias & 15) == 0 && (g->bias_bs & 3) == 0);
    int epi = tc::EPI_GENERIC;
    if (v4 && b4 && !v1_env && wide) {
      const bool plain = !g->Z && !g->colscale && !g->rowscale && !g->R && !g->Zin && !g->accumulate;
      if (!a_mn && !b_mn && bf_out && g->act == GA_ACT_GELU && !g->colscale && !g->rowscale && !g->R && !g->Zin && !g->accumulate && g->z_shadow != 1)
        epi = g->Z ? tc::EPI_BIAS_GELU_Z : tc::EPI_BIAS_GELU;
      else if (!a_mn && !b_mn && !bf_out && g->act == GA_ACT_NONE && g->R && (g->ldr & 3) == 0 && !g->Zin && !g->accumulate &&
               (!g->Z || g->z_shadow == 1) && (!g->colscale || (((uintptr_t)g->colscale & 15) == 0 && (g->colscale_bs & 3) == 0)))
        epi = tc::EPI_RES_F32_SHADOW;
      else if (!a_mn && b_mn && bf_out && g->Zin && (g->zmode == GA_ACT_GELU || g->zmode == GA_ACT_MUL) && (g->ldz & 3) == 0 && !g->bias && !g->Z && !g->colscale &&
               !g->rowscale && !g->R && !g->accumulate)
        epi = tc::EPI_ZIN_GELU;
      else if (!a_mn && bf_out && g->act == GA_ACT_NONE && plain)
        epi = tc::EPI_PLAIN_BF16;
      else if (!a_mn && !bf_out && g->act == GA_ACT_NONE && plain && ((uintptr_t)g->D & 15) == 0 && (g->d_bs & 3) == 0)
        epi = tc::EPI_PLAIN_F32;        // fp32 outputs of bf16 GEMMs: the heads' q / k / v / proj / fc linears, stem and downsample convs
    }
    if (((g->N & 3) == 0) && ((g->ldd & 3) == 0) && e.d_cs == 1 && !v1_env && wide && a_mn && b_mn && g->accumulate &&
        ((uintptr_t)g->D & 15) == 0 && (g->d_bs & 3) == 0)
      epi = tc::EPI_ACCUM;
    if (g->ln_xhat) {
      GA_REQUIRE(epi == tc::EPI_PLAIN_BF16 && !g->bias && g->N <= 128 && g->batch == 1 && g->ln_rstd && (g->ld_xhat & 3) == 0 &&
                     ((uintptr_t)g->ln_xhat & 7) == 0,
                 GA_ERR_UNSUPPORTED, "ga_gemm: the fused LayerNorm backward needs a plain bf16 GEMM with the whole row in one tile (32 < N <= 128, N %% 4 == 0)");
      if (b_mn) return tc::launch2<128, false, true, tc::EPI_LNBWD>(g, e, st);
      return tc::launch2<128, false, false, tc::EPI_LNBWD>(g, e, st);
    }
#define GA_TC2(BN_, AM, BM_, EP) return tc::launch2<BN_, AM, BM_, EP>(g, e, st)
#define GA_TC_BN(AM, BM_, EP) { if (wide256) GA_TC2(256, AM, BM_, EP); else GA_TC2(128, AM, BM_, EP); }
    GA_REQUIRE(!g->colsum || (epi == tc::EPI_ZIN_GELU && wide && !v1_env && g->batch == 1 && (((uintptr_t)g->colsum) & 15) == 0),
               GA_ERR_UNSUPPORTED, "ga_gemm: colsum is fused only into the tcgen05 x act' epilogue (bf16, N %% 4 == 0, N > 32)");
    if (wide && !v1_env) {
      switch (epi) {
        case tc::EPI_BIAS_GELU_Z: GA_TC_BN(false, false, tc::EPI_BIAS_GELU_Z)
        case tc::EPI_BIAS_GELU: GA_TC_BN(false, false, tc::EPI_BIAS_GELU)
        case tc::EPI_RES_F32_SHADOW: GA_TC_BN(false, false, tc::EPI_RES_F32_SHADOW)
        case tc::EPI_ZIN_GELU: GA_TC_BN(false, true, tc::EPI_ZIN_GELU)
        case tc::EPI_ACCUM: GA_TC_BN(true, true, tc::EPI_ACCUM)
        case tc::EPI_PLAIN_BF16:
          if (b_mn) GA_TC_BN(false, true, tc::EPI_PLAIN_BF16) else GA_TC_BN(false, false, tc::EPI_PLAIN_BF16)
        case tc::EPI_PLAIN_F32:
          if (b_mn) GA_TC_BN(false, true, tc::EPI_PLAIN_F32) else GA_TC_BN(false, false, tc::EPI_PLAIN_F32)
        default: break;
      }
    }
#define GA_TC_CASE(AM, BM_)                                                                      \
  if (a_mn == AM && b_mn == BM_) {                                                                \
    if (!wide) return tc::launch<64, AM, BM_>(g, e, st);                                          \
    if (v1_env) return tc::launch<128, AM, BM_>(g, e, st);                                        \
    GA_TC_BN(AM, BM_, tc::EPI_GENERIC)                                                            \
  }
    GA_TC_CASE(false, false)
    GA_TC_CASE(false, true)
    GA_TC_CASE(true, false)
    GA_TC_CASE(true, true)
#undef GA_TC_CASE
  }
  GA_REQUIRE(!g->colsum, GA_ERR_UNSUPPORTED, "ga_gemm: colsum is fused only into the tcgen05 x act' epilogue");
  GA_REQUIRE(!g->ln_xhat, GA_ERR_UNSUPPORTED, "ga_gemm: the fused LayerNorm backward exists on the tcgen05 path only");
  if (g->backend_used) *g->backend_used = GA_BACKEND_SIMT;
  dim3 grid((g->M + 63) / 64, (g->N + 63) / 64, g->batch);
  GA_REQUIRE(grid.y <= 65535 && grid.z <= 65535, GA_ERR_SHAPE, "ga_gemm(simt): grid too large");
  if (g->in_dtype == GA_F32)
    gemm_simt_kernel<float><<<grid, 256, 0, st>>>((const float*)g->A, g->a_rs, g->a_cs, g->a_bs, (const float*)g->B, g->b_rs,
                                                  g->b_cs, g->b_bs, g->K, e);
  else
    gemm_simt_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)g->A, g->a_rs, g->a_cs, g->a_bs, (const bf16*)g->B, g->b_rs,
                                                 g->b_cs, g->b_bs, g->K, e);
  ga_count_launch();
  return ga_check_launch("gemm_simt");
}
