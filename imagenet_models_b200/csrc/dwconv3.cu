// K1, second generation (bf16): channels-last depthwise 7x7 + bias + LayerNorm forward, and ONE fused backward kernel
// (data gradient + residual add + bf16 shadow, weight gradient, bias gradient).  Reference: ConvNeXtBlock.conv_dw + norm
// (ga_convnext.py:92-93,100,105-106; map_convnext.py:18-19,29-31) and their autograd.
//
// What changed against dwconv.cu (ncu, profiles/r01_ncu_dwconv7.txt: FFMA2 was 27 % of 221 M warp instructions; IMAD / IADD3 /
// MOV address arithmetic 25 %, the LayerNorm phases 30 %; 11 M shared-memory bank conflicts):
//  * a CTA owns a pixel tile for a SLICE of <= 192 channels, not for all of them: every shared-memory offset is a compile-time
//    immediate (templates on the slice width and the tile width), tiles stay large at C = 384 / 688, and the LayerNorm
//    statistics of a pixel are combined across the slices' CTAs through distributed shared memory (a thread-block cluster
//    along the channel axis: st.shared::cluster of the per-slice (sum, sum of squares), one barrier.cluster);
//  * a thread (row pair, 7-pixel strip, channel pair) computes TWO output rows: every staged halo value is unpacked once and
//    feeds 14 packed FFMA2 (fma.rn.f32x2) instead of 7;
//  * 7-pixel strips put the two half-warps that straddle a 48-pair row (C = 96) 1344 bytes = 64 mod 128 apart: no conflicts;
//  * LayerNorm: one pass (sum, sum of squares), partials in the (dead) halo buffer, 4 lanes per pixel + two shuffles;
//  * backward: the SAME staged dconv halo feeds the data gradient (flipped taps) and the weight gradient
//    (dw[tap] = sum_p' dconv[p' - tap + 3] x[p'], x read without halo) -- 14 FFMA2 per staged value, one kernel instead of
//    two; a CTA walks several tiles and keeps its [50][slice] weight-gradient sums in shared memory, so global atomics drop
//    from one set per tile to one set per CTA.
// fp32 tensors and LayerNorm-with-affine calls keep using dwconv.cu.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace dw3 {

typedef unsigned long long u64;

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
      "DW3_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DW3_DONE;\n"
      "bra DW3_WAIT;\n"
      "DW3_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ u64 pack2(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// two consecutive bf16 channels at a shared-window address (+ compile-time byte offset) as packed fp32x2: one LDS, two
// byte-permutes on the ALU pipe (the FMA pipe is the one this kernel saturates; a shift would be issued there as IMAD.SHL)
__device__ __forceinline__ u64 lds_bf2(uint32_t a) {
  u64 d;
  asm volatile(
      "{\n.reg .b32 u, l, h;\n"
      "ld.shared.b32 u, [%1];\n"
      "prmt.b32 l, u, 0, 0x1044;\n"
      "prmt.b32 h, u, 0, 0x3244;\n"
      "mov.b64 %0, {l, h};\n}"
      : "=l"(d) : "r"(a));
  return d;
}
// explicit shared-window accesses (32-bit addresses, immediates after unrolling; generic pointers would cost LEA/IADD3 pairs)
__device__ __forceinline__ void sts_f2(uint32_t a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u64(uint32_t a, u64 v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ u64 lds_u64(uint32_t a) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ u64 ldg_w2(const float* p) { return __ldg(reinterpret_cast<const unsigned long long*>(p)); }

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void st_cluster_f2(uint32_t local_addr, uint32_t rank, float a, float b) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(ra), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct Geo3 {
  int B, H, W, C;
  int tiles_x, tiles_y;
  int nsl;        // channel slices (forward: cluster size along x)
  int tpc;        // backward: tiles per CTA
  long long total_tiles;
};

__host__ __device__ constexpr int pp_pad(int P) { return P + ((4 - P % 16) + 16) % 16; }   // P rounded up to 4 mod 16: conflict-free 4-lane pixel sums

// ------------------------------------------------------------------------------------------------------ forward
// grid (nsl, tiles_x * tiles_y, B), cluster (nsl, 1, 1).  CB channels per slice, TW x TH output pixels per tile.
template <int CB, int TW, int TH>
__global__ void __launch_bounds__((TH / 2) * (TW / 7) * (CB / 2), (TH / 2) * (TW / 7) * (CB / 2) <= 384 ? 2 : 1)
dwconv7_ln_fwd3_kernel(const __grid_constant__ CUtensorMap tm, const float* __restrict__ w49c, const float* __restrict__ bias,
                       bf16* __restrict__ y, float* __restrict__ rstd_out, float eps, Geo3 g) {
  constexpr int P = CB / 2, NS = TW / 7, HW_ = TW + 6, NT = (TH / 2) * NS * P, NPIX = TH * TW, PP = pp_pad(P);
  constexpr int PIXB = CB * 2, ROWB = HW_ * PIXB;                       // bytes per staged pixel / halo row
  constexpr uint32_t TILE_BYTES = (uint32_t)((TH + 6) * HW_ * PIXB);
  static_assert(NPIX * PP * 8 <= (TH + 6) * HW_ * PIXB, "LayerNorm partials must fit the halo buffer they reuse");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* bar = (uint64_t*)sm;
  const uint32_t tile_s = smem_u32(sm) + 128;                                        // halo tile; after the conv: float2 part[NPIX][PP]
  const uint32_t xstat_s = tile_s + ((TILE_BYTES + 127) & ~127u);                    // float2 [8][NPIX]: per-slice (sum, sumsq), written by every CTA of the cluster
  const uint32_t stat_s = xstat_s + 8 * NPIX * 8;                                    // float2 [NPIX]: (-mean * rstd, rstd)

  const int sl = blockIdx.x, c0 = sl * CB;
  const int tx = blockIdx.y % g.tiles_x, ty = blockIdx.y / g.tiles_x, b = blockIdx.z;
  const int x0 = tx * TW, y0 = ty * TH;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, TILE_BYTES);
    tma_load_4d(tile_s, &tm, bar, c0, x0 - 3, y0 - 3, b);
  }
  const int pr = threadIdx.x % P, rs = threadIdx.x / P;
  const int strip = rs % NS, rg = rs / NS;
  const int c = c0 + 2 * pr;
  const bool cvalid = c < g.C;
  u64 acc0[7], acc1[7];
  {
    const u64 bv = (cvalid && bias) ? ldg_w2(bias + c) : 0ull;
#pragma unroll
    for (int i = 0; i < 7; ++i) { acc0[i] = bv; acc1[i] = bv; }
  }
  const float* wq = w49c + (cvalid ? c : 0);          // walks the 49 taps of this channel pair, one row of [49][C] at a time
  const uint32_t base = tile_s + (uint32_t)((rg * 2) * ROWB + strip * 7 * PIXB + pr * 4);
  mbar_wait(bar, 0);
  u64 wa[7], wb[7];
#pragma unroll
  for (int kx = 0; kx < 7; ++kx) wb[kx] = 0ull;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (r < 7) {      // out-of-range channel pairs (tail slice) read channel 0's taps: their halo values are TMA zero fill
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) { wa[kx] = ldg_w2(wq); wq += g.C; }
    }
#pragma unroll
    for (int jx = 0; jx < 13; ++jx) {
      const u64 in = lds_bf2(base + (uint32_t)(r * ROWB + jx * PIXB));   // r, jx unrolled: the offset is an LDS immediate
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const int i = jx - kx;
        if (i >= 0 && i < 7) {
          if (r < 7) acc0[i] = ffma2(in, wa[kx], acc0[i]);      // output row 2rg   uses ky = r
          if (r >= 1) acc1[i] = ffma2(in, wb[kx], acc1[i]);     // output row 2rg+1 uses ky = r - 1
        }
      }
    }
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) wb[kx] = wa[kx];
  }
  __syncthreads();                 // every thread is done with the halo: its buffer now holds the LayerNorm partials
  // ---- LayerNorm statistics of the slice: (sum, sum of squares) over this slice's channels, per pixel
  {
    const uint32_t p0 = tile_s + (uint32_t)((((rg * 2) * TW + strip * 7) * PP + pr) * 8);
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float a0 = lo2(acc0[i]), a1 = hi2(acc0[i]), b0 = lo2(acc1[i]), b1 = hi2(acc1[i]);
      sts_f2(p0 + (uint32_t)(i * PP * 8), a0 + a1, fmaf(a0, a0, a1 * a1));
      sts_f2(p0 + (uint32_t)((TW + i) * PP * 8), b0 + b1, fmaf(b0, b0, b1 * b1));
    }
  }
  __syncthreads();
  const float invC = 1.f / (float)g.C;
  const uint32_t my_rank = g.nsl > 1 ? cluster_rank() : 0u;
  for (int item = threadIdx.x; item < NPIX * 4; item += NT) {
    const int q = item >> 2, s4 = item & 3;
    float s = 0.f, qq = 0.f;
    const uint32_t pa = tile_s + (uint32_t)((q * PP + s4) * 8);
#pragma unroll
    for (int k = 0; k < (P + 3) / 4; ++k) {
      if (s4 + 4 * k < P) { const float2 v = lds_f2(pa + (uint32_t)(k * 32)); s += v.x; qq += v.y; }
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1); qq += __shfl_xor_sync(0xffffffffu, qq, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2); qq += __shfl_xor_sync(0xffffffffu, qq, 2);
    if (s4 == 0) {
      if (g.nsl > 1) {
        const uint32_t la = xstat_s + (uint32_t)((my_rank * NPIX + q) * 8);
        for (int rk = 0; rk < g.nsl; ++rk) st_cluster_f2(la, (uint32_t)rk, s, qq);
      } else {
        const float mean = s * invC;
        const float r_ = rsqrtf(fmaxf(fmaf(-mean, mean, qq * invC), 0.f) + eps);
        sts_f2(stat_s + (uint32_t)(q * 8), -mean * r_, r_);
        const int oy = y0 + q / TW, ox = x0 + q % TW;
        if (rstd_out && oy < g.H && ox < g.W) rstd_out[((size_t)b * g.H + oy) * g.W + ox] = r_;
      }
    }
  }
  if (g.nsl > 1) {
    cluster_sync();
    for (int q = threadIdx.x; q < NPIX; q += NT) {
      float s = 0.f, qq = 0.f;
      for (int rk = 0; rk < g.nsl; ++rk) { const float2 v = lds_f2(xstat_s + (uint32_t)((rk * NPIX + q) * 8)); s += v.x; qq += v.y; }
      const float mean = s * invC;
      const float r_ = rsqrtf(fmaxf(fmaf(-mean, mean, qq * invC), 0.f) + eps);
      sts_f2(stat_s + (uint32_t)(q * 8), -mean * r_, r_);
      const int oy = y0 + q / TW, ox = x0 + q % TW;
      if (sl == 0 && rstd_out && oy < g.H && ox < g.W) rstd_out[((size_t)b * g.H + oy) * g.W + ox] = r_;
    }
  }
  __syncthreads();
  // ---- normalise and store xhat (the affine is folded into fc1 by the caller)
  if (cvalid) {
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const int row = rg * 2 + o, oy = y0 + row;
      bf16* yrow = y + (((size_t)b * g.H + oy) * g.W + (x0 + strip * 7)) * g.C + c;
      const uint32_t sa = stat_s + (uint32_t)((row * TW + strip * 7) * 8);
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const float2 st = lds_f2(sa + (uint32_t)(i * 8));
        const u64 a = o == 0 ? acc0[i] : acc1[i];
        if (oy < g.H && (x0 + strip * 7 + i) < g.W)
          *reinterpret_cast<uint32_t*>(yrow + (size_t)i * g.C) = pack_bf16(fmaf(lo2(a), st.y, st.x), fmaf(hi2(a), st.y, st.x));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------ backward
// grid (ceil(total_tiles / tpc), nsl).  Thread (row, 7-pixel strip, channel pair), one output row each.
//   dx[p]   = sum_t dconv[p + 3 - t] w[t] + dres[p]            (fp32 or bf16 stream gradient, plus an optional bf16 shadow)
//   dw[t]   = sum_p x[p] dconv[p + 3 - t]      db = sum_p dconv[p]
// halo offset (ky, kx) of the staged dconv tile pairs with tap (6 - ky, 6 - kx) for both sums.
template <int CB, int TW, int TH, typename TR>
__global__ void __launch_bounds__(TH * (TW / 7) * (CB / 2), TH * (TW / 7) * (CB / 2) <= 384 ? 2 : 1)
dwconv7_bwd3_kernel(const __grid_constant__ CUtensorMap tmd, const __grid_constant__ CUtensorMap tmx, const float* __restrict__ w49c,
                    const TR* __restrict__ dres, TR* __restrict__ dx, bf16* __restrict__ dxs, const float* __restrict__ dxs_scale,
                    float* __restrict__ partial, int nparts,
                    Geo3 g) {
  constexpr int P = CB / 2, NS = TW / 7, HW_ = TW + 6, RS = TH * NS, NT = RS * P;
  constexpr int PIXB = CB * 2, ROWB = HW_ * PIXB;
  constexpr uint32_t HALO_BYTES = (uint32_t)((TH + 6) * HW_ * PIXB), CEN_BYTES = (uint32_t)(TH * TW * PIXB);
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* bar = (uint64_t*)sm;
  const uint32_t halo_s = smem_u32(sm) + 128;
  const uint32_t cen_s = halo_s + ((HALO_BYTES + 127) & ~127u);
  const uint32_t fold_s = cen_s + ((CEN_BYTES + 127) & ~127u);          // u64 [7][RS][P]
  const uint32_t accum_s = fold_s + 7 * RS * P * 8;                     // float2 [50][P]: this CTA's weight / bias gradient sums
  const uint32_t wsm_s = accum_s + 50 * P * 8;                          // float2 [49][P]: the slice's taps (ncu: the per-ky LDG of the
                                                                         // taps was 28 % long-scoreboard stall; the CTA walks several tiles)

  const int sl = blockIdx.y, c0 = sl * CB;
  const int pr = threadIdx.x % P, rs = threadIdx.x / P;
  const int strip = rs % NS, row = rs / NS;
  const int c = c0 + 2 * pr;
  const bool cvalid = c < g.C;
  for (int i = threadIdx.x; i < 50 * P; i += NT) sts_f2(accum_s + (uint32_t)(i * 8), 0.f, 0.f);
  for (int i = threadIdx.x; i < 49 * P; i += NT) {
    const int tap = i / P, cc = c0 + 2 * (i - tap * P);
    const u64 v = cc < g.C ? ldg_w2(w49c + (size_t)tap * g.C + cc) : 0ull;
    sts_u64(wsm_s + (uint32_t)(i * 8), v);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t wbase = wsm_s + (uint32_t)(pr * 8);
  const uint32_t hbase = halo_s + (uint32_t)(row * ROWB + strip * 7 * PIXB + pr * 4);
  const uint32_t cbase = cen_s + (uint32_t)((row * TW + strip * 7) * PIXB + pr * 4);
  const long long t_begin = (long long)blockIdx.x * g.tpc;
  const long long t_end = t_begin + g.tpc < g.total_tiles ? t_begin + g.tpc : g.total_tiles;
  uint32_t phase = 0;
  auto issue_tile = [&](long long t) {               // one elected thread: stage the dconv halo and the x centre of tile t
    const int tx = (int)(t % g.tiles_x);
    const int ty = (int)((t / g.tiles_x) % g.tiles_y);
    const int b = (int)(t / ((long long)g.tiles_x * g.tiles_y));
    mbar_expect_tx(bar, HALO_BYTES + CEN_BYTES);
    tma_load_4d(halo_s, &tmd, bar, c0, tx * TW - 3, ty * TH - 3, b);
    tma_load_4d(cen_s, &tmx, bar, c0, tx * TW, ty * TH, b);
  };
  if (threadIdx.x == 0 && t_begin < t_end) issue_tile(t_begin);
  for (long long t = t_begin; t < t_end; ++t) {
    const int tx = (int)(t % g.tiles_x);
    const int ty = (int)((t / g.tiles_x) % g.tiles_y);
    const int b = (int)(t / ((long long)g.tiles_x * g.tiles_y));
    const int x0 = tx * TW, y0 = ty * TH;
    // the residual-stream gradient of this thread's 7 pixels is only needed in the epilogue: pull its lines into L2 now
    // (ncu: 19 % long-scoreboard stall, most of it on these loads)
    if (dres && cvalid && y0 + row < g.H) {
      const TR* rp = dres + (((size_t)b * g.H + (y0 + row)) * g.W + (x0 + strip * 7)) * g.C + c;
#pragma unroll
      for (int i = 0; i < 7; ++i)
        if (x0 + strip * 7 + i < g.W) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (size_t)i * g.C));
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    u64 xc[7], acc[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) acc[i] = 0ull;
#pragma unroll
    for (int i = 0; i < 7; ++i) xc[i] = lds_bf2(cbase + (uint32_t)(i * PIXB));
    float ds0 = 0.f, ds1 = 0.f;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
      u64 wv[7], a7[7];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        wv[kx] = lds_u64(wbase + (uint32_t)((48 - (ky * 7 + kx)) * P * 8));      // flipped taps
        a7[kx] = 0ull;
      }
#pragma unroll
      for (int jx = 0; jx < 13; ++jx) {
        const u64 in = lds_bf2(hbase + (uint32_t)(ky * ROWB + jx * PIXB));
        if (ky == 3 && jx >= 3 && jx < 10) { ds0 += lo2(in); ds1 += hi2(in); }       // the tile's own dconv pixels
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int i = jx - kx;
          if (i >= 0 && i < 7) {
            acc[i] = ffma2(in, wv[kx], acc[i]);
            a7[kx] = ffma2(in, xc[i], a7[kx]);
          }
        }
      }
      // fold the 7 tap sums of this ky over the CTA's rows and strips, then into the CTA's running sums
      __syncthreads();                       // previous round's fold buffer fully consumed
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) sts_u64(fold_s + (uint32_t)(((kx * RS + rs) * P + pr) * 8), a7[kx]);
      __syncthreads();
      for (int idx = threadIdx.x; idx < 7 * P; idx += NT) {
        const int kx = idx / P, pp = idx - kx * P;
        float s0 = 0.f, s1 = 0.f;
        const uint32_t fa = fold_s + (uint32_t)((kx * RS * P + pp) * 8);
#pragma unroll
        for (int r = 0; r < RS; ++r) { const u64 v = lds_u64(fa + (uint32_t)(r * P * 8)); s0 += lo2(v); s1 += hi2(v); }
        const uint32_t aa = accum_s + (uint32_t)(((48 - (ky * 7 + kx)) * P + pp) * 8);
        const float2 o = lds_f2(aa);
        sts_f2(aa, o.x + s0, o.y + s1);
      }
    }
    // bias gradient: one more fold round
    __syncthreads();
    sts_u64(fold_s + (uint32_t)((rs * P + pr) * 8), pack2(ds0, ds1));
    __syncthreads();
    for (int pp = threadIdx.x; pp < P; pp += NT) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int r = 0; r < RS; ++r) { const u64 v = lds_u64(fold_s + (uint32_t)((r * P + pp) * 8)); s0 += lo2(v); s1 += hi2(v); }
      const uint32_t aa = accum_s + (uint32_t)((49 * P + pp) * 8);
      const float2 o = lds_f2(aa);
      sts_f2(aa, o.x + s0, o.y + s1);
    }
    // every thread passed the barrier above after its last read of the staged tiles: the next tile's TMA can overlap the epilogue
    if (threadIdx.x == 0 && t + 1 < t_end) issue_tile(t + 1);
    // data gradient + residual, stream dtype TR, optional bf16 shadow
    const int oy = y0 + row;
    if (cvalid && oy < g.H && dx) {
      const size_t off0 = (((size_t)b * g.H + oy) * g.W + (x0 + strip * 7)) * g.C + c;
      // the bf16 shadow may carry the consumer's DropPath factor of this image (it is the GEMM operand of the previous block's
      // backward, whose branch gradient is ps[b] * dy): saves that block a full scale-rows pass
      const float ssc = dxs_scale ? dxs_scale[b] : 1.f;
      float r0[7], r1[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        r0[i] = r1[i] = 0.f;
        if (dres && x0 + strip * 7 + i < g.W) {
          if (sizeof(TR) == 4) { const float2 v = *reinterpret_cast<const float2*>(dres + off0 + (size_t)i * g.C); r0[i] = v.x; r1[i] = v.y; }
          else { const uint32_t u = *reinterpret_cast<const uint32_t*>(dres + off0 + (size_t)i * g.C); r0[i] = bf16lo(u); r1[i] = bf16hi(u); }
        }
      }
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        if (x0 + strip * 7 + i < g.W) {
          const float v0 = lo2(acc[i]) + r0[i], v1 = hi2(acc[i]) + r1[i];
          if (sizeof(TR) == 4) *reinterpret_cast<float2*>(dx + off0 + (size_t)i * g.C) = make_float2(v0, v1);
          else *reinterpret_cast<uint32_t*>(dx + off0 + (size_t)i * g.C) = pack_bf16(v0, v1);
          if (dxs) *reinterpret_cast<uint32_t*>(dxs + off0 + (size_t)i * g.C) = pack_bf16(v0 * ssc, v1 * ssc);
        }
      }
    }
  }
  __syncthreads();         // accum updates of the last tile ordered before the flush
  // flush the CTA's sums: [50][C] slot (blockIdx.x % nparts)
  if (partial) {
    float* slot = partial + (size_t)(blockIdx.x % nparts) * 50 * g.C;
    for (int idx = threadIdx.x; idx < 50 * P; idx += NT) {
      const int tap = idx / P, pp = idx - tap * P;
      const int cc = c0 + 2 * pp;
      if (cc < g.C) {
        const float2 v = lds_f2(accum_s + (uint32_t)(idx * 8));
        atomicAdd(slot + (size_t)tap * g.C + cc, v.x);
        atomicAdd(slot + (size_t)tap * g.C + cc + 1, v.y);
      }
    }
  }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static int make_map(const void* x, int B, int H, int W, int C, int cb, int bw, int bh, CUtensorMap* out) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {(uint32_t)cb, (uint32_t)bw, (uint32_t)bh, 1};
  return ga_tensor_map(out, GA_BF16, 4, x, dims, strides, box, 0);
}

static int pick_cb(int C) {          // forward: slices are coupled through the cluster LayerNorm, keep them wide (cluster size <= 8)
  if (C == 96) return 96;
  if (C % 128 == 0 && C % 192 != 0) return 128;
  return 192;
}
// backward: slices are independent.  96-channel slices keep 384-thread CTAs, two per SM (measured: 7.5 ns per 10^6 elements at
// 56x56x96 against 10.0 with 192-channel slices and one 768-thread CTA per SM at 28x28x192)
static int pick_cb_bwd(int C) {
  if (C % 96 == 0) return 96;
  if (C % 128 == 0) return 128;
  return 96;
}

template <int CB, int TW, int TH>
static int launch_fwd(const void* x, const float* w49c, const float* bias, void* y, float* rstd, int B, int H, int W, int C, float eps,
                      cudaStream_t st) {
  constexpr int NT = (TH / 2) * (TW / 7) * (CB / 2), NPIX = TH * TW;
  constexpr size_t TILE = (size_t)(TH + 6) * (TW + 6) * CB * 2;
  constexpr size_t SMEM = 128 + 128 + ((TILE + 127) & ~(size_t)127) + (size_t)9 * NPIX * 8;
  Geo3 g;
  g.B = B; g.H = H; g.W = W; g.C = C;
  g.tiles_x = (W + TW - 1) / TW; g.tiles_y = (H + TH - 1) / TH;
  g.nsl = (C + CB - 1) / CB; g.tpc = 1; g.total_tiles = (long long)g.tiles_x * g.tiles_y * B;
  GA_REQUIRE(g.nsl <= 8, GA_ERR_UNSUPPORTED, "dwconv7 v3: C=%d needs more than 8 channel slices", C);
  GA_REQUIRE(g.tiles_x * g.tiles_y <= 65535 && B <= 65535, GA_ERR_SHAPE, "dwconv7 v3: grid too large");
  CUtensorMap tm;
  int rc = make_map(x, B, H, W, C, CB, TW + 6, TH + 6, &tm);
  if (rc) return rc;
  auto k = dwconv7_ln_fwd3_kernel<CB, TW, TH>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g.nsl, g.tiles_x * g.tiles_y, B);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = g.nsl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, tm, w49c, bias, (bf16*)y, rstd, eps, g);
  ga_count_launch();
  if (e != cudaSuccess) { ga_set_error("dwconv7_ln_fwd3: launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return GA_ERR_LAUNCH; }
  return ga_check_launch("dwconv7_ln_fwd3");
}

template <int CB, int TW, int TH, typename TR>
static int launch_bwd(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dxs, const float* dxs_scale, float* partial,
                      int nparts, int B, int H, int W, int C, cudaStream_t st) {
  constexpr int P = CB / 2, RS = TH * (TW / 7), NT = RS * P;
  constexpr size_t HALO = (size_t)(TH + 6) * (TW + 6) * CB * 2, CEN = (size_t)TH * TW * CB * 2;
  constexpr size_t SMEM = 128 + 128 + ((HALO + 127) & ~(size_t)127) + ((CEN + 127) & ~(size_t)127) + (size_t)7 * RS * P * 8 + (size_t)99 * P * 8;
  static_assert(SMEM <= 227 * 1024, "dwconv7 bwd3 tile does not fit shared memory");
  Geo3 g;
  g.B = B; g.H = H; g.W = W; g.C = C;
  g.tiles_x = (W + TW - 1) / TW; g.tiles_y = (H + TH - 1) / TH;
  g.nsl = (C + CB - 1) / CB;
  g.total_tiles = (long long)g.tiles_x * g.tiles_y * B;
  const int per_sm = NT <= 384 ? 2 : 1;
  long long want = (long long)ga_num_sms() * per_sm * 4 / g.nsl;           // ~4 CTAs per resident slot
  if (want < 1) want = 1;
  long long tpc = (g.total_tiles + want - 1) / want;
  if (tpc < 1) tpc = 1;
  if (tpc > 16) tpc = 16;
  g.tpc = (int)tpc;
  const long long gx = (g.total_tiles + tpc - 1) / tpc;
  GA_REQUIRE(gx <= 0x7fffffffLL && g.nsl <= 65535, GA_ERR_SHAPE, "dwconv7 bwd3: grid too large");
  CUtensorMap tmd, tmx;
  int rc = make_map(dconv, B, H, W, C, CB, TW + 6, TH + 6, &tmd);
  if (rc) return rc;
  rc = make_map(x, B, H, W, C, CB, TW, TH, &tmx);
  if (rc) return rc;
  auto k = dwconv7_bwd3_kernel<CB, TW, TH, TR>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  k<<<dim3((unsigned)gx, g.nsl), NT, SMEM, st>>>(tmd, tmx, w49c, (const TR*)dres, (TR*)dx, (bf16*)dxs, dxs_scale, partial, nparts, g);
  ga_count_launch();
  return ga_check_launch("dwconv7_bwd3");
}

}  // namespace dw3

// bf16 forward: xhat = LN(conv7x7(x) + bias) without affine, rstd saved.  Returns GA_ERR_UNSUPPORTED for shapes it does not take.
int ga_dwconv7_ln_fwd_v3(const void* x, const float* w49c, const float* bias, void* y, float* rstd, int B, int H, int W, int C,
                         float eps, cudaStream_t st) {
  if (C % 8 || C < 32) return GA_ERR_UNSUPPORTED;
  const int cb = dw3::pick_cb(C);
  const bool narrow = W <= 7;
#define DW3_FWD(CB_)                                                                                                   \
  if (cb == CB_) {                                                                                                     \
    if (narrow) return dw3::launch_fwd<CB_, 7, 8>(x, w49c, bias, y, rstd, B, H, W, C, eps, st);                         \
    if (CB_ == 96) return dw3::launch_fwd<CB_, 14, 8>(x, w49c, bias, y, rstd, B, H, W, C, eps, st);                     \
    return dw3::launch_fwd<CB_, 14, 4>(x, w49c, bias, y, rstd, B, H, W, C, eps, st);                                    \
  }
  DW3_FWD(96) DW3_FWD(128) DW3_FWD(192)
#undef DW3_FWD
  return GA_ERR_UNSUPPORTED;
}

int ga_dwconv7_bwd_v3(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dxs, const float* dxs_scale, float* partial,
                      int nparts, int B, int H, int W, int C, int res_dtype, cudaStream_t st) {
  if (C % 8 || C < 32) return GA_ERR_UNSUPPORTED;
  const int cb = dw3::pick_cb_bwd(C);
  const bool narrow = W <= 7;
#define DW3_BWD(CB_, TR_)                                                                                                             \
  if (cb == CB_) {                                                                                                                    \
    if (narrow) return dw3::launch_bwd<CB_, 7, 8, TR_>(dconv, x, dres, w49c, dx, dxs, dxs_scale, partial, nparts, B, H, W, C, st);                \
    return dw3::launch_bwd<CB_, 14, 4, TR_>(dconv, x, dres, w49c, dx, dxs, dxs_scale, partial, nparts, B, H, W, C, st);                           \
  }
  if (res_dtype == GA_F32) { DW3_BWD(96, float) DW3_BWD(128, float) }
  else { DW3_BWD(96, bf16) DW3_BWD(128, bf16) }
#undef DW3_BWD
  return GA_ERR_UNSUPPORTED;
}
