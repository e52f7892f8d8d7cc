// K1: channels-last depthwise 7x7 convolution fused with LayerNorm (forward), its data gradient (flipped taps,
// fused with the residual add) and its weight gradient.   Reference: ConvNeXtBlock.conv_dw + norm
// (ga_convnext.py:92-93,100,105-106; map_convnext.py:18-19,29-31) and their autograd.
//
// Layout: x is NHWC.  One CTA owns a TH x TW pixel tile for ALL channels (LayerNorm needs the whole row).
//  * The (TH+6) x (TW+6) x C input halo is staged in shared memory by TMA (4-D tensor map; out-of-bounds = zero =
//    the conv padding): one elected thread, one mbarrier.  Two CTAs are resident per SM, so one CTA's halo load
//    overlaps the other's FMA phase.
//  * Thread (row, channel PAIR) slides the 7-tap window along its output row: the halo row streams through one
//    LDS at a time (32-bit shared-window addresses, no generic pointers) and every tap is ONE packed FFMA2
//    (fma.rn.f32x2, two channels per issue slot) into TW fp32x2 accumulators held in registers.
//  * LayerNorm statistics: per-thread pair sums -> shared memory -> one warp per pixel (two rounds: mean, then
//    centred variance), independent of how rows map onto warps (no idle lanes at C=96).
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace dw {

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
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// packed fp32x2 arithmetic (Blackwell FFMA2: two FMAs per issue slot)
__device__ __forceinline__ u64 pack2(float lo, float hi) { return ((u64)__float_as_uint(hi) << 32) | (u64)__float_as_uint(lo); }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// two consecutive channels from shared memory (32-bit shared-window address) as packed fp32x2
template <typename T> __device__ __forceinline__ u64 lds_pair(uint32_t a);
template <> __device__ __forceinline__ u64 lds_pair<float>(uint32_t a) {
  u64 v;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
template <> __device__ __forceinline__ u64 lds_pair<bf16>(uint32_t a) {
  uint32_t u;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(u) : "r"(a));
  return ((u64)(u & 0xffff0000u) << 32) | (u64)(u << 16);
}
template <typename T> __device__ __forceinline__ u64 ldg_pair(const T* p);
template <> __device__ __forceinline__ u64 ldg_pair<float>(const float* p) { return *reinterpret_cast<const u64*>(p); }
template <> __device__ __forceinline__ u64 ldg_pair<bf16>(const bf16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return ((u64)(u & 0xffff0000u) << 32) | (u64)(u << 16);
}
template <typename T> __device__ __forceinline__ void stg_pair(T* p, float lo, float hi);
template <> __device__ __forceinline__ void stg_pair<float>(float* p, float lo, float hi) { *reinterpret_cast<float2*>(p) = make_float2(lo, hi); }
template <> __device__ __forceinline__ void stg_pair<bf16>(bf16* p, float lo, float hi) { *reinterpret_cast<uint32_t*>(p) = pack_bf16(lo, hi); }

struct Geo {
  int B, H, W, C;
  int TH;            // output rows per CTA tile
  int TR;            // rows computed per pass (threads = TR * P); the CTA makes ceil(TH/TR) passes over its halo
  int tiles_x, tiles_y;
  int cbox, nbox;    // channel box of the TMA load and number of boxes (nbox*cbox >= C)
  int box_stride;    // elements between consecutive channel boxes in smem (128-byte aligned)
  int P;             // channel pairs (C/2)
  int c0, Cfull;     // weight-gradient channel slice: this launch covers channels [c0, c0 + C) of Cfull
};

// stage the halo tile: one thread arms the barrier and issues one 4-D box per channel box
template <typename T, int TW>
__device__ __forceinline__ void load_halo(uint32_t tile_s, uint64_t* bar, const CUtensorMap* map, const Geo& g, int b, int y0, int x0) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t box_bytes = (uint32_t)((g.TH + 6) * (TW + 6) * g.cbox * sizeof(T));
    mbar_expect_tx(bar, box_bytes * g.nbox);
    for (int j = 0; j < g.nbox; ++j)
      tma_load_4d(tile_s + (uint32_t)((size_t)j * g.box_stride * sizeof(T)), map, bar, g.c0 + j * g.cbox, x0 - 3, y0 - 3, b);
  }
  mbar_wait(bar, 0);
}

// MODE 0: forward  y = LN(conv(x) + bias)  (xhat, optional affine), rstd saved
// MODE 1: dgrad    y = corr(x = dconv, flipped taps) + res
template <typename T, typename TO, int TW, int MODE, bool BIG>
__global__ void __launch_bounds__(BIG ? 512 : (MODE == 0 ? 384 : 448), BIG ? 1 : 2) dwconv7_kernel(const __grid_constant__ CUtensorMap tm, const float* __restrict__ w49c,
                                                         const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                         const float* __restrict__ ln_b, const TO* __restrict__ res,
                                                         TO* __restrict__ y, float* __restrict__ rstd_out, float eps, Geo g) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bar = (uint64_t*)sm;
  const uint32_t tile_s = smem_u32(sm + 128);
  const size_t tile_bytes = (size_t)g.nbox * g.box_stride * sizeof(T);
  float* part = (float*)(sm + 128 + tile_bytes);                   // [TR*TW][P]
  float* stat = part + (size_t)g.TR * TW * g.P;                    // [2][TR*TW]

  const int tile_id = blockIdx.x;
  const int tx = tile_id % g.tiles_x, ty = tile_id / g.tiles_x;
  const int b = blockIdx.y;
  const int x0 = tx * TW, y0 = ty * g.TH;

  load_halo<T, TW>(tile_s, bar, &tm, g, b, y0, x0);

  const int trow = threadIdx.x / g.P;            // row slot of this thread inside a pass
  const int pr = threadIdx.x - trow * g.P;       // channel pair
  const int c = pr * 2;
#pragma unroll 1
  for (int r0 = 0; r0 < g.TH; r0 += g.TR) {
  const int row = r0 + trow;                     // output row inside the tile
  const bool active = (trow < g.TR) && (row < g.TH);

  u64 acc[TW];
  {
    u64 init = 0ull;
    if (MODE != 1 && bias && active) init = *reinterpret_cast<const u64*>(bias + g.c0 + c);
#pragma unroll
    for (int i = 0; i < TW; ++i) acc[i] = init;
  }
  if (active) {
    const int box = c / g.cbox, cc = c - box * g.cbox;
    const uint32_t cbs = (uint32_t)(g.cbox * sizeof(T));
    uint32_t rowaddr = tile_s + (uint32_t)(((size_t)box * g.box_stride + (size_t)row * (TW + 6) * g.cbox + cc) * sizeof(T));
    const uint32_t row_pitch = (uint32_t)(TW + 6) * cbs;
    // taps walk forward (conv) or backward (data gradient = correlation with the flipped kernel)
    const float* wk = w49c + g.c0 + c + (MODE != 1 ? 0 : 48 * g.Cfull);
    const int wstep = (MODE != 1) ? g.Cfull : -g.Cfull;
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
      u64 wv[7];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) wv[kx] = __ldg(reinterpret_cast<const u64*>(wk + kx * wstep));
      wk += 7 * wstep;
      uint32_t a = rowaddr;
      rowaddr += row_pitch;
      // stream the TW+6 inputs of this halo row: one packed value live at a time, 7 FFMA2 each
#pragma unroll
      for (int jx = 0; jx < TW + 6; ++jx) {
        const u64 in = lds_pair<T>(a);
        a += cbs;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int i = jx - kx;
          if (i >= 0 && i < TW) acc[i] = ffma2(in, wv[kx], acc[i]);
        }
      }
    }
  }

  const int oy = y0 + row;
  if (MODE != 0) {       // 1: data gradient (+ residual); 2: conv + bias only (channel-sliced forward, LayerNorm done separately)
    if (active && oy < g.H) {
      const size_t off0 = (((size_t)b * g.H + oy) * g.W + x0) * g.Cfull + g.c0 + c;
      u64 r[TW];
      // all TW residual loads are issued back to back (one latency per row, not one per pixel), then add + store
#pragma unroll
      for (int i = 0; i < TW; ++i) r[i] = (res && x0 + i < g.W) ? ldg_pair<TO>(res + off0 + (size_t)i * g.Cfull) : 0ull;
      T* ysh = reinterpret_cast<T*>(rstd_out);      // MODE 1: the rstd slot carries an optional compute-dtype copy of dx
#pragma unroll
      for (int i = 0; i < TW; ++i) {
        if (x0 + i < g.W) {
          const float v0 = lo2(acc[i]) + lo2(r[i]), v1 = hi2(acc[i]) + hi2(r[i]);
          stg_pair<TO>(y + off0 + (size_t)i * g.Cfull, v0, v1);
          if (ysh) stg_pair<T>(ysh + off0 + (size_t)i * g.Cfull, v0, v1);
        }
      }
    }
    continue;
  }

  // ---- LayerNorm over channels (per pixel): round 1 mean, round 2 centred variance.
  // Each thread folds its two channels, adjacent pair-lanes fold once more by shuffle (P is even, so lane^1 is the same
  // row's neighbouring pair), even pairs write part[pair/2][pixel]; then ONE THREAD PER PIXEL sums the P/2 partials with
  // loads (nsl adjacent lanes per pixel when the pass has fewer pixels than threads, folded by shuffle).
  const float invC = 1.f / (float)g.C;
  const int npix = g.TR * TW;
  const int PH = g.P >> 1;
  const bool writer = active && !(pr & 1);
  // reduction mapping: nsl (power of two) adjacent lanes share one pixel and split its P/2 partials
  int nsl = 1;
  while (nsl < 32 && nsl * 2 * npix <= (int)blockDim.x && nsl * 2 <= PH) nsl <<= 1;
  const int rp = (int)threadIdx.x / nsl, rslice = (int)threadIdx.x & (nsl - 1);
  float* prow = part + (size_t)(pr >> 1) * npix + trow * TW;
#pragma unroll
  for (int i = 0; i < TW; ++i) {
    float v = lo2(acc[i]) + hi2(acc[i]);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (writer) prow[i] = v;
  }
  __syncthreads();
  {
    float a0 = 0.f, a1 = 0.f;
    if (rp < npix) {
      int k = rslice;
      for (; k + nsl < PH; k += 2 * nsl) { a0 += part[(size_t)k * npix + rp]; a1 += part[(size_t)(k + nsl) * npix + rp]; }
      if (k < PH) a0 += part[(size_t)k * npix + rp];
    }
    a0 += a1;
    for (int o = nsl >> 1; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    if (rp < npix && rslice == 0) stat[rp] = a0 * invC;
  }
  __syncthreads();
  float mean[TW];
#pragma unroll
  for (int i = 0; i < TW; ++i) {
    mean[i] = active ? stat[trow * TW + i] : 0.f;
    const float d0 = lo2(acc[i]) - mean[i], d1 = hi2(acc[i]) - mean[i];
    float v = d0 * d0 + d1 * d1;
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if (writer) prow[i] = v;
  }
  __syncthreads();
  {
    float a0 = 0.f, a1 = 0.f;
    if (rp < npix) {
      int k = rslice;
      for (; k + nsl < PH; k += 2 * nsl) { a0 += part[(size_t)k * npix + rp]; a1 += part[(size_t)(k + nsl) * npix + rp]; }
      if (k < PH) a0 += part[(size_t)k * npix + rp];
    }
    a0 += a1;
    for (int o = nsl >> 1; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    if (rp < npix && rslice == 0) {
      const float r = rsqrtf(a0 * invC + eps);
      stat[npix + rp] = r;
      const int py = y0 + r0 + rp / TW, px = x0 + rp % TW;
      if (rstd_out && r0 + rp / TW < g.TH && py < g.H && px < g.W) rstd_out[((size_t)b * g.H + py) * g.W + px] = r;
    }
  }
  __syncthreads();
  if (active && oy < g.H) {
    float lw0 = 1.f, lw1 = 1.f, lb0 = 0.f, lb1 = 0.f;
    if (ln_w) { lw0 = ln_w[c]; lw1 = ln_w[c + 1]; lb0 = ln_b[c]; lb1 = ln_b[c + 1]; }
#pragma unroll
    for (int i = 0; i < TW; ++i) {
      const int ox = x0 + i;
      if (ox < g.W) {
        const float r = stat[npix + trow * TW + i];
        stg_pair<TO>(y + (((size_t)b * g.H + oy) * g.W + ox) * g.C + c, (lo2(acc[i]) - mean[i]) * r * lw0 + lb0,
                     (hi2(acc[i]) - mean[i]) * r * lw1 + lb1);
      }
    }
  }
  __syncthreads();                       // `part` / `stat` are reused by the next pass
  }  // pass over row groups
}

// weight gradient: dw[tap][c] += sum_pixels dconv(p, c) * x(p + tap - 3, c);  dbias[c] += sum dconv(p, c)
// Thread (row, channel pair): for each ky the halo row streams through, 7 packed accumulators; rows are folded through
// shared memory once per ky, then one atomicAdd per (tap, channel) per CTA into partial slot (blockIdx % nparts).
template <typename T, int TW, bool BIG>
__global__ void __launch_bounds__(BIG ? 512 : 448, BIG ? 1 : 2) dwconv7_wgrad_kernel(const __grid_constant__ CUtensorMap tmx,
                                                               const __grid_constant__ CUtensorMap tmd, int dbox_stride,
                                                               float* __restrict__ partial, int nparts, Geo g) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bar = (uint64_t*)sm;
  const uint32_t tile_s = smem_u32(sm + 128);
  const size_t tile_bytes = (size_t)g.nbox * g.box_stride * sizeof(T);
  const uint32_t dtile_s = tile_s + (uint32_t)tile_bytes;          // dconv tile [box][TH][TW][cbox], no halo
  const size_t dtile_bytes = (size_t)g.nbox * dbox_stride * sizeof(T);
  u64* fold = (u64*)(sm + 128 + tile_bytes + dtile_bytes);         // [7][TR][P] packed pairs

  const int tile_id = blockIdx.x;
  const int tx = tile_id % g.tiles_x, ty = tile_id / g.tiles_x;
  const int b = blockIdx.y;
  const int x0 = tx * TW, y0 = ty * g.TH;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t hb = (uint32_t)((g.TH + 6) * (TW + 6) * g.cbox * sizeof(T)), db = (uint32_t)(g.TH * TW * g.cbox * sizeof(T));
    mbar_expect_tx(bar, (hb + db) * g.nbox);
    for (int j = 0; j < g.nbox; ++j) {
      tma_load_4d(tile_s + (uint32_t)((size_t)j * g.box_stride * sizeof(T)), &tmx, bar, g.c0 + j * g.cbox, x0 - 3, y0 - 3, b);
      tma_load_4d(dtile_s + (uint32_t)((size_t)j * dbox_stride * sizeof(T)), &tmd, bar, g.c0 + j * g.cbox, x0, y0, b);
    }
  }
  mbar_wait(bar, 0);

  const int trow = threadIdx.x / g.P;
  const int pr = threadIdx.x - trow * g.P;
  const int c = pr * 2;
  float* slot = partial + (size_t)((blockIdx.x + blockIdx.y * gridDim.x) % nparts) * 50 * g.Cfull + g.c0;
  const int box = c / g.cbox, cc = c - box * g.cbox;
  const uint32_t cbs = (uint32_t)(g.cbox * sizeof(T));
  const uint32_t row_pitch = (uint32_t)(TW + 6) * cbs;
  const uint32_t base = tile_s + (uint32_t)(((size_t)box * g.box_stride + cc) * sizeof(T));
  const uint32_t dbase = dtile_s + (uint32_t)(((size_t)box * dbox_stride + cc) * sizeof(T));
  float ds0 = 0.f, ds1 = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
    u64 a7[7];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) a7[kx] = 0ull;
    // every row group of the tile accumulates into the same 7 packed registers: one fold per ky, not per pass
#pragma unroll 1
    for (int r0 = 0; r0 < g.TH; r0 += g.TR) {
      const int row = r0 + trow;
      const int oy = y0 + row;
      if (!((trow < g.TR) && (row < g.TH) && oy < g.H)) continue;
      u64 d[TW];
      {
        uint32_t da = dbase + (uint32_t)(row * TW) * cbs;      // out-of-image pixels were zero-filled by TMA
#pragma unroll
        for (int i = 0; i < TW; ++i) {
          d[i] = lds_pair<T>(da);
          da += cbs;
          if (ky == 0) { ds0 += lo2(d[i]); ds1 += hi2(d[i]); }
        }
      }
      uint32_t a = base + (uint32_t)(row + ky) * row_pitch;
#pragma unroll
      for (int jx = 0; jx < TW + 6; ++jx) {
        const u64 in = lds_pair<T>(a);
        a += cbs;
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const int i = jx - kx;
          if (i >= 0 && i < TW) a7[kx] = ffma2(d[i], in, a7[kx]);
        }
      }
    }
    __syncthreads();  // previous round's fold buffer fully consumed (first round: nothing pending)
    if (trow < g.TR) {
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) fold[(kx * g.TR + trow) * g.P + pr] = a7[kx];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 7 * g.P; idx += blockDim.x) {
      const int kx = idx / g.P, pp = idx - kx * g.P;
      float s0 = 0.f, s1 = 0.f;
      for (int r = 0; r < g.TR; ++r) { const u64 v = fold[(kx * g.TR + r) * g.P + pp]; s0 += lo2(v); s1 += hi2(v); }
      float* dst = slot + (size_t)(ky * 7 + kx) * g.Cfull + pp * 2;
      atomicAdd(dst, s0);
      atomicAdd(dst + 1, s1);
    }
  }
  // bias gradient
  __syncthreads();
  if (trow < g.TR) fold[trow * g.P + pr] = pack2(ds0, ds1);
  __syncthreads();
  for (int pp = threadIdx.x; pp < g.P; pp += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    for (int r = 0; r < g.TR; ++r) { const u64 v = fold[r * g.P + pp]; s0 += lo2(v); s1 += hi2(v); }
    atomicAdd(slot + (size_t)49 * g.Cfull + pp * 2, s0);
    atomicAdd(slot + (size_t)49 * g.Cfull + pp * 2 + 1, s1);
  }
}

// out[j] += sum_p partial[p][j]
__global__ void reduce_parts_kernel(const float* __restrict__ partial, int nparts, int n, float* __restrict__ out0, int n0,
                                    float* __restrict__ out1) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * n + j];
  if (j < n0) { if (out0) out0[j] += s; }
  else if (out1) out1[j - n0] += s;
}

// ---- host side ---------------------------------------------------------------------------------------------------
static int make_x_map(const void* x, int B, int H, int W, int C, int dtype, int cbox, int tw, int th, CUtensorMap* out,
                      int halo = 6) {
  const uint64_t es = dtype == GA_BF16 ? 2 : 4;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * es, (uint64_t)W * C * es, (uint64_t)H * W * C * es};
  uint32_t box[4] = {(uint32_t)cbox, (uint32_t)(tw + halo), (uint32_t)(th + halo), 1};
  return ga_tensor_map(out, dtype, 4, x, dims, strides, box, 0);
}

struct Plan { Geo g; int tw, threads, dbox_stride; size_t smem; };

// choose tile width and rows per CTA: <=512 threads, <=110 KB shared memory (two CTAs per SM)
static int plan(int B, int H, int W, int C, int dtype, bool wgrad, Plan* p, int max_threads = 512) {
  const int es = dtype == GA_BF16 ? 2 : 4;
  GA_REQUIRE(C >= 8 && (C * es) % 16 == 0 && (C & 1) == 0, GA_ERR_ALIGN, "dwconv7: C=%d must be even with 16-byte rows", C);
  const int P = C / 2;
  GA_REQUIRE(P <= 512, GA_ERR_UNSUPPORTED, "dwconv7: C=%d too wide for one CTA row", C);
  if (P > max_threads) max_threads = 512;   // C up to 1024: one row per pass, registers spill a little
  int tw = (W % 14 == 0) ? 14 : ((W % 7 == 0) ? 7 : (W >= 12 ? 14 : (W >= 6 ? 7 : 4)));
  const char* et = getenv("GA_DW_TW");
  if (et) { int v = atoi(et); if (v == 4 || v == 7 || v == 14) tw = v; }
  const int align = 16 / es;                 // channel boxes: <=256 elements, 16-byte multiples
  const int nbox = (C + 255) / 256;
  const int cbox = (((C + nbox - 1) / nbox) + align - 1) / align * align;
  int tr = max_threads / P;
  if (tr < 1) tr = 1;
  if (tr > H) tr = H;
  int th = 1, box_stride = 0;
  size_t smem = 0;
  auto fit = [&](int tw_, size_t limit) {          // largest tile height (<= 14 rows) whose halo + scratch fit `limit`
    int best = 0;
    for (int t = 1; t <= 14 && t <= H; ++t) {
      const int trr = tr < t ? tr : t;
      const size_t box_bytes = (((size_t)(t + 6) * (tw_ + 6) * cbox * es) + 127) & ~(size_t)127;
      const size_t dbox = wgrad ? ((((size_t)t * tw_ * cbox * es) + 127) & ~(size_t)127) : 0;
      const size_t extra = wgrad ? (size_t)7 * trr * P * 8 + (size_t)nbox * dbox : ((size_t)trr * tw_ * P + 2 * (size_t)trr * tw_) * 4;
      if (128 + 128 + (size_t)nbox * box_bytes + extra + 64 <= limit) best = t;
    }
    return best;
  };
  // prefer two resident CTAs per SM (110 KB each); fall back to one big CTA when that leaves fewer than 3 rows per tile
  size_t limit = 110 * 1024;
  th = fit(tw, limit);
  if (th < 3 && th < H) {
    int th7 = (tw == 14) ? fit(7, limit) : 0;
    if (th7 >= 3) { tw = 7; th = th7; }
    else { limit = 225 * 1024; th = fit(tw, limit); if (th < 1 && tw == 14) { tw = 7; th = fit(7, limit); } if (th < 1) { tw = 4; th = fit(4, limit); } }
  }
  {
    // measured on B200 (scripts/kernel_bench.py dwconv, GA_DW_TH sweep 2..14): short tiles win for the conv / data-gradient
    // kernels (more CTAs in flight hide the halo-load latency; the extra halo rows are L2 hits: 56x56x96 fwd 360 -> 307 us,
    // 7x7x688 fwd 132 -> 88 us), taller ones for the weight gradient (fewer folds and atomics per pixel: 436 -> 374 us)
    if (!wgrad) {
      const int pref = H >= 28 ? 4 : 2;
      if (th > pref) th = pref;
    } else if (H >= 28 && fit(tw, 225 * 1024) >= 10) {
      th = 10;                                   // two passes of 5 rows: 240-thread CTAs, still two per SM at C <= 192
    }
  }
  {
    const char* eh = getenv("GA_DW_TH");
    if (eh) { int v = atoi(eh); if (v >= 1 && v <= 14) th = v; }
  }
  GA_REQUIRE(th >= 1, GA_ERR_UNSUPPORTED, "dwconv7: tile does not fit shared memory (C=%d W=%d)", C, W);
  // balance the passes: e.g. 7 rows with 5 row slots -> 2 passes of 4 rows
  if (tr > th) tr = th;
  {
    const int passes = (th + tr - 1) / tr;
    tr = (th + passes - 1) / passes;
    const size_t box_bytes = (((size_t)(th + 6) * (tw + 6) * cbox * es) + 127) & ~(size_t)127;
    box_stride = (int)(box_bytes / es);
    const size_t dbox = wgrad ? ((((size_t)th * tw * cbox * es) + 127) & ~(size_t)127) : 0;
    p->dbox_stride = (int)(dbox / es);
    const size_t extra = wgrad ? (size_t)7 * tr * P * 8 + (size_t)nbox * dbox : ((size_t)tr * tw * P + 2 * (size_t)tr * tw) * 4;
    smem = 128 + 128 + (size_t)nbox * box_bytes + extra + 64;
  }
  GA_REQUIRE(smem <= 227 * 1024, GA_ERR_UNSUPPORTED, "dwconv7: tile does not fit shared memory (C=%d W=%d)", C, W);
  p->g.B = B; p->g.H = H; p->g.W = W; p->g.C = C; p->g.TH = th; p->g.TR = tr;
  p->g.tiles_x = (W + tw - 1) / tw; p->g.tiles_y = (H + th - 1) / th;
  p->g.cbox = cbox; p->g.nbox = nbox; p->g.box_stride = box_stride; p->g.P = P; p->g.c0 = 0; p->g.Cfull = C;
  p->tw = tw; p->threads = ((tr * P + 31) / 32) * 32; p->smem = smem;
  return GA_OK;
}

template <typename T, typename TO, int MODE>
static int launch_conv(const Plan& p, const CUtensorMap& tm, const float* w, const float* bias, const float* ln_w,
                       const float* ln_b, const void* res, void* y, float* rstd, float eps, cudaStream_t st) {
  dim3 grid(p.g.tiles_x * p.g.tiles_y, p.g.B);
#define GA_DW_LAUNCH(TW_)                                                                                               \
  if (p.tw == TW_) {                                                                                                    \
    auto k = p.threads > 448 ? dwconv7_kernel<T, TO, TW_, MODE, true> : dwconv7_kernel<T, TO, TW_, MODE, false>;         \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);                                   \
    k<<<grid, p.threads, p.smem, st>>>(tm, w, bias, ln_w, ln_b, (const TO*)res, (TO*)y, rstd, eps, p.g);                 \
    ga_count_launch();                                                                                                  \
    return ga_check_launch("dwconv7");                                                                                  \
  }
  GA_DW_LAUNCH(14) GA_DW_LAUNCH(7) GA_DW_LAUNCH(4)
#undef GA_DW_LAUNCH
  ga_set_error("dwconv7: no kernel for tw=%d", p.tw);
  return GA_ERR_UNSUPPORTED;
}

}  // namespace dw

// second-generation bf16 kernels (dwconv3.cu); GA_ERR_UNSUPPORTED = shape not taken, fall through to the kernels above
int ga_dwconv7_ln_fwd_v3(const void* x, const float* w49c, const float* bias, void* y, float* rstd, int B, int H, int W, int C,
                         float eps, cudaStream_t st);
int ga_dwconv7_bwd_v3(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dxs, const float* dxs_scale, float* partial,
                      int nparts, int B, int H, int W, int C, int res_dtype, cudaStream_t st);
// GA_DW_V3=0 (read once, immutable afterwards) keeps the first-generation kernels for A/B timing in scripts/kernel_bench.py
static bool dw_v3_enabled() {
  static const bool on = [] { const char* e = getenv("GA_DW_V3"); return !(e && atoi(e) == 0); }();
  return on;
}

extern "C" int ga_dwconv7_ln_fwd(const void* x, const float* w49c, const float* bias, const float* ln_w, const float* ln_b,
                                 void* y, float* rstd, int B, int H, int W, int C, float eps, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && w49c && y && B > 0 && H > 0 && W > 0, GA_ERR_SHAPE, "ga_dwconv7_ln_fwd: bad arguments");
  GA_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w49c & 7) == 0 && ((uintptr_t)bias & 7) == 0, GA_ERR_ALIGN,
             "ga_dwconv7_ln_fwd: x must be 16-byte, weights 8-byte aligned");
  if (dtype == GA_BF16 && !ln_w && !ln_b && dw_v3_enabled()) {
    const int rc3 = ga_dwconv7_ln_fwd_v3(x, w49c, bias, y, rstd, B, H, W, C, eps, (cudaStream_t)s);
    if (rc3 != GA_ERR_UNSUPPORTED) return rc3;
  }
  dw::Plan p;
  int rc = dw::plan(B, H, W, C, dtype, false, &p, 384);
  if (rc == GA_ERR_UNSUPPORTED) {
    // the full-width halo does not fit one CTA (fp32 at C >= 976): conv + bias in channel slices, then a row LayerNorm pass
    int nslice = 2;
    for (;; ++nslice) {
      GA_REQUIRE(nslice <= 16, GA_ERR_UNSUPPORTED, "dwconv7: C=%d does not fit shared memory even in 16 channel slices", C);
      if (C % nslice || (C / nslice) % 8) continue;
      if (dw::plan(B, H, W, C / nslice, dtype, false, &p, 448) == GA_OK) break;
    }
    CUtensorMap tm;
    rc = dw::make_x_map(x, B, H, W, C, dtype, p.g.cbox, p.tw, p.g.TH, &tm);
    if (rc) return rc;
    p.g.Cfull = C;
    for (int sl = 0; sl < nslice; ++sl) {
      p.g.c0 = sl * (C / nslice);
      rc = dtype == GA_BF16 ? dw::launch_conv<bf16, bf16, 2>(p, tm, w49c, bias, nullptr, nullptr, nullptr, y, nullptr, 0.f, (cudaStream_t)s)
                            : dw::launch_conv<float, float, 2>(p, tm, w49c, bias, nullptr, nullptr, nullptr, y, nullptr, 0.f, (cudaStream_t)s);
      if (rc) return rc;
    }
    return ga_layernorm_fwd(y, ln_w, ln_b, y, nullptr, rstd, (long long)B * H * W, C, C, C, eps, dtype, s);
  }
  if (rc) return rc;
  CUtensorMap tm;
  rc = dw::make_x_map(x, B, H, W, C, dtype, p.g.cbox, p.tw, p.g.TH, &tm);
  if (rc) return rc;
  if (dtype == GA_BF16) return dw::launch_conv<bf16, bf16, 0>(p, tm, w49c, bias, ln_w, ln_b, nullptr, y, rstd, eps, (cudaStream_t)s);
  return dw::launch_conv<float, float, 0>(p, tm, w49c, bias, ln_w, ln_b, nullptr, y, rstd, eps, (cudaStream_t)s);
}

extern "C" int ga_dwconv7_bwd_parts(int B, int H, int W, int C) { return 32; }

extern "C" int ga_dwconv7_bwd2(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dx_shadow,
                               float* dw49c, float* dbias, float* dw_partial, int B, int H, int W, int C, int dtype, int res_dtype,
                               ga_stream_t s);
extern "C" int ga_dwconv7_bwd3(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dx_shadow,
                               const float* shadow_rowscale, float* dw49c, float* dbias, float* dw_partial, int B, int H, int W, int C,
                               int dtype, int res_dtype, ga_stream_t s);
extern "C" int ga_dwconv7_bwd(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, float* dw49c,
                              float* dbias, float* dw_partial, int B, int H, int W, int C, int dtype, int res_dtype,
                              ga_stream_t s) {
  return ga_dwconv7_bwd2(dconv, x, dres, w49c, dx, nullptr, dw49c, dbias, dw_partial, B, H, W, C, dtype, res_dtype, s);
}
extern "C" int ga_dwconv7_bwd2(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dx_shadow,
                               float* dw49c, float* dbias, float* dw_partial, int B, int H, int W, int C, int dtype, int res_dtype,
                               ga_stream_t s) {
  return ga_dwconv7_bwd3(dconv, x, dres, w49c, dx, dx_shadow, nullptr, dw49c, dbias, dw_partial, B, H, W, C, dtype, res_dtype, s);
}
extern "C" int ga_dwconv7_bwd3(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dx_shadow,
                               const float* shadow_rowscale, float* dw49c, float* dbias, float* dw_partial, int B, int H, int W, int C,
                               int dtype, int res_dtype, ga_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  GA_REQUIRE(dconv && w49c && B > 0, GA_ERR_SHAPE, "ga_dwconv7_bwd: bad arguments");
  GA_REQUIRE(((uintptr_t)dconv & 15) == 0 && ((uintptr_t)w49c & 7) == 0, GA_ERR_ALIGN, "ga_dwconv7_bwd: misaligned operands");
  int rc;
  // GA_DW_BWD3 = 0 (read once) keeps the two first-generation kernels for A/B timing (scripts/kernel_bench.py)
  static const int bwd3_mode = [] { const char* e = getenv("GA_DW_BWD3"); return e ? atoi(e) : -1; }();
  const bool bwd3 = bwd3_mode != 0;
  // dw_partial == NULL with dbias == dw49c + 49 C (one [50][C] accumulator, as the callers' gradient slabs are laid out):
  // the fused kernel adds its per-CTA sums straight into dw49c / dbias (fp32 atomics, ACCUMULATING: the caller zeroes them) --
  // no workspace clear and no reduction launch.  Only the fused bf16 kernel supports it; otherwise GA_ERR_UNSUPPORTED.
  const bool direct = !dw_partial && dw49c && dbias == dw49c + (size_t)49 * C;
  if (dtype == GA_BF16 && dx && x && (dw_partial || direct) && (dw49c || dbias) && ((uintptr_t)x & 15) == 0 && dw_v3_enabled() && bwd3) {
    // one fused kernel: data gradient (+ residual, + bf16 shadow), weight gradient and bias gradient from one staged dconv halo
    if (direct) {
      rc = ga_dwconv7_bwd_v3(dconv, x, dres, w49c, dx, dx_shadow, shadow_rowscale, dw49c, 1, B, H, W, C, res_dtype, st);
      GA_REQUIRE(rc != GA_ERR_UNSUPPORTED, GA_ERR_UNSUPPORTED, "ga_dwconv7_bwd: this shape needs the partial workspace");
      return rc;
    }
    const int nparts = ga_dwconv7_bwd_parts(B, H, W, C);
    cudaMemsetAsync(dw_partial, 0, (size_t)nparts * 50 * C * sizeof(float), st);
    rc = ga_dwconv7_bwd_v3(dconv, x, dres, w49c, dx, dx_shadow, shadow_rowscale, dw_partial, nparts, B, H, W, C, res_dtype, st);
    if (rc == GA_OK) {
      const int n = 50 * C;
      dw::reduce_parts_kernel<<<(n + 255) / 256, 256, 0, st>>>(dw_partial, nparts, n, dw49c, 49 * C, dbias);
      ga_count_launch();
      return ga_check_launch("dwconv7_wgrad_reduce");
    }
    if (rc != GA_ERR_UNSUPPORTED) return rc;
  }
  GA_REQUIRE(!shadow_rowscale, GA_ERR_UNSUPPORTED, "ga_dwconv7_bwd3: the scaled shadow is written by the fused bf16 kernel only");
  GA_REQUIRE(dw_partial || !(dw49c || dbias), GA_ERR_UNSUPPORTED, "ga_dwconv7_bwd: this configuration needs the partial workspace");
  if (dx) {
    dw::Plan p;
    int nslice = 1;
    for (;; ++nslice) {            // fp32 at C >= 976: cover the channels in equal slices
      GA_REQUIRE(nslice <= 16, GA_ERR_UNSUPPORTED, "dwconv7: C=%d does not fit shared memory even in 16 channel slices", C);
      if (C % nslice || (C / nslice) % 8) continue;
      rc = dw::plan(B, H, W, C / nslice, dtype, false, &p, 448);
      if (rc == GA_OK) break;
      if (rc != GA_ERR_UNSUPPORTED) return rc;
    }
    CUtensorMap tm;
    rc = dw::make_x_map(dconv, B, H, W, C, dtype, p.g.cbox, p.tw, p.g.TH, &tm);
    if (rc) return rc;
    p.g.Cfull = C;
    for (int sl = 0; sl < nslice; ++sl) {
    p.g.c0 = sl * (C / nslice);
    GA_REQUIRE(dtype == GA_BF16 || res_dtype == GA_F32, GA_ERR_UNSUPPORTED, "ga_dwconv7_bwd: fp32 gradients need an fp32 residual stream");
    if (dtype == GA_BF16 && res_dtype == GA_BF16) rc = dw::launch_conv<bf16, bf16, 1>(p, tm, w49c, nullptr, nullptr, nullptr, dres, dx, (float*)dx_shadow, 0.f, st);
    else if (dtype == GA_BF16) rc = dw::launch_conv<bf16, float, 1>(p, tm, w49c, nullptr, nullptr, nullptr, dres, dx, (float*)dx_shadow, 0.f, st);
    else rc = dw::launch_conv<float, float, 1>(p, tm, w49c, nullptr, nullptr, nullptr, dres, dx, (float*)dx_shadow, 0.f, st);
    if (rc) return rc;
    }
  }
  if (dw49c || dbias) {
    GA_REQUIRE(x && dw_partial, GA_ERR_SHAPE, "ga_dwconv7_bwd: weight gradient needs x and a partial workspace");
    GA_REQUIRE(((uintptr_t)x & 15) == 0, GA_ERR_ALIGN, "ga_dwconv7_bwd: x must be 16-byte aligned");
    // wide C in fp32 does not fit one CTA's shared memory: cover the channels in equal slices (multiples of 8)
    int nslice = 1;
    dw::Plan p;
    for (;; ++nslice) {
      if (C % nslice || (C / nslice) % 8) { GA_REQUIRE(nslice < 16, GA_ERR_UNSUPPORTED, "dwconv7 wgrad: cannot slice C=%d", C); continue; }
      rc = dw::plan(B, H, W, C / nslice, dtype, true, &p, 448);
      if (rc == GA_OK) break;
      GA_REQUIRE(nslice < 16, GA_ERR_UNSUPPORTED, "dwconv7 wgrad: C=%d does not fit shared memory", C);
    }
    const int nparts = ga_dwconv7_bwd_parts(B, H, W, C);
    cudaMemsetAsync(dw_partial, 0, (size_t)nparts * 50 * C * sizeof(float), st);
    CUtensorMap tm, tmd;
    rc = dw::make_x_map(x, B, H, W, C, dtype, p.g.cbox, p.tw, p.g.TH, &tm);
    if (rc) return rc;
    rc = dw::make_x_map(dconv, B, H, W, C, dtype, p.g.cbox, p.tw, p.g.TH, &tmd, 0);
    if (rc) return rc;
    dim3 grid(p.g.tiles_x * p.g.tiles_y, B);
    p.g.Cfull = C;
#define GA_DWW_LAUNCH(T_, TW_)                                                                      \
  {                                                                                                 \
    auto k = p.threads > 448 ? dw::dwconv7_wgrad_kernel<T_, TW_, true> : dw::dwconv7_wgrad_kernel<T_, TW_, false>; \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);               \
    k<<<grid, p.threads, p.smem, st>>>(tm, tmd, p.dbox_stride, dw_partial, nparts, p.g);             \
  }
    for (int sl = 0; sl < nslice; ++sl) {
      p.g.c0 = sl * (C / nslice);
      if (dtype == GA_BF16) {
        if (p.tw == 14) GA_DWW_LAUNCH(bf16, 14) else if (p.tw == 7) GA_DWW_LAUNCH(bf16, 7) else GA_DWW_LAUNCH(bf16, 4)
      } else {
        if (p.tw == 14) GA_DWW_LAUNCH(float, 14) else if (p.tw == 7) GA_DWW_LAUNCH(float, 7) else GA_DWW_LAUNCH(float, 4)
      }
      ga_count_launch();
      rc = ga_check_launch("dwconv7_wgrad");
      if (rc) return rc;
    }
#undef GA_DWW_LAUNCH
    const int n = 50 * C;
    dw::reduce_parts_kernel<<<(n + 255) / 256, 256, 0, st>>>(dw_partial, nparts, n, dw49c, 49 * C, dbias);
    ga_count_launch();
    rc = ga_check_launch("dwconv7_wgrad_reduce");
    if (rc) return rc;
  }
  return GA_OK;
}
