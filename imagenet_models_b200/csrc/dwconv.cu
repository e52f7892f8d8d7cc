// K1: channels-last depthwise 7x7 convolution fused with LayerNorm (forward), its data gradient (flipped taps,
// fused with the residual add) and its weight gradient.   Reference: ConvNeXtBlock.conv_dw + norm
// (ga_convnext.py:92-93,100,105-106; map_convnext.py:18-19,29-31) and their autograd.
//
// Layout: x is NHWC.  One CTA owns a TH x TW pixel tile for ALL channels (LayerNorm needs the whole row).
// The (TH+6) x (TW+6) x C input halo is staged in shared memory by TMA (4-D tensor map, out-of-bounds = zero
// = the conv padding), one elected thread, one mbarrier.  Thread (row, channel-group) slides a 7-tap window
// along its output row keeping TW x CPT fp32 accumulators in registers; LayerNorm statistics are reduced with
// warp shuffles + one small smem exchange (two rounds: mean, then centred variance).
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace dw {

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
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct Geo {
  int B, H, W, C;
  int TH;            // output rows per CTA
  int tiles_x, tiles_y;
  int cbox, nbox;    // channel box of the TMA load and number of boxes (nbox*cbox >= C)
  int box_stride;    // elements between consecutive channel boxes in smem (128-byte aligned)
  int cgpad;         // channel groups per row padded to a multiple of 32 (threads per output row)
};

// smem element (halo row hy, halo col hx, channel c)
template <typename T, int TW>
__device__ __forceinline__ const T* halo_ptr(const T* tile, const Geo& g, int hy, int hx, int c) {
  const int box = c / g.cbox, cc = c - box * g.cbox;
  return tile + (size_t)box * g.box_stride + ((size_t)hy * (TW + 6) + hx) * g.cbox + cc;
}

template <int CPT, typename T> struct LdC;
template <> struct LdC<1, float> { static __device__ __forceinline__ void ld(const float* p, float* v) { v[0] = p[0]; } };
template <> struct LdC<2, float> { static __device__ __forceinline__ void ld(const float* p, float* v) { float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y; } };
template <> struct LdC<1, bf16> { static __device__ __forceinline__ void ld(const bf16* p, float* v) { v[0] = __bfloat162float(p[0]); } };
template <> struct LdC<2, bf16> { static __device__ __forceinline__ void ld(const bf16* p, float* v) { uint32_t u = *reinterpret_cast<const uint32_t*>(p); v[0] = bf16lo(u); v[1] = bf16hi(u); } };
template <int CPT, typename T> struct StC;
template <> struct StC<1, float> { static __device__ __forceinline__ void st(float* p, const float* v) { p[0] = v[0]; } };
template <> struct StC<2, float> { static __device__ __forceinline__ void st(float* p, const float* v) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); } };
template <> struct StC<1, bf16> { static __device__ __forceinline__ void st(bf16* p, const float* v) { p[0] = __float2bfloat16_rn(v[0]); } };
template <> struct StC<2, bf16> { static __device__ __forceinline__ void st(bf16* p, const float* v) { *reinterpret_cast<uint32_t*>(p) = pack_bf16(v[0], v[1]); } };

// stage the halo tile: one thread arms the barrier and issues one 4-D box per channel box
template <typename T, int TW>
__device__ __forceinline__ void load_halo(T* tile, uint64_t* bar, const CUtensorMap* map, const Geo& g, int b, int y0, int x0) {
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t box_bytes = (uint32_t)((g.TH + 6) * (TW + 6) * g.cbox * sizeof(T));
    mbar_expect_tx(bar, box_bytes * g.nbox);
    for (int j = 0; j < g.nbox; ++j)
      tma_load_4d(tile + (size_t)j * g.box_stride, map, bar, j * g.cbox, x0 - 3, y0 - 3, b);
  }
  mbar_wait(bar, 0);
}

// MODE 0: forward  y = LN(conv(x) + bias)  (xhat, optional affine), rstd saved
// MODE 1: dgrad    y = corr(x = dconv, flipped taps) + res
constexpr int max_threads_for(int tw, int cpt) { return tw * cpt >= 28 ? 512 : (tw * cpt >= 14 ? 768 : 1024); }

template <typename T, typename TO, int TW, int CPT, int MODE>
__global__ void __launch_bounds__(max_threads_for(TW, CPT)) dwconv7_kernel(const __grid_constant__ CUtensorMap tm, const float* __restrict__ w49c,
                                                       const float* __restrict__ bias, const float* __restrict__ ln_w,
                                                       const float* __restrict__ ln_b, const TO* __restrict__ res,
                                                       TO* __restrict__ y, float* __restrict__ rstd_out, float eps, Geo g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bar = (uint64_t*)sm;
  T* tile = (T*)(sm + 128);
  const size_t tile_bytes = (size_t)g.nbox * g.box_stride * sizeof(T);
  float* red = (float*)(sm + 128 + ((tile_bytes + 15) & ~(size_t)15));   // [TH][TW][nwarps_per_row]
  const int wpr = g.cgpad / 32;                                     // warps per output row
  float* stat = red + g.TH * TW * wpr;                              // [TH][TW] (mean, then rstd)

  const int tile_id = blockIdx.x;
  const int tx = tile_id % g.tiles_x, ty = tile_id / g.tiles_x;
  const int b = blockIdx.y;
  const int x0 = tx * TW, y0 = ty * g.TH;

  load_halo<T, TW>(tile, bar, &tm, g, b, y0, x0);

  const int row = threadIdx.x / g.cgpad;         // output row inside the tile
  const int cg = threadIdx.x - row * g.cgpad;    // channel group
  const int c = cg * CPT;
  const bool active = (c < g.C);
  const int lane = threadIdx.x & 31, wrow = cg >> 5;

  float acc[TW][CPT];
#pragma unroll
  for (int i = 0; i < TW; ++i)
#pragma unroll
    for (int j = 0; j < CPT; ++j) acc[i][j] = 0.f;

  if (active) {
    if (MODE == 0 && bias) {
      float bv[CPT];
#pragma unroll
      for (int j = 0; j < CPT; ++j) bv[j] = bias[c + j];
#pragma unroll
      for (int i = 0; i < TW; ++i)
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[i][j] = bv[j];
    }
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky) {
      float wv[7][CPT];
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const int tap = (MODE == 0) ? (ky * 7 + kx) : ((6 - ky) * 7 + (6 - kx));
#pragma unroll
        for (int j = 0; j < CPT; ++j) wv[kx][j] = __ldg(w49c + (size_t)tap * g.C + c + j);
      }
      const T* rowp = halo_ptr<T, TW>(tile, g, row + ky, 0, c);
      float in[TW + 6][CPT];
#pragma unroll
      for (int i = 0; i < TW + 6; ++i) LdC<CPT, T>::ld(rowp + (size_t)i * g.cbox, in[i]);
#pragma unroll
      for (int i = 0; i < TW; ++i)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx)
#pragma unroll
          for (int j = 0; j < CPT; ++j) acc[i][j] = fmaf(in[i + kx][j], wv[kx][j], acc[i][j]);
    }
  }

  const int oy = y0 + row;
  if (MODE == 1) {
    if (active && oy < g.H) {
#pragma unroll
      for (int i = 0; i < TW; ++i) {
        const int ox = x0 + i;
        if (ox < g.W) {
          const size_t off = (((size_t)b * g.H + oy) * g.W + ox) * g.C + c;
          float v[CPT];
#pragma unroll
          for (int j = 0; j < CPT; ++j) v[j] = acc[i][j];
          if (res) {
            float r[CPT];
            LdC<CPT, TO>::ld(res + off, r);
#pragma unroll
            for (int j = 0; j < CPT; ++j) v[j] += r[j];
          }
          StC<CPT, TO>::st(y + off, v);
        }
      }
    }
    return;
  }

  // ---- LayerNorm over channels (per pixel): round 1 mean, round 2 centred variance
  const float invC = 1.f / (float)g.C;
#pragma unroll
  for (int i = 0; i < TW; ++i) {
    float s = 0.f;
    if (active) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) s += acc[i][j];
    }
    s = warp_sum(s);
    if (lane == 0) red[(row * TW + i) * wpr + wrow] = s;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < g.TH * TW; p += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < wpr; ++k) s += red[p * wpr + k];
    stat[p] = s * invC;
  }
  __syncthreads();
  float mean[TW];
#pragma unroll
  for (int i = 0; i < TW; ++i) {
    mean[i] = stat[row * TW + i];
    float s = 0.f;
    if (active) {
#pragma unroll
      for (int j = 0; j < CPT; ++j) { float d = acc[i][j] - mean[i]; s += d * d; }
    }
    s = warp_sum(s);
    if (lane == 0) red[(row * TW + i) * wpr + wrow] = s;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < g.TH * TW; p += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < wpr; ++k) s += red[p * wpr + k];
    const float r = rsqrtf(s * invC + eps);
    stat[p] = r;
    const int py = y0 + p / TW, px = x0 + p % TW;
    if (rstd_out && py < g.H && px < g.W) rstd_out[((size_t)b * g.H + py) * g.W + px] = r;
  }
  __syncthreads();
  if (active && oy < g.H) {
    float lw[CPT], lb[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) { lw[j] = ln_w ? ln_w[c + j] : 1.f; lb[j] = ln_w ? ln_b[c + j] : 0.f; }
#pragma unroll
    for (int i = 0; i < TW; ++i) {
      const int ox = x0 + i;
      if (ox < g.W) {
        const float r = stat[row * TW + i];
        float v[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) v[j] = (acc[i][j] - mean[i]) * r * lw[j] + lb[j];
        StC<CPT, TO>::st(y + (((size_t)b * g.H + oy) * g.W + ox) * g.C + c, v);
      }
    }
  }
}

// weight gradient: dw[tap][c] += sum_pixels dconv(p, c) * x(p + tap - 3, c);  dbias[c] += sum dconv(p, c)
// One CTA per tile; thread (row, channel) keeps 49 accumulators; rows are folded through smem, then one
// atomicAdd per (tap, channel) per CTA into partial slot (blockIdx % nparts).
template <typename T, int TW>
__global__ void __launch_bounds__(max_threads_for(TW, 1)) dwconv7_wgrad_kernel(const __grid_constant__ CUtensorMap tmx, const T* __restrict__ dconv,
                                                             float* __restrict__ partial, int nparts, Geo g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  uint64_t* bar = (uint64_t*)sm;
  T* tile = (T*)(sm + 128);
  const size_t tile_bytes = (size_t)g.nbox * g.box_stride * sizeof(T);
  float* fold = (float*)(sm + 128 + ((tile_bytes + 15) & ~(size_t)15));   // [7][TH][cgpad]

  const int tile_id = blockIdx.x;
  const int tx = tile_id % g.tiles_x, ty = tile_id / g.tiles_x;
  const int b = blockIdx.y;
  const int x0 = tx * TW, y0 = ty * g.TH;
  load_halo<T, TW>(tile, bar, &tmx, g, b, y0, x0);

  const int row = threadIdx.x / g.cgpad;
  const int c = threadIdx.x - row * g.cgpad;
  const bool active = (c < g.C);
  const int oy = y0 + row;

  float d[TW];
  float dsum = 0.f;
#pragma unroll
  for (int i = 0; i < TW; ++i) {
    const int ox = x0 + i;
    d[i] = (active && oy < g.H && ox < g.W) ? ld_f(dconv + (((size_t)b * g.H + oy) * g.W + ox) * g.C + c) : 0.f;
    dsum += d[i];
  }
  float* slot = partial + (size_t)((blockIdx.x + blockIdx.y * gridDim.x) % nparts) * 50 * g.C;
#pragma unroll 1
  for (int ky = 0; ky < 7; ++ky) {
    float a[7];
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) a[kx] = 0.f;
    if (active) {
      const T* rowp = halo_ptr<T, TW>(tile, g, row + ky, 0, c);
      float in[TW + 6];
#pragma unroll
      for (int i = 0; i < TW + 6; ++i) in[i] = ld_f(rowp + (size_t)i * g.cbox);
#pragma unroll
      for (int i = 0; i < TW; ++i)
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) a[kx] = fmaf(d[i], in[i + kx], a[kx]);
    }
    __syncthreads();  // previous round's fold buffer fully consumed
#pragma unroll
    for (int kx = 0; kx < 7; ++kx) fold[(kx * g.TH + row) * g.cgpad + c] = a[kx];
    __syncthreads();
    for (int idx = threadIdx.x; idx < 7 * g.cgpad; idx += blockDim.x) {
      const int kx = idx / g.cgpad, cc = idx - kx * g.cgpad;
      if (cc < g.C) {
        float s = 0.f;
        for (int r = 0; r < g.TH; ++r) s += fold[(kx * g.TH + r) * g.cgpad + cc];
        atomicAdd(slot + (size_t)(ky * 7 + kx) * g.C + cc, s);
      }
    }
  }
  // bias gradient
  __syncthreads();
  fold[row * g.cgpad + c] = dsum;
  __syncthreads();
  for (int cc = threadIdx.x; cc < g.cgpad; cc += blockDim.x) {
    if (cc < g.C) {
      float s = 0.f;
      for (int r = 0; r < g.TH; ++r) s += fold[r * g.cgpad + cc];
      atomicAdd(slot + (size_t)49 * g.C + cc, s);
    }
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
static int make_x_map(const void* x, int B, int H, int W, int C, int dtype, int cbox, int tw, int th, CUtensorMap* out) {
  const uint64_t es = dtype == GA_BF16 ? 2 : 4;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t strides[3] = {(uint64_t)C * es, (uint64_t)W * C * es, (uint64_t)H * W * C * es};
  uint32_t box[4] = {(uint32_t)cbox, (uint32_t)(tw + 6), (uint32_t)(th + 6), 1};
  return ga_tensor_map(out, dtype, 4, x, dims, strides, box, 0);
}

struct Plan { Geo g; int tw, cpt, threads; size_t smem; };

// choose tile width, channels per thread and rows per CTA
static int plan(int B, int H, int W, int C, int dtype, bool wgrad, Plan* p) {
  const int es = dtype == GA_BF16 ? 2 : 4;
  GA_REQUIRE(C >= 8 && (C * es) % 16 == 0, GA_ERR_ALIGN, "dwconv7: C=%d rows must be 16-byte multiples", C);
  int cpt = 1;
  if (!wgrad && C % 2 == 0) {
    const int pad2 = ((C / 2 + 31) / 32) * 64, pad1 = ((C + 31) / 32) * 32;  // padded channel slots
    if (pad2 <= pad1) cpt = 2;
  }
  const char* ev = getenv("GA_DW_CPT");
  if (ev && !wgrad) { int v = atoi(ev); if ((v == 1 || v == 2) && C % v == 0) cpt = v; }
  const int cgpad = ((C / cpt + 31) / 32) * 32;
  GA_REQUIRE(cgpad <= 1024, GA_ERR_UNSUPPORTED, "dwconv7: C=%d too wide for one CTA row", C);
  int tw = (W % 14 == 0) ? 14 : ((W % 7 == 0) ? 7 : (W >= 12 ? 14 : (W >= 6 ? 7 : 4)));
  if (cpt == 2 && tw == 14 && dtype == GA_F32) tw = 7;
  const char* et = getenv("GA_DW_TW");
  if (et) { int v = atoi(et); if (v == 4 || v == 7 || v == 14) tw = v; }
  while (cgpad > max_threads_for(tw, cpt) && tw > 4) tw = (tw == 14) ? 7 : 4;
  GA_REQUIRE(cgpad <= max_threads_for(tw, cpt), GA_ERR_UNSUPPORTED, "dwconv7: C=%d does not fit a CTA", C);
  const int align = 16 / es;                 // channel boxes: <=256 elements, 16-byte multiples
  const int nbox = (C + 255) / 256;
  const int cbox = (((C + nbox - 1) / nbox) + align - 1) / align * align;
  int th = 1, box_stride = 0;
  size_t smem = 0;
  const size_t limit = 200 * 1024;
  for (;;) {
    th = max_threads_for(tw, cpt) / cgpad;
    if (th > H) th = H;
    const char* eh = getenv("GA_DW_TH");
    if (eh) { int v = atoi(eh); if (v >= 1 && v * cgpad <= max_threads_for(tw, cpt)) th = v; }
    for (;; --th) {
      const size_t box_bytes = (((size_t)(th + 6) * (tw + 6) * cbox * es) + 127) & ~(size_t)127;
      box_stride = (int)(box_bytes / es);
      const size_t tile = (size_t)nbox * box_bytes;
      const size_t extra = wgrad ? (size_t)7 * th * cgpad * 4 : (size_t)th * tw * (cgpad / 32 + 1) * 4;
      smem = 128 + 128 + tile + extra + 64;
      if (smem <= limit || th == 1) break;
    }
    if (smem <= 227 * 1024 || tw == 4) break;
    tw = (tw == 14) ? 7 : 4;   // narrower tile when even one row does not fit (wide C in fp32)
  }
  GA_REQUIRE(smem <= 227 * 1024, GA_ERR_UNSUPPORTED, "dwconv7: tile does not fit shared memory (C=%d W=%d)", C, W);
  p->g.B = B; p->g.H = H; p->g.W = W; p->g.C = C; p->g.TH = th;
  p->g.tiles_x = (W + tw - 1) / tw; p->g.tiles_y = (H + th - 1) / th;
  p->g.cbox = cbox; p->g.nbox = nbox; p->g.cgpad = cgpad; p->g.box_stride = box_stride;
  p->tw = tw; p->cpt = cpt; p->threads = th * cgpad; p->smem = smem;
  return GA_OK;
}

template <typename T, typename TO, int MODE>
static int launch_conv(const Plan& p, const CUtensorMap& tm, const float* w, const float* bias, const float* ln_w,
                       const float* ln_b, const void* res, void* y, float* rstd, float eps, cudaStream_t st) {
  dim3 grid(p.g.tiles_x * p.g.tiles_y, p.g.B);
#define GA_DW_LAUNCH(TW_, CPT_)                                                                                         \
  if (p.tw == TW_ && p.cpt == CPT_) {                                                                                   \
    auto k = dwconv7_kernel<T, TO, TW_, CPT_, MODE>;                                                                    \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);                                   \
    k<<<grid, p.threads, p.smem, st>>>(tm, w, bias, ln_w, ln_b, (const TO*)res, (TO*)y, rstd, eps, p.g);                 \
    ga_count_launch();                                                                                                  \
    return ga_check_launch("dwconv7");                                                                                  \
  }
  GA_DW_LAUNCH(14, 1) GA_DW_LAUNCH(14, 2) GA_DW_LAUNCH(7, 1) GA_DW_LAUNCH(7, 2) GA_DW_LAUNCH(4, 1) GA_DW_LAUNCH(4, 2)
#undef GA_DW_LAUNCH
  ga_set_error("dwconv7: no kernel for tw=%d cpt=%d", p.tw, p.cpt);
  return GA_ERR_UNSUPPORTED;
}

}  // namespace dw

extern "C" int ga_dwconv7_ln_fwd(const void* x, const float* w49c, const float* bias, const float* ln_w, const float* ln_b,
                                 void* y, float* rstd, int B, int H, int W, int C, float eps, int dtype, ga_stream_t s) {
  GA_REQUIRE(x && w49c && y && B > 0 && H > 0 && W > 0, GA_ERR_SHAPE, "ga_dwconv7_ln_fwd: bad arguments");
  GA_REQUIRE(((uintptr_t)x & 15) == 0, GA_ERR_ALIGN, "ga_dwconv7_ln_fwd: x must be 16-byte aligned");
  dw::Plan p;
  int rc = dw::plan(B, H, W, C, dtype, false, &p);
  if (rc) return rc;
  CUtensorMap tm;
  rc = dw::make_x_map(x, B, H, W, C, dtype, p.g.cbox, p.tw, p.g.TH, &tm);
  if (rc) return rc;
  if (dtype == GA_BF16) return dw::launch_conv<bf16, bf16, 0>(p, tm, w49c, bias, ln_w, ln_b, nullptr, y, rstd, eps, (cudaStream_t)s);
  return dw::launch_conv<float, float, 0>(p, tm, w49c, bias, ln_w, ln_b, nullptr, y, rstd, eps, (cudaStream_t)s);
}

extern "C" int ga_dwconv7_bwd_parts(int B, int H, int W, int C) { return 32; }

extern "C" int ga_dwconv7_bwd(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, float* dw49c,
                              float* dbias, float* dw_partial, int B, int H, int W, int C, int dtype, int res_dtype,
                              ga_stream_t s) {
  cudaStream_t st = (cudaStream_t)s;
  GA_REQUIRE(dconv && w49c && B > 0, GA_ERR_SHAPE, "ga_dwconv7_bwd: bad arguments");
  int rc;
  if (dx) {
    dw::Plan p;
    rc = dw::plan(B, H, W, C, dtype, false, &p);
    if (rc) return rc;
    CUtensorMap tm;
    rc = dw::make_x_map(dconv, B, H, W, C, dtype, p.g.cbox, p.tw, p.g.TH, &tm);
    if (rc) return rc;
    GA_REQUIRE(dtype == GA_BF16 || res_dtype == GA_F32, GA_ERR_UNSUPPORTED, "ga_dwconv7_bwd: fp32 gradients need an fp32 residual stream");
    if (dtype == GA_BF16 && res_dtype == GA_BF16) rc = dw::launch_conv<bf16, bf16, 1>(p, tm, w49c, nullptr, nullptr, nullptr, dres, dx, nullptr, 0.f, st);
    else if (dtype == GA_BF16) rc = dw::launch_conv<bf16, float, 1>(p, tm, w49c, nullptr, nullptr, nullptr, dres, dx, nullptr, 0.f, st);
    else rc = dw::launch_conv<float, float, 1>(p, tm, w49c, nullptr, nullptr, nullptr, dres, dx, nullptr, 0.f, st);
    if (rc) return rc;
  }
  if (dw49c || dbias) {
    GA_REQUIRE(x && dw_partial, GA_ERR_SHAPE, "ga_dwconv7_bwd: weight gradient needs x and a partial workspace");
    dw::Plan p;
    rc = dw::plan(B, H, W, C, dtype, true, &p);
    if (rc) return rc;
    CUtensorMap tm;
    rc = dw::make_x_map(x, B, H, W, C, dtype, p.g.cbox, p.tw, p.g.TH, &tm);
    if (rc) return rc;
    const int nparts = ga_dwconv7_bwd_parts(B, H, W, C);
    cudaMemsetAsync(dw_partial, 0, (size_t)nparts * 50 * C * sizeof(float), st);
    dim3 grid(p.g.tiles_x * p.g.tiles_y, B);
#define GA_DWW_LAUNCH(T_, TW_)                                                                      \
  {                                                                                                 \
    auto k = dw::dwconv7_wgrad_kernel<T_, TW_>;                                                     \
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);               \
    k<<<grid, p.threads, p.smem, st>>>(tm, (const T_*)dconv, dw_partial, nparts, p.g);              \
  }
    if (dtype == GA_BF16) {
      if (p.tw == 14) GA_DWW_LAUNCH(bf16, 14) else if (p.tw == 7) GA_DWW_LAUNCH(bf16, 7) else GA_DWW_LAUNCH(bf16, 4)
    } else {
      if (p.tw == 14) GA_DWW_LAUNCH(float, 14) else if (p.tw == 7) GA_DWW_LAUNCH(float, 7) else GA_DWW_LAUNCH(float, 4)
    }
#undef GA_DWW_LAUNCH
    ga_count_launch();
    rc = ga_check_launch("dwconv7_wgrad");
    if (rc) return rc;
    const int n = 50 * C;
    dw::reduce_parts_kernel<<<(n + 255) / 256, 256, 0, st>>>(dw_partial, nparts, n, dw49c, 49 * C, dbias);
    ga_count_launch();
    rc = ga_check_launch("dwconv7_wgrad_reduce");
    if (rc) return rc;
  }
  return GA_OK;
}
