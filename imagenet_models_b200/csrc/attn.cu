// K6: CSWin stripe attention with LePE, forward and backward (GA/ga_cswin.py:59-136, img2windows :215, windows2img :225).
//
// Input is the qkv projection as token rows [B*R*R, 3C] (q | k | v column blocks).  With two branches, branch 0 takes
// channels [0, C/2) and attends inside R x split stripes, branch 1 takes [C/2, C) and attends inside split x R stripes;
// with one branch the stripe is the whole R x R map.  Heads are 32 channels wide everywhere in the reference's
// configurations.  out = softmax(scale q k^T) v + dw3x3(v) evaluated INSIDE the stripe (zero padding at its border).
//
// One CTA per (image, branch, stripe, head): K and V of the stripe (<= 128 tokens x 32) live in shared memory as fp32,
// one thread per query row keeps q, the running max / sum and the 32-wide accumulator in registers (online softmax,
// 4 keys per rescale).  The window gather / scatter of the reference is pure index arithmetic here.  No score matrix
// ever reaches HBM: algorithmic traffic is qkv in + out (+ 4 B/token/head of log-sum-exp).
// Backward recomputes the probabilities from the saved log-sum-exp (one pass per query row for dq, one per key row
// for dk / dv), adds the transposed LePE stencil to dv and accumulates the LePE weight gradient per CTA over a slice of
// the batch before one atomic flush.
#include <cuda.h>

#include "common.cuh"

struct AttnGeom {
  int units, n, n4, nt;
};

namespace {
constexpr int HD = 32;   // head width
constexpr int RS = 36;   // shared-memory row pitch in floats: per-thread float4 row reads are bank-conflict free
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Unit {
  int hs, ws, y0, x0, cb, hg;   // stripe shape and origin, first channel, global head index
};

__device__ __forceinline__ Unit decode_unit(int u, int R, int C, int split, int nbr) {
  Unit t;
  if (nbr == 1) {
    t.hs = R; t.ws = R; t.y0 = 0; t.x0 = 0; t.cb = u * HD; t.hg = u;
    return t;
  }
  const int hb = C / 2 / HD, nst = R / split;
  const int br = u / (nst * hb);
  const int rem = u - br * nst * hb;
  const int st = rem / hb, h = rem - st * hb;
  if (br == 0) { t.hs = R; t.ws = split; t.y0 = 0; t.x0 = st * split; }
  else { t.hs = split; t.ws = R; t.y0 = st * split; t.x0 = 0; }
  t.cb = br * (C / 2) + h * HD;
  t.hg = br * hb + h;
  return t;
}

__device__ __forceinline__ long long tok_row(const Unit& u, int b, int R, int j) {
  const int ry = j / u.ws, rx = j - ry * u.ws;
  return ((long long)b * R + u.y0 + ry) * R + u.x0 + rx;
}

// [n][32] tile of `src` (row pitch ld, first column col) -> shared rows of pitch RS, scaled by mul
template <typename T, int NT>
__device__ __forceinline__ void stage_tile(float* dst, const T* __restrict__ src, long long ld, int col, const Unit& u, int b, int R,
                                           int n, float mul) {
  for (int idx = threadIdx.x; idx < n * 8; idx += NT) {
    const int j = idx >> 3, part = idx & 7;
    float4 v = ld4(src + tok_row(u, b, R, j) * ld + col + part * 4);
    v.x *= mul; v.y *= mul; v.z *= mul; v.w *= mul;
    *reinterpret_cast<float4*>(dst + j * RS + part * 4) = v;
  }
}

__device__ __forceinline__ void load_row(float* r, const float* row) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float4 v = *reinterpret_cast<const float4*>(row + p * 4);
    r[p * 4] = v.x; r[p * 4 + 1] = v.y; r[p * 4 + 2] = v.z; r[p * 4 + 3] = v.w;
  }
}

__device__ __forceinline__ float dot_row(const float* a, const float* row) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float4 v = *reinterpret_cast<const float4*>(row + p * 4);
    s0 = fmaf(a[p * 4], v.x, s0); s1 = fmaf(a[p * 4 + 1], v.y, s1);
    s0 = fmaf(a[p * 4 + 2], v.z, s0); s1 = fmaf(a[p * 4 + 3], v.w, s1);
  }
  return s0 + s1;
}

__device__ __forceinline__ void axpy_row(float* acc, float a, const float* row) {
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float4 v = *reinterpret_cast<const float4*>(row + p * 4);
    acc[p * 4] = fmaf(a, v.x, acc[p * 4]); acc[p * 4 + 1] = fmaf(a, v.y, acc[p * 4 + 1]);
    acc[p * 4 + 2] = fmaf(a, v.z, acc[p * 4 + 2]); acc[p * 4 + 3] = fmaf(a, v.w, acc[p * 4 + 3]);
  }
}

// lepe[d] (+)= sum_tap w[tap][d] * rows[neighbour(i, tap)][d]; sign = +1: correlation (forward), -1: its adjoint
__device__ __forceinline__ void lepe_row(float* acc, const float* rows, const float* w, const Unit& u, int i, int sign) {
  const int ry = i / u.ws, rx = i - ry * u.ws;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int yy = ry + sign * (tap / 3 - 1), xx = rx + sign * (tap % 3 - 1);
    if (yy < 0 || yy >= u.hs || xx < 0 || xx >= u.ws) continue;
    const float* row = rows + (yy * u.ws + xx) * RS;
    const float* wt = w + tap * HD;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float4 v = *reinterpret_cast<const float4*>(row + p * 4);
      const float4 c = *reinterpret_cast<const float4*>(wt + p * 4);
      acc[p * 4] = fmaf(c.x, v.x, acc[p * 4]); acc[p * 4 + 1] = fmaf(c.y, v.y, acc[p * 4 + 1]);
      acc[p * 4 + 2] = fmaf(c.z, v.z, acc[p * 4 + 2]); acc[p * 4 + 3] = fmaf(c.w, v.w, acc[p * 4 + 3]);
    }
  }
}

template <typename T>
__device__ __forceinline__ void store_row(T* dst, const float* r, float mul) {
#pragma unroll
  for (int p = 0; p < 8; ++p) st4(dst + p * 4, make_float4(r[p * 4] * mul, r[p * 4 + 1] * mul, r[p * 4 + 2] * mul, r[p * 4 + 3] * mul));
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) cswin_attn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ lw,
                                                            const float* __restrict__ lb, T* __restrict__ out,
                                                            float* __restrict__ lse, int R, int C, int split, int nbr,
                                                            long long ldq, long long ldo, float qscale) {
  extern __shared__ __align__(16) float sm[];
  const Unit u = decode_unit(blockIdx.x, R, C, split, nbr);
  const int b = blockIdx.y;
  const int n = u.hs * u.ws, n4 = (n + 3) & ~3;
  float* Ks = sm;
  float* Vs = Ks + n4 * RS;
  float* wsm = Vs + n4 * RS;   // [9][32]
  float* bsm = wsm + 9 * HD;   // [32]
  stage_tile<T, NT>(Ks, qkv, ldq, C + u.cb, u, b, R, n, 1.f);
  stage_tile<T, NT>(Vs, qkv, ldq, 2 * C + u.cb, u, b, R, n, 1.f);
  for (int idx = threadIdx.x; idx < (n4 - n) * RS; idx += NT) { Ks[n * RS + idx] = 0.f; Vs[n * RS + idx] = 0.f; }
  for (int idx = threadIdx.x; idx < 9 * HD; idx += NT) wsm[idx] = lw[(u.cb + (idx & 31)) * 9 + (idx >> 5)];
  if (threadIdx.x < HD) bsm[threadIdx.x] = lb[u.cb + threadIdx.x];
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= n) return;
  const long long r = tok_row(u, b, R, i);
  float q[HD], acc[HD];
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    const float4 v = ld4(qkv + r * ldq + u.cb + p * 4);
    q[p * 4] = v.x * qscale; q[p * 4 + 1] = v.y * qscale; q[p * 4 + 2] = v.z * qscale; q[p * 4 + 3] = v.w * qscale;
    acc[p * 4] = acc[p * 4 + 1] = acc[p * 4 + 2] = acc[p * 4 + 3] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < n4; j0 += 4) {
    float s[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      s[jj] = dot_row(q, Ks + (j0 + jj) * RS);
      if (j0 + jj >= n) s[jj] = -INFINITY;
    }
    const float mx = fmaxf(fmaxf(m, fmaxf(s[0], s[1])), fmaxf(s[2], s[3]));
    const float corr = ex2(m - mx);
    l *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float p = ex2(s[jj] - mx);
      l += p;
      axpy_row(acc, p, Vs + (j0 + jj) * RS);
    }
    m = mx;
  }
  const float inv = 1.f / l;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = fmaf(acc[d], inv, bsm[d]);
  lepe_row(o, Vs, wsm, u, i, 1);
  store_row(out + r * ldo + u.cb, o, 1.f);
  if (lse) lse[r * (C / HD) + u.hg] = m + log2f(l);
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) cswin_attn_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ qkv,
                                                            const T* __restrict__ out, const float* __restrict__ lse,
                                                            const float* __restrict__ lw, const float* __restrict__ lb,
                                                            T* __restrict__ dqkv, float* __restrict__ dlw, float* __restrict__ dlb,
                                                            int B, int bper, int R, int C, int split, int nbr, long long ldq,
                                                            long long ldo, long long lddo, long long lddq, float scale) {
  extern __shared__ __align__(16) float sm[];
  const Unit u = decode_unit(blockIdx.x, R, C, split, nbr);
  const int n = u.hs * u.ws, n4 = (n + 3) & ~3;
  float* Qs = sm;
  float* Ks = Qs + n4 * RS;
  float* Vs = Ks + n4 * RS;
  float* Ds = Vs + n4 * RS;
  float* lse_s = Ds + n4 * RS;
  float* del_s = lse_s + n4;
  float* wsm = del_s + n4;     // [9][32]
  float* bsm = wsm + 9 * HD;   // [32]
  for (int idx = threadIdx.x; idx < 9 * HD; idx += NT) wsm[idx] = lw[(u.cb + (idx & 31)) * 9 + (idx >> 5)];
  if (threadIdx.x < HD) bsm[threadIdx.x] = lb[u.cb + threadIdx.x];
  constexpr int NG = NT / 32, TPG = (9 + NG - 1) / NG;
  const int ch = threadIdx.x & 31, grp = threadIdx.x >> 5;
  float dwacc[TPG];
#pragma unroll
  for (int t = 0; t < TPG; ++t) dwacc[t] = 0.f;
  float dbacc = 0.f;
  const int i = threadIdx.x;
  const int b0 = blockIdx.y * bper;
  const int b1 = b0 + bper < B ? b0 + bper : B;
  for (int b = b0; b < b1; ++b) {
    __syncthreads();
    stage_tile<T, NT>(Qs, qkv, ldq, u.cb, u, b, R, n, scale * LOG2E);
    stage_tile<T, NT>(Ks, qkv, ldq, C + u.cb, u, b, R, n, 1.f);
    stage_tile<T, NT>(Vs, qkv, ldq, 2 * C + u.cb, u, b, R, n, 1.f);
    stage_tile<T, NT>(Ds, dout, lddo, u.cb, u, b, R, n, 1.f);
    __syncthreads();
    const long long r = i < n ? tok_row(u, b, R, i) : 0;
    // delta_i = dO_i . (O_i - lepe_i): the softmax-weighted part of the output only
    if (i < n) {
      float lp[HD], dO[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) lp[d] = bsm[d];
      lepe_row(lp, Vs, wsm, u, i, 1);
      load_row(dO, Ds + i * RS);
      float dl = 0.f;
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const float4 o = ld4(out + r * ldo + u.cb + p * 4);
        dl = fmaf(dO[p * 4], o.x - lp[p * 4], dl); dl = fmaf(dO[p * 4 + 1], o.y - lp[p * 4 + 1], dl);
        dl = fmaf(dO[p * 4 + 2], o.z - lp[p * 4 + 2], dl); dl = fmaf(dO[p * 4 + 3], o.w - lp[p * 4 + 3], dl);
      }
      del_s[i] = dl;
      lse_s[i] = lse[r * (C / HD) + u.hg];
    }
    __syncthreads();
    if (i < n) {
      // query row i: dq_i = scale * sum_j ds_ij k_j,  ds_ij = p_ij (dO_i . v_j - delta_i)
      {
        float q[HD], dO[HD], dq[HD];
        load_row(q, Qs + i * RS);
        load_row(dO, Ds + i * RS);
#pragma unroll
        for (int d = 0; d < HD; ++d) dq[d] = 0.f;
        const float li = lse_s[i], di = del_s[i];
        for (int j = 0; j < n; ++j) {
          const float p = ex2(dot_row(q, Ks + j * RS) - li);
          const float ds = p * (dot_row(dO, Vs + j * RS) - di);
          axpy_row(dq, ds, Ks + j * RS);
        }
        store_row(dqkv + r * lddq + u.cb, dq, scale);
      }
      // key row j (= i): dk_j = sum_i ds_ij (scale q_i),  dv_j = sum_i p_ij dO_i + lepe^T(dO)_j
      {
        float k[HD], v[HD], dk[HD], dv[HD];
        load_row(k, Ks + i * RS);
        load_row(v, Vs + i * RS);
#pragma unroll
        for (int d = 0; d < HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
        for (int t = 0; t < n; ++t) {
          const float p = ex2(dot_row(k, Qs + t * RS) - lse_s[t]);
          const float ds = p * (dot_row(v, Ds + t * RS) - del_s[t]);
          axpy_row(dk, ds, Qs + t * RS);
          axpy_row(dv, p, Ds + t * RS);
        }
        lepe_row(dv, Ds, wsm, u, i, -1);
        store_row(dqkv + r * lddq + C + u.cb, dk, 1.f / LOG2E);
        store_row(dqkv + r * lddq + 2 * C + u.cb, dv, 1.f);
      }
    }
    // LePE weight gradient: thread (channel ch, tap group grp) walks the stripe
#pragma unroll
    for (int t = 0; t < TPG; ++t) {
      const int tap = grp + t * NG;
      if (tap >= 9) break;
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      float a = 0.f;
      for (int ry = 0; ry < u.hs; ++ry) {
        const int yy = ry + dy;
        if (yy < 0 || yy >= u.hs) continue;
        for (int rx = 0; rx < u.ws; ++rx) {
          const int xx = rx + dx;
          if (xx < 0 || xx >= u.ws) continue;
          a = fmaf(Ds[(ry * u.ws + rx) * RS + ch], Vs[(yy * u.ws + xx) * RS + ch], a);
        }
      }
      dwacc[t] += a;
    }
    if (grp == 0) {
      float a = 0.f;
      for (int j = 0; j < n; ++j) a += Ds[j * RS + ch];
      dbacc += a;
    }
  }
#pragma unroll
  for (int t = 0; t < TPG; ++t) {
    const int tap = grp + t * NG;
    if (tap < 9) atomicAdd(dlw + (u.cb + ch) * 9 + tap, dwacc[t]);
  }
  if (grp == 0) atomicAdd(dlb + u.cb + ch, dbacc);
}


int attn_geom(int B, int R, int C, int split, int nbr, AttnGeom* g) {
  GA_REQUIRE(B > 0 && R > 0 && C > 0 && (nbr == 1 || nbr == 2), GA_ERR_SHAPE, "ga_cswin_attn: bad arguments B=%d R=%d C=%d nbr=%d", B, R,
             C, nbr);
  GA_REQUIRE(C % (nbr * HD) == 0, GA_ERR_SHAPE, "ga_cswin_attn: heads are %d channels wide; C=%d with %d branch(es) does not split", HD, C, nbr);
  GA_REQUIRE(nbr == 1 || (split > 0 && R % split == 0), GA_ERR_SHAPE, "ga_cswin_attn: R=%d is not a multiple of split=%d", R, split);
  g->n = nbr == 1 ? R * R : R * split;
  GA_REQUIRE(g->n <= 128, GA_ERR_UNSUPPORTED, "ga_cswin_attn: %d tokens per stripe (max 128)", g->n);
  GA_REQUIRE(B <= 65535, GA_ERR_SHAPE, "ga_cswin_attn: batch too large");
  g->n4 = (g->n + 3) & ~3;
  g->nt = g->n <= 64 ? 64 : 128;
  g->units = nbr == 1 ? C / HD : 2 * (R / split) * (C / 2 / HD);
  return GA_OK;
}
}  // namespace

static bool attn_use_tc(int dtype, long long ld_a, long long ld_b, long long ld_c, long long ld_d);
int ga_attn_fwd_tc(const void* qkv, const float* lw, const float* lb, void* out, float* lse, int B, int R, int C, int split, int nbr,
                   long long ldq, long long ldo, float scale, const AttnGeom& g, cudaStream_t st);
int ga_attn_bwd_tc(const void* dout, const void* qkv, const void* out, const float* lse, const float* lw, const float* lb, void* dqkv,
                   float* dlw, float* dlb, int B, int R, int C, int split, int nbr, long long ldq, long long ldo, long long lddo,
                   long long lddq, float scale, const AttnGeom& g, int bper, int nchunk, cudaStream_t st);
int ga_attn_fwd_tc5(const void* qkv, const float* lw, const float* lb, void* out, float* lse, int B, int R, int C, int split, int nbr,
                    long long ldq, long long ldo, float scale, const AttnGeom& g, cudaStream_t st);

extern "C" int ga_cswin_attn_fwd(const void* qkv, const float* lepe_w, const float* lepe_b, void* out, float* lse, int B, int R,
                                 int C, int split, int nbr, long long ldq, long long ldo, float scale, int dtype, int backend,
                                 ga_stream_t s) {
  AttnGeom g;
  if (int rc = attn_geom(B, R, C, split, nbr, &g)) return rc;
  GA_REQUIRE(qkv && lepe_w && lepe_b && out, GA_ERR_SHAPE, "ga_cswin_attn_fwd: null argument");
  GA_REQUIRE((ldq & 3) == 0 && (ldo & 3) == 0 && ldq >= 3 * C && ldo >= C, GA_ERR_ALIGN, "ga_cswin_attn_fwd: bad pitches");
  cudaStream_t st = (cudaStream_t)s;
  if (attn_use_tc(dtype, ldq, ldo, 8, 8)) {
    // auto: the tcgen05 kernel where it is at least as fast as the register-fragment one (measured, scripts/attn_bench.py):
    // long stripes (65..112 tokens fill the M = 128 tile) with an even number of heads per branch (no idle warpgroup)
    const bool tc5_auto = backend == GA_BACKEND_AUTO && g.n > 64 && ((C / nbr / HD) & 1) == 0;
    if ((backend == GA_BACKEND_TCGEN05 || tc5_auto) && g.n <= 112) return ga_attn_fwd_tc5(qkv, lepe_w, lepe_b, out, lse, B, R, C, split, nbr, ldq, ldo, scale, g, st);
    return ga_attn_fwd_tc(qkv, lepe_w, lepe_b, out, lse, B, R, C, split, nbr, ldq, ldo, scale, g, st);
  }
  const size_t smem = (size_t)(2 * g.n4 * RS + 10 * HD) * sizeof(float);
  const dim3 grid(g.units, B);
#define GA_ATTN_FWD(T, NT)                                                                                        \
  cswin_attn_fwd_kernel<T, NT><<<grid, NT, smem, st>>>((const T*)qkv, lepe_w, lepe_b, (T*)out, lse, R, C, split, nbr, ldq, ldo, \
                                                       scale * LOG2E)
  if (dtype == GA_F32) { if (g.nt == 64) GA_ATTN_FWD(float, 64); else GA_ATTN_FWD(float, 128); }
  else { if (g.nt == 64) GA_ATTN_FWD(bf16, 64); else GA_ATTN_FWD(bf16, 128); }
#undef GA_ATTN_FWD
  ga_count_launch();
  return ga_check_launch("cswin_attn_fwd");
}

extern "C" int ga_cswin_attn_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const float* lepe_w,
                                 const float* lepe_b, void* dqkv, float* dlepe_w, float* dlepe_b, int B, int R, int C, int split,
                                 int nbr, long long ldq, long long ldo, long long lddo, long long lddq, float scale, int dtype,
                                 ga_stream_t s) {
  AttnGeom g;
  if (int rc = attn_geom(B, R, C, split, nbr, &g)) return rc;
  GA_REQUIRE(dout && qkv && out && lse && lepe_w && lepe_b && dqkv && dlepe_w && dlepe_b, GA_ERR_SHAPE, "ga_cswin_attn_bwd: null argument");
  GA_REQUIRE(((ldq | ldo | lddo | lddq) & 3) == 0, GA_ERR_ALIGN, "ga_cswin_attn_bwd: pitches must be multiples of 4");
  const size_t smem = (size_t)(4 * g.n4 * RS + 2 * g.n4 + 10 * HD) * sizeof(float);
  int nchunk = (ga_num_sms() * 8 + g.units - 1) / g.units;
  if (nchunk > B) nchunk = B;
  if (nchunk < 1) nchunk = 1;
  const int bper = (B + nchunk - 1) / nchunk;
  nchunk = (B + bper - 1) / bper;
  const dim3 grid(g.units, nchunk);
  cudaStream_t st = (cudaStream_t)s;
  if (attn_use_tc(dtype, ldq, ldo, lddo, lddq))
    return ga_attn_bwd_tc(dout, qkv, out, lse, lepe_w, lepe_b, dqkv, dlepe_w, dlepe_b, B, R, C, split, nbr, ldq, ldo, lddo, lddq, scale, g,
                          bper, nchunk, st);
#define GA_ATTN_BWD(T, NT)                                                                                                   \
  do {                                                                                                                       \
    static GaPerDevice attr;                                                                                                 \
    if (ga_first_on_device(attr)) cudaFuncSetAttribute(cswin_attn_bwd_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); \
    cswin_attn_bwd_kernel<T, NT><<<grid, NT, smem, st>>>((const T*)dout, (const T*)qkv, (const T*)out, lse, lepe_w, lepe_b, (T*)dqkv, \
                                                         dlepe_w, dlepe_b, B, bper, R, C, split, nbr, ldq, ldo, lddo, lddq, scale); \
  } while (0)
  if (dtype == GA_F32) { if (g.nt == 64) GA_ATTN_BWD(float, 64); else GA_ATTN_BWD(float, 128); }
  else { if (g.nt == 64) GA_ATTN_BWD(bf16, 64); else GA_ATTN_BWD(bf16, 128); }
#undef GA_ATTN_BWD
  ga_count_launch();
  return ga_check_launch("cswin_attn_bwd");
}

// ================================================================================================ tensor-core path (bf16)
// Same work decomposition, but the two matrix products of the stripe run on the tensor pipe.  A stripe is at most
// 128 x 128 x 32, far below one tcgen05 tile (M = 128 per issue, operands in canonical swizzled shared-memory layouts,
// accumulators in TMEM): the op is bound by the exponentials and the gather / scatter of 64-byte token rows, not by MMA
// rate, so the products are issued as register-fragment mma.sync.m16n8k16 (bf16 in, fp32 accumulate), which lets the
// softmax stay in the accumulator registers and feed the second product without a round trip (the S -> P fragment
// identity).  One warp owns 16 query rows (forward, dq) or 16 key rows (dk, dv, through the transposed products).
namespace tcattn {
constexpr int TP = 40;   // bf16 tile row pitch (80 B): ldmatrix row addresses fall in distinct bank groups

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragment (16 rows x 16 k) of a row-major tile
__device__ __forceinline__ void frag_a(uint32_t (&r)[4], const bf16* tile, int m0, int k0, int lane) {
  ldsm4(r, smem_u32(tile + (m0 + (lane & 15)) * TP + k0 + (lane >> 4) * 8));
}
// B fragments for C = A * tile^T: 8 tile rows n0.. as the n index, all 32 k: r0,r1 = k 0..15, r2,r3 = k 16..31
__device__ __forceinline__ void frag_b_rows(uint32_t (&r)[4], const bf16* tile, int n0, int lane) {
  ldsm4(r, smem_u32(tile + (n0 + (lane & 7)) * TP + (lane >> 3) * 8));
}
// B fragments for C = A * tile: tile rows j0..j0+15 as k, columns d0..d0+15 as n: r0,r1 = n-tile d0; r2,r3 = n-tile d0+8
__device__ __forceinline__ void frag_b_cols(uint32_t (&r)[4], const bf16* tile, int j0, int d0, int lane) {
  ldsm4t(r, smem_u32(tile + (j0 + (lane & 7) + ((lane >> 3) & 1) * 8) * TP + d0 + (lane >> 4) * 8));
}

// Tile staging: NP rows x 4 sixteen-byte chunks = exactly 2 chunks per thread (NP = 16 NKT rows, 32 NKT threads).  The token
// offsets of a thread's two chunks do not depend on the image, so they are computed once (stage_offsets).
template <int NT>
__device__ __forceinline__ void stage_offsets(int (&off)[2], const Unit& u, int R, int n) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int j = (threadIdx.x + k * NT) >> 2;
    const int ry = j / u.ws, rx = j - ry * u.ws;
    off[k] = j < n ? (u.y0 + ry) * R + u.x0 + rx : -1;
  }
}
template <int NT>
__device__ __forceinline__ void stage_load(uint4 (&v)[2], const bf16* __restrict__ src, long long ld, int col, long long img, const int (&off)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int part = (threadIdx.x + k * NT) & 3;
    v[k] = make_uint4(0, 0, 0, 0);
    if (off[k] >= 0) v[k] = *reinterpret_cast<const uint4*>(src + (img + off[k]) * ld + col + part * 8);
  }
}
template <int NT>
__device__ __forceinline__ void stage_store(bf16* dst, const uint4 (&v)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = threadIdx.x + k * NT;
    *reinterpret_cast<uint4*>(dst + (idx >> 2) * TP + (idx & 3) * 8) = v[k];
  }
}

// acc[nt][0..1] (+)= sum_tap w[tap][d] * tile[nbr(i,tap)][d] for the fragment's two rows; sign as in lepe_row
__device__ __forceinline__ void lepe_frag(float (&acc)[4][4], const bf16* tile, const float* w, const Unit& u, int i_lo, int n, int t,
                                          int sign) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = i_lo + h * 8;
    if (i >= n) continue;
    const int ry = i / u.ws, rx = i - ry * u.ws;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = ry + sign * (tap / 3 - 1), xx = rx + sign * (tap % 3 - 1);
      if (yy < 0 || yy >= u.hs || xx < 0 || xx >= u.ws) continue;
      const bf16* row = tile + (yy * u.ws + xx) * TP;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int d = nt * 8 + 2 * t;
        const uint32_t v = *reinterpret_cast<const uint32_t*>(row + d);
        const float2 c = *reinterpret_cast<const float2*>(w + tap * HD + d);
        acc[nt][2 * h] = fmaf(c.x, bf16lo(v), acc[nt][2 * h]);
        acc[nt][2 * h + 1] = fmaf(c.y, bf16hi(v), acc[nt][2 * h + 1]);
      }
    }
  }
}

__device__ __forceinline__ void store_frag(bf16* dst, long long ld, const float (&acc)[4][4], long long r_lo, long long r_hi, bool ok_lo,
                                           bool ok_hi, int t, float mul) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int d = nt * 8 + 2 * t;
    if (ok_lo) *reinterpret_cast<uint32_t*>(dst + r_lo * ld + d) = pack_bf16(acc[nt][0] * mul, acc[nt][1] * mul);
    if (ok_hi) *reinterpret_cast<uint32_t*>(dst + r_hi * ld + d) = pack_bf16(acc[nt][2] * mul, acc[nt][3] * mul);
  }
}

// NKT: 16-row tiles covering the stripe (4: <= 64 tokens, 7: <= 112, 8: <= 128); one warp per tile
template <int NKT>
__global__ void __launch_bounds__(32 * NKT, NKT <= 4 ? 1 : 3) attn_fwd_tc_kernel(const bf16* __restrict__ qkv, const float* __restrict__ lw,
                                                               const float* __restrict__ lb, bf16* __restrict__ out,
                                                               float* __restrict__ lse, int R, int C, int split, int nbr,
                                                               long long ldq, long long ldo, float c2) {
  constexpr int NT = 32 * NKT, NP = 16 * NKT;
  __shared__ __align__(16) bf16 Qs[NP * TP];
  __shared__ __align__(16) bf16 Ks[NP * TP];
  __shared__ __align__(16) bf16 Vs[NP * TP];
  __shared__ __align__(16) float wsm[9 * HD];
  __shared__ float bsm[HD];
  const Unit u = decode_unit(blockIdx.x, R, C, split, nbr);
  const int b = blockIdx.y;
  const int n = u.hs * u.ws;
  {
    int off[2];
    stage_offsets<NT>(off, u, R, n);
    uint4 tq[2], tk[2], tv[2];
    const long long img = (long long)b * R * R;
    stage_load<NT>(tq, qkv, ldq, u.cb, img, off);
    stage_load<NT>(tk, qkv, ldq, C + u.cb, img, off);
    stage_load<NT>(tv, qkv, ldq, 2 * C + u.cb, img, off);
    stage_store<NT>(Qs, tq); stage_store<NT>(Ks, tk); stage_store<NT>(Vs, tv);
  }
  for (int idx = threadIdx.x; idx < 9 * HD; idx += NT) wsm[idx] = lw[(u.cb + (idx & 31)) * 9 + (idx >> 5)];
  if (threadIdx.x < HD) bsm[threadIdx.x] = lb[u.cb + threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int m0 = warp * 16;
  if (m0 >= n) return;
  uint32_t qa[2][4];
  frag_a(qa[0], Qs, m0, 0, lane);
  frag_a(qa[1], Qs, m0, 16, lane);
  float s[2 * NKT][4];
  float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
  for (int jt = 0; jt < 2 * NKT; ++jt) {
    uint32_t kb[4];
    frag_b_rows(kb, Ks, jt * 8, lane);
    s[jt][0] = s[jt][1] = s[jt][2] = s[jt][3] = 0.f;
    mma16816(s[jt], qa[0], kb[0], kb[1]);
    mma16816(s[jt], qa[1], kb[2], kb[3]);
    const int j = jt * 8 + 2 * t;
#pragma unroll
    for (int e = 0; e < 4; ++e) s[jt][e] = (j + (e & 1) < n) ? s[jt][e] * c2 : -INFINITY;
    mx_lo = fmaxf(mx_lo, fmaxf(s[jt][0], s[jt][1]));
    mx_hi = fmaxf(mx_hi, fmaxf(s[jt][2], s[jt][3]));
  }
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1)); mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1)); mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
  float l_lo = 0.f, l_hi = 0.f;
  uint32_t pa[NKT][4];
#pragma unroll
  for (int jt = 0; jt < 2 * NKT; ++jt) {
    const float p0 = ex2(s[jt][0] - mx_lo), p1 = ex2(s[jt][1] - mx_lo), p2 = ex2(s[jt][2] - mx_hi), p3 = ex2(s[jt][3] - mx_hi);
    l_lo += p0 + p1; l_hi += p2 + p3;
    pa[jt >> 1][(jt & 1) * 2] = pack_bf16(p0, p1);
    pa[jt >> 1][(jt & 1) * 2 + 1] = pack_bf16(p2, p3);
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1); l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1); l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  float o[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < NKT; ++kt) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t vb[4];
      frag_b_cols(vb, Vs, kt * 16, half * 16, lane);
      mma16816(o[2 * half], pa[kt], vb[0], vb[1]);
      mma16816(o[2 * half + 1], pa[kt], vb[2], vb[3]);
    }
  }
  const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 bb = *reinterpret_cast<const float2*>(bsm + nt * 8 + 2 * t);
    o[nt][0] = fmaf(o[nt][0], inv_lo, bb.x); o[nt][1] = fmaf(o[nt][1], inv_lo, bb.y);
    o[nt][2] = fmaf(o[nt][2], inv_hi, bb.x); o[nt][3] = fmaf(o[nt][3], inv_hi, bb.y);
  }
  const int i_lo = m0 + g, i_hi = i_lo + 8;
  lepe_frag(o, Vs, wsm, u, i_lo, n, t, 1);
  const bool ok_lo = i_lo < n, ok_hi = i_hi < n;
  const long long r_lo = ok_lo ? tok_row(u, b, R, i_lo) : 0, r_hi = ok_hi ? tok_row(u, b, R, i_hi) : 0;
  store_frag(out + u.cb, ldo, o, r_lo, r_hi, ok_lo, ok_hi, t, 1.f);
  if (lse && t == 0) {
    if (ok_lo) lse[r_lo * (C / HD) + u.hg] = mx_lo + log2f(l_lo);
    if (ok_hi) lse[r_hi * (C / HD) + u.hg] = mx_hi + log2f(l_hi);
  }
}

template <int NKT>
__global__ void __launch_bounds__(32 * NKT, NKT <= 4 ? 4 : 2) attn_bwd_tc_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ qkv,
                                                               const bf16* __restrict__ out, const float* __restrict__ lse,
                                                               const float* __restrict__ lw, const float* __restrict__ lb,
                                                               bf16* __restrict__ dqkv, float* __restrict__ dlw, float* __restrict__ dlb,
                                                               int B, int bper, int R, int C, int split, int nbr, long long ldq,
                                                               long long ldo, long long lddo, long long lddq, float scale) {
  constexpr int NT = 32 * NKT, NP = 16 * NKT;
  __shared__ __align__(16) bf16 Qs[NP * TP];
  __shared__ __align__(16) bf16 Ks[NP * TP];
  __shared__ __align__(16) bf16 Vs[NP * TP];
  __shared__ __align__(16) bf16 Ds[NP * TP];
  __shared__ __align__(16) float lse_s[NP];
  __shared__ __align__(16) float del_s[NP];
  __shared__ __align__(16) float wsm[9 * HD];
  __shared__ float bsm[HD];
  const Unit u = decode_unit(blockIdx.x, R, C, split, nbr);
  const int n = u.hs * u.ws;
  const float c2 = scale * LOG2E;
  for (int idx = threadIdx.x; idx < 9 * HD; idx += NT) wsm[idx] = lw[(u.cb + (idx & 31)) * 9 + (idx >> 5)];
  if (threadIdx.x < HD) bsm[threadIdx.x] = lb[u.cb + threadIdx.x];
  __shared__ float red[10 * HD];
  for (int idx = threadIdx.x; idx < 10 * HD; idx += NT) red[idx] = 0.f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  // LePE weight gradient: thread (channel pair cp, position slice) keeps all 9 taps + the bias of its two channels
  constexpr int NS = NT / 16;
  const int cp = threadIdx.x & 15, slice = threadIdx.x >> 4;
  float dwacc[9][2], dbacc[2] = {0.f, 0.f};
#pragma unroll
  for (int q = 0; q < 9; ++q) dwacc[q][0] = dwacc[q][1] = 0.f;
  int off[2];
  stage_offsets<NT>(off, u, R, n);
  const int b0 = blockIdx.y * bper;
  const int b1 = b0 + bper < B ? b0 + bper : B;
  for (int b = b0; b < b1; ++b) {
    __syncthreads();
    {
      uint4 tq[2], tk[2], tv[2], td[2];
      const long long img = (long long)b * R * R;
      stage_load<NT>(tq, qkv, ldq, u.cb, img, off);
      stage_load<NT>(tk, qkv, ldq, C + u.cb, img, off);
      stage_load<NT>(tv, qkv, ldq, 2 * C + u.cb, img, off);
      stage_load<NT>(td, dout, lddo, u.cb, img, off);
      stage_store<NT>(Qs, tq); stage_store<NT>(Ks, tk); stage_store<NT>(Vs, tv); stage_store<NT>(Ds, td);
    }
    __syncthreads();
    // log-sum-exp per row; padded rows get +inf so their probabilities vanish in the transposed pass
    for (int i = threadIdx.x; i < NP; i += NT) {
      lse_s[i] = i < n ? lse[tok_row(u, b, R, i) * (C / HD) + u.hg] : INFINITY;
      del_s[i] = 0.f;
    }
    __syncthreads();
    const int m0 = warp * 16;
    if (m0 < n) {
      const int i_lo = m0 + g, i_hi = i_lo + 8;
      const bool ok_lo = i_lo < n, ok_hi = i_hi < n;
      const long long r_lo = ok_lo ? tok_row(u, b, R, i_lo) : 0, r_hi = ok_hi ? tok_row(u, b, R, i_hi) : 0;
      // ---- query rows m0..m0+15: dq = scale * dS K
      {
        uint32_t qa[2][4], da[2][4];
        frag_a(qa[0], Qs, m0, 0, lane); frag_a(qa[1], Qs, m0, 16, lane);
        frag_a(da[0], Ds, m0, 0, lane); frag_a(da[1], Ds, m0, 16, lane);
        const float ls_lo = lse_s[i_lo], ls_hi = lse_s[i_hi];
        // pass 1: P (kept as bf16 pairs) and delta_i = sum_j P_ij dP_ij  (= dO_i . (P V)_i, the softmax part of the output)
        uint32_t ppk[2 * NKT][2];
        float dl_lo = 0.f, dl_hi = 0.f;
#pragma unroll
        for (int jt = 0; jt < 2 * NKT; ++jt) {
          uint32_t kb[4], vb[4];
          frag_b_rows(kb, Ks, jt * 8, lane);
          frag_b_rows(vb, Vs, jt * 8, lane);
          float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(s, qa[0], kb[0], kb[1]); mma16816(s, qa[1], kb[2], kb[3]);
          mma16816(dp, da[0], vb[0], vb[1]); mma16816(dp, da[1], vb[2], vb[3]);
          const int j = jt * 8 + 2 * t;
          float p[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) p[e] = (j + (e & 1) < n) ? ex2(fmaf(s[e], c2, -(e < 2 ? ls_lo : ls_hi))) : 0.f;
          dl_lo = fmaf(p[0], dp[0], fmaf(p[1], dp[1], dl_lo));
          dl_hi = fmaf(p[2], dp[2], fmaf(p[3], dp[3], dl_hi));
          ppk[jt][0] = pack_bf16(p[0], p[1]);
          ppk[jt][1] = pack_bf16(p[2], p[3]);
        }
        dl_lo += __shfl_xor_sync(0xffffffffu, dl_lo, 1); dl_lo += __shfl_xor_sync(0xffffffffu, dl_lo, 2);
        dl_hi += __shfl_xor_sync(0xffffffffu, dl_hi, 1); dl_hi += __shfl_xor_sync(0xffffffffu, dl_hi, 2);
        if (t == 0) { del_s[i_lo] = dl_lo; del_s[i_hi] = dl_hi; }
        // pass 2: dS = P (dP - delta), dP recomputed on the tensor pipe
        uint32_t dsa[NKT][4];
#pragma unroll
        for (int jt = 0; jt < 2 * NKT; ++jt) {
          uint32_t vb[4];
          frag_b_rows(vb, Vs, jt * 8, lane);
          float dp[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(dp, da[0], vb[0], vb[1]); mma16816(dp, da[1], vb[2], vb[3]);
          dsa[jt >> 1][(jt & 1) * 2] = pack_bf16(bf16lo(ppk[jt][0]) * (dp[0] - dl_lo), bf16hi(ppk[jt][0]) * (dp[1] - dl_lo));
          dsa[jt >> 1][(jt & 1) * 2 + 1] = pack_bf16(bf16lo(ppk[jt][1]) * (dp[2] - dl_hi), bf16hi(ppk[jt][1]) * (dp[3] - dl_hi));
        }
        float dq[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;
#pragma unroll
        for (int kt = 0; kt < NKT; ++kt) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t kb[4];
            frag_b_cols(kb, Ks, kt * 16, half * 16, lane);
            mma16816(dq[2 * half], dsa[kt], kb[0], kb[1]);
            mma16816(dq[2 * half + 1], dsa[kt], kb[2], kb[3]);
          }
        }
        store_frag(dqkv + u.cb, lddq, dq, r_lo, r_hi, ok_lo, ok_hi, t, scale);
      }
    }
    __syncthreads();
    if (m0 < n) {
      const int i_lo = m0 + g, i_hi = i_lo + 8;
      const bool ok_lo = i_lo < n, ok_hi = i_hi < n;
      const long long r_lo = ok_lo ? tok_row(u, b, R, i_lo) : 0, r_hi = ok_hi ? tok_row(u, b, R, i_hi) : 0;
      // ---- key rows m0..m0+15 through the transposed products: dk = scale * dS^T Q, dv = P^T dO + lepe^T(dO)
      {
        uint32_t ka[2][4], va[2][4];
        frag_a(ka[0], Ks, m0, 0, lane); frag_a(ka[1], Ks, m0, 16, lane);
        frag_a(va[0], Vs, m0, 0, lane); frag_a(va[1], Vs, m0, 16, lane);
        uint32_t pa[NKT][4], dsa[NKT][4];
#pragma unroll
        for (int it = 0; it < 2 * NKT; ++it) {
          uint32_t qb[4], db[4];
          frag_b_rows(qb, Qs, it * 8, lane);
          frag_b_rows(db, Ds, it * 8, lane);
          float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(s, ka[0], qb[0], qb[1]); mma16816(s, ka[1], qb[2], qb[3]);
          mma16816(dp, va[0], db[0], db[1]); mma16816(dp, va[1], db[2], db[3]);
          const int i = it * 8 + 2 * t;
          const float2 ls = *reinterpret_cast<const float2*>(lse_s + i);
          const float2 dl = *reinterpret_cast<const float2*>(del_s + i);
          float p[4], ds[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            p[e] = ex2(fmaf(s[e], c2, -((e & 1) ? ls.y : ls.x)));
            ds[e] = p[e] * (dp[e] - ((e & 1) ? dl.y : dl.x));
          }
          pa[it >> 1][(it & 1) * 2] = pack_bf16(p[0], p[1]);
          pa[it >> 1][(it & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
          dsa[it >> 1][(it & 1) * 2] = pack_bf16(ds[0], ds[1]);
          dsa[it >> 1][(it & 1) * 2 + 1] = pack_bf16(ds[2], ds[3]);
        }
        float dk[4][4], dv[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f; dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f; }
#pragma unroll
        for (int kt = 0; kt < NKT; ++kt) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t qb[4], db[4];
            frag_b_cols(qb, Qs, kt * 16, half * 16, lane);
            frag_b_cols(db, Ds, kt * 16, half * 16, lane);
            mma16816(dk[2 * half], dsa[kt], qb[0], qb[1]);
            mma16816(dk[2 * half + 1], dsa[kt], qb[2], qb[3]);
            mma16816(dv[2 * half], pa[kt], db[0], db[1]);
            mma16816(dv[2 * half + 1], pa[kt], db[2], db[3]);
          }
        }
        lepe_frag(dv, Ds, wsm, u, i_lo, n, t, -1);
        store_frag(dqkv + C + u.cb, lddq, dk, r_lo, r_hi, ok_lo, ok_hi, t, scale);
        store_frag(dqkv + 2 * C + u.cb, lddq, dv, r_lo, r_hi, ok_lo, ok_hi, t, 1.f);
      }
    }
    for (int i = slice; i < n; i += NS) {
      const int ry = i / u.ws, rx = i - ry * u.ws;
      const uint32_t dd = *reinterpret_cast<const uint32_t*>(Ds + i * TP + 2 * cp);
      const float d0 = bf16lo(dd), d1 = bf16hi(dd);
      dbacc[0] += d0; dbacc[1] += d1;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = ry + tap / 3 - 1, xx = rx + tap % 3 - 1;
        const bool ok = yy >= 0 && yy < u.hs && xx >= 0 && xx < u.ws;
        const uint32_t vv = ok ? *reinterpret_cast<const uint32_t*>(Vs + (i + (tap / 3 - 1) * u.ws + tap % 3 - 1) * TP + 2 * cp) : 0u;
        dwacc[tap][0] = fmaf(d0, bf16lo(vv), dwacc[tap][0]);
        dwacc[tap][1] = fmaf(d1, bf16hi(vv), dwacc[tap][1]);
      }
    }
  }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    atomicAdd(red + tap * HD + 2 * cp, dwacc[tap][0]);
    atomicAdd(red + tap * HD + 2 * cp + 1, dwacc[tap][1]);
  }
  atomicAdd(red + 9 * HD + 2 * cp, dbacc[0]);
  atomicAdd(red + 9 * HD + 2 * cp + 1, dbacc[1]);
  __syncthreads();
  for (int idx = threadIdx.x; idx < 10 * HD; idx += NT) {
    const int tap = idx >> 5, ch = idx & 31;
    if (tap < 9) atomicAdd(dlw + (u.cb + ch) * 9 + tap, red[idx]);
    else atomicAdd(dlb + u.cb + ch, red[idx]);
  }
}
}  // namespace tcattn

static bool attn_use_tc(int dtype, long long ld_a, long long ld_b, long long ld_c, long long ld_d) {
  static int simt = -1;
  if (simt < 0) { const char* e = getenv("GA_ATTN_SIMT"); simt = (e && atoi(e)) ? 1 : 0; }
  return dtype == GA_BF16 && !simt && ((ld_a | ld_b | ld_c | ld_d) & 7) == 0;
}

int ga_attn_fwd_tc(const void* qkv, const float* lw, const float* lb, void* out, float* lse, int B, int R, int C, int split, int nbr,
                   long long ldq, long long ldo, float scale, const AttnGeom& g, cudaStream_t st) {
  const dim3 grid(g.units, B);
  const float c2 = scale * LOG2E;
#define GA_TCF(NKT) tcattn::attn_fwd_tc_kernel<NKT><<<grid, 32 * NKT, 0, st>>>((const bf16*)qkv, lw, lb, (bf16*)out, lse, R, C, split, nbr, ldq, ldo, c2)
  if (g.n <= 64) GA_TCF(4); else if (g.n <= 112) GA_TCF(7); else GA_TCF(8);
#undef GA_TCF
  ga_count_launch();
  return ga_check_launch("cswin_attn_fwd_tc");
}

int ga_attn_bwd_tc(const void* dout, const void* qkv, const void* out, const float* lse, const float* lw, const float* lb, void* dqkv,
                   float* dlw, float* dlb, int B, int R, int C, int split, int nbr, long long ldq, long long ldo, long long lddo,
                   long long lddq, float scale, const AttnGeom& g, int bper, int nchunk, cudaStream_t st) {
  const dim3 grid(g.units, nchunk);
#define GA_TCB(NKT)                                                                                                                  \
  tcattn::attn_bwd_tc_kernel<NKT><<<grid, 32 * NKT, 0, st>>>((const bf16*)dout, (const bf16*)qkv, (const bf16*)out, lse, lw, lb, (bf16*)dqkv, \
                                                             dlw, dlb, B, bper, R, C, split, nbr, ldq, ldo, lddo, lddq, scale)
  if (g.n <= 64) GA_TCB(4); else if (g.n <= 112) GA_TCB(7); else GA_TCB(8);
#undef GA_TCB
  ga_count_launch();
  return ga_check_launch("cswin_attn_bwd_tc");
}

// ================================================================================================ tcgen05 / TMEM / TMA forward
// The same unit of work on the 5th-generation tensor cores: one CTA = (image, branch, stripe, head), 128 threads.
//   * Q, K, V of the stripe arrive by TMA: one 4-D box {64 channels, ws, hs, 1} per tensor over the [B, R, R, 3C] qkv rows lands
//     as a K-major SWIZZLE_128B tile (128-byte rows = 64 bf16, of which this head's 32 are used; rows = stripe tokens in
//     (y, x) order), so the window gather of img2windows is the tensor map and the tile is directly a UMMA operand.
//   * S = Q K^T: two tcgen05.mma (M=128, N=112, K=16 each, descriptor advanced 32 B inside the swizzle atom) into TMEM columns
//     [0,112); every thread owns one query row = one TMEM lane, reads it back with tcgen05.ld, does the softmax in registers
//     and writes P (bf16) into shared memory in the K-major swizzled layout (two 64-wide k-blocks).
//   * O = P V: seven tcgen05.mma (M=128, N=64, K=16) with V as the MN-major operand straight from its TMA tile, into TMEM
//     columns [128,192); the epilogue scales by 1/l, adds the LePE stencil (read from the swizzled V tile) and stores.
// Rows / keys beyond the stripe's n tokens: V rows >= n are zero-filled before the TMA (0 * garbage must not make NaN), score
// columns >= n are masked by index.
namespace tc5 {
constexpr uint32_t TILE_BYTES = 128 * 128;      // 128 rows x 128 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
// SWIZZLE_128B descriptor (see gemm.cu): K-major lbo 16 / sbo 1024; MN-major lbo 8192 / sbo 1024
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// byte offset of the 16-byte chunk `ch16` (0..7) of row r inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz(int r, int ch16) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((ch16 ^ (r & 7)) << 4)); }

// One CTA = (image, branch, stripe, PAIR of heads), 256 threads: warpgroup g (threads 128 g ..) owns head 2*pair + g, thread
// (t & 127) owns query row / TMEM lane (t & 127).  The 64-channel TMA box feeds both heads (k-steps 0,1 -> head A; 2,3 -> head B).
// TMEM (256 columns): S_A [0,128), S_B [128,256); O_g overwrites the first 64 columns of S_g once the softmax has consumed it.
// Shared memory: Q, K, V tiles (48 KB) + P_A (32 KB); P_B overlays Q and K, which are dead after the S products.
__global__ void __launch_bounds__(256) attn_fwd_tc5_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                                                           const float* __restrict__ lw, const float* __restrict__ lb,
                                                           bf16* __restrict__ out, float* __restrict__ lse, int R, int C, int split,
                                                           int nbr, long long ldo, float c2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Qt = sm;
  uint8_t* Kt = Qt + TILE_BYTES;
  uint8_t* Vt = Kt + TILE_BYTES;
  uint8_t* Pa = Vt + TILE_BYTES;                 // two k-blocks of [128 x 64]
  float* wsm = (float*)(Pa + 2 * TILE_BYTES);    // [9][64]
  float* bsm = wsm + 9 * 64;                     // [64]
  uint64_t* bars = (uint64_t*)(bsm + 64);        // tma, s, o
  uint32_t* tmem_slot = (uint32_t*)(bars + 3);

  // unit -> (branch, stripe, head pair)
  const int hb = C / nbr / HD, hp = (hb + 1) >> 1;
  int br = 0, stp = 0, pair = blockIdx.x;
  if (nbr == 2) {
    const int nst = R / split;
    br = blockIdx.x / (nst * hp);
    const int rem = blockIdx.x - br * nst * hp;
    stp = rem / hp;
    pair = rem - stp * hp;
  }
  Unit u;
  if (nbr == 1) { u.hs = R; u.ws = R; u.y0 = 0; u.x0 = 0; }
  else if (br == 0) { u.hs = R; u.ws = split; u.y0 = 0; u.x0 = stp * split; }
  else { u.hs = split; u.ws = R; u.y0 = stp * split; u.x0 = 0; }
  const int b = blockIdx.y;
  const int n = u.hs * u.ws;
  const int t = threadIdx.x, wg = t >> 7, row = t & 127, warp4 = (t >> 5) & 3;
  const int head = 2 * pair + wg;                // head of this warpgroup inside the branch
  const bool hvalid = head < hb;
  const int cb0 = br * (C / nbr) + 2 * pair * HD;   // first channel of the pair
  u.cb = cb0 + wg * HD;
  u.hg = br * hb + head;

  for (int idx = t; idx < (128 - n) * 8; idx += 256) {
    const int r = n + (idx >> 3);
    *reinterpret_cast<uint4*>(Vt + r * 128 + (idx & 7) * 16) = make_uint4(0, 0, 0, 0);   // whole rows: swizzle stays inside the row
  }
  for (int idx = t; idx < 9 * 64; idx += 256) {
    const int c = cb0 + (idx & 63);
    wsm[idx] = c < (br + 1) * (C / nbr) ? lw[c * 9 + (idx >> 6)] : 0.f;
  }
  if (t < 64) bsm[t] = cb0 + t < (br + 1) * (C / nbr) ? lb[cb0 + t] : 0.f;
  if (t == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if ((t >> 5) == 0) tmem_alloc(tmem_slot, 256);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zero fill before async-proxy (TMA / UMMA) accesses
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base + (uint32_t)(wg * 128), tO = tS;

  if (t == 0) {
    const CUtensorMap* tm = (nbr == 2 && br == 1) ? &tm1 : &tm0;
    mbar_expect_tx(&bars[0], 3u * (uint32_t)n * 128u);
    tma_load_4d(smem_u32(Qt), tm, &bars[0], cb0, u.x0, u.y0, b);
    tma_load_4d(smem_u32(Kt), tm, &bars[0], C + cb0, u.x0, u.y0, b);
    tma_load_4d(smem_u32(Vt), tm, &bars[0], 2 * C + cb0, u.x0, u.y0, b);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint64_t qd = make_desc(smem_u32(Qt), 16, 1024), kd = make_desc(smem_u32(Kt), 16, 1024);
    constexpr uint32_t id_s = idesc_bf16(128, 112, false, false);
#pragma unroll
    for (int k = 0; k < 4; ++k)                 // k-steps 0,1: head A channels; 2,3: head B channels
      umma_bf16(tmem_base + (uint32_t)((k >> 1) * 128), qd + (uint64_t)((k * 32) >> 4), kd + (uint64_t)((k * 32) >> 4), id_s, (k & 1) ? 1u : 0u);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  // ---- softmax of row `row` of this warpgroup's head; the stripe has n <= 112 keys
  const uint32_t lane_addr = (uint32_t)(warp4 * 32) << 16;
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld32(tS + lane_addr + (uint32_t)(c * 32), r);
#pragma unroll
    for (int j = 0; j < 32; ++j) if (c * 32 + j < n) mx = fmaxf(mx, __uint_as_float(r[j]));
  }
  const float mxs = mx * c2;
  float l = 0.f;
  uint8_t* Pt = wg == 0 ? Pa : Qt;              // P_B overlays Q | K
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    tmem_ld32(tS + lane_addr + (uint32_t)(c * 32), r);
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float p0 = (c * 32 + j < n) ? ex2(fmaf(__uint_as_float(r[j]), c2, -mxs)) : 0.f;
      const float p1 = (c * 32 + j + 1 < n) ? ex2(fmaf(__uint_as_float(r[j + 1]), c2, -mxs)) : 0.f;
      l += p0 + p1;
      pk[j >> 1] = pack_bf16(p0, p1);
    }
    uint8_t* blk = Pt + (c >> 1) * TILE_BYTES;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(blk + swz(row, (c & 1) * 4 + q)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (t == 0) {
    constexpr uint32_t id_o = idesc_bf16(128, 64, false, true);
#pragma unroll
    for (int g2 = 0; g2 < 2; ++g2) {
      const uint8_t* Pg = g2 == 0 ? Pa : Qt;
#pragma unroll
      for (int ks = 0; ks < 7; ++ks) {          // 112 keys = 7 steps of 16
        const int kb = ks >> 2, kk = ks & 3;
        const uint64_t pd = make_desc(smem_u32(Pg + kb * TILE_BYTES), 16, 1024) + (uint64_t)((kk * 32) >> 4);
        const uint64_t vd = make_desc(smem_u32(Vt + kb * 8192), 8192, 1024) + (uint64_t)((kk * 16 * 128) >> 4);
        umma_bf16(tmem_base + (uint32_t)(g2 * 128), pd, vd, id_o, ks > 0 ? 1u : 0u);
      }
    }
    umma_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld32(tO + lane_addr + (uint32_t)(wg * 32), r);   // O columns wg*32.. = this head's 32 channels of the 64-wide product
  if (row < n && hvalid) {
    const float inv = 1.f / l;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = fmaf(__uint_as_float(r[d]), inv, bsm[wg * HD + d]);
    const int ry = row / u.ws, rx = row - ry * u.ws;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = ry + tap / 3 - 1, xx = rx + tap % 3 - 1;
      if (yy < 0 || yy >= u.hs || xx < 0 || xx >= u.ws) continue;
      const int j = yy * u.ws + xx;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[8];
        ld8_bf16(reinterpret_cast<const bf16*>(Vt + swz(j, wg * 4 + q)), v);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[q * 8 + e] = fmaf(wsm[tap * 64 + wg * HD + q * 8 + e], v[e], o[q * 8 + e]);
      }
    }
    const long long grow = tok_row(u, b, R, row);
#pragma unroll
    for (int q = 0; q < 4; ++q) st8_bf16(out + grow * ldo + u.cb + q * 8, o + q * 8);
    if (lse) lse[grow * (C / HD) + u.hg] = mxs + log2f(l);
  }
  tc_fence_before();
  __syncthreads();
  if ((t >> 5) == 0) tmem_dealloc(tmem_base, 256);
}
}  // namespace tc5

int ga_attn_fwd_tc5(const void* qkv, const float* lw, const float* lb, void* out, float* lse, int B, int R, int C, int split, int nbr,
                    long long ldq, long long ldo, float scale, const AttnGeom& g, cudaStream_t st) {
  GA_REQUIRE(g.n <= 112 && (ldq & 7) == 0 && (ldo & 7) == 0 && (((uintptr_t)qkv | (uintptr_t)out) & 15) == 0, GA_ERR_UNSUPPORTED,
             "ga_cswin_attn_fwd (tcgen05): stripes of at most 112 tokens, 16-byte aligned rows");
  CUtensorMap tm0, tm1;
  const uint64_t dims[4] = {(uint64_t)(3 * C), (uint64_t)R, (uint64_t)R, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)ldq * 2, (uint64_t)R * ldq * 2, (uint64_t)R * R * ldq * 2};
  const uint32_t box0[4] = {64, (uint32_t)(nbr == 1 ? R : split), (uint32_t)R, 1};
  const uint32_t box1[4] = {64, (uint32_t)R, (uint32_t)(nbr == 1 ? R : split), 1};
  int rc = ga_tensor_map(&tm0, GA_BF16, 4, qkv, dims, strides, box0, 1);
  if (rc) return rc;
  rc = ga_tensor_map(&tm1, GA_BF16, 4, qkv, dims, strides, box1, 1);
  if (rc) return rc;
  const size_t smem = 1024 + 5 * (size_t)tc5::TILE_BYTES + (10 * 64) * sizeof(float) + 64;
  static GaPerDevice attr;
  if (ga_first_on_device(attr)) cudaFuncSetAttribute(tc5::attn_fwd_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int hb = C / nbr / HD, hp = (hb + 1) / 2;
  const dim3 grid(nbr == 1 ? hp : 2 * (R / split) * hp, B);      // one CTA per pair of heads
  tc5::attn_fwd_tc5_kernel<<<grid, 256, smem, st>>>(tm0, tm1, lw, lb, (bf16*)out, lse, R, C, split, nbr, ldo, scale * LOG2E);
  ga_count_launch();
  return ga_check_launch("cswin_attn_fwd_tc5");
}
