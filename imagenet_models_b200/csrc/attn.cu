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
#include "common.cuh"

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

struct AttnGeom {
  int units, n, n4, nt;
};

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

extern "C" int ga_cswin_attn_fwd(const void* qkv, const float* lepe_w, const float* lepe_b, void* out, float* lse, int B, int R,
                                 int C, int split, int nbr, long long ldq, long long ldo, float scale, int dtype, ga_stream_t s) {
  AttnGeom g;
  if (int rc = attn_geom(B, R, C, split, nbr, &g)) return rc;
  GA_REQUIRE(qkv && lepe_w && lepe_b && out, GA_ERR_SHAPE, "ga_cswin_attn_fwd: null argument");
  GA_REQUIRE((ldq & 3) == 0 && (ldo & 3) == 0 && ldq >= 3 * C && ldo >= C, GA_ERR_ALIGN, "ga_cswin_attn_fwd: bad pitches");
  const size_t smem = (size_t)(2 * g.n4 * RS + 10 * HD) * sizeof(float);
  const dim3 grid(g.units, B);
  cudaStream_t st = (cudaStream_t)s;
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
#define GA_ATTN_BWD(T, NT)                                                                                                   \
  do {                                                                                                                       \
    static bool attr = false;                                                                                                \
    if (!attr) { cudaFuncSetAttribute(cswin_attn_bwd_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; } \
    cswin_attn_bwd_kernel<T, NT><<<grid, NT, smem, st>>>((const T*)dout, (const T*)qkv, (const T*)out, lse, lepe_w, lepe_b, (T*)dqkv, \
                                                         dlepe_w, dlepe_b, B, bper, R, C, split, nbr, ldq, ldo, lddo, lddq, scale); \
  } while (0)
  if (dtype == GA_F32) { if (g.nt == 64) GA_ATTN_BWD(float, 64); else GA_ATTN_BWD(float, 128); }
  else { if (g.nt == 64) GA_ATTN_BWD(bf16, 64); else GA_ATTN_BWD(bf16, 128); }
#undef GA_ATTN_BWD
  ga_count_launch();
  return ga_check_launch("cswin_attn_bwd");
}
