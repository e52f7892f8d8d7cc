// K7: fused multi-tensor AdamW + EMA over flat fp32 buffers, and the GA multi-branch loss (forward + gradient).
// Reference: create_optimizer_v2(... 'adamw') + ModelEmaV2.update (GA/train.py:466,499,760-761) and the loss
// expression at GA/train.py:735-745.  Both are pure HBM streams: 128-bit accesses, grid = k x SM count.
#include "common.cuh"

static inline int launch_ok(const char* n) { ga_count_launch(); return ga_check_launch(n); }

// torch.optim.AdamW (decoupled decay, eps outside the bias-corrected sqrt):
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// timm ModelEmaV2: ema = d*ema + (1-d)*p (after the step).  decay_flag is per 2^seg_shift-element segment.
__global__ void __launch_bounds__(256) adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ ema, bf16* __restrict__ p16,
                                                        const unsigned char* __restrict__ decay_flag, int seg_shift, long long n4,
                                                        float lr, float b1, float b2, float eps, float wd, float inv_bc1,
                                                        float inv_sqrt_bc2, float ema_decay, float grad_scale,
                                                        const float* __restrict__ hyper) {
  if (hyper) {   // step-dependent scalars read from device memory so a captured CUDA graph stays valid across steps
    lr = hyper[0]; inv_bc1 = 1.f / hyper[1]; inv_sqrt_bc2 = rsqrtf(hyper[2]); grad_scale = hyper[3];
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    const float decay = (decay_flag == nullptr || decay_flag[(i * 4) >> seg_shift]) ? (1.f - lr * wd) : 1.f;
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ga[4] = {gg.x, gg.y, gg.z, gg.w}, ma[4] = {mm.x, mm.y, mm.z, mm.w},
          va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * grad_scale;
      pa[k] *= decay;
      ma[k] = b1 * ma[k] + (1.f - b1) * gk;
      va[k] = b2 * va[k] + (1.f - b2) * gk * gk;
      const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
      pa[k] -= lr * inv_bc1 * ma[k] / denom;
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (ema) {
      float4 ee = reinterpret_cast<float4*>(ema)[i];
      ee.x = ema_decay * ee.x + (1.f - ema_decay) * pa[0]; ee.y = ema_decay * ee.y + (1.f - ema_decay) * pa[1];
      ee.z = ema_decay * ee.z + (1.f - ema_decay) * pa[2]; ee.w = ema_decay * ee.w + (1.f - ema_decay) * pa[3];
      reinterpret_cast<float4*>(ema)[i] = ee;
    }
    if (p16) {
      uint2 u; u.x = pack_bf16(pa[0], pa[1]); u.y = pack_bf16(pa[2], pa[3]);
      reinterpret_cast<uint2*>(p16)[i] = u;
    }
  }
}

static int adamw_launch(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, const unsigned char* decay_flag,
                        int seg_shift, long long n, float lr, float beta1, float beta2, float eps, float wd, float bias_c1, float bias_c2,
                        float ema_decay, float grad_scale, const float* hyper, ga_stream_t s) {
  GA_REQUIRE(p && g && m && v && n >= 0 && (n & 3) == 0, GA_ERR_ALIGN, "ga_adamw_ema: n=%lld must be a multiple of 4", n);
  GA_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)ema) & 15) == 0, GA_ERR_ALIGN,
             "ga_adamw_ema: buffers must be 16-byte aligned");
  GA_REQUIRE(seg_shift >= 2, GA_ERR_SHAPE, "ga_adamw_ema: segments must hold >= 4 elements");
  if (n == 0) return GA_OK;
  const long long n4 = n >> 2;
  long long blocks = (n4 + 255) / 256;
  const long long cap = (long long)ga_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  adamw_ema_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(p, g, m, v, ema, (bf16*)p_bf16, decay_flag, seg_shift, n4, lr, beta1,
                                                                   beta2, eps, wd, 1.f / bias_c1, rsqrtf(bias_c2), ema_decay, grad_scale, hyper);
  return launch_ok("adamw_ema");
}
extern "C" int ga_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16,
                            const unsigned char* decay_flag, int seg_shift, long long n, float lr, float beta1, float beta2,
                            float eps, float wd, float bias_c1, float bias_c2, float ema_decay, float grad_scale, ga_stream_t s) {
  return adamw_launch(p, g, m, v, ema, p_bf16, decay_flag, seg_shift, n, lr, beta1, beta2, eps, wd, bias_c1, bias_c2, ema_decay, grad_scale,
                      nullptr, s);
}
extern "C" int ga_adamw_ema_dev(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16,
                                const unsigned char* decay_flag, int seg_shift, long long n, const float* hyper, float beta1,
                                float beta2, float eps, float wd, float ema_decay, ga_stream_t s) {
  GA_REQUIRE(hyper, GA_ERR_SHAPE, "ga_adamw_ema_dev: hyper (lr, 1-b1^t, 1-b2^t, grad_scale) is NULL");
  return adamw_launch(p, g, m, v, ema, p_bf16, decay_flag, seg_shift, n, 0.f, beta1, beta2, eps, wd, 1.f, 1.f, ema_decay, 1.f, hyper, s);
}

// Gather per-tensor gradients into the flat gradient buffer: table[t] = {src pointer (may be NULL: zeros), flat offset,
// element count, first chunk}; chunk c of 4096 elements belongs to the tensor found by binary search over first_chunk.
struct GradEntry { const float* src; long long off; long long n; long long first_chunk; };
__global__ void __launch_bounds__(256) gather_grads_kernel(const GradEntry* __restrict__ table, int count, long long chunks,
                                                           float* __restrict__ flat) {
  for (long long c = blockIdx.x; c < chunks; c += gridDim.x) {
    int lo = 0, hi = count - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (table[mid].first_chunk <= c) lo = mid; else hi = mid - 1;
    }
    const GradEntry e = table[lo];
    const long long start = (c - e.first_chunk) * 4096;
    const long long len = e.n - start < 4096 ? e.n - start : 4096;
    float* dst = flat + e.off + start;
    const float* src = e.src ? e.src + start : nullptr;
    if (src && ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0)) {
      const long long n4 = len >> 2;
      for (long long i = threadIdx.x; i < n4; i += 256) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
      for (long long i = (n4 << 2) + threadIdx.x; i < len; i += 256) dst[i] = src[i];
    } else {
      for (long long i = threadIdx.x; i < len; i += 256) dst[i] = src ? src[i] : 0.f;
    }
  }
}
extern "C" int ga_gather_grads(const void* table, int count, long long chunks, float* flat, ga_stream_t s) {
  GA_REQUIRE(table && flat && count > 0 && chunks > 0, GA_ERR_SHAPE, "ga_gather_grads: bad arguments");
  long long blocks = chunks < 148LL * 16 ? chunks : 148LL * 16;
  gather_grads_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>((const GradEntry*)table, count, chunks, flat);
  return launch_ok("gather_grads");
}

// ema = d*ema + (1-d)*src over a flat buffer (buffers such as BatchNorm running statistics)
__global__ void ema_lerp_kernel(float* __restrict__ ema, const float* __restrict__ src, long long n, float d) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ema[i] = d * ema[i] + (1.f - d) * src[i];
}
extern "C" int ga_ema_lerp(float* ema, const float* src, long long n, float decay, ga_stream_t s) {
  if (n == 0) return GA_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ema_lerp_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(ema, src, n, decay);
  return launch_ok("ema_lerp");
}

// ---------------------------------------------------------------------------------------------- GA / MAP loss
// logits [nb][B][ncls] fp32.  loss = sum_k mean_b CE(out_k[b], y_b) + lam * sum_k mean_{b,c} KL term
//   KL_mean(logp_k || logq) with log_target: mean over B*ncls of  q*(logq - logp_k),  q = softmax(mean_k out) (detached)
// Optional aux logits (MAP's self-distillation heads, MAP/train.py:815-821):
//   + sum_k (1/(B*ncls)) sum_{b,c} p_k (logp_k - logpaux_k),  p_k = softmax(out_k) detached
// One CTA per sample: the log-softmaxes stay in smem; writes dlogits / daux and atomically adds the sample's loss.
__device__ __forceinline__ void log_softmax_row(float* row, int ncls, float* red) {
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) mx = fmaxf(mx, row[c]);
  mx = warp_max(mx);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  float se = 0.f;
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) se += expf(row[c] - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) row[c] -= lse;
  __syncthreads();
}

// Classification term per branch: hard labels (cross entropy), dense targets [B, ncls] (timm SoftTargetCrossEntropy: mixup /
// cutmix / label smoothing, GA/train.py:615-624) or BCE-with-logits on dense targets, mean over B*ncls (timm BinaryCrossEntropy,
// --bce-loss of every published recipe, GA/train.py:618-619).  mode: 0 = CE (hard or dense), 1 = BCE.
__global__ void __launch_bounds__(256) ga_loss_kernel(const float* __restrict__ logits, const float* __restrict__ aux,
                                                      const long long* __restrict__ target, const float* __restrict__ dense, int mode,
                                                      float* __restrict__ loss, float* __restrict__ dlogits, float* __restrict__ daux,
                                                      int nb, int B, int ncls, float lam, float grad_scale) {
  extern __shared__ float sm[];  // [(nb+2)][ncls]: nb branch rows, the mean row, one aux row
  __shared__ float red[32];
  const int b = blockIdx.x;
  float* lq = sm + (size_t)nb * ncls;
  float* la = lq + ncls;
  for (int c = threadIdx.x; c < ncls; c += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < nb; ++k) {
      const float v = logits[((size_t)k * B + b) * ncls + c];
      sm[(size_t)k * ncls + c] = v;
      a += v;
    }
    lq[c] = a / (float)nb;
  }
  __syncthreads();
  for (int k = 0; k <= nb; ++k) log_softmax_row(sm + (size_t)k * ncls, ncls, red);
  const int y = target ? (int)target[b] : -1;
  const float* trow = dense ? dense + (size_t)b * ncls : nullptr;
  float tsum = 1.f;
  if (trow && mode == 0) {                 // sum of the dense target row (1 for mixup / smoothing targets, kept general)
    float a = 0.f;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x) a += trow[c];
    tsum = block_sum(a, red);
    __syncthreads();
  }
  const float invB = 1.f / (float)B, inv_bc = 1.f / ((float)B * (float)ncls);
  float local = 0.f;
  for (int k = 0; k < nb; ++k) {
    const float* lp = sm + (size_t)k * ncls;
    for (int c = threadIdx.x; c < ncls; c += blockDim.x) {
      const float p = expf(lp[c]), qv = expf(lq[c]);
      local += lam * inv_bc * qv * (lq[c] - lp[c]);
      float gcls;
      if (mode == 1) {
        const float z = logits[((size_t)k * B + b) * ncls + c], t = trow ? trow[c] : (c == y ? 1.f : 0.f);
        local += (fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)))) * inv_bc;
        gcls = (1.f / (1.f + expf(-z)) - t) * inv_bc;
      } else if (trow) {
        local += -trow[c] * lp[c] * invB;
        gcls = (p * tsum - trow[c]) * invB;
      } else {
        if (c == y) local += -lp[c] * invB;
        gcls = (p - (c == y ? 1.f : 0.f)) * invB;          // dCE/dz = (p - onehot)/B
      }
      if (dlogits) {
        // d/dz_k of -lam/(B C) * sum_c q_c logp_k,c = -lam/(B C) * (q_c - p_c)
        const float gkl = -lam * inv_bc * (qv - p);
        dlogits[((size_t)k * B + b) * ncls + c] = (gcls + gkl) * grad_scale;
      }
    }
    if (aux) {
      __syncthreads();
      for (int c = threadIdx.x; c < ncls; c += blockDim.x) la[c] = aux[((size_t)k * B + b) * ncls + c];
      __syncthreads();
      log_softmax_row(la, ncls, red);
      for (int c = threadIdx.x; c < ncls; c += blockDim.x) {
        const float p = expf(lp[c]);
        local += inv_bc * p * (lp[c] - la[c]);
        if (daux) daux[((size_t)k * B + b) * ncls + c] = -inv_bc * (p - expf(la[c])) * grad_scale;
      }
    }
  }
  local = block_sum(local, red);
  if (threadIdx.x == 0) atomicAdd(loss, local);
}
static int loss_launch(const float* logits, const float* aux, const long long* target, const float* dense, int mode, float* loss,
                       float* dlogits, float* daux, int nb, int B, int ncls, float lam, float grad_scale, ga_stream_t s) {
  const size_t smem = (size_t)(nb + 2) * ncls * sizeof(float);
  GA_REQUIRE(smem <= 200 * 1024, GA_ERR_UNSUPPORTED, "ga_loss_fwd_bwd: (nb+2)*ncls too large for shared memory");
  cudaFuncSetAttribute(ga_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ga_loss_kernel<<<B, 256, smem, (cudaStream_t)s>>>(logits, aux, target, dense, mode, loss, dlogits, daux, nb, B, ncls, lam, grad_scale);
  return launch_ok("ga_loss");
}
extern "C" int ga_loss_fwd_bwd(const float* logits, const float* aux, const long long* target, float* loss, float* dlogits,
                               float* daux, int nb, int B, int ncls, float lam, float grad_scale, ga_stream_t s) {
  GA_REQUIRE(logits && target && loss && nb > 0 && B > 0 && ncls > 0, GA_ERR_SHAPE, "ga_loss_fwd_bwd: bad arguments");
  return loss_launch(logits, aux, target, nullptr, 0, loss, dlogits, daux, nb, B, ncls, lam, grad_scale, s);
}
extern "C" int ga_loss_dense_fwd_bwd(const float* logits, const float* aux, const float* dense_target, int bce, float* loss,
                                     float* dlogits, float* daux, int nb, int B, int ncls, float lam, float grad_scale, ga_stream_t s) {
  GA_REQUIRE(logits && dense_target && loss && nb > 0 && B > 0 && ncls > 0, GA_ERR_SHAPE, "ga_loss_dense_fwd_bwd: bad arguments");
  return loss_launch(logits, aux, nullptr, dense_target, bce ? 1 : 0, loss, dlogits, daux, nb, B, ncls, lam, grad_scale, s);
}

// ---------------------------------------------------------------------------------------------- LAMB (timm.optim.Lamb)
// The published recipes train with `--opt lamb` (GA/README.md:26).  timm's Lamb (absent dependency, restated in
// oracle/lamb_oracle.py):  global-norm clip g /= max(1, ||g|| / max_grad_norm);  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
// u = (m / bc1) / (sqrt(v) / sqrt(bc2) + eps) + wd p;  decayed tensors only: u *= ||p|| / ||u|| (1 when either norm is 0);
// p -= lr u.  Three passes over the flat state: (1) ||g||^2, (2) moments + update (written over the gradient buffer) +
// per-tensor ||p||^2, ||u||^2, (3) apply with the trust ratio (+ EMA).  Chunk table as in ga_gather_grads: one CTA works
// inside one tensor, so the per-tensor norms cost two atomics per CTA.
struct LambEntry { long long off; long long n; long long first_chunk; long long decay; };

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n4, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

__device__ __forceinline__ int lamb_find(const LambEntry* table, int count, long long c) {
  int lo = 0, hi = count - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].first_chunk <= c) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// hyper = {lr, 1-b1^t, 1-b2^t, grad_scale}; gnorm_sq: sum of squares of the UNSCALED flat gradient
__global__ void __launch_bounds__(256) lamb_update_kernel(const LambEntry* __restrict__ table, int count, long long chunks,
                                                          const float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, float* __restrict__ norms, const float* __restrict__ hyper,
                                                          const float* __restrict__ gnorm_sq, float b1, float b2, float eps, float wd,
                                                          float max_grad_norm) {
  __shared__ float red[32];
  const float inv_bc1 = 1.f / hyper[1], inv_sqrt_bc2 = rsqrtf(hyper[2]), gs = hyper[3];
  const float gn = sqrtf(*gnorm_sq) * gs;
  const float clip = (max_grad_norm > 0.f && gn > max_grad_norm) ? max_grad_norm / gn : 1.f;
  const float gmul = gs * clip;
  for (long long c = blockIdx.x; c < chunks; c += gridDim.x) {
    const int t = lamb_find(table, count, c);
    const LambEntry e = table[t];
    const long long start = (c - e.first_chunk) * 4096;
    const long long len = e.n - start < 4096 ? e.n - start : 4096;
    const long long base = e.off + start;
    const float wdt = e.decay ? wd : 0.f;
    float pn = 0.f, un = 0.f;
    for (long long i = threadIdx.x; i < len; i += 256) {
      const float gi = g[base + i] * gmul, pi = p[base + i];
      const float mi = b1 * m[base + i] + (1.f - b1) * gi;
      const float vi = b2 * v[base + i] + (1.f - b2) * gi * gi;
      m[base + i] = mi;
      v[base + i] = vi;
      const float u = (mi * inv_bc1) / (sqrtf(vi) * inv_sqrt_bc2 + eps) + wdt * pi;
      g[base + i] = u;
      pn += pi * pi;
      un += u * u;
    }
    pn = block_sum(pn, red);
    un = block_sum(un, red);
    if (threadIdx.x == 0 && e.decay) { atomicAdd(norms + 2 * t, pn); atomicAdd(norms + 2 * t + 1, un); }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) lamb_apply_kernel(const LambEntry* __restrict__ table, int count, long long chunks,
                                                         float* __restrict__ p, const float* __restrict__ u, float* __restrict__ ema,
                                                         const float* __restrict__ norms, const float* __restrict__ hyper, float ema_decay) {
  const float lr = hyper[0];
  for (long long c = blockIdx.x; c < chunks; c += gridDim.x) {
    const int t = lamb_find(table, count, c);
    const LambEntry e = table[t];
    float trust = 1.f;
    if (e.decay) {
      const float wn = sqrtf(norms[2 * t]), un = sqrtf(norms[2 * t + 1]);
      trust = (wn > 0.f && un > 0.f) ? wn / un : 1.f;
    }
    const long long start = (c - e.first_chunk) * 4096;
    const long long len = e.n - start < 4096 ? e.n - start : 4096;
    const long long base = e.off + start;
    const float step = lr * trust;
    for (long long i = threadIdx.x; i < len; i += 256) {
      const float pi = p[base + i] - step * u[base + i];
      p[base + i] = pi;
      if (ema) ema[base + i] = ema_decay * ema[base + i] + (1.f - ema_decay) * pi;
    }
  }
}

// Global-norm gradient clipping for the AdamW path (timm dispatch_clip_grad(mode='norm') = torch.nn.utils.clip_grad_norm_,
// GA/train.py:321-327,758): hyper[3] (the gradient scale the optimizer kernel applies) is multiplied by
// min(1, max_norm / (||g * hyper[3]|| + 1e-6)).  push_hyper() rewrites hyper[3] before every step.
__global__ void clip_scale_kernel(const float* __restrict__ sumsq, float* __restrict__ hyper, float max_norm, float* __restrict__ norm_out) {
  const float norm = sqrtf(*sumsq) * hyper[3];
  if (norm_out) *norm_out = norm;
  const float coef = max_norm / (norm + 1e-6f);
  if (coef < 1.f) hyper[3] *= coef;
}
extern "C" int ga_grad_clip_scale(const float* g, long long n, float max_norm, float* hyper, float* scratch /* [2]: sumsq, norm */,
                                  ga_stream_t s) {
  GA_REQUIRE(g && hyper && scratch && n > 0 && (n & 3) == 0 && max_norm > 0.f, GA_ERR_SHAPE, "ga_grad_clip_scale: bad arguments");
  cudaStream_t st = (cudaStream_t)s;
  cudaMemsetAsync(scratch, 0, 2 * sizeof(float), st);
  const long long n4 = n >> 2;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, n4, scratch);
  int rc = launch_ok("clip_sumsq");
  if (rc) return rc;
  clip_scale_kernel<<<1, 1, 0, st>>>(scratch, hyper, max_norm, scratch + 1);
  return launch_ok("clip_scale");
}

extern "C" int ga_lamb_ema(float* p, float* g, float* m, float* v, float* ema, const void* table, int count, long long chunks,
                           long long n_flat, float* scratch /* [1 + 2*count] */, const float* hyper, float beta1, float beta2,
                           float eps, float wd, float max_grad_norm, float ema_decay, ga_stream_t s) {
  GA_REQUIRE(p && g && m && v && table && scratch && hyper && count > 0 && chunks > 0 && (n_flat & 3) == 0, GA_ERR_SHAPE,
             "ga_lamb_ema: bad arguments");
  cudaStream_t st = (cudaStream_t)s;
  cudaMemsetAsync(scratch, 0, (size_t)(1 + 2 * count) * sizeof(float), st);
  const long long n4 = n_flat >> 2;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, st>>>(g, n4, scratch);
  int rc = launch_ok("lamb_sumsq");
  if (rc) return rc;
  long long cb = chunks < 148LL * 16 ? chunks : 148LL * 16;
  lamb_update_kernel<<<(unsigned)cb, 256, 0, st>>>((const LambEntry*)table, count, chunks, p, g, m, v, scratch + 1, hyper, scratch, beta1,
                                                   beta2, eps, wd, max_grad_norm);
  rc = launch_ok("lamb_update");
  if (rc) return rc;
  lamb_apply_kernel<<<(unsigned)cb, 256, 0, st>>>((const LambEntry*)table, count, chunks, p, g, ema, scratch + 1, hyper, ema_decay);
  return launch_ok("lamb_apply");
}
