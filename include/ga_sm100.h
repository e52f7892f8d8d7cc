/* ga_sm100.h -- C ABI of libga_sm100.so: the sm_100a kernels under the GA-ConvNeXt / GA-CSWin / MAP hot path.
 *
 * The reference (Lab-LVM/imagenet-models) has NO FFI: its hot path is torch.nn modules that dispatch to
 * cuDNN/cuBLAS/ATen.  This header is therefore the NEW seam underneath the
 * drop-in nn.Modules; every entry point cites the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *  - plain pointers + sizes, no torch types; every buffer (incl. workspaces) is owned by the caller;
 *  - activations are NHWC-contiguous ("channels_last" storage), dtype GA_F32 or GA_BF16;
 *    parameters and all statistics are fp32;
 *  - every call launches on the given stream and returns immediately (no sync, no allocation);
 *  - return 0 on success, non-zero otherwise; ga_last_error(buf, n) copies the calling thread's message;
 *  - no process-global mutable state: backends are per-call arguments, the only caches are the immutable tensor-map /
 *    function-attribute caches (mutex guarded);
 *    nothing throws or exits across this boundary.
 */
#ifndef GA_SM100_H
#define GA_SM100_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ga_stream_t; /* cudaStream_t */

enum { GA_F32 = 0, GA_BF16 = 1 };
enum { GA_ACT_NONE = 0, GA_ACT_GELU = 1, GA_ACT_RELU = 2, GA_ACT_MUL = 3 /* zmode only: D = acc * Zin */ };
enum { GA_BACKEND_AUTO = 0, GA_BACKEND_SIMT = 1, GA_BACKEND_TCGEN05 = 2 };

int ga_version(void);
/* message of the calling thread's last failed call, copied into buf (NUL-terminated, truncated to n); returns its length */
int ga_last_error(char* buf, size_t n);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long ga_launch_count(void);
/* bytes of caller-owned fp32 scratch an entry point needs: GA_WS_DWCONV7_BWD (B, H, W, C) -> dw_partial;
 * GA_WS_COLSTATS (M, C, 0, 0) -> ga_colstats / ga_bn_bwd_reduce workspace; GA_WS_LAYERNORM_BWD (M, C, 0, 0).  -1 = unknown op */
enum { GA_WS_DWCONV7_BWD = 1, GA_WS_COLSTATS = 2, GA_WS_LAYERNORM_BWD = 3 };
long long ga_workspace_bytes(int op, long long a, long long b, long long c, long long d);

/* ---- GEMM with fused epilogue: every nn.Linear / 1x1 / k=s conv on the path --------------------------------
 *   D[b][m][n] = epi( alpha * sum_k A[b](m,k) * B[b](n,k) )
 *   epi(v): v += bias[n]; Z = v (optional save); v = act(v); v *= colscale[n]; v *= rowscale[m / rows_per_scale];
 *           v += R[m][n];   or, when Zin is given,   v = v * act'(Zin[m][n])   (GELU' / ReLU mask)
 * Replaces: Mlp fc1+GELU / fc2 + gamma + shortcut (ga_convnext.py:107-111), stem/downsample convs (:127,:357),
 * Bottleneck 1x1 convs (:260,:269,:282), gram_contraction/embedding (:407,:418), ClassAttn q/k/v/proj (:163-167),
 * GroupConvMlp (:202,:205), fc (:422), torch.bmm in get_gram (:460) and all their autograd backward GEMMs.
 * bf16 operands with unit stride along m/n or k run on tcgen05 (TMA-fed, TMEM accumulators); fp32 operands and
 * irregular strides run on the fp32 SIMT kernel. */
typedef struct GaGemm {
  const void* A; long long a_rs, a_cs, a_bs; /* element strides of A: per m, per k, per batch */
  const void* B; long long b_rs, b_cs, b_bs; /* element strides of B: per n, per k, per batch */
  void* D; long long ldd, d_bs, d_cs;        /* row / batch / column strides of D (d_cs 0 or 1 = contiguous) */
  int M, N, K, batch;
  int in_dtype;   /* dtype of A and B */
  int out_dtype;  /* dtype of D, Z, R, Zin */
  int accumulate; /* D += ... (fp32 D only; atomics, enables split-K; no bias/act/R/Zin) */
  float alpha;
  const float* bias; long long bias_bs;
  int act;
  void* Z;                 /* optional pre-activation save, same ld / batch stride as D */
  const float* colscale; long long colscale_bs;
  const float* rowscale; int rows_per_scale;
  const void* R; long long ldr, r_bs;
  const void* Zin; long long ldz, z_bs; int zmode; /* GA_ACT_GELU: *gelu'(Zin); GA_ACT_RELU: *(Zin>0); GA_ACT_MUL: *Zin */
  int backend;    /* GA_BACKEND_* */
  int splits;     /* split-K factor for accumulate mode; 0 = auto */
  int z_shadow;   /* 1: Z receives a bf16 copy of the FINAL value (bf16 shadow of an fp32 residual stream);
                     2: Z receives act'(pre-activation) (GELU), to be applied in backward with zmode GA_ACT_MUL */
  int* backend_used; /* optional out: GA_BACKEND_* this call ran on */
  float* colsum;  /* optional, with Zin on the tcgen05 path only: colsum[n] += sum_m D[m,n] (fp32, 16-byte aligned; the bias
                     gradient of the layer whose dz this GEMM produces); GA_ERR_UNSUPPORTED when it cannot be fused */
  /* optional fused LayerNorm backward of the output rows (the dxhat GEMM of a ConvNeXt block, ga_convnext.py:105-107 backward):
     D = rstd * (P - mean_n(P) - xhat * mean_n(P * xhat)) with P = A B^T, xhat [M,N] in the operand dtype (row pitch ld_xhat),
     rstd [M] fp32.  tcgen05 path, bf16, 32 < N <= 128, no bias / activation; otherwise GA_ERR_UNSUPPORTED (nothing launched) */
  const void* ln_xhat; long long ld_xhat; const float* ln_rstd;
} GaGemm;
int ga_gemm(const GaGemm* p, ga_stream_t s);

/* ---- K1: depthwise 7x7 conv (+bias) fused with LayerNorm over C  (ga_convnext.py:100,105-106) ---------------
 * x [B,H,W,C] -> y = LN(conv7x7(x)+bias): xhat when ln_w == NULL (the affine is then folded into fc1 by the
 * caller), else xhat*ln_w+ln_b;  rstd[B*H*W] saved.  w49c is conv_dw.weight[C,1,7,7] re-laid out tap-major [49][C]. */
int ga_dwconv7_ln_fwd(const void* x, const float* w49c, const float* bias, const float* ln_w, const float* ln_b,
                      void* y, float* rstd, int B, int H, int W, int C, float eps, int dtype, ga_stream_t s);
/* LN backward on rows with xhat saved: dconv = rstd*(dxhat - mean(dxhat) - xhat*mean(dxhat*xhat)) */
int ga_ln_bwd_rows(const void* dxhat, const void* xhat, const float* rstd, void* dconv, long long M, int C,
                   int dtype, ga_stream_t s);
/* same with the stream gradient added: out = res + LN'(dxhat) in res_dtype (fp32 or `dtype`), plus an optional
 * `dtype` shadow of the sum -- the residual form of a pre-norm transformer block (ga_cswin.py:198, 209-210) */
int ga_ln_bwd_rows_res(const void* dxhat, const void* xhat, const float* rstd, const void* res, void* out, void* shadow,
                       long long M, int C, int dtype, int res_dtype, ga_stream_t s);
/* dwconv backward: dx = corr7(dconv, flipped w) + dres  (dx may be NULL);  dw49c += sum dconv*x(shifted);
 * dbias += sum dconv.  dw_partial: [ga_dwconv7_bwd_parts()][50][C] fp32 workspace. */
int ga_dwconv7_bwd_parts(int B, int H, int W, int C);
int ga_dwconv7_bwd(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx,
                   float* dw49c, float* dbias, float* dw_partial, int B, int H, int W, int C, int dtype,
                   int res_dtype /* dtype of dres and dx: the residual-stream gradient */, ga_stream_t s);

/* same, additionally writing dx_shadow (may be NULL): dx in the compute dtype `dtype`, the operand copy the previous
 * block's backward GEMMs read (saves a separate fp32 -> bf16 pass over the stream gradient) */
int ga_dwconv7_bwd2(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dx_shadow,
                    float* dw49c, float* dbias, float* dw_partial, int B, int H, int W, int C, int dtype,
                    int res_dtype, ga_stream_t s);
/* same; shadow_rowscale [B] (may be NULL): dx_shadow = dx * shadow_rowscale[image] -- the DropPath factor of the block that consumes
 * the shadow as its backward GEMM operand (timm DropPath on the residual branch, ga_convnext.py:111), so that block needs no
 * separate row-scaling pass.  GA_ERR_UNSUPPORTED when the fused bf16 kernel does not take the shape (caller scales itself).
 * dw_partial == NULL with dbias == dw49c + 49 C: the per-CTA sums are added straight into dw49c / dbias (fp32 atomics; no workspace
 * clear, no reduction launch); GA_ERR_UNSUPPORTED when the shape needs the workspace.  Nothing is launched before that answer. */
int ga_dwconv7_bwd3(const void* dconv, const void* x, const void* dres, const float* w49c, void* dx, void* dx_shadow,
                    const float* shadow_rowscale, float* dw49c, float* dbias, float* dw_partial, int B, int H, int W, int C,
                    int dtype, int res_dtype, ga_stream_t s);

/* ---- row LayerNorm over the last dim (LayerNorm2d on NHWC rows, nn.LayerNorm)  (ga_convnext.py:51-67,233,237) */
int ga_layernorm_fwd(const void* x, const float* w, const float* b, void* y, float* mean, float* rstd,
                     long long M, int C, long long ldx, long long ldy, float eps, int dtype, ga_stream_t s);
/* dx = LN backward (mean == NULL: x already holds xhat); dw/db accumulated (+=) via a [parts][2][C] workspace, or, with
 * partial == NULL, by fp32 atomics from the one kernel (no second launch; summation order, hence the last bits, not fixed) */
int ga_layernorm_bwd_parts(long long M, int C);
int ga_layernorm_bwd(const void* dy, const void* x, const float* w, const float* mean, const float* rstd,
                     void* dx, float* dw, float* db, float* partial, long long M, int C, long long lddy,
                     long long ldx, long long lddx, int dtype, ga_stream_t s);

/* ---- patch gathers (im2col is a permutation for kernel==stride)  (ga_convnext.py:127,357; Bottleneck 3x3 :273) */
/* NHWC [B,H,W,C] -> rows [B*(H/k)*(W/k), k*k*C] ordered (ky,kx,c); inverse = the backward scatter */
int ga_patchify(const void* x, void* y, int B, int H, int W, int C, int k, int inverse, int dtype, ga_stream_t s);
/* stem input: fp32 image [B,3,H,W] with element strides (sb,sc,sy,sx) (NCHW or channels_last storage)
 * -> patch rows [B*(H/k)*(W/k), k*k*3] ordered (ky,kx,c) in `dtype` */
int ga_stem_patchify(const float* x, void* y, int B, int H, int W, int k, long long sb, long long sc, long long sy,
                     long long sx, int dtype, ga_stream_t s);
/* 3x3 pad-1 im2col: NHWC rows (stride ldx) -> [B*H*W, ldy>=9C] ordered (tap,c); inverse = col2im in gather form */
int ga_im2col3(const void* x, void* y, int B, int H, int W, int C, long long ldx, long long ldy, int inverse,
               int dtype, ga_stream_t s);
/* the same with stride 1 or 2 (CSWin Merge_Block / deep stem, ga_cswin.py:256, 467, 472): output map ((H-1)/stride+1)^2 */
int ga_im2col3s(const void* x, void* y, int B, int H, int W, int C, int stride, long long ldx, long long ldy,
                int inverse, int dtype, ga_stream_t s);
/* CSWin deep-stem first conv input (ga_cswin.py:463): fp32 image [B,3,H,W] with element strides -> 3x3/pad-1 patch rows
 * [B*Ho*Wo, 32]: 27 columns ordered (ky,kx,c) + 5 zero columns */
int ga_stem_im2col3(const float* x, void* y, int B, int H, int W, int stride, long long sb, long long sc, long long sy,
                    long long sx, int dtype, ga_stream_t s);

/* ---- column statistics over rows of [M,C]: BatchNorm2d (ga_convnext.py:261,270,276,283,409,420), bias grads --- */
int ga_colstats_parts(long long M, int C);
/* sum[c] (=|+=) sum_m x[m,c]; sumsq[c] likewise with x^2 (either may be NULL); partial: [parts][2][C].
 * partial == NULL (here, in ga_colstats_shifted and in ga_bn_bwd_reduce): single-kernel mode -- every CTA adds its partial sums
 * to the outputs with fp32 atomics, so the outputs must be zero (or hold the value to add to; ga_colstats then needs
 * accumulate = 1) and the summation order is not fixed; with the workspace a second kernel reduces in a fixed order. */
int ga_colstats(const void* x, float* sum, float* sumsq, float* partial, long long M, int C, long long ldx,
                int accumulate, int dtype, ga_stream_t s);
/* the same sums of (x - pivot), pivot[c] = x[0,c] (written): BatchNorm statistics without the E[x^2] - E[x]^2 cancellation */
int ga_colstats_shifted(const void* x, float* pivot, float* sum, float* sumsq, float* partial, long long M, int C,
                        long long ldx, int dtype, ga_stream_t s);
/* training: mean/invstd from the sums (of x - pivot when pivot != NULL), running stats updated (momentum, unbiased var);
 * eval: from running stats.  scale = w*invstd, shift = b - mean*scale */
int ga_bn_finalize(const float* sum, const float* sumsq, const float* pivot, const float* w, const float* b, float* running_mean,
                   float* running_var, float* mean, float* invstd, float* scale, float* shift, long long M, int C,
                   float momentum, float eps, int training, ga_stream_t s);
/* y = act( x*scale[c] + shift[c]  (+ x2*scale2[c] + shift2[c]) ): BN apply (+ReLU) and the Bottleneck merge
 * relu(bn3(conv3) + bn(shortcut)) (ga_convnext.py:298-316); scale2 may be NULL (plain residual) */
int ga_affine_act(const void* x, const float* scale, const float* shift, const void* x2, const float* scale2,
                  const float* shift2, void* y, long long M, int C, long long ldx, long long ldx2, long long ldy,
                  int act, int dtype, ga_stream_t s);
/* BN backward pass 1: c1[c] = sum_m d, c2[c] = sum_m d*xhat with d = dy*(y>0 if relu), xhat=(x-mean)*invstd */
int ga_bn_bwd_reduce(const void* dy, const void* x, const void* y, const float* mean, const float* invstd, float* c1,
                     float* c2, float* partial, long long M, int C, long long lddy, long long ldx, long long ldy,
                     int relu, int dtype, ga_stream_t s);
/* BN backward pass 2: dx = scale[c]*(d - c1[c]/M - xhat*c2[c]/M); c1 == NULL (eval mode): dx = scale[c]*d */
int ga_bn_bwd_apply(const void* dy, const void* x, const void* y, const float* mean, const float* invstd,
                    const float* scale, const float* c1, const float* c2, void* dx, long long M, int C,
                    long long lddy, long long ldx, long long ldy, long long lddx, int relu, int dtype, ga_stream_t s);

/* ---- K3: multi-scale aggregation and SE gate  (ga_convnext.py:479-483, :305; map.py:322-331) ------------------
 * mode 0: avg-pool by integer factor (56->14, 28->14); 1: copy; 2: bilinear x2 (align_corners=False); 3: non-antialiased
 * bilinear shrink by an even factor (MAP MultiScale, map.py:328); 4: replication (adaptive_avg_pool2d to a larger map).  Writes channel
 * slice [coff, coff+C) of dst [B,Ho,Wo,ldd].  inverse=1: adjoint, src (contiguous) receives the gradient of the slice. */
int ga_aggregate(const void* src, void* dst, int B, int Hs, int Ws, int C, int Ho, int Wo, long long ldd, int coff,
                 int mode, int inverse, int dtype, ga_stream_t s);
/* timm SEModule: gate = sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2); y = x*gate.  pooled[B,C] hidden[B,R] gate[B,C] saved */
int ga_se_fwd(const void* x, const float* w1, const float* b1, const float* w2, const float* b2, void* y,
              float* pooled, float* hidden, float* gate, int B, int HW, int C, int R, long long ldx, long long ldy,
              int dtype, ga_stream_t s);
/* dx plus per-image pre-activation gradients dpre2[B,C], dh[B,R] (the four weight gradients are tiny GEMMs over them) */
int ga_se_bwd(const void* dy, const void* x, const float* w1, const float* w2, const float* hidden, const float* gate,
              void* dx, float* dpre2, float* dh, int B, int HW, int C, int R, long long lddy, long long ldx,
              long long lddx, int dtype, ga_stream_t s);

/* ---- K4: Gram -> upper triangle -> L2 normalise  (get_gram, ga_convnext.py:452-467; map.py:217-227) -----------
 * G [B,C,C] fp32 comes from ga_gemm (X^T X, alpha = 1/(div^2 HW)).  out[b, (t/glen)*gld + t%glen] = triu(G)_t / max(||triu||,1e-12)
 * with t the row-major i<=j enumeration; (glen, gld) let the caller pad each conv group to a 16-byte multiple. */
int ga_gram_triu_fwd(const float* G, void* out, float* norm, int B, int C, int glen, int gld, long long out_bs,
                     int out_dtype, int interleave /* n_tokens of map.GramToken's token interleave, map.py:225-227; 1 = none */,
                     ga_stream_t s);
/* backward through normalise + gather: S[b] = dG + dG^T (C x C, diagonal doubled); dX = alpha * X S is a ga_gemm */
int ga_gram_triu_bwd(const void* dout, const void* out, const float* norm, void* S, int B, int C, int glen, int gld,
                     long long out_bs, int io_dtype, int s_dtype, int interleave, ga_stream_t s);

/* the same vector as two bf16 terms hi + lo (~16 mantissa bits) for the embedding conv that feeds a train-mode BatchNorm over
 * the batch (gram_embedding ga_convnext.py:417-420, bp_reduction map.py:203-206): that BatchNorm amplifies operand rounding ~20x */
int ga_gram_triu_fwd_split(const float* G, void* hi, void* lo, float* norm, int B, int C, int glen, int gld,
                           long long out_bs, int interleave, ga_stream_t s);
int ga_gram_triu_bwd_split(const void* dout, const void* hi, const void* lo, const float* norm, void* S, int B, int C,
                           int glen, int gld, long long out_bs, int s_dtype, int interleave, ga_stream_t s);

/* ---- K5: attention pooling: Q query tokens against Q+N keys  (ClassAttn ga_convnext.py:170-183; map.py:100-144)
 * q [B,Q,E] fp32 pre-scaled; kv_cls [B,Q,2E] fp32 (k | v of the query tokens); kv_tok rows [B*N, ldt] (k at col 0,
 * v at col E); H heads.  out [B,Q,E] fp32; attn [B,H,Q,Q+N] fp32 saved for backward. */
/* ---- K6: CSWin stripe attention + LePE  (ga_cswin.py:59-136, img2windows :215, windows2img :225) --------------------
 * qkv rows [B*R*R, 3C] (q | k | v blocks, pitch ldq); heads are 32 channels.  nbr == 2: branch 0 = channels [0,C/2) over
 * R x split stripes, branch 1 = [C/2,C) over split x R stripes; nbr == 1: one R x R window.  <= 128 tokens per stripe.
 * lepe_w [C,9] / lepe_b [C]: the branches' get_v depthwise 3x3 weights, concatenated over channels.
 * out [B*R*R, C] = softmax(scale q k^T) v + dw3x3(v inside the stripe);  lse [B*R*R, C/32] fp32 (log2 units) or NULL.
 * backend (bf16, stripes of <= 112 tokens): GA_BACKEND_AUTO = tcgen05 kernel (Q/K/V by 4-D TMA boxes straight into UMMA operand
 * tiles, S and O accumulated in TMEM, two heads per CTA) for 65..112-token stripes with an even head count per branch, else the
 * register-fragment mma.sync kernel; GA_BACKEND_SIMT forces the mma.sync kernel, GA_BACKEND_TCGEN05 the tcgen05 one. */
int ga_cswin_attn_fwd(const void* qkv, const float* lepe_w, const float* lepe_b, void* out, float* lse, int B, int R,
                      int C, int split, int nbr, long long ldq, long long ldo, float scale, int dtype, int backend,
                      ga_stream_t s);
/* dqkv [B*R*R, 3C] is fully overwritten; dlepe_w / dlepe_b are accumulated (+=, atomics) */
int ga_cswin_attn_bwd(const void* dout, const void* qkv, const void* out, const float* lse, const float* lepe_w,
                      const float* lepe_b, void* dqkv, float* dlepe_w, float* dlepe_b, int B, int R, int C, int split,
                      int nbr, long long ldq, long long ldo, long long lddo, long long lddq, float scale, int dtype,
                      ga_stream_t s);

int ga_attnpool_fwd(const float* q, const float* kv_cls, const void* kv_tok, float* out, float* attn, int B, int Q,
                    int N, int H, int E, long long ldt, int dtype,
                    const float* drop_mask /* attention dropout (map.py:138): [B,H,Q,Q+N] multipliers 0 or 1/(1-p), or NULL */,
                    ga_stream_t s);
int ga_attnpool_bwd(const float* dout, const float* q, const float* kv_cls, const void* kv_tok, const float* attn,
                    float* dq, float* dkv_cls, void* dkv_tok, int B, int Q, int N, int H, int E, long long ldt,
                    long long lddt, int dtype, const float* drop_mask, ga_stream_t s);

/* ---- small fused pieces -------------------------------------------------------------------------------------*/
/* dst[m, c] = src[m, c] for c < C with row strides and dtype conversion (concat slices, casts) */
int ga_copy_cols(const void* src, void* dst, long long M, int C, long long lds, long long ldd, int src_dtype,
                 int dst_dtype, ga_stream_t s);
/* fp32 -> bf16 cast of a flat buffer (weight shadows) */
int ga_cast_bf16(const float* src, void* dst, long long n, ga_stream_t s);
/* dst[r, c] = src[r, c] * rowscale[r] * colscale[c]  (either may be NULL): folds layer-scale / LN affine into weights */
/* LayerNorm-affine fold into the Linear that follows (ga_convnext.py:105-107, ga_cswin.py:196-197, 210):
 * Wf[n,k] = W[n,k]*ln_w[k] in dst_dtype (row pitch ldw), bf[n] = bias[n] + sum_k W[n,k]*ln_b[k] (bias may be NULL) */
int ga_fold_ln(const float* W, const float* ln_w, const float* ln_b, const float* bias, void* Wf, float* bf, int N, int K,
               long long ldw, int dst_dtype, ga_stream_t s);
/* Operand preparation of every ConvNeXt block of a model in ONE launch (the per-step work that GA/ga_convnext.py:100,105-111
 * leaves to cuDNN/cuBLAS weight layouts): for block b, `table` holds GA_BLOCK_PREP_WORDS int64 words
 *   [0] conv_dw.weight [C,1,7,7] f32   [1] norm.weight [C]   [2] norm.bias [C]   [3] fc1.weight [4C,C]   [4] fc1.bias [4C]
 *   [5] fc2.weight [C,4C]   [6] gamma [C] or 0      (inputs, fp32)
 *   [7] taps [49,C] f32   [8] fc1.weight*diag(norm.weight) bf16 [4C,C]   [9] fc1.bias + fc1.weight norm.bias f32 [4C]
 *   [10] bf16 fc2.weight [C,4C]   [11] bf16 diag(gamma) fc2.weight [C,4C]      (outputs)
 *   [12] C (multiple of 8)   [13] first unit of the block = sum over earlier blocks of (5C + 49)
 * total_units = sum over all blocks of (5C + 49). */
#define GA_BLOCK_PREP_WORDS 14
int ga_block_weight_prep(const void* table, int nblocks, int total_units, ga_stream_t s);
int ga_scale_matrix(const float* src, const float* rowscale, const float* colscale, void* dst, int rows, int cols,
                    int dst_dtype, ga_stream_t s);
/* y[r,:] = x[r,:] * rowscale[r / rows_per_scale]  (DropPath mask applied to a gradient; timm drop_path) */
int ga_scale_rows(const void* x, const float* rowscale, void* y, long long M, int C, int rows_per_scale, int dtype,
                  ga_stream_t s);
/* dz = dy * gelu'(z) (act = GA_ACT_GELU, zy = saved pre-activation) or dy * (y > 0) (GA_ACT_RELU, zy = saved output) */
int ga_act_bwd(const void* dy, const void* zy, void* dz, long long M, int C, long long lddy, long long ldz,
               long long lddz, int act, int dtype, ga_stream_t s);
/* Finalise a Linear's gradients from G = dOut^T In (fp32 [N,K]) when a per-output scale g[N] and/or a per-input
 * affine (w[K], b_in[K]) was folded into the forward weights (layer-scale gamma, LayerNorm affine): the layer
 * computed out = g * (W (In*w + b_in) + bias).  s = column sums of dOut:
 *   dW[n,k] += g[n]*(G[n,k]*w[k] + s[n]*b_in[k]);  dg[n] += sum_k W[n,k]*w[k]*G[n,k] + bias[n]*s[n];
 *   dbias[n] += g[n]*s[n];  dw[k] += sum_n W[n,k]*g[n]*G[n,k];  db_in[k] += sum_n W[n,k]*g[n]*s[n]
 * (any output may be NULL; dg as written assumes b_in == NULL, which holds for the layer-scale use) */
int ga_linear_grad_finalize(const float* G, const float* s, const float* W, const float* bias, const float* g,
                            const float* w, const float* b_in, float* dW, float* dbias, float* dg, float* dw,
                            float* db_in, int N, int K, ga_stream_t st);

/* ---- loss: sum_k CE(out_k, y) + lam * sum_k KL_mean(logsm(out_k) || logsm(mean_k out).detach())  (GA/train.py:735-745)
 * plus, when aux != NULL (MAP's self-distillation heads, MAP/train.py:815-821, 833-837 with lam = dec_lam):
 *   sum_k KL_sum(logsm(aux_k) || logsm(out_k).detach()) / numel.
 * logits, aux: [nb][B][ncls] fp32; dlogits / daux same shape (may be NULL), scaled by grad_scale; loss[0] += value */
int ga_loss_fwd_bwd(const float* logits, const float* aux, const long long* target, float* loss, float* dlogits,
                    float* daux, int nb, int B, int ncls, float lam, float grad_scale, ga_stream_t s);
/* the same with DENSE targets [B, ncls] fp32 (mixup / cutmix / label smoothing: timm SoftTargetCrossEntropy, GA/train.py:615-624)
 * or, bce != 0, BCE-with-logits averaged over B*ncls (timm BinaryCrossEntropy, --bce-loss, GA/train.py:618-619) */
int ga_loss_dense_fwd_bwd(const float* logits, const float* aux, const float* dense_target, int bce, float* loss,
                          float* dlogits, float* daux, int nb, int B, int ncls, float lam, float grad_scale, ga_stream_t s);

/* ---- K7: fused multi-tensor AdamW + EMA over flat fp32 buffers  (GA/train.py:466,499,760-761; timm ModelEmaV2)
 * torch.optim.AdamW update of p from g (scaled by grad_scale) with state m, v; bias_c1 = 1-beta1^t, bias_c2 = 1-beta2^t.
 * Weight decay applies to segment i>>seg_shift where decay_flag != 0 (NULL: everywhere).  ema (optional):
 * ema = d*ema + (1-d)*p after the step.  p_bf16 (optional) receives the bf16 shadow of the updated parameters. */
int ga_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16, const unsigned char* decay_flag,
                 int seg_shift, long long n, float lr, float beta1, float beta2, float eps, float wd, float bias_c1,
                 float bias_c2, float ema_decay, float grad_scale, ga_stream_t s);
/* ema = d*ema + (1-d)*src (buffers: BatchNorm running statistics) */
/* ga_adamw_ema with the step-dependent scalars in device memory: hyper = {lr, 1-beta1^t, 1-beta2^t, grad_scale}
 * (so one captured CUDA graph of the training step stays valid while lr and t change) */
int ga_adamw_ema_dev(float* p, const float* g, float* m, float* v, float* ema, void* p_bf16,
                     const unsigned char* decay_flag, int seg_shift, long long n, const float* hyper, float beta1,
                     float beta2, float eps, float wd, float ema_decay, ga_stream_t s);
/* global-norm clipping for the AdamW path (torch.nn.utils.clip_grad_norm_ via timm dispatch_clip_grad, GA/train.py:321-327):
 * hyper[3] *= min(1, max_norm / (||g|| * hyper[3] + 1e-6));  scratch[0] = sum g^2, scratch[1] = the (scaled) norm */
int ga_grad_clip_scale(const float* g, long long n, float max_norm, float* hyper, float* scratch, ga_stream_t s);
/* input pipeline step in front of the stem (timm PrefetchLoader + Mixup, GA/train.py:545-557,598-626): uint8 NCHW batch ->
 * fp32 (x - mean*255) / (std*255), optionally mixed with the batch in reverse order: mode 1 mixup out = lam*a + (1-lam)*b,
 * mode 2 cutmix (b inside the box [y0,y1) x [x0,x1)).  One pass. */
int ga_prep_batch(const unsigned char* x, float* y, int B, int H, int W, const float* mean3, const float* std3, int mode,
                  float lam, int y0, int y1, int x0, int x1, ga_stream_t s);
/* Per-tensor gradients -> the flat gradient buffer the optimizer / all-reduce buckets read (replaces one accumulate
 * kernel per parameter).  table: device array of `count` records {const float* src (NULL = zeros); long long flat_offset;
 * long long numel; long long first_chunk}, chunks of 4096 elements numbered consecutively over the table. */
int ga_gather_grads(const void* table, int count, long long chunks, float* flat, ga_stream_t s);
/* LAMB as timm.optim.Lamb does it (GA/README.md:26 trains with --opt lamb; GA/train.py:466 create_optimizer_v2) + EMA:
 * global gradient-norm clip (max_grad_norm, 0 = off), Adam moments with bias correction, update + wd*p, per-tensor trust
 * ratio ||p||/||update|| for decayed tensors, p -= lr*trust*update.  table: device array of `count` records
 * {long long flat_offset, numel, first_chunk (4096-element chunks), decay (0/1)}; hyper = {lr, 1-b1^t, 1-b2^t, grad_scale}
 * in device memory; scratch: 1 + 2*count floats; g is overwritten with the update. */
int ga_lamb_ema(float* p, float* g, float* m, float* v, float* ema, const void* table, int count, long long chunks,
                long long n_flat, float* scratch, const float* hyper, float beta1, float beta2, float eps, float wd,
                float max_grad_norm, float ema_decay, ga_stream_t s);
int ga_ema_lerp(float* ema, const float* src, long long n, float decay, ga_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* GA_SM100_H */
