"""GPU: every kernel of libga_sm100.so through the C ABI against a plain PyTorch fp32 restatement of the same op.

Tolerances: fp32 path 1e-5 relative (L2), bf16 path 2e-2 relative (north_star).  Shapes include the awkward widths
of ga_convnext_tiny_688 (688, 172, 168, 2128) and ragged tails.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from imagenet_models_b200 import lib as L  # noqa: E402
from imagenet_models_b200 import ops  # noqa: E402

DEV = 'cuda'
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def tol(dtype):
    return 1e-5 if dtype == torch.float32 else 1.2e-2


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


# ------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (256, 384, 96), (300, 200, 200), (1000, 96, 384), (77, 688, 172), (513, 172, 2128),
                                   (64, 1000, 688), (130, 64, 48)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gemm_plain(M, N, K, dtype):
    A, B = rnd(M, K, dtype=dtype, seed=1), rnd(N, K, dtype=dtype, seed=2)
    D = ops.gemm(A, B)
    ref = A.float() @ B.float().t()
    assert rel(D.float(), ref) < tol(dtype)
    if dtype == torch.bfloat16 and K % 8 == 0:
        assert ops.LAST_GEMM_BACKEND == L.BACKEND_TCGEN05


@pytest.mark.parametrize('a_mn,b_mn', [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (384, 96, 1000), (200, 344, 304), (96, 384, 4096)])
def test_gemm_tc_majors(a_mn, b_mn, M, N, K):
    """MN-major operands (weight-gradient / data-gradient GEMMs read activations and weights in place)."""
    dtype = torch.bfloat16
    A = rnd(K, M, dtype=dtype, seed=3).t() if a_mn else rnd(M, K, dtype=dtype, seed=3)
    B = rnd(K, N, dtype=dtype, seed=4).t() if b_mn else rnd(N, K, dtype=dtype, seed=4)
    D = ops.gemm(A, B, out_dtype=torch.float32)
    assert ops.LAST_GEMM_BACKEND == L.BACKEND_TCGEN05
    ref = A.float() @ B.float().t()
    assert rel(D, ref) < 1e-2 * 0.2   # fp32 output of bf16 products: only accumulation-order error


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gemm_epilogues(dtype):
    M, N, K = 392, 384, 96
    A, B = rnd(M, K, dtype=dtype, seed=5), rnd(N, K, dtype=dtype, seed=6, scale=0.2)
    bias, gam = rnd(N, seed=7), rnd(N, seed=8)
    R = rnd(M, N, dtype=dtype, seed=9)
    rs = torch.tensor([1.25, 0.0], device=DEV)
    y, z = ops.gemm(A, B, bias=bias, act=L.ACT_GELU, save_z=True)
    zr = A.float() @ B.float().t() + bias
    assert rel(z.float(), zr) < tol(dtype) and rel(y.float(), F.gelu(zr)) < tol(dtype)
    y2 = ops.gemm(A, B, bias=bias, colscale=gam, rowscale=rs, rows_per_scale=196, residual=R)
    ref2 = zr * gam * rs.repeat_interleave(196)[:, None] + R.float()
    assert rel(y2.float(), ref2) < tol(dtype)
    # dgrad epilogue: acc * gelu'(Zin)
    zin = rnd(M, N, dtype=dtype, seed=10)
    y3 = ops.gemm(A, B, zin=zin, zmode=L.ACT_GELU)
    zz = zin.float().requires_grad_(True)
    F.gelu(zz).sum().backward()
    assert rel(y3.float(), (A.float() @ B.float().t()) * zz.grad) < tol(dtype)
    y4 = ops.gemm(A, B, bias=bias, act=L.ACT_RELU)
    assert rel(y4.float(), F.relu(zr)) < tol(dtype)
    # saved-derivative form: forward stores gelu'(z) next to gelu(z), backward multiplies by it (zmode ACT_MUL)
    y5, gd = ops.gemm(A, B, bias=bias, act=L.ACT_GELU, save_z='grad')
    zq = zr.clone().requires_grad_(True)
    F.gelu(zq).sum().backward()
    assert rel(y5.float(), F.gelu(zr)) < tol(dtype) and rel(gd.float(), zq.grad) < tol(dtype)
    y6 = ops.gemm(A, B, zin=gd, zmode=L.ACT_MUL)
    assert rel(y6.float(), (A.float() @ B.float().t()) * gd.float()) < tol(dtype)
    # the same with the bias gradient (column sums of the result) fused into the epilogue; several row tiles and column blocks
    M2, N2 = 1000, 768
    A2, Bt = rnd(M2, K, dtype=dtype, seed=16), rnd(K, N2, dtype=dtype, seed=17, scale=0.2)
    g2 = rnd(M2, N2, dtype=dtype, seed=18)
    sbuf = torch.zeros(N2, device=DEV)
    dz, s1 = ops.gemm_dz(A2, Bt.t(), g2, sbuf)
    ref = (A2.float() @ Bt.float()) * g2.float()
    assert rel(dz.float(), ref) < tol(dtype)
    assert rel(s1, ref.sum(0)) < (1e-5 if dtype == torch.float32 else 3e-3)
    if dtype == torch.bfloat16:
        assert s1.data_ptr() == sbuf.data_ptr()      # came out of the GEMM epilogue, not from a second pass


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gemm_wgrad_splitk_accumulate(dtype):
    T_, N, K = 5000, 384, 96      # dW[n,k] = sum_t dY[t,n] X[t,k]
    dY, X = rnd(T_, N, dtype=dtype, seed=11), rnd(T_, K, dtype=dtype, seed=12)
    dW = torch.zeros(N, K, device=DEV)
    ops.gemm(dY.t(), X.t(), dW, accumulate=True)
    ref = dY.float().t() @ X.float()
    assert rel(dW, ref) < (1e-5 if dtype == torch.float32 else 2e-3)
    ops.gemm(dY.t(), X.t(), dW, accumulate=True)     # accumulates on top
    assert rel(dW, 2 * ref) < (1e-5 if dtype == torch.float32 else 2e-3)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gemm_batched_and_strided(dtype):
    # Gram: per-image X^T X from NHWC rows
    Bn, HW, Cc = 3, 196, 192
    x = rnd(Bn * HW, Cc, dtype=dtype, seed=13)
    X3 = x.view(Bn, HW, Cc).transpose(1, 2)
    G = torch.empty(Bn, Cc, Cc, device=DEV)
    ops.gemm(X3, X3, G, alpha=0.5)
    ref = 0.5 * torch.bmm(X3.float(), X3.float().transpose(1, 2))
    assert rel(G, ref) < (1e-5 if dtype == torch.float32 else 2e-3)
    # grouped 1x1 conv with a channel-shuffled (strided-K) operand and an interleaved output
    Bm, Gp, Kc, Nc = 16, 4, 688, 172
    h = rnd(Bm, Gp * Kc, dtype=dtype, seed=14)
    W = rnd(Gp, Nc, Kc, dtype=dtype, seed=15, scale=0.1)
    A3 = h.view(Bm, Kc, Gp).permute(2, 0, 1)
    out = torch.empty(Bm, Gp * Nc, dtype=dtype, device=DEV)
    D3 = out.as_strided((Gp, Bm, Nc), (Nc, Gp * Nc, 1))
    ops.gemm(A3, W, D3)
    ref = torch.einsum('gmk,gnk->mgn', A3.float(), W.float()).reshape(Bm, Gp * Nc)
    assert rel(out.float(), ref) < tol(dtype)


def test_gemm_rejects_cpu():
    with pytest.raises(L.GaError):
        ops.gemm(torch.zeros(4, 4), torch.zeros(4, 4).cuda())


# ------------------------------------------------------------------------------------------------- K1 dwconv + LN
@pytest.mark.parametrize('Bn,H,Cc', [(2, 56, 96), (2, 28, 192), (3, 14, 384), (2, 7, 688), (2, 14, 192), (1, 9, 32), (2, 14, 128), (1, 7, 976), (1, 14, 1024)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_dwconv_ln_fwd_bwd(Bn, H, Cc, dtype):
    x = rnd(Bn * H * H, Cc, dtype=dtype, seed=20)
    w = rnd(Cc, 1, 7, 7, seed=21, scale=0.15)
    b = rnd(Cc, seed=22, scale=0.1)
    w49c = w.reshape(Cc, 49).t().contiguous()
    xhat = torch.empty_like(x)
    rstd = torch.empty(Bn * H * H, device=DEV)
    lib = L.load()
    L.check(lib.ga_dwconv7_ln_fwd(L.ptr(x), L.ptr(w49c), L.ptr(b), None, None, L.ptr(xhat), L.ptr(rstd), Bn, H, H, Cc, L.f(1e-6),
                                  L.dt(x), L.stream()), 'fwd')
    xr = x.float().view(Bn, H, H, Cc).permute(0, 3, 1, 2)
    conv = F.conv2d(xr, w, b, padding=3, groups=Cc).permute(0, 2, 3, 1).reshape(-1, Cc)
    ref = F.layer_norm(conv, (Cc,), None, None, 1e-6)
    assert rel(xhat.float(), ref) < tol(dtype)
    var = conv.var(dim=1, unbiased=False)
    assert rel(rstd, torch.rsqrt(var + 1e-6)) < (1e-5 if dtype == torch.float32 else 1e-2)
    # backward pieces: dgrad (+residual) and wgrad
    dconv = rnd(Bn * H * H, Cc, dtype=dtype, seed=23)
    dres = rnd(Bn * H * H, Cc, dtype=dtype, seed=24)
    dx = torch.empty_like(x)
    d49 = torch.zeros(49, Cc, device=DEV)
    db = torch.zeros(Cc, device=DEV)
    ws = torch.empty(lib.ga_dwconv7_bwd_parts(Bn, H, H, Cc) * 50 * Cc, device=DEV)
    L.check(lib.ga_dwconv7_bwd(L.ptr(dconv), L.ptr(x), L.ptr(dres), L.ptr(w49c), L.ptr(dx), L.ptr(d49), L.ptr(db), L.ptr(ws), Bn, H,
                               H, Cc, L.dt(x), L.dt(x), L.stream()), 'bwd')
    xg = xr.clone().requires_grad_(True)
    wg = w.clone().requires_grad_(True)
    bg = b.clone().requires_grad_(True)
    out = F.conv2d(xg, wg, bg, padding=3, groups=Cc)
    out.backward(dconv.float().view(Bn, H, H, Cc).permute(0, 3, 1, 2))
    dx_ref = xg.grad.permute(0, 2, 3, 1).reshape(-1, Cc) + dres.float()
    assert rel(dx.float(), dx_ref) < tol(dtype)
    assert rel(d49, wg.grad.reshape(Cc, 49).t()) < (2e-5 if dtype == torch.float32 else 5e-3)
    assert rel(db, bg.grad) < (2e-5 if dtype == torch.float32 else 5e-3)


@pytest.mark.parametrize('Bn,H,W,Cc', [(48, 56, 56, 96), (16, 28, 28, 192), (8, 14, 14, 384), (4, 7, 7, 688), (3, 12, 12, 192), (2, 24, 20, 96),
                                       (2, 14, 14, 512), (2, 9, 13, 256)])
def test_dwconv_bf16_second_generation_kernels(Bn, H, W, Cc):
    """dwconv3.cu at the training shapes' structure: fp32 residual-stream gradient + bf16 shadow (ga_dwconv7_bwd2), several tiles
    per CTA (batch 48 at 56x56), channel slices + cluster LayerNorm (C = 384, 512, 688), widths that are no multiple of 7."""
    dtype = torch.bfloat16
    M = Bn * H * W
    x = rnd(M, Cc, dtype=dtype, seed=30)
    w = rnd(Cc, 1, 7, 7, seed=31, scale=0.15)
    b = rnd(Cc, seed=32, scale=0.1)
    w49c = w.reshape(Cc, 49).t().contiguous()
    xhat = torch.empty_like(x)
    rstd = torch.empty(M, device=DEV)
    lib = L.load()
    L.check(lib.ga_dwconv7_ln_fwd(L.ptr(x), L.ptr(w49c), L.ptr(b), None, None, L.ptr(xhat), L.ptr(rstd), Bn, H, W, Cc, L.f(1e-6),
                                  L.dt(x), L.stream()), 'fwd')
    xr = x.float().view(Bn, H, W, Cc).permute(0, 3, 1, 2)
    conv = F.conv2d(xr, w, b, padding=3, groups=Cc).permute(0, 2, 3, 1).reshape(-1, Cc)
    ref = F.layer_norm(conv, (Cc,), None, None, 1e-6)
    assert rel(xhat.float(), ref) < tol(dtype)
    assert rel(rstd, torch.rsqrt(conv.var(dim=1, unbiased=False) + 1e-6)) < 1e-2
    dconv = rnd(M, Cc, dtype=dtype, seed=33)
    dres = rnd(M, Cc, dtype=torch.float32, seed=34)
    dx = torch.empty(M, Cc, device=DEV)
    dxs = torch.empty(M, Cc, device=DEV, dtype=dtype)
    d49 = torch.zeros(49, Cc, device=DEV)
    db = torch.zeros(Cc, device=DEV)
    ws = torch.empty(lib.ga_dwconv7_bwd_parts(Bn, H, W, Cc) * 50 * Cc, device=DEV)
    L.check(lib.ga_dwconv7_bwd2(L.ptr(dconv), L.ptr(x), L.ptr(dres), L.ptr(w49c), L.ptr(dx), L.ptr(dxs), L.ptr(d49), L.ptr(db), L.ptr(ws),
                                Bn, H, W, Cc, L.BF16, L.F32, L.stream()), 'bwd2')
    xg = xr.clone().requires_grad_(True)
    wg = w.clone().requires_grad_(True)
    bg = b.clone().requires_grad_(True)
    F.conv2d(xg, wg, bg, padding=3, groups=Cc).backward(dconv.float().view(Bn, H, W, Cc).permute(0, 3, 1, 2))
    dx_ref = xg.grad.permute(0, 2, 3, 1).reshape(-1, Cc) + dres
    assert rel(dx, dx_ref) < 1e-5                          # fp32 accumulation of exact bf16 products + fp32 residual
    assert rel(dxs.float(), dx_ref) < tol(dtype)
    assert rel(d49, wg.grad.reshape(Cc, 49).t()) < 1e-4
    assert rel(db, bg.grad) < 1e-4


@pytest.mark.parametrize('cname', ['c32_h9', 'c96_h14', 'c192_h14', 'c688_h7'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_convnext_block_vs_golden(cname, dtype, golden_dir):
    """Whole ConvNeXtBlock fwd+bwd through the autograd Function against the REFERENCE's outputs (tests/golden)."""
    import os
    from oracle import cases
    g = torch.load(os.path.join(golden_dir, 'ga_convnext_modules.pt'))[cname]
    Cc, H, Bn = cases.BLOCK_CASES[cname]
    P = {k: v.to(DEV).requires_grad_(True) for k, v in cases.block_state(Cc).items()}
    x, dy = cases.block_inputs(Cc, H, Bn)
    xr = x.to(DEV).permute(0, 2, 3, 1).reshape(-1, Cc).to(dtype).contiguous().requires_grad_(True)
    y, _ = ops.convnext_block(xr, P, (Bn, H, H), None, True)
    y.backward(dy.to(DEV).permute(0, 2, 3, 1).reshape(-1, Cc).to(dtype))
    t = 2e-5 if dtype == torch.float32 else 2e-2
    yref = g['y'].to(DEV).permute(0, 2, 3, 1).reshape(-1, Cc)
    dxref = g['dx'].to(DEV).permute(0, 2, 3, 1).reshape(-1, Cc)
    assert rel(y.float(), yref) < t
    assert rel(xr.grad.float(), dxref) < t
    for k, d in g['grads'].items():
        assert cases.digest_close(P[k].grad, d, t * (1 if dtype == torch.float32 else 1.5), 1e-5), k


# ------------------------------------------------------------------------------------------------- rows / columns
@pytest.mark.parametrize('M,Cc', [(1000, 96), (333, 688), (50, 2048), (64, 1536), (1001, 64), (77, 32), (515, 128), (9, 24)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_layernorm_fwd_bwd(M, Cc, dtype):
    x = rnd(M, Cc, dtype=dtype, seed=30).requires_grad_(True)
    w = (rnd(Cc, seed=31) * 0.2 + 1).requires_grad_(True)
    b = rnd(Cc, seed=32).requires_grad_(True)
    dy = rnd(M, Cc, dtype=dtype, seed=33)
    y = ops.layernorm(x, w, b, 1e-6)
    y.backward(dy)
    xr = x.detach().float().requires_grad_(True)
    wr, br = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    yr = F.layer_norm(xr, (Cc,), wr, br, 1e-6)
    yr.backward(dy.float())
    t = tol(dtype)
    assert rel(y.float(), yr) < t and rel(x.grad.float(), xr.grad) < t
    assert rel(w.grad, wr.grad) < (2e-5 if dtype == torch.float32 else 5e-3)
    assert rel(b.grad, br.grad) < (2e-5 if dtype == torch.float32 else 5e-3)


@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('training', [True, False])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_batchnorm(relu, training, dtype):
    M, Cc = 784, 172
    xp = ops.alloc_rows(M, Cc, dtype, DEV)
    xp.copy_(rnd(M, Cc, dtype=dtype, seed=40) * 1.5 + 0.3)
    x = xp.detach().requires_grad_(True)
    bn = {'weight': (rnd(Cc, seed=41) * 0.2 + 1).requires_grad_(True), 'bias': rnd(Cc, seed=42).requires_grad_(True),
          'running_mean': rnd(Cc, seed=43) * 0.1, 'running_var': rnd(Cc, seed=44).abs() + 0.5,
          'num_batches_tracked': torch.zeros((), dtype=torch.long, device=DEV)}
    rm0, rv0 = bn['running_mean'].clone(), bn['running_var'].clone()
    dy = rnd(M, Cc, dtype=dtype, seed=45)
    y = ops.batchnorm(x, bn, training, relu=relu)
    y.backward(dy)
    xr = x.detach().float().requires_grad_(True)
    wr, br = bn['weight'].detach().clone().requires_grad_(True), bn['bias'].detach().clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    yr = F.batch_norm(xr, rm, rv, wr, br, training, 0.1, 1e-5)
    if relu:
        yr = F.relu(yr)
    yr.backward(dy.float())
    t = tol(dtype)
    assert rel(y.float(), yr) < t
    assert rel(x.grad.float(), xr.grad) < (t if dtype == torch.float32 else 3e-2)
    assert rel(bn['weight'].grad, wr.grad) < (5e-5 if dtype == torch.float32 else 1e-2)
    assert rel(bn['bias'].grad, br.grad) < (5e-5 if dtype == torch.float32 else 1e-2)
    if training:
        assert rel(bn['running_mean'], rm) < (1e-5 if dtype == torch.float32 else 1e-2)
        assert rel(bn['running_var'], rv) < (1e-5 if dtype == torch.float32 else 1e-2)
        assert int(bn['num_batches_tracked']) == 1


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_patchify_im2col_aggregate(dtype):
    Bn, H, Cc = 2, 28, 96
    x = rnd(Bn * H * H, Cc, dtype=dtype, seed=50).requires_grad_(True)
    p = ops.patchify(x, (Bn, H, H, Cc), 2)
    xr = x.detach().float().view(Bn, H, H, Cc).permute(0, 3, 1, 2)
    ref = F.unfold(xr, 2, stride=2).view(Bn, Cc, 4, -1).permute(0, 3, 2, 1).reshape(-1, 4 * Cc)
    assert torch.equal(p.float(), ref)
    p.backward(p.detach())
    assert torch.equal(x.grad, x.detach())      # a permutation: adjoint == inverse
    # stem patchify from NCHW and channels_last storage
    img = rnd(2, 3, 32, 32, seed=51)
    ref = F.unfold(img, 4, stride=4).view(2, 3, 16, -1).permute(0, 3, 2, 1).reshape(-1, 48)
    for im in (img, img.contiguous(memory_format=torch.channels_last)):
        assert rel(ops.stem_patchify(im, 4, dtype).float(), ref) < (1e-7 if dtype == torch.float32 else 4e-3)
    # 3x3 im2col and its adjoint
    H2, C2 = 14, 172
    xs = ops.alloc_rows(Bn * H2 * H2, C2, dtype, DEV)
    xs.copy_(rnd(Bn * H2 * H2, C2, dtype=dtype, seed=52))
    xs = xs.detach().requires_grad_(True)
    col = ops.im2col3(xs, (Bn, H2, H2))
    xr = xs.detach().float().view(Bn, H2, H2, C2).permute(0, 3, 1, 2).requires_grad_(True)
    cref = F.unfold(xr, 3, padding=1).view(Bn, C2, 9, -1).permute(0, 3, 2, 1).reshape(-1, 9 * C2)
    assert torch.equal(col.float(), cref.detach())
    dcol = rnd(*col.shape, dtype=dtype, seed=53)
    col.backward(dcol)
    cref.backward(dcol.float())
    assert rel(xs.grad.float(), xr.grad.permute(0, 2, 3, 1).reshape(-1, C2)) < tol(dtype)
    # aggregation: pool 56->14, pool 28->14, copy, bilinear 7->14
    Bn = 2
    srcs = [rnd(Bn * 56 * 56, 96, dtype=dtype, seed=54), rnd(Bn * 28 * 28, 192, dtype=dtype, seed=55),
            rnd(Bn * 14 * 14, 384, dtype=dtype, seed=56), rnd(Bn * 7 * 7, 688, dtype=dtype, seed=57)]
    srcs = [s.requires_grad_(True) for s in srcs]
    spec = (Bn, 14, 14, [(56, 56, 96, 0), (28, 28, 192, 0), (14, 14, 384, 1), (7, 7, 688, 2)])
    cat = ops.aggregate(spec, srcs)
    refs = [s.detach().float().view(Bn, hw, hw, c).permute(0, 3, 1, 2).requires_grad_(True)
            for s, (hw, _, c, _) in zip(srcs, spec[3])]
    rcat = torch.cat((F.adaptive_avg_pool2d(refs[0], 14), F.adaptive_avg_pool2d(refs[1], 14), refs[2],
                      F.interpolate(refs[3], scale_factor=2, mode='bilinear')), 1)
    assert rel(cat.float(), rcat.permute(0, 2, 3, 1).reshape(-1, 1360)) < tol(dtype)
    dc = rnd(*cat.shape, dtype=dtype, seed=58)
    cat.backward(dc)
    rcat.backward(dc.float().view(Bn, 14, 14, 1360).permute(0, 3, 1, 2))
    for s, r in zip(srcs, refs):
        assert rel(s.grad.float(), r.grad.permute(0, 2, 3, 1).reshape(s.shape)) < tol(dtype)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_se_gate(dtype):
    Bn, HW, Cc, R = 3, 196, 172, 40
    xp = ops.alloc_rows(Bn * HW, Cc, dtype, DEV)
    xp.copy_(rnd(Bn * HW, Cc, dtype=dtype, seed=60))
    x = xp.detach().requires_grad_(True)
    ps = [rnd(R, Cc, 1, 1, seed=61, scale=0.1), rnd(R, seed=62, scale=0.1), rnd(Cc, R, 1, 1, seed=63, scale=0.2), rnd(Cc, seed=64, scale=0.1)]
    ps = [p.requires_grad_(True) for p in ps]
    y = ops.se_gate(x, *ps, Bn, HW)
    dy = rnd(Bn * HW, Cc, dtype=dtype, seed=65)
    y.backward(dy)
    xr = x.detach().float().view(Bn, HW, Cc).permute(0, 2, 1).reshape(Bn, Cc, 14, 14).requires_grad_(True)
    pr = [p.detach().clone().requires_grad_(True) for p in ps]
    s = xr.mean((2, 3), keepdim=True)
    s = F.conv2d(F.relu(F.conv2d(s, pr[0], pr[1])), pr[2], pr[3])
    yr = xr * torch.sigmoid(s)
    yr.backward(dy.float().view(Bn, 14, 14, Cc).permute(0, 3, 1, 2))
    t = tol(dtype)
    assert rel(y.float(), yr.permute(0, 2, 3, 1).reshape(-1, Cc)) < t
    assert rel(x.grad.float(), xr.grad.permute(0, 2, 3, 1).reshape(-1, Cc)) < t
    for a, b in zip(ps, pr):
        assert rel(a.grad, b.grad) < (5e-5 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize('cname', ['c192_h14', 'c24_h5'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_gram_vector(cname, dtype, golden_dir):
    import os
    from oracle import cases
    from oracle import ga_convnext_oracle as O
    Cc, H, Bn = cases.GRAM_CASES[cname]
    x = cases.gram_input(Cc, H, Bn)
    gold = torch.load(os.path.join(golden_dir, 'ga_convnext_modules.pt'))[f'{cname}/train0'].reshape(Bn, -1)
    xr = x.to(DEV).permute(0, 2, 3, 1).reshape(-1, Cc).to(dtype).contiguous().requires_grad_(True)
    out = ops.gram_vector(xr, Bn, H * H, float(H))      # fp32, unpadded (the reference's layout)
    assert out.dtype == torch.float32
    assert rel(out.cpu(), gold) < (1e-5 if dtype == torch.float32 else 6e-3)
    dout = rnd(*out.shape, seed=70)
    out.backward(dout)
    xo = x.clone().requires_grad_(True)
    O.gram_vector(xo, False).reshape(Bn, -1).backward(dout.cpu())
    assert rel(xr.grad.float().cpu(), xo.grad.permute(0, 2, 3, 1).reshape(-1, Cc)) < (2e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize('Q,N,H,E', [(1, 196, 8, 168), (3, 196, 12, 384), (1, 10, 8, 32)])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_attnpool(Q, N, H, E, dtype):
    nb, Bn = 2, 3
    q = rnd(nb, Bn, Q, E, seed=80, scale=0.3).requires_grad_(True)
    kvc = rnd(nb, Bn, Q, 2 * E, seed=81).requires_grad_(True)
    kvt = rnd(Bn * N, nb * 2 * E, dtype=dtype, seed=82).requires_grad_(True)
    out = ops.attnpool(q, kvc, kvt, N, H)
    dout = rnd(*out.shape, seed=83)
    out.backward(dout)
    qr, kr, tr = q.detach().clone().requires_grad_(True), kvc.detach().clone().requires_grad_(True), kvt.detach().float().requires_grad_(True)
    hd = E // H
    outs = []
    for k in range(nb):
        t = tr[:, k * 2 * E:(k + 1) * 2 * E].view(Bn, N, 2 * E)
        kk = torch.cat((kr[k][..., :E], t[..., :E]), 1).view(Bn, Q + N, H, hd).permute(0, 2, 1, 3)
        vv = torch.cat((kr[k][..., E:], t[..., E:]), 1).view(Bn, Q + N, H, hd).permute(0, 2, 1, 3)
        qq = qr[k].view(Bn, Q, H, hd).permute(0, 2, 1, 3)
        a = torch.softmax(qq @ kk.transpose(-1, -2), -1)
        outs.append((a @ vv).permute(0, 2, 1, 3).reshape(Bn, Q, E))
    ref = torch.stack(outs)
    ref.backward(dout)
    t = 1e-5 if dtype == torch.float32 else 1e-2
    assert rel(out, ref) < t
    assert rel(q.grad, qr.grad) < t and rel(kvc.grad, kr.grad) < t
    assert rel(kvt.grad.float(), tr.grad) < (t if dtype == torch.float32 else 2e-2)


def test_loss_and_adamw():
    from oracle import ga_convnext_oracle as O
    nb, Bn, ncls = 5, 6, 1000
    lg = rnd(nb, Bn, ncls, seed=90, scale=2.0).requires_grad_(True)
    y = torch.randint(0, ncls, (Bn,), device=DEV)
    loss = ops.ga_loss(lg, y, -0.8)
    loss.backward()
    lr = lg.detach().cpu().clone().requires_grad_(True)
    lref = O.ga_loss([lr[k] for k in range(nb)], y.cpu(), -0.8)
    lref.backward()
    assert abs(loss.item() - lref.item()) < 1e-4 * abs(lref.item())
    assert rel(lg.grad.cpu(), lr.grad) < 1e-5
    # fused AdamW + EMA vs torch.optim.AdamW + lerp
    n = 4096 * 3 + 8
    p = rnd(n, seed=91)
    g = rnd(n, seed=92)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05)
    m, v, ema = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV), p.clone()
    ema_r = p.clone()
    p16 = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    lib = L.load()
    for step in range(1, 4):
        pr.grad = g.clone() * step
        opt.step()
        ema_r.mul_(0.99).add_(pr.detach(), alpha=0.01)
        L.check(lib.ga_adamw_ema(L.ptr(p), L.ptr(g * step), L.ptr(m), L.ptr(v), L.ptr(ema), L.ptr(p16), None, 12, L.ll(n), L.f(1e-3),
                                 L.f(0.9), L.f(0.999), L.f(1e-8), L.f(0.05), L.f(1 - 0.9 ** step), L.f(1 - 0.999 ** step), L.f(0.99),
                                 L.f(1.0), L.stream()), 'adamw')
    assert rel(p, pr.detach()) < 1e-6 and rel(ema, ema_r) < 1e-6
    assert rel(p16.float(), p) < 4e-3


def test_attnpool_attention_dropout():
    """Attention dropout inside the pooling kernel (map.py:138): out = (softmax(qk) * mask) v, gradients through the mask."""
    Q, N, H, E, nb, Bn = 3, 196, 12, 384, 2, 3
    hd = E // H
    q = rnd(nb, Bn, Q, E, seed=84, scale=0.3).requires_grad_(True)
    kvc = rnd(nb, Bn, Q, 2 * E, seed=85).requires_grad_(True)
    kvt = rnd(Bn * N, nb * 2 * E, seed=86).requires_grad_(True)
    torch.manual_seed(5)
    mask = ops.dropout_mask((nb, Bn, H, Q, Q + N), 0.25, DEV)
    assert 0.6 < (mask > 0).float().mean().item() < 0.9 and abs(mask.max().item() - 1 / 0.75) < 1e-6
    out = ops.attnpool(q, kvc, kvt, N, H, mask)
    dout = rnd(*out.shape, seed=87)
    out.backward(dout)
    qr, kr, tr = q.detach().clone().requires_grad_(True), kvc.detach().clone().requires_grad_(True), kvt.detach().clone().requires_grad_(True)
    outs = []
    for k in range(nb):
        t = tr[:, k * 2 * E:(k + 1) * 2 * E].view(Bn, N, 2 * E)
        kk = torch.cat((kr[k][..., :E], t[..., :E]), 1).view(Bn, Q + N, H, hd).permute(0, 2, 1, 3)
        vv = torch.cat((kr[k][..., E:], t[..., E:]), 1).view(Bn, Q + N, H, hd).permute(0, 2, 1, 3)
        qq = qr[k].view(Bn, Q, H, hd).permute(0, 2, 1, 3)
        a = torch.softmax(qq @ kk.transpose(-1, -2), -1) * mask[k]
        outs.append((a @ vv).permute(0, 2, 1, 3).reshape(Bn, Q, E))
    ref = torch.stack(outs)
    ref.backward(dout)
    assert rel(out, ref) < 1e-5
    assert rel(q.grad, qr.grad) < 1e-5 and rel(kvc.grad, kr.grad) < 1e-5 and rel(kvt.grad, tr.grad) < 1e-5


@pytest.mark.parametrize('bce', [False, True])
@pytest.mark.parametrize('with_aux', [False, True])
def test_dense_target_loss(bce, with_aux):
    """Soft-target cross entropy / BCE-with-logits on dense targets (timm SoftTargetCrossEntropy, BinaryCrossEntropy) + the GA KL term
    (+ MAP's self-distillation term), against their plain PyTorch definitions."""
    nb, Bn, ncls, lam = 4, 6, 1000, -0.8
    lg = rnd(nb, Bn, ncls, seed=91, scale=2.0).requires_grad_(True)
    aux = rnd(nb, Bn, ncls, seed=92, scale=2.0).requires_grad_(True) if with_aux else None
    y = torch.randint(0, ncls, (Bn,), device=DEV)
    t = ops.smooth_one_hot(y, ncls, 0.1)
    t = 0.7 * t + 0.3 * t.flip(0)                      # a mixup target
    assert abs(t.sum(1).mean().item() - 1.0) < 1e-5
    loss = ops.ga_soft_loss(lg, t, lam, aux=aux, bce=bce)
    loss.backward()
    lr = lg.detach().clone().requires_grad_(True)
    ar = aux.detach().clone().requires_grad_(True) if with_aux else None
    ref = 0
    for k in range(nb):
        if bce:
            ref = ref + F.binary_cross_entropy_with_logits(lr[k], t)
        else:
            ref = ref + torch.sum(-t * F.log_softmax(lr[k], dim=-1), dim=-1).mean()
        if with_aux:
            ref = ref + F.kl_div(F.log_softmax(ar[k], dim=1), F.log_softmax(lr[k], dim=1).detach(), reduction='sum', log_target=True) / lr[k].numel()
    mean = F.log_softmax(lr.detach().mean(0), dim=-1)
    for k in range(nb):
        ref = ref + F.kl_div(F.log_softmax(lr[k], dim=-1), mean, reduction='mean', log_target=True) * lam
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-6
    assert rel(lg.grad, lr.grad) < 1e-5
    if with_aux:
        assert rel(aux.grad, ar.grad) < 1e-5
    with pytest.raises(L.GaError):
        ops.ga_loss(lg, t, lam)                          # dense targets must not be reinterpreted as int64 labels
    with pytest.raises(L.GaError):
        ops.ga_loss(lg, y.int(), lam)


@pytest.mark.parametrize('mode', [0, 1, 2])
def test_prep_batch_normalise_mixup_cutmix(mode):
    """uint8 -> normalised fp32 (+ mixup / cutmix with the reversed batch) in one kernel vs the timm PrefetchLoader / Mixup maths."""
    import ctypes as C
    Bn, H, W = 6, 32, 48
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (Bn, 3, H, W), dtype=torch.uint8, generator=g).cuda()
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    lam, box = 0.3, (5, 20, 8, 30)
    out = torch.empty(Bn, 3, H, W, device=DEV)
    L.check(L.load().ga_prep_batch(L.ptr(x), L.ptr(out), Bn, H, W, (C.c_float * 3)(*mean), (C.c_float * 3)(*std), mode, L.f(lam),
                                   box[0], box[1], box[2], box[3], L.stream()), 'ga_prep_batch')
    xf = x.float()
    if mode == 1:
        xf = lam * xf + (1 - lam) * xf.flip(0)
    elif mode == 2:
        xf = xf.clone()
        xf[:, :, box[0]:box[1], box[2]:box[3]] = x.float().flip(0)[:, :, box[0]:box[1], box[2]:box[3]]
    m = torch.tensor(mean, device=DEV).view(1, 3, 1, 1) * 255
    s = torch.tensor(std, device=DEV).view(1, 3, 1, 1) * 255
    assert torch.allclose(out, (xf - m) / s, atol=2e-5, rtol=1e-5)


def test_adamw_global_norm_clip_matches_torch():
    """ga_grad_clip_scale + ga_adamw_ema_dev against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW over three steps."""
    import torch.nn as nn
    from imagenet_models_b200.optim import FusedAdamWEma
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(37, 64), nn.GELU(), nn.Linear(64, 1000)).cuda()
    ref = nn.Sequential(nn.Linear(37, 64), nn.GELU(), nn.Linear(64, 1000)).cuda()
    ref.load_state_dict(net.state_dict())
    decay = [p for n, p in ref.named_parameters() if p.ndim > 1]
    no_decay = [p for n, p in ref.named_parameters() if p.ndim <= 1]
    topt = torch.optim.AdamW([{'params': decay, 'weight_decay': 0.05}, {'params': no_decay, 'weight_decay': 0.0}], lr=1e-3)
    opt = FusedAdamWEma(net, lr=1e-3, weight_decay=0.05, ema_decay=None)
    opt.clip_grad = 0.5
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        scale = 10.0 if step != 1 else 1e-3               # steps 0 and 2 exceed the norm bound, step 1 does not
        grads = [torch.randn(p.shape, generator=g).cuda() * scale for p in net.parameters()]
        opt.zero_grad()
        for p, r, gr in zip(net.parameters(), ref.parameters(), grads):
            p.grad = gr.clone()
            r.grad = gr.clone()
        opt.step()
        nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
        topt.step()
    for (n, p), r in zip(net.named_parameters(), ref.parameters()):
        assert (p - r).abs().max().item() <= 2e-6 + 1e-5 * r.abs().max().item(), n


def test_stage_with_droppath_prescaled_shadow_vs_oracle():
    """Three ConvNeXt blocks with DropPath factors (one sample dropped): the stream gradient's bf16 shadow leaves the fused depthwise
    backward already multiplied by the consumer block's factors (ga_dwconv7_bwd3); every parameter gradient must still match the
    oracle's x + ps * branch chain (ga_convnext.py:98-112 with timm DropPath)."""
    from oracle import cases
    from oracle import ga_convnext_oracle as O
    C, H, Bn, nblk = 96, 14, 4, 3
    g = torch.Generator().manual_seed(7)
    x = torch.randn(Bn, C, H, H, generator=g)
    dy = torch.randn(Bn, C, H, H, generator=g)
    states = [cases.block_state(C, seed=100 + i) for i in range(nblk)]
    for st in states:
        st['gamma'] = st['gamma'] * 0.3
    scales = [torch.tensor([1.25, 0.0, 1.25, 1.25]), torch.tensor([0.0, 1.25, 1.25, 1.25]), torch.tensor([1.25, 1.25, 1.25, 0.0])]
    # oracle (fp32 CPU)
    Ps = [{k: v.clone().requires_grad_(True) for k, v in st.items()} for st in states]
    xo = x.clone().requires_grad_(True)
    h = xo
    for P, sc in zip(Ps, scales):
        h = O.convnext_block(P, '', h, path_scale=sc)
    h.backward(dy)
    # implementation: fp32 stream + bf16 shadow, as the model runs it under autocast
    T = torch.bfloat16
    xr = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous().cuda().requires_grad_(True)
    xs = ops.to_dtype(xr, T)
    Pg = [{k: v.clone().cuda().requires_grad_(True) for k, v in st.items()} for st in states]
    sg = [s.cuda() for s in scales]
    y, ys = xr, xs
    for i in range(nblk):
        y, ys = ops.convnext_block(y, Pg[i], (Bn, H, H), sg[i], True, xs=ys, T=T, ps_prev=sg[i - 1] if i > 0 else None)
    y.backward(dy.permute(0, 2, 3, 1).reshape(-1, C).contiguous().cuda())
    assert rel(y.detach().cpu(), h.detach().permute(0, 2, 3, 1).reshape(-1, C)) < 1e-2
    assert rel(xr.grad.cpu(), xo.grad.permute(0, 2, 3, 1).reshape(-1, C)) < 1e-2
    for i in range(nblk):
        for k in Ps[i]:
            e = rel(Pg[i][k].grad.cpu().reshape(-1), Ps[i][k].grad.reshape(-1))
            assert e < 3e-2, (i, k, e)


def test_block_weight_prep_matches_per_block_ops():
    """ga_block_weight_prep (one launch for every ConvNeXt block of a model) against the per-block operand preparation it replaces
    (transpose copy, ga_fold_ln, ga_cast_bf16, ga_scale_matrix) -- bit-identical bf16 operands, fp32 folded bias to 1e-6."""
    torch.manual_seed(0)
    blocks = []
    for Cc, has_gamma in ((96, True), (192, True), (688, True), (384, False)):
        p = {'conv_dw.weight': torch.randn(Cc, 1, 7, 7, device=DEV), 'norm.weight': torch.randn(Cc, device=DEV),
             'norm.bias': torch.randn(Cc, device=DEV), 'mlp.fc1.weight': torch.randn(4 * Cc, Cc, device=DEV) * 0.1,
             'mlp.fc1.bias': torch.randn(4 * Cc, device=DEV), 'mlp.fc2.weight': torch.randn(Cc, 4 * Cc, device=DEV) * 0.1}
        if has_gamma:
            p['gamma'] = torch.randn(Cc, device=DEV)
        blocks.append(p)
    T = torch.bfloat16
    assert ops.BlockWeights.supported(blocks, T)
    bw = ops.BlockWeights(lambda: blocks)
    bw.refresh()
    torch.cuda.synchronize()
    for i, p in enumerate(blocks):
        w49c, w1f, b1f, w2c, w2s = bw.get(i)
        Cc = p['norm.weight'].numel()
        assert torch.equal(w49c, p['conv_dw.weight'].reshape(Cc, 49).t().contiguous())
        rf, rb = ops.fold_ln(p['mlp.fc1.weight'], p['mlp.fc1.bias'], p['norm.weight'], p['norm.bias'], T)
        assert torch.equal(w1f, rf)
        assert torch.allclose(b1f, rb, rtol=1e-6, atol=1e-6)
        assert torch.equal(w2c, ops.cast_like(p['mlp.fc2.weight'], T))
        g = p.get('gamma', torch.ones(Cc, device=DEV))
        assert torch.equal(w2s, ops.scale_matrix(p['mlp.fc2.weight'], g, None, T))
    # parameters updated in place are picked up by the next refresh; replaced tensors rebuild the table
    blocks[0]['mlp.fc2.weight'].mul_(2.0)
    blocks[1]['norm.weight'] = torch.randn(192, device=DEV)
    bw.refresh()
    assert torch.equal(bw.get(0)[3], ops.cast_like(blocks[0]['mlp.fc2.weight'], T))
    assert torch.equal(bw.get(1)[1], ops.fold_ln(blocks[1]['mlp.fc1.weight'], blocks[1]['mlp.fc1.bias'], blocks[1]['norm.weight'],
                                                 blocks[1]['norm.bias'], T)[0])


def test_zero_arena_hands_out_disjoint_zeroed_views():
    """ops.ZeroArena: one fill per step; views are zero, aligned, disjoint; exhaustion and first use fall back to torch.zeros."""
    A = ops.ZeroArena()
    A.begin(DEV)                                     # nothing known yet: every request is its own fill
    a = A.take(1000, torch.float32, DEV)
    b = A.take(77, torch.bfloat16, DEV)
    assert a.sum() == 0 and b.sum() == 0 and A.buf is None
    A.begin(DEV)                                     # sized by the previous step
    assert A.buf is not None and A.buf.numel() == A.need == 4096 + 256
    a = A.take(1000, torch.float32, DEV)
    b = A.take(77, torch.bfloat16, DEV)
    a.fill_(1.0)
    assert b.float().sum() == 0 and a.data_ptr() % 256 == 0 and b.data_ptr() % 256 == 0
    assert a.data_ptr() + 4000 <= b.data_ptr()
    c = A.take(10, torch.float32, DEV)               # beyond the arena: fallback, still zero
    assert c.sum() == 0 and not (A.buf.data_ptr() <= c.data_ptr() < A.buf.data_ptr() + A.buf.numel())


def test_weight_shadows_follow_the_optimizer():
    """optim.FlatState.flat16: the AdamW kernel writes the bf16 copy of every parameter; ops.cast_like returns that copy for a
    registered parameter (and its same-size views), casts afresh once the parameter was written from outside, and is
    re-registered by refresh_shadows()."""
    from imagenet_models_b200.optim import FusedAdamWEma
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(64, 48), torch.nn.Linear(48, 8)).to(DEV)
    opt = FusedAdamWEma(m, lr=1e-2, weight_decay=0.05, ema_decay=None)
    st = opt.state
    lo, hi = st.flat16.data_ptr(), st.flat16.data_ptr() + st.flat16.numel() * 2
    w = m[0].weight
    s0 = ops.cast_like(w, torch.bfloat16)
    assert lo <= s0.data_ptr() < hi and torch.equal(s0, w.detach().bfloat16())
    assert ops.cast_like(w.unsqueeze(0), torch.bfloat16).data_ptr() == s0.data_ptr()        # view of the parameter
    for p in m.parameters():
        p.grad = torch.randn_like(p)
    opt.step()
    torch.cuda.synchronize()
    s1 = ops.cast_like(w, torch.bfloat16)
    assert s1.data_ptr() == s0.data_ptr() and torch.equal(s1, w.detach().bfloat16())        # refreshed by the step kernel
    assert st.shadows_current()
    with torch.no_grad():
        w.copy_(torch.randn_like(w))                                                        # written behind the optimizer's back
    assert not st.shadows_current()
    s2 = ops.cast_like(w, torch.bfloat16)
    assert not (lo <= s2.data_ptr() < hi) and torch.equal(s2, w.detach().bfloat16())
    st.refresh_shadows()
    s3 = ops.cast_like(w, torch.bfloat16)
    assert lo <= s3.data_ptr() < hi and torch.equal(s3, w.detach().bfloat16()) and st.shadows_current()


@pytest.mark.parametrize('M,N,K', [(1000, 96, 384), (4096, 128, 512), (130, 64, 256), (802816 // 16, 96, 384)])
@pytest.mark.parametrize('b_mn', [True, False])
def test_gemm_fused_layernorm_backward(M, N, K, b_mn):
    """ga_gemm with ln_xhat / ln_rstd (EPI_LNBWD): D = LN'(A B^T) against the unfused pair (GEMM, then ga_ln_bwd_rows) and an
    fp32 torch evaluation of the same formula.  Rows beyond the last full 128-row tile, 2..4 live column slices, both operand
    layouts of B (the block's backward passes W1'.t(), an MN-major view)."""
    torch.manual_seed(M + N)
    bf = torch.bfloat16
    A = torch.randn(M, K, device=DEV, dtype=bf)
    W = (torch.randn(K, N, device=DEV) * 0.05).to(bf)            # [K, N]: W.t() is the MN-major [N, K] operand
    Bop = W.t() if b_mn else W.t().contiguous()
    xhat = torch.randn(M, N, device=DEV, dtype=bf)
    rstd = torch.rand(M, device=DEV) + 0.5
    fused = ops.gemm(A, Bop, ln_bwd=(xhat, rstd))
    assert L.BACKEND_TCGEN05 == ops.LAST_GEMM_BACKEND
    P = ops.gemm(A, Bop)                                          # bf16-rounded product, as the unfused path stores it
    unfused = torch.empty_like(P)
    L.check(L.load().ga_ln_bwd_rows(L.ptr(P), L.ptr(xhat), L.ptr(rstd), L.ptr(unfused), L.ll(M), N, L.dt(P), L.stream()), 'ga_ln_bwd_rows')
    Pf = A.float() @ W.float()
    xf = xhat.float()
    ref = rstd[:, None] * (Pf - Pf.mean(1, keepdim=True) - xf * (Pf * xf).mean(1, keepdim=True))
    err_f = ((fused.float() - ref).norm() / ref.norm()).item()
    err_u = ((unfused.float() - ref).norm() / ref.norm()).item()
    assert err_f <= 4e-3, err_f                                   # one bf16 rounding of the result
    assert err_f <= err_u * 1.05 + 1e-4, (err_f, err_u)           # never worse than the path that rounds the product first
    with pytest.raises(L.GaError):                                # a row wider than one tile cannot be fused
        ops.gemm(torch.randn(256, 64, device=DEV, dtype=bf), torch.randn(192, 64, device=DEV, dtype=bf),
                 ln_bwd=(torch.randn(256, 192, device=DEV, dtype=bf), torch.rand(256, device=DEV)))


@pytest.mark.parametrize('sub,K,N,act,has_bias,alpha', [(1, 688, 168, 'none', False, 0.1543), (1, 168, 688, 'none', True, 1.0),
                                                        (4, 172, 688, 'gelu', True, 1.0), (4, 688, 172, 'none', True, 1.0)])
def test_stacked_linear_matches_per_layer_linears(sub, K, N, act, has_bias, alpha):
    """ops.StackedLinearFn (one grouped GEMM per direction for several same-shaped layers whose weights stay separate parameters)
    against the per-layer ops.linear / grouped_linear it replaces in the GA heads: outputs, input gradient, every weight and bias
    gradient.  Shapes of the GA-ConvNeXt-T heads (q, proj, GroupConvMlp fc1 / fc2 with 172-wide groups: padded operand pitches)."""
    torch.manual_seed(sub * 1000 + K)
    nl, M = 5, 64
    G = nl * sub
    T = torch.bfloat16
    ACT = L.ACT_GELU if act == 'gelu' else L.ACT_NONE
    ws = [torch.nn.Parameter(torch.randn(sub * N, K, device=DEV) * 0.05) for _ in range(nl)]
    bs = [torch.nn.Parameter(torch.randn(sub * N, device=DEV) * 0.1) for _ in range(nl)] if has_bias else None
    a_pad = torch.zeros(G, M, ops.pad8(K), device=DEV, dtype=T)
    A3 = a_pad[:, :, :K].copy_(torch.randn(G, M, K, device=DEV)).detach().requires_grad_(True)
    odt = None if act == 'gelu' else torch.float32          # an activation keeps the operand dtype (its saved input is re-read in it)
    out = ops.stacked_linear(A3, ws, bs, act=ACT, out_dtype=odt, sub=sub, alpha=alpha)
    assert out.shape == (G, M, N)
    dout = torch.randn_like(out)
    out.backward(dout)
    got = [A3.grad.clone()] + [w.grad.clone() for w in ws] + ([b.grad.clone() for b in bs] if has_bias else [])
    A3.grad = None
    for p in ws + (bs or []):
        p.grad = None
    # reference: one grouped_linear per layer (sub groups each) on the same operands
    outs = []
    for i in range(nl):
        o = ops.grouped_linear(A3[i * sub:(i + 1) * sub], ws[i].view(sub, N, K) * alpha, bs[i] if has_bias else None, act=ACT,
                               out_dtype=odt)                                 # [M, sub*N]
        outs.append(o.view(M, sub, N).permute(1, 0, 2))
    ref = torch.cat(outs, 0)
    ref.backward(dout)
    want = [A3.grad] + [w.grad for w in ws] + ([b.grad for b in bs] if has_bias else [])
    # the reference rounds alpha * W to bf16, the stacked form applies alpha to the fp32 product: one operand rounding apart
    assert (out.float() - ref.float()).norm() <= 4e-3 * ref.float().norm()
    for g_, w_ in zip(got, want):
        assert (g_.float() - w_.float()).norm() <= 1e-2 * w_.float().norm() + 1e-6, ((g_.float() - w_.float()).norm() / w_.float().norm()).item()
