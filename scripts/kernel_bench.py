#!/usr/bin/env python3
"""Per-kernel micro-benchmarks at the GA-ConvNeXt-T (batch 256) shapes: achieved GB/s / TFLOP/s vs MEASURED_PEAKS.json.

  python scripts/kernel_bench.py [gemm|dwconv|rows|all] [--one NAME]     (CUDA-event timing, L2-exceeding inputs)
`--one NAME` runs a single case three times (for `ncu --set full -k regex:... -c 1`).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from imagenet_models_b200 import lib as L  # noqa: E402
from imagenet_models_b200 import ops  # noqa: E402

DEV = 'cuda'
B = 256


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d['hbm_gbs'], d['bf16_tflops']
    return 6650.0, 1590.0


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3    # us


def gemm_cases():
    bf = torch.bfloat16
    cases = {}

    def fwd(name, M, K, N, gelu=False, res=False):
        A = torch.randn(M, K, device=DEV, dtype=bf)
        W = torch.randn(N, K, device=DEV, dtype=bf) * 0.05
        bias = torch.randn(N, device=DEV)
        if gelu:
            out = torch.empty(M, N, device=DEV, dtype=bf)
            fn = lambda: ops.gemm(A, W, out, bias=bias, act=L.ACT_GELU, save_z='grad')
            byts = M * K * 2 + 2 * M * N * 2
        elif res:
            x = torch.randn(M, N, device=DEV)
            out = torch.empty(M, N, device=DEV)
            sh = torch.empty(M, N, device=DEV, dtype=bf)
            gam = torch.ones(N, device=DEV)
            fn = lambda: ops.gemm(A, W, out, bias=bias, colscale=gam, residual=x, shadow=sh)
            byts = M * K * 2 + M * N * (4 + 4 + 2)
        else:
            out = torch.empty(M, N, device=DEV, dtype=bf)
            fn = lambda: ops.gemm(A, W, out, bias=bias)
            byts = M * K * 2 + M * N * 2
        cases[name] = (fn, byts, 2.0 * M * K * N)

    def dgrad_gelu(name, M, K, N):
        dy = torch.randn(M, K, device=DEV, dtype=bf)
        W = torch.randn(K, N, device=DEV, dtype=bf) * 0.05    # B(n,k) = W[k,n]  (MN-major)
        z = torch.randn(M, N, device=DEV, dtype=bf)
        out = torch.empty(M, N, device=DEV, dtype=bf)
        cases[name] = (lambda: ops.gemm(dy, W.t(), out, zin=z, zmode=L.ACT_MUL), M * K * 2 + 2 * M * N * 2, 2.0 * M * K * N)

    def wgrad(name, M, N, K):
        dY = torch.randn(M, N, device=DEV, dtype=bf)
        X = torch.randn(M, K, device=DEV, dtype=bf)
        G = torch.zeros(N, K, device=DEV)
        cases[name] = (lambda: ops.gemm(dY.t(), X.t(), G, accumulate=True), M * (N + K) * 2, 2.0 * M * K * N)

    for s, (hw, c) in enumerate([(56, 96), (28, 192), (14, 384), (7, 688)]):
        M = B * hw * hw
        fwd(f's{s}_fc1_gelu', M, c, 4 * c, gelu=True)
        fwd(f's{s}_fc2_res', M, 4 * c, c, res=True)
        dgrad_gelu(f's{s}_dz', M, c, 4 * c)
        fwd(f's{s}_dxhat', M, 4 * c, c)
        wgrad(f's{s}_wgrad_fc2', M, c, 4 * c)
    if os.environ.get('KB_SPLIT_SWEEP'):
        for s_, (hw, c) in enumerate([(56, 96), (28, 192), (14, 384), (7, 688)]):
            M = B * hw * hw
            for (n_, k_, tag) in ((c, 4 * c, 'fc2'), (4 * c, c, 'fc1')):
                dY = torch.randn(M, n_, device=DEV, dtype=bf)
                X = torch.randn(M, k_, device=DEV, dtype=bf)
                G = torch.zeros(n_, k_, device=DEV)
                for sp in (0, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48):
                    cases[f's{s_}_wgrad_{tag}_sp{sp}'] = (lambda dY=dY, X=X, G=G, sp=sp: ops.gemm(dY.t(), X.t(), G, accumulate=True, splits=sp),
                                                         M * (n_ + k_) * 2, 2.0 * M * k_ * n_)
    fwd('kv_proj', B * 196, 688, 1680)
    fwd('bneck_conv1', B * 196, 2128, 172)
    return cases


def dwconv_cases():
    bf = torch.float32 if os.environ.get('DW_DTYPE') == 'f32' else torch.bfloat16
    DT = L.F32 if bf == torch.float32 else L.BF16
    lib = L.load()
    cases = {}
    for s, (hw, c) in enumerate([(56, 96), (28, 192), (14, 384), (7, 688)]):
        M = B * hw * hw
        x = torch.randn(M, c, device=DEV, dtype=bf)
        w = torch.randn(49, c, device=DEV)
        b = torch.randn(c, device=DEV)
        y = torch.empty_like(x)
        rstd = torch.empty(M, device=DEV)
        dres = torch.randn(M, c, device=DEV)
        dx = torch.empty(M, c, device=DEV)
        d49 = torch.zeros(49, c, device=DEV)
        db = torch.zeros(c, device=DEV)
        ws = torch.empty(lib.ga_dwconv7_bwd_parts(B, hw, hw, c) * 50 * c, device=DEV)

        def f_fwd(x=x, w=w, b=b, y=y, rstd=rstd, hw=hw, c=c):
            L.check(lib.ga_dwconv7_ln_fwd(L.ptr(x), L.ptr(w), L.ptr(b), None, None, L.ptr(y), L.ptr(rstd), B, hw, hw, c, L.f(1e-6),
                                          DT, L.stream()), 'fwd')

        def f_dgrad(x=x, w=w, dres=dres, dx=dx, hw=hw, c=c):
            L.check(lib.ga_dwconv7_bwd(L.ptr(x), None, L.ptr(dres), L.ptr(w), L.ptr(dx), None, None, None, B, hw, hw, c, DT,
                                       L.F32, L.stream()), 'dgrad')

        def f_wgrad(x=x, y=y, w=w, d49=d49, db=db, ws=ws, hw=hw, c=c):
            L.check(lib.ga_dwconv7_bwd(L.ptr(y), L.ptr(x), None, L.ptr(w), None, L.ptr(d49), L.ptr(db), L.ptr(ws), B, hw, hw, c,
                                       DT, L.F32, L.stream()), 'wgrad')
        dxs = torch.empty(M, c, device=DEV, dtype=bf)

        def f_bwd2(x=x, y=y, w=w, dres=dres, dx=dx, dxs=dxs, d49=d49, db=db, ws=ws, hw=hw, c=c):
            L.check(lib.ga_dwconv7_bwd2(L.ptr(y), L.ptr(x), L.ptr(dres), L.ptr(w), L.ptr(dx), L.ptr(dxs), L.ptr(d49), L.ptr(db), L.ptr(ws),
                                        B, hw, hw, c, DT, L.F32, L.stream()), 'bwd2')
        fl = 2.0 * 49 * M * c
        cases[f's{s}_dw_bwd2'] = (f_bwd2, M * c * (2 + 2 + 4 + 4 + 2), 2 * fl)      # the training call: dgrad + wgrad + dbias
        cases[f's{s}_dw_fwd'] = (f_fwd, M * c * 4 + M * 4, fl)
        cases[f's{s}_dw_dgrad'] = (f_dgrad, M * c * (2 + 4 + 4), fl)
        cases[f's{s}_dw_wgrad'] = (f_wgrad, M * c * 4, fl)
    return cases


def rows_cases():
    bf = torch.bfloat16
    lib = L.load()
    cases = {}
    M, c = B * 56 * 56, 96
    a = torch.randn(M, c, device=DEV, dtype=bf)
    b2 = torch.randn(M, c, device=DEV, dtype=bf)
    rstd = torch.rand(M, device=DEV) + 0.5
    o = torch.empty_like(a)
    cases['ln_bwd_rows_s0'] = (lambda: L.check(lib.ga_ln_bwd_rows(L.ptr(a), L.ptr(b2), L.ptr(rstd), L.ptr(o), L.ll(M), c, L.BF16,
                                                                  L.stream()), 'ln'), M * c * 6 + M * 4, 0.0)
    cases['colsum_s0'] = (lambda: ops.colsum(a), M * c * 2, 0.0)
    big = torch.randn(M, 4 * c, device=DEV, dtype=bf)
    cases['colsum_s0_4c'] = (lambda: ops.colsum(big), M * c * 8, 0.0)
    f32 = torch.randn(M, c, device=DEV)
    cases['convert_s0'] = (lambda: ops.convert(f32, bf), M * c * 6, 0.0)
    return cases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('what', nargs='?', default='all')
    ap.add_argument('--one', default=None)
    args = ap.parse_args()
    hbm, tf = peaks()
    cases = {}
    if args.what in ('gemm', 'all'):
        cases.update(gemm_cases())
    if args.what in ('dwconv', 'all'):
        cases.update(dwconv_cases())
    if args.what in ('rows', 'all'):
        cases.update(rows_cases())
    if args.one:
        for name in args.one.split(','):          # one launch each (for an ncu capture); several names: comma-separated
            cases[name][0]()
        torch.cuda.synchronize()
        return
    print(f'{"case":18s} {"us":>9s} {"GB/s":>8s} {"%hbm":>6s} {"TFLOP/s":>8s} {"%tc":>6s}')
    for name, (fn, byts, flops) in cases.items():
        us = timeit(fn)
        gbs = byts / us / 1e3
        tfs = flops / us / 1e6
        print(f'{name:18s} {us:9.1f} {gbs:8.0f} {gbs / hbm * 100:6.1f} {tfs:8.1f} {tfs / tf * 100:6.1f}')


if __name__ == '__main__':
    main()
