#!/usr/bin/env python3
"""K6 stripe-attention micro-benchmark: device time per launch (CUDA events) and fraction of the HBM roofline.

  python scripts/attn_bench.py [--batch 128] [--dtype bf16]
Algorithmic bytes per token: forward 4C (qkv in, out) ; backward 8C (dout, qkv, out in; dqkv out) elements.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

CASES = [('stage1', 56, 1, 2, 64), ('stage2', 28, 2, 2, 128), ('stage3', 14, 7, 2, 256), ('stage4', 7, 7, 1, 512),
         ('stage5', 14, 7, 2, 512), ('gram', 14, 7, 2, 192)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    from imagenet_models_b200 import ops
    dt = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    peak = 6500.0
    try:
        peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', peak)
    except Exception:
        pass
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for name, R, split, nbr, C in CASES:
        B = args.batch
        M = B * R * R
        qkv = torch.randn(M, 3 * C, device='cuda').to(dt)
        lw = torch.randn(C, 9, device='cuda') * 0.2
        lb = torch.randn(C, device='cuda') * 0.1
        dout = torch.randn(M, C, device='cuda').to(dt)
        dl = torch.zeros(C * 10, device='cuda')
        out, lse = ops._attn_fwd(qkv, lw, lb, B, R, C, split, nbr, True)
        res = {}
        for kind in ('fwd', 'bwd'):
            ts = []
            for _ in range(args.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if kind == 'fwd':
                    ops._attn_fwd(qkv, lw, lb, B, R, C, split, nbr, True)
                else:
                    ops._attn_bwd(dout, qkv, out, lse, lw, lb, dl[:C * 9], dl[C * 9:], B, R, C, split, nbr)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            ts.sort()
            res[kind] = ts[len(ts) // 2]
        es = qkv.element_size()
        fb, bb = M * 4 * C * es, M * 8 * C * es
        print(f'{name:7s} R={R:2d} split={split} C={C:3d} B={B}: fwd {res["fwd"]:8.1f} us ({fb / res["fwd"] / 1e3:7.1f} GB/s, {fb / res["fwd"] / 1e3 / peak:5.1%})'
              f'   bwd {res["bwd"]:8.1f} us ({bb / res["bwd"] / 1e3:7.1f} GB/s, {bb / res["bwd"] / 1e3 / peak:5.1%})')


if __name__ == '__main__':
    main()
