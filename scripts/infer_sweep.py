#!/usr/bin/env python3
"""Inference throughput sweep: images/s of the eval forward + top-k/loss over batch sizes,
bf16 autocast, CUDA-graph replay (EvalEngine), device-resident synthetic input.  One JSON line per (model, size, batch).

  python scripts/infer_sweep.py --model ga_convnext_base_976 --sizes 224 384 --batches 1 2 4 8 16 32 64 128 256
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='ga_convnext_base_976')
    ap.add_argument('--sizes', type=int, nargs='+', default=[224])
    ap.add_argument('--batches', type=int, nargs='+', default=[1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024])
    ap.add_argument('--iters', type=int, default=30)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--fwd-gflop', type=float, default=None, help='forward GFLOP per image (for the tensor-pipe fraction)')
    args = ap.parse_args()
    import imagenet_models_b200.ga_convnext  # noqa: F401
    import imagenet_models_b200.ga_cswin  # noqa: F401
    import imagenet_models_b200.map_convnext  # noqa: F401
    from imagenet_models_b200.engine import EvalEngine
    from imagenet_models_b200.registry import create_model
    torch.manual_seed(0)
    model = create_model(args.model).cuda().eval()
    reduce = 'mean' if args.model.startswith('map_') else 'sum'
    eng = EvalEngine(model, reduce, torch.bfloat16, cuda_graph=not args.no_graph)
    for S in args.sizes:
        for B in args.batches:
            x = torch.randn(B, 3, S, S, device='cuda').contiguous(memory_format=torch.channels_last)
            y = torch.randint(0, 1000, (B,), device='cuda')
            try:
                for _ in range(4):
                    eng(x, y)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    eng(x, y)
                e1.record()
                torch.cuda.synchronize()
            except Exception as exc:  # noqa: BLE001
                print(json.dumps({'model': args.model, 'size': S, 'batch': B, 'error': str(exc)[:200]}))
                continue
            ms = e0.elapsed_time(e1) / args.iters
            rec = {'model': args.model, 'size': S, 'batch': B, 'ms_per_batch': round(ms, 4), 'img_per_s': round(B / ms * 1e3, 1),
                   'cuda_graph': not args.no_graph}
            if args.fwd_gflop:
                rec['tflops'] = round(B / ms * args.fwd_gflop, 1)
            print(json.dumps(rec), flush=True)


if __name__ == '__main__':
    main()
