#!/usr/bin/env python3
"""Drop-in entry point for the reference's `MAP/validate.py` loop (validate(): MAP/validate.py:131-326) on the sm_100a path.

Replicas only: with N GPUs launch N processes (torchrun); each evaluates a disjoint slice and one final all-reduce
combines (loss, correct@1, correct@5, count) -- this replaces the reference's single-process nn.DataParallel (:191-192).
Branch logits are averaged for MAP models and summed for GA models, as the reference does.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imagenet_models_b200.ga_convnext  # noqa: F401,E402
import imagenet_models_b200.map_convnext  # noqa: F401,E402
import imagenet_models_b200.ga_cswin  # noqa: F401,E402
from imagenet_models_b200.engine import EvalEngine  # noqa: E402
from imagenet_models_b200.registry import create_model  # noqa: E402

parser = argparse.ArgumentParser(description='validation on B200 (flags follow MAP/validate.py:49-125)')
parser.add_argument('data', nargs='?', default='')
parser.add_argument('--model', '-m', default='ga_convnext_tiny_688')
parser.add_argument('-b', '--batch-size', default=256, type=int)
parser.add_argument('--img-size', default=224, type=int)
parser.add_argument('--num-batches', default=20, type=int, help='synthetic batches per process')
parser.add_argument('--checkpoint', default='', type=str)
parser.add_argument('--amp', action='store_true', default=False)
parser.add_argument('--channels-last', action='store_true', default=True)
parser.add_argument('--results-file', default='', type=str)
parser.add_argument('--no-graph', action='store_true', help='run every batch eagerly instead of replaying a CUDA graph')


def main():
    args = parser.parse_args()
    distributed = int(os.environ.get('WORLD_SIZE', '1')) > 1
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    rank, world = 0, 1
    if distributed:
        dist.init_process_group(backend='nccl', init_method='env://')
        rank, world = dist.get_rank(), dist.get_world_size()
    model = create_model(args.model, checkpoint_path=args.checkpoint).cuda().eval()
    reduce = 'mean' if args.model.startswith('map_') else 'sum'
    amp = torch.bfloat16 if args.amp else None
    g = torch.Generator(device='cuda').manual_seed(1234 + rank)
    B, S = args.batch_size, args.img_size
    x = torch.randn(B, 3, S, S, device='cuda', generator=g)
    if args.channels_last:
        x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, model.num_classes, (B,), device='cuda', generator=g)
    engine = EvalEngine(model, reduce, amp, cuda_graph=not args.no_graph)
    for _ in range(4):
        engine(x, y)                                         # warm-up forwards (MAP/validate.py:240-244) + graph capture
    if distributed:
        dist.all_reduce(torch.zeros(1, device='cuda'))       # NCCL communicator set-up stays out of the timed loop
    torch.cuda.synchronize()
    acc = torch.zeros(4, device='cuda')
    t0 = time.time()
    for _ in range(args.num_batches):
        acc += torch.stack([t.float() for t in engine(x, y)])   # stays on the device
    if distributed:
        dist.all_reduce(acc)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        n = acc[3].item()
        res = {'model': args.model, 'top1': round(100 * acc[1].item() / n, 4), 'top5': round(100 * acc[2].item() / n, 4),
               'loss': round(acc[0].item() / (args.num_batches * world), 4), 'img_size': S,
               'param_count': round(sum(p.numel() for p in model.parameters()) / 1e6, 2), 'images_per_sec': round(n / dt, 1)}
        print(f'--result\n{json.dumps(res, indent=4)}')
        if args.results_file:
            json.dump(res, open(args.results_file, 'w'))
    if distributed:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
