#!/usr/bin/env python3
"""Minimal probe (2 ranks): NCCL all-reduce captured in a CUDA graph on a forked side stream, replayed, then an eager collective.
  torchrun --nproc-per-node 2 scripts/nccl_graph_probe.py <variant>"""
import os
import sys

import torch
import torch.distributed as dist

v = sys.argv[1]
rank = int(os.environ['RANK'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
if 'noasync' not in v:
    os.environ['TORCH_NCCL_ASYNC_ERROR_HANDLING'] = '0'
dist.init_process_group('nccl')
x = torch.ones(1 << 20, device='cuda') * (rank + 1)
a = torch.ones(1 << 20, device='cuda')
side = torch.cuda.Stream()
for _ in range(3):                      # eager warm-up of the communicator
    dist.all_reduce(x)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
kw = {'capture_error_mode': 'thread_local'} if 'tl' in v else ({'capture_error_mode': 'relaxed'} if 'relaxed' in v else {})
with torch.cuda.graph(g, **kw):
    a.mul_(2.0)
    if 'side' in v:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            if 'async' in v.split('_'):
                w = dist.all_reduce(x, async_op=True)
                w.wait()
            else:
                dist.all_reduce(x)
        a.add_(1.0)
        torch.cuda.current_stream().wait_stream(side)
    else:
        dist.all_reduce(x)
    a.add_(x)
print(f'[{rank}] {v}: captured', flush=True)
for i in range(3):
    x.fill_(rank + 1.0)
    g.replay()
torch.cuda.synchronize()
print(f'[{rank}] {v}: replayed, x[0]={x[0].item()} a[0]={a[0].item()}', flush=True)
if 'delwork' in v:
    del w
dist.all_reduce(x)                      # eager collective after the replays
torch.cuda.synchronize()
print(f'[{rank}] {v}: eager after replay ok x[0]={x[0].item()}', flush=True)
dist.barrier()
print(f'[{rank}] {v}: barrier ok', flush=True)
dist.destroy_process_group()
print(f'[{rank}] {v}: done', flush=True)
