#!/usr/bin/env python3
"""Per-kernel device time of one training step (CUPTI through torch.profiler; cheaper than an ncu launch list).

  python scripts/profile_step.py [--batch 256] [--steps 2] [--out gpurun_out/step_profile.txt]
Numbers are for ranking kernels and spotting regressions; bench values never come from a profiled run.
"""
import argparse
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'step_profile.txt'))
    ap.add_argument('--fwd-only', action='store_true')
    ap.add_argument('--model', default='ga_convnext_tiny_688')
    ap.add_argument('--drop-path', type=float, default=0.2)
    ap.add_argument('--top', type=int, default=45)
    ap.add_argument('--sequence', default=None, help='also write the ordered kernel list of the last profiled step here')
    args = ap.parse_args()
    from imagenet_models_b200 import ops
    from imagenet_models_b200.optim import FusedAdamWEma
    from imagenet_models_b200.registry import create_model
    import imagenet_models_b200.ga_convnext  # noqa: F401
    import imagenet_models_b200.ga_cswin  # noqa: F401
    import imagenet_models_b200.map_convnext  # noqa: F401
    dev = torch.device('cuda')
    torch.manual_seed(0)
    model = create_model(args.model, drop_path_rate=args.drop_path).to(dev).train()
    opt = FusedAdamWEma(model, lr=1e-3, weight_decay=0.05, ema_decay=0.9998)
    x = torch.randn(args.batch, 3, 224, 224, device=dev)
    y = torch.randint(0, 1000, (args.batch,), device=dev)

    def step():
        if args.fwd_only:
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                model(x)
            return
        opt.zero_grad()
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = model(x)
        if isinstance(out[0], (list, tuple)):
            ops.ga_loss(torch.stack([o[0] for o in out]), y, -0.8, aux=torch.stack([o[1] for o in out])).backward()
        else:
            ops.ga_loss(torch.stack(out), y, -0.8).backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    shapes = collections.defaultdict(list)
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r'^void ', '', ev.name)
            name = re.sub(r'\(.*', '', name)
            agg[name][0] += 1
            agg[name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
            shapes[name].append(ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time)
    total = sum(v[1] for v in agg.values())
    lines = [f'batch {args.batch}, {args.steps} steps, total device time {total / 1e3 / args.steps:.3f} ms/step']
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
        top = sorted(shapes[k], reverse=True)[:3]
        lines.append(f'{t / total * 100:6.2f}%  {t / 1e3 / args.steps:8.3f} ms/step  n={c // args.steps:4d}  max {top[0]:8.1f} us  {k[:100]}')
    nk = sum(c for c, _ in agg.values()) // args.steps
    ours = sum(c for k, (c, _) in agg.items() if not (k.startswith('at::') or 'Memcpy' in k or 'Memset' in k or k.startswith('cub::'))) // args.steps
    lines.insert(1, f'{nk} kernel launches per step: {ours} from libga_sm100.so, {nk - ours} ATen / memcpy / memset')
    if args.sequence:
        evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        evs = evs[len(evs) - len(evs) // args.steps:]
        with open(args.sequence, 'w') as f:
            for ev in evs:
                f.write(f'{(ev.device_time if hasattr(ev, "device_time") else ev.cuda_time):9.1f}  {re.sub("^void ", "", ev.name).split(chr(40))[0][:110]}\n')
    txt = '\n'.join(lines)
    print(txt)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, 'w').write(txt + '\n')


if __name__ == '__main__':
    main()
