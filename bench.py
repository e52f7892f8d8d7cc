#!/usr/bin/env python3
"""Headline benchmark: GA-ConvNeXt-T (ga_convnext_tiny_688) training step, bf16, batch 256 per GPU, 224x224 synthetic.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      (the reference's CPU path = oracle port, on the host cores)

One step = forward (5 branch logits) + GA loss (CE + lam*KL) + backward + gradient all-reduce (N>1) + fused AdamW + EMA.
`value` times steps with the batch already resident in HBM; `e2e` times the same steps fed from pinned host memory
(uint8 images -> H2D -> normalise, as timm's PrefetchLoader does) with the loss read back every step.
`roofline` is the dominant kernel call site (largest share of the timed region among the tcgen05 GEMM launches):
algorithmic bytes / CUDA-event time measured on the launching stream inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = 'ga_convnext_tiny_688'
TRAIN_GFLOP_PER_IMG = 32.73      # BASELINE.md section 4 (3 x 10.909 forward)
TRAIN_MB_PER_IMG = 268.0         # GEMM-boundary-fusion convention, bf16
GA_LAM = -0.8
# dram__bytes_read+write per launch from profiles/r01_ncu_fc1_gelu.txt (ncu --set full) for the stage-0 fc1+GELU GEMM
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures (profiles/r01_ncu_gemm_sites.txt)
# BASELINE.json's metric is "train/infer images/sec/GPU at 1/2/4/8 B200 (224^2) + % roofline"; the bench contract wants the whole-job
# aggregate in `value`, so the line carries the training aggregate, `per_gpu` = value / n_gpus, and the roofline objects
METRIC = 'train images/sec, whole job (BASELINE.json metric: train/infer images/sec/GPU at 1/2/4/8 B200 (224^2) + % roofline; per_gpu = value / n_gpus)'
NCU_TRAFFIC = {(802816, 384, 96, 'gelu+z'): 1333.5e6, (802816, 384, 96, 'lin+zin'): 1360.3e6}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d['hbm_gbs'], d.get('bf16_tflops_sustained', d['bf16_tflops']), 'measured'
    return 6650.0, 1400.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        busy = sorted(float(r[0]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in rows)]
        return {'sm_mhz': statistics.median(busy), 'sm_max_mhz': float(rows[0][1]), 'power_w_max': max(float(r[2]) for r in rows),
                'samples': len(rows), 'reasons': reasons}


def cpu_reference_step_rate(batch, steps, warmup, threads=None):
    """The reference's CPU path (fp32, all host threads): oracle port of GA_ConvNeXt fwd + GA loss + bwd.  img/s."""
    import torch
    from oracle import ga_convnext_oracle as O
    torch.set_num_threads(threads or os.cpu_count())
    spec = O.SPECS[MODEL]
    P = O.make_state(spec, 7)
    leaves = {k: (v.requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in P.items()}
    g = torch.Generator().manual_seed(42)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    y = torch.randint(0, 1000, (batch,), generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for v in leaves.values():
            v.grad = None
        out = O.forward(leaves, spec, x, training=True)
        O.ga_loss(out, y, GA_LAM).backward()
        times.append(time.perf_counter() - t0)
    t = statistics.median(times[warmup:])
    return batch / t, t, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = 8
    rate, t, threads = cpu_reference_step_rate(batch, args.steps, args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'img/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{MODEL} training step (fwd + GA loss + bwd), fp32, 224x224, host CPU',
                   'sample': f'batch {batch} per step: a bounded sample of the batch-256 step'},
        'cpu_baseline': {'value': rate, 'unit': 'img/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{args.steps} steps of batch {batch}; oracle port of the reference modules (timm is absent, so the '
                                   f'reference itself cannot be imported on this box)'},
        'e2e': {'value': rate, 'unit': 'img/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='per-GPU batch')
    ap.add_argument('--model', default=MODEL, help='headline: ga_convnext_tiny_688; also map_convnext_tiny (config 4, use --batch 512)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='run every step eagerly instead of replaying the captured CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from imagenet_models_b200 import lib as L
    from imagenet_models_b200 import ops
    from imagenet_models_b200.engine import TrainEngine
    from imagenet_models_b200.registry import create_model
    import imagenet_models_b200.ga_convnext  # noqa: F401
    import imagenet_models_b200.map_convnext  # noqa: F401
    import imagenet_models_b200.ga_cswin  # noqa: F401

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', init_method='env://')
    dev = torch.device('cuda', local)
    L.load()

    torch.manual_seed(42 + rank)                       # random_seed(seed, rank), GA/train.py:402
    model = create_model(args.model).to(dev).train()
    gf_img, mb_img = (TRAIN_GFLOP_PER_IMG, TRAIN_MB_PER_IMG) if args.model == MODEL else ({'map_convnext_tiny': 30.4, 'ga_CSWin_64_12211_tiny_224': 38.3}.get(args.model), None)
    if world > 1:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    engine = TrainEngine(model, lr=1e-3, weight_decay=0.05, ema_decay=0.9998, ga_lam=GA_LAM, amp_dtype=torch.bfloat16,
                         cuda_graph=not args.no_graph, graph_warmup=3)
    B = args.batch
    x_dev = torch.randn(B, 3, 224, 224, device=dev)
    y_dev = torch.randint(0, 1000, (B,), device=dev)
    x_host = torch.randint(0, 256, (B, 3, 224, 224), dtype=torch.uint8).pin_memory()
    y_host = torch.randint(0, 1000, (B,)).pin_memory()
    loss_host = torch.zeros(1).pin_memory()

    def step_resident():
        engine.step(x_dev, y_dev)

    from imagenet_models_b200.engine import DevicePrefetcher
    prefetch = DevicePrefetcher(device=dev)

    def step_e2e():
        # timm PrefetchLoader semantics: each step takes the batch whose uint8 H2D copy + normalisation was issued on the side
        # stream one step earlier, and issues the next one (one H2D of the full batch per step, inside the timed region);
        # the loss is read back every step
        x, y = prefetch.get()
        prefetch.submit(x_host, y_host)
        loss = engine.step(x, y)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return loss_host.item()

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    warm = max(args.warmup, 3)
    for _ in range(warm + (2 if engine.cuda_graph else 0)):     # graph mode: 3 eager steps, capture, one replay before timing
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.launch_count()
    ms = timed(step_resident, args.steps)
    launches = L.launch_count() - n0 + (engine.graph_launches * args.steps if engine._graph is not None else 0)
    clocks = sampler.stop() if rank == 0 else None
    e2e = None
    if not args.no_e2e:
        prefetch.submit(x_host, y_host)
        for _ in range(2):
            step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        e2e = {'value': world * B * args.steps / (ms_e2e / 1e3), 'unit': 'img/s', 'ms_per_step': ms_e2e / args.steps,
               'h2d_bytes_per_step': x_host.numel() + y_host.numel() * 8, 'd2h_bytes_per_step': 4}

    # per-kernel roofline: an instrumented EAGER pass after the timed region (CUDA events around every ga_gemm launch on
    # the launching stream; a graph replay has no per-launch host hook).  All ranks run it so collectives stay matched.
    graph_used = engine._graph is not None
    engine.cuda_graph = False
    n_inst = min(args.steps, 5)
    step_resident()
    if rank == 0:
        ops.TIMER = ops.GemmTimer()
    ms_inst = timed(step_resident, n_inst)
    gemm_times = ops.TIMER.summary() if rank == 0 else {}
    ops.TIMER = None
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms / args.steps
    img_s = world * B * args.steps / (ms / 1e3)
    hbm, tf, src = peaks()
    per_gpu = img_s / world
    # dominant kernel call site: the GEMM shape with the largest total time inside the timed region
    key, (n_launch, t_ms) = max(gemm_times.items(), key=lambda kv: kv[1][1])
    nb, M, N, K, dt, kind, byts = key
    avg_us = t_ms / n_launch * 1e3
    achieved = byts / (avg_us * 1e-6) / 1e9
    gemm_total_ms = sum(v[1] for v in gemm_times.values()) / n_inst
    ms_step_inst = ms_inst / n_inst
    line = {
        'metric': METRIC, 'value': img_s, 'unit': 'img/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': warm, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': f'{args.model} training step (fwd + GA loss + bwd + all-reduce + fused AdamW + EMA), bf16 autocast '
                               f'(fp32 residual stream), batch {B}/GPU, 224x224', 'global_batch': B * world,
                   'parallelism': f'dp{world}', 'cuda_graph': graph_used, 'l2': 'activations per step (>10 GB) exceed the 126 MB L2; no explicit flush'},
        'per_gpu': img_s / world, 'gpu_launches': launches, 'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s', 'frac': achieved / hbm,
                     'traffic': NCU_TRAFFIC.get((M, N, K, kind)), 'peak_source': src,
                     'kernel': f'tc::gemm_tc2_kernel (tcgen05 persistent GEMM), call site M={M} N={N} K={K} {dt} epilogue {kind}',
                     'algorithmic_bytes_per_launch': byts, 'avg_launch_us': avg_us, 'launches_timed': n_launch,
                     'share_of_step': (t_ms / n_inst) / ms_step_inst, 'all_gemm_share_of_step': gemm_total_ms / ms_step_inst,
                     'measured': f'{n_inst} instrumented eager steps after the timed region ({ms_step_inst:.2f} ms/step with the events)'},
        'step_roofline': {'hbm_frac': per_gpu * mb_img / 1e3 / hbm if mb_img else None, 'tensor_frac': per_gpu * gf_img / 1e3 / tf if gf_img else None,
                          'note': f'{mb_img} MB/img (SURVEY 8d convention) and {gf_img} training GFLOP/img x img/s/GPU over the measured peaks'},
    }
    if e2e:
        line['e2e'] = e2e
    if not args.no_cpu_baseline and world == 1:
        rate, t, threads = cpu_reference_step_rate(8, 3, 1)
        line['cpu_baseline'] = {'value': rate, 'unit': 'img/s', 'cores': threads, 'kind': 'port',
                                'sample': '3 training steps (fwd + GA loss + bwd) of batch 8, fp32, after 1 warm-up; oracle port of '
                                          'the reference modules on all host threads'}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
