#!/usr/bin/env python3
"""Headline benchmark: GA-ConvNeXt-T (ga_convnext_tiny_688) training step, bf16, batch 256 per GPU, 224x224 synthetic.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                      (the reference's CPU path = oracle port, on the host cores)

One step = forward (5 branch logits) + GA loss (CE + lam*KL) + backward + gradient all-reduce (N>1) + fused AdamW + EMA.
`value` times steps with the batch already resident in HBM; `e2e` times the same steps fed from pinned host memory
(uint8 images -> H2D -> normalise, as timm's PrefetchLoader does) with the loss read back every step.
`roofline` is the dominant C-ABI call site of the step -- EVERY libga entry point is timed by CUDA events on the launching
stream (lib.CallTimer), keyed by entry point and shape; the one with the largest total time is reported with its algorithmic
bytes / measured duration, `top_calls` lists the next ones.  `infer` and `sustained` are extra legs.
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of the same call sites (profiles/)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures of the same call sites
# (profiles/r02_ncu_gemm_epilogue.txt, profiles/r02_ncu_dwconv3.txt)
NCU_TRAFFIC = {('ga_gemm', (802816, 384, 96, 'gelu')): 1334.0e6, ('ga_gemm', (802816, 384, 96, 'mul')): 1353.0e6,
               ('ga_dwconv7_bwd2', (256, 56, 56, 96)): 1194.6e6, ('ga_dwconv7_bwd3', (256, 56, 56, 96)): 1194.6e6,
               ('ga_dwconv7_bwd2', (256, 14, 14, 384)): 264.0e6, ('ga_dwconv7_bwd3', (256, 14, 14, 384)): 264.0e6}
KERNEL_OF = {'ga_gemm': 'tc::gemm_tc2_kernel (tcgen05 persistent GEMM)', 'ga_dwconv7_ln_fwd': 'dw::dwconv7 forward (dw7x7 + bias + LayerNorm)',
             'ga_dwconv7_bwd2': 'dw3::dwconv7_bwd3_kernel (fused depthwise 7x7 backward: data gradient + residual + bf16 shadow, weight and bias gradient)',
             'ga_dwconv7_bwd3': 'dw3::dwconv7_bwd3_kernel (fused depthwise 7x7 backward: data gradient + residual + bf16 shadow, weight and bias gradient)'}


def call_bytes(name, sig):
    """Algorithmic HBM bytes of one C-ABI call (SURVEY 8d convention: every operand read once, every result written once)."""
    if name == 'ga_gemm':
        nb, M, N, K, idt, odt, act, has_z, z_shadow, has_r, has_zin, acc, a_mn, b_mn = sig
        ei, eo = (2 if idt == 1 else 4), (2 if odt == 1 else 4)
        b = nb * (M * K + N * K) * ei + nb * M * N * eo
        if has_z:
            b += nb * M * N * (2 if z_shadow == 1 else eo)
        b += nb * M * N * eo * (has_r + has_zin)
        return b
    if name == 'ga_dwconv7_ln_fwd':
        B, H, W, C, dt = sig[:5]
        es = 2 if dt == 1 else 4
        return B * H * W * (2 * C * es + 4)
    if name in ('ga_dwconv7_bwd2', 'ga_dwconv7_bwd3'):
        B, H, W, C, dt, rdt = sig[:6]
        es, rs = (2 if dt == 1 else 4), (2 if rdt == 1 else 4)
        return B * H * W * C * (2 * es + 2 * rs + (es if rs != es else 0))     # R dconv, R x, R dres, W dx (+ W bf16 shadow)
    if name in ('ga_ln_bwd_rows', 'ga_ln_bwd_rows_res'):
        M, C, dt = sig[:3]
        return M * C * 3 * (2 if dt == 1 else 4)
    if name == 'ga_colstats':
        M, C = sig[:2]
        return M * C * 2
    if name in ('ga_adamw_ema_dev', 'ga_adamw_ema'):
        return sig[1] * 9 * 4
    return None


def call_label(name, sig):
    if name == 'ga_gemm':
        nb, M, N, K, idt, odt, act, has_z, z_shadow, has_r, has_zin, acc, a_mn, b_mn = sig
        epi = ('gelu' if act == 1 else 'relu' if act == 2 else 'lin') + ('+z' if has_z and z_shadow != 1 else '') + ('+res' if has_r else '') + \
            ('+shadow' if z_shadow == 1 else '') + ('*zin' if has_zin else '') + ('+acc' if acc else '')
        return f'ga_gemm M={M} N={N} K={K}' + (f' batch={nb}' if nb > 1 else '') + f' {"bf16" if idt == 1 else "f32"} epilogue {epi}'
    return f'{name}{tuple(sig)}'


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


REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def cpu_reference_step_rate(batch, steps, warmup, threads=None):
    """The reference's CPU path (fp32, all host threads), fwd + GA loss + bwd of config 1.  -> (img/s, s/step, threads, kind).
    kind 'reference': the UNMODIFIED reference module staged under baseline/_ref (git-ignored copy made by build(); it travels to
    the GPU box with the snapshot), imported through oracle/timm_shim because timm is not installed; kind 'port': the oracle
    restatement, used only when the staged copy is absent."""
    import warnings
    import torch
    import torch.nn.functional as F
    warnings.filterwarnings('ignore', message='.*reduction.*')     # the reference's own kl_div(reduction='mean') warning
    torch.set_num_threads(threads or os.cpu_count())
    g = torch.Generator().manual_seed(42)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    y = torch.randint(0, 1000, (batch,), generator=g)
    if os.path.exists(os.path.join(REF_DIR, 'GA', 'ga_convnext.py')):
        kind = 'reference'
        sys.path.insert(0, os.path.join(ROOT, 'oracle', 'timm_shim'))
        sys.path.insert(0, os.path.join(REF_DIR, 'GA'))
        import timm
        import ga_convnext  # noqa: F401  (the reference file, registers into the shim)
        torch.manual_seed(0)
        model = timm.create_model(MODEL).train()

        def step():
            for p in model.parameters():
                p.grad = None
            outs = model(x)
            output, loss = 0, 0                      # the loss expression of GA/train.py:735-745
            for o in outs:
                loss = loss + F.cross_entropy(o, y)
                output = output + o.data
            for o in outs:
                loss = loss + F.kl_div(F.log_softmax(o, -1), F.log_softmax(output.detach() / len(outs), -1), reduction='mean',
                                       log_target=True) * GA_LAM
            loss.backward()
    else:
        kind = 'port'
        from oracle import ga_convnext_oracle as O
        spec = O.SPECS[MODEL]
        P = O.make_state(spec, 7)
        leaves = {k: (v.requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in P.items()}

        def step():
            for v in leaves.values():
                v.grad = None
            O.ga_loss(O.forward(leaves, spec, x, training=True), y, GA_LAM).backward()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    t = statistics.median(times[warmup:])
    return batch / t, t, torch.get_num_threads(), kind


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch = 8
    rate, t, threads, kind = cpu_reference_step_rate(batch, args.steps, args.warmup)
    what = ('the unmodified reference module (baseline/_ref, through oracle/timm_shim)' if kind == 'reference'
            else 'oracle port of the reference modules (baseline/_ref not staged on this box)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'img/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{MODEL} training step (fwd + GA loss + bwd), fp32, 224x224, host CPU (BASELINE config 1)',
                   'sample': f'batch {batch} per step: a bounded sample of the batch-256 step'},
        'cpu_baseline': {'value': rate, 'unit': 'img/s', 'cores': threads, 'kind': kind,
                         'sample': f'{args.steps} steps of batch {batch}; {what}'},
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
    ap.add_argument('--no-infer', action='store_true', help='skip the inference leg (T-688 B=256, B-976 B=1024 eval throughput)')
    ap.add_argument('--sustained', type=float, default=5.0, help='seconds of back-to-back steps for the sustained figure (0 = skip)')
    ap.add_argument('--drop-path', type=float, default=0.2, help='drop_path_rate of the benchmarked model (recipe value)')
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
        os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')     # the step's NCCL calls are captured in a CUDA graph
        dist.init_process_group('nccl', init_method='env://')
    dev = torch.device('cuda', local)
    L.load()

    torch.manual_seed(42 + rank)                       # random_seed(seed, rank), GA/train.py:402
    model = create_model(args.model, drop_path_rate=args.drop_path).to(dev).train()    # recipe value 0.2 (throughput config)
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
    sustained = None
    if args.sustained > 0:
        n_s = max(args.steps, int(args.sustained * 1e3 / (ms / args.steps)) + 1)
        ms_s = timed(step_resident, n_s)
        sustained = {'value': world * B * n_s / (ms_s / 1e3), 'unit': 'img/s', 'steps': n_s, 'seconds': ms_s / 1e3,
                     'ms_per_step': ms_s / n_s}

    # per-call roofline: an instrumented EAGER pass after the timed region (CUDA events around every C-ABI call on the
    # launching stream; a graph replay has no per-launch host hook).  All ranks run it so collectives stay matched.
    graph_used = engine._graph is not None
    engine.cuda_graph = False
    n_inst = min(args.steps, 3)
    step_resident()
    timer = L.start_timing() if rank == 0 else None
    ms_inst = timed(step_resident, n_inst)
    calls = {}
    for (nm, sg), (n, t) in (timer.summary() if rank == 0 else {}).items():
        nm = 'ga_dwconv7_bwd2' if nm == 'ga_dwconv7_bwd3' else nm          # same kernel, with / without the DropPath-scaled shadow
        n0_, t0_ = calls.get((nm, sg), (0, 0.0))
        calls[(nm, sg)] = (n0_ + n, t0_ + t)
    L.stop_timing()

    # inference leg (BASELINE config 5 / validate.py loop body), CUDA-graph replay per batch, replicas only
    infer = None
    if not args.no_infer:
        from imagenet_models_b200.engine import EvalEngine
        del engine, model
        torch.cuda.empty_cache()
        infer = {}
        for name, ib in (('ga_convnext_tiny_688', 256), ('ga_convnext_base_976', 1024)):
            m = create_model(name).to(dev).eval()
            ev = EvalEngine(m, 'sum', torch.bfloat16)
            xi = torch.randn(ib, 3, 224, 224, device=dev).contiguous(memory_format=torch.channels_last)
            yi = torch.randint(0, 1000, (ib,), device=dev)
            for _ in range(4):
                ev(xi, yi)
            n_i = 10
            ms_i = timed(lambda: ev(xi, yi), n_i)
            infer[f'{name} B={ib}'] = {'value': world * ib * n_i / (ms_i / 1e3), 'unit': 'img/s', 'ms_per_batch': ms_i / n_i}
            del m, ev, xi
            torch.cuda.empty_cache()
    if rank != 0:
        _finish(world)
        return
    ms_step = ms / args.steps
    img_s = world * B * args.steps / (ms / 1e3)
    hbm, tf, src = peaks()
    per_gpu = img_s / world
    ms_step_inst = ms_inst / n_inst
    ranked = sorted(calls.items(), key=lambda kv: -kv[1][1])
    total_call_ms = sum(v[1] for v in calls.values()) / n_inst
    (name, sig), (n_launch, t_ms) = ranked[0]
    byts = call_bytes(name, sig)
    avg_us = t_ms / n_launch * 1e3
    achieved = byts / (avg_us * 1e-6) / 1e9 if byts else None
    top = []
    for (nm, sg), (n, t) in ranked[:8]:
        b_ = call_bytes(nm, sg)
        top.append({'call': call_label(nm, sg), 'calls_per_step': n / n_inst, 'share_of_step': (t / n_inst) / ms_step_inst,
                    'avg_us': t / n * 1e3, 'hbm_frac': (b_ / (t / n * 1e-3) / 1e9 / hbm) if b_ else None})
    fam = {}
    for (nm, sg), (n, t) in calls.items():
        k = 'gemm' if nm == 'ga_gemm' else 'dwconv7' if nm.startswith('ga_dwconv7') else 'other'
        fam[k] = fam.get(k, 0.0) + t / n_inst / ms_step_inst
    tkey = (name, tuple(sig[:4])) if name.startswith('ga_dwconv7') else None
    if name == 'ga_gemm':
        tkey = (name, (sig[1], sig[2], sig[3], 'gelu' if sig[6] == 1 else 'mul' if sig[10] else 'lin'))
    line = {
        'metric': METRIC, 'value': img_s, 'unit': 'img/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': warm, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16', 'data': 'synthetic',
        'config': {'workload': f'{args.model} training step (fwd + GA loss + bwd + all-reduce + fused AdamW + EMA), bf16 autocast '
                               f'(fp32 residual stream), batch {B}/GPU, 224x224, drop_path {args.drop_path}', 'global_batch': B * world,
                   'parallelism': f'dp{world}', 'cuda_graph': graph_used, 'l2': 'activations per step (>10 GB) exceed the 126 MB L2; no explicit flush'},
        'per_gpu': img_s / world, 'gpu_launches': launches, 'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s', 'frac': achieved / hbm if achieved else None,
                     'traffic': NCU_TRAFFIC.get(tkey), 'peak_source': src,
                     'kernel': f'{KERNEL_OF.get(name, name)}; call site {call_label(name, sig)}',
                     'algorithmic_bytes_per_launch': byts, 'avg_launch_us': avg_us, 'launches_timed': n_launch,
                     'share_of_step': (t_ms / n_inst) / ms_step_inst, 'family_share_of_step': fam,
                     'libga_share_of_step': total_call_ms / ms_step_inst, 'top_calls': top,
                     'measured': f'{n_inst} instrumented eager steps after the timed region ({ms_step_inst:.2f} ms/step with events around '
                                 f'every C-ABI call); the dominant call = largest total time over all libga entry points'},
        'step_roofline': {'hbm_frac': per_gpu * mb_img / 1e3 / hbm if mb_img else None, 'tensor_frac': per_gpu * gf_img / 1e3 / tf if gf_img else None,
                          'note': f'{mb_img} MB/img (SURVEY 8d convention) and {gf_img} training GFLOP/img x img/s/GPU over the measured peaks'},
    }
    if e2e:
        line['e2e'] = e2e
    if sustained:
        line['sustained'] = sustained
    if infer:
        line['infer'] = infer
    if not args.no_cpu_baseline and world == 1:
        rate, t, threads, kind = cpu_reference_step_rate(8, 3, 1)
        line['cpu_baseline'] = {'value': rate, 'unit': 'img/s', 'cores': threads, 'kind': kind,
                                'sample': '3 training steps (fwd + GA loss + bwd) of batch 8 (BASELINE config 1), fp32, after 1 warm-up, on all '
                                          'host threads; ' + ('the unmodified reference module staged in baseline/_ref' if kind == 'reference'
                                                              else 'oracle port (baseline/_ref not staged)')}
    print(json.dumps(line), flush=True)
    _finish(world)


def _finish(world):
    """Several ranks: leave without ncclCommDestroy.  destroy_process_group() blocks for ever once a CUDA graph holding NCCL
    kernels of that communicator exists (probed on 2 x B200, scripts/nccl_graph_probe.py: capture, replays, eager collectives and
    barrier all complete, destroy_process_group() never returns), so the ranks synchronise, flush and exit."""
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
