"""Shared body of the drop-in entry points GA/train.py, MAP/train.py and MAP/train_with_script.py.

The reference scripts (GA/train.py:69-309, MAP/train.py:60-380) drive timm's runtime: loaders, augmentation, schedulers, checkpoint
saver.  timm and ImageNet are not in this image, so the data here is synthetic (uint8 batches in pinned host memory, copied and
normalised -- and mixed, with --mixup / --cutmix -- by DevicePrefetcher exactly where timm's PrefetchLoader sits); everything on
the device side is the B200-native implementation: model, loss (CE / label smoothing / soft targets / BCE + the GA or MAP group
terms), backward, bucketed all-reduce, LAMB or AdamW (+ global-norm clip), EMA, BatchNorm buffer sync, cosine schedule.
Flags that only concern the absent data pipeline (--aa, --reprob, -j, --crop-pct, ...) are accepted and reported as ignored.
"""
import argparse
import logging
import os
import time

import torch
import torch.distributed as dist

from .engine import CosineSchedule, DevicePrefetcher, Mixup, TrainEngine, evaluate_batch
from .registry import create_model

_logger = logging.getLogger('train')


def build_parser(description, default_model, lam_flag, lam_default):
    p = argparse.ArgumentParser(description=description)
    p.add_argument('data_dir', nargs='?', default='', help='unused: no dataset on this box (synthetic batches)')
    p.add_argument('--model', default=default_model, type=str)
    p.add_argument('--num-classes', type=int, default=None)
    p.add_argument('-b', '--batch-size', type=int, default=128, help='per-process batch')
    p.add_argument('--img-size', type=int, default=224)
    p.add_argument('--epochs', type=int, default=1)
    p.add_argument('--steps-per-epoch', type=int, default=50, help='synthetic data: steps per epoch')
    p.add_argument('--opt', default='lamb', type=str, help="'lamb' (timm.optim.Lamb semantics, the recipes' optimizer) or 'adamw'")
    p.add_argument('--lr', type=float, default=5e-3)
    p.add_argument('--weight-decay', type=float, default=0.05)
    p.add_argument('--opt-eps', type=float, default=None)
    p.add_argument('--opt-betas', type=float, nargs=2, default=(0.9, 0.999))
    p.add_argument('--momentum', type=float, default=0.9, help='accepted for recipe compatibility (LAMB / AdamW use --opt-betas)')
    p.add_argument('--clip-grad', type=float, default=None, help='global-norm clip (--clip-mode norm)')
    p.add_argument('--clip-mode', type=str, default='norm')
    p.add_argument('--sched', default='cosine', type=str)
    p.add_argument('--warmup-epochs', '--warmup-epoch', type=int, default=0, dest='warmup_epochs')
    p.add_argument('--warmup-lr', type=float, default=1e-6)
    p.add_argument('--min-lr', type=float, default=1e-5)
    p.add_argument('--drop-path', type=float, default=None)
    p.add_argument('--drop', type=float, default=None)
    p.add_argument('--smoothing', type=float, default=0.1)
    p.add_argument('--bce-loss', action='store_true', default=False)
    p.add_argument('--bce-target-thresh', type=float, default=None)
    p.add_argument('--mixup', type=float, default=0.0)
    p.add_argument('--cutmix', type=float, default=0.0)
    p.add_argument('--mixup-prob', type=float, default=1.0)
    p.add_argument('--mixup-switch-prob', type=float, default=0.5)
    p.add_argument('--mixup-off-epoch', type=int, default=0)
    p.add_argument('--amp', action='store_true', default=False, help='bf16 autocast (the reference uses fp16 + GradScaler; bf16 needs no scaler)')
    p.add_argument('--channels-last', action='store_true', default=False)
    p.add_argument('--model-ema', action='store_true', default=False)
    p.add_argument('--model-ema-decay', type=float, default=0.9998)
    p.add_argument('--grad-accumulation', type=int, default=1)
    p.add_argument(lam_flag, default=lam_default, type=float, dest='group_lam')
    p.add_argument('--distill-tokens', default=0, type=float)
    p.add_argument('--token-distillation', default=1, type=float)
    p.add_argument('--no-ddp-bb', action='store_true', default=False, help='do not broadcast BatchNorm buffers before each forward')
    p.add_argument('--dist-bn', type=str, default='reduce', help="epoch-end BatchNorm sync: 'reduce', 'broadcast' or ''")
    p.add_argument('--seed', type=int, default=42)
    p.add_argument('--log-interval', type=int, default=50)
    p.add_argument('--initial-checkpoint', default='', type=str)
    p.add_argument('--resume', default='', type=str, help='last.pth.tar written by --output (weights, EMA, optimizer moments, epoch)')
    p.add_argument('--output', default='', type=str, help='directory for last.pth.tar')
    p.add_argument('--local_rank', default=0, type=int)
    p.add_argument('--no-cuda-graph', action='store_true', help='run every step eagerly (default: replay one captured CUDA graph per step)')
    return p


def run(args, unknown=()):
    logging.basicConfig(level=logging.INFO, format='%(message)s')
    if unknown:
        _logger.info('ignored flags (data pipeline / logging options without an effect on synthetic data): ' + ' '.join(unknown))
    if args.opt.lower() not in ('adamw', 'lamb'):
        raise SystemExit(f"--opt {args.opt}: adamw and lamb have fused sm_100a steps in this build")
    if args.distill_tokens > 0:
        raise SystemExit('--distill-tokens > 0 (DeiT-style distillation tokens) is not on the B200 path; the published MAP recipes use 0')
    distributed = int(os.environ.get('WORLD_SIZE', '1')) > 1
    local_rank = int(os.environ.get('LOCAL_RANK', args.local_rank))
    torch.cuda.set_device(local_rank)
    rank, world = 0, 1
    if distributed:
        os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')
        dist.init_process_group(backend='nccl', init_method='env://')
        rank, world = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(args.seed + rank)                      # timm random_seed(seed, rank), GA/train.py:402

    model = create_model(args.model, num_classes=args.num_classes, drop_path_rate=args.drop_path, drop_rate=args.drop,
                         checkpoint_path=args.initial_checkpoint).cuda()
    if args.channels_last:
        model = model.to(memory_format=torch.channels_last)
    if rank == 0:
        _logger.info(f'Model {args.model} created, param count:{sum(m.numel() for m in model.parameters())}')
    if distributed:                                          # DDP ctor broadcast of parameters and buffers from rank 0
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    opt = args.opt.lower()
    eps = args.opt_eps if args.opt_eps is not None else (1e-6 if opt == 'lamb' else 1e-8)
    engine = TrainEngine(model, lr=args.lr, weight_decay=args.weight_decay, betas=tuple(args.opt_betas), eps=eps,
                         ema_decay=args.model_ema_decay if args.model_ema else None, ga_lam=args.group_lam,
                         amp_dtype=torch.bfloat16 if args.amp else None, grad_accumulation=args.grad_accumulation,
                         cuda_graph=not args.no_cuda_graph, opt=opt, broadcast_buffers=not args.no_ddp_bb,
                         loss='bce' if args.bce_loss else 'ce', smoothing=args.smoothing, bce_target_thresh=args.bce_target_thresh,
                         clip_grad=args.clip_grad if args.clip_mode == 'norm' else None)
    start_epoch = 0
    if args.resume:
        ck = torch.load(args.resume, map_location='cuda', weights_only=False)
        model.load_state_dict(ck['state_dict'], strict=True)
        if engine.model_ema is not None and 'state_dict_ema' in ck:
            engine.model_ema.load_state_dict(ck['state_dict_ema'], strict=True)
        if 'optimizer' in ck:
            engine.opt.load_state_dict(ck['optimizer'])
        start_epoch = ck.get('epoch', -1) + 1
    sched = CosineSchedule(engine.opt, args.lr, args.epochs, args.warmup_epochs, args.warmup_lr, args.min_lr) if args.sched == 'cosine' else None
    mix = None
    if args.mixup > 0 or args.cutmix > 0:
        mix = Mixup(args.mixup, args.cutmix, args.mixup_prob, args.mixup_switch_prob, args.smoothing, model.num_classes, seed=args.seed + rank)
    B, S = args.batch_size, args.img_size
    g = torch.Generator().manual_seed(args.seed + rank)
    x_host = torch.randint(0, 256, (B, 3, S, S), dtype=torch.uint8, generator=g).pin_memory()
    y_host = torch.randint(0, model.num_classes, (B,), generator=g).pin_memory()
    prefetch = DevicePrefetcher(device=torch.device('cuda', local_rank), mixup=mix)
    reduce_kind = 'mean' if args.model.startswith('map_') else 'sum'
    for epoch in range(start_epoch, args.epochs):
        if sched is not None:
            sched.step(epoch)
        if mix is not None and args.mixup_off_epoch and epoch >= args.mixup_off_epoch:
            mix.enabled = False
        model.train()
        t0 = time.time()
        prefetch.submit(x_host, y_host)
        for it in range(args.steps_per_epoch):
            x, y = prefetch.get()
            if it + 1 < args.steps_per_epoch:
                prefetch.submit(x_host, y_host)
            if args.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            loss = engine.step(x, y)
            if it % args.log_interval == 0 or it == args.steps_per_epoch - 1:
                lv = loss.detach().clone()
                if distributed:                              # reduce_tensor(loss.data, world_size), GA/train.py:782
                    dist.all_reduce(lv)
                    lv /= world
                torch.cuda.synchronize()
                if not torch.isfinite(lv).item():            # MAP/train.py:887-891: NaN on any rank stops the run
                    if rank == 0:
                        print('nan occurs and exit')
                    raise SystemExit(0)
                dt = time.time() - t0
                if rank == 0:
                    _logger.info(f'Train: {epoch} [{it:>4d}/{args.steps_per_epoch}]  Loss: {lv.item():#.4g}  '
                                 f'Time: {dt / (it + 1):.3f}s, {B * world * (it + 1) / dt:>7.2f}/s  LR: {engine.opt.param_groups[0]["lr"]:.3e}')
        if distributed and args.dist_bn in ('broadcast', 'reduce'):   # GA/train.py:665-668
            engine.distribute_bn(args.dist_bn == 'reduce')
        model.eval()                                         # validate on one synthetic batch (GA sums the branch logits, MAP averages)
        xv = torch.randn(B, 3, S, S, device='cuda')
        yv = y_host.cuda()
        stats = torch.stack([t.float() for t in evaluate_batch(model, xv, yv, reduce_kind, torch.bfloat16 if args.amp else None)])
        if distributed:
            dist.all_reduce(stats)
            stats[0] /= world
        if rank == 0:
            n = stats[3].item()
            _logger.info(f'Test: Loss: {stats[0].item():.4f}  Acc@1: {100 * stats[1].item() / n:.3f}  Acc@5: {100 * stats[2].item() / n:.3f}')
            if args.output:
                os.makedirs(args.output, exist_ok=True)
                ck = {'epoch': epoch, 'arch': args.model, 'state_dict': model.state_dict(), 'optimizer': engine.opt.state_dict(),
                      'version': 2, 'args': args}
                if engine.model_ema is not None:
                    ck['state_dict_ema'] = engine.model_ema.state_dict()
                torch.save(ck, os.path.join(args.output, 'last.pth.tar'))
    if distributed:
        # no destroy_process_group(): it never returns once a CUDA graph with NCCL kernels of the communicator exists
        # (scripts/nccl_graph_probe.py); synchronise, flush and leave
        dist.barrier()
        torch.cuda.synchronize()
        import sys
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
