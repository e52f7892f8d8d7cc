#!/usr/bin/env python3
"""Drop-in entry point for the reference's `GA/train.py` on the sm_100a hot path.

Keeps the reference's launch convention (`torchrun --nproc_per_node=N GA/train.py <data> --model ga_convnext_tiny_688 ...`,
one process per GPU, NCCL, env:// rendezvous; GA/train.py:374-381) and the flags that reach the hot path.  The timm
runtime the reference imports (loaders, augmentation, schedulers, checkpoint saver) is not in this image, so data is
synthetic unless a tensor dataset is given; everything on the device side -- model, loss, backward, gradient
all-reduce, AdamW, EMA -- is the B200-native implementation (imagenet_models_b200).
"""
import argparse
import logging
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imagenet_models_b200.ga_convnext  # noqa: F401,E402  (registers the ga_convnext_* factories)
import imagenet_models_b200.map_convnext  # noqa: F401,E402  (registers map_convnext_tiny / _small)
import imagenet_models_b200.ga_cswin  # noqa: F401,E402  (GA_CSWinTransformer, ga_CSWin_64_12211_tiny_224)
from imagenet_models_b200.engine import TrainEngine, evaluate_batch  # noqa: E402
from imagenet_models_b200.registry import create_model  # noqa: E402

_logger = logging.getLogger('train')

parser = argparse.ArgumentParser(description='GA training on B200 (flags follow GA/train.py:69-309)')
parser.add_argument('data_dir', nargs='?', default='', help='unused unless --tensor-data is given (no dataset on this box)')
parser.add_argument('--model', default='ga_convnext_tiny_688', type=str)
parser.add_argument('--num-classes', type=int, default=None)
parser.add_argument('-b', '--batch-size', type=int, default=128, help='per-process batch')
parser.add_argument('--img-size', type=int, default=224)
parser.add_argument('--epochs', type=int, default=1)
parser.add_argument('--steps-per-epoch', type=int, default=50, help='synthetic data: steps per epoch')
parser.add_argument('--opt', default='adamw', type=str, help="'adamw' or 'lamb' (timm.optim.Lamb semantics, the recipes' optimizer)")
parser.add_argument('--lr', type=float, default=1e-3)
parser.add_argument('--weight-decay', type=float, default=0.05)
parser.add_argument('--opt-eps', type=float, default=1e-8)
parser.add_argument('--opt-betas', type=float, nargs=2, default=(0.9, 0.999))
parser.add_argument('--drop-path', type=float, default=None)
parser.add_argument('--amp', action='store_true', default=False, help='bf16 autocast (the reference uses fp16 + GradScaler)')
parser.add_argument('--channels-last', action='store_true', default=False)
parser.add_argument('--model-ema', action='store_true', default=False)
parser.add_argument('--model-ema-decay', type=float, default=0.9998)
parser.add_argument('--grad-accumulation', type=int, default=1)
parser.add_argument('--GA_lam', default=0, type=float)
parser.add_argument('--seed', type=int, default=42)
parser.add_argument('--log-interval', type=int, default=50)
parser.add_argument('--initial-checkpoint', default='', type=str)
parser.add_argument('--output', default='', type=str, help='directory for last.pth.tar (state_dict, state_dict_ema, epoch)')
parser.add_argument('--local_rank', default=0, type=int)
parser.add_argument('--no-cuda-graph', action='store_true', help='run every step eagerly (default: replay one captured CUDA graph per step)')


def main():
    logging.basicConfig(level=logging.INFO, format='%(message)s')
    args = parser.parse_args()
    if args.opt.lower() not in ('adamw', 'lamb'):
        raise SystemExit(f"--opt {args.opt}: adamw and lamb have fused sm_100a steps in this build")
    distributed = int(os.environ.get('WORLD_SIZE', '1')) > 1
    local_rank = int(os.environ.get('LOCAL_RANK', args.local_rank))
    torch.cuda.set_device(local_rank)
    rank, world = 0, 1
    if distributed:
        os.environ.setdefault('TORCH_NCCL_ASYNC_ERROR_HANDLING', '0')     # the step's NCCL calls are captured in a CUDA graph
        dist.init_process_group(backend='nccl', init_method='env://')
        rank, world = dist.get_rank(), dist.get_world_size()
    torch.manual_seed(args.seed + rank)                      # timm random_seed(seed, rank), GA/train.py:402

    model = create_model(args.model, num_classes=args.num_classes, drop_path_rate=args.drop_path,
                         checkpoint_path=args.initial_checkpoint).cuda()
    if args.channels_last:
        model = model.to(memory_format=torch.channels_last)
    if rank == 0:
        _logger.info(f'Model {args.model} created, param count:{sum(m.numel() for m in model.parameters())}')
    if distributed:                                          # DDP ctor broadcast of parameters and buffers from rank 0
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    engine = TrainEngine(model, lr=args.lr, weight_decay=args.weight_decay, betas=tuple(args.opt_betas), eps=args.opt_eps,
                         ema_decay=args.model_ema_decay if args.model_ema else None, ga_lam=args.GA_lam,
                         amp_dtype=torch.bfloat16 if args.amp else None, grad_accumulation=args.grad_accumulation,
                         cuda_graph=not args.no_cuda_graph, opt=args.opt.lower())
    B, S = args.batch_size, args.img_size
    g = torch.Generator(device='cuda').manual_seed(args.seed + rank)
    for epoch in range(args.epochs):
        model.train()
        t0 = time.time()
        for it in range(args.steps_per_epoch):
            x = torch.randn(B, 3, S, S, device='cuda', generator=g)
            y = torch.randint(0, model.num_classes, (B,), device='cuda', generator=g)
            if args.channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            loss = engine.step(x, y)
            if it % args.log_interval == 0 or it == args.steps_per_epoch - 1:
                lv = loss.detach().clone()
                if distributed:                              # reduce_tensor(loss.data, world_size), GA/train.py:782
                    dist.all_reduce(lv)
                    lv /= world
                torch.cuda.synchronize()
                dt = time.time() - t0
                if rank == 0:
                    _logger.info(f'Train: {epoch} [{it:>4d}/{args.steps_per_epoch}]  Loss: {lv.item():#.4g}  '
                                 f'Time: {dt / (it + 1):.3f}s, {B * world * (it + 1) / dt:>7.2f}/s  LR: {args.lr:.3e}')
        # validate on one synthetic batch (GA sums the branch logits, GA/train.py:848-851)
        model.eval()
        stats = torch.stack([t.float() for t in evaluate_batch(model, x, y, 'mean' if args.model.startswith('map_') else 'sum', torch.bfloat16 if args.amp else None)])
        if distributed:
            stats[1:] = stats[1:].clone()
            dist.all_reduce(stats)
            stats[0] /= world
        if rank == 0:
            n = stats[3].item()
            _logger.info(f'Test: Loss: {stats[0].item():.4f}  Acc@1: {100 * stats[1].item() / n:.3f}  Acc@5: {100 * stats[2].item() / n:.3f}')
            if args.output:
                os.makedirs(args.output, exist_ok=True)
                ck = {'epoch': epoch, 'arch': args.model, 'state_dict': model.state_dict()}
                if engine.model_ema is not None:
                    ck['state_dict_ema'] = engine.model_ema.state_dict()
                torch.save(ck, os.path.join(args.output, 'last.pth.tar'))
    if distributed:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
