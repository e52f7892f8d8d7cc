"""The per-step pieces of GA/train.py and MAP/validate.py on the sm_100a kernels.

  TrainEngine.step  = train_one_epoch's loop body (GA/train.py:722-774): forward under bf16 autocast, GA loss
                      (sum of branch losses + lam * KL to the branch mean, :735-745), backward with the bucketed
                      gradient all-reduce overlapped (the DDP of :505-515), fused AdamW + EMA (:466, :499, :760-761).
  evaluate_batch    = validate()'s loop body: GA sums branch logits (GA/train.py:848-851), MAP averages them
                      (MAP/validate.py:275-279); top-1 / top-5 via torch.topk exactly like timm.utils.accuracy.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops
from .optim import FusedAdamWEma, FusedLambEma, GradBuckets


DDP_IN_GRAPH_DEFAULT = 0      # see TrainEngine(ddp_in_graph=...)


class TrainEngine:
    def __init__(self, model, lr=1e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, ema_decay: Optional[float] = 0.9998,
                 ga_lam: float = -0.8, amp_dtype=torch.bfloat16, grad_accumulation: int = 1, bucket_mb: float = 25.0,
                 cuda_graph: bool = False, graph_warmup: int = 3, opt: str = 'adamw', broadcast_buffers: bool = True,
                 ddp_in_graph: Optional[bool] = None, loss: str = 'ce', smoothing: float = 0.0, bce_target_thresh: Optional[float] = None,
                 clip_grad: Optional[float] = None):
        """cuda_graph: after `graph_warmup` eager steps the whole step (zero-grad, forward, loss, backward, the bucketed gradient
        all-reduce on its side stream, gradient gather, optimizer + EMA) is captured once into a CUDA graph and replayed, which
        removes the ~1.5k per-step kernel launches from the CPU's critical path.  Needs fixed batch shapes; drop-path masks
        are drawn inside the graph from the graph-safe generator."""
        self.model = model
        if opt == 'lamb':          # timm.optim.Lamb, the optimizer of the published recipes (GA/README.md:26)
            self.opt = FusedLambEma(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, ema_decay=ema_decay)
        elif opt == 'adamw':
            self.opt = FusedAdamWEma(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, ema_decay=ema_decay)
        else:
            raise ValueError(f"optimizer '{opt}': 'adamw' and 'lamb' are fused here")
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        if self.world > 1:
            # DistributedDataParallel construction (GA/train.py:505-515) copies rank 0's parameters and buffers to every rank
            # (the trainers seed with seed + rank, so the initialisations differ); the EMA copy and the bf16 shadows follow
            st = self.opt.state
            dist.broadcast(st.flat, 0)
            if st.bufflat is not None and st.bufflat.numel():
                dist.broadcast(st.bufflat, 0)
            if getattr(self.opt, 'ema_flat', None) is not None:
                self.opt.ema_flat.copy_(st.flat)
                if getattr(self.opt, 'ema_bufflat', None) is not None and st.bufflat is not None:
                    self.opt.ema_bufflat.copy_(st.bufflat)
            st.refresh_shadows()
        self.buckets = GradBuckets(self.opt.state, bucket_mb=bucket_mb) if self.world > 1 else None
        # DistributedDataParallel(broadcast_buffers=True) (GA/train.py:514, --no-ddp-bb turns it off): rank 0's BatchNorm running
        # statistics replace every rank's before each forward.  All float buffers live in one flat tensor -> one small broadcast.
        self.broadcast_buffers = bool(broadcast_buffers) and self.world > 1
        # classification term: 'ce' hard labels (LabelSmoothingCrossEntropy when smoothing > 0; SoftTargetCrossEntropy when the
        # loader hands dense mixup targets), 'bce' = timm BinaryCrossEntropy (--bce-loss of the published recipes)
        assert loss in ('ce', 'bce')
        self.loss_kind, self.smoothing, self.bce_thresh = loss, smoothing, bce_target_thresh
        self.clip_grad = clip_grad            # global-norm clip (timm dispatch_clip_grad mode 'norm'); LAMB clips inside its own step
        if clip_grad is not None and opt == 'adamw':
            self.opt.clip_grad = float(clip_grad)
        self.ga_lam, self.amp_dtype, self.accum = ga_lam, amp_dtype, max(1, grad_accumulation)
        self.micro = 0
        self.cuda_graph = bool(cuda_graph) and self.accum == 1
        # several ranks + CUDA graph: True captures the bucketed NCCL all-reduce (side stream) and the optimizer inside the step
        # graph; False replays forward / backward / gather and then runs ONE all-reduce of the flat gradient and the optimizer
        # eagerly.  GA_DDP_IN_GRAPH=0/1 overrides the default.
        if ddp_in_graph is None:
            import os
            ddp_in_graph = os.environ.get('GA_DDP_IN_GRAPH', '%d' % DDP_IN_GRAPH_DEFAULT) == '1'
        self.ddp_in_graph = bool(ddp_in_graph)
        self.graph_warmup, self._calls, self._graph = graph_warmup, 0, None
        self.graph_launches = 0                     # kernels captured per replay (ga_launch_count delta at capture)

    @property
    def model_ema(self):
        return self.opt.ema_model

    def distribute_bn(self, reduce: bool = True):
        """timm.utils.distribute_bn at epoch end (GA/train.py:665-668): average (reduce=True) or broadcast rank 0's BatchNorm
        running statistics, for the model and its EMA copy."""
        if self.world == 1:
            return
        flats = [self.opt.state.bufflat] + ([self.opt.ema_bufflat] if self.opt.ema_model is not None else [])
        for f in flats:
            if reduce:
                dist.all_reduce(f)
                f.div_(self.world)
            else:
                dist.broadcast(f, 0)

    def _loss(self, out, y):
        pairs = isinstance(out[0], (list, tuple))            # MAP train mode: [main, self-distillation] pairs per group
        main = torch.stack([o[0] for o in out]) if pairs else torch.stack(out)
        aux = torch.stack([o[1] for o in out]) if pairs else None
        dense = y.dtype != torch.int64
        if not dense and self.loss_kind == 'ce' and self.smoothing == 0.0:
            return ops.ga_loss(main, y, self.ga_lam, aux=aux)
        if not dense:
            y = ops.smooth_one_hot(y, main.shape[-1], self.smoothing, self.bce_thresh if self.loss_kind == 'bce' else None)
        elif self.loss_kind == 'bce' and self.bce_thresh is not None:
            y = y.gt(self.bce_thresh).float()
        return ops.ga_soft_loss(main, y, self.ga_lam, aux=aux, bce=self.loss_kind == 'bce')

    def _forward_backward(self, x, y, broadcast=True):
        if self.broadcast_buffers and broadcast:
            dist.broadcast(self.opt.state.bufflat, 0)
        if self.amp_dtype is not None:
            with torch.autocast('cuda', dtype=self.amp_dtype):
                out = self.model(x)
        else:
            out = self.model(x)
        loss = self._loss(out, y)
        loss.backward()
        return loss

    def _graph_step(self, x, y):
        from . import lib as L
        in_graph = self.world == 1 or self.ddp_in_graph
        if self._graph is None:
            self._sx, self._sy = torch.empty_like(x), torch.empty_like(y)
            self._sx.copy_(x)
            self._sy.copy_(y)
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            # several ranks, ddp_in_graph: the bucketed all-reduce is captured too -- every bucket's NCCL call sits on the side
            # stream, forked from the capture stream by the grad-ready hook that completes the bucket and joined before the
            # optimizer, so the replayed step overlaps the reduction with the rest of backward exactly like the eager path
            # (GA/train.py:505-515).  thread_local: the NCCL watchdog thread's event queries must not invalidate the capture.
            kw = {'capture_error_mode': 'thread_local'} if (self.world > 1 and in_graph) else {}
            with torch.cuda.graph(g, **kw):
                self.opt.zero_grad()
                if self.buckets is not None:
                    self.buckets.enabled = in_graph
                    if in_graph:
                        self.buckets.prepare()
                self._sloss = self._forward_backward(self._sx, self._sy, broadcast=in_graph).detach()
                if self.buckets is not None and in_graph:
                    self.buckets.finish()
                else:
                    self.opt.state.gather()
                if in_graph:
                    self.opt.step(gathered=True, device_hyper=True)
            self.graph_launches = L.launch_count() - n0
            self._graph = g
            from . import ops
            ops.ZEROS.reset()
        else:
            self._sx.copy_(x, non_blocking=True)
            self._sy.copy_(y, non_blocking=True)
        if in_graph:
            self.opt.push_hyper(1.0 / self.world)      # lr, bias corrections and the 1/world gradient scale, read by the captured optimizer
            self._graph.replay()
        else:
            if self.broadcast_buffers:
                dist.broadcast(self.opt.state.bufflat, 0)
            self._graph.replay()
            dist.all_reduce(self.opt.state.grad)
            self.opt.step(grad_scale=1.0 / self.world, gathered=True)
        return self._sloss

    def step(self, x, y):
        """One micro-step; the optimizer (and the all-reduce) runs every `grad_accumulation` micro-steps.  Returns the loss tensor."""
        self._calls += 1
        if self._graph is None and not self.opt.state.shadows_current():
            self.opt.state.refresh_shadows()        # parameters were written from outside (load_state_dict before training)
        if self.cuda_graph and self._calls > self.graph_warmup:
            return self._graph_step(x, y)
        first = self.micro % self.accum == 0
        last = (self.micro + 1) % self.accum == 0
        if first:
            self.opt.zero_grad()
        if self.buckets is not None:
            # the reference all-reduces on every micro-step (no no_sync(), GA/train.py:750-761); reducing once on the
            # update step moves 1/accum of the bytes for the same result
            self.buckets.enabled = last
            if last:
                self.buckets.prepare()
        if self.broadcast_buffers:
            dist.broadcast(self.opt.state.bufflat, 0)
        if self.amp_dtype is not None:
            with torch.autocast('cuda', dtype=self.amp_dtype):
                out = self.model(x)
        else:
            out = self.model(x)
        loss = self._loss(out, y)
        (loss / self.accum if self.accum > 1 else loss).backward()
        if last:
            scale = self.buckets.finish() if self.buckets is not None else 1.0
            self.opt.step(grad_scale=scale, gathered=self.buckets is not None)
        self.micro += 1
        # detached: a caller holding the loss must not keep this step's autograd graph (and its AccumulateGrad nodes, bound to
        # the eager stream) alive into the CUDA-graph capture of a later step
        return loss.detach()


class Mixup:
    """timm.data.Mixup (mode 'batch') on the device, GA/train.py:545-557: per batch one lambda ~ Beta(alpha, alpha); mixup blends
    the batch with itself in reverse order, cutmix pastes a random box from it (lambda corrected to the box area); the targets
    become lam * smooth_one_hot(y) + (1 - lam) * smooth_one_hot(y.flip(0)).  The host RNG is numpy's, as in timm.
    draw() -> (mode, lam, (y0, y1, x0, x1)) for DevicePrefetcher / ga_prep_batch; targets() builds the dense targets.
    timm is absent from this image, so this follows its published algorithm (parity unpinned)."""

    def __init__(self, mixup_alpha=0.8, cutmix_alpha=1.0, prob=1.0, switch_prob=0.5, label_smoothing=0.1, num_classes=1000, seed=None):
        import numpy as np
        self.mixup_alpha, self.cutmix_alpha, self.prob, self.switch_prob = mixup_alpha, cutmix_alpha, prob, switch_prob
        self.smoothing, self.num_classes = label_smoothing, num_classes
        self.rng = np.random.RandomState(seed)
        self.enabled = True

    def draw(self, H, W):
        r = self.rng
        if not self.enabled or r.rand() >= self.prob or (self.mixup_alpha <= 0 and self.cutmix_alpha <= 0):
            return 0, 1.0, (0, 0, 0, 0)
        use_cutmix = self.cutmix_alpha > 0 and (self.mixup_alpha <= 0 or r.rand() < self.switch_prob)
        lam = float(r.beta(self.cutmix_alpha, self.cutmix_alpha) if use_cutmix else r.beta(self.mixup_alpha, self.mixup_alpha))
        if not use_cutmix:
            return 1, lam, (0, 0, 0, 0)
        ratio = (1.0 - lam) ** 0.5                                     # timm rand_bbox
        ch, cw = int(H * ratio), int(W * ratio)
        cy, cx = r.randint(0, H), r.randint(0, W)
        y0, y1 = max(cy - ch // 2, 0), min(cy + ch // 2, H)
        x0, x1 = max(cx - cw // 2, 0), min(cx + cw // 2, W)
        lam = 1.0 - (y1 - y0) * (x1 - x0) / float(H * W)               # correct_lam
        return 2, lam, (y0, y1, x0, x1)

    def targets(self, y, lam):
        t = ops.smooth_one_hot(y, self.num_classes, self.smoothing)
        return t if lam == 1.0 else t * lam + t.flip(0) * (1.0 - lam)


class DevicePrefetcher:
    """timm `PrefetchLoader` semantics (create_loader(..., use_prefetcher=True), GA/train.py:598-626): the NEXT batch's uint8
    host-to-device copy and its normalisation (x - mean*255) / (std*255) -- and, with a Mixup, the mixing with the reversed
    batch -- run on a side stream under the current step, as ONE kernel (ga_prep_batch) behind the copy.
    submit(pinned uint8 [B,3,H,W], pinned int64 [B]) -> later get() -> (float x, y or dense targets)."""

    def __init__(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), device='cuda', mixup: Optional[Mixup] = None):
        import ctypes as C
        self.device = torch.device(device)
        self.mean = (C.c_float * 3)(*mean)
        self.std = (C.c_float * 3)(*std)
        self.stream = torch.cuda.Stream(device=self.device)
        self.mixup = mixup
        self._pending = None

    def submit(self, x_u8: torch.Tensor, y: torch.Tensor):
        from . import lib as L
        assert self._pending is None, 'one batch in flight'
        assert x_u8.dtype == torch.uint8 and x_u8.dim() == 4 and x_u8.shape[1] == 3 and x_u8.is_contiguous()
        Bn, _, H, W = x_u8.shape
        mode, lam, box = (0, 1.0, (0, 0, 0, 0)) if self.mixup is None else self.mixup.draw(H, W)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            xd = x_u8.to(self.device, non_blocking=True)
            yd = y.to(self.device, non_blocking=True)
            x = torch.empty(Bn, 3, H, W, dtype=torch.float32, device=self.device)
            L.check(L.load().ga_prep_batch(L.ptr(xd), L.ptr(x), Bn, H, W, self.mean, self.std, mode, L.f(lam), box[0], box[1], box[2],
                                           box[3], L.stream()), 'ga_prep_batch')
            if self.mixup is not None:
                yd = self.mixup.targets(yd, lam if mode else 1.0)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (x, yd, ev, xd)

    def get(self):
        x, y, ev, xd = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        x.record_stream(cur)
        y.record_stream(cur)
        return x, y


class CosineSchedule:
    """timm CosineLRScheduler as the recipes use it (--sched cosine --warmup-epochs W --warmup-lr L0 --min-lr Lmin, stepped per
    epoch; GA/train.py:518-531): linear warm-up from warmup_lr to lr over `warmup_epochs`, then
    min_lr + 0.5 (lr - min_lr) (1 + cos(pi * epoch / epochs)).  timm is absent: its published formula, parity unpinned."""

    def __init__(self, optimizer, base_lr, epochs, warmup_epochs=0, warmup_lr=1e-6, min_lr=1e-5):
        self.opt, self.base, self.epochs = optimizer, base_lr, max(1, epochs)
        self.warm, self.warm_lr, self.min_lr = warmup_epochs, warmup_lr, min_lr
        self.step(0)

    def lr_at(self, epoch):
        import math
        if epoch < self.warm:
            return self.warm_lr + epoch * (self.base - self.warm_lr) / self.warm
        return self.min_lr + 0.5 * (self.base - self.min_lr) * (1 + math.cos(math.pi * min(epoch, self.epochs) / self.epochs))

    def step(self, epoch):
        self.opt.param_groups[0]['lr'] = self.lr_at(epoch)


@torch.no_grad()
def evaluate_batch(model, x, y, reduce: str = 'sum', amp_dtype=torch.bfloat16):
    """-> (loss, correct@1, correct@5, count) as device tensors (no host sync).  reduce: 'sum' (GA) | 'mean' (MAP)."""
    if amp_dtype is not None:
        with torch.autocast('cuda', dtype=amp_dtype):
            outs = model(x)
    else:
        outs = model(x)
    logits = torch.stack([o.float() for o in outs]).sum(0)
    if reduce == 'mean':
        logits = logits / len(outs)
    loss = torch.nn.functional.cross_entropy(logits, y)
    _, pred = logits.topk(5, 1, True, True)           # timm.utils.accuracy
    hit = pred.eq(y.view(-1, 1))
    return loss, hit[:, :1].sum(), hit.sum(), torch.full((), y.numel(), dtype=torch.int64, device=y.device)   # no H2D copy: graph-safe


class EvalEngine:
    """validate()'s loop body with the forward captured in a CUDA graph per input shape (MAP/validate.py:250-311,
    GA/train.py:838-868).  Small-batch inference is launch-bound (a T-688 forward is ~450 kernels): a replay costs one
    launch.  __call__(x, y) -> (loss, correct@1, correct@5, count) device tensors, valid until the next call."""

    def __init__(self, model, reduce: str = 'sum', amp_dtype=torch.bfloat16, cuda_graph: bool = True, graph_warmup: int = 2):
        self.model, self.reduce, self.amp_dtype = model.eval(), reduce, amp_dtype
        self.cuda_graph, self.graph_warmup = cuda_graph, graph_warmup
        self._graphs = {}          # (shape, dtype) -> (graph, static x, static y, static outputs)
        self._seen = {}

    def __call__(self, x, y):
        if not self.cuda_graph:
            return evaluate_batch(self.model, x, y, self.reduce, self.amp_dtype)
        key = (tuple(x.shape), x.dtype, tuple(x.stride()))
        ent = self._graphs.get(key)
        if ent is None:
            n = self._seen.get(key, 0)
            self._seen[key] = n + 1
            if n < self.graph_warmup:
                return evaluate_batch(self.model, x, y, self.reduce, self.amp_dtype)
            sx, sy = torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device), torch.empty_like(y)
            sx.copy_(x)
            sy.copy_(y)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = evaluate_batch(self.model, sx, sy, self.reduce, self.amp_dtype)
            ent = self._graphs[key] = (g, sx, sy, out)
        else:
            ent[1].copy_(x, non_blocking=True)
            ent[2].copy_(y, non_blocking=True)
        ent[0].replay()
        return ent[3]
