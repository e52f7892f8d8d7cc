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


class TrainEngine:
    def __init__(self, model, lr=1e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, ema_decay: Optional[float] = 0.9998,
                 ga_lam: float = -0.8, amp_dtype=torch.bfloat16, grad_accumulation: int = 1, bucket_mb: float = 25.0,
                 cuda_graph: bool = False, graph_warmup: int = 3, opt: str = 'adamw', broadcast_buffers: bool = True):
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
        self.buckets = GradBuckets(self.opt.state, bucket_mb=bucket_mb) if self.world > 1 else None
        # DistributedDataParallel(broadcast_buffers=True) (GA/train.py:514, --no-ddp-bb turns it off): rank 0's BatchNorm running
        # statistics replace every rank's before each forward.  All float buffers live in one flat tensor -> one small broadcast.
        self.broadcast_buffers = bool(broadcast_buffers) and self.world > 1
        self.ga_lam, self.amp_dtype, self.accum = ga_lam, amp_dtype, max(1, grad_accumulation)
        self.micro = 0
        self.cuda_graph = bool(cuda_graph) and self.accum == 1
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

    def _forward_backward(self, x, y):
        if self.broadcast_buffers:
            dist.broadcast(self.opt.state.bufflat, 0)
        if self.amp_dtype is not None:
            with torch.autocast('cuda', dtype=self.amp_dtype):
                out = self.model(x)
        else:
            out = self.model(x)
        if isinstance(out[0], (list, tuple)):            # MAP train mode: [main, self-distillation] pairs per group
            loss = ops.ga_loss(torch.stack([o[0] for o in out]), y, self.ga_lam, aux=torch.stack([o[1] for o in out]))
        else:
            loss = ops.ga_loss(torch.stack(out), y, self.ga_lam)
        loss.backward()
        return loss

    def _graph_step(self, x, y):
        from . import lib as L
        if self._graph is None:
            self._sx, self._sy = torch.empty_like(x), torch.empty_like(y)
            self._sx.copy_(x)
            self._sy.copy_(y)
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            # several ranks: the bucketed all-reduce is captured too -- every bucket's NCCL call sits on the side stream, forked
            # from the capture stream by the grad-ready hook that completes the bucket and joined before the optimizer, so the
            # replayed step overlaps the reduction with the rest of backward exactly like the eager path (GA/train.py:505-515).
            # thread_local: the NCCL watchdog thread's event queries must not invalidate the capture.
            kw = {'capture_error_mode': 'thread_local'} if self.world > 1 else {}
            with torch.cuda.graph(g, **kw):
                self.opt.zero_grad()
                if self.buckets is not None:
                    self.buckets.enabled = True
                    self.buckets.prepare()
                self._sloss = self._forward_backward(self._sx, self._sy).detach()
                if self.buckets is not None:
                    self.buckets.finish()
                else:
                    self.opt.state.gather()
                self.opt.step(gathered=True, device_hyper=True)
            self.graph_launches = L.launch_count() - n0
            self._graph = g
        else:
            self._sx.copy_(x, non_blocking=True)
            self._sy.copy_(y, non_blocking=True)
        self.opt.push_hyper(1.0 / self.world)          # lr, bias corrections and the 1/world gradient scale, read by the captured optimizer
        self._graph.replay()
        return self._sloss

    def step(self, x, y):
        """One micro-step; the optimizer (and the all-reduce) runs every `grad_accumulation` micro-steps.  Returns the loss tensor."""
        self._calls += 1
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
        if isinstance(out[0], (list, tuple)):            # MAP train mode: [main, self-distillation] pairs per group
            loss = ops.ga_loss(torch.stack([o[0] for o in out]), y, self.ga_lam, aux=torch.stack([o[1] for o in out]))
        else:
            loss = ops.ga_loss(torch.stack(out), y, self.ga_lam)
        (loss / self.accum if self.accum > 1 else loss).backward()
        if last:
            scale = self.buckets.finish() if self.buckets is not None else 1.0
            self.opt.step(grad_scale=scale, gathered=self.buckets is not None)
        self.micro += 1
        # detached: a caller holding the loss must not keep this step's autograd graph (and its AccumulateGrad nodes, bound to
        # the eager stream) alive into the CUDA-graph capture of a later step
        return loss.detach()


class DevicePrefetcher:
    """timm `PrefetchLoader` semantics (create_loader(..., use_prefetcher=True), GA/train.py:598-626): the NEXT batch's uint8
    host-to-device copy and its normalisation (x - mean*255) / (std*255) run on a side stream under the current step, so PCIe
    time disappears from the step.  submit(pinned uint8 [B,3,H,W], pinned int64 [B]) -> later get() -> (float x, y)."""

    def __init__(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), device='cuda'):
        self.device = torch.device(device)
        self.mean = torch.tensor(mean, device=self.device).view(1, 3, 1, 1) * 255
        self.std = torch.tensor(std, device=self.device).view(1, 3, 1, 1) * 255
        self.stream = torch.cuda.Stream(device=self.device)
        self._pending = None

    def submit(self, x_u8: torch.Tensor, y: torch.Tensor):
        assert self._pending is None, 'one batch in flight'
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            x = x_u8.to(self.device, non_blocking=True).float().sub_(self.mean).div_(self.std)
            yd = y.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (x, yd, ev)

    def get(self):
        x, y, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        x.record_stream(cur)
        y.record_stream(cur)
        return x, y


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
