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
from .optim import FusedAdamWEma, GradBuckets


class TrainEngine:
    def __init__(self, model, lr=1e-3, weight_decay=0.05, betas=(0.9, 0.999), eps=1e-8, ema_decay: Optional[float] = 0.9998,
                 ga_lam: float = -0.8, amp_dtype=torch.bfloat16, grad_accumulation: int = 1, bucket_mb: float = 25.0):
        self.model = model
        self.opt = FusedAdamWEma(model, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, ema_decay=ema_decay)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.buckets = GradBuckets(self.opt.state, bucket_mb=bucket_mb) if self.world > 1 else None
        self.ga_lam, self.amp_dtype, self.accum = ga_lam, amp_dtype, max(1, grad_accumulation)
        self.micro = 0

    @property
    def model_ema(self):
        return self.opt.ema_model

    def step(self, x, y):
        """One micro-step; the optimizer (and the all-reduce) runs every `grad_accumulation` micro-steps.  Returns the loss tensor."""
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
            self.opt.step(grad_scale=scale)
        self.micro += 1
        return loss


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
    return loss, hit[:, :1].sum(), hit.sum(), torch.tensor(y.numel(), device=y.device)
