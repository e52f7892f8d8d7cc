"""CPU oracle for the LAMB step  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference trains with `--opt lamb` (GA/README.md:26, GA/train_with_script.py:13-19) through timm's
`create_optimizer_v2` (GA/train.py:466).  The algorithm lives in a third-party dependency that is ABSENT from
/root/reference and from this image: **timm** (`timm/optim/lamb.py`; pinned `timm>=0.4.5` by GA/README.md:13 and
`timm==0.9.2` by MAP/README.MD:16).  This file restates its published algorithm (You et al., "Large Batch Optimization
for Deep Learning", as implemented by timm.optim.Lamb with its defaults: bias_correction=True, grad_averaging=True,
max_grad_norm=1.0, trust_clip=False, always_adapt=False; eps 1e-6) per tensor in plain PyTorch.

Parity status: **UNPINNED** -- neither timm nor any golden vector of its Lamb is available offline; the fused kernel is
checked against this restatement only (tests/test_engine_gpu.py).
"""
import math
from typing import List

import torch


def lamb_step(params: List[torch.Tensor], grads: List[torch.Tensor], exp_avg: List[torch.Tensor], exp_avg_sq: List[torch.Tensor],
              decay: List[bool], step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-6, weight_decay: float = 0.01,
              max_grad_norm: float = 1.0) -> None:
    """One in-place LAMB step over tensors; `decay[i]` False = the no-weight-decay group (biases, 1-D tensors), which timm
    updates without the trust ratio.  `step` is the 1-based step count."""
    beta1, beta2 = betas
    gnorm = torch.sqrt(sum(g.double().pow(2).sum() for g in grads)).float()
    clip = gnorm / max_grad_norm if (max_grad_norm and gnorm > max_grad_norm) else torch.tensor(1.0)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    for p, g, m, v, d in zip(params, grads, exp_avg, exp_avg_sq, decay):
        g = g / clip
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        update = (m / bc1).div_(denom)
        wd = weight_decay if d else 0.0
        if wd != 0:
            update.add_(p, alpha=wd)
            w_norm, u_norm = p.norm(2.0), update.norm(2.0)
            trust = (w_norm / u_norm) if (w_norm > 0 and u_norm > 0) else torch.tensor(1.0)
            update.mul_(trust)
        p.add_(update, alpha=-lr)
