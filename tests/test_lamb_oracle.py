"""CPU: the LAMB restatement (oracle/lamb_oracle.py) against an independent float64 transcription of the published update rule
(You et al. 2019, as configured by timm.optim.Lamb: bias correction, global-norm clip at 1.0, trust ratio on decayed tensors)."""
import math

import torch

from oracle.lamb_oracle import lamb_step


def _reference(params, grads, m, v, decay, t, lr, b1, b2, eps, wd, max_norm):
    gn = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
    clip = gn / max_norm if gn > max_norm else 1.0
    out = []
    for p, g, mi, vi, d in zip(params, grads, m, v, decay):
        g = g.double() / clip
        mi.mul_(b1).add_((1 - b1) * g)
        vi.mul_(b2).add_((1 - b2) * g * g)
        upd = (mi / (1 - b1 ** t)) / ((vi / (1 - b2 ** t)).sqrt() + eps)
        if d:
            upd = upd + wd * p
            wn, un = float(p.norm()), float(upd.norm())
            if wn > 0 and un > 0:
                upd = upd * (wn / un)
        out.append(p - lr * upd)
    return out


def test_lamb_oracle_matches_independent_transcription():
    g = torch.Generator().manual_seed(0)
    shapes = [(64, 37), (64,), (5, 64), (5,)]
    decay = [True, False, True, False]
    p32 = [torch.randn(s, generator=g) for s in shapes]
    p64 = [p.double().clone() for p in p32]
    m32, v32 = [torch.zeros(s) for s in shapes], [torch.zeros(s) for s in shapes]
    m64, v64 = [torch.zeros(s, dtype=torch.float64) for s in shapes], [torch.zeros(s, dtype=torch.float64) for s in shapes]
    for t in range(1, 4):
        grads = [torch.randn(s, generator=g) * (2.0 if t == 1 else 0.05) for s in shapes]     # step 1 clips, later steps do not
        lamb_step(p32, grads, m32, v32, decay, t, lr=5e-3, eps=1e-6, weight_decay=0.05, max_grad_norm=1.0)
        p64 = _reference(p64, grads, m64, v64, decay, t, 5e-3, 0.9, 0.999, 1e-6, 0.05, 1.0)
    for a, b in zip(p32, p64):
        assert torch.allclose(a.double(), b, rtol=1e-5, atol=1e-6)
    assert not torch.allclose(p32[1], torch.zeros(64))
