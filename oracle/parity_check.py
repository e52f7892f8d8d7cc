"""Whole-model training-step parity at the BASELINE.json shapes  --  TEST INFRASTRUCTURE (used by tests/ and scripts/ only).

measure(key, dtype) runs the CUDA implementation on fixture `key` of oracle/cases.py GA_PARITY_CASES and returns every error
the parity contract names (north_star: 1e-5 fp32 / 2e-2 bf16 on logits and gradients):

  logits            vs the REFERENCE's train-mode logits (tests/golden/ga_convnext_parity.pt)
  loss              vs the reference's loss
  tail gradients    parameters after the Bottleneck (no ReLU between them and the loss): vs the REFERENCE's gradients
  all gradients     vs the ORACLE (pinned to the reference by make_golden.py) evaluated at the implementation's own ReLU
                    decisions.  A ReLU network's gradient is discontinuous in its pre-activations: the reference run on inputs
                    perturbed by 1e-7 flips single decisions of the 2.2 M in the Bottleneck and its own upstream gradients move
                    by 1e-3 (`ref_self_noise` in the fixture), so no two implementations can agree to 1e-5 -- or, in bf16, where
                    ~0.5 % of the decisions sit inside the rounding noise, to 2e-2 -- unless the decisions are the same.
  raw gradients     upstream parameters vs the reference's gradients, reported together with the number of differing decisions
"""
import os

import numpy as np
import torch

from . import cases
from . import ga_convnext_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'ga_convnext_parity.pt')


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def measure(key, dtype, fixture=None):
    from imagenet_models_b200 import ops
    from imagenet_models_b200.registry import create_model
    import imagenet_models_b200.ga_convnext  # noqa: F401
    name, B, profile, kind = cases.GA_PARITY_CASES[key]
    g = (fixture or torch.load(GOLDEN))[key]
    spec = O.SPECS[name]
    P = O.make_state(spec, cases.STATE_SEED, profile=profile)
    x, y = cases.parity_inputs(kind, B)
    m = create_model(name).cuda()
    m.load_state_dict({k: v.cuda() for k, v in P.items()}, strict=True)
    m.compute_dtype = dtype
    m.train()
    ops.RELU_TAP = []
    try:
        out = m(x.cuda())
        taps = ops.RELU_TAP
    finally:
        ops.RELU_TAP = None
    loss = ops.ga_loss(torch.stack(out), y.cuda(), cases.GA_LAM)
    loss.backward()
    torch.cuda.synchronize()
    res = {'key': key, 'dtype': str(dtype).split('.')[-1], 'B': B}
    res['logits'] = max(rel(a.detach().cpu(), b) for a, b in zip(out, g['train_logits']))
    res['loss'] = abs(loss.item() - g['loss'].item()) / abs(g['loss'].item())
    grads = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()}
    # ---- raw comparison with the reference's gradient digests
    raw = {k: (cases.digest_rel_err(grads[k], g['grads'][k]), g['grads'][k][0]) for k in grads}
    tail = {k: v for k, v in raw.items() if k.startswith(cases.GA_TAIL_PREFIXES)}
    up = {k: v for k, v in raw.items() if not k.startswith(cases.GA_TAIL_PREFIXES)}
    live = lambda d: {k: e for k, (e, n) in d.items() if n > ZERO_NORM}   # noqa: E731
    res['tail_grads'] = live(tail)
    res['upstream_grads_raw'] = live(up)
    # ---- ReLU decisions: ours vs the reference's
    bn_taps = [t for kind, t in taps if kind == 'bn']
    se_taps = [t for kind, t in taps if kind == 'se']
    assert len(bn_taps) == 3 and len(se_taps) == 1, (len(bn_taps), len(se_taps))
    H = W = 14
    masks = {}
    for kname, t in zip(('bn1', 'bn2', 'out'), bn_taps):
        masks[kname] = t.view(B, H, W, -1).permute(0, 3, 1, 2).cpu()
    masks['se'] = se_taps[0].view(B, -1, 1, 1).cpu()
    flips = 0
    for kname, packed in g['relu_masks'].items():                  # the three large ReLUs (the SE one is not in the fixture)
        refm = torch.from_numpy(np.unpackbits(packed.numpy())[:masks[kname].numel()].astype(bool)).view(masks[kname].shape)
        flips += int((refm != masks[kname]).sum())
    res['relu_flips_vs_reference'] = flips
    res['relu_decisions'] = sum(v.numel() for v in masks.values())
    # ---- oracle at our decisions
    Po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in P.items()}
    torch.set_num_threads(os.cpu_count())
    o_out = O.forward(Po, spec, x, training=True, relu_masks=masks)
    O.ga_loss(o_out, y, cases.GA_LAM).backward()
    res['logits_vs_pinned_oracle'] = max(rel(a.detach().cpu(), b.detach()) for a, b in zip(out, o_out))
    pinned = {}
    for k in grads:
        ref = Po[k].grad
        if ref.norm().item() > ZERO_NORM:
            pinned[k] = rel(grads[k], ref)
    res['grads_pinned'] = pinned
    sd = m.state_dict()
    res['running'] = max(rel(sd[k].cpu(), v) for k, v in g['running'].items())
    return res


# parameters whose true gradient is exactly zero (a conv bias feeding a train-mode BatchNorm): both sides hold rounding noise only
ZERO_NORM = 1e-3


def summarise(res):
    def stats(d):
        v = sorted(d.values())
        worst = max(d.items(), key=lambda kv: kv[1])
        return {'n': len(v), 'median': v[len(v) // 2], 'max': v[-1], 'argmax': worst[0]}
    return {'key': res['key'], 'dtype': res['dtype'], 'B': res['B'], 'logits': res['logits'], 'loss': res['loss'],
            'logits_vs_pinned_oracle': res['logits_vs_pinned_oracle'], 'running_stats': res['running'],
            'tail_grads_vs_reference': stats(res['tail_grads']), 'all_grads_vs_pinned_oracle': stats(res['grads_pinned']),
            'upstream_grads_vs_reference_raw': stats(res['upstream_grads_raw']),
            'relu_flips_vs_reference': res['relu_flips_vs_reference'], 'relu_decisions': res['relu_decisions']}
