"""Whole-model training-step parity at the BASELINE.json shapes  --  TEST INFRASTRUCTURE (used by tests/ and scripts/ only).

measure(family, key, dtype) runs the CUDA implementation on a fixture of oracle/cases.py ({GA,MAP,CSWIN}_PARITY_CASES) and
returns every error the parity contract names (north_star: 1e-5 fp32 / 2e-2 bf16 on logits and gradients):

  logits            vs the REFERENCE's train-mode logits (tests/golden/*_parity.pt)
  loss              vs the reference's loss
  tail gradients    parameters with no ReLU between them and the loss: vs the REFERENCE's gradients
  all gradients     vs the ORACLE (pinned to the reference by make_golden.py) evaluated at the implementation's own ReLU
                    decisions.  A ReLU network's gradient is discontinuous in its pre-activations: the reference run on inputs
                    perturbed by 1e-7 flips single decisions of the 3.2 M in the Bottleneck and its own upstream gradients move
                    by 6e-4 (`ref_self_noise` in the fixture), so no two implementations can agree to 1e-5 -- or, in bf16, where
                    ~0.25 % of the decisions sit inside the rounding noise, to 2e-2 -- unless the decisions are the same.
  raw gradients     upstream parameters vs the reference's gradients, reported together with the number of differing decisions
"""
import os

import numpy as np
import torch

from . import cases

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
# parameters whose true gradient is exactly zero (a conv bias feeding a train-mode BatchNorm): both sides hold rounding noise only
ZERO_NORM = 1e-3


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


class _GA:
    golden, cases_, tail = 'ga_convnext_parity.pt', cases.GA_PARITY_CASES, cases.GA_TAIL_PREFIXES

    def __init__(self):
        from . import ga_convnext_oracle as O
        self.O = O

    def spec(self, name):
        return self.O.SPECS[name]

    def model(self, name):
        import imagenet_models_b200.ga_convnext  # noqa: F401
        from imagenet_models_b200.registry import create_model
        return create_model(name)

    def labels(self, y, spec):
        return y

    def impl_loss(self, out, y):
        from imagenet_models_b200 import ops
        return ops.ga_loss(torch.stack(out), y, cases.GA_LAM)

    def flat(self, out):
        return list(out)

    def oracle_loss(self, out, y):
        return self.O.ga_loss(out, y, cases.GA_LAM)

    def masks(self, taps, B):
        bn = [t for kind, t in taps if kind == 'bn']
        se = [t for kind, t in taps if kind == 'se']
        assert len(bn) == 3 and len(se) == 1, (len(bn), len(se))
        m = {k: t.view(B, 14, 14, -1).permute(0, 3, 1, 2).cpu() for k, t in zip(('bn1', 'bn2', 'out'), bn)}
        m['se'] = se[0].view(B, -1, 1, 1).cpu()
        return m


class _MAP(_GA):
    golden, cases_, tail = 'map_convnext_parity.pt', cases.MAP_PARITY_CASES, cases.MAP_TAIL_PREFIXES

    def __init__(self):
        from . import map_convnext_oracle as O
        self.O = O

    def model(self, name):
        import imagenet_models_b200.map_convnext  # noqa: F401
        from imagenet_models_b200.registry import create_model
        m = create_model(name)
        m.head.drop = m.head.attn_drop = 0.0           # parity contract: every drop rate 0 (the fixture zeroes the reference's Dropouts)
        return m

    def impl_loss(self, out, y):
        from imagenet_models_b200 import ops
        return ops.ga_loss(torch.stack([o[0] for o in out]), y, cases.MAP_DEC_LAM, aux=torch.stack([o[1] for o in out]))

    def flat(self, out):
        return [t for pair in out for t in pair]

    def oracle_loss(self, out, y):
        return self.O.map_loss(out, y, cases.MAP_DEC_LAM)

    def masks(self, taps, B):
        mm = [t for kind, t in taps if kind == 'gemm']
        assert len(mm) == 4, len(mm)
        return {f'mlp{g}': t.view(B, -1, t.shape[1]).permute(0, 2, 1).unsqueeze(-1).cpu() for g, t in enumerate(mm)}


class _CSWIN(_GA):
    golden, cases_, tail = 'ga_cswin_parity.pt', cases.CSWIN_PARITY_CASES, cases.CSWIN_TAIL_PREFIXES

    def __init__(self):
        from . import ga_cswin_oracle as O
        self.O = O

    def model(self, name):
        import imagenet_models_b200.ga_cswin as GC
        s = self.O.SPECS[name]
        return GC.GA_CSWinTransformer(img_size=224, patch_size=4, num_classes=s.num_classes, embed_dim=s.embed_dim, depth=list(s.depth),
                                      split_size=list(s.split_size), num_heads=list(s.num_heads), dims=list(s.dims),
                                      stage3_naggre=s.naggre, gram_dim=s.gram_dim)

    def labels(self, y, spec):
        return y % spec.num_classes

    def masks(self, taps, B):
        assert not taps, 'GA-CSWin has no ReLU'
        return None


FAMILIES = {'ga': _GA, 'map': _MAP, 'cswin': _CSWIN}


def load_fixture(family):
    return torch.load(os.path.join(GOLDEN_DIR, FAMILIES[family].golden))


def measure(family, key, dtype, fixture=None):
    from imagenet_models_b200 import ops
    F_ = FAMILIES[family]()
    name, B, profile, kind = F_.cases_[key]
    g = (fixture or load_fixture(family))[key]
    spec = F_.spec(name)
    P = F_.O.make_state(spec, cases.STATE_SEED, profile=profile)
    x, y = cases.parity_inputs(kind, B)
    y = F_.labels(y, spec)
    m = F_.model(name).cuda()
    m.load_state_dict({k: v.cuda() for k, v in P.items()}, strict=True)
    m.compute_dtype = dtype
    m.train()
    ops.RELU_TAP = []
    try:
        out = m(x.cuda())
        taps = ops.RELU_TAP
    finally:
        ops.RELU_TAP = None
    loss = F_.impl_loss(out, y.cuda())
    loss.backward()
    torch.cuda.synchronize()
    res = {'family': family, 'key': key, 'dtype': str(dtype).split('.')[-1], 'B': B}
    res['logits'] = max(rel(a.detach().cpu(), b) for a, b in zip(F_.flat(out), g['train_logits']))
    res['loss'] = abs(loss.item() - g['loss'].item()) / abs(g['loss'].item())
    grads = {k: p.grad.detach().float().cpu() for k, p in m.named_parameters()}
    # ---- raw comparison with the reference's gradient digests
    raw = {k: (cases.digest_rel_err(grads[k], g['grads'][k]), g['grads'][k][0]) for k in grads}
    live = lambda d: {k: e for k, (e, n) in d.items() if n > ZERO_NORM}   # noqa: E731
    res['tail_grads'] = live({k: v for k, v in raw.items() if k.startswith(F_.tail)})
    res['upstream_grads_raw'] = live({k: v for k, v in raw.items() if not k.startswith(F_.tail)})
    # ---- ReLU decisions: ours vs the reference's
    masks = F_.masks(taps, B)
    flips = 0
    for kname, packed in g.get('relu_masks', {}).items():
        refm = torch.from_numpy(np.unpackbits(packed.numpy())[:masks[kname].numel()].astype(bool)).view(masks[kname].shape)
        flips += int((refm != masks[kname]).sum())
    res['relu_flips_vs_reference'] = flips
    res['relu_decisions'] = sum(packed.numel() * 8 for packed in g.get('relu_masks', {}).values())
    # ---- oracle at our decisions (CPU, all host threads)
    res['grads_pinned'] = {}
    res['logits_vs_pinned_oracle'] = None
    if masks is not None:
        Po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in P.items()}
        torch.set_num_threads(os.cpu_count())
        o_out = F_.O.forward(Po, spec, x, training=True, relu_masks=masks)
        F_.oracle_loss(o_out, y).backward()
        res['logits_vs_pinned_oracle'] = max(rel(a.detach().cpu(), b.detach()) for a, b in zip(F_.flat(out), F_.flat(o_out)))
        for k in grads:
            ref = Po[k].grad
            if ref is not None and ref.norm().item() > ZERO_NORM:
                res['grads_pinned'][k] = rel(grads[k], ref)
    else:                                   # no ReLU anywhere: the reference's gradients ARE the pinned ones
        res['grads_pinned'] = dict(res['tail_grads'])
    sd = m.state_dict()
    res['running'] = max([rel(sd[k].cpu(), v) for k, v in g['running'].items()] or [0.0])
    return res


def summarise(res):
    def stats(d):
        if not d:
            return None
        v = sorted(d.values())
        worst = max(d.items(), key=lambda kv: kv[1])
        return {'n': len(v), 'median': v[len(v) // 2], 'max': v[-1], 'argmax': worst[0]}
    return {'family': res['family'], 'key': res['key'], 'dtype': res['dtype'], 'B': res['B'], 'logits': res['logits'], 'loss': res['loss'],
            'logits_vs_pinned_oracle': res['logits_vs_pinned_oracle'], 'running_stats': res['running'],
            'tail_grads_vs_reference': stats(res['tail_grads']), 'all_grads_vs_pinned_oracle': stats(res['grads_pinned']),
            'upstream_grads_vs_reference_raw': stats(res['upstream_grads_raw']),
            'relu_flips_vs_reference': res['relu_flips_vs_reference'], 'relu_decisions': res['relu_decisions']}
