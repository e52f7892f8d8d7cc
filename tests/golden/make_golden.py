#!/usr/bin/env python3
"""Generate the committed golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference modules are imported from where they lie (never copied) through `oracle/timm_shim`.
For each case the script
  1. builds the reference module, loads `oracle.make_state(...)` with strict=True (pins key names/shapes),
  2. runs the reference forward / backward on seeded inputs,
  3. asserts the oracle restatement agrees (<= 2e-5 relative), and
  4. stores the REFERENCE outputs (not the oracle's) in tests/golden/*.pt.
Inputs and weights are regenerated from seeds on the test side, so fixtures hold outputs only.
"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'timm_shim'))
sys.path.insert(0, '/root/reference/GA')
sys.path.insert(0, '/root/reference/MAP')

from oracle import ga_convnext_oracle as O  # noqa: E402
from oracle import cases  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def grad_digest(named_grads):
    """Per-tensor (L2 norm, strided sample of <=2048 values): small enough to commit, sharp enough to catch errors."""
    return {k: cases.digest(g) for k, g in named_grads.items()}


def close(a, b, rtol, atol=3e-4):
    """||a-b|| <= rtol*||b|| + atol.  atol absorbs parameters whose true gradient is exactly zero
    (conv biases feeding a train-mode BatchNorm), where both sides only hold rounding noise."""
    return (a.double() - b.double()).norm().item() <= rtol * b.double().norm().item() + atol


def golden_ga_model():
    import ga_convnext as R
    import timm
    out = {}
    for name, B in cases.GA_MODEL_CASES:
        spec = O.SPECS[name]
        torch.manual_seed(0)
        ref = timm.create_model(name)
        P = O.make_state(spec, seed=cases.STATE_SEED)
        missing = ref.load_state_dict(P, strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        assert sum(p.numel() for p in ref.parameters()) == cases.PARAM_COUNTS[name]
        x, y = cases.ga_inputs(B)
        # ---- eval
        ref.eval()
        with torch.no_grad():
            r_eval = ref(x)
            o_eval = O.forward({k: v.clone() for k, v in P.items()}, spec, x, training=False)
        for a, b in zip(o_eval, r_eval):
            assert rel(a, b) < 2e-5, rel(a, b)
        # ---- train (drop rates 0): forward, GA loss (CE, lam=-0.8), backward
        ref.train()
        r_train = ref(x)
        # the exact expression of GA/train.py:735-745
        output, loss = 0, 0
        for o in r_train:
            loss = loss + F.cross_entropy(o, y)
            output = output + o.data
        for o in r_train:
            loss = loss + F.kl_div(F.log_softmax(o + 0), F.log_softmax((output.detach() / len(r_train)) + 0),
                                   reduction='mean', log_target=True) * cases.GA_LAM
        loss.backward()
        r_grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
        r_state = {k: v.detach().clone() for k, v in ref.state_dict().items() if 'running' in k}

        Po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
              for k, v in P.items()}
        o_train = O.forward(Po, spec, x, training=True)
        o_loss = O.ga_loss(o_train, y, cases.GA_LAM)
        o_loss.backward()
        assert rel(o_loss.detach(), loss.detach()) < 1e-6
        for a, b in zip(o_train, r_train):
            assert rel(a.detach(), b.detach()) < 2e-5, rel(a.detach(), b.detach())
        for k, g in r_grads.items():
            assert close(Po[k].grad, g, 5e-5), (k, rel(Po[k].grad, g))
        worst = max(rel(Po[k].grad, g) for k, g in r_grads.items() if g.norm() > 1e-2)
        for k, v in r_state.items():
            assert rel(Po[k], v) < 1e-5, k
        # the reference's OWN bf16-autocast deviation from its fp32 result on this fixture (CPU autocast): the
        # yardstick for the bf16 parity tolerance in train mode, where BatchNorm batch statistics amplify rounding
        self_err = {}
        for mode in ('eval', 'train'):
            m2 = timm.create_model(name)
            m2.load_state_dict(P, strict=True)
            m2.train(mode == 'train')
            with torch.no_grad(), torch.autocast('cpu', dtype=torch.bfloat16):
                o16 = m2(x)
            base = r_eval if mode == 'eval' else [t.detach() for t in r_train]
            self_err[mode] = max(rel(a.float(), b) for a, b in zip(o16, base))
        # per-parameter deviation of the reference's bf16-autocast GRADIENTS from its fp32 gradients (train mode)
        m3 = timm.create_model(name)
        m3.load_state_dict(P, strict=True)
        m3.train()
        with torch.autocast('cpu', dtype=torch.bfloat16):
            o3 = m3(x)
            l3 = sum(F.cross_entropy(o.float(), y) for o in o3)
            mean3 = sum(o.detach().float() for o in o3) / len(o3)
            for o in o3:
                l3 = l3 + F.kl_div(F.log_softmax(o.float(), -1), F.log_softmax(mean3, -1), reduction='mean', log_target=True) * cases.GA_LAM
        l3.backward()
        self_err['grads'] = {k: rel(p.grad.float(), r_grads[k]) for k, p in m3.named_parameters()}
        big = sorted(self_err['grads'].values())
        print(f'   reference bf16-autocast gradient self error: median {big[len(big)//2]:.2e}, max {big[-1]:.2e}')
        print(f'{name} B={B}: oracle==reference  (loss {loss.item():.6f}, worst grad rel {worst:.2e}, '
              f"reference bf16-autocast self error eval {self_err['eval']:.2e} train {self_err['train']:.2e})")
        out[f'{name}/B{B}'] = dict(ref_bf16_self_err=self_err,
            eval_logits=[t.clone() for t in r_eval], train_logits=[t.detach().clone() for t in r_train],
            loss=loss.detach().clone(), grads=grad_digest(r_grads), running=r_state)
    torch.save(out, os.path.join(HERE, 'ga_convnext_model.pt'))


def golden_ga_modules():
    """Small-shape module-level vectors from the reference classes (kernel-level parity fixtures)."""
    import ga_convnext as R
    out = {}
    for cname, (C, H, B) in cases.BLOCK_CASES.items():
        blk = R.ConvNeXtBlock(C, ls_init_value=1.0)
        P = cases.block_state(C, seed=cases.STATE_SEED)
        blk.load_state_dict(P, strict=True)
        x, dy = cases.block_inputs(C, H, B)
        x.requires_grad_(True)
        yv = blk(x)
        yv.backward(dy)
        grads = {k: p.grad.clone() for k, p in blk.named_parameters()}
        Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        xo = x.detach().clone().requires_grad_(True)
        yo = O.convnext_block(Po, '', xo)
        yo.backward(dy)
        assert rel(yo.detach(), yv.detach()) < 1e-5
        assert rel(xo.grad, x.grad) < 1e-5
        for k in grads:
            assert rel(Po[k].grad, grads[k]) < 1e-5, k
        out[cname] = dict(y=yv.detach().clone(), dx=x.grad.clone(), grads=grad_digest(grads))
        print(f'block {cname}: oracle==reference')

    # get_gram on a fake `self` (training / eval branches), ga_convnext.py:452-467
    class _G:
        pass
    for cname, (C, H, B) in cases.GRAM_CASES.items():
        x = cases.gram_input(C, H, B)
        for training in (False, True):
            g = _G()
            g.training = training
            import numpy as np
            idx = np.zeros(((C + 1) * C // 2))
            n = 0
            for i in range(C):
                for j in range(C):
                    if j >= i:
                        idx[n] = i * C + j
                        n += 1
            g.gram_index = idx
            r = R.GA_ConvNeXt.get_gram(g, x, B, C)
            o = O.gram_vector(x, training)
            assert rel(o, r) < 1e-6
            out[f'{cname}/train{int(training)}'] = r.clone()
        print(f'gram {cname}: oracle==reference')

    # ClassAttn + LayerScaleBlockClassAttn
    for cname, (C, E, N, B) in cases.CLASSATTN_CASES.items():
        m = R.LayerScaleBlockClassAttn(C, num_heads=8, mlp_block_groups=4, dim_embed=E)
        P = cases.ga_block_state(C, E, seed=cases.STATE_SEED)
        m.load_state_dict(P, strict=True)
        tokens, cls, dy = cases.ga_block_inputs(C, N, B)
        tokens.requires_grad_(True)
        cls.requires_grad_(True)
        r = m(tokens, cls[:, None, :])[:, 0]
        r.backward(dy)
        spec = O.GASpec((1,), (C,), E, 0)
        Po = {'ga.0.' + k: v.clone().requires_grad_(True) for k, v in P.items()}
        to, co = tokens.detach().clone().requires_grad_(True), cls.detach().clone().requires_grad_(True)
        o = O.ga_block(Po, spec, 0, to, co)
        o.backward(dy)
        assert rel(o.detach(), r.detach()) < 1e-5
        assert rel(to.grad, tokens.grad) < 1e-5 and rel(co.grad, cls.grad) < 1e-5
        grads = {k: p.grad.clone() for k, p in m.named_parameters()}
        for k in grads:
            assert rel(Po['ga.0.' + k].grad, grads[k]) < 2e-5, k
        out[cname] = dict(y=r.detach().clone(), dtokens=tokens.grad.clone(), dcls=cls.grad.clone(), grads=grad_digest(grads))
        print(f'ga_block {cname}: oracle==reference')
    torch.save(out, os.path.join(HERE, 'ga_convnext_modules.pt'))


def golden_map_model():
    """map_convnext_tiny through the unmodified reference (MAP/models/map_convnext.py + map.py)."""
    import types
    import timm
    import models.map_convnext  # noqa: F401  (registers map_convnext_* into the shim registry)
    from oracle import map_convnext_oracle as MO
    out = {}
    for name, B in cases.MAP_MODEL_CASES:
        spec = MO.SPECS[name]
        ref = timm.create_model(name)
        assert sum(p.numel() for p in ref.parameters()) == MO.PARAM_COUNTS[name]      # MAP/README.MD:308,373
        P = MO.make_state(spec, seed=cases.STATE_SEED)
        ref.load_state_dict(P, strict=True)
        for m in ref.modules():                    # parity contract: every drop rate 0 (map.py:149 hard-codes 0.05)
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        x, y = cases.ga_inputs(B)
        ref.eval()
        with torch.no_grad():
            r_eval = ref(x)
            o_eval = MO.forward({k: v.clone() for k, v in P.items()}, spec, x, training=False)
        for a, b in zip(o_eval, r_eval):
            assert rel(a, b) < 2e-5, rel(a, b)
        ref.train()
        r_train = ref(x)
        args = types.SimpleNamespace(distill_tokens=0, token_distillation=False, dec_lam=cases.MAP_DEC_LAM)
        sys.path.insert(0, '/root/reference/MAP')
        # the exact expression of MAP/train.py:792-839 (copied call, not copied code: import would pull timm.data)
        loss = 0
        agg = 0
        for yh, ym in r_train:
            agg = agg + yh
            loss = loss + F.cross_entropy(yh, y) + F.kl_div(F.log_softmax(ym, dim=1), F.log_softmax(yh, dim=1).detach(),
                                                            reduction='sum', log_target=True) / yh.numel()
        for yh, ym in r_train:
            loss = loss + F.kl_div(F.log_softmax(yh, dim=1), F.log_softmax(agg.detach() / len(r_train), dim=1),
                                   reduction='mean', log_target=True) * args.dec_lam
        loss.backward()
        r_grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None}
        r_state = {k: v.detach().clone() for k, v in ref.state_dict().items() if 'running' in k}
        Po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
              for k, v in P.items()}
        o_train = MO.forward(Po, spec, x, training=True)
        o_loss = MO.map_loss(o_train, y, cases.MAP_DEC_LAM)
        o_loss.backward()
        assert rel(o_loss.detach(), loss.detach()) < 1e-6
        for (a1, a2), (b1, b2) in zip(o_train, r_train):
            assert rel(a1.detach(), b1.detach()) < 2e-5 and rel(a2.detach(), b2.detach()) < 2e-5
        for k, g in r_grads.items():
            assert close(Po[k].grad, g, 5e-5), (k, rel(Po[k].grad, g))
        for k, v in r_state.items():
            assert rel(Po[k], v) < 1e-5, k
        self_err = {}
        for mode in ('eval', 'train'):
            m2 = timm.create_model(name)
            m2.load_state_dict(P, strict=True)
            for m in m2.modules():
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
            m2.train(mode == 'train')
            with torch.no_grad(), torch.autocast('cpu', dtype=torch.bfloat16):
                o16 = m2(x)
            if mode == 'eval':
                self_err[mode] = max(rel(a.float(), b) for a, b in zip(o16, r_eval))
            else:
                self_err[mode] = max(rel(a[0].float(), b[0].detach()) for a, b in zip(o16, r_train))
        print(f'{name} B={B}: oracle==reference  (loss {loss.item():.6f}; reference bf16-autocast self error '
              f"eval {self_err['eval']:.2e} train {self_err['train']:.2e})")
        out[f'{name}/B{B}'] = dict(ref_bf16_self_err=self_err, eval_logits=[t.clone() for t in r_eval],
                                   train_logits=[[a.detach().clone(), b.detach().clone()] for a, b in r_train],
                                   loss=loss.detach().clone(), grads=grad_digest(r_grads), running=r_state)
    torch.save(out, os.path.join(HERE, 'map_convnext_model.pt'))


def golden_parity(family):
    """Whole-model training-step fixtures at the BASELINE.json shapes (oracle/cases.py *_PARITY_CASES): GA-ConvNeXt config 1
    (fp32, batch 8, fp64 Gram branch) and its batch-16 bf16 fixture, MAP-ConvNeXt-T and GA-CSWin-T at batch 8; mutually
    different images, trained-magnitude residual-branch scales.  Stores the reference's logits, loss, gradient digests,
    BatchNorm statistics, ReLU decisions, and its own noise floor (1e-7 input perturbation) and bf16-autocast deviation."""
    import numpy as np
    import timm
    from oracle import parity_check as PC
    fam = PC.FAMILIES[family]()
    if family == 'ga':
        import ga_convnext  # noqa: F401
    elif family == 'map':
        import models.map_convnext  # noqa: F401

    def build(name):
        if family == 'cswin':
            import ga_cswin as R
            s_ = fam.spec(name)
            return R.GA_CSWinTransformer(img_size=224, patch_size=4, num_classes=s_.num_classes, embed_dim=s_.embed_dim,
                                         depth=list(s_.depth), split_size=list(s_.split_size), num_heads=list(s_.num_heads),
                                         dims=list(s_.dims), stage3_naggre=s_.naggre, gram_dim=s_.gram_dim)
        ref = timm.create_model(name)
        if family == 'map':
            for m in ref.modules():                # parity contract: every drop rate 0 (map.py:149 hard-codes 0.05)
                if isinstance(m, torch.nn.Dropout):
                    m.p = 0.0
        return ref

    def relu_modules(ref):
        if family == 'ga':
            return {'bn1': ref.stages[4].act1, 'bn2': ref.stages[4].act2, 'out': ref.stages[4].act3}
        if family == 'map':
            return {f'mlp{g}': ref.head.mmcap.mmcap[g].attention[0].mlp.act for g in range(len(ref.head.mmcap.mmcap))}
        return {}

    def ref_loss(outs, y):
        if family == 'map':                        # the expression of MAP/train.py:792-839 (distill_tokens == 0)
            loss, agg = 0, 0
            for yh, ym in outs:
                agg = agg + yh
                loss = loss + F.cross_entropy(yh, y) + F.kl_div(F.log_softmax(ym, dim=1), F.log_softmax(yh, dim=1).detach(),
                                                                reduction='sum', log_target=True) / yh.numel()
            for yh, ym in outs:
                loss = loss + F.kl_div(F.log_softmax(yh, dim=1), F.log_softmax(agg.detach() / len(outs), dim=1),
                                       reduction='mean', log_target=True) * cases.MAP_DEC_LAM
            return loss
        output, loss = 0, 0                        # the loss expression of GA/train.py:735-745
        for o in outs:
            loss = loss + F.cross_entropy(o, y)
            output = output + o.data
        for o in outs:
            loss = loss + F.kl_div(F.log_softmax(o + 0), F.log_softmax((output.detach() / len(outs)) + 0),
                                   reduction='mean', log_target=True) * cases.GA_LAM
        return loss

    def run_ref(name, P, x, y):
        ref = build(name)
        ref.load_state_dict(P, strict=True)
        ref.train()
        masks = {}
        for key, act in relu_modules(ref).items():
            act.register_forward_hook(lambda m, i, o, key=key: masks.__setitem__(key, (o.detach() > 0)))
        outs = ref(x)
        loss = ref_loss(outs, y)
        loss.backward()
        grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters() if p.grad is not None}
        return ref, outs, loss, grads, dict(masks)   # copy: later forwards re-fire the hooks

    out = {}
    for key, (name, B, profile, kind) in fam.cases_.items():
        spec = fam.spec(name)
        P = fam.O.make_state(spec, seed=cases.STATE_SEED, profile=profile)
        x, y = cases.parity_inputs(kind, B)
        y = fam.labels(y, spec)
        ref, r_train, loss, r_grads, r_masks = run_ref(name, P, x, y)
        r_flat = [t.detach() for t in fam.flat(r_train)]
        r_state = {k: v.detach().clone() for k, v in ref.state_dict().items() if 'running' in k}
        ref.eval()
        ref.load_state_dict(P, strict=True)
        with torch.no_grad():
            r_eval = ref(x)
        # oracle == reference, with the oracle evaluated at the reference's ReLU decisions (they agree anyway unless a
        # pre-activation sits within fp32 rounding of zero)
        Po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
              for k, v in P.items()}
        kw = {'relu_masks': r_masks} if r_masks else {}
        o_train = fam.O.forward(Po, spec, x, training=True, **kw)
        o_loss = fam.oracle_loss(o_train, y)
        o_loss.backward()
        assert rel(o_loss.detach(), loss.detach()) < 1e-6
        for a, b in zip(fam.flat(o_train), r_flat):
            assert rel(a.detach(), b) < 2e-5, rel(a.detach(), b)
        worst = 0.0
        for k, g in r_grads.items():
            assert close(Po[k].grad, g, 5e-5), (k, rel(Po[k].grad, g))
            if g.norm() > 1e-3:
                worst = max(worst, rel(Po[k].grad, g))
        # the reference against ITSELF on inputs perturbed by 1e-7 relative: the noise floor of any fp32 comparison here,
        # and how many ReLU decisions flip
        gen = torch.Generator().manual_seed(5)
        _, p_train, _, p_grads, p_masks = run_ref(name, P, x * (1 + 1e-7 * torch.randn(x.shape, generator=gen)), y)
        noise_g = sorted(rel(p_grads[k], g) for k, g in r_grads.items() if g.norm() > 1e-3)
        flips = sum(int((p_masks[k] != r_masks[k]).sum()) for k in r_masks)
        self_noise = dict(logits=max(rel(a.detach(), b) for a, b in zip(fam.flat(p_train), r_flat)),
                          grads_median=noise_g[len(noise_g) // 2], grads_max=noise_g[-1], relu_flips=flips,
                          relu_decisions=sum(m.numel() for m in r_masks.values()))
        # information only: how far the reference's OWN bf16 autocast (CPU) is from its fp32 result on this fixture
        m2 = build(name)
        m2.load_state_dict(P, strict=True)
        m2.train()
        with torch.no_grad(), torch.autocast('cpu', dtype=torch.bfloat16):
            o16 = m2(x)
        self_err = max(rel(a.float(), b) for a, b in zip(fam.flat(o16), r_flat))
        print(f'{family}/{key}: {name} B={B} {profile}/{kind}: oracle==reference (loss {loss.item():.6f}, worst gradient {worst:.2e}); '
              f'reference vs itself at 1e-7 input noise: logits {self_noise["logits"]:.2e}, gradients median '
              f'{self_noise["grads_median"]:.2e} max {self_noise["grads_max"]:.2e}, {flips} of {self_noise["relu_decisions"]} ReLU '
              f'decisions flipped; reference bf16-autocast train logits are {self_err:.2e} from its fp32 logits')
        out[key] = dict(eval_logits=[t.clone() for t in r_eval], train_logits=[t.clone() for t in r_flat],
                        loss=loss.detach().clone(), grads=grad_digest(r_grads), running=r_state,
                        relu_masks={k: torch.from_numpy(np.packbits(m.numpy().reshape(-1))) for k, m in r_masks.items()},
                        ref_bf16_train_self_err=self_err, ref_self_noise=self_noise)
    torch.save(out, os.path.join(HERE, fam.golden))


def golden_ga_384():
    """BASELINE config 5 runs GA-ConvNeXt-B at 384x384 as well.  The reference hard-codes AdaptiveAvgPool2d(14) (ga_convnext.py:397,
    = H/16 at 224) and cannot run other sizes; with that one module ATTRIBUTE replaced by AdaptiveAvgPool2d(H/16) on the
    otherwise unmodified reference instance it does.  Eval logits of ga_convnext_base_976 and ga_convnext_tiny_688 at 384x384."""
    import ga_convnext  # noqa: F401
    import timm
    out = {}
    for name, B in (('ga_convnext_base_976', 1), ('ga_convnext_tiny_688', 2)):
        spec = O.SPECS[name]
        ref = timm.create_model(name)
        P = O.make_state(spec, seed=cases.STATE_SEED, profile='trained')
        ref.load_state_dict(P, strict=True)
        ref.avg_pool = torch.nn.AdaptiveAvgPool2d(384 // 16)
        ref.eval()
        x, _ = cases.ga_inputs_diverse(B, size=384)
        with torch.no_grad():
            r = ref(x)
            o = O.forward({k: v.clone() for k, v in P.items()}, spec, x, training=False)
        for a, b in zip(o, r):
            assert rel(a, b) < 2e-5, rel(a, b)
        print(f'{name} at 384x384 (pool target 24): oracle==reference')
        out[f'{name}/B{B}/384'] = [t.clone() for t in r]
    torch.save(out, os.path.join(HERE, 'ga_convnext_384.pt'))


def golden_cswin():
    """GA-CSWin (GA/ga_cswin.py) through the unmodified reference: CSWinBlock cases + two whole-model cases."""
    import ga_cswin as R
    from oracle import ga_cswin_oracle as CO
    out = {}
    # ---- CSWinBlock / LePEAttention (rows a15, a16): every (tokens/stripe, heads) geometry of the T configuration
    for cname, (dim, reso, split, heads, last, B) in cases.CSWIN_BLOCK_CASES.items():
        blk = R.CSWinBlock(dim=dim, reso=reso, num_heads=heads, split_size=split, qkv_bias=True, last_stage=last)
        S = {}
        CO.block_shapes(S, '', dim, reso, split, last)
        P = cases._fill(S, cases.STATE_SEED)
        blk.load_state_dict(P, strict=True)
        x, dy = cases.cswin_block_inputs(dim, reso, B)
        x.requires_grad_(True)
        yv = blk(x)
        yv.backward(dy)
        grads = {k: p.grad.clone() for k, p in blk.named_parameters()}
        Po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        xo = x.detach().clone().requires_grad_(True)
        yo = CO.cswin_block(Po, '', xo, reso, split, heads, last)
        yo.backward(dy)
        assert rel(yo.detach(), yv.detach()) < 1e-5, rel(yo.detach(), yv.detach())
        assert rel(xo.grad, x.grad) < 1e-5
        for k in grads:
            assert close(Po[k].grad, grads[k], 2e-5), (k, rel(Po[k].grad, grads[k]))
        out['block/' + cname] = dict(y=cases.digest(yv.detach()), dx=cases.digest(x.grad), grads=grad_digest(grads))
        print(f'cswin block {cname}: oracle==reference')
    # ---- whole models
    for name, B in cases.CSWIN_MODEL_CASES:
        spec = CO.SPECS[name]
        torch.manual_seed(0)
        ref = R.GA_CSWinTransformer(img_size=spec.img_size, patch_size=4, num_classes=spec.num_classes, embed_dim=spec.embed_dim,
                                    depth=list(spec.depth), split_size=list(spec.split_size), num_heads=list(spec.num_heads),
                                    dims=list(spec.dims), stage3_naggre=spec.naggre, gram_dim=spec.gram_dim)
        P = CO.make_state(spec, seed=cases.STATE_SEED)
        res = ref.load_state_dict(P, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        if name in CO.PARAM_COUNTS:
            assert sum(p.numel() for p in ref.parameters()) == CO.PARAM_COUNTS[name], sum(p.numel() for p in ref.parameters())
        x, y = cases.ga_inputs(B)
        y = y % spec.num_classes
        ref.eval()
        with torch.no_grad():
            r_eval = ref(x)
            o_eval = CO.forward({k: v.clone() for k, v in P.items()}, spec, x, training=False)
        for a, b in zip(o_eval, r_eval):
            assert rel(a, b) < 2e-5, rel(a, b)
        ref.train()
        r_train = ref(x)
        output, loss = 0, 0           # the loss expression of GA/train.py:735-745
        for o in r_train:
            loss = loss + F.cross_entropy(o, y)
            output = output + o.data
        for o in r_train:
            loss = loss + F.kl_div(F.log_softmax(o + 0), F.log_softmax((output.detach() / len(r_train)) + 0),
                                   reduction='mean', log_target=True) * cases.GA_LAM
        loss.backward()
        r_grads = {k: p.grad.detach().clone() for k, p in ref.named_parameters()}
        r_state = {k: v.detach().clone() for k, v in ref.state_dict().items() if 'running' in k}
        Po = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone())
              for k, v in P.items()}
        o_train = CO.forward(Po, spec, x, training=True)
        o_loss = CO.ga_loss(o_train, y, cases.GA_LAM)
        o_loss.backward()
        assert rel(o_loss.detach(), loss.detach()) < 1e-6
        for a, b in zip(o_train, r_train):
            assert rel(a.detach(), b.detach()) < 2e-5, rel(a.detach(), b.detach())
        for k, g in r_grads.items():
            assert close(Po[k].grad, g, 5e-5), (k, rel(Po[k].grad, g))
        for k, v in r_state.items():
            assert rel(Po[k], v) < 1e-5, k
        self_err = {}
        for mode in ('eval', 'train'):
            ref.train(mode == 'train')
            ref.load_state_dict(P, strict=True)
            with torch.no_grad(), torch.autocast('cpu', dtype=torch.bfloat16):
                o16 = ref(x)
            base = r_eval if mode == 'eval' else [t.detach() for t in r_train]
            self_err[mode] = max(rel(a.float(), b) for a, b in zip(o16, base))
        print(f'{name} B={B}: oracle==reference  (loss {loss.item():.6f}; reference bf16-autocast self error '
              f"eval {self_err['eval']:.2e} train {self_err['train']:.2e})")
        out[f'{name}/B{B}'] = dict(ref_bf16_self_err=self_err, eval_logits=[t.clone() for t in r_eval],
                                   train_logits=[t.detach().clone() for t in r_train], loss=loss.detach().clone(),
                                   grads=grad_digest(r_grads), running=r_state)
    torch.save(out, os.path.join(HERE, 'ga_cswin.pt'))


if __name__ == '__main__':
    torch.set_num_threads(8)
    which = sys.argv[1:] or ['modules', 'model']
    if 'modules' in which:
        golden_ga_modules()
    if 'model' in which:
        golden_ga_model()
    if 'map' in which or 'model' in which:
        golden_map_model()
    if 'cswin' in which:
        golden_cswin()
    if 'ga384' in which:
        golden_ga_384()
    for fam in ('ga', 'map', 'cswin'):
        if f'parity_{fam}' in which or 'parity' in which:
            golden_parity(fam)
