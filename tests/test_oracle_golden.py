"""CPU: the oracle restatement against the committed reference outputs (tests/golden/*.pt).

The fixtures were produced by tests/golden/make_golden.py from the unmodified reference modules.
"""
import os

import pytest
import torch

from oracle import cases
from oracle import ga_convnext_oracle as O


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope='module')
def gmod(golden_dir):
    return torch.load(os.path.join(golden_dir, 'ga_convnext_modules.pt'))


@pytest.fixture(scope='module')
def gmodel(golden_dir):
    return torch.load(os.path.join(golden_dir, 'ga_convnext_model.pt'))


def test_param_count_kats():
    # BASELINE.md section 2 (code-derived counts of the six registered variants)
    for name, n in cases.PARAM_COUNTS.items():
        shapes = O.state_shapes(O.SPECS[name])
        total = 0
        for key, (shape, kind) in shapes.items():
            if kind in ('w', 'b', 'g'):
                k = 1
                for s in shape:
                    k *= s
                total += k
        assert total == n, (name, total, n)


def test_state_key_count():
    # 407 state_dict entries for GA-T
    assert len(O.state_shapes(O.SPECS['ga_convnext_tiny_688'])) == 407


@pytest.mark.parametrize('cname', list(cases.BLOCK_CASES))
def test_block_vs_reference(gmod, cname):
    C, H, B = cases.BLOCK_CASES[cname]
    P = {k: v.requires_grad_(True) for k, v in cases.block_state(C).items()}
    x, dy = cases.block_inputs(C, H, B)
    x.requires_grad_(True)
    y = O.convnext_block(P, '', x)
    y.backward(dy)
    g = gmod[cname]
    assert rel(y.detach(), g['y']) < 1e-5
    assert rel(x.grad, g['dx']) < 1e-5
    for k, d in g['grads'].items():
        assert cases.digest_close(P[k].grad, d, 2e-5, 1e-6), k


@pytest.mark.parametrize('cname', list(cases.GRAM_CASES))
@pytest.mark.parametrize('training', [False, True])
def test_gram_vs_reference(gmod, cname, training):
    C, H, B = cases.GRAM_CASES[cname]
    out = O.gram_vector(cases.gram_input(C, H, B), training)
    assert rel(out, gmod[f'{cname}/train{int(training)}']) < 1e-6
    # unit L2 norm per image, upper-triangle length
    assert out.shape[1] == C * (C + 1) // 2
    assert torch.allclose(out.flatten(1).norm(dim=1), torch.ones(B), atol=1e-5)


@pytest.mark.parametrize('cname', list(cases.CLASSATTN_CASES))
def test_ga_block_vs_reference(gmod, cname):
    C, E, N, B = cases.CLASSATTN_CASES[cname]
    P = {'ga.0.' + k: v.requires_grad_(True) for k, v in cases.ga_block_state(C, E).items()}
    tokens, cls, dy = cases.ga_block_inputs(C, N, B)
    tokens.requires_grad_(True)
    cls.requires_grad_(True)
    y = O.ga_block(P, O.GASpec((1,), (C,), E, 0), 0, tokens, cls)
    y.backward(dy)
    g = gmod[cname]
    assert rel(y.detach(), g['y']) < 1e-5
    assert rel(tokens.grad, g['dtokens']) < 1e-5
    assert rel(cls.grad, g['dcls']) < 1e-5
    for k, d in g['grads'].items():
        assert cases.digest_close(P['ga.0.' + k].grad, d, 3e-5, 1e-6), k


def test_model_eval_vs_reference(gmodel):
    name, B = cases.GA_MODEL_CASES[0]
    spec = O.SPECS[name]
    P = O.make_state(spec, cases.STATE_SEED)
    x, _ = cases.ga_inputs(B)
    with torch.no_grad():
        out = O.forward(P, spec, x, training=False)
    for a, b in zip(out, gmodel[f'{name}/B{B}']['eval_logits']):
        assert rel(a, b) < 2e-5
        # bit-exact top-5 indices on (near-)identical logits
        assert torch.equal(a.topk(5).indices, b.topk(5).indices)


def test_channel_shuffle_is_transpose():
    t = torch.arange(24.).reshape(2, 12)
    s = O.channel_shuffle(t, 4)
    assert s[0].tolist() == [0, 4, 8, 1, 5, 9, 2, 6, 10, 3, 7, 11]
    # reference formula: view (C/g, g) -> permute -> flatten
    assert torch.equal(s, t.reshape(2, 3, 4).permute(0, 2, 1).reshape(2, 12))


def test_loss_matches_manual():
    g = torch.Generator().manual_seed(0)
    outs = [torch.randn(4, 10, generator=g) for _ in range(5)]
    y = torch.randint(0, 10, (4,), generator=g)
    base = sum(torch.nn.functional.cross_entropy(o, y) for o in outs)
    assert torch.allclose(O.ga_loss(outs, y, 0.0), base)
    assert not torch.allclose(O.ga_loss(outs, y, -0.8), base)


def test_create_model_loads_a_timm_format_checkpoint(tmp_path):
    """timm's CheckpointSaver writes an argparse.Namespace and optimizer state beside the weights (GA/train.py:649); such a
    file must load through create_model(checkpoint_path=...) with strict key matching, EMA weights preferred."""
    import argparse
    from imagenet_models_b200.registry import create_model
    import imagenet_models_b200.ga_convnext  # noqa: F401
    m = create_model('ga_convnext_tiny_688')
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    ema = {('module.' + k): (v + 1 if v.is_floating_point() else v.clone()) for k, v in sd.items()}     # DDP-prefixed, different values
    ck = {'epoch': 3, 'arch': 'ga_convnext_tiny_688', 'state_dict': sd, 'state_dict_ema': ema, 'version': 2,
          'args': argparse.Namespace(model='ga_convnext_tiny_688', lr=5e-3), 'optimizer': {'state': {}, 'param_groups': [{'lr': 5e-3}]},
          'metric': 83.2}
    path = tmp_path / 'checkpoint-3.pth.tar'
    torch.save(ck, path)
    m2 = create_model('ga_convnext_tiny_688', checkpoint_path=str(path))
    for k, v in m2.state_dict().items():
        assert torch.equal(v, ema['module.' + k]), k
