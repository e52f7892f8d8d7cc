"""GA_ConvNeXt drop-in module: structure (CPU) and whole-model parity against the reference fixtures (GPU)."""
import os

import pytest
import torch

from imagenet_models_b200 import ga_convnext as M
from imagenet_models_b200 import lib as L
from imagenet_models_b200 import ops
from imagenet_models_b200.registry import create_model, list_models
from oracle import cases
from oracle import ga_convnext_oracle as O


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------- CPU: structure
def test_registry_names():
    assert set(cases.PARAM_COUNTS) <= set(list_models())    # the six factories of ga_convnext.py:572-613
    with pytest.raises(RuntimeError):
        create_model('ga_convnext_tiny')                    # README name is not a registered model


@pytest.mark.parametrize('name', ['ga_convnext_tiny_688', 'ga_convnext_small_768', 'ga_convnext_base_976'])
def test_state_dict_contract(name):
    m = create_model(name, drop_path_rate=None, global_pool=None)   # None-valued kwargs are dropped like timm does
    sd = m.state_dict()
    want = O.state_shapes(O.SPECS[name])
    assert sorted(sd.keys()) == sorted(want.keys())      # load_state_dict matches by key, not order
    for k, (shape, _) in want.items():
        assert tuple(sd[k].shape) == shape, k
    assert sum(p.numel() for p in m.parameters()) == cases.PARAM_COUNTS[name]
    m.load_state_dict(O.make_state(O.SPECS[name], 3), strict=True)
    assert m.num_classes == 1000 and m.default_cfg['input_size'] == (3, 224, 224)


def test_init_matches_reference_bitwise():
    """Same seed -> same initial weights as the reference constructor (needs /root/reference; build box only)."""
    if not os.path.isdir('/root/reference/GA'):
        pytest.skip('reference sources not present on this box')
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, 'oracle', 'timm_shim'))
    sys.path.insert(0, '/root/reference/GA')
    import timm
    import ga_convnext  # noqa: F401  (registers into the shim)
    torch.manual_seed(123)
    ref = timm.create_model('ga_convnext_tiny_688', drop_path_rate=0.1)
    torch.manual_seed(123)
    mine = create_model('ga_convnext_tiny_688', drop_path_rate=0.1)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    for k in rs:
        assert torch.equal(rs[k], ms[k]), k
    # drop-path schedule (ga_convnext.py:362, 413, 376)
    rates = [b.drop_prob for s in mine.stages[:4] for b in s.blocks]
    want = torch.linspace(0, 0.1, 19).tolist()[:18]
    assert rates == pytest.approx(want)
    assert mine.gram_layer[0].blocks[0].drop_prob == pytest.approx(0.1)
    assert mine.stages[4].drop_prob == pytest.approx(0.1)


def test_deepcopy_and_cpu_forward_raises():
    import copy
    m = create_model('ga_convnext_tiny_688')
    m2 = copy.deepcopy(m)                       # ModelEmaV2 does this (GA/train.py:499)
    assert list(m2.state_dict().keys()) == list(m.state_dict().keys())
    with pytest.raises(L.GaError):
        m(torch.zeros(1, 3, 224, 224))          # no CPU fallback


def test_library_exports_every_declared_symbol():
    import ctypes
    assert os.path.exists(L.LIB_PATH), 'run __graft_entry__.build() first'
    lib = ctypes.CDLL(L.LIB_PATH)
    names = L.exported_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), n
    assert lib.ga_version() >= 100


# ------------------------------------------------------------------------------------------------- GPU: parity
@pytest.fixture(scope='module')
def gmodel(golden_dir):
    return torch.load(os.path.join(golden_dir, 'ga_convnext_model.pt'))


def _build(name, dtype):
    m = create_model(name).cuda()
    m.load_state_dict({k: v.cuda() for k, v in O.make_state(O.SPECS[name], cases.STATE_SEED).items()}, strict=True)
    m.compute_dtype = dtype
    return m


@pytest.mark.gpu
@pytest.mark.parametrize('dtype,tol', [(torch.float32, 2e-5), (torch.bfloat16, 2e-2)])
def test_eval_logits_vs_reference(gmodel, dtype, tol):
    name, B = cases.GA_MODEL_CASES[0]
    g = gmodel[f'{name}/B{B}']
    m = _build(name, dtype).eval()
    x, _ = cases.ga_inputs(B)
    with torch.no_grad():
        out = m(x.cuda())
        out_cl = m(x.cuda().contiguous(memory_format=torch.channels_last))
    assert len(out) == 5
    for a, b, c in zip(out, g['eval_logits'], out_cl):
        assert a.shape == (B, 1000) and a.dtype == torch.float32
        assert rel(a.cpu(), b) < tol, rel(a.cpu(), b)
        assert torch.equal(a, c)                       # channels_last input storage gives identical results
        if dtype == torch.float32:
            assert torch.equal(a.cpu().topk(5).indices, b.topk(5).indices)   # bit-exact top-5 on identical logits


@pytest.mark.gpu
def test_train_step_vs_reference_fp32_batch2(gmodel):
    """forward + GA loss (CE, lam=-0.8) + backward against the reference's logits, loss, gradients, BN statistics, fp32, on the
    batch-2 fixture.  Train-mode BatchNorm over a batch of 2 is a sign function (the gram_embedding BatchNorm sees [2, C, 1, 1]),
    so this fixture only serves the fp32 path; the bf16 training contract (2e-2 on logits and every gradient, asserted directly)
    is tests/test_parity_baseline_shapes.py on the batch-8 / batch-16 fixtures."""
    from imagenet_models_b200 import ops
    name, B = cases.GA_MODEL_CASES[0]
    g = gmodel[f'{name}/B{B}']
    m = _build(name, torch.float32).train()
    x, y = cases.ga_inputs(B)
    out = m(x.cuda())
    errs = [rel(a.detach().cpu(), b) for a, b in zip(out, g['train_logits'])]
    print('fp32 train logits vs reference:', errs)
    assert max(errs) < 5e-5, errs
    loss = ops.ga_loss(torch.stack(out), y.cuda(), cases.GA_LAM)
    assert abs(loss.item() - g['loss'].item()) < 1e-4 * abs(g['loss'].item())
    loss.backward()
    worst = []
    for k, p in m.named_parameters():
        assert p.grad is not None, k
        if not cases.digest_close(p.grad, g['grads'][k], 5e-5, 3e-4):
            worst.append((k, g['grads'][k][0], p.grad.double().norm().item()))
    assert not worst, worst[:10]
    sd = m.state_dict()
    for k, v in g['running'].items():
        assert rel(sd[k].cpu(), v) < 1e-5, k
    assert int(sd['stages.4.bn1.num_batches_tracked']) == 1


@pytest.mark.gpu
def test_autocast_selects_bf16_and_grads_are_fp32():
    name, B = cases.GA_MODEL_CASES[0]
    m = _build(name, None).train()
    x, _ = cases.ga_inputs(B)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        out = m(x.cuda())
    assert out[0].dtype == torch.float32
    sum(o.float().sum() for o in out).backward()
    assert all(p.grad.dtype == torch.float32 for p in m.parameters())
    assert ops.LAST_GEMM_BACKEND in (L.BACKEND_SIMT, L.BACKEND_TCGEN05)


@pytest.mark.gpu
@pytest.mark.parametrize('name,size', [('ga_convnext_small_768', 224), ('ga_convnext_base_976', 224), ('ga_convnext_tiny_688', 288)])
def test_other_variants_eval_vs_oracle(name, size):
    """The committed reference fixtures cover tiny_688 at 224 (and 384, tests/test_parity_baseline_shapes.py); the deeper / wider
    factories (27-block stage, 4 taps, C=976) and a 288 input (pool target H/16 = 18) are checked against the pinned oracle
    directly, fp32 and bf16."""
    spec = O.SPECS[name]
    x, _ = cases.ga_inputs(1, size=size)
    with torch.no_grad():
        ref = O.forward(O.make_state(spec, cases.STATE_SEED), spec, x, training=False)
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
        m = _build(name, dtype).eval()
        with torch.no_grad():
            out = m(x.cuda())
        assert len(out) == 5 and all(o.shape == (1, 1000) and torch.isfinite(o).all() for o in out)
        for a, b in zip(out, ref):
            assert rel(a.cpu(), b) < tol, (name, dtype, rel(a.cpu(), b))
