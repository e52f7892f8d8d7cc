"""TrainEngine on the GPU: the CUDA-graph step must train exactly like the eager step (same losses, same weights, EMA and
BatchNorm statistics), and the adopted-gradient gather must equal per-parameter accumulation."""
import copy

import pytest
import torch

from oracle import cases
from oracle import ga_cswin_oracle as CO


def _model():
    import imagenet_models_b200.ga_cswin as GC
    spec = CO.SPECS['ga_cswin_test']
    torch.manual_seed(3)
    m = GC.GA_CSWinTransformer(img_size=224, patch_size=4, num_classes=spec.num_classes, embed_dim=spec.embed_dim,
                               depth=list(spec.depth), split_size=list(spec.split_size), num_heads=list(spec.num_heads),
                               dims=list(spec.dims), stage3_naggre=spec.naggre, gram_dim=spec.gram_dim).cuda()
    m.load_state_dict({k: v.cuda() for k, v in CO.make_state(spec, cases.STATE_SEED).items()}, strict=True)
    return m.train(), spec


@pytest.mark.gpu
def test_graph_step_matches_eager_step():
    """AdamW turns rounding-level gradient noise into +-lr updates, so trajectories of two runs drift apart chaotically;
    the comparison is therefore made one step at a time from bit-identical weights (warm-up steps run with lr = 0)."""
    from imagenet_models_b200.engine import TrainEngine
    m0, spec = _model()
    m1 = copy.deepcopy(m0)
    e0 = TrainEngine(m0, lr=0.0, weight_decay=0.05, ema_decay=0.99, cuda_graph=False)
    e1 = TrainEngine(m1, lr=0.0, weight_decay=0.05, ema_decay=0.99, cuda_graph=True, graph_warmup=2)
    g = torch.Generator().manual_seed(0)

    def batch():
        return (torch.randn(4, 3, 224, 224, generator=g).cuda(), torch.randint(0, spec.num_classes, (4,), generator=g).cuda())
    for _ in range(2):                                   # eager in both engines; lr = 0 keeps the weights bit-identical
        x, y = batch()
        e0.step(x, y)
        e1.step(x, y)
    for p0, p1 in zip(m0.parameters(), m1.parameters()):
        assert torch.equal(p0, p1)
    # ---- first captured step, lr = 1e-3
    e0.opt.param_groups[0]['lr'] = e1.opt.param_groups[0]['lr'] = 1e-3
    x, y = batch()
    l0, l1 = e0.step(x, y).item(), e1.step(x, y).item()
    assert e1._graph is not None and e1.graph_launches > 100
    assert abs(l0 - l1) <= 1e-4 * abs(l0), (l0, l1)
    g0, g1 = e0.opt.state.grad, e1.opt.state.grad
    # run-to-run noise of the bf16 chain (atomic accumulation order flips occasional bf16 roundings) is ~3e-3 at batch 4;
    # a stale pointer or a missed kernel in the captured graph would be O(1)
    assert (g0 - g1).norm().item() <= 1e-2 * g0.norm().item()
    for p0, p1 in zip(m0.parameters(), m1.parameters()):
        assert (p0 - p1).abs().max().item() <= 2.1e-3 + 1e-6           # at most +-lr each, whatever the noise did
    before = [p.detach().clone() for p in m1.parameters()]
    # ---- second replay with lr = 0: the captured optimizer must read the new lr from device memory -> weights frozen
    e0.opt.param_groups[0]['lr'] = e1.opt.param_groups[0]['lr'] = 0.0
    x, y = batch()
    l0, l1 = e0.step(x, y).item(), e1.step(x, y).item()
    assert abs(l0 - l1) <= 3e-2 * abs(l0), (l0, l1)
    for b, p in zip(before, m1.parameters()):
        assert torch.equal(b, p)
    # ---- and it did train at lr = 1e-3: weights differ from the initial state
    init = CO.make_state(spec, cases.STATE_SEED)
    assert (m1.state_dict()['fc.0.weight'].cpu() - init['fc.0.weight']).abs().max().item() > 1e-4
    assert e0.opt.step_count == e1.opt.step_count == 4
    for b0, b1 in zip(m0.buffers(), m1.buffers()):                      # BatchNorm statistics advanced identically
        if b0.is_floating_point():
            assert (b0 - b1).norm().item() <= 1e-3 * b0.norm().item() + 1e-6
        else:
            assert torch.equal(b0, b1)


@pytest.mark.gpu
def test_gather_equals_parameter_gradients():
    from imagenet_models_b200 import ops
    from imagenet_models_b200.optim import FusedAdamWEma
    m, spec = _model()
    opt = FusedAdamWEma(m, lr=1e-3, ema_decay=None)
    x = torch.randn(2, 3, 224, 224, device='cuda')
    y = torch.randint(0, spec.num_classes, (2,), device='cuda')
    opt.zero_grad()
    assert all(p.grad is None for p in m.parameters())
    ops.ga_loss(torch.stack(m(x)), y, -0.8).backward()
    opt.state.grad.fill_(float('nan'))
    opt.state.gather()
    torch.cuda.synchronize()
    for p, o in zip(opt.state.params, opt.state.offsets):
        assert torch.equal(opt.state.grad[o:o + p.numel()], p.grad.reshape(-1))


@pytest.mark.gpu
def test_fused_lamb_matches_the_restated_timm_algorithm():
    """ga_lamb_ema (3 launches over the flat state) vs oracle/lamb_oracle.py over 4 steps, incl. the global-norm clip,
    the no-decay group (biases: no trust ratio) and the EMA."""
    import torch.nn as nn
    from imagenet_models_b200.optim import FusedLambEma
    from oracle.lamb_oracle import lamb_step
    torch.manual_seed(0)
    net = nn.Sequential(nn.Linear(37, 64), nn.GELU(), nn.Linear(64, 5000), nn.LayerNorm(5000), nn.Linear(5000, 3)).cuda()
    ref_p = [p.detach().cpu().clone() for p in net.parameters()]
    names = [n for n, _ in net.named_parameters()]
    decay = [not (p.ndim <= 1 or n.endswith('.bias')) for n, p in zip(names, ref_p)]
    m = [torch.zeros_like(p) for p in ref_p]
    v = [torch.zeros_like(p) for p in ref_p]
    ema = [p.clone() for p in ref_p]
    opt = FusedLambEma(net, lr=5e-3, weight_decay=0.05, ema_decay=0.9, max_grad_norm=1.0)
    g = torch.Generator().manual_seed(1)
    for step in range(1, 5):
        scale = 3.0 if step == 1 else 0.01            # step 1 exceeds max_grad_norm (clip active), later steps do not
        grads = [torch.randn(p.shape, generator=g) * scale for p in ref_p]
        opt.zero_grad()
        for p, gr in zip(net.parameters(), grads):
            p.grad = gr.cuda()
        opt.step()
        lamb_step(ref_p, grads, m, v, decay, step, lr=5e-3, eps=1e-6, weight_decay=0.05, max_grad_norm=1.0)
        for e, p in zip(ema, ref_p):
            e.mul_(0.9).add_(p, alpha=0.1)
    for n, p, r in zip(names, net.parameters(), ref_p):
        assert (p.detach().cpu() - r).norm().item() <= 2e-5 * r.norm().item() + 1e-7, n
    for n, p, r in zip(names, opt.ema_model.parameters(), ema):
        assert (p.detach().cpu() - r).norm().item() <= 2e-5 * r.norm().item() + 1e-7, n
    assert opt.step_count == 4


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['ga_convnext_tiny_768', 'ga_convnext_small_688', 'ga_convnext_small_768', 'ga_convnext_base_976',
                                  'ga_convnext_base_1024', 'map_convnext_small', 'ga_CSWin_64_12211_tiny_224'])
def test_every_other_factory_trains(name):
    """Shape coverage beyond the fixtures: every registered factory runs two bf16 training steps (all gradients finite and
    non-zero somewhere, weights move) and an fp32 eval forward."""
    import imagenet_models_b200.ga_convnext  # noqa: F401
    import imagenet_models_b200.ga_cswin  # noqa: F401
    import imagenet_models_b200.map_convnext  # noqa: F401
    from imagenet_models_b200.engine import TrainEngine, evaluate_batch
    from imagenet_models_b200.registry import create_model
    torch.manual_seed(0)
    m = create_model(name).cuda().train()
    w0 = m.fc[0].weight.detach().clone() if hasattr(m, 'fc') else next(m.parameters()).detach().clone()
    eng = TrainEngine(m, lr=1e-3, ema_decay=0.99)
    x = torch.randn(2, 3, 224, 224, device='cuda')
    y = torch.randint(0, 1000, (2,), device='cuda')
    for _ in range(2):
        loss = eng.step(x, y)
    assert torch.isfinite(loss).item()
    g = eng.opt.state.grad
    assert torch.isfinite(g).all().item() and g.abs().max().item() > 0
    w1 = m.fc[0].weight if hasattr(m, 'fc') else next(m.parameters())
    assert (w1.detach() - w0).abs().max().item() > 0
    m.eval()
    out = evaluate_batch(m, x, y, 'mean' if name.startswith('map_') else 'sum', amp_dtype=None)
    assert all(torch.isfinite(t.float()).all().item() for t in out)


@pytest.mark.gpu
def test_device_prefetcher_matches_direct_normalisation():
    from imagenet_models_b200.engine import DevicePrefetcher
    pf = DevicePrefetcher()
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randint(0, 256, (4, 3, 32, 32), dtype=torch.uint8, generator=g).pin_memory(),
                torch.randint(0, 1000, (4,), generator=g).pin_memory()) for _ in range(3)]
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1) * 255
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1) * 255
    pf.submit(*batches[0])
    for i in range(3):
        x, y = pf.get()
        if i + 1 < 3:
            pf.submit(*batches[i + 1])                 # next copy runs under whatever the main stream does with x
        ref = (batches[i][0].float() - mean) / std
        assert torch.allclose(x.cpu(), ref, atol=1e-6) and torch.equal(y.cpu(), batches[i][1])
