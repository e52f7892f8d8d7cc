"""Host-side logic of the step runtime that needs no GPU: the zero arena's bookkeeping, the one-draw DropPath factors, the
operand-preparation eligibility rule, the weight-shadow registry and the BatchNorm-counter collection."""
import torch

from imagenet_models_b200 import ops
from imagenet_models_b200.ga_convnext import _path_scales


def test_zero_arena_falls_back_without_a_device_buffer():
    A = ops.ZeroArena()
    A.begin('cpu')                                   # no CUDA device: never holds a buffer, every request is its own zeros
    a = A.take(10, torch.float32, 'cpu')
    b = A.take(7, torch.bfloat16, 'cpu')
    assert A.buf is None and a.shape == (10,) and b.dtype == torch.bfloat16 and a.sum() == 0 and b.float().sum() == 0
    assert A.used == 256 + 256                       # demand is recorded in 256-byte units for the next step's sizing
    A.begin('cpu')
    assert A.need == 512 and A.used == 0 and A.buf is None
    z = ops.zeros((3, 5), torch.float32, 'cpu')      # module-level helper, shape as tuple or int
    assert z.shape == (3, 5) and ops.zeros(4, torch.float32, 'cpu').shape == (4,)


def test_path_scales_one_draw_for_many_blocks():
    torch.manual_seed(0)
    probs = [0.0, 0.1, 0.0, 0.5, 0.2]
    out = _path_scales(probs, True, 4096, 'cpu')
    assert out[0] is None and out[2] is None
    for p, m in zip(probs, out):
        if p == 0.0:
            continue
        keep = 1.0 - p
        vals = torch.unique(m)
        assert m.shape == (4096,) and m.is_contiguous()
        assert all(abs(v.item()) < 1e-6 or abs(v.item() - 1.0 / keep) < 1e-5 for v in vals)      # 0 or 1 / keep
        assert abs((m > 0).float().mean().item() - keep) < 0.03                                   # kept with probability keep
        assert abs(m.mean().item() - 1.0) < 0.05                                                  # unbiased
    assert all(m is None for m in _path_scales(probs, False, 8, 'cpu'))                           # eval: identity
    assert _path_scales([], True, 8, 'cpu') == []


def test_block_weights_eligibility():
    def blk(C, hidden):
        return {'norm.weight': torch.ones(C), 'mlp.fc1.weight': torch.zeros(hidden, C)}
    assert ops.BlockWeights.supported([blk(96, 384), blk(688, 2752)], torch.bfloat16)
    assert not ops.BlockWeights.supported([blk(96, 384)], torch.float32)          # fp32 keeps the per-block fp32 preparation
    assert not ops.BlockWeights.supported([blk(100, 400)], torch.bfloat16)        # C % 8 != 0: no 16-byte operand rows
    assert not ops.BlockWeights.supported([blk(96, 192)], torch.bfloat16)         # not a 4x MLP


def test_weight_shadow_registry_matches_by_pointer_version_and_size():
    w = torch.nn.Parameter(torch.randn(8, 16))
    flat16 = torch.zeros(8 * 16, dtype=torch.bfloat16)
    ops.register_weight_shadows([w], [0], flat16)
    ent = ops._weight_shadows[w.data_ptr()]
    assert ent[0]() is w and ent[1] == w._version and ent[2].data_ptr() == flat16.data_ptr()
    with torch.no_grad():
        w.add_(1.0)                                  # written from outside the optimizer: the version no longer matches
    assert ops._weight_shadows[w.data_ptr()][1] != w._version
    ops.register_weight_shadows([w], [0], flat16)    # re-registration (FlatState.refresh_shadows) makes it current again
    assert ops._weight_shadows[w.data_ptr()][1] == w._version
    del w
    ops.register_weight_shadows([], [], flat16)      # dead parameters are dropped at the next registration
    assert all(v[0]() is not None for v in ops._weight_shadows.values())


def test_bn_counters_are_collected_and_bumped_once():
    bn = {'num_batches_tracked': torch.zeros((), dtype=torch.long)}
    bn2 = {'num_batches_tracked': torch.tensor(5)}

    class Stop(Exception):
        pass

    def fake_apply(*a, **k):
        raise Stop                                   # the test is about the counters, not the kernel behind BatchNormFn

    orig = ops.BatchNormFn.apply
    ops.BatchNormFn.apply = fake_apply
    try:
        with ops.collect_bn_counters():
            for d in (bn, bn2):
                try:
                    ops.batchnorm(torch.zeros(2, 4), dict(d, weight=None, bias=None, running_mean=None, running_var=None), True)
                except Stop:
                    pass
            assert bn['num_batches_tracked'].item() == 0 and len(ops.BN_COUNTERS) == 2      # deferred while collecting
        assert bn['num_batches_tracked'].item() == 1 and bn2['num_batches_tracked'].item() == 6 and ops.BN_COUNTERS is None
        try:
            ops.batchnorm(torch.zeros(2, 4), dict(bn, weight=None, bias=None, running_mean=None, running_var=None), True)
        except Stop:
            pass
        assert bn['num_batches_tracked'].item() == 2                                         # outside: advanced immediately
        try:
            ops.batchnorm(torch.zeros(2, 4), dict(bn, weight=None, bias=None, running_mean=None, running_var=None), False)
        except Stop:
            pass
        assert bn['num_batches_tracked'].item() == 2                                         # eval: untouched
    finally:
        ops.BatchNormFn.apply = orig
