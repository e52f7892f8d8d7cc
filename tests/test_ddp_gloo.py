"""CPU, world_size 2 over gloo: the bucketed gradient all-reduce (imagenet_models_b200.optim.GradBuckets) that replaces
DistributedDataParallel on the training path (GA/train.py:505-515).  Runs here without a GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(16, 64), nn.GELU(), nn.Linear(64, 64), nn.GELU(), nn.Linear(64, 10))


def _worker(rank, world, port, bucket_mb, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from imagenet_models_b200.optim import FlatState, GradBuckets
    net = _net()
    state = FlatState(net)
    buckets = GradBuckets(state, bucket_mb=bucket_mb, overlap=False)
    g = torch.Generator().manual_seed(100 + rank)            # different data per rank (seed + rank, GA/train.py:402)
    results = []
    for step in range(2):
        x = torch.randn(8, 16, generator=g)
        state.zero_grad()
        buckets.prepare()
        net(x).square().mean().backward()
        scale = buckets.finish()
        # gradients are adopted by autograd (no per-parameter accumulate) and collected bucket by bucket into the flat buffer
        results.append((state.grad.clone() * scale, [p.grad is not None for p in state.params]))
    q.put((rank, [r[0].tolist() for r in results], all(all(r[1]) for r in results), len(buckets.buckets)))   # plain lists: no shm handles
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('bucket_mb', [25.0, 0.004])        # one bucket / many small buckets
def test_bucketed_allreduce_matches_mean_of_rank_gradients(bucket_mb):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, bucket_mb, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort(key=lambda t: t[0])
    # reference: mean over ranks of single-process gradients on each rank's data
    refs = []
    for step in range(2):
        acc = None
        for rank in range(world):
            net = _net()
            g = torch.Generator().manual_seed(100 + rank)
            for s in range(step + 1):
                x = torch.randn(8, 16, generator=g)
            net.zero_grad()
            net(x).square().mean().backward()
            flat = torch.cat([torch.nn.functional.pad(p.grad.reshape(-1), (0, (-p.numel()) % 16)) for p in net.parameters()])
            acc = flat if acc is None else acc + flat
        refs.append(acc / world)
    for rank, grads, views_ok, nb in out:
        assert views_ok, 'every parameter must have received a gradient'
        assert nb >= 1 and (bucket_mb > 1 or nb > 1)
        for step in range(2):
            assert torch.allclose(torch.tensor(grads[step]), refs[step], atol=1e-6), (rank, step)
    assert out[0][1][1] == out[1][1][1]            # both ranks hold identical reduced gradients


def test_bucket_layout_is_reverse_registration_order():
    from imagenet_models_b200.optim import FlatState, GradBuckets
    net = _net()
    state = FlatState(net)
    b = GradBuckets(state, bucket_mb=0.004, overlap=False)
    order = [i for bk in b.buckets for i in bk['members']]
    assert order == list(reversed(range(len(state.params))))  # backward produces the last layer's gradients first
    covered = sorted((bk['lo'], bk['hi']) for bk in b.buckets)
    assert covered[0][0] == 0 and covered[-1][1] == state.numel
    for (lo0, hi0), (lo1, hi1) in zip(covered, covered[1:]):
        assert hi0 == lo1


def test_flat_state_keeps_values_and_decay_flags():
    from imagenet_models_b200.optim import FlatState
    net = _net()
    before = [p.detach().clone() for p in net.parameters()]
    st = FlatState(net)
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p, b) and p.data_ptr() >= st.flat.data_ptr()
    flags = st.decay_flags(True)
    # weights decay, biases do not (timm filter_bias_and_bn)
    for name, p, o in zip(st.names, st.params, st.offsets):
        assert int(flags[o >> 4]) == (0 if name.endswith('.bias') else 1)


def _engine_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from imagenet_models_b200.engine import TrainEngine
    torch.manual_seed(1000 + rank)                           # the trainers seed with seed + rank (GA/train.py:402): different initialisations
    net = nn.Sequential(nn.Linear(16, 32), nn.BatchNorm1d(32), nn.GELU(), nn.Linear(32, 10))
    net[1].running_mean.add_(float(rank) + 1.0)              # buffers differ too
    before = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()
    eng = TrainEngine(net, lr=1e-3, ema_decay=0.99, amp_dtype=None, cuda_graph=False)
    after = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    ema = torch.cat([p.detach().reshape(-1) for p in eng.model_ema.parameters()])
    q.put((rank, before.tolist(), after.tolist(), ema.tolist(), net[1].running_mean.tolist(), eng.model_ema[1].running_mean.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_engine_construction_copies_rank0_parameters_and_buffers():
    """TrainEngine with several ranks starts every replica (and its EMA copy) from rank 0's parameters and buffers, as
    DistributedDataParallel does when it wraps a module (GA/train.py:505-515)."""
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_engine_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, b0, a0, e0, rm0, erm0), (_, b1, a1, e1, rm1, erm1) = out
    assert b0 != b1                                          # the initialisations did differ
    assert a0 == b0 and a1 == b0                             # both replicas hold rank 0's parameters
    assert e0 == b0 and e1 == b0                             # ... and so do their EMA copies
    assert rm0 == rm1 == erm0 == erm1 and abs(rm0[0] - 1.0) < 1e-6      # buffers (rank 0 had added 1.0)
