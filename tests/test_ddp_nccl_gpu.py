"""Two ranks over NCCL on two GPUs (skipped on a one-GPU box): the data-parallel training step of imagenet_models_b200.engine in
its three forms -- 'eager' (bucketed all-reduce overlapped with backward from grad-ready hooks), 'graph' (forward / backward /
gather replayed as a CUDA graph, then one all-reduce of the flat gradient and the optimizer) and 'graph+nccl' (the buckets'
NCCL calls captured inside the step graph on a forked side stream) -- against ONE process on the concatenated batch.
Reference behaviour: DistributedDataParallel gradient averaging (GA/train.py:505-515).  BatchNorm runs on its running statistics
(model.eval() with gradients enabled), the only mode in which a sharded batch and the whole batch define the same function; fp32
compute so the comparison is about the collective, not about rounding."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from oracle import cases

B_PER_RANK, STEPS, LR = 4, 4, 1e-3
MODES = ('eager', 'graph', 'graph+nccl')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    import imagenet_models_b200.ga_convnext  # noqa: F401
    from imagenet_models_b200.registry import create_model
    from oracle import ga_convnext_oracle as O
    name = 'ga_convnext_tiny_688'
    m = create_model(name).cuda()
    P = O.make_state(O.SPECS[name], cases.STATE_SEED, profile='trained')
    m.load_state_dict({k: v.cuda() for k, v in P.items()}, strict=True)
    return m.eval()                     # BatchNorm on running statistics, no drop path; gradients still flow


def _batches(world):
    return [cases.ga_inputs_diverse(B_PER_RANK * world, seed=1000 + s) for s in range(STEPS)]


def _worker(rank, world, port, q, outdir):
    import time

    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      TORCH_NCCL_ASYNC_ERROR_HANDLING='0')
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world)
    from imagenet_models_b200.engine import TrainEngine
    batches = _batches(world)
    out = {'rank': rank}
    progress = os.environ.get('GA_TEST_PROGRESS')          # optional: a file prefix that receives one line per finished stage

    def note(msg):
        if progress:
            with open(f'{progress}.rank{rank}', 'a') as f:
                f.write(f'{time.time():.1f} {msg}\n')
    note('process group up')
    try:
        for mode in MODES:
            note(f'{mode}: start')
            m = _model()
            eng = TrainEngine(m, lr=LR, weight_decay=0.05, ema_decay=0.99, ga_lam=cases.GA_LAM, amp_dtype=None, cuda_graph=mode != 'eager',
                              graph_warmup=2, bucket_mb=25.0, ddp_in_graph=mode == 'graph+nccl')
            m.eval()
            first = None
            for x, y in batches:
                eng.step(x[rank * B_PER_RANK:(rank + 1) * B_PER_RANK].cuda(), y[rank * B_PER_RANK:(rank + 1) * B_PER_RANK].cuda())
                if first is None:
                    first = (eng.opt.state.grad / world).cpu().numpy()     # the optimizer folds 1/world into its gradient scale
            torch.cuda.synchronize()
            note(f'{mode}: {len(batches)} steps done')
            # the arrays (190 MB each) go through files: a multiprocessing queue would still be feeding them through its pipe
            # when this process takes its hard exit below
            np.save(os.path.join(outdir, f'grad0_{mode}_{rank}.npy'), first)
            np.save(os.path.join(outdir, f'flat_{mode}_{rank}.npy'), eng.opt.state.flat.cpu().numpy())
            out[mode] = {'graph': eng._graph is not None, 'nbuckets': len(eng.buckets.buckets)}
    except Exception as e:  # noqa: BLE001  (report instead of leaving the peer in a collective for ever)
        out['error'] = repr(e)
    note('results written: ' + ('error ' + out['error'] if 'error' in out else 'ok'))
    q.put(out)
    time.sleep(2.0)            # let the queue's feeder thread hand the (small) result over before the hard exit
    os._exit(0)                # not destroy_process_group(): it blocks once a CUDA graph holds NCCL kernels (scripts/nccl_graph_probe.py)


@pytest.mark.gpu
def test_two_rank_nccl_steps_match_one_process_on_the_whole_batch(tmp_path):
    import numpy as np
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs (gpurun --gpus 2)')
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    try:
        outs = sorted([q.get(timeout=300) for _ in range(world)], key=lambda o: o['rank'])
    finally:
        for p in procs:
            p.join(timeout=20)
            if p.is_alive():
                p.kill()               # the exact processes this test started
    assert all('error' not in o for o in outs), [o.get('error') for o in outs]
    for o in outs:
        for mode in MODES:
            o[mode]['grad0'] = np.load(tmp_path / f"grad0_{mode}_{o['rank']}.npy")
            o[mode]['flat'] = np.load(tmp_path / f"flat_{mode}_{o['rank']}.npy")
    # single process, whole batch
    from imagenet_models_b200.engine import TrainEngine
    torch.cuda.set_device(0)
    m = _model()
    eng = TrainEngine(m, lr=LR, weight_decay=0.05, ema_decay=0.99, ga_lam=cases.GA_LAM, amp_dtype=None, cuda_graph=False)
    m.eval()
    ref_grad0 = None
    for x, y in _batches(world):
        eng.step(x.cuda(), y.cuda())
        if ref_grad0 is None:
            ref_grad0 = eng.opt.state.grad.cpu().clone()
    ref_flat = eng.opt.state.flat.cpu()
    for mode in MODES:
        assert outs[0][mode]['graph'] == (mode != 'eager') and outs[0][mode]['nbuckets'] >= 2
        for o in outs:
            g0 = torch.from_numpy(o[mode]['grad0'])
            # step 1 starts from identical weights: reduced gradient == whole-batch gradient (summation order only)
            err = (g0 - ref_grad0).norm().item() / ref_grad0.norm().item()
            assert err <= 2e-5, (mode, err)
            # later steps start from weights that differ by optimizer round-off (AdamW turns 1e-7 gradient noise into +-lr on
            # near-zero gradients), so the bound is per step: nobody moves more than lr per step away from the single process
            assert (torch.from_numpy(o[mode]['flat']) - ref_flat).abs().max().item() <= STEPS * LR * 1.05, mode
        assert (torch.from_numpy(outs[0][mode]['flat']) - torch.from_numpy(outs[1][mode]['flat'])).abs().max().item() == 0.0, mode   # replicas stay bit-identical
