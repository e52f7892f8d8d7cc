"""The incumbent on the same GPU ("GPU eager baseline"): the oracle's plain-PyTorch restatement of the
reference modules, run by torch eager (cuDNN / cuBLAS kernels) under bf16 autocast with a channels_last input.  The reference
sources cannot travel to the GPU box, the restatement (pinned to them on CPU) can.  Not a parity test: it records the
throughput the hand-written kernels have to beat in gpurun_out/eager_baseline.json and checks the run was sane.

  EAGER_BASELINE_BATCH=256 python -m pytest tests/test_gpu_eager_oracle_baseline.py -m gpu -q -s
"""
import json
import os

import pytest
import torch

from oracle import cases
from oracle import ga_convnext_oracle as O


@pytest.mark.gpu
def test_record_torch_eager_training_throughput():
    name = 'ga_convnext_tiny_688'
    B = int(os.environ.get('EAGER_BASELINE_BATCH', '64'))
    spec = O.SPECS[name]
    P = {k: (v.cuda().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.cuda())
         for k, v in O.make_state(spec, cases.STATE_SEED).items()}
    params = [v for v in P.values() if v.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.05, fused=True)
    x = torch.randn(B, 3, 224, 224, device='cuda').contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (B,), device='cuda')

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = O.forward(P, spec, x, training=True)
        loss = O.ga_loss([o.float() for o in out], y, cases.GA_LAM)
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 5
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert torch.isfinite(loss).item()
    rec = {'what': 'torch eager (cuDNN/cuBLAS) training step of the oracle restatement: fwd + GA loss + bwd + fused AdamW, '
                   'bf16 autocast, channels_last input', 'model': name, 'batch': B, 'ms_per_step': ms,
           'img_per_s': B / ms * 1e3, 'torch': torch.__version__, 'gpu': torch.cuda.get_device_name(0)}
    root = os.environ.get('GRAFT_REPO_ROOT', os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.makedirs(os.path.join(root, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(root, 'gpurun_out', 'eager_baseline.json'), 'w') as f:
        json.dump(rec, f)
    print(json.dumps(rec))
    assert rec['img_per_s'] > 0
