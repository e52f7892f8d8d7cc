#!/usr/bin/env python3
"""Measured parity errors of the CUDA path on the BASELINE-shape fixtures (oracle/parity_check.py), as JSON.
  python scripts/parity_report.py [out.json]     (GPU box; prints one summary per (fixture, dtype))"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import parity_check as PC  # noqa: E402

if __name__ == '__main__':
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'parity_report.json')
    rows, full = [], {}
    for family, key in (('ga', 'config1'), ('ga', 'bf16'), ('map', 'map'), ('cswin', 'cswin')):
        fx = PC.load_fixture(family)
        for dtype in (torch.float32, torch.bfloat16):
            res = PC.measure(family, key, dtype, fx)
            s = PC.summarise(res)
            s['reference_self_noise'] = fx[key]['ref_self_noise']
            s['reference_bf16_autocast_self_error'] = fx[key]['ref_bf16_train_self_err']
            rows.append(s)
            full[f'{family}/{key}/{s["dtype"]}'] = {k: res[k] for k in ('tail_grads', 'grads_pinned', 'upstream_grads_raw')}
            print(json.dumps(s))
            sys.stdout.flush()
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump({'summary': rows, 'per_tensor': full}, open(out, 'w'), indent=1)
