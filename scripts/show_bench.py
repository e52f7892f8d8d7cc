#!/usr/bin/env python3
"""Print the key fields of bench.py JSON lines: python scripts/show_bench.py file.json [...]"""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    if 'unavailable' in d:
        print(path, d)
        continue
    e = d.get('e2e') or {}
    r = d.get('roofline') or {}
    print(f"{path}: {d['value']:.1f} {d['unit']} ({d['ms_per_step']:.2f} ms/step, n_gpus {d['n_gpus']}) e2e {e.get('value')} "
          f"roofline {r.get('frac')} [{(r.get('kernel') or '')[-48:]}] launches {d.get('gpu_launches')} clocks {d.get('clocks', {}).get('reasons')}")
