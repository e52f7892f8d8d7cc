#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: isolates ONE training step (the kernels
between two consecutive adamw_ema launches) and prints per-kernel totals and the share of time spent in libga_sm100.so kernels.

  python scripts/summarise_launches.py gpurun_out/launches.csv [step_index] > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import re
import sys

FOREIGN = ('at::', 'cub::', 'thrust::', 'nccl', 'cutlass', 'cudnn', 'cublas', 'distribution_', 'elementwise_kernel', 'reduce_kernel',
           'CatArrayBatchedCopy', 'index_', 'gatherTopK', 'bitonic', 'fill')


def is_own(name: str) -> bool:
    """Kernels of libga_sm100.so: everything that is not an ATen / CUB / NCCL kernel."""
    return not any(f in name for f in FOREIGN)


def main():
    path = sys.argv[1]
    want = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        rows.append((r['Kernel Name'], ns))
    ends = [i for i, (k, _) in enumerate(rows) if 'adamw_ema' in k]
    if len(ends) <= want:
        raise SystemExit(f'only {len(ends)} optimizer launches in the list')
    lo, hi = ends[want - 1] + 1, ends[want] + 1
    while hi < len(rows) and 'ema_lerp' in rows[hi][0]:
        hi += 1
    while lo < len(rows) and 'ema_lerp' in rows[lo][0]:
        lo += 1
    step = rows[lo:hi]
    tot = sum(ns for _, ns in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    own_n = own_t = 0
    for k, ns in step:
        name = re.sub(r'\(.*', '', k).replace('void ', '')
        agg[name][0] += 1
        agg[name][1] += ns
        if is_own(name):
            own_n += 1
            own_t += ns
    print(f'one training step = launches {lo}..{hi - 1} of the list: {len(step)} launches, {tot / 1e6:.3f} ms serialised')
    print('(cold-cache, serialised per-launch times under the profiler: compare SHARES, not absolutes)')
    print(f'libga_sm100.so kernels: {own_n} launches, {own_t / tot * 100:.1f}% of the time; ATen glue / fills: {len(step) - own_n} launches, '
          f'{(tot - own_t) / tot * 100:.1f}%\n')
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f'{t / tot * 100:6.2f}%  {t / 1e6:8.3f} ms  n={n:4d}  {name[:130]}')


if __name__ == '__main__':
    main()
