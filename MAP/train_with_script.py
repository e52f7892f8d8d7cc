#!/usr/bin/env python3
"""Drop-in for the reference's `MAP/train_with_script.py:12-20`: named recipes -> MAP/train.py arguments.

  python MAP/train_with_script.py convnext_tiny -m map_convnext_tiny        (torchrun for several GPUs)
Only the recipes whose backbone is on the B200 path are listed; the strings are the reference's hyper-parameters with the
data / logging options that have no effect on synthetic data removed.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from train import _parse_args  # noqa: E402
from imagenet_models_b200.train_loop import run  # noqa: E402

setting_dict = dict(
    convnext_tiny="imageNet --drop-path .1 -b 128 --smoothing 0.1 --bce-loss --opt lamb --opt-eps 1e-8 --momentum 0.8 --weight-decay 0.05 "
                  "--sched cosine --epochs 300 --lr 5e-3 --warmup-lr 1e-6 --mixup .8 --cutmix 1.0 --amp --channels-last --model-ema "
                  "--model-ema-decay 0.9999",
    convnext_small="imageNet --drop-path .4 -b 128 --smoothing 0.1 --bce-loss --opt lamb --opt-eps 1e-8 --momentum 0.8 --weight-decay 0.05 "
                   "--sched cosine --epochs 300 --lr 5e-3 --warmup-lr 1e-6 --mixup .8 --cutmix 1.0 --amp --channels-last --model-ema "
                   "--model-ema-decay 0.9999",
)

if __name__ == '__main__':
    ap = argparse.ArgumentParser(description='recipe runner (MAP/train_with_script.py)')
    ap.add_argument('setup', type=str, choices=sorted(setting_dict))
    ap.add_argument('-m', '--model-name', type=str, default=None)
    ap.add_argument('--dec-lam', default=-0.8, type=float)
    ap.add_argument('--epochs', type=int, default=None, help='override the recipe (synthetic runs)')
    ap.add_argument('--steps-per-epoch', type=int, default=None)
    a = ap.parse_args()
    argv = setting_dict[a.setup].split() + ['--dec-lam', str(a.dec_lam), '--model', a.model_name or ('map_' + a.setup)]
    if a.epochs is not None:
        argv += ['--epochs', str(a.epochs)]
    if a.steps_per_epoch is not None:
        argv += ['--steps-per-epoch', str(a.steps_per_epoch)]
    args, unknown = _parse_args(argv)
    run(args, unknown)
