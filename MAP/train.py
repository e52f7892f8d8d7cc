#!/usr/bin/env python3
"""Drop-in entry point for the reference's `MAP/train.py` (run(): MAP/train.py:382; train_one_epoch :842-960) on the sm_100a path.

The training step is the same engine as GA/train.py; what is MAP-specific is the model family (train-mode outputs are
[main, self-distillation] pairs per group) and multi_group_loss (MAP/train.py:792-839) with --dec-lam, which the fused loss kernel
evaluates (distill_tokens == 0 branch, the published configuration).  A non-finite loss on any rank stops the run like
MAP/train.py:887-891.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imagenet_models_b200.map_convnext  # noqa: F401,E402
import imagenet_models_b200.ga_convnext  # noqa: F401,E402
from imagenet_models_b200.train_loop import build_parser, run  # noqa: E402

parser = build_parser('MAP training on B200 (flags follow MAP/train.py:60-380)', 'map_convnext_tiny', '--dec-lam', -0.8)


def _parse_args(argv=None):
    return parser.parse_known_args(argv)


if __name__ == '__main__':
    args, unknown = _parse_args()
    run(args, unknown)
