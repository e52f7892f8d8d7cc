#!/usr/bin/env python3
"""Drop-in entry point for the reference's `GA/train.py` on the sm_100a hot path.

Keeps the reference's launch convention (`torchrun --nproc_per_node=N GA/train.py <data> --model ga_convnext_tiny_688 ...`,
one process per GPU, NCCL, env:// rendezvous; GA/train.py:374-381) and the flags of the published recipe (GA/README.md:26:
--opt lamb --lr 5e-3 --weight-decay .05 --sched cosine --bce-loss --smoothing 0.1 --mixup .8 --cutmix 1.0 --GA_lam -0.8
--model-ema --amp --channels-last).  The body is imagenet_models_b200/train_loop.py.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imagenet_models_b200.ga_convnext  # noqa: F401,E402  (registers the ga_convnext_* factories)
import imagenet_models_b200.map_convnext  # noqa: F401,E402  (registers map_convnext_tiny / _small)
import imagenet_models_b200.ga_cswin  # noqa: F401,E402  (GA_CSWinTransformer, ga_CSWin_64_12211_tiny_224)
from imagenet_models_b200.train_loop import build_parser, run  # noqa: E402

parser = build_parser('GA training on B200 (flags follow GA/train.py:69-309)', 'ga_convnext_tiny_688', '--GA_lam', 0.0)

if __name__ == '__main__':
    args, unknown = parser.parse_known_args()
    run(args, unknown)
