"""timm-compatible model registry (reference boundary: `@register_model` + `timm.create_model`).

When the real `timm` is importable the entrypoints are registered there as well, so `timm.create_model(name)`
in the reference's train.py / validate.py resolves to these modules.  timm is absent from this image, so the
local registry below is what the entry points in this repo use.
"""
from __future__ import annotations

from typing import Callable, Dict

_ENTRYPOINTS: Dict[str, Callable] = {}

try:  # pragma: no cover - timm is not installed in the build image
    from timm.models import register_model as _timm_register
except Exception:  # noqa: BLE001
    _timm_register = None


def register_model(fn: Callable) -> Callable:
    _ENTRYPOINTS[fn.__name__] = fn
    if _timm_register is not None:
        try:
            _timm_register(fn)
        except Exception:  # noqa: BLE001  (already registered under the same name)
            pass
    return fn


def list_models():
    return sorted(_ENTRYPOINTS)


def is_model(name: str) -> bool:
    return name in _ENTRYPOINTS


def create_model(model_name: str, pretrained: bool = False, checkpoint_path: str = '', **kwargs):
    """timm.create_model semantics for the subset the reference uses: None-valued kwargs are dropped
    (GA/train.py:407-420 passes drop_rate, drop_connect_rate, drop_block_rate, global_pool, bn_* ... as None)."""
    if model_name not in _ENTRYPOINTS:
        raise RuntimeError(f'Unknown model ({model_name}); known: {list_models()}')
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    model = _ENTRYPOINTS[model_name](pretrained=pretrained, **kwargs)
    if checkpoint_path:
        import torch
        # timm's CheckpointSaver pickles an argparse.Namespace ('args') and optimizer state next to the weights
        # (GA/train.py:649): that needs the full unpickler, like timm.models.load_checkpoint does for trusted files
        ckpt = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
        sd = ckpt.get('state_dict_ema') or ckpt.get('state_dict') or ckpt.get('model') or ckpt
        sd = {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}
        model.load_state_dict(sd, strict=True)
    return model


def build_model_with_cfg(model_cls, variant, pretrained=False, **kwargs):
    """The three things timm's helper does for the reference: pop cfg kwargs, refuse unreachable weights, construct."""
    for k in ('pretrained_cfg', 'pretrained_cfg_overlay', 'default_cfg', 'features_only', 'pretrained_strict',
              'pretrained_filter_fn', 'kwargs_filter', 'feature_cfg'):
        kwargs.pop(k, None)
    if pretrained:
        raise RuntimeError(f'{variant}: no pretrained weights are published for the GA models (default_cfg url is empty)')
    return model_cls(**kwargs)
