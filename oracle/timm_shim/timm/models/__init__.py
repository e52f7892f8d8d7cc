"""timm.models subset: registry + named_apply + build_model_with_cfg (timm 0.9.2 semantics)."""
from .registry import register_model, create_model, _ENTRYPOINTS  # noqa: F401
from . import layers  # noqa: F401


def register_notrace_module(cls):
    return cls


def named_apply(fn, module, name='', depth_first=True, include_root=False):
    # timm/models/_manipulate.py: children first (depth-first), root excluded by default.
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        child_name = '.'.join((name, child_name)) if name else child_name
        named_apply(fn, child, name=child_name, depth_first=depth_first, include_root=True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module


def build_model_with_cfg(model_cls, variant, pretrained=False, **kwargs):
    for k in ('pretrained_cfg', 'pretrained_cfg_overlay', 'default_cfg', 'features_only',
              'pretrained_strict', 'pretrained_filter_fn', 'kwargs_filter', 'feature_cfg'):
        kwargs.pop(k, None)
    assert not pretrained, 'no checkpoints are reachable from this image'
    return model_cls(**kwargs)
