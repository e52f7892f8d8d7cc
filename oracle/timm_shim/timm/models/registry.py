_ENTRYPOINTS = {}


def register_model(fn):
    _ENTRYPOINTS[fn.__name__] = fn
    return fn


def create_model(model_name, pretrained=False, checkpoint_path='', **kwargs):
    # timm.create_model drops None-valued kwargs before calling the entrypoint.
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _ENTRYPOINTS[model_name](pretrained=pretrained, **kwargs)
