"""Model registry: name -> factory (reference: lvae/models/registry.py:1-15)."""
_registry = {}


def register_model(func):
    name = func.__name__
    if name in _registry:
        print(f'Warning: model function *{name}* is multiply defined.')
    _registry[name] = func
    return func


def get_model(name, *args, **kwargs):
    if name not in _registry:
        raise KeyError(f'unknown model {name!r}; registered: {sorted(_registry)}')
    return _registry[name](*args, **kwargs)
