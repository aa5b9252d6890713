import importlib

_REGISTRY = {}


def register(id, entry_point=None, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, kw = _REGISTRY[id]
    if callable(entry_point):
        cls = entry_point
    else:
        mod_name, attr = entry_point.split(":")
        cls = getattr(importlib.import_module(mod_name), attr)
    # gym 0.20: no TimeLimit wrapper unless max_episode_steps is registered.
    return cls(**{**kw.get("kwargs", {}), **kwargs})
