"""Stub of gymnasium.envs.registration: the reference registers two ids at import time."""
registry = {}


def register(id, entry_point=None, vector_entry_point=None, **kwargs):
    registry[id] = dict(entry_point=entry_point, vector_entry_point=vector_entry_point, **kwargs)


def make_vec(id, num_envs=1, vectorization_mode=None, vector_kwargs=None, **kwargs):
    """What gymnasium.make_vec does for a spec that has a ``vector_entry_point``: import "module:Class" and call it with
    ``num_envs`` and the caller's keyword arguments."""
    import importlib
    spec = registry[id]
    target = spec["vector_entry_point"]
    if target is None:
        raise ValueError(f"{id} has no vector_entry_point")
    if isinstance(target, str):
        mod, _, name = target.partition(":")
        target = getattr(importlib.import_module(mod), name)
    return target(num_envs=num_envs, **dict(spec.get("kwargs") or {}, **kwargs))
