"""Stub of gymnasium.envs.registration: the reference registers two ids at import time."""
registry = {}


def register(id, entry_point=None, **kwargs):
    registry[id] = dict(entry_point=entry_point, **kwargs)
