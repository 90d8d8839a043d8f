"""Stub of gymnasium.vector (TEST INFRASTRUCTURE ONLY): the shape of gymnasium >= 1.0's ``VectorEnv`` base class that a
vector env written against the gymnasium API subclasses — class-level attribute slots, ``reset`` / ``step`` to
override, ``close`` -> ``close_extras``, ``unwrapped``."""


class VectorEnv:
    metadata = {}
    spec = None
    render_mode = None
    closed = False
    observation_space = None
    action_space = None
    single_observation_space = None
    single_action_space = None
    num_envs = None

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, actions):
        raise NotImplementedError

    def render(self):
        raise NotImplementedError

    def close(self, **kwargs):
        if self.closed:
            return
        self.close_extras(**kwargs)
        self.closed = True

    def close_extras(self, **kwargs):
        pass

    @property
    def unwrapped(self):
        return self
