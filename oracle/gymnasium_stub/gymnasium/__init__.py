"""Minimal stand-in for the third-party ``gymnasium`` package (NOT part of the reference).

TEST INFRASTRUCTURE ONLY.  The reference (`/root/reference/src/gym_trading_env`) imports four
names from gymnasium (`environments.py:1-2`, `__init__.py:1`): ``Env``, ``spaces.Discrete``,
``spaces.Box`` and ``envs.registration.register``.  gymnasium is not installed in this image and
there is no network, so `oracle/make_golden.py` puts this directory on ``sys.path`` to let the
UNMODIFIED reference import and run.  It is only used when the real gymnasium is missing.
"""
from . import spaces  # noqa: F401
from . import vector  # noqa: F401
from .envs import registration  # noqa: F401
from .envs.registration import make_vec, register  # noqa: F401


class Env:
    """Shape of gymnasium.Env that TradingEnv relies on: reset(seed, options) is a no-op hook."""

    metadata = {}
    np_random = None

    def reset(self, *, seed=None, options=None):
        # gymnasium seeds self.np_random here; the reference never draws from it
        # (environments.py:167,174 use the global numpy RNG), so nothing to do.
        return None

    def close(self):
        pass


def make(*args, **kwargs):  # pragma: no cover - not used by the harness
    raise NotImplementedError("gymnasium stub: construct TradingEnv directly")
