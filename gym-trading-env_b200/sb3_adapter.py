"""stable-baselines3 ``VecEnv`` view of a ``TradingVectorEnv``.

The reference's callers train with SB3 on ``gym.make("TradingEnv", ...)`` (luckymodel/envs/env.py:67-95,
luckymodel/scripts/train_RPPO.py:16-38), which SB3 wraps into a ``DummyVecEnv`` of single Python envs.  This adapter
gives the same ``VecEnv`` surface — ``reset() -> obs``, ``step_async`` / ``step_wait() -> (obs, rewards, dones, infos)``,
``infos[i]["terminal_observation"]`` / ``["TimeLimit.truncated"]`` for the envs whose episode ended — on top of the
batched CUDA env, so ``PPO("MlpPolicy", SB3VecEnv(env))`` runs N lockstep envs on the GPU.

stable-baselines3 is not part of this image: the class derives from ``stable_baselines3.common.vec_env.VecEnv`` when
it can be imported and is duck-typed otherwise.  Observations / rewards are returned as numpy arrays (what SB3's
rollout buffers take); the per-env ``infos`` list is what bounds this adapter to SB3-sized batches (thousands of
envs, not millions) — use ``TradingVectorEnv`` directly for on-device policies.
"""
from __future__ import annotations

import numpy as np
import torch

try:                                                     # pragma: no cover - not installed in this image
    from stable_baselines3.common.vec_env import VecEnv as _Base
except Exception:                                        # noqa: BLE001
    _Base = object


class SB3VecEnv(_Base):
    """``SB3VecEnv(TradingVectorEnv(..., final_obs=True))``.

    ``final_obs=True`` makes the env keep the terminal observations SB3 bootstraps from on truncation; without it
    ``terminal_observation`` is omitted.  ``info_keys`` selects which History columns (``env.infos`` keys) are copied
    into every per-env info dict each step (default: none — they cost one device->host copy per key)."""

    def __init__(self, env, info_keys=()):
        self.env = env
        env.output = "numpy"                             # obs / reward / flags through pinned host buffers, one sync per step
        env._host = None
        self.num_envs = env.num_envs
        self.observation_space = env.single_observation_space
        self.action_space = env.single_action_space
        self.render_mode = getattr(env, "render_mode", None)
        self.info_keys = tuple(info_keys)
        self._actions = None
        if _Base is not object:                          # pragma: no cover
            _Base.__init__(self, self.num_envs, self.observation_space, self.action_space)

    @staticmethod
    def _np(x):
        return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)

    def reset(self):
        obs, _ = self.env.reset()
        return self._np(obs).copy()

    def step_async(self, actions):
        self._actions = np.asarray(actions, dtype=np.int64).reshape(self.num_envs)

    def step_wait(self):
        obs, reward, term, trunc, infos = self.env.step(self._actions)
        term, trunc = self._np(term).astype(bool), self._np(trunc).astype(bool)
        dones = term | trunc
        out = [{} for _ in range(self.num_envs)]
        for key in self.info_keys:
            col = self._np(infos[key])
            for i in range(self.num_envs):
                out[i][key] = col[i]
        ended = np.flatnonzero(dones)
        if ended.size:
            final = None
            if getattr(self.env, "final_obs", None) is not None:
                final = self._np(self.env.final_obs[torch.as_tensor(ended, device=self.env.device)])
            for j, i in enumerate(ended):
                out[i]["TimeLimit.truncated"] = bool(trunc[i] and not term[i])
                if final is not None:
                    out[i]["terminal_observation"] = final[j]
        return self._np(obs).copy(), self._np(reward).astype(np.float32), dones, out

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.env.close()

    def seed(self, seed=None):
        return [None] * self.num_envs                    # episode starts come from the env's Philox seed (constructor)

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        return [indices] if isinstance(indices, int) else list(indices)

    def get_attr(self, attr_name, indices=None):
        return [getattr(self.env, attr_name) for _ in self._indices(indices)]

    def set_attr(self, attr_name, value, indices=None):
        setattr(self.env, attr_name, value)

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        res = getattr(self.env, method_name)(*method_args, **method_kwargs)
        return [res for _ in self._indices(indices)]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def get_images(self):
        return [None] * self.num_envs

    def render(self, mode=None):
        return None
