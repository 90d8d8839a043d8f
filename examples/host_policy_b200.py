"""A policy that lives on the HOST (numpy in, numpy out — how the reference's examples and stable-baselines3 drive the
env, examples/example_vectorized_environment.py:58-62) on the B200 path.

    python examples/host_policy_b200.py            # needs a B200 and the built library (see README)

output="hybrid": actions come from pinned host memory (int8: Discrete(P) always fits), reward / terminated / truncated
land in pinned host memory every step, the observation windows stay in HBM.  One C call per step (gte_step_host); for a
small windows=None batch host_io="server" keeps a kernel resident so that a step costs no kernel launch at all.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_trading_env_b200 as gte  # noqa: E402


def run(env, steps, policy):
    env.reset()
    actions = env.pinned_actions()                       # pinned int8 [N]: filled in place, read by the GPU in place
    total, t0 = 0.0, time.perf_counter()
    for _ in range(steps):
        policy(actions)
        obs, reward, terminated, truncated, infos = env.step(actions)     # numpy reward / flags, CUDA obs
        total += float(reward.sum())
    dt = time.perf_counter() - t0
    print(f"{env.num_envs:>8,} envs: {1e6 * dt / steps:8.1f} us per step, {env.num_envs * steps / dt:.3e} env-steps/s, "
          f"mean reward {total / (env.num_envs * steps):+.2e}")


def main():
    df = gte.make_gbm_ohlcv(50_000, seed=0)
    rng = np.random.default_rng(0)
    kw = dict(positions=[-1, 0, 1], trading_fees=0.01 / 100, borrow_interest_rate=0.0003 / 100, max_episode_duration=500,
              verbose=0, output="hybrid")

    def random_policy(a):
        a[...] = rng.integers(0, 3, size=a.shape)

    # small batch, one-row observations: the resident server kernel answers each step through mapped host memory
    run(gte.TradingVectorEnv(df, windows=None, num_envs=4096, host_io="server", **kw), 2000, random_policy)
    # the same through the default mechanism (mapped memory, one launch per step)
    run(gte.TradingVectorEnv(df, windows=None, num_envs=4096, **kw), 2000, random_policy)
    # a large windowed batch: copy engines, one copy per direction, the flags as the list of ended envs
    run(gte.TradingVectorEnv(df, windows=64, num_envs=1 << 20, **kw), 100, random_policy)


if __name__ == "__main__":
    main()
