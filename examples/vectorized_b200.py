"""The reference's vectorised example (examples/example_vectorized_environment.py) on the B200 path:
same feature recipe and constructor arguments, N envs advanced in lockstep by CUDA kernels.

    python examples/vectorized_b200.py            # needs a B200 and the built library (see README)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_trading_env_b200 as gte  # noqa: E402


def main(num_envs=4096, steps=1000):
    df = gte.make_gbm_ohlcv(50_000, seed=0)                    # or: your OHLCV DataFrame with "feature*" columns
    env = gte.TradingVectorEnv(
        df, name="SYNTH", windows=64,
        positions=[-1, -0.5, 0, 0.5, 1, 1.5, 2],               # from -1 (short) to +2 (leveraged long)
        initial_position="random", trading_fees=0.01 / 100, borrow_interest_rate=0.0003 / 100,
        reward_function=gte.log_return_reward(clip=(-0.002, 0.005)),      # the callers' clipped log-return
        portfolio_initial_value=1000, max_episode_duration=500, num_envs=num_envs, seed=0, verbose=0)
    obs, infos = env.reset()
    print("observation batch", tuple(obs.shape), obs.dtype, "on", obs.device)
    n_actions = env.single_action_space.n
    total_reward = torch.zeros((), dtype=torch.float64, device=env.device)
    for _ in range(steps):
        actions = torch.randint(0, n_actions, (num_envs,), device=env.device)     # your policy goes here
        obs, reward, terminated, truncated, infos = env.step(actions)
        total_reward += reward.sum()
    m = {k: float(v) for k, v in env.get_metrics().items()}
    print(f"{num_envs * steps:,} env-steps; {int(m['episodes'])} episodes finished "
          f"({int(m['terminated'])} stopped out); mean portfolio return per episode "
          f"{100 * m['sum_portfolio_return'] / max(m['episodes'], 1):.2f} %, market {100 * m['sum_market_return'] / max(m['episodes'], 1):.2f} %")
    print("mean reward per env-step", float(total_reward) / (num_envs * steps))
    print("current valuations (first 4 envs)", infos["portfolio_valuation"][:4].tolist())


if __name__ == "__main__":
    main()
