"""Drive the UNMODIFIED reference (`/root/reference/src/gym_trading_env`) in lockstep and record it.

TEST INFRASTRUCTURE ONLY — runs in the build container (where `/root/reference` is mounted),
never on the GPU box and never from the product path.  `oracle/make_golden.py` uses it to write
the golden fixtures under `tests/golden/`; `tests/test_oracle_vs_reference.py` uses it (skipped
when `/root/reference` is absent) to cross-check the C restatement live.

It implements the lockstep vector-step contract of SURVEY.md §8(a) on top of N independent
reference `TradingEnv` objects:

    obs, r, term, trunc, info = env_i.step(a_i)         # reference semantics
    if term or trunc:  obs, info = env_i.reset()        # same call (in-place auto-reset)

Episode starts / initial positions / dataset choices come from the reference's own global
`np.random` stream (environments.py:167,174,385); they are *read back* after every reset and
stored as the `plan[N, E, 3]` tensor (record-and-replay, hazard H2) that the oracle port and the
CUDA path consume.

The one documented normalisation (hazard H3): with ``normalize_dyn=True`` the env's private
dynamic-feature columns are zeroed before each `reset()` so that stale values from earlier
episodes of the same env object do not leak into later windows.
"""
from __future__ import annotations

import glob
import os
import sys
import tempfile
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "gym_trading_env"))


def import_reference():
    """Import the reference package, adding the gymnasium stub if the real one is missing.

    The reference turns every warning into an error process-wide at import
    (environments.py:15, hazard H5); the filters are restored afterwards.
    """
    try:
        import gymnasium  # noqa: F401
    except Exception:
        stub = os.path.join(_HERE, "gymnasium_stub")
        if stub not in sys.path:
            sys.path.insert(0, stub)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    saved = warnings.filters[:]
    import gym_trading_env.environments as envmod  # noqa: E402
    warnings.filters[:] = saved
    return envmod


def _zero_dyn(env):
    nb = env._nb_static_features
    env._obs_array[:, nb:] = 0


def run_lockstep(dfs, n_envs, actions, *, positions, windows, trading_fees, borrow_interest_rate,
                 portfolio_initial_value, initial_position, max_episode_duration,
                 dynamic_features=True, normalize_dyn=True, np_seed=0,
                 multi_dataset=False, episodes_between_dataset_switch=1, max_episodes=None,
                 reward_function=None, limit_orders=None):
    """Run N reference envs for K lockstep iterations; return a dict of recorded arrays.

    dfs: list of DataFrames (one unless multi_dataset).  actions: int64 [K, N]; a negative
    action is passed to the reference as ``None`` (= hold, environments.py:234).
    limit_orders: list of (position, factor): after EVERY reset (which clears the env's orders, :168)
    ``env.add_limit_order(position, factor * close[start], persistent=True)`` is called in that order; the
    limits are returned as ``limit_plan[N, E, n_orders]``.
    """
    envmod = import_reference()
    K, N = actions.shape
    assert N == n_envs
    kwargs = dict(positions=list(positions), windows=windows, trading_fees=trading_fees,
                  borrow_interest_rate=borrow_interest_rate,
                  portfolio_initial_value=portfolio_initial_value, initial_position=initial_position,
                  max_episode_duration=max_episode_duration, verbose=0)
    if not dynamic_features:
        kwargs["dynamic_feature_functions"] = []
    if reward_function is not None:
        kwargs["reward_function"] = reward_function       # a caller-style Python callback over History

    np.random.seed(np_seed)
    tmpdir = None
    envs = []
    if multi_dataset:
        tmpdir = tempfile.TemporaryDirectory()
        for k, df in enumerate(dfs):
            d = df.copy()
            d["ds_id"] = float(k)          # non-feature column: identifies the dataset in env.df
            d.to_pickle(os.path.join(tmpdir.name, f"ds_{k:03d}.pkl"))
        for _ in range(N):
            envs.append(envmod.MultiDatasetTradingEnv(
                dataset_dir=os.path.join(tmpdir.name, "*.pkl"),
                episodes_between_dataset_switch=episodes_between_dataset_switch, **kwargs))
    else:
        assert len(dfs) == 1
        for _ in range(N):
            envs.append(envmod.TradingEnv(df=dfs[0], **kwargs))

    def ds_of(env):
        return int(env.df["ds_id"].iloc[0]) if multi_dataset else 0

    def pos_index(env):
        return list(positions).index(env._position)

    obs_shape = envs[0].observation_space.shape
    E = max_episodes or 64
    plan = np.full((N, E, 3), -1, dtype=np.int32)
    limit_plan = np.full((N, E, len(limit_orders or [])), np.nan, dtype=np.float64)
    cursor = np.zeros(N, dtype=np.int64)

    def do_reset(i):
        env = envs[i]
        if normalize_dyn and dynamic_features:
            _zero_dyn(env)
        obs, info = env.reset()
        e = cursor[i]
        if e >= E:
            raise RuntimeError(f"env {i} needs more than max_episodes={E} plan slots")
        plan[i, e] = (env._idx, pos_index(env), ds_of(env))
        for j, (lpos, factor) in enumerate(limit_orders or []):
            limit = float(factor * env._price_array[env._idx])
            env.add_limit_order(lpos, limit, persistent=True)          # environments.py:227-231
            limit_plan[i, e, j] = limit
        cursor[i] += 1
        return np.array(obs, dtype=np.float32, copy=True), info

    rec = {
        "obs0": np.zeros((N,) + obs_shape, np.float32),
        "obs": np.zeros((K, N) + obs_shape, np.float32),         # post-reset where a reset happened
        "step_obs": np.zeros((K, N) + obs_shape, np.float32),    # what step() itself returned (terminal obs)
        "idx": np.zeros((K, N), np.int32), "step": np.zeros((K, N), np.int32),
        "position": np.zeros((K, N), np.float64), "real_position": np.zeros((K, N), np.float64),
        "valuation": np.zeros((K, N), np.float64), "reward": np.zeros((K, N), np.float64),
        "terminated": np.zeros((K, N), np.uint8), "truncated": np.zeros((K, N), np.uint8),
        "asset": np.zeros((K, N), np.float64), "fiat": np.zeros((K, N), np.float64),
        "interest_asset": np.zeros((K, N), np.float64), "interest_fiat": np.zeros((K, N), np.float64),
        "market_return": np.full((K, N), np.nan), "portfolio_return": np.full((K, N), np.nan),
        # state after the (possible) in-place reset
        "post_idx": np.zeros((K, N), np.int32), "post_step": np.zeros((K, N), np.int32),
        "post_pos_idx": np.zeros((K, N), np.int32), "post_dataset": np.zeros((K, N), np.int32),
        "post_asset": np.zeros((K, N), np.float64), "post_fiat": np.zeros((K, N), np.float64),
        "post_interest_asset": np.zeros((K, N), np.float64),
        "post_interest_fiat": np.zeros((K, N), np.float64),
    }
    init_state = {k: np.zeros(N, np.float64) for k in ("asset", "fiat")}
    for i in range(N):
        rec["obs0"][i], _ = do_reset(i)
        init_state["asset"][i] = envs[i]._portfolio.asset
        init_state["fiat"][i] = envs[i]._portfolio.fiat

    for k in range(K):
        for i, env in enumerate(envs):
            a = int(actions[k, i])
            obs, reward, term, trunc, info = env.step(a if a >= 0 else None)
            pf = env._portfolio
            rec["step_obs"][k, i] = obs
            rec["idx"][k, i] = info["idx"]; rec["step"][k, i] = info["step"]
            rec["position"][k, i] = info["position"]; rec["real_position"][k, i] = info["real_position"]
            rec["valuation"][k, i] = info["portfolio_valuation"]
            rec["reward"][k, i] = float(reward)
            rec["terminated"][k, i] = term; rec["truncated"][k, i] = trunc
            rec["asset"][k, i] = pf.asset; rec["fiat"][k, i] = pf.fiat
            rec["interest_asset"][k, i] = pf.interest_asset; rec["interest_fiat"][k, i] = pf.interest_fiat
            if term or trunc:
                h = env.historical_info
                # numeric core of calculate_metrics (environments.py:279-283)
                rec["market_return"][k, i] = h["data_close", -1] / h["data_close", 0] - 1
                rec["portfolio_return"][k, i] = h["portfolio_valuation", -1] / h["portfolio_valuation", 0] - 1
                obs, _ = do_reset(i)
            rec["obs"][k, i] = obs
            pf = env._portfolio
            rec["post_idx"][k, i] = env._idx; rec["post_step"][k, i] = env._step
            rec["post_pos_idx"][k, i] = pos_index(env); rec["post_dataset"][k, i] = ds_of(env)
            rec["post_asset"][k, i] = pf.asset; rec["post_fiat"][k, i] = pf.fiat
            rec["post_interest_asset"][k, i] = pf.interest_asset
            rec["post_interest_fiat"][k, i] = pf.interest_fiat

    rec["plan"] = plan[:, :int(cursor.max())].copy()
    rec["limit_plan"] = limit_plan[:, :int(cursor.max())].copy()
    rec["init_asset"] = init_state["asset"]; rec["init_fiat"] = init_state["fiat"]
    if tmpdir is not None:
        tmpdir.cleanup()
    return rec
