"""Generate the golden fixtures under tests/golden/ by EXECUTING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Every fixture holds the inputs (float32 static features exactly as `_set_df` casts them,
environments.py:141; float64 close prices), the constructor parameters, the action stream, the
recorded reset plan (episode start / initial position / dataset, read back from the reference
after each `reset()`, hazard H2) and the per-step outputs of N reference `TradingEnv` objects
driven in lockstep by `oracle/ref_harness.py`.  The fixtures travel to the GPU box; the
reference does not.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import pandas as pd

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
sys.path.insert(0, _ROOT)
sys.path.insert(0, _HERE)

import gym_trading_env_b200 as gte  # noqa: E402
import ref_harness as rh            # noqa: E402

OUT = os.path.join(_ROOT, "tests", "golden")
RECORD_KEYS = ["obs0", "obs", "step_obs", "idx", "step", "position", "real_position", "valuation", "reward",
               "terminated", "truncated", "asset", "fiat", "interest_asset", "interest_fiat",
               "market_return", "portfolio_return", "post_idx", "post_step", "post_pos_idx",
               "post_dataset", "post_asset", "post_fiat", "post_interest_asset",
               "post_interest_fiat", "plan", "init_asset", "init_fiat", "limit_plan"]


def btc_frame(rows=2500):
    """The bundled real-data fixture with the recipe of examples/example_environnement.py:11-23."""
    df = pd.read_csv("/root/reference/examples/data/BTC_USD-Hourly.csv", parse_dates=["date"], index_col="date")
    df.sort_index(inplace=True)
    df.dropna(inplace=True)
    df.drop_duplicates(inplace=True)
    df["feature_close"] = df["close"].pct_change()
    df["feature_open"] = df["open"] / df["close"]
    df["feature_high"] = df["high"] / df["close"]
    df["feature_low"] = df["low"] / df["close"]
    df["feature_volume"] = df["Volume USD"] / df["Volume USD"].rolling(7 * 24).max()
    df.dropna(inplace=True)
    df = df.drop(columns=[c for c in df.columns if df[c].dtype == object])
    return df.iloc[5000:5000 + rows].copy()


def stack_series(dfs):
    arrs = [gte.frame_to_arrays(d) for d in dfs]
    t_max = max(a.length for a in arrs)
    feats = np.zeros((len(arrs), t_max, arrs[0].features.shape[1]), np.float32)
    price = np.ones((len(arrs), t_max), np.float64)
    for k, a in enumerate(arrs):
        feats[k, :a.length] = a.features
        price[k, :a.length] = a.price
    return feats, price, np.array([a.length for a in arrs], np.int32)


# reward callbacks exactly as the reference's callers write them
def reward_clipped_log_return(history):          # luckymodel/envs/env.py:16-18
    log_return = np.log(history["portfolio_valuation", -1] / history["portfolio_valuation", -2])
    return np.clip(log_return, -0.002, 0.005)


def reward_scaled_log_return(history):           # luckymodel/scripts/test_env.py:20-22
    return 100 * np.log(history["portfolio_valuation", -1] / history["portfolio_valuation", -2])


def reward_simple_return(history):               # the commented-out variant of luckymodel/envs/env.py:19, unclipped
    return (history["portfolio_valuation", -1] - history["portfolio_valuation", -2]) / history["portfolio_valuation", -2]


REWARDS = {"clipped_log_return": (reward_clipped_log_return, dict(reward_kind=0, reward_scale=1.0, reward_clip=(-0.002, 0.005))),
           "scaled_log_return": (reward_scaled_log_return, dict(reward_kind=0, reward_scale=100.0, reward_clip=(None, None))),
           "simple_return": (reward_simple_return, dict(reward_kind=1, reward_scale=1.0, reward_clip=(None, None)))}


def make_case(name, dfs, *, n_envs, n_steps, action_seed, hold_fraction=0.0, reward=None, **kw):
    positions = kw["positions"]
    rng = np.random.default_rng(action_seed)
    actions = rng.integers(0, len(positions), size=(n_steps, n_envs)).astype(np.int64)
    if hold_fraction > 0:
        actions[rng.random(actions.shape) < hold_fraction] = -1
    rec = rh.run_lockstep(dfs, n_envs, actions, reward_function=REWARDS[reward][0] if reward else None, **kw)
    feats, price, lengths = stack_series(dfs)
    params = dict(kw)
    if reward:
        params["reward"] = REWARDS[reward][1]
    params.update(n_envs=n_envs, n_steps=n_steps, action_seed=action_seed, name=name,
                  reference="ten2net/Gym-Trading-Env src/gym_trading_env (unmodified), numpy %s" % np.__version__)
    out = {k: rec[k] for k in RECORD_KEYS}
    tmax = feats.shape[1]
    hi, lo = np.ones((len(dfs), tmax)), np.ones((len(dfs), tmax))
    for k_, d_ in enumerate(dfs):
        hi[k_, :len(d_)], lo[k_, :len(d_)] = d_["high"].to_numpy(), d_["low"].to_numpy()
    out.update(high=hi, low=lo)
    out.update(features=feats, price=price, lengths=lengths, actions=actions,
               positions=np.array(positions, np.float64), params=np.array(json.dumps(params)))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    ep = int(rec["terminated"].sum() + rec["truncated"].sum() - (rec["terminated"] & rec["truncated"]).sum())
    print(f"{name:28s} N={n_envs:3d} K={n_steps:5d} episodes_finished={ep:4d} terminated={int(rec['terminated'].sum()):3d} "
          f"plan={rec['plan'].shape} {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    os.makedirs(OUT, exist_ok=True)
    common = dict(trading_fees=0.01 / 100, borrow_interest_rate=0.0003 / 100, portfolio_initial_value=1000)
    gbm = gte.make_gbm_ohlcv(1200, seed=0)

    # C1: single env, windows=None, 'max' duration -> runs to the end of the data (truncation at T-1) twice
    make_case("c1_single_nowindow", [gbm], n_envs=1, n_steps=2600, action_seed=11,
              positions=[-1, 0, 0.5, 1], windows=None, initial_position="random",
              max_episode_duration="max", max_episodes=8, **common)
    # C2-like: several lockstep envs, windows=None, bounded episodes
    make_case("c2_lockstep_nowindow", [gbm], n_envs=16, n_steps=300, action_seed=12,
              positions=[-1, 0, 0.5, 1], windows=None, initial_position="random",
              max_episode_duration=50, max_episodes=16, **common)
    # C3/C5-like: windows + leveraged/short positions + duration-limited episodes (normalised H3)
    make_case("c3_windows_leveraged", [gbm], n_envs=8, n_steps=220, action_seed=13,
              positions=[-3, -2, -1, 0, 1, 2, 3], windows=8, initial_position="random",
              max_episode_duration=40, max_episodes=16, **common)
    # same shape, RAW reference (stale dynamic rows leak across episodes): pins the oracle's faithful mode
    make_case("raw_stale_dynamic_rows", [gbm.iloc[:260]], n_envs=4, n_steps=200, action_seed=14,
              positions=[-3, -2, -1, 0, 1, 2, 3], windows=8, initial_position="random",
              max_episode_duration=40, normalize_dyn=False, max_episodes=16, **common)
    # windows=64 on the full window geometry of the headline configs, fixed initial position, hold actions
    make_case("w64_fixed_start_holds", [gbm], n_envs=4, n_steps=150, action_seed=15, hold_fraction=0.3,
              positions=[-3, -2, -1, 0, 1, 2, 3], windows=64, initial_position=0,
              max_episode_duration=100, max_episodes=8, **common)
    # termination: volatile series + extreme leverage so that the 0.7 stop (environments.py:246) fires
    vol = gte.make_gbm_ohlcv(800, seed=3, sigma=0.03)
    make_case("termination_stop", [vol], n_envs=8, n_steps=250, action_seed=16,
              positions=[-5, 0, 5], windows=None, initial_position="random",
              max_episode_duration="max", max_episodes=64, trading_fees=0.001,
              borrow_interest_rate=0.0001, portfolio_initial_value=1000)
    # no dynamic features (dynamic_feature_functions=[])
    make_case("no_dynamic_features", [gbm], n_envs=4, n_steps=120, action_seed=17,
              positions=[0, 1], windows=16, initial_position="random", max_episode_duration=30,
              dynamic_features=False, max_episodes=16, trading_fees=0.0, borrow_interest_rate=0.0,
              portfolio_initial_value=1000)
    # real data: bundled BTC/USD hourly CSV, example recipe; caller config of luckymodel/envs/env.py:67-82
    btc = btc_frame(2500)
    make_case("btc_luckymodel_config", [btc], n_envs=4, n_steps=300, action_seed=18,
              positions=[0, 0.5, 1], windows=24, initial_position="random",
              max_episode_duration=200, max_episodes=8, **common)
    # real data: the example script's config (examples/example_environnement.py:30-43)
    make_case("btc_example_config", [btc], n_envs=4, n_steps=300, action_seed=19,
              positions=[-1, -0.5, 0, 0.5, 1, 1.5, 2], windows=5, initial_position="random",
              max_episode_duration=100, max_episodes=8, **common)
    # the reward callbacks the reference's callers use (fused device catalogue, "next" row of SURVEY.md §8f)
    for rname in REWARDS:
        make_case("reward_" + rname, [vol], n_envs=6, n_steps=160, action_seed=30, reward=rname,
                  positions=[-2, -1, 0, 1, 2], windows=None, initial_position="random",
                  max_episode_duration=60, max_episodes=16, **common)
    # persistent limit orders (add_limit_order, environments.py:217-231), re-added after every reset; a hold-heavy
    # action stream so that the orders, not the actions, move the position most of the time
    make_case("limit_orders_persistent", [vol], n_envs=6, n_steps=200, action_seed=31, hold_fraction=0.7,
              positions=[-1, 0, 1, 2], windows=4, initial_position="random", max_episode_duration=60, max_episodes=16,
              limit_orders=[(2, 0.99), (-1, 1.012), (0, 1.003)], **common)
    # MultiDatasetTradingEnv: ragged datasets, least-used rotation, switch every episode / every 3 episodes
    multi = [gte.make_gbm_ohlcv(T, seed=20 + k) for k, T in enumerate([300, 420, 360, 500])]
    make_case("multi_dataset_k1", multi, n_envs=6, n_steps=260, action_seed=20,
              positions=[-1, 0, 1, 2], windows=4, initial_position="random", max_episode_duration=30,
              multi_dataset=True, episodes_between_dataset_switch=1, max_episodes=16, **common)
    make_case("multi_dataset_k3", multi, n_envs=6, n_steps=260, action_seed=21,
              positions=[-1, 0, 1, 2], windows=4, initial_position="random", max_episode_duration=30,
              multi_dataset=True, episodes_between_dataset_switch=3, max_episodes=16, **common)


if __name__ == "__main__":
    main()
