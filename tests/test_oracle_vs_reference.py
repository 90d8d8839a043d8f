"""CPU, build container only: the C restatement against the UNMODIFIED reference executed live
(skipped where /root/reference is absent, e.g. on the GPU box — the committed goldens cover that)."""
import numpy as np
import pytest

import helpers as H
import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="/root/reference not mounted")


@pytest.mark.parametrize("windows,duration,normalize", [(None, "max", True), (8, 40, True), (8, 40, False), (16, 25, True)])
def test_oracle_matches_live_reference(windows, duration, normalize):
    import gym_trading_env_b200 as gte
    import oracle as orc
    df = gte.make_gbm_ohlcv(700, seed=9)
    arr = gte.frame_to_arrays(df)
    pos = [-2, -1, 0, 0.5, 1, 2]
    rng = np.random.default_rng(3)
    K, N = 260, 5
    acts = rng.integers(0, len(pos), size=(K, N))
    kw = dict(positions=pos, windows=windows, trading_fees=1e-4, borrow_interest_rate=3e-6,
              portfolio_initial_value=1000, max_episode_duration=duration)
    rec = rh.run_lockstep([df], N, acts, initial_position="random", normalize_dyn=normalize, np_seed=4,
                          max_episodes=32, **kw)
    env = orc.OracleVecEnv(arr.features, arr.price, num_envs=N, plan=rec["plan"], dyn_mode=1 if normalize else 0, **kw)
    H.assert_bits(env.reset(), rec["obs0"], "reset obs")
    for k in range(K):
        env.step(acts[k], want_final_obs=True)
        H.assert_bits(env.obs, rec["obs"][k], f"step {k} obs")
        H.assert_bits(env.final_obs, rec["step_obs"][k], f"step {k} terminal obs")
        H.assert_bits(env.valuation, rec["valuation"][k], f"step {k} valuation")
        H.assert_bits(env.terminated, rec["terminated"][k], f"step {k} terminated")
        H.assert_bits(env.truncated, rec["truncated"][k], f"step {k} truncated")
        H.assert_close64(env.reward, rec["reward"][k], f"step {k} reward")
        H.assert_bits(env.asset, rec["post_asset"][k], f"step {k} asset")
        H.assert_bits(env.idx, rec["post_idx"][k], f"step {k} idx")


def test_reference_quirks_the_survey_relies_on():
    """D-duration episode = D-1 steps; T-row dataset = T-1 steps; reset draws choice() then randint()."""
    import gym_trading_env_b200 as gte
    df = gte.make_gbm_ohlcv(120, seed=1)
    acts = np.zeros((130, 1), dtype=np.int64)
    rec = rh.run_lockstep([df], 1, acts, positions=[0, 1], windows=None, trading_fees=0, borrow_interest_rate=0,
                          portfolio_initial_value=1000, initial_position=0, max_episode_duration="max", max_episodes=8)
    first_end = int(np.argmax(rec["truncated"][:, 0]))
    assert first_end == 118 and rec["idx"][first_end, 0] == 119          # T-1 steps from idx 0
    rec = rh.run_lockstep([df], 1, acts[:40], positions=[0, 1], windows=None, trading_fees=0, borrow_interest_rate=0,
                          portfolio_initial_value=1000, initial_position=0, max_episode_duration=10, max_episodes=8)
    ends = np.flatnonzero(rec["truncated"][:, 0])
    assert ends[0] == 8 and np.all(np.diff(ends) == 9)                   # D-1 = 9 steps per episode
