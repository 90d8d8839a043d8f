"""GPU: the CUDA path, called through the C-ABI (ctypes -> libgte_b200.so), against
(a) the golden vectors recorded from the reference's own code (with the H3 normalisation of oracle/ref_harness.py:
dynamic-feature columns zeroed before each reset) and (b) the CPU oracle on larger
seeded inputs, with the same step-by-step comparison the oracle itself passes on the CPU.

Bars (BASELINE.json north_star): position/step indices, flags and observations bit-exact;
valuation and reward within 1e-12 relative (helpers.RTOL) — and, stronger, every fp64 portfolio
quantity is also required to be BIT-exact (only `log` may differ, by <= 1 ulp)."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

DEVICE_GOLDEN = [n for n in H.GOLDEN_NAMES if n != "raw_stale_dynamic_rows"]   # raw H3 leak: oracle-only


def _variants(g):
    p = g["params"]
    F = g["features"].shape[2] + (2 if p.get("dynamic_features", True) else 0)
    if p["windows"] is not None and (p["windows"] * F * 4) % 16 == 0:
        return ["generic", "vec", "tma"]
    return ["generic"]


@pytest.mark.parametrize("name", DEVICE_GOLDEN)
def test_cuda_matches_reference_golden(name):
    g = H.load_golden(name)
    for variant in _variants(g):
        env = H.make_device_env(g, obs_variant=variant)
        stats = H.replay_golden(H.DeviceAdapter(env), g, exact_money=True)
        env.check_errors()
        assert stats["steps"] == g["actions"].size
        assert stats["reward_bit_mismatch"] <= 0.25 * stats["steps"], variant


def _compare_with_oracle(dev, orc_env, K, n_pos, seed, hold=0.05):
    import torch
    rng = np.random.default_rng(seed)
    obs_d, _ = dev.reset()
    obs_o = orc_env.reset()
    H.assert_bits(obs_d.cpu().numpy(), obs_o, "reset obs")
    c = lambda t: t.cpu().numpy()   # noqa: E731
    total_eps = 0
    for k in range(K):
        a = rng.integers(0, n_pos, size=dev.num_envs)
        a[rng.random(dev.num_envs) < hold] = -1
        dev.step(torch.as_tensor(a, device=dev.device))
        orc_env.step(a)
        w = f"step {k}"
        H.assert_bits(c(dev._obs), orc_env.obs, f"{w} obs")
        H.assert_bits(c(dev._terminated), orc_env.terminated, f"{w} terminated")
        H.assert_bits(c(dev._truncated), orc_env.truncated, f"{w} truncated")
        H.assert_bits(c(dev._info_idx), orc_env.info_idx, f"{w} idx")
        H.assert_bits(c(dev._info_step), orc_env.info_step, f"{w} step")
        H.assert_bits(c(dev._valuation), orc_env.valuation, f"{w} valuation")
        H.assert_bits(c(dev._real_position), orc_env.real_position, f"{w} real_position")
        H.assert_close64(c(dev._reward), orc_env.reward, f"{w} reward")
        for nm in ("asset", "fiat", "interest_asset", "interest_fiat"):
            H.assert_bits(c(getattr(dev, "_" + nm)), getattr(orc_env, nm), f"{w} {nm}")
        H.assert_bits(c(dev._pos_idx), orc_env.pos_idx, f"{w} pos_idx")
        H.assert_bits(c(dev._ep_start), orc_env.ep_start, f"{w} ep_start")
        H.assert_bits(c(dev._step), orc_env.step_, f"{w} step state")
        H.assert_bits(c(dev._dataset_idx), orc_env.dataset_idx, f"{w} dataset")
        m = c(dev._metrics_step)
        assert m[0] == orc_env.metrics[0] and m[1] == orc_env.metrics[1] and m[2] == orc_env.metrics[2]
        assert m[5] == orc_env.metrics[5]
        np.testing.assert_allclose(m[3:5], orc_env.metrics[3:5], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(m[6], orc_env.metrics[6], rtol=1e-9, atol=1e-11)
        total_eps += int(m[0])
    np.testing.assert_allclose(c(dev._metrics_total)[0], total_eps)
    dev.check_errors()
    return total_eps


@pytest.mark.parametrize("variant", ["tma", "vec", "generic"])
def test_cuda_matches_oracle_philox_resets_c3_shape(variant):
    """C3/C5 shape at reduced N: W=64, 8 static + 2 dynamic features, leveraged positions, D=100,
    auto-reset with in-kernel Philox starts (no plan) — the oracle draws from the same stream."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    df = gte.make_gbm_ohlcv(6000, seed=5)
    arr = gte.frame_to_arrays(df)
    pos = [-3, -2, -1, 0, 1, 2, 3]
    kw = dict(positions=pos, windows=64, trading_fees=1e-4, borrow_interest_rate=3e-6,
              portfolio_initial_value=1000, max_episode_duration=100)
    N = 3000     # not a multiple of the CTA size: exercises the ragged tail
    dev = gte.TradingVectorEnv(arr, num_envs=N, seed=77, env_id_offset=12345, obs_variant=variant, verbose=0, debug_outputs=True, **kw)
    o = orc.OracleVecEnv(arr.features, arr.price, num_envs=N, seed=77, env_id_offset=12345, **kw)
    eps = _compare_with_oracle(dev, o, K=230, n_pos=len(pos), seed=1)
    assert eps >= 2 * N


def test_cuda_matches_oracle_windows_none_c2_shape():
    import gym_trading_env_b200 as gte
    import oracle as orc
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(5000, seed=6))
    pos = [-1, 0, 0.5, 1]
    kw = dict(positions=pos, windows=None, trading_fees=1e-4, borrow_interest_rate=3e-6,
              portfolio_initial_value=1000, max_episode_duration="max")
    dev = gte.TradingVectorEnv(arr, num_envs=4096, seed=3, verbose=0, debug_outputs=True, **kw)
    o = orc.OracleVecEnv(arr.features, arr.price, num_envs=4096, seed=3, **kw)
    _compare_with_oracle(dev, o, K=120, n_pos=len(pos), seed=2)


def test_cuda_matches_oracle_multi_dataset_rotation():
    """C4 shape at reduced size: ragged datasets, per-env dataset index, in-kernel least-used rotation."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    lens = [900, 1200, 1000, 1500, 800]
    series = [gte.frame_to_arrays(gte.make_gbm_ohlcv(T, seed=40 + k)) for k, T in enumerate(lens)]
    tmax = max(lens)
    feats = np.zeros((len(lens), tmax, 8), np.float32)
    price = np.ones((len(lens), tmax))
    for k, s in enumerate(series):
        feats[k, :s.length], price[k, :s.length] = s.features, s.price
    pos = [-1, 0, 1, 2]
    for k_switch in (1, 2):
        kw = dict(positions=pos, windows=16, trading_fees=1e-4, borrow_interest_rate=3e-6,
                  portfolio_initial_value=1000, max_episode_duration=25)
        dev = gte.MultiDatasetTradingVectorEnv(datasets=series, episodes_between_dataset_switch=k_switch,
                                               num_envs=1500, seed=9, verbose=0, debug_outputs=True, **kw)
        o = orc.OracleVecEnv(feats, price, np.array(lens), num_envs=1500, seed=9, multi_dataset=True,
                             episodes_between_dataset_switch=k_switch, **kw)
        _compare_with_oracle(dev, o, K=150, n_pos=len(pos), seed=3)
        used = np.bincount(dev._dataset_idx.cpu().numpy(), minlength=len(lens))
        assert (used > 0).all()


def test_termination_rule_and_threshold_parameter():
    """H1: done when valuation/initial <= 0.7 (this fork, environments.py:246); 0.0 = upstream rule."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(1500, seed=3, sigma=0.03))
    pos = [-5, 0, 5]
    kw = dict(positions=pos, windows=None, trading_fees=1e-3, borrow_interest_rate=1e-4,
              portfolio_initial_value=1000, max_episode_duration="max")
    for ratio in (0.7, 0.0):
        dev = gte.TradingVectorEnv(arr, num_envs=512, seed=1, done_valuation_ratio=ratio, verbose=0,
                                   debug_outputs=True, **kw)
        o = orc.OracleVecEnv(arr.features, arr.price, num_envs=512, seed=1, done_ratio=ratio, **kw)
        _compare_with_oracle(dev, o, K=200, n_pos=3, seed=4, hold=0.0)
        term = float(dev.get_metrics()["terminated"].item())
        assert (term > 50) if ratio == 0.7 else True


def test_cabi_argument_errors_and_host_validation():
    import gym_trading_env_b200 as gte
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(500, seed=0))
    env = gte.TradingVectorEnv(arr, positions=[0, 1], windows=8, num_envs=16, verbose=0)
    env.reset()
    env_np = gte.TradingVectorEnv(arr, positions=[0, 1], windows=8, num_envs=16, verbose=0, output="numpy")
    env_np.reset()
    with pytest.raises(IndexError):
        env_np.step(np.full(16, 2))                    # flagged in-kernel, raised when the results reach the host
    with pytest.raises(ValueError):
        env.step(np.zeros(3, dtype=np.int64))
    import torch
    assert int(env.error_flag) == 0
    env.step(torch.full((16,), 5, dtype=torch.int64, device=env.device))   # device actions: flagged in-kernel
    assert int(env.error_flag) == 1                                        # refreshed by the step itself, no extra launch
    with pytest.raises(IndexError):
        env.check_errors()
    with pytest.raises(NotImplementedError):
        gte.TradingVectorEnv(arr, reward_function=lambda h: 0.0, num_envs=2)
    with pytest.raises(AssertionError):
        gte.TradingVectorEnv(arr, positions=[0, 1], initial_position=0.5, num_envs=2)
    with pytest.raises(ValueError):
        gte.TradingVectorEnv(arr, windows=400, max_episode_duration=200, num_envs=2)
    # the C entry points reject inconsistent structs instead of launching
    import ctypes as C
    from gym_trading_env_b200 import _cabi
    bad = _cabi.GteParams()
    C.memmove(C.byref(bad), C.byref(env._P), C.sizeof(bad))
    bad.n_positions = 1000
    rc = env._lib.gte_reset(C.byref(bad), C.byref(env._D), C.byref(env._S), None, 0, 0, None)
    assert rc == -1 and b"n_positions" in env._lib.gte_last_error()


def test_numpy_output_mode_and_infos():
    import gym_trading_env_b200 as gte
    import oracle as orc
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(2000, seed=2))
    kw = dict(positions=[-1, 0, 1], windows=16, trading_fees=1e-4, borrow_interest_rate=3e-6,
              portfolio_initial_value=1000, max_episode_duration=50)
    env = gte.TradingVectorEnv(arr, num_envs=256, seed=5, output="numpy", verbose=0, **kw)
    o = orc.OracleVecEnv(arr.features, arr.price, num_envs=256, seed=5, **kw)
    obs, infos = env.reset()
    assert isinstance(obs, np.ndarray) and obs.dtype == np.float32 and obs.shape == (256, 16, 10)
    H.assert_bits(obs, o.reset(), "reset obs (numpy)")
    assert np.array_equal(infos["portfolio_valuation"].cpu().numpy(), np.full(256, 1000.0))
    rng = np.random.default_rng(0)
    for _ in range(60):
        a = rng.integers(0, 3, size=256)
        obs, rew, term, trunc, infos = env.step(a)
        o.step(a)
        H.assert_bits(obs, o.obs, "obs (numpy)")
        assert term.dtype == np.bool_ and rew.dtype == np.float64
        H.assert_close64(rew, o.reward, "reward")
    # infos are History's last row of the CURRENT (post-reset) state
    H.assert_bits(infos["idx"].cpu().numpy(), o.idx, "infos idx")
    H.assert_bits(infos["position_index"].cpu().numpy(), o.pos_idx, "infos position_index")
    dist = np.stack([infos[f"portfolio_distribution_{k}"].cpu().numpy() for k in
                     ("asset", "fiat", "borrowed_asset", "borrowed_fiat", "interest_asset", "interest_fiat")])
    assert np.array_equal(dist[0] - dist[2], o.asset) and np.array_equal(dist[1] - dist[3], o.fiat)
    assert np.array_equal(dist[4], o.interest_asset) and np.array_equal(dist[5], o.interest_fiat)


@pytest.mark.parametrize("mode", ["chunks3", "graph", "graph_chunks", "hybrid"])
def test_step_obs_pipelined_graph_and_hybrid_modes_match_oracle(mode):
    """gte_step_obs with env-range pipelining over two streams, CUDA-graph replay (device-resident
    Philox tick) and the hybrid host/device output mode must all produce the plain two-launch results."""
    import torch
    import gym_trading_env_b200 as gte
    import oracle as orc
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(4000, seed=8))
    pos = [-2, -1, 0, 1, 2]
    kw = dict(positions=pos, windows=32, trading_fees=1e-4, borrow_interest_rate=3e-6,
              portfolio_initial_value=1000, max_episode_duration=40)
    N = 2500
    extra = {"chunks3": dict(n_chunks=3), "graph": dict(cuda_graph=True),
             "graph_chunks": dict(cuda_graph=True, n_chunks=4), "hybrid": dict(output="hybrid")}[mode]
    dev = gte.TradingVectorEnv(arr, num_envs=N, seed=21, verbose=0, **extra, **kw)
    o = orc.OracleVecEnv(arr.features, arr.price, num_envs=N, seed=21, **kw)
    obs, _ = dev.reset()
    H.assert_bits(obs.cpu().numpy(), o.reset(), "reset obs")
    rng = np.random.default_rng(5)
    for k in range(130):
        a = rng.integers(0, len(pos), size=N)
        if mode == "hybrid":
            obs, rew, term, trunc, _ = dev.step(a)
            assert isinstance(rew, np.ndarray) and isinstance(obs, torch.Tensor)
            rew, term, trunc = rew.copy(), term.copy(), trunc.copy()
        else:
            obs, rew, term, trunc, _ = dev.step(torch.as_tensor(a, device=dev.device))
            rew, term, trunc = rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        o.step(a)
        H.assert_bits(obs.cpu().numpy(), o.obs, f"step {k} obs ({mode})")
        H.assert_bits(term, o.terminated.astype(bool), f"step {k} terminated")
        H.assert_bits(trunc, o.truncated.astype(bool), f"step {k} truncated")
        H.assert_close64(rew, o.reward, f"step {k} reward")
        H.assert_bits(dev._valuation.cpu().numpy(), o.valuation, f"step {k} valuation")
        H.assert_bits(dev._ep_start.cpu().numpy(), o.ep_start, f"step {k} ep_start (Philox tick)")
        m = dev._metrics_step.cpu().numpy()
        assert m[0] == o.metrics[0] and m[5] == o.metrics[5]
        np.testing.assert_allclose(m[6], o.metrics[6], rtol=1e-9, atol=1e-11)
    assert int(dev._tick_dev.item()) == 2 + 130
    dev.check_errors()


def test_reward_catalogue_rejects_python_callbacks_and_accepts_device_functors():
    import gym_trading_env_b200 as gte
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(500, seed=0))
    with pytest.raises(NotImplementedError):
        gte.TradingVectorEnv(arr, reward_function=lambda h: 0.0, num_envs=2)
    with pytest.raises(ValueError):
        gte.DeviceReward("sharpe")
    env = gte.TradingVectorEnv(arr, reward_function=gte.log_return_reward(scale=100.0, clip=(-0.2, 0.5)), num_envs=8, verbose=0)
    env.reset()
    _, rew, *_ = env.step(np.ones(8, dtype=np.int64))
    assert float(rew.max()) <= 0.5 and float(rew.min()) >= -0.2


def test_tracked_history_reproduces_the_reference_history_rows():
    """The device-side History log of a tracked env against the golden C1 run:
    the rows step() adds and the rows reset() writes, including across an end-of-data truncation."""
    g = H.load_golden("c1_single_nowindow")
    env = H.make_device_env(g)
    env.track([0], max_steps=4000)
    env.reset()
    import torch
    K = g["actions"].shape[0]
    for k in range(K):
        env.step(torch.as_tensor(g["actions"][k], device=env.device))
    df = env.tracked_history(0)
    steps = df[~df["new_episode"]].reset_index(drop=True)
    assert len(steps) == K and int(df["new_episode"].sum()) == g["plan"].shape[1]
    H.assert_bits(steps["idx"].to_numpy(np.int32), g["idx"][:, 0], "history idx")
    H.assert_bits(steps["step"].to_numpy(np.int32), g["step"][:, 0], "history step")
    H.assert_bits(steps["portfolio_valuation"].to_numpy(), g["valuation"][:, 0], "history valuation")
    H.assert_bits(steps["position"].to_numpy(), g["position"][:, 0], "history position")
    H.assert_bits(steps["real_position"].to_numpy(), g["real_position"][:, 0], "history real_position")
    H.assert_bits(steps["position_index"].to_numpy(), g["actions"][:, 0], "history position_index (raw action)")
    H.assert_close64(steps["reward"].to_numpy(), g["reward"][:, 0], "history reward")
    a, f = g["asset"][:, 0], g["fiat"][:, 0]
    H.assert_bits(steps["portfolio_distribution_asset"].to_numpy(), np.maximum(a, 0) + 0.0, "distribution asset")
    H.assert_bits(steps["portfolio_distribution_borrowed_fiat"].to_numpy(), np.maximum(-f, 0) + 0.0, "distribution borrowed fiat")
    H.assert_bits(steps["data_close"].to_numpy(), g["price"][0][g["idx"][:, 0]], "data_close")
    starts = df[df["new_episode"]]
    assert (starts["portfolio_valuation"] == 1000.0).all() and (starts["step"] == 0).all()
    assert starts["idx"].tolist() == g["plan"][0, :, 0].tolist()


def test_final_obs_mode_keeps_the_terminal_observation():
    """gymnasium SAME_STEP semantics (hazard H8): with final_obs=True the observation step() itself returned
    for an ended episode (golden `step_obs`) is kept in env.final_obs, obs is the reset observation."""
    import torch
    g = H.load_golden("c3_windows_leveraged")
    env = H.make_device_env(g, final_obs=True)
    obs, _ = env.reset()
    H.assert_bits(obs.cpu().numpy(), g["obs0"], "reset obs")
    seen = 0
    for k in range(g["actions"].shape[0]):
        obs, rew, term, trunc, _ = env.step(torch.as_tensor(g["actions"][k], device=env.device))
        H.assert_bits(obs.cpu().numpy(), g["obs"][k], f"step {k} obs (post-reset)")
        ended = (g["terminated"][k] | g["truncated"][k]).astype(bool)
        H.assert_bits(term.cpu().numpy() | trunc.cpu().numpy(), ended, f"step {k} ended")
        if ended.any():
            H.assert_bits(env.final_obs.cpu().numpy()[ended], g["step_obs"][k][ended], f"step {k} final_obs")
            seen += int(ended.sum())
        H.assert_bits(env._valuation.cpu().numpy(), g["valuation"][k], f"step {k} valuation")
        H.assert_bits(env.idx.cpu().numpy(), g["post_idx"][k], f"step {k} post idx")
    assert seen >= 30


@pytest.mark.parametrize("windows", [8, None])
def test_rollout_equals_stepping_one_by_one(windows):
    """rollout() = ONE gte_rollout call enqueueing K iterations: same results as K step() calls, with every
    observation kept (keep_obs) and with only the last one gathered."""
    import torch
    import gym_trading_env_b200 as gte
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(1500, seed=4))
    kw = dict(positions=[-1, 0, 1], windows=windows, trading_fees=1e-4, max_episode_duration=30, num_envs=500, seed=2, verbose=0)
    a, b, c = gte.TradingVectorEnv(arr, **kw), gte.TradingVectorEnv(arr, **kw), gte.TradingVectorEnv(arr, **kw)
    a.reset(); b.reset(); c.reset()
    g = torch.Generator(device=a.device); g.manual_seed(0)
    acts = torch.randint(0, 3, (40, 500), generator=g, device=a.device)
    out = a.rollout(acts, keep_obs=True)
    last = c.rollout(acts[:25])                              # observation of the last iteration only ...
    last2 = c.rollout(acts[25:])                             # ... and the ring stays current across calls
    for k in range(40):
        obs, rew, term, trunc, _ = b.step(acts[k])
        assert torch.equal(out["obs"][k].view(torch.int32), obs.view(torch.int32))
        assert torch.equal(out["reward"][k], rew) and torch.equal(out["terminated"][k], term) and torch.equal(out["truncated"][k], trunc)
        assert torch.equal(out["valuation"][k], b._valuation)
    assert torch.equal(torch.cat([last["reward"], last2["reward"]]), out["reward"])
    assert torch.equal(torch.cat([last["terminated"], last2["terminated"]]), out["terminated"])
    for e in (a, c):
        assert torch.equal(e._obs.view(torch.int32), b._obs.view(torch.int32))
        # counts are exact; the fp64 sums are folded in a fixed order PER LAUNCH SHAPE (one fused launch when the
        # observation is gathered, the plain step kernel when it is not), so they agree to rounding, not to the bit
        assert torch.equal(e._metrics_total[:3], b._metrics_total[:3]) and torch.equal(e._metrics_total[5], b._metrics_total[5])
        torch.testing.assert_close(e._metrics_total, b._metrics_total, rtol=1e-12, atol=1e-12)
        for name in ("_asset", "_fiat", "_pos_idx", "_step", "_ep_start", "_reward", "_terminated", "_valuation"):
            assert torch.equal(getattr(e, name), getattr(b, name)), name
    assert "obs" not in last and out["obs"].shape[0] == 40


def test_limit_order_api_validation():
    import gym_trading_env_b200 as gte
    env = gte.TradingVectorEnv(gte.make_gbm_ohlcv(400, seed=1), positions=[0, 1], num_envs=8, verbose=0)
    env.reset()
    with pytest.raises(NotImplementedError):
        env.add_limit_order(1, 100.0)                    # persistent=False: the reference itself raises when it executes
    with pytest.raises(ValueError):
        env.add_limit_order(0.5, 100.0, persistent=True)
    close0 = float(env.infos["data_close"][0])
    env.add_limit_order(1, close0, persistent=True)      # at the money: executes on the next bars
    for _ in range(5):
        env.step(np.full(8, -1))                         # hold: only the order can move the position
    assert int((env._pos_idx == 1).sum()) > 0
