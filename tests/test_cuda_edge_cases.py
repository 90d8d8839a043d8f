"""GPU: edge shapes of the path against the oracle — single env, env counts that are not multiples of the
tile / group / CTA sizes, tiny and odd windows (every gather variant and the fallbacks), one static
feature, the maximum number of positions and datasets, a dataset barely longer than one episode."""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu


def _run(series_list, n_envs, K, *, positions, windows, duration, multi=False, k_switch=1, dyn=True, variant="auto", seed=5):
    import gym_trading_env_b200 as gte
    import oracle as orc
    kw = dict(positions=positions, windows=windows, trading_fees=1e-4, borrow_interest_rate=3e-6,
              portfolio_initial_value=1000, max_episode_duration=duration)
    tmax = max(s.length for s in series_list)
    ns = series_list[0].features.shape[1]
    feats = np.zeros((len(series_list), tmax, ns), np.float32)
    price = np.ones((len(series_list), tmax))
    for k, s in enumerate(series_list):
        feats[k, :s.length], price[k, :s.length] = s.features, s.price
    lens = np.array([s.length for s in series_list])
    dkw = dict(kw, num_envs=n_envs, seed=seed, verbose=0, obs_variant=variant, debug_outputs=True)
    if not dyn:
        dkw["dynamic_feature_functions"] = []
    if multi:
        dev = gte.MultiDatasetTradingVectorEnv(datasets=series_list, episodes_between_dataset_switch=k_switch, **dkw)
    else:
        dev = gte.TradingVectorEnv(series_list[0], **dkw)
    o = orc.OracleVecEnv(feats, price, lens, num_envs=n_envs, seed=seed, multi_dataset=multi,
                         episodes_between_dataset_switch=k_switch, dynamic_features=dyn, **kw)
    obs, _ = dev.reset()
    H.assert_bits(obs.cpu().numpy(), o.reset(), "reset obs")
    rng = np.random.default_rng(seed)
    for k in range(K):
        a = rng.integers(-1, len(positions), size=n_envs)          # -1 = hold
        dev.step(torch.as_tensor(a, device=dev.device))
        o.step(a)
        H.assert_bits(dev._obs.cpu().numpy(), o.obs, f"step {k} obs")
        H.assert_bits(dev._valuation.cpu().numpy(), o.valuation, f"step {k} valuation")
        H.assert_bits(dev._terminated.cpu().numpy(), o.terminated, f"step {k} terminated")
        H.assert_bits(dev._truncated.cpu().numpy(), o.truncated, f"step {k} truncated")
        H.assert_bits(dev._ep_start.cpu().numpy(), o.ep_start, f"step {k} ep_start")
        H.assert_bits(dev._dataset_idx.cpu().numpy(), o.dataset_idx, f"step {k} dataset")
        H.assert_close64(dev._reward.cpu().numpy(), o.reward, f"step {k} reward")
        assert dev._metrics_step[0].item() == o.metrics[0]
    dev.check_errors()
    return dev


def _series(T, seed, n_features=8):
    import gym_trading_env_b200 as gte
    s = gte.frame_to_arrays(gte.make_gbm_ohlcv(T, seed=seed))
    if n_features != 8:
        s = gte.SeriesArrays(np.ascontiguousarray(s.features[:, :n_features]), s.price, s.feature_names[:n_features], {}, None)
    return s


@pytest.mark.parametrize("n_envs", [1, 3, 31, 33, 127, 257, 1000])
def test_env_counts_off_the_tile_sizes(n_envs):
    dev = _run([_series(900, 1)], n_envs, 70, positions=[-1, 0, 1, 2], windows=16, duration=20)
    assert dev.obs_variant == "tma"


@pytest.mark.parametrize("windows,n_features,variant", [
    (2, 8, "auto"), (2, 8, "generic"), (4, 2, "vec"), (6, 8, "tma"),      # tiny windows
    (5, 5, "auto"), (7, 3, "auto"), (9, 8, "auto"),                        # odd windows / row sizes -> vec or generic
    (128, 8, "auto"), (200, 8, "auto"), (1, 8, "auto"),                    # long windows (smaller TMA groups), W=1
    (64, 1, "auto"), (64, 14, "auto"),                                     # 1 and 14 static features
])
def test_window_and_feature_shapes(windows, n_features, variant):
    _run([_series(1200, 2, n_features)], 300, 45, positions=[-2, 0, 0.5, 2], windows=windows, duration=30, variant=variant)


def test_no_dynamic_features_all_variants():
    for variant in ("tma", "vec", "generic"):
        _run([_series(800, 3)], 200, 50, positions=[0, 1], windows=8, duration=25, dyn=False, variant=variant)


def test_maximum_positions_and_datasets():
    pos = [round(-3 + 6 * i / 63, 6) for i in range(64)]                   # GTE_MAX_POSITIONS
    _run([_series(700, 4)], 150, 40, positions=pos, windows=8, duration=20)
    many = [_series(260 + 7 * k, 100 + k) for k in range(64)]              # GTE_MAX_DATASETS, ragged lengths
    dev = _run(many, 256, 90, positions=[-1, 0, 1], windows=4, duration=12, multi=True, k_switch=1)
    assert len(set(dev._dataset_idx.cpu().numpy().tolist())) > 30


def test_dataset_barely_longer_than_an_episode_and_max_duration():
    # T - D - 2(W-1) = 1: every episode starts at the same row; 'max': episodes run to the end of the data
    _run([_series(60, 6)], 64, 80, positions=[0, 1], windows=8, duration=45)
    _run([_series(80, 7)], 64, 200, positions=[-1, 1], windows=8, duration="max")
    _run([_series(50, 8)], 17, 120, positions=[-1, 1], windows=None, duration="max")


@pytest.mark.parametrize("variant,n_envs", [("tma", 1000), ("tma", 37), ("vec", 333), ("generic", 333), ("nowindow", 333)])
def test_no_write_lands_outside_the_output_and_state_buffers(variant, n_envs):
    """Every buffer the kernels write is re-homed between 4 KiB guard bands filled with a byte pattern;
    after resets and steps with auto-resets the bands must be untouched (compute-sanitizer is closed on
    this GPU pool, so this is the out-of-bounds check)."""
    import gym_trading_env_b200 as gte
    windows = None if variant == "nowindow" else 16
    env = gte.TradingVectorEnv(_series(700, 11), positions=[-1, 0, 1, 2], windows=windows, trading_fees=1e-4,
                               borrow_interest_rate=3e-6, max_episode_duration=15, num_envs=n_envs, seed=3, verbose=0,
                               debug_outputs=True, obs_variant="auto" if variant == "nowindow" else variant)
    GUARD = 4096
    bands = []
    names = ["_obs", "_dyn_ring", "_ring_clock", "_reward", "_terminated", "_truncated", "_valuation", "_real_position", "_info_idx",
             "_info_step", "_pre_reset_portfolio", "_asset", "_fiat", "_interest_asset", "_interest_fiat", "_pos_idx",
             "_step", "_ep_start", "_dataset_idx", "_plan_cursor", "_ds_used", "_ds_episodes", "_metrics_step",
             "_metrics_total", "_metric_partials", "_error_flag", "_tick_dev", "_block_counter"]
    for name in names:
        t = getattr(env, name)
        nbytes = t.numel() * t.element_size()
        pad = (-nbytes) % 16
        big = torch.full((GUARD + nbytes + pad + GUARD,), 0xA5, dtype=torch.uint8, device=env.device)
        view = big[GUARD:GUARD + nbytes].view(t.dtype).view(t.shape)
        view.copy_(t)
        setattr(env, name, view)
        bands.append((name, big, nbytes))
    env._build_structs()
    env.reset()
    rng = np.random.default_rng(0)
    for _ in range(60):
        env.step(torch.as_tensor(rng.integers(0, 4, size=n_envs), device=env.device))
    env.infos["portfolio_valuation"]
    torch.cuda.synchronize()
    assert float(env.get_metrics()["episodes"].item()) > n_envs          # auto-resets happened
    for name, big, nbytes in bands:
        assert bool((big[:GUARD] == 0xA5).all()), f"write below {name}"
        assert bool((big[GUARD + nbytes:] == 0xA5).all()), f"write above {name}"


def test_long_soak_against_the_oracle_with_every_episode_end_kind():
    """1 500 lockstep iterations of 2 048 envs on a volatile series with extreme leverage: thousands of
    valuation stops, duration truncations and end-of-data truncations, every one followed by an in-kernel
    Philox reset — compared with the oracle every 25 iterations (state bit-exact implies the steps between)."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(400, seed=21, sigma=0.02))
    pos = [-4, -1, 0, 1, 4]
    kw = dict(positions=pos, windows=16, trading_fees=5e-4, borrow_interest_rate=5e-5,
              portfolio_initial_value=1000, max_episode_duration="max")
    N, K = 2048, 1500
    dev = gte.TradingVectorEnv(arr, num_envs=N, seed=17, verbose=0, **kw)
    o = orc.OracleVecEnv(arr.features, arr.price, num_envs=N, seed=17, **kw)
    dev.reset()
    o.reset()
    g = torch.Generator(device=dev.device)
    g.manual_seed(5)
    acts = torch.randint(0, len(pos), (K, N), generator=g, device=dev.device, dtype=torch.int64)
    acts_h = acts.cpu().numpy()
    c = lambda t: t.cpu().numpy()   # noqa: E731
    for k in range(K):
        dev.step(acts[k])
        o.step(acts_h[k], want_obs=(k % 25 == 24))
        if k % 25 == 24:
            H.assert_bits(c(dev._obs), o.obs, f"step {k} obs")
            for nm in ("asset", "fiat", "interest_asset", "interest_fiat"):
                H.assert_bits(c(getattr(dev, "_" + nm)), getattr(o, nm), f"step {k} {nm}")
            H.assert_bits(c(dev._ep_start), o.ep_start, f"step {k} ep_start")
            H.assert_bits(c(dev._step), o.step_, f"step {k} step")
            H.assert_bits(c(dev._pos_idx), o.pos_idx, f"step {k} pos_idx")
            H.assert_bits(c(dev._valuation), o.valuation, f"step {k} valuation")
    m = c(dev._metrics_total)
    assert m[1] > 1000 and m[2] > 50 and m[0] <= m[1] + m[2]    # valuation stops AND end-of-data truncations occurred
    dev.check_errors()


@pytest.mark.parametrize("variant", ["tma", "vec", "generic"])
def test_masked_reset_in_the_middle_of_a_run_matches_the_oracle(variant):
    """reset(options={"mask": ...}) restarts only the chosen envs: their windows restart (zero rows before the new
    episode start), everybody else's dynamic-feature history is untouched — with the time-indexed ring a reset env's
    first row lands in the slot of the CURRENT iteration."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    s = _series(900, 21)
    n, positions = 203, [-2, -1, 0, 1, 2]
    kw = dict(positions=positions, windows=12, trading_fees=1e-4, borrow_interest_rate=3e-6,
              portfolio_initial_value=1000, max_episode_duration=50)
    dev = gte.TradingVectorEnv(s, num_envs=n, seed=9, verbose=0, obs_variant=variant, debug_outputs=True, **kw)
    o = orc.OracleVecEnv(s.features[None], s.price[None], np.array([s.length]), num_envs=n, seed=9, **kw)
    obs, _ = dev.reset()
    H.assert_bits(obs.cpu().numpy(), o.reset(), "reset obs")
    rng = np.random.default_rng(1)
    for k in range(90):
        if k in (7, 8, 30, 61):                                     # also two resets in consecutive iterations
            mask = rng.random(n) < 0.3
            obs, _ = dev.reset(options={"mask": mask})
            H.assert_bits(obs.cpu().numpy(), o.reset(mask=mask), f"masked reset before step {k}")
        a = rng.integers(-1, len(positions), size=n)
        dev.step(torch.as_tensor(a, device=dev.device))
        o.step(a)
        H.assert_bits(dev._obs.cpu().numpy(), o.obs, f"step {k} obs")
        H.assert_bits(dev._valuation.cpu().numpy(), o.valuation, f"step {k} valuation")
        H.assert_bits(dev._ep_start.cpu().numpy(), o.ep_start, f"step {k} ep_start")
    dev.check_errors()


def test_checkpoint_round_trip_resumes_bit_identically():
    """state_dict() / load_state_dict(): portfolio state, dynamic-feature ring + its clock, rotation state and the
    Philox tick — a restored env produces the same observations, rewards and random restarts as the original."""
    import gym_trading_env_b200 as gte
    series = [_series(T, 30 + k) for k, T in enumerate([400, 520, 460])]
    kw = dict(positions=[-1, 0, 1, 2], windows=16, trading_fees=1e-4, borrow_interest_rate=3e-6,
              max_episode_duration=25, num_envs=300, seed=4, verbose=0, episodes_between_dataset_switch=2)
    a = gte.MultiDatasetTradingVectorEnv(datasets=series, **kw)
    b = gte.MultiDatasetTradingVectorEnv(datasets=series, **kw)
    a.reset(); b.reset()
    g = torch.Generator(device=a.device); g.manual_seed(3)
    acts = torch.randint(0, 4, (80, 300), generator=g, device=a.device)
    for k in range(37):
        a.step(acts[k])
    snap = a.state_dict()
    obs_at_snap = a._obs.clone()
    b.load_state_dict(snap)
    b._launch_obs()
    assert torch.equal(b._obs.view(torch.int32), obs_at_snap.view(torch.int32))
    for k in range(37, 80):
        oa, ra, ta, tra, _ = a.step(acts[k])
        ob, rb, tb, trb, _ = b.step(acts[k])
        assert torch.equal(oa.view(torch.int32), ob.view(torch.int32)), k
        assert torch.equal(ra, rb) and torch.equal(ta, tb) and torch.equal(tra, trb), k
        assert torch.equal(a._dataset_idx, b._dataset_idx) and torch.equal(a._ep_start, b._ep_start), k
    assert torch.equal(a._metrics_total, b._metrics_total)
    assert float(a._metrics_total[0]) > 300                          # several random restarts were replayed identically
