"""GPU: oracle parity at BASELINE.json's FULL configuration sizes, and of the host-facing paths at scale.

  C4  : MultiDatasetTradingEnv semantics on 32 datasets x 1 000 000 rows with 2^20 envs, windows=64 — a block of 128
        envs inside the run against the CPU oracle (Philox resets, least-used dataset rotation, in-place auto-reset),
        plus the gathered windows of a sample of ALL envs against the rows of their own dataset's table;
  C4+ : the same check on 64 datasets x 1M rows, whose window tables (2.56 GB per copy) push the int64 table offsets
        past 2^31;
  C5  : a block of 128 envs INSIDE a 2^21-env shard against the oracle;
  f4  : the `dataset_dir` glob -> read_pickle -> preprocess loader against the multi-dataset goldens recorded from the
        reference's MultiDatasetTradingEnv (environments.py:365-400);
  host: the "hybrid" path (gte_step_host: copy-engine and zero-copy mechanisms, int8 / int64 wire actions) and the
        "numpy" path at 65 536 envs against the oracle.
Bars: observations / indices / flags / fp64 portfolio state and valuation bit-exact, reward within 1e-12 (helpers.RTOL).
"""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

FEES = dict(trading_fees=0.01 / 100, borrow_interest_rate=0.0003 / 100, portfolio_initial_value=1000)


def _acts(n, k, n_pos, device, seed=99):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randint(0, n_pos, (k, n), generator=g, device=device, dtype=torch.int64)


def _compare_block(env, o, lo, n_s, acts, K, what):
    """Step `env` (all envs, device actions) and the oracle (envs lo..lo+n_s) K times; compare the block every step."""
    obs, _ = env.reset()
    H.assert_bits(obs[lo:lo + n_s].cpu().numpy(), o.reset(), f"{what} reset obs")
    sl = slice(lo, lo + n_s)
    c = lambda t: t[sl].cpu().numpy()   # noqa: E731
    eps = 0
    for k in range(K):
        env.step(acts[k])
        o.step(acts[k][sl].cpu().numpy())
        w = f"{what} step {k}"
        H.assert_bits(c(env._obs), o.obs, f"{w} obs")
        H.assert_bits(c(env._terminated), o.terminated, f"{w} terminated")
        H.assert_bits(c(env._truncated), o.truncated, f"{w} truncated")
        H.assert_bits(c(env._info_idx), o.info_idx, f"{w} idx")
        H.assert_bits(c(env._info_step), o.info_step, f"{w} step")
        H.assert_bits(c(env._valuation), o.valuation, f"{w} valuation")
        H.assert_bits(c(env._real_position), o.real_position, f"{w} real_position")
        H.assert_close64(c(env._reward), o.reward, f"{w} reward")
        for nm in ("asset", "fiat", "interest_asset", "interest_fiat"):
            H.assert_bits(c(getattr(env, "_" + nm)), getattr(o, nm), f"{w} {nm}")
        H.assert_bits(c(env._pos_idx), o.pos_idx, f"{w} pos_idx")
        H.assert_bits(c(env._ep_start), o.ep_start, f"{w} ep_start")
        H.assert_bits(c(env._step), o.step_, f"{w} step state")
        H.assert_bits(c(env._dataset_idx), o.dataset_idx, f"{w} dataset")
        eps += int(o.metrics[0])
    env.check_errors()
    return eps


def _check_windows_against_tables(env, W, n_sample=32768):
    """P1 across datasets: the static columns of a sample of ALL envs' windows are rows idx-W+1..idx of THEIR dataset."""
    dev, N, ns = env.device, env.num_envs, env._n_static
    sel = torch.arange(0, N, max(1, N // n_sample), device=dev)
    idx = (env._ep_start + env._step).long()[sel]
    ds = env._dataset_idx.long()[sel]
    rows = idx[:, None] - (W - 1) + torch.arange(W, device=dev)[None, :]
    want = env._features[ds[:, None], rows]                               # [n_s, W, ns]
    got = env._obs[sel][:, :, :ns]
    assert torch.equal(got.view(torch.int32), want.view(torch.int32)), "static window != rows of the env's dataset"
    return ds


def _multi_series(n_ds, rows):
    import gym_trading_env_b200 as gte
    out = []
    for k in range(n_ds):
        f, p = gte.make_gbm_arrays(rows, seed=k)
        out.append(gte.SeriesArrays(f, p, [f"feature_{j}" for j in range(f.shape[1])], {}, None))
    return out


@pytest.mark.parametrize("n_ds,n_envs", [(32, 1 << 20), (64, 1 << 16)])
def test_c4_full_size_tables_match_the_oracle(n_ds, n_envs):
    """BASELINE config 4 at its full size (32 x 1M rows, 2^20 envs), and a 64 x 1M-row variant whose table offsets
    exceed 2^31 bytes: oracle block + table-row property over all datasets, with resets and rotation."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    rows, W, lo, n_s, K = 1_000_000, 64, n_envs - 4096 - 160, 128, 44
    series = _multi_series(n_ds, rows)
    pos = [-1, 0, 0.5, 1]
    kw = dict(positions=pos, windows=W, max_episode_duration=12, **FEES)
    env = gte.MultiDatasetTradingVectorEnv(datasets=series, episodes_between_dataset_switch=1, num_envs=n_envs,
                                           seed=5, verbose=0, debug_outputs=True, **kw)
    assert env.obs_variant == "tma"
    assert env._window_ds_stride * n_ds > (1 << 31) or n_ds == 32
    feats = np.stack([s.features for s in series])
    price = np.stack([s.price for s in series])
    del series
    o = orc.OracleVecEnv(feats, price, num_envs=n_s, seed=5, env_id_offset=lo, multi_dataset=True,
                         episodes_between_dataset_switch=1, **kw)
    acts = _acts(n_envs, K, len(pos), env.device)
    eps = _compare_block(env, o, lo, n_s, acts, K, f"C4 {n_ds}x1M")
    assert eps >= 3 * n_s                                               # every env restarted (and rotated) several times
    ds = _check_windows_against_tables(env, W)
    assert len(torch.unique(ds)) == n_ds                                # the sample really spans every dataset
    used = torch.bincount(env._dataset_idx.long(), minlength=n_ds)
    assert (used > 0).all() and used.max() < 2 * used.float().mean()    # least-used rotation keeps the spread even


def test_c5_shard_block_matches_the_oracle():
    """BASELINE config 5's per-GPU shard (2^21 envs, positions -3..3, windows=64, T=100 000): 128 envs inside the run
    against the oracle stepping exactly those global env ids, with resets (D=25 so that episodes end inside the run)."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    n_envs, lo, n_s, K = 1 << 21, 1_234_560, 128, 60
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(100_000, seed=0))
    pos = [-3, -2, -1, 0, 1, 2, 3]
    kw = dict(positions=pos, windows=64, max_episode_duration=25, **FEES)
    env = gte.TradingVectorEnv(series, num_envs=n_envs, seed=11, verbose=0, debug_outputs=True, **kw)
    o = orc.OracleVecEnv(series.features, series.price, num_envs=n_s, seed=11, env_id_offset=lo, **kw)
    eps = _compare_block(env, o, lo, n_s, _acts(n_envs, K, len(pos), env.device), K, "C5 shard")
    assert eps >= 2 * n_s


@pytest.mark.parametrize("name", ["multi_dataset_k1", "multi_dataset_k3"])
def test_dataset_dir_loader_matches_the_reference_goldens(name, tmp_path):
    """MultiDatasetTradingEnv(dataset_dir, preprocess=...) (environments.py:365-400): glob -> sorted paths ->
    read_pickle -> preprocess, on the very frames the goldens were recorded from (oracle/make_golden.py), written as
    pickles WITHOUT their feature columns so that `preprocess` has to produce them, as in the reference's examples
    (examples/example_multi_environnement.py)."""
    import gym_trading_env_b200 as gte
    g = H.load_golden(name)
    p = g["params"]
    lens = [300, 420, 360, 500]
    assert list(g["lengths"]) == lens
    for k, T in enumerate(lens):
        df = gte.make_gbm_ohlcv(T, seed=20 + k)
        raw = df.rename(columns={c: c.replace("feature_", "raw_") for c in df.columns})
        assert not any("feature" in c for c in raw.columns)
        raw.to_pickle(tmp_path / f"series_{k}.pkl")
    (tmp_path / "notes.txt").write_text("not a dataset")

    def preprocess(df):
        return df.rename(columns={c: c.replace("raw_", "feature_") for c in df.columns})

    kw = H.env_kwargs(g)
    env = gte.MultiDatasetTradingVectorEnv(str(tmp_path / "*.pkl"), preprocess=preprocess,
                                           episodes_between_dataset_switch=p["episodes_between_dataset_switch"],
                                           num_envs=p["n_envs"], reset_plan=g["plan"], verbose=0, debug_outputs=True, **kw)
    assert env.dataset_names == [f"series_{k}.pkl" for k in range(4)]
    for k, T in enumerate(lens):                                         # staged arrays == what the reference staged
        H.assert_bits(env._features[k, :T].cpu().numpy(), g["features"][k, :T], f"dataset {k} features")
        H.assert_bits(env._price[k, :T].cpu().numpy(), g["price"][k, :T], f"dataset {k} price")
    stats = H.replay_golden(H.DeviceAdapter(env), g, exact_money=True)
    assert stats["episodes"] > 0
    with pytest.raises(FileNotFoundError):
        gte.MultiDatasetTradingVectorEnv(str(tmp_path / "*.parquet"), num_envs=2, **kw)


@pytest.mark.parametrize("mode,dtype", [("copy", np.int8), ("copy", np.int64), ("mapped", np.int8), ("mapped", np.int64),
                                        ("numpy", np.int16)])
def test_host_paths_at_65536_envs_match_the_oracle(mode, dtype):
    """C3 size through the host-facing paths: hybrid via gte_step_host with the copy engines and with mapped (zero-copy)
    host memory, narrow and gymnasium-typed actions, and the full "numpy" mode — all against the oracle."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    from gym_trading_env_b200 import _cabi
    N, K = 65_536, 36
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(20_000, seed=2))
    pos = [-1, 0, 0.5, 1]
    kw = dict(positions=pos, windows=64, max_episode_duration=15, **FEES)
    if mode == "numpy":
        env = gte.TradingVectorEnv(series, num_envs=N, seed=3, verbose=0, output="numpy", **kw)
    else:
        env = gte.TradingVectorEnv(series, num_envs=N, seed=3, verbose=0, output="hybrid", host_io=mode, **kw)
    o = orc.OracleVecEnv(series.features, series.price, num_envs=N, seed=3, threads=8, **kw)
    obs, _ = env.reset()
    H.assert_bits(obs if isinstance(obs, np.ndarray) else obs.cpu().numpy(), o.reset(), "reset obs")
    rng = np.random.default_rng(7)
    pin = env.pinned_actions(dtype)
    assert pin.dtype == dtype
    for k in range(K):
        a = rng.integers(0, len(pos), size=N)
        a[rng.random(N) < 0.05] = -1
        if k % 3 == 2:
            got = env.step(a.astype(dtype))                              # pageable array: staged
        else:
            pin[...] = a
            got = env.step(pin)                                          # pinned: used in place
        o.step(a)
        obs_h = got[0] if isinstance(got[0], np.ndarray) else got[0].cpu().numpy()
        assert isinstance(got[1], np.ndarray) and got[1].dtype == np.float64
        assert got[2].dtype == np.bool_ and got[3].dtype == np.bool_
        H.assert_bits(obs_h, o.obs, f"step {k} obs")
        H.assert_close64(got[1], o.reward, f"step {k} reward")
        H.assert_bits(got[2].view(np.uint8), o.terminated, f"step {k} terminated")
        H.assert_bits(got[3].view(np.uint8), o.truncated, f"step {k} truncated")
        H.assert_bits(env._valuation.cpu().numpy(), o.valuation, f"step {k} valuation")
        H.assert_bits(env._asset.cpu().numpy(), o.asset, f"step {k} asset")
        assert np.array_equal(np.asarray(env.infos["reward"]), got[1])
    if mode != "numpy":
        assert env._io_mode_used.value == _cabi.IO_MODES[mode]
    assert float(env.get_metrics()["episodes"].item()) >= 2 * N
    # a bad action surfaces as IndexError with the results of the very step that carried it
    bad = np.zeros(N, dtype=dtype)
    bad[N // 2] = len(pos)
    with pytest.raises(IndexError):
        env.step(bad)
    env.step(np.zeros(N, dtype=dtype))                                   # the flag was cleared: the env keeps working


def test_strict_actions_reject_negative_indices_other_than_hold():
    """The reference's positions[-2] would index from the end of the list (environments.py:234); here every negative
    action is a hold unless strict_actions=True, which raises for anything below -1."""
    import gym_trading_env_b200 as gte
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(2000, seed=1))
    kw = dict(positions=[-1, 0, 1], windows=8, max_episode_duration=50, **FEES)
    lax = gte.TradingVectorEnv(series, num_envs=64, seed=1, verbose=0, output="hybrid", **kw)
    strict = gte.TradingVectorEnv(series, num_envs=64, seed=1, verbose=0, output="hybrid", strict_actions=True, **kw)
    lax.reset(); strict.reset()
    a = np.full(64, -1, np.int8)
    r0, r1 = lax.step(a), strict.step(a)                                # -1 = hold in both
    assert np.array_equal(r0[1], r1[1])
    a[5] = -2
    lax.step(a)                                                          # hold
    with pytest.raises(IndexError):
        strict.step(a)
    dev = gte.TradingVectorEnv(series, num_envs=64, seed=1, verbose=0, strict_actions=True, **kw)
    dev.reset()
    dev.step(torch.full((64,), -3, dtype=torch.int64, device=dev.device))
    with pytest.raises(IndexError):
        dev.check_errors()


def test_reset_plan_rows_are_range_checked_and_exhaustion_is_flagged():
    """A plan row outside the dataset / position list is clamped and flagged instead of reading out of bounds, and a
    plan that runs out of episodes is flagged instead of silently replaying (C-ABI hardening)."""
    import gym_trading_env_b200 as gte
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(600, seed=1))
    kw = dict(positions=[0, 1], windows=4, max_episode_duration=10, **FEES)
    plan = np.zeros((8, 2, 3), np.int32)
    plan[:, :, 0] = 50
    env = gte.TradingVectorEnv(series, num_envs=8, reset_plan=plan, verbose=0, **kw)
    env.reset()
    a = torch.zeros(8, dtype=torch.int64, device=env.device)
    for _ in range(9):
        env.step(a)                                                      # second plan episode consumed by the auto-reset
    env.check_errors()
    for _ in range(9):
        env.step(a)                                                      # third episode: plan exhausted
    with pytest.raises(RuntimeError):
        env.check_errors()
    bad = plan.copy()
    bad[3, 0] = (10_000_000, 7, 0)                                       # start row and position index out of range
    env2 = gte.TradingVectorEnv(series, num_envs=8, reset_plan=bad, verbose=0, **kw)
    env2.reset()
    with pytest.raises(ValueError):
        env2.check_errors()
    assert 0 <= int(env2._ep_start[3]) <= 598 and 0 <= int(env2._pos_idx[3]) <= 1


@pytest.mark.parametrize("n_envs,windows,pos", [(2500, 64, [-3, -2, -1, 0, 1, 2, 3]), (65_536, 64, [-1, 0, 0.5, 1]),
                                                (100_000, 64, [-1, 0, 1]), (9_000, 16, [0, 1])])
def test_fused_single_launch_equals_two_plain_launches(n_envs, windows, pos):
    """gte_step_obs runs small batches as ONE fused launch (transition + gather per CTA); n_chunks=1 forces the two plain
    launches.  Both must produce the same bits: observations, rewards, flags, portfolio state, Philox resets."""
    import gym_trading_env_b200 as gte
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(20_000, seed=4))
    kw = dict(positions=pos, windows=windows, max_episode_duration=20, num_envs=n_envs, seed=21, verbose=0,
              debug_outputs=True, **FEES)
    fused, plain = gte.TradingVectorEnv(series, n_chunks=0, **kw), gte.TradingVectorEnv(series, n_chunks=1, **kw)
    assert fused.launches_per_step == 1 and plain.launches_per_step == 2
    o0, _ = fused.reset()
    o1, _ = plain.reset()
    assert torch.equal(o0.view(torch.int32), o1.view(torch.int32))
    acts = _acts(n_envs, 50, len(pos), fused.device, seed=5)
    acts[:, ::17] = -1                                                   # holds
    for k in range(50):
        fused.step(acts[k])
        plain.step(acts[k])
        assert torch.equal(fused._obs.view(torch.int32), plain._obs.view(torch.int32)), k
        for nm in ("reward", "valuation", "real_position", "terminated", "truncated", "asset", "fiat", "interest_asset",
                   "interest_fiat", "pos_idx", "step", "ep_start", "info_idx", "info_step", "dyn_ring"):
            assert torch.equal(getattr(fused, "_" + nm), getattr(plain, "_" + nm)), (k, nm)
        assert torch.equal(fused._metrics_step[:3], plain._metrics_step[:3])
        torch.testing.assert_close(fused._metrics_step, plain._metrics_step, rtol=1e-9, atol=1e-9)
    assert float(fused._metrics_total[0]) >= 2 * n_envs
    assert int(fused._ring_clock) == int(plain._ring_clock) == 50 and int(fused._tick_dev) == int(plain._tick_dev)
    fused.check_errors()


def test_dataset_dir_with_mixed_files_and_shuffled_columns_equals_in_memory_datasets(tmp_path):
    """Ragged ingestion end to end: a directory holding a pickle and a CSV of different lengths whose columns come in
    different orders gives the same env as the reconciled in-memory frames."""
    import gym_trading_env_b200 as gte
    a, b = gte.make_gbm_ohlcv(400, seed=31), gte.make_gbm_ohlcv(650, seed=32)
    a.to_pickle(tmp_path / "0_a.pkl")
    shuffled = b[list(reversed(b.columns))]
    shuffled.to_pickle(tmp_path / "1_b.pkl")
    kw = dict(positions=[-1, 0, 1], windows=8, max_episode_duration=25, num_envs=512, seed=9, verbose=0, **FEES)
    disk = gte.MultiDatasetTradingVectorEnv(str(tmp_path / "*.pkl"), **kw)
    mem = gte.MultiDatasetTradingVectorEnv(datasets=[a, b], **kw)
    assert disk.dataset_names == ["0_a.pkl", "1_b.pkl"] and disk._lengths_np.tolist() == [400, 650]
    o0, _ = disk.reset()
    o1, _ = mem.reset()
    assert torch.equal(o0.view(torch.int32), o1.view(torch.int32))
    acts = _acts(512, 60, 3, disk.device, seed=3)
    for k in range(60):
        r0, r1 = disk.step(acts[k]), mem.step(acts[k])
        assert torch.equal(r0[0].view(torch.int32), r1[0].view(torch.int32)) and torch.equal(r0[1], r1[1])
        assert torch.equal(disk._dataset_idx, mem._dataset_idx)
    assert set(disk._dataset_idx.cpu().tolist()) == {0, 1}


def test_resident_server_kernel_matches_the_oracle_and_survives_interruptions():
    """host_io="server" (C2 shape: 4096 envs, windows=None): a resident kernel answers each step through mapped host
    memory.  Same bits as the oracle, across the events that make the kernel leave and come back: an info read and a
    masked reset (other entry points stop it), an idle pause longer than its time-out, and close()."""
    import time
    import gym_trading_env_b200 as gte
    import oracle as orc
    from gym_trading_env_b200 import _cabi
    N = 4096
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(5000, seed=6))
    pos = [-1, 0, 0.5, 1]
    kw = dict(positions=pos, windows=None, max_episode_duration=40, **FEES)
    env = gte.TradingVectorEnv(series, num_envs=N, seed=3, verbose=0, output="hybrid", host_io="server", **kw)
    o = orc.OracleVecEnv(series.features, series.price, num_envs=N, seed=3, **kw)
    obs, _ = env.reset()
    H.assert_bits(obs.cpu().numpy(), o.reset(), "reset obs")
    rng = np.random.default_rng(1)
    pin = env.pinned_actions()
    mask = np.zeros(N, dtype=bool)
    mask[::7] = True
    for k in range(160):
        a = rng.integers(0, len(pos), size=N)
        a[rng.random(N) < 0.05] = -1
        pin[...] = a
        obs, rew, term, trunc, infos = env.step(pin)
        assert env._io_mode_used.value == _cabi.IO_SERVER
        o.step(a)
        H.assert_bits(obs.cpu().numpy(), o.obs, f"step {k} obs")
        H.assert_close64(rew, o.reward, f"step {k} reward")
        H.assert_bits(term.view(np.uint8), o.terminated, f"step {k} terminated")
        H.assert_bits(trunc.view(np.uint8), o.truncated, f"step {k} truncated")
        H.assert_bits(env._valuation.cpu().numpy(), o.valuation, f"step {k} valuation")
        H.assert_bits(env._asset.cpu().numpy(), o.asset, f"step {k} asset")
        H.assert_bits(env._step.cpu().numpy(), o.step_, f"step {k} step state")
        if k == 40:                                                      # gte_info stops the server; the next step relaunches it
            H.assert_bits(infos["idx"].cpu().numpy(), o.idx, "infos idx")
        if k == 80:                                                      # so does a (masked) reset
            env.reset(options={"mask": mask})
            o.reset(mask=mask)
        if k == 120:
            time.sleep(0.02)                                             # longer than the idle time-out: it left by itself
    m = env._metrics_total.cpu().numpy()
    assert m[0] >= 3 * N
    env.close()
    assert env.closed
    # two server-mode envs on one device take turns: each call finds the other env's kernel resident, stops it and
    # launches its own — slow, but every step still matches the oracle
    e1 = gte.TradingVectorEnv(series, num_envs=512, seed=5, verbose=0, output="hybrid", host_io="server", **kw)
    e2 = gte.TradingVectorEnv(series, num_envs=300, seed=6, verbose=0, output="hybrid", host_io="server", **kw)
    o1 = orc.OracleVecEnv(series.features, series.price, num_envs=512, seed=5, **kw)
    o2 = orc.OracleVecEnv(series.features, series.price, num_envs=300, seed=6, **kw)
    e1.reset(); e2.reset(); o1.reset(); o2.reset()
    for k in range(30):
        a1, a2 = rng.integers(0, len(pos), size=512), rng.integers(0, len(pos), size=300)
        r1 = e1.step(a1.astype(np.int8))
        r2 = e2.step(a2.astype(np.int16))
        o1.step(a1); o2.step(a2)
        H.assert_close64(r1[1], o1.reward, f"env1 step {k} reward")
        H.assert_close64(r2[1], o2.reward, f"env2 step {k} reward")
        H.assert_bits(r1[0].cpu().numpy(), o1.obs, f"env1 step {k} obs")
        H.assert_bits(r2[0].cpu().numpy(), o2.obs, f"env2 step {k} obs")
    e1.close(); e2.close()
    bad = gte.TradingVectorEnv(series, num_envs=64, seed=3, verbose=0, output="hybrid", host_io="server", **kw)
    bad.reset()
    with pytest.raises(IndexError):
        bad.step(np.full(64, 9, np.int8))
    bad.step(np.zeros(64, np.int8))
    bad.close()


@pytest.mark.parametrize("host_io", ["auto", "copy", "server"])
def test_numpy_mode_with_one_row_observations_matches_the_oracle(host_io):
    """output="numpy", windows=None (the gymnasium / SB3 shape of the reference's default env): actions from the host,
    observations + reward + flags back on the host in one C call.  "auto" (mapped at this size) and "server" have the
    step kernel write the observation rows straight into pinned host memory; "copy" goes through the copy engines."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    from gym_trading_env_b200 import _cabi
    N = 2048
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(4000, seed=8))
    pos = [-1, 0, 1]
    kw = dict(positions=pos, windows=None, max_episode_duration=30, **FEES)
    env = gte.TradingVectorEnv(series, num_envs=N, seed=4, verbose=0, output="numpy", host_io=host_io, **kw)
    o = orc.OracleVecEnv(series.features, series.price, num_envs=N, seed=4, **kw)
    obs, _ = env.reset()
    assert isinstance(obs, np.ndarray)
    H.assert_bits(obs, o.reset(), "reset obs")
    rng = np.random.default_rng(2)
    for k in range(90):
        a = rng.integers(0, len(pos), size=N)
        obs, rew, term, trunc, infos = env.step(a.astype(np.int8) if k % 2 else a)
        o.step(a)
        assert isinstance(obs, np.ndarray) and obs.shape == (N, 10) and obs.dtype == np.float32
        H.assert_bits(obs, o.obs, f"step {k} obs")
        H.assert_close64(rew, o.reward, f"step {k} reward")
        H.assert_bits(term.view(np.uint8), o.terminated, f"step {k} terminated")
        H.assert_bits(trunc.view(np.uint8), o.truncated, f"step {k} truncated")
    want = {"auto": _cabi.IO_MAPPED, "copy": _cabi.IO_COPY, "server": _cabi.IO_SERVER}[host_io]
    assert env._io_mode_used.value == want
    env.close()


def test_step_async_wait_pipelined_equals_synchronous_steps():
    """step_async / step_wait with TWO iterations in flight (the copy of iteration k's results runs under iteration
    k+1): same results, in order, as synchronous hybrid steps and as the oracle; pinned and pageable action arrays."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    N, K = 40_000, 50
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(20_000, seed=2))
    pos = [-1, 0, 0.5, 1]
    kw = dict(positions=pos, windows=64, max_episode_duration=15, **FEES)
    env = gte.TradingVectorEnv(series, num_envs=N, seed=3, verbose=0, output="hybrid", **kw)
    o = orc.OracleVecEnv(series.features, series.price, num_envs=N, seed=3, threads=8, **kw)
    obs, _ = env.reset()
    H.assert_bits(obs.cpu().numpy(), o.reset(), "reset obs")
    rng = np.random.default_rng(5)
    acts = rng.integers(0, len(pos), size=(K, N))
    pins = [env.pinned_actions(), env.pinned_actions()]
    with pytest.raises(RuntimeError):
        env.step_wait()
    pins[0][...] = acts[0]
    env.step_async(pins[0])
    for k in range(K):
        if k + 1 < K:                                                    # iteration k+1 goes out before k is read
            if k % 5 == 4:
                env.step_async(acts[k + 1].astype(np.int16))             # pageable: staged per wire set
            else:
                pins[(k + 1) & 1][...] = acts[k + 1]
                env.step_async(pins[(k + 1) & 1])
        _, rew, term, trunc, _ = env.step_wait()
        o.step(acts[k])
        H.assert_close64(rew, o.reward, f"step {k} reward")
        H.assert_bits(term.view(np.uint8), o.terminated, f"step {k} terminated")
        H.assert_bits(trunc.view(np.uint8), o.truncated, f"step {k} truncated")
    H.assert_bits(env._obs.cpu().numpy(), o.obs, "last obs")
    H.assert_bits(env._asset.cpu().numpy(), o.asset, "last asset")
    env.step_async(acts[0]); env.step_async(acts[1])
    with pytest.raises(RuntimeError):
        env.step_async(acts[2])                                          # two already in flight
    env.step_wait(); env.step_wait()
    r = env.step(acts[2].astype(np.int8))                                # the synchronous call still works afterwards
    assert r[1].shape == (N,)
    env.close()


def test_sparse_flag_wire_reconstructs_the_dense_flags_including_bursts():
    """The sparse flag wire (copy engines): terminated / truncated rebuilt on the host from the list of ended envs are
    identical to the dense bytes — in ordinary iterations, and in the iteration where EVERY env's episode ends at once
    (more entries than the list holds: the dense bytes are fetched inside the same call), synchronously and through
    step_async / step_wait."""
    import gym_trading_env_b200 as gte
    N = 50_000
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(20_000, seed=2))
    pos = [-1, 0, 0.5, 1]
    kw = dict(positions=pos, windows=8, max_episode_duration=12, num_envs=N, seed=3, verbose=0, output="hybrid",
              host_io="copy", **FEES)
    sparse = gte.TradingVectorEnv(series, sparse_flags=True, **kw)
    dense = gte.TradingVectorEnv(series, sparse_flags=False, **kw)
    apipe = gte.TradingVectorEnv(series, sparse_flags=True, **kw)
    assert sparse.sparse_flags and not dense.sparse_flags
    sparse.reset(); dense.reset(); apipe.reset()
    rng = np.random.default_rng(5)
    acts = rng.integers(0, len(pos), size=(40, N)).astype(np.int8)
    bursts = 0
    apipe.step_async(acts[0])
    for k in range(40):
        if k + 1 < 40:
            apipe.step_async(acts[k + 1])
        ra = apipe.step_wait()
        rs, rd = sparse.step(acts[k]), dense.step(acts[k])
        for r in (rs, ra):
            assert np.array_equal(r[1], rd[1]), k
            assert np.array_equal(r[2], rd[2]) and np.array_equal(r[3], rd[3]), k
            assert r[2].dtype == np.bool_ and r[3].dtype == np.bool_
        n_end = int((rd[2] | rd[3]).sum())
        assert int(sparse._host["sparse"].n_ended[0]) == n_end
        bursts += n_end > sparse._host["sparse"].entries.size
    assert bursts >= 2            # all envs start together: every 11th iteration all 50 000 episodes end at once
    assert int((rd[2] | rd[3]).sum()) < N


def test_float32_reward_wire_is_the_numpy_cast_of_the_fp64_reward():
    """reward_wire="f32" (opt-in, lossy): the rewards that cross PCIe are exactly numpy's float32 cast of the fp64 rewards
    of the lossless wire, flags and error word are unchanged (sparse and dense flag wires, bursts included), the fp64
    rewards stay on the device, and a small batch answered through mapped memory hands out the same dtype."""
    import gym_trading_env_b200 as gte
    N = 50_000
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(20_000, seed=2))
    pos = [-1, 0, 0.5, 1]
    kw = dict(positions=pos, windows=8, max_episode_duration=12, num_envs=N, seed=3, verbose=0, output="hybrid",
              host_io="copy", **FEES)
    ref = gte.TradingVectorEnv(series, sparse_flags=False, **kw)
    envs = [gte.TradingVectorEnv(series, sparse_flags=sp, reward_wire="f32", **kw) for sp in (True, False)]
    apipe = gte.TradingVectorEnv(series, sparse_flags=True, reward_wire="f32", **kw)
    for e in envs + [ref, apipe]:
        e.reset()
    rng = np.random.default_rng(5)
    acts = rng.integers(0, len(pos), size=(26, N)).astype(np.int8)
    apipe.step_async(acts[0])
    for k in range(26):
        if k + 1 < 26:
            apipe.step_async(acts[k + 1])
        rd = ref.step(acts[k])
        want = rd[1].astype(np.float32)
        for r in [e.step(acts[k]) for e in envs] + [apipe.step_wait()]:
            assert r[1].dtype == np.float32
            H.assert_bits(r[1], want, f"step {k} f32 reward")
            assert np.array_equal(r[2], rd[2]) and np.array_equal(r[3], rd[3]), k
        for e in envs:
            H.assert_bits(e.reward_device.cpu().numpy(), rd[1], f"step {k} device fp64 reward")
    assert np.count_nonzero(want) > N // 2
    with pytest.raises(ValueError):
        gte.TradingVectorEnv(series, reward_wire="f32", **dict(kw, output="torch"))
    small = dict(kw, num_envs=512, host_io="auto")
    a, b = gte.TradingVectorEnv(series, reward_wire="f32", **small), gte.TradingVectorEnv(series, **small)
    a.reset(); b.reset()
    for k in range(5):
        ra, rb = a.step(acts[k, :512]), b.step(acts[k, :512])
        assert ra[1].dtype == np.float32
        H.assert_bits(ra[1], rb[1].astype(np.float32), "mapped batch, f32 out")


@pytest.mark.parametrize("windows", [None, 8])
def test_persistent_rollout_with_dataset_rotation_and_narrow_actions_equals_stepping(windows):
    """gte_rollout as ONE persistent launch (state in registers for all K iterations) on a multi-dataset env: dataset
    switches, Philox resets and ring slots derived inside the kernel give the same bits as K separate step() calls."""
    import gym_trading_env_b200 as gte
    frames = [gte.make_gbm_ohlcv(T, seed=50 + k) for k, T in enumerate([700, 900, 800])]
    kw = dict(positions=[-1, 0, 1, 2], windows=windows, max_episode_duration=17, num_envs=3000, seed=13, verbose=0,
              episodes_between_dataset_switch=2, **FEES)
    a, b = gte.MultiDatasetTradingVectorEnv(datasets=frames, **kw), gte.MultiDatasetTradingVectorEnv(datasets=frames, **kw)
    a.reset(); b.reset()
    acts = _acts(3000, 70, 4, a.device, seed=8)
    acts[::3, ::5] = -1
    out = a.rollout(acts[:40], keep_obs=windows is None)
    out2 = a.rollout(acts[40:])
    for k in range(70):
        obs, rew, term, trunc, _ = b.step(acts[k])
        o = out if k < 40 else out2
        kk = k if k < 40 else k - 40
        assert torch.equal(o["reward"][kk], rew) and torch.equal(o["terminated"][kk], term) and torch.equal(o["truncated"][kk], trunc), k
        assert torch.equal(o["valuation"][kk], b._valuation), k
        if windows is None and k < 40:
            assert torch.equal(o["obs"][kk].view(torch.int32), obs.view(torch.int32)), k
    for nm in ("obs", "asset", "fiat", "interest_asset", "interest_fiat", "pos_idx", "step", "ep_start", "dataset_idx", "dyn_ring",
               "ds_used", "ds_episodes"):
        x, y = getattr(a, "_" + nm), getattr(b, "_" + nm)
        assert torch.equal(x.view(torch.int32) if x.dtype == torch.float32 else x, y.view(torch.int32) if y.dtype == torch.float32 else y), nm
    assert int(a._ring_clock) == int(b._ring_clock) == 70 and int(a._tick_dev) == int(b._tick_dev)
    assert torch.equal(a._metrics_total[:3], b._metrics_total[:3])
    torch.testing.assert_close(a._metrics_total, b._metrics_total, rtol=1e-12, atol=1e-12)
    assert set(a._dataset_idx.cpu().tolist()) == {0, 1, 2}


def test_result_block_slots_are_recycled_once_their_iteration_has_ended():
    """gte_step_host_begin keeps at most 4 result blocks IN FLIGHT per device; blocks whose iteration has ended give their
    slot back, so any number of envs (or wire sets rebuilt by enable_result_relay) can take turns."""
    import gym_trading_env_b200 as gte
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(5_000, seed=2))
    kw = dict(positions=[-1, 0, 1], windows=4, max_episode_duration=30, num_envs=2048, seed=3, verbose=0, output="hybrid",
              host_io="copy", **FEES)
    envs = [gte.TradingVectorEnv(series, **kw) for _ in range(4)]          # 4 envs x 2 wire sets = 8 result blocks
    ref = gte.TradingVectorEnv(series, **kw)
    acts = np.random.default_rng(1).integers(0, 3, size=(6, 2048)).astype(np.int8)
    for e in envs + [ref]:
        e.reset()
    for k in range(6):
        want = ref.step(acts[k])
        for e in envs:
            e.step_async(acts[k])
            got = e.step_wait()
            H.assert_bits(got[1], want[1], f"step {k} reward")
    envs[0].step_async(acts[0]); envs[0].step_async(acts[1]); envs[1].step_async(acts[0]); envs[1].step_async(acts[1])
    with pytest.raises(RuntimeError):                                      # a fifth block in flight is refused, loudly
        envs[2].step_async(acts[0])
    for e in (envs[0], envs[0], envs[1], envs[1]):
        e.step_wait()
    for e in envs + [ref]:
        e.close()


def test_reward_host_count_stops_the_results_copy_where_the_relay_takes_over():
    """GteHostIO.reward_host_count (the result relay's hook in the C-ABI): the call's own device-to-host copy delivers the
    first `count` rewards and the header / list / flags; rewards [count, N) stay on the device (for the caller to route
    through a peer GPU) — the host array keeps whatever it held there."""
    import gym_trading_env_b200 as gte
    N, count = 50_000, 20_480
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(20_000, seed=2))
    kw = dict(positions=[-1, 0, 0.5, 1], windows=8, max_episode_duration=12, num_envs=N, seed=3, verbose=0, output="hybrid",
              host_io="copy", **FEES)
    for sparse in (True, False):
        a, b = gte.TradingVectorEnv(series, sparse_flags=sparse, **kw), gte.TradingVectorEnv(series, sparse_flags=sparse, **kw)
        a.reset(); b.reset()
        acts = np.random.default_rng(5).integers(0, 4, size=(14, N)).astype(np.int8)
        a.step(acts[0]); b.step(acts[0])
        a._io.reward_host_count = count
        for k in range(1, 14):                                          # (iteration 11 ends every episode at once: dense flags)
            a._host["reward"][count:] = -7.0
            ra, rb = a.step(acts[k]), b.step(acts[k])
            H.assert_bits(ra[1][:count], rb[1][:count], f"step {k}: the rewards this call's own copy delivers")
            assert np.all(ra[1][count:] == -7.0), "rewards past reward_host_count must not be copied by this call"
            H.assert_bits(a.reward_device.cpu().numpy(), rb[1], f"step {k}: all rewards are on the device")
            assert np.array_equal(ra[2], rb[2]) and np.array_equal(ra[3], rb[3]), k
        a.close(); b.close()


def test_claimed_tiles_gather_passes_the_parity_suites_at_every_size():
    """The gather's claimed-tiles form (tiles handed out from a device counter instead of an even split; taken from
    ~150k envs on, so the full-size C4 / C5 tests above already run it) forced on for EVERY windowed batch
    (GTE_TMA_DYN=2): the golden / oracle parity files and the edge cases — ragged last tiles, batches smaller than the grid,
    CTAs whose second tile is already past the end, env ranges — must not notice."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GTE_TMA_DYN="2")
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
           os.path.join(root, "tests", "test_cuda_parity.py"), os.path.join(root, "tests", "test_cuda_edge_cases.py")]
    r = subprocess.run(cmd, cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout and "failed" not in r.stdout
