"""CPU: host-side logic — data staging rules, window-table layout, env sharding, import shim."""
import numpy as np
import pandas as pd
import pytest

import gym_trading_env_b200 as gte
from gym_trading_env_b200.data import build_window_tables, window_table_classes


def test_frame_to_arrays_follows_set_df_rules():
    df = gte.make_gbm_ohlcv(400, seed=1)
    df["my_feature_x"] = np.arange(400, dtype=np.float64) / 3.0       # name CONTAINS "feature" (environments.py:130)
    df["label"] = "x"
    a = gte.frame_to_arrays(df)
    assert a.feature_names == [c for c in df.columns if "feature" in c] and a.feature_names[-1] == "my_feature_x"
    assert a.features.dtype == np.float32 and a.price.dtype == np.float64
    # numpy's own fp64->fp32 cast, as `_set_df` does (:141)
    assert np.array_equal(a.features, np.array(df[a.feature_names], dtype=np.float32))
    assert np.array_equal(a.price, df["close"].to_numpy())
    assert set(a.info) == {"open", "high", "low", "close", "volume"}
    with pytest.raises(ValueError):
        gte.frame_to_arrays(df.drop(columns=["close"]))


def test_gbm_generator_is_deterministic_and_has_eight_features():
    a, b = gte.make_gbm_ohlcv(1000, seed=3), gte.make_gbm_ohlcv(1000, seed=3)
    assert a.equals(b) and len(a) == 1000 and not a.isna().any().any()
    assert len([c for c in a.columns if "feature" in c]) == 8
    assert isinstance(a.index, pd.DatetimeIndex)
    f, p = gte.make_gbm_arrays(5000, seed=2)
    assert f.shape == (5000, 8) and f.dtype == np.float32 and p.shape == (5000,) and np.isfinite(f).all()


@pytest.mark.parametrize("ns,nd,W", [(8, 2, 64), (5, 2, 24), (3, 0, 16), (7, 2, 16), (8, 0, 4)])
def test_window_tables_put_every_window_on_a_16_byte_boundary(ns, nd, W):
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((3, 50, ns)).astype(np.float32)
    tables, shifts, ds_stride = build_window_tables(feats, nd)
    F, rb = ns + nd, 4 * (ns + nd)
    assert ds_stride % 16 == 0 and set(tables) == set(window_table_classes(rb))
    for ds in range(3):
        for r0 in range(0, 50 - W):
            off = r0 * rb
            c = (off >> 2) & 3
            start = shifts[c] + ds * ds_stride + off
            assert start % 16 == 0                                        # TMA / 128-bit load alignment
            win = tables[c][start:start + W * rb].view(np.float32).reshape(W, F)
            assert np.array_equal(win[:, :ns], feats[ds, r0:r0 + W]) and not win[:, ns:].any()


def test_shard_envs_partitions_the_index_range():
    from gym_trading_env_b200.vector_env import shard_envs
    for total, world in [(1 << 24, 8), (65536, 4), (10, 3), (7, 8)]:
        spans = [shard_envs(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_import_shim_points_at_the_hyphenated_package_dir():
    import os
    assert os.path.basename(os.path.dirname(gte.__file__)) == "gym-trading-env_b200"
    assert gte.__version__


def test_constructor_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gte.TradingVectorEnv(gte.make_gbm_ohlcv(300, seed=0), num_envs=4)


def test_sb3_view_exposes_the_vecenv_surface_without_touching_the_gpu():
    """The stable-baselines3 adapter is plain host code over an env object: check its contract on a stand-in env
    (the GPU test replays a reference golden through it)."""
    import torch
    import gym_trading_env_b200 as gte

    class FakeEnv:
        num_envs, device, render_mode, output, _host, final_obs = 3, torch.device("cpu"), "logs", "torch", None, None
        single_observation_space, single_action_space = "obs-space", "act-space"

        def reset(self):
            return np.zeros((3, 2), np.float32), {}

        def step(self, a):
            assert a.dtype == np.int64 and a.shape == (3,)
            self.final_obs = torch.full((3, 2), 7.0)
            return (np.ones((3, 2), np.float32), np.array([0.5, 0.25, 0.0]), np.array([False, True, False]),
                    np.array([False, False, True]), {"idx": np.array([4, 5, 6])})

        def close(self):
            self.closed = True

    env = FakeEnv()
    v = gte.SB3VecEnv(env, info_keys=("idx",))
    assert env.output == "numpy" and v.num_envs == 3 and v.observation_space == "obs-space"
    assert v.reset().shape == (3, 2)
    obs, rew, dones, infos = v.step([0, 1, 0])
    assert rew.dtype == np.float32 and dones.tolist() == [False, True, True]
    assert [i["idx"] for i in infos] == [4, 5, 6]
    assert "terminal_observation" not in infos[0] and infos[1]["TimeLimit.truncated"] is False
    assert infos[2]["TimeLimit.truncated"] is True and infos[2]["terminal_observation"].tolist() == [7.0, 7.0]
    assert v.env_is_wrapped(object) == [False] * 3 and v.get_attr("num_envs", 1) == [3]
    v.close()
    assert env.closed


def test_dataset_files_are_loaded_and_reconciled_by_column_name(tmp_path):
    """f4 ingestion (reference: MultiDatasetTradingEnv glob + read_pickle + preprocess, environments.py:365-391;
    examples/example_environnement.py:11-14 for the CSV form): pickles and CSVs, ragged lengths, shuffled column
    order, a non-numeric column — staged into arrays that share one feature schema; a dataset with another feature
    set is refused by name."""
    import pandas as pd
    a = gte.make_gbm_ohlcv(300, seed=1)
    b = gte.make_gbm_ohlcv(450, seed=2)
    b = b[list(reversed(b.columns))].copy()                      # same columns, other order
    b["exchange"] = "binance"                                    # non-numeric: ignored
    a.to_pickle(tmp_path / "a.pkl")
    b.reset_index().rename(columns={"index": "date"}).to_csv(tmp_path / "b.csv", index=False)
    fa, fb = gte.load_frame(str(tmp_path / "a.pkl")), gte.load_frame(str(tmp_path / "b.csv"))
    assert isinstance(fb.index, pd.DatetimeIndex) and fb.index.is_monotonic_increasing and len(fb) == 450
    sa, sb = gte.frame_to_arrays(fa), gte.frame_to_arrays(fb)
    assert sb.feature_names == list(reversed(sa.feature_names)) and "exchange" not in sb.info
    ra, rb = gte.reconcile_series([sa, sb], ["a.pkl", "b.csv"])
    assert ra.feature_names == rb.feature_names == sa.feature_names and (ra.length, rb.length) == (300, 450)
    want = gte.frame_to_arrays(gte.make_gbm_ohlcv(450, seed=2))
    np.testing.assert_allclose(rb.features, want.features, rtol=1e-6)      # CSV text round trip: to float32 rounding
    np.testing.assert_allclose(rb.price, want.price, rtol=1e-12)
    assert set(ra.info) == set(rb.info) == {"open", "high", "low", "close", "volume"}
    c = gte.frame_to_arrays(gte.make_gbm_ohlcv(200, seed=3).drop(columns=["feature_vol_64"]))
    with pytest.raises(ValueError, match="c.pkl.*feature_vol_64"):
        gte.reconcile_series([sa, c], ["a.pkl", "c.pkl"])
    with pytest.raises(ValueError, match="unsupported dataset file type"):
        gte.load_frame(str(tmp_path / "x.xlsx"))


def test_sparse_flag_mirror_patches_and_falls_back_to_dense_bytes():
    """Host side of the sparse flag wire (pure numpy): entries (terminated << 31 | truncated << 30 | env) are patched into
    persistent arrays, the previous iteration's entries are cleared, and a burst larger than the list takes the dense bytes."""
    from gym_trading_env_b200 import _cabi
    from gym_trading_env_b200.vector_env import _SparseFlags
    n = 5000
    toff, uoff, eoff, nbytes = _cabi.host_result_layout(n)
    cap = _cabi.host_result_ended_cap(n)
    assert cap == 1024 and _cabi.host_result_sparse_bytes(n) == 8 * n + 32 + 4 * cap == toff and nbytes % 8 == 0
    assert _cabi.host_result_ended_cap(1 << 21) == 65536 and eoff == 8 * n and uoff == toff + n
    block = np.zeros(nbytes, np.uint8)
    m = _SparseFlags(n, block)
    hdr = block[eoff:eoff + 32].view(np.uint32)
    entries = block[eoff + 32:eoff + 32 + 4 * cap].view(np.uint32)

    def publish(term_ids, trunc_ids, both_ids=()):
        e = [i | 0x80000000 for i in term_ids] + [i | 0x40000000 for i in trunc_ids] + [i | 0xC0000000 for i in both_ids]
        entries[:len(e)] = np.array(e, np.uint32)
        hdr[2] = len(e)

    publish([3, 4999], [10, 11], [77])
    t, u = m.update()
    assert np.flatnonzero(t).tolist() == [3, 77, 4999] and np.flatnonzero(u).tolist() == [10, 11, 77]
    publish([], [5])
    t2, u2 = m.update()
    assert t2 is t and u2 is u and not t.any() and np.flatnonzero(u).tolist() == [5]      # same arrays, old entries cleared
    # burst: more ends than the list holds -> the dense bytes (fetched by gte_step_host in the same call) are taken
    dense_t, dense_u = block[toff:toff + n].view(np.bool_), block[uoff:uoff + n].view(np.bool_)
    dense_u[:] = True
    dense_t[::2] = True
    hdr[2] = n
    t, u = m.update()
    assert u.all() and t[::2].all() and not t[1::2].any()
    publish([1], [])
    t, u = m.update()                                                                     # after a burst everything is cleared first
    assert np.flatnonzero(t).tolist() == [1] and not u.any()


def test_hostbind_helpers_parse_cpu_lists_and_device_order(monkeypatch):
    from gym_trading_env_b200 import hostbind
    assert hostbind._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and hostbind._parse_cpulist("") == set()
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "4,6")
    assert hostbind._physical_index(1) == 6 and hostbind._physical_index(5) == 5
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES")
    assert hostbind._physical_index(3) == 3
    info = hostbind.bind_host_to_gpu(0, enable=False)
    assert info == {"gpu": 0, "bound": False, "how": "disabled"}
    info = hostbind.bind_host_to_gpu(0)                         # no GPU / no NUMA information here: a no-op that says so
    assert info["gpu"] == 0 and "how" in info


# ---- result relay (relay.py): the planner and the shared block, without a device -----------------------------------
def test_relay_plan_pairs_slow_links_with_fast_ones_and_leaves_even_links_alone():
    from gym_trading_env_b200.relay import QUANTUM, parse_forced_plan, plan_relay
    N = 1 << 21
    bw = [12.9, 12.9, 12.9, 12.9, 22.5, 22.7, 23.0, 22.1]          # profiles/r02_host_io_probe_n8.json
    plan = plan_relay(bw, N)
    assert sorted(plan) == [0, 1, 2, 3] and sorted(f for f, _ in plan.values()) == [4, 5, 6, 7]
    for s, (f, x) in plan.items():
        assert x % QUANTUM == 0 and 0 < x < N // 2
        t_s, t_f = (N - x) / bw[s], (N + x) / bw[f]                  # both links finish together
        assert abs(t_s - t_f) / t_f < 0.01
    assert plan_relay([50.0, 51.0, 49.5, 50.2], N) == {}              # within 20 %: nothing to balance
    assert plan_relay([10.0], N) == {} and plan_relay([], N) == {}
    assert plan_relay([10.0, 100.0], N)[0] == (1, N // 2 // QUANTUM * QUANTUM)      # never more than half
    assert plan_relay([0.0, 10.0], N) == {}
    assert plan_relay(bw, N) == plan_relay(list(bw), N)               # pure: the same plan on every rank
    assert parse_forced_plan("0>1:0.3, 2>3:0.25", 1 << 18) == {0: (1, 77824), 2: (3, 65536)}
    assert parse_forced_plan("", 1 << 18) == {} and parse_forced_plan("0>1:0.0001", 1 << 18) == {}


def test_relay_shared_block_is_one_memory_seen_through_two_mappings():
    from gym_trading_env_b200.relay import SharedPinnedBlock

    class NoDevice:                                                    # page-locking needs a GPU; the mapping does not
        def __init__(self): self.calls = []
        def gte_host_register(self, p, n): self.calls.append(("reg", n)); return 0
        def gte_host_unregister(self, p): self.calls.append(("unreg",)); return 0

    lib = NoDevice()
    own = SharedPinnedBlock(lib, 1 << 16)
    peer = SharedPinnedBlock(lib, 1 << 16, path=own.path)
    assert own.array.flags.writeable and not own.array.any()
    own.array[5] = 7
    peer.array[1000:1004].view(np.uint32)[0] = 0xdeadbeef
    assert peer.array[5] == 7 and own.array[1000:1004].view(np.uint32)[0] == 0xdeadbeef
    own.drop_fd()                                                      # the mappings outlive the descriptor
    peer.array[9] = 4
    assert own.array[9] == 4
    own.close(); peer.close(); own.close()
    assert lib.calls == [("reg", 1 << 16), ("reg", 1 << 16), ("unreg",), ("unreg",)]

