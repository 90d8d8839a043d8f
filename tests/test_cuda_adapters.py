"""GPU tests of the caller-facing adapters around the hot path: the
stable-baselines3 VecEnv view and `save_for_render`, both checked against goldens recorded from the reference."""
from __future__ import annotations

import numpy as np
import pandas as pd
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def test_sb3_vecenv_view_matches_the_reference_golden():
    """SB3's VecEnv contract on top of the CUDA env: numpy obs / float32 rewards / dones, and for every env whose
    episode ended `terminal_observation` = what the reference's step() returned, `TimeLimit.truncated` = its flag."""
    import gym_trading_env_b200 as gte
    g = H.load_golden("c3_windows_leveraged")
    env = H.make_device_env(g, final_obs=True)
    venv = gte.SB3VecEnv(env, info_keys=("portfolio_valuation", "idx"))
    n = g["params"]["n_envs"]
    assert venv.num_envs == n and venv.action_space.n == len(g["positions"])
    assert venv.observation_space.shape == g["obs0"].shape[1:]
    obs = venv.reset()
    assert isinstance(obs, np.ndarray) and obs.dtype == np.float32
    H.assert_bits(obs, g["obs0"], "reset obs")
    ended_seen = 0
    for k in range(g["actions"].shape[0]):
        venv.step_async(g["actions"][k])
        obs, rew, dones, infos = venv.step_wait()
        H.assert_bits(obs, g["obs"][k], f"step {k} obs")
        assert rew.dtype == np.float32
        np.testing.assert_allclose(rew, g["reward"][k].astype(np.float32), rtol=1e-6, atol=1e-12)
        want_done = (g["terminated"][k] | g["truncated"][k]).astype(bool)
        assert np.array_equal(dones, want_done)
        assert isinstance(infos, list) and len(infos) == n
        for i in range(n):
            assert infos[i]["idx"] == g["post_idx"][k, i]
            if want_done[i]:
                ended_seen += 1
                H.assert_bits(infos[i]["terminal_observation"], g["step_obs"][k, i], f"step {k} env {i} terminal obs")
                assert infos[i]["TimeLimit.truncated"] == bool(g["truncated"][k, i] and not g["terminated"][k, i])
            else:
                assert "terminal_observation" not in infos[i]
    assert ended_seen > 0
    assert venv.env_is_wrapped(object) == [False] * n and venv.get_attr("num_envs", [0, 1]) == [n, n]
    venv.close()


def test_save_for_render_writes_the_frame_the_reference_renderer_loads(tmp_path):
    """The pickle holds one episode: market columns joined with the History columns on the date index
    (environments.py:296-307); valuations / positions equal the golden rows of that episode."""
    import gym_trading_env_b200 as gte
    g = H.load_golden("c1_single_nowindow")
    T = int(g["lengths"][0])
    index = pd.date_range("2000-01-01", periods=T, freq="h").values
    close = np.ascontiguousarray(g["price"][0, :T])
    info = {"open": close * 1.001, "high": close * 1.002, "low": close * 0.998, "close": close}
    srs = gte.SeriesArrays(np.ascontiguousarray(g["features"][0, :T]), close,
                           [f"feature_{j}" for j in range(g["features"].shape[2])], info, index)
    kw = H.env_kwargs(g)
    env = gte.TradingVectorEnv(srs, num_envs=1, reset_plan=g["plan"], verbose=0, debug_outputs=True, name="golden", **kw)
    env.track([0], max_steps=4000)
    env.reset()
    K = g["actions"].shape[0]
    import torch
    for k in range(K):
        env.step(torch.as_tensor(g["actions"][k], device=env.device))
    path = env.save_for_render(dir=str(tmp_path), which=0, episode=0)
    df = pd.read_pickle(path)
    first_end = int(np.flatnonzero(g["terminated"][:, 0] | g["truncated"][:, 0])[0])
    assert len(df) == first_end + 2                                    # the reset row + one row per step
    for c in ("open", "high", "low", "close", "feature_0", "portfolio_valuation", "position", "real_position", "reward",
              "idx", "step", "position_index", "data_close", "data_high", "portfolio_distribution_fiat"):
        assert c in df.columns, c
    assert df.index.name == "date" and df.index.is_monotonic_increasing
    np.testing.assert_array_equal(df["idx"].to_numpy()[1:], g["idx"][:first_end + 1, 0])
    np.testing.assert_allclose(df["portfolio_valuation"].to_numpy()[1:], g["valuation"][:first_end + 1, 0], rtol=H.RTOL, atol=H.ATOL)
    np.testing.assert_array_equal(df["position"].to_numpy()[1:], g["position"][:first_end + 1, 0])
    np.testing.assert_array_equal(df["data_close"].to_numpy(), df["close"].to_numpy())
    assert path.startswith(str(tmp_path)) and "/golden_" in path


def test_per_iteration_metric_allreduce_accumulates_the_step_metrics():
    """enable_metric_allreduce(): the side-stream exchange (snapshot -> NCCL all-reduce -> accumulate) sees every
    iteration's metric vector exactly once, also in the hybrid output mode (single-rank NCCL group here; the
    multi-rank sums are checked by bench.py on every torchrun launch)."""
    import torch
    import torch.distributed as dist
    import gym_trading_env_b200 as gte
    if dist.is_initialized():
        pytest.skip("a process group already exists in this process")
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29641", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        g = H.load_golden("c3_windows_leveraged")
        for output in ("torch", "hybrid"):
            env = H.make_device_env(g, output=output)
            env.enable_metric_allreduce()
            env.reset()
            per_step = []
            for k in range(g["actions"].shape[0]):
                a = g["actions"][k] if output == "hybrid" else torch.as_tensor(g["actions"][k], device=env.device)
                env.step(a)
                per_step.append(env._metrics_step.clone())
            env.wait_metric_allreduce()
            torch.cuda.synchronize()
            assert torch.equal(env.global_metrics_step, per_step[-1])
            want = torch.zeros(8, dtype=torch.float64, device=env.device)
            for v in per_step:
                want += v
            assert torch.equal(env.global_metrics_total, want)
            assert torch.equal(env.global_metrics_total, env._metrics_total)
            assert float(env.global_metrics_total[0]) == float((g["terminated"] | g["truncated"]).sum())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("output", ["numpy", "hybrid"])
def test_narrow_integer_host_actions_are_widened_on_the_device(output):
    """int8 / int16 / int32 numpy actions (pageable or from pinned_actions(dtype)) cross PCIe as they are and give the
    results of the int64 stream, hold (-1) included."""
    import gym_trading_env_b200 as gte
    g = H.load_golden("c3_windows_leveraged")
    K = 60
    for dt in (np.int8, np.int16, np.int32):
        ref = H.make_device_env(g, output=output)
        env = H.make_device_env(g, output=output)
        ref.reset(); env.reset()
        pin = env.pinned_actions(dt)
        assert pin.dtype == dt
        for k in range(K):
            a = g["actions"][k].copy()
            a[k % a.size] = -1                                       # a hold
            want = ref.step(a)
            if k % 2:
                pin[...] = a
                got = env.step(pin)                                  # pinned, no staging copy
            else:
                got = env.step(a.astype(dt))                         # pageable
            for w, x in zip(want[:4], got[:4]):
                w = w.cpu().numpy() if hasattr(w, "cpu") else w
                x = x.cpu().numpy() if hasattr(x, "cpu") else x
                assert np.array_equal(w, x), (dt, k)
    u8 = H.make_device_env(g, output=output)
    u8.reset()
    u8.step(g["actions"][0].astype(np.uint8))                        # other integer dtypes are widened to int64 on the host
    with pytest.raises(IndexError):                                  # 255 is not a position: flagged like any bad index
        u8.step(np.full(g["params"]["n_envs"], 255, np.uint8))
        u8.check_errors()
    assert env.pinned_actions().dtype == np.int8                     # the default wire type holds Discrete(P)
    with pytest.raises(ValueError):
        env.pinned_actions(np.uint8)


def test_result_relay_between_two_processes_is_bit_exact():
    """The multi-GPU result relay (relay.py, gte_relay_*): two processes — on this box's one GPU, so the NVLink hop is a
    local copy, everything else (CUDA IPC mapping, stream wait on the sequence word, the peer's copy engine writing into
    the sender's shared pinned result block) is the real thing — each step a relayed and a plain env and compare reward /
    flags bit for bit, synchronously and with two iterations in flight, in both directions at once."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, GTE_HOST_SPIN_TIMEOUT_S="20")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", os.path.join(root, "tools", "relay_check.py"), "--same-gpu", "--envs", "65536",
           "--plan", "0>1:0.3,1>0:0.2"]
    r = subprocess.run(cmd, cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"relay_check": "ok"' in r.stdout
