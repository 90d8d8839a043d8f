"""The gymnasium side of the drop-in boundary (reference: src/gym_trading_env/__init__.py:3-14 registers its
envs, callers use gym.make / gym.make_vec): when gymnasium is importable the batched env IS a
``gymnasium.vector.VectorEnv`` and `register()` adds vector entry points that `make_vec` resolves.  gymnasium is not in
this image, so both tests run in a subprocess with the test stub (oracle/gymnasium_stub) on the path — the real package
takes the same code path."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "oracle", "gymnasium_stub")


def _run(code):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([STUB, ROOT, os.environ.get("PYTHONPATH", "")]))
    p = subprocess.run([sys.executable, "-c", textwrap.dedent(code)], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    return p.stdout


def test_vector_env_subclasses_gymnasium_and_registers_vector_entry_points():
    out = _run("""
        import gymnasium
        assert "gymnasium_stub" in gymnasium.__file__          # the stand-in (the real package is absent from the image)
        import gym_trading_env_b200 as gte
        from gym_trading_env_b200 import vector_env as ve
        assert issubclass(gte.TradingVectorEnv, gymnasium.vector.VectorEnv)
        assert issubclass(gte.MultiDatasetTradingVectorEnv, gte.TradingVectorEnv)
        assert ve._Discrete is gymnasium.spaces.Discrete and ve._Box is gymnasium.spaces.Box
        gte.register()
        reg = gymnasium.envs.registration.registry
        assert reg["TradingEnv-B200-v0"]["vector_entry_point"] == "gym_trading_env_b200.vector_env:TradingVectorEnv"
        assert reg["MultiDatasetTradingEnv-B200-v0"]["vector_entry_point"].endswith(":MultiDatasetTradingVectorEnv")
        assert reg["TradingEnv-B200-v0"]["disable_env_checker"] is True and reg["TradingEnv-B200-v0"]["order_enforce"] is False
        import torch
        if not torch.cuda.is_available():
            # make_vec resolves the entry point and calls the class: without a GPU the constructor refuses loudly
            df = gte.make_gbm_ohlcv(300, seed=0)
            try:
                gymnasium.make_vec("TradingEnv-B200-v0", num_envs=4, df=df)
            except RuntimeError as e:
                assert "CUDA" in str(e) or "libgte_b200" in str(e), e
            else:
                raise AssertionError("constructed without a CUDA device")
        print("ok")
    """)
    assert out.strip().endswith("ok")


@pytest.mark.gpu
def test_make_vec_builds_a_stepping_vector_env_on_the_gpu():
    out = _run("""
        import numpy as np, torch, gymnasium
        import gym_trading_env_b200 as gte
        gte.register()
        df = gte.make_gbm_ohlcv(2000, seed=0)
        env = gymnasium.make_vec("TradingEnv-B200-v0", num_envs=256, df=df, positions=[-1, 0, 1], windows=8,
                                 trading_fees=1e-4, max_episode_duration=40, verbose=0)
        assert isinstance(env, gymnasium.vector.VectorEnv) and env.unwrapped is env and env.num_envs == 256
        assert isinstance(env.single_action_space, gymnasium.spaces.Discrete) and env.single_action_space.n == 3
        assert env.observation_space.shape == (256, 8, 10) and env.single_observation_space.shape == (8, 10)
        obs, infos = env.reset(seed=3)
        assert obs.shape == (256, 8, 10)
        for k in range(50):
            obs, rew, term, trunc, infos = env.step(torch.randint(0, 3, (256,), device=obs.device))
        assert float(env.get_metrics()["episodes"]) >= 256
        env.close()
        assert env.closed
        print("ok")
    """)
    assert out.strip().endswith("ok")


@pytest.mark.gpu
def test_infos_carry_every_reference_info_column():
    """The reference's info dict (environments.py:186-197, 253-264): idx, step, date, position_index (the step's action,
    None for a hold), position, real_position, data_<every numeric non-feature column>, portfolio_valuation,
    portfolio_distribution_*, reward."""
    import numpy as np
    import torch
    import gym_trading_env_b200 as gte
    df = gte.make_gbm_ohlcv(3000, seed=2)
    pos = [-1, 0, 0.5, 1]
    env = gte.TradingVectorEnv(df, positions=pos, windows=8, trading_fees=1e-4, borrow_interest_rate=3e-6,
                               max_episode_duration=30, num_envs=512, seed=4, verbose=0)
    obs, infos = env.reset()
    want = {"idx", "step", "date", "position_index", "position", "real_position", "data_open", "data_high", "data_low",
            "data_close", "data_volume", "portfolio_valuation", "portfolio_distribution_asset",
            "portfolio_distribution_fiat", "portfolio_distribution_borrowed_asset", "portfolio_distribution_borrowed_fiat",
            "portfolio_distribution_interest_asset", "portfolio_distribution_interest_fiat", "reward"}
    assert want <= set(infos.keys())
    # the reset row: position_index = positions.index(position), real_position = position, valuation = initial
    assert torch.equal(infos["position_index"], env._pos_idx)
    assert torch.equal(infos["portfolio_valuation"], torch.full((512,), 1000.0, dtype=torch.float64, device=env.device))
    rng = np.random.default_rng(0)
    for k in range(45):
        a = rng.integers(0, len(pos), size=512)
        a[rng.random(512) < 0.2] = -1
        act = torch.as_tensor(a, device=env.device)
        obs, rew, term, trunc, infos = env.step(act)
        idx = infos["idx"].cpu().numpy()
        for c in ("open", "high", "low", "close", "volume"):
            assert np.array_equal(infos["data_" + c].cpu().numpy(), df[c].to_numpy()[idx]), c
        assert np.array_equal(infos["date"], df.index.values[idx])
        stepped = (infos["step"] > 0).cpu().numpy()
        pi = infos["position_index"].cpu().numpy()
        assert np.array_equal(pi[stepped], a[stepped])                   # the raw action, hold -> -1
        assert np.array_equal(pi[~stepped], env._pos_idx.cpu().numpy()[~stepped])        # the reset row
        assert (~stepped).sum() == int((term | trunc).sum())
        assert np.array_equal(infos["position"].cpu().numpy(), np.array(pos)[env._pos_idx.cpu().numpy()])
        assert torch.equal(infos["reward"], rew)
    # host-output mode: numpy actions
    henv = gte.TradingVectorEnv(df, positions=pos, windows=8, max_episode_duration=30, num_envs=64, seed=4, verbose=0,
                                output="hybrid")
    henv.reset()
    a = np.array([1, -1, 3, 0] * 16, dtype=np.int8)
    _, rew, term, trunc, infos = henv.step(a)
    assert np.array_equal(infos["position_index"].cpu().numpy(), a.astype(np.int32))
    assert infos["reward"] is rew
