"""GPU: BASELINE.json's FULL sizes (C3: 65 536 envs; C5 shard: 2^21 envs; windows=64, T=100 000, 8+2
features, D=720), where the scalar oracle would take minutes, checked through size-independent
properties instead — each one exact (bitwise) unless stated:

  P1 gather       : the static columns of every observation are exactly rows idx-63..idx of the feature
                    table (recomputed with torch indexing), the newest dynamic row is exactly
                    (float32(position), float32(real_position)) of the step outputs, rows before the
                    episode start are zero;
  P2 valuation    : the kernel's valuation equals the reference formula recomputed from the SoA state
                    with un-fused torch fp64 ops in the same order (asset*p + fiat - ia*p - if);
  P3 bookkeeping  : step/idx advance by one except where an episode ended, where the env restarts inside
                    [W-1, T-D-(W-1)); flags agree with their definition; metrics are the sums of the
                    per-env outputs (counts exact, fp64 sums to 1e-9);
  P4 sharding     : an arbitrary env slice run alone with env_id_offset reproduces the full run bit for
                    bit (Philox keyed by the global env id -> results independent of the GPU count);
  P5 determinism  : a second run from the same seed is bit-identical;
  P6 oracle sample: 64 randomly chosen envs of the full run match the CPU oracle run on exactly those
                    global env ids (bit-exact obs/valuation, reward to 1e-12).
"""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

POS = [-3, -2, -1, 0, 1, 2, 3]
KW = dict(positions=POS, windows=64, trading_fees=0.01 / 100, borrow_interest_rate=0.0003 / 100,
          portfolio_initial_value=1000, max_episode_duration=720)
T = 100_000


@pytest.fixture(scope="module")
def series():
    import gym_trading_env_b200 as gte
    return gte.frame_to_arrays(gte.make_gbm_ohlcv(T, seed=0))


def _actions(n, k, device):
    g = torch.Generator(device=device)
    g.manual_seed(99)
    return torch.randint(0, len(POS), (k, n), generator=g, device=device, dtype=torch.int64)


def _check_step(env, prev_idx, prev_step, first):
    """P1-P3 on the current outputs of `env` (everything stays on the device)."""
    dev = env.device
    N, W, ns = env.num_envs, 64, env._n_static
    idx = (env._ep_start + env._step).long()
    # ---- P1: static window == table rows; sampled to bound the temporary (N_s x 64 x 8 floats)
    sel = torch.arange(0, N, max(1, N // 131072), device=dev)
    rows = idx[sel, None] - (W - 1) + torch.arange(W, device=dev)[None, :]
    want = env._features[0][rows]                                             # [n_s, 64, 8]
    got = env._obs[sel]
    assert torch.equal(got[:, :, :ns].view(torch.int32), want.view(torch.int32)), "static window != table rows"
    # newest dynamic row, for every env
    pos_f32 = torch.tensor(POS, dtype=torch.float64, device=dev)[env._pos_idx.long()].float()
    assert torch.equal(env._obs[:, W - 1, ns].view(torch.int32), pos_f32.view(torch.int32))
    started = env._step == 0                                                  # reset rows hold (position, position)
    rp_f32 = torch.where(started, pos_f32, env._real_position.float())
    assert torch.equal(env._obs[:, W - 1, ns + 1].view(torch.int32), rp_f32.view(torch.int32))
    # rows before the episode start are zero
    before = rows < env._ep_start[sel, None]
    assert (got[:, :, ns:][before] == 0).all()
    # ---- P2: valuation of envs that did not reset this step, from the SoA state (same op order, no FMA in torch)
    p = env._price[0][idx]
    val = ((env._asset * p + env._fiat) + (-env._interest_asset) * p) + (-env._interest_fiat)
    live = ~started
    assert torch.equal(val[live], env._valuation[live]), "valuation != formula(state)"
    # ---- P3: bookkeeping
    term, trunc = env._terminated.bool(), env._truncated.bool()
    ended = term | trunc
    if not first:
        assert torch.equal(env._step[~ended], prev_step[~ended] + 1) and torch.equal(idx[~ended], prev_idx[~ended] + 1)
        assert (env._step[ended] == 0).all()
        assert (idx[ended] >= W - 1).all() and (idx[ended] < T - 720 - (W - 1)).all()
        assert torch.equal(term, (env._valuation / 1000.0) <= 0.7)
        assert torch.equal(trunc, (prev_idx + 1 >= T - 1) | (prev_step + 1 >= 719))
        assert (env._reward[term] == 0).all()
    m = env._metrics_step
    assert m[0].item() == ended.sum().item() and m[1].item() == term.sum().item() and m[2].item() == trunc.sum().item()
    torch.testing.assert_close(m[6], env._reward.sum(), rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(m[3], (env._valuation[ended] / 1000.0 - 1.0).sum(), rtol=1e-9, atol=1e-9)
    return idx.clone(), env._step.clone()


@pytest.mark.parametrize("n_envs,n_steps", [(65_536, 40), (1 << 21, 6)])
def test_full_size_properties(series, n_envs, n_steps):
    import gym_trading_env_b200 as gte
    env = gte.TradingVectorEnv(series, num_envs=n_envs, seed=7, verbose=0, debug_outputs=True, **KW)
    assert env.obs_variant == "tma"
    acts = _actions(n_envs, n_steps, env.device)
    env.reset()
    # shorten the first episodes so that auto-resets happen inside the few steps we can afford:
    # pretend every env is already 700..718 steps into its episode (ep_start stays put)
    env._step.copy_(torch.randint(700, 719, (n_envs,), device=env.device, dtype=torch.int32))
    idx0 = (env._ep_start + env._step).clone()
    step0 = env._step.clone()
    # P4/P5/P6 companions start from the same (modified) state
    lo, n_slice = 12_352, 4_096         # a multiple of 32: the ring is stored per 32-env tile
    sl = gte.TradingVectorEnv(series, num_envs=n_slice, seed=7, env_id_offset=lo, verbose=0, debug_outputs=True, **KW)
    sl.reset()
    for name in ("asset", "fiat", "interest_asset", "interest_fiat", "pos_idx", "step", "ep_start"):
        getattr(sl, "_" + name).copy_(getattr(env, "_" + name)[lo:lo + n_slice])
    sl._dyn_ring.copy_(env._dyn_ring[lo // 32:(lo + n_slice) // 32])       # one ring block per tile of 32 envs

    prev_idx, prev_step = idx0, step0
    total_eps = 0
    for k in range(n_steps):
        env.step(acts[k])
        sl.step(acts[k][lo:lo + n_slice])
        prev_idx, prev_step = _check_step(env, prev_idx, prev_step, first=False)
        total_eps += int(env._metrics_step[0].item())
        # P4: the slice, keyed by global env ids, reproduces the full run
        for name in ("asset", "fiat", "pos_idx", "step", "ep_start", "reward", "valuation", "terminated", "truncated"):
            assert torch.equal(getattr(sl, "_" + name), getattr(env, "_" + name)[lo:lo + n_slice]), (k, name)
        assert torch.equal(sl._obs.view(torch.int32), env._obs[lo:lo + n_slice].view(torch.int32)), k
    assert total_eps > n_envs * 0.3 * min(1.0, n_steps / 19)      # the shortened episodes really end and restart
    env.check_errors()
    # P5: determinism
    env2 = gte.TradingVectorEnv(series, num_envs=n_envs, seed=7, verbose=0, **KW)
    env2.reset()
    env2._step.copy_(step0)
    for k in range(min(n_steps, 3)):
        env2.step(acts[k])
    env3 = gte.TradingVectorEnv(series, num_envs=n_envs, seed=7, verbose=0, **KW)
    env3.reset()
    env3._step.copy_(step0)
    for k in range(min(n_steps, 3)):
        env3.step(acts[k])
    assert torch.equal(env2._obs.view(torch.int32), env3._obs.view(torch.int32))
    assert torch.equal(env2._valuation, env3._valuation) and torch.equal(env2._reward, env3._reward)
    assert torch.equal(env2._metrics_total, env3._metrics_total)            # deterministic reduction order


def test_sampled_envs_of_a_full_size_run_match_the_oracle(series):
    """P6: 128 envs spread over a 65 536-env run (C3 size) vs the CPU oracle stepping exactly those
    global env ids (a contiguous block so the oracle's Philox keys line up), 60 steps with resets."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    n_envs, lo, n_s, K = 65_536, 40_000, 128, 60
    kw = dict(KW, max_episode_duration=25)
    env = gte.TradingVectorEnv(series, num_envs=n_envs, seed=11, verbose=0, debug_outputs=True, **kw)
    o = orc.OracleVecEnv(series.features, series.price, num_envs=n_s, seed=11, env_id_offset=lo, **kw)
    obs, _ = env.reset()
    H.assert_bits(obs[lo:lo + n_s].cpu().numpy(), o.reset(), "reset obs")
    acts = _actions(n_envs, K, env.device)
    for k in range(K):
        env.step(acts[k])
        o.step(acts[k][lo:lo + n_s].cpu().numpy())
        H.assert_bits(env._obs[lo:lo + n_s].cpu().numpy(), o.obs, f"step {k} obs")
        H.assert_bits(env._valuation[lo:lo + n_s].cpu().numpy(), o.valuation, f"step {k} valuation")
        H.assert_bits(env._ep_start[lo:lo + n_s].cpu().numpy(), o.ep_start, f"step {k} ep_start")
        H.assert_close64(env._reward[lo:lo + n_s].cpu().numpy(), o.reward, f"step {k} reward")
        H.assert_bits(env._terminated[lo:lo + n_s].cpu().numpy(), o.terminated, f"step {k} terminated")
