"""Stress the one-launch (fused) form against the two-launch form: many seeds / sizes, report the first mismatching field."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gym_trading_env_b200 as gte

series = gte.frame_to_arrays(gte.make_gbm_ohlcv(20_000, seed=4))
FEES = dict(trading_fees=1e-4, borrow_interest_rate=3e-6, portfolio_initial_value=1000)
bad = 0
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for rnd in range(rounds):
    for n_envs, pos in ((100_000, [-1, 0, 1]), (65_536, [-1, 0, 0.5, 1]), (37_777, [-3, -2, -1, 0, 1, 2, 3])):
        kw = dict(positions=pos, windows=64, max_episode_duration=20, num_envs=n_envs, seed=21 + rnd, verbose=0,
                  debug_outputs=True, **FEES)
        fused, plain = gte.TradingVectorEnv(series, n_chunks=0, **kw), gte.TradingVectorEnv(series, n_chunks=1, **kw)
        fused.reset(); plain.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(rnd)
        acts = torch.randint(0, len(pos), (60, n_envs), generator=g, device="cuda")
        for k in range(60):
            fused.step(acts[k]); plain.step(acts[k])
            for nm in ("obs", "reward", "valuation", "real_position", "terminated", "truncated", "asset", "fiat",
                       "interest_asset", "interest_fiat", "pos_idx", "step", "ep_start", "info_idx", "info_step", "dyn_ring"):
                a, b = getattr(fused, "_" + nm), getattr(plain, "_" + nm)
                a = a.view(torch.int32) if a.dtype == torch.float32 else a
                b = b.view(torch.int32) if b.dtype == torch.float32 else b
                if not torch.equal(a, b):
                    diff = (a != b).nonzero()
                    print(f"MISMATCH round {rnd} envs {n_envs} step {k} field {nm}: {diff.shape[0]} elements, first {diff[:4].tolist()}", flush=True)
                    bad += 1
                    break
            else:
                if not torch.equal(fused._metrics_step[:3], plain._metrics_step[:3]):
                    print(f"MISMATCH round {rnd} envs {n_envs} step {k} metrics {fused._metrics_step.tolist()} {plain._metrics_step.tolist()}", flush=True)
                    bad += 1
                continue
            break
        del fused, plain
print("stress done, mismatches:", bad)
