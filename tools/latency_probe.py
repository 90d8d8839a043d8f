"""Small-N latency probe: per-iteration time of env.step() from Python vs env.rollout() (one C call) for several N."""
import sys, time
import torch
sys.path.insert(0, ".")
import gym_trading_env_b200 as gte

f, p = gte.make_gbm_arrays(100_000, seed=0)
arr = gte.SeriesArrays(f, p, [f"feature_{j}" for j in range(f.shape[1])], {}, None)
for windows in (None, 64):
    for N in (4096, 65536):
        env = gte.TradingVectorEnv(arr, positions=[-1, 0, 0.5, 1], windows=windows, trading_fees=1e-4, borrow_interest_rate=3e-6,
                                   max_episode_duration=720 if windows else "max", num_envs=N, seed=1, verbose=0)
        env.reset()
        K = 1000
        acts = torch.randint(0, 4, (K, N), device=env.device)
        for mode in ("step", "rollout", "rollout_keep"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            if mode == "step":
                for k in range(K):
                    env.step(acts[k])
            else:
                env.rollout(acts, keep_obs=(mode == "rollout_keep" and windows is None))
            t_issue = time.perf_counter() - t0
            e1.record()
            torch.cuda.synchronize()
            print(f"windows={windows} N={N} {mode:13s} device {1e3 * e0.elapsed_time(e1) / K:7.2f} us/iter   host issue {1e6 * t_issue / K:7.2f} us/iter", flush=True)
