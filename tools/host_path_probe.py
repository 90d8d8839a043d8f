"""Where the time of one host-policy step goes at small N (run on the GPU box):

    python tools/host_path_probe.py [--envs 4096 8192 ...] [--windows 0]

For every env count: us per step of (a) the raw ctypes call of gte_step_host with pre-built arguments, mapped and
copy-engine IO; (b) env.step(pinned array) through the Python wrapper; (c) the device-resident loop
(gte_step_obs + stream synchronize) as the no-host-IO floor; (d) the launch floor of an empty stream synchronize.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import gym_trading_env_b200 as gte  # noqa: E402
from gym_trading_env_b200 import _cabi  # noqa: E402


def timed(fn, n, warm=50):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, nargs="+", default=[4096, 8192, 16384, 32768, 65536])
    ap.add_argument("--windows", type=int, default=0)
    ap.add_argument("--iters", type=int, default=3000)
    args = ap.parse_args()
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(100_000, seed=0))
    out = []
    for N in args.envs:
        kw = dict(positions=[-1, 0, 0.5, 1], windows=args.windows or None, trading_fees=1e-4, borrow_interest_rate=3e-6,
                  portfolio_initial_value=1000, max_episode_duration="max" if not args.windows else 720, num_envs=N,
                  seed=1, verbose=0)
        row = {"envs": N, "windows": args.windows}
        for mode in ("mapped", "copy") + (("server",) if not args.windows and N <= 32768 else ()):
            env = gte.TradingVectorEnv(series, output="hybrid", host_io=mode, **kw)
            env.reset()
            pin = env.pinned_actions()
            pin[...] = np.random.default_rng(0).integers(0, 4, N)
            row[f"env_step_{mode}_us"] = timed(lambda: env.step(pin), args.iters)
            # raw C call with the wrapper's own pre-built arguments
            io, f = env._io, env._fast_args
            io.actions = pin.ctypes.data
            env._P.action_bytes = 1
            stream = torch.cuda.current_stream().cuda_stream
            lib = env._lib
            row[f"raw_call_{mode}_us"] = timed(lambda: lib.gte_step_host(f[0], f[1], f[2], env._io_ref, f[3], f[4], 1, 0,
                                                                       env._io_mode_ref, stream), args.iters)
            env._P.action_bytes = 0
        dev = gte.TradingVectorEnv(series, **kw)
        dev.reset()
        a = torch.randint(0, 4, (N,), device="cuda", dtype=torch.int64)

        def dev_step():
            dev.step(a)
            torch.cuda.current_stream().synchronize()
        row["device_loop_sync_us"] = timed(dev_step, args.iters)
        row["device_loop_async_us"] = timed(lambda: dev.step(a), args.iters)
        row["sync_only_us"] = timed(lambda: torch.cuda.current_stream().synchronize(), args.iters)
        out.append(row)
        print(json.dumps(row), flush=True)
    return out


if __name__ == "__main__":
    main()
