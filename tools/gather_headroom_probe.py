"""How far is the window gather from what HBM gives a write-dominated stream?  (run on the GPU box)
fill / copy of the observation buffer with torch, the gather with and without the dynamic-feature ring reads."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gym_trading_env_b200 as gte

def timed(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

series = gte.frame_to_arrays(gte.make_gbm_ohlcv(100_000, seed=0))
N = 1 << 21
out = {}
for name, dyn in (("with_ring_F10", None), ("no_ring_F8", [])):
    kw = dict(positions=[-3, -2, -1, 0, 1, 2, 3], windows=64, trading_fees=1e-4, borrow_interest_rate=3e-6, max_episode_duration=720,
              num_envs=N, seed=1, verbose=0)
    if dyn is not None:
        kw["dynamic_feature_functions"] = dyn
    env = gte.TradingVectorEnv(series, **kw)
    env.reset()
    a = torch.randint(0, 7, (N,), device="cuda")
    for _ in range(70): env.step(a)                       # fill the ring
    ms = timed(env._launch_obs)
    nbytes = env._obs.numel() * 4
    out[name] = {"gather_ms": ms, "obs_GB": nbytes / 1e9, "obs_write_TBps": nbytes / ms / 1e9,
                 "algorithmic_TBps": (nbytes + (63 * 8 * N if dyn is None else 0)) / ms / 1e9}
    if name == "with_ring_F10":
        buf = env._obs
        out["fill_TBps"] = buf.numel() * 4 / timed(lambda: buf.zero_()) / 1e9
        half = buf.view(-1)[: buf.numel() // 2]
        other = buf.view(-1)[buf.numel() // 2:]
        out["copy_TBps_rw"] = 2 * half.numel() * 4 / timed(lambda: other.copy_(half)) / 1e9
    del env
    torch.cuda.empty_cache()
print(json.dumps(out))
