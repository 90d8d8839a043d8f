"""End-to-end time of a host-policy step at 8 GPUs against the FRACTION of the slow ranks' rewards the relay ships
(forced plans: ranks 0-3 send to ranks 4-7), one step at a time and with two iterations in flight.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/relay_sweep.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
    import gym_trading_env_b200 as gte
    N, steps = 1 << 21, 40
    env = gte.TradingVectorEnv(gte.make_gbm_ohlcv(100_000, seed=0), positions=[-3, -2, -1, 0, 1, 2, 3], windows=64,
                               trading_fees=0.01 / 100, borrow_interest_rate=0.0003 / 100, max_episode_duration=720,
                               num_envs=N, env_id_offset=rank * N, seed=0, verbose=0, output="hybrid")
    env.reset()
    rng = np.random.default_rng(rank)
    sets, keep = [], []
    for _ in range(4):
        t = torch.empty(N, dtype=torch.int8, pin_memory=True)       # pinned: read by the GPU in place
        t.numpy()[...] = rng.integers(0, 7, size=N)
        keep.append(t)
        sets.append(t.numpy())

    def timed(fn):
        fn(5)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        fn(steps)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return 1e3 * t.item() / steps

    def sync(n):
        for k in range(n):
            env.step(sets[k % 4])

    def piped(n):
        env.step_async(sets[0])
        for k in range(n):
            if k + 1 < n:
                env.step_async(sets[(k + 1) % 4])
            env.step_wait()

    half = world // 2
    for frac in (0.0, 0.10, 0.15, 0.20, 0.25, 0.30, 0.35):
        x = int(frac * N) // 1024 * 1024
        if x > 0:
            env.enable_result_relay(plan={s: (s + half, x) for s in range(half)})
        else:
            env.close_extras()                                         # fraction 0: leave the relay (collective)
        row = {"fraction": frac, "ms_one_step_at_a_time": round(timed(sync), 4), "ms_two_in_flight": round(timed(piped), 4)}
        if rank == 0:
            print(json.dumps(row), flush=True)
    env.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
