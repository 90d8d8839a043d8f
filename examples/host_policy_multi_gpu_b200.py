"""A HOST policy on several B200s of one node, one process per GPU, with the result relay.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/host_policy_multi_gpu_b200.py

Every rank owns a shard of the envs (Philox keyed by the global env id: results do not depend on the GPU count).  Each
step moves int8 actions host -> device and fp64 rewards + exact flags device -> host; the observation windows stay in HBM.
With eight ranks copying at once the GPUs' PCIe links are not equally fast on this kind of host (12 vs 20 GB/s measured):
`enable_result_relay()` measures that once and lets the slow links ship part of their reward bytes over NVLink through the
GPUs with the fast links — lossless; afterwards every rank must call step() the same number of times (a collective).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gym_trading_env_b200 as gte  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    gte.bind_host_to_gpu(local)                          # first thing: host thread + pinned buffers near the GPU (if the host says where)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{local}"))
    n_per_gpu, steps = 1 << 21, 100
    env = gte.TradingVectorEnv(gte.make_gbm_ohlcv(100_000, seed=0), positions=[-3, -2, -1, 0, 1, 2, 3], windows=64,
                               trading_fees=0.01 / 100, borrow_interest_rate=0.0003 / 100, max_episode_duration=720,
                               num_envs=n_per_gpu, env_id_offset=rank * n_per_gpu, seed=0, verbose=0, output="hybrid")
    env.reset()
    if world > 1:
        plan = env.enable_result_relay()                 # collective; an empty plan (equal links) changes nothing
        if rank == 0:
            print("relay plan:", plan["plan"], "measured GB/s:", plan["measured_d2h_gbs_all_ranks_copying"])
    actions = env.pinned_actions()                       # pinned int8 [N]
    rng = np.random.default_rng(rank)
    sets = [rng.integers(0, 7, size=actions.shape).astype(np.int8) for _ in range(4)]
    for k in range(5):
        actions[...] = sets[k % 4]
        env.step(actions)
    if world > 1:
        dist.barrier()
    total, t0 = 0.0, time.perf_counter()
    for k in range(steps):
        actions[...] = sets[k % 4]                       # the "policy"
        obs, reward, terminated, truncated, infos = env.step(actions)     # numpy reward / flags on the host, CUDA obs
        total += float(reward[:1024].sum())
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{world} GPU(s) x {n_per_gpu:,} envs: {1e3 * dt.item() / steps:.3f} ms per step, "
              f"{world * n_per_gpu * steps / dt.item():.3e} env-steps/s end to end")
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
