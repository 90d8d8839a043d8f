"""Result relay, checked end to end by two (or more) processes: every rank steps a relayed env and a plain env with the
same seed and actions and compares reward / flags bit for bit, synchronously and with two iterations in flight.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
        tools/relay_check.py [--same-gpu] [--envs 262144] [--plan "0>1:0.3,1>0:0.2"] [--auto]

--same-gpu puts every rank on cuda:0 (the mechanism — IPC mapping, stream wait, shared pinned block — is the same; the
NVLink hop becomes a local copy) so the check also runs on a one-GPU box; the job then talks over gloo."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--same-gpu", action="store_true")
    ap.add_argument("--envs", type=int, default=1 << 18)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--windows", type=int, default=8)
    ap.add_argument("--plan", default="0>1:0.3")
    ap.add_argument("--auto", action="store_true", help="measure and plan instead of --plan")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = 0 if args.same_gpu else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo" if args.same_gpu else "nccl", rank=rank, world_size=world,
                            **({} if args.same_gpu else {"device_id": torch.device(f"cuda:{dev}")}))
    import gym_trading_env_b200 as gte
    from gym_trading_env_b200.relay import parse_forced_plan
    N = args.envs
    series = gte.frame_to_arrays(gte.make_gbm_ohlcv(20_000, seed=2))
    pos = [-1, 0, 0.5, 1]
    kw = dict(positions=pos, windows=args.windows, max_episode_duration=12, num_envs=N, seed=3 + rank, verbose=0,
              output="hybrid", host_io="copy", trading_fees=0.01 / 100, borrow_interest_rate=0.0003 / 100,
              env_id_offset=rank * N)
    out = {"rank": rank}
    for sparse in (True, False):
        plain = gte.TradingVectorEnv(series, sparse_flags=sparse, **kw)
        relayed = gte.TradingVectorEnv(series, sparse_flags=sparse, **kw)
        plan = None if args.auto else parse_forced_plan(args.plan, N)
        desc = relayed.enable_result_relay(plan=plan)
        out["plan"] = desc
        plain.reset(); relayed.reset()
        rng = np.random.default_rng(5 + rank)
        acts = rng.integers(0, len(pos), size=(args.steps, N)).astype(np.int8)
        # synchronous steps
        for k in range(args.steps // 2):
            a, b = plain.step(acts[k]), relayed.step(acts[k])
            assert a[1].tobytes() == b[1].tobytes(), f"rank {rank} sparse={sparse} step {k}: rewards differ"
            assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), f"rank {rank} step {k}: flags differ"
        # two iterations in flight
        k0 = args.steps // 2
        relayed.step_async(acts[k0])
        for k in range(k0, args.steps):
            if k + 1 < args.steps:
                relayed.step_async(acts[k + 1])
            b = relayed.step_wait()
            a = plain.step(acts[k])
            assert a[1].tobytes() == b[1].tobytes(), f"rank {rank} sparse={sparse} pipelined step {k}: rewards differ"
            assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]), f"rank {rank} pipelined step {k}: flags differ"
        assert np.count_nonzero(a[1]) > N // 2
        assert torch.equal(plain._obs.view(torch.int32), relayed._obs.view(torch.int32))
        # timing, for orientation only
        dist.barrier()
        t0 = time.perf_counter()
        for k in range(20):
            relayed.step(acts[k % args.steps])
        t1 = time.perf_counter()
        for k in range(20):
            plain.step(acts[k % args.steps])
        t2 = time.perf_counter()
        out[f"us_per_step_relayed_sparse{int(sparse)}"] = round((t1 - t0) / 20 * 1e6, 1)
        out[f"us_per_step_plain_sparse{int(sparse)}"] = round((t2 - t1) / 20 * 1e6, 1)
        relayed.close(); plain.close()
    allout = [None] * world
    dist.all_gather_object(allout, out)
    if rank == 0:
        print(json.dumps({"relay_check": "ok", "ranks": allout}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
