"""CPU, world_size 2 over gloo: the multi-GPU contract of the path — envs shard by
index with no data-path collective, per-env results do not depend on the number of ranks (Philox is
keyed by the GLOBAL env id), and the only exchange is an all-reduce(sum) of the 8-double metric
vector.  The per-rank engine here is the CPU oracle; the NCCL version of the same code path is
bench.py under torchrun."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, K = 96, 70
KW = dict(positions=[-1, 0, 1, 2], windows=8, trading_fees=1e-4, borrow_interest_rate=3e-6,
          portfolio_initial_value=1000, max_episode_duration=20)


def _run(env_off, n, actions):
    import gym_trading_env_b200 as gte
    import oracle as orc
    arr = gte.frame_to_arrays(gte.make_gbm_ohlcv(600, seed=4))
    env = orc.OracleVecEnv(arr.features, arr.price, num_envs=n, seed=5, env_id_offset=env_off, **KW)
    env.reset()
    vals, total = [], np.zeros(8)
    for k in range(K):
        env.step(actions[k, env_off:env_off + n])
        vals.append(env.valuation.copy())
        total += env.metrics
    return np.stack(vals), env.ep_start.copy(), total


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from gym_trading_env_b200.vector_env import shard_envs
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    actions = np.random.default_rng(0).integers(0, 4, size=(K, N_TOTAL))
    off, n = shard_envs(N_TOTAL, rank, world)
    vals, starts, metrics = _run(off, n, actions)
    t = torch.from_numpy(metrics.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)          # the path's only collective
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), vals=vals, starts=starts, off=off, n=n, reduced=t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_reproduce_the_single_rank_run(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    actions = np.random.default_rng(0).integers(0, 4, size=(K, N_TOTAL))
    vals, starts, metrics = _run(0, N_TOTAL, actions)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    assert int(parts[0]["n"]) + int(parts[1]["n"]) == N_TOTAL and int(parts[1]["off"]) == int(parts[0]["n"])
    got_vals = np.concatenate([p["vals"] for p in parts], axis=1)
    got_starts = np.concatenate([p["starts"] for p in parts])
    assert got_vals.tobytes() == vals.tobytes()            # per-env results independent of the sharding
    assert np.array_equal(got_starts, starts)              # Philox keyed by global env id
    for p in parts:
        assert np.array_equal(p["reduced"][:3], metrics[:3]) and p["reduced"][5] == metrics[5]
        np.testing.assert_allclose(p["reduced"], metrics, rtol=1e-12, atol=1e-12)
    assert metrics[0] > N_TOTAL                             # several episodes per env finished
