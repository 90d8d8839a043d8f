"""Under torchrun (one rank per GPU): how fast pinned host<->device copies go when ALL ranks copy at once — the limit
behind the end-to-end ("hybrid") number at 8 GPUs — and what the host looks like (NUMA, affinity).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 \
        tools/host_io_scaling_probe.py

Rank 0 prints one JSON line: per-direction aggregate GB/s for the C5 wire sizes (2 MB / 17 MB in, 21 MB out per rank and
iteration), every rank alone vs all ranks together, with this repo's host binding on.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gym_trading_env_b200.hostbind import bind_host_to_gpu  # noqa: E402

local_rank = int(os.environ.get("LOCAL_RANK", "0"))
bind = bind_host_to_gpu(local_rank)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N = 1 << 21
    sizes = {"h2d_int8_actions": N, "h2d_int64_actions": 8 * N, "d2h_results": 10 * N + 8}
    res = {}
    for name, nbytes in sizes.items():
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.zero_()
        devb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        h2d = name.startswith("h2d")

        def once():
            if h2d:
                devb.copy_(host, non_blocking=True)
            else:
                host.copy_(devb, non_blocking=True)

        def run(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                once()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n

        run(5)
        # every rank alone (the others idle)
        alone = [0.0] * world
        for r in range(world):
            if world > 1:
                dist.barrier()
            if r == rank:
                alone[r] = run(30)
        # all ranks together
        if world > 1:
            dist.barrier()
        together = run(30)
        t = torch.tensor([alone[rank], together], dtype=torch.float64, device=dev)
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
        else:
            g = [t]
        a = [float(x[0]) for x in g]
        tg = [float(x[1]) for x in g]
        res[name] = {"bytes_per_rank": nbytes,
                     "alone_gbs_per_rank": [round(nbytes / x / 1e9, 1) for x in a],
                     "together_gbs_per_rank": [round(nbytes / x / 1e9, 1) for x in tg],
                     "together_aggregate_gbs": round(sum(nbytes / x for x in tg) / 1e9, 1),
                     "together_ms_max": round(1e3 * max(tg), 3)}
    binds = [None] * world
    if world > 1:
        dist.all_gather_object(binds, bind)
    else:
        binds = [bind]
    if rank == 0:
        def sh(cmd):
            try:
                return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
            except Exception as e:  # noqa: BLE001
                return repr(e)
        out = {"world": world, "copies": res, "host_bind": binds, "nproc": os.cpu_count(),
               "numa_nodes": sh("ls -d /sys/devices/system/node/node* | wc -l"),
               "lscpu": sh("lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)'"),
               "mem_gb": sh("free -g | awk '/Mem/{print $2}'"), "topo": sh("nvidia-smi topo -m | head -14")}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
