"""bench.py — env-steps/s of the batched trading-env hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c5|c3|c2|c4]

Metric (BASELINE.json): env-steps/sec, whole box, device-timed; % of HBM roofline.
A "step" is ONE lockstep iteration of the hot path over this rank's envs: the fused step kernel
(trade / interest / valuation / reward / flags / in-place auto-reset) + the observation-window gather
(+ the NCCL allreduce of the episode-metric vector when N > 1).  Default workload = BASELINE config 5
sharded: 2^21 envs per GPU (2^24 over 8), positions -3..3, windows=64, 8 static + 2 dynamic
features, max_episode_duration=720, fees 0.01 %, borrow 0.0003 %/step, synthetic GBM T=100 000.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with everything resident in HBM;
`e2e` = the same metric through the public VectorEnv API with HOST numpy buffers (pinned H2D of the
actions, D2H of obs/reward/flags inside the timed region); `roofline` = the window-gather kernel
(the dominant kernel) timed live with CUDA events; `cpu_baseline` = the scalar C oracle port
(oracle/gte_oracle.c) on this box's host cores on a bounded sample of the same workload.
`--impl reference` times that CPU port alone (all host threads) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "env-steps/sec (whole box, device-timed)"
UNIT = "env-steps/s"

WORKLOADS = {
    # name: (envs per GPU, positions, windows, max_episode_duration, n_datasets, rows per dataset)
    "c5": dict(envs=1 << 21, positions=[-3, -2, -1, 0, 1, 2, 3], windows=64, duration=720, n_datasets=1, rows=100_000,
               label="C5 shard: 2^21 envs/GPU (2^24 over 8), positions -3..3, windows=64, 8+2 features, D=720, GBM T=100k"),
    "c3": dict(envs=65_536, positions=[-1, 0, 0.5, 1], windows=64, duration=720, n_datasets=1, rows=100_000,
               label="C3: 65,536 envs, windows=64, 8+2 features, D=720, GBM T=100k"),
    "c2": dict(envs=4096, positions=[-1, 0, 0.5, 1], windows=None, duration="max", n_datasets=1, rows=100_000,
               label="C2: 4096 lockstep envs, windows=None, GBM T=100k"),
    "c4": dict(envs=1 << 20, positions=[-1, 0, 0.5, 1], windows=64, duration=720, n_datasets=32, rows=1_000_000,
               label="C4: MultiDataset 32 x 1M rows, per-env dataset index, 2^20 envs/GPU, windows=64, D=720"),
}
FEE, RATE, V0 = 0.01 / 100, 0.0003 / 100, 1000.0


def algorithmic_bytes(windows, n_static=8, n_dyn=2):
    """compulsory HBM bytes per env-step (dataset rows treated as cache-resident)."""
    W = 1 if windows is None else windows
    step = 2 * 44 + 8 + 10                                   # state r+w, action, reward+flags
    obs = W * (n_static + n_dyn) * 4 + (W - 1) * n_dyn * 4   # window written + prior dynamic rows re-read
    return step, obs


def make_series(wl, gte):
    if wl["n_datasets"] == 1:
        return [gte.frame_to_arrays(gte.make_gbm_ohlcv(wl["rows"], seed=0))]
    out = []
    for k in range(wl["n_datasets"]):
        f, p = gte.make_gbm_arrays(wl["rows"], seed=k)
        out.append(gte.SeriesArrays(f, p, [f"feature_{j}" for j in range(f.shape[1])], {}, None))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t_end = time.time() + 2.0                 # wait for the first row: nvidia-smi needs tens of ms to start
            while not self.rows and time.time() < t_end:
                time.sleep(0.005)
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Rows sampled inside [t0, t1] (host clock; one sampling period of slack on both sides); when the timed
        region is shorter than the sampler's period, the rows taken under the same load since start()."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 is None or (t0 - 0.025 <= t <= t1 + 0.025)]
        window = "timed region"
        if not rows:
            rows, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than the 20 ms sampling period)"
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


CPU_SAMPLE_ENVS = 4096
CPU_SAMPLE_NOTE = ("timed after a 60 000-iteration burn-in (steady state of the private columns); the %d-env sample's state and observation batch are cache-resident on the host (10 MB of observations against "
                   "2 MB of L2 per core), which flatters the CPU; each env keeps its own private dynamic-feature columns like the "
                   "reference (3.2 GB, pre-faulted)" % CPU_SAMPLE_ENVS)


def make_cpu_port(wl, n_sample, threads, seed=0, burn_in=60_000):
    """The scalar C port of the reference (oracle/gte_oracle.c) set up on a bounded sample of the workload."""
    import gym_trading_env_b200 as gte
    import oracle as orc
    series = make_series(dict(wl, n_datasets=1, rows=min(wl["rows"], 100_000)), gte)[0]
    env = orc.OracleVecEnv(series.features, series.price, num_envs=n_sample, positions=wl["positions"],
                           windows=wl["windows"], trading_fees=FEE, borrow_interest_rate=RATE,
                           portfolio_initial_value=V0, max_episode_duration=wl["duration"], seed=seed,
                           threads=threads, dyn_mode=0)
    env.prefault()
    env.reset()
    rng = np.random.default_rng(1234)
    acts = rng.integers(0, len(wl["positions"]), size=(16, n_sample))
    # burn-in to the steady state: every env keeps PRIVATE dynamic-feature columns (like a reference env object), and
    # only after ~100 episodes has it written rows all over them — a fresh sample runs ~20 % faster than the long-run rate
    if burn_in:
        env.rollout_timed(acts, burn_in)
    return env, acts


def cpu_port_rate(wl, n_sample, seconds, threads, seed=0, repeats=5, burn_in=60_000):
    """env-steps/s of the C port on `threads` pinned host threads (pthreads created and timed inside C):
    median of `repeats` equal slices of ~`seconds` of work, with the spread."""
    env, acts = make_cpu_port(wl, n_sample, threads, seed, burn_in=burn_in)
    env.rollout_timed(acts, 50)                                     # warm-up
    per_iter = max(env.rollout_timed(acts, 50) / 50, 1e-7)
    iters = int(max(50, min(200000, seconds / repeats / per_iter)))
    rates = [n_sample * iters / env.rollout_timed(acts, iters) for _ in range(repeats)]
    return {"value": statistics.median(rates), "min": min(rates), "max": max(rates), "repeats": repeats,
            "iters_per_repeat": iters, "spread": (max(rates) - min(rates)) / statistics.median(rates)}


# ---- the UNMODIFIED Python reference, when the driver-style install under baseline/_ref is present -------------
def _pyref_worker(args):
    """One host process: n reference TradingEnv objects stepped in a sync-style lockstep loop for `seconds`."""
    n_envs, seconds, wl, seed = args
    import warnings
    sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    try:
        import gymnasium  # noqa: F401
    except Exception:  # noqa: BLE001
        sys.path.insert(0, os.path.join(ROOT, "oracle", "gymnasium_stub"))     # 4-name stand-in (see its docstring)
    import gym_trading_env_b200 as gte
    from gym_trading_env.environments import TradingEnv
    warnings.resetwarnings()
    df = gte.make_gbm_ohlcv(min(wl["rows"], 100_000), seed=0)
    np.random.seed(seed)
    envs = [TradingEnv(df=df, positions=wl["positions"], windows=wl["windows"], trading_fees=FEE,
                       borrow_interest_rate=RATE, portfolio_initial_value=V0,
                       max_episode_duration=wl["duration"], verbose=0) for _ in range(n_envs)]
    obs = [e.reset()[0] for e in envs]
    rng = np.random.default_rng(seed)
    n_pos = len(wl["positions"])
    steps = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        acts = rng.integers(0, n_pos, size=n_envs)
        for i, e in enumerate(envs):                      # what gymnasium's SyncVectorEnv does
            o, r, term, trunc, info = e.step(int(acts[i]))
            if term or trunc:
                o, info = e.reset()
            obs[i] = o
        np.stack(obs)
        steps += n_envs
    return steps, time.perf_counter() - t0


def python_reference_rate(wl, workload_name, seconds=6.0):
    """env-steps/s of the unmodified reference (pip-installed into baseline/_ref) on this box's host cores:
    one sync-style loop on one core, and one such loop per core (separate processes, hard time-outs).
    None when the install is absent."""
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "gym_trading_env")):
        return None
    cores = len(os.sched_getaffinity(0))
    out = {"driver": "shim: gymnasium is not installed, so its two vector drivers are re-created — 'one_core' = the sequential "
                     "SyncVectorEnv loop over reference TradingEnv objects in one process, 'all_cores' = one such worker "
                     "process per core (AsyncVectorEnv-style); the env code is the unmodified reference (4-name gymnasium "
                     "stub from oracle/gymnasium_stub)", "envs_per_process": 8, "cores": cores}

    def run(n_proc):
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", workload_name, "--pyref-worker", str(seconds)]
        procs = [subprocess.Popen(cmd + ["--pyref-seed", str(k)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                  text=True) for k in range(n_proc)]
        res = []
        for p in procs:
            try:
                o, _ = p.communicate(timeout=seconds + 180)
                res.append(json.loads(o.strip().splitlines()[-1]))
            except Exception:  # noqa: BLE001
                p.kill()
        if len(res) != n_proc:
            raise RuntimeError(f"{n_proc - len(res)} of {n_proc} reference workers failed")
        return sum(r["steps"] for r in res) / max(r["seconds"] for r in res)
    try:
        out["one_core"] = run(1)
        out["all_cores"] = run(cores)
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)[:200]
    return out


def run_reference_arm(args, wl, rank):
    """`--impl reference`: the reference algorithm's CPU port (oracle/gte_oracle.c, scalar C, faithful per-env private
    dynamic-feature columns) on all host cores — pinned pthreads created and timed inside C, the same driver and the
    same sample the GPU arm's `cpu_baseline` leg uses.  Rank 0 only.  One bench "step" = `inner` lockstep iterations
    of the sample; `value` follows the MEDIAN step time (min / max in `spread`)."""
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n_sample = CPU_SAMPLE_ENVS
    env, acts = make_cpu_port(wl, n_sample, cores)
    inner = 1000                                  # lockstep iterations per bench "step" (bounded sample, ~50 ms)
    for _ in range(max(args.warmup, 3)):
        env.rollout_timed(acts, inner)
    times = [env.rollout_timed(acts, inner) for _ in range(args.steps)]
    med = statistics.median(times)
    value = n_sample * inner / med
    sample = (f"{n_sample} envs x {inner} lockstep iterations per step, {cores} pinned host threads (pthreads inside C), "
              f"scalar C port of the reference; {CPU_SAMPLE_NOTE}")
    spread = {"min": n_sample * inner / max(times), "max": n_sample * inner / min(times),
              "rel": (max(times) - min(times)) / med, "steps": len(times)}
    _emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * med,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"], "sample": sample}, "spread": spread,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "single_core_value": cpu_port_rate(wl, 256, 2.0, 1, repeats=3, burn_in=0)["value"],
                         "python_reference": python_reference_rate(wl, args.workload)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


_RESULT_FD = None


def _claim_stdout():
    """Keep stdout for the ONE result line: everything else that writes to fd 1 — NCCL's version banner, library
    chatter, worker output — is sent to stderr from here on."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed iterations (default 200; 2000 / 5000 for the small c3 / c2 workloads, whose "
                         "iterations take tens of microseconds, so that the clocks can be sampled during the timed region)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c5", choices=list(WORKLOADS))
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--obs-variant", default="auto")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cuda-graph", action="store_true", help="replay one captured lockstep iteration (small N)")
    ap.add_argument("--pyref-worker", type=float, default=None, help=argparse.SUPPRESS)
    ap.add_argument("--pyref-seed", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-relay", action="store_true", help="multi-GPU e2e legs without enable_result_relay()")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the short C2 / C3 / C4 legs of the default run")
    ap.add_argument("--repeats", type=int, default=3, help="the timed region of K steps is repeated at least this many times; the median counts")
    ap.add_argument("--min-timed-seconds", type=float, default=0.3,
                    help="keep repeating the K-step region (up to 25 times) until this much device time is covered")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"c3": 2000, "c2": 5000}.get(args.workload, 200) if args.impl == "b200" else 200
    wl = dict(WORKLOADS[args.workload])
    if args.envs_per_gpu:
        wl["envs"] = args.envs_per_gpu
    if args.pyref_worker is not None:                      # internal: one reference worker process
        st, dt = _pyref_worker((8, args.pyref_worker, wl, args.pyref_seed))
        print(json.dumps({"steps": st, "seconds": dt}), flush=True)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _claim_stdout()

    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return

    # first thing in the process: keep this rank's host thread (and the pinned buffers it will allocate) on the
    # CPU cores / memory node of its GPU — eight ranks share one host in the scaling runs
    from gym_trading_env_b200.hostbind import bind_host_to_gpu
    host_bind = bind_host_to_gpu(local_rank)

    import torch
    import torch.distributed as dist
    import gym_trading_env_b200 as gte

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Ctx(torch=torch, dist=dist, gte=gte, dev=dev, rank=rank, world=world, local_rank=local_rank, args=args)

    sampler = ClockSampler(local_rank)
    env, actions = build_env(ctx, wl, args.workload)
    if world > 1:
        # C5: NCCL all-reduce(sum) of the 8 fp64 episode metrics EVERY iteration, issued by the env between its two
        # kernels on a side stream (overlaps the gather; the next iteration only waits for the 64-byte snapshot)
        env.enable_metric_allreduce()
    if rank == 0:
        sampler.start()                      # already running (and past its start-up) when the timed region begins
    m = measure(ctx, env, actions, wl, args.workload, args.steps, max(args.warmup, 3), repeats=args.repeats, sampler=sampler)
    N = wl["envs"]
    if world > 1:
        # the per-iteration all-reduces must add up to the all-reduced local totals (counts exactly, fp sums to rounding)
        torch.cuda.synchronize()
        tot = env._metrics_total.clone()
        dist.all_reduce(tot)
        got = env.global_metrics_total
        if not (torch.equal(got[:3], tot[:3]) and torch.equal(got[5], tot[5]) and torch.allclose(got, tot, rtol=1e-9, atol=1e-9)):
            raise RuntimeError(f"per-iteration metric all-reduce disagrees with the local totals: {got.tolist()} vs {tot.tolist()}")

    latency = latency_probe(ctx, env, actions, wl, m) if (wl["envs"] < 2 ** 20 and world == 1 and not args.cuda_graph) else None
    e2e = e2e_legs(ctx, env, actions, wl) if not args.no_e2e else {}
    env.close()
    del env

    # ---- the other BASELINE configs, short legs in the same process (N=1 default run only) ----
    configs = None
    if rank == 0 and world == 1 and args.workload == "c5" and not args.no_configs and not args.envs_per_gpu:
        configs = {}
        for name in ("c2", "c3", "c4"):
            try:
                configs[name] = config_leg(ctx, name)
            except Exception as e:  # noqa: BLE001
                configs[name] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N=1 only): scalar C port of the reference on the host cores ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_leg(args.workload)

    if rank == 0:
        out = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": m["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["label"], "envs_per_gpu": N, "obs_variant": m["obs_variant"], "chunks": m["chunks"],
                       "cuda_graph": bool(args.cuda_graph), "l2_policy": m["l2_policy"],
                       "parallelism": f"env-sharded x{world}, dataset replicated, NCCL allreduce of 8 fp64 metrics per iteration",
                       "host_bind": host_bind},
            "repeats": m["repeats"], "spread": m["spread"],
            "clocks": m["clocks"], "e2e": e2e.get("e2e"), "e2e_gymnasium_dtypes": e2e.get("e2e_gymnasium_dtypes"),
            "e2e_pipelined": e2e.get("e2e_pipelined"), "e2e_no_relay": e2e.get("e2e_no_relay"),
            "e2e_f32_reward": e2e.get("e2e_f32_reward"),
            "e2e_other_host_io": e2e.get("e2e_other_host_io"), "e2e_server": e2e.get("e2e_server"),
            "e2e_full_obs_to_host": e2e.get("e2e_full_obs_to_host"),
            "gpu_launches": m["launches_per_step"] * args.steps * m["repeats"]["n"], "roofline": m["roofline"], "cpu_baseline": cpu,
        }
        if latency is not None:
            out["latency"] = latency
        if configs is not None:
            out["configs"] = configs
        _emit(out)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_leg(workload_name):
    """The `cpu_baseline` object of the GPU arm: the SAME measurement as `--impl reference`, run the same way — in a
    fresh process (a process that holds a CUDA context, torch's thread pools and gigabytes of pinned memory times the
    very same C code ~20 % slower), 12 bench steps of 1000 lockstep iterations after the burn-in."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload_name,
           "--steps", "12", "--warmup", "3"]
    try:
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        d = json.loads(p.stdout.strip().splitlines()[-1])
        cb = d["cpu_baseline"]
        cb["spread"] = d.get("spread")
        cb["how"] = "subprocess: " + " ".join(cmd[1:])
        return cb
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                "sample": "failed: " + repr(e)[:200]}


class Ctx:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def build_env(ctx, wl, workload_name, **over):
    gte, torch = ctx.gte, ctx.torch
    N = wl["envs"]
    series = make_series(wl, gte)
    kw = dict(positions=wl["positions"], windows=wl["windows"], trading_fees=FEE, borrow_interest_rate=RATE,
              portfolio_initial_value=V0, max_episode_duration=wl["duration"], num_envs=N, device=ctx.dev,
              seed=2024, env_id_offset=ctx.rank * N, obs_variant=ctx.args.obs_variant, verbose=0,
              cuda_graph=ctx.args.cuda_graph)
    kw.update(over)
    if wl["n_datasets"] > 1:
        env = gte.MultiDatasetTradingVectorEnv(datasets=series, **kw)
    else:
        env = gte.TradingVectorEnv(series[0], **kw)
    gen = torch.Generator(device=ctx.dev)
    gen.manual_seed(1234 + ctx.rank)
    actions = torch.randint(0, len(wl["positions"]), (N_ACTION_SETS, N), generator=gen, device=ctx.dev, dtype=torch.int64)
    env.reset()
    return env, actions


N_ACTION_SETS = 8


def _barrier(ctx):
    ctx.torch.cuda.synchronize()
    if ctx.world > 1:
        ctx.dist.barrier()
    ctx.torch.cuda.synchronize()


def measure(ctx, env, actions, wl, workload_name, steps, warmup, repeats=3, sampler=None):
    """Device-timed throughput of `steps` lockstep iterations (CUDA events, barrier + synchronize on both sides, max over
    ranks), repeated `repeats` times: `value` follows the MEDIAN repetition, the spread is reported.  The dominant
    kernel is timed live inside the timed regions with its own CUDA events (see `roofline.timing`)."""
    torch, dist, world = ctx.torch, ctx.dist, ctx.world
    N, n_sets = wl["envs"], actions.shape[0]
    for k in range(warmup):
        env.step(actions[k % n_sets])
    _barrier(ctx)
    # per-kernel CUDA events INSIDE the timed region: on every 4th (8th) iteration env.step() issues its two kernels
    # as two calls (the very kernels gte_step_obs launches) with events on the launching stream around each; the
    # event records cost ~3 us per iteration they bracket, hence not every iteration
    launches_per_step = env.launches_per_step
    fused = wl["windows"] is not None and launches_per_step == 1      # transition + gather in ONE launch (small batches)
    env._kernel_events = [] if (wl["windows"] is not None and not fused and not ctx.args.cuda_graph) else None
    env._kernel_events_every = 1 if steps < 64 else (4 if N >= 2 ** 20 else 8)
    times = []
    t_host0 = time.time()
    rep = 0
    while rep < repeats:
        rep += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _barrier(ctx)
        e0.record()
        for k in range(steps):
            env.step(actions[k % n_sets])
        env.wait_metric_allreduce()          # the timed region ends when the last all-reduce has landed
        e1.record()
        _barrier(ctx)
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=ctx.dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t.item()))
        if rep == 1 and ctx.args.min_timed_seconds > 0:
            # a short K (the driver passes --steps 20: 22 ms) would leave the clock sampler (20 ms period) with one or
            # two samples and the median with three points: repeat the region until ~min_timed_seconds are covered
            # (same count on every rank: the decision uses the max-over-ranks time)
            repeats = max(repeats, min(25, int(ctx.args.min_timed_seconds * 1e3 / max(times[0], 1e-3)) + 1))
    clocks = sampler.stop(t_host0, time.time()) if (sampler is not None and ctx.rank == 0) else None
    env.check_errors()
    ms = statistics.median(times)
    value = world * N * steps / (ms * 1e-3)
    spread = {"ms_per_step_min": min(times) / steps, "ms_per_step_max": max(times) / steps,
              "rel": (max(times) - min(times)) / ms}

    # ---- roofline of the dominant kernel (window gather; the whole launch when there is only one) ----
    a_step, a_obs = algorithmic_bytes(wl["windows"])
    env_events = env._kernel_events
    if env_events:
        step_ms = [ev[0].elapsed_time(ev[1]) for ev in env_events]
        obs_ms = [ev[1].elapsed_time(ev[2]) for ev in env_events]
    else:                                    # one launch per iteration = the whole step
        step_ms, obs_ms = [ms / steps], [ms / steps]
    every = env._kernel_events_every
    env._kernel_events = None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    obs_ms_avg = sum(obs_ms) / max(len(obs_ms), 1)
    step_ms_avg = sum(step_ms) / max(len(step_ms), 1)
    split = wl["windows"] is not None and bool(env_events)               # per-kernel events were recorded
    dom_bytes = a_obs if split else a_step + a_obs                       # else one timing for the whole iteration
    achieved = (dom_bytes * N) / (obs_ms_avg * 1e-3) / 1e9
    traffic, traffic_step = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj.get(f"{workload_name}:{env.obs_variant}{':fused' if fused else ''}:{N}")
        traffic_step = tj.get(f"{workload_name}:step:{N}")
    whole = (a_step + a_obs) * N * steps / (ms * 1e-3) / 1e9
    # C4: the 1.28 GB feature table is not L2-resident, so the static window read is HBM traffic too
    # (SURVEY.md §8d "5 218 B" accounting) — reported beside the conservative figure
    a_static = (wl["windows"] or 1) * 8 * 4 if wl["n_datasets"] > 1 else 0
    kname = {"tma": "obs_tma_coop_kernel", "vec": "obs_vec_kernel", "generic": "obs_generic_kernel"}[env.obs_variant]
    if wl["windows"] is None:
        kname = "step_kernel (windows=None: writes the one-row observation itself)"
    elif fused:
        kname += "<fused> (ONE launch per iteration: every CTA advances its envs, then gathers their windows; timed as a whole)"
    elif not split:
        kname = "step_kernel + " + kname + " (graph replay: timed together)"
    timing = ("CUDA events on the launching stream around the launches of every %s iteration of the timed regions"
              % {1: "single", 4: "4th", 8: "8th"}[every]) if split else "the timed regions themselves (one launch per iteration)"
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_env": dom_bytes, "kernel_ms": obs_ms_avg,
                "timing": timing, "step_kernel_ms": step_ms_avg if split else None,
                "step_kernel": None if not split else {
                    "algorithmic_bytes_per_env": a_step, "achieved": a_step * N / (step_ms_avg * 1e-3) / 1e9,
                    "frac": a_step * N / (step_ms_avg * 1e-3) / 1e9 / peak, "traffic": traffic_step},
                "whole_step": {"algorithmic_bytes_per_env_step": a_step + a_obs, "achieved": whole,
                               "frac": whole / peak, "frac_of_nominal_8TBs": whole / 8000.0,
                               "with_static_window_read": None if not a_static else {
                                   "algorithmic_bytes_per_env_step": a_step + a_obs + a_static,
                                   "achieved": whole * (a_step + a_obs + a_static) / (a_step + a_obs),
                                   "frac": whole * (a_step + a_obs + a_static) / (a_step + a_obs) / peak}}}
    return {"value": value, "ms_per_step": ms / steps, "spread": spread, "clocks": clocks, "roofline": roofline,
            "repeats": {"n": len(times), "steps_each": steps, "statistic": "median"},
            "launches_per_step": launches_per_step, "obs_variant": env.obs_variant, "chunks": env.chunks,
            "l2_policy": "inputs larger than L2 (obs %.0f MB + state/ring %.0f MB per step vs 126 MB L2)"
                         % (env._obs.numel() * 4 / 1e6, (N * 44 + env._dyn_ring.numel()) / 1e6)}


def latency_probe(ctx, env, actions, wl, m):
    """Launch / latency floor at small N (BASELINE config 2): the same iterations enqueued by ONE host call
    (env.rollout -> gte_rollout: a C loop of kernel launches, no Python between iterations), and the cost of an
    empty stream round trip on this box."""
    torch = ctx.torch
    K, N, n_sets = 2000, wl["envs"], actions.shape[0]
    acts_k = actions[torch.arange(K, device=ctx.dev) % n_sets]
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    us_open = float("inf")
    for _ in range(4):                   # the GPU idled while the clocks were read: let the SM clock ramp up again
        r0.record()
        env.rollout(acts_k, keep_obs=False)
        r1.record()
        torch.cuda.synchronize()
        us_open = min(us_open, 1e3 * r0.elapsed_time(r1) / K)
    # the floor under any synchronous host call: one step() + stream synchronize (launch -> completion -> host wake-up)
    stream = torch.cuda.current_stream()
    for _ in range(50):
        env.step(actions[0]); stream.synchronize()
    t0 = time.perf_counter()
    for k in range(500):
        env.step(actions[k % n_sets]); stream.synchronize()
    sync_us = 1e6 * (time.perf_counter() - t0) / 500
    return {"step_call_us": 1e3 * m["ms_per_step"],
            "rollout_us_per_iteration": us_open, "rollout_env_steps_per_s": N / (us_open * 1e-6),
            "step_plus_stream_sync_us": sync_us,
            "note": "step_call = env.step() per iteration from Python, back to back (one ctypes call per iteration); rollout = "
                    "%d iterations enqueued by one gte_rollout call (open-loop actions; only the last observation is "
                    "gathered, so with windows this is the step kernel's rate), best of 4; step_plus_stream_sync = one "
                    "iteration followed by a stream synchronize (launch -> completion -> host wake-up): the floor under "
                    "any synchronous host-policy step on this box" % K}


def e2e_legs(ctx, env, actions, wl):
    """e2e: the public API with HOST numpy actions in and HOST numpy reward / terminated / truncated out, wall clock
    around the step() calls.  "hybrid" (headline): observations stay device-resident for the policy's forward pass;
    "numpy": the full observation windows also cross PCIe (2.5 KB per env-step: PCIe-bound)."""
    torch, dist, world, dev, args = ctx.torch, ctx.dist, ctx.world, ctx.dev, ctx.args
    from gym_trading_env_b200._cabi import host_result_layout, host_result_sparse_bytes

    def io_mode_used():
        """1 = copy engines, 2 = mapped, 3 = server; a relayed env always steps through the copy engines (begin / end)"""
        return 1 if env._relay is not None else env._io_mode_used.value

    def sparse_wire():
        if env._relay is not None:
            return bool(env.sparse_flags)
        return (env._host or {}).get("sparse") is not None and env._io_mode_used.value == 1

    def d2h_bytes(mode):
        """bytes of the result block that cross PCIe per step (the sparse prefix when the env's flag wire is sparse)"""
        sparse = mode == "hybrid" and sparse_wire()
        b = host_result_sparse_bytes(N) if sparse else host_result_layout(N)[3]
        if mode == "hybrid" and env.reward_wire == "f32" and io_mode_used() == 1:
            b -= 4 * N                                       # float32 instead of fp64 rewards on the wire
        return b
    N, n_sets = wl["envs"], actions.shape[0]
    wire = {}
    for dt in (torch.int8, torch.int64):                 # action sets staged in pinned host memory, both wire widths
        t = torch.empty(actions.shape, dtype=dt, pin_memory=True)
        t.copy_(actions)
        wire[dt] = (t, t.numpy())

    def time_e2e(mode, host_io, dt, n_it):
        env.output, env.host_io, env._host = mode, host_io, None
        acts_h = wire[dt][1]
        rows = [acts_h[i] for i in range(n_sets)]        # the caller's pinned action arrays, one per action set
        for k in range(3):
            env.step(rows[k % n_sets])                   # allocates + warms the pinned buffers
        _barrier(ctx)
        t0 = time.perf_counter()
        for k in range(n_it):
            env.step(rows[k % n_sets])
        env.wait_metric_allreduce()
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        tt = torch.tensor([dt_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ab = acts_h.dtype.itemsize
        return {"value": world * N * n_it / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": N * ab,
                "d2h_bytes_per_step": d2h_bytes(mode) + (env._obs.numel() * 4 if mode == "numpy" else 0),
                "flag_wire": "sparse (list of ended envs)" if sparse_wire() else "dense bytes",
                "reward_wire": env.reward_wire,
                "steps": n_it, "us_per_step": 1e6 * float(tt.item()) / n_it,
                "timing": "host wall clock around the step() calls, max over ranks", "mode": mode,
                "host_io": {1: "copy", 2: "mapped", 3: "server"}.get(io_mode_used(), "?"),
                "action_dtype": str(acts_h.dtype)}

    out = {}
    n_it = min(max(args.steps, 60), 200) if N >= 2 ** 20 else 2000
    relay = None
    if world > 1 and N >= 2 ** 20 and not args.no_relay:
        # multi-GPU: first the plain wire for comparison (every rank's result block over its own PCIe link), then the
        # result relay — part of the slow links' reward bytes reach the host through a peer GPU (lossless, collective)
        out["e2e_no_relay"] = time_e2e("hybrid", "auto", torch.int8, n_it)
        out["e2e_no_relay"]["note"] = "the headline wire WITHOUT enable_result_relay(): every rank's results over its own PCIe link"
        env.output, env._host = "hybrid", None
        try:
            relay = env.enable_result_relay()
        except Exception as exc:  # noqa: BLE001 — the plain wire's number above still stands
            relay = {"errors": [repr(exc)], "plan": {}, "measured_d2h_gbs_all_ranks_copying": []}
    # headline: the documented default wire format of a host policy — int8 actions (Discrete(P) fits, widened by the
    # step kernel: lossless), fp64 rewards + terminated + truncated + error flag back in ONE block
    if relay is not None and not relay["plan"]:
        out["e2e"] = out.pop("e2e_no_relay")                 # links within 20 % of each other: nothing to balance, one measurement
    else:
        out["e2e"] = time_e2e("hybrid", "auto", torch.int8, n_it)
    out["e2e"]["note"] = ("VectorEnv(output='hybrid').step(pinned int8 numpy actions) -> numpy f64 reward / bool terminated / bool "
                          "truncated every step through ONE gte_step_host call (one copy per direction, or mapped host memory at "
                          "small N); the observation tensor stays in HBM for the policy's forward pass")
    if relay is not None:
        out["e2e"]["relay"] = relay
        out["e2e"]["note"] += ("; env.enable_result_relay() called once: reward bytes balanced over the GPUs' PCIe links through peer "
                               "GPUs (NVLink + the peer's copy engine) when the links' measured rates differ by more than 20 %, see `relay`")
    out["e2e_gymnasium_dtypes"] = time_e2e("hybrid", "auto", torch.int64, n_it)
    out["e2e_gymnasium_dtypes"]["note"] = "same call with gymnasium's own dtypes on the wire (int64 actions in, f64 reward + bool flags out)"
    if N >= 2 ** 20:
        # the same wire, two iterations in flight (step_async / step_wait): iteration k's result copy runs under iteration
        # k+1 — what a caller that does not need step k's rewards to choose step k+1's actions can do
        env.output, env.host_io, env._host = "hybrid", "auto", None
        rows = [wire[torch.int8][1][i] for i in range(n_sets)]
        def pipelined(n):
            env.step_async(rows[0])
            for k in range(n):
                if k + 1 < n:
                    env.step_async(rows[(k + 1) % n_sets])
                env.step_wait()
        pipelined(4)
        _barrier(ctx)
        t0 = time.perf_counter()
        pipelined(n_it)
        env.wait_metric_allreduce()
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        out["e2e_pipelined"] = {"value": world * N * n_it / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": N,
                                "d2h_bytes_per_step": host_result_sparse_bytes(N) if env.sparse_flags else host_result_layout(N)[3],
                                "steps": n_it,
                                "us_per_step": 1e6 * float(tt.item()) / n_it, "mode": "hybrid", "host_io": "copy",
                                "action_dtype": "int8", "timing": "host wall clock around the calls, max over ranks",
                                "note": "step_async(a[k+1]) before step_wait() of iteration k: two iterations in flight, every "
                                        "step's actions still come from pinned host memory and every step's reward / flags still "
                                        "land on the host inside the timed region"}
    if relay is not None:
        env.close_extras()                                   # collective: leaves the relay, frees the wire sets
    if N >= 2 ** 20:
        env.reward_wire = "f32"
        out["e2e_f32_reward"] = time_e2e("hybrid", "auto", torch.int8, n_it)
        out["e2e_f32_reward"]["note"] = ("OPT-IN, LOSSY: reward_wire='f32' — the rewards cross PCIe as numpy's float32 cast of the fp64 "
                                         "rewards (what stable-baselines3 keeps anyway); flags exact; the fp64 rewards stay on the device")
        env.reward_wire = "f64"
    if world == 1:
        other = "mapped" if out["e2e"]["host_io"] == "copy" else "copy"
        out["e2e_other_host_io"] = time_e2e("hybrid", other, torch.int8, min(n_it, 50) if N >= 2 ** 20 else n_it)
        out["e2e_other_host_io"]["note"] = ("the OTHER host-IO mechanism forced, for comparison (mapped = the step kernel reads the "
                                            "actions from, and writes its results into, pinned host memory; copy = copy engines)")
        if wl["windows"] is None and N <= 32768:
            out["e2e_server"] = time_e2e("hybrid", "server", torch.int8, n_it)
            out["e2e_server"]["note"] = ("host_io='server': a RESIDENT kernel answers every step through mapped host memory — no kernel "
                                         "launch, driver call or interrupt per step (opt-in: it occupies its SMs while it waits)")
            env._lib.gte_serve_stop()
        out["e2e_full_obs_to_host"] = time_e2e("numpy", "auto", torch.int8, args.e2e_steps if N >= 2 ** 20 else 50)
        out["e2e_full_obs_to_host"]["note"] = ("VectorEnv(output='numpy'): the full observation batch is also delivered into pinned host "
                                               "memory every step (same single C call); bounded by PCIe (~52 GB/s) for windowed batches, "
                                               "reported for transparency")
    env.output, env.host_io, env._host = "torch", "auto", None
    return out


def config_leg(ctx, name):
    """A short leg of another BASELINE config inside the default run: value, time per iteration, whole-step fraction of
    the HBM roofline, the dominant kernel's fraction, and for the small configs the latency floor and the host path."""
    wl = dict(WORKLOADS[name])
    steps = {"c2": 3000, "c3": 1500, "c4": 60}[name]
    env, actions = build_env(ctx, wl, name)
    m = measure(ctx, env, actions, wl, name, steps, 5, repeats=3)
    r = m["roofline"]
    leg = {"workload": wl["label"], "value": m["value"], "unit": UNIT, "ms_per_step": m["ms_per_step"], "steps": steps,
           "repeats": m["repeats"], "spread": m["spread"], "launches_per_step": m["launches_per_step"],
           "whole_step_frac": r["whole_step"]["frac"], "whole_step_frac_with_static_window_read":
               (r["whole_step"]["with_static_window_read"] or {}).get("frac"),
           "kernel": r["kernel"], "kernel_ms": r["kernel_ms"], "kernel_frac": r["frac"], "step_kernel_ms": r["step_kernel_ms"],
           "traffic": r["traffic"], "algorithmic_bytes_per_env_step": r["whole_step"]["algorithmic_bytes_per_env_step"]}
    if wl["envs"] < 2 ** 20:
        leg["latency"] = latency_probe(ctx, env, actions, wl, m)
        e = e2e_legs(ctx, env, actions, wl)
        leg["e2e"] = {k: {kk: v[kk] for kk in ("value", "us_per_step", "host_io", "action_dtype", "h2d_bytes_per_step", "d2h_bytes_per_step")}
                      for k, v in e.items() if v}
    env.close()
    return leg


if __name__ == "__main__":
    main()
