"""Result relay for a one-process-per-GPU job with a HOST policy (``output="hybrid"``): ranks whose GPU has a slow
device-to-host path under load ship the tail of their reward block over NVLink to a peer GPU with a fast path, whose
copy engine writes it into the sender's result block (host memory both processes map).  Lossless: the same bytes arrive,
by another road.

Why (measured, `profiles/r02_host_io_probe_n8.json`): on this pool's 8 x B200 nodes, with all eight ranks copying results
to the host at once, GPUs 0-3 get 12.9 GB/s each and GPUs 4-7 get 22.6 GB/s each (56 GB/s alone).  The result copy of
the slow ranks — 17 MB per iteration at 2^21 envs — is then the longest thing in the iteration (1.32 ms against 1.11 ms
of kernels), and every rank waits for them.  Balancing the bytes over the links in proportion to their measured
bandwidth shortens the longest copy.

The data path is `libgte_b200.so`'s (``gte_relay_*``, include/gte_b200.h): copy engines plus one stream memory
operation, no kernel, no host thread, no NCCL call per step.  This module only plans (who sends how much to whom, from a
bandwidth measurement every rank takes at the same time) and exchanges the handles once (``torch.distributed`` object
collectives: any backend).

``TradingVectorEnv.enable_result_relay()`` is a COLLECTIVE call, and afterwards so is every ``step()`` /
``step_async()``: a peer enqueues its part of iteration k when it enqueues its own iteration k.
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
import time

import numpy as np

from . import _cabi

QUANTUM = 1024                      # envs: relayed tails are a multiple of this
MIN_RATIO = 1.2                     # pairs whose measured bandwidths differ by less are left alone
MAX_FRACTION = 0.5                  # a sender never ships more than half of its rewards (the peer buffers are sized for it)


def plan_relay(bandwidth, n_envs, min_ratio=MIN_RATIO, quantum=QUANTUM, max_fraction=MAX_FRACTION):
    """Who ships how many of its per-env result words to whom.

    ``bandwidth[r]``: device-to-host GB/s rank r gets while every rank copies.  Ranks are paired slowest-with-fastest;
    a pair (s, f) moves ``x`` of s's ``n_envs`` words to f's link so that both finish together:
    ``(n - x) / bw_s = (n + x) / bw_f``  ->  ``x = n (bw_f - bw_s) / (bw_f + bw_s)``.
    Returns ``{sender: (peer, x)}`` with x a multiple of ``quantum`` and > 0 only; pure function (same answer on every
    rank from the same all-gathered list)."""
    order = sorted(range(len(bandwidth)), key=lambda r: (bandwidth[r], r))
    plan = {}
    for i in range(len(order) // 2):
        s, f = order[i], order[len(order) - 1 - i]
        bs, bf = float(bandwidth[s]), float(bandwidth[f])
        if bs <= 0 or bf < min_ratio * bs:
            continue
        x = n_envs * (bf - bs) / (bf + bs)
        x = int(min(x, max_fraction * n_envs)) // quantum * quantum
        if x > 0:
            plan[s] = (f, x)
    return plan


def parse_forced_plan(text, n_envs, quantum=QUANTUM):
    """``GTE_RELAY_FORCE="0>1:0.3,2>3:0.25"`` -> {0: (1, x), 2: (3, x)} (tests / experiments)."""
    plan = {}
    for item in filter(None, (t.strip() for t in text.split(","))):
        pair, _, frac = item.partition(":")
        s, _, f = pair.partition(">")
        x = int(float(frac) * n_envs) // quantum * quantum
        if x > 0:
            plan[int(s)] = (int(f), x)
    return plan


class SharedPinnedBlock:
    """A host buffer both processes map, page-locked for this process's GPU (``gte_host_register``): the owner's result
    block, which a peer process maps as well so that ITS copy engine can write into it.  Backed by an anonymous
    ``memfd`` (RAM, not limited by the size of a container's /dev/shm); the peer opens it through the owner's
    ``/proc/<pid>/fd/<n>`` entry, which stays valid until the owner calls :meth:`drop_fd`."""

    def __init__(self, lib, nbytes, path=None):
        self.lib, self.nbytes, self.registered, self.fd = lib, int(nbytes), False, -1
        create = path is None
        fd = os.memfd_create("gte_relay") if create else os.open(path, os.O_RDWR)
        try:
            if create:
                os.ftruncate(fd, self.nbytes)
            self.map = mmap.mmap(fd, self.nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        except BaseException:
            os.close(fd)
            raise
        if create:
            self.fd, self.path = fd, f"/proc/{os.getpid()}/fd/{fd}"
        else:
            os.close(fd)
        self.array = np.frombuffer(self.map, dtype=np.uint8)
        if create:
            self.array[:] = 0                                    # touches every page
        self.ptr = self.array.ctypes.data
        _cabi.check(lib.gte_host_register(C.c_void_p(self.ptr), self.nbytes), "gte_host_register")
        self.registered = True

    def drop_fd(self):
        if self.fd >= 0:
            os.close(self.fd)
            self.fd = -1

    def close(self):
        self.drop_fd()
        if self.registered:
            self.lib.gte_host_unregister(C.c_void_p(self.ptr))
            self.registered = False


def measure_d2h_together(torch, dist, device, nbytes_of_rank, reps=6, group=None):
    """GB/s every rank gets for device-to-host copies of ITS planned size while all ranks copy at the same time (CUDA
    events around `reps` back-to-back copies behind a barrier).  Returns the all-gathered list."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = int(nbytes_of_rank[rank])
    src = torch.empty(max(n, 8), dtype=torch.uint8, device=device)
    dst = torch.empty(max(n, 8), dtype=torch.uint8, pin_memory=True)
    s = torch.cuda.Stream(device=device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s):
        dst.copy_(src, non_blocking=True)
        s.synchronize()
        dist.barrier(group=group)
        a.record(s)
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        b.record(s)
        s.synchronize()
    gbs = n * reps / (a.elapsed_time(b) * 1e-3) / 1e9
    out = [None] * world
    dist.all_gather_object(out, float(gbs), group=group)
    return [float(v) for v in out]


class ResultRelay:
    """Per-env state of the relay: the plan, this rank's role(s), the shared result blocks and the opened peer buffers."""

    def __init__(self, env, group=None, plan=None, calibrate_rounds=2, verbose=False):
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("enable_result_relay() needs an initialised torch.distributed job (one process per GPU)")
        self.env, self.lib, self.group = env, env._lib, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        N = self.N = env.num_envs
        counts = [None] * self.world
        dist.all_gather_object(counts, int(N), group=group)
        if len(set(counts)) != 1:
            raise ValueError(f"the result relay needs the same num_envs on every rank, got {counts}")
        supported = [None] * self.world
        dist.all_gather_object(supported, int(self.lib.gte_relay_supported()), group=group)
        self.block_bytes = _cabi.host_result_layout(N)[3]
        self.sparse_bytes = _cabi.host_result_sparse_bytes(N) if env.sparse_flags else self.block_bytes
        self.measured = []
        forced = os.environ.get("GTE_RELAY_FORCE")
        if plan is None and forced is not None:
            plan = parse_forced_plan(forced, N)
        if plan is None:
            # round 1: everybody copies a full block — decides WHICH pairs relay; round 2 copies the planned split and
            # moves each pair's split half-way towards what the rates seen then ask for.  (Measured on the 8 x B200 node:
            # the even load overstates a slow link's rate — its copy runs alone once the fast links have finished — 12.2
            # GB/s against 9.9 GB/s when the fast links stay busy for as long.)  Back-to-back copies are not how the job
            # will load the links, though, and more rounds drift towards the split that suits two iterations in flight (every
            # link busy all the time: 34-37 % shipped, 1.26 ms pipelined / 1.32-1.35 ms one step at a time) and away from
            # the one that suits one step at a time (27-30 %: 1.26 ms / 1.32 ms pipelined) — measured, profiles/r02_tuning.md §6.
            plan, sizes = {}, [self.sparse_bytes] * self.world
            for rnd in range(max(1, int(calibrate_rounds))):
                bw = measure_d2h_together(torch, dist, env.device, sizes, group=group)
                self.measured.append(bw)
                if rnd == 0:
                    plan = plan_relay(bw, N) if all(supported) else {}
                else:
                    for s, (f, x) in list(plan.items()):
                        want = N * (bw[f] - bw[s]) / (bw[f] + bw[s])
                        x = int(max(0.0, min((x + want) / 2, 0.5 * N))) // QUANTUM * QUANTUM
                        if x > 0:
                            plan[s] = (f, x)
                        else:
                            del plan[s]
                if not plan:
                    break
                sizes = [self.sparse_bytes] * self.world
                for s, (f, x) in plan.items():
                    sizes[s] -= 8 * x
                    sizes[f] += 8 * x
        bad = [s for s, (f, x) in plan.items() if not (0 <= s < self.world and 0 <= f < self.world and s != f and 0 < x < N)]
        if bad:
            raise ValueError(f"invalid relay plan {plan}")
        if plan and not all(supported):
            raise RuntimeError("the result relay needs stream memory operations (cuStreamWaitValue32) on every rank")
        self.plan = dict(plan)
        self.count = 0                                            # iterations enqueued since the relay was enabled
        self.peer = self.plan.get(self.rank)                      # (rank I send to, how many envs) or None
        self.senders = sorted(s for s, (f, _) in self.plan.items() if f == self.rank)      # ranks I serve
        if len(self.senders) > 8:
            raise ValueError("a rank serves at most 8 senders")
        self.blocks, self.peer_bufs, self.serve = [], [None, None], {}
        with torch.cuda.device(env.device):
            self._setup(torch, dist)
        if verbose and self.rank == 0:
            print(f"[gte relay] measured GB/s {self.measured}; plan {self.plan}", flush=True)

    # ------------------------------------------------------------------ one-time exchange of names and handles
    def _everyone_ok(self, dist, err):
        """Exchange this rank's error (or None); True when nobody failed.  A failure anywhere (no memfd, IPC refused in
        this container ...) switches the relay off on EVERY rank instead of leaving ranks waiting for one another."""
        errs = [None] * self.world
        dist.all_gather_object(errs, None if err is None else f"rank {self.rank}: {err!r}", group=self.group)
        self.errors = [e for e in errs if e is not None]
        return not self.errors

    def _setup(self, torch, dist):
        lib, N, rank = self.lib, self.N, self.rank
        self.errors, self._own, self.relay_words, self.expect, self.seq_base = [], {}, None, [0, 0], 0
        if not self.plan:
            return
        mine, err = {"paths": None, "handles": {}}, None
        try:
            # a sender's two result blocks (one per wire set of step_async / step_wait) live in shared memory, 64 more
            # bytes at the end hold the relay's completion word
            if self.peer is not None:
                self.blocks = [SharedPinnedBlock(lib, self.block_bytes + 64) for _ in range(2)]
                mine["paths"] = [b.path for b in self.blocks]
            # a peer allocates, per sender and wire set, the device buffer the sender's copy engine writes into
            for s in self.senders:
                self._own[s] = []
                for k in range(2):
                    base, handle = C.c_void_p(), (C.c_ubyte * 64)()
                    _cabi.check(lib.gte_relay_alloc(8 * self.plan[s][1], C.byref(base), handle), "gte_relay_alloc")
                    self._own[s].append(base.value)
                    mine["handles"][(s, k)] = bytes(handle)
        except Exception as e:  # noqa: BLE001
            err = e
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        if not self._everyone_ok(dist, err):
            return self._switch_off()
        try:
            if self.peer is not None:                              # open the peer's two buffers for me
                f = self.peer[0]
                for k in range(2):
                    h = (C.c_ubyte * 64).from_buffer_copy(everyone[f]["handles"][(rank, k)])
                    base = C.c_void_p()
                    _cabi.check(lib.gte_relay_open(h, C.byref(base)), "gte_relay_open")
                    self.peer_bufs[k] = base.value
            for lane, s in enumerate(self.senders):                # map each sender's result blocks
                blks = [SharedPinnedBlock(lib, self.block_bytes + 64, path=p) for p in everyone[s]["paths"]]
                self.serve[s] = {"lane": lane, "blocks": blks, "own": self._own[s],
                                 "seq": [b.ptr + self.block_bytes for b in blks]}
        except Exception as e:  # noqa: BLE001
            err = e
        ok = self._everyone_ok(dist, err)                          # (also: every mapping exists, the descriptors can go)
        for b in self.blocks:
            b.drop_fd()
        if not ok:
            return self._switch_off()
        if self.peer is not None:
            self.relay_words = [b.array[self.block_bytes:self.block_bytes + 4].view(np.uint32) for b in self.blocks]
        # self-test: ONE round trip per pair (sequence number 1, wire set 0) before the job depends on the road — IPC mapping,
        # peer access, the stream wait and the peer's write into the shared block all have to work; if any pair fails, every
        # rank goes back to the plain path instead of waiting for a word that never comes in the middle of a run
        err = None
        try:
            for s in self.senders:
                e = self.serve[s]
                _cabi.check(lib.gte_relay_serve(e["lane"], C.c_void_p(e["own"][0]), 16, 1,
                                                C.c_void_p(e["blocks"][0].ptr + 8 * (N - self.plan[s][1])), C.c_void_p(e["seq"][0])),
                            "gte_relay_serve")
            if self.peer is not None:
                _cabi.check(lib.gte_relay_push(C.c_void_p(self.peer_bufs[0]), C.c_void_p(self.env._result_block.data_ptr()), 16, 1,
                                               None, None), "gte_relay_push")
                t0 = time.monotonic()
                while self.relay_words[0][0] != 1:
                    if time.monotonic() - t0 > 5.0:
                        raise RuntimeError(f"no answer from rank {self.peer[0]} within 5 s")
        except Exception as e:  # noqa: BLE001
            err = e
        self.seq_base = 1                                          # the job's sequence numbers start behind the self-test's
        if not self._everyone_ok(dist, err):
            for e in self.serve.values():                          # a serve stream may still be waiting: let it go
                lib.gte_relay_unblock(C.c_void_p(e["own"][0]), 1)
            torch.cuda.synchronize()
            return self._switch_off()

    def _switch_off(self):
        self._release()
        self.plan, self.peer, self.senders = {}, None, []

    # ------------------------------------------------------------------ used by TradingVectorEnv
    def host_block(self, k):
        """torch uint8 tensor over wire set k's result block in shared memory (sender ranks), else None."""
        import torch
        if self.peer is None:
            return None
        return torch.from_numpy(self.blocks[k].array[:self.block_bytes])

    @property
    def own_count(self):
        """How many leading rewards this rank's own copy delivers (GteHostIO.reward_host_count); 0 = all."""
        return self.N - self.peer[1] if self.peer is not None else 0

    def before_begin(self):
        """Called first thing in step_async: enqueue, for every sender this rank serves, the wait for its data of this
        iteration and the two copies into its result block."""
        k, seq = self.count & 1, (self.count + 1 + self.seq_base) & 0xffffffff
        for s in self.senders:
            e, x = self.serve[s], self.plan[s][1]
            rc = self.lib.gte_relay_serve(e["lane"], C.c_void_p(e["own"][k]), 8 * x, seq,
                                          C.c_void_p(e["blocks"][k].ptr + 8 * (self.N - x)), C.c_void_p(e["seq"][k]))
            if rc:
                _cabi.check(rc, "gte_relay_serve")

    def after_begin(self, k_set, dev_block_ptr, step_done_event):
        """Called right behind gte_step_host_begin: ship this iteration's reward tail to the peer."""
        seq = (self.count + 1 + self.seq_base) & 0xffffffff
        if self.peer is not None:
            x = self.peer[1]
            rc = self.lib.gte_relay_push(C.c_void_p(self.peer_bufs[k_set]), C.c_void_p(dev_block_ptr + 8 * (self.N - x)), 8 * x, seq,
                                         C.c_void_p(step_done_event), None)
            if rc:
                _cabi.check(rc, "gte_relay_push")
            self.expect[k_set] = seq
        self.count += 1

    def wait(self, k_set):
        """Called behind gte_step_host_end on a sender: the peer's copy engine has written the tail when the relay word
        of the block holds this iteration's number."""
        if self.peer is None:
            return
        w, seq = self.relay_words[k_set], self.expect[k_set]
        if w[0] == seq:
            return
        limit = float(os.environ.get("GTE_HOST_SPIN_TIMEOUT_S", "20"))
        t0, spins = time.monotonic(), 0
        while w[0] != seq:
            spins += 1
            if (spins & 0xfff) == 0 and time.monotonic() - t0 > limit:
                raise RuntimeError(f"result relay: rank {self.rank} waited {limit:.0f} s for rank {self.peer[0]} to deliver "
                                   f"iteration {seq} (is every rank calling step() in lockstep?)")

    def describe(self):
        return {"errors": self.errors, "plan": {str(s): {"via": f, "envs": x, "fraction": round(x / self.N, 4)} for s, (f, x) in sorted(self.plan.items())},
                "measured_d2h_gbs_all_ranks_copying": [[round(v, 2) for v in bw] for bw in self.measured]}

    def _release(self):
        for k in range(2):
            if self.peer_bufs[k] is not None:
                self.lib.gte_relay_release(C.c_void_p(self.peer_bufs[k]), 1)
                self.peer_bufs[k] = None
        for e in self.serve.values():
            for b in e["blocks"]:
                b.close()
        self.serve = {}
        try:                                       # senders unmap before the owners free (a collective, like the set-up)
            import torch.distributed as dist
            if dist.is_initialized() and self.world > 1:
                dist.barrier(group=self.group)
        except Exception:  # noqa: BLE001
            pass
        for bufs in self._own.values():
            for p in bufs:
                self.lib.gte_relay_release(C.c_void_p(p), 0)
        self._own = {}
        for b in self.blocks:
            b.close()
        self.blocks = []

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self.plan:
            self._release()
        self.plan, self.peer, self.senders = {}, None, []
