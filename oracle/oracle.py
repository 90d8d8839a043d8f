"""ctypes wrapper around the C restatement `oracle/gte_oracle.c` (TEST INFRASTRUCTURE ONLY).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import this
module; the product package never does (it fails loudly when the CUDA library is missing).

`OracleVecEnv` exposes the same lockstep vector-step contract as the CUDA path
(SURVEY.md §8(a)): `reset()` then `step(actions)` with in-place auto-reset, episode starts either
from an injected `plan[N,E,3]` (record-and-replay of the reference's RNG draws) or from the same
Philox4x32-10 stream the device uses.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgte_oracle.so")
N_METRICS = 8
METRIC_NAMES = ["episodes", "terminated", "truncated", "sum_portfolio_return", "sum_market_return",
                "sum_episode_length", "sum_reward", "reserved"]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gte_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgte_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _OrcEnv(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("n_positions", C.c_int32), ("windows", C.c_int32),
        ("n_static", C.c_int32), ("n_dyn", C.c_int32), ("max_episode_duration", C.c_int32),
        ("n_datasets", C.c_int32), ("initial_position_idx", C.c_int32),
        ("episodes_between_switch", C.c_int32), ("dyn_mode", C.c_int32),
        ("plan_episodes", C.c_int32), ("multi_dataset", C.c_int32),
        ("reward_kind", C.c_int32), ("n_limit_positions", C.c_int32),
        ("t_stride", C.c_int64), ("env_id_offset", C.c_int64), ("seed", C.c_uint64),
        ("fee", C.c_double), ("rate", C.c_double), ("v0", C.c_double), ("done_ratio", C.c_double),
        ("reward_scale", C.c_double), ("reward_lo", C.c_double), ("reward_hi", C.c_double),
        ("positions", C.c_void_p), ("features", C.c_void_p), ("price", C.c_void_p), ("lengths", C.c_void_p),
        ("asset", C.c_void_p), ("fiat", C.c_void_p), ("interest_asset", C.c_void_p),
        ("interest_fiat", C.c_void_p), ("prev_val", C.c_void_p),
        ("pos_idx", C.c_void_p), ("step", C.c_void_p), ("ep_start", C.c_void_p), ("dataset_idx", C.c_void_p),
        ("plan_cursor", C.c_void_p), ("ds_used", C.c_void_p), ("ds_episodes", C.c_void_p),
        ("dyn_cols", C.c_void_p), ("touched_lo", C.c_void_p), ("touched_hi", C.c_void_p),
        ("plan", C.c_void_p),
        ("high", C.c_void_p), ("low", C.c_void_p), ("limit_price", C.c_void_p), ("limit_seq", C.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        assert _lib.orc_struct_size() == C.sizeof(_OrcEnv), "OrcEnv layout mismatch"
        d = C.c_double
        _lib.orc_valorisation.restype = d
        _lib.orc_valorisation.argtypes = [d] * 5
        _lib.orc_real_position.restype = d
        _lib.orc_real_position.argtypes = [d] * 5
        _lib.orc_position.restype = d
        _lib.orc_position.argtypes = [d] * 5
        _lib.orc_trade_to_position.argtypes = [C.c_void_p, d, d, d]
        _lib.orc_update_interest.argtypes = [C.c_void_p, d]
        _lib.orc_target_portfolio.argtypes = [C.c_void_p, d, d, d]
        _lib.orc_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        _lib.orc_init_datasets.argtypes = [C.c_void_p, C.c_uint64]
        _lib.orc_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        _lib.orc_step.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64] + [C.c_void_p] * 11
        _lib.orc_step_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64] + [C.c_void_p] * 11
        _lib.orc_rollout_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64] + [C.c_void_p] * 5
        _lib.orc_rollout_threads.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64] + [C.c_void_p] * 5 + [C.c_int, C.c_int]
        _lib.orc_rollout_threads.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


# ---- scalar helpers (property tests) -------------------------------------------------------------

def trade_to_position(state, position, price, fee):
    s = np.array(state, dtype=np.float64)
    lib().orc_trade_to_position(_p(s), float(position), float(price), float(fee))
    return s


def update_interest(state, rate):
    s = np.array(state, dtype=np.float64)
    lib().orc_update_interest(_p(s), float(rate))
    return s


def target_portfolio(position, value, price):
    s = np.zeros(4, dtype=np.float64)
    lib().orc_target_portfolio(_p(s), float(position), float(value), float(price))
    return s


def valorisation(state, price):
    return lib().orc_valorisation(*[float(x) for x in state], float(price))


def position_of(state, price):
    return lib().orc_position(*[float(x) for x in state], float(price))


def philox(seed, tick, env_id):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(int(seed), int(tick), int(env_id), _p(out))
    return out


# ---- the vector env ------------------------------------------------------------------------------

class OracleVecEnv:
    """N reference-semantics envs advanced in lockstep on the CPU (scalar C, one env at a time)."""

    def __init__(self, features, price, lengths=None, *, num_envs, positions, windows=None,
                 trading_fees=0.0, borrow_interest_rate=0.0, portfolio_initial_value=1000.0,
                 initial_position="random", max_episode_duration="max", dynamic_features=True,
                 done_ratio=0.7, seed=0, env_id_offset=0, plan=None, multi_dataset=False,
                 episodes_between_dataset_switch=1, dyn_mode=1, threads=1,
                 reward_kind=0, reward_scale=1.0, reward_clip=(-np.inf, np.inf), high=None, low=None):
        features = np.ascontiguousarray(features, dtype=np.float32)
        price = np.ascontiguousarray(price, dtype=np.float64)
        if features.ndim == 2:
            features = features[None]
            price = price[None]
        self.features, self.price = features, price
        n_ds, t_stride, n_static = features.shape
        self.lengths = np.ascontiguousarray(
            lengths if lengths is not None else np.full(n_ds, t_stride), dtype=np.int32)
        self.positions = np.ascontiguousarray(positions, dtype=np.float64)
        N = int(num_envs)
        self.num_envs, self.windows = N, windows
        self.n_static, self.n_dyn = n_static, (2 if dynamic_features else 0)
        self.F = self.n_static + self.n_dyn
        self.obs_shape = (N, self.F) if windows is None else (N, int(windows), self.F)
        self.plan = None if plan is None else np.ascontiguousarray(plan, dtype=np.int32)
        f64 = lambda: np.zeros(N, np.float64)  # noqa: E731
        i32 = lambda: np.zeros(N, np.int32)    # noqa: E731
        self.asset, self.fiat, self.interest_asset, self.interest_fiat, self.prev_val = f64(), f64(), f64(), f64(), f64()
        self.pos_idx, self.step_, self.ep_start, self.dataset_idx = i32(), i32(), i32(), i32()
        self.plan_cursor, self.ds_episodes, self.touched_lo, self.touched_hi = i32(), i32(), i32(), i32()
        self.ds_used = np.zeros(N, np.uint64)
        self.dyn_cols = np.zeros((N, t_stride, max(self.n_dyn, 1)), np.float32)
        self.tick = 0
        self.threads = int(threads)
        e = _OrcEnv()
        e.n_envs, e.n_positions = N, len(self.positions)
        e.windows = 0 if windows is None else int(windows)
        e.n_static, e.n_dyn = self.n_static, self.n_dyn
        e.max_episode_duration = -1 if max_episode_duration == "max" else int(max_episode_duration)
        e.n_datasets = n_ds
        e.initial_position_idx = -1 if initial_position == "random" else list(positions).index(initial_position)
        e.episodes_between_switch = int(episodes_between_dataset_switch)
        e.dyn_mode = int(dyn_mode)
        e.plan_episodes = 0 if self.plan is None else self.plan.shape[1]
        e.multi_dataset = int(bool(multi_dataset))
        e.t_stride, e.env_id_offset, e.seed = t_stride, int(env_id_offset), int(seed)
        e.fee, e.rate = float(trading_fees), float(borrow_interest_rate)
        e.v0, e.done_ratio = float(portfolio_initial_value), float(done_ratio)
        e.reward_kind, e.reward_scale = int(reward_kind), float(reward_scale)
        e.reward_lo, e.reward_hi = float(reward_clip[0]), float(reward_clip[1])
        e.positions, e.features, e.price, e.lengths = _p(self.positions), _p(features), _p(price), _p(self.lengths)
        e.asset, e.fiat = _p(self.asset), _p(self.fiat)
        e.interest_asset, e.interest_fiat, e.prev_val = _p(self.interest_asset), _p(self.interest_fiat), _p(self.prev_val)
        e.pos_idx, e.step, e.ep_start, e.dataset_idx = _p(self.pos_idx), _p(self.step_), _p(self.ep_start), _p(self.dataset_idx)
        e.plan_cursor, e.ds_used, e.ds_episodes = _p(self.plan_cursor), _p(self.ds_used), _p(self.ds_episodes)
        e.dyn_cols, e.touched_lo, e.touched_hi = _p(self.dyn_cols), _p(self.touched_lo), _p(self.touched_hi)
        e.plan = _p(self.plan)
        # limit orders (environments.py:217-231): needs the high / low columns
        self.high = None if high is None else np.ascontiguousarray(np.asarray(high, np.float64).reshape(n_ds, t_stride))
        self.low = None if low is None else np.ascontiguousarray(np.asarray(low, np.float64).reshape(n_ds, t_stride))
        self.limit_price = np.full((N, len(self.positions)), np.nan, np.float64)
        self.limit_seq = np.zeros(len(self.positions), np.int32)
        e.high, e.low, e.limit_price, e.limit_seq = _p(self.high), _p(self.low), _p(self.limit_price), _p(self.limit_seq)
        e.n_limit_positions = 0
        self._e = e
        self._lib = lib()
        self._lib.orc_init_datasets(C.byref(e), self._next_tick())
        # outputs (persistent, overwritten every step — same ownership rule as the device env)
        self.obs = np.zeros(self.obs_shape, np.float32)
        self.final_obs = np.zeros(self.obs_shape, np.float32)
        self.reward, self.valuation, self.real_position = f64(), f64(), f64()
        self.terminated, self.truncated = np.zeros(N, np.uint8), np.zeros(N, np.uint8)
        self.info_idx, self.info_step = i32(), i32()
        self.final_state = np.zeros((N, 4), np.float64)
        self.metrics = np.zeros(N_METRICS, np.float64)

    def _next_tick(self):
        t = self.tick
        self.tick += 1
        return t

    @property
    def idx(self):
        return self.ep_start + self.step_

    def add_limit_order(self, position, limit, persistent=True, env_ids=None):
        """TradingEnv.add_limit_order (environments.py:227-231) for the given envs (default: all)."""
        assert persistent, "the reference's non-persistent branch raises RuntimeError when it executes (:223)"
        assert self.high is not None and self.low is not None, "limit orders need the high / low columns"
        pk = list(self.positions).index(position)
        n = self._e.n_limit_positions
        if pk not in self.limit_seq[:n]:
            self.limit_seq[n] = pk
            self._e.n_limit_positions = n + 1
        ids = slice(None) if env_ids is None else np.asarray(env_ids)
        self.limit_price[ids, pk] = limit

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self._lib.orc_reset(C.byref(self._e), _p(m), self._next_tick(), _p(self.obs))
        return self.obs

    def step(self, actions, want_obs=True, want_final_obs=False):
        a = np.ascontiguousarray(actions, dtype=np.int64)
        assert a.shape == (self.num_envs,)
        tick = self._next_tick()
        self.metrics[:] = 0
        obs = _p(self.obs) if want_obs else None
        fobs = _p(self.final_obs) if want_final_obs else None
        if self.threads <= 1:
            self._lib.orc_step(C.byref(self._e), _p(a), tick, obs, _p(self.reward),
                               _p(self.terminated), _p(self.truncated), _p(self.valuation),
                               _p(self.real_position), _p(self.info_idx), _p(self.info_step),
                               _p(self.final_state), fobs, _p(self.metrics))
        else:
            # one env slice per host thread (ctypes releases the GIL); per-thread metric partials
            N, nt = self.num_envs, self.threads
            parts = np.zeros((nt, N_METRICS), np.float64)
            bounds = [(N * t) // nt for t in range(nt + 1)]

            def work(t):
                self._lib.orc_step_range(C.byref(self._e), bounds[t], bounds[t + 1], _p(a), tick, obs,
                                         _p(self.reward), _p(self.terminated), _p(self.truncated),
                                         _p(self.valuation), _p(self.real_position), _p(self.info_idx),
                                         _p(self.info_step), _p(self.final_state), fobs, _p(parts[t]))
            ths = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
            [t.start() for t in ths]
            [t.join() for t in ths]
            self.metrics[:] = parts.sum(0)
        return self.obs, self.reward, self.terminated, self.truncated

    def prefault(self):
        """Touch every page of the per-env private dynamic columns (the reference materialises each env's arrays at
        construction): keeps first-touch page faults out of a timed region."""
        self.dyn_cols.fill(0)

    def rollout_timed(self, action_sets, iters, pin=True):
        """`iters` lockstep iterations on `self.threads` PINNED pthreads created inside C (orc_rollout_threads), each
        rolling its own env slice forward; returns the wall seconds measured in C (bench.py's CPU baseline)."""
        a = np.ascontiguousarray(action_sets, dtype=np.int64)
        assert a.ndim == 2 and a.shape[1] == self.num_envs
        nt = max(1, self.threads)
        parts = np.zeros((nt, N_METRICS), np.float64)
        tick0 = self.tick
        self.tick += iters
        dt = self._lib.orc_rollout_threads(C.byref(self._e), _p(a), a.shape[0], int(iters), tick0, _p(self.obs),
                                           _p(self.reward), _p(self.terminated), _p(self.truncated), _p(parts), nt, int(pin))
        self.metrics[:] = parts.sum(0)
        return dt

    def rollout(self, action_sets, iters):
        """`iters` lockstep iterations with actions cycling through action_sets[A, N]; each host thread
        rolls its own env slice forward inside C (envs are independent)."""
        a = np.ascontiguousarray(action_sets, dtype=np.int64)
        assert a.ndim == 2 and a.shape[1] == self.num_envs
        N, nt = self.num_envs, max(1, self.threads)
        parts = np.zeros((nt, N_METRICS), np.float64)
        bounds = [(N * t) // nt for t in range(nt + 1)]
        tick0 = self.tick
        self.tick += iters

        def work(t):
            self._lib.orc_rollout_range(C.byref(self._e), bounds[t], bounds[t + 1], _p(a), a.shape[0], int(iters),
                                        tick0, _p(self.obs), _p(self.reward), _p(self.terminated),
                                        _p(self.truncated), _p(parts[t]))
        if nt == 1:
            work(0)
        else:
            ths = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
            [t.start() for t in ths]
            [t.join() for t in ths]
        self.metrics[:] = parts.sum(0)
        return self.metrics
