// gte_device.cuh — device-side building blocks shared by the step / reset / info kernels.
//
// Money math = the reference's Portfolio (src/gym_trading_env/utils/portfolio.py) restated in IEEE
// fp64 with explicit round-to-nearest intrinsics (__dmul_rn/__dadd_rn/__ddiv_rn), which nvcc never
// contracts into FMA, in exactly the Python operation order — so asset/fiat/interest/valuation/
// real_position are bit-identical to the CPU reference.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gte_b200.h"

namespace gte {

constexpr int kStepThreads = GTE_STEP_THREADS;
constexpr int kMaxPartialRows = GTE_MAX_PARTIAL_ROWS;     // metric_partials rows the step kernel may use (grid cap)

// Programmatic dependent launch: wait until the preceding kernel of the stream has completed and its writes are visible
// (returns at once for a kernel that was not launched with the attribute).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct Portfolio {        // signed state of utils/portfolio.py:2-6
    double asset, fiat, ia, ifi;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
// IEEE division.  The inline fast path of the fp64 divide hands a ZERO numerator to a ~250-instruction subroutine,
// and zero numerators are everyday values here (a flat position: asset = 0, target = 0), so one flat env per warp
// would send the whole warp there several times per step.  0 / b for a finite non-zero b is a zero whose sign is
// the XOR of the operand signs: answer that directly and divide 1 / b instead (result discarded) — bit-identical.
__device__ __forceinline__ double ddiv(double a, double b) {
    const bool zero_num = (a == 0.0) && (b != 0.0) && (fabs(b) <= 1.7976931348623157e308);   // false for NaN / inf / 0
    double num = zero_num ? 1.0 : a;
    asm("" : "+d"(num));      // opaque: otherwise the compiler divides the original (zero) numerator, whose quotient is dead
    const double q = __ddiv_rn(num, b);
    const double z = __longlong_as_double((__double_as_longlong(a) ^ __double_as_longlong(b)) & (long long)0x8000000000000000ull);
    return zero_num ? z : q;
}

// Portfolio.valorisation (portfolio.py:7-13): sum([asset*price, fiat, -ia*price, -if]) left to right.
__device__ __forceinline__ double valorisation(const Portfolio& s, double price) {
    double v = dmul(s.asset, price);
    v = dadd(v, s.fiat);
    v = dadd(v, dmul(-s.ia, price));
    v = dadd(v, -s.ifi);
    return v;
}

// Portfolio.real_position (portfolio.py:14-15), with the valuation already at hand.
__device__ __forceinline__ double real_position(const Portfolio& s, double price, double val) {
    return ddiv(dmul(dsub(s.asset, s.ia), price), val);
}

// Portfolio.trade_to_position (portfolio.py:18-43).  `val` = valorisation(s, price) on entry.
__device__ __forceinline__ void trade_to_position(Portfolio& s, double target, double price,
                                                  double fee, double val) {
    const double current = ddiv(dmul(s.asset, price), val);                  // :20 position(price)
    double r = 1.0;                                                          // :21
    if (target <= 0.0 && current < 0.0) {                                    // :22-23
        const double q = ddiv(target, current);
        r = (q < 1.0) ? q : 1.0;                                             // min(1, q): first arg on ties
    } else if (target >= 1.0 && current > 1.0) {                             // :24-25
        const double q = ddiv(dsub(target, 1.0), dsub(current, 1.0));
        r = (q < 1.0) ? q : 1.0;
    }
    if (r < 1.0) {                                                           // :26-30 repay interest
        const double k = dsub(1.0, r);
        s.asset = dsub(s.asset, dmul(k, s.ia));
        s.fiat = dsub(s.fiat, dmul(k, s.ifi));
        s.ia = dmul(r, s.ia);
        s.ifi = dmul(r, s.ifi);
        val = valorisation(s, price);                                        // :33 recomputed after repay
    }
    double trade = dsub(ddiv(dmul(target, val), price), s.asset);            // :33
    if (trade > 0.0) {                                                       // :34-38 buy
        trade = ddiv(trade, dadd(dsub(1.0, fee), dmul(fee, target)));
        s.asset = dadd(s.asset, dmul(trade, dsub(1.0, fee)));
        s.fiat = dadd(s.fiat, dmul(-trade, price));
    } else {                                                                 // :39-43 sell
        trade = ddiv(trade, dsub(1.0, dmul(fee, target)));
        s.asset = dadd(s.asset, trade);
        s.fiat = dadd(s.fiat, dmul(dmul(-trade, price), dsub(1.0, fee)));
    }
}

// Portfolio.update_interest (portfolio.py:44-46).  Python's max(0, -x) returns the FIRST argument on
// ties, so -0.0 never appears: written as a compare-select, not fmax.
__device__ __forceinline__ void update_interest(Portfolio& s, double rate) {
    const double na = -s.asset, nf = -s.fiat;
    s.ia = dmul((na > 0.0) ? na : 0.0, rate);
    s.ifi = dmul((nf > 0.0) ? nf : 0.0, rate);
}

// TargetPortfolio.__init__ (portfolio.py:59-66)
__device__ __forceinline__ Portfolio target_portfolio(double position, double value, double price) {
    Portfolio s;
    s.asset = ddiv(dmul(position, value), price);
    s.fiat = dmul(dsub(1.0, position), value);
    s.ia = 0.0;
    s.ifi = 0.0;
    return s;
}

// ---- Philox4x32-10 (Salmon et al., SC'11; Random123 known-answer vectors in tests/) -------------
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ void philox_draw(uint64_t seed, uint64_t tick, uint64_t env_id, uint32_t r[4]) {
    r[0] = (uint32_t)tick; r[1] = (uint32_t)(tick >> 32);
    r[2] = (uint32_t)env_id; r[3] = (uint32_t)(env_id >> 32);
    philox4x32_10(r, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// MultiDatasetTradingEnv.next_dataset (environments.py:380-391): uniform among the least-used
// datasets == among those not yet used in the current round (use counts differ by at most one).
__device__ __forceinline__ int next_dataset(const GteParams& P, const GteState& S, int i, uint32_t u) {
    const uint64_t full = (P.n_datasets >= 64) ? ~0ull : ((1ull << P.n_datasets) - 1ull);
    uint64_t used = S.ds_used[i];
    if ((used & full) == full) used = 0;
    const int n_free = P.n_datasets - __popcll(used & full);
    int pick = (int)(((uint64_t)u * (uint64_t)n_free) >> 32);                // :385
    int ds = 0;
    for (int d = 0; d < P.n_datasets; ++d) {
        if (!((used >> d) & 1ull)) {
            if (pick == 0) { ds = d; break; }
            --pick;
        }
    }
    S.ds_used[i] = used | (1ull << ds);                                      // :388
    S.ds_episodes[i] = 0;                                                    // :381
    return ds;
}

// Dynamic-feature ring (time-indexed, one block per 32-env tile: include/gte_b200.h): the row env i observes at
// lockstep iteration c lives in slot c % W.  All envs write the same slot in one iteration: a warp stores one
// full 128-byte line of real_position values and one full 32-byte sector of position indices.
__device__ __forceinline__ int ring_slot_of(const GteParams& P, uint64_t clock) {
    const int W = P.windows > 0 ? P.windows : 1;
    return (int)(clock % (uint64_t)W);
}
__device__ __forceinline__ const uint8_t* ring_tile(const GteState& S, int W, int64_t env) {
    return S.dyn_ring + (env >> 5) * (int64_t)GTE_RING_TILE_BYTES(W);
}
__device__ __forceinline__ void ring_store(const GteParams& P, const GteState& S, int i, int slot, float rp, int pos_idx) {
    const int W = P.windows > 0 ? P.windows : 1;
    const int e = i & 31;
    uint8_t* tile = S.dyn_ring + (int64_t)(i >> 5) * GTE_RING_TILE_BYTES(W);
    *reinterpret_cast<float*>(tile + GTE_RING_RP_OFFSET(slot, e)) = rp;
    tile[GTE_RING_POS_OFFSET(W, slot, e)] = (uint8_t)pos_idx;
}

struct EnvRegs {          // the per-env state a step keeps in registers
    Portfolio pf;
    int pos_idx, step, ep_start, ds;
};

// TradingEnv.reset (environments.py:163-199) preceded by MultiDatasetTradingEnv.reset (:393-400).
// Rare path (once per episode): plan/rotation bookkeeping goes straight to global memory.
__device__ __forceinline__ void reset_env(const GteParams& P, const GteData& D, const GteState& S,
                                       int i, uint64_t tick, int ring_slot, EnvRegs& e) {
    uint32_t r[4];
    philox_draw(P.seed, tick, (uint64_t)(P.env_id_offset + i), r);
    const bool have_plan = P.plan_episodes > 0 && S.reset_plan != nullptr;
    int plan_start = 0, plan_pos = 0, plan_ds = 0;
    if (have_plan) {
        const int cur = S.plan_cursor[i];
        if (cur >= P.plan_episodes) atomicOr(S.error_flag, GTE_E_PLAN_EXHAUSTED);     // replays from the start
        const int32_t* pl = S.reset_plan + ((int64_t)i * P.plan_episodes + (cur % P.plan_episodes)) * 3;
        plan_start = pl[0]; plan_pos = pl[1]; plan_ds = pl[2];
        S.plan_cursor[i] = cur + 1;
        // a plan row must name a dataset, a position and a start row the reference could have drawn
        // (environments.py:167-177): clamp and flag instead of reading out of bounds
        bool bad = plan_pos < 0 || plan_pos >= P.n_positions || plan_ds < 0 || plan_ds >= P.n_datasets;
        plan_pos = min(max(plan_pos, 0), P.n_positions - 1);
        plan_ds = P.multi_dataset ? min(max(plan_ds, 0), P.n_datasets - 1) : e.ds;
        const int Tp = D.lengths[plan_ds], lo = P.windows > 0 ? P.windows - 1 : 0;
        if (plan_start < lo || plan_start > Tp - 2) { bad = true; plan_start = min(max(plan_start, lo), Tp - 2); }
        if (bad) atomicOr(S.error_flag, GTE_E_PLAN_RANGE);
    }
    if (P.multi_dataset) {                                                   // :394-398
        const int n = S.ds_episodes[i] + 1;
        if (n % P.episodes_between_switch == 0) {
            if (have_plan) S.ds_episodes[i] = 0;
            else e.ds = next_dataset(P, S, i, r[3]);
        } else {
            S.ds_episodes[i] = n;
        }
        // a plan records the dataset every episode ran on (incl. the one drawn in __init__, :378)
        if (have_plan) e.ds = plan_ds;
    }
    const int T = D.lengths[e.ds];
    e.step = 0;                                                              // :166
    if (have_plan) e.pos_idx = plan_pos;                                     // :167
    else if (P.initial_position_idx >= 0) e.pos_idx = P.initial_position_idx;
    else e.pos_idx = (int)(((uint64_t)r[2] * (uint64_t)P.n_positions) >> 32);
    int start = P.windows > 0 ? P.windows - 1 : 0;                           // :171-172
    if (have_plan) start = plan_start;
    else if (P.max_episode_duration >= 0) {                                  // :173-177 randint(low, high)
        const int64_t lo = start, hi = (int64_t)T - P.max_episode_duration - start;
        const uint64_t x = ((uint64_t)r[1] << 32) | r[0];
        start = (int)(lo + (int64_t)__umul64hi(x, (uint64_t)(hi - lo)));
    }
    e.ep_start = start;
    const double position = P.positions[e.pos_idx];
    const double price = D.price[(int64_t)e.ds * P.t_stride + start];
    e.pf = target_portfolio(position, P.v0, price);                          // :179-183
    for (int k = 0; k < P.n_limit_positions; ++k)                            // self._limit_orders = {} (:168)
        S.limit_price[(int64_t)i * P.n_positions + S.limit_seq[k]] = __longlong_as_double(0x7ff8000000000000ll);
    if (P.n_dyn > 0)                                                         // first obs row: (position, position) :191-192
        ring_store(P, S, i, ring_slot, (float)position, e.pos_idx);
}

}  // namespace gte
