/*
 * gte_oracle.c — CPU restatement of Gym-Trading-Env's per-step simulation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gym-trading-env_b200/) may import, link
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs use it, and there only as the checker / the timed CPU baseline.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this restatement
 * is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed unmodified in the build container
 * by oracle/make_golden.py (fixtures committed under tests/golden/, checked by
 * tests/test_oracle_golden.py), and live against /root/reference when it is mounted
 * (tests/test_oracle_vs_reference.py).
 *
 * Each function cites the reference lines it follows; paths are relative to
 * /root/reference/src/gym_trading_env/.  All money math is IEEE fp64, round-to-nearest, in the
 * reference's exact operation order; build with -ffp-contract=off so no FMA is formed.
 *
 * Plain scalar C on purpose: one env at a time, like the reference, so it also serves as the
 * "port" CPU baseline (one thread per env slice; bench.py splits envs over host threads).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>
#include <time.h>

#define ORC_N_METRICS 8
/* metric slots (sums over finished episodes unless noted) */
enum { M_EPISODES = 0, M_TERMINATED, M_TRUNCATED, M_SUM_PORTFOLIO_RETURN, M_SUM_MARKET_RETURN,
       M_SUM_EPISODE_LENGTH, M_SUM_REWARD /* all envs, this call */, M_RESERVED };

typedef struct OrcEnv {
    /* ---- parameters: TradingEnv.__init__ (environments.py:79-110) ---- */
    int32_t n_envs;
    int32_t n_positions;
    int32_t windows;                 /* 0 = None */
    int32_t n_static;                /* static "feature*" columns (:130) */
    int32_t n_dyn;                   /* 0 or 2: the two default dynamic features (:20-24) */
    int32_t max_episode_duration;    /* -1 = 'max' */
    int32_t n_datasets;
    int32_t initial_position_idx;    /* -1 = 'random' (:167) */
    int32_t episodes_between_switch; /* MultiDatasetTradingEnv (:370) */
    int32_t dyn_mode;                /* 0 = faithful private columns (stale rows leak, H3); 1 = zeroed on reset */
    int32_t plan_episodes;           /* E of plan[N,E,3]; 0 = no plan -> Philox */
    int32_t multi_dataset;           /* 1 = MultiDatasetTradingEnv semantics (:393-400) */
    int32_t reward_kind;             /* 0 = log-return (basic_reward_function :17-18), 1 = simple return */
    int32_t n_limit_positions;       /* positions carrying persistent limit orders (:217-231), 0 = none */
    int64_t t_stride;                /* rows allocated per dataset */
    int64_t env_id_offset;           /* global id of env 0 (multi-GPU sharding) */
    uint64_t seed;
    double fee, rate, v0, done_ratio;
    double reward_scale, reward_lo, reward_hi;   /* reward = clip(scale * f, lo, hi): the callers' variants
                                                    (luckymodel/envs/env.py:16-18, scripts/test_env.py:20-22) */
    /* ---- data: _set_df (:128-143) ---- */
    const double* positions;         /* [P] */
    const float* features;           /* [n_datasets, t_stride, n_static] */
    const double* price;             /* [n_datasets, t_stride] */
    const int32_t* lengths;          /* [n_datasets] */
    /* ---- state, structure of arrays ---- */
    double* asset; double* fiat; double* interest_asset; double* interest_fiat;   /* portfolio.py:2-6 */
    double* prev_val;                /* history["portfolio_valuation", -1] of the previous row */
    int32_t* pos_idx; int32_t* step; int32_t* ep_start; int32_t* dataset_idx;
    int32_t* plan_cursor; uint64_t* ds_used; int32_t* ds_episodes;
    float* dyn_cols;                 /* [N, t_stride, n_dyn]: each env's private dynamic columns (:135-138,153-154) */
    int32_t* touched_lo; int32_t* touched_hi;   /* rows of dyn_cols written since the last zeroing */
    const int32_t* plan;             /* [N, E, 3] (start idx, position idx, dataset idx) or NULL */
    const double* high; const double* low;   /* [n_datasets, t_stride] "high"/"low" columns (limit orders, :221) */
    double* limit_price;             /* [N, P] limit per target position, NaN = none (add_limit_order :227-231) */
    const int32_t* limit_seq;        /* [n_limit_positions] position indices in dict insertion order (:220) */
} OrcEnv;

/* ------------------------------------------------------------------ Portfolio (utils/portfolio.py) */

/* Portfolio.valorisation, portfolio.py:7-13: sum([asset*price, fiat, -ia*price, -if]) left to right */
double orc_valorisation(double asset, double fiat, double ia, double ifi, double price) {
    double s = asset * price;
    s = s + fiat;
    s = s + (-ia) * price;
    s = s + (-ifi);
    return s;
}

/* Portfolio.real_position, portfolio.py:14-15 */
double orc_real_position(double asset, double fiat, double ia, double ifi, double price) {
    return ((asset - ia) * price) / orc_valorisation(asset, fiat, ia, ifi, price);
}

/* Portfolio.position, portfolio.py:16-17 */
double orc_position(double asset, double fiat, double ia, double ifi, double price) {
    return (asset * price) / orc_valorisation(asset, fiat, ia, ifi, price);
}

/* Portfolio.trade_to_position, portfolio.py:18-43.  s = {asset, fiat, interest_asset, interest_fiat} */
void orc_trade_to_position(double* s, double position, double price, double fee) {
    double asset = s[0], fiat = s[1], ia = s[2], ifi = s[3];
    double current = orc_position(asset, fiat, ia, ifi, price);                       /* :20 */
    double r = 1.0;                                                                   /* :21 */
    if (position <= 0 && current < 0) {                                               /* :22 */
        double q = position / current;                                                /* :23 min(1, q) */
        r = (q < 1.0) ? q : 1.0;
    } else if (position >= 1 && current > 1) {                                        /* :24 */
        double q = (position - 1.0) / (current - 1.0);                                /* :25 */
        r = (q < 1.0) ? q : 1.0;
    }
    if (r < 1.0) {                                                                    /* :26-30 */
        asset = asset - (1.0 - r) * ia;
        fiat = fiat - (1.0 - r) * ifi;
        ia = r * ia;
        ifi = r * ifi;
    }
    double trade = (position * orc_valorisation(asset, fiat, ia, ifi, price)) / price - asset;   /* :33 */
    if (trade > 0) {                                                                  /* :34-38 */
        trade = trade / ((1.0 - fee) + fee * position);
        double af = (-trade) * price;
        asset = asset + trade * (1.0 - fee);
        fiat = fiat + af;
    } else {                                                                          /* :39-43 */
        trade = trade / (1.0 - fee * position);
        double af = (-trade) * price;
        asset = asset + trade;
        fiat = fiat + af * (1.0 - fee);
    }
    s[0] = asset; s[1] = fiat; s[2] = ia; s[3] = ifi;
}

/* Portfolio.update_interest, portfolio.py:44-46: max(0, -x) returns the FIRST argument on ties */
void orc_update_interest(double* s, double rate) {
    s[2] = ((-s[0] > 0) ? -s[0] : 0.0) * rate;
    s[3] = ((-s[1] > 0) ? -s[1] : 0.0) * rate;
}

/* TargetPortfolio.__init__, portfolio.py:59-66 */
void orc_target_portfolio(double* s, double position, double value, double price) {
    s[0] = (position * value) / price;
    s[1] = (1.0 - position) * value;
    s[2] = 0.0;
    s[3] = 0.0;
}

/* ------------------------------------------------------------------ counter-based RNG (production resets) */
/* Philox4x32-10 (Salmon et al., SC'11).  Same constants in the CUDA path (csrc/gte_philox.cuh). */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int i = 0; i < 10; ++i) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void orc_philox(uint64_t seed, uint64_t tick, uint64_t env_id, uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)tick, (uint32_t)(tick >> 32), (uint32_t)env_id, (uint32_t)(env_id >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    memcpy(out, c, sizeof(c));
}

static inline uint64_t mulhi64(uint64_t a, uint64_t b) { return (uint64_t)(((__uint128_t)a * b) >> 64); }

/* ------------------------------------------------------------------ observation (environments.py:152-160) */

static void orc_write_dyn(const OrcEnv* e, int i, int row, float f0, float f1) {
    if (e->n_dyn == 0) return;
    float* d = e->dyn_cols + ((int64_t)i * e->t_stride + row) * e->n_dyn;   /* :153-154 write-back, fp64->fp32 */
    d[0] = f0;
    if (e->n_dyn > 1) d[1] = f1;
    if (row < e->touched_lo[i]) e->touched_lo[i] = row;
    if (row > e->touched_hi[i]) e->touched_hi[i] = row;
}

static void orc_zero_dyn(const OrcEnv* e, int i) {
    if (e->n_dyn == 0) return;
    if (e->touched_hi[i] >= e->touched_lo[i]) {
        float* d = e->dyn_cols + ((int64_t)i * e->t_stride + e->touched_lo[i]) * e->n_dyn;
        memset(d, 0, sizeof(float) * (size_t)(e->touched_hi[i] - e->touched_lo[i] + 1) * e->n_dyn);
    }
    e->touched_lo[i] = 0x7fffffff;
    e->touched_hi[i] = -1;
}

/* _get_obs, :156-160: row idx (windows=None) or rows idx+1-W .. idx */
void orc_get_obs(const OrcEnv* e, int i, float* obs) {
    int F = e->n_static + e->n_dyn;
    int idx = e->ep_start[i] + e->step[i];
    int ds = e->dataset_idx[i];
    int W = e->windows > 0 ? e->windows : 1;
    for (int w = 0; w < W; ++w) {
        int row = idx + 1 - W + w;
        const float* st = e->features + ((int64_t)ds * e->t_stride + row) * e->n_static;
        memcpy(obs + w * F, st, sizeof(float) * (size_t)e->n_static);
        const float* d = e->dyn_cols + ((int64_t)i * e->t_stride + row) * e->n_dyn;
        memcpy(obs + w * F + e->n_static, d, sizeof(float) * (size_t)e->n_dyn);
    }
}

/* ------------------------------------------------------------------ reset */

/* MultiDatasetTradingEnv.next_dataset, :380-391: uniform among the least-used datasets.
 * Use counts never differ by more than one, so "least used" == "not yet used this round". */
static int orc_next_dataset(OrcEnv* e, int i, uint32_t u) {
    uint64_t full = (e->n_datasets >= 64) ? ~0ull : ((1ull << e->n_datasets) - 1ull);
    uint64_t used = e->ds_used[i];
    if ((used & full) == full) used = 0;
    int n_free = e->n_datasets - __builtin_popcountll(used & full);
    int pick = (int)(((uint64_t)u * (uint64_t)n_free) >> 32);            /* :385 randint(size) */
    int ds = 0;
    for (int d = 0; d < e->n_datasets; ++d) {
        if (!((used >> d) & 1ull)) { if (pick == 0) { ds = d; break; } --pick; }
    }
    e->ds_used[i] = used | (1ull << ds);                                 /* :388 */
    e->ds_episodes[i] = 0;                                               /* :381 */
    return ds;
}

/* MultiDatasetTradingEnv.__init__ draws one dataset before any episode (:378). */
void orc_init_datasets(OrcEnv* e, uint64_t tick) {
    for (int i = 0; i < e->n_envs; ++i) {
        e->ds_used[i] = 0; e->ds_episodes[i] = 0; e->dataset_idx[i] = 0;
        e->touched_lo[i] = 0x7fffffff; e->touched_hi[i] = -1;
        if (e->multi_dataset && e->plan_episodes == 0) {
            uint32_t r[4];
            orc_philox(e->seed, tick, (uint64_t)(e->env_id_offset + i), r);
            e->dataset_idx[i] = orc_next_dataset(e, i, r[3]);
        }
    }
}

/* TradingEnv.reset (:163-199) preceded by MultiDatasetTradingEnv.reset (:393-400) */
void orc_reset_env(OrcEnv* e, int i, uint64_t tick) {
    uint32_t r[4];
    orc_philox(e->seed, tick, (uint64_t)(e->env_id_offset + i), r);
    int have_plan = e->plan_episodes > 0 && e->plan != 0;
    const int32_t* pl = 0;
    if (have_plan) {
        pl = e->plan + ((int64_t)i * e->plan_episodes + (e->plan_cursor[i] % e->plan_episodes)) * 3;
        e->plan_cursor[i] += 1;
    }
    /* dataset (:394-398) */
    int switched = 0;
    if (e->multi_dataset) {
        e->ds_episodes[i] += 1;
        if (e->ds_episodes[i] % e->episodes_between_switch == 0) {
            if (have_plan) e->ds_episodes[i] = 0;
            else e->dataset_idx[i] = orc_next_dataset(e, i, r[3]);
            switched = 1;                                  /* _set_df: fresh zero dynamic columns */
        }
        /* a plan records the dataset every episode ran on (incl. the one drawn in __init__, :378) */
        if (have_plan) e->dataset_idx[i] = pl[2];
    }
    if (switched || e->dyn_mode == 1) orc_zero_dyn(e, i);
    int ds = e->dataset_idx[i];
    int T = e->lengths[ds];
    e->step[i] = 0;                                                      /* :166 */
    int pidx;                                                            /* :167 */
    if (have_plan) pidx = pl[1];
    else if (e->initial_position_idx >= 0) pidx = e->initial_position_idx;
    else pidx = (int)(((uint64_t)r[2] * (uint64_t)e->n_positions) >> 32);
    int start = e->windows > 0 ? e->windows - 1 : 0;                     /* :171-172 */
    if (have_plan) start = pl[0];
    else if (e->max_episode_duration >= 0) {                             /* :173-177 randint(low, high) */
        int64_t lo = start, hi = (int64_t)T - e->max_episode_duration - start;
        uint64_t x = ((uint64_t)r[1] << 32) | r[0];
        start = (int)(lo + (int64_t)mulhi64(x, (uint64_t)(hi - lo)));
    }
    e->ep_start[i] = start;
    e->pos_idx[i] = pidx;
    double position = e->positions[pidx];
    double price = e->price[(int64_t)ds * e->t_stride + start];
    double s[4];
    orc_target_portfolio(s, position, e->v0, price);                     /* :179-183 */
    e->asset[i] = s[0]; e->fiat[i] = s[1]; e->interest_asset[i] = s[2]; e->interest_fiat[i] = s[3];
    e->prev_val[i] = e->v0;                                              /* :194 */
    for (int k = 0; k < e->n_limit_positions; ++k)                       /* self._limit_orders = {} (:168) */
        e->limit_price[(int64_t)i * e->n_positions + e->limit_seq[k]] = NAN;
    orc_write_dyn(e, i, start, (float)position, (float)position);        /* :191-192 + :153-154 */
}

void orc_reset(OrcEnv* e, const uint8_t* mask, uint64_t tick, float* obs) {
    int F = e->n_static + e->n_dyn;
    int64_t obs_sz = (int64_t)(e->windows > 0 ? e->windows : 1) * F;
    for (int i = 0; i < e->n_envs; ++i) {
        if (mask && !mask[i]) continue;
        orc_reset_env(e, i, tick);
        if (obs) orc_get_obs(e, i, obs + i * obs_sz);
    }
}

/* ------------------------------------------------------------------ step (environments.py:233-272) */

/* One lockstep iteration over envs [lo, hi) with in-place auto-reset (SURVEY.md §8a contract).
 * Outputs are the TERMINAL-step values for reward/flags/valuation/idx/step; obs and state are
 * post-reset where a reset happened.  final_obs (optional) receives what step() itself returned. */
void orc_step_range(OrcEnv* e, int lo, int hi, const int64_t* actions, uint64_t tick,
                    float* obs, double* reward, uint8_t* terminated, uint8_t* truncated,
                    double* valuation, double* real_position, int32_t* info_idx, int32_t* info_step,
                    double* final_state /* [N,4] or NULL */, float* final_obs, double* metrics) {
    int F = e->n_static + e->n_dyn;
    int64_t obs_sz = (int64_t)(e->windows > 0 ? e->windows : 1) * F;
    for (int i = lo; i < hi; ++i) {
        int ds = e->dataset_idx[i];
        const double* price = e->price + (int64_t)ds * e->t_stride;
        int T = e->lengths[ds];
        int idx = e->ep_start[i] + e->step[i];
        double s[4] = {e->asset[i], e->fiat[i], e->interest_asset[i], e->interest_fiat[i]};
        int64_t a = actions[i];
        if (a >= 0) {                                                    /* :234 (None = hold) */
            double target = e->positions[a];
            if (target != e->positions[e->pos_idx[i]]) {                 /* :213-215 value compare */
                orc_trade_to_position(s, target, price[idx], e->fee);    /* :204-211, price BEFORE advancing */
                e->pos_idx[i] = (int32_t)a;
            }
        }
        idx += 1;                                                        /* :235 */
        int step = e->step[i] + 1;                                       /* :236 */
        if (e->n_limit_positions > 0) {                                  /* _take_action_order_limit, :217-223 */
            double hi = e->high[(int64_t)ds * e->t_stride + idx], lo = e->low[(int64_t)ds * e->t_stride + idx];
            for (int k = 0; k < e->n_limit_positions; ++k) {
                int pk = e->limit_seq[k];
                double lim = e->limit_price[(int64_t)i * e->n_positions + pk];
                if (lim == lim && e->positions[pk] != e->positions[e->pos_idx[i]] && lim <= hi && lim >= lo) {  /* :221 */
                    orc_trade_to_position(s, e->positions[pk], lim, e->fee);   /* :222 _trade(position, price=limit) */
                    e->pos_idx[i] = pk;
                }
            }
        }
        double p = price[idx];                                           /* :239 */
        orc_update_interest(s, e->rate);                                 /* :240 */
        double val = orc_valorisation(s[0], s[1], s[2], s[3], p);        /* :241 */
        int done = (val / e->v0) <= e->done_ratio;                       /* :246 (0.7 in this fork) */
        int trunc = idx >= T - 1;                                        /* :248 */
        if (e->max_episode_duration >= 0 && step >= e->max_episode_duration - 1) trunc = 1;   /* :250 */
        double rp = ((s[0] - s[2]) * p) / val;                           /* :259 real_position */
        double rew = 0.0;                                                /* :263 */
        if (!done) {                                                     /* :265-267 reward_function(history) */
            double x = e->reward_kind == 1 ? (val - e->prev_val[i]) / e->prev_val[i]
                                           : log(val / e->prev_val[i]);  /* :17-18 */
            x = e->reward_scale * x;
            x = (x < e->reward_lo) ? e->reward_lo : x;                   /* np.clip */
            x = (x > e->reward_hi) ? e->reward_hi : x;
            rew = x;
        }
        e->asset[i] = s[0]; e->fiat[i] = s[1]; e->interest_asset[i] = s[2]; e->interest_fiat[i] = s[3];
        e->step[i] = step;
        e->prev_val[i] = val;
        double position = e->positions[e->pos_idx[i]];
        orc_write_dyn(e, i, idx, (float)position, (float)rp);            /* :153-154 via :20-24 */
        reward[i] = rew; terminated[i] = (uint8_t)done; truncated[i] = (uint8_t)trunc;
        if (valuation) valuation[i] = val;
        if (real_position) real_position[i] = rp;
        if (info_idx) info_idx[i] = idx;
        if (info_step) info_step[i] = step;
        if (final_state) { final_state[4 * i] = s[0]; final_state[4 * i + 1] = s[1];
                           final_state[4 * i + 2] = s[2]; final_state[4 * i + 3] = s[3]; }
        if (final_obs) orc_get_obs(e, i, final_obs + i * obs_sz);
        if (metrics) metrics[M_SUM_REWARD] += rew;
        if (done || trunc) {                                             /* :269-271 + vector auto-reset */
            if (metrics) {
                metrics[M_EPISODES] += 1.0;
                metrics[M_TERMINATED] += done ? 1.0 : 0.0;
                metrics[M_TRUNCATED] += trunc ? 1.0 : 0.0;
                metrics[M_SUM_PORTFOLIO_RETURN] += val / e->v0 - 1.0;                  /* :282 */
                metrics[M_SUM_MARKET_RETURN] += p / price[e->ep_start[i]] - 1.0;       /* :281 */
                metrics[M_SUM_EPISODE_LENGTH] += (double)step;
            }
            orc_reset_env(e, i, tick);
        }
        if (obs) orc_get_obs(e, i, obs + i * obs_sz);                    /* :272 / reset :199 */
    }
}

void orc_step(OrcEnv* e, const int64_t* actions, uint64_t tick,
              float* obs, double* reward, uint8_t* terminated, uint8_t* truncated,
              double* valuation, double* real_position, int32_t* info_idx, int32_t* info_step,
              double* final_state, float* final_obs, double* metrics) {
    orc_step_range(e, 0, e->n_envs, actions, tick, obs, reward, terminated, truncated, valuation,
                   real_position, info_idx, info_step, final_state, final_obs, metrics);
}

/* K lockstep iterations over the env slice [lo, hi) with actions[(k % n_sets), :] — envs are
 * independent, so host threads can each roll their own slice forward without meeting every
 * iteration (used only by bench.py's CPU-baseline legs). */
void orc_rollout_range(OrcEnv* e, int lo, int hi, const int64_t* actions, int n_sets, int iters, uint64_t tick0,
                       float* obs, double* reward, uint8_t* terminated, uint8_t* truncated, double* metrics) {
    for (int k = 0; k < iters; ++k)
        orc_step_range(e, lo, hi, actions + (int64_t)(k % n_sets) * e->n_envs, tick0 + (uint64_t)k, obs, reward,
                       terminated, truncated, 0, 0, 0, 0, 0, 0, metrics);
}

/* ------------------------------------------------------------------ CPU baseline driver (bench.py only) */
/* The same rollout on n_threads host threads, each PINNED to one of the process's allowed CPUs.  Envs are independent,
 * so there is no barrier between iterations (like the per-env worker processes of an AsyncVectorEnv): the envs are cut
 * into chunks of ORC_CHUNK_ENVS that the threads claim from a shared counter and roll forward all `iters` iterations —
 * a core that is briefly taken by something else then simply claims fewer chunks instead of holding the whole run up.
 * A chunk is 64 envs so that two threads never write the same cache line of the per-env byte arrays (terminated /
 * truncated: 16-env chunks cost the windows=None case a factor 3 in false sharing).
 * All threads start together behind a barrier; the return value is the wall time in seconds from that barrier to the
 * last thread's exit, measured here so that no interpreter overhead is inside it.
 * metrics: [n_threads][ORC_N_METRICS] partial sums. */
#define ORC_CHUNK_ENVS 64
typedef struct OrcWork {
    OrcEnv* e; const int64_t* actions; int n_sets, iters; uint64_t tick0;
    float* obs; double* reward; uint8_t* terminated; uint8_t* truncated; double* metrics;
    int cpu; pthread_barrier_t* bar; int* next_chunk; int n_chunks;
} OrcWork;

static void* orc_worker(void* arg) {
    OrcWork* w = (OrcWork*)arg;
    if (w->cpu >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(w->cpu, &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    pthread_barrier_wait(w->bar);
    for (;;) {
        int c = __atomic_fetch_add(w->next_chunk, 1, __ATOMIC_RELAXED);
        if (c >= w->n_chunks) break;
        int lo = c * ORC_CHUNK_ENVS, hi = lo + ORC_CHUNK_ENVS;
        if (hi > w->e->n_envs) hi = w->e->n_envs;
        orc_rollout_range(w->e, lo, hi, w->actions, w->n_sets, w->iters, w->tick0, w->obs, w->reward,
                          w->terminated, w->truncated, w->metrics);
    }
    return 0;
}

static double orc_now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

double orc_rollout_threads(OrcEnv* e, const int64_t* actions, int n_sets, int iters, uint64_t tick0,
                           float* obs, double* reward, uint8_t* terminated, uint8_t* truncated,
                           double* metrics, int n_threads, int pin) {
    enum { MAX_T = 256 };
    if (n_threads < 1) n_threads = 1;
    if (n_threads > MAX_T) n_threads = MAX_T;
    pthread_t th[MAX_T];
    OrcWork w[MAX_T];
    int cpus[1024], n_cpus = 0;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (pin && sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < 1024 && c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &allowed)) cpus[n_cpus++] = c;
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, 0, (unsigned)n_threads + 1);
    int next_chunk = 0;
    int n_chunks = (e->n_envs + ORC_CHUNK_ENVS - 1) / ORC_CHUNK_ENVS;
    for (int t = 0; t < n_threads; ++t) {
        w[t].e = e; w[t].actions = actions; w[t].n_sets = n_sets; w[t].iters = iters; w[t].tick0 = tick0;
        w[t].obs = obs; w[t].reward = reward; w[t].terminated = terminated; w[t].truncated = truncated;
        w[t].metrics = metrics + (int64_t)t * ORC_N_METRICS;
        w[t].cpu = n_cpus > 0 ? cpus[t % n_cpus] : -1;
        w[t].bar = &bar; w[t].next_chunk = &next_chunk; w[t].n_chunks = n_chunks;
        pthread_create(&th[t], 0, orc_worker, &w[t]);
    }
    pthread_barrier_wait(&bar);
    double t0 = orc_now();
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
    double dt = orc_now() - t0;
    pthread_barrier_destroy(&bar);
    return dt;
}

int orc_struct_size(void) { return (int)sizeof(OrcEnv); }
