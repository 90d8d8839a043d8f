// gte_step.cu — fused per-env transition kernel, reset kernel and info kernel (sm_100a).
//
// One thread per env, structure-of-arrays state in HBM (coalesced 8/4-byte accesses), one 256-env
// tile per CTA (a few consecutive tiles when N is huge, so the per-CTA metric partials stay bounded).
// Replaces TradingEnv.step (environments.py:233-272) for N envs in lockstep; see
// include/gte_b200.h for the boundary and DESIGN.md for the data layout / roofline.
#include <cstdlib>

#include "gte_device.cuh"
#include "gte_launch.h"

namespace gte {

constexpr int kChunkFirst = 1, kChunkLast = 2;

struct StepThreadOut {       // what the gather phase of the fused kernel needs from the step phase
    int idx, ep_start, ds;
};

struct MetricAcc {           // per-thread accumulators (deterministic: fixed tile order per thread)
    double sum_pr = 0.0, sum_mr = 0.0, sum_rew = 0.0;
    int episodes = 0, terminated = 0, truncated = 0, sum_len = 0;
};

// The transition of ONE env (the reference's step(), line by line).
__device__ __forceinline__ StepThreadOut step_env(const GteParams& P, const GteData& D, const GteState& S,
                                                  const int64_t* __restrict__ actions,
                                                  const GteStepOut& O, uint64_t tick, int autoreset,
                                                  int i, MetricAcc& acc) {
    EnvRegs e;
    e.pf.asset = S.asset[i];
    e.pf.fiat = S.fiat[i];
    e.pf.ia = S.interest_asset[i];
    e.pf.ifi = S.interest_fiat[i];
    e.pos_idx = S.pos_idx[i];
    e.step = S.step[i];
    e.ep_start = S.ep_start[i];
    e.ds = (P.n_datasets > 1) ? S.dataset_idx[i] : 0;
    int64_t a = actions[i];

    const double* __restrict__ price = D.price + (int64_t)e.ds * P.t_stride;
    const int T = D.lengths[e.ds];
    int idx = e.ep_start + e.step;
    if (idx + 1 >= T) {              // stepping past the end of the data without a reset (caller bug)
        atomicOr(S.error_flag, 2);
        idx = T - 2;
    }
    if (a >= (int64_t)P.n_positions) { atomicOr(S.error_flag, 1); a = -1; }

    const double p0 = __ldg(price + idx);                                    // price BEFORE advancing (:204-207)
    const double p1 = __ldg(price + idx + 1);
    // history["portfolio_valuation", -2]: the previous row's valuation.  The state on entry is exactly
    // the state that row was valued with, at the same price, so it is recomputed bit-identically
    // instead of being stored; the first row of an episode holds portfolio_initial_value (:194).
    const double val0 = valorisation(e.pf, p0);
    const double prev_val = (e.step == 0) ? P.v0 : val0;

    if (a >= 0) {                                                            // :234 (None = hold)
        const double target = P.positions[a];
        if (target != P.positions[e.pos_idx]) {                              // :213-215 value compare
            trade_to_position(e.pf, target, p0, P.fee, val0);                // :204-211
            e.pos_idx = (int)a;
        }
    }
    idx += 1;                                                                // :235
    e.step += 1;                                                             // :236
    update_interest(e.pf, P.rate);                                           // :240
    const double val = valorisation(e.pf, p1);                               // :241
    const bool done = ddiv(val, P.v0) <= P.done_ratio;                       // :246
    bool trunc = idx >= T - 1;                                               // :248
    if (P.max_episode_duration >= 0 && e.step >= P.max_episode_duration - 1) trunc = true;   // :250
    const double rp = real_position(e.pf, p1, val);                          // :259
    const double rew = done ? 0.0 : log(ddiv(val, prev_val));                // :263-267 -> :17-18

    O.reward[i] = rew;
    O.terminated[i] = (uint8_t)done;
    O.truncated[i] = (uint8_t)trunc;
    if (O.valuation) O.valuation[i] = val;
    if (O.real_position) O.real_position[i] = rp;
    if (O.info_idx) O.info_idx[i] = idx;
    if (O.info_step) O.info_step[i] = e.step;
    if (O.pre_reset_portfolio) {
        const int64_t N = P.n_envs;
        O.pre_reset_portfolio[i] = e.pf.asset;
        O.pre_reset_portfolio[N + i] = e.pf.fiat;
        O.pre_reset_portfolio[2 * N + i] = e.pf.ia;
        O.pre_reset_portfolio[3 * N + i] = e.pf.ifi;
    }
    if (P.n_dyn > 0) {                                                       // _get_obs write-back (:153-154)
        const int W = P.windows > 0 ? P.windows : 1;
        float2* ring = reinterpret_cast<float2*>(S.dyn_ring) + (int64_t)i * W + (idx % W);
        *ring = make_float2((float)P.positions[e.pos_idx], (float)rp);
    }
    acc.sum_rew = dadd(acc.sum_rew, rew);
    if (done || trunc) {                                                     // :269-271 calculate_metrics
        acc.episodes += 1;
        acc.terminated += done ? 1 : 0;
        acc.truncated += trunc ? 1 : 0;
        acc.sum_len += e.step;
        acc.sum_pr = dadd(acc.sum_pr, dsub(ddiv(val, P.v0), 1.0));           // :282
        acc.sum_mr = dadd(acc.sum_mr, dsub(ddiv(p1, __ldg(price + e.ep_start)), 1.0));   // :281
        if (autoreset) {
            reset_env(P, D, S, i, tick, e);                                  // in-place auto-reset
            idx = e.ep_start;
            if (P.n_datasets > 1) S.dataset_idx[i] = e.ds;
        }
    }
    S.asset[i] = e.pf.asset;
    S.fiat[i] = e.pf.fiat;
    S.interest_asset[i] = e.pf.ia;
    S.interest_fiat[i] = e.pf.ifi;
    S.pos_idx[i] = e.pos_idx;
    S.step[i] = e.step;
    S.ep_start[i] = e.ep_start;
    StepThreadOut r;
    r.idx = idx; r.ep_start = e.ep_start; r.ds = e.ds;
    return r;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dadd(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// CTA-level metric reduction -> metric_partials[blockIdx.x]; the last CTA to arrive folds all
// partial rows in a fixed order (deterministic) into metrics_step / metrics_total.
__device__ void reduce_metrics(const MetricAcc& acc, const GteStepOut& O, const GteState& S, int chunk_flags) {
    __shared__ double s_part[kStepThreads / 32][GTE_N_METRICS];
    __shared__ double s_fold[32][GTE_N_METRICS];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const double w_rew = warp_sum(acc.sum_rew);
    const int w_ep = __reduce_add_sync(0xffffffffu, acc.episodes);
    double w_pr = 0.0, w_mr = 0.0;
    int w_term = 0, w_trunc = 0, w_len = 0;
    if (w_ep > 0) {                                      // warp-uniform: episode ends are rare
        w_pr = warp_sum(acc.sum_pr);
        w_mr = warp_sum(acc.sum_mr);
        w_term = __reduce_add_sync(0xffffffffu, acc.terminated);
        w_trunc = __reduce_add_sync(0xffffffffu, acc.truncated);
        w_len = __reduce_add_sync(0xffffffffu, acc.sum_len);
    }
    if (lane == 0) {
        s_part[warp][GTE_M_EPISODES] = (double)w_ep;
        s_part[warp][GTE_M_TERMINATED] = (double)w_term;
        s_part[warp][GTE_M_TRUNCATED] = (double)w_trunc;
        s_part[warp][GTE_M_SUM_PORTFOLIO_RETURN] = w_pr;
        s_part[warp][GTE_M_SUM_MARKET_RETURN] = w_mr;
        s_part[warp][GTE_M_SUM_EPISODE_LENGTH] = (double)w_len;
        s_part[warp][GTE_M_SUM_REWARD] = w_rew;
        s_part[warp][GTE_M_RESERVED] = 0.0;
    }
    __syncthreads();
    if (threadIdx.x < GTE_N_METRICS) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kStepThreads / 32; ++w) t = dadd(t, s_part[w][threadIdx.x]);
        O.metric_partials[(int64_t)blockIdx.x * GTE_N_METRICS + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(O.block_counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int m = threadIdx.x % GTE_N_METRICS, g = threadIdx.x / GTE_N_METRICS;   // 32 groups x 8 metrics
    double t = 0.0;
    for (int b = g; b < (int)gridDim.x; b += kStepThreads / GTE_N_METRICS)
        t = dadd(t, __ldcg(O.metric_partials + (int64_t)b * GTE_N_METRICS + m));
    s_fold[g][m] = t;
    __syncthreads();
    if (threadIdx.x < GTE_N_METRICS) {
        double tot = 0.0;
        for (int k = 0; k < kStepThreads / GTE_N_METRICS; ++k) tot = dadd(tot, s_fold[k][threadIdx.x]);
        // env chunks of one lockstep iteration run as consecutive launches: the first one overwrites
        // metrics_step, later ones add to it in launch order (deterministic)
        O.metrics_step[threadIdx.x] = (chunk_flags & kChunkFirst) ? tot : dadd(O.metrics_step[threadIdx.x], tot);
        if (O.metrics_total) O.metrics_total[threadIdx.x] = dadd(O.metrics_total[threadIdx.x], tot);
    }
    if (threadIdx.x == 0) {
        *O.block_counter = 0u;                           // self-resetting for the next launch
        // every CTA of this iteration has read the tick by now: advance the Philox event counter
        if (chunk_flags & kChunkLast) *S.tick = *S.tick + 1ull;
    }
}

// Each CTA owns `tiles_per_cta` CONSECUTIVE 256-env tiles (1 unless N > 256 x kMaxPartialRows), so the
// grid is as wide as the problem (latency hiding comes from CTA-level parallelism, not a loop) while
// the metric partials stay bounded.
template <int MIN_CTAS>
__global__ void __launch_bounds__(kStepThreads, MIN_CTAS)
step_kernel(const GteParams P, const GteData D, const GteState S, const int64_t* __restrict__ actions,
            const GteStepOut O, int autoreset, int tiles_per_cta, int env_begin, int env_end, int chunk_flags) {
    MetricAcc acc;
    const uint64_t tick = *S.tick;
    const int64_t base0 = (int64_t)env_begin + (int64_t)blockIdx.x * tiles_per_cta * kStepThreads;
    for (int t = 0; t < tiles_per_cta; ++t) {
        const int64_t i = base0 + (int64_t)t * kStepThreads + threadIdx.x;
        if (i < env_end) step_env(P, D, S, actions, O, tick, autoreset, (int)i, acc);
    }
    reduce_metrics(acc, O, S, chunk_flags);
}

// TradingEnv.reset for the masked envs (environments.py:163-199, :393-400); `first` also performs
// the dataset draw of MultiDatasetTradingEnv.__init__ (:377-378).
__global__ void __launch_bounds__(kStepThreads)
reset_kernel(const GteParams P, const GteData D, const GteState S, const uint8_t* __restrict__ mask,
             int first) {
    const int i = blockIdx.x * kStepThreads + threadIdx.x;
    if (i >= P.n_envs) return;
    const uint64_t tick = *S.tick;
    if (first) {
        S.ds_used[i] = 0;
        S.ds_episodes[i] = 0;
        S.plan_cursor[i] = 0;
        int ds = 0;
        if (P.multi_dataset && !(P.plan_episodes > 0 && S.reset_plan != nullptr)) {
            uint32_t r[4];
            philox_draw(P.seed, tick, (uint64_t)(P.env_id_offset + i), r);
            ds = next_dataset(P, S, i, r[3]);
        }
        S.dataset_idx[i] = ds;
        return;
    }
    if (mask != nullptr && mask[i] == 0) return;
    EnvRegs e;
    e.ds = S.dataset_idx[i];
    reset_env(P, D, S, i, tick, e);
    S.asset[i] = e.pf.asset;
    S.fiat[i] = e.pf.fiat;
    S.interest_asset[i] = e.pf.ia;
    S.interest_fiat[i] = e.pf.ifi;
    S.pos_idx[i] = e.pos_idx;
    S.step[i] = e.step;
    S.ep_start[i] = e.ep_start;
    S.dataset_idx[i] = e.ds;
}

__global__ void bump_tick_kernel(uint64_t* tick) { *tick = *tick + 1ull; }

// History's last row from the current state (environments.py:253-264, portfolio.py:49-57).
__global__ void __launch_bounds__(kStepThreads)
info_kernel(const GteParams P, const GteData D, const GteState S, const GteInfo I) {
    const int i = blockIdx.x * kStepThreads + threadIdx.x;
    if (i >= P.n_envs) return;
    Portfolio pf;
    pf.asset = S.asset[i]; pf.fiat = S.fiat[i]; pf.ia = S.interest_asset[i]; pf.ifi = S.interest_fiat[i];
    const int step = S.step[i], ep_start = S.ep_start[i], ds = S.dataset_idx[i], pidx = S.pos_idx[i];
    const int idx = ep_start + step;
    const double price = D.price[(int64_t)ds * P.t_stride + idx];
    const double position = P.positions[pidx];
    // the reset row holds portfolio_initial_value and real_position = position (:192-194)
    const double val = (step == 0) ? P.v0 : valorisation(pf, price);
    const double rp = (step == 0) ? position : real_position(pf, price, val);
    if (I.idx) I.idx[i] = idx;
    if (I.step) I.step[i] = step;
    if (I.position_index) I.position_index[i] = pidx;
    if (I.dataset_idx) I.dataset_idx[i] = ds;
    if (I.position) I.position[i] = position;
    if (I.real_position) I.real_position[i] = rp;
    if (I.portfolio_valuation) I.portfolio_valuation[i] = val;
    if (I.data_close) I.data_close[i] = price;
    if (I.distribution) {                                                    // portfolio.py:49-57
        const int64_t N = P.n_envs;
        const double na = -pf.asset, nf = -pf.fiat;
        I.distribution[i] = (pf.asset > 0.0) ? pf.asset : 0.0;
        I.distribution[N + i] = (pf.fiat > 0.0) ? pf.fiat : 0.0;
        I.distribution[2 * N + i] = (na > 0.0) ? na : 0.0;
        I.distribution[3 * N + i] = (nf > 0.0) ? nf : 0.0;
        I.distribution[4 * N + i] = pf.ia;
        I.distribution[5 * N + i] = pf.ifi;
    }
}

// ---- host launchers -------------------------------------------------------------------------------

static int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

static int step_tiles_per_cta(int n_envs) {
    const int64_t tiles = ((int64_t)n_envs + kStepThreads - 1) / kStepThreads;
    return (int)((tiles + kMaxPartialRows - 1) / kMaxPartialRows);
}

int step_grid(int n_envs) {
    const int64_t tiles = ((int64_t)n_envs + kStepThreads - 1) / kStepThreads;
    const int tpc = step_tiles_per_cta(n_envs);
    return (int)((tiles + tpc - 1) / tpc);
}

cudaError_t launch_step_range(const GteParams& P, const GteData& D, const GteState& S, const int64_t* actions,
                              const GteStepOut& O, int autoreset, int env_begin, int env_end, int chunk_flags,
                              cudaStream_t stream) {
    static const int min_ctas = [] { const char* e = getenv("GTE_STEP_MIN_CTAS"); return e ? atoi(e) : 4; }();
    const int n = env_end - env_begin;
    const int grid = step_grid(n), tpc = step_tiles_per_cta(n);
    // same shared-memory carve-out as the gather kernel, so CTAs of both can be resident on one SM
    static const bool carveout_set = [] {
        cudaFuncSetAttribute(step_kernel<3>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(step_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return true;
    }();
    (void)carveout_set;
    if (min_ctas >= 4)
        step_kernel<4><<<grid, kStepThreads, 0, stream>>>(P, D, S, actions, O, autoreset, tpc, env_begin, env_end, chunk_flags);
    else
        step_kernel<3><<<grid, kStepThreads, 0, stream>>>(P, D, S, actions, O, autoreset, tpc, env_begin, env_end, chunk_flags);
    return cudaGetLastError();
}

cudaError_t launch_step(const GteParams& P, const GteData& D, const GteState& S, const int64_t* actions,
                        const GteStepOut& O, int autoreset, cudaStream_t stream) {
    return launch_step_range(P, D, S, actions, O, autoreset, 0, P.n_envs, kChunkFirst | kChunkLast, stream);
}

cudaError_t launch_reset(const GteParams& P, const GteData& D, const GteState& S, const uint8_t* mask,
                         int first, cudaStream_t stream) {
    const int grid = (P.n_envs + kStepThreads - 1) / kStepThreads;
    reset_kernel<<<grid, kStepThreads, 0, stream>>>(P, D, S, mask, first);
    bump_tick_kernel<<<1, 1, 0, stream>>>(S.tick);
    return cudaGetLastError();
}

// ---- one lockstep iteration = step + gather, env chunks pipelined over two streams ---------------
// The step kernel is latency-bound (dependent loads + fp64 divide/log chains), the gather is
// HBM-bound; running the step of chunk c+1 beside the gather of chunk c hides the former.
struct AuxStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork[16] = {};
    cudaEvent_t join = nullptr;
    int device = -1;
};
static AuxStream g_aux[16];

static cudaError_t aux_for_current_device(AuxStream** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    AuxStream& a = g_aux[dev & 15];
    if (a.stream == nullptr) {
        if ((e = cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
        for (auto& ev : a.fork)
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming)) != cudaSuccess) return e;
        a.device = dev;
    }
    *out = &a;
    return cudaSuccess;
}

int default_chunks(int n_envs) {
    // Measured on B200 (profiles/r01_tuning.md): running the step kernel of range c+1 beside the gather of
    // range c does not beat two plain launches — the gather already saturates HBM writes and the extra
    // launches add tails — so pipelining is opt-in (n_chunks > 1 or GTE_CHUNKS).
    static const int forced = [] { const char* e = getenv("GTE_CHUNKS"); return e ? atoi(e) : 0; }();
    (void)n_envs;
    if (forced > 0) return forced > 16 ? 16 : forced;
    return 1;
}

cudaError_t launch_step_obs(const GteParams& P, const GteData& D, const GteState& S, const int64_t* actions,
                            const GteStepOut& O, float* obs, int autoreset, int variant, int n_chunks,
                            cudaStream_t stream) {
    if (n_chunks <= 0) n_chunks = default_chunks(P.n_envs);
    if (n_chunks > 16) n_chunks = 16;
    cudaError_t e;
    if (n_chunks == 1) {
        if ((e = launch_step(P, D, S, actions, O, autoreset, stream)) != cudaSuccess) return e;
        return launch_obs_range(P, D, S, obs, variant, 0, P.n_envs, stream);
    }
    AuxStream* aux = nullptr;
    if ((e = aux_for_current_device(&aux)) != cudaSuccess) return e;
    const int64_t per = (((int64_t)P.n_envs + n_chunks - 1) / n_chunks + kStepThreads - 1) / kStepThreads * kStepThreads;
    n_chunks = (int)(((int64_t)P.n_envs + per - 1) / per);         // drop empty tail chunks
    for (int c = 0; c < n_chunks; ++c) {
        const int b = (int)(c * per);
        const int en = (int)((c + 1) * per < P.n_envs ? (c + 1) * per : P.n_envs);
        const int flags = (c == 0 ? kChunkFirst : 0) | (c == n_chunks - 1 ? kChunkLast : 0);
        if ((e = launch_step_range(P, D, S, actions, O, autoreset, b, en, flags, stream)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(aux->fork[c], stream)) != cudaSuccess) return e;
        if ((e = cudaStreamWaitEvent(aux->stream, aux->fork[c], 0)) != cudaSuccess) return e;
        if ((e = launch_obs_range(P, D, S, obs, variant, b, en, aux->stream)) != cudaSuccess) return e;
    }
    if ((e = cudaEventRecord(aux->join, aux->stream)) != cudaSuccess) return e;
    return cudaStreamWaitEvent(stream, aux->join, 0);
}

cudaError_t launch_info(const GteParams& P, const GteData& D, const GteState& S, const GteInfo& I,
                        cudaStream_t stream) {
    const int grid = (P.n_envs + kStepThreads - 1) / kStepThreads;
    info_kernel<<<grid, kStepThreads, 0, stream>>>(P, D, S, I);
    return cudaGetLastError();
}

}  // namespace gte
