// gte_obs.cu — observation-window gather kernels (sm_100a).  Replaces TradingEnv._get_obs
// (environments.py:152-160): obs[env] = rows idx-W+1..idx of [static features | dynamic features].
//
// Three variants:
//   generic : scalar gather, any shape (windows=None, odd row sizes); a power-of-two lane group per env.
//   vec     : one warp per env; 128-bit ld.global.nc from the 16-byte-aligned window table (rows in the
//             reference's own [t, F] layout, dynamic columns zero), dynamic columns patched in
//             registers from the per-env ring, 128-bit coalesced st.global.  No shared memory.
//   tma     : one warp per env per pipeline stage; cp.async.bulk (TMA, 1-D) pulls the whole window
//             global->shared, lanes patch the dynamic columns in shared memory, cp.async.bulk pushes
//             the finished window shared->global.  The LSU only touches the 8-byte ring entries.
#include <cstdlib>

#include "gte_device.cuh"
#include "gte_launch.h"

namespace gte {

// floor(e / d) for e*d < 2^32 via one umulhi; magic = ceil(2^32 / d)
__device__ __forceinline__ uint32_t fast_div(uint32_t e, uint32_t magic) { return __umulhi(e, magic); }
static inline uint32_t div_magic(uint32_t d) { return (uint32_t)(((1ull << 32) + d - 1) / d); }

struct ObsShape {
    int W, F, ns, nd;            // window rows, row floats, static cols, dynamic cols
    uint32_t magicF;             // ceil(2^32 / F)
    int row_bytes, win_bytes, n_vec;
};

static ObsShape make_shape(const GteParams& P) {
    ObsShape s;
    s.W = P.windows > 0 ? P.windows : 1;
    s.ns = P.n_static;
    s.nd = P.n_dyn;
    s.F = s.ns + s.nd;
    s.magicF = div_magic((uint32_t)s.F);
    s.row_bytes = s.F * 4;
    s.win_bytes = s.W * s.row_bytes;
    s.n_vec = s.win_bytes / 16;
    return s;
}

// ------------------------------------------------------------------------------------------ generic
template <int G>
__global__ void __launch_bounds__(256)
obs_generic_kernel(const GteParams P, const GteData D, const GteState S, float* __restrict__ obs,
                   const ObsShape sh, const int env_begin, const int env_end) {
    const int groups_per_block = 256 / G;
    const int lane = threadIdx.x % G;
    const int64_t n_groups = (int64_t)gridDim.x * groups_per_block;
    const int per_env = sh.W * sh.F;
    for (int64_t env = env_begin + (int64_t)blockIdx.x * groups_per_block + threadIdx.x / G; env < env_end;
         env += n_groups) {
        const int ep_start = S.ep_start[env];
        const int idx = ep_start + S.step[env];
        const int ds = S.dataset_idx[env];
        const int r0 = idx + 1 - sh.W;
        const float* __restrict__ feat = D.features + ((int64_t)ds * P.t_stride + r0) * sh.ns;
        const float* __restrict__ ring = S.dyn_ring + env * sh.W * 2;
        float* __restrict__ out = obs + env * per_env;
        for (int e = lane; e < per_env; e += G) {
            const int w = e / sh.F, c = e - w * sh.F;
            float v;
            if (c < sh.ns) {
                v = __ldg(feat + (int64_t)w * sh.ns + c);
            } else {
                const int r = r0 + w;
                v = (r >= ep_start) ? ring[(r % sh.W) * 2 + (c - sh.ns)] : 0.0f;
            }
            out[e] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float4 ld_nc_v4(const void* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs_v4(void* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// byte address of the 16B-aligned copy of the window that starts at row r0 of dataset ds
__device__ __forceinline__ const char* window_src(const GteData& D, const ObsShape& sh, int ds, int r0) {
    const int64_t off = (int64_t)r0 * sh.row_bytes;
    const int c = (int)((off >> 2) & 3);
    // select without dynamic indexing (keeps the kernel parameters in the constant bank)
    const float* t = (c == 0) ? D.window_table[0] : (c == 1) ? D.window_table[1]
                   : (c == 2) ? D.window_table[2] : D.window_table[3];
    return reinterpret_cast<const char*>(t) + (int64_t)ds * D.window_table_ds_stride + off;
}

// ------------------------------------------------------------------------------------------ vec
template <bool PAIR>     // PAIR: n_dyn == 2 and F even -> the dynamic pair is an aligned float2
__global__ void __launch_bounds__(256)
obs_vec_kernel(const GteParams P, const GteData D, const GteState S, float* __restrict__ obs,
               const ObsShape sh, const int env_begin, const int env_end) {
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t env = env_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); env < env_end;
         env += n_warps) {
        const int ep_start = S.ep_start[env];
        const int idx = ep_start + S.step[env];
        const int ds = S.dataset_idx[env];
        const int r0 = idx + 1 - sh.W;
        const int s0 = ((r0 % sh.W) + sh.W) % sh.W;            // ring slot of window row 0
        const char* __restrict__ src = window_src(D, sh, ds, r0);
        const float2* __restrict__ ring = reinterpret_cast<const float2*>(S.dyn_ring) + env * sh.W;
        char* __restrict__ dst = reinterpret_cast<char*>(obs) + env * (int64_t)sh.win_bytes;
        for (int j = lane; j < sh.n_vec; j += 32) {
            float4 v = ld_nc_v4(src + 16 * j);
            if (sh.nd > 0) {
                float* pv = reinterpret_cast<float*>(&v);
#pragma unroll
                for (int k = 0; k < 4; k += (PAIR ? 2 : 1)) {
                    const uint32_t e = 4u * j + k;
                    const int w = (int)fast_div(e, sh.magicF);
                    const int c = (int)e - w * sh.F;
                    if (c >= sh.ns) {
                        const int r = r0 + w;
                        int slot = s0 + w;
                        if (slot >= sh.W) slot -= sh.W;
                        if (PAIR) {
                            const float2 d = (r >= ep_start) ? __ldg(ring + slot) : make_float2(0.f, 0.f);
                            pv[k] = d.x;
                            pv[k + 1] = d.y;
                        } else {
                            const float* rf = reinterpret_cast<const float*>(ring + slot);
                            pv[k] = (r >= ep_start) ? __ldg(rf + (c - sh.ns)) : 0.0f;
                        }
                    }
                }
            }
            st_cs_v4(dst + 16 * j, v);
        }
    }
}

// ------------------------------------------------------------------------------------------ tma
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA 1-D bulk copy shared -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
// L2 eviction policies (createpolicy): the observation stream is written once and never re-read by this
// path (evict_first), the window tables are re-read by every env (evict_last)
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* gdst, const void* smem_src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ float2 ld_nc_v2_hint(const float2* p, uint64_t policy) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
                 : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct EnvMeta { int ep_start, step, ds; };   // raw loaded words: nothing is computed from them until use,
                                              // so a prefetch never stalls the warp on its own load

__device__ __forceinline__ EnvMeta load_meta(const GteState& S, int64_t env, int n_envs) {
    EnvMeta m;
    m.ep_start = 0; m.step = 0; m.ds = 0;
    if (env < n_envs) {
        m.ep_start = __ldg(S.ep_start + env);
        m.step = __ldg(S.step + env);
        m.ds = __ldg(S.dataset_idx + env);
    }
    return m;
}

// One warp per env per pipeline stage.  STAGES window buffers per warp: the env being finished,
// DEPTH = STAGES-2 windows loading ahead, one buffer draining its store.  RPL = ring entries per
// lane = ceil(W/32).  Per-env metadata and ring entries are prefetched into registers one full
// iteration before they are consumed.
template <int RPL, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
obs_tma_kernel(const GteParams P, const GteData D, const GteState S, float* __restrict__ obs,
               const ObsShape sh, const int env_begin, const int env_end, const int hints) {
    constexpr int DEPTH = STAGES - 2;
    const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[WARPS][STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wbuf = smem_raw + (size_t)warp * STAGES * sh.win_bytes;
    const int N = env_end;               // exclusive upper bound of this launch's env range

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int64_t n_warps = (int64_t)gridDim.x * WARPS;
    const int64_t env0 = env_begin + (int64_t)blockIdx.x * WARPS + warp;

    auto issue_load = [&](int64_t env, const EnvMeta& m, int stage) {   // lane 0 only
        if (env < N) {
            mbar_expect_tx(&bars[warp][stage], (uint32_t)sh.win_bytes);
            const void* src = window_src(D, sh, m.ds, m.ep_start + m.step + 1 - sh.W);
            if (hints & 2) bulk_g2s_hint(wbuf + (size_t)stage * sh.win_bytes, src, (uint32_t)sh.win_bytes, &bars[warp][stage], pol_last);
            else bulk_g2s(wbuf + (size_t)stage * sh.win_bytes, src, (uint32_t)sh.win_bytes, &bars[warp][stage]);
        }
    };
    auto load_ring = [&](int64_t env, float2 (&d)[RPL]) {
        const float2* __restrict__ ring = reinterpret_cast<const float2*>(S.dyn_ring) + env * sh.W;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int s = lane + 32 * q;
            d[q] = make_float2(0.f, 0.f);
            if (sh.nd > 0 && env < N && s < sh.W) d[q] = (hints & 4) ? ld_nc_v2_hint(ring + s, pol_first) : __ldg(ring + s);
        }
    };

    // mq[k] belongs to env + k*n_warps; mq[0..DEPTH] have their window load in flight.
    EnvMeta mq[DEPTH + 2];
#pragma unroll
    for (int k = 0; k < DEPTH + 2; ++k) mq[k] = load_meta(S, env0 + (int64_t)k * n_warps, N);
    EnvMeta m_pend = load_meta(S, env0 + (int64_t)(DEPTH + 2) * n_warps, N);
    float2 d_cur[RPL], d_pend[RPL];
    load_ring(env0, d_cur);
    load_ring(env0 + n_warps, d_pend);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k <= DEPTH; ++k) issue_load(env0 + (int64_t)k * n_warps, mq[k], k);
    }

    int it = 0;
    for (int64_t env = env0; env < N; env += n_warps, ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = (uint32_t)(it / STAGES) & 1u;
        unsigned char* buf = wbuf + (size_t)stage * sh.win_bytes;
        const EnvMeta m = mq[0];
        const int r0 = m.ep_start + m.step + 1 - sh.W;
        const int s0 = ((r0 % sh.W) + sh.W) % sh.W;

        mbar_wait(&bars[warp][stage], parity);           // window (static cols + zero dyn cols) has landed

        if (sh.nd > 0) {
            float* fbuf = reinterpret_cast<float*>(buf);
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                const int s = lane + 32 * q;
                int w = s - s0;
                if (w < 0) w += sh.W;
                if (s < sh.W && r0 + w >= m.ep_start) {
                    float* qd = fbuf + w * sh.F + sh.ns;
                    qd[0] = d_cur[q].x;
                    if (sh.nd > 1) qd[1] = d_cur[q].y;
                }
            }
            fence_proxy_async();                          // generic-proxy smem writes -> visible to TMA
        }
        __syncwarp();
        if (lane == 0) {
            char* dst = reinterpret_cast<char*>(obs) + env * (int64_t)sh.win_bytes;
            if (hints & 1) bulk_s2g_hint(dst, buf, (uint32_t)sh.win_bytes, pol_first);
            else bulk_s2g(dst, buf, (uint32_t)sh.win_bytes);
            bulk_commit();
            // refill the buffer whose store was committed in the PREVIOUS iteration: allow only the
            // store just committed to be still reading shared memory.
            bulk_wait_read<1>();
            issue_load(env + (int64_t)(DEPTH + 1) * n_warps, mq[DEPTH + 1], (it + DEPTH + 1) % STAGES);
        }
        // rotate: consume the prefetches issued one iteration ago, then issue the next ones
#pragma unroll
        for (int k = 0; k < DEPTH + 1; ++k) mq[k] = mq[k + 1];
        mq[DEPTH + 1] = m_pend;
        m_pend = load_meta(S, env + (int64_t)(DEPTH + 3) * n_warps, N);
#pragma unroll
        for (int q = 0; q < RPL; ++q) d_cur[q] = d_pend[q];
        load_ring(env + 2 * n_warps, d_pend);
    }
    if (lane == 0) bulk_wait_read<0>();                   // smem must outlive the last store's reads
}

using ObsKernelFn = void (*)(const GteParams, const GteData, const GteState, float*, const ObsShape, int, int, int);

template <int STAGES, int WARPS>
static ObsKernelFn tma_kernel_for(int rpl) {
    return rpl <= 1 ? obs_tma_kernel<1, STAGES, WARPS>
                    : (rpl <= 2 ? obs_tma_kernel<2, STAGES, WARPS> : obs_tma_kernel<4, STAGES, WARPS>);
}

struct TmaConfig { int stages, warps; };

// default pipeline shape; GTE_TMA_STAGES / GTE_TMA_WARPS override it for tuning runs
static TmaConfig tma_config() {
    static TmaConfig cfg = [] {
        TmaConfig c{3, 8};
        if (const char* e = getenv("GTE_TMA_STAGES")) c.stages = atoi(e);
        if (const char* e = getenv("GTE_TMA_WARPS")) c.warps = atoi(e);
        if (c.stages < 2 || c.stages > 6) c.stages = 3;
        if (c.warps != 4 && c.warps != 8) c.warps = 8;
        return c;
    }();
    return cfg;
}

static ObsKernelFn tma_kernel(const TmaConfig& c, int rpl) {
    if (c.warps == 4) {
        switch (c.stages) {
            case 2: return tma_kernel_for<2, 4>(rpl);
            case 3: return tma_kernel_for<3, 4>(rpl);
            case 5: return tma_kernel_for<5, 4>(rpl);
            case 6: return tma_kernel_for<6, 4>(rpl);
            default: return tma_kernel_for<4, 4>(rpl);
        }
    }
    switch (c.stages) {
        case 2: return tma_kernel_for<2, 8>(rpl);
        case 3: return tma_kernel_for<3, 8>(rpl);
        case 5: return tma_kernel_for<5, 8>(rpl);
        case 6: return tma_kernel_for<6, 8>(rpl);
        default: return tma_kernel_for<4, 8>(rpl);
    }
}

// ------------------------------------------------------------------------------------------ launch
bool obs_vec_supported(const GteParams& P, const GteData& D) {
    const ObsShape sh = make_shape(P);
    if (P.windows <= 0 || sh.win_bytes % 16 != 0) return false;
    if ((int64_t)sh.W * sh.F * (int64_t)sh.F >= (1ll << 31)) return false;     // fast_div range
    for (int r = 0; r < 4; ++r) {                       // every residue class a window start can have
        const int c = (int)((((int64_t)r * sh.row_bytes) >> 2) & 3);
        if (D.window_table[c] == nullptr) return false;
    }
    return D.window_table_ds_stride % 16 == 0;
}

static size_t tma_smem_bytes(const ObsShape& sh) {
    const TmaConfig c = tma_config();
    return (size_t)c.warps * c.stages * sh.win_bytes;
}

bool obs_tma_supported(const GteParams& P, const GteData& D) {
    if (!obs_vec_supported(P, D)) return false;
    const ObsShape sh = make_shape(P);
    return sh.W <= 128 && tma_smem_bytes(sh) <= 200 * 1024;
}

cudaError_t launch_obs_range(const GteParams& P, const GteData& D, const GteState& S, float* obs, int variant,
                             int env_begin, int env_end, cudaStream_t stream) {
    const ObsShape sh = make_shape(P);
    const int n_envs = env_end - env_begin;
    if (n_envs <= 0) return cudaSuccess;
    if (variant == GTE_OBS_AUTO)
        variant = obs_tma_supported(P, D) ? GTE_OBS_TMA : (obs_vec_supported(P, D) ? GTE_OBS_VEC : GTE_OBS_GENERIC);
    const int sms = num_sms();
    if (variant == GTE_OBS_GENERIC) {
        const int per_env = sh.W * sh.F;
        int G = 32;
        while (G > 1 && G / 2 >= per_env) G /= 2;
        const int64_t need = ((int64_t)n_envs * G + 255) / 256;
        const int grid = (int)(need < (int64_t)sms * 16 ? need : (int64_t)sms * 16);
        switch (G) {
            case 32: obs_generic_kernel<32><<<grid, 256, 0, stream>>>(P, D, S, obs, sh, env_begin, env_end); break;
            case 16: obs_generic_kernel<16><<<grid, 256, 0, stream>>>(P, D, S, obs, sh, env_begin, env_end); break;
            case 8: obs_generic_kernel<8><<<grid, 256, 0, stream>>>(P, D, S, obs, sh, env_begin, env_end); break;
            case 4: obs_generic_kernel<4><<<grid, 256, 0, stream>>>(P, D, S, obs, sh, env_begin, env_end); break;
            case 2: obs_generic_kernel<2><<<grid, 256, 0, stream>>>(P, D, S, obs, sh, env_begin, env_end); break;
            default: obs_generic_kernel<1><<<grid, 256, 0, stream>>>(P, D, S, obs, sh, env_begin, env_end); break;
        }
        return cudaGetLastError();
    }
    if (variant == GTE_OBS_VEC) {
        if (!obs_vec_supported(P, D)) return cudaErrorInvalidValue;
        const int64_t need = ((int64_t)n_envs + 7) / 8;
        const int grid = (int)(need < (int64_t)sms * 8 ? need : (int64_t)sms * 8);
        if (sh.nd == 2 && sh.F % 2 == 0) obs_vec_kernel<true><<<grid, 256, 0, stream>>>(P, D, S, obs, sh, env_begin, env_end);
        else obs_vec_kernel<false><<<grid, 256, 0, stream>>>(P, D, S, obs, sh, env_begin, env_end);
        return cudaGetLastError();
    }
    if (variant == GTE_OBS_TMA) {
        if (!obs_tma_supported(P, D)) return cudaErrorInvalidValue;
        const size_t smem = tma_smem_bytes(sh);
        const TmaConfig cfg = tma_config();
        ObsKernelFn kern = tma_kernel(cfg, (sh.W + 31) / 32);
        static ObsKernelFn configured_kern = nullptr;      // opt-in to > 48 KB dynamic smem once per kernel
        static size_t configured_smem = 0;
        if (kern != configured_kern || smem > configured_smem) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured_kern = kern;
            configured_smem = smem;
        }
        int per_sm = (int)((227 * 1024) / (smem + 1024 + 8 * cfg.warps * cfg.stages));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2048 / (cfg.warps * 32)) per_sm = 2048 / (cfg.warps * 32);
        const int64_t need = ((int64_t)n_envs + cfg.warps - 1) / cfg.warps;
        const int grid = (int)(need < (int64_t)sms * per_sm ? need : (int64_t)sms * per_sm);
        static const int hints = [] { const char* e = getenv("GTE_TMA_HINTS"); return e ? atoi(e) : 0; }();
        kern<<<grid, cfg.warps * 32, smem, stream>>>(P, D, S, obs, sh, env_begin, env_end, hints);
        return cudaGetLastError();
    }
    return cudaErrorInvalidValue;
}

}  // namespace gte
