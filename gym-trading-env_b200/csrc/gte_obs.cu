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
    int w_mask;                  // W-1 when W is a power of two, else -1
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
    s.w_mask = (s.W & (s.W - 1)) == 0 ? s.W - 1 : -1;
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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// What the gather needs to know about one env, computed ONCE per env by the lane that owns it in the
// current 32-env block (in parallel across lanes) and then broadcast with shuffles: the per-env loop
// of a warp is a latency chain, so nothing expensive (64-bit address math, modulo) stays inside it.
struct EnvPre {
    unsigned long long src;      // 16-byte-aligned address of the env's window in the window table
    int r0;                      // absolute row of window row 0
    int s0;                      // ring slot of window row 0
    int ep_start;                // rows before it read as zero
};

// One warp per GROUP of G consecutive envs per pipeline stage, everything bulky moved by the TMA engine.
// The TMA unit of an SM serves bulk operations at a fixed cost per OPERATION (~50 cycles measured: three
// operations per env capped the kernel at ~1.0 ms regardless of bytes), so operations are merged wherever
// the bytes are contiguous: consecutive envs have contiguous rings and contiguous output windows.
//   * per group, lane 0 issues onto the stage's mbarrier: G cp.async.bulk window loads (one per env, from
//     the L2-resident window table) and ONE cp.async.bulk for the G rings (G*W*8 contiguous bytes, HBM);
//   * when the barrier flips, lanes patch the dynamic columns smem->smem (rows before the episode start
//     stay zero), fence.proxy.async once, and lane 0 issues ONE cp.async.bulk shared->global for the G
//     finished windows (G*W*F*4 contiguous bytes);
//   * STAGES group buffers per warp: the group being finished, DEPTH = STAGES-2 groups loading ahead, one
//     store draining.  No register-staged global loads in the loop: a warp works through 32 CONSECUTIVE
//     envs, whose (ep_start, step, dataset) words are fetched with one coalesced load per array a whole
//     32-env block ahead and turned into addresses / ring slots once per env by the owning lane.
template <int STAGES, int WARPS, int G>
__global__ void __launch_bounds__(WARPS * 32)
obs_tma_kernel(const GteParams P, const GteData D, const GteState S, float* __restrict__ obs,
               const ObsShape sh, const int env_begin, const int env_end) {
    constexpr int DEPTH = STAGES - 2;
    constexpr int GROUPS = 32 / G;                               // groups per 32-env block
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[WARPS][STAGES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t win_bytes = (uint32_t)sh.win_bytes;
    const uint32_t ring_bytes = sh.nd > 0 ? (uint32_t)sh.W * 8u : 0u;
    const uint32_t stage_bytes = (win_bytes + ring_bytes) * G;   // [G windows][G rings]
    unsigned char* wbuf = smem_raw + (size_t)warp * STAGES * stage_bytes;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bars[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const int64_t n_warps = (int64_t)gridDim.x * WARPS;
    const int64_t warp_global = (int64_t)blockIdx.x * WARPS + warp;
    // first env of block k of this warp (blocks of 32 consecutive envs, strided over the grid)
    auto block_env0 = [&](int k) -> int64_t { return env_begin + ((warp_global + (int64_t)k * n_warps) << 5); };
    auto load_block = [&](int k) {
        EnvPre p;
        p.src = 0ull; p.r0 = 0; p.s0 = 0; p.ep_start = 0;
        const int64_t env = block_env0(k) + lane;
        if (env < env_end) {
            const int ep = __ldg(S.ep_start + env), st = __ldg(S.step + env), ds = __ldg(S.dataset_idx + env);
            p.ep_start = ep;
            p.r0 = ep + st + 1 - sh.W;
            p.s0 = sh.w_mask >= 0 ? (p.r0 & sh.w_mask) : (((p.r0 % sh.W) + sh.W) % sh.W);
            p.src = (unsigned long long)window_src(D, sh, ds, p.r0);
        }
        return p;
    };

    EnvPre pc = load_block(0), pn = load_block(1);               // current / next 32-env block (one env per lane)

    // all lanes call (shuffles); lane 0 issues the loads of group gi of block ki; cur_k = block of pc
    auto issue = [&](int ki, int gi, int cur_k, int stage) {
        const int64_t env0 = block_env0(ki) + gi * G;
        const int64_t left = (int64_t)env_end - env0;
        const int n_valid = left >= G ? G : (left > 0 ? (int)left : 0);
        unsigned char* sbuf = wbuf + (size_t)stage * stage_bytes;
        uint64_t* bar = &bars[warp][stage];
        if (lane == 0 && n_valid > 0) {
            mbar_expect_tx(bar, (uint32_t)n_valid * (win_bytes + ring_bytes));
            if (ring_bytes)                                       // the G rings are contiguous: one operation
                bulk_g2s(sbuf + (size_t)G * win_bytes, S.dyn_ring + env0 * (int64_t)sh.W * 2,
                         (uint32_t)n_valid * ring_bytes, bar);
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const unsigned long long src = __shfl_sync(FULL, ki != cur_k ? pn.src : pc.src, gi * G + g);
            if (lane == 0 && g < n_valid) bulk_g2s(sbuf + (size_t)g * win_bytes, reinterpret_cast<const void*>(src), win_bytes, bar);
        }
    };

#pragma unroll
    for (int q0 = 0; q0 <= DEPTH; ++q0) issue(q0 / GROUPS, q0 % GROUPS, 0, q0);

    int q = 0;                                                   // groups finished by this warp so far
    for (int k = 0; block_env0(k) < env_end; ++k) {
        const int64_t env_k = block_env0(k);
        for (int gi = 0; gi < GROUPS; ++gi, ++q) {
            const int64_t env0 = env_k + gi * G;
            const int64_t left = (int64_t)env_end - env0;
            if (left <= 0) break;                                // ragged tail of the last block (warp-uniform)
            const int n_valid = left >= G ? G : (int)left;
            const int stage = q % STAGES;
            const uint32_t parity = (uint32_t)(q / STAGES) & 1u;
            unsigned char* sbuf = wbuf + (size_t)stage * stage_bytes;

            mbar_wait(&bars[warp][stage], parity);               // the group's windows + rings have landed

            if (ring_bytes) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const int r0 = __shfl_sync(FULL, pc.r0, gi * G + g);
                    const int s0 = __shfl_sync(FULL, pc.s0, gi * G + g);
                    const int ep_start = __shfl_sync(FULL, pc.ep_start, gi * G + g);
                    if (g < n_valid) {
                        float* fbuf = reinterpret_cast<float*>(sbuf + (size_t)g * win_bytes);
                        const float2* rbuf = reinterpret_cast<const float2*>(sbuf + (size_t)G * win_bytes + (size_t)g * ring_bytes);
                        for (int s = lane; s < sh.W; s += 32) {
                            int w = s - s0;
                            if (w < 0) w += sh.W;
                            if (r0 + w >= ep_start) {            // rows before the episode start stay zero
                                const float2 d = rbuf[s];
                                float* qd = fbuf + w * sh.F + sh.ns;
                                qd[0] = d.x;
                                qd[1] = d.y;
                            }
                        }
                    }
                }
                fence_proxy_async();                             // generic-proxy smem writes -> visible to TMA
            }
            __syncwarp();
            if (lane == 0) {
                // the G finished windows are contiguous in smem and in the output: one operation
                bulk_s2g(reinterpret_cast<char*>(obs) + env0 * (int64_t)win_bytes, sbuf, (uint32_t)n_valid * win_bytes);
                bulk_commit();
                // the stage refilled below is the one whose store was committed in the PREVIOUS iteration:
                // allow only the store just committed to be still reading shared memory
                bulk_wait_read<1>();
            }
            __syncwarp();
            const int qi = gi + DEPTH + 1;                       // group to start loading now
            issue(k + qi / GROUPS, qi % GROUPS, k, (q + DEPTH + 1) % STAGES);
        }
        pc = pn;                                                 // loaded a whole block ago: no stall
        pn = load_block(k + 2);
    }
    if (lane == 0) bulk_wait_read<0>();                          // smem must outlive the last store's reads
}

using ObsKernelFn = void (*)(const GteParams, const GteData, const GteState, float*, const ObsShape, int, int);

struct TmaConfig { int stages, warps, group; };

// default pipeline shape; GTE_TMA_STAGES / GTE_TMA_WARPS / GTE_TMA_GROUP override it for tuning runs
static TmaConfig tma_config() {
    static TmaConfig cfg = [] {
        TmaConfig c{3, 2, 4};
        if (const char* e = getenv("GTE_TMA_STAGES")) c.stages = atoi(e);
        if (const char* e = getenv("GTE_TMA_WARPS")) c.warps = atoi(e);
        if (const char* e = getenv("GTE_TMA_GROUP")) c.group = atoi(e);
        if (c.stages != 3 && c.stages != 4 && c.stages != 6) c.stages = 3;
        if (c.warps != 1 && c.warps != 2 && c.warps != 4 && c.warps != 8) c.warps = 2;
        if (c.group != 1 && c.group != 2 && c.group != 4 && c.group != 8) c.group = 4;
        return c;
    }();
    return cfg;
}

template <int STAGES, int WARPS>
static ObsKernelFn tma_kernel_g(int g) {
    switch (g) {
        case 1: return obs_tma_kernel<STAGES, WARPS, 1>;
        case 2: return obs_tma_kernel<STAGES, WARPS, 2>;
        case 8: return obs_tma_kernel<STAGES, WARPS, 8>;
        default: return obs_tma_kernel<STAGES, WARPS, 4>;
    }
}
template <int STAGES>
static ObsKernelFn tma_kernel_w(int w, int g) {
    switch (w) {
        case 1: return tma_kernel_g<STAGES, 1>(g);
        case 4: return tma_kernel_g<STAGES, 4>(g);
        case 8: return tma_kernel_g<STAGES, 8>(g);
        default: return tma_kernel_g<STAGES, 2>(g);
    }
}
static ObsKernelFn tma_kernel(const TmaConfig& c) {
    return c.stages == 3 ? tma_kernel_w<3>(c.warps, c.group)
                         : (c.stages == 4 ? tma_kernel_w<4>(c.warps, c.group) : tma_kernel_w<6>(c.warps, c.group));
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
    return (size_t)c.warps * c.stages * c.group * ((size_t)sh.win_bytes + (sh.nd > 0 ? (size_t)sh.W * 8 : 0));
}

bool obs_tma_supported(const GteParams& P, const GteData& D) {
    if (!obs_vec_supported(P, D)) return false;
    const ObsShape sh = make_shape(P);
    // the ring of one env (W*8 bytes) must itself be a 16-byte multiple for cp.async.bulk
    return (sh.nd == 0 || (sh.W * 8) % 16 == 0) && (sh.nd == 0 || sh.nd == 2) && tma_smem_bytes(sh) <= 200 * 1024;
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
        ObsKernelFn kern = tma_kernel(cfg);
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
        if (per_sm > 32) per_sm = 32;
        const int64_t need = (((int64_t)n_envs + 31) / 32 + cfg.warps - 1) / cfg.warps;   // one 32-env block per warp at least
        const int grid = (int)(need < (int64_t)sms * per_sm ? need : (int64_t)sms * per_sm);
        kern<<<grid, cfg.warps * 32, smem, stream>>>(P, D, S, obs, sh, env_begin, env_end);
        return cudaGetLastError();
    }
    return cudaErrorInvalidValue;
}

}  // namespace gte
