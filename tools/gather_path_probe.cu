// gather_path_probe.cu — what bounds the C5 gather (2^21 windows of 2560 B assembled from an L2-resident table)?
// A stripped pipeline (producer warp: TMA loads of each env's window from a 4 MB table at pseudo-random rows; 4 consumer
// warps; output by TMA bulk store or by LSU stores), tiles of 32 envs claimed from a counter.  Variants differ in the bytes
// LOADED per env (2560 = the full row with dynamic placeholders, 2048 = static columns only, 0 = none) and in the store path.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gather_path_probe tools/gather_path_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int WS = 3, G = 4, WIN = 2560, TILE = 32, GROUPS = TILE / G;

// STORE: 0 = one TMA bulk store per group (10 KB) by a consumer thread; 1 = LSU: 128 threads LDS.64 + STG.64
template <int STORE>
__global__ void __launch_bounds__(160) gather_probe(const char* __restrict__ table, size_t table_rows, int row_bytes, int load_bytes,
                                                    char* __restrict__ out, int n_tiles, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[WS], empty[WS];
    __shared__ int tile_s[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < WS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], STORE == 0 ? 1 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 4) {
        int t = blockIdx.x, q = 0;
        for (int k = 0;; ++k) {
            const bool last = t >= n_tiles;
            for (int gi = 0; gi < GROUPS; ++gi, ++q) {
                const int stage = q % WS, use = q / WS;
                if (use > 0) mbar_wait(&empty[stage], (uint32_t)(use - 1) & 1u);
                if (gi == 0 && lane == 0) tile_s[k & 3] = last ? -1 : t;
                __syncwarp();
                if (last) { if (lane == 0) mbar_arrive(&full[stage]); break; }
                if (lane == 0) {
                    if (load_bytes > 0) {
                        mbar_expect_tx(&full[stage], (uint32_t)(G * load_bytes));
                        for (int g = 0; g < G; ++g) {
                            const unsigned env = (unsigned)t * TILE + gi * G + g;
                            const size_t row = (((size_t)env * 2654435761u) % (table_rows - 64)) & ~(size_t)1;   // even rows: 16-byte aligned for both row sizes
                            bulk_g2s(smem + (size_t)stage * G * WIN + (size_t)g * WIN, table + row * (size_t)row_bytes, (uint32_t)load_bytes, &full[stage]);
                        }
                    } else {
                        mbar_arrive(&full[stage]);
                    }
                }
            }
            if (last) break;
            t = __shfl_sync(0xffffffffu, lane == 0 ? (int)(gridDim.x + atomicAdd(counter, 1u)) : 0, 0);
        }
    } else {
        int q = 0;
        size_t e0 = 0;
        for (int k = 0;; ++k) {
            bool stop = false;
            for (int gi = 0; gi < GROUPS; ++gi, ++q) {
                const int stage = q % WS;
                mbar_wait(&full[stage], (uint32_t)(q / WS) & 1u);
                if (gi == 0) {
                    const int t = *reinterpret_cast<volatile int*>(&tile_s[k & 3]);
                    if (t < 0) { stop = true; break; }
                    e0 = (size_t)t * TILE;
                }
                char* dst = out + (e0 + gi * G) * (size_t)WIN;
                unsigned char* src = smem + (size_t)stage * G * WIN;
                if (STORE == 0) {
                    // touch the stage like the patch does (2 floats per row), then one bulk store by one thread
                    for (int j = tid; j < G * 64; j += 128) reinterpret_cast<float*>(src)[(j / 64) * (WIN / 4) + (j % 64) * 10 + 8] = (float)j;
                    fence_proxy_async();
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    if (tid == 0) {
                        bulk_s2g(dst, src, G * WIN);
                        bulk_commit();
                        bulk_wait_read<0>();
                        mbar_arrive(&empty[stage]);
                    }
                } else {
                    const double* s8 = reinterpret_cast<const double*>(src);
                    double* d8 = reinterpret_cast<double*>(dst);
                    for (int j = tid; j < G * WIN / 8; j += 128) d8[j] = s8[j];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                }
            }
            if (stop) break;
        }
    }
}

template <typename F>
static double timed_ms(F launch, int reps = 10) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main() {
    const size_t n_env = 1u << 21;
    const size_t bytes = n_env * WIN;
    char *out = nullptr, *table = nullptr;
    unsigned int* counter = nullptr;
    const size_t table_rows = 100000;
    CK(cudaMalloc(&out, bytes));
    CK(cudaMalloc(&table, table_rows * 40 + 4096));
    CK(cudaMemset(table, 1, table_rows * 40 + 4096));
    CK(cudaMalloc(&counter, 4));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int smem = WS * G * WIN;
    CK(cudaFuncSetAttribute(gather_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(gather_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int n_tiles = (int)(n_env / TILE);
    auto report = [&](const char* name, double ms) { printf("{\"variant\": \"%s\", \"ms\": %.4f, \"out_TBps\": %.3f}\n", name, ms, bytes / ms / 1e9); fflush(stdout); };
    struct V { const char* name; int store, row_bytes, load_bytes; };
    const V vs[] = {
        {"load 2560 B/env (40 B rows) + TMA store", 0, 40, 2560},
        {"load 2048 B/env (32 B rows) + TMA store", 0, 32, 2048},
        {"load 1280 B/env + TMA store", 0, 32, 1280},
        {"no load + TMA store", 0, 32, 0},
        {"load 2560 B/env (40 B rows) + LSU store", 1, 40, 2560},
        {"load 2048 B/env (32 B rows) + LSU store", 1, 32, 2048},
        {"no load + LSU store", 1, 32, 0},
    };
    for (int cpsm : {4, 6}) {
        for (const V& v : vs) {
            char nm[200];
            snprintf(nm, sizeof nm, "%s, %d CTAs/SM", v.name, cpsm);
            report(nm, timed_ms([&] {
                cudaMemsetAsync(counter, 0, 4);
                if (v.store == 0) gather_probe<0><<<sms * cpsm, 160, smem>>>(table, table_rows, v.row_bytes, v.load_bytes, out, n_tiles, counter);
                else gather_probe<1><<<sms * cpsm, 160, smem>>>(table, table_rows, v.row_bytes, v.load_bytes, out, n_tiles, counter);
            }));
        }
    }
    return 0;
}
