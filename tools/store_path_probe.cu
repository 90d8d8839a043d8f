// store_path_probe.cu — how fast can one B200 WRITE the gather's output (2^21 windows of 2560 B, 5.4 GB) through the
// different store paths?  All variants write pseudo-distinct data (no constant fill), persistent grid unless stated.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/store_path_probe tools/store_path_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// (a) plain grid-stride STG.128 fill, non-persistent: what torch's fill does
__global__ void stg_fill(float4* out, size_t n_vec) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n_vec; i += stride) { float f = (float)(i & 1023); out[i] = make_float4(f, f + 1, f + 2, f + 3); }
}

// (b) persistent CTAs, chunk-strided like the gather: CTA c writes chunks c, c+grid, ... of CHUNK bytes with STG.128 from registers
template <int CS>
__global__ void stg_chunks(char* out, size_t n_chunks, int chunk) {
    const int nv = chunk / 16;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        float4* dst = reinterpret_cast<float4*>(out + c * (size_t)chunk);
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            float f = (float)((c + v) & 1023);
            float4 x = make_float4(f, f + 1, f + 2, f + 3);
            if (CS) __stcs(dst + v, x); else dst[v] = x;
        }
    }
}

// (c) persistent CTAs, data staged in shared memory (as after the patch), stored with LDS.128 + STG.128 by all threads
__global__ void stg_from_smem(char* out, size_t n_chunks, int chunk, int stages) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < stages * chunk / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)(i + blockIdx.x);
    __syncthreads();
    const int nv = chunk / 16;
    int q = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++q) {
        const float4* src = reinterpret_cast<const float4*>(smem + (size_t)(q % stages) * chunk);
        float4* dst = reinterpret_cast<float4*>(out + c * (size_t)chunk);
        for (int v = threadIdx.x; v < nv; v += blockDim.x) dst[v] = src[v];
    }
}

// (d) persistent CTAs, one thread issues a bulk store per chunk from rotating smem stages, DEPTH stores in flight
template <int DEPTH>
__global__ void tma_store(char* out, size_t n_chunks, int chunk, int stages) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < stages * chunk / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)(i + blockIdx.x);
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        int q = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++q) {
            bulk_s2g(out + c * (size_t)chunk, smem + (size_t)(q % stages) * chunk, (uint32_t)chunk);
            bulk_commit();
            bulk_wait_read<DEPTH>();
        }
        bulk_wait_read<0>();
    }
}

// (e) like (d) but ISSUERS threads (one per warp) each issue bulk stores of chunk/ISSUERS bytes
__global__ void tma_store_multi(char* out, size_t n_chunks, int chunk, int stages, int issuers) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < stages * chunk / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)(i + blockIdx.x);
    fence_proxy_async();
    __syncthreads();
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        const int part = chunk / issuers;
        int q = 0;
        for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++q) {
            bulk_s2g(out + c * (size_t)chunk + (size_t)w * part, smem + (size_t)(q % stages) * chunk + (size_t)w * part, (uint32_t)part);
            bulk_commit();
            bulk_wait_read<1>();
        }
        bulk_wait_read<0>();
    }
}

// (f) like (d)/(b) with another chunk -> CTA mapping.  MAP 0: strided (chunk = cta + k*grid, the gather's); 1: blocked (CTA b
// owns chunks [b*per, (b+1)*per)); 2: blocked in SEG segments (the buffer cut into SEG parts, strided inside a part)
__device__ __forceinline__ size_t map_chunk(int map, size_t k, size_t n_chunks, int seg) {
    const size_t g = gridDim.x, b = blockIdx.x;
    if (map == 0) return b + k * g;
    const size_t per = (n_chunks + g - 1) / g;
    if (map == 1) return k < per ? b * per + k : n_chunks;
    // map 2: CTAs split into `seg` teams; team t owns the t-th part of the buffer and strides inside it
    const size_t team = b % seg, idx = b / seg, team_size = (g + seg - 1) / seg;
    const size_t part = (n_chunks + seg - 1) / seg;
    const size_t c = idx + k * team_size;
    return c < part ? team * part + c : n_chunks;
}
__global__ void tma_store_map(char* out, size_t n_chunks, int chunk, int stages, int map, int seg) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < stages * chunk / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)(i + blockIdx.x);
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        int q = 0;
        for (size_t k = 0;; ++k, ++q) {
            const size_t c = map_chunk(map, k, n_chunks, seg);
            if (c >= n_chunks) break;
            bulk_s2g(out + c * (size_t)chunk, smem + (size_t)(q % stages) * chunk, (uint32_t)chunk);
            bulk_commit();
            bulk_wait_read<1>();
        }
        bulk_wait_read<0>();
    }
}
__global__ void stg_chunks_map(char* out, size_t n_chunks, int chunk, int map, int seg) {
    const int nv = chunk / 16;
    for (size_t k = 0;; ++k) {
        const size_t c = map_chunk(map, k, n_chunks, seg);
        if (c >= n_chunks) break;
        float4* dst = reinterpret_cast<float4*>(out + c * (size_t)chunk);
        for (int v = threadIdx.x; v < nv; v += blockDim.x) {
            float f = (float)((c + v) & 1023);
            dst[v] = make_float4(f, f + 1, f + 2, f + 3);
        }
    }
}

// (g) persistent CTAs, chunks claimed from a global counter (work stealing): does the static assignment cost the bandwidth?
__global__ void tma_store_dyn(char* out, size_t n_chunks, int chunk, int stages, unsigned long long* counter, int batch) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < stages * chunk / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = (float)(i + blockIdx.x);
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        int q = 0;
        for (;;) {
            const size_t c0 = (size_t)atomicAdd(counter, (unsigned long long)batch);
            if (c0 >= n_chunks) break;
            for (int j = 0; j < batch && c0 + j < n_chunks; ++j, ++q) {
                bulk_s2g(out + (c0 + j) * (size_t)chunk, smem + (size_t)(q % stages) * chunk, (uint32_t)chunk);
                bulk_commit();
                bulk_wait_read<1>();
            }
        }
        bulk_wait_read<0>();
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
    const size_t n_env = 1u << 21, win = 2560;
    const size_t bytes = n_env * win;
    char* out = nullptr;
    CK(cudaMalloc(&out, bytes));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    auto report = [&](const char* name, double ms) { printf("{\"variant\": \"%s\", \"ms\": %.4f, \"TBps\": %.3f}\n", name, ms, bytes / ms / 1e9); fflush(stdout); };
    report("stg_fill grid-stride (sms*16 x 256)", timed_ms([&] { stg_fill<<<sms * 16, 256>>>(reinterpret_cast<float4*>(out), bytes / 16); }));
    report("stg_fill grid-stride (sms*32 x 512)", timed_ms([&] { stg_fill<<<sms * 32, 512>>>(reinterpret_cast<float4*>(out), bytes / 16); }));
    for (int chunk : {2560, 10240, 20480}) {
        const size_t nc = bytes / chunk;
        char nm[160];
        for (int cpsm : {4, 8}) {
            snprintf(nm, sizeof nm, "stg_chunks regs chunk=%d ctas/sm=%d threads=128", chunk, cpsm);
            report(nm, timed_ms([&] { stg_chunks<0><<<sms * cpsm, 128>>>(out, nc, chunk); }));
        }
        snprintf(nm, sizeof nm, "stg_chunks regs st.cs chunk=%d ctas/sm=4 threads=128", chunk);
        report(nm, timed_ms([&] { stg_chunks<1><<<sms * 4, 128>>>(out, nc, chunk); }));
        snprintf(nm, sizeof nm, "stg_chunks regs chunk=%d ctas/sm=4 threads=256", chunk);
        report(nm, timed_ms([&] { stg_chunks<0><<<sms * 4, 256>>>(out, nc, chunk); }));
    }
    CK(cudaFuncSetAttribute(stg_from_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(tma_store<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(tma_store<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(tma_store<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(tma_store_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    {
        const int chunk = 10240;
        const size_t nc = bytes / chunk;
        char nm[160];
        for (int threads : {128, 256}) {
            snprintf(nm, sizeof nm, "stg_from_smem chunk=10240 stages=3 ctas/sm=4 threads=%d", threads);
            report(nm, timed_ms([&] { stg_from_smem<<<sms * 4, threads, 3 * chunk>>>(out, nc, chunk, 3); }));
        }
        snprintf(nm, sizeof nm, "tma_store depth=1 chunk=10240 stages=3 ctas/sm=4 (the gather's store warp)");
        report(nm, timed_ms([&] { tma_store<1><<<sms * 4, 64, 3 * chunk>>>(out, nc, chunk, 3); }));
        snprintf(nm, sizeof nm, "tma_store depth=2 chunk=10240 stages=4 ctas/sm=4");
        report(nm, timed_ms([&] { tma_store<2><<<sms * 4, 64, 4 * chunk>>>(out, nc, chunk, 4); }));
        snprintf(nm, sizeof nm, "tma_store depth=4 chunk=10240 stages=5 ctas/sm=4");
        report(nm, timed_ms([&] { tma_store<4><<<sms * 4, 64, 5 * chunk>>>(out, nc, chunk, 5); }));
        snprintf(nm, sizeof nm, "tma_store depth=1 chunk=10240 stages=3 ctas/sm=6");
        report(nm, timed_ms([&] { tma_store<1><<<sms * 6, 64, 3 * chunk>>>(out, nc, chunk, 3); }));
        snprintf(nm, sizeof nm, "tma_store depth=1 chunk=10240 stages=2 ctas/sm=8");
        report(nm, timed_ms([&] { tma_store<1><<<sms * 8, 64, 2 * chunk>>>(out, nc, chunk, 2); }));
        for (int issuers : {2, 4}) {
            snprintf(nm, sizeof nm, "tma_store_multi issuers=%d chunk=10240 stages=3 ctas/sm=4", issuers);
            report(nm, timed_ms([&] { tma_store_multi<<<sms * 4, 128, 3 * chunk>>>(out, nc, chunk, 3, issuers); }));
        }
    }
    {
        const int chunk = 2560;
        const size_t nc = bytes / chunk;
        report("tma_store depth=4 chunk=2560 stages=8 ctas/sm=4", timed_ms([&] { tma_store<4><<<sms * 4, 64, 8 * chunk>>>(out, nc, chunk, 8); }));
    }
    {
        const int chunk = 20480;
        const size_t nc = bytes / chunk;
        report("tma_store depth=1 chunk=20480 stages=2 ctas/sm=4", timed_ms([&] { tma_store<1><<<sms * 4, 64, 2 * chunk>>>(out, nc, chunk, 2); }));
    }
    CK(cudaFuncSetAttribute(tma_store_map, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    {
        const int chunk = 10240;
        const size_t nc = bytes / chunk;
        char nm[160];
        for (int map : {0, 1}) {
            snprintf(nm, sizeof nm, "tma_store_map map=%s chunk=10240 stages=3 ctas/sm=4", map ? "blocked" : "strided");
            report(nm, timed_ms([&] { tma_store_map<<<sms * 4, 64, 3 * chunk>>>(out, nc, chunk, 3, map, 1); }));
            snprintf(nm, sizeof nm, "stg_chunks_map map=%s chunk=10240 ctas/sm=4 threads=128", map ? "blocked" : "strided");
            report(nm, timed_ms([&] { stg_chunks_map<<<sms * 4, 128>>>(out, nc, chunk, map, 1); }));
        }
        for (int seg : {2, 4, 8, 16, 37, 74, 148}) {
            snprintf(nm, sizeof nm, "tma_store_map map=teams seg=%d chunk=10240 stages=3 ctas/sm=4", seg);
            report(nm, timed_ms([&] { tma_store_map<<<sms * 4, 64, 3 * chunk>>>(out, nc, chunk, 3, 2, seg); }));
        }
        report("stg_chunks_map map=teams seg=8 chunk=10240 ctas/sm=4 threads=128", timed_ms([&] { stg_chunks_map<<<sms * 4, 128>>>(out, nc, chunk, 2, 8); }));
        report("stg_chunks_map map=blocked chunk=10240 ctas/sm=8 threads=128", timed_ms([&] { stg_chunks_map<<<sms * 8, 128>>>(out, nc, chunk, 1, 1); }));
        report("stg_chunks_map map=blocked chunk=10240 ctas/sm=16 threads=128", timed_ms([&] { stg_chunks_map<<<sms * 16, 128>>>(out, nc, chunk, 1, 1); }));
    }
    {
        char nm[160];
        for (int mult : {4, 8, 16, 64}) {
            snprintf(nm, sizeof nm, "stg_fill grid-stride (sms*%d x 512)", mult);
            report(nm, timed_ms([&] { stg_fill<<<sms * mult, 512>>>(reinterpret_cast<float4*>(out), bytes / 16); }));
        }
        for (int chunk : {8192, 10240}) {
            const size_t nc = bytes / chunk;
            snprintf(nm, sizeof nm, "stg_chunks regs chunk=%d ctas/sm=4 threads=512", chunk);
            report(nm, timed_ms([&] { stg_chunks<0><<<sms * 4, 512>>>(out, nc, chunk); }));
            snprintf(nm, sizeof nm, "stg_chunks regs chunk=%d ONE CHUNK PER CTA (grid=n_chunks) threads=128", chunk);
            report(nm, timed_ms([&] { stg_chunks<0><<<(unsigned)nc, 128>>>(out, nc, chunk); }));
            snprintf(nm, sizeof nm, "stg_chunks regs chunk=%d ONE CHUNK PER CTA (grid=n_chunks) threads=512", chunk);
            report(nm, timed_ms([&] { stg_chunks<0><<<(unsigned)nc, 512>>>(out, nc, chunk); }));
        }
        const size_t nc = bytes / 10240;
        report("tma_store depth=1 chunk=10240 stages=1 ONE CHUNK PER CTA (grid=n_chunks)", timed_ms([&] { tma_store<1><<<(unsigned)nc, 64, 10240>>>(out, nc, 10240, 1); }));
    }
    {
        CK(cudaFuncSetAttribute(tma_store_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        unsigned long long* counter = nullptr;
        CK(cudaMalloc(&counter, 8));
        const int chunk = 10240;
        const size_t nc = bytes / chunk;
        char nm[160];
        for (int batch : {1, 3, 12}) {
            snprintf(nm, sizeof nm, "tma_store_dyn WORK STEALING batch=%d chunk=10240 stages=3 ctas/sm=4", batch);
            report(nm, timed_ms([&] { cudaMemsetAsync(counter, 0, 8); tma_store_dyn<<<sms * 4, 64, 3 * chunk>>>(out, nc, chunk, 3, counter, batch); }));
        }
        CK(cudaFree(counter));
    }
    CK(cudaFree(out));
    return 0;
}
