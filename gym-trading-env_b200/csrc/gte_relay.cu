// gte_relay.cu — the multi-GPU result relay (include/gte_b200.h, "result relay"): part of a rank's result block reaches
// the host through a PEER GPU's PCIe link.  Copy engines and one stream memory operation only: no kernel, no host thread.
//
//   sender GPU                         peer GPU                              host (shared memory, pinned in both processes)
//   reward[count..N) --NVLink DMA-->   payload
//   seq              --NVLink DMA-->   header.seq
//                                      stream waits header.seq >= seq
//                                      payload   ---------PCIe DMA------->   sender's result block, rewards [count..N)
//                                      header.seq --------PCIe DMA------->   sender's relay word   (polled by the sender's host)
#include <cstring>

#include <cuda.h>

#include "gte_launch.h"

namespace gte {

namespace {

using WaitValue32 = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

struct RelayCtx {
    cudaStream_t push = nullptr;
    cudaStream_t lane[8] = {};
    uint32_t* seq_ring = nullptr;        // pinned: the source of the sequence-word copies (slot = seq & 63)
    WaitValue32 wait_value = nullptr;
    int supported = -1;
};
RelayCtx g_relay[16];

cudaError_t relay_ctx(RelayCtx** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    RelayCtx& r = g_relay[dev & 15];
    if (r.supported < 0) {
        r.supported = 0;
        // stream memory operations (v2) are part of every CUDA 12 driver: finding the entry point is the whole check
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess && fn != nullptr) {
            r.wait_value = reinterpret_cast<WaitValue32>(fn);
            r.supported = 1;
        }
        (void)cudaGetLastError();
    }
    *out = &r;
    return cudaSuccess;
}

cudaError_t ensure_stream(cudaStream_t* s) {
    return *s != nullptr ? cudaSuccess : cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
}

}  // namespace

bool relay_supported() {
    RelayCtx* r = nullptr;
    return relay_ctx(&r) == cudaSuccess && r->supported == 1;
}

cudaError_t relay_alloc(int64_t bytes, void** dev_base, void* ipc_handle) {
    cudaError_t e;
    void* p = nullptr;
    const size_t total = (size_t)GTE_RELAY_HEADER_BYTES + (size_t)bytes;
    if ((e = cudaMalloc(&p, total)) != cudaSuccess) return e;
    if ((e = cudaMemset(p, 0, total)) != cudaSuccess) { cudaFree(p); return e; }
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) { cudaFree(p); return e; }
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(ipc_handle, &h, sizeof(h));
    *dev_base = p;
    return cudaSuccess;
}

cudaError_t relay_open(const void* ipc_handle, void** dev_base) {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, ipc_handle, sizeof(h));
    return cudaIpcOpenMemHandle(dev_base, h, cudaIpcMemLazyEnablePeerAccess);
}

cudaError_t relay_release(void* dev_base, bool opened) {
    return opened ? cudaIpcCloseMemHandle(dev_base) : cudaFree(dev_base);
}

cudaError_t relay_push(void* peer_base, const void* src_dev, int64_t bytes, uint32_t seq, cudaEvent_t after, cudaEvent_t done) {
    RelayCtx* r = nullptr;
    cudaError_t e;
    if ((e = relay_ctx(&r)) != cudaSuccess) return e;
    if ((e = ensure_stream(&r->push)) != cudaSuccess) return e;
    if (r->seq_ring == nullptr) {
        if ((e = cudaHostAlloc(reinterpret_cast<void**>(&r->seq_ring), 64 * sizeof(uint32_t), cudaHostAllocPortable)) != cudaSuccess) return e;
        std::memset(r->seq_ring, 0, 64 * sizeof(uint32_t));
    }
    if (after != nullptr && (e = cudaStreamWaitEvent(r->push, after, 0)) != cudaSuccess) return e;
    char* base = static_cast<char*>(peer_base);
    if ((e = cudaMemcpyAsync(base + GTE_RELAY_HEADER_BYTES, src_dev, (size_t)bytes, cudaMemcpyDefault, r->push)) != cudaSuccess) return e;
    // the sequence word follows the payload in stream order: whoever sees it sees the payload (a ring of 64 source
    // slots: a slot is rewritten 64 pushes later, long after its copy ran — the caller allows one push in flight)
    uint32_t* slot = r->seq_ring + (seq & 63u);
    *slot = seq;
    if ((e = cudaMemcpyAsync(base, slot, sizeof(uint32_t), cudaMemcpyDefault, r->push)) != cudaSuccess) return e;
    return done != nullptr ? cudaEventRecord(done, r->push) : cudaSuccess;
}

cudaError_t relay_serve(int lane, void* own_base, int64_t bytes, uint32_t seq, void* host_dst, void* host_seq) {
    RelayCtx* r = nullptr;
    cudaError_t e;
    if ((e = relay_ctx(&r)) != cudaSuccess) return e;
    if (r->supported != 1) return cudaErrorNotSupported;
    cudaStream_t s = nullptr;
    if ((e = ensure_stream(&r->lane[lane])) != cudaSuccess) return e;
    s = r->lane[lane];
    // wait (on the device, no SM involved) until the sender's copy engine has published this iteration's number
    if (r->wait_value(reinterpret_cast<CUstream>(s), reinterpret_cast<CUdeviceptr>(own_base), seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
        return cudaErrorUnknown;
    const char* base = static_cast<const char*>(own_base);
    if ((e = cudaMemcpyAsync(host_dst, base + GTE_RELAY_HEADER_BYTES, (size_t)bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    return cudaMemcpyAsync(host_seq, base, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
}

// Release a serve stream that is waiting for a sequence word which will never come (a failed set-up): write the word from
// this side.
cudaError_t relay_unblock(void* own_base, uint32_t seq) {
    cudaStream_t s = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(own_base, &seq, sizeof(seq), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    return e;
}

}  // namespace gte
