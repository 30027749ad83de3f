// nldpc_spec_host.cuh — host-side pieces shared by the translation units that instantiate specialised kernels:
// the constant-memory weight arena (one per translation unit: each has its own copy of c_wb) and the pack kernel.
#pragma once
#include "nldpc_spec_kernel.cuh"

#include <mutex>
#include <vector>

namespace nldpc {


namespace {

// True while `st` is being captured into a CUDA graph.  The constant arena hands out ranges at LAUNCH time (event queries,
// host-side bookkeeping), which a replayed graph would not repeat.  Captured launches therefore either read their weights
// from global memory (Neural: the LDG variant of the same kernel) or use the fixed range ConstArena::acquire_captured hands
// out (Boosted forward / backward sweeps: same specialised kernels as eager launches).
inline bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &status) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return status != cudaStreamCaptureStatusNone;
}

// {w[i], b[i]} -> constant arena (written through its global address; visible to the launches that follow
// on the same stream: the constant cache is invalidated at kernel boundaries)
__global__ void pack_wb_kernel(const float *__restrict__ w, const float *__restrict__ b, float2 *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float2(w ? w[i] : 1.0f, b ? b[i] : 0.0f);   // absent weights: x * 1 + 0 is exact
}

// Ring allocator over the constant arena, per device.  A range is reused only after the launch that read it
// has finished: a launch on another stream that wants an overlapping range first waits on that launch's event.
struct ConstArena {
    struct Pending { int off, len; cudaEvent_t ev; cudaStream_t st; };
    std::mutex mu;
    int head = 0;
    std::vector<Pending> pend;
    std::vector<cudaEvent_t> pool;
    float2 *base = nullptr;

    // returns offset (float2 units) or -1 when `len` does not fit at all
    int acquire(int len, cudaStream_t st, cudaError_t *err) {
        *err = cudaSuccess;
        if (len > kConstFloat2) return -1;
        std::lock_guard<std::mutex> lk(mu);
        if (!base) {
            *err = cudaGetSymbolAddress((void **)&base, c_wb);
            if (*err != cudaSuccess) return -1;
        }
        if (head + len > kConstFloat2) head = 0;
        const int off = head;
        head += (len + 1) & ~1;   // keep 16-byte alignment
        for (size_t i = 0; i < pend.size();) {
            Pending &p = pend[i];
            const bool overlap = p.off < off + len && off < p.off + p.len;
            if (overlap) {
                if (p.st != st) {
                    *err = cudaStreamWaitEvent(st, p.ev, 0);
                    if (*err != cudaSuccess) return -1;
                }
                pool.push_back(p.ev);
                pend[i] = pend.back();
                pend.pop_back();
            } else {
                i++;
            }
        }
        return off;
    }
    // Launch being captured into a CUDA graph: every captured launch uses the range at offset 0 — inside a graph the
    // pack -> consumer pairs are ordered by the capture stream, and a replay is ordered against eager launches of the stream
    // it is replayed on.  What nothing orders is a replay against launches of this library running CONCURRENTLY on another
    // stream of the device (the ring's events cannot be re-recorded by a replay): callers must not do that.
    // -1 when the arena has not been set up by an earlier eager launch (the caller then uses the table-driven kernel).
    int acquire_captured(int len) {
        if (len > kConstFloat2) return -1;
        std::lock_guard<std::mutex> lk(mu);
        return base ? 0 : -1;
    }
    // call after the consumer kernel has been enqueued on `st`
    cudaError_t release_after(int off, int len, cudaStream_t st) {
        std::lock_guard<std::mutex> lk(mu);
        cudaEvent_t ev;
        if (!pool.empty()) { ev = pool.back(); pool.pop_back(); }
        else {
            cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        cudaError_t e = cudaEventRecord(ev, st);
        if (e != cudaSuccess) return e;
        pend.push_back({off, len, ev, st});
        return cudaSuccess;
    }
};

ConstArena &arena_for_current_device() {
    static ConstArena arenas[64];
    int dev = 0;
    cudaGetDevice(&dev);
    return arenas[dev & 63];
}


template <class K>
cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace
}  // namespace nldpc
