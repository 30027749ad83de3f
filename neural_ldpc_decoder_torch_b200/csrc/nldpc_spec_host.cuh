// nldpc_spec_host.cuh — host-side pieces shared by the translation units that instantiate specialised kernels:
// the constant-memory weight arena (one per translation unit: each has its own copy of c_wb) and the pack kernel.
#pragma once
#include "nldpc_spec_kernel.cuh"

#include <mutex>
#include <vector>

namespace nldpc {


namespace {

// True while `st` is being captured into a CUDA graph.  The constant arena hands out ring ranges at LAUNCH time (event
// queries, host-side bookkeeping), which a replayed graph would not repeat.  Captured launches therefore either read their
// weights from global memory (Neural: the LDG variant of the same kernel) or use the reserved range
// ConstArena::acquire_captured hands out (Boosted forward / backward sweeps: same specialised kernels as eager launches).
inline bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &status) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return status != cudaStreamCaptureStatusNone;
}

// {w[i], b[i]} pairs -> a staging buffer in GLOBAL memory; upload_wb then copies the range into the __constant__ array with
// cudaMemcpyToSymbolAsync (device to device, stream ordered, capturable as a memcpy node).  Device code never writes the
// constant bank itself: CUDA leaves that undefined and gives no guarantee that the constant caches see it.
__global__ void pack_wb_kernel(const float *__restrict__ w, const float *__restrict__ b, float2 *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float2(w ? w[i] : 1.0f, b ? b[i] : 0.0f);   // absent weights: x * 1 + 0 is exact
}

// Allocator over the constant arena (and the staging buffer, which mirrors it offset for offset), per device.
//   [0, ring_end)            ring for eager launches.  A range is reused only after the launch that read it has finished: a
//                            launch on another stream that wants an overlapping range first waits on that launch's event.
//   [ring_end, kConstFloat2) ranges of launches captured into CUDA graphs: handed out top-down while capturing, never given
//                            to an eager launch afterwards (a replay records no event the ring could wait on).
struct ConstArena {
    struct Pending { int off, len; cudaEvent_t ev; cudaStream_t st; };
    std::mutex mu;
    int head = 0;
    int ring_end = kConstFloat2;
    std::vector<Pending> pend;
    std::vector<cudaEvent_t> pool;
    float2 *stage = nullptr;      // global-memory mirror of the arena (pack target, source of the symbol copy)

    cudaError_t ensure_stage() {
        if (stage) return cudaSuccess;
        return cudaMalloc((void **)&stage, sizeof(float2) * kConstFloat2);
    }

    // eager launch: returns offset (float2 units) or -1 when `len` does not fit the ring at all
    int acquire(int len, cudaStream_t st, cudaError_t *err) {
        *err = cudaSuccess;
        std::lock_guard<std::mutex> lk(mu);
        if (len > ring_end) return -1;
        if ((*err = ensure_stage()) != cudaSuccess) return -1;
        if (head + len > ring_end) head = 0;
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
    // Launch being captured into a CUDA graph: a fixed range at the top of the arena, shared by all captured launches of the
    // same length class (inside a graph the upload -> consumer pairs are ordered by the capture stream, and replays of
    // graphs on one stream are ordered by that stream) and taken away from the eager ring for good, so no eager launch on any
    // stream can overwrite weights a replay is reading.  The ring shrinks to what is left; eager launches that no longer fit
    // it use the LDG / table-driven kernels.  -1 when `len` does not fit or a pending eager launch still occupies the range
    // and cannot be waited for during capture (the caller then uses the kernel variant without the arena).
    int acquire_captured(int len, cudaStream_t st, cudaError_t *err) {
        *err = cudaSuccess;
        std::lock_guard<std::mutex> lk(mu);
        if (len > kConstFloat2 || !stage) return -1;      // (no cudaMalloc while a capture is open: an eager launch comes first)
        const int off = (kConstFloat2 - len) & ~1;
        if (off < ring_end) {
            // Eager launches enqueued earlier may still be reading the part of the ring being taken.  Nothing may be queried or
            // synchronised while a capture is open, so the graph itself waits for them: an external event-wait node (a no-op
            // on later replays, the event is never recorded again: it does not go back to the pool).
            for (size_t i = 0; i < pend.size();) {
                Pending &p = pend[i];
                if (p.off + p.len > off) {
                    if ((*err = cudaStreamWaitEvent(st, p.ev, cudaEventWaitExternal)) != cudaSuccess) return -1;
                    pend[i] = pend.back();
                    pend.pop_back();
                } else {
                    i++;
                }
            }
            ring_end = off;
            if (head > ring_end) head = 0;
        }
        return off;
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


// {w, b} rows of one launch -> arena range [off, off + len): pack into the staging mirror, then a stream-ordered
// device-to-device copy into the constant array (which also invalidates the constant caches for the consumer kernel)
inline cudaError_t upload_wb(ConstArena &arena, const float *w, const float *b, int off, int len, cudaStream_t st) {
    pack_wb_kernel<<<(len + 255) / 256, 256, 0, st>>>(w, b, arena.stage + off, len);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(c_wb, arena.stage + off, sizeof(float2) * (size_t)len, sizeof(float2) * (size_t)off,
                                   cudaMemcpyDeviceToDevice, st);
}

// Neural forward kernels (NLDPC_CN_PAIR): the same T * E {w, b} values in the pair layout of WbPlan — float offsets within an
// iteration's 2 * E floats from the generated table G::wb_pair_off(), copied to the device once (never during a capture:
// captured Neural launches read their weights with LDG and do not come here).
__global__ void pack_wb_paired_kernel(const float *__restrict__ w, const float *__restrict__ b, const int *__restrict__ tab,
                                      float *__restrict__ dst, int E, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = i / E, e = i - t * E;
    float *it = dst + (size_t)t * 2 * E;
    it[tab[e]] = w ? w[i] : 1.0f;
    it[tab[E + e]] = b ? b[i] : 0.0f;
}
template <class G>
cudaError_t upload_wb_paired(ConstArena &arena, const float *w, const float *b, int off, int len, cudaStream_t st) {
    static std::mutex mu;
    static int *tabs[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int *tab = nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!tabs[dev & 63]) {
            cudaError_t e = cudaMalloc((void **)&tabs[dev & 63], sizeof(int) * 2 * G::E);
            if (e != cudaSuccess) return e;
            e = cudaMemcpy(tabs[dev & 63], G::wb_pair_off(), sizeof(int) * 2 * G::E, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) {
                cudaFree(tabs[dev & 63]);
                tabs[dev & 63] = nullptr;
                return e;
            }
        }
        tab = tabs[dev & 63];
    }
    pack_wb_paired_kernel<<<(len + 255) / 256, 256, 0, st>>>(w, b, tab, reinterpret_cast<float *>(arena.stage + off), G::E, len);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(c_wb, arena.stage + off, sizeof(float2) * (size_t)len, sizeof(float2) * (size_t)off,
                                   cudaMemcpyDeviceToDevice, st);
}

// descriptors of the looped checks -> this translation unit's c_desc, once per device.  Synchronous on purpose: when the call
// returns the table is in place for launches on ANY stream (the flag is shared by all of them).
template <class G>
cudaError_t ensure_loop_desc(int graph_slot, bool capturing) {
    static_assert(G::kLoopDescWords <= kDescStride, "descriptor slot too small");
    static std::mutex mu;
    static bool done[64] = {};
    if (G::kLoopChecks == 0) return cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev & 63]) return cudaSuccess;
    if (capturing) return cudaErrorStreamCaptureUnsupported;      // first use inside a capture: the caller falls back
    cudaError_t e = cudaMemcpyToSymbol(c_desc, G::loop_desc(), sizeof(uint32_t) * G::kLoopDescWords,
                                       sizeof(uint32_t) * (size_t)graph_slot * kDescStride, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) done[dev & 63] = true;
    return e;
}

// weights only (Boosted decoders): n plain floats at arena offset `off` (float2 units), absent weights = 1.0
__global__ void pack_w_kernel(const float *__restrict__ w, float *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = w ? w[i] : 1.0f;
}
inline cudaError_t upload_w(ConstArena &arena, const float *w, int off, int n, cudaStream_t st) {
    pack_w_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<float *>(arena.stage + off), n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(c_wb, arena.stage + off, sizeof(float2) * (size_t)((n + 1) / 2), sizeof(float2) * (size_t)off,
                                   cudaMemcpyDeviceToDevice, st);
}

// the same for the specialised Boosted FORWARD kernels: iteration t's E weights at float offset t * pitch (pitch even, kWPitch)
__global__ void pack_w_pitched_kernel(const float *__restrict__ w, float *__restrict__ dst, int E, int pitch, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = i / pitch, e = i - t * pitch;
    dst[i] = (w && e < E) ? w[t * E + e] : 1.0f;
}
inline cudaError_t upload_w_pitched(ConstArena &arena, const float *w, int off, int T, int E, int pitch, cudaStream_t st) {
    const int n = T * pitch;
    pack_w_pitched_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, reinterpret_cast<float *>(arena.stage + off), E, pitch, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(c_wb, arena.stage + off, sizeof(float2) * (size_t)((n + 1) / 2), sizeof(float2) * (size_t)off,
                                   cudaMemcpyDeviceToDevice, st);
}

template <class K>
cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace
}  // namespace nldpc
