// nldpc_spec.cu — registry + instantiation of the specialised kernels for the built-in codes
// (5G NR BG2 set 0 at Z=16, 802.16e N=576 R=3/4 at Z=24).
#include "nldpc_spec.cuh"

#include <algorithm>

#include "generated/nldpc_graph_bg2z16.cuh"
#include "generated/nldpc_graph_wimaxz24.cuh"

#include <mutex>
#include <vector>

namespace nldpc {


namespace {

// {w[i], b[i]} -> constant arena (written through its global address; visible to the launches that follow
// on the same stream: the constant cache is invalidated at kernel boundaries)
__global__ void pack_wb_kernel(const float *__restrict__ w, const float *__restrict__ b, float2 *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_float2(w[i], b ? b[i] : 0.0f);
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

template <class G>
bool graph_matches(const int32_t *bg, int M, int N, int Z) {
    if (M != G::M || N != G::N || Z != G::Z) return false;
    const int32_t *ref = G::basegraph();
    for (int i = 0; i < M * N; i++) {
        const bool a = bg[i] == -1, b = ref[i] == -1;
        if (a != b) return false;
        if (!a && (bg[i] % Z) != (ref[i] % Z)) return false;
    }
    return true;
}

template <class K>
cudaError_t set_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <class G>
int prepare() {
    cudaError_t e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, true>, SpecCfg<G>::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, false>, SpecCfg<G>::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, false, true>, SpecCfg<G>::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, false, false>, SpecCfg<G>::kSmemBytes)) != cudaSuccess) return (int)e;
    return 0;
}

template <class G>
int launch_neural(const DecodeArgs &a, int sm_count, cudaStream_t st) {
    using Cfg = SpecCfg<G>;
    const int n_units = (a.B + Cfg::Shape::kCw - 1) / Cfg::Shape::kCw;
    const int ctas = (n_units + Cfg::kGroups - 1) / Cfg::kGroups;
    const int grid = std::min(ctas, sm_count * 2);
    const bool every = a.soft_mode == 1 || a.hard_mode == 1;
    DecodeArgs args = a;
    // weights -> constant arena (uniform-datapath reads in the kernel); LDG variant if they do not fit
    ConstArena &arena = arena_for_current_device();
    const int len = a.T * G::E;
    cudaError_t err;
    const int off = arena.acquire(len, st, &err);
    if (err != cudaSuccess) return (int)err;
    args.wb_off = off;
    if (off >= 0) {
        pack_wb_kernel<<<(len + 255) / 256, 256, 0, st>>>(a.w, a.b, arena.base + off, len);
        if (every) nldpc_spec_neural_kernel<G, true, true><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(args);
        else nldpc_spec_neural_kernel<G, false, true><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(args);
        err = cudaGetLastError();
        if (err != cudaSuccess) return (int)err;
        return (int)arena.release_after(off, len, st);
    }
    if (every) nldpc_spec_neural_kernel<G, true, false><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(args);
    else nldpc_spec_neural_kernel<G, false, false><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(args);
    return (int)cudaGetLastError();
}

}  // namespace

int spec_find(const int32_t *bg, int M, int N, int Z) {
    if (graph_matches<gen::Bg2Z16>(bg, M, N, Z)) return 0;
    if (graph_matches<gen::WimaxZ24>(bg, M, N, Z)) return 1;
    return -1;
}

int spec_prepare(int id) {
    switch (id) {
        case 0: return prepare<gen::Bg2Z16>();
        case 1: return prepare<gen::WimaxZ24>();
        default: return -1;
    }
}

int spec_cw_per_cta(int id) { return id == 0 ? SpecCfg<gen::Bg2Z16>::kCwPerCta : (id == 1 ? SpecCfg<gen::WimaxZ24>::kCwPerCta : 0); }
int spec_threads(int id) { return id == 0 ? SpecCfg<gen::Bg2Z16>::kThreads : (id == 1 ? SpecCfg<gen::WimaxZ24>::kThreads : 0); }

int spec_launch_neural(int id, const DecodeArgs &a, int sm_count, cudaStream_t st) {
    switch (id) {
        case 0: return launch_neural<gen::Bg2Z16>(a, sm_count, st);
        case 1: return launch_neural<gen::WimaxZ24>(a, sm_count, st);
        default: return -1;
    }
}

}  // namespace nldpc
