// nldpc_spec.cu — registry + instantiation of the specialised kernels for the built-in codes
// (5G NR BG2 set 0 at Z=16, 802.16e N=576 R=3/4 at Z=24).
#include "nldpc_spec.cuh"

#include <algorithm>

#include "generated/nldpc_graph_bg2z16.cuh"
#include "generated/nldpc_graph_wimaxz24.cuh"

#include "nldpc_spec_host.cuh"
#include "nldpc_spec_backward.cuh"

namespace nldpc {

namespace {

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

template <class G>
int prepare() {
    cudaError_t e;
    using Every = typename KernelCfg<G, true, false>::type;
    using Last = typename KernelCfg<G, false, false>::type;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, true>, Every::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, false>, Every::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, false, true>, Last::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, false, false>, Last::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, true, 0, false, false, true>, Every::kSmemBytes)) != cudaSuccess) return (int)e;   // soft only
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, false, 0, false, false, true>, Every::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, true, 0, false, true>, Every::kSmemBytes)) != cudaSuccess) return (int)e;      // + training dump
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, false, 0, false, true>, Every::kSmemBytes)) != cudaSuccess) return (int)e;
    return 0;
}

template <class G>
int launch_neural(const DecodeArgs &a, int sm_count, cudaStream_t st) {
    using Every = typename KernelCfg<G, true, false>::type;       // list mode may carry output staging rows
    using Last = typename KernelCfg<G, false, false>::type;
    const int n_units = (a.B + Last::Shape::kCw - 1) / Last::Shape::kCw;
    if (a.hist_v2c && a.hist_fmt != 1) return -1;      // slot-major training dump: the table-driven kernel writes that format
    const bool every = a.soft_mode == 1 || a.hard_mode == 1 || a.hist_v2c != nullptr;
    const int grid = std::min(n_units, sm_count * (every ? Every::kCtasPerSm : Last::kCtasPerSm));
    DecodeArgs args = a;
    // weights -> constant arena (uniform-datapath reads in the kernel); LDG variant if they do not fit
    ConstArena &arena = arena_for_current_device();
    const int len = a.T * G::E;
    cudaError_t err = cudaSuccess;
    const bool capturing = stream_is_capturing(st);      // CUDA graph capture: fixed arena range (ConstArena::acquire_captured)
    // (a captured Neural decode reads its weights with LDG instead: no arena range is taken away from the eager ring for it)
    const int off = capturing ? -1 : arena.acquire(len, st, &err);
    if (err != cudaSuccess) return (int)err;
    args.wb_off = off;
    // A group notes the units it has to decode a second time (exact zeros, see run_unit) in a 64-bit mask, so one launch gives
    // a group at most 64 units: larger batches run as several launches over consecutive unit ranges (~150 k codewords each on
    // BG2; the arguments stay the same, only [unit_begin, unit_end) moves).
    const int groups = every ? Every::kGroups : Last::kGroups;
    const long long per_launch = 64ll * grid * groups;
    auto launch_all = [&](auto kernel_every, auto kernel_last) -> cudaError_t {
        for (long long u0 = 0; u0 < n_units; u0 += per_launch) {
            args.unit_begin = (int)u0;
            args.unit_end = (int)std::min<long long>(n_units, u0 + per_launch);
            if (every) kernel_every<<<grid, Every::kThreads, Every::kSmemBytes, st>>>(args);
            else kernel_last<<<grid, Last::kThreads, Last::kSmemBytes, st>>>(args);
        }
        return cudaGetLastError();
    };
    // every-iteration instantiations: + training dump (Dumping<>), soft outputs only (SoftOnly<>: what forward() asks for), or both outputs
    const bool dump = a.hist_v2c != nullptr, soft_only = a.hard_mode == 0;
    if (dump && !soft_only) return -1;      // (a training dump with hard decisions: not a combination the module ever asks for)
    if (off >= 0) {
#if NLDPC_CN_PAIR
        if ((err = upload_wb_paired<G>(arena, a.w, a.b, off, len, st)) != cudaSuccess) return (int)err;
#else
        if ((err = upload_wb(arena, a.w, a.b, off, len, st)) != cudaSuccess) return (int)err;
#endif
        err = dump        ? launch_all(nldpc_spec_neural_kernel<G, true, true, 0, false, true>, nldpc_spec_neural_kernel<G, false, true>)
              : soft_only ? launch_all(nldpc_spec_neural_kernel<G, true, true, 0, false, false, true>, nldpc_spec_neural_kernel<G, false, true>)
                          : launch_all(nldpc_spec_neural_kernel<G, true, true>, nldpc_spec_neural_kernel<G, false, true>);
        if (err != cudaSuccess || capturing) return (int)err;
        return (int)arena.release_after(off, len, st);
    }
    return (int)(dump        ? launch_all(nldpc_spec_neural_kernel<G, true, false, 0, false, true>, nldpc_spec_neural_kernel<G, false, false>)
                 : soft_only ? launch_all(nldpc_spec_neural_kernel<G, true, false, 0, false, false, true>, nldpc_spec_neural_kernel<G, false, false>)
                             : launch_all(nldpc_spec_neural_kernel<G, true, false>, nldpc_spec_neural_kernel<G, false, false>));
}

}  // namespace

int spec_find(const int32_t *bg, int M, int N, int Z) {
    if (graph_matches<gen::Bg2Z16>(bg, M, N, Z)) return 0;
    if (graph_matches<gen::WimaxZ24>(bg, M, N, Z)) return 1;
    return -1;
}

int spec_prepare(int id) {
    switch (id) {
        case 0: { int rc = prepare<gen::Bg2Z16>(); if (!rc) rc = spec_boosted_prepare_bg2(); return rc ? rc : spec_train_prepare_bg2(); }
        case 1: { int rc = prepare<gen::WimaxZ24>(); if (!rc) rc = spec_boosted_prepare_wimax(); return rc ? rc : spec_train_prepare_wimax(); }
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

int spec_launch_boosted(int id, const DecodeArgs &a, int sm_count, cudaStream_t st) {
    if (a.hist_v2c) {       // training dump: the training variant (check-packed records) or the table-driven kernel (slot-major)
        switch (id) {
            case 0: return spec_train_launch_bg2(a, sm_count, st);
            case 1: return spec_train_launch_wimax(a, sm_count, st);
            default: return -1;
        }
    }
    switch (id) {
        case 0: return spec_boosted_launch_bg2(a, sm_count, st);
        case 1: return spec_boosted_launch_wimax(a, sm_count, st);
        default: return -1;
    }
}

int spec_backward_scratch_rows(int id) {
    switch (id) {
        case 0: return gen::Bg2Z16::kXRegs > 0 ? gen::Bg2Z16::kXRegs : 1;
        case 1: return gen::WimaxZ24::kXRegs > 0 ? gen::WimaxZ24::kXRegs : 1;
        default: return 0;
    }
}

bool spec_backward_covers(int id, int mode, int T, bool has_cn_w, bool has_vn_w, bool ucn, int qbit) {
    switch (id) {
        case 0: return spec_bwd_covers<gen::Bg2Z16>(mode, T, has_cn_w, has_vn_w, ucn, qbit);
        case 1: return spec_bwd_covers<gen::WimaxZ24>(mode, T, has_cn_w, has_vn_w, ucn, qbit);
        default: return false;
    }
}

size_t spec_dump_bytes_per_cw_iter(int id, int mode) {
    switch (id) {
        case 0: return mode == 2 ? BwdStage<gen::Bg2Z16, 2>::kCwBytes : BwdStage<gen::Bg2Z16, 0>::kCwBytes;
        case 1: return mode == 2 ? BwdStage<gen::WimaxZ24, 2>::kCwBytes : BwdStage<gen::WimaxZ24, 0>::kCwBytes;
        default: return 0;
    }
}

int spec_launch_backward(int id, const BwdArgs &a, int sm_count, cudaStream_t st) {
    if (a.mode == 0) {
        switch (id) {
            case 0: return spec_bwd_launch<gen::Bg2Z16, false>(a, 0, sm_count, st);
            case 1: return spec_bwd_launch<gen::WimaxZ24, false>(a, 1, sm_count, st);
            default: return -1;
        }
    }
    switch (id) {
        case 0: return spec_boosted_backward_bg2(a, sm_count, st);
        case 1: return spec_boosted_backward_wimax(a, sm_count, st);
        default: return -1;
    }
}

}  // namespace nldpc
