// nldpc_spec_boosted.cuh — instantiation helper for the specialised Boosted (MS / QMS q=5, no UCN) kernels of one code.
// Included by one translation unit per code so the 8 variants of each compile in parallel.
#pragma once
#include <algorithm>
#include <type_traits>

#include "nldpc_spec_host.cuh"

namespace nldpc {
namespace {

template <class G, int MODE, bool kXo>
int boosted_prepare_one() {
    cudaError_t e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, true, true, MODE, kXo>, KernelCfg<G, true, kXo>::type::kSmemBytes)) != cudaSuccess) return (int)e;
    if ((e = set_smem(nldpc_spec_neural_kernel<G, false, true, MODE, kXo>, KernelCfg<G, false, kXo>::type::kSmemBytes)) != cudaSuccess) return (int)e;
    return 0;
}

template <class G>
int boosted_prepare() {
    int rc;
    if ((rc = boosted_prepare_one<G, 1, false>())) return rc;
    if ((rc = boosted_prepare_one<G, 1, true>())) return rc;
    if ((rc = boosted_prepare_one<G, 2, false>())) return rc;
    if ((rc = boosted_prepare_one<G, 2, true>())) return rc;
    return 0;
}

template <class G, int MODE, bool kXo>
int boosted_launch_one(const DecodeArgs &args, int sm_count, cudaStream_t st) {
    const bool every = args.soft_mode == 1 || args.hard_mode == 1 || args.llr_all != nullptr;
    auto launch = [&](auto every_tag) {
        constexpr bool kEvery = decltype(every_tag)::value;
        using Cfg = typename KernelCfg<G, kEvery, kXo>::type;   // list mode keeps xa_origin rows on chip (throughput mode re-reads it, xo_global) and may stage its outputs
        const int n_units = (args.B + Cfg::Shape::kCw - 1) / Cfg::Shape::kCw;
                const int grid = std::min(n_units, sm_count * Cfg::kCtasPerSm);
        nldpc_spec_neural_kernel<G, kEvery, true, MODE, kXo><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(args);
    };
    if (every) launch(std::true_type{});
    else launch(std::false_type{});
    return (int)cudaGetLastError();
}

// 0 launched, >0 cudaError_t, -1 configuration not covered (caller uses the table-driven kernel)
template <class G>
int boosted_launch(const DecodeArgs &a, int sm_count, cudaStream_t st) {
    const bool qms5 = a.decoder_type == 2 && a.qbit == 5, ms = a.decoder_type == 1;
    if (!(qms5 || ms) || a.compute_ucn || a.llr_init || a.xin_init || a.xin_out || a.app_init) return -1;
    if (a.hist_v2c) return -1;      // training dumps: the training variant (nldpc_spec_train.cuh), tried first by the dispatcher
    const bool capturing = stream_is_capturing(st);      // CUDA graph capture: fixed arena range, no launch-time bookkeeping
    ConstArena &arena = arena_for_current_device();
    const int n_w = a.T * kWPitch<G>, len = (n_w + 1) / 2;      // plain floats in the arena (wb_at<true>), even pitch per iteration; len in float2 units
    cudaError_t err = cudaSuccess;
    const int off = capturing ? arena.acquire_captured(len, st, &err) : arena.acquire(len, st, &err);
    if (err != cudaSuccess) return (int)err;
    if (off < 0) return -1;
    DecodeArgs args = a;
    args.wb_off = 2 * off;                               // float units
    if ((err = upload_w_pitched(arena, a.w, off, a.T, G::E, kWPitch<G>, st)) != cudaSuccess) return (int)err;   // cn_w (or 1.0)
    int rc;
    if (ms) rc = a.vn_w ? boosted_launch_one<G, 1, true>(args, sm_count, st) : boosted_launch_one<G, 1, false>(args, sm_count, st);
    else rc = a.vn_w ? boosted_launch_one<G, 2, true>(args, sm_count, st) : boosted_launch_one<G, 2, false>(args, sm_count, st);
    if (rc != 0 || capturing) return rc;
    return (int)arena.release_after(off, len, st);
}

}  // namespace
}  // namespace nldpc
