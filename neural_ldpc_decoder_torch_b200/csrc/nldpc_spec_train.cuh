// nldpc_spec_train.cuh — instantiation helper for the training variant of the specialised Boosted forward (MS / QMS q=5, no
// UCN): every-iteration outputs + check-packed training dump [+ fused multi-iteration BCE], extension checks as descriptor
// loops (Train<G>, nldpc_spec_kernel.cuh).  One translation unit per code, together with that code's Boosted backward sweeps.
#pragma once
#include <algorithm>
#include <type_traits>

#include "nldpc_spec_host.cuh"

namespace nldpc {
namespace {

template <class G, int MODE, bool kXo>
int train_prepare_one() {
    return (int)set_smem(nldpc_spec_neural_kernel<G, true, true, MODE, kXo, true>, KernelCfg<G, true, kXo, true>::type::kSmemBytes);
}

template <class G>
int train_prepare() {
    int rc;
    if ((rc = train_prepare_one<G, 1, false>())) return rc;
    if ((rc = train_prepare_one<G, 1, true>())) return rc;
    if ((rc = train_prepare_one<G, 2, false>())) return rc;
    if ((rc = train_prepare_one<G, 2, true>())) return rc;
    return 0;
}

template <class G, int MODE, bool kXo>
int train_launch_one(const DecodeArgs &args, int sm_count, cudaStream_t st) {
    using Cfg = typename KernelCfg<G, true, kXo, true>::type;
    const int n_units = (args.B + Cfg::Shape::kCw - 1) / Cfg::Shape::kCw;
    const int grid = std::min(n_units, sm_count * Cfg::kCtasPerSm);
    nldpc_spec_neural_kernel<G, true, true, MODE, kXo, true><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(args);
    return (int)cudaGetLastError();
}

// 0 launched, >0 cudaError_t, -1 configuration not covered (caller uses the table-driven kernel)
template <class G>
int train_launch(const DecodeArgs &a, int graph_slot, int sm_count, cudaStream_t st) {
    const bool qms5 = a.decoder_type == 2 && a.qbit == 5, ms = a.decoder_type == 1;
    if (!(qms5 || ms) || a.compute_ucn || a.llr_init || a.xin_init || a.xin_out || a.app_init) return -1;
    // the training variant writes check-packed records only, every-iteration soft outputs (or dL/dout), no hard decisions
    if (!a.hist_v2c || a.hist_fmt != 1 || a.soft_mode != 1 || a.hard_mode != 0) return -1;
    if (a.ybits && (!a.coef || !a.loss_acc)) return -1;
    const bool capturing = stream_is_capturing(st);      // CUDA graph capture: fixed arena range, no launch-time bookkeeping
    cudaError_t err = ensure_loop_desc<G>(graph_slot, capturing);
    if (err == cudaErrorStreamCaptureUnsupported && capturing) return -1;      // first use inside a capture
    if (err != cudaSuccess) return (int)err;
    ConstArena &arena = arena_for_current_device();
    const int n_w = a.T * kWPitch<G>, len = (n_w + 1) / 2;      // plain floats in the arena (wb_at<true>), even pitch per iteration; len in float2 units
    const int off = capturing ? arena.acquire_captured(len, st, &err) : arena.acquire(len, st, &err);
    if (err != cudaSuccess) return (int)err;
    if (off < 0) return -1;
    DecodeArgs args = a;
    args.wb_off = 2 * off;                               // float units
    args.desc_base = graph_slot * kDescStride;
    if ((err = upload_w_pitched(arena, a.w, off, a.T, G::E, kWPitch<G>, st)) != cudaSuccess) return (int)err;   // cn_w (or 1.0)
    int rc;
    if (ms) rc = a.vn_w ? train_launch_one<G, 1, true>(args, sm_count, st) : train_launch_one<G, 1, false>(args, sm_count, st);
    else rc = a.vn_w ? train_launch_one<G, 2, true>(args, sm_count, st) : train_launch_one<G, 2, false>(args, sm_count, st);
    if (rc != 0 || capturing) return rc;
    return (int)arena.release_after(off, len, st);
}

}  // namespace
}  // namespace nldpc
