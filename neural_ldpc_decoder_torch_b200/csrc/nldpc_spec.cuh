// nldpc_spec.cuh — registry of the specialised (fully unrolled, graph-as-immediates) kernels.
#pragma once
#include "nldpc_common.cuh"

namespace nldpc {
// returns the registry index of a kernel specialised for exactly this base graph and lifting size, or -1
int spec_find(const int32_t *bg, int M, int N, int Z);
int spec_prepare(int id);                 // one-time cudaFuncSetAttribute; 0 or cudaError_t
int spec_cw_per_cta(int id);
int spec_threads(int id);
// 0 = launched, >0 = cudaError_t, <0 = configuration not covered (caller falls back to the generic kernel)
int spec_launch_neural(int id, const DecodeArgs &a, int sm_count, cudaStream_t st);
int spec_launch_boosted(int id, const DecodeArgs &a, int sm_count, cudaStream_t st);
// scratch the specialised backward needs: sm_count * spec_backward_scratch_rows(id) * kSpecBwdScratchLanes floats
constexpr int kSpecBwdScratchLanes = 384;
int spec_backward_scratch_rows(int id);   // chain-state rows per CTA (>= 1), 0 if there is no specialised backward
// Will spec_launch_backward take this configuration?  (mode 0 Neural / 1 MS / 2 QMS.)  The training-mode forward asks before it
// writes its dump: the specialised sweep reads check-packed records (DecodeArgs::hist_fmt 1), the table-driven one slot-major rows.
bool spec_backward_covers(int id, int mode, int T, bool has_cn_w, bool has_vn_w, bool ucn, int qbit);
size_t spec_dump_bytes_per_cw_iter(int id, int mode);   // check-packed records of one codeword and iteration
int spec_launch_backward(int id, const BwdArgs &a, int sm_count, cudaStream_t st);   // 0 / cudaError_t / -1 not covered
// per-code translation units
int spec_boosted_prepare_bg2();
int spec_boosted_launch_bg2(const DecodeArgs &a, int sm_count, cudaStream_t st);
int spec_boosted_prepare_wimax();
int spec_boosted_launch_wimax(const DecodeArgs &a, int sm_count, cudaStream_t st);
int spec_train_prepare_bg2();
int spec_train_launch_bg2(const DecodeArgs &a, int sm_count, cudaStream_t st);
int spec_train_prepare_wimax();
int spec_train_launch_wimax(const DecodeArgs &a, int sm_count, cudaStream_t st);
int spec_boosted_backward_bg2(const BwdArgs &a, int sm_count, cudaStream_t st);
int spec_boosted_backward_wimax(const BwdArgs &a, int sm_count, cudaStream_t st);
}  // namespace nldpc
