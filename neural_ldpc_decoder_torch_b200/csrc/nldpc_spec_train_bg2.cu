// training variant of the specialised Boosted forward + the Boosted backward sweeps for 5G NR BG2 set 0, Z = 16
#include "generated/nldpc_graph_bg2z16.cuh"
#include "nldpc_spec_train.cuh"
#include "nldpc_spec_backward.cuh"
#include "nldpc_spec.cuh"
namespace nldpc {
int spec_train_prepare_bg2() { return train_prepare<gen::Bg2Z16>(); }
int spec_train_launch_bg2(const DecodeArgs &a, int sm_count, cudaStream_t st) { return train_launch<gen::Bg2Z16>(a, 0, sm_count, st); }
int spec_boosted_backward_bg2(const BwdArgs &a, int sm_count, cudaStream_t st) { return spec_bwd_launch<gen::Bg2Z16, true>(a, 0, sm_count, st); }
}  // namespace nldpc
