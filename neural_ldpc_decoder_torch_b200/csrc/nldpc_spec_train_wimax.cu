// training variant of the specialised Boosted forward + the Boosted backward sweeps for 802.16e N=576 R=3/4, Z = 24
#include "generated/nldpc_graph_wimaxz24.cuh"
#include "nldpc_spec_train.cuh"
#include "nldpc_spec_backward.cuh"
#include "nldpc_spec.cuh"
namespace nldpc {
int spec_train_prepare_wimax() { return train_prepare<gen::WimaxZ24>(); }
int spec_train_launch_wimax(const DecodeArgs &a, int sm_count, cudaStream_t st) { return train_launch<gen::WimaxZ24>(a, 1, sm_count, st); }
int spec_boosted_backward_wimax(const BwdArgs &a, int sm_count, cudaStream_t st) { return spec_bwd_launch<gen::WimaxZ24, true>(a, 1, sm_count, st); }
}  // namespace nldpc
