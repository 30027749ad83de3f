// specialised Boosted kernels for 802.16e N=576 R=3/4, Z = 24
#include "generated/nldpc_graph_wimaxz24.cuh"
#include "nldpc_spec_boosted.cuh"
#include "nldpc_spec.cuh"
namespace nldpc {
int spec_boosted_prepare_wimax() { return boosted_prepare<gen::WimaxZ24>(); }
int spec_boosted_launch_wimax(const DecodeArgs &a, int sm_count, cudaStream_t st) { return boosted_launch<gen::WimaxZ24>(a, sm_count, st); }
}  // namespace nldpc
