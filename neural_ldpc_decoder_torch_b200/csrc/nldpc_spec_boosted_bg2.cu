// specialised Boosted kernels for 5G NR BG2 set 0, Z = 16
#include "generated/nldpc_graph_bg2z16.cuh"
#include "nldpc_spec_boosted.cuh"
#include "nldpc_spec.cuh"
namespace nldpc {
int spec_boosted_prepare_bg2() { return boosted_prepare<gen::Bg2Z16>(); }
int spec_boosted_launch_bg2(const DecodeArgs &a, int sm_count, cudaStream_t st) { return boosted_launch<gen::Bg2Z16>(a, sm_count, st); }
}  // namespace nldpc
