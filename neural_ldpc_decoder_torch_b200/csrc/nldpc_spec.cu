// nldpc_spec.cu — specialised kernels (placeholder registry until the generated kernels land).
#include "nldpc_spec.cuh"
namespace nldpc {
int spec_find(const int32_t *, int, int, int) { return -1; }
int spec_prepare(int) { return 0; }
int spec_cw_per_cta(int) { return 0; }
int spec_threads(int) { return 0; }
int spec_launch_neural(int, const DecodeArgs &, int, cudaStream_t) { return -1; }
}  // namespace nldpc
