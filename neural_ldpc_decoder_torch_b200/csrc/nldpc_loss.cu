// nldpc_loss.cu — fused multi-iteration BCE-with-logits loss and its gradient
// (LDPCDecoderLoss.forward, /root/reference/src/boosted_neural_ldpc_decoder/LDPCDecoderLoss.py:73-108, BCE branch):
//     L = sum_t c_t * mean_i bce(out_t[i], y[i]),   c_t = etha^{coeff_t} / sum_t etha^{coeff_t}
//     dL/dout_t[i] = c_t * (sigmoid(out_t[i]) - y[i]) / n
// One pass over the T iteration outputs (read once, gradient written once) instead of ~6 elementwise kernels per iteration.
#include <cuda_runtime.h>
#include <stdint.h>

namespace nldpc {

__global__ void __launch_bounds__(256) multi_iter_bce_kernel(const float *__restrict__ soft, const float *__restrict__ y,
                                                             const float *__restrict__ coef, int T, size_t n, float inv_n,
                                                             float *__restrict__ loss, float *__restrict__ gout) {
    __shared__ double red[8];
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float yi = __ldg(y + i);
        for (int t = 0; t < T; t++) {
            const float x = __ldcs(soft + (size_t)t * n + i);
            const float c = __ldg(coef + t);
            const float e = expf(-fabsf(x));
            // binary_cross_entropy_with_logits: max(x,0) - x*y + log(1 + exp(-|x|))
            const float l = fmaxf(x, 0.0f) - x * yi + log1pf(e);
            acc += (double)(c * l);
            if (gout) {
                const float s = (x >= 0.0f) ? 1.0f / (1.0f + e) : e / (1.0f + e);      // sigmoid(x)
                __stcs(gout + (size_t)t * n + i, c * (s - yi) * inv_n);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) atomicAdd(loss, (float)(v * (double)inv_n));
    }
}

int launch_multi_iter_bce(const float *soft, const float *y, const float *coef, int T, size_t n, float *loss, float *gout, int sm_count,
                          cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    if (n == 0) return 0;
    const size_t want = (n + 255) / 256;
    const int grid = (int)(want < (size_t)sm_count * 16 ? want : (size_t)sm_count * 16);
    multi_iter_bce_kernel<<<grid, 256, 0, st>>>(soft, y, coef, T, n, (float)(1.0 / (double)n), loss, gout);
    return (int)cudaGetLastError();
}

}  // namespace nldpc
