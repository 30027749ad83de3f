// nldpc_loss.cu — fused multi-iteration BCE-with-logits loss and its gradient
// (LDPCDecoderLoss.forward, /root/reference/src/boosted_neural_ldpc_decoder/LDPCDecoderLoss.py:73-108, BCE branch):
//     L = sum_t c_t * mean_i bce(out_t[i], y[i]),   c_t = etha^{coeff_t} / sum_t etha^{coeff_t}
//     dL/dout_t[i] = c_t * (sigmoid(out_t[i]) - y[i]) / n
// One pass over the T iteration outputs (read once, gradient written once) instead of ~6 elementwise kernels per iteration.
#include <cuda_runtime.h>
#include <stdint.h>

namespace nldpc {

__device__ __forceinline__ float bce_term(float x, float yi, float c, float inv_n, float *g) {
    const float e = __expf(-fabsf(x));
    // binary_cross_entropy_with_logits: max(x,0) - x*y + log(1 + exp(-|x|))
    const float l = fmaxf(x, 0.0f) - x * yi + log1pf(e);
    const float r = __fdividef(1.0f, 1.0f + e);
    const float s = (x >= 0.0f) ? r : e * r;                              // sigmoid(x)
    *g = c * (s - yi) * inv_n;
    return c * l;
}

// VEC = 4: 16-byte loads / stores (n % 4 == 0 and 16 B aligned bases), else scalar.  The per-thread partial sum is fp32 over
// one element group (<= 4 T terms), then fp64 across groups.
template <int VEC>
__global__ void __launch_bounds__(256) multi_iter_bce_kernel(const float *__restrict__ soft, const float *__restrict__ y,
                                                             const float *__restrict__ coef, const float *__restrict__ gscale, int T,
                                                             size_t n, float inv_n_host, float *__restrict__ loss,
                                                             float *__restrict__ gout) {
    __shared__ double red[8];
    double acc = 0.0;
    const float inv_n = gscale ? inv_n_host * __ldg(gscale) : inv_n_host;      // upstream dL (a device scalar) folded into the gradient
    const size_t n_grp = n / VEC;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_grp; i += (size_t)gridDim.x * blockDim.x) {
        float part = 0.0f;
        if constexpr (VEC == 4) {
            const float4 yi = __ldg(reinterpret_cast<const float4 *>(y) + i);
            for (int t = 0; t < T; t++) {
                const float4 x = __ldcs(reinterpret_cast<const float4 *>(soft + (size_t)t * n) + i);
                const float c = __ldg(coef + t);
                float4 g;
                part += bce_term(x.x, yi.x, c, inv_n, &g.x);
                part += bce_term(x.y, yi.y, c, inv_n, &g.y);
                part += bce_term(x.z, yi.z, c, inv_n, &g.z);
                part += bce_term(x.w, yi.w, c, inv_n, &g.w);
                if (gout) __stcs(reinterpret_cast<float4 *>(gout + (size_t)t * n) + i, g);
            }
        } else {
            const float yi = __ldg(y + i);
            for (int t = 0; t < T; t++) {
                float g;
                part += bce_term(__ldcs(soft + (size_t)t * n + i), yi, __ldg(coef + t), inv_n, &g);
                if (gout) __stcs(gout + (size_t)t * n + i, g);
            }
        }
        acc += (double)part;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0 && loss) atomicAdd(loss, (float)(v * (double)inv_n_host));
    }
}

int launch_multi_iter_bce(const float *soft, const float *y, const float *coef, const float *gscale, int T, size_t n, float *loss,
                          float *gout, int sm_count, cudaStream_t st) {
    if (loss) {
        cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    if (n == 0) return 0;
    const bool vec = (n % 4 == 0) && (((uintptr_t)soft | (uintptr_t)y | (uintptr_t)gout) % 16 == 0);
    const size_t want = ((vec ? n / 4 : n) + 255) / 256;
    const int grid = (int)(want < (size_t)sm_count * 16 ? want : (size_t)sm_count * 16);
    if (vec) multi_iter_bce_kernel<4><<<grid, 256, 0, st>>>(soft, y, coef, gscale, T, n, (float)(1.0 / (double)n), loss, gout);
    else multi_iter_bce_kernel<1><<<grid, 256, 0, st>>>(soft, y, coef, gscale, T, n, (float)(1.0 / (double)n), loss, gout);
    return (int)cudaGetLastError();
}

// labels y [n_cw][NZ] fp32 -> bits [n_cw][ceil(NZ/8)], bit i of a codeword = (y[i] != 0), LSB first: the packing of the decode
// kernels' hard decisions, which is also how the fused training forward (DecodeArgs::ybits) reads its labels
__global__ void __launch_bounds__(256) pack_labels_kernel(const float *__restrict__ y, size_t n_cw, int NZ, int hb, uint8_t *__restrict__ bits) {
    const size_t total = n_cw * (size_t)hb;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t cw = i / hb;
        const int byte = (int)(i - cw * hb);
        const float *src = y + cw * (size_t)NZ + (size_t)byte * 8;
        unsigned v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (byte * 8 + k < NZ && __ldcs(src + k) != 0.0f) v |= 1u << k;
        bits[i] = (uint8_t)v;
    }
}

int launch_pack_labels(const float *y, size_t n_cw, int NZ, uint8_t *bits, cudaStream_t st) {
    const int hb = (NZ + 7) / 8;
    const size_t want = (n_cw * (size_t)hb + 255) / 256;
    const int grid = (int)(want < (size_t)148 * 16 ? want : (size_t)148 * 16);
    pack_labels_kernel<<<grid, 256, 0, st>>>(y, n_cw, NZ, hb, bits);
    return (int)cudaGetLastError();
}

}  // namespace nldpc
