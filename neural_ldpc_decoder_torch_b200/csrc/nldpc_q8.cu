// nldpc_q8.cu — narrow channel-LLR transports of the host-buffer APIs: one-byte codes x = scale * q (q int8) and fp16 values.
// The Boosted pipeline quantises its channel LLRs before they reach the decoder (boosted AWGNPassedDatagen.py:165-166,
// Functions.Cal_MSA_Q, Functions.py:70-83: multiples of 0.5 in +-7.5 for q_bit = 5), so an int8 code per LLR carries them
// without loss and the host -> device transfer — the bound of the end-to-end path — shrinks 4x.  This kernel expands a
// chunk to the fp32 layout the decode kernels stage with TMA (1 B read + 4 B written per LLR; HBM bound, ~1 % of a decode).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nldpc {

__global__ void __launch_bounds__(256) q8_to_f32_kernel(const int8_t *__restrict__ in, float *__restrict__ out, size_t n, float scale) {
    const size_t n16 = n / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const int4 v = __ldcs(reinterpret_cast<const int4 *>(in) + i);
        const int w[4] = {v.x, v.y, v.z, v.w};
        float4 *o = reinterpret_cast<float4 *>(out) + 4 * i;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float4 f;
            f.x = scale * (float)(int8_t)(w[k] & 0xff);
            f.y = scale * (float)(int8_t)((w[k] >> 8) & 0xff);
            f.z = scale * (float)(int8_t)((w[k] >> 16) & 0xff);
            f.w = scale * (float)(int8_t)((w[k] >> 24) & 0xff);
            o[k] = f;
        }
    }
    for (size_t i = n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = scale * (float)in[i];
}

// fp16 channel LLRs -> the fp32 layout the decode kernels stage (exact: every fp16 value is an fp32 value)
__global__ void __launch_bounds__(256) f16_to_f32_kernel(const __half *__restrict__ in, float *__restrict__ out, size_t n) {
    const size_t n8 = n / 8;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(in) + i);
        const __half2 *h = reinterpret_cast<const __half2 *>(&v);
        float4 *o = reinterpret_cast<float4 *>(out) + 2 * i;
        const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
        o[0] = make_float4(a.x, a.y, b.x, b.y);
        o[1] = make_float4(c.x, c.y, d.x, d.y);
    }
    for (size_t i = n8 * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = __half2float(in[i]);
}

int launch_f16_to_f32(const void *in, float *out, size_t n, int sm_count, cudaStream_t st) {
    if (n == 0) return 0;
    if ((((uintptr_t)in | (uintptr_t)out) % 16) != 0) return (int)cudaErrorMisalignedAddress;
    const size_t want = (n / 8 + 255) / 256 + 1;
    const int grid = (int)(want < (size_t)sm_count * 8 ? want : (size_t)sm_count * 8);
    f16_to_f32_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __half *>(in), out, n);
    return (int)cudaGetLastError();
}

int launch_q8_to_f32(const int8_t *in, float *out, size_t n, float scale, int sm_count, cudaStream_t st) {
    if (n == 0) return 0;
    const bool aligned = (((uintptr_t)in | (uintptr_t)out) % 16) == 0;
    if (!aligned) return (int)cudaErrorMisalignedAddress;
    const size_t want = (n / 16 + 255) / 256 + 1;
    const int grid = (int)(want < (size_t)sm_count * 8 ? want : (size_t)sm_count * 8);
    q8_to_f32_kernel<<<grid, 256, 0, st>>>(in, out, n, scale);
    return (int)cudaGetLastError();
}

}  // namespace nldpc
