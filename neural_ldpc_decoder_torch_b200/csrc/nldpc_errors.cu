// nldpc_errors.cu — per-iteration bit / frame error counts of a decode
// (Functions.evaluate_ber_fer, /root/reference/src/boosted_neural_ldpc_decoder/Functions.py:86-102):
//     bit_t   = #{ ((out_t[c][i] < 0) ? 1.0 : 0.0) != y[c][i] }          (:90, :93-94)
//     frame_t = #{ c : codeword c has at least one such position }       (:98-99)
// The reference makes 4 elementwise passes and 2 reductions per iteration over [B, N*Z]; here every soft output is read
// ONCE (16-byte streaming loads), the label row of a codeword sits in registers across the T iterations, and the counts
// are exact 64-bit integers.  HBM bound: 4*N*Z*(T + 1) bytes per codeword.
// The packed variant counts on the packed hard decisions the decode kernels write (bit i of byte i/8 = out[i] < 0).
#include <cuda_runtime.h>
#include <stdint.h>

namespace nldpc {

constexpr int kErrThreads = 256;
constexpr int kErrMaxT = 256;      // per-CTA shared counters; larger T is served in slices by the launcher

__device__ __forceinline__ unsigned wrong4(const float4 x, const float4 y) {
    return (unsigned)(((x.x < 0.0f) ? 1.0f : 0.0f) != y.x) + (unsigned)(((x.y < 0.0f) ? 1.0f : 0.0f) != y.y) +
           (unsigned)(((x.z < 0.0f) ? 1.0f : 0.0f) != y.z) + (unsigned)(((x.w < 0.0f) ? 1.0f : 0.0f) != y.w);
}

__device__ __forceinline__ void flush_counts(const unsigned long long *sh, int T, unsigned long long *counts, int t_total, int t0) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) {
        const unsigned long long v = sh[i];
        if (v) atomicAdd(counts + (size_t)(i / T) * t_total + t0 + (i % T), v);
    }
}

// One warp per codeword (grid-stride).  KV = float4 groups per lane (ceil(NZ / 128)); rows must be 16 B aligned.
template <int KV>
__global__ void __launch_bounds__(kErrThreads) count_errors_vec_kernel(const float *__restrict__ soft, size_t iter_stride,
                                                                        const float *__restrict__ y, int T, int B, int NZ,
                                                                        unsigned long long *__restrict__ counts, int t_total, int t0) {
    __shared__ unsigned long long sh[2 * kErrMaxT];
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) sh[i] = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int n4 = NZ >> 2;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < B; c += warps) {
        const float4 *yr = reinterpret_cast<const float4 *>(y + (size_t)c * NZ);
        float4 yv[KV];
#pragma unroll
        for (int k = 0; k < KV; k++) {
            const int i = lane + 32 * k;
            yv[k] = (i < n4) ? __ldcs(yr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int t = 0; t < T; t++) {
            const float4 *xr = reinterpret_cast<const float4 *>(soft + (size_t)t * iter_stride + (size_t)c * NZ);
            float4 xv[KV];
#pragma unroll
            for (int k = 0; k < KV; k++) {
                const int i = lane + 32 * k;
                xv[k] = (i < n4) ? __ldcs(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);      // padding: (0 < 0) = 0 == label 0
            }
            unsigned cnt = 0;
#pragma unroll
            for (int k = 0; k < KV; k++) cnt += wrong4(xv[k], yv[k]);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && cnt) {
                atomicAdd(&sh[t], (unsigned long long)cnt);
                atomicAdd(&sh[T + t], 1ull);
            }
        }
    }
    flush_counts(sh, T, counts, t_total, t0);
}

// Any N*Z / alignment: scalar loads, labels re-read per iteration (L1 / L2 hits).
__global__ void __launch_bounds__(kErrThreads) count_errors_scalar_kernel(const float *__restrict__ soft, size_t iter_stride,
                                                                           const float *__restrict__ y, int T, int B, int NZ,
                                                                           unsigned long long *__restrict__ counts, int t_total, int t0) {
    __shared__ unsigned long long sh[2 * kErrMaxT];
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) sh[i] = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < B; c += warps) {
        const float *yr = y + (size_t)c * NZ;
        for (int t = 0; t < T; t++) {
            const float *xr = soft + (size_t)t * iter_stride + (size_t)c * NZ;
            unsigned cnt = 0;
            for (int i = lane; i < NZ; i += 32) cnt += (unsigned)(((__ldcs(xr + i) < 0.0f) ? 1.0f : 0.0f) != __ldg(yr + i));
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && cnt) {
                atomicAdd(&sh[t], (unsigned long long)cnt);
                atomicAdd(&sh[T + t], 1ull);
            }
        }
    }
    flush_counts(sh, T, counts, t_total, t0);
}

// Packed decisions: hard [T][B][hb] bytes (hb = ceil(NZ / 8)), labels packed the same way or NULL (= all-zero codeword).
// Bits past N*Z in the last byte are ignored.
__global__ void __launch_bounds__(kErrThreads) count_errors_packed_kernel(const uint8_t *__restrict__ hard, size_t iter_stride,
                                                                           const uint8_t *__restrict__ yp, int T, int B, int NZ,
                                                                           unsigned long long *__restrict__ counts, int t_total, int t0) {
    __shared__ unsigned long long sh[2 * kErrMaxT];
    for (int i = threadIdx.x; i < 2 * T; i += blockDim.x) sh[i] = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int hb = (NZ + 7) >> 3;
    const unsigned tail = (NZ & 7) ? ((1u << (NZ & 7)) - 1u) : 0xffu;
    const bool words = (hb % 4 == 0) && (((uintptr_t)hard | (uintptr_t)yp | (uintptr_t)iter_stride) % 4 == 0);
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < B; c += warps) {
        for (int t = 0; t < T; t++) {
            const uint8_t *hr = hard + (size_t)t * iter_stride + (size_t)c * hb;
            const uint8_t *yr = yp ? yp + (size_t)c * hb : nullptr;
            unsigned cnt = 0;
            if (words) {
                const int nw = hb >> 2;
                for (int i = lane; i < nw; i += 32) {
                    unsigned v = __ldcs(reinterpret_cast<const unsigned *>(hr) + i);
                    if (yr) v ^= __ldg(reinterpret_cast<const unsigned *>(yr) + i);
                    if (i == nw - 1) v &= (tail << 24) | 0x00ffffffu;
                    cnt += __popc(v);
                }
            } else {
                for (int i = lane; i < hb; i += 32) {
                    unsigned v = hr[i];
                    if (yr) v ^= yr[i];
                    if (i == hb - 1) v &= tail;
                    cnt += __popc(v);
                }
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && cnt) {
                atomicAdd(&sh[t], (unsigned long long)cnt);
                atomicAdd(&sh[T + t], 1ull);
            }
        }
    }
    flush_counts(sh, T, counts, t_total, t0);
}

static int err_grid(int B, int sm_count) {
    const long long want = ((long long)B * 32 + kErrThreads - 1) / kErrThreads;       // one warp per codeword
    const long long cap = (long long)sm_count * 8;                                      // 8 CTAs x 8 warps = 64 warps per SM
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <int KV>
static void launch_vec(const float *soft, size_t stride, const float *y, int T, int B, int NZ, unsigned long long *counts, int t_total,
                       int t0, int grid, cudaStream_t st) {
    count_errors_vec_kernel<KV><<<grid, kErrThreads, 0, st>>>(soft, stride, y, T, B, NZ, counts, t_total, t0);
}

// counts [2][T]: OVERWRITTEN.  Returns a cudaError_t value.
int launch_count_errors(const float *soft, size_t iter_stride, const float *y, int T, int B, int NZ, unsigned long long *counts,
                        int sm_count, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 2 * (size_t)T, st);
    if (e != cudaSuccess) return (int)e;
    if (B == 0 || NZ == 0) return 0;
    const int grid = err_grid(B, sm_count);
    const bool vec = (NZ % 4 == 0) && (iter_stride % 4 == 0) && (((uintptr_t)soft | (uintptr_t)y) % 16 == 0) && (NZ <= 8 * 128);
    for (int t0 = 0; t0 < T; t0 += kErrMaxT) {
        const int tn = (T - t0 < kErrMaxT) ? T - t0 : kErrMaxT;
        const float *s = soft + (size_t)t0 * iter_stride;
        if (!vec) {
            count_errors_scalar_kernel<<<grid, kErrThreads, 0, st>>>(s, iter_stride, y, tn, B, NZ, counts, T, t0);
            continue;
        }
        switch ((NZ + 127) / 128) {
            case 1: launch_vec<1>(s, iter_stride, y, tn, B, NZ, counts, T, t0, grid, st); break;
            case 2: launch_vec<2>(s, iter_stride, y, tn, B, NZ, counts, T, t0, grid, st); break;
            case 3: launch_vec<3>(s, iter_stride, y, tn, B, NZ, counts, T, t0, grid, st); break;
            case 4: launch_vec<4>(s, iter_stride, y, tn, B, NZ, counts, T, t0, grid, st); break;
            case 5: launch_vec<5>(s, iter_stride, y, tn, B, NZ, counts, T, t0, grid, st); break;
            case 6: launch_vec<6>(s, iter_stride, y, tn, B, NZ, counts, T, t0, grid, st); break;
            case 7: launch_vec<7>(s, iter_stride, y, tn, B, NZ, counts, T, t0, grid, st); break;
            default: launch_vec<8>(s, iter_stride, y, tn, B, NZ, counts, T, t0, grid, st); break;
        }
    }
    return (int)cudaGetLastError();
}

int launch_count_errors_packed(const uint8_t *hard, size_t iter_stride, const uint8_t *yp, int T, int B, int NZ,
                               unsigned long long *counts, int sm_count, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * 2 * (size_t)T, st);
    if (e != cudaSuccess) return (int)e;
    if (B == 0 || NZ == 0) return 0;
    const int grid = err_grid(B, sm_count);
    for (int t0 = 0; t0 < T; t0 += kErrMaxT) {
        const int tn = (T - t0 < kErrMaxT) ? T - t0 : kErrMaxT;
        count_errors_packed_kernel<<<grid, kErrThreads, 0, st>>>(hard + (size_t)t0 * iter_stride, iter_stride, yp, tn, B, NZ, counts, T, t0);
    }
    return (int)cudaGetLastError();
}

}  // namespace nldpc
