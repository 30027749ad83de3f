// Companion to icache.cu: do warps that stream DIFFERENT unrolled code (each far larger than the L1.5 instruction cache)
// still get the same per-warp instruction delivery as warps that share one stream?  Warp w runs copy (w % kStreams) of a
// K-instruction straight-line FFMA body.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xptxas -O1 -o icache_streams icache_streams.cu && ./icache_streams
#include <cstdio>
#include <cuda_runtime.h>

constexpr int K = 6144;

template <int COPY>
__device__ __forceinline__ void body(float &a0, float &a1, float &a2, float &a3) {
    const float m = 1.0000001f + COPY * 1e-7f, c = 1e-9f;
#pragma unroll
    for (int k = 0; k < K / 4; k++) {
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(m), "f"(c));
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(m), "f"(c));
    }
}

__global__ void __launch_bounds__(512, 1) run(float *out, int reps, int streams, long long *cycles) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    const int s = (threadIdx.x >> 5) % streams;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        switch (s) {
            case 0: body<0>(a0, a1, a2, a3); break;
            case 1: body<1>(a0, a1, a2, a3); break;
            case 2: body<2>(a0, a1, a2, a3); break;
            default: body<3>(a0, a1, a2, a3); break;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int reps = 600;
    for (int warps : {8, 16}) {
        for (int streams : {1, 2, 4}) {
            run<<<148, warps * 32>>>(out, 2, streams, cyc);
            cudaDeviceSynchronize();
            run<<<148, warps * 32>>>(out, reps, streams, cyc);
            cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double sum = 0;
            for (int i = 0; i < 148; i++) sum += (double)h[i];
            sum /= 148;
            const double per_warp = sum / ((double)reps * K);
            printf("body %d instr (%d KB) x %d distinct streams, warps/SM %2d : %.2f cycles/instr/warp -> %.3f IPC per sub-partition\n", K,
                   K * 16 / 1024, streams, warps, per_warp, (warps / 4.0) / per_warp);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
