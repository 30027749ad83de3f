// Instruction-fetch microbenchmark (sm_100a): a fully unrolled straight-line body of K independent-chain FFMAs is run
// `reps` times by every warp of a 1-CTA-per-SM grid; prints cycles per issued instruction per SM sub-partition versus
// the body size.  Answers: at which code size does a loop body stop being served by the near instruction caches, and
// what does each level cost?  (Used to size the unrolled LDPC iteration bodies; see DESIGN.md.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xptxas -O1 -o icache icache.cu && ./icache
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void __launch_bounds__(512, 1) body(float *out, int reps, long long *cycles) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    const float m = 1.0000001f, c = 1e-9f;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int k = 0; k < K / 4; k++) {
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a0) : "f"(m), "f"(c));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1) : "f"(m), "f"(c));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a2) : "f"(m), "f"(c));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a3) : "f"(m), "f"(c));
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int K>
void run(int warps, float *out, long long *cyc) {
    const int reps = (1 << 22) / K;     // ~4M instructions per warp
    body<K><<<148, warps * 32>>>(out, 2, cyc);
    cudaDeviceSynchronize();
    body<K><<<148, warps * 32>>>(out, reps, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; i++) s += (double)h[i];
    s /= 148;
    const double per_warp = s / ((double)reps * K);
    printf("body %6d instr (%4d KB)  warps/SM %2d : %.2f cycles/instr/warp  -> %.3f IPC per sub-partition\n", K, K * 16 / 1024, warps,
           per_warp, (warps / 4.0) / per_warp);
}

int main() {
    float *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&cyc, 148 * 8);
    for (int warps : {4, 8, 16}) {
        run<2048>(warps, out, cyc);
        run<3072>(warps, out, cyc);
        run<4096>(warps, out, cyc);
        run<5120>(warps, out, cyc);
        run<6144>(warps, out, cyc);
        run<10240>(warps, out, cyc);
        run<16384>(warps, out, cyc);
    }
    cudaError_t e = cudaGetLastError();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
