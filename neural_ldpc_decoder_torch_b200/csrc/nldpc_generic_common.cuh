// nldpc_generic_common.cuh — pieces shared by the table-driven kernels (Neural and Boosted).
#pragma once
#include "nldpc_common.cuh"

namespace nldpc {

#define NLDPC_DEG_SWITCH(d, F)                                                                                     \
    switch (d) {                                                                                                   \
        case 1: F(1); break;   case 2: F(2); break;   case 3: F(3); break;   case 4: F(4); break;                  \
        case 5: F(5); break;   case 6: F(6); break;   case 7: F(7); break;   case 8: F(8); break;                  \
        case 9: F(9); break;   case 10: F(10); break; case 11: F(11); break; case 12: F(12); break;                \
        case 13: F(13); break; case 14: F(14); break; case 15: F(15); break; case 16: F(16); break;                \
        case 17: F(17); break; case 18: F(18); break; case 19: F(19); break; case 20: F(20); break;                \
        case 21: F(21); break; case 22: F(22); break; case 23: F(23); break; case 24: F(24); break;                \
        case 25: F(25); break; case 26: F(26); break; case 27: F(27); break; case 28: F(28); break;                \
        case 29: F(29); break; case 30: F(30); break; case 31: F(31); break; case 32: F(32); break;                \
        default: break;                                                                                            \
    }

// VN update of one variable block of degree D (NeuralLDPCDecoder.py:56-58):
//   v2c[k] = x + (((0 + c[0]) + c[1]) + ... skipping k ...), ascending check row; returns the full
//   sequential column total (the `llr @ W_output` marginal of the PREVIOUS iteration, :94).
template <int D>
__device__ __forceinline__ float vn_block(float *__restrict__ slabz, int Z, const int *__restrict__ rows, float x) {
    float c[D];
    int off[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
        off[k] = __ldg(rows + k) * Z;
        c[k] = slabz[off[k]];
    }
    float s[D];
    float p = 0.0f;
#pragma unroll
    for (int k = 0; k < D; k++) {
        s[k] = p;
        p = addf(p, c[k]);
    }
#pragma unroll
    for (int k = 0; k < D; k++) {
#pragma unroll
        for (int m = k + 1; m < D; m++) s[k] = addf(s[k], c[m]);
    }
#pragma unroll
    for (int k = 0; k < D; k++) slabz[off[k]] = addf(x, s[k]);
    return p;
}

struct EmitCtx {
    float *soft;     // base of this iteration's [B][NZ] block for codeword 0, or nullptr
    uint32_t *hbits; // this codeword's packed-bit words in shared memory, or nullptr
    size_t cw_off;   // b * NZ
};

__device__ __forceinline__ void emit(const EmitCtx &ec, int q, float v) {
    if (ec.soft) st_global_stream(ec.soft + ec.cw_off + q, v);
    if (ec.hbits && v < 0.0f) atomicOr(ec.hbits + (q >> 5), 1u << (q & 31));
}

}  // namespace nldpc
