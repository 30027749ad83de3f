// nldpc_generic.cu — table-driven decode kernel for ANY base graph / lifting size that fits on chip.
//
// One thread per (codeword, lane z); a CTA holds `cw_per_cta` codeword slabs in shared memory for all
// T iterations.  Node degrees are dispatched (warp-uniformly) to fully unrolled templates so the
// exact-order fp32 sums of the VN update live in registers.  The specialised kernels generated for the
// built-in graphs (csrc/generated/) follow the same scheme with every table folded into immediates.
//
// Exactness contract (SURVEY.md Appendix A / oracle/nldpc_oracle.c): every + and * is a single fp32
// rounding in the reference's order; min / compare / sign logic is exact by construction.
#include "nldpc_generic_common.cuh"

namespace nldpc {


// CN update of one check of degree D, NeuralLDPCDecoder.py:66-91, in the check-lane domain h.
template <int D>
__device__ __forceinline__ void cn_check_neural(float *__restrict__ slab, int h, const GraphDev &g, int e0,
                                                const float *__restrict__ wt, const float *__restrict__ bt,
                                                const EmitCtx &ec) {
    float u[D];
    int addr[D];
    unsigned par = 0;
#pragma unroll
    for (int k = 0; k < D; k++) {
        int zz = h + __ldg(g.e_shift + e0 + k);
        zz = (zz >= g.Z) ? zz - g.Z : zz;             // gather u[h] = v2c[(h+s) mod Z]   (:59-63)
        addr[k] = __ldg(g.e_row + e0 + k) * g.Z + zz;
        u[k] = slab[addr[k]];
        par ^= (u[k] > 0.0f) ? 1u : 0u;               // parity of #positive (:77-79)
    }
    // min over the OTHER edges of |u| with exact zeros replaced by 10000, capped at 10000 (:74-75)
    float suf[D + 1];
    suf[D] = 10000.0f;
#pragma unroll
    for (int k = D - 1; k >= 0; k--) {
        float a = fabsf(u[k]);
        a = (a > 0.0f) ? a : 10000.0f;
        suf[k] = fminf(suf[k + 1], a);
    }
    float pre = 10000.0f;
#pragma unroll
    for (int k = 0; k < D; k++) {
        const float mag = fminf(pre, suf[k + 1]);
        float a = fabsf(u[k]);
        a = (a > 0.0f) ? a : 10000.0f;
        pre = fminf(pre, a);
        const unsigned npos_odd = par ^ ((u[k] > 0.0f) ? 1u : 0u);   // others' positives
        const int e = e0 + k;
        float m = addf(mulf(mag, __ldg(wt + e)), __ldg(bt + e));      // |o|*w + b  (:89)
        m = (m > 0.0f) ? m : 0.0f;                                    // ReLU (:90)
        const float c2v = npos_odd ? m : -m;                          // o = mag * (npos even ? -1 : +1) (:79-80, :91)
        const int j1 = __ldg(g.e_col1 + e);
        if (j1 < 0) {
            slab[addr[k]] = c2v;                                      // scatter back in place (:82-86)
        } else {
            // degree-1 variable block: marginal = xa + (0 + c2v)  (:94-98), lane z = (h+s) mod Z
            emit(ec, addr[k], addf(u[k], addf(0.0f, c2v)));           // addr == j*Z + z because row j < N
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
nldpc_generic_neural_kernel(const GraphDev g, const DecodeArgs a, const int cw_per_cta, const int use_tma) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *slabs = reinterpret_cast<float *>(smem_raw);
    const int NZ = g.N * g.Z;
    const int hwords = (NZ + 31) >> 5;
    uint32_t *hbits_all = reinterpret_cast<uint32_t *>(slabs + (size_t)cw_per_cta * g.slab_stride);
    uint64_t *bar = reinterpret_cast<uint64_t *>(hbits_all + (((size_t)cw_per_cta * hwords + 1) & ~(size_t)1));

    const int tid = threadIdx.x;
    const int L = cw_per_cta * g.Z;
    const int cw = tid / g.Z;
    const int z = tid - cw * g.Z;
    float *slab = slabs + (size_t)cw * g.slab_stride;
    const int n_tiles = (a.B + cw_per_cta - 1) / cw_per_cta;
    const int nb = (NZ + 7) >> 3;

    if (use_tma && tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b0 = tile * cw_per_cta;
        const int ncw = min(cw_per_cta, a.B - b0);
        const bool active = (tid < L) && (cw < ncw);
        // ---- load channel LLRs (one 1-D bulk TMA per codeword) and clear the message slots ----
        if (use_tma) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(bar, (uint32_t)(ncw * NZ * sizeof(float)));
                for (int c = 0; c < ncw; c++)
                    tma_load_1d(slabs + (size_t)c * g.slab_stride, a.xa + (size_t)(b0 + c) * NZ, (uint32_t)(NZ * sizeof(float)), bar);
            }
        } else {
            for (int i = tid; i < ncw * NZ; i += blockDim.x) {
                const int c = i / NZ, q = i - c * NZ;
                slabs[(size_t)c * g.slab_stride + q] = __ldg(a.xa + (size_t)(b0 + c) * NZ + q);
            }
        }
        for (int i = tid; i < cw_per_cta * g.S * g.Z; i += blockDim.x) {
            const int c = i / (g.S * g.Z), q = i - c * (g.S * g.Z);
            slabs[(size_t)c * g.slab_stride + NZ + q] = 0.0f;                        // llr = zeros (:49)
        }
        for (int i = tid; i < cw_per_cta * hwords; i += blockDim.x) hbits_all[i] = 0u;
        if (use_tma) { mbar_wait(bar, phase); phase ^= 1; }
        __syncthreads();

        for (int t = 0; t < a.T; t++) {
            // ---------------- VN phase (+ marginal of iteration t-1 for blocks of degree >= 2) ----------------
            {
                const bool soft_prev = t > 0 && a.soft_mode == 1;                    // NLDPC_OUT_ALL
                const bool hard_prev = t > 0 && a.hard_mode == 1;
                EmitCtx ec;
                ec.soft = soft_prev ? a.soft + (size_t)(t - 1) * a.B * NZ : nullptr;
                ec.hbits = hard_prev ? hbits_all + (size_t)cw * hwords : nullptr;
                ec.cw_off = (size_t)(b0 + cw) * NZ;
                if (active) {
                    for (int c = 0; c < g.n_vcols; c++) {
                        const int j = __ldg(g.vcol_j + c);
                        const int p0 = __ldg(g.vcol_ptr + c);
                        const int d = __ldg(g.vcol_ptr + c + 1) - p0;
                        const float x = slab[j * g.Z + z];
                        float tot = 0.0f;
#define NLDPC_VN_CASE(D) tot = vn_block<D>(slab + z, g.Z, g.vcol_row + p0, x)
                        NLDPC_DEG_SWITCH(d, NLDPC_VN_CASE)
#undef NLDPC_VN_CASE
                        if (soft_prev || hard_prev) emit(ec, j * g.Z + z, addf(x, tot));   // out = xa + tot (:96)
                    }
                    if (a.hist_v2c) {   // training dump: the v2c every CN phase reads (nldpc_backward.cu)
                        float *hv = a.hist_v2c + (((size_t)t * a.B + (b0 + cw)) * g.S) * g.Z + z;
                        for (int s = 0; s < g.S; s++) hv[(size_t)s * g.Z] = slab[(g.N + s) * g.Z + z];
                    }
                }
                __syncthreads();
                if (hard_prev) {   // flush iteration t-1's packed decisions, then clear for iteration t
                    uint8_t *dst = a.hard + ((size_t)(t - 1) * a.B + b0) * nb;
                    for (int i = tid; i < ncw * nb; i += blockDim.x) {
                        const int c = i / nb, q = i - c * nb;
                        dst[(size_t)c * nb + q] = reinterpret_cast<const uint8_t *>(hbits_all + (size_t)c * hwords)[q];
                    }
                    __syncthreads();
                    for (int i = tid; i < cw_per_cta * hwords; i += blockDim.x) hbits_all[i] = 0u;
                    __syncthreads();
                }
            }
            // ---------------- CN phase (+ marginal of iteration t for degree-1 blocks) ----------------
            {
                const bool last = (t == a.T - 1);
                const bool soft_now = a.soft_mode == 1 || (a.soft_mode == 2 && last);
                const bool hard_now = a.hard_mode == 1 || (a.hard_mode == 2 && last);
                EmitCtx ec;
                ec.soft = soft_now ? a.soft + (a.soft_mode == 1 ? (size_t)t * a.B * NZ : 0) : nullptr;
                ec.hbits = hard_now ? hbits_all + (size_t)cw * hwords : nullptr;
                ec.cw_off = (size_t)(b0 + cw) * NZ;
                const float *wt = a.w + (size_t)t * g.E, *bt = a.b + (size_t)t * g.E;
                if (active) {
                    for (int i = 0; i < g.M; i++) {
                        const int e0 = __ldg(g.row_ptr + i);
                        const int d = __ldg(g.row_ptr + i + 1) - e0;
#define NLDPC_CN_CASE(D) cn_check_neural<D>(slab, z, g, e0, wt, bt, ec)
                        NLDPC_DEG_SWITCH(d, NLDPC_CN_CASE)
#undef NLDPC_CN_CASE
                    }
                }
                __syncthreads();
            }
        }
        // ---------------- final marginal (iteration T-1) of the blocks of degree >= 2 ----------------
        {
            const bool soft_now = a.soft_mode != 0, hard_now = a.hard_mode != 0;
            EmitCtx ec;
            ec.soft = soft_now ? a.soft + (a.soft_mode == 1 ? (size_t)(a.T - 1) * a.B * NZ : 0) : nullptr;
            ec.hbits = hard_now ? hbits_all + (size_t)cw * hwords : nullptr;
            ec.cw_off = (size_t)(b0 + cw) * NZ;
            if (active && (soft_now || hard_now)) {
                for (int c = 0; c < g.n_vcols; c++) {
                    const int j = __ldg(g.vcol_j + c);
                    const int p0 = __ldg(g.vcol_ptr + c), p1 = __ldg(g.vcol_ptr + c + 1);
                    float tot = 0.0f;
                    for (int k = p0; k < p1; k++) tot = addf(tot, slab[__ldg(g.vcol_row + k) * g.Z + z]);
                    emit(ec, j * g.Z + z, addf(slab[j * g.Z + z], tot));
                }
            }
            __syncthreads();
            if (hard_now) {
                uint8_t *dst = a.hard + ((a.hard_mode == 1 ? (size_t)(a.T - 1) * a.B : 0) + b0) * nb;
                for (int i = tid; i < ncw * nb; i += blockDim.x) {
                    const int c = i / nb, q = i - c * nb;
                    dst[(size_t)c * nb + q] = reinterpret_cast<const uint8_t *>(hbits_all + (size_t)c * hwords)[q];
                }
            }
            __syncthreads();
        }
    }
}


int generic_prepare(size_t smem_bytes) {
    (void)smem_bytes;   // several graphs may coexist: always allow the full 227 KB
    return (int)cudaFuncSetAttribute(nldpc_generic_neural_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
}

int generic_launch_neural(const GraphDev &g, const DecodeArgs &a, int cw_per_cta, int threads, size_t smem_bytes, int use_tma,
                          int grid, cudaStream_t st) {
    nldpc_generic_neural_kernel<<<grid, threads, smem_bytes, st>>>(g, a, cw_per_cta, use_tma);
    return (int)cudaGetLastError();
}

}  // namespace nldpc
