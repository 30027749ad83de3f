// nldpc_generic_boosted.cu — table-driven kernel for the loop body of BoostedNeuralLDPCDecoder.forward
// (/root/reference/src/boosted_neural_ldpc_decoder/BoostedNeuralLDPCDecoder.py:320-531) for ANY base graph / lifting size
// and every decoder type (SP / MS / QMS with all q-bit grids), weight-sharing folding and the UCN indicator.
//
// Same execution model as nldpc_generic.cu (one thread per (codeword, lane z), messages in shared memory for all
// executed iterations, in place).  Slab rows of Z floats per codeword:
//     [0,N)            xin  : the compounding channel input (xa_input: * vn_w, re-quantised every iteration, :325-337)
//     [N,N+S)          messages of edges whose variable block has degree >= 2
//     [N+S,N+S+N)      xo   : xa_origin (quantised once if QMS, :517-518)
//     [N+S+N,N+S+2N)   app  : previous iteration's output, only when the UCN indicator is evaluated (:339-346)
// Exactness: fp32, one rounding per operation, the reference's operation order (see oracle/nldpc_oracle.c).
#include <algorithm>

#include "nldpc_generic_common.cuh"

namespace nldpc {

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

// _quantize_message forward value (:187-214); rintf == torch.round (half to even)
__device__ __forceinline__ float quantf(float x, int qbit) {
    switch (qbit) {
        case 6: return clampf(rintf(x), -15.5f, 15.5f);
        case 5: return clampf(mulf(rintf(mulf(x, 2.0f)), 0.5f), -7.5f, 7.5f);
        case -5: return clampf(rintf(x), -15.0f, 15.0f);
        case 4: return clampf(rintf(x), -7.0f, 7.0f);
        case 3: return clampf(mulf(rintf(mulf(x, 0.5f)), 2.0f), -6.0f, 6.0f);
        default: return x;
    }
}

struct BoostedCtx {
    int decoder_type, qbit;      // NLDPC_DEC_*; decoder_qms_qbit
    float lo, hi;                // allowed_llr_range
    const float *cn_w, *ucn_w;   // this iteration's rows [E] or nullptr
    bool ucn_mix, compute_ucn;
    int app_row0;                // first slab row holding the APP used by the UCN indicator (xin rows at "t == 0")
    int xo_row0, app_store_row0; // xo rows; rows receiving this iteration's output (or -1)
    bool want_c2v1;              // compute the c2v of degree-1 edges even when nothing is emitted (llr_last)
    float *llr_last;             // &llr_last[b][z][0] (stride llr_pitch per lane) or nullptr
    int llr_pitch;
    int E;
    uint8_t *mask_out;           // training dump: &hist_mask[t][b][0] of the iteration being emitted, or nullptr
    uint8_t *ucn_out;            // training dump: &hist_ucn[t][b][0] ([M][Z]) or nullptr
};

// out = clamp(xo + tot, range) (:520-521) -> soft/hard outputs and, when tracked, the APP rows
__device__ __forceinline__ void emit_boosted(const EmitCtx &ec, const BoostedCtx &bc, float *__restrict__ slab, int Z, int q, float tot) {
    const float sum = addf(slab[bc.xo_row0 * Z + q], tot);
    const float v = clampf(sum, bc.lo, bc.hi);
    if (bc.mask_out) bc.mask_out[q] = (sum >= bc.lo && sum <= bc.hi) ? 1 : 0;
    emit(ec, q, v);
    if (bc.app_store_row0 >= 0) slab[bc.app_store_row0 * Z + q] = v;
}

template <int D>
__device__ __forceinline__ void cn_check_boosted(float *__restrict__ slab, int h, const GraphDev &g, int e0, int row_i,
                                                 const BoostedCtx &bc, const EmitCtx &ec, bool emit_now) {
    const bool is_qms = bc.decoder_type == 2, is_sp = bc.decoder_type == 0;
    float u[D];
    int addr[D], zz[D];
    unsigned par = 0, ucn_par = 0;
#pragma unroll
    for (int k = 0; k < D; k++) {
        int z2 = h + __ldg(g.e_shift + e0 + k);
        z2 = (z2 >= g.Z) ? z2 - g.Z : z2;                         // gather (:380-384)
        zz[k] = z2;
        addr[k] = __ldg(g.e_row + e0 + k) * g.Z + z2;
        float v = slab[addr[k]];
        v = is_qms ? quantf(v, bc.qbit) : clampf(v, bc.lo, bc.hi);             // (:386-389)
        if (!is_sp) v = addf(v, mulf(0.0001f, 1.0f - ((fabsf(v) > 0.0f) ? 1.0f : 0.0f)));   // exact zeros -> +1e-4 (:391-393)
        u[k] = v;
        par ^= (v > 0.0f) ? 1u : 0u;
        if (bc.compute_ucn) {                                     // unsatisfied-check indicator (:339-368)
            const int j = __ldg(g.e_colj + e0 + k);
            const float app = -slab[(bc.app_row0 + j) * g.Z + z2];
            ucn_par ^= (app > 0.0f) ? 0u : 1u;                   // sign -1 unless -APP > 0; product < 0 <=> odd number of -1
        }
    }
    const float s_ucn = ucn_par ? 1.0f : 0.0f;
    if (bc.ucn_out) bc.ucn_out[row_i * g.Z + h] = (uint8_t)ucn_par;
    float o[D];
    if (is_sp) {                                                  // (:400-408) product of tanh(-u/2) over the others
        float th[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            float t = tanhf(mulf(-0.5f, u[k]));
            th[k] = addf(t, 1.0f - ((fabsf(t) > 0.0f) ? 1.0f : 0.0f));
        }
#pragma unroll
        for (int k = 0; k < D; k++) {
            float p = 1.0f;
#pragma unroll
            for (int q = 0; q < D; q++)
                if (q != k) p = mulf(p, th[q]);
            p = clampf(p, -1.0f + 1e-7f, 1.0f - 1e-7f);
            o[k] = mulf(-2.0f, atanhf(p));
        }
    } else {                                                      // (:409-423)
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
            float mag = fminf(pre, suf[k + 1]);
            float a = fabsf(u[k]);
            a = (a > 0.0f) ? a : 10000.0f;
            pre = fminf(pre, a);
            mag = addf(mag, mulf(-0.0001f, addf(-((mag > 0.0001f) ? 1.0f : 0.0f), 1.0f)));      // (:416)
            const unsigned npos_odd = par ^ ((u[k] > 0.0f) ? 1u : 0u);
            o[k] = mulf(mag, npos_odd ? 1.0f : -1.0f);           // x3 * sign(-prod) (:417-423)
        }
    }
#pragma unroll
    for (int k = 0; k < D; k++) {
        const int e = e0 + k;
        const int j1 = __ldg(g.e_col1 + e);
        if (j1 >= 0 && !emit_now && !bc.want_c2v1) continue;
        const float a = fabsf(o[k]);
        float pre;                                                // (:431-503)
        if (!bc.cn_w) pre = a;
        else if (bc.ucn_mix) pre = addf(mulf(mulf(a, __ldg(bc.cn_w + e)), 1.0f - s_ucn), mulf(mulf(a, __ldg(bc.ucn_w + e)), s_ucn));
        else pre = mulf(a, __ldg(bc.cn_w + e));
        float m = mulf(pre, (pre > 0.0f) ? 1.0f : 0.0f);                                         // (:505)
        m = is_qms ? quantf(m, bc.qbit) : clampf(m, bc.lo, bc.hi);                               // (:507-510)
        const float sg = (o[k] > 0.0f) ? 1.0f : ((o[k] < 0.0f) ? -1.0f : 0.0f);
        const float c2v = mulf(m, sg);                                                           // (:512)
        if (bc.llr_last) bc.llr_last[(size_t)zz[k] * bc.llr_pitch + e] = c2v;
        if (j1 < 0) slab[addr[k]] = c2v;
        else if (emit_now) emit_boosted(ec, bc, slab, g.Z, j1 * g.Z + zz[k], addf(0.0f, c2v));
    }
}

struct BoostedLayout {
    int rows;          // slab rows per codeword
    int stride;        // floats between codeword slabs (== Z mod 32)
    int xo_row0, app_row0;
};

__global__ void __launch_bounds__(256, 1)
nldpc_generic_boosted_kernel(const GraphDev g, const DecodeArgs a, const BoostedLayout lay, const int cw_per_cta) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *slabs = reinterpret_cast<float *>(smem_raw);
    const int NZ = g.N * g.Z, Z = g.Z;
    const int hwords = (NZ + 31) >> 5;
    uint32_t *hbits_all = reinterpret_cast<uint32_t *>(slabs + (size_t)cw_per_cta * lay.stride);

    const int tid = threadIdx.x;
    const int L = cw_per_cta * Z;
    const int cw = tid / Z;
    const int z = tid - cw * Z;
    float *slab = slabs + (size_t)cw * lay.stride;
    const int n_tiles = (a.B + cw_per_cta - 1) / cw_per_cta;
    const int nb = (NZ + 7) >> 3;
    const bool is_qms = a.decoder_type == 2;
    const bool track_app = a.compute_ucn != 0;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b0 = tile * cw_per_cta;
        const int ncw = min(cw_per_cta, a.B - b0);
        const bool active = (tid < L) && (cw < ncw);
        const int b = b0 + cw;
        // ---- load: xo <- xa (quantised once if QMS), xin <- xin_init or xa, messages <- llr_init or 0, app <- app_init ----
        if (active) {
            for (int j = 0; j < g.N; j++) {
                const float x = __ldg(a.xa + (size_t)b * NZ + j * Z + z);
                slab[(lay.xo_row0 + j) * Z + z] = is_qms ? quantf(x, a.qbit) : x;                // (:517-518), idempotent
                slab[j * Z + z] = a.xin_init ? __ldg(a.xin_init + (size_t)b * NZ + j * Z + z) : x;
                if (track_app && a.app_init) slab[(lay.app_row0 + j) * Z + z] = __ldg(a.app_init + (size_t)b * NZ + j * Z + z);
            }
            for (int e = 0; e < g.E; e++) {
                if (__ldg(g.e_col1 + e) >= 0) continue;
                slab[__ldg(g.e_row + e) * Z + z] = a.llr_init ? __ldg(a.llr_init + ((size_t)b * Z + z) * g.E + e) : 0.0f;
            }
        }
        for (int i = tid; i < cw_per_cta * hwords; i += blockDim.x) hbits_all[i] = 0u;
        __syncthreads();

        for (int t = 0; t < a.T; t++) {
            // ---------------- channel-input update + VN phase (+ marginal of iteration t-1 for blocks of degree >= 2) -------
            {
                const bool soft_prev = t > 0 && a.soft_mode == 1;
                const bool hard_prev = t > 0 && a.hard_mode == 1;
                const bool app_prev = t > 0 && track_app;
                EmitCtx ec;
                ec.soft = soft_prev ? a.soft + (size_t)(t - 1) * a.B * NZ : nullptr;
                ec.hbits = hard_prev ? hbits_all + (size_t)cw * hwords : nullptr;
                ec.cw_off = (size_t)b * NZ;
                BoostedCtx bc{};
                bc.lo = a.llr_lo; bc.hi = a.llr_hi; bc.xo_row0 = lay.xo_row0;
                bc.app_store_row0 = app_prev ? lay.app_row0 : -1;
                const bool dump_prev = t > 0 && a.hist_mask != nullptr;
                bc.mask_out = dump_prev ? a.hist_mask + ((size_t)(t - 1) * a.B + b) * NZ : nullptr;
                if (active) {
                    if (a.hist_xin && t == 0)
                        for (int j = 0; j < g.N; j++) a.hist_xin[((size_t)b * g.N + j) * Z + z] = slab[j * Z + z];
                    // marginal of the previous iteration must read the OLD c2v: do it inside the block loop before the
                    // in-place v2c overwrite (vn_block returns the column total of the c2v it loaded)
                    const float *vw = a.vn_w ? a.vn_w + (size_t)t * g.N : nullptr;
                    for (int j = 0; j < g.N; j++) {                                              // (:325-337)
                        float x = slab[j * Z + z];
                        if (vw) x = mulf(x, __ldg(vw + j));
                        if (is_qms) x = quantf(x, a.qbit);
                        slab[j * Z + z] = x;
                        if (a.hist_xin) a.hist_xin[(((size_t)(t + 1) * a.B + b) * g.N + j) * Z + z] = x;
                    }
                    for (int c = 0; c < g.n_vcols; c++) {
                        const int j = __ldg(g.vcol_j + c);
                        const int p0 = __ldg(g.vcol_ptr + c);
                        const int d = __ldg(g.vcol_ptr + c + 1) - p0;
                        const float x = slab[j * Z + z];
                        float tot = 0.0f;
#define NLDPC_VN_CASE(D) tot = vn_block<D>(slab + z, Z, g.vcol_row + p0, x)
                        NLDPC_DEG_SWITCH(d, NLDPC_VN_CASE)
#undef NLDPC_VN_CASE
                        if (soft_prev || hard_prev || app_prev || dump_prev) emit_boosted(ec, bc, slab, Z, j * Z + z, tot);
                    }
                    if (a.hist_v2c) {
                        float *hv = a.hist_v2c + (((size_t)t * a.B + b) * g.S) * Z + z;
                        for (int s = 0; s < g.S; s++) hv[(size_t)s * Z] = slab[(g.N + s) * Z + z];
                    }
                }
                __syncthreads();
                if (hard_prev) {
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
                ec.cw_off = (size_t)b * NZ;
                BoostedCtx bc{};
                bc.decoder_type = a.decoder_type; bc.qbit = a.qbit; bc.lo = a.llr_lo; bc.hi = a.llr_hi;
                bc.cn_w = a.w ? a.w + (size_t)t * g.E : nullptr;
                bc.ucn_w = a.b ? a.b + (size_t)t * g.E : nullptr;
                bc.ucn_mix = a.ucn_mix != 0 && bc.cn_w && bc.ucn_w;
                bc.compute_ucn = track_app;
                // APP of the indicator: the previous output, or the (updated) channel input for the very first iteration
                bc.app_row0 = (t == 0 && !a.app_init) ? 0 : lay.app_row0;
                bc.xo_row0 = lay.xo_row0;
                bc.app_store_row0 = track_app ? lay.app_row0 : -1;
                bc.want_c2v1 = (last && a.llr_last != nullptr) || a.llr_all != nullptr;
                bc.llr_pitch = a.llr_pitch;
                bc.llr_last = a.llr_all ? a.llr_all + ((size_t)t * a.B + b) * Z * a.llr_pitch
                                        : ((last && a.llr_last) ? a.llr_last + (size_t)b * Z * a.llr_pitch : nullptr);
                bc.E = g.E;
                bc.mask_out = a.hist_mask ? a.hist_mask + ((size_t)t * a.B + b) * NZ : nullptr;
                bc.ucn_out = (a.hist_ucn && track_app) ? a.hist_ucn + ((size_t)t * a.B + b) * g.M * Z : nullptr;
                const bool emit_now = soft_now || hard_now || track_app || a.hist_mask != nullptr;
                if (active) {
                    for (int i = 0; i < g.M; i++) {
                        const int e0 = __ldg(g.row_ptr + i);
                        const int d = __ldg(g.row_ptr + i + 1) - e0;
#define NLDPC_CN_CASE(D) cn_check_boosted<D>(slab, z, g, e0, i, bc, ec, emit_now)
                        NLDPC_DEG_SWITCH(d, NLDPC_CN_CASE)
#undef NLDPC_CN_CASE
                    }
                }
                __syncthreads();
            }
        }
        // ---------------- final marginal (last iteration) of the blocks of degree >= 2, state hand-back ----------------
        {
            const bool soft_now = a.soft_mode != 0, hard_now = a.hard_mode != 0;
            EmitCtx ec;
            ec.soft = soft_now ? a.soft + (a.soft_mode == 1 ? (size_t)(a.T - 1) * a.B * NZ : 0) : nullptr;
            ec.hbits = hard_now ? hbits_all + (size_t)cw * hwords : nullptr;
            ec.cw_off = (size_t)b * NZ;
            BoostedCtx bc{};
            bc.lo = a.llr_lo; bc.hi = a.llr_hi; bc.xo_row0 = lay.xo_row0; bc.app_store_row0 = -1;
            bc.mask_out = a.hist_mask ? a.hist_mask + ((size_t)(a.T - 1) * a.B + b) * NZ : nullptr;
            if (active) {
                if (soft_now || hard_now || a.hist_mask) {
                    for (int c = 0; c < g.n_vcols; c++) {
                        const int j = __ldg(g.vcol_j + c);
                        const int p0 = __ldg(g.vcol_ptr + c), p1 = __ldg(g.vcol_ptr + c + 1);
                        float tot = 0.0f;
                        for (int k = p0; k < p1; k++) tot = addf(tot, slab[__ldg(g.vcol_row + k) * Z + z]);
                        emit_boosted(ec, bc, slab, Z, j * Z + z, tot);
                    }
                }
                if (a.xin_out)
                    for (int j = 0; j < g.N; j++) a.xin_out[(size_t)b * NZ + j * Z + z] = slab[j * Z + z];
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

// host side ---------------------------------------------------------------------------------------------------
int generic_boosted_prepare() {
    return (int)cudaFuncSetAttribute(nldpc_generic_boosted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
}

// returns 0, a cudaError_t, or -2 when one codeword does not fit
int generic_launch_boosted(const GraphDev &g, const DecodeArgs &a, int sm_count, cudaStream_t st) {
    BoostedLayout lay;
    lay.xo_row0 = g.N + g.S;
    lay.app_row0 = lay.xo_row0 + g.N;
    lay.rows = lay.app_row0 + (a.compute_ucn ? g.N : 0);
    int stride = lay.rows * g.Z;
    while ((stride & 31) != (g.Z & 31)) stride++;
    lay.stride = stride;
    const int hwords = (g.N * g.Z + 31) / 32;
    const size_t per_cw = (size_t)stride * 4 + (size_t)hwords * 4;
    if (g.Z > 256 || per_cw + 64 > (size_t)kSmemBudget) return -2;
    const size_t half_budget = (size_t)(kSmemBudget - 2048) / 2;
    int cw = (int)std::min<size_t>((half_budget - 64) / per_cw, (size_t)(256 / g.Z));
    if (cw < 1) cw = (int)std::min<size_t>(((size_t)kSmemBudget - 64) / per_cw, (size_t)(256 / g.Z));
    if (cw < 1) cw = 1;
    const int threads = ((cw * g.Z + 31) / 32) * 32;
    const size_t smem = (size_t)cw * stride * 4 + (size_t)cw * hwords * 4 + 16;
    const int n_tiles = (a.B + cw - 1) / cw;
    const int ctas_per_sm = std::max(1, (int)((size_t)kSmemBudget / (smem + 1024)));
    const int grid = std::min(n_tiles, sm_count * ctas_per_sm);
    nldpc_generic_boosted_kernel<<<grid, threads, smem, st>>>(g, a, lay, cw);
    return (int)cudaGetLastError();
}

}  // namespace nldpc
