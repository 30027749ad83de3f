// nldpc_common.cuh — shared device helpers (mbarrier / 1-D bulk TMA / exact fp32 primitives) and the
// device-side Tanner graph descriptor.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nldpc {

constexpr int kMaxDeg = 32;        // max check / variable degree supported by the generic kernel
constexpr int kSmemBudget = 227 * 1024;

// Device tables of one lifted Tanner graph (all int32 arrays live in one device allocation).
// Shared-memory state of ONE codeword ("slab"): rows of Z floats,
//   rows [0, N)        channel LLR of variable block j          (xa[b][j][:])
//   rows [N, N+S)      messages of the S edges whose variable block has degree >= 2 ("stored" edges);
//                      in place: c2v after the CN phase, v2c after the VN phase.
// Edges of degree-1 variable blocks need no storage: their v2c is the channel LLR and their c2v only
// feeds that block's own marginal, which the CN phase writes directly.
struct GraphDev {
    int M, N, Z, E, S;
    int n_vcols;           // number of variable blocks with degree >= 2
    int slab_stride;       // floats between consecutive codeword slabs (== Z mod 32 -> conflict-free lanes)
    const int *vcol_j;     // [n_vcols]   variable block index
    const int *vcol_ptr;   // [n_vcols+1] offsets into vcol_row
    const int *vcol_row;   // slab row (N + slot) of each edge of the block, ascending check row
    const int *row_ptr;    // [M+1] row-major edge ranges per check
    const int *e_row;      // [E] slab row read by the CN phase: N+slot (stored) or j (degree-1 block)
    const int *e_shift;    // [E] circulant shift (mod Z)
    const int *e_col1;     // [E] variable block j if it has degree 1, else -1
    const int *e_colj;     // [E] variable block j of the edge
};

struct DecodeArgs {
    const float *xa;   // [B][N][Z]
    const float *w;    // Neural: weights_var [T][E];  Boosted: cn_w [T][E] or nullptr
    const float *b;    // Neural: biases_var  [T][E];  Boosted: ucn_w [T][E] or nullptr
    const float *vn_w; // Boosted: [T][N] or nullptr
    int B, T;
    int soft_mode; float *soft;
    int hard_mode; uint8_t *hard;
    float *llr_last;   // Boosted: [B][Z][E] or nullptr
    float *llr_all;    // Boosted: [T][B][Z][E] (self.llr[t+1] of every executed iteration) or nullptr
    int llr_pitch;     // row pitch (floats) of llr_last / llr_all: [..][b][z][e] at ((.. * B + b) * Z + z) * llr_pitch + e  (>= E)
    int unit_begin, unit_end;   // specialised kernels: work units [unit_begin, unit_end) of this launch (unit_end 0 = all)
    int wb_off;        // specialised kernels: offset (float2 units) of this launch's weights in the constant arena, -1 = use w/b pointers
    // boosted config
    int decoder_type, qbit, compute_ucn, ucn_mix;
    float llr_lo, llr_hi;
    // boosted state for runs that do not start from a zero state (nullptr = default)
    const float *llr_init;   // [B][Z][E] c2v entering the first executed iteration
    const float *xin_init;   // [B][N][Z] compounding channel input entering the first executed iteration
    float *xin_out;          // [B][N][Z] receives it after the last one
    const float *app_init;   // [B][N*Z] previous output for the UCN indicator of the first executed iteration
    // training dump for the backward kernel (nullptr = off), see nldpc_backward.cu
    float *hist_v2c;         // [T][B][S][Z]
    float *hist_xin;         // [T+1][B][N][Z] (Boosted)
    uint8_t *hist_mask;      // [T][B][N*Z]   (Boosted)
    uint8_t *hist_ucn;       // [T][B][M][Z]  (Boosted, UCN)
    // hist_fmt 1 (specialised kernels only; nldpc_spec_backward.cuh reads it): hist_v2c holds the CN inputs of every check
    // as check-packed per-lane records [T][B][record][Z][P] (fp16 for QMS q=5, fp32 otherwise; degree-1 edges included),
    // hist_xin only rows 1..T-1 and only with VN weights, hist_mask may be nullptr
    int hist_fmt;
    int desc_base;           // specialised training-mode forward: this graph's first word in the constant descriptor table (set by the launcher)
    // fused multi-iteration BCE (training forward, Boosted): with `ybits` set the kernel writes dL/dout instead of out to `soft`
    // (clamp mask folded in) and accumulates sum_t c_t * sum_i bce(out_t[i], y[i]) into *loss_acc
    const uint8_t *ybits;    // [B][ceil(N*Z/8)] label bits, bit i of a codeword = y[i] != 0 (same packing as `hard`)
    const float *coef;       // [T] c_t
    float ginv;              // upstream dL / (B*N*Z)
    double *loss_acc;
};

// host-side launch helpers of the table-driven kernel (nldpc_generic.cu)
int generic_prepare(size_t smem_bytes);
int generic_launch_neural(const GraphDev &g, const DecodeArgs &a, int cw_per_cta, int threads, size_t smem_bytes, int use_tma,
                          int grid, cudaStream_t st);
// launch arguments of the backward kernel (nldpc_backward.cu)
struct BwdArgs {
    const float *xa;        // [B][N][Z]
    const float *w, *b;     // Neural: w/b [T][E];  Boosted: cn_w / ucn_w [T][E] or nullptr
    const float *vn_w;      // Boosted: [T][N] or nullptr
    const float *gout;      // [T][B][N*Z]
    int hist_fmt;           // 0: slot-major fp32 dump (below), 1: check-packed dump (see DecodeArgs::hist_fmt)
    const float *hist_v2c;  // [T][B][S][Z]   v2c of the stored edges entering the CN phase of iteration t
    const float *hist_xin;  // [T+1][B][N][Z] channel-input state: [0] = before iteration 0, [t+1] = after iteration t's update (Boosted)
    const uint8_t *hist_mask;   // [T][B][N*Z] 1 where the output clamp passed the gradient (Boosted); nullptr: already folded into gout
    const uint8_t *hist_ucn;    // [T][B][M][Z] unsatisfied-check indicator (Boosted, ucn_mix) or nullptr
    float *gw, *gb;         // [T][E] (+=)   Neural: weights/biases;  Boosted: cn_w / ucn_w rows
    float *gvn;             // [T][N] (+=)   Boosted VN weights or nullptr
    float *scratch;         // specialised kernel: [grid][kXRegs][threads] VN-weight chain state of the looped blocks, or nullptr
    int B, T;
    int mode;               // 0 Neural, 1 Boosted MS, 2 Boosted QMS
    int qbit;
    float lo, hi;
    int ucn_mix;
};
int backward_prepare();
int backward_launch(const GraphDev &g, const BwdArgs &a, int sm_count, cudaStream_t st);   // -2: does not fit
int generic_boosted_prepare();
int generic_launch_boosted(const GraphDev &g, const DecodeArgs &a, int sm_count, cudaStream_t st);   // -2: does not fit

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk) ------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared, `bytes` multiple of 16, both addresses 16-byte aligned; completes on `bar` (complete_tx).
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk store (bytes multiple of 16, 16-byte aligned), tracked by the bulk async-group.
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- exact fp32 primitives: one IEEE rounding each, never contracted into FMA -----------------------------
__device__ __forceinline__ float addf(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float mulf(float a, float b) { return __fmul_rn(a, b); }

__device__ __forceinline__ void st_global_stream(float *p, float v) { __stcs(p, v); }

}  // namespace nldpc
