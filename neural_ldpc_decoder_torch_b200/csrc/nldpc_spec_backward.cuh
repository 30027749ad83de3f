// nldpc_spec_backward.cuh — specialised (graph-as-immediates) version of the backward sweep of nldpc_backward.cu for the
// built-in codes: Neural (weights + biases) and Boosted MS / QMS q=5 without UCN (CN rows, optional VN rows).
//
// Same closed form as the table-driven kernel; what changes is the machinery:
//   * one thread per (codeword, lane z), gradient messages in the forward's slab layout (rotation = pointer choice),
//   * the forward's per-iteration v2c / channel-input state / clamp mask come from the HBM dump the training-mode forward
//     wrote (coalesced 64 B rows per codeword),
//   * per-edge weight gradients are accumulated hierarchically and WITHOUT atomics in the inner loop (fp32 atomicAdd on
//     shared memory is an ATOMS.CAST.SPIN loop on sm_100), and without leaving the SM: every lane stores its per-edge
//     term to its own column of a small WARP-PRIVATE shared-memory buffer [16 rows][32 lanes] (one conflict-free STS per
//     edge); when 16 rows are full the warp sums each row (lane l: row l & 15, half l >> 4: four LDS.128, 15 adds, one
//     shuffle) and adds the 16 sums into a per-CTA [T][rows] table in shared memory (one shared atomic per 16 rows per
//     warp) -> one global atomicAdd per (t, row) per CTA at the end of the kernel.  (Round 1 kept the rows in an
//     "L2-resident" global scratch [row][thread]; ncu showed 3 GB of them written back to DRAM per launch.)  With the
//     fold private to the warp nothing is shared across groups any more, so the phases synchronise per group only,
//   * the VN-weight chain (d xa_input) is lane-private and lives in registers.
#pragma once
#include <algorithm>

#include "nldpc_spec.cuh"
#include "nldpc_spec_host.cuh"

namespace nldpc {

__device__ __forceinline__ float bwd_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp-private fold of per-lane weight-gradient terms (see the header comment).  Rows are identified by their POSITION in
// the (fixed) order in which an iteration produces them; `RowOrder` below records the row each position stands for.
constexpr int kFoldRows = 16, kFoldStride = 36;      // 36 floats per row: LDS.128 of a quarter warp hits 32 distinct banks
struct Folder {
    float *col;               // &buf[warp][0][lane]
    const float *rd;          // &buf[warp][lane & 15][(lane >> 4) * 16]
    float *tot_t;             // &tot[t][0], indexed by position
    int nslot, pos0;          // rows buffered / position of the first of them (warp-uniform)
    int ln;
    __device__ __forceinline__ void begin(float *tot_row) {
        tot_t = tot_row;
        nslot = 0;
        pos0 = 0;
    }
    __device__ __forceinline__ void flush() {
        if (nslot == 0) return;
        __syncwarp();
        const float4 *p = reinterpret_cast<const float4 *>(rd);
        const float4 a = p[0], b = p[1], c4 = p[2], d = p[3];
        float sum = ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
        sum += ((c4.x + c4.y) + (c4.z + c4.w)) + ((d.x + d.y) + (d.z + d.w));
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (ln < nslot && sum != 0.0f) atomicAdd(tot_t + pos0 + ln, sum);
        __syncwarp();
        pos0 += nslot;
        nslot = 0;
    }
    __device__ __forceinline__ void reserve(int n) {
        if (nslot + n > kFoldRows) flush();
    }
    __device__ __forceinline__ void put(float v) {
        col[nslot * kFoldStride] = v;
        nslot++;
    }
};

template <class G, int MODE>
struct BwdLane {
    static constexpr int Z = G::Z;
    float *lane;              // &slab[z]
    float *rot[Z];            // &slab[(z + s) mod Z]
    const char *hv;           // this codeword's check-packed records of iteration t (DecodeArgs::hist_fmt 1)
    const float *gt;          // &gout[t][b][0]
    const uint8_t *mk;        // &hist_mask[t][b][0], or nullptr when the clamp mask is already folded into gout
    int wb_base;              // constant-arena offset of {w,b}[t][0]
    Folder fold;              // weight-gradient terms of this iteration -> per-CTA totals
    float *chn;               // &scratch[cta][0][tid]: VN-weight chain state of the looped degree-1 blocks (kVn), by register index
    const float *xprev;       // channel-input state entering iteration t: &hist_xin[t][b][0], or &xa[b][0] for t = 0 (kVn)
    const float *vw;          // w_VN[t] (kVn)
    bool last_iter;           // t == T - 1: the chain starts from zero
    float lo, hi;
    int z;
    bool valid;
    float dxr[G::kXRegs > 0 ? G::kXRegs : 1];      // d xa_input of the register-resident degree-1 blocks (this iteration)

    // upstream gradient through the output clamp mask.  Loads are unconditional (rows of codeword 0 stand in for the
    // padding lanes of the last tile) so that they stay branch-free and can be hoisted; the select zeroes padding lanes.
    __device__ __forceinline__ float g_at(int q) const {
        const float gv = __ldg(gt + q);
        bool keep = valid;
        if (mk) keep = keep && (__ldg(mk + q) != 0);
        return keep ? gv : 0.0f;
    }
};

// VN phase backwards: dc2v[e] = G[j] + (sum of the block's dv2c_{t+1} - own).  The upstream gradients of all blocks are
// fetched first (independent loads in flight together), then the lane-private row updates run.
template <class G, int MODE>
struct VnBwdLoad {
    BwdLane<G, MODE> &c;
    float *g;                 // [N] per-thread (only the blocks of degree >= 2 are touched)
    template <int J, int XROW, int... R>
    __device__ __forceinline__ void col() {
        g[J] = c.g_at(J * G::Z + c.z);
    }
};
template <class G, int MODE>
struct VnBwd {
    BwdLane<G, MODE> &c;
    const float *g;
    template <int J, int XROW, int... R>
    __device__ __forceinline__ void col() {
        const float gj = g[J];
        float tot = 0.0f;
        ((tot += c.lane[R * G::Z]), ...);
        ((c.lane[R * G::Z] = gj + (tot - c.lane[R * G::Z])), ...);
    }
};

// sum of the block's dv2c_t (after the CN phase) -> d xa_input contribution of the blocks of degree >= 2
template <class G, int MODE>
struct VnChainSum {
    BwdLane<G, MODE> &c;
    float *dx;                // [N] per-thread, indexed by block
    template <int J, int XROW, int... R>
    __device__ __forceinline__ void col() {
        float tot = 0.0f;
        ((tot += c.lane[R * G::Z]), ...);
        dx[J] = tot;
    }
};

// ---- operand staging with cp.async ------------------------------------------------------------------------------------
// ptxas, at 255 registers, sinks register-destination loads issued a check ahead back to 10-45 instructions before their
// first use (measured in SASS), which leaves every check exposed to the full L2 / HBM latency.  An asynchronous
// global -> shared copy has no register destination and cannot be sunk past its wait: each lane copies ITS OWN operands of
// check i+1 into a private staging slot [stage][entry][thread] (conflict-free, no cross-lane synchronisation needed) while
// check i computes, then waits for the group and reads them back with LDS.
__device__ __forceinline__ void cp_async4(const float *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#ifndef NLDPC_BWD_SLOTS
#define NLDPC_BWD_SLOTS 3   // staging slots of the fp16-record loops: operands are fetched NLDPC_BWD_SLOTS - 1 checks ahead
#endif
#ifndef NLDPC_BWD_L2_PREFETCH
// 1: pull the next iteration's records / gradients / channel inputs of a codeword into L2 a whole iteration ahead.  Measured in
// both rounds and left off: 21.3 vs 19.3 ms per sweep in round 2 (the scratch rows it had competed with in round 1 are gone;
// it still loses — the streamed lines are evicted again before the cp.async two checks ahead of the compute asks for them).
#define NLDPC_BWD_L2_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async8(const void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(const void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

// Staging slot of one check, per thread.  The forward wrote the check's CN inputs as ONE contiguous record per check lane
// (check-packed dump, DecodeArgs::hist_fmt 1; degree-1 edges included), so a check costs one to three copies instead of one
// per edge, and no rotated addresses:
//   fp16 records (QMS q=5): per WARP two arrays of 16-byte entries [entry][lane]: the record (4 / 8 / 16 halfs), or — checks
//     with a degree-1 edge, whose records have at most 8 halfs — the record in entry 0 and, in the space of the warp's entry-1
//     array, 4-byte arrays [comp][lane] {upstream gradient, mask word, xprev} (component-major: consecutive lanes, consecutive
//     banks).  Warp-private on purpose: warps drift apart, and a lane's extras alias OTHER lanes' entry-1 records;
//   fp32 records: kMaxRowDeg + 3 4-byte entries [entry][thread]: D values, then gradient, mask word, xprev.
// The fp16 slots are small enough for three of them: the loops fetch two checks ahead.
template <class G, int MODE>
struct BwdStage {
    static constexpr bool kHalf = MODE == 2;
    static constexpr int kEntF = kHalf ? 8 : G::kMaxRowDeg + 3;      // floats per thread and slot
    static constexpr int kSlots = kHalf ? NLDPC_BWD_SLOTS : 2;
    static constexpr size_t kCwBytes = kHalf ? (size_t)G::kDumpH * G::Z * 2 : (size_t)G::kDumpF * G::Z * 4;   // per codeword and iteration
    // `stg` = this thread's place in a slot: fp32 &slot[0][tid]; fp16 &slot[warp][entry 0][lane] (256 floats per warp).
    // address of extra `comp` (0 gradient, 1 mask word, 2 xprev) of this thread
    template <int kThreads>
    __device__ static __forceinline__ const float *extra(const float *stg, int comp) {
        if constexpr (kHalf) return stg + 128 + comp * 32 - 3 * (int)(threadIdx.x & 31);
        else return stg + (G::kMaxRowDeg + comp) * kThreads;
    }
    template <int kThreads>
    __device__ static __forceinline__ const float *thread_base(const float *stage) {
        if constexpr (kHalf) return stage + (int)(threadIdx.x >> 5) * 256 + 4 * (int)(threadIdx.x & 31);
        else return stage + threadIdx.x;
    }
};

// issue the copies of one record: D values of the check starting at element offset OFF (per lane) of this codeword-iteration
template <class G, int MODE, int kThreads, int D>
__device__ __forceinline__ void record_issue(const BwdLane<G, MODE> &c, const float *stg, int off) {
    using St = BwdStage<G, MODE>;
    if constexpr (St::kHalf) {
        constexpr int P = G::dump_slots_h(D);
        const char *src = c.hv + ((size_t)off * G::Z + (size_t)c.z * P) * 2;
        if constexpr (P == 4) cp_async8(stg, src);
        else cp_async16(stg, src);
        if constexpr (P == 16) cp_async16(stg + 128, src + 16);
    } else {
        const float *src = reinterpret_cast<const float *>(c.hv) + (size_t)off * G::Z + (size_t)c.z * D;
#pragma unroll
        for (int k = 0; k < D; k++) cp_async4(stg + k * kThreads, src + k);
    }
}
template <class G, int MODE, int kThreads, int D>
__device__ __forceinline__ void record_fetch(const float *stg, float *pv) {
    using St = BwdStage<G, MODE>;
    if constexpr (St::kHalf) {
        constexpr int P = G::dump_slots_h(D);
        uint32_t w[P / 2];
        if constexpr (P == 4) {
            const uint2 v = *reinterpret_cast<const uint2 *>(stg);
            w[0] = v.x; w[1] = v.y;
        } else {
#pragma unroll
            for (int i = 0; i < P / 8; i++) {
                const uint4 v = *reinterpret_cast<const uint4 *>(stg + i * 128);
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
        }
#pragma unroll
        for (int k = 0; k < D; k++) {
            const __half2 h = *reinterpret_cast<const __half2 *>(&w[k >> 1]);
            pv[k] = (k & 1) ? __high2float(h) : __low2float(h);
        }
    } else {
#pragma unroll
        for (int k = 0; k < D; k++) pv[k] = stg[k * kThreads];
    }
}

// unrolled checks: record + (for a degree-1 edge) the upstream gradient of its block with the clamp-mask word
template <class G, int MODE, int kThreads, class... Es>
__device__ __forceinline__ void cn_check_bwd_issue(const BwdLane<G, MODE> &c, const float *stg) {
    using St = BwdStage<G, MODE>;
    constexpr int D = sizeof...(Es);
    constexpr int shf[D] = {Es::shift...};
    constexpr int eix[D] = {Es::e...};
    constexpr int col1[D] = {Es::col1...};
    constexpr int n_deg1 = ((Es::col1 >= 0 ? 1 : 0) + ...);
    static_assert(n_deg1 <= 1 && D <= G::kMaxRowDeg, "staging holds one degree-1 edge per check");
    static_assert(!(St::kHalf && n_deg1 > 0 && G::dump_slots_h(D) > 8), "fp16 slot: the extras share the second entry");
    record_issue<G, MODE, kThreads, D>(c, stg, St::kHalf ? G::dump_off_h(eix[0]) : G::dump_off_f(eix[0]));
#pragma unroll
    for (int k = 0; k < D; k++) {
        if (col1[k] >= 0) {
            const int zz = (int)(c.rot[shf[k]] - (c.lane - c.z));
            const int q = col1[k] * G::Z + zz;
            cp_async4(St::template extra<kThreads>(stg, 0), c.gt + q);
            if (c.mk) cp_async4(St::template extra<kThreads>(stg, 1), c.mk + (q & ~3));
        }
    }
    cp_async_commit();
}

template <class G, int MODE, int kThreads, class... Es>
__device__ __forceinline__ void cn_check_bwd_fetch(const BwdLane<G, MODE> &c, const float *stg, float *pv, float *pg) {
    using St = BwdStage<G, MODE>;
    constexpr int D = sizeof...(Es);
    constexpr int shf[D] = {Es::shift...};
    constexpr int col1[D] = {Es::col1...};
    record_fetch<G, MODE, kThreads, D>(stg, pv);   // (padding lanes carry codeword 0's values: harmless, every gradient they meet is zero)
#pragma unroll
    for (int k = 0; k < D; k++) {
        if (col1[k] >= 0) {
            const float gv = *St::template extra<kThreads>(stg, 0);
            bool keep = c.valid;
            if (c.mk) {
                const int zz = (int)(c.rot[shf[k]] - (c.lane - c.z));
                const uint32_t w = __float_as_uint(*St::template extra<kThreads>(stg, 1));
                keep = keep && (((w >> (8 * ((col1[k] * G::Z + zz) & 3))) & 0xffu) != 0);
            }
            pg[k] = keep ? gv : 0.0f;
        }
    }
}

// One check, backwards — the arithmetic, independent of where operands come from.  Forward recap (per edge k):
// mag_k = min(min_{j != k} |u_j|, 10000) (zeros masked to 10000), sign_k negative iff an even number of the OTHER inputs is
// positive, out_k = sign_k * relu(mag_k * w + b) (Neural) or the conditioned Boosted form.  mag_k takes only two values per
// check (m1 for k != i1, m2 for k == i1), so everything that depends on it alone is computed once per variant; signs are
// carried as XOR masks on the raw words.
//   in : pv[k] forward CN input (raw), dc[k] upstream gradient of the edge's c2v, wb[k] = {w, b}
//   out: gw[k] (and gb[k], Neural) per-lane weight-gradient terms, du[k] gradient w.r.t. the raw CN input
template <int MODE, int D>
__device__ __forceinline__ void cn_bwd_math(const float *pv, const float *dc, const float2 *wb, float lo, float hi, float *gw, float *gb,
                                            float *du) {
    constexpr uint32_t kSign = 0x80000000u;
    float u[D], av[D];
    uint32_t sb[D];           // sign word of u_k ("not positive" <=> bit 31 set; a Neural zero counts as not positive)
    bool pass[D];             // the input conditioning (clamp / quantiser window) lets the gradient through
    uint32_t xall = ((D - 1) & 1) ? 0u : kSign;
    float m1 = 3.0e38f, m2 = 3.0e38f;
#pragma unroll
    for (int k = 0; k < D; k++) {
        float v = pv[k];
        if constexpr (MODE == 2) {
            // QMS q=5 as in the forward fast path (cn_check_boosted_core): the recorded CN input is already on the 0.5 grid
            // (zeros are +0.0, standing for the reference's "+1e-4, then mag - 1e-4 = 0"; no gradient flows through them:
            // relu'(0) = 0), so the re-quantisation is the clamp — taken on the magnitude below, the sign is the raw word's.
            pass[k] = fabsf(v) <= 7.5f;
        } else if constexpr (MODE == 1) {
            pass[k] = (v >= lo && v <= hi);
            v = clamp_rng(v, lo, hi);
            v = (v == 0.0f) ? 0.0001f : v;
        } else {
            pass[k] = true;
        }
        u[k] = v;
        float a = fabsf(v);
        if constexpr (MODE == 2) a = fminf(a, 7.5f);
        if constexpr (MODE == 0) {
            sb[k] = (v == 0.0f) ? kSign : __float_as_uint(v);
            a = (a > 0.0f) ? a : 10000.0f;
        } else {
            sb[k] = __float_as_uint(v);
        }
        av[k] = a;
        xall ^= sb[k];
        m2 = fminf(m2, fmaxf(m1, a));       // two smallest magnitudes (m2 == m1 when the minimum occurs twice)
        m1 = fminf(m1, a);
    }
    // FIRST-index tie rule of torch.min on CPU: i1 = first k attaining m1, i2 = first k != i1 attaining m2
    int i1 = 0, i2 = 0;
#pragma unroll
    for (int k = D - 1; k >= 0; k--) i1 = (av[k] == m1) ? k : i1;
#pragma unroll
    for (int k = D - 1; k >= 0; k--) i2 = (av[k] == m2 && k != i1) ? k : i2;
    // the two variants of everything that depends on the others' minimum alone
    const bool capA = !(m1 < 10000.0f), capB = !(m2 < 10000.0f);
    float madjA = fminf(m1, 10000.0f), madjB = fminf(m2, 10000.0f);
    if constexpr (MODE == 1) {
        madjA = (madjA > 0.0001f) ? madjA : addf(madjA, -0.0001f);
        madjB = (madjB > 0.0001f) ? madjB : addf(madjB, -0.0001f);
    }
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < D; k++) {
        const bool first = (k == i1);
        const float msel = first ? madjB : madjA;
        const uint32_t smask = (xall ^ sb[k]) & kSign;                  // bit 31 set <=> sign_k negative
        float dm;
        if constexpr (MODE == 0) {
            const float pre = addf(mulf(msel, wb[k].x), wb[k].y);
            const float base = (pre > 0.0f) ? dc[k] : 0.0f;
            const float coef = __uint_as_float(__float_as_uint(base) ^ smask);      // dc * sign_k
            gw[k] = coef * msel;
            gb[k] = coef;
            dm = coef * wb[k].x;
        } else {
            const float o = __uint_as_float(__float_as_uint(msel) ^ smask);         // madj * sign_k
            const float pre = mulf(fabsf(o), wb[k].x);
            bool live = pre > 0.0f;
            if constexpr (MODE == 2) live = live && (pre <= 7.5f);
            else live = live && (pre >= lo && pre <= hi);
            const float base = live ? dc[k] : 0.0f;
            gw[k] = base * o;                                                       // dc * sign(o) * |o|
            dm = __uint_as_float(__float_as_uint(base * wb[k].x) ^ smask);          // d out / d madj = w * sign_k
        }
        if (first) s2 += dm;
        else s1 += dm;
    }
    s1 = capA ? 0.0f : s1;      // a capped / all-masked minimum is a constant
    s2 = capB ? 0.0f : s2;
#pragma unroll
    for (int k = 0; k < D; k++) {
        float d = (k == i1) ? s1 : ((k == i2) ? s2 : 0.0f);
        d = __uint_as_float(__float_as_uint(d) ^ (__float_as_uint(u[k]) & kSign));      // * sign(u_k)
        if constexpr (MODE == 0) d = (u[k] == 0.0f) ? 0.0f : d;
        du[k] = pass[k] ? d : 0.0f;
    }
}

// unrolled form: every table entry of the check is an immediate
template <class G, int MODE, bool kVn, int kThreads, class... Es>
__device__ __forceinline__ void cn_check_bwd_core(BwdLane<G, MODE> &c, const float *pv, const float *pg) {
    constexpr int D = sizeof...(Es);
    constexpr int rows[D] = {Es::row...};
    constexpr int shf[D] = {Es::shift...};
    constexpr int eix[D] = {Es::e...};
    constexpr int col1[D] = {Es::col1...};
    constexpr int Z = G::Z;
    float dc[D], gw[D], gb[D], du[D];
    float2 wb[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
        if (col1[k] < 0) dc[k] = c.rot[shf[k]][rows[k] * Z];
        else dc[k] = pg[k];
        wb[k] = wb_at<MODE != 0>(c.wb_base + eix[k]);
    }
    cn_bwd_math<MODE, D>(pv, dc, wb, c.lo, c.hi, gw, gb, du);
    // fold order (RowOrder::chk mirrors it): the check's weight rows, then (Neural) its bias rows
    c.fold.reserve(D);
#pragma unroll
    for (int k = 0; k < D; k++) c.fold.put(gw[k]);
    if constexpr (MODE == 0) {
        c.fold.reserve(D);
#pragma unroll
        for (int k = 0; k < D; k++) c.fold.put(gb[k]);
    }
#pragma unroll
    for (int k = 0; k < D; k++) {
        if (col1[k] < 0) c.rot[shf[k]][rows[k] * Z] = du[k];          // dv2c_t, variable-lane domain
        else if constexpr (kVn) c.dxr[rows[k] < 0 ? -rows[k] - 1 : 0] = du[k];   // (register-resident: G::kDeg1Smem == 0)
    }
}

// ---- looped form: checks with D stored edges + ONE trailing degree-1 register block, table entries at run time ----------
// The unrolled iteration body of BG2 is 12 K instructions = 190 KB and the sweep is instruction-delivery bound;
// truncation experiments put the cliff between 80 KB and 160 KB.  The 38 "extension" checks of BG2 differ only in their
// table entries, so they run as four short loops (one per stored-edge count) over one-word descriptors in constant memory:
// ~4 more address instructions per edge, but a body of ~5 K instructions.
template <class G, int MODE, bool kVn, int kThreads, int D>
__device__ __forceinline__ void cn_loop_issue(const BwdLane<G, MODE> &c, const float *stg, int w0, int rec_off) {
    using St = BwdStage<G, MODE>;
    constexpr int Z = G::Z;
    static_assert(D + 1 <= G::kMaxRowDeg && (!St::kHalf || G::dump_slots_h(D + 1) <= 8), "staging entries");
    record_issue<G, MODE, kThreads, D + 1>(c, stg, rec_off);       // D stored edges + the degree-1 edge
    const int q = (int)(c_desc[w0 + D] & 0xff) * Z + c.z;          // degree-1 block J, identity circulant
    cp_async4(St::template extra<kThreads>(stg, 0), c.gt + q);
    if (c.mk) cp_async4(St::template extra<kThreads>(stg, 1), c.mk + (q & ~3));
    if constexpr (kVn) cp_async4(St::template extra<kThreads>(stg, 2), c.xprev + q);
    cp_async_commit();
}

template <class G, int MODE, bool kVn, int kThreads, int D>
__device__ __forceinline__ void cn_loop_compute(BwdLane<G, MODE> &c, const float *stg, int w0, float chain_prev) {
    using St = BwdStage<G, MODE>;
    constexpr int Z = G::Z, NE = D + 1;
    float *slab0 = c.lane - c.z;
    float pv[NE], dc[NE], gw[NE], gb[NE], du[NE];
    float2 wb[NE];
    float *msg[D];
    record_fetch<G, MODE, kThreads, NE>(stg, pv);
#pragma unroll
    for (int k = 0; k < D; k++) {
        const uint32_t w = c_desc[w0 + k];
        msg[k] = slab0 + (w & 0xff) * Z + rot_lane<G>(c.z, (w >> 8) & 0xff);
        dc[k] = *msg[k];
        wb[k] = wb_at<MODE != 0>(c.wb_base + (int)(w >> 16));
    }
    const uint32_t w1 = c_desc[w0 + D];
    const int J = w1 & 0xff, ridx = (w1 >> 8) & 0xff;
    const int q = J * Z + c.z;
    {
        const float gv = *St::template extra<kThreads>(stg, 0);
        bool keep = c.valid;
        if (c.mk) keep = keep && (((__float_as_uint(*St::template extra<kThreads>(stg, 1)) >> (8 * (q & 3))) & 0xffu) != 0);
        dc[D] = keep ? gv : 0.0f;
    }
    wb[D] = wb_at<MODE != 0>(c.wb_base + (int)(w1 >> 16));
    cn_bwd_math<MODE, NE>(pv, dc, wb, c.lo, c.hi, gw, gb, du);
    // fold order (RowOrder::cls mirrors it): weight rows, (Neural) bias rows, (kVn) the VN row of block J
    c.fold.reserve(NE);
#pragma unroll
    for (int k = 0; k < NE; k++) c.fold.put(gw[k]);
    if constexpr (MODE == 0) {
        c.fold.reserve(NE);
#pragma unroll
        for (int k = 0; k < NE; k++) c.fold.put(gb[k]);
    }
#pragma unroll
    for (int k = 0; k < D; k++) *msg[k] = du[k];
    if constexpr (kVn) {     // VN-weight chain step of block J (VnChainStep), inline: the block belongs to this check alone
        float dx = chain_prev + du[D];
        const float xp = *St::template extra<kThreads>(stg, 2);
        const float w = __ldg(c.vw + J);
        if constexpr (MODE == 2) {
            if (!(fabsf(mulf(xp, w)) <= 7.5f)) dx = 0.0f;
        }
        c.fold.reserve(1);
        c.fold.put(dx * xp);
        __stcg(c.chn + ridx * kThreads, dx * w);
    }
}

template <class G, int MODE, bool kVn, int kThreads>
struct CnBwdLoops {
    using St = BwdStage<G, MODE>;
    BwdLane<G, MODE> &c;
    const float *stg;         // &stage[0][0][tid]
    int base;                 // first descriptor word of this graph in c_desc
    // VN-weight chain state of the check's degree-1 block (own slot, written by this thread one iteration ago; L2, not the
    // non-coherent L1).  Loaded ahead into loop-carried registers: ptxas cannot sink it across the back edge.
    template <int D>
    __device__ __forceinline__ float chain_state(int w0) const {
        if constexpr (!kVn) return 0.0f;
        if (c.last_iter) return 0.0f;
        return __ldcg(c.chn + ((c_desc[w0 + D] >> 8) & 0xff) * kThreads);
    }
    // one class: COUNT checks with D stored edges each, records REC apart starting at OFFH / OFFF; operands are fetched
    // kSlots - 1 checks ahead
    template <int D, int FIRST, int COUNT, int OFFH, int OFFF>
    __device__ __forceinline__ void cls() {
        constexpr int kStage = St::kEntF * kThreads, kAhead = St::kSlots - 1;
        constexpr int REC = St::kHalf ? G::dump_slots_h(D + 1) : D + 1, OFF = St::kHalf ? OFFH : OFFF;
#pragma unroll
        for (int j = 0; j < kAhead; j++)
            if (j < COUNT) cn_loop_issue<G, MODE, kVn, kThreads, D>(c, stg + j * kStage, base + FIRST + j * (D + 1), OFF + j * REC);
        float chain_cur = chain_state<D>(base + FIRST);
        float chain_next = COUNT > 1 ? chain_state<D>(base + FIRST + (D + 1)) : 0.0f;
        int slot_c = 0, slot_i = kAhead % St::kSlots;        // slot being computed / being filled
#pragma unroll 1
        for (int i = 0; i < COUNT; i++) {
            const int w0 = base + FIRST + i * (D + 1);
            float chain_next2 = 0.0f;
            if (i + 2 < COUNT) chain_next2 = chain_state<D>(w0 + 2 * (D + 1));
            if (i + kAhead < COUNT) {
                cn_loop_issue<G, MODE, kVn, kThreads, D>(c, stg + slot_i * kStage, w0 + kAhead * (D + 1), OFF + (i + kAhead) * REC);
                cp_async_wait<kAhead>();
            } else if (kAhead > 1 && i + 1 < COUNT) {
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            cn_loop_compute<G, MODE, kVn, kThreads, D>(c, stg + slot_c * kStage, w0, chain_cur);
            chain_cur = chain_next;
            chain_next = chain_next2;
            slot_c = slot_c + 1 == St::kSlots ? 0 : slot_c + 1;
            slot_i = slot_i + 1 == St::kSlots ? 0 : slot_i + 1;
        }
    }
};

template <class G, int MODE, bool kVn, int kThreads>
struct CnBwd {
    using St = BwdStage<G, MODE>;
    BwdLane<G, MODE> &c;
    const float *stg;         // &stage[0][0][tid]
    int n_ld = 0, n_chk = 0;  // (compile-time after inlining: the sequence is straight-line)
    template <int SLOT, class... Es>
    __device__ __forceinline__ void ld() {
        cn_check_bwd_issue<G, MODE, kThreads, Es...>(c, stg + SLOT * St::kEntF * kThreads);
        n_ld++;
    }
    template <int SLOT, class... Es>
    __device__ __forceinline__ void chk() {
        if (n_ld - n_chk >= 2) cp_async_wait<1>();       // the next check's group may stay in flight
        else cp_async_wait<0>();
        n_chk++;
        float pv[G::kMaxRowDeg], pg[G::kMaxRowDeg];
        cn_check_bwd_fetch<G, MODE, kThreads, Es...>(c, stg + SLOT * St::kEntF * kThreads, pv, pg);
        cn_check_bwd_core<G, MODE, kVn, kThreads, Es...>(c, pv, pg);
    }
};

// VN-weight chain per block: dx = chain_{t+1} + sum_e dv2c_t[e];  dz = dx * [|xin_{t-1} * w| <= 7.5] (QMS);
// grad_wVN[t][j] += dz * xin_{t-1};  chain_t = dz * w   (BoostedNeuralLDPCDecoder.py:325-337)
// Runs in batches of kChainBatch blocks: a batch's channel-input loads are issued together, then consumed.
constexpr int kChainBatch = 13;
template <class G, int MODE, int BATCH>
struct VnChainLoad {
    BwdLane<G, MODE> &c;
    const float *xprev;       // &hist_xin[t][b][0]
    float *xp;                // [kChainBatch]
    template <int J, int DEST>
    __device__ __forceinline__ void put() {
        if constexpr (J / kChainBatch == BATCH) xp[J % kChainBatch] = __ldg(xprev + J * G::Z + c.z);
    }
};
template <class G, int MODE, int kThreads, int BATCH>
struct VnChainStep {
    BwdLane<G, MODE> &c;
    float *chain;             // [N] per-thread
    const float *dx_blocks;   // [N] per-thread: sums for blocks of degree >= 2
    const float *xpb;         // [kChainBatch]
    const float *vw;          // w_VN[t]
    template <int J, int DEST>
    __device__ __forceinline__ void put() {
        if constexpr (J / kChainBatch == BATCH) {
            float dx = chain[J] + (DEST < 0 ? c.dxr[DEST < 0 ? -DEST - 1 : 0] : dx_blocks[J]);
            const float xp = xpb[J % kChainBatch];
            const float w = __ldg(vw + J);
            if constexpr (MODE == 2) {
                if (!(fabsf(mulf(xp, w)) <= 7.5f)) dx = 0.0f;
            }
            c.fold.reserve(1);                                   // (RowOrder::put mirrors it)
            c.fold.put(dx * xp);                                 // Boosted rows: [E] CN weights, then [N] VN weights
            chain[J] = dx * w;
        }
    }
};
template <class G, int MODE, int kThreads, int BATCH>
__device__ __forceinline__ void vn_chain_batches(BwdLane<G, MODE> &c, float *chain, const float *dxb, const float *xprev, const float *vw) {
    if constexpr (BATCH * kChainBatch < G::N) {
        float xp[kChainBatch];
        VnChainLoad<G, MODE, BATCH> l{c, xprev, xp};
        G::blocks_rest(l);      // (the degree-1 blocks of looped checks are stepped inside cn_loop_compute)
        VnChainStep<G, MODE, kThreads, BATCH> st{c, chain, dxb, xp, vw};
        G::blocks_rest(st);
        vn_chain_batches<G, MODE, kThreads, BATCH + 1>(c, chain, dxb, xprev, vw);
    }
}

// position -> row of the fold order, built once per CTA by one thread running the same traversal as the sweep
template <class G, int MODE, bool kVn>
struct RowOrder {
    uint16_t *order;
    int base;                 // first descriptor word of this graph in c_desc
    int n = 0;
    template <int SLOT, class... Es>
    __device__ __forceinline__ void ld() {}
    template <int SLOT, class... Es>
    __device__ __forceinline__ void chk() {
        ((order[n++] = (uint16_t)Es::e), ...);
        if constexpr (MODE == 0) ((order[n++] = (uint16_t)(G::E + Es::e)), ...);
    }
    template <int D, int FIRST, int COUNT, int OFFH, int OFFF>
    __device__ __forceinline__ void cls() {
        for (int i = 0; i < COUNT; i++) {
            const int w0 = base + FIRST + i * (D + 1);
            for (int k = 0; k <= D; k++) order[n++] = (uint16_t)(c_desc[w0 + k] >> 16);
            if constexpr (MODE == 0)
                for (int k = 0; k <= D; k++) order[n++] = (uint16_t)(G::E + (c_desc[w0 + k] >> 16));
            if constexpr (kVn) order[n++] = (uint16_t)(G::E + (c_desc[w0 + D] & 0xff));
        }
    }
    template <int J, int DEST>
    __device__ __forceinline__ void put() {
        if constexpr (kVn) order[n++] = (uint16_t)(G::E + J);
    }
};

template <class G>
struct SpecBwdCfg {
    using Fwd = SpecCfg<G, false>;
    using Shape = typename Fwd::Shape;
#ifdef NLDPC_FORCE_GROUPS
    static constexpr int kGroups = Shape::kLanes == 32 ? 8 : Fwd::kGroups;
#else
    static constexpr int kGroups = Fwd::kGroups;            // same CTA shape as the forward
#endif
    static constexpr int kThreads = kGroups * Shape::kLanes;
    static constexpr int kWarps = kThreads / 32;
    static constexpr int kCwPerCta = kGroups * Shape::kCw;
    static_assert(kThreads % 32 == 0 && kThreads <= kSpecBwdScratchLanes, "workspace scratch rows hold one float per thread");
    // the sweep keeps only the MESSAGE rows of a codeword on chip (the channel rows of the forward's slab are not needed:
    // channel inputs come from the dump), stride == Z (mod 32) as in the forward
    static constexpr int kSlabF = slab_floats(G::S * G::Z, 0, G::Z);
    static constexpr size_t slab_bytes() { return (size_t)kCwPerCta * kSlabF * 4; }
    __host__ __device__ static constexpr int rows(int mode, bool vn) { return mode == 0 ? 2 * G::E : G::E + (vn ? G::N : 0); }
    __host__ __device__ static constexpr size_t stage_floats(int mode) {
        return mode == 2 ? (size_t)BwdStage<G, 2>::kSlots * BwdStage<G, 2>::kEntF * kThreads : (size_t)BwdStage<G, 0>::kSlots * BwdStage<G, 0>::kEntF * kThreads;
    }
    static constexpr size_t fold_bytes() { return (size_t)kWarps * kFoldRows * kFoldStride * 4; }
    static constexpr size_t order_bytes(int mode, bool vn) { return ((size_t)rows(mode, vn) * 2 + 15) & ~(size_t)15; }
    // message slabs + operand staging (2 stages) + fold buffers + per-CTA totals [T][rows] + position -> row table
    static constexpr size_t smem_bytes(int T, int mode, bool vn) {
        return slab_bytes() + stage_floats(mode) * 4 + fold_bytes() + (size_t)T * rows(mode, vn) * 4 + order_bytes(mode, vn) + 64;
    }
};

template <class G, int MODE, bool kVn>
__global__ void __launch_bounds__(SpecBwdCfg<G>::kThreads, 1) nldpc_spec_backward_kernel(const BwdArgs a, const int wb_off, const int desc_base) {
    using Cfg = SpecBwdCfg<G>;
    using Shape = typename Cfg::Shape;
    constexpr int Z = G::Z, NZ = G::N * G::Z, E = G::E, N = G::N;
    constexpr int kThreads = Cfg::kThreads, kRows = Cfg::rows(MODE, kVn);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *slabs = reinterpret_cast<float *>(smem_raw);
    float *stage = slabs + (size_t)Cfg::kCwPerCta * Cfg::kSlabF;   // [kSlots][kEntF][kThreads] (BwdStage)
    float *foldb = stage + Cfg::stage_floats(MODE);                // [kWarps][kFoldRows][kFoldStride]
    float *tot = foldb + Cfg::kWarps * kFoldRows * kFoldStride;     // [T][kRows], by fold position
    uint16_t *order = reinterpret_cast<uint16_t *>(tot + (size_t)a.T * kRows);      // [kRows] position -> row

    const int tid = threadIdx.x, warp = tid >> 5, ln = tid & 31;
    const int grp = tid / Shape::kLanes, gl = tid - grp * Shape::kLanes;
    int cwl, z;
    Shape::map(gl, cwl, z);      // (Z = 24: the bank-conflict-free lane mapping of the forward, GroupShape::map)
    const int cw_in_cta = grp * Shape::kCw + cwl;
    // row indices of the graph program count the forward's channel rows too: point kXRows rows before this codeword's messages
    float *slab = slabs + (size_t)cw_in_cta * Cfg::kSlabF - G::kXRows * Z;
    // per CTA: the chain state of the looped degree-1 blocks (kVn), one row per register index
    float *scr_cta = a.scratch + (size_t)blockIdx.x * (G::kXRegs > 0 ? G::kXRegs : 1) * kThreads;

    BwdLane<G, MODE> c;
    c.lane = slab + z;
    c.z = z;
    c.lo = a.lo;
    c.hi = a.hi;
    c.chn = scr_cta + tid;
    c.fold.col = foldb + (size_t)warp * kFoldRows * kFoldStride + ln;
    c.fold.rd = foldb + (size_t)warp * kFoldRows * kFoldStride + (ln & 15) * kFoldStride + (ln >> 4) * 16;
    c.fold.ln = ln;
#pragma unroll
    for (int s = 0; s < Z; s++) c.rot[s] = slab + ((z + s) % Z);

    for (int i = tid; i < a.T * kRows; i += kThreads) tot[i] = 0.0f;
    if (tid == 0) {
        RowOrder<G, MODE, kVn> ro{order, desc_base};
        G::checks_pipelined_rest(ro);
        if constexpr (G::kLoopChecks > 0) G::loop_classes(ro);
        if constexpr (kVn) G::blocks_rest(ro);
    }
    __syncthreads();

    // phases synchronise per group (one codeword group = the lanes that share slabs); nothing else is shared between warps
    // inside the sweep (fold buffers, staging slots and chain-state rows are private, the totals take atomics)
    auto phase_sync = [&]() { group_sync<Shape::kLanes>(grp); };

    const int n_tiles = (a.B + Cfg::kCwPerCta - 1) / Cfg::kCwPerCta;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = tile * Cfg::kCwPerCta + cw_in_cta;
        c.valid = b < a.B;
        const size_t bb = (size_t)(c.valid ? b : 0);
        for (int q = G::kXRows; q < G::kXRows + G::S; q++) c.lane[q * Z] = 0.0f;      // dv2c_T = 0 (own lane)
        float chain[kVn ? N : 1], dxb[kVn ? N : 1];
        if constexpr (kVn) {
#pragma unroll
            for (int j = 0; j < N; j++) chain[j] = 0.0f;
        }
        phase_sync();
        for (int t = a.T - 1; t >= 0; t--) {
            c.hv = reinterpret_cast<const char *>(a.hist_v2c) + ((size_t)t * a.B + bb) * BwdStage<G, MODE>::kCwBytes;
            c.gt = a.gout + ((size_t)t * a.B + bb) * NZ;
            c.mk = (MODE != 0 && a.hist_mask) ? a.hist_mask + ((size_t)t * a.B + bb) * NZ : nullptr;
            c.wb_base = wb_off + t * E;
            c.last_iter = (t == a.T - 1);
            c.fold.begin(tot + (size_t)t * kRows);
#if NLDPC_BWD_L2_PREFETCH
            // The records, upstream gradients and channel inputs of the NEXT iteration of the sweep (t - 1) stream from HBM
            // exactly once; pull this codeword's lines into L2 a whole iteration (~15 K instructions) ahead, so that the
            // cp.async two checks ahead of the compute finds them there.  (Experiment knob, off: see NLDPC_BWD_L2_PREFETCH.)
            if (t > 0 && c.valid) {
                const char *p0 = reinterpret_cast<const char *>(a.hist_v2c) + ((size_t)(t - 1) * a.B + bb) * BwdStage<G, MODE>::kCwBytes;
                for (int off = z * 128; off < (int)BwdStage<G, MODE>::kCwBytes; off += Z * 128) prefetch_l2(p0 + off);
                const char *p1 = reinterpret_cast<const char *>(a.gout + ((size_t)(t - 1) * a.B + bb) * NZ);
                for (int off = z * 128; off < NZ * 4; off += Z * 128) prefetch_l2(p1 + off);
                if constexpr (kVn) {
                    const char *p2 = reinterpret_cast<const char *>(t - 1 == 0 ? a.xa + bb * NZ : a.hist_xin + ((size_t)(t - 1) * a.B + bb) * NZ);
                    for (int off = z * 128; off < NZ * 4; off += Z * 128) prefetch_l2(p2 + off);
                }
            }
#endif
            if constexpr (kVn) {
                c.xprev = (t == 0) ? a.xa + bb * NZ : a.hist_xin + ((size_t)t * a.B + bb) * NZ;      // (row 0 is the raw input)
                c.vw = a.vn_w + (size_t)t * N;
            }
            {
                float g[N];
                VnBwdLoad<G, MODE> l{c, g};
                G::vcols(l);
                VnBwd<G, MODE> f{c, g};
                G::vcols(f);
            }
            phase_sync();
            {
                const float *stg = BwdStage<G, MODE>::template thread_base<kThreads>(stage);
                CnBwd<G, MODE, kVn, kThreads> f{c, stg};
                G::checks_pipelined_rest(f);
                if constexpr (G::kLoopChecks > 0) {
                    c.fold.flush();      // the loops carry the fold counters in registers: enter them from a known state
                    CnBwdLoops<G, MODE, kVn, kThreads> l{c, stg, desc_base};
                    G::loop_classes(l);
                }
            }
            phase_sync();
            if constexpr (kVn) {
                VnChainSum<G, MODE> s{c, dxb};
                G::vcols(s);
                vn_chain_batches<G, MODE, kThreads, 0>(c, chain, dxb, c.xprev, c.vw);
            }
            c.fold.flush();
        }
        phase_sync();
    }
    __syncthreads();
    // one global atomic per (t, row) per CTA
    for (int i = tid; i < a.T * kRows; i += kThreads) {
        const int t = i / kRows, r = order[i - t * kRows];
        const float v = tot[i];
        if (v == 0.0f) continue;
        if (r < E) atomicAdd(a.gw + (size_t)t * E + r, v);
        else if (MODE == 0) atomicAdd(a.gb + (size_t)t * E + (r - E), v);
        else atomicAdd(a.gvn + (size_t)t * N + (r - E), v);
    }
}

namespace {

template <class G, int MODE, bool kVn>
int spec_bwd_launch_one(const BwdArgs &a, int wb_off, int graph_slot, int sm_count, cudaStream_t st, bool capturing) {
    using Cfg = SpecBwdCfg<G>;
    const size_t smem = Cfg::smem_bytes(a.T, MODE, kVn);
    static bool prepared[64] = {};   // per translation unit, variant and device; the attribute is idempotent
    int dev = 0;
    cudaGetDevice(&dev);
    if (!prepared[dev & 63]) {
        if (capturing) return -1;            // one-off set-up belongs to an eager (warm-up) launch
        cudaError_t e = set_smem(nldpc_spec_backward_kernel<G, MODE, kVn>, kSmemBudget);
        if (e != cudaSuccess) return (int)e;
        prepared[dev & 63] = true;
    }
    cudaError_t e = ensure_loop_desc<G>(graph_slot, capturing);
    if (e == cudaErrorStreamCaptureUnsupported && capturing) return -1;
    if (e != cudaSuccess) return (int)e;
    const int n_tiles = (a.B + Cfg::kCwPerCta - 1) / Cfg::kCwPerCta;
    const int grid = std::min(n_tiles, sm_count);
    nldpc_spec_backward_kernel<G, MODE, kVn><<<grid, Cfg::kThreads, smem, st>>>(a, wb_off, graph_slot * kDescStride);
    return (int)cudaGetLastError();
}

// the configurations the specialised sweep takes (nldpc_spec.cuh: spec_backward_covers)
template <class G>
bool spec_bwd_covers(int mode, int T, bool has_cn_w, bool has_vn_w, bool ucn, int qbit) {
    if (G::kDeg1Smem != 0) return false;                      // the lane-private chain assumes identity circulants on degree-1 blocks
    if (mode != 0 && (ucn || !has_cn_w)) return false;        // no CN weights: nothing but VN rows to learn; keep it simple
    if (mode == 2 && qbit != 5) return false;
    return SpecBwdCfg<G>::smem_bytes(T, mode, mode != 0 && has_vn_w) <= (size_t)kSmemBudget;
}

// 0 launched, >0 cudaError_t, -1 not covered (caller uses the table-driven kernel).
// kBoosted selects which kernels this translation unit instantiates: every kernel must live in exactly ONE unit, because
// it reads that unit's constant arena (a second instantiation elsewhere would be folded with this one by the linker and
// read the other unit's, never written, arena).
template <class G, bool kBoosted>
int spec_bwd_launch(const BwdArgs &a, int graph_slot, int sm_count, cudaStream_t st) {
    if (!a.scratch || kBoosted != (a.mode != 0) || a.hist_fmt != 1) return -1;      // (the sweep reads the check-packed dump)
    if (!spec_bwd_covers<G>(a.mode, a.T, a.w != nullptr, a.gvn != nullptr, a.ucn_mix || a.hist_ucn, a.qbit)) return -1;
    const bool capturing = stream_is_capturing(st);           // CUDA graph capture: see ConstArena::acquire_captured
    ConstArena &arena = arena_for_current_device();
    const int n_w = a.T * G::E;
    const int len = kBoosted ? (n_w + 1) / 2 : n_w;      // float2 units: Boosted launches store plain floats (wb_at)
    cudaError_t err = cudaSuccess;
    const int off = capturing ? arena.acquire_captured(len, st, &err) : arena.acquire(len, st, &err);
    if (err != cudaSuccess) return (int)err;
    if (off < 0) return -1;
    if ((err = kBoosted ? upload_w(arena, a.w, off, n_w, st) : upload_wb(arena, a.w, a.b, off, len, st)) != cudaSuccess) return (int)err;
    int rc;
    if constexpr (!kBoosted) {
        rc = spec_bwd_launch_one<G, 0, false>(a, off, graph_slot, sm_count, st, capturing);
    } else {
        const int offf = 2 * off;                        // float units
        if (a.mode == 1) rc = a.gvn ? spec_bwd_launch_one<G, 1, true>(a, offf, graph_slot, sm_count, st, capturing) : spec_bwd_launch_one<G, 1, false>(a, offf, graph_slot, sm_count, st, capturing);
        else rc = a.gvn ? spec_bwd_launch_one<G, 2, true>(a, offf, graph_slot, sm_count, st, capturing) : spec_bwd_launch_one<G, 2, false>(a, offf, graph_slot, sm_count, st, capturing);
    }
    if (capturing) return rc;
    const cudaError_t rel = arena.release_after(off, len, st);
    if (rc != 0) return rc;
    return (int)rel;
}

}  // namespace
}  // namespace nldpc
