// nldpc_spec_kernel.cuh — hand-written templates that turn a generated compile-time graph program
// (csrc/generated/nldpc_graph_*.cuh) into a fully unrolled decode kernel for sm_100a.
//
// Execution model ("codeword group"): Z lanes per codeword, one thread per (codeword, lane).  A group is
// lcm(Z,32) consecutive threads: 1 warp = 2 codewords for Z=16, 3 warps = 4 codewords for Z=24.  Groups are
// fully independent: each owns its codewords' shared-memory slabs, its own mbarrier (bulk-TMA loads of the
// channel LLRs) and synchronises only inside the group (__syncwarp / a named barrier) between the VN and CN
// phases.  A group loops over its share of the batch (persistent); there is no CTA-wide barrier in steady
// state, so warps drift apart and FMA-heavy VN phases overlap ALU/LSU-heavy CN phases of other warps.
//
// Per codeword slab (floats): rows [0,N) channel LLR, rows [N,N+S) messages of edges whose variable block
// has degree >= 2 (in place: c2v after CN, v2c after VN).  The slab stride is == Z (mod 32), so lane
// (cw, z) sits in bank (cw*Z + z + const) mod 32: conflict-free for the un-rotated VN accesses and, because
// the circulant rotation permutes lanes inside a codeword's own Z-block, for the rotated CN accesses too
// when Z divides 32.
//
// Exactness contract: see nldpc_generic.cu / oracle/nldpc_oracle.c.
#pragma once
#include <cuda_fp16.h>

#include <type_traits>

#include "nldpc_common.cuh"

namespace nldpc {

// Learned weights live in constant memory for the duration of a launch: {w, b} pairs [T][E], read through the
// uniform datapath (LDCU) so they cost no LSU bandwidth.  The arena is a ring shared by in-flight launches
// (see ConstArena in nldpc_spec.cu); launches whose weights do not fit use the LDG variant (kConstW = false).
constexpr int kConstFloat2 = 7680;                 // 60 KB of the 64 KB constant bank
__constant__ float2 c_wb[kConstFloat2];   // one copy per translation unit that instantiates specialised kernels
// The Boosted decoders have weights only (no biases): their launches store plain floats in the arena (half the space: T * E
// floats, so an eager launch and a CUDA-graph-captured one of the training configuration fit side by side) and index it in
// float units; the Neural decoder stores {w, b} pairs and indexes in float2 units.
template <bool kScalarW>
__device__ __forceinline__ float2 wb_at(int idx) {
    if constexpr (kScalarW) return make_float2(reinterpret_cast<const float *>(c_wb)[idx], 0.0f);
    else return c_wb[idx];
}

// Runtime descriptors of the checks that run as LOOPS (G::loop_desc(), one 32-bit word per edge; backward sweep and the
// training-mode forward), per translation unit like c_wb; graph slot * kDescStride is the base of a code's table
constexpr int kDescStride = 512;
__constant__ uint32_t c_desc[2 * kDescStride];

template <class G>
__device__ __forceinline__ int rot_lane(int z, int s) {      // (z + s) mod Z
    int zz = z + s;
    if constexpr ((G::Z & (G::Z - 1)) == 0) return zz & (G::Z - 1);
    else return zz >= G::Z ? zz - G::Z : zz;
}

// Training-mode forward: a graph type wrapped in Train<> (below, after slab_floats) keeps the channel LLRs of ALL blocks in
// shared rows — the register-resident ones of the decode kernels cannot be indexed by a loop — so that the structurally
// identical "extension" checks run as loops over runtime descriptors like in the backward sweep.  The fully unrolled
// every-iteration body with dump, loss and export code in it was 12.6 K instructions = 197 KB and instruction-delivery bound
// (ncu: no_instruction 3.3 of 8 stall cycles per issue, issue slots 25 % busy; profiles/r02_ncu_train_forward_unrolled_summary.txt).
template <class G, class = void>
struct train_traits {
    static constexpr bool on = false;
};
template <class G>
struct train_traits<G, std::void_t<decltype(G::kTrainVariant)>> {
    static constexpr bool on = G::kTrainVariant;
};
// Neural decoder: the every-iteration kernel that also writes the training dump is its own instantiation (Dumping<G>), so that
// the list-mode decode carries no per-check branch (basic-block boundaries between checks cost the cross-check scheduling)
template <class G0>
struct Dumping : G0 {
    static constexpr bool kDumpVariant = true;
};
template <class G, class = void>
struct dump_traits {
    static constexpr bool on = false;
};
template <class G>
struct dump_traits<G, std::void_t<decltype(G::kDumpVariant)>> {
    static constexpr bool on = G::kDumpVariant;
};

// List mode as forward() uses it asks for soft outputs only: SoftOnly<G> compiles the hard-decision code out of the 52
// emission sites per iteration (predicated off they still issue: 312 of 3 613 instructions per warp-iteration)
template <class G0>
struct SoftOnly : G0 {
    static constexpr bool kNoHardVariant = true;
};
template <class G, class = void>
struct nohard_traits {
    static constexpr bool on = false;
};
template <class G>
struct nohard_traits<G, std::void_t<decltype(G::kNoHardVariant)>> {
    static constexpr bool on = G::kNoHardVariant;
};

#ifdef NLDPC_DEBUG_COUNT
__device__ unsigned g_dbg_restarts;
extern "C" unsigned nldpc_debug_restarts() { unsigned v = 0; cudaMemcpyFromSymbol(&v, g_dbg_restarts, 4); return v; }
#endif
// scheduling knobs (experiments: tools/build_variant.sh)
#ifndef NLDPC_TRAIN_GANG
#define NLDPC_TRAIN_GANG 2  // looped checks of the training forward processed together (independent arithmetic chains side by side)
#endif
#ifndef NLDPC_VN_QOUTER
#define NLDPC_VN_QOUTER 1   // 1: VN chain pairs written operand-major
#endif
#ifndef NLDPC_LLR_VEC
#define NLDPC_LLR_VEC 1     // Boosted state export with 16-byte stores after the CN phase (LlrExport) when the rows allow it
#endif
#ifndef NLDPC_LLR_VEC_LAST
#define NLDPC_LLR_VEC_LAST 0     // ... also for the single export of the last-iteration (throughput) kernels: OFF.  With it the
#endif                           // WiMAX throughput kernel wrote its packed decisions to a wrong address although the export
                                 // itself never ran (decode_hard passes no llr pointer) — not understood, and one export per
                                 // decode is not worth finding out: the every-iteration kernels are where the 16.5 GB go.
                                 // (Seen with the first version of the export code only: on the final code a variant build
                                 // with this switch ON passes all Boosted / staging GPU tests.  Unexplained, hence still OFF.)
#ifndef NLDPC_PIPE_CN
#define NLDPC_PIPE_CN 1     // 1: the next check's inputs are loaded before the current check computes
#endif
#ifndef NLDPC_PIPE_VN
#define NLDPC_PIPE_VN 1     // 1: the next variable block's messages are loaded before the current block computes
#endif

// one edge of a check row: slab row accessed by the CN phase, circulant shift, row-major edge index
// (index into the weight vectors), and the variable block when it has degree 1 (message not stored).
template <int ROW, int SHIFT, int EIDX, int COL1>
struct Ed {
    static constexpr int row = ROW, shift = SHIFT, e = EIDX, col1 = COL1;
};

// ---- packed helpers ----------------------------------------------------------------------------------------------
// f2 = two fp32 values {low word, high word} in one 64-bit register pair; add2 is PTX add.rn.f32x2 (SASS FADD2):
// two individually rounded IEEE additions in one issue slot.  It is used to advance TWO adjacent chains of the
// exact-order VN sums of the same codeword at once.
// NOTE: there is deliberately no packed multiply: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into ONE FFMA2
// (single rounding) even with explicit .rn and -fmad=false, which would break the bit-exactness of |o|*w + b.
using f2 = unsigned long long;
__device__ __forceinline__ f2 pack2(float a, float b) { return (f2)__float_as_uint(a) | ((f2)__float_as_uint(b) << 32); }
__device__ __forceinline__ float lo(f2 v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(f2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// packed multiply (SASS FMUL2).  Only ever used where NO addition consumes the product (see the note above): the sign of a
// c2v message, relu(m) * (+-1.0).
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {      // FADD2 with a negated operand
    f2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
#ifndef NLDPC_VN_QMS_TOTAL      // QMS q=5 VN phase: (x + total) - m[k] instead of the ordered chains (exact either way)
#define NLDPC_VN_QMS_TOTAL 1
#endif
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));   // one FMNMX3
    return d;
}

// decoder arithmetic selected at compile time: 0 = NeuralLDPCDecoder, 1 = Boosted MS, 2 = Boosted QMS with q_bit = 5
// (other q-bit grids, SP and the UCN indicator run on the table-driven kernel)
__device__ __forceinline__ float clamp_rng(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float quant5(float x) { return clamp_rng(mulf(rintf(mulf(x, 2.0f)), 0.5f), -7.5f, 7.5f); }   // (:190-191)
// The same value on the full-rate FADD pipe: clamp first (identical result: rint(2x)/2 is monotone and fixes +-7.5), then
// round to the 0.5 grid by adding and subtracting 1.5 * 2^22, whose ulp is 0.5 and whose mantissa is even in that unit, so
// the sum rounds half-to-even exactly like rint(2x).  Differs from quant5 only in the SIGN of a zero result (always +0.0
// here, -0.0 from rint for -0.25 <= x < 0): used where the value is next compared, min-ed or added to a +0-started sum.
__device__ __forceinline__ float quant5_grid(float x) {
    constexpr float kMagic = 6291456.0f;
    return addf(addf(clamp_rng(x, -7.5f, 7.5f), kMagic), -kMagic);
}
template <int MODE>
__device__ __forceinline__ float condition(float x, float lo, float hi) {        // QMS: quantise, MS: clamp (:386-389, :507-510)
    if constexpr (MODE == 2) return quant5(x);
    else return clamp_rng(x, lo, hi);
}

// Boosted output: clamp(xa_origin + total, range) (:520-521); the training dump records whether the clamp passes gradient
// Fused training forward (DecodeArgs::ybits): the value that leaves is dL/dout = c_t * upstream / n * (sigmoid(out) - y), zeroed
// where the clamp blocks the gradient, and the loss terms max(x,0) - x*y + log(1 + exp(-|x|)) (LDPCDecoderLoss.py:73-108, BCE
// branch; y is 0 or 1) are summed per lane: accA the max() part, accL log2(1 + exp(-|x|)); the kernel folds both with c_t
// after each phase.  MUFU approximations (ex2 / rcp / lg2, relative error ~2^-22): this is the floating-point part of the
// path, compared with the reference's autograd at 3e-5.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <class L>
__device__ __forceinline__ float bce_fused(L &c, int q, float out, bool pass) {
    const bool y1 = (c.yb[q >> 5] >> (q & 31)) & 1u;
    const float e = ex2_approx(-1.4426950408889634f * fabsf(out));
    const float d = 1.0f + e;
    const float r = rcp_approx(d);
    float s = (out >= 0.0f) ? r : e * r;             // sigmoid(out)
    s = y1 ? s - 1.0f : s;
    c.accA += fmaxf(y1 ? -out : out, 0.0f);
    c.accL += lg2_approx(d);
    return pass ? s * c.cg : 0.0f;
}
template <class L>
__device__ __forceinline__ float boosted_out(L &c, int q, float xo, float tot) {
    const float sum = addf(xo, tot);
    const float out = clamp_rng(sum, c.lo, c.hi);
    if constexpr (train_traits<typename L::GraphT>::on) {
        const bool pass = sum >= c.lo && sum <= c.hi;
        if (c.yb) return bce_fused(c, q, out, pass);      // fused training: dL/dout leaves instead of out (all lanes of a launch)
        if (c.mask) c.mask[q] = pass ? 1 : 0;             // two-call training: dL/dout arrives from autograd, the mask gates it
    }
    return out;
}

// Where xa_origin lives when VN weights make the on-chip channel value (xa_input) drift away from it — the `kXo` parameter:
//   0  no VN weights: xa_origin == xa_input, nothing extra;
//   1  its own N shared rows per codeword (list mode: every iteration produces marginals, the rows are read T times);
//   2  not kept at all (throughput mode: marginals only after the last iteration): the raw value is read again from global
//      memory — an L2 hit, the codeword was bulk-loaded moments ago — and quantised if QMS (:517-518); the slab stays at
//      its plain size, so the CTA keeps all its codewords (BG2: 8 groups instead of 7).
template <int MODE, class L>
__device__ __forceinline__ float xo_global(const L &c, int q) {
    const float v = __ldg(c.xa_cw + q);
    if constexpr (MODE == 2) return quant5_grid(v);      // (+0.0 where quant5 gives -0.0: absorbed by the sum it enters, xo + (0 + ...))
    else return v;
}

// List mode (soft outputs after every iteration): a graph type wrapped in Staged<> makes the emit functions write each
// marginal into a staging row that follows the codeword's slab (N rows of Z floats at float offset kStageOff, an
// immediate), and the kernel ships the finished [N*Z] row of a codeword-iteration to global memory with ONE bulk TMA store
// (cp.async.bulk.global.shared::cta) instead of N per-lane 4-byte STG instructions per iteration.
template <class G0, int kOff>
struct Staged : G0 {
    static constexpr bool kStageOut = true;
    static constexpr int kStageOff = kOff;
};
template <class G, class = void>
struct stage_traits {
    static constexpr bool on = false;
    static constexpr int off = 0;
};
template <class G>
struct stage_traits<G, std::void_t<decltype(G::kStageOut)>> {
    static constexpr bool on = G::kStageOut;
    static constexpr int off = G::kStageOff;
};

template <int Z>
struct GroupShape {
    static constexpr int kLanes = (Z == 16 || Z == 32) ? 32 : (Z == 24 ? 96 : 0);
    static constexpr int kCw = kLanes / Z;      // codewords per group
    static constexpr int kWarps = kLanes / 32;
    static_assert(kLanes != 0, "specialised kernels exist for Z in {16, 24, 32}");
    // Which (codeword, lane z) a thread of the group works on.  Z | 32: consecutive threads = consecutive lanes of a codeword.
    // Z = 24 (4 codewords on 3 warps): warp w of the group takes lanes 8w .. 8w+7 of ALL four codewords.  With the slab stride
    // == 24 (mod 32) the four codewords' rows start 24 banks apart, so the 4 x 8 lanes of a warp fall on four disjoint 8-bank
    // windows — for the un-rotated accesses and, because a rotation moves all four 8-lane runs by the same amount (a run that
    // wraps at z = 24 continues 8 banks further, exactly where the next codeword's window ended), for every circulant shift.
    // The codeword-major mapping left warps straddling codewords: 35 % of all shared wavefronts were bank-conflict replays (ncu).
    __device__ static __forceinline__ void map(int gl, int &cwl, int &z) {
        if constexpr (Z == 24) {
            cwl = (gl & 31) >> 3;
            z = ((gl >> 5) << 3) | (gl & 7);
        } else {
            cwl = gl / Z;
            z = gl - cwl * Z;
        }
    }
};

template <int kLanes>
__device__ __forceinline__ void group_sync(int group_in_cta) {
    if constexpr (kLanes == 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(group_in_cta + 1), "n"(kLanes) : "memory");
    }
}

// OR of `pred` over the lanes of the group; for multi-warp groups this is a reducing named barrier (bar.red.or), i.e. it
// also orders the group's shared-memory accesses like group_sync
template <int kLanes>
__device__ __forceinline__ bool group_any(int group_in_cta, bool pred) {
    if constexpr (kLanes == 32) {
        return __any_sync(0xffffffffu, pred);
    } else {
        int r;
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t"
            "setp.ne.s32 p, %1, 0;\n\t"
            "bar.red.or.pred q, %2, %3, p;\n\t"
            "selp.s32 %0, 1, 0, q;\n\t}"
            : "=r"(r)
            : "r"((int)pred), "r"(group_in_cta + 1), "n"(kLanes)
            : "memory");
        return r != 0;
    }
}

// -----------------------------------------------------------------------------------------------------------
// Per-thread state of the Neural decode (NeuralLDPCDecoder.py:44-100).
template <class G>
struct NeuralLane {
    using GraphT = G;
    static constexpr int Z = G::Z, N = G::N, NZ = G::N * G::Z;
    float *lane;             // &slab[z]               un-rotated accesses (VN phase)
    float *rot[Z];           // &slab[(z + s) mod Z]   rotated accesses   (CN phase), indexed by the immediate shift
    const float *wt, *bt;    // weights_var[t], biases_var[t]          (kConstW == false)
    int wb_base;             // offset of {w,b}[t][0] in the constant arena (kConstW == true)
    float *soft;             // &soft_t[b][0] of the iteration being emitted, or nullptr (also when b >= B)
    uint8_t *hb;             // this codeword's hard-decision staging bytes in shared memory (N*Z/8), or nullptr
    float xreg[G::kXRegs > 0 ? G::kXRegs : 1];   // channel LLRs of identity-circulant degree-1 blocks (lane-private)
    int z;
    int grp;                 // group within the CTA (named barrier id for multi-warp groups)
    bool gl0;                // lane 0 of the group: issues the bulk loads / stores
    bool valid;
    float zmin;              // smallest |CN input| this lane has seen in the current unit (0 => the unit needs the zero-safe CN phase)
    // Boosted decoder (MODE != 0)
    int xo_off;              // float offset from the xin rows to the xo rows (0: xa_origin and xa_input are the same rows)
    const float *xa_cw;      // &xa[b][0] in global memory (kXo == 2)
    float lo, hi;            // allowed_llr_range
    float *llr_last;         // &llr_last[b][0][0] ([Z][llr_pitch]) while the last iteration's CN phase runs, else nullptr
    float *llr_d1;           // same tensor, used by the degree-1 edges of the unrolled CN phase (== llr_last, except while the
                             // stored edges leave through LlrExport: then llr_last is nullptr and only these keep their store)
    int llr_pitch;           // row pitch of the state tensors (floats, >= E; DecodeArgs::llr_pitch)
    float *llr_row;          // inline vector export (llr_inline<G>): &llr[..][b][z][0] of this lane's own row, or nullptr
    float llr_q[3];          // ... the open group of 4 consecutive edges
    bool llr_inl;            // ... launch-uniform: this launch exports that way (every lane takes the extra warp sync)
    bool llr_sc;             // warp-uniform: some lane may hold llr_last / llr_d1 (the scalar export runs behind ONE uniform branch per
                             // check; as a per-edge predicate it cost 9 % of the executed instructions of launches that never store)
    uint8_t *mask;           // training dump: &hist_mask[t_emit][b][0] of the iteration being emitted, or nullptr
    char *dump;              // training dump, check-packed format: this codeword's records of the running iteration, or nullptr
    const uint32_t *yb;      // fused loss: this codeword's packed label bits (shared memory), or nullptr
    float cg;                // fused loss: c_t * upstream / n of the iteration being emitted
    float accA, accL;        // fused loss: partial sums of the running phase (see bce_fused)

    // ---- emission of one marginal value -------------------------------------------------------------------
    // un-rotated: this lane holds bit (J, z)
    template <int J>
    __device__ __forceinline__ void emit(float v) {
        if constexpr (stage_traits<G>::on) {
            if (soft) lane[stage_traits<G>::off + J * Z] = v;
        } else {
            if (soft) st_global_stream(soft + J * Z + z, v);
        }
        if constexpr (train_traits<G>::on || nohard_traits<G>::on) return;      // (training / soft-only list mode: no hard decisions)
        if (hb) {
            if constexpr (Z == 16 || Z == 32) {
                const unsigned bal = __ballot_sync(0xffffffffu, v < 0.0f);
                if constexpr (Z == 16) {
                    if (z == 0) reinterpret_cast<uint16_t *>(hb)[J] = (uint16_t)((threadIdx.x & 16) ? (bal >> 16) : (bal & 0xffffu));
                } else {
                    if (z == 0) reinterpret_cast<uint32_t *>(hb)[J] = bal;
                }
            } else {
                // Z = 24: the warp holds lanes 8w .. 8w+7 of four codewords (GroupShape::map): 8 ballot bits = one whole byte
                // of the codeword's packed row
                static_assert(Z % 8 == 0, "byte-aligned lane runs");
                const unsigned bal = __ballot_sync(0xffffffffu, v < 0.0f);
                if ((threadIdx.x & 7) == 0) hb[(J * Z + z) >> 3] = (uint8_t)(bal >> (threadIdx.x & 24));
            }
        }
    }
    // rotated by SHIFT: this lane (check lane h = z) holds bit (J, (h + SHIFT) mod Z)
    // (J and SHIFT are compile-time constants after unrolling/inlining)
    __device__ __forceinline__ void emit_rot(const int J, const int SHIFT, float v) {
        const int zz = (int)(rot[SHIFT] - (lane - z));      // (z + SHIFT) mod Z
        if constexpr (stage_traits<G>::on) {
            if (soft) rot[SHIFT][stage_traits<G>::off + J * Z] = v;
        } else {
            if (soft) st_global_stream(soft + J * Z + zz, v);
        }
        if constexpr (train_traits<G>::on || nohard_traits<G>::on) return;
        if (hb) {
            if constexpr (Z == 16 || Z == 32) {
                const unsigned bal = __ballot_sync(0xffffffffu, v < 0.0f);
                if constexpr (Z == 16) {
                    unsigned w = (threadIdx.x & 16) ? (bal >> 16) : (bal & 0xffffu);
                    w = ((w << SHIFT) | (w >> (16 - SHIFT))) & 0xffffu;      // bit h -> bit (h + SHIFT) mod 16
                    if (z == 0) reinterpret_cast<uint16_t *>(hb)[J] = (uint16_t)w;
                } else {
                    if (z == 0) reinterpret_cast<uint32_t *>(hb)[J] = __funnelshift_l(bal, bal, SHIFT);
                }
            } else {
                if (v < 0.0f) atomicOr(reinterpret_cast<unsigned *>(hb) + ((J * Z + zz) >> 5), 1u << ((J * Z + zz) & 31));
            }
        }
    }
};

// all lanes of the group: wait until the bulk stores issued from the staging rows have READ them (they may be overwritten)
template <class G>
__device__ __forceinline__ void stage_wait(const NeuralLane<G> &c) {
    if constexpr (stage_traits<G>::on) {
        if (c.gl0) tma_store_wait_read();
        group_sync<GroupShape<G::Z>::kLanes>(c.grp);
    }
}

// ---- VN phase functors (one `col<J, R...>()` call per variable block of degree >= 2) --------------------------
// iteration 0: all c2v are zero -> v2c = xa + 0  (:49, :56-58)
template <class G>
struct VnFirst {
    NeuralLane<G> &c;
    template <int J, int XROW, int... R>
    __device__ __forceinline__ void col() {
        const float v = addf(c.lane[XROW * G::Z], 0.0f);
        ((c.lane[R * G::Z] = v), ...);
    }
};

// iterations >= 1: v2c[k] = x + (((0 + c[0]) + c[1]) + ... skipping k); kEmit: marginal of the previous iteration
template <class G, bool kEmit, int MODE = 0, int kXo = 0>
struct VnStep {
    NeuralLane<G> &c;
    float mm[2][G::kMaxColDeg];     // messages of the block being computed / the block being fetched (registers)
    float xx[2];
    template <int SLOT, int J, int XROW, int... R>
    __device__ __forceinline__ void pld() {
        constexpr int D = sizeof...(R);
        constexpr int rows[D] = {R...};
#pragma unroll
        for (int k = 0; k < D; k++) mm[SLOT][k] = c.lane[rows[k] * G::Z];
        xx[SLOT] = c.lane[XROW * G::Z];
    }
    template <int J, int XROW, int... R>
    __device__ __forceinline__ void col() {
        pld<0, J, XROW, R...>();
        pcol<0, J, XROW, R...>();
    }
    template <int SLOT, int J, int XROW, int... R>
    __device__ __forceinline__ void pcol() {
        constexpr int D = sizeof...(R);
        constexpr int rows[D] = {R...};
        const float *m = mm[SLOT];
        const float x = xx[SLOT];
#if NLDPC_VN_QMS_TOTAL
        if constexpr (MODE == 2) {
            // QMS q=5: the channel value and every message are multiples of 0.5 with |.| <= 7.5, so every partial sum of the
            // reference's ordered chain (:49, :56-58) is exact — and so is ANY order.  v2c[k] = (x + total) - m[k]: D/2 packed
            // adds for the total and D/2 packed subtractions instead of D^2/4 chain steps (BG2: blocks of degree up to 23).
            // Zeros: messages may be -0.0, the total starts from +0.0 like the reference's sums, x is +0.0 for zeros
            // (quant5_grid), so x + total is never -0.0 and (x + total) - m[k] is +0.0 whenever it is zero — as x + (+0-started
            // sum of the others) is.
            f2 acc = pack2(0.0f, 0.0f);
#pragma unroll
            for (int k = 0; k + 1 < D; k += 2) acc = add2(acc, pack2(m[k], m[k + 1]));
            float p = addf(lo(acc), hi(acc));
            if constexpr (D & 1) p = addf(p, m[D - 1]);
            const float xt = addf(x, p);
#pragma unroll
            for (int k = 0; k + 1 < D; k += 2) {
                const f2 s = sub2(pack2(xt, xt), pack2(m[k], m[k + 1]));
                c.lane[rows[k] * G::Z] = lo(s);
                c.lane[rows[k + 1] * G::Z] = hi(s);
            }
            if constexpr (D & 1) c.lane[rows[D - 1] * G::Z] = addf(xt, -m[D - 1]);
            if constexpr (kEmit)
                c.template emit<J>(boosted_out(c, J * G::Z + c.z, kXo == 2 ? xo_global<MODE>(c, J * G::Z + c.z) : (kXo == 1 ? c.lane[c.xo_off + J * G::Z] : x), p));   // Boosted :520-521
            return;
        }
#endif
        float pre[D];
        float p = 0.0f;                       // running prefix ((0 + m0) + m1) + ...
#pragma unroll
        for (int k = 0; k < D; k++) {
            pre[k] = p;
            p = addf(p, m[k]);
        }
        // chains k and k+1 advance together on the packed pipe: chain k is pre[k] + m[k+1] + m[k+2] + ...,
        // chain k+1 is pre[k+1] + m[k+2] + ...; from m[k+2] on both add the same operand (FADD2, broadcast).
#if NLDPC_VN_QOUTER
        // the same chains, written operand-major: all live chain pairs take m[q] before any takes m[q+1], so consecutive
        // instructions are independent (each chain is a serial dependency of 4-cycle adds)
        constexpr int P = D / 2;
        f2 sp[P > 0 ? P : 1];
#pragma unroll
        for (int kk = 0; kk < P; kk++) sp[kk] = pack2(addf(pre[2 * kk], m[2 * kk + 1]), pre[2 * kk + 1]);
#pragma unroll
        for (int q = 2; q < D; q++) {
            const f2 mq = pack2(m[q], m[q]);
#pragma unroll
            for (int kk = 0; kk < P; kk++)
                if (2 * kk + 2 <= q) sp[kk] = add2(sp[kk], mq);
        }
#pragma unroll
        for (int kk = 0; kk < P; kk++) {
            const f2 s = add2(pack2(x, x), sp[kk]);
            c.lane[rows[2 * kk] * G::Z] = lo(s);
            c.lane[rows[2 * kk + 1] * G::Z] = hi(s);
        }
#else
#pragma unroll
        for (int k = 0; k + 1 < D; k += 2) {
            f2 s = pack2(addf(pre[k], m[k + 1]), pre[k + 1]);
#pragma unroll
            for (int q = k + 2; q < D; q++) s = add2(s, pack2(m[q], m[q]));
            s = add2(pack2(x, x), s);
            c.lane[rows[k] * G::Z] = lo(s);
            c.lane[rows[k + 1] * G::Z] = hi(s);
        }
#endif
        if constexpr (D & 1) {
            const float v = addf(x, pre[D - 1]);
            c.lane[rows[D - 1] * G::Z] = v;
        }
        if constexpr (kEmit) {
            if constexpr (MODE == 0) c.template emit<J>(addf(x, p));       // out = xa + llr @ W_output (:94-96)
            else c.template emit<J>(boosted_out(c, J * G::Z + c.z, kXo == 2 ? xo_global<MODE>(c, J * G::Z + c.z) : (kXo == 1 ? c.lane[c.xo_off + J * G::Z] : x), p));   // Boosted :520-521
        }
    }
};

template <class G, class F>
__device__ __forceinline__ void run_vcols(F &f) {
#if NLDPC_PIPE_VN
    G::vcols_pipelined(f);
#else
    G::vcols(f);
#endif
}

// marginal only (after the last CN phase)
template <class G, int MODE = 0, int kXo = 0>
struct Marginal {
    NeuralLane<G> &c;
    template <int J, int XROW, int... R>
    __device__ __forceinline__ void col() {
        float p = 0.0f;
        ((p = addf(p, c.lane[R * G::Z])), ...);
        if constexpr (MODE == 0) c.template emit<J>(addf(c.lane[XROW * G::Z], p));
        else c.template emit<J>(boosted_out(c, J * G::Z + c.z, kXo == 2 ? xo_global<MODE>(c, J * G::Z + c.z) : c.lane[(kXo == 1 ? c.xo_off + J * G::Z : XROW * G::Z)], p));
    }
};

// where each block's channel LLR lives: shared row DEST >= 0 or lane register -(1 + DEST)
template <class G, int DEST>
__device__ __forceinline__ float &xa_ref(NeuralLane<G> &c) {
    if constexpr (DEST >= 0) return c.lane[DEST * G::Z];
    else if constexpr (train_traits<G>::on) return c.lane[(G::kXRows + G::S + (-DEST - 1)) * G::Z];   // Train<>: behind the message rows
    else return c.xreg[-DEST - 1];
}

// after the bulk-TMA load: move this lane's raw channel LLRs from the staging area (message rows) to their places.
// MODE 0 also screens for exact zeros; Boosted quantises xa_origin once if QMS (:517-518) and, with VN weights (kXo),
// keeps it in its own rows while xa_input starts as the raw value.
template <class G, int MODE, int kXo>
struct PlaceXa {
    NeuralLane<G> &c;
    const float *stage;      // &staging[z]
    float zm;
    template <int J, int DEST>
    __device__ __forceinline__ void put() {
        const float v = stage[J * G::Z];
        if constexpr (MODE == 0) {
            zm = fminf(zm, fabsf(v));
            xa_ref<G, DEST>(c) = v;
        } else {
            const float xq = (MODE == 2) ? quant5_grid(v) : v;      // (+0.0 zeros: the CN phase reads raw sign words)
            if constexpr (kXo != 0) {
                if constexpr (kXo == 1) c.lane[c.xo_off + J * G::Z] = xq;
                xa_ref<G, DEST>(c) = v;
            } else {
                xa_ref<G, DEST>(c) = xq;
            }
        }
    }
};

// xa_input *= w_VN(t), re-quantised if QMS (:325-337); every element is private to its lane
template <class G, int MODE>
struct ScaleXin {
    NeuralLane<G> &c;
    const float *vw;
    template <int J, int DEST>
    __device__ __forceinline__ void put() {
        float x = mulf(xa_ref<G, DEST>(c), __ldg(vw + J));
        // quant5_grid: +0.0 where quant5 gives -0.0.  The value is only ever added to +0-started sums, quantised again by the CN
        // phase, or multiplied — a zero's sign reaches no output.
        if constexpr (MODE == 2) x = quant5_grid(x);
        xa_ref<G, DEST>(c) = x;
    }
};

// training dump of the channel-input state (hist_xin[.][b][J][z])
template <class G>
struct DumpXin {
    NeuralLane<G> &c;
    float *dst;              // &hist_xin[t][b][0][z]
    template <int J, int DEST>
    __device__ __forceinline__ void put() {
        dst[J * G::Z] = xa_ref<G, DEST>(c);
    }
};

// refill the lane registers from global memory (out-of-line paths that cannot receive them by value)
template <class G>
struct ReloadXreg {
    NeuralLane<G> &c;
    const float *xa_lane;    // &xa[b][0][z]
    template <int J, int DEST>
    __device__ __forceinline__ void put() {
        if constexpr (DEST < 0) c.xreg[-DEST - 1] = __ldg(xa_lane + J * G::Z);
    }
};

// ---- CN phase (one `chk<Ed...>()` call per check row), NeuralLDPCDecoder.py:59-91 -----------------------------
// kEmit: also produce the marginals of degree-1 variable blocks (their c2v is needed for nothing else).
// kZeroSafe = false is the fast path: it assumes that no CN input of this warp is exactly 0.  The VN phase tracks
// min |v2c| (half an instruction per edge) and the channel LLRs are screened once per codeword; when a zero is seen
// the warp runs the out-of-line kZeroSafe = true phase, which applies the reference's "exact zero -> magnitude 10000,
// not positive" rule (:74, :78) at two extra instructions per edge.  Exact-zero v2c only occur for punctured /
// quantised channel values, so the hot loop does not pay for them, and the whole CN phase stays one basic block.
// gather of one check's inputs (:59-63): raw[k] = v2c of edge k at this lane's rotated position (or the lane register of a
// degree-1 identity block).  Separate from the arithmetic so that the NEXT check's loads can be issued before this check's
// stores (software pipelining, NLDPC_PIPE): ptxas cannot hoist them itself — every access goes through one of Z rotated base
// pointers and a store through one may alias a load through another as far as it can tell.
// The same stage also fetches the edges' {w, b} pairs: the constant load (LDC) has its own latency, and issued right before
// the multiply that needs it the multiply stalled on it (ncu: short_scoreboard on FMUL, 13 % of all stall samples).
// training dump, check-packed (hist_fmt 1): the check's CN inputs as this check lane read them, one contiguous record per lane
// ([record][Z][P]).  QMS q=5: every CN input is a sum of multiples of 0.5 (quantised channel value + quantised messages,
// |sum| <= 7.5 * (kMaxColDeg + 1)), i.e. exact in fp16 -> records of 4 / 8 / 16 halfs; otherwise D floats.
template <class G, int MODE, int D>
__device__ __forceinline__ void dump_record(const NeuralLane<G> &c, const float *raw, int off) {      // off: elements per lane
    if constexpr (MODE == 2) {
        constexpr int P = G::dump_slots_h(D);
        uint32_t w[P / 2];
#pragma unroll
        for (int i = 0; i < P / 2; i++) {
            const __half2 h = __floats2half2_rn(2 * i < D ? raw[2 * i] : 0.0f, 2 * i + 1 < D ? raw[2 * i + 1] : 0.0f);
            w[i] = *reinterpret_cast<const uint32_t *>(&h);
        }
        char *p = c.dump + ((size_t)off * G::Z + (size_t)c.z * P) * 2;
        if constexpr (P == 4) {
            __stcs(reinterpret_cast<uint2 *>(p), make_uint2(w[0], w[1]));
        } else {
#pragma unroll
            for (int i = 0; i < P / 8; i++) __stcs(reinterpret_cast<uint4 *>(p) + i, make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
        }
    } else {
        float *p = reinterpret_cast<float *>(c.dump) + (size_t)off * G::Z + (size_t)c.z * D;
#pragma unroll
        for (int k = 0; k < D; k++) __stcs(p + k, raw[k]);
    }
}
template <class G, int MODE, class... Es>
__device__ __forceinline__ void dump_check(const NeuralLane<G> &c, const float *raw) {
    constexpr int D = sizeof...(Es);
    constexpr int eix[D] = {Es::e...};
    constexpr int off = MODE == 2 ? G::dump_off_h(eix[0]) : G::dump_off_f(eix[0]);
    static_assert(off >= 0, "check not in the dump table");
    dump_record<G, MODE, D>(c, raw, off);
}

// ---- pair layout of the Neural decoder's {w, b} (NLDPC_CN_PAIR) ----------------------------------------------------
// The STORED edges of a check are taken two at a time.  A pair (a, b) owns two consecutive float2 entries of the arena,
// (w_a, w_b) and (b_a, b_b): the two products go to a register pair with scalar FMULs, ONE packed add (FADD2) adds the b
// pair straight from a uniform register pair, and after the ReLU one packed multiply (FMUL2) by (+-1.0, +-1.0) applies
// the two signs: 10 instructions per pair instead of 12 (two LOP3 fewer on the half-rate ALU pipe).  The leftover stored
// edge and the unstored degree-1 edges keep a (w, b) entry.  Offsets = tools/gen_kernels.py: G::wb_pair_off() (host side).
#ifndef NLDPC_CN_PAIR
#define NLDPC_CN_PAIR 1
#endif
#ifndef NLDPC_CN_SIGNMUL      // 1: sign applied by the packed multiply (FMA pipe); 0: OR-ed in (LOP3, ALU pipe)
#define NLDPC_CN_SIGNMUL 0
#endif
// Boosted decoders (plain float weights in the arena, MODE != 0): an iteration's E weights start at an EVEN float offset
// (pitch kWPitch), so the weights of edges (e, e + 1), e even, are one aligned 64-bit constant: QMS q=5 multiplies two stored
// edges of a check with ONE FMUL2 (uniform register pair operand) and quantises both with two FADD2 (NLDPC_QMS_PAIR).
#ifndef NLDPC_QMS_PAIR
#define NLDPC_QMS_PAIR 1
#endif
template <class G>
constexpr int kWPitch = (G::E + 1) & ~1;
// edge k is the first of a QMS pair: even edge index, the next edge belongs to the same check, both are computed now
template <bool kEmit, int D>
__host__ __device__ constexpr bool qms_pair_first(const int *eix, const int *col1, int k) {
    return (eix[k] & 1) == 0 && k + 1 < D && (kEmit || (col1[k] < 0 && col1[k + 1] < 0));
}
template <int D>
struct WbPlanT {
    int entry[D];      // float2 entry (relative to the iteration's base): pair -> entry of (w_a, w_b), +1 = (b_a, b_b); single -> (w, b)
    int role[D];       // 0 single, 1 first of a pair, 2 second of a pair
    int mate[D];       // the other edge of the pair
};
template <class... Es>
struct WbPlan {
    static constexpr int D = sizeof...(Es);
    __host__ __device__ static constexpr WbPlanT<D> make() {
        constexpr int col1[D] = {Es::col1...};
        constexpr int eix[D] = {Es::e...};
        WbPlanT<D> p{};
        int entry = eix[0], prev = -1;
        for (int k = 0; k < D; k++) {
            p.role[k] = 0;
            p.mate[k] = -1;
            if (col1[k] >= 0) continue;
            if (prev < 0) {
                prev = k;
            } else {
                p.entry[prev] = p.entry[k] = entry;
                p.role[prev] = 1;
                p.role[k] = 2;
                p.mate[prev] = k;
                p.mate[k] = prev;
                entry += 2;
                prev = -1;
            }
        }
        if (prev >= 0) p.entry[prev] = entry++;
        for (int k = 0; k < D; k++)
            if (col1[k] >= 0) p.entry[k] = entry++;
        return p;
    }
    __host__ __device__ static constexpr bool consecutive() {
        constexpr int eix[D] = {Es::e...};
        for (int k = 0; k < D; k++)
            if (eix[k] != eix[0] + k) return false;
        return true;
    }
    static_assert(consecutive(), "the edges of a check are consecutive in the weight vectors");
};

template <class G, bool kEmit, bool kConstW, bool kScalarW, bool kPairW, class... Es>
__device__ __forceinline__ void cn_load(const NeuralLane<G> &c, float *raw, float2 *wb) {
    constexpr int D = sizeof...(Es);
    constexpr int rows[D] = {Es::row...};
    constexpr int shf[D] = {Es::shift...};
    constexpr int eix[D] = {Es::e...};
    constexpr int col1[D] = {Es::col1...};
#pragma unroll
    for (int k = 0; k < D; k++)
        raw[k] = rows[k] >= 0 ? c.rot[shf[k]][rows[k] * G::Z]
                 : (train_traits<G>::on ? c.lane[(G::kXRows + G::S + (rows[k] < 0 ? -rows[k] - 1 : 0)) * G::Z] : c.xreg[rows[k] < 0 ? -rows[k] - 1 : 0]);
    if constexpr (NLDPC_CN_PAIR && !kScalarW) {
        // pair layout: wb[first of a pair] = (w_a, w_b), wb[second] = (b_a, b_b), wb[single] = (w, b)
        constexpr WbPlanT<D> pl = WbPlan<Es...>::make();
#pragma unroll
        for (int k = 0; k < D; k++) {
            if (col1[k] >= 0 && !kEmit) continue;
            if constexpr (kConstW) {
                wb[k] = c_wb[c.wb_base + pl.entry[k] + (pl.role[k] == 2 ? 1 : 0)];
            } else {
                if (pl.role[k] == 1) wb[k] = make_float2(__ldg(c.wt + eix[k]), __ldg(c.wt + eix[pl.mate[k] >= 0 ? pl.mate[k] : 0]));
                else if (pl.role[k] == 2) wb[k] = make_float2(__ldg(c.bt + eix[pl.mate[k] >= 0 ? pl.mate[k] : 0]), __ldg(c.bt + eix[k]));
                else wb[k] = make_float2(__ldg(c.wt + eix[k]), __ldg(c.bt + eix[k]));
            }
        }
        return;
    }
    if constexpr (NLDPC_QMS_PAIR && kScalarW && kConstW && kPairW) {
        // wb[first of a pair] = (w_k, w_k+1): one LDCU.64; everything else (w, 0)
        const int hb = c.wb_base >> 1;      // (wb_base is even: even arena offset + t * kWPitch)
#pragma unroll
        for (int k = 0; k < D; k++) {
            if (col1[k] >= 0 && !kEmit) continue;
            if (k > 0 && qms_pair_first<kEmit, D>(eix, col1, k - 1)) continue;      // second of a pair
            if (qms_pair_first<kEmit, D>(eix, col1, k)) wb[k] = c_wb[hb + eix[k] / 2];
            else wb[k] = wb_at<true>(c.wb_base + eix[k]);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < D; k++) {
        if (col1[k] >= 0 && !kEmit) continue;                            // unstored edge, marginal not wanted now
        if constexpr (kConstW) wb[k] = wb_at<kScalarW>(c.wb_base + eix[k]);
        else wb[k] = make_float2(__ldg(c.wt + eix[k]), __ldg(c.bt + eix[k]));
    }
}

template <class G, bool kEmit, bool kConstW, bool kZeroSafe, class... Es>
__device__ __forceinline__ void cn_check_core(NeuralLane<G> &c, const float *raw, const float2 *wb) {
    constexpr int D = sizeof...(Es);
    constexpr int rows[D] = {Es::row...};
    constexpr int shf[D] = {Es::shift...};
    constexpr int eix[D] = {Es::e...};
    constexpr int col1[D] = {Es::col1...};
    if constexpr (kEmit && dump_traits<G>::on) {       // (training dump of the Neural decoder: Dumping<G> instantiation)
        if (c.dump) dump_check<G, 0, Es...>(c, raw);
    }
    float u[D];
#pragma unroll
    for (int k = 0; k < D; k++) u[k] = (kZeroSafe && raw[k] == 0.0f) ? -10000.0f : raw[k];
    // min over the other edges, capped at 10000 (:74-75): pairwise prefix/suffix minima with 3-input FMNMX
    constexpr int H = (D + 1) / 2;
    float se[H + 1];
    se[H] = 10000.0f;
#pragma unroll
    for (int q = H - 1; q >= 0; q--) {
        if (2 * q + 1 < D) se[q] = fmin3(fabsf(u[2 * q]), fabsf(u[2 * q + 1]), se[q + 1]);
        else se[q] = fminf(fabsf(u[2 * q]), se[q + 1]);
    }
    // zero screen of the fast path: se[0] is the smallest |input| of the whole check, so one FMNMX per check keeps the
    // smallest CN input of the unit; the kernel looks at it once, after the last iteration (run_unit)
    if constexpr (!kZeroSafe) c.zmin = fminf(c.zmin, se[0]);
    unsigned x = (D & 1) ? 0x80000000u : 0u;
#pragma unroll
    for (int k = 0; k < D; k++) x ^= __float_as_uint(u[k]);
    float pe = 10000.0f;
#if NLDPC_CN_PAIR
    {
        constexpr WbPlanT<D> pl = WbPlan<Es...>::make();
        float mg[D], r[D], sg[D], cv[D];
#pragma unroll
        for (int k = 0; k < D; k++) {
            const int q = k >> 1;
            if ((k & 1) == 0) {
                if (k + 1 < D) mg[k] = fmin3(pe, fabsf(u[k + 1]), se[q + 1]);
                else mg[k] = fminf(pe, se[q + 1]);
            } else {
                mg[k] = fmin3(pe, fabsf(u[k - 1]), se[q + 1]);
                pe = fmin3(pe, fabsf(u[k - 1]), fabsf(u[k]));
            }
        }
        // sign as a factor: x1 = +-1.0 with the sign of the product of all inputs (negated for odd D, see x), and
        // sg[k] = x1 with the sign of input k taken out again = -(product of the signs of the OTHER inputs) (:77-80, :89-91)
        // NLDPC_CN_SIGNMUL == 2: the ReLU leaves the ALU pipe too: m + |m| = 2 relu(m) exactly (+0 for m <= 0, -0 included),
        // and the sign factor is +-0.5 (exact scaling)
        constexpr bool kReluAdd = NLDPC_CN_SIGNMUL == 2;
        const unsigned x1 = (x & 0x80000000u) | (kReluAdd ? 0x3f000000u : 0x3f800000u);
#pragma unroll
        for (int k = 0; k < D; k++) {
            if (col1[k] >= 0 && !kEmit) continue;                        // unstored edge, marginal not wanted now
            if constexpr (NLDPC_CN_SIGNMUL) sg[k] = __uint_as_float(x1 ^ (__float_as_uint(u[k]) & 0x80000000u));
            // |o| * w + b, ReLU
            if (pl.role[k] == 0) {
                const float m = addf(mulf(mg[k], wb[k].x), wb[k].y);
                r[k] = kReluAdd ? addf(m, fabsf(m)) : fmaxf(m, 0.0f);
            }
        }
#pragma unroll
        for (int k = 0; k < D; k++) {
            if (pl.role[k] != 1) continue;
            const int m2 = pl.mate[k] >= 0 ? pl.mate[k] : 0;
            const f2 s = add2(pack2(mulf(mg[k], wb[k].x), mulf(mg[m2], wb[k].y)), pack2(wb[m2].x, wb[m2].y));
            r[k] = kReluAdd ? addf(lo(s), fabsf(lo(s))) : fmaxf(lo(s), 0.0f);
            r[m2] = kReluAdd ? addf(hi(s), fabsf(hi(s))) : fmaxf(hi(s), 0.0f);
            if constexpr (NLDPC_CN_SIGNMUL) {
                const f2 v = mul2(pack2(r[k], r[m2]), pack2(sg[k], sg[m2]));
                cv[k] = lo(v);
                cv[m2] = hi(v);
            }
        }
#pragma unroll
        for (int k = 0; k < D; k++) {
            if (col1[k] >= 0 && !kEmit) continue;
            float c2v;
            if constexpr (NLDPC_CN_SIGNMUL) c2v = pl.role[k] == 0 ? mulf(r[k], sg[k]) : cv[k];
            else c2v = __uint_as_float(__float_as_uint(r[k]) | ((x ^ __float_as_uint(u[k])) & 0x80000000u));
            if (col1[k] < 0) {
                c.rot[shf[k]][rows[k] * G::Z] = c2v;                     // scatter back (:82-86), in place
            } else {
                // degree-1 block col1: out = xa + (0 + c2v) at lane (h + s) mod Z (:94-98)
                if constexpr (kEmit) c.emit_rot(col1[k], shf[k], addf(raw[k], addf(0.0f, c2v)));
            }
        }
        return;
    }
#endif
#pragma unroll
    for (int k = 0; k < D; k++) {
        const int q = k >> 1;
        float mag;
        if ((k & 1) == 0) {
            if (k + 1 < D) mag = fmin3(pe, fabsf(u[k + 1]), se[q + 1]);
            else mag = fminf(pe, se[q + 1]);
        } else {
            mag = fmin3(pe, fabsf(u[k - 1]), se[q + 1]);
            pe = fmin3(pe, fabsf(u[k - 1]), fabsf(u[k]));
        }
        if (col1[k] >= 0 && !kEmit) continue;                            // unstored edge, marginal not wanted now
        // |o| * w + b, ReLU, sign: negative iff the number of positive OTHER inputs is even (:77-80, :89-91)
        float m = addf(mulf(mag, wb[k].x), wb[k].y);
        m = fmaxf(m, 0.0f);
        const unsigned sb = (x ^ __float_as_uint(u[k])) & 0x80000000u;
        const float c2v = __uint_as_float(__float_as_uint(m) | sb);
        if (col1[k] < 0) {
            c.rot[shf[k]][rows[k] * G::Z] = c2v;                         // scatter back (:82-86), in place
        } else {
            // degree-1 block col1: out = xa + (0 + c2v) at lane (h + s) mod Z (:94-98)
            if constexpr (kEmit) c.emit_rot(col1[k], shf[k], addf(raw[k], addf(0.0f, c2v)));
        }
    }
}

// Boosted MS / QMS check update, BoostedNeuralLDPCDecoder.py:380-526 (no UCN mixing): condition the inputs (quantise or
// clamp), nudge exact zeros to +1e-4, min over the others, mag - 1e-4 [mag <= 1e-4], o = mag * sgn, |o| * W_cn, ReLU,
// condition again, * sign(o) (sign(0) = 0).  kXo: xa_origin lives in its own rows (VN weights make xa_input drift).
// Inline vector export of self.llr[t + 1] (see LlrExport below for the problem).  When a lane group is one warp (Z | 32) and the
// degree-1 blocks are lane-private, every check can export its own messages right after it has scattered them: one warp sync,
// then lane z reads the D stored messages of ITS row back from the slab, has the degree-1 message of its row still in a register
// (nothing is parked across the phase — 38 registers per lane for BG2, which spilled), and closes groups of 4 consecutive edges
// with one 16-byte store each (the open group, <= 3 values, rides in NeuralLane::llr_q across checks).
template <class G>
__device__ __forceinline__ constexpr bool llr_inline() {
    return NLDPC_LLR_VEC && GroupShape<G::Z>::kLanes == 32 && G::kDeg1Smem == 0 && G::kXRegs > 0 && !train_traits<G>::on;
}

template <class G, bool kEmit, int MODE, int kXo, class... Es>
__device__ __forceinline__ void cn_check_boosted_core(NeuralLane<G> &c, const float *raw, const float2 *wb) {
    constexpr int D = sizeof...(Es);
    float cv[kEmit ? D : 1]; // (kEmit: the messages, for the state export behind the loop)
    constexpr int rows[D] = {Es::row...};
    constexpr int shf[D] = {Es::shift...};
    constexpr int eix[D] = {Es::e...};
    constexpr int col1[D] = {Es::col1...};
    if constexpr (kEmit && train_traits<G>::on) {
        if (c.dump) dump_check<G, MODE, Es...>(c, raw);
    }
    float u[D];
#pragma unroll
    for (int k = 0; k < D; k++) {
        if constexpr (MODE == 2) {
            // QMS q=5.  (1) Every CN input is ALREADY on the 0.5 grid: it is a sum of the quantised channel value and quantised
            // messages (|sum| <= 7.5 * 24, exact in fp32), so the reference's re-quantisation (:386-389) is just the clamp to
            // +-7.5 — and the clamp is free: the min network below starts from 7.5 instead of 10000 (min over the others of
            // min(|x|, 7.5)), the sign comes from the raw word.  (2) Every non-zero input is >= 0.5 in magnitude, so the
            // reference's zero handling — 0 -> +1e-4 (:391-393), then mag - 1e-4 where mag <= 1e-4 (:416), i.e. 1e-4 - 1e-4 = 0 —
            // is "a zero counts as positive and as magnitude 0"; zeros are +0.0 here (the +0-started VN sums and quant5_grid
            // never produce -0.0), so neither step needs an instruction.  Conditioning cost per edge: none.
            u[k] = raw[k];
        } else {
            const float v = condition<MODE>(raw[k], c.lo, c.hi);
            u[k] = (v == 0.0f) ? 0.0001f : v;                            // x + 1e-4 * [x == 0] (:391-393)
        }
    }
    constexpr int H = (D + 1) / 2;
    float se[H + 1];
    constexpr float kCap = MODE == 2 ? 7.5f : 10000.0f;      // (:74-75 cap; QMS: the clamp of the re-quantisation, see above)
    se[H] = kCap;
#pragma unroll
    for (int q = H - 1; q >= 0; q--) {
        if (2 * q + 1 < D) se[q] = fmin3(fabsf(u[2 * q]), fabsf(u[2 * q + 1]), se[q + 1]);
        else se[q] = fminf(fabsf(u[2 * q]), se[q + 1]);
    }
    unsigned x = (D & 1) ? 0x80000000u : 0u;
#pragma unroll
    for (int k = 0; k < D; k++) x ^= __float_as_uint(u[k]);
    float pe = kCap;
    float pm[D];      // (NLDPC_QMS_PAIR: quantised magnitudes, written pairwise)
#pragma unroll
    for (int k = 0; k < D; k++) {
        const int q = k >> 1;
        float mag;
        if ((k & 1) == 0) {
            if (k + 1 < D) mag = fmin3(pe, fabsf(u[k + 1]), se[q + 1]);
            else mag = fminf(pe, se[q + 1]);
        } else {
            mag = fmin3(pe, fabsf(u[k - 1]), se[q + 1]);
            pe = fmin3(pe, fabsf(u[k - 1]), fabsf(u[k]));
        }
        if (col1[k] >= 0 && !kEmit) continue;
        float c2v;
        if constexpr (MODE == 2 && NLDPC_QMS_PAIR) {
            // (arithmetic as in the scalar branch below; two stored edges share the multiply and the two quantiser adds)
            constexpr float kMagic = 6291456.0f;
            if (qms_pair_first<kEmit, D>(eix, col1, k)) {
                // mag of edge k + 1 (the min network above, one step ahead)
                float mag1;
                if (((k + 1) & 1) == 0) {
                    if (k + 2 < D) mag1 = fmin3(pe, fabsf(u[k + 2]), se[((k + 1) >> 1) + 1]);
                    else mag1 = fminf(pe, se[((k + 1) >> 1) + 1]);
                } else {
                    mag1 = fmin3(pe, fabsf(u[k]), se[((k + 1) >> 1) + 1]);
                }
                const f2 pr = mul2(pack2(mag, mag1), pack2(wb[k].x, wb[k].y));
                const float a0 = fminf(fmaxf(lo(pr), 0.0f), 7.5f), a1 = fminf(fmaxf(hi(pr), 0.0f), 7.5f);
                const f2 qv = add2(add2(pack2(a0, a1), pack2(kMagic, kMagic)), pack2(-kMagic, -kMagic));
                pm[k] = lo(qv);
                pm[k + 1] = hi(qv);
            } else if (!(k > 0 && qms_pair_first<kEmit, D>(eix, col1, k - 1))) {
                const float m = fminf(fmaxf(mulf(mag, wb[k].x), 0.0f), 7.5f);
                pm[k] = addf(addf(m, kMagic), -kMagic);
            }
            const unsigned sb = (x ^ __float_as_uint(u[k])) & 0x80000000u;
            c2v = __uint_as_float(__float_as_uint(pm[k]) | sb);
            if constexpr (kEmit) c2v = (mag == 0.0f) ? 0.0f : c2v;
        } else if constexpr (MODE == 2) {
            const float wk = wb[k].x;
            // madj == mag >= 0 (see above); m >= 0 so only the upper clamp of the quantiser can bind
            float m = fmaxf(mulf(mag, wk), 0.0f);                             // |o| * W, ReLU (:431-505)
            constexpr float kMagic = 6291456.0f;
            m = addf(addf(fminf(m, 7.5f), kMagic), -kMagic);                  // quantise (:507-510)
            const unsigned sb = (x ^ __float_as_uint(u[k])) & 0x80000000u;    // negative iff #positive others is even (:417-423)
            c2v = __uint_as_float(__float_as_uint(m) | sb);
            // m * sign(o) with sign(0) = 0 (:512): mag == 0 gives +0.0, not -0.0.  A -0.0 message is absorbed by the
            // +0-started sums that consume it, so only iterations whose messages are exported (self.llr) pay for it.
            if constexpr (kEmit) c2v = (mag == 0.0f) ? 0.0f : c2v;
        } else {
            const float wk = wb[k].x;
            const float madj = (mag > 0.0001f) ? mag : addf(mag, -0.0001f);   // (:416)
            float m = fmaxf(mulf(fabsf(madj), wk), 0.0f);                     // |o| * W, ReLU (:431-505)
            m = condition<MODE>(m, c.lo, c.hi);                               // (:507-510)
            // sign(o) = sign(madj) * sgn;  sgn is negative iff the number of positive OTHER inputs is even (:417-423)
            const unsigned sb = (x ^ __float_as_uint(u[k]) ^ __float_as_uint(madj)) & 0x80000000u;
            c2v = __uint_as_float(__float_as_uint(m) | sb);
            c2v = (madj == 0.0f) ? 0.0f : c2v;                                // m * sign(0) (:512)
        }
        if constexpr (kEmit) {
            cv[k] = c2v;
        }
        if (col1[k] < 0) {
            c.rot[shf[k]][rows[k] * G::Z] = c2v;
        } else if constexpr (kEmit) {
            const int q = col1[k] * G::Z + (int)(c.rot[shf[k]] - (c.lane - c.z));
            const float xo = kXo == 2 ? xo_global<MODE>(c, q) : (kXo == 1 ? c.rot[shf[k]][c.xo_off + col1[k] * G::Z] : raw[k]);
            c.emit_rot(col1[k], shf[k], boosted_out(c, q, xo, addf(0.0f, c2v)));   // (:513-526)
        }
    }
    if constexpr (kEmit) {
        if (c.llr_sc) {
#pragma unroll
            for (int k = 0; k < D; k++) {
                float *const st = col1[k] >= 0 ? c.llr_d1 : c.llr_last;
                if (st) st[(size_t)(c.rot[shf[k]] - (c.lane - c.z)) * c.llr_pitch + eix[k]] = cv[k];   // self.llr[T][b][z][e]
            }
        }
    }
    if constexpr (kEmit && llr_inline<G>()) {
        if (c.llr_inl) {
            __syncwarp();
            if (c.llr_row) {
#pragma unroll
                for (int k = 0; k < D; k++) {
                    const float v = col1[k] >= 0 ? cv[k] : c.lane[rows[k] * G::Z];
                    if ((eix[k] & 3) == 3)
                        __stcs(reinterpret_cast<float4 *>(c.llr_row + eix[k] - 3), make_float4(c.llr_q[0], c.llr_q[1], c.llr_q[2], v));
                    else
                        c.llr_q[eix[k] & 3] = v;
                }
            }
        }
    }
}

#ifndef NLDPC_PIPE_CN
#define NLDPC_PIPE_CN 1     // 1: the next check's inputs are loaded before the current check computes
#endif
// ---- Boosted state export, vector form (self.llr[t + 1], BoostedNeuralLDPCDecoder.py:512) ---------------------------------
// The CN phase scatters each c2v to its VARIABLE lane, so a 4-byte store per lane and edge hits 16 different rows of the
// [B][Z][pitch] tensor: one 32-byte sector per 4 bytes, L2-request bound (0.7 TB/s).  After the phase (and its group sync) lane
// z owns row z of the tensor in the variable-lane view of the slab: it walks the edges in weight order (G::checks) and writes 4
// consecutive ones with ONE 16-byte store.  Needs 16-byte rows (llr_pitch % 4 == 0: WiMAX E = 88 as is, BG2's E = 197 padded to
// 200 by the caller).  Degree-1 edges with an identity circulant (BG2: 38 of 197) have no slab row — their message is computed
// by the lane that owns the tensor row, which keeps its 4-byte store inside the CN phase (NeuralLane::llr_d1); parking the 38
// values in registers until here made every BG2 kernel spill (measured).  They are HOLES of this pass: an aligned group of 4
// with holes leaves as 8- / 4-byte pieces (`holes` is a compile-time constant after inlining: edges and groups are immediates).
// Must run before the next VN phase (in place).
template <class G>
struct LlrExport {
    NeuralLane<G> &c;
    float *dst;              // &llr[..][b][z][0], 16-byte aligned
    float q[4];
    unsigned holes;
    __device__ __forceinline__ void first_deg1() {}
    __device__ __forceinline__ void flush(int e0, int n) {      // elements e0 .. e0 + n - 1 of the aligned group starting at e0
        if (n == 4 && holes == 0) {
            __stcs(reinterpret_cast<float4 *>(dst + e0), make_float4(q[0], q[1], q[2], q[3]));
        } else {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const bool a = 2 * h < n && !((holes >> (2 * h)) & 1), b = 2 * h + 1 < n && !((holes >> (2 * h + 1)) & 1);
                if (a && b) __stcs(reinterpret_cast<float2 *>(dst + e0 + 2 * h), make_float2(q[2 * h], q[2 * h + 1]));
                else if (a) __stcs(dst + e0 + 2 * h, q[2 * h]);
                else if (b) __stcs(dst + e0 + 2 * h + 1, q[2 * h + 1]);
            }
        }
        holes = 0;
    }
    template <class... Es>
    __device__ __forceinline__ void chk() {
        constexpr int D = sizeof...(Es);
        constexpr int rows[D] = {Es::row...};
        constexpr int eix[D] = {Es::e...};
        constexpr int col1[D] = {Es::col1...};
        static_assert(G::kDeg1Smem == 0, "degree-1 blocks are lane-private (identity circulants)");
#pragma unroll
        for (int k = 0; k < D; k++) {
            if (col1[k] >= 0) holes |= 1u << (eix[k] & 3);
            else q[eix[k] & 3] = c.lane[rows[k] * G::Z];
            if ((eix[k] & 3) == 3) flush(eix[k] - 3, 4);
        }
    }
    __device__ __forceinline__ void finish() {
        if ((G::E & 3) != 0) flush(G::E & ~3, G::E & 3);
    }
};
template <class G>
__device__ __forceinline__ void llr_export(NeuralLane<G> &c, float *cw_base) {      // cw_base = &llr[..][b][0][0]
    LlrExport<G> ex{c, cw_base + (size_t)c.z * c.llr_pitch, {0.0f, 0.0f, 0.0f, 0.0f}, 0u};
    G::checks(ex);
    ex.finish();
}

// ---- Train<> variant: the "extension" checks (D stored edges + ONE trailing degree-1 block with an identity circulant; 38 of
// BG2's 42) as loops over the runtime descriptors of G::loop_classes — same arithmetic as cn_check_boosted_core, table entries
// from constant memory instead of immediates.  word k < D: slab row | shift << 8 | edge << 16; word D: block J | x-row index << 8
// | edge << 16.  The block's marginal leaves from here (every iteration: this is list mode), its xa_origin is fetched from
// global memory one check ahead into a loop-carried register when VN weights make it differ from the on-chip xa_input.
template <class G, int MODE, int kXo, bool kLlr>      // kLlr: the launch exports self.llr (compile-time: predicated off, the
struct CnBoostedLoops {                              // export still issued 4 % of the kernel's instructions)
    NeuralLane<G> &c;
    int base;                 // first descriptor word of this graph in c_desc
    // operands of one looped check (D stored edges + the degree-1 edge)
    template <int D>
    struct Ops {
        float raw[D + 1], wk[D + 1];
        int moff[D];              // message slot of edge k, floats from the codeword's slab base
        int zl[D];                // its variable lane (self.llr is [lane][edge])
        uint32_t w1;              // descriptor of the degree-1 edge: block J | x-row index << 8 | edge << 16
        float xo_raw;             // xa_origin of block J (kXo == 2: re-read from global memory / L2)
    };
    template <int D>
    __device__ __forceinline__ void load(Ops<D> &o, int w0) const {
        constexpr int Z = G::Z;
        const float *slab0 = c.lane - c.z;
#pragma unroll
        for (int k = 0; k < D; k++) {
            const uint32_t w = c_desc[w0 + k];
            const int zz = rot_lane<G>(c.z, (w >> 8) & 0xff);
            o.moff[k] = (int)(w & 0xff) * Z + zz;
            o.zl[k] = zz;
            o.raw[k] = slab0[o.moff[k]];
            o.wk[k] = wb_at<true>(c.wb_base + (int)(w >> 16)).x;
        }
        o.w1 = c_desc[w0 + D];
        o.raw[D] = c.lane[(G::kXRows + G::S + ((o.w1 >> 8) & 0xff)) * Z];        // xa_input of block J, own lane
        o.wk[D] = wb_at<true>(c.wb_base + (int)(o.w1 >> 16)).x;
        if constexpr (kXo == 2) o.xo_raw = __ldg(c.xa_cw + (int)(o.w1 & 0xff) * Z + c.z);
        else o.xo_raw = 0.0f;
    }
    // the arithmetic of one check: c2v of its D + 1 edges (cn_check_boosted_core with run-time table entries)
    template <int D>
    __device__ __forceinline__ void math(const Ops<D> &o, float *c2v) const {
        constexpr int NE = D + 1;
        const float *raw = o.raw, *wk = o.wk;
        float u[NE];
#pragma unroll
        for (int k = 0; k < NE; k++) {
            if constexpr (MODE == 2) {
                u[k] = raw[k];             // already on the grid; the clamp to +-7.5 is the min network's cap (cn_check_boosted_core)
            } else {
                const float v = condition<MODE>(raw[k], c.lo, c.hi);
                u[k] = (v == 0.0f) ? 0.0001f : v;
            }
        }
        constexpr int H = (NE + 1) / 2;
        float se[H + 1];
        constexpr float kCap = MODE == 2 ? 7.5f : 10000.0f;
        se[H] = kCap;
#pragma unroll
        for (int q = H - 1; q >= 0; q--) {
            if (2 * q + 1 < NE) se[q] = fmin3(fabsf(u[2 * q]), fabsf(u[2 * q + 1]), se[q + 1]);
            else se[q] = fminf(fabsf(u[2 * q]), se[q + 1]);
        }
        unsigned x = (NE & 1) ? 0x80000000u : 0u;
#pragma unroll
        for (int k = 0; k < NE; k++) x ^= __float_as_uint(u[k]);
        float pe = kCap;
#pragma unroll
        for (int k = 0; k < NE; k++) {
            const int q = k >> 1;
            float mag;
            if ((k & 1) == 0) {
                if (k + 1 < NE) mag = fmin3(pe, fabsf(u[k + 1]), se[q + 1]);
                else mag = fminf(pe, se[q + 1]);
            } else {
                mag = fmin3(pe, fabsf(u[k - 1]), se[q + 1]);
                pe = fmin3(pe, fabsf(u[k - 1]), fabsf(u[k]));
            }
            if constexpr (MODE == 2) {
                float m = fmaxf(mulf(mag, wk[k]), 0.0f);
                constexpr float kMagic = 6291456.0f;
                m = addf(addf(fminf(m, 7.5f), kMagic), -kMagic);
                const unsigned sb = (x ^ __float_as_uint(u[k])) & 0x80000000u;
                const float v = __uint_as_float(__float_as_uint(m) | sb);
                c2v[k] = (mag == 0.0f) ? 0.0f : v;
            } else {
                const float madj = (mag > 0.0001f) ? mag : addf(mag, -0.0001f);
                float m = fmaxf(mulf(fabsf(madj), wk[k]), 0.0f);
                m = condition<MODE>(m, c.lo, c.hi);
                const unsigned sb = (x ^ __float_as_uint(u[k]) ^ __float_as_uint(madj)) & 0x80000000u;
                const float v = __uint_as_float(__float_as_uint(m) | sb);
                c2v[k] = (madj == 0.0f) ? 0.0f : v;
            }
        }
    }
    // scatter the messages back, export, and emit the marginal of block J (degree 1: out = xa_origin + (0 + c2v), :513-526;
    // un-rotated: this lane holds bit (J, z))
    template <int D>
    __device__ __forceinline__ void finish(const Ops<D> &o, const float *c2v, int w0, int rec_off) {
        constexpr int Z = G::Z, NE = D + 1;
        float *slab0 = c.lane - c.z;
        if (c.dump) dump_record<G, MODE, NE>(c, o.raw, rec_off);
#pragma unroll
        for (int k = 0; k < NE; k++) {
            if constexpr (kLlr) {
                const int zlane = k < D ? o.zl[k < D ? k : 0] : c.z;
                if (c.llr_last) c.llr_last[zlane * c.llr_pitch + (int)(c_desc[w0 + k] >> 16)] = c2v[k];      // self.llr[t + 1][b][z][e] (:512)
            }
            if (k < D) slab0[o.moff[k < D ? k : 0]] = c2v[k];
        }
        const int qbit = (int)(o.w1 & 0xff) * Z + c.z;
        float xo;
        if constexpr (kXo == 2) xo = (MODE == 2) ? quant5_grid(o.xo_raw) : o.xo_raw;      // (a zero's sign is absorbed by the sum)
        else xo = o.raw[D];
        const float v = boosted_out(c, qbit, xo, addf(0.0f, c2v[D]));
        if (c.soft) st_global_stream(c.soft + qbit, v);
    }
    // Checks of a class are processed in PAIRS: both checks' operands are loaded first (they touch different edges), then the
    // two independent arithmetic chains sit next to each other in one basic block — the per-check chain (descriptor ->
    // address -> load -> min network -> multiply -> quantise -> sign -> store) alone leaves a warp waiting on fixed
    // latencies (ncu: `wait` 1.13 cycles per issue at 2 warps per scheduler)
    template <int D, int GANG>
    __device__ __forceinline__ void gang(int w0, int rec_off, int rec) {
        constexpr int NE = D + 1;
        Ops<D> o[GANG];
        float cv[GANG][NE];
#pragma unroll
        for (int g = 0; g < GANG; g++) load<D>(o[g], w0 + g * NE);
#pragma unroll
        for (int g = 0; g < GANG; g++) math<D>(o[g], cv[g]);
#pragma unroll
        for (int g = 0; g < GANG; g++) finish<D>(o[g], cv[g], w0 + g * NE, rec_off + g * rec);
    }
    template <int D, int FIRST, int COUNT, int OFFH, int OFFF>
    __device__ __forceinline__ void cls() {
        constexpr int NE = D + 1;
        constexpr int REC = MODE == 2 ? G::dump_slots_h(NE) : NE, OFF = MODE == 2 ? OFFH : OFFF;
        constexpr int kGang = NLDPC_TRAIN_GANG;
        const int w00 = base + FIRST;
        int i = 0;
#pragma unroll 1
        for (; i + kGang <= COUNT; i += kGang) gang<D, kGang>(w00 + i * NE, OFF + i * REC, REC);
        if constexpr (kGang > 2) {
            if (i + 2 <= COUNT) {
                gang<D, 2>(w00 + i * NE, OFF + i * REC, REC);
                i += 2;
            }
        }
        if (i < COUNT) gang<D, 1>(w00 + i * NE, OFF + i * REC, REC);
    }
};

template <class G, bool kEmit, int MODE, int kXo>
struct CnBoosted {
    NeuralLane<G> &c;
    float raw[2][G::kMaxRowDeg];
    __device__ __forceinline__ void first_deg1() { stage_wait(c); }      // (generated: before the first check that emits)
    float2 wb[2][G::kMaxRowDeg];
    template <class... Es>
    __device__ __forceinline__ void chk() {
        cn_load<G, kEmit, true, true, MODE == 2, Es...>(c, raw[0], wb[0]);
        cn_check_boosted_core<G, kEmit, MODE, kXo, Es...>(c, raw[0], wb[0]);
    }
    template <int SLOT, class... Es>
    __device__ __forceinline__ void ld() {
        cn_load<G, kEmit, true, true, MODE == 2, Es...>(c, raw[SLOT], wb[SLOT]);
    }
    template <int SLOT, class... Es>
    __device__ __forceinline__ void chk() {
        cn_check_boosted_core<G, kEmit, MODE, kXo, Es...>(c, raw[SLOT], wb[SLOT]);
    }
};

template <class G, bool kEmit, bool kConstW, bool kZeroSafe>
struct CnNeural {
    NeuralLane<G> &c;
    float raw[2][G::kMaxRowDeg];
    __device__ __forceinline__ void first_deg1() { stage_wait(c); }      // (generated: before the first check that emits)
    float2 wb[2][G::kMaxRowDeg];
    template <class... Es>
    __device__ __forceinline__ void chk() {
        cn_load<G, kEmit, kConstW, false, false, Es...>(c, raw[0], wb[0]);
        cn_check_core<G, kEmit, kConstW, kZeroSafe, Es...>(c, raw[0], wb[0]);
    }
    template <int SLOT, class... Es>
    __device__ __forceinline__ void ld() {
        cn_load<G, kEmit, kConstW, false, false, Es...>(c, raw[SLOT], wb[SLOT]);
    }
    template <int SLOT, class... Es>
    __device__ __forceinline__ void chk() {
        cn_check_core<G, kEmit, kConstW, kZeroSafe, Es...>(c, raw[SLOT], wb[SLOT]);
    }
};

template <class G, class F>
__device__ __forceinline__ void run_checks(F &f) {
#if NLDPC_PIPE_CN
    G::checks_pipelined(f);
#else
    G::checks(f);
#endif
}

// -----------------------------------------------------------------------------------------------------------
constexpr int slab_floats(int base, int extra, int z) {      // slab stride == Z (mod 32), multiple of 4 floats
    int s = base + extra;
    while ((s % 32) != (z % 32) || (s % 4)) s++;
    return s;
}

// see train_traits: all channel LLRs in shared rows (kXRegs extra rows behind the message rows), extension checks as loops
template <class G0>
struct Train : G0 {
    static constexpr bool kTrainVariant = true;
    static constexpr int kSlab = slab_floats(G0::kSlab, G0::kXRegs * G0::Z, G0::Z);
};

#ifndef NLDPC_PREFETCH
// 1: throughput-mode kernels bulk-load the NEXT unit's channel LLRs into a landing buffer during the current unit's decode.
// Measured (round 2) and left off: 47.19 vs 47.43 M cw/s on BG2, 66.3 vs 69.8 on WiMAX — the other warp of the scheduler
// already covers the ~1 us a group waits for its bulk load, and the extra group barrier + 53 KB of shared memory cost more.
#define NLDPC_PREFETCH 0
#endif
#ifndef NLDPC_STAGE_OUT
#define NLDPC_STAGE_OUT 1   // 1: list-mode soft outputs leave through shared staging rows + bulk TMA stores (see Staged<>)
#endif

// kXoRows: N extra rows per codeword for xa_origin; kStage: N extra rows staging the soft output row (list mode);
// kPf: a landing buffer of N*Z floats per codeword OUTSIDE the slab, into which the NEXT work unit's channel LLRs are
// bulk-loaded while the current unit decodes (throughput mode, where the shared memory left over allows it)
template <class G, bool kXoRows = false, bool kStage = false, bool kPf = false>
struct SpecCfg {
    using Shape = GroupShape<G::Z>;
    static constexpr int kNZ = G::N * G::Z;
    static constexpr int kSlabF = (kXoRows || kStage) ? slab_floats(G::kSlab, (kXoRows ? kNZ : 0) + (kStage ? kNZ : 0), G::Z) : G::kSlab;
    static constexpr int kXoOff = kXoRows ? (G::kXRows + G::S) * G::Z : 0;   // N xo rows follow the message rows
    static constexpr int kStageOff = (G::kXRows + G::S) * G::Z + (kXoRows ? kNZ : 0);   // then the N staging rows
    static_assert(!kStage || ((kStageOff * 4) % 16 == 0 && (kSlabF * 4) % 16 == 0 && (kNZ * 4) % 16 == 0), "bulk-store source alignment");
    static constexpr int kHardBytes = (G::N * G::Z + 7) / 8;
    static constexpr int kHardStride = (kHardBytes + 15) & ~15;           // per-codeword staging, 16 B multiple
    static constexpr bool kPrefetch = kPf;
    static constexpr int kPerCw = kSlabF * 4 + kHardStride + (kPf ? kNZ * 4 : 0);      // shared bytes per codeword
    // CTA shape.  Warp k of a CTA runs on SM sub-partition k mod 4 and the kernel is issue-bound per sub-partition, so what
    // counts is how evenly the resident warps spread over the four of them, not how many there are (measured on BG2:
    // 2 CTAs x 5 warps put 4/2/2/2 warps on the sub-partitions and ran 20 % SLOWER than 2 x 4 warps although 25 % more
    // codewords were resident).  Score a configuration (c CTAs per SM, g groups per CTA) as
    //     resident warps * u(m) / m,   m = warps on the busiest sub-partition,  u = issue utilisation at m warps
    // and take the best one that fits shared memory.  More than 2 warps per sub-partition do NOT pay off here: every warp
    // streams through ~60 KB of unrolled code and more resident warps mean more instruction-cache pressure (BG2: 1 CTA x
    // 10 warps 36.3 M cw/s vs 8 warps 43.2 M; WiMAX: 15 warps 57.2 M vs 12 warps 59.2 M vs 2 x 6 warps 55.4 M).
    // what fits: every CTA needs its codewords, one mbarrier per group and 16 B; the SM has 228 KB, of which each resident
    // CTA also takes 1 KB for the system, and one CTA may ask for at most 227 KB
    static constexpr size_t cta_bytes(int g) { return (size_t)g * Shape::kCw * kPerCw + (size_t)g * 8 + 16; }
    static constexpr int score(int ctas, int g) {
        if (g <= 0 || g * Shape::kLanes > 512) return -1;
        if (cta_bytes(g) > (size_t)kSmemBudget || ctas * (cta_bytes(g) + 1024) > (size_t)228 * 1024) return -1;
        const int w = g * Shape::kWarps;
        const int m = ctas * ((w + 3) / 4);
        const int u = m <= 1 ? 35 : (m == 2 ? 59 : (m == 3 ? 50 : 45));      // percent, from the measurements below
        return ctas * w * u * 12 / m;
    }
    static constexpr int best(bool want_groups) {
        int bs = -1, bc = 1, bg = 1;
        for (int c = 1; c <= 2; c++)
            for (int g = 1; g <= 16; g++)
                if (score(c, g) > bs) { bs = score(c, g); bc = c; bg = g; }
        return want_groups ? bg : bc;
    }
#ifdef NLDPC_FORCE_GROUPS       // experiments only (tools/build_variant.sh)
    static constexpr int kCtasPerSm = Shape::kLanes == 32 ? 1 : best(false);
    static constexpr int kGroups = Shape::kLanes == 32 ? NLDPC_FORCE_GROUPS : best(true);
#else
    static constexpr int kCtasPerSm = best(false);
    static constexpr int kGroups = best(true);
#endif
    static constexpr int kThreads = kGroups * Shape::kLanes;
    static constexpr int kCwPerCta = kGroups * Shape::kCw;
    static constexpr size_t kSmemBytes = cta_bytes(kGroups);
    static_assert(G::S >= G::N, "the raw codeword is staged in the message rows");
};

// The configuration a kernel variant runs with (kernel and launchers must agree): xa_origin rows only in list mode with VN
// weights (see xo_global); output staging in list mode whenever it does not cost resident codewords.
// The training variant (kTrain: every-iteration outputs + training dump [+ fused loss], Boosted only) runs on Train<G>: no xo
// rows (xa_origin is re-read from global memory where VN weights make it differ) and no output staging, so that its larger
// slab still leaves the CTA all its codewords.
template <class G0, bool kEvery, bool kXo, bool kTrain = false>
struct KernelCfg {
    using G = std::conditional_t<kTrain, Train<G0>, G0>;
    static constexpr bool kXoRows = kXo && kEvery && !kTrain;
    using Plain = SpecCfg<G, kXoRows, false>;
    using Stage = SpecCfg<G, kXoRows, true>;
    using Pf = SpecCfg<G, kXoRows, false, true>;
    static constexpr bool kStage = NLDPC_STAGE_OUT && kEvery && !kTrain &&
                                   (Stage::kCtasPerSm * Stage::kCwPerCta >= Plain::kCtasPerSm * Plain::kCwPerCta);
    // throughput mode: prefetch the next unit's channel LLRs when the landing buffer costs no resident codewords and leaves
    // the CTA shape alone (BG2: 179.0 + 53.2 KB of the 227 KB)
    static constexpr bool kPrefetch = NLDPC_PREFETCH && !kEvery && !kTrain && Pf::kCtasPerSm == Plain::kCtasPerSm &&
                                      Pf::kGroups == Plain::kGroups;
    using type = std::conditional_t<kStage, Stage, std::conditional_t<kPrefetch, Pf, Plain>>;
};
// kEvery: outputs are produced after every iteration (drop-in list mode / per-iteration hard decisions);
// otherwise only after the last one (throughput mode) and the loop body carries no output code at all.
template <class G0, bool kEvery, bool kConstW, int MODE = 0, bool kXo = false, bool kTrain = false, bool kNoHard = false>
__global__ void __launch_bounds__(KernelCfg<G0, kEvery, kXo, kTrain && MODE != 0>::type::kThreads, KernelCfg<G0, kEvery, kXo, kTrain && MODE != 0>::type::kCtasPerSm)
nldpc_spec_neural_kernel(const DecodeArgs a) {
    // kTrain: Boosted -> the Train<> variant (all channel LLRs in shared rows, extension checks as loops, dump + fused loss);
    //         Neural  -> the list-mode kernel plus the training dump (Dumping<>)
    constexpr bool kTrainWrap = kTrain && MODE != 0;
    using Cfg = typename KernelCfg<G0, kEvery, kXo, kTrainWrap>::type;
    constexpr bool kStage = KernelCfg<G0, kEvery, kXo, kTrainWrap>::kStage;
    static_assert(!kNoHard || (kEvery && MODE == 0), "the soft-only variant is a Neural every-iteration kernel");
    using G1 = std::conditional_t<kTrainWrap, Train<G0>,
                                  std::conditional_t<kTrain, Dumping<SoftOnly<G0>>, std::conditional_t<kNoHard, SoftOnly<G0>, G0>>>;
    using G = std::conditional_t<kStage, Staged<G1, Cfg::kStageOff>, G1>;      // (same graph program; emit() writes staging rows)
    constexpr int kXoMode = !kXo ? 0 : ((kEvery && !kTrain) ? 1 : 2);
    static_assert(!kTrain || kEvery, "the training variants are every-iteration kernels");
    constexpr bool kDumps = kTrain;
    static_assert(MODE == 0 || kConstW, "the Boosted variants read their weights from the constant arena");
    static_assert(MODE != 0 || !kXo, "xo rows only exist for the Boosted decoder with VN weights");
    using Shape = typename Cfg::Shape;
    constexpr int Z = G::Z, NZ = G::N * G::Z;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *slabs = reinterpret_cast<float *>(smem_raw);
    uint8_t *hstage = smem_raw + (size_t)Cfg::kCwPerCta * Cfg::kSlabF * 4;
    uint64_t *bars = reinterpret_cast<uint64_t *>(hstage + (size_t)Cfg::kCwPerCta * Cfg::kHardStride);
    constexpr bool kPrefetch = Cfg::kPrefetch;
    // landing buffer of the prefetch, [codeword in CTA][N*Z] (behind the barriers, 16-byte aligned: 8 * kGroups + 16 is)
    float *pfbuf = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(bars) + ((size_t)Cfg::kGroups * 8 + 15) / 16 * 16);
    int pf_unit = -1;           // the unit whose channel LLRs are in (or on their way into) the landing buffer

    // group within the CTA, through a shuffle from lane 0: the value is the same, but the compiler now KNOWS it is warp-uniform,
    // so everything derived from it (work-unit loop, iteration counter, weight offsets in the constant arena) lives on the
    // uniform datapath: the per-edge {w, b} become LDCU loads into uniform registers consumed directly by FMUL / FADD instead
    // of indexed LDC loads into vector registers whose latency the multiply waited for (ncu: short_scoreboard on FMUL).
    const int grp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) / Shape::kWarps;
    const int gl = threadIdx.x - grp * Shape::kLanes;       // lane within the group
    int cwl, z;                                             // codeword within the group, lane within the codeword
    Shape::map(gl, cwl, z);
    const int cw_in_cta = grp * Shape::kCw + cwl;
    float *slab = slabs + (size_t)cw_in_cta * Cfg::kSlabF;
    uint64_t *bar = bars + grp;

    if (gl == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    group_sync<Shape::kLanes>(grp);

    NeuralLane<G> c;
    c.lane = slab + z;
    c.z = z;
    c.grp = grp;
    c.gl0 = gl == 0;
    c.xo_off = Cfg::kXoOff;
    c.lo = a.llr_lo;
    c.hi = a.llr_hi;
    c.llr_last = nullptr;
    c.llr_d1 = nullptr;
    c.llr_row = nullptr;
    c.llr_inl = false;
    c.llr_sc = false;
    c.llr_pitch = a.llr_pitch;
    c.mask = nullptr;
    c.dump = nullptr;
    c.yb = nullptr;
    c.cg = 0.0f;
    c.accA = c.accL = 0.0f;
#pragma unroll
    for (int s = 0; s < Z; s++) c.rot[s] = slab + ((z + s) % Z);
    uint8_t *hb_mine = hstage + (size_t)cw_in_cta * Cfg::kHardStride;
    // check-packed training dump: bytes per codeword and iteration
    constexpr size_t kDumpCw = (MODE == 2) ? (size_t)G::kDumpH * Z * 2 : (size_t)G::kDumpF * Z * 4;
    const bool fused = kTrainWrap && a.ybits != nullptr;      // fused BCE: `soft` receives dL/dout (see bce_fused)
    static_assert(Cfg::kHardBytes % 4 == 0 || MODE == 0, "packed labels are staged as 32-bit words");

    const int n_units = (a.B + Shape::kCw - 1) / Shape::kCw;
    const int unit_stride = gridDim.x * Cfg::kGroups;
    const int unit_end_all = a.unit_end > 0 ? min(a.unit_end, n_units) : n_units;      // this launch's units are [unit_begin, unit_end_all)
    uint32_t phase = 0;
    const bool soft_all = a.soft_mode == 1, hard_all = a.hard_mode == 1;
    const bool soft_any = a.soft_mode != 0, hard_any = a.hard_mode != 0;
    // Boosted decode variants (not the training variant, which exports from its loops): self.llr leaves through LlrExport when the
    // caller's rows are 16-byte aligned
    // (graphs with lane-private degree-1 blocks and one-warp groups export inline instead, llr_inline<G>: BG2)
    constexpr bool kLlrInl = MODE != 0 && !kTrainWrap && llr_inline<G>();
    constexpr bool kLlrVec = NLDPC_LLR_VEC && MODE != 0 && !kTrainWrap && G::kDeg1Smem == 0 && !kLlrInl;
    const bool llr_vec_ok = (kLlrVec || kLlrInl) && (a.llr_pitch & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.llr_all) | reinterpret_cast<uintptr_t>(a.llr_last)) & 15) == 0;

    // phase barrier: group-local (a CTA-wide lockstep variant, so that all warps stream the same code, bought nothing:
    // 42.4 vs 43.2 M cw/s in round 1)
    auto phase_sync = [&]() __attribute__((always_inline)) { group_sync<Shape::kLanes>(grp); };
    // The decode of one work unit.  kSafe = false is the fast path, which assumes that no CN input of the group is exactly
    // zero: the channel LLRs are screened once, the VN phase tracks min |v2c| (half an instruction per edge), and when a zero
    // shows up the unit is abandoned (return false) and decoded again from its channel LLRs with kSafe = true, whose CN phase
    // applies the reference's "exact zero -> magnitude 10000, not positive" rule (:74, :78) at two more instructions per edge.
    // Exact zeros need punctured / quantised inputs or an exact fp32 cancellation (a few dozen units per 65536 codewords at
    // 2 dB), so the second attempt is noise; what matters is that the hot loop contains NO call and NO second code path: with
    // the group index known to be warp-uniform (see above) its iteration counter and weight offsets then stay on the uniform
    // datapath.  Outputs an abandoned attempt has already written are rewritten with identical values.
    auto run_unit = [&](const int unit, auto safe_tag) __attribute__((always_inline)) -> bool {
        const int b0 = unit * Shape::kCw;
        const int b = b0 + cwl;
        c.valid = b < a.B;
        c.xa_cw = a.xa + (size_t)min(b, a.B - 1) * NZ;      // (padding lanes re-read the last codeword; their outputs are dropped)
        const int ncw = max(0, min(Shape::kCw, a.B - b0));
        constexpr bool kSafe = decltype(safe_tag)::value;
        // ---- bulk-TMA the group's channel LLRs (one 1-D copy per codeword) ----
        // destination: the message rows of the slabs (staging, not yet live) or, with prefetching, the landing buffer
        auto issue_load = [&](int u) __attribute__((always_inline)) {
            const int ub0 = u * Shape::kCw;
            const int n = max(0, min(Shape::kCw, a.B - ub0));
            if (gl == 0 && n > 0) {
                fence_proxy_async();
                mbar_arrive_expect_tx(bar, (uint32_t)(n * NZ * sizeof(float)));
                for (int q = 0; q < n; q++) {
                    float *dst = kPrefetch ? pfbuf + (size_t)(grp * Shape::kCw + q) * NZ
                                           : slabs + (size_t)(grp * Shape::kCw + q) * Cfg::kSlabF + G::kXRows * Z;
                    tma_load_1d(dst, a.xa + (size_t)(ub0 + q) * NZ, (uint32_t)(NZ * sizeof(float)), bar);
                }
            }
        };
        if (!kPrefetch || pf_unit != unit) issue_load(unit);
        if constexpr (Z != 16 && Z != 32 && (G::kXRegs + G::kDeg1Smem > 0)) {   // atomicOr staging (rotated emission) must start from zero
            if (hard_any) for (int q = z; q < Cfg::kHardStride / 4; q += Z) reinterpret_cast<unsigned *>(hb_mine)[q] = 0u;
        }
        if (ncw > 0) {
            mbar_wait(bar, phase);
            phase ^= 1;
        }

        // the raw codeword sits in the message rows (staging); every lane moves its own elements to their places
        // (shared rows / registers).  Lane z only ever touches element z of a row in the VN phase, so no barrier is needed
        // before iteration 0 overwrites the staging area.
        {
            PlaceXa<G, MODE, kXoMode> pl{c, kPrefetch ? pfbuf + (size_t)cw_in_cta * NZ + z : c.lane + G::kXRows * Z, 1.0f};
            G::blocks(pl);
            if constexpr (kPrefetch) {
                // the landing buffer has been read by every lane of the group: start the NEXT unit's load now, it has this
                // unit's whole decode (~50 us) to arrive (pass 2 visits scattered units: no prefetch there)
                group_sync<Shape::kLanes>(grp);
                const int next = unit + unit_stride;
                if (!kSafe && next < unit_end_all) {
                    issue_load(next);
                    pf_unit = next;
                } else {
                    pf_unit = -1;
                }
            }
            if constexpr (MODE == 0 && !kSafe) {
                if (group_any<Shape::kLanes>(grp, pl.zm == 0.0f)) return false;
            }
        }
        c.zmin = 10000.0f;
        // per-iteration pieces shared by both output modes
        auto xin_update = [&](int t) __attribute__((always_inline)) {
            if constexpr (kXo) {
                ScaleXin<G, MODE> sc{c, a.vn_w + (size_t)t * G::N};
                G::blocks(sc);
            }
        };
        // inline state export around an emitting CN phase (llr_inline<G>): c.llr_last / c.llr_sc have just been set for the phase
        auto llr_inline_begin = [&]() __attribute__((always_inline)) {
            if constexpr (kLlrInl) {
                c.llr_inl = llr_vec_ok && c.llr_sc;
                if (c.llr_inl) {
                    c.llr_row = c.llr_last ? c.llr_last + (size_t)c.z * a.llr_pitch : nullptr;
                    c.llr_last = nullptr;
                    c.llr_d1 = nullptr;
                    c.llr_sc = false;
                }
            }
        };
        auto llr_inline_end = [&]() __attribute__((always_inline)) {
            if constexpr (kLlrInl) {
                if (c.llr_row) {      // the open group: E % 4 trailing edges
#pragma unroll
                    for (int i = 0; i < (G::E & 3); i++) __stcs(c.llr_row + (G::E & ~3) + i, c.llr_q[i]);
                }
                c.llr_row = nullptr;
                c.llr_inl = false;
            }
        };
        auto cn_run = [&](auto emit_tag) __attribute__((always_inline)) {
            constexpr bool kEmitNow = decltype(emit_tag)::value;
            if constexpr (MODE == 0) {
                CnNeural<G, kEmitNow, kConstW, kSafe> f{c};
                run_checks<G>(f);
            } else if constexpr (kTrainWrap) {
                // the checks that differ structurally stay unrolled, the extension checks run as descriptor loops
                CnBoosted<G, true, MODE, kXoMode> f{c};
                G::checks_pipelined_rest(f);
                if constexpr (G::kLoopChecks > 0) {
                    if (a.llr_all != nullptr || a.llr_last != nullptr) {
                        CnBoostedLoops<G, MODE, kXoMode, true> l{c, a.desc_base};
                        G::loop_classes(l);
                    } else {
                        CnBoostedLoops<G, MODE, kXoMode, false> l{c, a.desc_base};
                        G::loop_classes(l);
                    }
                }
            } else {
                CnBoosted<G, kEmitNow, MODE, kXoMode> f{c};
                run_checks<G>(f);
            }
        };
        float *soft_cw = (soft_any && c.valid) ? a.soft + (size_t)b * NZ : nullptr;     // + t*B*NZ in ALL mode
        float lacc = 0.0f;                      // fused loss: sum_t c_t * (terms of this lane), this unit
        // fold the loss terms the phase that just ended has produced (they belong to the output of iteration t_out)
        auto loss_fold = [&](int t_out) __attribute__((always_inline)) {
            if (fused) {
                lacc += __ldg(a.coef + t_out) * (c.accA + 0.6931471805599453f * c.accL);
                c.accA = c.accL = 0.0f;
            }
        };
        if (fused) {
            // this codeword's packed labels -> the (unused: training asks for no hard decisions) hard-decision staging bytes
            const uint32_t *src = reinterpret_cast<const uint32_t *>(a.ybits + (size_t)min(b, a.B - 1) * Cfg::kHardBytes);
            for (int q = z; q < Cfg::kHardBytes / 4; q += Z) reinterpret_cast<uint32_t *>(hb_mine)[q] = __ldg(src + q);
            c.yb = reinterpret_cast<const uint32_t *>(hb_mine);        // (first read: after the phase barrier of iteration 0)
        }
        const size_t soft_iter = (size_t)a.B * NZ;
        uint8_t *hb_cw = hard_any ? hb_mine : nullptr;
        auto flush_hard = [&](int t_out) __attribute__((always_inline)) {
            // group-cooperative copy of the staged packed decisions to global memory (16 B per lane-step)
            group_sync<Shape::kLanes>(grp);
            uint8_t *dst = a.hard + ((hard_all ? (size_t)t_out * a.B : 0) + b0) * Cfg::kHardBytes;
            const uint8_t *src = hstage + (size_t)(grp * Shape::kCw) * Cfg::kHardStride;
            if constexpr (Cfg::kHardBytes % 4 == 0) {
                constexpr int W = Cfg::kHardBytes / 4;
                for (int i = gl; i < ncw * W; i += Shape::kLanes) {
                    const int q = i / W, r = i - q * W;
                    reinterpret_cast<uint32_t *>(dst + (size_t)q * Cfg::kHardBytes)[r] =
                        reinterpret_cast<const uint32_t *>(src + (size_t)q * Cfg::kHardStride)[r];
                }
            } else {
                for (int i = gl; i < ncw * Cfg::kHardBytes; i += Shape::kLanes) {
                    const int q = i / Cfg::kHardBytes, r = i - q * Cfg::kHardBytes;
                    dst[(size_t)q * Cfg::kHardBytes + r] = src[(size_t)q * Cfg::kHardStride + r];
                }
            }
            if constexpr (Z != 16 && Z != 32 && (G::kXRegs + G::kDeg1Smem > 0)) {
                group_sync<Shape::kLanes>(grp);
                for (int q = z; q < Cfg::kHardStride / 4; q += Z) reinterpret_cast<unsigned *>(hb_mine)[q] = 0u;
            }
            group_sync<Shape::kLanes>(grp);
        };

        // list mode with output staging: the finished [N*Z] row of every codeword of the group -> soft[t_out][b], one bulk
        // TMA store each (the rows sit behind the codewords' slabs, see Staged<>)
        auto flush_soft = [&](int t_out) __attribute__((always_inline)) {
            if constexpr (kStage) {
                if (soft_any && (soft_all || t_out == a.T - 1)) {
                    fence_proxy_async();                     // this lane's staging writes -> visible to the async proxy
                    group_sync<Shape::kLanes>(grp);
                    if (gl == 0) {
                        float *dst = a.soft + ((soft_all ? (size_t)t_out * a.B : 0) + b0) * NZ;
                        for (int q = 0; q < ncw; q++)
                            tma_store_1d(dst + (size_t)q * NZ, slabs + (size_t)(grp * Shape::kCw + q) * Cfg::kSlabF + Cfg::kStageOff,
                                         (uint32_t)(NZ * sizeof(float)));
                        tma_store_commit();
                    }
                }
            }
        };

        if constexpr (kEvery) {
            for (int t = 0; t < a.T; t++) {
                c.wt = a.w + (size_t)t * G::E;
                c.bt = a.b + (size_t)t * G::E;
                c.wb_base = a.wb_off + t * (MODE != 0 ? kWPitch<G> : G::E);
                // training dump for the backward sweep: check-packed records written by the CN phase (hist_fmt 1; launches that
                // want the slot-major format of the table-driven sweep run on the table-driven forward, see the launchers)
                const bool dump = kDumps && a.hist_v2c != nullptr && c.valid;
                xin_update(t);
                // channel-input state after this iteration's update: what the VN-weight chain of the sweep reads (rows 1..T-1)
                if constexpr (kTrainWrap && kXo) {
                    if (dump && t + 1 < a.T) {
                        DumpXin<G> d{c, a.hist_xin + ((size_t)(t + 1) * a.B + b) * NZ + z};
                        G::blocks(d);
                    }
                }
                if (t == 0) {
                    VnFirst<G> f{c};
                    G::vcols(f);
                } else {
                    c.soft = (soft_all && soft_cw) ? soft_cw + (size_t)(t - 1) * soft_iter : nullptr;
                    c.hb = hard_all ? hb_cw : nullptr;
                    if constexpr (kTrainWrap) {
                        c.mask = (dump && a.hist_mask) ? a.hist_mask + ((size_t)(t - 1) * a.B + b) * NZ : nullptr;
                        if (fused) c.cg = __ldg(a.coef + (t - 1)) * a.ginv;
                    }
                    stage_wait(c);       // (graphs without degree-1 blocks emit only here; a no-op wait otherwise)
                    VnStep<G, true, MODE, kXoMode> f{c};
                    run_vcols<G>(f);
                    loss_fold(t - 1);
                    flush_soft(t - 1);
                    if (hard_all) flush_hard(t - 1);
                }
                phase_sync();
                const bool last = t == a.T - 1;
                c.soft = (soft_cw && (soft_all || last)) ? soft_cw + (soft_all ? (size_t)t * soft_iter : 0) : nullptr;
                c.hb = (hard_all || last) ? hb_cw : nullptr;
                if constexpr (kDumps) c.dump = dump ? reinterpret_cast<char *>(a.hist_v2c) + ((size_t)t * a.B + b) * kDumpCw : nullptr;
                if constexpr (kTrainWrap) {
                    c.mask = (dump && a.hist_mask) ? a.hist_mask + ((size_t)t * a.B + b) * NZ : nullptr;
                    if (fused) c.cg = __ldg(a.coef + t) * a.ginv;
                }
                c.llr_last = !c.valid ? nullptr
                             : (a.llr_all ? a.llr_all + ((size_t)t * a.B + b) * Z * a.llr_pitch
                                          : ((last && a.llr_last) ? a.llr_last + (size_t)b * Z * a.llr_pitch : nullptr));
                c.llr_d1 = c.llr_last;
                c.llr_sc = a.llr_all != nullptr || (last && a.llr_last != nullptr);
                float *llr_vec = nullptr;      // vector export after the phase instead of the scalar stores inside it (LlrExport)
                if constexpr (kLlrVec) {
                    if (llr_vec_ok) {
                        llr_vec = c.llr_last;
                        c.llr_last = nullptr;
                        c.llr_sc = c.llr_sc && G::kXRegs > 0;      // (degree-1 edges, if any, keep their scalar store)
                    }
                }
                llr_inline_begin();      // ... or check by check inside the phase (llr_inline<G>)
                cn_run(std::true_type{});
                llr_inline_end();
                loss_fold(t);
                phase_sync();
                if constexpr (kLlrVec) {
                    if (llr_vec) llr_export<G>(c, llr_vec);
                }
            }
        } else {
            for (int t = 0; t < a.T; t++) {
                c.wt = a.w + (size_t)t * G::E;
                c.bt = a.b + (size_t)t * G::E;
                c.wb_base = a.wb_off + t * (MODE != 0 ? kWPitch<G> : G::E);
                xin_update(t);
                if (t == 0) {
                    VnFirst<G> f{c};
                    G::vcols(f);
                } else {
                    VnStep<G, false, MODE, kXoMode> f{c};
                    run_vcols<G>(f);
                }
                phase_sync();
                if (t < a.T - 1) {
                    cn_run(std::false_type{});
                } else {
                    c.soft = soft_cw;
                    c.hb = hb_cw;
                    c.llr_last = (a.llr_last && c.valid) ? a.llr_last + (size_t)b * Z * a.llr_pitch : nullptr;
                    c.llr_d1 = c.llr_last;
                    c.llr_sc = a.llr_last != nullptr;
                    float *llr_vec = nullptr;
                    if constexpr (kLlrVec && NLDPC_LLR_VEC_LAST) {
                        if (llr_vec_ok) {
                            llr_vec = c.llr_last;
                            c.llr_last = nullptr;
                            c.llr_sc = c.llr_sc && G::kXRegs > 0;
                        }
                    }
                    llr_inline_begin();
                    cn_run(std::true_type{});
                    llr_inline_end();
                    if constexpr (kLlrVec && NLDPC_LLR_VEC_LAST) {
                        if (llr_vec_ok && a.llr_last) {      // (launch-uniform: the barrier is taken by every lane of the group)
                            phase_sync();
                            if (llr_vec) llr_export<G>(c, llr_vec);
                        }
                    }
                }
                phase_sync();
            }
        }
        // Neural fast path: did any CN phase of this unit see an exact zero?  (Then what it computed from there on is not the
        // reference's result: abandon the unit; pass 2 decodes it again with the zero-safe CN phase and rewrites its outputs.)
        if constexpr (MODE == 0 && !kSafe) {
            if (group_any<Shape::kLanes>(grp, c.zmin == 0.0f)) return false;
        }
        // marginal of the last iteration for the blocks of degree >= 2
        {
            c.soft = soft_cw ? soft_cw + (soft_all ? (size_t)(a.T - 1) * soft_iter : 0) : nullptr;
            c.hb = hb_cw;
            if constexpr (kTrainWrap) c.mask = (a.hist_v2c && a.hist_mask && c.valid) ? a.hist_mask + ((size_t)(a.T - 1) * a.B + b) * NZ : nullptr;
            c.dump = nullptr;
            if constexpr (kEvery) stage_wait(c);
            Marginal<G, MODE, kXoMode> f{c};
            G::vcols(f);
            if constexpr (kEvery) flush_soft(a.T - 1);
            if (hard_any) flush_hard(a.T - 1);
        }
        if constexpr (kTrainWrap) {
            if (fused) {      // (cg of the last iteration is still set from its CN phase)
                loss_fold(a.T - 1);
                float v = c.valid ? lacc : 0.0f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) == 0) atomicAdd(a.loss_acc, (double)v);
                c.yb = nullptr;
            }
        }
        return true;
    };      // run_unit

    // Units are dealt out CTA-major: group g of CTA b takes units b + gridDim.x * (g + kGroups * k).  A batch too small to fill
    // the machine then spreads over all SMs with fewer busy warps each (the host launches min(n_units, resident CTAs) CTAs),
    // which is what a latency-bound small decode wants; for large batches the order of units is irrelevant.
    // Pass 1 runs every unit on the fast path and notes the abandoned ones in a 64-bit mask (the host never gives a group more
    // than 64 units per launch, see spec_units_per_launch); pass 2 — cold code BEHIND the hot loop, not inside it — decodes those.
    const int unit_first = a.unit_begin + (int)blockIdx.x + grp * (int)gridDim.x;
    const int unit_end = unit_end_all;
    unsigned long long failed = 0ull;
    {
        int k = 0;
        for (int unit = unit_first; unit < unit_end; unit += unit_stride, k++) {
            if (!run_unit(unit, std::false_type{})) failed |= 1ull << k;
            group_sync<Shape::kLanes>(grp);   // all lanes done with the slabs before the next TMA overwrites them
        }
    }
    if constexpr (MODE == 0) {
        if (failed != 0ull) {
            int k = 0;
            for (int unit = unit_first; unit < unit_end; unit += unit_stride, k++) {
                if (!((failed >> k) & 1ull)) continue;
                if constexpr (kStage) {
                    if (gl == 0) tma_store_wait_all();       // the abandoned attempt's bulk stores land before they are rewritten
                    group_sync<Shape::kLanes>(grp);
                }
#ifdef NLDPC_DEBUG_COUNT
                if (gl == 0) atomicAdd(&g_dbg_restarts, 1u);
#endif
                run_unit(unit, std::true_type{});
                group_sync<Shape::kLanes>(grp);
            }
        }
    }
    if constexpr (kStage) {
        if (gl == 0) tma_store_wait_all();      // the staging rows must outlive the bulk stores that read them
    }
}

}  // namespace nldpc
