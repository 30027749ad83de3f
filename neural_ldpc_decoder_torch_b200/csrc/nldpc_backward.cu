// nldpc_backward.cu — weight gradients of the decode loop: closed form of autograd through
// NeuralLDPCDecoder.forward (NeuralLDPCDecoder.py:54-98) and the MS / QMS branches of BoostedNeuralLDPCDecoder.forward
// (BoostedNeuralLDPCDecoder.py:320-531), SURVEY.md Appendix B.  Table-driven (any base graph that fits on chip).
//
// Two launches:
//   (A) the forward kernel re-run in "training dump" mode: it writes, per iteration, the v2c of every stored edge
//       (what the CN phase read), and for the Boosted decoder the channel-input state and the output-clamp mask, to an
//       HBM workspace.  HBM is >98 % idle in the forward, so spilling 10 KB per codeword-iteration is cheap, while a
//       c2v history on chip (126 KB per BG2 codeword at T=10) would leave one codeword per SM.
//   (B) this kernel: walks the iterations backwards with the gradient messages in shared memory (same slab layout and
//       thread mapping as the forward), recomputing each check's min / argmin / signs from the dumped v2c.
//
//   dc2v[e=(i,j)] = G_t[j] + sum_{e2 in col(j), e2 != e} dv2c_{t+1}[e2]                       (VN phase, backwards)
//   coef = dc2v * sign(o) * [pre > 0]  (and the quantiser / clamp straight-through masks for Boosted)
//   grad_w[t][e] += coef * |o| ;  grad_b[t][e] += coef ;  d|o| = coef * w
//   the min sends d|o| * sgn to the FIRST other edge (ascending column) attaining it: du[e*] += ... * sign(u[e*])
//   dv2c_t[e][z] = du[e][(z - s_e) mod Z]
// Gradients are sums over B*Z lanes accumulated with fp32 atomics: equal to autograd within fp32 summation-order noise.
#include <algorithm>

#include "nldpc_generic_common.cuh"

namespace nldpc {

__device__ __forceinline__ float clampf_b(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ float qlim(int qbit) {
    switch (qbit) {
        case 6: return 15.5f;
        case 5: return 7.5f;
        case -5: return 15.0f;
        case 4: return 7.0f;
        case 3: return 6.0f;
        default: return -1.0f;   // no quantisation: identity, gradient 1
    }
}
__device__ __forceinline__ float quantf_b(float x, int qbit) {
    switch (qbit) {
        case 6: return clampf_b(rintf(x), -15.5f, 15.5f);
        case 5: return clampf_b(mulf(rintf(mulf(x, 2.0f)), 0.5f), -7.5f, 7.5f);
        case -5: return clampf_b(rintf(x), -15.0f, 15.0f);
        case 4: return clampf_b(rintf(x), -7.0f, 7.0f);
        case 3: return clampf_b(mulf(rintf(mulf(x, 0.5f)), 2.0f), -6.0f, 6.0f);
        default: return x;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


template <int D>
__device__ __forceinline__ void cn_check_bwd(float *__restrict__ slab, float *__restrict__ acc_w, float *__restrict__ acc_b,
                                             int h, const GraphDev &g, int e0, int row_i, const BwdArgs &a, int t, size_t b_idx,
                                             bool active) {
    const int Z = g.Z;
    float u[D], dc[D];
    int addr[D];
    bool stored[D];
    const float *hv = a.hist_v2c + (((size_t)t * a.B + b_idx) * g.S) * Z;
    const float *xin = a.mode == 0 ? a.xa + b_idx * g.N * Z : a.hist_xin + (((size_t)(t + 1) * a.B + b_idx) * g.N) * Z;
#pragma unroll
    for (int k = 0; k < D; k++) {
        int zz = h + __ldg(g.e_shift + e0 + k);
        zz = (zz >= Z) ? zz - Z : zz;
        const int row = __ldg(g.e_row + e0 + k);
        stored[k] = row >= g.N;
        addr[k] = row * Z + zz;
        float v = 0.0f, d = 0.0f;
        if (active) {
            v = stored[k] ? __ldg(hv + (size_t)(row - g.N) * Z + zz) : __ldg(xin + (size_t)row * Z + zz);
            d = slab[addr[k]];                          // dc2v (stored edge) or G_t[j] (degree-1 block: row j < N)
        }
        u[k] = v;
        dc[k] = d;
    }
    // ---- recompute the forward of this check ----
    const bool is_qms = a.mode == 2, boosted = a.mode != 0;
    const float lim = is_qms ? qlim(a.qbit) : 0.0f;
    float uq[D], smask[D];      // conditioned CN inputs; straight-through mask of the pre-CN quantiser / clamp
    unsigned par = 0;
#pragma unroll
    for (int k = 0; k < D; k++) {
        float v = u[k];
        float mk = 1.0f;
        if (boosted) {
            if (is_qms) {
                if (lim >= 0.0f) mk = (fabsf(v) <= lim) ? 1.0f : 0.0f;
                v = quantf_b(v, a.qbit);
            } else {
                mk = (v >= a.lo && v <= a.hi) ? 1.0f : 0.0f;
                v = clampf_b(v, a.lo, a.hi);
            }
            v = addf(v, mulf(0.0001f, 1.0f - ((fabsf(v) > 0.0f) ? 1.0f : 0.0f)));
        }
        uq[k] = v;
        smask[k] = mk;
        par ^= (v > 0.0f) ? 1u : 0u;
    }
    // two smallest magnitudes with FIRST-index tie rule (torch.min on CPU)
    float m1 = 3.0e38f, m2 = 3.0e38f;
    int i1 = -1, i2 = -1;
#pragma unroll
    for (int k = 0; k < D; k++) {
        float av = fabsf(uq[k]);
        av = (av > 0.0f) ? av : 10000.0f;
        if (av < m1) { m2 = m1; i2 = i1; m1 = av; i1 = k; }
        else if (av < m2) { m2 = av; i2 = k; }
    }
    float s1 = 0.0f, s2 = 0.0f;     // gradient collected by the argmin edges i1 (from every k != i1) and i2 (from k == i1)
    const float *wt = a.w ? a.w + (size_t)t * g.E : nullptr;
    const float *bt = a.b ? a.b + (size_t)t * g.E : nullptr;
    float s_ucn = 0.0f;
    if (boosted && a.ucn_mix && a.hist_ucn && active) s_ucn = a.hist_ucn[(((size_t)t * a.B + b_idx) * g.M + row_i) * Z + h] ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 0; k < D; k++) {
        const int e = e0 + k;
        const float others_min = (k == i1) ? m2 : m1;
        const bool from_cap = !(others_min < 10000.0f);                // min came from a masked (constant) entry: no gradient
        float mag = fminf(others_min, 10000.0f);
        const unsigned npos_odd = par ^ ((uq[k] > 0.0f) ? 1u : 0u);
        const float sgn = npos_odd ? 1.0f : -1.0f;
        float coef = 0.0f, dmag = 0.0f, gwv = 0.0f, gbv = 0.0f;
        if (!boosted) {
            const float wk = __ldg(wt + e), bk = __ldg(bt + e);
            const float pre = addf(mulf(mag, wk), bk);
            coef = (pre > 0.0f) ? dc[k] * sgn : 0.0f;                 // dc2v * sign(o) * relu'
            gwv = coef * mag;
            gbv = coef;
            dmag = coef * wk * 1.0f;                                  // d|o|/do * do/dmag = sign(o) * sgn = 1
        } else {
            const float mag_adj = addf(mag, mulf(-0.0001f, addf(-((mag > 0.0001f) ? 1.0f : 0.0f), 1.0f)));
            const float o = mulf(mag_adj, sgn);
            const float ao = fabsf(o);
            const float so = (o > 0.0f) ? 1.0f : ((o < 0.0f) ? -1.0f : 0.0f);
            float wk = 1.0f, wc = 1.0f, wu = 1.0f;
            float pre;
            if (!wt) pre = ao;
            else if (a.ucn_mix && bt) {
                wc = __ldg(wt + e); wu = __ldg(bt + e);
                pre = addf(mulf(mulf(ao, wc), 1.0f - s_ucn), mulf(mulf(ao, wu), s_ucn));
                wk = wc * (1.0f - s_ucn) + wu * s_ucn;
            } else {
                wk = __ldg(wt + e);
                pre = mulf(ao, wk);
            }
            const float m0 = (pre > 0.0f) ? pre : 0.0f;
            float pm;                                                  // straight-through mask of the post quantiser / clamp
            if (is_qms) pm = (lim < 0.0f || fabsf(m0) <= lim) ? 1.0f : 0.0f;
            else pm = (m0 >= a.lo && m0 <= a.hi) ? 1.0f : 0.0f;
            coef = (pre > 0.0f) ? dc[k] * so * pm : 0.0f;             // d pre
            if (wt) {
                if (a.ucn_mix && bt) { gwv = coef * ao * (1.0f - s_ucn); gbv = coef * ao * s_ucn; }
                else gwv = coef * ao;
            }
            dmag = coef * wk * so * sgn;                               // d|o| * sign(o) * d o / d mag_adj
        }
        if (!from_cap) {
            if (k == i1) s2 += dmag; else s1 += dmag;
        }
        // per-edge weight gradients: reduce over the warp, one shared-memory atomic per warp
        if (wt) {
            const float r = warp_sum(gwv);
            if ((threadIdx.x & 31) == 0) atomicAdd(acc_w + e, r);
        }
        if (bt && (!boosted || a.ucn_mix)) {
            const float r = warp_sum(gbv);
            if ((threadIdx.x & 31) == 0) atomicAdd(acc_b + e, r);
        }
    }
    // ---- gradient w.r.t. the CN inputs (only the argmin edges receive any), back through conditioning, to dv2c ----
#pragma unroll
    for (int k = 0; k < D; k++) {
        if (!stored[k]) continue;                       // degree-1 block: its v2c is the channel input (no weight upstream... see VN chain)
        float du = 0.0f;
        if (k == i1) du = s1;
        else if (k == i2) du = s2;
        const float su = (uq[k] > 0.0f) ? 1.0f : ((uq[k] < 0.0f) ? -1.0f : 0.0f);
        du = du * su * smask[k];
        if (active) slab[addr[k]] = du;                 // becomes dv2c_t in the variable-lane domain
    }
    // degree-1 blocks with a VN weight chain need du as well: hand it back through the G rows (row j < N)
    if (a.gvn) {
#pragma unroll
        for (int k = 0; k < D; k++) {
            if (stored[k]) continue;
            float du = 0.0f;
            if (k == i1) du = s1;
            else if (k == i2) du = s2;
            const float su = (uq[k] > 0.0f) ? 1.0f : ((uq[k] < 0.0f) ? -1.0f : 0.0f);
            if (active) slab[addr[k]] = du * su * smask[k];   // overwrite G_t[j] (already consumed) with d xin_t[j] from this edge
        }
    }
}

__global__ void __launch_bounds__(256, 1) nldpc_backward_kernel(const GraphDev g, const BwdArgs a, const int cw_per_cta, const int stride) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *slabs = reinterpret_cast<float *>(smem_raw);
    float *acc_w = slabs + (size_t)cw_per_cta * stride;   // [E]
    float *acc_b = acc_w + g.E;                           // [E]
    float *acc_v = acc_b + g.E;                           // [N]
    const int Z = g.Z, NZ = g.N * g.Z;
    const int tid = threadIdx.x;
    const int L = cw_per_cta * Z;
    const int cw = tid / Z;
    const int z = tid - cw * Z;
    float *slab = slabs + (size_t)cw * stride;
    const int n_tiles = (a.B + cw_per_cta - 1) / cw_per_cta;
    const bool boosted = a.mode != 0, is_qms = a.mode == 2;
    const float lim = is_qms ? qlim(a.qbit) : -1.0f;

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b0 = tile * cw_per_cta;
        const int ncw = min(cw_per_cta, a.B - b0);
        const bool active = (tid < L) && (cw < ncw);
        const size_t b = (size_t)(b0 + cw);
        const bool lane_ok = tid < L;
        if (lane_ok)
            for (int s = 0; s < g.S; s++) slab[(g.N + s) * Z + z] = 0.0f;      // dv2c_T = 0
        // d xin chain (Boosted VN weights): per block j, kept in the extra rows [N+S, N+S+N)
        if (a.gvn && lane_ok)
            for (int j = 0; j < g.N; j++) slab[(g.N + g.S + j) * Z + z] = 0.0f;
        __syncthreads();

        for (int t = a.T - 1; t >= 0; t--) {
            for (int i = tid; i < 2 * g.E + g.N; i += blockDim.x) acc_w[i] = 0.0f;
            // G_t rows (through the output clamp mask for Boosted, :520-521)
            if (lane_ok) {
                for (int j = 0; j < g.N; j++) {
                    float gv = 0.0f;
                    if (active) {
                        gv = __ldg(a.gout + ((size_t)t * a.B + b) * NZ + j * Z + z);
                        if (boosted && !a.hist_mask[((size_t)t * a.B + b) * NZ + j * Z + z]) gv = 0.0f;
                    }
                    slab[j * Z + z] = gv;
                }
            }
            __syncthreads();
            // ---- VN phase backwards: dc2v[e] = G[j] + (sum of the block's dv2c_{t+1} - own) ----
            if (lane_ok) {
                for (int c = 0; c < g.n_vcols; c++) {
                    const int j = __ldg(g.vcol_j + c);
                    const int p0 = __ldg(g.vcol_ptr + c), p1 = __ldg(g.vcol_ptr + c + 1);
                    float tot = 0.0f;
                    for (int k = p0; k < p1; k++) tot += slab[__ldg(g.vcol_row + k) * Z + z];
                    const float gj = slab[j * Z + z];
                    for (int k = p0; k < p1; k++) {
                        const int r = __ldg(g.vcol_row + k) * Z + z;
                        slab[r] = gj + (tot - slab[r]);
                    }
                }
            }
            __syncthreads();
            // ---- CN phase backwards ----
            for (int i = 0; i < g.M; i++) {
                const int e0 = __ldg(g.row_ptr + i);
                const int d = __ldg(g.row_ptr + i + 1) - e0;
#define NLDPC_CNB_CASE(D) cn_check_bwd<D>(slab, acc_w, acc_b, z, g, e0, i, a, t, b, active)
                NLDPC_DEG_SWITCH(d, NLDPC_CNB_CASE)
#undef NLDPC_CNB_CASE
            }
            __syncthreads();
            // ---- Boosted: channel-input chain.  d xin_t[j] = sum_{e in col(j)} dv2c_t[e] + chain_{t+1}[j];
            //      xin_t = q(xin_{t-1} * wVN_t): dz = d xin_t * [|xin_{t-1} * w| <= lim] (QMS only), grad_wVN += dz * xin_{t-1},
            //      chain_t = dz * wVN_t   (BoostedNeuralLDPCDecoder.py:325-337)
            if (a.gvn) {     // (every thread of the CTA takes part: warp_sum below needs full warps)
                const float *vw = a.vn_w + (size_t)t * g.N;
                if (lane_ok) {
                    // blocks of degree >= 2: their G row is not a gradient message; add the stored dv2c_t of the block's edges
                    for (int c = 0; c < g.n_vcols; c++) {
                        const int j = __ldg(g.vcol_j + c);
                        const int p0 = __ldg(g.vcol_ptr + c), p1 = __ldg(g.vcol_ptr + c + 1);
                        float tot = 0.0f;
                        for (int k = p0; k < p1; k++) tot += slab[__ldg(g.vcol_row + k) * Z + z];
                        slab[j * Z + z] = tot;            // G_t[j] has been consumed by the VN phase above
                    }
                }
                for (int j = 0; j < g.N; j++) {
                    // row j now holds sum_e dv2c_t[e] of block j (degree-1 blocks: left there by the CN phase)
                    float dx = 0.0f, xprev = 0.0f;
                    if (lane_ok) dx = slab[(g.N + g.S + j) * Z + z] + slab[j * Z + z];
                    if (active) xprev = __ldg(a.hist_xin + (((size_t)t * a.B + b) * g.N + j) * Z + z);
                    const float wv = __ldg(vw + j);
                    float dz = dx;
                    if (is_qms && lim >= 0.0f && !(fabsf(mulf(xprev, wv)) <= lim)) dz = 0.0f;
                    const float gv = warp_sum(active ? dz * xprev : 0.0f);
                    if ((tid & 31) == 0) atomicAdd(acc_v + j, gv);
                    if (lane_ok) slab[(g.N + g.S + j) * Z + z] = dz * wv;          // chain_t
                }
            }
            __syncthreads();
            for (int i = tid; i < g.E; i += blockDim.x) {
                if (a.gw && acc_w[i] != 0.0f) atomicAdd(a.gw + (size_t)t * g.E + i, acc_w[i]);
                if (a.gb && acc_b[i] != 0.0f) atomicAdd(a.gb + (size_t)t * g.E + i, acc_b[i]);
            }
            if (a.gvn)
                for (int i = tid; i < g.N; i += blockDim.x)
                    if (acc_v[i] != 0.0f) atomicAdd(a.gvn + (size_t)t * g.N + i, acc_v[i]);
            __syncthreads();
        }
    }
}

int backward_prepare() {
    return (int)cudaFuncSetAttribute(nldpc_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
}

// returns 0, cudaError_t, or -2 when one codeword does not fit
int backward_launch(const GraphDev &g, const BwdArgs &a, int sm_count, cudaStream_t st) {
    int rows = g.N + g.S + (a.gvn ? g.N : 0);
    int stride = rows * g.Z;
    while ((stride & 31) != (g.Z & 31)) stride++;
    const size_t extra = (size_t)(2 * g.E + g.N) * 4 + 64;
    const size_t per_cw = (size_t)stride * 4;
    if (g.Z > 256 || per_cw + extra > (size_t)kSmemBudget) return -2;
    const size_t half_budget = (size_t)(kSmemBudget - 2048) / 2;
    int cw = (int)std::min<size_t>((half_budget - extra) / per_cw, (size_t)(256 / g.Z));
    if (cw < 1) cw = (int)std::min<size_t>(((size_t)kSmemBudget - extra) / per_cw, (size_t)(256 / g.Z));
    if (cw < 1) cw = 1;
    const int threads = ((cw * g.Z + 31) / 32) * 32;
    const size_t smem = (size_t)cw * per_cw + extra;
    const int n_tiles = (a.B + cw - 1) / cw;
    const int ctas_per_sm = std::max(1, (int)((size_t)kSmemBudget / (smem + 1024)));
    const int grid = std::min(n_tiles, sm_count * ctas_per_sm);
    nldpc_backward_kernel<<<grid, threads, smem, st>>>(g, a, cw, stride);
    return (int)cudaGetLastError();
}

}  // namespace nldpc
