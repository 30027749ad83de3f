/*
 * nldpc_oracle.c — CPU restatement (TEST INFRASTRUCTURE, not product code) of the reference's
 * iterative neural belief-propagation decode.
 *
 *   reference: ShapeLayer/neural-ldpc-decoder-torch
 *     src/neural_ldpc_decoder/NeuralLDPCDecoder.py:44-100          (NeuralLDPCDecoder.forward)
 *     src/boosted_neural_ldpc_decoder/BoostedNeuralLDPCDecoder.py:187-214, 260-538
 *     src/neural_ldpc_decoder/ConnectingMatrix.py:68-140            (edge orders, shift direction)
 *
 * The reference expresses the Tanner-graph message passing as dense 0/1 matmuls and an
 * [B,Z,E,E] tile; this file restates the same arithmetic sparsely (O(E*Z) per iteration) with
 * the SAME fp32 operations in the SAME order, so results are bit-identical to the reference on
 * CPU (pinned by the .npz fixtures under tests/golden/, generated from the live reference by tools/gen_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (neural_ldpc_decoder_torch_b200/csrc) never does.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -pthread; NO -ffast-math: every + and *
 * below must be one IEEE fp32 rounding, exactly as in ATen's elementwise kernels / MKL sgemm
 * with 0/1 matrices, whose k-loop is a sequential fp32 accumulation starting from +0).
 *
 * Conventions:
 *   edge e = (i,j) for bg[i,j] != -1, indexed ROW-MAJOR (i outer, j inner): this is the index of
 *   weights_var[t], biases_var[t], per-edge boosted weights and self.llr[..][:, :, e];
 *   shift s_e = bg[i,j] mod Z;  gather  u[e][h] = v2c[e][(h+s_e) mod Z]   (Lift_Matrix1^T)
 *                               scatter op[e][z] = o[e][(z-s_e) mod Z]     (Lift_Matrix2)
 *   channel input xa[B][N][Z], outputs out[t][B][N*Z] with bit index j*Z+z.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ---- tiny parallel-for over codewords (pthreads; thread count = NLDPC_ORACLE_THREADS or all online cores) ---- */
typedef void (*range_fn)(void *ctx, int begin, int end);
typedef struct { range_fn fn; void *ctx; int begin, end; } pf_job_t;
static void *pf_tramp(void *p) { pf_job_t *j = (pf_job_t *)p; j->fn(j->ctx, j->begin, j->end); return NULL; }
int nldpc_oracle_num_threads(void) {
    const char *env = getenv("NLDPC_ORACLE_THREADS");
    long n = env ? atol(env) : sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}
static void parallel_for(int n, range_fn fn, void *ctx) {
    int nt = nldpc_oracle_num_threads();
    if (nt > n) nt = n > 0 ? n : 1;
    if (nt <= 1) { fn(ctx, 0, n); return; }
    pthread_t th[256]; pf_job_t jobs[256];
    int started[256];
    for (int t = 0; t < nt; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].begin = (int)((long)n * t / nt); jobs[t].end = (int)((long)n * (t + 1) / nt);
        started[t] = (pthread_create(&th[t], NULL, pf_tramp, &jobs[t]) == 0);
        if (!started[t]) fn(ctx, jobs[t].begin, jobs[t].end);
    }
    for (int t = 0; t < nt; t++) if (started[t]) pthread_join(th[t], NULL);
}

typedef struct {
    int M, N, Z, E;
    int *erow, *ecol, *eshift; /* [E] row-major edges */
    int *row_ptr;              /* [M+1]: row i owns edges row_ptr[i]..row_ptr[i+1]-1 */
    int *col_ptr, *col_edges;  /* [N+1], [E]: edges of column j in ascending row */
} graph_t;

static int graph_build(graph_t *g, const int32_t *bg, int M, int N, int Z) {
    int E = 0;
    for (int i = 0; i < M * N; i++) E += (bg[i] != -1);
    g->M = M; g->N = N; g->Z = Z; g->E = E;
    g->erow = malloc(sizeof(int) * (E + 1)); g->ecol = malloc(sizeof(int) * (E + 1));
    g->eshift = malloc(sizeof(int) * (E + 1));
    g->row_ptr = malloc(sizeof(int) * (M + 1));
    g->col_ptr = malloc(sizeof(int) * (N + 1)); g->col_edges = malloc(sizeof(int) * (E + 1));
    if (!g->erow || !g->ecol || !g->eshift || !g->row_ptr || !g->col_ptr || !g->col_edges) return -1;
    int e = 0;
    for (int i = 0; i < M; i++) {
        g->row_ptr[i] = e;
        for (int j = 0; j < N; j++) {
            int v = bg[i * N + j];
            if (v == -1) continue;
            g->erow[e] = i; g->ecol[e] = j;
            g->eshift[e] = ((v % Z) + Z) % Z; /* ConnectingMatrix.py:73,82 */
            e++;
        }
    }
    g->row_ptr[M] = e;
    int k = 0;
    for (int j = 0; j < N; j++) {
        g->col_ptr[j] = k;
        for (int q = 0; q < E; q++) if (g->ecol[q] == j) g->col_edges[k++] = q; /* ascending rm == ascending row */
    }
    g->col_ptr[N] = k;
    return 0;
}

static void graph_free(graph_t *g) {
    free(g->erow); free(g->ecol); free(g->eshift); free(g->row_ptr); free(g->col_ptr); free(g->col_edges);
}

static inline float signf_(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }

/* VN update, NeuralLDPCDecoder.py:56-58 / Boosted :376-378:
 * v2c[e] = xin[j] + ( ((0 + c2v[e1]) + c2v[e2]) + ... ), others of the same column, ascending row, skipping e. */
static void vn_update(const graph_t *g, const float *xin /*[N][Z]*/, const float *c2v /*[E][Z]*/, float *v2c) {
    const int Z = g->Z;
    for (int j = 0; j < g->N; j++) {
        const int *ce = g->col_edges + g->col_ptr[j];
        const int d = g->col_ptr[j + 1] - g->col_ptr[j];
        for (int k = 0; k < d; k++) {
            for (int z = 0; z < Z; z++) {
                float acc = 0.0f;
                for (int k2 = 0; k2 < d; k2++) if (k2 != k) acc = acc + c2v[ce[k2] * Z + z];
                v2c[ce[k] * Z + z] = xin[j * Z + z] + acc;
            }
        }
    }
}

/* column marginal: tot = sequential sum of ALL column messages from +0 (llr @ W_output), NeuralLDPCDecoder.py:94 */
static void col_total(const graph_t *g, const float *c2v, float *tot /*[N][Z]*/) {
    const int Z = g->Z;
    for (int j = 0; j < g->N; j++)
        for (int z = 0; z < Z; z++) {
            float acc = 0.0f;
            for (int k = g->col_ptr[j]; k < g->col_ptr[j + 1]; k++) acc = acc + c2v[g->col_edges[k] * Z + z];
            tot[j * Z + z] = acc;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* NeuralLDPCDecoder.forward (NeuralLDPCDecoder.py:44-100).  out: [T][B][N*Z].                 */
typedef struct {
    const graph_t *g; const float *xa, *w, *b; int B, T; float *out; int rc;
    int last_only; /* out is [B][N*Z]: only the last iteration is kept (full-size parity runs) */
} neural_ctx_t;

static void neural_range(void *p, int begin, int end) {
    neural_ctx_t *c = (neural_ctx_t *)p;
    const graph_t g = *c->g;
    const int E = g.E, Z = g.Z, M = g.M, NZ = g.N * g.Z, B = c->B;
    float *c2v = malloc(sizeof(float) * E * Z), *v2c = malloc(sizeof(float) * E * Z);
    float *u = malloc(sizeof(float) * E * Z), *o = malloc(sizeof(float) * E * Z), *tot = malloc(sizeof(float) * NZ);
    if (!c2v || !v2c || !u || !o || !tot) { c->rc = -1; goto done; }
    for (int cw = begin; cw < end; cw++) {
        const float *x = c->xa + (size_t)cw * NZ;
        memset(c2v, 0, sizeof(float) * E * Z); /* :49 */
        for (int t = 0; t < c->T; t++) {
            vn_update(&g, x, c2v, v2c);                                              /* :56-58 */
            for (int e = 0; e < E; e++)                                              /* :59-63 */
                for (int h = 0; h < Z; h++) u[e * Z + h] = v2c[e * Z + (h + g.eshift[e]) % Z];
            for (int i = 0; i < M; i++) {                                            /* :66-80 */
                for (int h = 0; h < Z; h++) {
                    for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; e++) {
                        /* masked tile entries (incl. self) are 0 -> |0| + 10000: the min is capped at 10000 (:74-75) */
                        float mag = 10000.0f;
                        int npos = 0;
                        for (int e2 = g.row_ptr[i]; e2 < g.row_ptr[i + 1]; e2++) {
                            if (e2 == e) continue;
                            float v = u[e2 * Z + h];
                            float a = fabsf(v);
                            a = a + 10000.0f * (1.0f - (a > 0.0f ? 1.0f : 0.0f));
                            if (a < mag) mag = a;
                            npos += (v > 0.0f);                                      /* :77-78: -x < 0 */
                        }
                        float prod = (npos & 1) ? -1.0f : 1.0f;                      /* prod of (1 - 2[x>0]) */
                        o[e * Z + h] = mag * signf_(-prod);                          /* :79-80 */
                    }
                }
            }
            const float *wt = c->w + (size_t)t * E, *bt = c->b + (size_t)t * E;
            for (int e = 0; e < E; e++)                                              /* :82-91 */
                for (int z = 0; z < Z; z++) {
                    float op = o[e * Z + ((z - g.eshift[e]) % Z + Z) % Z];
                    float pre = fabsf(op) * wt[e];
                    pre = pre + bt[e];
                    float m = pre * (pre > 0.0f ? 1.0f : 0.0f);
                    c2v[e * Z + z] = m * signf_(op);
                }
            col_total(&g, c2v, tot);                                                 /* :94-98 */
            if (c->last_only && t != c->T - 1) continue;
            float *ot = c->out + (c->last_only ? (size_t)cw : (size_t)t * B + cw) * NZ;
            for (int q = 0; q < NZ; q++) ot[q] = x[q] + tot[q];
        }
    }
done:
    free(c2v); free(v2c); free(u); free(o); free(tot);
}

int nldpc_oracle_neural_forward(const int32_t *bg, int M, int N, int Z,
                                const float *xa /*[B][N][Z]*/, const float *w /*[T][E]*/, const float *b /*[T][E]*/,
                                int B, int T, float *out) {
    graph_t g;
    if (graph_build(&g, bg, M, N, Z)) return -1;
    neural_ctx_t c = { &g, xa, w, b, B, T, out, 0, 0 };
    parallel_for(B, neural_range, &c);
    graph_free(&g);
    return c.rc;
}

/* the same, keeping only the last iteration's output: out_last [B][N*Z] */
int nldpc_oracle_neural_forward_last(const int32_t *bg, int M, int N, int Z,
                                     const float *xa, const float *w, const float *b, int B, int T, float *out_last) {
    graph_t g;
    if (graph_build(&g, bg, M, N, Z)) return -1;
    neural_ctx_t c = { &g, xa, w, b, B, T, out_last, 0, 1 };
    parallel_for(B, neural_range, &c);
    graph_free(&g);
    return c.rc;
}

/* ------------------------------------------------------------------------------------------ */
/* BoostedNeuralLDPCDecoder._quantize_message forward value (:187-214).
 * rintf = round-half-to-even = torch.round.  q_bit values other than 6,5,-5,4,3: identity.      */
static inline float clampf_(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline float quantize_(float x, int q) {
    switch (q) {
    case 6:  return clampf_(rintf(x), -15.5f, 15.5f);
    case 5:  return clampf_(rintf(x * 2.0f) / 2.0f, -7.5f, 7.5f);
    case -5: return clampf_(rintf(x), -15.0f, 15.0f);
    case 4:  return clampf_(rintf(x), -7.0f, 7.0f);
    case 3:  return clampf_(rintf(x / 2.0f) * 2.0f, -6.0f, 6.0f);
    default: return x;
    }
}

void nldpc_oracle_quantize(const float *x, float *y, long n, int qbit) {
    for (long i = 0; i < n; i++) y[i] = quantize_(x[i], qbit);
}

/* One iteration of BoostedNeuralLDPCDecoder.forward's loop body (:320-531), with the sharing
 * types already folded (by the caller) into per-column / per-edge weight rows:
 *   decoder_type: 0 SP, 1 MS, 2 QMS (struct/DecoderType.py)
 *   vn_w   [N] or NULL  -> xin *= vn_w (cumulative state!, :325-334), then quantised if QMS (:336)
 *   cn_w   [E] or NULL  (NULL = sharing type 0, no multiply, :433)
 *   ucn_w  [E] or NULL  (only used when ucn_mix != 0: W = ucn ? ucn_w : cn_w, :436-488)
 *   ucn_app [B][N*Z] or NULL: APP used for the unsatisfied-check indicator when UCN sharing > 0:
 *            previous iteration's output (t>0); NULL means "use xin" (t==0) (:339-346).  The
 *            indicator is computed whenever compute_ucn != 0 (the reference computes it even when unused).
 *   xin    [B][N][Z] in/out: the compounding, (re)quantised channel input (xa_input)
 *   xo     [B][N][Z] in/out: xa_origin, re-quantised every iteration if QMS (:517-518)
 *   llr_in [B][E][Z]  c2v of the previous iteration (self.llr[curr_iter], stored here edge-major)
 *   llr_out[B][E][Z]  -> self.llr[curr_iter+1];   out [B][N*Z] -> self.outputs[curr_iter]
 */
typedef struct {
    const graph_t *g; int decoder_type, qbit; float llr_lo, llr_hi;
    const float *vn_w, *cn_w, *ucn_w; int compute_ucn, ucn_mix; const float *ucn_app;
    float *xin, *xo; const float *llr_in; float *llr_out, *out; int rc;
} boosted_ctx_t;

static void boosted_range(void *p, int begin, int end) {
    boosted_ctx_t *c = (boosted_ctx_t *)p;
    const graph_t g = *c->g;
    const int E = g.E, Z = g.Z, M = g.M, N = g.N, NZ = g.N * g.Z, qbit = c->qbit;
    const int is_qms = (c->decoder_type == 2), is_sp = (c->decoder_type == 0);
    const float llr_lo = c->llr_lo, llr_hi = c->llr_hi;
    const float *vn_w = c->vn_w, *cn_w = c->cn_w, *ucn_w = c->ucn_w;
    float *v2c = malloc(sizeof(float) * E * Z), *u = malloc(sizeof(float) * E * Z);
    float *o = malloc(sizeof(float) * E * Z), *tot = malloc(sizeof(float) * NZ);
    unsigned char *ucn_chk = malloc((size_t)M * Z);
    if (!v2c || !u || !o || !tot || !ucn_chk) { c->rc = -1; goto done; }
    for (int cw = begin; cw < end; cw++) {
        float *x = c->xin + (size_t)cw * NZ, *xorig = c->xo + (size_t)cw * NZ;
        const float *c2v = c->llr_in + (size_t)cw * E * Z;
        float *c2v_new = c->llr_out + (size_t)cw * E * Z;
        if (vn_w) for (int j = 0; j < N; j++) for (int z = 0; z < Z; z++) x[j * Z + z] = x[j * Z + z] * vn_w[j]; /* :327-334 */
        if (is_qms) for (int q = 0; q < NZ; q++) x[q] = quantize_(x[q], qbit);                     /* :336-337 */
        memset(ucn_chk, 0, (size_t)M * Z);
        if (c->compute_ucn) {                                                                     /* :339-368 */
            const float *app = c->ucn_app ? c->ucn_app + (size_t)cw * NZ : x;
            for (int i = 0; i < M; i++)
                for (int h = 0; h < Z; h++) {
                    int nneg = 0; /* sign = (-app > 0) ? +1 : -1 ; product < 0  <=> odd number of -1 */
                    for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; e++) {
                        float a = -app[g.ecol[e] * Z + (h + g.eshift[e]) % Z];
                        nneg += !(a > 0.0f);
                    }
                    ucn_chk[i * Z + h] = (unsigned char)(nneg & 1);
                }
        }
        vn_update(&g, x, c2v, v2c);                                                               /* :376-378 */
        for (int e = 0; e < E; e++)                                                               /* :380-393 */
            for (int h = 0; h < Z; h++) {
                float v = v2c[e * Z + (h + g.eshift[e]) % Z];
                v = is_qms ? quantize_(v, qbit) : clampf_(v, llr_lo, llr_hi);
                if (!is_sp) v = v + 0.0001f * (1.0f - (fabsf(v) > 0.0f ? 1.0f : 0.0f));
                u[e * Z + h] = v;
            }
        for (int i = 0; i < M; i++)
            for (int h = 0; h < Z; h++)
                for (int e = g.row_ptr[i]; e < g.row_ptr[i + 1]; e++) {
                    if (is_sp) {                                                                  /* :400-408 */
                        /* product over the others in ascending column order; masked entries are 1.
                         * torch.prod's internal order / SLEEF tanh differ from libm, so SP is only
                         * tolerance-comparable. */
                        float pr = 1.0f;
                        for (int e2 = g.row_ptr[i]; e2 < g.row_ptr[i + 1]; e2++) {
                            if (e2 == e) continue;
                            float th = tanhf(-0.5f * u[e2 * Z + h]);
                            if (!(fabsf(th) > 0.0f)) th = th + 1.0f;
                            pr = pr * th;
                        }
                        pr = clampf_(pr, -1.0f + 1e-7f, 1.0f - 1e-7f);
                        o[e * Z + h] = -2.0f * atanhf(pr);
                    } else {                                                                      /* :409-423 */
                        float mag = 10000.0f; /* the (masked) self entry is always in the min */
                        int npos = 0;
                        for (int e2 = g.row_ptr[i]; e2 < g.row_ptr[i + 1]; e2++) {
                            if (e2 == e) continue;
                            float v = u[e2 * Z + h];
                            float a = fabsf(v);
                            a = a + 10000.0f * (1.0f - (a > 0.0f ? 1.0f : 0.0f));
                            if (a < mag) mag = a;
                            npos += (v > 0.0f);
                        }
                        mag = mag + (-0.0001f) * (-(mag > 0.0001f ? 1.0f : 0.0f) + 1.0f);         /* :416 */
                        float prod = (npos & 1) ? -1.0f : 1.0f;
                        o[e * Z + h] = mag * signf_(-1.0f * prod);
                    }
                }
        for (int e = 0; e < E; e++)                                                               /* :425-512 */
            for (int z = 0; z < Z; z++) {
                const int h = ((z - g.eshift[e]) % Z + Z) % Z;
                float op = o[e * Z + h];
                float a = fabsf(op), pre;
                if (!cn_w) pre = a;
                else if (c->ucn_mix) {
                    float s = ucn_chk[g.erow[e] * Z + h] ? 1.0f : 0.0f;
                    pre = (a * cn_w[e]) * (1.0f - s) + (a * ucn_w[e]) * s;
                } else pre = a * cn_w[e];
                float m = pre * (pre > 0.0f ? 1.0f : 0.0f);
                m = is_qms ? quantize_(m, qbit) : clampf_(m, llr_lo, llr_hi);
                c2v_new[e * Z + z] = m * signf_(op);
            }
        col_total(&g, c2v_new, tot);                                                              /* :513-526 */
        if (is_qms) for (int q = 0; q < NZ; q++) xorig[q] = quantize_(xorig[q], qbit);
        float *ot = c->out + (size_t)cw * NZ;
        for (int q = 0; q < NZ; q++) ot[q] = clampf_(xorig[q] + tot[q], llr_lo, llr_hi);
    }
done:
    free(v2c); free(u); free(o); free(tot); free(ucn_chk);
}

int nldpc_oracle_boosted_step(const int32_t *bg, int M, int N, int Z, int decoder_type, int qbit,
                              float llr_lo, float llr_hi,
                              const float *vn_w, const float *cn_w, const float *ucn_w,
                              int compute_ucn, int ucn_mix, const float *ucn_app,
                              float *xin, float *xo, const float *llr_in, float *llr_out, float *out, int B) {
    graph_t g;
    if (graph_build(&g, bg, M, N, Z)) return -1;
    boosted_ctx_t c = { &g, decoder_type, qbit, llr_lo, llr_hi, vn_w, cn_w, ucn_w, compute_ucn, ucn_mix, ucn_app,
                        xin, xo, llr_in, llr_out, out, 0 };
    parallel_for(B, boosted_range, &c);
    graph_free(&g);
    return c.rc;
}

/* hard decisions, reference predicate `out < 0` (Functions.py:90), packed little-endian bit order:
 * bit q of codeword cw lives in byte q/8, bit q%8 (== np.packbits(out<0, axis=1, bitorder='little')). */
void nldpc_oracle_pack_hard(const float *out /*[B][NZ]*/, int B, int NZ, uint8_t *packed /*[B][(NZ+7)/8]*/) {
    const int nb = (NZ + 7) / 8;
    memset(packed, 0, (size_t)B * nb);
    for (int cw = 0; cw < B; cw++)
        for (int q = 0; q < NZ; q++)
            if (out[(size_t)cw * NZ + q] < 0.0f) packed[(size_t)cw * nb + q / 8] |= (uint8_t)(1u << (q % 8));
}
