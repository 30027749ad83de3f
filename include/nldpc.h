/*
 * nldpc.h — C ABI of libnldpc_b200.so: B200 (sm_100a) neural belief-propagation LDPC decode.
 *
 * This is the drop-in boundary for ONE hot path of ShapeLayer/neural-ldpc-decoder-torch: the
 * iteration loops of
 *     NeuralLDPCDecoder.forward         src/neural_ldpc_decoder/NeuralLDPCDecoder.py:44-100
 *     BoostedNeuralLDPCDecoder.forward  src/boosted_neural_ldpc_decoder/BoostedNeuralLDPCDecoder.py:260-538
 * and their backward (autograd of the above + LDPCDecoderLoss.py:73-108).  The reference has no
 * FFI of its own (pure PyTorch); these entry points are what a ctypes / torch.library binding on
 * the reference side binds.  Plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - All `*_dev` pointers are device pointers on the graph's device; `*_host` are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All device entry
 *     points are asynchronous on that stream and never synchronise.
 *   - Return value: 0 on success; >0 = cudaError_t; <0 = NLDPC_E_*.  nldpc_last_error() returns a
 *     thread-local human readable message for the last failure on the calling thread.
 *   - Edge index e is the reference's ROW-MAJOR edge index (check row outer, variable column inner,
 *     ConnectingMatrix.py:78-85): the layout of weights_var[t] / biases_var[t] / weight_CN_t.
 *   - xa is [B][N][Z] fp32 (channel LLR, LLR>0 <=> bit 1 as in the reference), outputs are
 *     [.., B][N*Z] fp32 with bit index j*Z+z.  Packed hard decisions: bit q of codeword b is
 *     bit (q%8) of byte hard[b*ceil(N*Z/8) + q/8], value (out < 0) — the reference predicate of
 *     Functions.evaluate_ber_fer (Functions.py:90).
 */
#ifndef NLDPC_H_
#define NLDPC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLDPC_ABI_VERSION 4

/* error codes (negative); positive return values are cudaError_t */
#define NLDPC_OK 0
#define NLDPC_E_INVALID (-1)     /* bad argument (null pointer, non-positive size, bad enum) */
#define NLDPC_E_UNSUPPORTED (-2) /* graph too large for the on-chip design (see nldpc_graph_create) */
#define NLDPC_E_NOMEM (-3)
#define NLDPC_E_NODEVICE (-4)    /* no CUDA device / not an sm_100 device */

typedef struct nldpc_graph nldpc_graph_t; /* opaque: per-device Tanner-graph tables */

/* output selection for the per-iteration soft outputs / hard decisions */
#define NLDPC_OUT_NONE 0
#define NLDPC_OUT_ALL 1  /* every iteration: soft [T][B][N*Z], hard [T][B][ceil(N*Z/8)] */
#define NLDPC_OUT_LAST 2 /* last iteration only: soft [B][N*Z], hard [B][ceil(N*Z/8)] */

/* decoder arithmetic of the Boosted decoder (struct/DecoderType.py) */
#define NLDPC_DEC_SP 0
#define NLDPC_DEC_MS 1
#define NLDPC_DEC_QMS 2

const char *nldpc_last_error(void);
int nldpc_abi_version(void);

/* Replaces ConnectingMatrix.__init__/_init_conn_matrix + ConnectingMatrixTorch.__init__
 * (neural ConnectingMatrix.py:4-140, ConnectingMatrixTorch.py:7-46; boosted :5-163 / :7-54):
 * builds the edge/shift tables for `basegraph` ([M][N] row-major int32, -1 = no edge, entry used
 * modulo Z) on CUDA device `device` instead of dense 0/1 matrices.
 * NLDPC_E_UNSUPPORTED if one codeword's state ((N + #edges in columns of degree>=2) * Z floats)
 * does not fit the 227 KB of shared memory, or a node degree exceeds 32. */
int nldpc_graph_create(const int32_t *basegraph, int M, int N, int Z, int device, nldpc_graph_t **out);
void nldpc_graph_destroy(nldpc_graph_t *g);
/* info[0..7] = {M, N, Z, E, stored_slots, codewords_per_cta, threads_per_cta, specialised(0/1)} */
int nldpc_graph_info(const nldpc_graph_t *g, int32_t info[8]);

/* Replaces the loop of NeuralLDPCDecoder.forward (NeuralLDPCDecoder.py:54-98) for B codewords and
 * T iterations: VN update (:56-58), circulant gather (:59-63), min-sum CN with the reference's
 * zero/10000 masking and sign rule (:66-80), scatter (:82-86), learned |m|*w+b, ReLU, sign (:89-91)
 * and the marginal (:94-98) in ONE kernel; messages never leave shared memory.
 *   w_dev, b_dev : [T][E] fp32 (weights_var / biases_var stacked)
 *   soft_dev     : per soft_mode (may be NULL with NLDPC_OUT_NONE)
 *   hard_dev     : per hard_mode (may be NULL with NLDPC_OUT_NONE), packed (out<0) bits
 * Results are bit-identical to the reference's CPU fp32 results (same operations, same order). */
int nldpc_neural_forward(const nldpc_graph_t *g, const float *xa_dev, const float *w_dev, const float *b_dev,
                         int B, int T, int soft_mode, float *soft_dev, int hard_mode, uint8_t *hard_dev,
                         void *stream);

/* Host-buffer convenience used for end-to-end timing and by non-torch callers: copies xa from host
 * memory (pinned or pageable) in chunks, decodes, copies the selected results back; H2D / kernel / D2H
 * of consecutive chunks overlap on internal streams.  Synchronous: returns when the host buffers are
 * filled.  Shapes as in nldpc_neural_forward with host pointers (w_host/b_host are [T][E]). */
int nldpc_neural_decode_host(const nldpc_graph_t *g, const float *xa_host, const float *w_host, const float *b_host,
                             int B, int T, int soft_mode, float *soft_host, int hard_mode, uint8_t *hard_host);

/* The same with the channel LLRs in a narrower host format: the path is host->device-link bound (3 328 B per BG2 codeword in
 * fp32), and receivers usually hold quantised LLRs anyway.  NLDPC_LLR_F16: IEEE half values; NLDPC_LLR_Q8: int8 codes,
 * x = scale * q.  Each chunk is expanded to fp32 on the device and decoded by the same kernels: the results are bit-identical
 * to nldpc_neural_decode_host on the widened values (exact whenever the caller's LLRs are representable in the format). */
#define NLDPC_LLR_F16 1
#define NLDPC_LLR_Q8 2
int nldpc_neural_decode_host_narrow(const nldpc_graph_t *g, const void *x_host, int x_format, float scale, const float *w_host,
                                    const float *b_host, int B, int T, int soft_mode, float *soft_host, int hard_mode,
                                    uint8_t *hard_host);

/* Backward of nldpc_neural_forward w.r.t. w and b (closed form of autograd through
 * NeuralLDPCDecoder.py:54-98).  The forward is re-run in a training-dump mode that
 * spills the per-iteration v2c to `workspace_dev` (HBM is idle in this kernel), then the iterations are
 * walked backwards with the gradient messages in shared memory.
 *   gout_dev : [T][B][N*Z] upstream gradients dL/dout_t (zeros where an iteration is unused)
 *   gw_dev, gb_dev : [T][E] fp32, OVERWRITTEN with the batch-summed gradients
 *   workspace_dev  : at least nldpc_backward_workspace_bytes(g, B, T, 0) bytes
 *   have_dump      : non-zero when nldpc_neural_forward_train already filled the workspace for these inputs
 *                    (the forward re-run is then skipped) */
size_t nldpc_backward_workspace_bytes(const nldpc_graph_t *g, int B, int T, int boosted);
int nldpc_neural_backward(const nldpc_graph_t *g, const float *xa_dev, const float *w_dev, const float *b_dev,
                          const float *gout_dev, int B, int T, float *gw_dev, float *gb_dev, void *workspace_dev,
                          size_t workspace_bytes, int have_dump, void *stream);
/* nldpc_neural_forward with every-iteration soft outputs [T][B][N*Z] that also writes the training dump. */
int nldpc_neural_forward_train(const nldpc_graph_t *g, const float *xa_dev, const float *w_dev, const float *b_dev, int B,
                               int T, float *soft_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* Configuration of the Boosted decoder loop body, sharing types already folded by the caller into
 * per-iteration rows (NULL = "no weight of that kind"):
 *   vn_w  [T][N]  multiplies the (compounding) channel input per column  (:325-334, sharing 2/3)
 *   cn_w  [T][E]  check-node weight per edge                               (:431-503, sharing 1-4)
 *   ucn_w [T][E]  weight used instead of cn_w on unsatisfied checks        (:436-488, ucn == cn type) */
typedef struct nldpc_boosted_cfg {
    int32_t decoder_type; /* NLDPC_DEC_* */
    int32_t qbit;         /* decoder_qms_qbit: 6, 5, -5, 4, 3; anything else = no quantisation (:187-214) */
    float llr_lo, llr_hi; /* allowed_llr_range (default -20, 20) */
    int32_t compute_ucn;  /* UCN sharing > 0: evaluate the unsatisfied-check indicator (:339-374) */
    int32_t ucn_mix;      /* UCN type == CN type in {1,2,3}: W = unsatisfied ? ucn_w : cn_w (:436-488) */
    /* Optional state for runs that do not start from the zero state (the module is stateful: self.llr / self.outputs,
     * :94-101; partial target_iter runs read what earlier calls left).  NULL = default. */
    const float *llr_init_dev; /* [B][Z][E] c2v entering the first executed iteration (self.llr[t0]); NULL = zeros */
    const float *xin_init_dev; /* [B][N][Z] compounding channel input entering the first iteration; NULL = xa */
    float *xin_out_dev;        /* [B][N][Z] receives it after the last executed iteration; NULL = not wanted */
    const float *app_init_dev; /* [B][N*Z] previous output used by the UCN indicator of the first executed iteration;
                                  NULL = use the channel input (the curr_iter == 0 rule, :340-341) */
    /* Training: when non-NULL (forward only, soft_mode = ALL) the kernel also spills what nldpc_boosted_backward needs;
     * at least nldpc_backward_workspace_bytes(g, B, T, 1) bytes.  Pass the same buffer to the backward with
     * have_dump = 1 + nldpc_boosted_dump_format(g, cfg, T, cn_w != NULL, vn_w != NULL) of THIS call (1 = slot-major rows of the
     * table-driven kernels, 2 = check-packed records of the specialised kernels).  With a dump buffer the call must ask for
     * soft_mode = NLDPC_OUT_ALL and hard_mode = NLDPC_OUT_NONE. */
    void *train_dump_dev;
    size_t train_dump_bytes;
    /* Optional [T][B][Z][E] fp32: receives self.llr[t + 1] of EVERY executed iteration (the reference stores each one,
     * :512) from the same single launch; NULL = only llr_last_dev (if given) is written. */
    float *llr_all_dev;
    /* Row pitch (in floats) of llr_all_dev AND llr_last_dev: element [..][b][z][e] lives at ((.. * B + b) * Z + z) * llr_pitch + e.
     * 0 = E (the reference's dense [B][Z][E]).  A pitch that is a multiple of 4 (16-byte rows: WiMAX E = 88 as is, BG2 E = 197
     * padded to 200) lets the specialised kernels export the state with 16-byte stores from the variable-lane view of the
     * messages instead of one 4-byte store per lane and edge (BoostedNeuralLDPCDecoder.py:512 is the tensor being filled). */
    int32_t llr_pitch;
} nldpc_boosted_cfg_t;

/* Replaces the loop of BoostedNeuralLDPCDecoder.forward (:320-531) for iterations 0..T-1 from a
 * zero message state (the target_iter=None / range(T) call of train/…py:278 and the validation loop), or for T consecutive
 * iterations from the state given in cfg.  vn_w/cn_w/ucn_w rows are indexed by EXECUTED iteration (0..T-1).
 *   llr_last_dev : optional [B][Z][E] fp32, receives self.llr[t_last + 1] (c2v of the last executed iteration). */
int nldpc_boosted_forward(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const float *xa_dev,
                          const float *vn_w_dev, const float *cn_w_dev, const float *ucn_w_dev,
                          int B, int T, int soft_mode, float *soft_dev, int hard_mode, uint8_t *hard_dev,
                          float *llr_last_dev, void *stream);

/* Backward of nldpc_boosted_forward (MS / QMS decoders, runs from the zero state) w.r.t. the folded weight rows:
 * autograd through BoostedNeuralLDPCDecoder.py:320-531 incl. the straight-through quantisers (:202-214), the
 * +-range clamps and the compounding VN-weight chain (:325-337).
 *   gvn_dev [T][N], gcn_dev [T][E], gucn_dev [T][E]: OVERWRITTEN (pass NULL where the weight kind is absent)
 *   workspace_dev: at least nldpc_backward_workspace_bytes(g, B, T, 1) bytes */
int nldpc_boosted_backward(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const float *xa_dev,
                           const float *vn_w_dev, const float *cn_w_dev, const float *ucn_w_dev, const float *gout_dev,
                           int B, int T, float *gvn_dev, float *gcn_dev, float *gucn_dev, void *workspace_dev,
                           size_t workspace_bytes, int have_dump, void *stream);

/* Which training-dump format nldpc_boosted_forward writes for this configuration: 1 = check-packed records (specialised
 * kernels), 0 = slot-major rows (table-driven kernels).  A caller that keeps the dump for nldpc_boosted_backward passes
 * have_dump = 1 + this value (of the FORWARD call's cfg, state pointers included). */
int nldpc_boosted_dump_format(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, int T, int has_cn_w, int has_vn_w);

/* ---- fused training step (replaces the loop body of train/train_BoostedNeuralLDPCDecoder.py:278-290: model(x) ->
 * LDPCDecoderLoss(BCE)(outputs, y) -> loss.backward()) for the configurations the specialised kernels cover (built-in codes,
 * MS / QMS q=5, CN weights [+ VN weights], no UCN, zero initial state).  Everything else: NLDPC_E_UNSUPPORTED, use the
 * unfused entry points above.
 *   forward : ONE launch runs the T iterations, evaluates L = sum_t coef[t] * mean_i bce_with_logits(out_t[i], y[i])
 *             (LDPCDecoderLoss.py:73-108) and writes dL/dout (clamp mask folded in, scaled by gscale) plus the check-packed
 *             dump into the workspace; *loss_sum_dev receives sum_t coef[t] * sum_i bce (divide by B*N*Z for L).
 *   backward: the sweep over the workspace; gvn [T][N] (iff vn_w), gcn [T][E] are overwritten.
 *   ybits   : labels as bits, [B][ceil(N*Z/8)], bit i of a codeword = (y[i] != 0), LSB first (nldpc_pack_labels). */
size_t nldpc_boosted_train_workspace_bytes(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, int B, int T, int has_cn_w,
                                           int has_vn_w); /* 0 = configuration not covered */
int nldpc_boosted_train_forward(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const float *xa_dev,
                                const float *vn_w_dev, const float *cn_w_dev, int B, int T, const uint8_t *ybits_dev,
                                const float *coef_dev, float gscale, double *loss_sum_dev, void *workspace_dev,
                                size_t workspace_bytes, void *stream);
int nldpc_boosted_train_backward(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const float *xa_dev,
                                 const float *vn_w_dev, const float *cn_w_dev, int B, int T, float *gvn_dev, float *gcn_dev,
                                 void *workspace_dev, size_t workspace_bytes, void *stream);
int nldpc_pack_labels(const float *y_dev, size_t n_codewords, int NZ, uint8_t *bits_dev, void *stream);

/* Fused multi-iteration BCE-with-logits loss and gradient: replaces the loop of LDPCDecoderLoss.forward
 * (LDPCDecoderLoss.py:73-108, BCE branch) over the T iteration outputs.
 *   soft_dev [T][n] logits (n = B*N*Z), y_dev [n] labels, coef_dev [T] = etha^{c_t} / sum_t etha^{c_t}
 *   loss_dev: 1 float, OVERWRITTEN with sum_t coef_t * mean_i bce(soft[t][i], y[i])
 *   gout_dev: [T][n] or NULL, receives dL/dsoft = coef_t * (sigmoid(x) - y) / n   (current device, asynchronous) */
int nldpc_multi_iter_bce(const float *soft_dev, const float *y_dev, const float *coef_dev, int T, size_t n_per_iter,
                         float *loss_dev, float *gout_dev, void *stream);

/* Gradient of the same loss alone, with the upstream gradient folded in (what autograd's backward of
 * LDPCDecoderLoss.forward produces, LDPCDecoderLoss.py:73-108):
 *   gout_dev [T][n] = gscale * coef_t * (sigmoid(soft[t][i]) - y[i]) / n,  gscale_dev: 1 float on the device (dL/dloss) or NULL = 1.
 * Lets the forward call above run without gout (reads only) and keeps no [T][n] tensor alive between forward and backward. */
int nldpc_multi_iter_bce_grad(const float *soft_dev, const float *y_dev, const float *coef_dev, const float *gscale_dev, int T,
                              size_t n_per_iter, float *gout_dev, void *stream);

/* Per-iteration bit / frame error counts: replaces Functions.evaluate_ber_fer (Functions.py:86-102) — the reference's
 * 4 elementwise passes + 2 reductions per iteration become one read of the T soft outputs.
 *   soft_dev    : iteration t at soft_dev + t * iter_stride (in floats), each [B][NZ] (NZ = N*Z), e.g. the [T][B][NZ]
 *                 tensor of nldpc_neural_forward / nldpc_boosted_forward with iter_stride = B*NZ
 *   y_dev       : [B][NZ] fp32 labels (the reference compares ((out < 0) ? 1.0 : 0.0) != y as floats, :90-93 — note the
 *                 predicate is inverted w.r.t. the true decision; kept as is)
 *   counts_dev  : [2][T] uint64, OVERWRITTEN: counts[0][t] = differing positions, counts[1][t] = codewords with >= 1
 * Current device, asynchronous on `stream`. */
int nldpc_count_errors(const float *soft_dev, size_t iter_stride, const float *y_dev, int T, int B, int NZ,
                       uint64_t *counts_dev, void *stream);
/* The same on the packed hard decisions the decode entry points write (bit i%8 of byte i/8 = out[i] < 0;
 * [T][B][ceil(NZ/8)] with iteration t at hard_dev + t * iter_stride_bytes).  y_packed_dev: labels packed the same way,
 * or NULL = the all-zero codeword.  Bits past NZ in a row's last byte are ignored. */
int nldpc_count_errors_packed(const uint8_t *hard_dev, size_t iter_stride_bytes, const uint8_t *y_packed_dev, int T, int B,
                              int NZ, uint64_t *counts_dev, void *stream);

/* Tail of one training step on the flat weight vector, one launch: replaces clip_grad_norm_ + Adam.step + the clamp of
 * _apply_constraints (train/train_BoostedNeuralLDPCDecoder.py:291-294, BoostedNeuralLDPCDecoder.py:153-179).
 *   param_dev, exp_avg_dev, exp_avg_sq_dev : [n] fp32, UPDATED in place
 *   grad_dev   : [n_norm] fp32, n_norm >= n: the gradient norm of clip_grad_norm_(model.parameters()) runs over all n_norm
 *                entries, the optimiser updates the first n (get_trainable_parameters()); all receive the clipped gradient
 *   lr_dev     : optional device scalar read at run time instead of `lr` (a replayed CUDA graph then follows a schedule)
 *   state_dev  : 2 floats: [0] step count so far (incremented by the kernel — the launch is CUDA-graph replayable),
 *                [1] receives the total gradient norm before clipping
 *   grad_scale : multiplies the gradient first (1 / world_size after a SUM all-reduce); max_norm <= 0 disables clipping
 *   lr, beta1, beta2, eps : torch.optim.Adam defaults 1e-3, 0.9, 0.999, 1e-8 (no weight decay, no amsgrad) */
int nldpc_clip_adam_clamp(float *param_dev, float *grad_dev, float *exp_avg_dev, float *exp_avg_sq_dev, float *state_dev, int n,
                          int n_norm, float grad_scale, float max_norm, double lr, const float *lr_dev, double beta1, double beta2,
                          double eps, float clamp_lo, float clamp_hi, void *stream);

/* Host-buffer decode of the Boosted decoder with ONE-BYTE channel LLRs: x = scale * q.  The Boosted pipeline quantises its
 * channel LLRs before the decoder sees them (boosted AWGNPassedDatagen.py:165-166 -> Functions.Cal_MSA_Q, Functions.py:70-83:
 * multiples of 0.5 in +-7.5 for q_bit = 5, i.e. q = 2 x in [-15, 15], scale = 0.5), so the int8 code is lossless and the
 * host -> device transfer (the bound of the end-to-end path) is 4x smaller than with fp32.  Stateless decode from the zero
 * message state (cfg's state / dump pointers must be NULL), T iterations, results as nldpc_boosted_forward with HOST
 * pointers; H2D, int8 -> fp32 expansion, decode and D2H of consecutive chunks overlap on internal streams.  Synchronous.
 *   xq_host [B][N][Z] int8 (pinned or pageable); vn_w_host [T][N] / cn_w_host [T][E] / ucn_w_host [T][E] or NULL */
int nldpc_boosted_decode_host_q8(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const int8_t *xq_host, float scale,
                                 const float *vn_w_host, const float *cn_w_host, const float *ucn_w_host, int B, int T,
                                 int soft_mode, float *soft_host, int hard_mode, uint8_t *hard_host);

#ifdef __cplusplus
}
#endif
#endif /* NLDPC_H_ */
