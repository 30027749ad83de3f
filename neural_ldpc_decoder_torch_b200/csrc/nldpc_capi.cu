// nldpc_capi.cu — extern "C" entry points of libnldpc_b200.so (see include/nldpc.h).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/nldpc.h"
#include "nldpc_common.cuh"
#include "nldpc_spec.cuh"

using namespace nldpc;

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                    \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess) {                                                                          \
            cudaGetLastError();                                                                           \
            return fail((int)_e, std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
        }                                                                                                 \
    } while (0)

// The entry points run on the graph's device and leave the calling thread's current device as they found it (a process
// that drives several GPUs from one thread must not have its default device changed under it).
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
        else if (err == cudaSuccess) prev = -1;      // already there: nothing to restore
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define ON_DEVICE(dev)                                                                                    \
    DeviceGuard _dev_guard(dev);                                                                          \
    if (_dev_guard.err != cudaSuccess) {                                                                  \
        cudaGetLastError();                                                                               \
        return fail((int)_dev_guard.err, std::string("cudaSetDevice: ") + cudaGetErrorString(_dev_guard.err)); \
    }

struct nldpc_graph {
    int device = 0;
    int M = 0, N = 0, Z = 0, E = 0, S = 0;
    GraphDev dev{};
    int *tables = nullptr;       // one device allocation holding every int table
    int cw_per_cta = 0, threads = 0, use_tma = 0;
    size_t smem_bytes = 0;
    int sm_count = 0;
    int spec_id = -1;            // index into the specialised-kernel registry, -1 = generic only
    std::vector<int32_t> bg;     // host copy
    // host-API staging
    cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t events[3] = {nullptr, nullptr, nullptr};
    // grow-only device workspace of the host-buffer API (per stream: xa chunk, soft chunk, hard chunk; + weights)
    void *ws[3][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};
    size_t ws_bytes[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};      // [3] = int8 LLR chunk (nldpc_boosted_decode_host_q8)
    void *ws_wb = nullptr;
    size_t ws_wb_bytes = 0;
};

struct WsLayout { size_t v2c, xin, mask, ucn, scratch, total; };
static WsLayout ws_layout(const nldpc_graph *g, int B, int T, int boosted) {
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    WsLayout l{};
    size_t off = 0;
    // the CN inputs of every iteration: slot-major rows [S][Z] fp32 (table-driven kernels) or check-packed records (specialised
    // kernels: E fp32 values per lane, degree-1 edges included, or padded fp16 records for QMS q=5) — sized for the larger one
    size_t per_cw = (size_t)g->S * g->Z * 4;
    if (g->spec_id >= 0) per_cw = std::max(per_cw, std::max(spec_dump_bytes_per_cw_iter(g->spec_id, 0), spec_dump_bytes_per_cw_iter(g->spec_id, 2)));
    l.v2c = off; off = up(off + (size_t)T * B * per_cw);
    if (boosted) {
        l.xin = off; off = up(off + (size_t)(T + 1) * B * g->N * g->Z * 4);
        l.mask = off; off = up(off + (size_t)T * B * g->N * g->Z);
        l.ucn = off; off = up(off + (size_t)T * B * g->M * g->Z);
    }
    if (g->spec_id >= 0) {   // partial-sum rows of the specialised backward sweep
        l.scratch = off;
        off = up(off + (size_t)g->sm_count * spec_backward_scratch_rows(g->spec_id) * kSpecBwdScratchLanes * 4);
    }
    l.total = off;
    return l;
}
static int boosted_dispatch(const nldpc_graph *g, const DecodeArgs &a, cudaStream_t st);
static bool force_generic();

// Which training-dump format a forward call with these arguments writes (and the backward sweep must be given):
// 1 = check-packed records, written by the specialised forward for the specialised sweep; 0 = slot-major rows (table-driven
// kernels).  Forward and backward decide with the same predicate; callers that keep a dump between the two calls pass the
// format along (have_dump = 1 + format).
static int neural_dump_fmt(const nldpc_graph *g, int T) {
    return (g->spec_id >= 0 && !force_generic() && spec_backward_covers(g->spec_id, 0, T, true, false, false, 0)) ? 1 : 0;
}
static int boosted_dump_fmt(const nldpc_graph *g, const nldpc_boosted_cfg_t *cfg, int T, bool has_cn_w, bool has_vn_w) {
    if (g->spec_id < 0 || force_generic()) return 0;
    const bool qms5 = cfg->decoder_type == NLDPC_DEC_QMS && cfg->qbit == 5, ms = cfg->decoder_type == NLDPC_DEC_MS;
    // what the specialised FORWARD takes (boosted_launch in nldpc_spec_boosted.cuh) ...
    if (!(qms5 || ms) || cfg->compute_ucn || cfg->llr_init_dev || cfg->xin_init_dev || cfg->xin_out_dev || cfg->app_init_dev) return 0;
    // ... and the specialised sweep
    return spec_backward_covers(g->spec_id, qms5 ? 2 : 1, T, has_cn_w, has_vn_w, cfg->ucn_mix != 0, cfg->qbit) ? 1 : 0;
}
extern "C" int nldpc_boosted_dump_format(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, int T, int has_cn_w, int has_vn_w) {
    if (!g || !cfg || T <= 0) return 0;
    return boosted_dump_fmt(g, cfg, T, has_cn_w != 0, has_vn_w != 0);
}

extern "C" const char *nldpc_last_error(void) { return g_err.c_str(); }
extern "C" int nldpc_abi_version(void) { return NLDPC_ABI_VERSION; }

extern "C" int nldpc_graph_create(const int32_t *basegraph, int M, int N, int Z, int device, nldpc_graph_t **out) {
    if (!basegraph || !out || M <= 0 || N <= 0 || Z <= 0) return fail(NLDPC_E_INVALID, "nldpc_graph_create: bad argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        return fail(NLDPC_E_NODEVICE, "nldpc_graph_create: no such CUDA device");
    }
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(NLDPC_E_NODEVICE, "nldpc_graph_create: device is not sm_100 (this library is built for sm_100a only)");
    ON_DEVICE(device);

    // ---- host tables (row-major edges; column lists in ascending check row) ----
    std::vector<int> erow, ecol, eshift, row_ptr(M + 1, 0);
    for (int i = 0; i < M; i++) {
        row_ptr[i] = (int)erow.size();
        for (int j = 0; j < N; j++) {
            const int v = basegraph[(size_t)i * N + j];
            if (v == -1) continue;
            if (v < -1) return fail(NLDPC_E_INVALID, "nldpc_graph_create: basegraph entries must be >= -1");
            erow.push_back(i); ecol.push_back(j); eshift.push_back(v % Z);
        }
    }
    const int E = (int)erow.size();
    row_ptr[M] = E;
    if (E == 0) return fail(NLDPC_E_INVALID, "nldpc_graph_create: graph has no edges");
    std::vector<int> col_deg(N, 0);
    for (int e = 0; e < E; e++) col_deg[ecol[e]]++;
    for (int i = 0; i < M; i++)
        if (row_ptr[i + 1] - row_ptr[i] > kMaxDeg) return fail(NLDPC_E_UNSUPPORTED, "nldpc_graph_create: check degree > 32");
    for (int j = 0; j < N; j++) {
        if (col_deg[j] > kMaxDeg) return fail(NLDPC_E_UNSUPPORTED, "nldpc_graph_create: variable degree > 32");
        // a block no check touches has no edge that would produce its output (the kernels emit marginals edge-side)
        if (col_deg[j] == 0) return fail(NLDPC_E_UNSUPPORTED, "nldpc_graph_create: variable block without any edge (all -1 column)");
    }
    // stored slots: edges of variable blocks with degree >= 2, numbered column-major (block by block)
    std::vector<int> slot(E, -1), vcol_j, vcol_ptr(1, 0), vcol_row;
    int S = 0;
    for (int j = 0; j < N; j++) {
        if (col_deg[j] < 2) continue;
        vcol_j.push_back(j);
        for (int e = 0; e < E; e++)
            if (ecol[e] == j) { slot[e] = S++; vcol_row.push_back(N + slot[e]); }
        vcol_ptr.push_back((int)vcol_row.size());
    }
    std::vector<int> e_row(E), e_col1(E);
    for (int e = 0; e < E; e++) {
        e_row[e] = slot[e] >= 0 ? N + slot[e] : ecol[e];
        e_col1[e] = slot[e] >= 0 ? -1 : ecol[e];
    }
    // slab stride == Z (mod 32): thread (cw, z) then maps to bank (cw*Z + z + const) mod 32 -> conflict-free
    int stride = (N + S) * Z;
    while ((stride & 31) != (Z & 31)) stride++;
    const int NZ = N * Z;
    const int use_tma = ((NZ % 4) == 0) && ((stride % 4) == 0);
    const int hwords = (NZ + 31) / 32;
    // per-CTA budget: aim for 2 CTAs / SM
    const size_t per_cw = (size_t)stride * 4 + (size_t)hwords * 4;
    if (Z > 256 || per_cw + 64 > (size_t)kSmemBudget)
        return fail(NLDPC_E_UNSUPPORTED, "nldpc_graph_create: one codeword's message state does not fit in 227 KB of shared memory");
    const size_t half_budget = (size_t)(kSmemBudget - 2048) / 2;
    int cw = (int)std::min<size_t>((half_budget - 64) / per_cw, (size_t)(256 / Z));
    if (cw < 1) cw = (int)std::min<size_t>(((size_t)kSmemBudget - 64) / per_cw, (size_t)(256 / Z));
    if (cw < 1) cw = 1;
    const int threads = ((cw * Z + 31) / 32) * 32;

    nldpc_graph *g = new (std::nothrow) nldpc_graph();
    if (!g) return fail(NLDPC_E_NOMEM, "nldpc_graph_create: out of host memory");
    g->device = device; g->M = M; g->N = N; g->Z = Z; g->E = E; g->S = S;
    g->bg.assign(basegraph, basegraph + (size_t)M * N);
    g->cw_per_cta = cw; g->threads = threads; g->use_tma = use_tma;
    g->smem_bytes = (size_t)cw * stride * 4 + (((size_t)cw * hwords + 1) & ~(size_t)1) * 4 + 16;
    g->sm_count = prop.multiProcessorCount;

    std::vector<int> all;
    auto push = [&](const std::vector<int> &v) { size_t o = all.size(); all.insert(all.end(), v.begin(), v.end()); return o; };
    const size_t o_vj = push(vcol_j), o_vp = push(vcol_ptr), o_vr = push(vcol_row), o_rp = push(row_ptr), o_er = push(e_row),
                 o_es = push(eshift), o_c1 = push(e_col1), o_cj = push(ecol);
    cudaError_t ce = cudaMalloc(&g->tables, all.size() * sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemcpy(g->tables, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        if (g->tables) cudaFree(g->tables);
        delete g;
        return fail((int)ce, std::string("nldpc_graph_create: table upload failed: ") + cudaGetErrorString(ce));
    }
    GraphDev &d = g->dev;
    d.M = M; d.N = N; d.Z = Z; d.E = E; d.S = S; d.n_vcols = (int)vcol_j.size(); d.slab_stride = stride;
    d.vcol_j = g->tables + o_vj; d.vcol_ptr = g->tables + o_vp; d.vcol_row = g->tables + o_vr;
    d.row_ptr = g->tables + o_rp; d.e_row = g->tables + o_er; d.e_shift = g->tables + o_es; d.e_col1 = g->tables + o_c1; d.e_colj = g->tables + o_cj;

    ce = (cudaError_t)generic_prepare(g->smem_bytes);
    if (ce == cudaSuccess) ce = (cudaError_t)generic_boosted_prepare();
    if (ce == cudaSuccess) ce = (cudaError_t)backward_prepare();
    if (ce != cudaSuccess) {
        cudaGetLastError();
        cudaFree(g->tables);
        delete g;
        return fail((int)ce, std::string("nldpc_graph_create: cudaFuncSetAttribute: ") + cudaGetErrorString(ce));
    }
    g->spec_id = spec_find(g->bg.data(), M, N, Z);
    if (g->spec_id >= 0) {
        int rc = spec_prepare(g->spec_id);
        if (rc != 0) { cudaFree(g->tables); delete g; return fail(rc, "nldpc_graph_create: specialised kernel setup failed"); }
    }
    *out = g;
    return NLDPC_OK;
}

extern "C" void nldpc_graph_destroy(nldpc_graph_t *g) {
    if (!g) return;
    DeviceGuard guard(g->device);
    for (int i = 0; i < 3; i++) {
        if (g->streams[i]) cudaStreamDestroy(g->streams[i]);
        if (g->events[i]) cudaEventDestroy(g->events[i]);
    }
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 4; k++)
            if (g->ws[i][k]) cudaFree(g->ws[i][k]);
    if (g->ws_wb) cudaFree(g->ws_wb);
    if (g->tables) cudaFree(g->tables);
    delete g;
}

extern "C" int nldpc_graph_info(const nldpc_graph_t *g, int32_t info[8]) {
    if (!g || !info) return fail(NLDPC_E_INVALID, "nldpc_graph_info: null argument");
    info[0] = g->M; info[1] = g->N; info[2] = g->Z; info[3] = g->E; info[4] = g->S;
    info[5] = g->spec_id >= 0 ? spec_cw_per_cta(g->spec_id) : g->cw_per_cta;
    info[6] = g->spec_id >= 0 ? spec_threads(g->spec_id) : g->threads;
    info[7] = g->spec_id >= 0 ? 1 : 0;
    return NLDPC_OK;
}

static int check_modes(int soft_mode, const void *soft, int hard_mode, const void *hard) {
    if (soft_mode < 0 || soft_mode > 2 || hard_mode < 0 || hard_mode > 2) return fail(NLDPC_E_INVALID, "bad output mode");
    if ((soft_mode != 0 && !soft) || (hard_mode != 0 && !hard)) return fail(NLDPC_E_INVALID, "output pointer is NULL for a requested output");
    return 0;
}

// NLDPC_FORCE_GENERIC=1 routes the built-in graphs through the table-driven kernel too (tests compare both).
static bool force_generic() {
    const char *e = getenv("NLDPC_FORCE_GENERIC");
    return e && e[0] == '1';
}

// Neural forward dispatch: specialised kernel when the graph has one, else the table-driven kernel
static int neural_dispatch(const nldpc_graph *g, const DecodeArgs &a, cudaStream_t st);

extern "C" int nldpc_neural_forward(const nldpc_graph_t *g, const float *xa_dev, const float *w_dev, const float *b_dev,
                                    int B, int T, int soft_mode, float *soft_dev, int hard_mode, uint8_t *hard_dev,
                                    void *stream) {
    if (!g || B < 0 || T <= 0) return fail(NLDPC_E_INVALID, "nldpc_neural_forward: bad argument");
    if (B == 0) return NLDPC_OK;   // empty batch: nothing to do (pointers may be NULL)
    if (!xa_dev || !w_dev || !b_dev) return fail(NLDPC_E_INVALID, "nldpc_neural_forward: NULL input pointer");
    if (int rc = check_modes(soft_mode, soft_dev, hard_mode, hard_dev)) return rc;
    ON_DEVICE(g->device);
    DecodeArgs a{};
    a.xa = xa_dev; a.w = w_dev; a.b = b_dev; a.B = B; a.T = T;
    a.soft_mode = soft_mode; a.soft = soft_dev; a.hard_mode = hard_mode; a.hard = hard_dev;
    return neural_dispatch(g, a, (cudaStream_t)stream);
}

static int neural_dispatch(const nldpc_graph *g, const DecodeArgs &a, cudaStream_t st) {
    const int B = a.B;
    if (g->spec_id >= 0 && !force_generic()) {
        int rc = spec_launch_neural(g->spec_id, a, g->sm_count, st);
        if (rc > 0) return fail(rc, std::string("nldpc_neural_forward (specialised): ") + cudaGetErrorString((cudaError_t)rc));
        if (rc == 0) return NLDPC_OK;
        // rc < 0: configuration not covered by the specialised kernel -> generic
    }
    const int n_tiles = (B + g->cw_per_cta - 1) / g->cw_per_cta;
    const int ctas_per_sm = std::max(1, (int)((size_t)kSmemBudget / (g->smem_bytes + 1024)));
    const int grid = std::min(n_tiles, g->sm_count * ctas_per_sm);
    CUDA_TRY((cudaError_t)generic_launch_neural(g->dev, a, g->cw_per_cta, g->threads, g->smem_bytes, g->use_tma, grid, st));
    return NLDPC_OK;
}

// ---- host-buffer API: chunked, H2D / decode / D2H of consecutive chunks overlap on 3 streams ----
static int ensure_streams(nldpc_graph *g) {
    for (int i = 0; i < 3; i++) {
        if (!g->streams[i]) CUDA_TRY(cudaStreamCreateWithFlags(&g->streams[i], cudaStreamNonBlocking));
        if (!g->events[i]) CUDA_TRY(cudaEventCreateWithFlags(&g->events[i], cudaEventDisableTiming));
    }
    return 0;
}

static int ws_reserve(void **p, size_t *have, size_t need) {
    if (need <= *have) return 0;
    if (*p) { CUDA_TRY(cudaFree(*p)); *p = nullptr; *have = 0; }
    CUDA_TRY(cudaMalloc(p, need));
    *have = need;
    return 0;
}

namespace nldpc {
int launch_q8_to_f32(const int8_t *in, float *out, size_t n, float scale, int sm_count, cudaStream_t st);
int launch_f16_to_f32(const void *in, float *out, size_t n, int sm_count, cudaStream_t st);
}
// x_format: 0 fp32 (copied straight into the decode kernel's input buffer), NLDPC_LLR_F16 / NLDPC_LLR_Q8: the chunk arrives in the
// narrow format and is expanded on the device (1-2 % of a decode) — the host -> device link is the bound of this path
static int neural_decode_host_impl(const nldpc_graph_t *gc, const void *xa_host_v, int x_format, float scale, const float *w_host,
                                   const float *b_host, int B, int T, int soft_mode, float *soft_host, int hard_mode, uint8_t *hard_host) {
    nldpc_graph *g = const_cast<nldpc_graph *>(gc);
    const float *xa_host = reinterpret_cast<const float *>(xa_host_v);
    const size_t in_bytes = x_format == NLDPC_LLR_F16 ? 2 : (x_format == NLDPC_LLR_Q8 ? 1 : 4);
    if (!g || B < 0 || T <= 0) return fail(NLDPC_E_INVALID, "nldpc_neural_decode_host: bad argument");
    if (B == 0) return NLDPC_OK;
    if (!xa_host || !w_host || !b_host) return fail(NLDPC_E_INVALID, "nldpc_neural_decode_host: NULL input pointer");
    if (int rc = check_modes(soft_mode, soft_host, hard_mode, hard_host)) return rc;
    ON_DEVICE(g->device);
    if (int rc = ensure_streams(g)) return rc;
    const size_t NZ = (size_t)g->N * g->Z, nb = (NZ + 7) / 8, E = (size_t)g->E;
    // chunks: big enough to fill the GPU (~2 waves of resident codewords), small enough that the H2D copy of
    // chunk k+1, the decode of chunk k and the D2H copy of chunk k-1 overlap on three streams
    int chunk_cfg = 4096;     // measured: 2048 / 4096 / 8192 / 16384 / 32768 -> 14.6 / 14.9 / 13.4 / 14.3 / 13.2 M cw/s on a 49 GB/s PCIe link
    if (const char *e = getenv("NLDPC_HOST_CHUNK")) {       // tuning knob (codewords per pipeline chunk)
        const int v = atoi(e);
        if (v >= 256) chunk_cfg = v;
    }
    const int chunk = std::min(B, chunk_cfg);
    // Chunk schedule: full chunks, then a tail that halves down to 512 codewords.  The copy engine is busy from t = 0 whatever
    // the chunk size, so only the END of the pipeline is exposed (decode + D2H of the last chunk, after the last H2D byte has
    // arrived): a small last chunk shortens exactly that.
    std::vector<int> sched;
    for (int rem = B; rem > 0;) {
        int n = std::min(chunk, rem);
        if (rem <= chunk && rem > 512) n = std::max(512, ((rem / 2 + 255) / 256) * 256);
        n = std::min(n, rem);
        sched.push_back(n);
        rem -= n;
    }
    const int nchunk = (int)sched.size();
    const size_t soft_per_cw = soft_mode == NLDPC_OUT_ALL ? (size_t)T * NZ : (soft_mode == NLDPC_OUT_LAST ? NZ : 0);
    const size_t hard_per_cw = hard_mode == NLDPC_OUT_ALL ? (size_t)T * nb : (hard_mode == NLDPC_OUT_LAST ? nb : 0);
    const int nbuf = std::min(nchunk, 3);
    if (int rc = ws_reserve(&g->ws_wb, &g->ws_wb_bytes, 2 * (size_t)T * E * 4)) return rc;
    for (int i = 0; i < nbuf; i++) {
        if (int rc = ws_reserve(&g->ws[i][0], &g->ws_bytes[i][0], (size_t)chunk * NZ * 4)) return rc;
        if (soft_per_cw) if (int rc = ws_reserve(&g->ws[i][1], &g->ws_bytes[i][1], (size_t)chunk * soft_per_cw * 4)) return rc;
        if (hard_per_cw) if (int rc = ws_reserve(&g->ws[i][2], &g->ws_bytes[i][2], (size_t)chunk * hard_per_cw)) return rc;
        if (x_format != 0) if (int rc = ws_reserve(&g->ws[i][3], &g->ws_bytes[i][3], (size_t)chunk * NZ * in_bytes)) return rc;
    }
    float *d_w = (float *)g->ws_wb, *d_b = d_w + (size_t)T * E;
#define HTRY(expr)                                                                                 \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            cudaGetLastError(); cudaDeviceSynchronize();                                           \
            return fail((int)_e, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
        }                                                                                          \
    } while (0)
    HTRY(cudaMemcpyAsync(d_w, w_host, (size_t)T * E * 4, cudaMemcpyHostToDevice, g->streams[0]));
    HTRY(cudaMemcpyAsync(d_b, b_host, (size_t)T * E * 4, cudaMemcpyHostToDevice, g->streams[0]));
    HTRY(cudaEventRecord(g->events[0], g->streams[0]));
    for (int i = 1; i < nbuf; i++) HTRY(cudaStreamWaitEvent(g->streams[i], g->events[0], 0));
    for (int c = 0, b0 = 0; c < nchunk; b0 += sched[c], c++) {
        const int s = c % nbuf;
        cudaStream_t st = g->streams[s];
        float *d_xa = (float *)g->ws[s][0], *d_soft = (float *)g->ws[s][1];
        uint8_t *d_hard = (uint8_t *)g->ws[s][2];
        const int nbw = sched[c];
        if (x_format == 0) {
            HTRY(cudaMemcpyAsync(d_xa, xa_host + (size_t)b0 * NZ, (size_t)nbw * NZ * 4, cudaMemcpyHostToDevice, st));
        } else {
            HTRY(cudaMemcpyAsync(g->ws[s][3], (const char *)xa_host_v + (size_t)b0 * NZ * in_bytes, (size_t)nbw * NZ * in_bytes,
                                 cudaMemcpyHostToDevice, st));
            const int erc = x_format == NLDPC_LLR_F16 ? launch_f16_to_f32(g->ws[s][3], d_xa, (size_t)nbw * NZ, g->sm_count, st)
                                                      : launch_q8_to_f32((const int8_t *)g->ws[s][3], d_xa, (size_t)nbw * NZ, scale, g->sm_count, st);
            if (erc) { cudaDeviceSynchronize(); return fail(erc, std::string("nldpc_neural_decode_host_narrow: ") + cudaGetErrorString((cudaError_t)erc)); }
        }
        int rc = nldpc_neural_forward(g, d_xa, d_w, d_b, nbw, T, soft_mode, d_soft, hard_mode, d_hard, st);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        // the chunk's device layout is [T][nbw][..]; the host layout is [T][B][..]
        if (soft_mode == NLDPC_OUT_ALL) {
            HTRY(cudaMemcpy2DAsync(soft_host + (size_t)b0 * NZ, (size_t)B * NZ * 4, d_soft, (size_t)nbw * NZ * 4,
                                   (size_t)nbw * NZ * 4, T, cudaMemcpyDeviceToHost, st));
        } else if (soft_mode == NLDPC_OUT_LAST) {
            HTRY(cudaMemcpyAsync(soft_host + (size_t)b0 * NZ, d_soft, (size_t)nbw * NZ * 4, cudaMemcpyDeviceToHost, st));
        }
        if (hard_mode == NLDPC_OUT_ALL) {
            HTRY(cudaMemcpy2DAsync(hard_host + (size_t)b0 * nb, (size_t)B * nb, d_hard, (size_t)nbw * nb, (size_t)nbw * nb, T,
                                   cudaMemcpyDeviceToHost, st));
        } else if (hard_mode == NLDPC_OUT_LAST) {
            HTRY(cudaMemcpyAsync(hard_host + (size_t)b0 * nb, d_hard, (size_t)nbw * nb, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int i = 0; i < nbuf; i++) HTRY(cudaStreamSynchronize(g->streams[i]));
#undef HTRY
    return NLDPC_OK;
}

extern "C" int nldpc_neural_decode_host(const nldpc_graph_t *g, const float *xa_host, const float *w_host, const float *b_host,
                                        int B, int T, int soft_mode, float *soft_host, int hard_mode, uint8_t *hard_host) {
    return neural_decode_host_impl(g, xa_host, 0, 1.0f, w_host, b_host, B, T, soft_mode, soft_host, hard_mode, hard_host);
}

extern "C" int nldpc_neural_decode_host_narrow(const nldpc_graph_t *g, const void *x_host, int x_format, float scale, const float *w_host,
                                               const float *b_host, int B, int T, int soft_mode, float *soft_host, int hard_mode,
                                               uint8_t *hard_host) {
    if (x_format != NLDPC_LLR_F16 && x_format != NLDPC_LLR_Q8) return fail(NLDPC_E_INVALID, "nldpc_neural_decode_host_narrow: x_format must be NLDPC_LLR_F16 or NLDPC_LLR_Q8");
    return neural_decode_host_impl(g, x_host, x_format, scale, w_host, b_host, B, T, soft_mode, soft_host, hard_mode, hard_host);
}

// ---- backward ------------------------------------------------------------------------------------------------
extern "C" size_t nldpc_backward_workspace_bytes(const nldpc_graph_t *g, int B, int T, int boosted) {
    if (!g || B <= 0 || T <= 0) return 0;
    return ws_layout(g, B, T, boosted).total;
}

extern "C" int nldpc_neural_backward(const nldpc_graph_t *g, const float *xa_dev, const float *w_dev, const float *b_dev,
                                     const float *gout_dev, int B, int T, float *gw_dev, float *gb_dev, void *workspace_dev,
                                     size_t workspace_bytes, int have_dump, void *stream) {
    if (!g || B < 0 || T <= 0 || !gw_dev || !gb_dev) return fail(NLDPC_E_INVALID, "nldpc_neural_backward: bad argument");
    ON_DEVICE(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(gw_dev, 0, (size_t)T * g->E * 4, st));
    CUDA_TRY(cudaMemsetAsync(gb_dev, 0, (size_t)T * g->E * 4, st));
    if (B == 0) return NLDPC_OK;
    if (!xa_dev || !w_dev || !b_dev || !gout_dev) return fail(NLDPC_E_INVALID, "nldpc_neural_backward: NULL input pointer");
    const WsLayout l = ws_layout(g, B, T, 0);
    if (!workspace_dev || workspace_bytes < l.total)
        return fail(NLDPC_E_INVALID, "nldpc_neural_backward: workspace too small (see nldpc_backward_workspace_bytes)");
    // (A) forward re-run in training-dump mode, unless the forward pass already wrote the dump (nldpc_neural_forward_train)
    DecodeArgs a{};
    a.xa = xa_dev; a.w = w_dev; a.b = b_dev; a.B = B; a.T = T; a.wb_off = -1;
    a.hist_v2c = reinterpret_cast<float *>((char *)workspace_dev + l.v2c);
    a.hist_fmt = neural_dump_fmt(g, T);
    if (!have_dump)
        if (int rc = neural_dispatch(g, a, st)) return rc;
    // (B) backward sweep
    BwdArgs ba{};
    ba.xa = xa_dev; ba.w = w_dev; ba.b = b_dev; ba.gout = gout_dev; ba.hist_v2c = a.hist_v2c; ba.hist_fmt = a.hist_fmt;
    ba.gw = gw_dev; ba.gb = gb_dev; ba.B = B; ba.T = T; ba.mode = 0;
    if (g->spec_id >= 0 && !force_generic()) {
        ba.scratch = reinterpret_cast<float *>((char *)workspace_dev + l.scratch);
        const int src = spec_launch_backward(g->spec_id, ba, g->sm_count, st);
        if (src > 0) return fail(src, std::string("nldpc_neural_backward (specialised): ") + cudaGetErrorString((cudaError_t)src));
        if (src == 0) return NLDPC_OK;
    }
    const int rc = backward_launch(g->dev, ba, g->sm_count, st);
    if (rc == -2) return fail(NLDPC_E_UNSUPPORTED, "nldpc_neural_backward: graph does not fit on chip");
    if (rc != 0) return fail(rc, std::string("nldpc_neural_backward: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

extern "C" int nldpc_neural_forward_train(const nldpc_graph_t *g, const float *xa_dev, const float *w_dev, const float *b_dev, int B,
                                          int T, float *soft_dev, void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!g || B < 0 || T <= 0) return fail(NLDPC_E_INVALID, "nldpc_neural_forward_train: bad argument");
    if (B == 0) return NLDPC_OK;
    if (!xa_dev || !w_dev || !b_dev || !soft_dev) return fail(NLDPC_E_INVALID, "nldpc_neural_forward_train: NULL pointer");
    const WsLayout l = ws_layout(g, B, T, 0);
    if (!workspace_dev || workspace_bytes < l.total)
        return fail(NLDPC_E_INVALID, "nldpc_neural_forward_train: workspace too small (see nldpc_backward_workspace_bytes)");
    ON_DEVICE(g->device);
    DecodeArgs a{};
    a.xa = xa_dev; a.w = w_dev; a.b = b_dev; a.B = B; a.T = T; a.soft_mode = NLDPC_OUT_ALL; a.soft = soft_dev;
    a.hist_v2c = reinterpret_cast<float *>((char *)workspace_dev + l.v2c);
    a.hist_fmt = neural_dump_fmt(g, T);
    return neural_dispatch(g, a, (cudaStream_t)stream);
}

extern "C" int nldpc_boosted_backward(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const float *xa_dev,
                                      const float *vn_w_dev, const float *cn_w_dev, const float *ucn_w_dev, const float *gout_dev,
                                      int B, int T, float *gvn_dev, float *gcn_dev, float *gucn_dev, void *workspace_dev,
                                      size_t workspace_bytes, int have_dump, void *stream) {
    if (!g || !cfg || B < 0 || T <= 0) return fail(NLDPC_E_INVALID, "nldpc_boosted_backward: bad argument");
    if (cfg->decoder_type != NLDPC_DEC_MS && cfg->decoder_type != NLDPC_DEC_QMS)
        return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_backward: only the MS and QMS decoders have a backward");
    if (cfg->llr_init_dev || cfg->xin_init_dev || cfg->app_init_dev)
        return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_backward: runs that continue from stored state are forward-only");
    if ((vn_w_dev && !gvn_dev) || (cn_w_dev && !gcn_dev) || (cfg->ucn_mix && !gucn_dev))
        return fail(NLDPC_E_INVALID, "nldpc_boosted_backward: missing gradient output");
    ON_DEVICE(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (gvn_dev) CUDA_TRY(cudaMemsetAsync(gvn_dev, 0, (size_t)T * g->N * 4, st));
    if (gcn_dev) CUDA_TRY(cudaMemsetAsync(gcn_dev, 0, (size_t)T * g->E * 4, st));
    if (gucn_dev) CUDA_TRY(cudaMemsetAsync(gucn_dev, 0, (size_t)T * g->E * 4, st));
    if (B == 0) return NLDPC_OK;
    if (!xa_dev || !gout_dev) return fail(NLDPC_E_INVALID, "nldpc_boosted_backward: NULL input pointer");
    const WsLayout l = ws_layout(g, B, T, 1);
    if (!workspace_dev || workspace_bytes < l.total)
        return fail(NLDPC_E_INVALID, "nldpc_boosted_backward: workspace too small (see nldpc_backward_workspace_bytes)");
    char *ws = (char *)workspace_dev;
    DecodeArgs a{};
    a.xa = xa_dev; a.w = cn_w_dev; a.b = ucn_w_dev; a.vn_w = vn_w_dev; a.B = B; a.T = T; a.wb_off = -1;
    a.decoder_type = cfg->decoder_type; a.qbit = cfg->qbit; a.compute_ucn = cfg->compute_ucn; a.ucn_mix = cfg->ucn_mix;
    a.llr_lo = cfg->llr_lo; a.llr_hi = cfg->llr_hi;
    a.hist_v2c = reinterpret_cast<float *>(ws + l.v2c); a.hist_xin = reinterpret_cast<float *>(ws + l.xin);
    a.hist_mask = reinterpret_cast<uint8_t *>(ws + l.mask); a.hist_ucn = reinterpret_cast<uint8_t *>(ws + l.ucn);
    // have_dump: 0 = none (the forward is re-run here), 1 = slot-major dump, 2 = check-packed dump (1 + nldpc_boosted_dump_format
    // of the forward call that wrote it)
    a.hist_fmt = have_dump ? (have_dump == 2 ? 1 : 0) : boosted_dump_fmt(g, cfg, T, cn_w_dev != nullptr, vn_w_dev != nullptr);
    int rc = 0;
    if (!have_dump)
        if ((rc = boosted_dispatch(g, a, st))) return rc;
    BwdArgs ba{};
    ba.hist_fmt = a.hist_fmt;
    ba.xa = xa_dev; ba.w = cn_w_dev; ba.b = cfg->ucn_mix ? ucn_w_dev : nullptr; ba.vn_w = vn_w_dev; ba.gout = gout_dev;
    ba.hist_v2c = a.hist_v2c; ba.hist_xin = a.hist_xin; ba.hist_mask = a.hist_mask; ba.hist_ucn = cfg->ucn_mix ? a.hist_ucn : nullptr;
    ba.gw = gcn_dev; ba.gb = cfg->ucn_mix ? gucn_dev : nullptr; ba.gvn = vn_w_dev ? gvn_dev : nullptr;
    ba.B = B; ba.T = T; ba.mode = cfg->decoder_type == NLDPC_DEC_QMS ? 2 : 1; ba.qbit = cfg->qbit; ba.lo = cfg->llr_lo; ba.hi = cfg->llr_hi;
    ba.ucn_mix = cfg->ucn_mix;
    if (g->spec_id >= 0 && !force_generic()) {
        ba.scratch = reinterpret_cast<float *>(ws + l.scratch);
        const int src = spec_launch_backward(g->spec_id, ba, g->sm_count, st);
        if (src > 0) return fail(src, std::string("nldpc_boosted_backward (specialised): ") + cudaGetErrorString((cudaError_t)src));
        if (src == 0) return NLDPC_OK;
    }
    rc = backward_launch(g->dev, ba, g->sm_count, st);
    if (rc == -2) return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_backward: graph does not fit on chip");
    if (rc != 0) return fail(rc, std::string("nldpc_boosted_backward: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

extern "C" int nldpc_boosted_forward(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const float *xa_dev,
                                     const float *vn_w_dev, const float *cn_w_dev, const float *ucn_w_dev, int B, int T,
                                     int soft_mode, float *soft_dev, int hard_mode, uint8_t *hard_dev, float *llr_last_dev,
                                     void *stream) {
    if (!g || !cfg || B < 0 || T <= 0) return fail(NLDPC_E_INVALID, "nldpc_boosted_forward: bad argument");
    if (B == 0) return NLDPC_OK;
    if (!xa_dev) return fail(NLDPC_E_INVALID, "nldpc_boosted_forward: NULL input pointer");
    if (cfg->decoder_type < 0 || cfg->decoder_type > 2) return fail(NLDPC_E_INVALID, "nldpc_boosted_forward: bad decoder_type");
    if (!(cfg->llr_lo <= cfg->llr_hi)) return fail(NLDPC_E_INVALID, "nldpc_boosted_forward: bad llr range");
    if (cfg->ucn_mix && (!cn_w_dev || !ucn_w_dev || !cfg->compute_ucn))
        return fail(NLDPC_E_INVALID, "nldpc_boosted_forward: ucn_mix needs cn_w, ucn_w and compute_ucn");
    if (int rc = check_modes(soft_mode, soft_dev, hard_mode, hard_dev)) return rc;
    ON_DEVICE(g->device);
    DecodeArgs a{};
    a.xa = xa_dev; a.w = cn_w_dev; a.b = ucn_w_dev; a.vn_w = vn_w_dev; a.B = B; a.T = T;
    a.soft_mode = soft_mode; a.soft = soft_dev; a.hard_mode = hard_mode; a.hard = hard_dev; a.llr_last = llr_last_dev;
    a.llr_all = cfg->llr_all_dev;
    if (cfg->llr_pitch != 0 && cfg->llr_pitch < g->E) return fail(NLDPC_E_INVALID, "nldpc_boosted_forward: llr_pitch smaller than E");
    a.llr_pitch = cfg->llr_pitch > 0 ? cfg->llr_pitch : g->E;
    a.wb_off = -1;
    a.decoder_type = cfg->decoder_type; a.qbit = cfg->qbit; a.compute_ucn = cfg->compute_ucn; a.ucn_mix = cfg->ucn_mix;
    a.llr_lo = cfg->llr_lo; a.llr_hi = cfg->llr_hi;
    a.llr_init = cfg->llr_init_dev; a.xin_init = cfg->xin_init_dev; a.xin_out = cfg->xin_out_dev; a.app_init = cfg->app_init_dev;
    if (cfg->train_dump_dev) {   // training mode: spill what the backward kernel needs (nldpc_backward_workspace_bytes(.., 1))
        if (soft_mode != NLDPC_OUT_ALL) return fail(NLDPC_E_INVALID, "nldpc_boosted_forward: the training dump needs soft_mode = ALL");
        const WsLayout l = ws_layout(g, B, T, 1);
        if (cfg->train_dump_bytes < l.total) return fail(NLDPC_E_INVALID, "nldpc_boosted_forward: training dump buffer too small");
        char *ws = (char *)cfg->train_dump_dev;
        a.hist_v2c = reinterpret_cast<float *>(ws + l.v2c); a.hist_xin = reinterpret_cast<float *>(ws + l.xin);
        a.hist_mask = reinterpret_cast<uint8_t *>(ws + l.mask); a.hist_ucn = reinterpret_cast<uint8_t *>(ws + l.ucn);
        a.hist_fmt = boosted_dump_fmt(g, cfg, T, cn_w_dev != nullptr, vn_w_dev != nullptr);
    }
    return boosted_dispatch(g, a, (cudaStream_t)stream);
}

// ---- fused training step pieces (training.FusedTrainer) ---------------------------------------------------------------
// forward + multi-iteration BCE + dL/dout in ONE launch (no [T][B][N*Z] output, loss or mask tensor crosses HBM twice), the
// sweep reads the workspace.  Covered: what the specialised kernels cover (built-in codes, MS / QMS q=5, CN weights, no UCN,
// zero initial state); everything else returns NLDPC_E_UNSUPPORTED and the caller keeps the unfused path.
struct TrainLayout { size_t dump, xin, gout, scratch, loss, total; };
static TrainLayout train_layout(const nldpc_graph *g, int mode, int B, int T, bool has_vn_w) {
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    TrainLayout l{};
    size_t off = 0;
    l.dump = off; off = up(off + (size_t)T * B * spec_dump_bytes_per_cw_iter(g->spec_id, mode));
    l.xin = off; if (has_vn_w) off = up(off + (size_t)T * B * g->N * g->Z * 4);       // rows 1..T-1 are used
    l.gout = off; off = up(off + (size_t)T * B * g->N * g->Z * 4);
    l.scratch = off; off = up(off + (size_t)g->sm_count * spec_backward_scratch_rows(g->spec_id) * kSpecBwdScratchLanes * 4);
    l.loss = off; off = up(off + 8);
    l.total = off;
    return l;
}
static int train_mode(const nldpc_graph *g, const nldpc_boosted_cfg_t *cfg, int T, bool has_cn_w, bool has_vn_w) {
    if (!boosted_dump_fmt(g, cfg, T, has_cn_w, has_vn_w)) return 0;
    return cfg->decoder_type == NLDPC_DEC_QMS ? 2 : 1;
}

extern "C" size_t nldpc_boosted_train_workspace_bytes(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, int B, int T,
                                                      int has_cn_w, int has_vn_w) {
    if (!g || !cfg || B <= 0 || T <= 0) return 0;
    const int mode = train_mode(g, cfg, T, has_cn_w != 0, has_vn_w != 0);
    return mode ? train_layout(g, mode, B, T, has_vn_w != 0).total : 0;
}

extern "C" int nldpc_boosted_train_forward(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const float *xa_dev,
                                           const float *vn_w_dev, const float *cn_w_dev, int B, int T, const uint8_t *ybits_dev,
                                           const float *coef_dev, float gscale, double *loss_sum_dev, void *workspace_dev,
                                           size_t workspace_bytes, void *stream) {
    if (!g || !cfg || B <= 0 || T <= 0 || !xa_dev || !ybits_dev || !coef_dev || !loss_sum_dev)
        return fail(NLDPC_E_INVALID, "nldpc_boosted_train_forward: bad argument");
    const int mode = train_mode(g, cfg, T, cn_w_dev != nullptr, vn_w_dev != nullptr);
    if (!mode) return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_train_forward: configuration not covered by the fused training kernels");
    if (!(cfg->llr_lo <= cfg->llr_hi)) return fail(NLDPC_E_INVALID, "nldpc_boosted_train_forward: bad llr range");
    const TrainLayout l = train_layout(g, mode, B, T, vn_w_dev != nullptr);
    if (!workspace_dev || workspace_bytes < l.total)
        return fail(NLDPC_E_INVALID, "nldpc_boosted_train_forward: workspace too small (see nldpc_boosted_train_workspace_bytes)");
    ON_DEVICE(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace_dev;
    CUDA_TRY(cudaMemsetAsync(loss_sum_dev, 0, sizeof(double), st));
    DecodeArgs a{};
    a.xa = xa_dev; a.w = cn_w_dev; a.vn_w = vn_w_dev; a.B = B; a.T = T; a.wb_off = -1;
    a.soft_mode = NLDPC_OUT_ALL; a.soft = reinterpret_cast<float *>(ws + l.gout);
    a.decoder_type = cfg->decoder_type; a.qbit = cfg->qbit; a.llr_lo = cfg->llr_lo; a.llr_hi = cfg->llr_hi;
    a.hist_v2c = reinterpret_cast<float *>(ws + l.dump); a.hist_xin = reinterpret_cast<float *>(ws + l.xin); a.hist_fmt = 1;
    a.ybits = ybits_dev; a.coef = coef_dev; a.ginv = (float)((double)gscale / ((double)B * g->N * g->Z)); a.loss_acc = loss_sum_dev;
    const int src = spec_launch_boosted(g->spec_id, a, g->sm_count, st);
    if (src > 0) return fail(src, std::string("nldpc_boosted_train_forward: ") + cudaGetErrorString((cudaError_t)src));
    if (src < 0) return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_train_forward: the specialised forward declined the launch (constant arena full)");
    return NLDPC_OK;
}

extern "C" int nldpc_boosted_train_backward(const nldpc_graph_t *g, const nldpc_boosted_cfg_t *cfg, const float *xa_dev,
                                            const float *vn_w_dev, const float *cn_w_dev, int B, int T, float *gvn_dev,
                                            float *gcn_dev, void *workspace_dev, size_t workspace_bytes, void *stream) {
    if (!g || !cfg || B <= 0 || T <= 0 || !xa_dev || !gcn_dev || (vn_w_dev && !gvn_dev))
        return fail(NLDPC_E_INVALID, "nldpc_boosted_train_backward: bad argument");
    const int mode = train_mode(g, cfg, T, cn_w_dev != nullptr, vn_w_dev != nullptr);
    if (!mode) return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_train_backward: configuration not covered by the fused training kernels");
    const TrainLayout l = train_layout(g, mode, B, T, vn_w_dev != nullptr);
    if (!workspace_dev || workspace_bytes < l.total)
        return fail(NLDPC_E_INVALID, "nldpc_boosted_train_backward: workspace too small (see nldpc_boosted_train_workspace_bytes)");
    ON_DEVICE(g->device);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace_dev;
    if (gvn_dev) CUDA_TRY(cudaMemsetAsync(gvn_dev, 0, (size_t)T * g->N * 4, st));
    CUDA_TRY(cudaMemsetAsync(gcn_dev, 0, (size_t)T * g->E * 4, st));
    BwdArgs ba{};
    ba.xa = xa_dev; ba.w = cn_w_dev; ba.vn_w = vn_w_dev; ba.gout = reinterpret_cast<const float *>(ws + l.gout);
    ba.hist_fmt = 1; ba.hist_v2c = reinterpret_cast<const float *>(ws + l.dump); ba.hist_xin = reinterpret_cast<const float *>(ws + l.xin);
    ba.hist_mask = nullptr;      // the forward folded the clamp mask into gout
    ba.gw = gcn_dev; ba.gvn = vn_w_dev ? gvn_dev : nullptr;
    ba.B = B; ba.T = T; ba.mode = mode; ba.qbit = cfg->qbit; ba.lo = cfg->llr_lo; ba.hi = cfg->llr_hi;
    ba.scratch = reinterpret_cast<float *>(ws + l.scratch);
    const int src = spec_launch_backward(g->spec_id, ba, g->sm_count, st);
    if (src > 0) return fail(src, std::string("nldpc_boosted_train_backward: ") + cudaGetErrorString((cudaError_t)src));
    if (src < 0) return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_train_backward: the specialised sweep declined the launch");
    return NLDPC_OK;
}

namespace nldpc {
int launch_pack_labels(const float *y, size_t n_cw, int NZ, uint8_t *bits, cudaStream_t st);
}
extern "C" int nldpc_pack_labels(const float *y_dev, size_t n_codewords, int NZ, uint8_t *bits_dev, void *stream) {
    if (NZ <= 0 || ((!y_dev || !bits_dev) && n_codewords > 0)) return fail(NLDPC_E_INVALID, "nldpc_pack_labels: bad argument");
    if (n_codewords == 0) return NLDPC_OK;
    const int rc = launch_pack_labels(y_dev, n_codewords, NZ, bits_dev, (cudaStream_t)stream);
    if (rc != 0) return fail(rc, std::string("nldpc_pack_labels: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

static int boosted_dispatch(const nldpc_graph *g, const DecodeArgs &a, cudaStream_t st) {
    if (g->spec_id >= 0 && !force_generic()) {
        const int src = spec_launch_boosted(g->spec_id, a, g->sm_count, st);
        if (src > 0) return fail(src, std::string("nldpc_boosted_forward (specialised): ") + cudaGetErrorString((cudaError_t)src));
        if (src == 0) return NLDPC_OK;
        // src < 0: configuration not covered by the specialised kernels (SP, other q-bit grids, UCN, stateful runs)
    }
    if (a.hist_v2c && a.hist_fmt == 1)      // (only when the constant arena cannot take the weights: T * E > 7680)
        return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_forward: the specialised forward declined a launch whose training dump the specialised sweep expects");
    const int rc = generic_launch_boosted(g->dev, a, g->sm_count, st);
    if (rc == -2) return fail(NLDPC_E_UNSUPPORTED, "nldpc_boosted_forward: one codeword's boosted state does not fit in shared memory");
    if (rc != 0) return fail(rc, std::string("nldpc_boosted_forward: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

namespace nldpc {
int launch_multi_iter_bce(const float *soft, const float *y, const float *coef, const float *gscale, int T, size_t n, float *loss,
                          float *gout, int sm_count, cudaStream_t st);
}

extern "C" int nldpc_multi_iter_bce(const float *soft_dev, const float *y_dev, const float *coef_dev, int T, size_t n_per_iter,
                                    float *loss_dev, float *gout_dev, void *stream) {
    if (!soft_dev || !y_dev || !coef_dev || !loss_dev || T <= 0) return fail(NLDPC_E_INVALID, "nldpc_multi_iter_bce: bad argument");
    int dev = 0, sms = 148;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int rc = launch_multi_iter_bce(soft_dev, y_dev, coef_dev, nullptr, T, n_per_iter, loss_dev, gout_dev, sms, (cudaStream_t)stream);
    if (rc != 0) return fail(rc, std::string("nldpc_multi_iter_bce: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

extern "C" int nldpc_multi_iter_bce_grad(const float *soft_dev, const float *y_dev, const float *coef_dev, const float *gscale_dev, int T,
                                         size_t n_per_iter, float *gout_dev, void *stream) {
    if (!soft_dev || !y_dev || !coef_dev || !gout_dev || T <= 0) return fail(NLDPC_E_INVALID, "nldpc_multi_iter_bce_grad: bad argument");
    int dev = 0, sms = 148;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int rc = launch_multi_iter_bce(soft_dev, y_dev, coef_dev, gscale_dev, T, n_per_iter, nullptr, gout_dev, sms, (cudaStream_t)stream);
    if (rc != 0) return fail(rc, std::string("nldpc_multi_iter_bce_grad: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

namespace nldpc {
int launch_count_errors(const float *soft, size_t iter_stride, const float *y, int T, int B, int NZ, unsigned long long *counts,
                        int sm_count, cudaStream_t st);
int launch_count_errors_packed(const uint8_t *hard, size_t iter_stride, const uint8_t *yp, int T, int B, int NZ,
                               unsigned long long *counts, int sm_count, cudaStream_t st);
}

extern "C" int nldpc_count_errors(const float *soft_dev, size_t iter_stride, const float *y_dev, int T, int B, int NZ,
                                  uint64_t *counts_dev, void *stream) {
    if (!counts_dev || T <= 0 || B < 0 || NZ < 0 || ((!soft_dev || !y_dev) && B > 0 && NZ > 0))
        return fail(NLDPC_E_INVALID, "nldpc_count_errors: bad argument");
    int dev = 0, sms = 148;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int rc = launch_count_errors(soft_dev, iter_stride, y_dev, T, B, NZ, (unsigned long long *)counts_dev, sms, (cudaStream_t)stream);
    if (rc != 0) return fail(rc, std::string("nldpc_count_errors: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

extern "C" int nldpc_count_errors_packed(const uint8_t *hard_dev, size_t iter_stride_bytes, const uint8_t *y_packed_dev, int T, int B,
                                         int NZ, uint64_t *counts_dev, void *stream) {
    if (!counts_dev || T <= 0 || B < 0 || NZ < 0 || (!hard_dev && B > 0 && NZ > 0))
        return fail(NLDPC_E_INVALID, "nldpc_count_errors_packed: bad argument");
    int dev = 0, sms = 148;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int rc = launch_count_errors_packed(hard_dev, iter_stride_bytes, y_packed_dev, T, B, NZ, (unsigned long long *)counts_dev, sms,
                                              (cudaStream_t)stream);
    if (rc != 0) return fail(rc, std::string("nldpc_count_errors_packed: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

namespace nldpc {
int launch_clip_adam_clamp(float *p, float *g, float *m, float *v, float *state, int n, int n_norm, float grad_scale, float max_norm,
                           double lr, const float *lr_dev, double beta1, double beta2, double eps, float lo, float hi, cudaStream_t st);
}

extern "C" int nldpc_clip_adam_clamp(float *param_dev, float *grad_dev, float *exp_avg_dev, float *exp_avg_sq_dev, float *state_dev, int n,
                                     int n_norm, float grad_scale, float max_norm, double lr, const float *lr_dev, double beta1,
                                     double beta2, double eps, float clamp_lo, float clamp_hi, void *stream) {
    if (n < 0 || n_norm < n || (n_norm > 0 && !grad_dev) || (n > 0 && (!param_dev || !exp_avg_dev || !exp_avg_sq_dev)) || !state_dev ||
        !(clamp_lo <= clamp_hi))
        return fail(NLDPC_E_INVALID, "nldpc_clip_adam_clamp: bad argument");
    const int rc = launch_clip_adam_clamp(param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, state_dev, n, n_norm, grad_scale, max_norm, lr,
                                          lr_dev, beta1, beta2, eps, clamp_lo, clamp_hi, (cudaStream_t)stream);
    if (rc != 0) return fail(rc, std::string("nldpc_clip_adam_clamp: ") + cudaGetErrorString((cudaError_t)rc));
    return NLDPC_OK;
}

namespace nldpc {
int launch_q8_to_f32(const int8_t *in, float *out, size_t n, float scale, int sm_count, cudaStream_t st);
}

extern "C" int nldpc_boosted_decode_host_q8(const nldpc_graph_t *gc, const nldpc_boosted_cfg_t *cfg, const int8_t *xq_host, float scale,
                                            const float *vn_w_host, const float *cn_w_host, const float *ucn_w_host, int B, int T,
                                            int soft_mode, float *soft_host, int hard_mode, uint8_t *hard_host) {
    nldpc_graph *g = const_cast<nldpc_graph *>(gc);
    if (!g || !cfg || B < 0 || T <= 0) return fail(NLDPC_E_INVALID, "nldpc_boosted_decode_host_q8: bad argument");
    if (B == 0) return NLDPC_OK;
    if (!xq_host) return fail(NLDPC_E_INVALID, "nldpc_boosted_decode_host_q8: NULL input pointer");
    if (cfg->llr_init_dev || cfg->xin_init_dev || cfg->xin_out_dev || cfg->app_init_dev || cfg->train_dump_dev)
        return fail(NLDPC_E_INVALID, "nldpc_boosted_decode_host_q8: stateless decode only (state / dump pointers must be NULL)");
    if (int rc = check_modes(soft_mode, soft_host, hard_mode, hard_host)) return rc;
    ON_DEVICE(g->device);
    if (int rc = ensure_streams(g)) return rc;
    const size_t NZ = (size_t)g->N * g->Z, nb = (NZ + 7) / 8, E = (size_t)g->E, N = (size_t)g->N;
    int chunk_cfg = 8192;      // the decode, not the copy, is the long stage here: larger chunks than the fp32 Neural path
    if (const char *e = getenv("NLDPC_HOST_CHUNK")) {
        const int v = atoi(e);
        if (v >= 256) chunk_cfg = v;
    }
    const int chunk = std::min(B, chunk_cfg);
    std::vector<int> sched;                   // full chunks, then a tail that halves down to 512 codewords (see nldpc_neural_decode_host)
    for (int rem = B; rem > 0;) {
        int n = std::min(chunk, rem);
        if (rem <= chunk && rem > 512) n = std::max(512, ((rem / 2 + 255) / 256) * 256);
        n = std::min(n, rem);
        sched.push_back(n);
        rem -= n;
    }
    const int nchunk = (int)sched.size();
    const size_t soft_per_cw = soft_mode == NLDPC_OUT_ALL ? (size_t)T * NZ : (soft_mode == NLDPC_OUT_LAST ? NZ : 0);
    const size_t hard_per_cw = hard_mode == NLDPC_OUT_ALL ? (size_t)T * nb : (hard_mode == NLDPC_OUT_LAST ? nb : 0);
    const int nbuf = std::min(nchunk, 3);
    const size_t wbytes = ((size_t)T * N + 2 * (size_t)T * E) * 4;
    if (int rc = ws_reserve(&g->ws_wb, &g->ws_wb_bytes, wbytes)) return rc;
    for (int i = 0; i < nbuf; i++) {
        if (int rc = ws_reserve(&g->ws[i][0], &g->ws_bytes[i][0], (size_t)chunk * NZ * 4)) return rc;
        if (soft_per_cw) if (int rc = ws_reserve(&g->ws[i][1], &g->ws_bytes[i][1], (size_t)chunk * soft_per_cw * 4)) return rc;
        if (hard_per_cw) if (int rc = ws_reserve(&g->ws[i][2], &g->ws_bytes[i][2], (size_t)chunk * hard_per_cw)) return rc;
        if (int rc = ws_reserve(&g->ws[i][3], &g->ws_bytes[i][3], (size_t)chunk * NZ)) return rc;
    }
    float *d_vn = (float *)g->ws_wb, *d_cn = d_vn + (size_t)T * N, *d_ucn = d_cn + (size_t)T * E;
#define HTRY(expr)                                                                                 \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            cudaGetLastError(); cudaDeviceSynchronize();                                           \
            return fail((int)_e, std::string(#expr) + ": " + cudaGetErrorString(_e));              \
        }                                                                                          \
    } while (0)
    if (vn_w_host) HTRY(cudaMemcpyAsync(d_vn, vn_w_host, (size_t)T * N * 4, cudaMemcpyHostToDevice, g->streams[0]));
    if (cn_w_host) HTRY(cudaMemcpyAsync(d_cn, cn_w_host, (size_t)T * E * 4, cudaMemcpyHostToDevice, g->streams[0]));
    if (ucn_w_host) HTRY(cudaMemcpyAsync(d_ucn, ucn_w_host, (size_t)T * E * 4, cudaMemcpyHostToDevice, g->streams[0]));
    HTRY(cudaEventRecord(g->events[0], g->streams[0]));
    for (int i = 1; i < nbuf; i++) HTRY(cudaStreamWaitEvent(g->streams[i], g->events[0], 0));
    for (int c = 0, b0 = 0; c < nchunk; b0 += sched[c], c++) {
        const int s = c % nbuf;
        cudaStream_t st = g->streams[s];
        float *d_xa = (float *)g->ws[s][0], *d_soft = (float *)g->ws[s][1];
        uint8_t *d_hard = (uint8_t *)g->ws[s][2];
        int8_t *d_q = (int8_t *)g->ws[s][3];
        const int nbw = sched[c];
        HTRY(cudaMemcpyAsync(d_q, xq_host + (size_t)b0 * NZ, (size_t)nbw * NZ, cudaMemcpyHostToDevice, st));
        HTRY((cudaError_t)nldpc::launch_q8_to_f32(d_q, d_xa, (size_t)nbw * NZ, scale, g->sm_count, st));
        int rc = nldpc_boosted_forward(g, cfg, d_xa, vn_w_host ? d_vn : nullptr, cn_w_host ? d_cn : nullptr, ucn_w_host ? d_ucn : nullptr,
                                       nbw, T, soft_mode, d_soft, hard_mode, d_hard, nullptr, st);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        if (soft_mode == NLDPC_OUT_ALL) {
            HTRY(cudaMemcpy2DAsync(soft_host + (size_t)b0 * NZ, (size_t)B * NZ * 4, d_soft, (size_t)nbw * NZ * 4,
                                   (size_t)nbw * NZ * 4, T, cudaMemcpyDeviceToHost, st));
        } else if (soft_mode == NLDPC_OUT_LAST) {
            HTRY(cudaMemcpyAsync(soft_host + (size_t)b0 * NZ, d_soft, (size_t)nbw * NZ * 4, cudaMemcpyDeviceToHost, st));
        }
        if (hard_mode == NLDPC_OUT_ALL) {
            HTRY(cudaMemcpy2DAsync(hard_host + (size_t)b0 * nb, (size_t)B * nb, d_hard, (size_t)nbw * nb, (size_t)nbw * nb, T,
                                   cudaMemcpyDeviceToHost, st));
        } else if (hard_mode == NLDPC_OUT_LAST) {
            HTRY(cudaMemcpyAsync(hard_host + (size_t)b0 * nb, d_hard, (size_t)nbw * nb, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int i = 0; i < nbuf; i++) HTRY(cudaStreamSynchronize(g->streams[i]));
#undef HTRY
    return NLDPC_OK;
}
