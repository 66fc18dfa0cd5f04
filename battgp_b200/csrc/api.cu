// extern "C" surface of libbattgp_b200.so (include/battgp_b200.h) and the look-ahead Cholesky driver.
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "common.cuh"

namespace bgp {

static thread_local char g_err[512] = "";
void set_error(const char* what, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

int cov_build(Ctx*, const bgp_kernel_spec*, const double*, int64_t, int64_t, const double*, int64_t, int64_t, double*,
              int64_t, int, cudaStream_t);
int cov_diag(Ctx*, const bgp_kernel_spec*, const double*, int64_t, int64_t, double*, cudaStream_t);
int trsv_lower(Ctx*, const double*, int64_t, int64_t, const double*, double*, double*, cudaStream_t);
int trsv_lower_t(Ctx*, const double*, int64_t, int64_t, const double*, double*, double*, cudaStream_t);
int dot(Ctx*, const double*, const double*, int64_t, double*, double*, cudaStream_t);
int predict_tail(Ctx*, int64_t, int64_t, const double*, int64_t, const double*, const double*, int64_t, const double*,
                 double, double*, double*, cudaStream_t);
int gemv_t(Ctx*, const double*, int64_t, int64_t, int64_t, const double*, double*, double, cudaStream_t);
int rowsumsq(Ctx*, const double*, int64_t, int64_t, int64_t, double*, int, cudaStream_t);
int64_t oz_slice_buffer_bytes(int64_t rows, int64_t K);
int oz_slice(Ctx*, const double*, int64_t, int64_t, int64_t, void*, cudaStream_t, const int32_t* blkmap = nullptr, int64_t blkrows = 0);
int oz_gemm(Ctx*, const void*, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int64_t, int64_t, double, double*,
            int64_t, int, int64_t, int64_t, cudaStream_t, int tiles_per_cta = 0);
int oz2_residues(Ctx*, const double*, int64_t, int64_t, int64_t, int8_t*, int32_t*, cudaStream_t);
int oz2_crt(Ctx*, const int32_t*, int64_t, int64_t, const int32_t*, const int32_t*, double, double*, int64_t, cudaStream_t);
int64_t oz2_gemm_work_bytes(int64_t, int64_t, int64_t);
int oz2_gemm(Ctx*, const double*, int64_t, int64_t, const double*, int64_t, int64_t, int64_t, double, double*, int64_t, void*, cudaStream_t);
int potri(Ctx*, double*, int64_t, int64_t, const double*, double*, int64_t, cudaStream_t);
int lml_grad(Ctx*, const bgp_kernel_spec*, const double*, int64_t, int64_t, const double*, int64_t, const double*,
             double*, cudaStream_t);

int fault_eval(Ctx*, const double*, const double*, int64_t, int64_t, int64_t, double, double, double*, double*, double*, double*, double*,
               double*, double*, cudaStream_t);
void leaf_clk_read(long long* out);
__global__ void init_scalars_kernel(int32_t* info, double* scal) {
    if (threadIdx.x == 0) { *info = INT_MAX; scal[0] = 0.0; }
}

__global__ void init_scalars_at_kernel(int32_t* info, double* logdet) {
    if (threadIdx.x == 0) { *info = INT_MAX; *logdet = 0.0; }
}
// lml = -0.5 z.z - 0.5 logdet - 0.5 n log(2 pi), everything on the device (bgp_lml_dev)
__global__ void lml_finish_kernel(const double* zz, const double* logdet, double n, double* out) {
    if (threadIdx.x == 0) *out = -0.5 * zz[0] - 0.5 * logdet[0] - 0.5 * n * 1.8378770664093454835606594728112;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Panel schedule of the look-ahead factorisation.  Panel k covers columns [start[k], start[k+1]).  With the "nb" knob set
// every panel has that width; otherwise the width follows the rows still to be factorised (m = R - start[k]): wide panels
// while the trailing update is long enough to hide the panel chain behind it (and wide K amortises the fixed per-tile cost
// of the int8 path), narrower ones once the chain  diag-block -> panel TRSM -> next column block  is what is exposed.
struct PanelSchedule {
    std::vector<int64_t> start;      // npanels + 1 entries
    int64_t ozbytes = 0;             // one digit-plane buffer for the rows below a panel (int8 path)
    int64_t trsm_bytes = 0;          // scratch of the int8 path inside the panel TRSM
    int64_t npanels() const { return (int64_t)start.size() - 1; }
    int64_t width(int64_t k) const { return start[k + 1] - start[k]; }
};

static inline int64_t panel_width(const Ctx* ctx, int64_t m, int64_t idx) {
    if (ctx->nb > 0) return ctx->nb;
    int64_t w = (m >= ctx->sched_t1024) ? 1024 : 512;
    if (ctx->ozaki) {
        if (ctx->sched_t2048 > 0 && m >= ctx->sched_t2048) w = 2048;
        if (ctx->sched_t4096 > 0 && m >= ctx->sched_t4096) w = 4096;
    }
    if (idx == 0 && ctx->sched_w0 > 0 && w > ctx->sched_w0) w = ctx->sched_w0;
    if (idx == 1 && ctx->sched_w1 > 0 && w > ctx->sched_w1) w = ctx->sched_w1;
    return w;
}

static inline int64_t round256(int64_t b) { return ((b + 255) / 256) * 256; }

// R rows (n of them the symmetric matrix, the rest ride along), n columns
static void make_schedule(const Ctx* ctx, int64_t R, int64_t n, PanelSchedule& s) {
    s.start.clear();
    s.ozbytes = s.trsm_bytes = 0;
    for (int64_t k0 = 0, idx = 0; k0 < n; idx++) {
        int64_t w = panel_width(ctx, R - k0, idx);
        if (w > n - k0) w = n - k0;
        s.start.push_back(k0);
        k0 += w;
        const int64_t below = R - k0;
        if (below > 0 && w % 64 == 0) {
            const int64_t ob = round256(oz_slice_buffer_bytes(below, w));
            // X1 [below, w/2] and L21 [w/2, w/2] digit planes (top split of trsm_rlt_rec; deeper levels are smaller)
            const int64_t tb = round256(oz_slice_buffer_bytes(below, w / 2)) + round256(oz_slice_buffer_bytes(w / 2, w / 2));
            if (k0 < n && ob > s.ozbytes) s.ozbytes = ob;      // the last panel is never sliced
            if (tb > s.trsm_bytes) s.trsm_bytes = tb;
        }
    }
    s.start.push_back(n);
}

// Right-looking over NB-wide panels with one panel of look-ahead (see potrf.cu header comment).
// A is (n + mx) x n: the first n rows are the symmetric matrix (lower triangle), the mx EXTRA rows ride along -- they take
// part in every panel solve and trailing update, so they come out as X L^-T (predictive-variance solve fused into the
// factorisation: the M = 300 query rows are processed by the big tiles of the trailing updates instead of a separate
// recursive TRSM).
static int potrf_driver(Ctx* ctx, double* A, int64_t n, int64_t mx, int64_t lda, double* dinv, cudaStream_t mainst) {
    const int64_t R = n + mx;
    // programmatic dependent launch of the leaf / DMMA GEMM chain pays off while the factorisation is latency-bound
    // (profiles/probe_r01_pdl.jsonl: N=4096 -4 %, N=8192 -2 %, N>=16384 +0.5 %)
    struct PdlChain {
        Ctx* c;
        int prev;
        PdlChain(Ctx* c_, bool on) : c(c_), prev(c_->pdl_chain) { c->pdl_chain = on ? 1 : 0; }
        ~PdlChain() { c->pdl_chain = prev; }
    } pdl_chain(ctx, R < ctx->pdl_chain_rows);
    PanelSchedule S;
    make_schedule(ctx, R, n, S);
    const int64_t npanels = S.npanels();
    // small matrices: one two-stream leaf chain over the whole matrix (potrf_chain2) instead of panels with look-ahead
    const bool whole_chain = ctx->leaf_chain && R <= ctx->chain_whole_max && n <= ctx->leaf_chain_max;
    if (!ctx->lookahead || npanels <= 2 || whole_chain) {
        int rc = potrf_rec(ctx, A, n, lda, dinv, 0, mainst);
        if (!rc && mx > 0) rc = trsm_rlt_rec(ctx, A, n, lda, dinv, A + n * lda, mx, lda, mainst);
        return rc;
    }
    cudaStream_t P = ctx->panel_stream;
    // int8/tcgen05 trailing updates (ozaki.cu): two slice buffers (panel k is still being read by T_k on the caller's
    // stream while panel k+1 is sliced on the panel stream)
    const int64_t ozbytes = S.ozbytes;
    bool oz = ctx->ozaki && ctx->ws && ((lda & 1) == 0) && (((uintptr_t)A & 15) == 0) && (((uintptr_t)ctx->ws & 255) == 0) &&
              ozbytes > 0 && ctx->ws_bytes >= 2 * ozbytes;
    for (int64_t k = 0; oz && k + 1 < npanels; k++) oz = (S.width(k) % 64 == 0);
    void* ozbuf[2] = {ctx->ws, reinterpret_cast<char*>(ctx->ws) + ozbytes};
    // the panel TRSM (stream P) gets its own scratch behind the two panel buffers
    struct TrsmScratch {
        Ctx* c;
        TrsmScratch(Ctx* c_, void* p, int64_t b) : c(c_) { c->ws_trsm = p; c->ws_trsm_bytes = b; }
        ~TrsmScratch() { c->ws_trsm = nullptr; c->ws_trsm_bytes = 0; }
    } trsm_scratch(ctx, (oz && ctx->ws_bytes > 2 * ozbytes) ? reinterpret_cast<char*>(ctx->ws) + 2 * ozbytes : nullptr,
                   (oz && ctx->ws_bytes > 2 * ozbytes) ? ctx->ws_bytes - 2 * ozbytes : 0);
    BGP_CUDA_OK(cudaEventRecord(ctx->ev_fork, mainst));
    BGP_CUDA_OK(cudaStreamWaitEvent(P, ctx->ev_fork, 0));
    // "trace" knob: timed events around every P_k / T_k, printed to stderr after the factorisation (diagnostics only)
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s) {
        if (!ctx->trace) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        tev.push_back(e);
    };
    mark(mainst);
    for (int64_t k = 0; k < npanels; k++) {
        const int64_t k0 = S.start[k];
        const int64_t nbk = S.width(k);
        const int64_t wprev = k >= 1 ? S.width(k - 1) : 0;
        const int64_t below = R - k0 - nbk;                    // rows under the diagonal block, extra rows included
        double* Akk = A + k0 * lda + k0;
        double* dinv_k = dinv + (k0 / LEAF) * (int64_t)LEAF * LEAF;
        int rc;
        // ---- P_k (panel stream): bring column block k up to date with panel k-1, factor it, solve the rows below
        mark(P);
        if (k >= 1) {
            if (k >= 2) BGP_CUDA_OK(cudaStreamWaitEvent(P, ctx->ev_trail[(k - 2) % 3], 0));
            if (oz) {
                // panel k-1 was sliced (rows k0.. of it are rows 0.. of its slice buffer)
                if ((rc = oz_gemm(ctx, ozbuf[(k - 1) & 1], R - k0, 0, ozbuf[(k - 1) & 1], R - k0, 0, R - k0, nbk, wprev, -1.0, Akk, lda,
                                  1, 0, 0, P, ctx->oz_tpc))) return rc;
            } else {
                const double* Lp = A + k0 * lda + (k0 - wprev);  // rows k0.., columns of panel k-1
                GemmArgs g{Lp, lda, Lp, lda, Akk, lda, (int)(R - k0), (int)nbk, (int)wprev, -1.0, 1.0, 1, 0, 0};
                if ((rc = gemm_nt(ctx, g, P))) return rc;
            }
        }
        mark(P);
        if ((rc = potrf_rec(ctx, Akk, nbk, lda, dinv_k, k0, P))) return rc;
        mark(P);
        if (below > 0 && (rc = trsm_rlt_rec(ctx, Akk, nbk, lda, dinv_k, Akk + nbk * lda, below, lda, P))) return rc;
        mark(P);
        // digit planes of the rows below: read by the column-block update of P_k+1 and by T_k
        if (oz && k + 1 < npanels && below > 0 && (rc = oz_slice(ctx, Akk + nbk * lda, below, nbk, lda, ozbuf[k & 1], P))) return rc;
        BGP_CUDA_OK(cudaEventRecord(ctx->ev_panel[k % 2], P));
        mark(P);
        // ---- T_k (caller's stream): rank-nbk update of everything right of column block k+1
        if (k + 2 < npanels) {
            const int64_t t0 = S.start[k + 2];
            const int64_t wnext = S.width(k + 1);
            BGP_CUDA_OK(cudaStreamWaitEvent(mainst, ctx->ev_panel[k % 2], 0));
            mark(mainst);
            if (oz) {
                const bool prof = ctx->prof != 0;
                if (prof) {
                    if (ctx->prof_ev.size() < 2 * (ctx->prof_used + 1)) {
                        cudaEvent_t a, b;
                        BGP_CUDA_OK(cudaEventCreate(&a));
                        BGP_CUDA_OK(cudaEventCreate(&b));
                        ctx->prof_ev.push_back(a); ctx->prof_ev.push_back(b);
                        ctx->prof_flop.push_back(0.0);
                    }
                    BGP_CUDA_OK(cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used], mainst));
                }
                // tail panels: the panel chain is what is exposed there, so the (short) trailing update runs as one persistent
                // CTA per SM on all but oz_reserve SMs and the high-priority panel kernels start without waiting for a CTA to retire
                const bool tail = ctx->sched_tail > 0 && (R - k0) < ctx->sched_tail;
                ctx->oz_reserve_now = tail ? ctx->oz_reserve : 0;
                rc = oz_gemm(ctx, ozbuf[k & 1], below, wnext, ozbuf[k & 1], below, wnext, R - t0, n - t0, nbk, -1.0,
                             A + t0 * lda + t0, lda, 1, 0, 0, mainst, tail ? 0 : ctx->oz_tpc);
                ctx->oz_reserve_now = 0;
                if (rc) return rc;
                if (prof) {
                    BGP_CUDA_OK(cudaEventRecord(ctx->prof_ev[2 * ctx->prof_used + 1], mainst));
                    const double c = (double)(n - t0);          // lower triangle of the square part + the full extra rows
                    ctx->prof_flop[ctx->prof_used] = 2.0 * (c * (c + 1.0) * 0.5 + (double)(R - n) * c) * (double)nbk;
                    ctx->prof_used++;
                }
            } else {
                const double* Lt = A + t0 * lda + k0;
                GemmArgs g{Lt, lda, Lt, lda, A + t0 * lda + t0, lda, (int)(R - t0), (int)(n - t0), (int)nbk, -1.0, 1.0, 1, 0, 0};
                if ((rc = gemm_nt(ctx, g, mainst))) return rc;
            }
            BGP_CUDA_OK(cudaEventRecord(ctx->ev_trail[k % 3], mainst));
            mark(mainst);
        } else if (ctx->trace) { mark(mainst); mark(mainst); }
    }
    BGP_CUDA_OK(cudaEventRecord(ctx->ev_join, P));
    BGP_CUDA_OK(cudaStreamWaitEvent(mainst, ctx->ev_join, 0));
    if (ctx->trace) {
        // per panel: 5 marks on P (start, after column-block update, after diag, after TRSM, after slice) + 2 on main (T_k)
        BGP_CUDA_OK(cudaStreamSynchronize(mainst));
        auto at = [&](size_t i) { float ms = 0.f; cudaEventElapsedTime(&ms, tev[0], tev[i]); return ms; };
        for (int64_t k = 0; k < npanels; k++) {
            const size_t b = 1 + 7 * (size_t)k;
            fprintf(stderr, "{\"potrf_trace\": %lld, \"w\": %lld, \"P_start\": %.3f, \"colupd\": %.3f, \"diag\": %.3f, \"trsm\": %.3f, "
                            "\"slice\": %.3f, \"P_end\": %.3f, \"T_start\": %.3f, \"T_end\": %.3f}\n",
                    (long long)k, (long long)S.width(k), at(b), at(b + 1) - at(b), at(b + 2) - at(b + 1), at(b + 3) - at(b + 2),
                    at(b + 4) - at(b + 3), at(b + 4), at(b + 5), at(b + 6));
        }
        for (auto e : tev) cudaEventDestroy(e);
    }
    return 0;
}

}  // namespace bgp

using namespace bgp;

#define CTX_OR_FAIL(c)                      \
    if (!(c)) return BGP_E_ARG;             \
    Ctx* ctx = reinterpret_cast<Ctx*>(c);   \
    DeviceGuard guard__(ctx->device);       \
    if (!guard__.ok) { set_error("cudaSetDevice", cudaGetLastError()); return BGP_E_CUDA; }

extern "C" {

int bgp_version(void) { return BGP_VERSION; }
void bgp_debug_leaf_clk(long long* out) { bgp::leaf_clk_read(out); }
// diagnostics (tools/oz_timeline.py): CTA 0 of the int8 kernel writes 16 clock64 stamps per tile into buf (device, cap tiles)
void bgp_debug_oz_timeline(bgp_ctx* p, long long* buf, int cap) {
    if (!p) return;
    Ctx* c = reinterpret_cast<Ctx*>(p);
    c->oz_dbg = cap > 0 ? buf : nullptr;
    c->oz_dbg_cap = cap > 0 ? cap : 0;
}
const char* bgp_last_error(void) { return g_err; }

int bgp_grad_slots(const bgp_kernel_spec* s) {
    if (!s || s->nterms < 1 || s->nterms > BGP_MAX_TERMS) return BGP_E_SPEC;
    int slots = 1;
    for (int t = 0; t < s->nterms; t++) {
        const bgp_term& T = s->terms[t];
        slots += 1;
        if (T.type == BGP_WIENER) continue;
        if (T.ndims < 1 || T.ndims > BGP_MAX_DIMS) return BGP_E_SPEC;
        slots += T.ndims;
        if (T.type == BGP_PERIODIC) slots += T.ndims;
    }
    return slots;
}

int bgp_ctx_create(int device, bgp_ctx** out) {
    if (!out) return BGP_E_ARG;
    *out = nullptr;
    DeviceGuard guard(device);
    if (!guard.ok) { set_error("cudaSetDevice", cudaGetLastError()); return BGP_E_CUDA; }
    Ctx* c = new (std::nothrow) Ctx();
    if (!c) return BGP_E_ARG;
    c->device = device;
    int lo = 0, hi = 0;
    BGP_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    BGP_CUDA_OK(cudaStreamCreateWithPriority(&c->panel_stream, cudaStreamNonBlocking, hi));
    BGP_CUDA_OK(cudaStreamCreateWithPriority(&c->leaf_stream, cudaStreamNonBlocking, hi));
    for (auto& e : c->ev_chain) BGP_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    BGP_CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    BGP_CUDA_OK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (auto& e : c->ev_panel) BGP_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : c->ev_trail) BGP_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    BGP_CUDA_OK(cudaMalloc(&c->d_info, 256));
    BGP_CUDA_OK(cudaMalloc(&c->d_scal, 512 * sizeof(double)));
    BGP_CUDA_OK(cudaMalloc(&c->d_scratch, SCRATCH_BYTES));
    *out = reinterpret_cast<bgp_ctx*>(c);
    return 0;
}

void bgp_ctx_destroy(bgp_ctx* p) {
    if (!p) return;
    Ctx* c = reinterpret_cast<Ctx*>(p);
    DeviceGuard guard(c->device);
    if (c->panel_stream) { cudaStreamSynchronize(c->panel_stream); cudaStreamDestroy(c->panel_stream); }
    if (c->leaf_stream) { cudaStreamSynchronize(c->leaf_stream); cudaStreamDestroy(c->leaf_stream); }
    for (auto& e : c->ev_chain) if (e) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    for (auto& e : c->ev_panel) if (e) cudaEventDestroy(e);
    for (auto& e : c->ev_trail) if (e) cudaEventDestroy(e);
    if (c->d_info) cudaFree(c->d_info);
    if (c->d_scal) cudaFree(c->d_scal);
    if (c->d_scratch) cudaFree(c->d_scratch);
    for (auto& e : c->prof_ev) cudaEventDestroy(e);
    delete c;
}

int bgp_ctx_set(bgp_ctx* p, const char* key, int value) {
    if (!p || !key) return BGP_E_ARG;
    Ctx* c = reinterpret_cast<Ctx*>(p);
    if (!strcmp(key, "nb")) {
        if (value != 0 && (value < LEAF || value % LEAF != 0 || value > 8192)) return BGP_E_ARG;
        c->nb = value;
        return 0;
    }
    if (!strcmp(key, "lookahead")) { c->lookahead = value ? 1 : 0; return 0; }
    if (!strcmp(key, "trace")) { c->trace = value ? 1 : 0; return 0; }
    if (!strcmp(key, "pdl")) { c->pdl = value ? 1 : 0; return 0; }
    if (!strncmp(key, "sched_", 6)) {
        if (value < 0 || (!strncmp(key, "sched_w", 7) && value % LEAF != 0)) return BGP_E_ARG;
        if (!strcmp(key, "sched_t1024")) { c->sched_t1024 = value; return 0; }
        if (!strcmp(key, "sched_t2048")) { c->sched_t2048 = value; return 0; }
        if (!strcmp(key, "sched_t4096")) { c->sched_t4096 = value; return 0; }
        if (!strcmp(key, "sched_w0")) { c->sched_w0 = value; return 0; }
        if (!strcmp(key, "sched_w1")) { c->sched_w1 = value; return 0; }
        if (!strcmp(key, "sched_tail")) { c->sched_tail = value; return 0; }
        return BGP_E_ARG;
    }
    if (!strcmp(key, "leaf_chain")) { c->leaf_chain = value ? 1 : 0; return 0; }
    if (!strcmp(key, "chain_whole_max")) { if (value < 0 || value > 8192) return BGP_E_ARG; c->chain_whole_max = value; return 0; }
    if (!strcmp(key, "pdl_chain_rows")) { if (value < 0) return BGP_E_ARG; c->pdl_chain_rows = value; return 0; }
    if (!strcmp(key, "chain_split")) { c->chain_split = value ? 1 : 0; return 0; }
    if (!strcmp(key, "chain_cfg")) { c->chain_cfg = value ? 1 : 0; return 0; }
    if (!strcmp(key, "leaf_chain_max")) { if (value < 256 || value > 8192) return BGP_E_ARG; c->leaf_chain_max = value; return 0; }
    if (!strcmp(key, "ozaki")) { c->ozaki = value ? 1 : 0; return 0; }
    if (!strcmp(key, "oz_cluster")) { if (value != 1 && value != 2 && value != 4) return BGP_E_ARG; c->oz_cluster = value; return 0; }
    if (!strcmp(key, "oz_kfence")) { c->oz_kfence = value ? 1 : 0; return 0; }
    if (!strcmp(key, "oz_dbg_epi")) { if (value < 0 || value > 3) return BGP_E_ARG; c->oz_dbg_epi = value; return 0; }
    if (!strcmp(key, "oz_backoff")) { c->oz_backoff = value ? 1 : 0; return 0; }
    if (!strcmp(key, "oz_relay")) { c->oz_relay = value ? 1 : 0; return 0; }
    if (!strcmp(key, "oz_order")) { if (value < 0 || value > 2) return BGP_E_ARG; c->oz_order = value; return 0; }
    if (!strcmp(key, "oz_collector")) { c->oz_collector = value ? 1 : 0; return 0; }
    if (!strcmp(key, "oz_l2hint")) { if (value < 0 || value > 3) return BGP_E_ARG; c->oz_l2hint = value; return 0; }
    if (!strcmp(key, "oz_reserve")) { if (value < 0 || value > 128) return BGP_E_ARG; c->oz_reserve = value; return 0; }
    if (!strcmp(key, "oz_group")) { if (value < 1 || value > 1024) return BGP_E_ARG; c->oz_group = value; return 0; }
    if (!strcmp(key, "oz_tpc_gemm")) { if (value < 0 || value > 4096) return BGP_E_ARG; c->oz_tpc_gemm = value; return 0; }
    if (!strcmp(key, "oz_tpc")) { if (value < 0 || value > 4096) return BGP_E_ARG; c->oz_tpc = value; return 0; }
    if (!strcmp(key, "gemm_cfg")) { if (value < 0 || value > 7) return BGP_E_ARG; c->gemm_cfg = value; return 0; }
    return BGP_E_ARG;
}

int bgp_panel_schedule(int64_t rows, int64_t n, int ozaki, int nb, int64_t* starts, int cap) {
    if (rows < n || n < 0 || nb < 0 || (nb > 0 && nb % LEAF != 0) || cap < 0 || (cap > 0 && !starts)) return BGP_E_ARG;
    Ctx c;                       // defaults of a fresh context; no CUDA call is made
    c.ozaki = ozaki ? 1 : 0;
    c.nb = nb;
    PanelSchedule S;
    make_schedule(&c, rows, n, S);
    const int np = (int)S.npanels();
    for (int i = 0; i <= np && i < cap; i++) starts[i] = S.start[i];
    return np;
}

int64_t bgp_potrf_workspace_bytes(const bgp_ctx* p, int64_t n) {
    if (!p || n <= 0) return 0;
    const Ctx* c = reinterpret_cast<const Ctx*>(p);
    // n counts every row (augmented ones included): the schedule of an [n, n - mx] factorisation is a prefix of this one
    PanelSchedule S;
    make_schedule(c, n, n, S);
    if (S.npanels() <= 2) return 0;
    return 2 * S.ozbytes + S.trsm_bytes;
}

int bgp_ctx_set_workspace(bgp_ctx* p, void* ptr, int64_t bytes) {
    if (!p || bytes < 0 || (bytes > 0 && !ptr)) return BGP_E_ARG;
    Ctx* c = reinterpret_cast<Ctx*>(p);
    c->ws = bytes ? ptr : nullptr;
    c->ws_bytes = bytes;
    return 0;
}

int64_t bgp_ctx_launches(const bgp_ctx* p) { return p ? reinterpret_cast<const Ctx*>(p)->launches : 0; }

int bgp_ctx_kernel_profile(bgp_ctx* p, int enable) {
    if (!p) return BGP_E_ARG;
    Ctx* c = reinterpret_cast<Ctx*>(p);
    c->prof = enable ? 1 : 0;
    c->prof_used = 0;
    return 0;
}

int bgp_ctx_kernel_profile_read(bgp_ctx* p, double* ms, double* flop, int64_t* nlaunch) {
    CTX_OR_FAIL(p);
    double tms = 0.0, tfl = 0.0;
    for (size_t i = 0; i < ctx->prof_used; i++) {
        BGP_CUDA_OK(cudaEventSynchronize(ctx->prof_ev[2 * i + 1]));
        float e = 0.f;
        BGP_CUDA_OK(cudaEventElapsedTime(&e, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        tms += e;
        tfl += ctx->prof_flop[i];
    }
    if (ms) *ms = tms;
    if (flop) *flop = tfl;
    if (nlaunch) *nlaunch = (int64_t)ctx->prof_used;
    ctx->prof_used = 0;
    return 0;
}

int bgp_cov_build(bgp_ctx* c, const bgp_kernel_spec* spec, const double* X1, int64_t n1, int64_t ldx1, const double* X2,
                  int64_t n2, int64_t ldx2, double* out, int64_t ldo, int symmetric, void* stream) {
    CTX_OR_FAIL(c);
    return cov_build(ctx, spec, X1, n1, ldx1, X2, n2, ldx2, out, ldo, symmetric, (cudaStream_t)stream);
}

int bgp_cov_diag(bgp_ctx* c, const bgp_kernel_spec* spec, const double* X, int64_t n, int64_t ldx, double* out,
                 void* stream) {
    CTX_OR_FAIL(c);
    return cov_diag(ctx, spec, X, n, ldx, out, (cudaStream_t)stream);
}

int bgp_gemm_nt(bgp_ctx* c, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
                const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int tri, int64_t roff, int64_t coff,
                void* stream) {
    CTX_OR_FAIL(c);
    if (M < 0 || N < 0 || K < 0 || M > INT_MAX || N > INT_MAX || K > INT_MAX) return BGP_E_ARG;
    if (M == 0 || N == 0) return 0;
    if (!C || ldc < N || (K > 0 && (!A || !B || lda < K || ldb < K))) return BGP_E_ARG;
    GemmArgs g{A, lda, B, ldb, C, ldc, (int)M, (int)N, (int)K, alpha, beta, tri ? 1 : 0, roff, coff};
    return gemm_nt(ctx, g, (cudaStream_t)stream);
}

int64_t bgp_potrf_dinv_elems(int64_t n) { return n <= 0 ? 0 : ((n + LEAF - 1) / LEAF) * (int64_t)LEAF * LEAF; }

int bgp_potrf(bgp_ctx* c, double* A, int64_t n, int64_t lda, double* dinv, double* logdet_host, void* stream) {
    return bgp_potrf_aug(c, A, n, 0, lda, dinv, logdet_host, stream);
}

int bgp_potrf_aug(bgp_ctx* c, double* A, int64_t n, int64_t mx, int64_t lda, double* dinv, double* logdet_host, void* stream) {
    CTX_OR_FAIL(c);
    if (n < 0 || mx < 0 || n + mx > INT_MAX) return BGP_E_ARG;
    if (n == 0) { if (logdet_host) *logdet_host = 0.0; return 0; }
    if (!A || !dinv || lda < n) return BGP_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    init_scalars_kernel<<<1, 32, 0, st>>>(ctx->d_info, ctx->d_scal);
    BGP_LAUNCH_OK(ctx);
    int rc = potrf_driver(ctx, A, n, mx, lda, dinv, st);
    if (rc) return rc;
    int32_t info = 0;
    double ld = 0.0;
    BGP_CUDA_OK(cudaMemcpyAsync(&info, ctx->d_info, sizeof(info), cudaMemcpyDeviceToHost, st));
    BGP_CUDA_OK(cudaMemcpyAsync(&ld, ctx->d_scal, sizeof(ld), cudaMemcpyDeviceToHost, st));
    BGP_CUDA_OK(cudaStreamSynchronize(st));
    if (logdet_host) *logdet_host = ld;
    return info == INT_MAX ? 0 : (int)info;
}

int bgp_potrf_async(bgp_ctx* c, double* A, int64_t n, int64_t mx, int64_t lda, double* dinv, int32_t* info_dev, double* logdet_dev,
                    void* stream) {
    CTX_OR_FAIL(c);
    if (n < 0 || mx < 0 || n + mx > INT_MAX || !info_dev || !logdet_dev) return BGP_E_ARG;
    if (n == 0) return 0;
    if (!A || !dinv || lda < n) return BGP_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    init_scalars_at_kernel<<<1, 32, 0, st>>>(info_dev, logdet_dev);
    BGP_LAUNCH_OK(ctx);
    // single stream, no look-ahead, nothing read back: the recursive factorisation with the leaf outputs redirected
    int32_t* si = ctx->d_info;
    double* ss = ctx->d_scal;
    ctx->d_info = info_dev;
    ctx->d_scal = logdet_dev;
    int rc = potrf_rec(ctx, A, n, lda, dinv, 0, st);
    if (!rc && mx > 0) rc = trsm_rlt_rec(ctx, A, n, lda, dinv, A + n * lda, mx, lda, st);
    ctx->d_info = si;
    ctx->d_scal = ss;
    return rc;
}

int bgp_lml_dev(bgp_ctx* c, const double* z, int64_t n, const double* logdet_dev, double* lml_dev, void* stream) {
    CTX_OR_FAIL(c);
    if (!lml_dev || !logdet_dev || n <= 0 || !z) return BGP_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = dot(ctx, z, z, n, ctx->d_scal + 8, ctx->d_scal + 1, st);
    if (rc) return rc;
    lml_finish_kernel<<<1, 32, 0, st>>>(ctx->d_scal + 1, logdet_dev, (double)n, lml_dev);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

int bgp_potrf_block(bgp_ctx* c, double* A, int64_t nb, int64_t lda, double* dinv, int32_t* info_dev, double* logdet_dev,
                    void* stream) {
    CTX_OR_FAIL(c);
    if (!A || !dinv || nb <= 0 || lda < nb) return BGP_E_ARG;
    // temporarily redirect the leaf outputs to the caller's device scalars
    int32_t* si = ctx->d_info;
    double* ss = ctx->d_scal;
    if (info_dev) ctx->d_info = info_dev;
    if (logdet_dev) ctx->d_scal = logdet_dev;
    int rc = potrf_rec(ctx, A, nb, lda, dinv, 0, (cudaStream_t)stream);
    ctx->d_info = si;
    ctx->d_scal = ss;
    return rc;
}

int bgp_potrs_vec(bgp_ctx* c, const double* L, int64_t n, int64_t ldl, const double* dinv, const double* y, double* z,
                  double* alpha, void* stream) {
    CTX_OR_FAIL(c);
    if (n == 0) return 0;
    if (!L || !dinv || !y || !z || !alpha || n < 0 || ldl < n) return BGP_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (z != y) BGP_CUDA_OK(cudaMemcpyAsync(z, y, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    int rc = trsv_lower(ctx, L, n, ldl, dinv, z, z, st);
    if (rc) return rc;
    BGP_CUDA_OK(cudaMemcpyAsync(alpha, z, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return trsv_lower_t(ctx, L, n, ldl, dinv, alpha, alpha, st);
}

int bgp_trsm_rlt(bgp_ctx* c, const double* L, int64_t n, int64_t ldl, const double* dinv, double* X, int64_t m,
                 int64_t ldx, void* stream) {
    CTX_OR_FAIL(c);
    if (n == 0 || m == 0) return 0;
    if (!L || !dinv || !X || n < 0 || m < 0 || ldl < n || ldx < n || m > INT_MAX) return BGP_E_ARG;
    // stand-alone call (no factorisation in flight on this context): the int8 path of the big updates may use the whole
    // caller-provided workspace
    void* p0 = ctx->ws_trsm;
    const int64_t b0 = ctx->ws_trsm_bytes;
    if (!p0 && ctx->ws) { ctx->ws_trsm = ctx->ws; ctx->ws_trsm_bytes = ctx->ws_bytes; }
    const int tpc = ctx->oz_tpc;
    ctx->oz_tpc = 0;
    const int rc = trsm_rlt_rec(ctx, L, n, ldl, dinv, X, m, ldx, (cudaStream_t)stream);
    ctx->oz_tpc = tpc;
    ctx->ws_trsm = p0;
    ctx->ws_trsm_bytes = b0;
    return rc;
}

int bgp_predict_tail(bgp_ctx* c, int64_t m, int64_t n, const double* Kq, int64_t ldk, const double* alpha,
                     const double* V, int64_t ldv, const double* kdiag, double min_var, double* mean, double* var,
                     void* stream) {
    CTX_OR_FAIL(c);
    if (m < 0 || n < 0) return BGP_E_ARG;
    if (mean && (!Kq || !alpha || ldk < n)) return BGP_E_ARG;
    if (var && (!V || !kdiag || ldv < n)) return BGP_E_ARG;
    return predict_tail(ctx, m, n, Kq, ldk, alpha, V, ldv, kdiag, min_var, mean, var, (cudaStream_t)stream);
}

int bgp_lml(bgp_ctx* c, const double* z, int64_t n, double logdet, double* lml_host, void* stream) {
    CTX_OR_FAIL(c);
    if (!lml_host || n < 0 || (n > 0 && !z)) return BGP_E_ARG;
    double zz = 0.0;
    if (n > 0) {
        cudaStream_t st = (cudaStream_t)stream;
        int rc = dot(ctx, z, z, n, ctx->d_scal + 8, ctx->d_scal + 1, st);
        if (rc) return rc;
        BGP_CUDA_OK(cudaMemcpyAsync(&zz, ctx->d_scal + 1, sizeof(double), cudaMemcpyDeviceToHost, st));
        BGP_CUDA_OK(cudaStreamSynchronize(st));
    }
    *lml_host = -0.5 * zz - 0.5 * logdet - 0.5 * (double)n * log(2.0 * M_PI);
    return 0;
}

int bgp_trsv(bgp_ctx* c, const double* L, int64_t n, int64_t ldl, const double* dinv, double* b, int trans, void* stream) {
    CTX_OR_FAIL(c);
    if (n == 0) return 0;
    if (!L || !dinv || !b || n < 0 || ldl < n) return BGP_E_ARG;
    return trans ? trsv_lower_t(ctx, L, n, ldl, dinv, b, b, (cudaStream_t)stream)
                 : trsv_lower(ctx, L, n, ldl, dinv, b, b, (cudaStream_t)stream);
}

int bgp_gemv_t(bgp_ctx* c, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* v, double* y,
               double alpha, void* stream) {
    CTX_OR_FAIL(c);
    if (rows < 0 || cols < 0 || (rows > 0 && cols > 0 && (!A || !v || !y || lda < cols))) return BGP_E_ARG;
    return gemv_t(ctx, A, rows, cols, lda, v, y, alpha, (cudaStream_t)stream);
}

int bgp_rowsumsq(bgp_ctx* c, const double* V, int64_t m, int64_t n, int64_t ldv, double* out, int accumulate, void* stream) {
    CTX_OR_FAIL(c);
    if (m < 0 || n < 0 || (m > 0 && (!V || !out || ldv < n))) return BGP_E_ARG;
    return rowsumsq(ctx, V, m, n, ldv, out, accumulate, (cudaStream_t)stream);
}

int64_t bgp_oz_slice_bytes(int64_t rows, int64_t K) { return (rows <= 0 || K <= 0) ? 0 : ((oz_slice_buffer_bytes(rows, K) + 255) / 256) * 256; }

int bgp_oz_slice(bgp_ctx* c, const double* P, int64_t rows, int64_t K, int64_t ld, void* buf, int64_t buf_bytes, void* stream) {
    CTX_OR_FAIL(c);
    if (rows <= 0) return 0;
    if (!P || !buf || K <= 0 || K % 64 || ld < K || buf_bytes < oz_slice_buffer_bytes(rows, K) || ((uintptr_t)buf & 255)) return BGP_E_ARG;
    return oz_slice(ctx, P, rows, K, ld, buf, (cudaStream_t)stream);
}

int bgp_oz_slice_gather(bgp_ctx* c, const double* P, int64_t rows, int64_t K, int64_t ld, const int32_t* blkmap, int64_t blkrows,
                        void* buf, int64_t buf_bytes, void* stream) {
    CTX_OR_FAIL(c);
    if (rows <= 0) return 0;
    if (!P || !buf || !blkmap || blkrows <= 0 || K <= 0 || K % 64 || ld < K || buf_bytes < oz_slice_buffer_bytes(rows, K) ||
        ((uintptr_t)buf & 255)) return BGP_E_ARG;
    return oz_slice(ctx, P, rows, K, ld, buf, (cudaStream_t)stream, blkmap, blkrows);
}

int bgp_oz_gemm(bgp_ctx* c, const void* bufA, int64_t rowsA, int64_t arow0, const void* bufB, int64_t rowsB, int64_t brow0,
                int64_t M, int64_t N, int64_t K, double alpha, double* C, int64_t ldc, int tri, int64_t roff, int64_t coff,
                void* stream) {
    CTX_OR_FAIL(c);
    if (M < 0 || N < 0 || M > INT_MAX || N > INT_MAX) return BGP_E_ARG;
    if (M == 0 || N == 0) return 0;
    if (!bufA || !bufB || !C || K <= 0 || K % 64 || K > 16384 || ldc < N || arow0 < 0 || brow0 < 0 || arow0 + M > ((rowsA + 127) / 128) * 128 ||
        brow0 + N > ((rowsB + 127) / 128) * 128) return BGP_E_ARG;
    return oz_gemm(ctx, bufA, rowsA, arow0, bufB, rowsB, brow0, M, N, K, alpha, C, ldc, tri ? 1 : 0, roff, coff, (cudaStream_t)stream,
                   ctx->oz_tpc_gemm);
}

int64_t bgp_gemm_nt_i8_work_bytes(int64_t M, int64_t N, int64_t K) {
    return oz_slice_buffer_bytes(M, K) + oz_slice_buffer_bytes(N, K) + 512;
}

int bgp_gemm_nt_i8(bgp_ctx* c, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, const double* B,
                   int64_t ldb, double* C, int64_t ldc, int tri, int64_t roff, int64_t coff, void* work, int64_t work_bytes,
                   void* stream) {
    CTX_OR_FAIL(c);
    // K <= 16384: the int32 accumulators hold 7 digit pairs of at most 2^14 per k (csrc/ozaki.cu OZ_MAX_K)
    if (M < 0 || N < 0 || K <= 0 || K % 64 || K > 16384 || M > INT_MAX || N > INT_MAX) return BGP_E_ARG;
    if (M == 0 || N == 0) return 0;
    if (!A || !B || !C || !work || lda < K || ldb < K || ldc < N || work_bytes < bgp_gemm_nt_i8_work_bytes(M, N, K)) return BGP_E_ARG;
    if ((uintptr_t)work & 255) return BGP_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    char* wa = reinterpret_cast<char*>(work);
    char* wb = wa + ((oz_slice_buffer_bytes(M, K) + 255) / 256) * 256;
    int rc = oz_slice(ctx, A, M, K, lda, wa, st);
    if (rc) return rc;
    const bool same = (A == B && lda == ldb && M == N);
    if (!same && (rc = oz_slice(ctx, B, N, K, ldb, wb, st))) return rc;
    return oz_gemm(ctx, wa, M, 0, same ? wa : wb, N, 0, M, N, K, alpha, C, ldc, tri ? 1 : 0, roff, coff, st);
}

int bgp_oz2_residues(bgp_ctx* c, const double* A, int64_t rows, int64_t K, int64_t ld, int8_t* residues, int32_t* expo, void* stream) {
    CTX_OR_FAIL(c);
    if (rows < 0 || K < 0 || (rows > 0 && K > 0 && (!A || !residues || !expo || ld < K))) return BGP_E_ARG;
    return oz2_residues(ctx, A, rows, K, ld, residues, expo, (cudaStream_t)stream);
}

int bgp_oz2_crt(bgp_ctx* c, const int32_t* G, int64_t M, int64_t N, const int32_t* ea, const int32_t* eb, double alpha, double* C,
                int64_t ldc, void* stream) {
    CTX_OR_FAIL(c);
    if (M < 0 || N < 0 || (M > 0 && N > 0 && (!G || !ea || !eb || !C || ldc < N))) return BGP_E_ARG;
    return oz2_crt(ctx, G, M, N, ea, eb, alpha, C, ldc, (cudaStream_t)stream);
}

int64_t bgp_oz2_gemm_work_bytes(int64_t M, int64_t N, int64_t K) { return (M <= 0 || N <= 0 || K <= 0) ? 0 : oz2_gemm_work_bytes(M, N, K); }

int bgp_oz2_gemm(bgp_ctx* c, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double* C, int64_t ldc, void* work, int64_t work_bytes, void* stream) {
    CTX_OR_FAIL(c);
    if (M < 0 || N < 0 || K <= 0 || K % 64 || M > INT_MAX || N > INT_MAX) return BGP_E_ARG;
    if (M == 0 || N == 0) return 0;
    if (!A || !B || !C || !work || lda < K || ldb < K || ldc < N || work_bytes < oz2_gemm_work_bytes(M, N, K) || ((uintptr_t)work & 255))
        return BGP_E_ARG;
    return oz2_gemm(ctx, A, M, lda, B, N, ldb, K, alpha, C, ldc, work, (cudaStream_t)stream);
}

int bgp_fault_eval(bgp_ctx* c, const double* r0, const double* r0var, int64_t M, int64_t C, int64_t ld, double band, double threshold,
                   double* p_outside, double* p_above, double* p_below, double* r0_mean, double* p_threshold, double* cells_var,
                   double* weakest_link, void* stream) {
    CTX_OR_FAIL(c);
    if (M < 0 || C < 2 || C > 16 || ld < C) return BGP_E_ARG;
    if (M == 0) return 0;
    if (!r0 || !r0var || !p_outside || !p_above || !p_below || !r0_mean || !p_threshold || !cells_var || !weakest_link) return BGP_E_ARG;
    return fault_eval(ctx, r0, r0var, M, C, ld, band, threshold, p_outside, p_above, p_below, r0_mean, p_threshold, cells_var, weakest_link,
                      (cudaStream_t)stream);
}

int bgp_potri(bgp_ctx* c, double* L, int64_t n, int64_t ldl, const double* dinv, double* work, int64_t ldw,
              void* stream) {
    CTX_OR_FAIL(c);
    if (n == 0) return 0;
    if (!L || !dinv || !work || n < 0 || ldl < n || ldw < n || n > INT_MAX) return BGP_E_ARG;
    return potri(ctx, L, n, ldl, dinv, work, ldw, (cudaStream_t)stream);
}

int bgp_lml_grad(bgp_ctx* c, const bgp_kernel_spec* spec, const double* X, int64_t n, int64_t ldx, const double* Kinv,
                 int64_t ldk, const double* alpha, double* grad_dev, void* stream) {
    CTX_OR_FAIL(c);
    if (!spec || !grad_dev || n < 0 || (n > 0 && (!X || !Kinv || !alpha || ldk < n))) return BGP_E_ARG;
    return lml_grad(ctx, spec, X, n, ldx, Kinv, ldk, alpha, grad_dev, (cudaStream_t)stream);
}

}  // extern "C"
