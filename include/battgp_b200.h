/*
 * battgp_b200 -- C-ABI of the B200-native exact-GP engine (fp64, sm_100a).
 *
 * This is the drop-in boundary for BattGP's `full_gp` hot path.  The reference has NO FFI: its hot path is
 * Python calling GPyTorch.  Each entry point below therefore cites the reference *call site*
 * (file:line under /root/reference) whose GPyTorch work it replaces; the Python facade in
 * battgp_b200/gpytorch/ binds these with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - all matrices are fp64, ROW-MAJOR, leading dimension `ld*` in elements; device pointers come from
 *     `tensor.data_ptr()`; nothing is allocated behind the caller's back except through `bgp_ctx` (streams,
 *     events, two small device scalars) -- workspaces are passed in explicitly so torch's caching allocator
 *     owns all large memory (battgp_full.py:102-120 frees models with empty_cache()).
 *   - symmetric matrices / Cholesky factors use the LOWER triangle; the strictly upper triangle is never
 *     read and is left unspecified.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *   - return value: 0 ok; >0 LAPACK-style 1-based index of the first non-positive pivot; <0 argument or CUDA
 *     error (BGP_E_*).  No entry point synchronises the device unless stated.
 *   - no global mutable state; a `bgp_ctx` must not be used from two host threads at once.
 */
#ifndef BATTGP_B200_H
#define BATTGP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGP_VERSION 100

#define BGP_E_ARG   (-1)   /* bad argument (null pointer, negative size, ld too small, misalignment) */
#define BGP_E_CUDA  (-2)   /* a CUDA runtime call or kernel launch failed; see bgp_last_error()         */
#define BGP_E_SPEC  (-3)   /* malformed kernel spec                                                    */

/* ---- covariance description ------------------------------------------------------------------------
 * k(x,x') = sum_i outputscale_i * k_i(x[dims_i], x'[dims_i])  (+ noise on the diagonal of the train block)
 * replaces: ScaleKernel(WienerKernel[0]) + ScaleKernel(RBFKernel ARD [1,2,3])   cell_gp.py:32-36
 *           GaussianLikelihood noise                                             cell_gp.py:27
 *           ScaleKernel(RBFKernel())                                             standard_models.py:24
 * MATERN52 / PERIODIC: BASELINE.json config 3 (not in the reference; GPyTorch public formulas). */
enum { BGP_WIENER = 0, BGP_RBF = 1, BGP_MATERN52 = 2, BGP_PERIODIC = 3 };
#define BGP_MAX_TERMS 4
#define BGP_MAX_DIMS  8

typedef struct {
    int32_t type;                       /* BGP_WIENER ...                                         */
    int32_t ndims;                      /* number of active dims (WIENER: 1)                      */
    int32_t dims[BGP_MAX_DIMS];         /* column indices into X (GPyTorch active_dims)           */
    double  outputscale;
    double  lengthscale[BGP_MAX_DIMS];  /* per active dim (isotropic: repeat)                     */
    double  period[BGP_MAX_DIMS];       /* PERIODIC only                                          */
} bgp_term;

typedef struct {
    int32_t  nterms;
    int32_t  _pad;
    bgp_term terms[BGP_MAX_TERMS];
    double   noise;                     /* sigma_n^2, added where global row == col (train block) */
} bgp_kernel_spec;

/* number of hyper-parameter gradient slots written by bgp_lml_grad for a spec:
 * [noise, then per term: outputscale, lengthscale[ndims], (period[ndims] if PERIODIC)] */
int bgp_grad_slots(const bgp_kernel_spec* spec);

/* ---- context ----------------------------------------------------------------------------------------*/
typedef struct bgp_ctx bgp_ctx;
int  bgp_version(void);
const char* bgp_last_error(void);                         /* thread-local text of the last BGP_E_CUDA */
int  bgp_ctx_create(int device, bgp_ctx** out);           /* creates the panel (high-priority) stream + events */
void bgp_ctx_destroy(bgp_ctx* ctx);
/* tuning knobs: "nb" uniform outer panel width (multiple of 128; 0 = automatic schedule: the width follows the rows still to
 * be factorised -- 512 below "sched_t1024" rows, 1024 from there, 2048 from "sched_t2048" and 4096 from "sched_t4096" on the
 * int8 path, "sched_w0"/"sched_w1" cap the first two panels), "lookahead" 0/1, "ozaki" 0/1 (trailing updates on the int8
 * tcgen05 path, fp64-accurate), "oz_tpc"/"oz_tpc_gemm"/"oz_cluster"/"oz_group" (tiles per CTA inside bgp_potrf / in stand-alone bgp_oz_gemm calls, 0 = one persistent CTA per SM; cluster size; tile rows per raster group of that kernel), "gemm_cfg" (probing),
 * "pdl" 0/1 (programmatic dependent launch of the dependent-kernel chains: triangular sweeps, leaf + small GEMMs), "trace" 0/1 (bgp_potrf prints a per-panel event timeline to stderr; diagnostics). returns 0 or BGP_E_ARG */
int  bgp_ctx_set(bgp_ctx* ctx, const char* key, int value);
/* Optional scratch for bgp_potrf's int8/tcgen05 trailing updates ("ozaki" knob, csrc/ozaki.cu): the caller (torch)
 * owns the memory; bgp_potrf uses the path only when at least bgp_potrf_workspace_bytes(ctx, n) bytes are set. */
int64_t bgp_potrf_workspace_bytes(const bgp_ctx* ctx, int64_t n);
/* host-only introspection (no CUDA call): the panel boundaries bgp_potrf_aug uses for a [rows, n] factorisation with the
 * default schedule knobs ("ozaki" 0/1, "nb" 0 = automatic).  Writes min(npanels + 1, cap) column offsets, returns npanels. */
int  bgp_panel_schedule(int64_t rows, int64_t n, int ozaki, int nb, int64_t* starts, int cap);
int  bgp_ctx_set_workspace(bgp_ctx* ctx, void* ptr, int64_t bytes);
/* counts kernels launched through this context since creation (bench.py's gpu_launches) */
int64_t bgp_ctx_launches(const bgp_ctx* ctx);
/* Measurement hook (bench.py's roofline): while enabled, bgp_potrf / bgp_potrf_aug bracket every trailing-update launch of the
 * int8/tcgen05 kernel (csrc/ozaki.cu oz_mma_kernel) with timed CUDA events on the stream it is launched on.
 * bgp_ctx_kernel_profile_read waits for the recorded launches, returns their summed duration (ms), summed algorithmic flop
 * (2 x computed lower-triangle area x panel width) and count, and clears the record.  Enabling also clears it. */
int bgp_ctx_kernel_profile(bgp_ctx* ctx, int enable);
int bgp_ctx_kernel_profile_read(bgp_ctx* ctx, double* ms, double* flop, int64_t* launches);

/* ---- K1+K2+K3: fused covariance build -------------------------------------------------------------
 * replaces WienerKernel.forward (wiener_kernel.py:10-32), RBFKernel/ScaleKernel/+ (cell_gp.py:33-36) and the
 * noise add (cell_gp.py:27) in ONE pass over the output.
 * X1 [n1,ldx1], X2 [n2,ldx2] row-major device arrays.  out [n1, ldo].
 *   symmetric != 0 : X2 is ignored (X2 := X1, n2 := n1); only tiles touching the LOWER triangle are written
 *                    and spec->noise is added on the diagonal.   (train block, ExactGP.__call__)
 *   symmetric == 0 : full n1 x n2 cross-covariance, no noise.     (kernel.forward(x1, x2): recursive_gp.py:52-57,
 *                    spatiotemporal_gp.py:157-162, K_*N at predict battcellgp_full.py:173) */
int bgp_cov_build(bgp_ctx* ctx, const bgp_kernel_spec* spec,
                  const double* X1, int64_t n1, int64_t ldx1,
                  const double* X2, int64_t n2, int64_t ldx2,
                  double* out, int64_t ldo, int symmetric, void* stream);

/* diag k(x,x) (no noise): kernel.forward(x, x, diag=True) / the k_** term of the predictive variance */
int bgp_cov_diag(bgp_ctx* ctx, const bgp_kernel_spec* spec, const double* X, int64_t n, int64_t ldx,
                 double* out, void* stream);

/* ---- dense building block (exposed for tests/bench; also the yardstick vs cuBLAS dgemm) ------------
 * C[M,N] = alpha * A[M,K] * B[N,K]^T + beta * C      (FP64 DMMA tensor pipe)
 * tri != 0: only elements with (col + coff) <= (row + roff) are computed/stored (lower-triangular SYRK). */
int bgp_gemm_nt(bgp_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha,
                const double* A, int64_t lda, const double* B, int64_t ldb,
                double beta, double* C, int64_t ldc, int tri, int64_t roff, int64_t coff, void* stream);

/* EXPERIMENTAL: same contraction with C += alpha * A * B^T evaluated by int8 slicing on the tcgen05 tensor cores
 * (Ozaki scheme, fp64-accurate; csrc/ozaki.cu).  K must be a multiple of 64 and <= 16384 (BGP_E_ARG otherwise).  work: device scratch of
 * bgp_gemm_nt_i8_work_bytes(M, N, K) bytes, 256-byte aligned. */
int64_t bgp_gemm_nt_i8_work_bytes(int64_t M, int64_t N, int64_t K);
int bgp_gemm_nt_i8(bgp_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha,
                   const double* A, int64_t lda, const double* B, int64_t ldb,
                   double* C, int64_t ldc, int tri, int64_t roff, int64_t coff,
                   void* work, int64_t work_bytes, void* stream);

/* The two halves of bgp_gemm_nt_i8, for callers that re-use one sliced panel for many products (sharded trailing update):
 * bgp_oz_slice: rows x K fp64 (K % 64 == 0, K <= 16384: the int32 accumulators of bgp_oz_gemm hold 7 K 2^14) -> 7 int8 digit planes + per-row scales in buf (bgp_oz_slice_bytes, 256-B aligned).
 * bgp_oz_gemm : C[M,N] += alpha * A B^T with A = rows arow0.. (multiple of 128) of bufA, B = rows brow0.. (multiple of 64) of bufB. */
int64_t bgp_oz_slice_bytes(int64_t rows, int64_t K);
int bgp_oz_slice(bgp_ctx* ctx, const double* P, int64_t rows, int64_t K, int64_t ld, void* buf, int64_t buf_bytes, void* stream);
/* like bgp_oz_slice, but logical row block b (blkrows rows each) is read from source block blkmap[b] (device int32 array):
 * slices a rank-major all-gather buffer directly into stripe order (battgp_b200/sharded.py). */
int bgp_oz_slice_gather(bgp_ctx* ctx, const double* P, int64_t rows, int64_t K, int64_t ld, const int32_t* blkmap, int64_t blkrows,
                        void* buf, int64_t buf_bytes, void* stream);
int bgp_oz_gemm(bgp_ctx* ctx, const void* bufA, int64_t rowsA, int64_t arow0, const void* bufB, int64_t rowsB, int64_t brow0,
                int64_t M, int64_t N, int64_t K, double alpha, double* C, int64_t ldc, int tri, int64_t roff, int64_t coff,
                void* stream);

/* Components of the modular (Chinese-remainder) int8 emulation -- 16 products instead of 36 (csrc/ozaki2.cu, DESIGN.md section 5;
 * design groundwork: the tensor-core kernel between them is next, bgp_potrf does not use them yet).
 * bgp_oz2_residues: A [rows, K] fp64 (ld) -> residues [16][rows][K] int8 (symmetric residues of trunc(a 2^(55-e_row)) for the
 *   moduli 256,255,253,251,247,241,239,233,229,227,223,217,211,199,197,193) and expo [rows] (2^e >= 2 max|row|; INT_MIN for a
 *   row holding NaN/Inf).
 * bgp_oz2_crt: G [16][M][N] int32 (G_j = A_j B_j^T) -> C[i][j] += alpha * 2^(ea_i + eb_j - 110) * C'_ij, C' reconstructed exactly. */
int bgp_oz2_residues(bgp_ctx* ctx, const double* A, int64_t rows, int64_t K, int64_t ld, int8_t* residues, int32_t* expo, void* stream);
int bgp_oz2_crt(bgp_ctx* ctx, const int32_t* G, int64_t M, int64_t N, const int32_t* ea, const int32_t* eb, double alpha, double* C,
                int64_t ldc, void* stream);
/* EXPERIMENTAL: C[M,N] += alpha A B^T through the modular scheme end to end (first, multicast-free form of its tensor-core
 * kernel, csrc/next/ozaki2_mma.cu): residues -> 16 int8 products (128x256 tiles, two moduli per TMEM pass) -> reconstruction.
 * K % 64 == 0; work: bgp_oz2_gemm_work_bytes bytes, 256-byte aligned. */
int64_t bgp_oz2_gemm_work_bytes(int64_t M, int64_t N, int64_t K);
int bgp_oz2_gemm(bgp_ctx* ctx, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double* C, int64_t ldc, void* work, int64_t work_bytes, void* stream);

/* ---- K4: Cholesky -------------------------------------------------------------------------------------
 * replaces torch.linalg.cholesky_ex inside GPyTorch's psd_safe_cholesky, reached from ExactGP.__call__
 * (battcellgp_full.py:173, standard_models.py:41) and ExactMarginalLogLikelihood (training.py:40).
 * A [n, lda] lower triangle in, L out (in place).  dinv: workspace of bgp_potrf_dinv_elems(n) doubles that
 * receives the inverses of the 128x128 diagonal blocks of L (needed by the solves below).
 * Synchronises `stream` at the end (it has to return info).  logdet (host, may be NULL) = 2*sum(log L_ii). */
int64_t bgp_potrf_dinv_elems(int64_t n);
int bgp_potrf(bgp_ctx* ctx, double* A, int64_t n, int64_t lda, double* dinv, double* logdet_host, void* stream);
/* Same, for an augmented (n + mx) x n matrix: rows n.. hold mx extra right-hand-side ROWS X (e.g. K_*N of the M = 300
 * query points, battgp_full.py:98) which leave as X L^-T -- the predictive-variance solve of bgp_trsm_rlt fused into the
 * panel solves / trailing updates of the factorisation.  Workspace: bgp_potrf_workspace_bytes(ctx, n + mx). */
int bgp_potrf_aug(bgp_ctx* ctx, double* A, int64_t n, int64_t mx, int64_t lda, double* dinv, double* logdet_host, void* stream);

/* Fully asynchronous form for SMALL systems (the reference's default N = 1000 per cell, 9 GPs per battery: battgp_full.py:41-60,
 * config.py:29): single stream, no look-ahead, info (INT_MAX = ok, else 1-based failing pivot) and logdet stay on the device,
 * no host synchronisation -- several independent GPs can be in flight on different streams (one bgp_ctx each) or be captured in
 * a CUDA graph (battgp_b200/batch.py).  bgp_lml_dev is bgp_lml without the read-back: lml_dev <- LML. */
int bgp_potrf_async(bgp_ctx* ctx, double* A, int64_t n, int64_t mx, int64_t lda, double* dinv, int32_t* info_dev, double* logdet_dev,
                    void* stream);
int bgp_lml_dev(bgp_ctx* ctx, const double* z, int64_t n, const double* logdet_dev, double* lml_dev, void* stream);

/* ---- K5: alpha = K^-1 y via two triangular sweeps (HBM-bound) -------------------------------------------
 * replaces cholesky_solve for the mean cache of DefaultPredictionStrategy / inv_quad of the mll.
 * y [n] in, z = L^-1 y written to z [n] (may be NULL), alpha = L^-T z written to alpha [n]. */
int bgp_potrs_vec(bgp_ctx* ctx, const double* L, int64_t n, int64_t ldl, const double* dinv,
                  const double* y, double* z, double* alpha, void* stream);

/* ---- K7: X <- X * L^-T  (X [m, ldx] row-major, m right-hand sides stored as ROWS) -----------------------
 * replaces solve_triangular / root_inv_decomposition in the predictive covariance (battcellgp_full.py:173-180,
 * standard_models.py:43-48). */
int bgp_trsm_rlt(bgp_ctx* ctx, const double* L, int64_t n, int64_t ldl, const double* dinv,
                 double* X, int64_t m, int64_t ldx, void* stream);

/* ---- K6+K7 fused tail: mean = Kq alpha ; var = kdiag - rowsum(V*V), clamped at min_var --------------------
 * Kq [m, ldk] = K_*N (before the solve), V [m, ldv] = K_*N L^-T.  mean/var [m].  Either half may be skipped by
 * passing NULL (Kq,mean) or (V,var).  replaces MultivariateNormal.mean/.variance (battcellgp_full.py:175,180). */
int bgp_predict_tail(bgp_ctx* ctx, int64_t m, int64_t n, const double* Kq, int64_t ldk, const double* alpha,
                     const double* V, int64_t ldv, const double* kdiag, double min_var,
                     double* mean, double* var, void* stream);

/* ---- K8: LML = -0.5 z.z - 0.5 logdet - 0.5 n log(2 pi)  (z = L^-1 y) -------------------------------------
 * replaces ExactMarginalLogLikelihood.__call__ (training.py:27,40); the facade divides by n.  Synchronises. */
int bgp_lml(bgp_ctx* ctx, const double* z, int64_t n, double logdet, double* lml_host, void* stream);

/* ---- K9: analytic LML gradient ----------------------------------------------------------------------------
 * bgp_potri: L (lower, in place) -> K^-1 (lower triangle), via trtri + lauum on the DMMA pipe.
 * bgp_lml_grad: grad[slot] = 0.5 * sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij/dtheta_slot, recomputing dK tile-wise
 * from X (no N^2 temporaries).  replaces loss.backward() through cholesky (training.py:41,140).
 * grad_dev: device array of bgp_grad_slots(spec) doubles (zeroed by the call). */
int bgp_potri(bgp_ctx* ctx, double* L, int64_t n, int64_t ldl, const double* dinv, double* work, int64_t ldw,
              void* stream);
int bgp_lml_grad(bgp_ctx* ctx, const bgp_kernel_spec* spec, const double* X, int64_t n, int64_t ldx,
                 const double* Kinv, int64_t ldk, const double* alpha, double* grad_dev, void* stream);

/* ---- output side: fault evaluation of a battery on the device -----------------------------------------------
 * replaces get_fault_evaluation / calc_outside_band_probabilities(band_mean_without_eval_cell=True) / calc_over_threshold_probability /
 * calc_r0_cells_var (/root/reference/src/batt_models/fault_evaluation.py:20-104) and the weakest-link statistic
 * (fault_probabilities.py:88-95) on r0, r0var [M, ld] (M query times x C cells, 2 <= C <= 16, row-major):
 * [M, C] outputs p_outside, p_above, p_below, r0_mean (Hodges-Lehmann location of the other cells), p_threshold; [M] outputs cells_var, weakest_link. */
int bgp_fault_eval(bgp_ctx* ctx, const double* r0, const double* r0var, int64_t M, int64_t C, int64_t ld, double band, double threshold,
                   double* p_outside, double* p_above, double* p_below, double* r0_mean, double* p_threshold, double* cells_var,
                   double* weakest_link, void* stream);

/* ---- multi-GPU building blocks (block-row-cyclic sharded Cholesky; the NCCL exchange lives in Python/
 * torch.distributed, see battgp_b200/sharded.py) --------------------------------------------------------------
 * factor one nb x nb diagonal block (nb multiple of 128, <= 4096): potrf + all 128-block inverses.  async. */
int bgp_potrf_block(bgp_ctx* ctx, double* A, int64_t nb, int64_t lda, double* dinv, int32_t* info_dev,
                    double* logdet_dev, void* stream);
/* in-place single right-hand-side solve with one factored block: L x = b (trans = 0) or L^T x = b (trans != 0) */
int bgp_trsv(bgp_ctx* ctx, const double* L, int64_t n, int64_t ldl, const double* dinv, double* b, int trans, void* stream);
/* y[0:cols] += alpha * A^T v for a row-block A [rows, lda] (back-substitution contributions of a block row of L) */
int bgp_gemv_t(bgp_ctx* ctx, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* v, double* y,
               double alpha, void* stream);
/* out[i] = (accumulate ? out[i] : 0) + |V[i,:]|^2  (partial predictive-variance sums of a column shard) */
int bgp_rowsumsq(bgp_ctx* ctx, const double* V, int64_t m, int64_t n, int64_t ldv, double* out, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BATTGP_B200_H */
