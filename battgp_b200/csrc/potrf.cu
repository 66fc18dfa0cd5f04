// Blocked Cholesky (K4) and the triangular solve with many right-hand sides (K7).
//
// Storage: row-major, lower triangle.  Everything above the 128x128 leaf is expressed as NT GEMMs on the
// DMMA pipe (gemm_nt.cu):
//   potrf_rec(A)      : A11 = L11 L11^T ; A21 <- A21 L11^-T ; A22 -= A21 A21^T (tri) ; recurse on A22
//   trsm_rlt_rec(X,L) : X1 <- X1 L11^-T ; X2 -= X1 L21^T ; X2 <- X2 L22^-T
//   leaf              : one CTA factors a <=128x128 diagonal block in shared memory AND inverts it; the
//                       inverse goes to `dinv` so every leaf-level solve is a GEMM with inv(L_kk)^T.
// The outer driver (bgp_potrf in api.cu) is right-looking over NB-wide panels with a one-panel look-ahead:
// the next panel is updated and factored on a high-priority stream while the bulk of the trailing update
// runs on the caller's stream.
#include <climits>
#include "common.cuh"

namespace bgp {

constexpr int SP = LEAF + 4;    // shared-memory pitch of the leaf block: == 4 (mod 16) -> conflict-free DMMA fragment loads
constexpr int BP = 36;          // pitch of the 32-wide scratch blocks (== 4 mod 16)

__device__ __forceinline__ void dmma_leaf(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// One warp, lane i = row i of the 32x32 diagonal block at D (pitch SP): in-place lower Cholesky held in registers (column
// broadcasts by shuffle).  Writes the factor back to D (zeros above the diagonal) and 1/L_jj to rdiag[0..31].
// Returns sum_j log d_j (each lane evaluates ONE log, then a warp reduction); *bad = first failing local column + 1.
__device__ __forceinline__ double warp_potrf32(double* D, double* rdiag, double* colbuf, int lane, int* bad) {
    double a[32];
#pragma unroll
    for (int k = 0; k < 32; k++) a[k] = (k <= lane) ? D[lane * SP + k] : 0.0;
    double my_d = 1.0;
    int fail = 0;
    // The chain from one pivot to the next is kept free of shared memory: lane j+1 holds row j+1, so its own column value
    // l = a[j] rinv gives the next pivot d_{j+1} = a[j+1] - l^2 locally; that is shuffled out and the rsqrt started at once.
    // The broadcast of the whole column through shared memory (every lane needs l_k of every row k for its rank-1 update)
    // runs underneath the rsqrt; it recomputes a[j+1] on lane j+1 from the same operands, i.e. to the same bits.
    double d = shfl_d(a[0], 0);
    double rinv = rsqrt(d);
#pragma unroll
    for (int j = 0; j < 32; j++) {
        if (!(d > 0.0) && fail == 0) fail = j + 1;
        const double l = (lane == j) ? d * rinv : a[j] * rinv;
        a[j] = l;
        if (lane == j) { my_d = d; rdiag[j] = rinv; }
        double d_next = 1.0, rinv_next = 1.0;
        if (j + 1 < 32) {
            d_next = shfl_d(fma(-l, l, a[j + 1]), j + 1);
            rinv_next = rsqrt(d_next);
        }
        // broadcast the column through shared memory: every lane then reads l_k with one (conflict-free, broadcast) LDS.64
        colbuf[(j & 1) * 32 + lane] = l;
        __syncwarp();
#pragma unroll
        for (int k = j + 1; k < 32; k++) a[k] = fma(-l, colbuf[(j & 1) * 32 + k], a[k]);
        d = d_next;
        rinv = rinv_next;
    }
    *bad = fail;
#pragma unroll
    for (int k = 0; k < 32; k++) D[lane * SP + k] = (k <= lane) ? a[k] : 0.0;
    double lg = log(my_d);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lg += __shfl_xor_sync(0xffffffffu, lg, o);
    return lg;
}

// One warp: X (pitch BP, zeros above the diagonal) = inverse of the 32x32 lower-triangular factor at D, one column per lane:
//   x_i = (delta_ic - sum_{k<i} L_ik x_k) / L_ii   (x_k = 0 for k < c falls out)
__device__ __forceinline__ void warp_trtri32(const double* D, const double* rdiag, double* X, int lane) {
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; i++) {
        double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < i; k += 2) {
            s0 = fma(-D[i * SP + k], x[k], s0);
            if (k + 1 < i) s1 = fma(-D[i * SP + k + 1], x[k + 1], s1);
        }
        x[i] = (s0 + s1) * rdiag[i];
    }
#pragma unroll
    for (int i = 0; i < 32; i++) X[i * BP + lane] = (i >= lane) ? x[i] : 0.0;
}

// One thread = one row p (32 values at P, pitch SP) of the panel below the diagonal block:  p <- p * L^-T  by forward
// substitution against the factor at D (broadcast shared-memory reads):  x_j = (p_j - sum_{k<j} x_k L_jk) / L_jj
__device__ __forceinline__ void row_trsm32(double* P, const double* D, const double* rdiag) {
    double x[32];
#pragma unroll
    for (int j = 0; j < 32; j++) x[j] = P[j];
#pragma unroll
    for (int j = 0; j < 32; j++) {
        double s0 = x[j], s1 = 0.0;
#pragma unroll
        for (int k = 0; k < j; k += 2) {
            s0 = fma(-x[k], D[j * SP + k], s0);
            if (k + 1 < j) s1 = fma(-x[k + 1], D[j * SP + k + 1], s1);
        }
        x[j] = (s0 + s1) * rdiag[j];
    }
#pragma unroll
    for (int j = 0; j < 32; j++) P[j] = x[j];
}

// per-phase cycle counters of the leaf kernel (development aid: compile with -DBGP_LEAF_PROFILE, read with bgp_debug_leaf_clk)
__device__ long long g_leaf_clk[8];
#ifdef BGP_LEAF_PROFILE
#define LEAF_CLK(i) do { if (tid == 0) { long long t_ = clock64(); g_leaf_clk[i] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define LEAF_CLK(i) do { (void)t_prev; } while (0)
#endif

// One CTA (8 warps): S = lower(A[0:n,0:n]) (identity-padded to 128); S <- chol(S) by 32-wide panels:
//   warp 0 factors the diagonal 32x32 block in registers; then, concurrently, warp 0 inverts it (needed only for the
//   final block inverse) while warps 1-3 solve the rows below by per-row substitution; the trailing update is an in-smem
//   DMMA product by all warps.  A <- L; then L^-1 by block columns (DMMA products) -> dinv (dense 128x128, zeros above
//   the diagonal).  info: atomicMin of the 1-based global index of the first non-positive pivot.  logdet += log|A|.
__global__ void __launch_bounds__(256)
leaf_potrf_trtri_kernel(double* A, int64_t lda, int n, double* dinv, int32_t* info, int64_t gofs, double* logdet, int mode) {
    // mode 0: factor + full inverse.  The leaf chain (potrf_chain2) splits the two: mode 1 = factor only and store the inverses
    // of the four diagonal 32-blocks (all the next block's substitution solve needs) into dinv's diagonal sub-blocks;
    // mode 2 = load the factor and those four blocks back and complete dinv (off the critical path, on the side stream).
    extern __shared__ double sm[];
    double* S = sm;                          // [128][SP]
    double* T = S + LEAF * SP;               // [96][BP]   scratch for the inverse phase
    double* Dg = T + 96 * BP;                // [4][32][BP] inverses of the diagonal 32-blocks
    __shared__ double rdiag[32];
    __shared__ double colbuf[64];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;
    pdl_wait();
    pdl_launch_dependents();
    long long t_prev = clock64();

    {
        // 64 threads x double2 cover one 128-wide row; 4 row groups; ALL 32 loads of a thread are in flight before the first
        // store (one round trip to L2/HBM for the whole 128 x 128 block instead of four)
        const int jc = (tid & 63) * 2, rg = tid >> 6;
        const bool vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
        double2 v[32];
#pragma unroll
        for (int u = 0; u < 32; u++) {
            const int i = u * 4 + rg;
            v[u] = make_double2(0.0, 0.0);
            if (i < n && jc <= i) {
                const double* p = A + (int64_t)i * lda + jc;
                if (vec) v[u] = *reinterpret_cast<const double2*>(p);
                else { v[u].x = p[0]; v[u].y = p[1]; }
            }
        }
#pragma unroll
        for (int u = 0; u < 32; u++) {
            const int i = u * 4 + rg;
            double a0 = (jc <= i) ? v[u].x : 0.0, a1 = (jc + 1 <= i) ? v[u].y : 0.0;
            if (i >= n) { a0 = (jc == i) ? 1.0 : 0.0; a1 = (jc + 1 == i) ? 1.0 : 0.0; }
            S[i * SP + jc] = a0;
            S[i * SP + jc + 1] = a1;
        }
    }
    __syncthreads();
    LEAF_CLK(0);

    double logsum = 0.0;
    if (mode == 2) {                      // the factor is already in S: fetch the diagonal-block inverses
#pragma unroll 4
        for (int idx = tid; idx < 4 * 32 * 32; idx += 256) {
            const int b = idx >> 10, i = (idx >> 5) & 31, j = idx & 31;
            Dg[b * 32 * BP + i * BP + j] = dinv[(b * 32 + i) * LEAF + b * 32 + j];
        }
        __syncthreads();
    }
    for (int kb = 0; kb < 4 && mode != 2; kb++) {
        const int c0 = kb * 32, r0 = c0 + 32, mt = (LEAF - r0) / 8;     // mt row tiles below the diagonal block
        if (warp == 0) {
            int bad;
            logsum += warp_potrf32(S + c0 * SP + c0, rdiag, colbuf, lane, &bad);
            if (bad && lane == 0) atomicMin(info, (int32_t)min((int64_t)INT_MAX, gofs + c0 + bad));
        }
        __syncthreads();
        LEAF_CLK(1);
        if (warp == 0) {
            warp_trtri32(S + c0 * SP + c0, rdiag, Dg + kb * 32 * BP, lane);
        } else if (warp <= 3) {
            const int r = r0 + (warp - 1) * 32 + lane;
            if (r < LEAF) row_trsm32(S + r * SP + c0, S + c0 * SP + c0, rdiag);
        }
        __syncthreads();
        LEAF_CLK(2);
        // trailing update  S[i][j] -= sum_k P[i][k] P[j][k]  over lower 8x8 tiles of the remaining block
        int cnt = 0;
        for (int ti = 0; ti < mt; ti++) {
            for (int tj = 0; tj <= ti; tj++, cnt++) {
                if ((cnt & 7) != warp) continue;
                const double* ap = S + (r0 + ti * 8 + fr) * SP + c0 + fk;
                const double* bp = S + (r0 + tj * 8 + fr) * SP + c0 + fk;
                double* cp = S + (r0 + ti * 8 + fr) * SP + r0 + tj * 8 + 2 * fk;
                double n0 = 0.0, n1 = 0.0;
#pragma unroll
                for (int k4 = 0; k4 < 32; k4 += 4) dmma_leaf(n0, n1, ap[k4], bp[k4]);
                cp[0] -= n0;
                cp[1] -= n1;
            }
        }
        __syncthreads();
        LEAF_CLK(3);
    }
    if (mode != 2) {
        if (tid == 0 && logdet != nullptr) atomicAdd(logdet, logsum);
        // write L back (lower triangle only)
#pragma unroll 8
        for (int idx = tid; idx < n * LEAF; idx += 256) {
            const int i = idx >> 7, j = idx & (LEAF - 1);
            if (j <= i) A[(int64_t)i * lda + j] = S[i * SP + j];
        }
        if (mode == 1) {                  // only the diagonal 32-block inverses leave the kernel
#pragma unroll 4
            for (int idx = tid; idx < 4 * 32 * 32; idx += 256) {
                const int b = idx >> 10, i = (idx >> 5) & 31, j = idx & 31;
                dinv[(b * 32 + i) * LEAF + b * 32 + j] = Dg[b * 32 * BP + i * BP + j];
            }
            return;
        }
    }
    __syncthreads();
    LEAF_CLK(4);
    // ---- inverse: X = L^-1 in place.  Diagonal 32-blocks come from Dg; block column j (2,1,0):
    //      X[r0:, j] = -X_trail * L[r0:, j] * X_jj   with X_trail = inv(L[r0:, r0:]) already in place
#pragma unroll 8
    for (int idx = tid; idx < 4 * 32 * 32; idx += 256) {
        const int b = idx >> 10, i = (idx >> 5) & 31, j = idx & 31;
        S[(b * 32 + i) * SP + b * 32 + j] = Dg[b * 32 * BP + i * BP + j];
    }
    __syncthreads();
    for (int jb = 2; jb >= 0; jb--) {
        const int cb = jb * 32, r0 = cb + 32, mt = (LEAF - r0) / 8;
        // (i) T = X_trail * B   (NN; X_trail lower-triangular: k < (ti+1)*8)
        for (int ti = warp; ti < mt; ti += 8) {
            double acc[4][2];
#pragma unroll
            for (int t = 0; t < 4; t++) acc[t][0] = acc[t][1] = 0.0;
            const double* ap = S + (r0 + ti * 8 + fr) * SP + r0 + fk;
            const double* bp = S + (r0 + fk) * SP + cb + fr;
            for (int k4 = 0; k4 < (ti + 1) * 8; k4 += 4) {
                const double av = ap[k4];
#pragma unroll
                for (int t = 0; t < 4; t++) dmma_leaf(acc[t][0], acc[t][1], av, bp[k4 * SP + t * 8]);
            }
            double* cp = T + (ti * 8 + fr) * BP + 2 * fk;
#pragma unroll
            for (int t = 0; t < 4; t++) { cp[t * 8] = acc[t][0]; cp[t * 8 + 1] = acc[t][1]; }
        }
        __syncthreads();
        // (ii) B <- -T * X_jj   (NN; X_jj = Dg[jb], explicit zeros above its diagonal)
        for (int ti = warp; ti < mt; ti += 8) {
            double acc[4][2];
#pragma unroll
            for (int t = 0; t < 4; t++) acc[t][0] = acc[t][1] = 0.0;
            const double* ap = T + (ti * 8 + fr) * BP + fk;
            const double* bp = Dg + jb * 32 * BP + fk * BP + fr;
#pragma unroll
            for (int k4 = 0; k4 < 32; k4 += 4) {
                const double av = ap[k4];
#pragma unroll
                for (int t = 0; t < 4; t++) dmma_leaf(acc[t][0], acc[t][1], av, bp[k4 * BP + t * 8]);
            }
            double* cp = S + (r0 + ti * 8 + fr) * SP + cb + 2 * fk;
#pragma unroll
            for (int t = 0; t < 4; t++) { cp[t * 8] = -acc[t][0]; cp[t * 8 + 1] = -acc[t][1]; }
        }
        __syncthreads();
    }
    LEAF_CLK(5);
#pragma unroll 8
    for (int idx = tid; idx < LEAF * LEAF; idx += 256) {
        const int i = idx >> 7, j = idx & (LEAF - 1);
        // mode 2: the diagonal 32-blocks are already there and may be being read by the next block's solve: leave them alone
        if (mode != 2 || (i >> 5) != (j >> 5)) dinv[idx] = (j <= i) ? S[i * SP + j] : 0.0;
    }
    __syncthreads();
    LEAF_CLK(6);
}

void leaf_clk_read(long long* out) { cudaMemcpyFromSymbol(out, g_leaf_clk, sizeof(long long) * 8); long long z[8] = {0}; cudaMemcpyToSymbol(g_leaf_clk, z, sizeof(z)); }

static int launch_leaf(Ctx* ctx, double* A, int64_t lda, int n, double* dinv, int64_t gofs, cudaStream_t st, int mode = 0) {
    constexpr int SMEM = (LEAF * SP + 96 * BP + 4 * 32 * BP) * sizeof(double);
    static thread_local uint64_t attr_done = 0;
    const uint64_t bit = 1ull << (ctx->device & 63);
    if (!(attr_done & bit)) {
        BGP_CUDA_OK(cudaFuncSetAttribute(leaf_potrf_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done |= bit;
    }
    BGP_CUDA_OK(launch_pdl(ctx->pdl && ctx->pdl_chain, leaf_potrf_trtri_kernel, dim3(1), dim3(256), SMEM, st, A, lda, n, dinv, ctx->d_info, gofs,
                           ctx->d_scal, mode));
    BGP_LAUNCH_OK(ctx);
    return 0;
}

int64_t oz_slice_buffer_bytes(int64_t rows, int64_t K);
int oz_slice(Ctx*, const double*, int64_t, int64_t, int64_t, void*, cudaStream_t, const int32_t* blkmap = nullptr, int64_t blkrows = 0);
int oz_gemm(Ctx*, const void*, int64_t, int64_t, const void*, int64_t, int64_t, int64_t, int64_t, int64_t, double, double*,
            int64_t, int, int64_t, int64_t, cudaStream_t, int tiles_per_cta);

// C[M,N] += alpha * A[M,K] B[N,K]^T on the int8/tcgen05 path, K processed in chunks so that the digit planes of both
// operands fit the scratch region [ws_trsm, +ws_trsm_bytes) of the context.  Returns 1 if the product was issued (or
// failed: *rc), 0 if the caller must use the DMMA kernel (path disabled, too small to pay off, misaligned, no scratch).
//   a_upper : A[i][k] == 0 for k < i  (upper-triangular operand: rows beyond a chunk's last column contribute nothing)
//   same_ab : B is A (SYRK): sliced once, N follows the row restriction
int oz_gemm_kchunked(Ctx* ctx, const double* A, int64_t lda, int64_t M, const double* B, int64_t ldb, int64_t N, int64_t K,
                     double alpha, double* C, int64_t ldc, int tri, bool a_upper, bool same_ab, cudaStream_t st, int tiles_per_cta,
                     int* rc) {
    *rc = 0;
    if (!ctx->ozaki || !ctx->ws_trsm || K < 512 || M < 1024 || N < 256) return 0;
    if ((lda | ldb) & 1 || (((uintptr_t)A | (uintptr_t)B) & 15)) return 0;
    const int64_t Kmain = (K / 64) * 64;
    int64_t KC = 2048;
    if (KC > Kmain) KC = Kmain;
    auto need = [&](int64_t kc) {
        return ((oz_slice_buffer_bytes(M, kc) + 255) / 256) * 256 + (same_ab ? 0 : ((oz_slice_buffer_bytes(N, kc) + 255) / 256) * 256);
    };
    while (KC > 512 && need(KC) > ctx->ws_trsm_bytes) KC /= 2;
    if (need(KC) > ctx->ws_trsm_bytes) return 0;
    char* wa = reinterpret_cast<char*>(ctx->ws_trsm);
    char* wb = wa + ((oz_slice_buffer_bytes(M, KC) + 255) / 256) * 256;
    for (int64_t kc = 0; kc < Kmain; kc += KC) {
        const int64_t kw = (Kmain - kc < KC) ? Kmain - kc : KC;
        const int64_t m_eff = a_upper ? ((kc + kw < M) ? kc + kw : M) : M;
        const int64_t n_eff = same_ab ? m_eff : N;
        if ((*rc = oz_slice(ctx, A + kc, m_eff, kw, lda, wa, st))) return 1;
        if (!same_ab && (*rc = oz_slice(ctx, B + kc, N, kw, ldb, wb, st))) return 1;
        if ((*rc = oz_gemm(ctx, wa, m_eff, 0, same_ab ? wa : wb, n_eff, 0, m_eff, n_eff, kw, alpha, C, ldc, tri, 0, 0, st, tiles_per_cta)))
            return 1;
    }
    if (Kmain < K) {      // ragged tail of K (< 64 columns) on the DMMA kernel
        GemmArgs g{A + Kmain, lda, B + Kmain, ldb, C, ldc, (int)M, (int)N, (int)(K - Kmain), alpha, 1.0, tri, 0, 0};
        *rc = gemm_nt(ctx, g, st);
    }
    return 1;
}

static int oz_try_update(Ctx* ctx, const double* A, int64_t lda, int64_t M, const double* B, int64_t ldb, int64_t N, int64_t K,
                         double* C, int64_t ldc, cudaStream_t st, int* rc) {
    return oz_gemm_kchunked(ctx, A, lda, M, B, ldb, N, K, -1.0, C, ldc, 0, false, false, st, ctx->oz_tpc, rc);
}

static inline int64_t split_point(int64_t n) {
    // largest multiple of LEAF that is >= n/2 and < n
    int64_t h = ((n / 2 + LEAF - 1) / LEAF) * LEAF;
    if (h >= n) h -= LEAF;
    return h;
}

// X [m, n] <- X * L^-T, L [n, n] lower (row-major), dinv = inverses of L's 128-blocks starting at L's first block.
// L's offset inside the factor must be a multiple of 128 so that dinv blocks line up.
int trsm_rlt_rec(Ctx* ctx, const double* L, int64_t n, int64_t ldl, const double* dinv, double* X, int64_t m,
                 int64_t ldx, cudaStream_t st) {
    if (n <= 0 || m <= 0) return 0;
    if (n <= LEAF) {
        // in place: C aliases A; safe because one CTA owns all n <= 128 columns of its rows
        GemmArgs g{X, ldx, dinv, LEAF, X, ldx, (int)m, (int)n, (int)n, 1.0, 0.0, 0, 0, 0};
        return gemm_nt_cfg(ctx, g, m >= 148 * 2 * 128 ? 1 : 2, st);
    }
    const int64_t n1 = split_point(n), n2 = n - n1;
    int rc = trsm_rlt_rec(ctx, L, n1, ldl, dinv, X, m, ldx, st);
    if (rc) return rc;
    if (!oz_try_update(ctx, X, ldx, m, L + n1 * ldl, ldl, n2, n1, X + n1, ldx, st, &rc)) {
        GemmArgs g{X, ldx, L + n1 * ldl, ldl, X + n1, ldx, (int)m, (int)n2, (int)n1, -1.0, 1.0, 0, 0, 0};
        rc = gemm_nt(ctx, g, st);
    }
    if (rc) return rc;
    return trsm_rlt_rec(ctx, L + n1 * ldl + n1, n2, ldl, dinv + (n1 / LEAF) * (int64_t)LEAF * LEAF, X + n1, m, ldx, st);
}

// X [m <= 128 rows, 128] <- X L^-T by blocked forward substitution against the 128 x 128 factor L (lower, row-major) and the
// inverses Dg of its four diagonal 32-blocks (found in the diagonal sub-blocks of `dinv`):
//     X_b = (X_b - sum_{c<b} X_c L_bc^T) Dg_b^T          b = 0..3
// The serial solve between two leaves of potrf_chain2: it needs only what the factor-only leaf (mode 1) leaves behind, so the
// full block inverse stays off the critical path.  One CTA per 16 rows (8 CTAs for a block), 4 warps; warp w owns the
// 8-column tile w of the 32-column block being solved, for both 8-row tiles.
constexpr int XP = LEAF + 4;    // pitch of the X rows in shared memory
__global__ void __launch_bounds__(128)
leaf_trsm_subst_kernel(double* X, int64_t ldx, int m, const double* L, int64_t ldl, const double* dinv) {
    extern __shared__ double sm[];
    double* Ls = sm;                         // [128][SP]  lower triangle of L (upper part never read)
    double* Xs = Ls + LEAF * SP;             // [16][XP]
    double* Ys = Xs + 16 * XP;               // [16][BP]
    double* Dgs = Ys + 16 * BP;              // [4][32][BP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int fr = lane >> 2, fk = lane & 3;
    const int r0 = blockIdx.x * 16;
    pdl_wait();
    pdl_launch_dependents();
    // loads: the factor's lower triangle (double2, 64 lanes per row), the CTA's 16 rows of X, the four diagonal inverses
    for (int idx = tid; idx < LEAF * (LEAF / 2); idx += 128) {
        const int i = idx >> 6, j = (idx & 63) * 2;
        if (j <= i) {
            const double2 v = *reinterpret_cast<const double2*>(L + (int64_t)i * ldl + j);
            Ls[i * SP + j] = v.x;
            Ls[i * SP + j + 1] = v.y;
        }
    }
    for (int idx = tid; idx < 16 * (LEAF / 2); idx += 128) {
        const int i = idx >> 6, j = (idx & 63) * 2;
        double2 v = make_double2(0.0, 0.0);
        if (r0 + i < m) v = *reinterpret_cast<const double2*>(X + (int64_t)(r0 + i) * ldx + j);
        Xs[i * XP + j] = v.x;
        Xs[i * XP + j + 1] = v.y;
    }
    for (int idx = tid; idx < 4 * 32 * 32; idx += 128) {
        const int b = idx >> 10, i = (idx >> 5) & 31, j = idx & 31;
        Dgs[b * 32 * BP + i * BP + j] = dinv[(b * 32 + i) * LEAF + b * 32 + j];
    }
    __syncthreads();
    for (int b = 0; b < 4; b++) {
        const int c0 = b * 32 + warp * 8;                     // this warp's 8 columns of block b
        // (i) Y = X_b - sum_{k < 32 b} X[:, k] L[c, k]
        double acc[2][2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            acc[t][0] = Xs[(t * 8 + fr) * XP + c0 + 2 * fk];
            acc[t][1] = Xs[(t * 8 + fr) * XP + c0 + 2 * fk + 1];
        }
        double n0[2] = {0.0, 0.0}, n1[2] = {0.0, 0.0};
        for (int k4 = 0; k4 < b * 32; k4 += 4) {
            const double bv = Ls[(c0 + fr) * SP + k4 + fk];
#pragma unroll
            for (int t = 0; t < 2; t++) dmma_leaf(n0[t], n1[t], Xs[(t * 8 + fr) * XP + k4 + fk], bv);
        }
#pragma unroll
        for (int t = 0; t < 2; t++) {
            Ys[(t * 8 + fr) * BP + warp * 8 + 2 * fk] = acc[t][0] - n0[t];
            Ys[(t * 8 + fr) * BP + warp * 8 + 2 * fk + 1] = acc[t][1] - n1[t];
        }
        __syncthreads();
        // (ii) X_b = Y Dg_b^T   (Dg_b lower triangular with explicit zeros above its diagonal)
        double x0[2] = {0.0, 0.0}, x1[2] = {0.0, 0.0};
#pragma unroll
        for (int k4 = 0; k4 < 32; k4 += 4) {
            const double bv = Dgs[b * 32 * BP + (warp * 8 + fr) * BP + k4 + fk];
#pragma unroll
            for (int t = 0; t < 2; t++) dmma_leaf(x0[t], x1[t], Ys[(t * 8 + fr) * BP + k4 + fk], bv);
        }
#pragma unroll
        for (int t = 0; t < 2; t++) {
            Xs[(t * 8 + fr) * XP + c0 + 2 * fk] = x0[t];
            Xs[(t * 8 + fr) * XP + c0 + 2 * fk + 1] = x1[t];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < 16 * (LEAF / 2); idx += 128) {
        const int i = idx >> 6, j = (idx & 63) * 2;
        if (r0 + i < m) *reinterpret_cast<double2*>(X + (int64_t)(r0 + i) * ldx + j) = make_double2(Xs[i * XP + j], Xs[i * XP + j + 1]);
    }
}

static int launch_trsm_subst(Ctx* ctx, double* X, int64_t ldx, int m, const double* L, int64_t ldl, const double* dinv, cudaStream_t st) {
    constexpr int SMEM = (LEAF * SP + 16 * XP + 16 * BP + 4 * 32 * BP) * sizeof(double);
    static thread_local uint64_t attr_done = 0;
    const uint64_t bit = 1ull << (ctx->device & 63);
    if (!(attr_done & bit)) {
        BGP_CUDA_OK(cudaFuncSetAttribute(leaf_trsm_subst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done |= bit;
    }
    BGP_CUDA_OK(launch_pdl(ctx->pdl && ctx->pdl_chain, leaf_trsm_subst_kernel, dim3((m + 15) / 16), dim3(128), SMEM, st, X, ldx, m, L, ldl, dinv));
    BGP_LAUNCH_OK(ctx);
    return 0;
}

// Right-looking Cholesky of an n x n block over its 128-wide leaves on TWO streams (diagonal blocks of the look-ahead
// factorisation and whole small matrices, where the chain of dependent kernels -- not the flops -- is what takes the time).
// Per leaf k only this is serial (stream s1):
//     leaf k (factor + inverse)  ->  the 128 rows of block k+1:  X = A[k+1,k] inv(L_kk)^T  ->  A[k+1,k+1] -= X X^T  ->  leaf k+1
// Everything else that panel k touches -- the solve of the rows from block k+2 on and the rank-128 update of the blocks
// below / right of (k+1,k+1) -- runs on the context's second high-priority stream beside leaf k+1 and is only waited for
// where its results are read (the recursive form puts all of it on the critical path: 24 GEMM launches between the 8
// leaves of a 1024 block).  ev_chain: [0..1] leaf done, [2..3] X done, [4..5] side update done (by parity of k), [6] fork/join.
// chain_split (off: measured no net gain) additionally takes the block inverse off the chain: factor-only leaf (mode 1), solve
// of block k+1 by blocked substitution (leaf_trsm_subst_kernel), inverse completed on the side stream (mode 2).
static int potrf_chain2(Ctx* ctx, double* A, int64_t n, int64_t lda, double* dinv, int64_t gofs, cudaStream_t s1) {
    cudaStream_t s2 = ctx->leaf_stream;
    const int64_t nl = (n + LEAF - 1) / LEAF;
    const bool split = ctx->chain_split && ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    int rc;
    BGP_CUDA_OK(cudaEventRecord(ctx->ev_chain[6], s1));
    BGP_CUDA_OK(cudaStreamWaitEvent(s2, ctx->ev_chain[6], 0));
    for (int64_t k = 0; k < nl; k++) {
        const int64_t k0 = k * LEAF, nk = (n - k0 < LEAF) ? n - k0 : LEAF;
        double* dk = dinv + k * (int64_t)LEAF * LEAF;
        // split form: the leaf on the critical stream only factors (and leaves the four diagonal 32-block inverses); the
        // full block inverse is completed on the side stream, where its consumers are
        const bool last = (k == nl - 1);
        if ((rc = launch_leaf(ctx, A + k0 * lda + k0, lda, (int)nk, dk, gofs + k0, s1, (split && !last) ? 1 : 0))) return rc;
        if (last) break;
        const int64_t r1 = k0 + LEAF, n1 = (n - r1 < LEAF) ? n - r1 : LEAF, r2 = r1 + n1;
        BGP_CUDA_OK(cudaEventRecord(ctx->ev_chain[k & 1], s1));
        // block (k+1, k) and (k+1, k+1) carry the side stream's updates with panel k-1
        if (k >= 1) BGP_CUDA_OK(cudaStreamWaitEvent(s1, ctx->ev_chain[4 + ((k - 1) & 1)], 0));
        {
            double* X = A + r1 * lda + k0;
            if (split) {
                if ((rc = launch_trsm_subst(ctx, X, lda, (int)n1, A + k0 * lda + k0, lda, dk, s1))) return rc;
            } else {
                GemmArgs g{X, lda, dk, LEAF, X, lda, (int)n1, (int)LEAF, (int)LEAF, 1.0, 0.0, 0, 0, 0};
                if ((rc = gemm_nt_cfg(ctx, g, ctx->chain_cfg ? 8 : 2, s1))) return rc;
            }
            BGP_CUDA_OK(cudaEventRecord(ctx->ev_chain[2 + (k & 1)], s1));
            GemmArgs u{X, lda, X, lda, A + r1 * lda + r1, lda, (int)n1, (int)n1, (int)LEAF, -1.0, 1.0, 1, 0, 0};
            if ((rc = gemm_nt_cfg(ctx, u, ctx->chain_cfg ? 9 : 0, s1))) return rc;
        }
        if (split) {                                                                    // complete inv(L_kk) beside the chain
            BGP_CUDA_OK(cudaStreamWaitEvent(s2, ctx->ev_chain[k & 1], 0));
            if ((rc = launch_leaf(ctx, A + k0 * lda + k0, lda, (int)nk, dk, gofs + k0, s2, 2))) return rc;
        }
        if (r2 < n) {
            const int64_t m2 = n - r2;
            BGP_CUDA_OK(cudaStreamWaitEvent(s2, ctx->ev_chain[k & 1], 0));             // inv(L_kk) is there
            double* X2 = A + r2 * lda + k0;
            GemmArgs g{X2, lda, dk, LEAF, X2, lda, (int)m2, (int)LEAF, (int)LEAF, 1.0, 0.0, 0, 0, 0};
            if ((rc = gemm_nt_cfg(ctx, g, 2, s2))) return rc;
            BGP_CUDA_OK(cudaStreamWaitEvent(s2, ctx->ev_chain[2 + (k & 1)], 0));       // rows of block k+1 solved
            // rows r2.., columns r1..: lower part only (global column <= global row)
            GemmArgs u{X2, lda, A + r1 * lda + k0, lda, A + r2 * lda + r1, lda, (int)m2, (int)(n - r1), (int)LEAF, -1.0, 1.0, 1, r2, r1};
            if ((rc = gemm_nt(ctx, u, s2))) return rc;
        }
        BGP_CUDA_OK(cudaEventRecord(ctx->ev_chain[4 + (k & 1)], s2));
    }
    BGP_CUDA_OK(cudaEventRecord(ctx->ev_chain[6], s2));
    BGP_CUDA_OK(cudaStreamWaitEvent(s1, ctx->ev_chain[6], 0));
    return 0;
}

// In-place Cholesky of the n x n block at A; gofs = global index of its first row (for info).
int potrf_rec(Ctx* ctx, double* A, int64_t n, int64_t lda, double* dinv, int64_t gofs, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n <= LEAF) return launch_leaf(ctx, A, lda, (int)n, dinv, gofs, st);
    if (ctx->leaf_chain && ctx->leaf_stream && n <= ctx->leaf_chain_max) return potrf_chain2(ctx, A, n, lda, dinv, gofs, st);
    const int64_t n1 = split_point(n), n2 = n - n1;
    int rc = potrf_rec(ctx, A, n1, lda, dinv, gofs, st);
    if (rc) return rc;
    double* A21 = A + n1 * lda;
    double* A22 = A21 + n1;
    rc = trsm_rlt_rec(ctx, A, n1, lda, dinv, A21, n2, lda, st);
    if (rc) return rc;
    GemmArgs g{A21, lda, A21, lda, A22, lda, (int)n2, (int)n2, (int)n1, -1.0, 1.0, 1, 0, 0};
    rc = gemm_nt(ctx, g, st);
    if (rc) return rc;
    return potrf_rec(ctx, A22, n2, lda, dinv + (n1 / LEAF) * (int64_t)LEAF * LEAF, gofs + n1, st);
}

}  // namespace bgp
