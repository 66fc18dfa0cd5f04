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

constexpr int LP = LEAF + 1;   // shared-memory pitch (doubles) of the leaf block

// One CTA: S = lower(A[0:n,0:n]); S <- chol(S); A <- S; dinv <- inv(S) (dense 128x128, zeros above diag,
// identity padding when n < 128).  info: atomicMin of the 1-based global index of the first bad pivot.
__global__ void __launch_bounds__(256)
leaf_potrf_trtri_kernel(double* A, int64_t lda, int n, double* dinv, int32_t* info, int64_t gofs, double* logdet) {
    extern __shared__ double S[];          // [LEAF][LP]; strictly-upper part later holds inv(L)^T
    __shared__ double s_dinv[LEAF];        // 1 / L_ii
    const int tid = threadIdx.x;

    for (int idx = tid; idx < LEAF * LEAF; idx += 256) {
        const int i = idx >> 7, j = idx & (LEAF - 1);
        double v = 0.0;
        if (i < n) { if (j <= i) v = A[(int64_t)i * lda + j]; }
        else if (i == j) v = 1.0;
        S[i * LP + j] = v;
    }

    // right-looking, one barrier per column: iteration j first finishes (scales) column j-1, then applies the
    // rank-1 update of the still unscaled column j:  S[i][k] -= S[i][j] S[k][j] / d_j
    const int tx = tid & 31, ty = tid >> 5;
    double d_prev = 1.0;
    double logsum = 0.0;
    bool failed = false;
    for (int j = 0; j <= n; j++) {
        __syncthreads();
        if (j > 0) {
            const double r = sqrt(d_prev), rinv = 1.0 / r;
            for (int i = j - 1 + tid; i < n; i += 256) S[i * LP + (j - 1)] = (i == j - 1) ? r : S[i * LP + (j - 1)] * rinv;
        }
        if (j == n) break;
        const double d = S[j * LP + j];
        if (!(d > 0.0) && !failed) {      // also catches NaN
            failed = true;
            if (tid == 0) atomicMin(info, (int32_t)min((int64_t)INT_MAX, gofs + j + 1));
        }
        logsum += log(d);
        const double dinv_j = 1.0 / d;
        for (int i = j + 1 + ty; i < n; i += 8) {
            const double lij = S[i * LP + j] * dinv_j;
            for (int k = j + 1 + tx; k <= i; k += 32) S[i * LP + k] -= lij * S[k * LP + j];
        }
        d_prev = d;
    }
    __syncthreads();
    if (tid == 0 && logdet != nullptr) atomicAdd(logdet, logsum);   // sum log d_j = 2 sum log L_jj

    // write L back (lower triangle only)
    for (int idx = tid; idx < n * LEAF; idx += 256) {
        const int i = idx >> 7, j = idx & (LEAF - 1);
        if (j <= i) A[(int64_t)i * lda + j] = S[i * LP + j];
    }
    if (tid < LEAF) s_dinv[tid] = 1.0 / S[tid * LP + tid];
    __syncthreads();

    // inverse by forward substitution, one column per thread pair (k split by parity):
    //   X[j][j] = 1/L[j][j];  X[i][j] = -(sum_{k=j}^{i-1} L[i][k] X[k][j]) / L[i][i]
    // X[i][j] (i > j) is kept transposed in the unused strictly-upper triangle: S[j][i].
    {
        const int j = tid >> 1, par = tid & 1;
        const int jw = (tid >> 5) * 16;            // first column handled by this warp (warp-uniform loop bounds)
        for (int i = jw + 1; i < LEAF; i++) {
            double s = 0.0;
            if (i > j) {
                // k = j term uses X[j][j] = s_dinv[j]
                if (par == 0) s = S[i * LP + j] * s_dinv[j];
                for (int k = j + 1 + par; k < i; k += 2) s = fma(S[i * LP + k], S[j * LP + k], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (i > j && par == 0) S[j * LP + i] = -s * s_dinv[i];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < LEAF * LEAF; idx += 256) {
        const int i = idx >> 7, j = idx & (LEAF - 1);
        double v = 0.0;
        if (j < i) v = S[j * LP + i];
        else if (j == i) v = s_dinv[i];
        dinv[idx] = v;
    }
}

static int launch_leaf(Ctx* ctx, double* A, int64_t lda, int n, double* dinv, int64_t gofs, cudaStream_t st) {
    constexpr int SMEM = LEAF * LP * sizeof(double);
    static thread_local uint64_t attr_done = 0;
    const uint64_t bit = 1ull << (ctx->device & 63);
    if (!(attr_done & bit)) {
        BGP_CUDA_OK(cudaFuncSetAttribute(leaf_potrf_trtri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done |= bit;
    }
    leaf_potrf_trtri_kernel<<<1, 256, SMEM, st>>>(A, lda, n, dinv, ctx->d_info, gofs, ctx->d_scal);
    BGP_LAUNCH_OK(ctx);
    return 0;
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
        return gemm_nt_cfg(ctx, g, m >= 148 * 128 ? 1 : 2, st);
    }
    const int64_t n1 = split_point(n), n2 = n - n1;
    int rc = trsm_rlt_rec(ctx, L, n1, ldl, dinv, X, m, ldx, st);
    if (rc) return rc;
    GemmArgs g{X, ldx, L + n1 * ldl, ldl, X + n1, ldx, (int)m, (int)n2, (int)n1, -1.0, 1.0, 0, 0, 0};
    rc = gemm_nt(ctx, g, st);
    if (rc) return rc;
    return trsm_rlt_rec(ctx, L + n1 * ldl + n1, n2, ldl, dinv + (n1 / LEAF) * (int64_t)LEAF * LEAF, X + n1, m, ldx, st);
}

// In-place Cholesky of the n x n block at A; gofs = global index of its first row (for info).
int potrf_rec(Ctx* ctx, double* A, int64_t n, int64_t lda, double* dinv, int64_t gofs, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n <= LEAF) return launch_leaf(ctx, A, lda, (int)n, dinv, gofs, st);
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
