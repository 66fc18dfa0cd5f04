// K5 / K6 / K8: the HBM-bound O(N^2) tail of the fit -- alpha = L^-T L^-1 y by two blocked triangular sweeps
// (each reads L once), the predictive mean / variance reductions and the dot product for the LML.
//
// Forward sweep, block k (128 rows):  x_k = inv(L_kk) b_k ;  b[k+1:] -= L[k+1:, k] x_k
// One launch per block: every CTA updates 128 rows of b with coalesced row-dot-products against x_k; CTA 0 owns
// the rows of block k+1 and, once they are final, also applies inv(L_{k+1,k+1}) so the next launch finds x_{k+1}
// ready.  Backward sweep is the mirror image on L^T (column sums over the row-block of L, coalesced along columns).
#include "common.cuh"

namespace bgp {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// x[0:nb] = Dinv[0:nb,0:nb] * b[0:nb]   (first block of the forward sweep), one CTA of 256 threads
__global__ void __launch_bounds__(256) trsv_first_kernel(const double* __restrict__ dinv, const double* __restrict__ b,
                                                         double* __restrict__ x, int nb) {
    __shared__ double sb[LEAF];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < LEAF) sb[tid] = tid < nb ? b[tid] : 0.0;
    __syncthreads();
    for (int r = warp; r < nb; r += 8) {
        double s = 0.0;
        for (int c = lane; c <= r; c += 32) s = fma(dinv[r * LEAF + c], sb[c], s);
        s = warp_sum(s);
        if (lane == 0) x[r] = s;
    }
}

// RW rows per warp, 128 columns: issue all loads of the RW x 128 slab (4 per lane and row) before anything is reduced
template <int RW>
__device__ __forceinline__ void rows_load(const double* __restrict__ base, int64_t ld, int nrows, int ncols, int lane,
                                          double (&a)[RW][4]) {
#pragma unroll
    for (int u = 0; u < RW; u++) {
        const double* lp = base + (int64_t)u * ld;
        const bool rok = u < nrows;
        a[u][0] = (rok && lane < ncols) ? lp[lane] : 0.0;
        a[u][1] = (rok && lane + 32 < ncols) ? lp[lane + 32] : 0.0;
        a[u][2] = (rok && lane + 64 < ncols) ? lp[lane + 64] : 0.0;
        a[u][3] = (rok && lane + 96 < ncols) ? lp[lane + 96] : 0.0;
    }
}
// lane u < RW returns the dot product of row u with x (x0..x3 = this lane's four entries of x)
template <int RW>
__device__ __forceinline__ double rows_reduce(const double (&a)[RW][4], double x0, double x1, double x2, double x3, int lane) {
    double mine = 0.0;
#pragma unroll
    for (int u = 0; u < RW; u++) {
        const double s = warp_sum(a[u][0] * x0 + a[u][1] * x1 + a[u][2] * x2 + a[u][3] * x3);
        if (lane == u) mine = s;
    }
    return mine;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The sweeps are chains of ~N/128 dependent launches.  Steps after the first of a sweep are launched with programmatic
// dependent launch (common.cuh): a step's CTAs become resident while earlier steps still run, pull their slab of L into
// registers (L is not written during a sweep) and the next diagonal-block inverse towards L2, and only then wait for the
// predecessor -- launch latency and the DRAM latency of the slab leave the dependent chain.
//
// forward step for block k (rows k0..k0+nbk-1 solved, x holds x_k at x[k0..]):
//   rows i >= k0+nbk:  b[i] -= L[i, k0:k0+nbk] . x_k ;  CTA 0 then x[k0+nbk ..] = Dinv_{k+1} * b[k0+nbk ..]
// 16 warps x 8 rows = 128 rows per CTA.
constexpr int TRSV_RW = 8;
__global__ void __launch_bounds__(512)
trsv_fwd_step_kernel(const double* __restrict__ L, int64_t ldl, int64_t n, int64_t k0, int nbk,
                     const double* __restrict__ dinv_next, double* b, double* x) {
    __shared__ double sx[LEAF];
    __shared__ double sb[LEAF];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t r0 = k0 + nbk + (int64_t)blockIdx.x * LEAF;
    const int64_t wr0 = r0 + warp * TRSV_RW;
    const int nrows = (int)max((int64_t)0, min((int64_t)TRSV_RW, n - wr0));
    pdl_launch_dependents();
    double a[TRSV_RW][4];
    rows_load<TRSV_RW>(L + wr0 * ldl + k0, ldl, nrows, nbk, lane, a);
    if (blockIdx.x == 0) {
        const char* dp = reinterpret_cast<const char*>(dinv_next);
        for (int i = tid; i < LEAF * LEAF * 8 / 128; i += 512) prefetch_l2(dp + (size_t)i * 128);
    }
    pdl_wait();
    if (tid < LEAF) sx[tid] = tid < nbk ? x[k0 + tid] : 0.0;
    __syncthreads();
    const double mine = rows_reduce<TRSV_RW>(a, sx[lane], sx[lane + 32], sx[lane + 64], sx[lane + 96], lane);
    if (lane < nrows) {
        const double nb_ = b[wr0 + lane] - mine;
        b[wr0 + lane] = nb_;
        if (blockIdx.x == 0) sb[warp * TRSV_RW + lane] = nb_;
    }
    if (blockIdx.x != 0) return;
    __syncthreads();
    // x_{k+1} = Dinv_{k+1} * b_{k+1}: again 8 rows per warp
    const int nnext = (int)min((int64_t)LEAF, n - r0);
    const int rr0 = warp * TRSV_RW;
    const int nr = max(0, min(TRSV_RW, nnext - rr0));
    double t[TRSV_RW][4];
    rows_load<TRSV_RW>(dinv_next + rr0 * LEAF, LEAF, nr, nnext, lane, t);
    const double m2 = rows_reduce<TRSV_RW>(t, lane < nnext ? sb[lane] : 0.0, lane + 32 < nnext ? sb[lane + 32] : 0.0,
                                           lane + 64 < nnext ? sb[lane + 64] : 0.0, lane + 96 < nnext ? sb[lane + 96] : 0.0, lane);
    if (lane < nr) x[r0 + rr0 + lane] = m2;
}

// x[k0 + i] = sum_r Dinv[r][i] * b[k0 + r]   (last block of the backward sweep = first to be solved)
__global__ void __launch_bounds__(128) trsv_last_kernel(const double* __restrict__ dinv, const double* __restrict__ b,
                                                        double* __restrict__ x, int nb) {
    __shared__ double sb[LEAF];
    const int tid = threadIdx.x;
    sb[tid] = tid < nb ? b[tid] : 0.0;
    __syncthreads();
    if (tid < nb) {
        double s = 0.0;
        for (int r = tid; r < nb; r++) s = fma(dinv[r * LEAF + tid], sb[r], s);
        x[tid] = s;
    }
}

// backward step for block k (x_k at x[k0..k0+nbk-1] solved): columns c < k0:  b[c] -= sum_r L[k0+r][c] x_k[r];
// CTA 0 owns the 128 columns just left of k0 and then applies inv(L_{k-1,k-1})^T.
// 512 threads = 128 columns x 4 row-quarters (32 independent loads per thread, all issued up front).
__global__ void __launch_bounds__(512)
trsv_bwd_step_kernel(const double* __restrict__ L, int64_t ldl, int64_t k0, int nbk,
                     const double* __restrict__ dinv_prev, double* b, double* x) {
    __shared__ double sx[LEAF];
    __shared__ double part[4][LEAF];
    __shared__ double sb[LEAF];
    const int tid = threadIdx.x;
    const int cl = tid & (LEAF - 1), q = tid >> 7;
    const int64_t c = k0 - LEAF - (int64_t)blockIdx.x * LEAF + cl;   // CTA 0: columns k0-128 .. k0-1 (always >= 0)
    pdl_launch_dependents();
    double v[32];
    {
        const double* lp = L + (k0 + q * 32) * ldl + c;
#pragma unroll
        for (int r = 0; r < 32; r++) v[r] = (q * 32 + r < nbk) ? lp[(int64_t)r * ldl] : 0.0;
    }
    if (blockIdx.x == 0) {
        const char* dp = reinterpret_cast<const char*>(dinv_prev);
        for (int i = tid; i < LEAF * LEAF * 8 / 128; i += 512) prefetch_l2(dp + (size_t)i * 128);
    }
    pdl_wait();
    if (tid < LEAF) sx[tid] = tid < nbk ? x[k0 + tid] : 0.0;
    __syncthreads();
    {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < 32; r += 2) { s0 = fma(v[r], sx[q * 32 + r], s0); s1 = fma(v[r + 1], sx[q * 32 + r + 1], s1); }
        part[q][cl] = s0 + s1;
    }
    __syncthreads();
    double nbv = 0.0;
    if (q == 0) {
        nbv = b[c] - ((part[0][cl] + part[1][cl]) + (part[2][cl] + part[3][cl]));
        b[c] = nbv;
    }
    if (blockIdx.x != 0) return;
    if (q == 0) sb[cl] = nbv;
    __syncthreads();
    // x_{k-1}[i] = sum_{r >= i} DinvPrev[r][i] * b_{k-1}[r]   (block k-1 is always a full 128 block; zeros above diag)
    {
        const double* dp = dinv_prev + (q * 32) * LEAF + cl;
        double v[32];
#pragma unroll
        for (int r = 0; r < 32; r++) v[r] = dp[r * LEAF];
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < 32; r += 2) { s0 = fma(v[r], sb[q * 32 + r], s0); s1 = fma(v[r + 1], sb[q * 32 + r + 1], s1); }
        part[q][cl] = s0 + s1;
    }
    __syncthreads();
    if (q == 0) x[k0 - LEAF + cl] = (part[0][cl] + part[1][cl]) + (part[2][cl] + part[3][cl]);
}

int trsv_lower(Ctx* ctx, const double* L, int64_t n, int64_t ldl, const double* dinv, double* b, double* x,
               cudaStream_t st) {
    // solves L x = b; b is destroyed (holds the updated right-hand side)
    const int64_t nblk = (n + LEAF - 1) / LEAF;
    trsv_first_kernel<<<1, 256, 0, st>>>(dinv, b, x, (int)min((int64_t)LEAF, n));
    BGP_LAUNCH_OK(ctx);
    for (int64_t k = 0; k + 1 < nblk; k++) {
        const int64_t k0 = k * LEAF;
        const int64_t rows = n - k0 - LEAF;
        const unsigned grid = (unsigned)((rows + LEAF - 1) / LEAF);
        // the first step of the sweep is serialised normally: whatever produced L must be complete before a slab is read
        BGP_CUDA_OK(launch_pdl(ctx->pdl != 0 && k > 0, trsv_fwd_step_kernel, dim3(grid), dim3(512), 0, st, L, ldl, n, k0, (int)LEAF,
                               dinv + (k + 1) * (int64_t)LEAF * LEAF, b, x));
        BGP_LAUNCH_OK(ctx);
    }
    return 0;
}

int trsv_lower_t(Ctx* ctx, const double* L, int64_t n, int64_t ldl, const double* dinv, double* b, double* x,
                 cudaStream_t st) {
    // solves L^T x = b; b is destroyed
    const int64_t nblk = (n + LEAF - 1) / LEAF;
    const int64_t kl = (nblk - 1) * LEAF;
    const int nbl = (int)(n - kl);
    trsv_last_kernel<<<1, 128, 0, st>>>(dinv + (nblk - 1) * (int64_t)LEAF * LEAF, b + kl, x + kl, nbl);
    BGP_LAUNCH_OK(ctx);
    for (int64_t k = nblk - 1; k >= 1; k--) {
        const int64_t k0 = k * LEAF;
        const int nbk = (int)min((int64_t)LEAF, n - k0);
        const unsigned grid = (unsigned)(k0 / LEAF);
        BGP_CUDA_OK(launch_pdl(ctx->pdl != 0 && k < nblk - 1, trsv_bwd_step_kernel, dim3(grid), dim3(512), 0, st, L, ldl, k0, nbk,
                               dinv + (k - 1) * (int64_t)LEAF * LEAF, b, x));
        BGP_LAUNCH_OK(ctx);
    }
    return 0;
}

// mean[i] = Kq[i,:] . alpha ;  var[i] = max(kdiag[i] - |V[i,:]|^2, min_var).  One CTA per query row.
__global__ void __launch_bounds__(256)
predict_tail_kernel(int64_t n, const double* __restrict__ Kq, int64_t ldk, const double* __restrict__ alpha,
                    const double* __restrict__ V, int64_t ldv, const double* __restrict__ kdiag, double min_var,
                    double* __restrict__ mean, double* __restrict__ var) {
    __shared__ double red[2][8];
    const int64_t i = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double sm = 0.0, sv = 0.0;
    if (mean) {
        const double* kp = Kq + i * ldk;
        for (int64_t j = tid; j < n; j += 256) sm = fma(kp[j], alpha[j], sm);
    }
    if (var) {
        const double* vp = V + i * ldv;
        for (int64_t j = tid; j < n; j += 256) { const double v = vp[j]; sv = fma(v, v, sv); }
    }
    sm = warp_sum(sm);
    sv = warp_sum(sv);
    if (lane == 0) { red[0][warp] = sm; red[1][warp] = sv; }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; w++) { a += red[0][w]; b += red[1][w]; }
        if (mean) mean[i] = a;
        if (var) var[i] = fmax(kdiag[i] - b, min_var);
    }
}

// deterministic two-stage dot product: partial[blockIdx] then a single-CTA finish into out[0]
__global__ void __launch_bounds__(256) dot_partial_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                          int64_t n, double* __restrict__ partial) {
    __shared__ double red[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double s = 0.0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + tid; j < n; j += (int64_t)gridDim.x * 256) s = fma(a[j], b[j], s);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < 8; w++) t += red[w]; partial[blockIdx.x] = t; }
}
__global__ void dot_finish_kernel(const double* __restrict__ partial, int np, double* __restrict__ out) {
    if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < np; i++) t += partial[i]; out[0] = t; }
}

int dot(Ctx* ctx, const double* a, const double* b, int64_t n, double* partial /*>=128*/, double* out, cudaStream_t st) {
    const int np = (int)min((int64_t)128, (n + 255) / 256 > 0 ? (n + 255) / 256 : 1);
    dot_partial_kernel<<<np, 256, 0, st>>>(a, b, n, partial);
    BGP_LAUNCH_OK(ctx);
    dot_finish_kernel<<<1, 32, 0, st>>>(partial, np, out);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

// y[c] += alpha * sum_r A[r][c] v[r]   (A [rows, cols] row-major): 128 columns x 4 row-quarters per CTA
__global__ void __launch_bounds__(512)
gemv_t_kernel(const double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda, const double* __restrict__ v,
              double* __restrict__ y, double alpha) {
    __shared__ double part[4][LEAF];
    const int tid = threadIdx.x, cl = tid & (LEAF - 1), q = tid >> 7;
    const int64_t c = (int64_t)blockIdx.x * LEAF + cl;
    const int64_t rq = (rows + 3) / 4, r0 = q * rq, r1 = min(rows, r0 + rq);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (c < cols) {
        const double* ap = A + c;
        int64_t r = r0;
        for (; r + 3 < r1; r += 4) {
            const double a0 = ap[r * lda], a1 = ap[(r + 1) * lda], a2 = ap[(r + 2) * lda], a3 = ap[(r + 3) * lda];
            s0 = fma(a0, v[r], s0); s1 = fma(a1, v[r + 1], s1); s2 = fma(a2, v[r + 2], s2); s3 = fma(a3, v[r + 3], s3);
        }
        for (; r < r1; r++) s0 = fma(ap[r * lda], v[r], s0);
    }
    part[q][cl] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (q == 0 && c < cols) y[c] += alpha * ((part[0][cl] + part[1][cl]) + (part[2][cl] + part[3][cl]));
}

int gemv_t(Ctx* ctx, const double* A, int64_t rows, int64_t cols, int64_t lda, const double* v, double* y, double alpha,
           cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    gemv_t_kernel<<<(unsigned)((cols + LEAF - 1) / LEAF), 512, 0, st>>>(A, rows, cols, lda, v, y, alpha);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

// out[i] = (accumulate ? out[i] : 0) + sum_j V[i][j]^2
__global__ void __launch_bounds__(256)
rowsumsq_kernel(const double* __restrict__ V, int64_t n, int64_t ldv, double* __restrict__ out, int accumulate) {
    __shared__ double red[8];
    const int64_t i = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double* vp = V + i * ldv;
    double s = 0.0;
    for (int64_t j = tid; j < n; j += 256) { const double v = vp[j]; s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += red[w];
        out[i] = (accumulate ? out[i] : 0.0) + t;
    }
}

int rowsumsq(Ctx* ctx, const double* V, int64_t m, int64_t n, int64_t ldv, double* out, int accumulate, cudaStream_t st) {
    if (m <= 0) return 0;
    rowsumsq_kernel<<<(unsigned)m, 256, 0, st>>>(V, n, ldv, out, accumulate);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

int predict_tail(Ctx* ctx, int64_t m, int64_t n, const double* Kq, int64_t ldk, const double* alpha, const double* V,
                 int64_t ldv, const double* kdiag, double min_var, double* mean, double* var, cudaStream_t st) {
    if (m <= 0) return 0;
    predict_tail_kernel<<<(unsigned)m, 256, 0, st>>>(n, Kq, ldk, alpha, V, ldv, kdiag, min_var, mean, var);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

}  // namespace bgp
