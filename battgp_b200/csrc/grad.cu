// K9: analytic LML gradient.
//   potri    : L -> K^-1 via Z = L^-T (upper, recursive; all contractions are NT GEMMs on the DMMA pipe with the
//              zero half of the triangular operand skipped) and K^-1 = Z Z^T (lower).
//   lml_grad : grad_theta = 0.5 sum_ij (alpha_i alpha_j - Kinv_ij) dK_ij/dtheta, with dK recomputed tile-wise from X
//              (no N^2 temporaries; one streaming read of the lower triangle of K^-1).
// Replaces autograd through torch.linalg.cholesky in loss.backward() (/root/reference/src/gp/training.py:41,140).
#include "common.cuh"
#include "kernspec.cuh"

namespace bgp {

// ------------------------------------------------------------------------------------------------ potri
__global__ void transpose_leaf_kernel(const double* __restrict__ dinv, double* __restrict__ Z, int64_t ldz, int nb) {
    // Z[i][j] = dinv[j][i] for an nb x nb block (dinv pitch LEAF); 32x32 tiles through shared memory
    __shared__ double t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 256 threads: ty in 0..7
    for (int r = ty; r < 32; r += 8) t[r][tx] = dinv[(by + r) * LEAF + bx + tx];
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = bx + r, j = by + tx;
        if (i < nb && j < nb) Z[(int64_t)i * ldz + j] = t[tx][r];
    }
}

// Z (n x n block of the work matrix, lower part already zero) <- L^-T where L is the n x n diagonal block of the factor
static int trtri_u_rec(Ctx* ctx, const double* L, int64_t n, int64_t ldl, const double* dinv, double* Z, int64_t ldz,
                       cudaStream_t st) {
    if (n <= 0) return 0;
    if (n <= LEAF) {
        transpose_leaf_kernel<<<dim3(4, 4), 256, 0, st>>>(dinv, Z, ldz, (int)n);
        BGP_LAUNCH_OK(ctx);
        return 0;
    }
    int64_t n1 = ((n / 2 + LEAF - 1) / LEAF) * LEAF;
    if (n1 >= n) n1 -= LEAF;
    const int64_t n2 = n - n1;
    int rc = trtri_u_rec(ctx, L, n1, ldl, dinv, Z, ldz, st);
    if (rc) return rc;
    // Z12 = -Z11 * L21^T   (Z11 upper-triangular: skip k < row).  Z12 is still zero (memset in potri), so the int8 path
    // can accumulate into it chunk by chunk.
    if (!oz_gemm_kchunked(ctx, Z, ldz, n1, L + n1 * ldl, ldl, n2, n1, -1.0, Z + n1, ldz, 0, true, false, st, 0, &rc)) {
        GemmArgs g{Z, ldz, L + n1 * ldl, ldl, Z + n1, ldz, (int)n1, (int)n2, (int)n1, -1.0, 0.0, 0, 0, 0, 1, 0};
        rc = gemm_nt(ctx, g, st);
    }
    if (rc) return rc;
    // Z12 <- Z12 * L22^-T
    const double* dinv2 = dinv + (n1 / LEAF) * (int64_t)LEAF * LEAF;
    if ((rc = trsm_rlt_rec(ctx, L + n1 * ldl + n1, n2, ldl, dinv2, Z + n1, n1, ldz, st))) return rc;
    return trtri_u_rec(ctx, L + n1 * ldl + n1, n2, ldl, dinv2, Z + n1 * ldz + n1, ldz, st);
}

int potri(Ctx* ctx, double* L, int64_t n, int64_t ldl, const double* dinv, double* work, int64_t ldw, cudaStream_t st) {
    BGP_CUDA_OK(cudaMemset2DAsync(work, ldw * sizeof(double), 0, n * sizeof(double), n, st));
    // the int8 path (if enabled) slices into the context workspace; everything runs on one stream, so the chunked
    // products and the TRSM updates can share the whole region
    struct Scratch {
        Ctx* c; void* p0; int64_t b0;
        explicit Scratch(Ctx* c_) : c(c_), p0(c_->ws_trsm), b0(c_->ws_trsm_bytes) { c->ws_trsm = c->ws; c->ws_trsm_bytes = c->ws_bytes; }
        ~Scratch() { c->ws_trsm = p0; c->ws_trsm_bytes = b0; }
    } scratch(ctx);
    const int tpc_saved = ctx->oz_tpc;
    ctx->oz_tpc = 0;                              // nothing runs concurrently here: fully persistent int8 kernels
    int rc = trtri_u_rec(ctx, L, n, ldl, dinv, work, ldw, st);
    if (!rc) {
        // Kinv(lower) = Z Z^T, Z upper: contributions only from k >= row
        int rc2 = 0;
        // the int8 path accumulates: clear the factor first (it is no longer needed once Z = L^-T exists)
        bool done = false;
        if (ctx->ozaki && ctx->ws && n >= 4096) {
            cudaError_t e = cudaMemset2DAsync(L, ldl * sizeof(double), 0, n * sizeof(double), n, st);
            if (e != cudaSuccess) { set_error("cudaMemset2DAsync", e); rc = BGP_E_CUDA; }
            else if (oz_gemm_kchunked(ctx, work, ldw, n, work, ldw, n, n, 1.0, L, ldl, 1, true, true, st, 0, &rc2)) { done = true; rc = rc2; }
        }
        if (!rc && !done) {
            GemmArgs g{work, ldw, work, ldw, L, ldl, (int)n, (int)n, (int)n, 1.0, 0.0, 1, 0, 0, 1, 0};
            rc = gemm_nt(ctx, g, st);
        }
    }
    ctx->oz_tpc = tpc_saved;
    return rc;
}

// ------------------------------------------------------------------------------------------------ lml_grad
constexpr int GM = 64, GN = 128;

__device__ __forceinline__ double warp_sum_g(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// accumulator layout per CTA: [0] noise, [1 + f] outputscale (at flast f), [1 + FMAX + f] lengthscale f,
// [1 + 2 FMAX + f] period f
template <int FMAX>
__global__ void __launch_bounds__(256)
lml_grad_kernel(DevSpec sp, const double* __restrict__ X, int64_t n, int64_t ldx, const double* __restrict__ Kinv,
                int64_t ldk, const double* __restrict__ alpha, double* __restrict__ partial) {
    constexpr int NA = 1 + 3 * FMAX;
    __shared__ double rowf[GM][FMAX];
    __shared__ double colf[GN][FMAX + 1];
    __shared__ double arow[GM];
    __shared__ double acol[GN];
    __shared__ double red[8][NA];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t rt = (int64_t)gridDim.x - 1 - blockIdx.x;     // heaviest tile-rows first
    const int64_t m0 = rt * GM;

    for (int idx = tid; idx < GM * FMAX; idx += 256) {
        const int r = idx / FMAX, f = idx % FMAX;
        double v = 0.0;
        if (f < sp.nfeat && m0 + r < n) v = prescale(sp, f, X[(m0 + r) * ldx + sp.fdim[f]]);
        rowf[r][f] = v;
    }
    if (tid < GM) arow[tid] = (m0 + tid < n) ? alpha[m0 + tid] : 0.0;

    double acc[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) acc[a] = 0.0;

    const int cl[4] = {2 * lane, 2 * lane + 1, 64 + 2 * lane, 64 + 2 * lane + 1};
    const int64_t last_col = min(n - 1, m0 + GM - 1);
    for (int64_t c0 = 0; c0 <= last_col; c0 += GN) {
        __syncthreads();
        for (int idx = tid; idx < GN * FMAX; idx += 256) {
            const int r = idx / FMAX, f = idx % FMAX;
            double v = 0.0;
            if (f < sp.nfeat && c0 + r < n) v = prescale(sp, f, X[(c0 + r) * ldx + sp.fdim[f]]);
            colf[r][f] = v;
        }
        if (tid < GN) acol[tid] = (c0 + tid < n) ? alpha[c0 + tid] : 0.0;
        __syncthreads();
        double cf[4][FMAX], ac[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            ac[c] = acol[cl[c]];
#pragma unroll
            for (int f = 0; f < FMAX; f++) cf[c][f] = colf[cl[c]][f];
        }
        for (int r = 0; r < 8; r++) {
            const int lr = warp * 8 + r;
            const int64_t gi = m0 + lr;
            if (gi >= n) break;
            double rf[FMAX];
#pragma unroll
            for (int f = 0; f < FMAX; f++) rf[f] = rowf[lr][f];
            const double ai = arow[lr];
            const double* krow = Kinv + gi * ldk + c0;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int64_t gj = c0 + cl[c];
                if (gj > gi) continue;                    // lower triangle only (also excludes gj >= n)
                const double w = (gj == gi) ? 0.5 : 1.0;  // 0.5 * (2 for the mirrored element)
                const double q = w * (ai * ac[c] - krow[cl[c]]);
                if (gj == gi) acc[0] += q;
                // forward pass over features: geometry factors, term values at flast
                double gfac[FMAX], hfac[FMAX], tval[FMAX], tfac[FMAX];
                double s = 0.0;
#pragma unroll
                for (int f = 0; f < FMAX; f++) {
                    gfac[f] = hfac[f] = tval[f] = tfac[f] = 0.0;
                    if (f < sp.nfeat) {
                        const int ty = sp.ftype[f];
                        const double diff = rf[f] - cf[c][f];
                        if (ty == BGP_WIENER) {
                            const double m = fmin(rf[f], cf[c][f]);
                            const double m2 = m * m;
                            tval[f] = m2 * m * (1.0 / 3.0) + fabs(diff) * m2 * 0.5;
                        } else {
                            if (ty == BGP_PERIODIC) {
                                double sn, cs;
                                sincos(diff, &sn, &cs);
                                gfac[f] = sn * sn;
                                hfac[f] = sn * cs * diff;
                                s = fma(gfac[f], sp.faux[f], s);
                            } else {
                                gfac[f] = diff * diff;
                                s += gfac[f];
                            }
                            if (sp.flast[f]) {
                                if (ty == BGP_RBF) { tval[f] = exp(-0.5 * s); tfac[f] = tval[f]; }
                                else if (ty == BGP_PERIODIC) { tval[f] = exp(-2.0 * s); tfac[f] = tval[f]; }
                                else {
                                    const double rr = sqrt(fmax(s, 1e-30));
                                    const double s5r = 2.23606797749978969641 * rr;
                                    const double e = exp(-s5r);
                                    tval[f] = (1.0 + s5r + (5.0 / 3.0) * rr * rr) * e;
                                    tfac[f] = (5.0 / 3.0) * (1.0 + s5r) * e;
                                }
                                s = 0.0;
                            }
                        }
                    }
                }
                // backward pass: distribute the term factor to the term's features
                double cur = 0.0;
#pragma unroll
                for (int f = FMAX - 1; f >= 0; f--) {
                    if (f < sp.nfeat) {
                        const int ty = sp.ftype[f];
                        if (ty == BGP_WIENER) { acc[1 + f] = fma(q, tval[f], acc[1 + f]); continue; }
                        if (sp.flast[f]) { cur = tfac[f]; acc[1 + f] = fma(q, tval[f], acc[1 + f]); }
                        if (ty == BGP_PERIODIC) {
                            acc[1 + FMAX + f] = fma(q, cur * gfac[f], acc[1 + FMAX + f]);
                            acc[1 + 2 * FMAX + f] = fma(q, cur * hfac[f], acc[1 + 2 * FMAX + f]);
                        } else {
                            acc[1 + FMAX + f] = fma(q, cur * gfac[f], acc[1 + FMAX + f]);
                        }
                    }
                }
            }
        }
    }
    // CTA reduction (fixed order -> deterministic)
#pragma unroll
    for (int a = 0; a < NA; a++) {
        const double v = warp_sum_g(acc[a]);
        if (lane == 0) red[warp][a] = v;
    }
    __syncthreads();
    if (tid < NA) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += red[w][tid];
        partial[(int64_t)blockIdx.x * NA + tid] = t;
    }
}

// grad[slot] = scale[slot] * sum_b partial[b][src[slot]]
struct SlotMap { int nslots; int src[64]; double scale[64]; };

__global__ void __launch_bounds__(256) grad_finish_kernel(const double* __restrict__ partial, int nblocks, int na, SlotMap mp,
                                                          double* __restrict__ grad) {
    __shared__ double red[8];
    const int slot = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double s = 0.0;
    for (int b = tid; b < nblocks; b += 256) s += partial[(int64_t)b * na + mp.src[slot]];
    s = warp_sum_g(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int w = 0; w < 8; w++) t += red[w]; grad[slot] = t * mp.scale[slot]; }
}

int lml_grad(Ctx* ctx, const bgp_kernel_spec* spec, const double* X, int64_t n, int64_t ldx, const double* Kinv,
             int64_t ldk, const double* alpha, double* grad, cudaStream_t st) {
    DevSpec d;
    int rc = make_devspec(spec, &d);
    if (rc) return rc;
    const int FMAX = d.nfeat <= 4 ? 4 : (d.nfeat <= 8 ? 8 : 16);
    const int NA = 1 + 3 * FMAX;
    // public slot order: noise, then per term: outputscale, lengthscale[ndims], (period[ndims] if PERIODIC)
    // factors: d k / d l_f = tfac * gfac / l (RBF, Matern; gfac = ((a-b)/l)^2);  PERIODIC: d/dl = k * 2 sin^2 / l^2,
    //          d/dp = k * (4/l) sin cos u / p with u = pi (a-b)/p
    SlotMap mp;
    int slot = 0, f = 0;
    mp.src[slot] = 0; mp.scale[slot] = 1.0; slot++;
    for (int t = 0; t < spec->nterms; t++) {
        const bgp_term& T = spec->terms[t];
        const int nd = (T.type == BGP_WIENER) ? 1 : T.ndims;
        const int flast = f + nd - 1;
        mp.src[slot] = 1 + flast; mp.scale[slot] = 1.0; slot++;
        if (T.type != BGP_WIENER) {
            for (int k = 0; k < nd; k++) {
                mp.src[slot] = 1 + FMAX + f + k;
                mp.scale[slot] = (T.type == BGP_PERIODIC)
                                     ? T.outputscale * 2.0 / (T.lengthscale[k] * T.lengthscale[k])
                                     : T.outputscale / T.lengthscale[k];
                slot++;
            }
            if (T.type == BGP_PERIODIC)
                for (int k = 0; k < nd; k++) {
                    mp.src[slot] = 1 + 2 * FMAX + f + k;
                    mp.scale[slot] = T.outputscale * 4.0 / (T.lengthscale[k] * T.period[k]);
                    slot++;
                }
        }
        f += nd;
    }
    mp.nslots = slot;
    if (n == 0) {
        BGP_CUDA_OK(cudaMemsetAsync(grad, 0, slot * sizeof(double), st));
        return 0;
    }
    const int64_t nblocks = (n + GM - 1) / GM;
    if ((size_t)nblocks * NA * sizeof(double) > SCRATCH_BYTES) return BGP_E_ARG;
    double* partial = ctx->d_scratch;
    if (FMAX == 4) lml_grad_kernel<4><<<(unsigned)nblocks, 256, 0, st>>>(d, X, n, ldx, Kinv, ldk, alpha, partial);
    else if (FMAX == 8) lml_grad_kernel<8><<<(unsigned)nblocks, 256, 0, st>>>(d, X, n, ldx, Kinv, ldk, alpha, partial);
    else lml_grad_kernel<16><<<(unsigned)nblocks, 256, 0, st>>>(d, X, n, ldx, Kinv, ldk, alpha, partial);
    BGP_LAUNCH_OK(ctx);
    grad_finish_kernel<<<slot, 256, 0, st>>>(partial, (int)nblocks, NA, mp, grad);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

}  // namespace bgp
