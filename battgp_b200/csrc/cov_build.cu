// K1+K2+K3: fused covariance build.
//   out[i][j] = sum_terms outputscale * k_term(x1_i, x2_j)  (+ noise where i == j in symmetric mode)
// One pass over the output, HBM-write bound (8 B per element; X is O(N) and staged in shared memory).
// Replaces WienerKernel.forward's cdist + Python column loop + 4 elementwise passes
// (/root/reference/src/gp/wiener_kernel.py:10-32), GPyTorch's RBF GEMM/clamp/div/exp passes, the two ScaleKernel
// multiplies, the kernel sum and the likelihood's noise add (cell_gp.py:27-36) -- ~11 N^2 passes -> 1.
//
// CTA = 256 threads -> 64 x 128 output tile; warp w owns rows 8w..8w+7, lane l owns columns {2l, 2l+1, 64+2l,
// 64+2l+1} so every store instruction of a warp covers 512 contiguous bytes.  Features are pre-scaled once per CTA
// (x/l for RBF/Matern, x*pi/p for Periodic, raw t for Wiener); the column features live in registers.
#include "common.cuh"
#include "kernspec.cuh"

namespace bgp {

constexpr int TM = 64, TN = 128;

// exp(x) for x <= 0 (RBF / Matern / Periodic arguments are never positive): range reduction by 2^k with the 2^52+2^51
// rounding trick, degree-13 Taylor polynomial on |r| <= ln2/2 (truncation 4e-18 relative), exponent patched in by integer add.
// ~19 FP64 instructions; the only special case is x < -708 (result near / below the smallest normal), which takes
// libdevice's exp().
__device__ __forceinline__ double exp_nonpos(double x) {
    const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.60590438368216145994e-10;                 // 1/13!
    p = fma(p, r, 2.08767569878680989792e-09);
    p = fma(p, r, 2.50521083854417187751e-08);
    p = fma(p, r, 2.75573192239858906526e-07);
    p = fma(p, r, 2.75573192239858906526e-06);
    p = fma(p, r, 2.48015873015873015873e-05);
    p = fma(p, r, 1.98412698412698412698e-04);
    p = fma(p, r, 1.38888888888888888889e-03);
    p = fma(p, r, 8.33333333333333333333e-03);
    p = fma(p, r, 4.16666666666666666667e-02);
    p = fma(p, r, 1.66666666666666666667e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double res = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    if (x < -708.0) return exp(x);        // (sub)normal boundary: rare, take libdevice's careful path
    return res;
}

// value of the kernel for one (row, col) pair given pre-scaled features; FMAX-unrolled, uniform branches
template <int FMAX>
__device__ __forceinline__ double eval_pair(const DevSpec& sp, const double (&rf)[FMAX], const double (&cf)[FMAX]) {
    double ksum = 0.0, acc = 0.0;
#pragma unroll
    for (int f = 0; f < FMAX; f++) {
        if (f < sp.nfeat) {
            const int ty = sp.ftype[f];
            const double diff = rf[f] - cf[f];
            if (ty == BGP_WIENER) {
                const double m = fmin(rf[f], cf[f]);
                const double m2 = m * m;
                ksum += sp.fos[f] * (m2 * m * (1.0 / 3.0) + fabs(diff) * m2 * 0.5);
            } else {
                if (ty == BGP_PERIODIC) { const double s = sin(diff); acc = fma(s * s, sp.faux[f], acc); }
                else acc = fma(diff, diff, acc);
                if (sp.flast[f]) {
                    double v;
                    if (ty == BGP_RBF) v = exp_nonpos(-0.5 * acc);
                    else if (ty == BGP_PERIODIC) v = exp_nonpos(-2.0 * acc);
                    else {  // Matern-5/2: r = sqrt(clamp(sqdist, 1e-30))
                        const double r = sqrt(fmax(acc, 1e-30));
                        const double s5r = 2.23606797749978969641 * r;
                        v = (1.0 + s5r + (5.0 / 3.0) * r * r) * exp_nonpos(-s5r);
                    }
                    ksum = fma(sp.fos[f], v, ksum);
                    acc = 0.0;
                }
            }
        }
    }
    return ksum;
}

// BattGP's own kernel structure (cell_gp.py:32-36): feature 0 = integrated Wiener on t, features 1..3 = one RBF-ARD term.
// Same arithmetic as eval_pair<4> with every spec-dependent branch resolved at compile time.
__device__ __forceinline__ double eval_pair_battgp(double osw, double osr, const double (&rf)[4], const double (&cf)[4]) {
    const double m = fmin(rf[0], cf[0]);
    const double m2 = m * m;
    const double w = m2 * m * (1.0 / 3.0) + fabs(rf[0] - cf[0]) * m2 * 0.5;
    const double d1 = rf[1] - cf[1], d2 = rf[2] - cf[2], d3 = rf[3] - cf[3];
    const double acc = fma(d3, d3, fma(d2, d2, d1 * d1));
    return fma(osr, exp_nonpos(-0.5 * acc), osw * w);
}

template <int FMAX, bool BATTGP>
__global__ void __launch_bounds__(256)
cov_build_kernel(DevSpec sp, const double* __restrict__ X1, int64_t n1, int64_t ldx1,
                 const double* __restrict__ X2, int64_t n2, int64_t ldx2,
                 double* __restrict__ out, int64_t ldo, int symmetric, int vec_ok) {
    const int64_t m0 = (int64_t)blockIdx.y * TM, c0 = (int64_t)blockIdx.x * TN;
    if (symmetric && c0 > m0 + TM - 1) return;
    __shared__ double rowf[TM][FMAX];
    __shared__ double colf[TN][FMAX + 1];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < TM * FMAX; idx += 256) {
        const int r = idx / FMAX, f = idx % FMAX;
        double v = 0.0;
        if (f < sp.nfeat && m0 + r < n1) v = prescale(sp, f, X1[(m0 + r) * ldx1 + sp.fdim[f]]);
        rowf[r][f] = v;
    }
    for (int idx = tid; idx < TN * FMAX; idx += 256) {
        const int r = idx / FMAX, f = idx % FMAX;
        double v = 0.0;
        if (f < sp.nfeat && c0 + r < n2) v = prescale(sp, f, X2[(c0 + r) * ldx2 + sp.fdim[f]]);
        colf[r][f] = v;
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    double cf[4][FMAX];
    const int cl[4] = {2 * lane, 2 * lane + 1, 64 + 2 * lane, 64 + 2 * lane + 1};
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int f = 0; f < FMAX; f++) cf[c][f] = colf[cl[c]][f];

#pragma unroll 2
    for (int r = 0; r < 8; r++) {
        const int lr = warp * 8 + r;
        const int64_t grow = m0 + lr;
        if (grow >= n1) break;
        double rf[FMAX];
#pragma unroll
        for (int f = 0; f < FMAX; f++) rf[f] = rowf[lr][f];
        double v[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if constexpr (BATTGP) v[c] = eval_pair_battgp(sp.fos[0], sp.fos[3], rf, cf[c]);
            else v[c] = eval_pair<FMAX>(sp, rf, cf[c]);
            if (symmetric && grow == c0 + cl[c]) v[c] += sp.noise;
        }
        double* orow = out + grow * ldo + c0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int64_t gc = c0 + cl[2 * h];
            if (vec_ok && gc + 1 < n2) {
                *reinterpret_cast<double2*>(orow + cl[2 * h]) = make_double2(v[2 * h], v[2 * h + 1]);
            } else {
                if (gc < n2) orow[cl[2 * h]] = v[2 * h];
                if (gc + 1 < n2) orow[cl[2 * h] + 1] = v[2 * h + 1];
            }
        }
    }
}

__global__ void cov_diag_kernel(DevSpec sp, const double* __restrict__ X, int64_t n, int64_t ldx, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int f = 0; f < sp.nfeat; f++) {
        if (sp.ftype[f] == BGP_WIENER) {
            const double t = X[i * ldx + sp.fdim[f]];
            s += sp.fos[f] * t * t * t * (1.0 / 3.0);      // wiener_kernel.py:32 with x1 == x2
        } else if (sp.flast[f]) {
            s += sp.fos[f];                                  // stationary terms: k(x,x) = 1
        }
    }
    out[i] = s;
}

int cov_build(Ctx* ctx, const bgp_kernel_spec* spec, const double* X1, int64_t n1, int64_t ldx1, const double* X2,
              int64_t n2, int64_t ldx2, double* out, int64_t ldo, int symmetric, cudaStream_t st) {
    DevSpec d;
    int rc = make_devspec(spec, &d);
    if (rc) return rc;
    if (symmetric) { X2 = X1; n2 = n1; ldx2 = ldx1; }
    if (n1 == 0 || n2 == 0) return 0;
    int maxdim = 0;
    for (int f = 0; f < d.nfeat; f++) maxdim = d.fdim[f] > maxdim ? d.fdim[f] : maxdim;
    if (!X1 || !X2 || !out || n1 < 0 || n2 < 0 || ldx1 <= maxdim || ldx2 <= maxdim || ldo < n2) return BGP_E_ARG;
    const int64_t gy = (n1 + TM - 1) / TM, gx = (n2 + TN - 1) / TN;
    if (gy > 65535) return BGP_E_ARG;
    const int vec_ok = ((ldo & 1) == 0 && ((uintptr_t)out & 15) == 0) ? 1 : 0;
    dim3 grid((unsigned)gx, (unsigned)gy);
    const bool battgp = d.nfeat == 4 && d.ftype[0] == BGP_WIENER && d.ftype[1] == BGP_RBF && d.ftype[2] == BGP_RBF &&
                        d.ftype[3] == BGP_RBF && d.flast[0] && !d.flast[1] && !d.flast[2] && d.flast[3];
    if (battgp)
        cov_build_kernel<4, true><<<grid, 256, 0, st>>>(d, X1, n1, ldx1, X2, n2, ldx2, out, ldo, symmetric, vec_ok);
    else if (d.nfeat <= 4)
        cov_build_kernel<4, false><<<grid, 256, 0, st>>>(d, X1, n1, ldx1, X2, n2, ldx2, out, ldo, symmetric, vec_ok);
    else if (d.nfeat <= 8)
        cov_build_kernel<8, false><<<grid, 256, 0, st>>>(d, X1, n1, ldx1, X2, n2, ldx2, out, ldo, symmetric, vec_ok);
    else
        cov_build_kernel<16, false><<<grid, 256, 0, st>>>(d, X1, n1, ldx1, X2, n2, ldx2, out, ldo, symmetric, vec_ok);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

int cov_diag(Ctx* ctx, const bgp_kernel_spec* spec, const double* X, int64_t n, int64_t ldx, double* out, cudaStream_t st) {
    DevSpec d;
    int rc = make_devspec(spec, &d);
    if (rc) return rc;
    if (n == 0) return 0;
    if (!X || !out || n < 0) return BGP_E_ARG;
    cov_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, X, n, ldx, out);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

}  // namespace bgp
