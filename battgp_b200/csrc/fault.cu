// Fault evaluation on the device: what /root/reference/src/batt_models/fault_evaluation.py computes on the
// host from the predicted cell resistances r0 [M times, C cells] and their variances -- the outputs of the exact-GP path --
// so that a battery's fault probabilities leave the GPU instead of 2 x M x C intermediate values.
//   r0_mean[i][c]   Hodges-Lehmann location of the OTHER cells at time i (median of the Walsh averages)   fault_evaluation.py:11-17,59-71
//   P_above/below   1 - Phi((r0_mean + band - r0)/std), Phi((r0_mean - band - r0)/std)                    :79-85
//   P_threshold     1 - Phi((thr - r0)/std)                                                                 :94-101
//   cells_var[i]    population variance over the cells                                                     :103-104
//   weakest[i]      1 - prod_c (1 - P_above - P_below)                                                      fault_probabilities.py:88-95
// One thread per (time, cell); C <= 16.
#include "common.cuh"

namespace bgp {

constexpr int FAULT_MAXC = 16;

__device__ __forceinline__ double normal_cdf_dev(double x, double mean, double std) {
    return 0.5 + 0.5 * erf((x - mean) / (1.4142135623730951 * std));          // np.sqrt(2) * std, as in fault_evaluation.py:7-8
}

__global__ void fault_eval_kernel(const double* __restrict__ r0, const double* __restrict__ r0var, int M, int C, int64_t ld, double band,
                                  double thr, double* p_out, double* p_above, double* p_below, double* r0_mean, double* p_thr,
                                  double* cells_var, double* weakest) {
    const int i = blockIdx.x * (blockDim.x / FAULT_MAXC) + threadIdx.x / FAULT_MAXC;
    const int c = threadIdx.x % FAULT_MAXC;
    __shared__ double s_band[16][FAULT_MAXC];                 // blockDim.x == 256: 16 time rows per block
    const int lr = threadIdx.x / FAULT_MAXC;
    const bool live = i < M && c < C;
    double pa = 0.0, pb = 0.0;
    if (live) {
        const double* row = r0 + (int64_t)i * ld;
        // Walsh averages of the other cells, sorted by insertion (at most 15 * 16 / 2 = 120 values)
        double w[FAULT_MAXC * (FAULT_MAXC - 1) / 2];
        int cnt = 0;
        for (int a = 0; a < C; a++) {
            if (a == c) continue;
            for (int b = a; b < C; b++) {
                if (b == c) continue;
                const double v = (row[a] + row[b]) / 2;
                int p = cnt++;
                while (p > 0 && w[p - 1] > v) { w[p] = w[p - 1]; p--; }
                w[p] = v;
            }
        }
        const double med = (cnt & 1) ? w[cnt / 2] : (w[cnt / 2 - 1] + w[cnt / 2]) / 2;   // np.median
        const double x = row[c], sd = sqrt(r0var[(int64_t)i * ld + c]);
        pa = 1.0 - normal_cdf_dev(med + band, x, sd);
        pb = normal_cdf_dev(med - band, x, sd);
        const int64_t o = (int64_t)i * C + c;
        r0_mean[o] = med;
        p_above[o] = pa;
        p_below[o] = pb;
        p_out[o] = pa + pb;
        p_thr[o] = 1.0 - normal_cdf_dev(thr, x, sd);
    }
    s_band[lr][c] = live ? pa + pb : 0.0;
    __syncthreads();
    if (i < M && c == 0) {
        const double* row = r0 + (int64_t)i * ld;
        double mean = 0.0;
        for (int a = 0; a < C; a++) mean += row[a];
        mean /= C;
        double v = 0.0, prod = 1.0;
        for (int a = 0; a < C; a++) { const double d = row[a] - mean; v += d * d; prod *= 1.0 - s_band[lr][a]; }
        cells_var[i] = v / C;
        weakest[i] = 1.0 - prod;
    }
}

int fault_eval(Ctx* ctx, const double* r0, const double* r0var, int64_t M, int64_t C, int64_t ld, double band, double thr, double* p_out,
               double* p_above, double* p_below, double* r0_mean, double* p_thr, double* cells_var, double* weakest, cudaStream_t st) {
    if (M <= 0) return 0;
    const int rows_per_block = 256 / FAULT_MAXC;
    fault_eval_kernel<<<(unsigned)((M + rows_per_block - 1) / rows_per_block), 256, 0, st>>>(r0, r0var, (int)M, (int)C, ld, band, thr, p_out,
                                                                                            p_above, p_below, r0_mean, p_thr, cells_var, weakest);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

}  // namespace bgp
