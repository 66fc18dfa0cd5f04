// Components of the MODULAR int8 emulation of fp64 products ("Ozaki scheme II", DESIGN.md section 5, tools/ozaki2_model.py):
//   C += alpha * A B^T  from 16 int8 x int8 -> int32 products (one per modulus) instead of the 36 digit-plane products of
//   ozaki.cu, followed by a Chinese-remainder reconstruction.
// This file holds the two arithmetic kernels every variant of that scheme needs, bit-compatible with the CPU model:
//   oz2_residue_kernel : fp64 rows -> per-row scale exponent + symmetric int8 residues of trunc(a * 2^(BETA - e)) for the
//                        16 moduli (plain [modulus][row][K] layout; the swizzled operand image comes with the MMA kernel)
//   oz2_crt_kernel     : 16 int32 accumulators per element -> exact 128-bit integer -> double -> C += alpha * 2^(ea+eb-2 BETA) * C'
// A first form of the tensor-core kernel that goes between them is next/ozaki2_mma.cu (bgp_oz2_gemm, experimental); bgp_potrf
// does not use this scheme yet.
#include <climits>
#include "common.cuh"

namespace bgp {

constexpr int OZ2_NMOD = 16;
constexpr int OZ2_BETA = 55;
__host__ __device__ constexpr int oz2_modulus(int j) {
    constexpr int m[OZ2_NMOD] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193};
    return m[j];
}

struct Oz2Consts {
    uint32_t w[OZ2_NMOD][4];     // CRT weights W_j = (P/p_j) * ((P/p_j)^-1 mod p_j) mod P, 32-bit limbs (little endian)
    uint32_t p[4];               // P = prod p_j
    double wf[OZ2_NMOD];         // W_j / P (fp64) for the quotient estimate
};
__constant__ Oz2Consts c_oz2;

// ---- host-side big-integer helpers (5 x 32-bit limbs are enough: P < 2^126) --------------------------------------
struct U160 { uint32_t v[5]; };
static U160 u160_mul_small(const U160& a, uint32_t m) {
    U160 r; uint64_t c = 0;
    for (int i = 0; i < 5; i++) { const uint64_t t = (uint64_t)a.v[i] * m + c; r.v[i] = (uint32_t)t; c = t >> 32; }
    return r;
}
static uint32_t u160_divmod_small(U160& a, uint32_t d) {      // a /= d, returns remainder
    uint64_t rem = 0;
    for (int i = 4; i >= 0; i--) { const uint64_t t = (rem << 32) | a.v[i]; a.v[i] = (uint32_t)(t / d); rem = t % d; }
    return (uint32_t)rem;
}
static double u160_to_double(const U160& a) {
    double r = 0.0;
    for (int i = 4; i >= 0; i--) r = r * 4294967296.0 + (double)a.v[i];
    return r;
}

void oz2_host_consts(uint32_t (&w)[OZ2_NMOD][4], uint32_t (&pl)[4], double (&wf)[OZ2_NMOD]);

static int oz2_upload_consts(Ctx* ctx) {
    static thread_local uint64_t done = 0;
    const uint64_t bit = 1ull << (ctx->device & 63);
    if (done & bit) return 0;
    Oz2Consts h = {};
    oz2_host_consts(h.w, h.p, h.wf);
    BGP_CUDA_OK(cudaMemcpyToSymbol(c_oz2, &h, sizeof(h)));
    done |= bit;
    return 0;
}

// CRT constants for the moduli above (host arithmetic on 160-bit integers)
void oz2_host_consts(uint32_t (&w)[OZ2_NMOD][4], uint32_t (&pl)[4], double (&wf)[OZ2_NMOD]) {
    struct { uint32_t (&w)[OZ2_NMOD][4]; uint32_t (&p)[4]; double (&wf)[OZ2_NMOD]; } h{w, pl, wf};
    U160 P = {{1, 0, 0, 0, 0}};
    for (int j = 0; j < OZ2_NMOD; j++) P = u160_mul_small(P, (uint32_t)oz2_modulus(j));
    for (int k = 0; k < 4; k++) h.p[k] = P.v[k];
    const double Pf = u160_to_double(P);
    for (int j = 0; j < OZ2_NMOD; j++) {
        const uint32_t p = (uint32_t)oz2_modulus(j);
        U160 M = P;
        u160_divmod_small(M, p);                               // M_j = P / p_j (exact)
        U160 t = M;
        const uint32_t mr = u160_divmod_small(t, p);           // M_j mod p_j
        uint32_t inv = 0;
        for (uint32_t y = 1; y < p; y++) if ((uint64_t)mr * y % p == 1) { inv = y; break; }
        const U160 W = u160_mul_small(M, inv);                 // < P: already reduced
        for (int k = 0; k < 4; k++) h.w[j][k] = W.v[k];
        h.wf[j] = u160_to_double(W) / Pf;
    }
}

// ---- residues -------------------------------------------------------------------------------------------------------
// One warp per row.  e = exponent with 2^e >= 2 max|a| (0 for an all-zero row); v = trunc(a * 2^(BETA - e)), |v| < 2^54;
// residue of v for modulus p from three 20-bit chunks: v = c0 + c1 2^20 + c2 2^40 (sign handled separately).
template <int J>
__device__ __forceinline__ void oz2_store_residues(int8_t* __restrict__ out, int64_t plane, int64_t idx, uint32_t c0, uint32_t c1,
                                                   uint32_t c2, bool neg) {
    if constexpr (J < OZ2_NMOD) {
        constexpr uint32_t p = (uint32_t)oz2_modulus(J);
        constexpr uint32_t m20 = (uint32_t)((1ull << 20) % p), m40 = (uint32_t)((1ull << 40) % p);
        uint32_t r = (c0 + c1 * m20 + c2 * m40) % p;           // < 2^20 + 2^28 + 2^22
        if (neg && r) r = p - r;                               // [0, p)
        const int s = (r >= (p + 1) / 2) ? (int)r - (int)p : (int)r;       // [-p/2, p/2)
        out[J * plane + idx] = (int8_t)s;
        oz2_store_residues<J + 1>(out, plane, idx, c0, c1, c2, neg);
    }
}

__global__ void __launch_bounds__(256)
oz2_residue_kernel(const double* __restrict__ A, int64_t rows, int64_t K, int64_t ld, int8_t* __restrict__ out,
                   int32_t* __restrict__ expo) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double* a = A + row * ld;
    // row maximum through the bit patterns (|x| ordering == unsigned ordering of the payload); NaN/Inf poison the row
    unsigned long long mx = 0;
    for (int64_t k = lane; k < K; k += 32) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(a[k]) & 0x7fffffffffffffffull;
        mx = b > mx ? b : mx;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_xor_sync(0xffffffffu, mx, o); mx = t > mx ? t : mx; }
    int e = 0;
    const bool bad = mx >= 0x7ff0000000000000ull;
    if (mx != 0 && !bad) {
        int fe;
        (void)frexp(__longlong_as_double((long long)mx), &fe);       // max = f 2^fe, 0.5 <= f < 1
        e = fe + 1;
    }
    if (lane == 0) expo[row] = bad ? INT_MIN : e;
    const int64_t plane = rows * K;
    for (int64_t k = lane; k < K; k += 32) {
        const long long v = bad ? 0ll : __double2ll_rz(scalbn(a[k], OZ2_BETA - e));
        const bool neg = v < 0;
        const unsigned long long u = (unsigned long long)(neg ? -v : v);
        oz2_store_residues<0>(out, plane, row * K + k, (uint32_t)(u & 0xFFFFFu), (uint32_t)((u >> 20) & 0xFFFFFu), (uint32_t)(u >> 40), neg);
    }
}

// ---- reconstruction ----------------------------------------------------------------------------------------------------
template <int J>
__device__ __forceinline__ void oz2_accumulate(const int32_t* __restrict__ G, int64_t plane, int64_t idx, unsigned long long (&s)[4],
                                               double& est) {
    if constexpr (J < OZ2_NMOD) {
        constexpr int p = oz2_modulus(J);
        int t = G[J * plane + idx] % p;                        // (-p, p)
        t += (t < 0) ? p : 0;                                  // [0, p)
#pragma unroll
        for (int k = 0; k < 4; k++) s[k] += (unsigned long long)c_oz2.w[J][k] * (unsigned)t;      // IMAD.WIDE, < 2^44 in total
        est += (double)t * c_oz2.wf[J];
        oz2_accumulate<J + 1>(G, plane, idx, s, est);
    }
}

// G: [16][M][N] int32 accumulators; C[i][j] += alpha * 2^(ea_i + eb_j - 2 BETA) * C'_ij
__global__ void __launch_bounds__(256)
oz2_crt_kernel(const int32_t* __restrict__ G, int64_t M, int64_t N, const int32_t* __restrict__ ea, const int32_t* __restrict__ eb,
               double alpha, double* __restrict__ C, int64_t ldc) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const int64_t i = idx / N, j = idx % N;
    unsigned long long s[4] = {0, 0, 0, 0};
    double est = 0.0;
    oz2_accumulate<0>(G, M * N, idx, s, est);
    const long long q = __double2ll_rn(est);                   // S = C' + q P, |C'| < P / 8
    long long r[4];
#pragma unroll
    for (int k = 0; k < 4; k++) r[k] = (long long)s[k] - q * (long long)c_oz2.p[k];      // |.| < 2^45
#pragma unroll
    for (int k = 0; k < 3; k++) {                              // carry-normalise: limbs 0..2 into [0, 2^32)
        const long long c = r[k] >> 32;
        r[k] -= c * 4294967296ll;
        r[k + 1] += c;
    }
    long long hi = r[3] * 4294967296ll + r[2];
    unsigned long long lo = ((unsigned long long)r[1] << 32) + (unsigned long long)r[0];
    const bool neg = hi < 0;
    if (neg) {                                                 // two's-complement negate of the 128-bit value
        hi = ~hi + (lo == 0 ? 1 : 0);
        lo = ~lo + 1ull;
    }
    double mag = __ull2double_rn((unsigned long long)hi) * 18446744073709551616.0 + __ull2double_rn(lo);
    if (neg) mag = -mag;
    const int ei = ea[i], ej = eb[j];
    double val;
    if (ei == INT_MIN || ej == INT_MIN) val = __longlong_as_double(0x7ff8000000000000ll);      // NaN/Inf row: poison
    else val = scalbn(mag, ei + ej - 2 * OZ2_BETA);
    double* c = C + i * ldc + j;
    *c = *c + alpha * val;
}

int oz2_residues(Ctx* ctx, const double* A, int64_t rows, int64_t K, int64_t ld, int8_t* out, int32_t* expo, cudaStream_t st) {
    if (rows <= 0 || K <= 0) return 0;
    oz2_residue_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(A, rows, K, ld, out, expo);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

int oz2_crt(Ctx* ctx, const int32_t* G, int64_t M, int64_t N, const int32_t* ea, const int32_t* eb, double alpha, double* C,
            int64_t ldc, cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    int rc = oz2_upload_consts(ctx);
    if (rc) return rc;
    oz2_crt_kernel<<<(unsigned)((M * N + 255) / 256), 256, 0, st>>>(G, M, N, ea, eb, alpha, C, ldc);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

}  // namespace bgp
