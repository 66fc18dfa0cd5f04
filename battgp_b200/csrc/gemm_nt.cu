// FP64 GEMM "NT" on the Blackwell DMMA tensor pipe:  C[M,N] = alpha * A[M,K] * B[N,K]^T + beta * C
//
// Every dense contraction of the exact-GP path has this shape once L is stored row-major/lower
// (DESIGN.md "GEMM shapes"): the trailing update A22 -= L21 L21^T (SYRK, tri mask), the panel solve
// L21 = A21 * inv(L11)^T, the recursive TRSM update X2 -= X1 L21^T, POTRI's products.
//
// Design (sm_100a): tcgen05 has no f64 kind, so the FP64 tensor path is mma.sync.m8n8k4.f64 (SASS
// DMMA.8x8x4, the only native f64 MMA shape on sm_100a -- wider PTX shapes decompose into it).  One DMMA
// occupies an SM sub-partition's FP64 pipe for 16 cycles (measured 37.1 TF/s chip peak,
// profiles/fp64_peak_r01.txt), so the kernel's job is to keep 8 resident warps issuing back-to-back
// independent DMMAs: 128x128 CTA tile, 64x32 warp tile (32 independent accumulators per k-step), operands
// staged by a 4-deep cp.async (LDGSTS) pipeline in 16-wide K chunks with a 20-double row pitch
// (bank-conflict-free 64-bit fragment loads), one __syncthreads per 128 DMMAs per warp.
#include "common.cuh"

namespace bgp {


__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int BM, int BN, int WM, int WN, int BK, int STAGES, int MINB, bool ALIGNED>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, MINB)
gemm_nt_kernel(GemmArgs g, int tiles_m, int tiles_n, int c_vec) {
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr int PITCH = BK + 4;      // 20 or 36 doubles: == 4 (mod 16) -> conflict-free 64-bit fragment loads
    constexpr int SEG = BK / 2;        // 16-byte segments per tile row
    constexpr int WARPS_N = BN / WN;
    constexpr int MI = WM / 8, NI = WN / 8;
    extern __shared__ __align__(16) double smem[];
    pdl_wait();
    pdl_launch_dependents();
    double* As = smem;
    double* Bs = smem + (size_t)STAGES * BM * PITCH;

    // grouped raster: 8 tile-rows share their B tiles in L2 while sweeping the columns
    constexpr int GROUP = 8;
    const int pid = blockIdx.x;
    const int per_group = GROUP * tiles_n;
    const int gid = pid / per_group;
    const int first_m = gid * GROUP;
    const int gsize = min(tiles_m - first_m, GROUP);
    const int tm = first_m + (pid % per_group) % gsize;
    const int tn = (pid % per_group) / gsize;
    const int m0 = tm * BM, n0 = tn * BN;
    if (g.tri && ((int64_t)n0 + g.coff > (int64_t)m0 + BM - 1 + g.roff)) return;

    const int tid = threadIdx.x;
    const int M = g.M, N = g.N, K = g.K;
    const double* __restrict__ A = g.A;
    const double* __restrict__ B = g.B;
    const int64_t lda = g.lda, ldb = g.ldb;

    // Per-thread copy slots: slot i of a thread always covers the same (row, 16-byte column) of the A / B tile, so
    // the source pointers and row predicates are computed once; a full chunk is then ASLOTS+BSLOTS branch-free
    // cp.async per thread (zero-fill for rows outside the matrix), only the ragged last K chunk takes the slow path.
    constexpr int RSTEP = NT / SEG;
    constexpr int ASLOTS = BM * SEG / NT, BSLOTS = BN * SEG / NT;
    static_assert(NT % SEG == 0 && (BM * SEG) % NT == 0 && (BN * SEG) % NT == 0, "tile/thread mismatch");
    const int lr = tid / SEG, lc = (tid % SEG) * 2;
    const double* a_base = A + (int64_t)(m0 + lr) * lda + lc;
    const double* b_base = B + (int64_t)(n0 + lr) * ldb + lc;
    const int64_t a_rstep = (int64_t)RSTEP * lda, b_rstep = (int64_t)RSTEP * ldb;
    unsigned a_mask = 0, b_mask = 0;
#pragma unroll
    for (int i = 0; i < ASLOTS; i++) a_mask |= (m0 + lr + i * RSTEP < M) ? (1u << i) : 0u;
#pragma unroll
    for (int i = 0; i < BSLOTS; i++) b_mask |= (n0 + lr + i * RSTEP < N) ? (1u << i) : 0u;
    const int sm_off = lr * PITCH + lc;

    auto load_chunk = [&](int stage, int k0) {
        double* as = As + (size_t)stage * BM * PITCH;
        double* bs = Bs + (size_t)stage * BN * PITCH;
        if (ALIGNED && k0 + BK <= K) {
#pragma unroll
            for (int i = 0; i < ASLOTS; i++) {
                const bool ok = (a_mask >> i) & 1u;
                cp_async16(as + sm_off + i * RSTEP * PITCH, ok ? a_base + i * a_rstep + k0 : A, ok ? 16 : 0);
            }
#pragma unroll
            for (int i = 0; i < BSLOTS; i++) {
                const bool ok = (b_mask >> i) & 1u;
                cp_async16(bs + sm_off + i * RSTEP * PITCH, ok ? b_base + i * b_rstep + k0 : B, ok ? 16 : 0);
            }
        } else if (ALIGNED) {
            for (int s = tid; s < BM * SEG; s += NT) {
                const int r = s / SEG, c = (s % SEG) * 2;
                const int gr = m0 + r, gk = k0 + c;
                int bytes = 0;
                if (gr < M) { const int rem = K - gk; bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0); }
                const double* src = bytes ? A + (int64_t)gr * lda + gk : A;
                cp_async16(as + r * PITCH + c, src, bytes);
            }
            for (int s = tid; s < BN * SEG; s += NT) {
                const int r = s / SEG, c = (s % SEG) * 2;
                const int gr = n0 + r, gk = k0 + c;
                int bytes = 0;
                if (gr < N) { const int rem = K - gk; bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0); }
                const double* src = bytes ? B + (int64_t)gr * ldb + gk : B;
                cp_async16(bs + r * PITCH + c, src, bytes);
            }
        } else {
            for (int s = tid; s < BM * BK; s += NT) {
                const int r = s / BK, c = s % BK;
                const int gr = m0 + r, gk = k0 + c;
                const int bytes = (gr < M && gk < K) ? 8 : 0;
                const double* src = bytes ? A + (int64_t)gr * lda + gk : A;
                cp_async8(as + r * PITCH + c, src, bytes);
            }
            for (int s = tid; s < BN * BK; s += NT) {
                const int r = s / BK, c = s % BK;
                const int gr = n0 + r, gk = k0 + c;
                const int bytes = (gr < N && gk < K) ? 8 : 0;
                const double* src = bytes ? B + (int64_t)gr * ldb + gk : B;
                cp_async8(bs + r * PITCH + c, src, bytes);
            }
        }
    };

    const int warp = tid >> 5, lane = tid & 31;
    const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
    const int fr = lane >> 2, fk = lane & 3;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int KT = (K + BK - 1) / BK;
    int kt0 = 0;
    if (g.kskip) {
        const int64_t kb = (int64_t)m0 + g.kofs;
        kt0 = kb > 0 ? (int)(kb / BK) : 0;
        if (kt0 > KT) kt0 = KT;
    }
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (kt0 + s < KT) load_chunk((kt0 + s) % STAGES, (kt0 + s) * BK);
        cp_async_commit();
    }
    for (int kt = kt0; kt < KT; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_chunk(nk % STAGES, nk * BK);
            cp_async_commit();
        }
        const double* as = As + (size_t)(kt % STAGES) * BM * PITCH + (wm0 + fr) * PITCH + fk;
        const double* bs = Bs + (size_t)(kt % STAGES) * BN * PITCH + (wn0 + fr) * PITCH + fk;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; i++) a[i] = as[i * 8 * PITCH + kk];
#pragma unroll
            for (int j = 0; j < NI; j++) b[j] = bs[j * 8 * PITCH + kk];
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: each thread owns 2 consecutive columns of one row per 8x8 MMA tile
    const double alpha = g.alpha, beta = g.beta;
    double* __restrict__ C = g.C;
    const int64_t ldc = g.ldc;
#pragma unroll
    for (int i = 0; i < MI; i++) {
        const int row = m0 + wm0 + i * 8 + fr;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < NI; j++) {
            const int col = n0 + wn0 + j * 8 + fk * 2;
            bool ok0 = col < N, ok1 = col + 1 < N;
            if (g.tri) {
                ok0 = ok0 && ((int64_t)col + g.coff <= (int64_t)row + g.roff);
                ok1 = ok1 && ((int64_t)col + 1 + g.coff <= (int64_t)row + g.roff);
            }
            if (!ok0 && !ok1) continue;
            double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
            double* p = C + (int64_t)row * ldc + col;
            if (ok0 && ok1 && c_vec) {
                if (beta != 0.0) {
                    const double2 o = *reinterpret_cast<const double2*>(p);
                    v0 += beta * o.x;
                    v1 += beta * o.y;
                }
                *reinterpret_cast<double2*>(p) = make_double2(v0, v1);
            } else {
                if (ok0) { if (beta != 0.0) v0 += beta * p[0]; p[0] = v0; }
                if (ok1) { if (beta != 0.0) v1 += beta * p[1]; p[1] = v1; }
            }
        }
    }
}

template <int BM, int BN, int WM, int WN, int BK, int STAGES, int MINB>
static int launch_cfg(Ctx* ctx, const GemmArgs& g, cudaStream_t st) {
    constexpr int NT = (BM / WM) * (BN / WN) * 32;
    constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * (BK + 4) * sizeof(double);
    const int tiles_m = (g.M + BM - 1) / BM, tiles_n = (g.N + BN - 1) / BN;
    const bool aligned = ((g.lda | g.ldb) & 1) == 0 && (((uintptr_t)g.A | (uintptr_t)g.B) & 15) == 0;
    const int c_vec = ((g.ldc & 1) == 0 && ((uintptr_t)g.C & 15) == 0) ? 1 : 0;
    auto kern = aligned ? gemm_nt_kernel<BM, BN, WM, WN, BK, STAGES, MINB, true>
                        : gemm_nt_kernel<BM, BN, WM, WN, BK, STAGES, MINB, false>;
    // the opt-in is per function and per device; remember which devices have it
    static thread_local uint64_t attr_done[2] = {0, 0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_done[aligned] & dev_bit)) {
        BGP_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
        attr_done[aligned] |= dev_bit;
    }
    BGP_CUDA_OK(launch_pdl(ctx->pdl && ctx->pdl_chain, kern, dim3(tiles_m * tiles_n), dim3(NT), SMEM, st, g, tiles_m, tiles_n, c_vec));
    BGP_LAUNCH_OK(ctx);
    return 0;
}

// cfg: 0 auto, 1 = 128x128, 2 = 64x128 (in-place safe for N<=128, short M), 3 = 64x64, 5 = 128x64 (default big),
// 8 = 16x128 and 9 = 32x32 (the two serial GEMMs between the leaves of potrf_chain2: many small CTAs instead of 2-4 big ones);
// 4..7 are experimental variants selectable through bgp_ctx_set("gemm_cfg", k)
int gemm_nt_cfg(Ctx* ctx, const GemmArgs& g, int cfg, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return 0;
    if (cfg == 0) {
        // 128x64 tiles, 2 CTAs/SM (16 warps/SM hide each other's barriers): 33.9 TF/s at 8192^3 vs 29.8 for 128x128
        const int64_t t = (int64_t)((g.M + 127) / 128) * ((g.N + 63) / 64);
        cfg = (t >= 240) ? (ctx->gemm_cfg ? ctx->gemm_cfg : 5) : 3;
        // short-and-wide products (the M = 300 query block of the predictive variance): 64-row tiles waste less
        if (g.M <= 1024 && (((g.M + 127) / 128) * 128 - g.M) >= 64 && (int64_t)((g.M + 63) / 64) * ((g.N + 127) / 128) >= 240)
            cfg = 2;
    }
    switch (cfg) {
        case 1: return launch_cfg<128, 128, 64, 32, 16, 4, 1>(ctx, g, st);
        case 2: return launch_cfg<64, 128, 32, 32, 16, 3, 2>(ctx, g, st);    // 8 warps, 2 CTAs / SM
        case 4: return launch_cfg<128, 128, 32, 32, 16, 4, 1>(ctx, g, st);   // 16 warps
        case 5: return launch_cfg<128, 64, 32, 32, 16, 3, 2>(ctx, g, st);    // 2 CTAs / SM
        case 6: return launch_cfg<128, 128, 64, 32, 32, 3, 1>(ctx, g, st);   // BK = 32
        case 8: return launch_cfg<16, 128, 16, 32, 16, 3, 1>(ctx, g, st);    // latency form of the in-place leaf solve: 8 CTAs per 128 rows
        case 9: return launch_cfg<32, 32, 16, 16, 16, 4, 1>(ctx, g, st);     // latency form of a 128 x 128 update: 10-16 CTAs
        default: return launch_cfg<64, 64, 32, 32, 16, 4, 1>(ctx, g, st);
    }
}

int gemm_nt(Ctx* ctx, const GemmArgs& g, cudaStream_t st) { return gemm_nt_cfg(ctx, g, 0, st); }

}  // namespace bgp
