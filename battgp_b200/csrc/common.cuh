// Shared declarations for the battgp_b200 CUDA engine (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "../../include/battgp_b200.h"

namespace bgp {

constexpr int LEAF = 128;
constexpr size_t SCRATCH_BYTES = 4u << 20;          // diagonal-block size of the leaf factorisation / stored inverses

struct Ctx {
    int device = 0;
    cudaStream_t panel_stream = nullptr;    // high priority: panel factorisation (look-ahead)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t leaf_stream = nullptr;     // high priority: off-critical-path updates of the leaf chain (potrf_chain2)
    cudaEvent_t ev_chain[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int leaf_chain = 1;                     // diagonal blocks / small matrices (129..leaf_chain_max rows): right-looking leaf chain on two streams
    int pdl_chain_rows = 12000;             // bgp_potrf chains its leaf / DMMA GEMM launches by programmatic dependent launch below this many rows
    int chain_split = 0;                    // leaf chain: factor-only leaf + substitution solve on the chain, block inverse on the side stream (measured: -4 % at N=1024, +6 % at 4096, +3 % at 40k: off)
    int chain_cfg = 1;                      // leaf chain: small-tile (many-CTA) forms of the two serial GEMMs between leaves
    int leaf_chain_max = 5120;
    int chain_whole_max = 5120;                // matrices up to this many rows (extra rows included) skip the panel schedule: one leaf chain over the whole matrix
    cudaEvent_t ev_panel[2] = {nullptr, nullptr};
    cudaEvent_t ev_trail[3] = {nullptr, nullptr, nullptr};
    int32_t* d_info = nullptr;              // first failing pivot (1-based), INT_MAX when none
    double* d_scal = nullptr;               // [0] logdet accumulator, [1] dot result, [8..] dot partials
    double* d_scratch = nullptr;            // SCRATCH_BYTES of reduction partials (lml_grad)
    int nb = 0;                             // 0 = automatic schedule (api.cu panel_width), else uniform panel width
    int lookahead = 1;
    int pdl = 1;                            // 1: chain kernels (triangular sweeps, leaf, DMMA GEMMs) use programmatic dependent launch
    int pdl_chain = 1;                      // leaf/GEMM launches only: cleared by bgp_potrf for large matrices (measured neutral to -0.5 %)
    int trace = 0;                          // 1: bgp_potrf prints per-panel event timings to stderr (diagnostics)
    // automatic panel schedule (nb == 0): the width of a panel follows the rows still to be factorised (api.cu panel_width)
    int sched_t1024 = 9000;                 // remaining rows >= this: 1024-wide panels (else 512)
    int sched_t2048 = 17000;                // int8 path only: 2048-wide panels from here up (0 = never)
    int sched_t4096 = 0;                    // int8 path only: 4096-wide panels from here up (0 = never)
    int sched_w0 = 0;                       // cap on the width of the first panel (its chain has nothing to hide behind)
    int sched_w1 = 0;                       // cap on the width of the second panel
    int oz_cluster = 1;                     // CTAs per cluster sharing the A operand by multicast (1, 2 or 4) when fully persistent
    int oz_tpc = 2;                         // tiles per CTA of the int8 kernel inside bgp_potrf (0 = fully persistent)
    int ozaki = 0;                          // 1: big trailing updates of bgp_potrf go through the int8/tcgen05 path
    void* ws = nullptr;                     // caller-provided scratch (bgp_ctx_set_workspace)
    int64_t ws_bytes = 0;
    void* ws_trsm = nullptr;                // tail of ws used by the TRSM updates (set by bgp_potrf for its panels)
    int64_t ws_trsm_bytes = 0;
    int oz_reserve = 0;                     // SMs left free by the persistent trailing update of the tail panels (sched_tail)
    int sched_tail = 0;                     // rows still to factorise below which the trailing updates run persistent on nsm - oz_reserve SMs (0 = never)
    int oz_reserve_now = 0;                 // set by bgp_potrf around such a launch
    int oz_tpc_gemm = 0;                    // tiles per CTA of stand-alone bgp_oz_gemm calls (0 = fully persistent)
    int oz_kfence = 1;                      // int8 kernel: tcgen05.fence::after_thread_sync after every operand-stage wait
    int oz_dbg_epi = 0;                     // int8 kernel experiments (results are garbage): parts of the epilogue switched off
    int oz_backoff = 0;                     // int8 kernel: nanosleep back-off in the producer / relay / epilogue barrier polls
    int oz_relay = 1;                       // int8 kernel: relay warp + named barrier instead of an mbarrier wait by the MMA-issuing warp
    int oz_order = 2;                       // int8 kernel: MMA issue order within a k-block (0: by A slice; 1: widest last per k-step; 2: seven N=256 instructions last)
    int oz_collector = 0;                   // int8 kernel: A-collector reuse (UTCIMMA .A_KEEP/.A_REUSE) between the windows of one A slice (measured: no gain -- the kernel is bound by the L2 -> SM fill, not by the shared-memory port)
    int oz_l2hint = 0;                      // L2 cache-policy variant of the int8 kernel (ozaki.cu OzArgs::l2hint)
    int oz_group = 8;                       // tile rows per raster group of the int8 kernel
    long long* oz_dbg = nullptr;            // diagnostics: clock64 timeline of CTA 0 (bgp_debug_oz_timeline)
    int oz_dbg_cap = 0;
    int gemm_cfg = 0;                       // 0 = default big-tile config, else forced variant (probing)
    int64_t launches = 0;
    // bgp_ctx_kernel_profile: timed CUDA events around every trailing-update launch of bgp_potrf (bench.py's roofline)
    int prof = 0;
    std::vector<cudaEvent_t> prof_ev;       // pairs (start, stop), re-used across calls
    std::vector<double> prof_flop;          // algorithmic flop of launch i
    size_t prof_used = 0;                   // launches recorded since the last read
};

void set_error(const char* what, cudaError_t e);

#define BGP_CUDA_OK(call)                                                                   \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) { ::bgp::set_error(#call, e__); return BGP_E_CUDA; }        \
    } while (0)

#define BGP_LAUNCH_OK(ctx)                                                                  \
    do {                                                                                    \
        (ctx)->launches++;                                                                  \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) { ::bgp::set_error("kernel launch", e__); return BGP_E_CUDA; } \
    } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may be scheduled while its predecessor on the
// stream is still running; it must execute pdl_wait() before touching anything the predecessor produces (or consumes) --
// every kernel launched through launch_pdl() does so as its first statement, so only launch latency overlaps.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// ---- gemm_nt.cu
struct GemmArgs {
    const double* A; int64_t lda;
    const double* B; int64_t ldb;
    double* C; int64_t ldc;
    int M, N, K;
    double alpha, beta;
    int tri;                 // 0: full; 1: only (col + coff) <= (row + roff)
    int64_t roff, coff;
    int kskip = 0;           // 1: A[i][k] == 0 for k < i + kofs (A upper-triangular): start K at the tile's first row
    int64_t kofs = 0;
};
int gemm_nt(Ctx* ctx, const GemmArgs& g, cudaStream_t st);
// cfg: 0 auto, 1 = 128x128 tiles, 2 = 64x128 (C may alias A when N <= 128), 3 = 64x64
int gemm_nt_cfg(Ctx* ctx, const GemmArgs& g, int cfg, cudaStream_t st);

// ---- potrf.cu
int oz_gemm_kchunked(Ctx* ctx, const double* A, int64_t lda, int64_t M, const double* B, int64_t ldb, int64_t N, int64_t K,
                     double alpha, double* C, int64_t ldc, int tri, bool a_upper, bool same_ab, cudaStream_t st, int tiles_per_cta,
                     int* rc);
int potrf_rec(Ctx* ctx, double* A, int64_t n, int64_t lda, double* dinv, int64_t gofs, cudaStream_t st);
int trsm_rlt_rec(Ctx* ctx, const double* L, int64_t n, int64_t ldl, const double* dinv, double* X, int64_t m,
                 int64_t ldx, cudaStream_t st);

}  // namespace bgp
