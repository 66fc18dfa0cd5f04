// DRAFT -- NOT BUILT (battgp_b200/build.py SOURCES does not list it), never run on a GPU.
// Compile check:  cd battgp_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -I../../include -I. -c next/ozaki2_mma_mc.cu
//
// Step 2 of the modular int8 kernel: next/ozaki2_mma.cu (parity-green) + OPERAND MULTICAST, nothing else
// changed, so that a failure can only come from the cluster plumbing.
//
// A 2x2 cluster (cluster dim x = 4; rank = 2 r + c) owns a 256 x 512 super-tile: CTA (r, c) computes the 128 x 256 tile at
// tile row 2 R + r, tile column 2 Cc + c.
//   A stage (2 planes x 8 KB) is the same for (r,0) and (r,1): each loads ONE plane... no -- each loads HALF of every plane
//     (rows 64 c .. 64 c + 63: 4 KB, contiguous in the swizzled image because a 128-row plane is 16 atoms of 8 rows x 64 B)
//     and multicasts it to both; B stage (2 planes x 16 KB) is the same for (0,c) and (1,c): CTA (r,c) loads 128-row block r of
//     each plane (8 KB) and multicasts it to both.  Fill per CTA and k-step: 2 x 4 + 2 x 8 = 24 KB instead of 48 KB.
//   full barrier  : unchanged -- expect_tx counts the bytes landing in THIS CTA's smem, whoever sent them.
//   empty barrier : count 4 -- every CTA's MMA warp commits with a multicast arrive on the empty barriers of all four CTAs, so a
//     producer only overwrites a stage (in its own and its partners' smem) once the whole cluster has consumed it.
//   No per-CTA early exit: a CTA whose tile is above the diagonal still runs the k-loops (its partners multicast into its
//     smem and wait for its arrivals); only a super-tile that is entirely above the diagonal is skipped by all four.
//   cluster barrier before the first multicast (peers' mbarriers initialised) and before exit (no peer still targets us).
#include "ozaki2_mma.cu"

namespace bgp {
namespace oz2draft {

__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// grid = 4 * super_m * super_n CTAs (cluster dimension 4), super_m = ceil(tiles_m / 2), super_n = ceil(tiles_n / 2);
// the operand images and T must be padded to whole super-tiles (rows of A to 256, rows of B to 512).
__global__ void __launch_bounds__(192, 1) mma_mc_kernel(Args g, int super_m, int super_n) {
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int r = (int)(crank >> 1), c = (int)(crank & 1);
    const int sid = blockIdx.x >> 2;
    const int sm_ = sid % super_m, sn_ = sid / super_m;
    const int m0 = (2 * sm_ + r) * BM, n0 = (2 * sn_ + c) * BN;
    // whole super-tile above the diagonal: all four CTAs leave together (uniform per cluster)
    if (g.tri && ((int64_t)(2 * sn_) * BN + g.coff > (int64_t)(2 * sm_) * BM + 2 * BM - 1 + g.roff)) return;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES);
    const uint32_t tfull = smem_u32(bars + 2 * STAGES), tempty = smem_u32(bars + 2 * STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 4); }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                        // peers' barriers are initialised
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int KB = g.K / BK;
    const uint16_t mask_a = (uint16_t)(3u << (2 * r));         // (r,0), (r,1)
    const uint16_t mask_b = (uint16_t)((1u << c) | (1u << (2 + c)));     // (0,c), (1,c)

    if (warp == 0) {
        const int64_t arb = (g.arow0 + m0) >> 7;
        const int64_t brb = (g.brow0 + n0) >> 7;               // first of the tile's two 128-row blocks of B
        uint32_t it = 0;
        for (int pass = 0; pass < PASSES; pass++)
            for (int kb = 0; kb < KB; kb++, it++) {
                const int st = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(empty0 + 8 * st, ph ^ 1);            // the whole cluster has consumed this stage
                if (elect_one()) {
                    mbar_expect_tx(full0 + 8 * st, A_STAGE + B_STAGE);
#pragma unroll
                    for (int pl = 0; pl < 2; pl++) {
                        // A: rows 64 c .. 64 c + 63 of plane pl (8 atoms of 512 B = 4096 B at offset c * 4096) -> (r,0) and (r,1)
                        bulk_g2s_mc(smem_u32(sA + st * A_STAGE + pl * TILE + c * 4096),
                                    g.sa + (((int64_t)kb * g.nrb_a + arb) * NMOD + 2 * pass + pl) * TILE + c * 4096, 4096,
                                    full0 + 8 * st, mask_a);
                        // B: 128-row block r of plane pl -> (0,c) and (1,c)
                        const int64_t b = (brb + r < g.nrb_b) ? brb + r : g.nrb_b - 1;
                        bulk_g2s_mc(smem_u32(sB + st * B_STAGE + pl * 2 * TILE + r * TILE),
                                    g.sb + (((int64_t)kb * g.nrb_b + b) * NMOD + 2 * pass + pl) * TILE, TILE, full0 + 8 * st, mask_b);
                    }
                }
                __syncwarp();
            }
    } else if (warp == 1) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        const uint64_t dzero = desc_sw64(0);
        uint32_t it = 0;
        for (int pass = 0; pass < PASSES; pass++) {
            mbar_wait(tempty, (pass & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kb = 0; kb < KB; kb++, it++) {
                const int st = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(full0 + 8 * st, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t da0 = dzero + (uint64_t)(smem_u32(sA + st * A_STAGE) >> 4);
                    const uint64_t db0 = dzero + (uint64_t)(smem_u32(sB + st * B_STAGE) >> 4);
#pragma unroll
                    for (int ks = 0; ks < BK / 32; ks++)
#pragma unroll
                        for (int pl = 0; pl < 2; pl++)
                            mma_i8(tmem_base + (uint32_t)pl * BN, da0 + (uint64_t)((pl * TILE + ks * 32) >> 4),
                                   db0 + (uint64_t)((pl * 2 * TILE + ks * 32) >> 4), idesc, (kb > 0 || ks > 0) ? 1u : 0u);
                    commit_mc(empty0 + 8 * st, (uint16_t)0xF);  // arrive on the empty barrier of all four CTAs
                    if (kb == KB - 1) commit(tfull);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        for (int pass = 0; pass < PASSES; pass++) {
            mbar_wait(tfull, pass & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int pl = 0; pl < 2; pl++) {
                const int plane = 2 * pass + pl;
                uint8_t* trow = g.T + ((int64_t)plane * g.tm + row) * g.tn + n0;
#pragma unroll 1
                for (int h = 0; h < BN / 32; h++) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pl * BN + h * 32), v);
                    uint32_t pk[8];
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        pk[i] = mod_plane<0>(plane, (int)v[4 * i]) | (mod_plane<0>(plane, (int)v[4 * i + 1]) << 8) |
                                (mod_plane<0>(plane, (int)v[4 * i + 2]) << 16) | (mod_plane<0>(plane, (int)v[4 * i + 3]) << 24);
                    uint4* dst = reinterpret_cast<uint4*>(trow + h * 32);
                    dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty) : "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                        // no peer still multicasts into / arrives on this CTA
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// Host side (to be merged into oz2_gemm): pad mp to 256 and np to 512, launch with
//   cudaLaunchAttributeClusterDimension {4,1,1}, grid = 4 * super_m * super_n, 192 threads, the same dynamic smem as mma_kernel.

}  // namespace oz2draft
}  // namespace bgp
