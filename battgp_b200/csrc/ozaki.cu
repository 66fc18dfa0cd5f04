// FP64-accurate NT GEMM on the 5th-gen tensor cores by integer slicing (Ozaki scheme I).
//
// tcgen05 has no f64 kind, so the FP64 DMMA pipe (37 TF/s) bounds gemm_nt.cu.  Here every fp64 operand row is scaled by a
// power of two 2^e (row maximum), rounded to a 56-bit signed integer q = rint(a 2^(55-e)) and written in radix 256 with
// SEVEN signed 8-bit digits (two's-complement radix-256: q = sum_s d_s 256^(6-s), d_s in [-128, 127] -- the digits are just
// the bytes of q + 0x80..80 with their top bits flipped),
//     a = 2^e * ( d0/2^7 + d1/2^15 + ... + d6/2^55 )            (55 bits + sign, every step exact),
// and  A B^T = sum_{s,t} 2^(eA_i + eB_j - 14 - 8(s+t)) (A_s B_t^T)  is evaluated with EXACT int8 x int8 -> int32
// tcgen05.mma (kind::i8) products.  Pairs with s+t > 6 are dropped (<= 6 K 2^-56 relative to rowscale*colscale, the size of
// fp64's own rounding), leaving 28 products whose partial sums with equal s+t share one int32 accumulator in TMEM
// (7 accumulators x 64 columns of a 128 x 64 tile; |sum| <= 7 K 2^14 < 2^31 for K <= OZ_MAX_K).  The epilogue reads the
// accumulators with tcgen05.ld, recombines them in fp64 and applies C += alpha * (...).
// (Round 1 used eight balanced 7-bit digits = 36 products for the same 55 bits.)
//
// Kernel structure (persistent, 224 threads, one CTA per SM or a few tiles per CTA):
//   warp 0  : producer   -- cp.async.bulk (TMA bulk engine) copies of pre-swizzled slice tiles into a 2-stage smem ring,
//                           completion on mbarriers (expect_tx); with clusters the A stage is multicast to the N-adjacent CTAs
//   warp 1  : MMA issuer -- one elected lane issues 20 wide tcgen05.mma per 64-wide k-block, tcgen05.commit frees the stage
//   warps 2-5: epilogue  -- software-pipelined tcgen05.ld of the accumulators (one TMEM lane quadrant each), fp64
//                           recombination, TMEM released, then the C update (C was prefetched into L2 while the MMAs ran)
//   warp 6  : relay      -- waits on the ring's `full` mbarriers and releases the MMA warp through a named barrier (bar.sync),
//                           so that the issuing warp never has an mbarrier wait queued behind its UTCIMMAs
// The slice kernel writes the operand slices to global memory already in the 64-byte-swizzled shared-memory image the
// UMMA descriptors expect, in [k-block][128-row block][slice] order, so a tile's slices are contiguous bulk copies.
#include <climits>
#include "common.cuh"

namespace bgp {

constexpr int OZ_S = 7;
constexpr int OZ_BM = 128, OZ_BN = 64, OZ_BK = 64;
constexpr int OZ_STAGES = 2;
constexpr int OZ_THREADS = 224;                       // warp 0 producer, 1 MMA issuer, 2-5 epilogue, 6 relay (see oz_mma_kernel)
constexpr int OZ_A_STAGE = OZ_S * OZ_BM * OZ_BK;     // 57344 B
constexpr int OZ_B_STAGE = OZ_S * OZ_BN * OZ_BK;     // 28672 B
constexpr int OZ_SLICE_TILE = OZ_BM * OZ_BK;         // 8192 B: one slice of a 128-row block for one k-block
constexpr int OZ_MAX_K = 16384;                      // 7 pairs * K * 2^14 must stay below 2^31
constexpr long long OZ_DIGIT_BIAS = 0x0080808080808080ll;

// ------------------------------------------------------------------------------------------------ slicing
__device__ __forceinline__ int oz_swz64(int r8, int kb64) {
    // Swizzle<2,4,3>: 16-byte chunk index (bits 4-5) ^= row bits 1-2 (address bits 7-8)
    const int chunk = (kb64 >> 4) ^ ((r8 >> 1) & 3);
    return r8 * 64 + chunk * 16 + (kb64 & 15);
}

// grid.x = ceil(rows_pad / 8); block = 512 threads = 8 rows x 64 threads (2 warps walk a row in 1 KB steps)
__global__ void __launch_bounds__(512)
oz_slice_kernel(const double* __restrict__ P, int64_t rows, int64_t K, int64_t ld, int8_t* __restrict__ out,
                double* __restrict__ ex, int64_t nrb, const int32_t* __restrict__ blkmap, int64_t blkrows) {
    __shared__ unsigned int smax[8];
    const int tid = threadIdx.x, r8 = tid >> 6, ct = tid & 63;
    const int64_t row = (int64_t)blockIdx.x * 8 + r8;
    // optional gather: logical row block b (blkrows rows) is read from source block blkmap[b] -- lets the sharded GP slice the
    // rank-major all-gather buffer straight into stripe order without a reordering copy
    const int64_t srow = (blkmap != nullptr && row < rows) ? (int64_t)blkmap[row / blkrows] * blkrows + row % blkrows : row;
    if (ct == 0) smax[r8] = 0u;
    __syncthreads();
    const int64_t ngroups = K / 4;                 // a thread handles groups of 4 consecutive values (32 B): a warp's two
    const bool live = row < rows;                  // 16-byte loads cover 1 KB of the row contiguously
    // row maximum of |a| from the HIGH words of the bit patterns (sign cleared): they order non-negative doubles down to the
    // top 20 mantissa bits, which is all the scale needs, AND let Inf / NaN win (fmax would drop a NaN), so a poisoned row is
    // detected instead of being sliced into garbage digits.  One AND + one integer max per value.
    unsigned int mb = 0u;
    if (live) {
#pragma unroll 4
        for (int64_t g = ct; g < ngroups; g += 64) {
            const uint4* p = reinterpret_cast<const uint4*>(P + srow * ld + g * 4);
            const uint4 v0 = p[0], v1 = p[1];
            mb = max(max(mb, max(v0.y & 0x7FFFFFFFu, v0.w & 0x7FFFFFFFu)), max(v1.y & 0x7FFFFFFFu, v1.w & 0x7FFFFFFFu));
        }
    }
    atomicMax(&smax[r8], mb);
    __syncthreads();
    const unsigned int rb_hi = smax[r8];
    const bool poisoned = (rb_hi >> 20) == 0x7FFu;                          // Inf or NaN somewhere in the row
    // upper bound of the row maximum given its high word (low word all ones); a row whose high words are all zero
    // (|a| < 2^-1042 everywhere) is treated as a zero row
    const double rowmax = rb_hi ? __hiloint2double((int)rb_hi, (int)0xFFFFFFFFu) : 0.0;
    int e = 0;
    if (!poisoned && rowmax > 0.0) {
        const double f = frexp(rowmax, &e);                                 // rowmax = f * 2^e, f in [0.5, 1)
        if (f > 0.99) e += 1;                                               // digits reach +127/128 (1 + 1/256 + ...) = 0.996 only
    }
    // row scale 2^e; NaN for a poisoned row, so that every product involving it comes out NaN like in fp64 arithmetic
    if (ct == 0 && row < nrb * 128) ex[row] = poisoned ? __longlong_as_double(0x7FF8000000000000ll) : scalbn(1.0, e);
    const int sh = 55 - e, sh1 = sh / 2;
    const double sc1 = __longlong_as_double((long long)(1023 + sh1) << 52);            // 2^sh1, |sh1| <= 540
    const double sc2 = __longlong_as_double((long long)(1023 + (sh - sh1)) << 52);     // 2^(sh - sh1)
    const int64_t rb = row >> 7;
    const int rin = (int)(row & 127);
#pragma unroll 2
    for (int64_t g = ct; g < ngroups; g += 64) {
        uint32_t w[OZ_S];
#pragma unroll
        for (int s = 0; s < OZ_S; s++) w[s] = 0u;
        if (live && !poisoned) {
            const double2* p = reinterpret_cast<const double2*>(P + srow * ld + g * 4);      // 16-byte aligned: ld even, P aligned
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int i2 = 0; i2 < 2; i2++) {
                const double2 v = p[i2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    // q = rint(a 2^(55-e)), |q| <= 0.99 2^55.  The power of two is applied as two exact multiplications
                    // (sc1 sc2 = 2^(55-e), each within the normal range for every e): bit-identical to scalbn -- the
                    // only inexact case is a product that underflows, and those values round to q = 0 either way.
                    const long long q = __double2ll_rn(((h == 0) ? v.x : v.y) * sc1 * sc2);
                    const unsigned long long dq = (unsigned long long)((q + OZ_DIGIT_BIAS) ^ OZ_DIGIT_BIAS);   // byte 6-s = digit s
                    lo[i2 * 2 + h] = (uint32_t)dq;
                    hi[i2 * 2 + h] = (uint32_t)(dq >> 32);
                }
            }
            // byte b = 6 - s of the four values -> bytes 0..3 of the plane's word: three byte permutes (PRMT) per plane
#pragma unroll
            for (int s = 0; s < OZ_S; s++) {
                const int b = OZ_S - 1 - s;
                const uint32_t sel = (uint32_t)((b & 3) | (((b & 3) + 4) << 4));             // byte (b&3) of x, then of y
                const uint32_t t01 = (b >= 4) ? __byte_perm(hi[0], hi[1], sel) : __byte_perm(lo[0], lo[1], sel);
                const uint32_t t23 = (b >= 4) ? __byte_perm(hi[2], hi[3], sel) : __byte_perm(lo[2], lo[3], sel);
                w[s] = __byte_perm(t01, t23, 0x5410);
            }
        }
        const int64_t kb = (g * 4) >> 6;
        const int kin = (int)((g * 4) & 63);
        // 16 consecutive lanes fill one 64-byte row of a plane's k-block image (the swizzle permutes its 16-byte chunks)
        int8_t* base = out + ((kb * nrb + rb) * OZ_S) * (int64_t)OZ_SLICE_TILE + (rin >> 3) * 512 + oz_swz64(rin & 7, kin);
#pragma unroll
        for (int s = 0; s < OZ_S; s++) *reinterpret_cast<uint32_t*>(base + (int64_t)s * OZ_SLICE_TILE) = w[s];
    }
}

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a wrong descriptor must end in a trap (CUDA error), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 24); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
// the same for waits that are expected to be long (the epilogue waiting for a whole tile of MMAs): back off between polls so
// that the pollers stay out of the MIO queue the UTCIMMAs and their operand reads go through
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, uint32_t ns) {
    for (uint32_t it = 0; it < (1u << 24); ++it) {
        if (mbar_try_wait(bar, parity)) return;
        __nanosleep(ns);
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// operand planes are re-read by the other CTAs of the raster group and by the next raster step: keep them in L2 ahead of
// the streaming C traffic
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
    // K-major, SWIZZLE_64B (layout type 4), SBO = 512 B (8 rows x 64 B), LBO = 0, descriptor version 1 (sm_100)
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// A-collector forms: KEEP loads the A operand into the tensor core's collector buffer and keeps it, REUSE takes it from there
// (no shared-memory read of A) and releases it.  SASS: UTCIMMA gdesc[..].A_KEEP / .A_REUSE.
__device__ __forceinline__ void oz_mma_i8_keep(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void oz_mma_i8_reuse(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void oz_named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void oz_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void oz_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
// asynchronous TMEM load of 32 columns of this warp's lane quadrant; the registers are valid after tmem_ld_wait(v)
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
// waits for every outstanding tcgen05.ld of this thread; the "+r" operands tie the consumers of v to the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                   "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                   "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.u32 %0, 1, 0, px;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
// exact int32 -> double without the (slow) I2F.F64 path: 2^52 + 2^31 + v has v + 2^31 in its low mantissa word
__device__ __forceinline__ double i32_to_f64(uint32_t v) {
    return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;   // 2^52 + 2^31
}

struct OzArgs {
    const int8_t* sa; int64_t nrb_a; int64_t arow0;     // slices of A, 128-row blocks in its buffer, first row (mult. of 128)
    const int8_t* sb; int64_t nrb_b; int64_t brow0;     // slices of B, first row (multiple of 64)
    const double* exa; const double* exb;               // per-row scales 2^e, indexed like the slice buffers
    double* C; int64_t ldc;
    int M, N, K;
    double alpha;
    int tri; int64_t roff, coff;
    int debug_noload;      // experiment: only the first OZ_STAGES k-blocks are really loaded
    int group;             // raster: tile rows per group (oz_decode)
    int kfence;            // 1: tcgen05.fence::after_thread_sync after every stage wait (0: only after the TMEM-empty wait of a tile)
    int dbg_epi;           // experiments (wrong results): 1 = no store phase, 2 = smem transpose only, 3 = global traffic only
    int backoff;           // 1: producer / relay / epilogue warps sleep between failed polls of their mbarriers
    int relay;             // 1: a relay warp watches the `full` barriers and releases the MMA warp through a named barrier
    int order;             // MMA issue order within a k-block: 0 = by A slice, 1 = widest last per k-step, 2 = seven widest last per k-block
    int collector;         // 1: A-collector reuse between the two MMA windows of an A slice
    int l2hint;            // 0: default L2 policy; 1: operand loads evict_last; 2: + streaming C accesses; 3: as 2, no C prefetch
    int64_t brb_max;       // last valid 128-row block of B (cluster tiles past N read a valid block; stores are masked)
    long long* dbg;        // diagnostics: per-tile clock64 stamps of CTA 0 (bgp_debug_oz_timeline), 16 slots per tile
    int dbg_cap;
};

// Grouped raster over (tile row, group of CS tile columns): consecutive positions walk down `group` tile rows, then move one
// column group to the right, so concurrently running CTAs share A row panels and B column panels in L2.  This CTA's own tile
// is column tng*CS + rank.  A position is skipped when its FIRST tile lies above the diagonal (tri); the other CTAs of a
// live cluster position always run (their stores are masked).
template <int CS>
__device__ __forceinline__ bool oz_decode(const OzArgs& g, int pid, int tiles_m, int tiles_ng, int rank, int& m0, int& n0) {
    const int GROUP = g.group;
    const int per_group = GROUP * tiles_ng;
    const int gid = pid / per_group;
    const int first_m = gid * GROUP;
    const int gsize = min(tiles_m - first_m, GROUP);
    const int tm = first_m + (pid % per_group) % gsize;
    const int tng = (pid % per_group) / gsize;
    m0 = tm * OZ_BM;
    n0 = (tng * CS + rank) * OZ_BN;
    const int nfirst = tng * CS * OZ_BN;
    return !(g.tri && ((int64_t)nfirst + g.coff > (int64_t)m0 + OZ_BM - 1 + g.roff));
}

constexpr int OZ_TBUF = 32 * 33;                  // doubles per epilogue warp (32 x 32 transpose, padded)
// diagnostics: CTA 0 stamps clock64 per tile (slot layout in tools/oz_probe.py)
#define OZ_STAMP(slot) do { if (dbg_on && lane == 0 && tile_it < (uint32_t)g.dbg_cap) g.dbg[(size_t)tile_it * 16 + (slot)] = clock64(); } while (0)

// The tile loop.  CS = 1: plain persistent CTA (tiles_per_cta > 0: consecutive raster positions, so CTAs keep retiring and the
// high-priority panel stream of the look-ahead Cholesky finds free SMs; tiles_per_cta == 0: grid-stride, one CTA per SM).
// CS = 2 / 4: clusters of N-adjacent CTAs share the A operand -- each CTA loads 1/CS of the A stage and MULTICASTS it to the
// whole cluster, so the L2 -> SM traffic per CTA and k-block drops from 84 KB to 28 + 56/CS KB.  A stage may only be
// overwritten once EVERY CTA of the cluster has consumed it: the MMA warps commit with a multicast arrive on the `empty`
// barriers of all CS CTAs (count = CS).  Always grid-stride.
template <int CS>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_mma_kernel(OzArgs g, int tiles_m, int tiles_n, int tiles_per_cta) {
    uint32_t crank = 0;
    if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int tiles_ng = (tiles_n + CS - 1) / CS;               // column groups per tile row
    const int total = tiles_m * tiles_ng;
    const int ncl = (int)gridDim.x / CS, cid = (int)blockIdx.x / CS;
    const int pid_step = (CS == 1 && tiles_per_cta > 0) ? 1 : ncl;
    const int pid_begin = (CS == 1 && tiles_per_cta > 0) ? cid * tiles_per_cta : cid;
    const int pid_end = (CS == 1 && tiles_per_cta > 0) ? min(total, pid_begin + tiles_per_cta) : total;
    if (CS == 1) {
        // leave before touching TMEM if none of this CTA's positions is a real tile (cluster CTAs never leave early:
        // peers multicast into their shared memory)
        bool any = false;
        for (int pid = pid_begin; pid < pid_end && !any; pid += pid_step) { int a, b; any = oz_decode<CS>(g, pid, tiles_m, tiles_ng, 0, a, b); }
        if (!any) return;
    }
    extern __shared__ uint8_t oz_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(oz_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                       // [stage][7 slices][128 x 64 B]
    uint8_t* sB = smem + OZ_STAGES * OZ_A_STAGE;              // [stage][7 slices][64 x 64 B]
    double* tbuf_all = reinterpret_cast<double*>(sB + OZ_STAGES * OZ_B_STAGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tbuf_all + 4 * OZ_TBUF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + OZ_STAGES);
    const uint32_t tfull = smem_u32(bars + 2 * OZ_STAGES), tempty = smem_u32(bars + 2 * OZ_STAGES + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < OZ_STAGES; i++) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, CS); }
        mbar_init(tfull, 1);
        mbar_init(tempty, 4);                                   // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CS > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // peers' barriers are initialised
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int KB = g.K / OZ_BK;
    const bool dbg_on = g.dbg != nullptr && blockIdx.x == 0;

    // Producer and MMA warps run their loops warp-uniformly (all 32 lanes wait on the barriers, one elected lane
    // issues): operands then live in uniform registers and no per-lane "waterfall" code is generated around UTCIMMA.
    if (warp == 0) {
        uint32_t it = 0, tile_it = 0;                           // k-block counter across tiles
        const uint64_t pol = l2_policy_evict_last();
        for (int pid = pid_begin; pid < pid_end; pid += pid_step) {
            int m0, n0;
            if (!oz_decode<CS>(g, pid, tiles_m, tiles_ng, (int)crank, m0, n0)) continue;
            OZ_STAMP(8);
            const int64_t arb = (g.arow0 + m0) >> 7;
            const int64_t brb = min((g.brow0 + n0) >> 7, g.brb_max);
            const int bhalf = (int)(((g.brow0 + n0) >> 6) & 1);
            for (int kb = 0; kb < KB; kb++, it++) {
                const int st = it % OZ_STAGES;
                const uint32_t ph = (it / OZ_STAGES) & 1;
                if (g.backoff) mbar_wait_backoff(empty0 + 8 * st, ph ^ 1, 64);
                else mbar_wait(empty0 + 8 * st, ph ^ 1);
                if (elect_one()) {
                    if (CS == 1 && g.debug_noload && it >= OZ_STAGES) {
                        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + 8 * st) : "memory");
                    } else {
                        mbar_expect_tx(full0 + 8 * st, OZ_A_STAGE + OZ_B_STAGE);
                        const int8_t* asrc = g.sa + ((int64_t)kb * g.nrb_a + arb) * OZ_S * OZ_SLICE_TILE;
                        if (CS == 1) {
                            if (g.l2hint) bulk_g2s_hint(smem_u32(sA + st * OZ_A_STAGE), asrc, OZ_A_STAGE, full0 + 8 * st, pol);
                            else bulk_g2s(smem_u32(sA + st * OZ_A_STAGE), asrc, OZ_A_STAGE, full0 + 8 * st);
                        } else {
                            constexpr int APART = OZ_A_STAGE / CS;  // this CTA's share of the A stage, delivered to all CS CTAs
                            bulk_g2s_mc(smem_u32(sA + st * OZ_A_STAGE + crank * APART), asrc + (int64_t)crank * APART, APART,
                                        full0 + 8 * st, (uint16_t)((1u << CS) - 1));
                        }
                        const int8_t* bsrc = g.sb + ((int64_t)kb * g.nrb_b + brb) * OZ_S * OZ_SLICE_TILE + bhalf * 4096;
                        if (g.l2hint) {
#pragma unroll
                            for (int s = 0; s < OZ_S; s++)
                                bulk_g2s_hint(smem_u32(sB + st * OZ_B_STAGE + s * 4096), bsrc + (int64_t)s * OZ_SLICE_TILE, 4096, full0 + 8 * st, pol);
                        } else {
#pragma unroll
                            for (int s = 0; s < OZ_S; s++)
                                bulk_g2s(smem_u32(sB + st * OZ_B_STAGE + s * 4096), bsrc + (int64_t)s * OZ_SLICE_TILE, 4096, full0 + 8 * st);
                        }
                    }
                }
                __syncwarp();
            }
            OZ_STAMP(9);
            tile_it++;
        }
    } else if (warp == 1) {
        // kind::i8, D = S32, A/B = signed int8, both K-major, M = 128.  For a fixed A slice s the B slices t = 0..6-s are
        // adjacent 64-row tiles in shared memory AND their accumulators c = s+t are adjacent 64-column blocks in TMEM, so
        // they are issued as ONE wide MMA (N up to 256): 10 instead of 28 instructions per k-step and, more importantly,
        // each A slice is read from shared memory 1-2 times instead of 7-s times (the 128 B/cycle smem port, not the
        // tensor pipe, limits narrow MMAs: tools/microbench/i8_peak.cu).
        const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
        const uint64_t dzero = oz_desc(0);
        uint32_t it = 0, tile_it = 0;
        for (int pid = pid_begin; pid < pid_end; pid += pid_step) {
            int m0, n0;
            if (!oz_decode<CS>(g, pid, tiles_m, tiles_ng, (int)crank, m0, n0)) continue;
            OZ_STAMP(0);
            mbar_wait(tempty, (tile_it & 1) ^ 1);                // epilogue has drained the previous tile's accumulators
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            OZ_STAMP(1);
            for (int kb = 0; kb < KB; kb++, it++) {
                const int st = it % OZ_STAGES;
                const uint32_t ph = (it / OZ_STAGES) & 1;
                // The stage's `full` barrier is watched by the relay warp (warp 6), which releases this warp through a named
                // barrier.  An mbarrier wait issued by THIS warp sits in its in-order MIO queue behind the 20 UTCIMMAs of
                // the previous k-block and stalls the issue of the next ones by 120-250 cycles per k-block; bar.sync does
                // not (tools/microbench/issue_wait.cu: 2186 vs 2069 cycles per k-block, named barrier = no sync at all).
                if (g.relay) oz_named_sync(1, 64);
                else mbar_wait(full0 + 8 * st, ph);
                if (g.kfence) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (kb == 0) OZ_STAMP(2);
                const uint64_t da0 = dzero + (uint64_t)(smem_u32(sA + st * OZ_A_STAGE) >> 4);
                const uint64_t db0 = dzero + (uint64_t)(smem_u32(sB + st * OZ_B_STAGE) >> 4);
                if (elect_one()) {
                    if (g.order == 0) {
#pragma unroll
                        for (int ks = 0; ks < OZ_BK / 32; ks++) {
#pragma unroll
                            for (int s = 0; s < OZ_S; s++) {
                                const uint64_t da = da0 + (uint64_t)((s * OZ_SLICE_TILE + ks * 32) >> 4);
                                const uint32_t acc = (ks > 0 || s > 0) ? 1u : (kb > 0 ? 1u : 0u);
#pragma unroll
                                for (int t0 = 0; t0 + s < OZ_S; t0 += 4) {
                                    const int nt = (OZ_S - s - t0) < 4 ? (OZ_S - s - t0) : 4;     // B slices in this MMA
                                    const uint64_t db = db0 + (uint64_t)((t0 * 4096 + ks * 32) >> 4);
                                    const uint32_t idesc = idesc0 | ((uint32_t)((nt * OZ_BN) >> 3) << 17);
                                    const uint32_t dcol = tmem_base + (uint32_t)(s + t0) * OZ_BN;
                                    // the two windows of A slices 0..2 read the same A operand back to back: the second one
                                    // may take it from the collector instead of shared memory (knob, off: no gain measured)
                                    if (g.collector && s + 4 < OZ_S) {
                                        if (t0 == 0) oz_mma_i8_keep(dcol, da, db, idesc, acc);
                                        else oz_mma_i8_reuse(dcol, da, db, idesc, acc);
                                    } else {
                                        oz_mma_i8(dcol, da, db, idesc, acc);
                                    }
                                }
                            }
                        }
                    } else if (g.order == 1) {
                        // Same 10 instructions per k-step, ordered so that a k-step ends with its widest ones (N = 256).
                        constexpr int NI = 10;
                        constexpr int OS[NI] = {0, 0, 6, 5, 4, 2, 1, 1, 2, 3};
                        constexpr int OT[NI] = {0, 4, 0, 0, 0, 4, 4, 0, 0, 0};
                        constexpr int ON[NI] = {4, 3, 1, 2, 3, 1, 2, 4, 4, 4};
#pragma unroll
                        for (int ks = 0; ks < OZ_BK / 32; ks++) {
#pragma unroll
                            for (int i = 0; i < NI; i++) {
                                const uint64_t da = da0 + (uint64_t)((OS[i] * OZ_SLICE_TILE + ks * 32) >> 4);
                                const uint64_t db = db0 + (uint64_t)((OT[i] * 4096 + ks * 32) >> 4);
                                const uint32_t acc = (ks > 0 || i > 1) ? 1u : (kb > 0 ? 1u : 0u);
                                oz_mma_i8(tmem_base + (uint32_t)(OS[i] + OT[i]) * OZ_BN, da, db,
                                          idesc0 | ((uint32_t)((ON[i] * OZ_BN) >> 3) << 17), acc);
                            }
                        }
                    } else {
                        // The 20 instructions of a k-block (2 k-steps x 10 windows) ordered so that the block ENDS with seven
                        // N = 256 instructions.  Why the order matters: the MIO queue is in order, so the wait on the next
                        // stage's `full` barrier, issued after the 20 UTCIMMAs, returns only once they have all been handed
                        // to the tensor pipe, and a few hundred cycles pass before the next UTCIMMA arrives there
                        // (tools/microbench/i8_peak.cu, commit_wait modes: a fixed ~560 cycles per barrier wait, whatever
                        // the distance to the phase waited for).  The pipe runs through that gap only on the <= 3
                        // instructions it holds itself: ~140 cycles of work when the 64/128-column instructions of A slices
                        // 5 and 6 come last (order 0: 2610-2730 cycles per k-block for 1792 of MMA work), ~384 when three
                        // N = 256 instructions do (orders 1 and 2: 2120-2180).  An early non-blocking probe of the barrier in
                        // front of the last four instructions bought nothing more (profiles/probe_r02_oz_order.jsonl).
                        // Slice 0 of k-step 0 stays first: with kb = 0 its two windows initialise all 7 accumulators.
                        constexpr int NI = 20;
                        constexpr int OK[NI] = {0, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 1};
                        constexpr int OS[NI] = {0, 0, 0, 6, 6, 5, 5, 4, 4, 2, 2, 1, 1, 0, 1, 1, 2, 2, 3, 3};
                        constexpr int OT[NI] = {0, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0};
                        constexpr int ON[NI] = {4, 3, 3, 1, 1, 2, 2, 3, 3, 1, 1, 2, 2, 4, 4, 4, 4, 4, 4, 4};
#pragma unroll
                        for (int i = 0; i < NI; i++) {
                            const uint64_t da = da0 + (uint64_t)((OS[i] * OZ_SLICE_TILE + OK[i] * 32) >> 4);
                            const uint64_t db = db0 + (uint64_t)((OT[i] * 4096 + OK[i] * 32) >> 4);
                            const uint32_t acc = (i > 1) ? 1u : (kb > 0 ? 1u : 0u);
                            oz_mma_i8(tmem_base + (uint32_t)(OS[i] + OT[i]) * OZ_BN, da, db,
                                      idesc0 | ((uint32_t)((ON[i] * OZ_BN) >> 3) << 17), acc);
                        }
                    }
                    if (CS == 1) oz_commit(empty0 + 8 * st);
                    else oz_commit_mc(empty0 + 8 * st, (uint16_t)((1u << CS) - 1));
                    if (kb == KB - 1) oz_commit(tfull);
                }
                __syncwarp();
            }
            OZ_STAMP(3);
            tile_it++;
        }
    } else if (warp == 6) {
        // relay: waits for every stage to be filled (TMA complete_tx on `full`) and hands the MMA warp on through named barrier 1
        if (g.relay) {
            uint32_t it = 0;
            for (int pid = pid_begin; pid < pid_end; pid += pid_step) {
                int m0, n0;
                if (!oz_decode<CS>(g, pid, tiles_m, tiles_ng, (int)crank, m0, n0)) continue;
                for (int kb = 0; kb < KB; kb++, it++) {
                    if (g.backoff) mbar_wait_backoff(full0 + 8 * (it % OZ_STAGES), (it / OZ_STAGES) & 1, 32);
                    else mbar_wait(full0 + 8 * (it % OZ_STAGES), (it / OZ_STAGES) & 1);
                    oz_named_sync(1, 64);
                }
            }
        }
    } else {
        const int q = warp & 3;                                   // TMEM lane quadrant this warp may read
        double* tbuf = tbuf_all + (warp - 2) * OZ_TBUF;
        uint32_t tile_it = 0;
        for (int pid = pid_begin; pid < pid_end; pid += pid_step) {
            int m0, n0;
            if (!oz_decode<CS>(g, pid, tiles_m, tiles_ng, (int)crank, m0, n0)) continue;
            const int row_l = q * 32 + lane;
            const double sa = (m0 + row_l < g.M) ? g.alpha * g.exa[g.arow0 + m0 + row_l] * (1.0 / 16384.0) : 0.0;
            // pull this warp's 32 x 64 block of C towards L2 while the MMAs of the tile run (lane = row, 4 lines of 128 B)
            if (m0 + row_l < g.M && g.l2hint != 3) {
                const double* crow = g.C + (int64_t)(m0 + row_l) * g.ldc + n0;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (n0 + i * 16 < g.N && (!g.tri || ((int64_t)n0 + i * 16 + g.coff <= (int64_t)m0 + row_l + g.roff))) prefetch_l2(crow + i * 16);
            }
            if (warp == 2) OZ_STAMP(4);
            if (g.backoff) mbar_wait_backoff(tfull, tile_it & 1, 256);
            else mbar_wait(tfull, tile_it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (warp == 2) OZ_STAMP(5);
            // drain: 2 column halves x 7 accumulators, the TMEM load of step i+1 in flight while step i is recombined
            double acc[2][32];
            uint32_t v[2][32];
            const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
            tmem_ld32_async(tq, v[0]);
#pragma unroll
            for (int i = 0; i < 2 * OZ_S; i++) {
                const int h = i / OZ_S, c = i % OZ_S;
                tmem_ld_wait(v[i & 1]);
                if (i + 1 < 2 * OZ_S) tmem_ld32_async(tq + (uint32_t)(((i + 1) % OZ_S) * OZ_BN + ((i + 1) / OZ_S) * 32), v[(i + 1) & 1]);
                const double w = 1.0 / (double)(1ull << (8 * c));
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const double x = i32_to_f64(v[i & 1][j]);
                    acc[h][j] = (c == 0) ? x : fma(x, w, acc[h][j]);
                }
            }
            // all TMEM reads of this warp are done: let the MMA warp start the next tile
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty) : "memory");
            if (warp == 2) OZ_STAMP(6);
            // rows of this warp's 32-row block that exist and (tri) lie on or below the diagonal for a given column: [r_lo, 32)
            // clipped to r_hi -- two small integers instead of a 64-bit predicate per element (the epilogue warp is alone on
            // its scheduler, so its instruction count is what the store phase costs)
            const int rbase = m0 + q * 32;
            const int r_hi = min(32, g.M - rbase);
            if (g.dbg_epi == 1) { tile_it++; continue; }                               // experiment: no store phase at all
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (g.dbg_epi == 3) {                                                   // experiment: global traffic without the smem transpose
                    const int col = n0 + h * 32 + lane;
                    if (col < g.N && rbase + 32 <= g.M) {
                        double* cp = g.C + (int64_t)rbase * g.ldc + col;
                        double cold[32];
#pragma unroll
                        for (int r = 0; r < 32; r++) cold[r] = cp[(int64_t)r * g.ldc];
#pragma unroll
                        for (int r = 0; r < 32; r++) cp[(int64_t)r * g.ldc] = cold[r] + acc[h][r] * sa;
                    }
                    continue;
                }
                // TMEM hands a thread one ROW (lane), global memory wants a warp on one row segment: transpose through smem
#pragma unroll
                for (int j = 0; j < 32; j++) tbuf[lane * 33 + j] = acc[h][j] * sa;     // row scale applied here
                __syncwarp();
                if (g.dbg_epi == 2) {                                                   // experiment: the smem transpose without global traffic
                    double sum = 0.0;
#pragma unroll
                    for (int r = 0; r < 32; r++) sum += tbuf[r * 33 + lane];
                    if (sum == 12345.678) g.C[0] = sum;
                    __syncwarp();
                    continue;
                }
                // now lane = column
                const int col = n0 + h * 32 + lane;
                const bool cok = col < g.N;
                const double sb = cok ? g.exb[g.brow0 + col] : 0.0;
                int r_lo = 0;
                if (g.tri) {
                    const int64_t rmin = (int64_t)col + g.coff - g.roff - rbase;      // first row with col + coff <= row + roff
                    r_lo = rmin > 32 ? 32 : (rmin < 0 ? 0 : (int)rmin);
                }
                const int r_end = cok ? r_hi : 0;
                double* cp = g.C + (int64_t)rbase * g.ldc + col;
                double cold[32];
                if (__all_sync(0xffffffffu, r_lo == 0 && r_end == 32)) {                // interior tile: no predicates at all
                    if (g.l2hint >= 2) {                                                 // C is touched once: streaming (evict-first) accesses
#pragma unroll
                        for (int r = 0; r < 32; r++) cold[r] = __ldcs(cp + (int64_t)r * g.ldc);
#pragma unroll
                        for (int r = 0; r < 32; r++) __stcs(cp + (int64_t)r * g.ldc, cold[r] + tbuf[r * 33 + lane] * sb);
                    } else {
#pragma unroll
                        for (int r = 0; r < 32; r++) cold[r] = cp[(int64_t)r * g.ldc];
                        if (dbg_on && warp == 2 && h == 0 && cold[31] == 12345.678) OZ_STAMP(13);   // never true: orders stamp 10 after the loads
                        if (warp == 2 && h == 0) OZ_STAMP(10);
#pragma unroll
                        for (int r = 0; r < 32; r++) cp[(int64_t)r * g.ldc] = cold[r] + tbuf[r * 33 + lane] * sb;
                        if (warp == 2 && h == 0) OZ_STAMP(11);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 32; r++) cold[r] = (r >= r_lo && r < r_end) ? cp[(int64_t)r * g.ldc] : 0.0;
#pragma unroll
                    for (int r = 0; r < 32; r++)
                        if (r >= r_lo && r < r_end) cp[(int64_t)r * g.ldc] = cold[r] + tbuf[r * 33 + lane] * sb;
                }
                __syncwarp();
            }
            if (warp == 2) OZ_STAMP(7);
            tile_it++;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CS > 1) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");   // no peer still targets this CTA
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------ host side
static inline int64_t oz_rows_pad(int64_t rows) { return ((rows + 127) / 128) * 128; }

int64_t oz_slice_buffer_bytes(int64_t rows, int64_t K) { return oz_rows_pad(rows) * K * OZ_S + oz_rows_pad(rows) * (int64_t)sizeof(double); }

// slices rows x K of P into buf (digits) + exponents stored right behind them
int oz_slice(Ctx* ctx, const double* P, int64_t rows, int64_t K, int64_t ld, void* buf, cudaStream_t st, const int32_t* blkmap,
             int64_t blkrows) {
    if (K % OZ_BK != 0 || (ld & 1) || ((uintptr_t)P & 15)) return BGP_E_ARG;
    const int64_t rp = oz_rows_pad(rows), nrb = rp / 128;
    int8_t* dig = reinterpret_cast<int8_t*>(buf);
    double* ex = reinterpret_cast<double*>(dig + rp * K * OZ_S);
    oz_slice_kernel<<<(unsigned)(rp / 8), 512, 0, st>>>(P, rows, K, ld, dig, ex, nrb, blkmap, blkrows > 0 ? blkrows : 1);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

int oz_gemm(Ctx* ctx, const void* bufA, int64_t rowsA_total, int64_t arow0, const void* bufB, int64_t rowsB_total, int64_t brow0,
            int64_t M, int64_t N, int64_t K, double alpha, double* C, int64_t ldc, int tri, int64_t roff, int64_t coff,
            cudaStream_t st, int tiles_per_cta) {
    if (M <= 0 || N <= 0) return 0;
    // K is bounded by the int32 accumulators: 7 digit pairs of at most 2^14 each per k
    if (K % OZ_BK != 0 || K > OZ_MAX_K || (arow0 & 127) || (brow0 & 63)) return BGP_E_ARG;
    const int64_t rpa = oz_rows_pad(rowsA_total), rpb = oz_rows_pad(rowsB_total);
    OzArgs g;
    g.sa = reinterpret_cast<const int8_t*>(bufA); g.nrb_a = rpa / 128; g.arow0 = arow0;
    g.sb = reinterpret_cast<const int8_t*>(bufB); g.nrb_b = rpb / 128; g.brow0 = brow0;
    g.exa = reinterpret_cast<const double*>(g.sa + rpa * K * OZ_S);
    g.exb = reinterpret_cast<const double*>(g.sb + rpb * K * OZ_S);
    g.C = C; g.ldc = ldc; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.alpha = alpha; g.tri = tri; g.roff = roff; g.coff = coff;
    g.debug_noload = (ctx->gemm_cfg == 7) ? 1 : 0;
    g.group = ctx->oz_group > 0 ? ctx->oz_group : 8;
    g.l2hint = ctx->oz_l2hint;
    g.collector = ctx->oz_collector;
    g.order = ctx->oz_order;
    g.relay = ctx->oz_relay;
    g.backoff = ctx->oz_backoff;
    g.dbg_epi = ctx->oz_dbg_epi;
    g.kfence = ctx->oz_kfence;
    g.brb_max = g.nrb_b - 1;
    g.dbg = ctx->oz_dbg; g.dbg_cap = ctx->oz_dbg_cap;
    const int tiles_m = (int)((M + OZ_BM - 1) / OZ_BM), tiles_n = (int)((N + OZ_BN - 1) / OZ_BN);
    const uint64_t bit = 1ull << (ctx->device & 63);
    constexpr int SMEM = OZ_STAGES * (OZ_A_STAGE + OZ_B_STAGE) + 4 * OZ_TBUF * (int)sizeof(double) + 1024 + 256;
    static thread_local uint64_t attr_done = 0;
    static thread_local int sm_count[64] = {0};
    if (!(attr_done & bit)) {
        BGP_CUDA_OK(cudaFuncSetAttribute(oz_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        BGP_CUDA_OK(cudaFuncSetAttribute(oz_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        BGP_CUDA_OK(cudaFuncSetAttribute(oz_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        BGP_CUDA_OK(cudaDeviceGetAttribute(&sm_count[ctx->device & 63], cudaDevAttrMultiProcessorCount, ctx->device));
        attr_done |= bit;
    }
    // tiles per CTA: enough to amortise the prologue and pipeline across tiles, few enough that CTAs keep retiring so
    // the high-priority panel stream of the look-ahead Cholesky still finds free SMs (a fully persistent grid would
    // hold every SM until the whole update is done).  tiles_per_cta <= 0 -> fully persistent (one CTA per SM).
    const int nsm = sm_count[ctx->device & 63];
    const int tpc = tiles_per_cta > 0 ? tiles_per_cta : 0;
    const int cs = (tpc == 0) ? ctx->oz_cluster : 1;
    if (cs == 2 || cs == 4) {
        auto kern = (cs == 2) ? oz_mma_kernel<2> : oz_mma_kernel<4>;
        const int tiles_ng = (tiles_n + cs - 1) / cs;
        const int total_g = tiles_m * tiles_ng;
        int ncl = nsm / cs;
        if (ncl > total_g) ncl = total_g;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(ncl * cs));
        cfg.blockDim = dim3(OZ_THREADS);
        cfg.dynamicSmemBytes = SMEM;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        BGP_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, g, tiles_m, tiles_n, 0));
        ctx->launches++;
        return 0;
    }
    const int total = tiles_m * tiles_n;
    int nfree = nsm - (tpc == 0 ? ctx->oz_reserve_now : 0);
    if (nfree < 1) nfree = 1;
    const int grid = tpc > 0 ? (total + tpc - 1) / tpc : (total < nfree ? total : nfree);
    oz_mma_kernel<1><<<grid, OZ_THREADS, SMEM, st>>>(g, tiles_m, tiles_n, tpc);
    BGP_LAUNCH_OK(ctx);
    return 0;
}

}  // namespace bgp
