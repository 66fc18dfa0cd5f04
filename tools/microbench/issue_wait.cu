// Microbenchmark: what does a per-k-block synchronisation cost the thread that issues tcgen05.mma?  (csrc/ozaki.cu's MMA warp
// waits on the operand ring's `full` barrier once per 64-wide k-block = 20 UTCIMMAs = ~1870 cycles of tensor work.)
// The 7-plane / 28-pair schedule of the kernel with static operands in shared memory, one CTA per SM, warp 0 issues
// (all 32 lanes run the loop, lane 0 issues -- like the kernel), and per k-block:
//   mode 0 : nothing (a commit + wait every 8 k-blocks only bounds the queue)
//   mode 1 : mbarrier.try_wait by the issuing warp on the commit of k-block it-2 (what a 2-stage ring imposes)
//   mode 2 : bar.sync (named barrier, 64 threads) with a partner warp that is always ready
//   mode 3 : the partner warp waits on the commit of k-block it-2 (mbarrier), then releases the issuer through the named barrier
//   modes 4-7 : as 0-3 with the k-block's instructions ordered widest-last (order 2 of the kernel)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_wait issue_wait.cu      Output: JSON lines.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 26); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ uint64_t kmajor_sw64_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

constexpr int A_BYTES = 7 * 128 * 64;
constexpr int B_BYTES = 7 * 64 * 64;

__global__ void __launch_bounds__(64, 1) issue_wait_kernel(int mode, int iters, unsigned long long* cycles, uint32_t seed) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + B_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
    uint32_t x = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
        x = x * 1664525u + 1013904223u;
        reinterpret_cast<uint32_t*>(smem)[i] = x ^ (x >> 13);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; i++) mbar_init(smem_u32(bar + i), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    const int sync_mode = mode & 3;
    const bool wide_last = mode >= 4;
    // barriers: 0 = queue bound, 2/3 = per-k-block commits (parity ring of 2)
    if (warp == 0) {
        const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da0 = kmajor_sw64_desc(smem_u32(sA)), db0 = kmajor_sw64_desc(smem_u32(sB));
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            if (sync_mode == 1 && it >= 2) mbar_wait(smem_u32(bar + 2 + (it & 1)), ((it >> 1) - 1) & 1);
            if (sync_mode >= 2) named_sync(1, 64);
            if (lane == 0) {
                if (!wide_last) {
#pragma unroll
                    for (int ks = 0; ks < 2; ks++)
#pragma unroll
                        for (int s = 0; s < 7; s++) {
                            const int cnt = 7 - s, first = cnt < 4 ? cnt : 4;
#pragma unroll
                            for (int w = 0; w < 2; w++) {
                                const int t = w == 0 ? 0 : first, nt = w == 0 ? first : cnt - first;
                                if (nt > 0)
                                    mma_i8(tmem + (uint32_t)(s + t) * 64, da0 + (uint64_t)((s * 8192 + ks * 32) >> 4),
                                           db0 + (uint64_t)((t * 4096 + ks * 32) >> 4), idesc0 | ((uint32_t)((nt * 64) >> 3) << 17),
                                           (it > 0 || ks > 0 || s > 0) ? 1u : 0u);
                            }
                        }
                } else {
                    constexpr int NI = 20;
                    constexpr int OK[NI] = {0, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 1};
                    constexpr int OS[NI] = {0, 0, 0, 6, 6, 5, 5, 4, 4, 2, 2, 1, 1, 0, 1, 1, 2, 2, 3, 3};
                    constexpr int OT[NI] = {0, 4, 4, 0, 0, 0, 0, 0, 0, 4, 4, 4, 4, 0, 0, 0, 0, 0, 0, 0};
                    constexpr int ON[NI] = {4, 3, 3, 1, 1, 2, 2, 3, 3, 1, 1, 2, 2, 4, 4, 4, 4, 4, 4, 4};
#pragma unroll
                    for (int i = 0; i < NI; i++)
                        mma_i8(tmem + (uint32_t)(OS[i] + OT[i]) * 64, da0 + (uint64_t)((OS[i] * 8192 + OK[i] * 32) >> 4),
                               db0 + (uint64_t)((OT[i] * 4096 + OK[i] * 32) >> 4), idesc0 | ((uint32_t)((ON[i] * 64) >> 3) << 17),
                               (it > 0 || i > 1) ? 1u : 0u);
                }
                if (sync_mode == 1 || sync_mode == 3) commit(smem_u32(bar + 2 + (it & 1)));
                if ((it & 7) == 7 || it == iters - 1) commit(smem_u32(bar));
            }
            __syncwarp();
            if ((it & 7) == 7 || it == iters - 1) {          // bound the queue: wait for everything issued so far
                mbar_wait(smem_u32(bar), phase);
                phase ^= 1;
            }
        }
        if (lane == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    } else {
        // partner warp
        if (sync_mode >= 2)
            for (int it = 0; it < iters; it++) {
                if (sync_mode == 3 && it >= 2) mbar_wait(smem_u32(bar + 2 + (it & 1)), ((it >> 1) - 1) & 1);
                named_sync(1, 64);
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main(int argc, char** argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const int smem = A_BYTES + B_BYTES + 2048;
    cudaFuncSetAttribute(issue_wait_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    unsigned long long* cyc; cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * sms);
    const char* names[8] = {"no_sync", "mbarrier_wait_by_issuer", "named_barrier_partner_ready", "partner_waits_mbarrier_then_named_barrier",
                            "widest_last/no_sync", "widest_last/mbarrier_wait_by_issuer", "widest_last/named_barrier_partner_ready",
                            "widest_last/partner_waits_mbarrier_then_named_barrier"};
    for (int mode = 0; mode < 8; mode++) {
        const double macs_it = 2.0 * 28 * 128 * 64 * 32;
        const int iters = 4000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        issue_wait_kernel<<<sms, 64, smem>>>(mode, 200, cyc, 1u);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\", \"mode\": %d}\n", cudaGetErrorString(cudaGetLastError()), mode); return 1; }
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0); issue_wait_kernel<<<sms, 64, smem>>>(mode, iters, cyc, 2u + r); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\", \"mode\": %d}\n", cudaGetErrorString(cudaGetLastError()), mode); return 1; }
        cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double cmax = 0; for (int i = 0; i < sms; i++) if ((double)h[i] > cmax) cmax = (double)h[i];
        printf("{\"op\": \"issue_wait\", \"mode\": \"%s\", \"cycles_per_kblock\": %.1f, \"mac_per_clk_per_sm\": %.1f, \"ms\": %.3f}\n",
               names[mode], cmax / iters, macs_it * iters / cmax, best);
        fflush(stdout);
    }
    return 0;
}
