// Microbenchmark: tcgen05.mma kind::i8 issue-rate peak on B200 -- the roofline denominator of the int8 (Ozaki) trailing
// updates, which MEASURED_PEAKS.json lacks (it records bf16 only).  One CTA per SM, operands resident in shared memory
// (random int8, so the data path toggles like the real kernel's), accumulators in TMEM, no global traffic in the loop.
//   mode 0/1/2 : back-to-back M=128, K=32 MMAs of width N = 256 / 128 / 64
//   mode 3     : the 12-instruction wide-N schedule of csrc/ozaki.cu (8 A planes x 8 B planes, 36 plane pairs per k-step)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_peak i8_peak.cu     Output: JSON lines.
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

constexpr int A_BYTES = 8 * 128 * 64;     // 8 planes of 128 rows x 64 B (K-major, 64-byte swizzle image)
constexpr int B_BYTES = 8 * 64 * 64;      // 8 planes of 64 rows x 64 B: also read as one 256-row or two 128-row operands

__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int mode, int iters, unsigned long long* cycles, uint32_t seed) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + B_BYTES);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 10);
    uint32_t x = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) {
        x = x * 1664525u + 1013904223u;
        uint32_t v = x ^ (x >> 13);
        // digits in [-64, 64] like the slicing kernel's
        uint32_t w = 0;
        for (int b = 0; b < 4; b++) { int d = (int)((v >> (8 * b)) & 127) - 64; w |= ((uint32_t)d & 255u) << (8 * b); }
        reinterpret_cast<uint32_t*>(smem)[i] = w;
    }
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int i = 0; i < 10; i++) mbar_init(smem_u32(bar + i), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc0 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da0 = kmajor_sw64_desc(smem_u32(sA)), db0 = kmajor_sw64_desc(smem_u32(sB));
        const uint32_t b = smem_u32(bar);
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++) {
            if (mode <= 2) {
                const int n = mode == 0 ? 256 : (mode == 1 ? 128 : 64);
                const uint32_t idesc = idesc0 | ((uint32_t)(n >> 3) << 17);
#pragma unroll
                for (int u = 0; u < 16; u++) {
                    const uint64_t da = da0 + (uint64_t)((((u & 7) * 8192) + ((u >> 3) & 1) * 32) >> 4);
                    mma_i8(tmem + (uint32_t)((u & 1) * 256), da, db0 + (uint64_t)((((u >> 1) & 1) * 32) >> 4), idesc, it > 0 || u > 1);
                }
            } else if (mode >= 4) {
                // the 7-plane schedule of csrc/ozaki.cu (28 plane pairs, s + t <= 6).  mode 4: windows of 4 B planes + the rest
                // (N = 256+192, 256+128, 256+64, 256, 192, 128, 64); mode 5: balanced windows (256+192, 192+192, 192+128, 256, 192, 128, 64)
#pragma unroll
                for (int ks = 0; ks < 2; ks++)
#pragma unroll
                    for (int s = 0; s < 7; s++) {
                        const int cnt = 7 - s;
                        const int first = (mode != 5) ? (cnt < 4 ? cnt : 4) : (cnt == 7 ? 4 : (cnt >= 5 ? 3 : cnt));
#pragma unroll
                        for (int w = 0; w < 2; w++) {
                            const int t = w == 0 ? 0 : first;
                            const int nt = w == 0 ? first : cnt - first;
                            if (nt > 0)
                                mma_i8(tmem + (uint32_t)(s + t) * 64, da0 + (uint64_t)((s * 8192 + ks * 32) >> 4),
                                       db0 + (uint64_t)((t * 4096 + ks * 32) >> 4), idesc0 | ((uint32_t)((nt * 64) >> 3) << 17),
                                       (it > 0 || ks > 0 || s > 0) ? 1u : 0u);
                        }
                    }
            } else {
#pragma unroll
                for (int ks = 0; ks < 2; ks++)
#pragma unroll
                    for (int s = 0; s < 8; s++)
#pragma unroll
                        for (int t = 0; t + s < 8; t += 4) {
                            const int nt = (8 - s - t) < 4 ? (8 - s - t) : 4;
                            mma_i8(tmem + (uint32_t)(s + t) * 64, da0 + (uint64_t)((s * 8192 + ks * 32) >> 4),
                                   db0 + (uint64_t)((t * 4096 + ks * 32) >> 4), idesc0 | ((uint32_t)((nt * 64) >> 3) << 17),
                                   (it > 0 || ks > 0 || s > 0) ? 1u : 0u);
                        }
            }
            if (mode == 6) commit(smem_u32(bar + 1));          // like the kernel: one commit per k-block (the barrier is never waited on)
            if (mode >= 7) {                                   // commit per k-block AND wait until k-block it - D has completed
                const int D = mode - 6;                        // (what a D-stage operand ring imposes on the issuing thread)
                if (it >= D) mbar_wait(smem_u32(bar + 2 + (it % D)), ((it / D) - 1) & 1);     // one phase in flight per barrier
                commit(smem_u32(bar + 2 + (it % D)));
            }
            if ((it & 7) == 7 || it == iters - 1) {          // bound the queue: wait for everything issued so far
                commit(b);
                mbar_wait(b, phase);
                phase ^= 1;
            }
        }
        cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main(int argc, char** argv) {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const int smem = A_BYTES + B_BYTES + 2048;
    cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    unsigned long long* cyc; cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    unsigned long long* h = (unsigned long long*)malloc(sizeof(unsigned long long) * sms);
    const double secs = argc > 1 ? atof(argv[1]) : 1.0;     // sustained window per mode (power cap sets the clock)
    const char* names[13] = {"N256", "N128", "N64", "ozaki_wide_n_schedule_8planes_36pairs", "ozaki_7planes_28pairs_windows_4+rest",
                            "ozaki_7planes_28pairs_balanced_windows", "ozaki_7planes_commit_per_kblock", "ozaki_7planes_commit_wait_kblock-1", "ozaki_7planes_commit_wait_kblock-2",
                            "ozaki_7planes_commit_wait_kblock-3", "ozaki_7planes_commit_wait_kblock-4", "ozaki_7planes_commit_wait_kblock-5", "ozaki_7planes_commit_wait_kblock-6"};
    const int first_mode = argc > 2 ? atoi(argv[2]) : 0;
    for (int mode = first_mode; mode < 13; mode++) {
        // MACs per loop iteration and CTA
        const double macs_it = mode == 0 ? 16.0 * 128 * 256 * 32 : mode == 1 ? 16.0 * 128 * 128 * 32 : mode == 2 ? 16.0 * 128 * 64 * 32
                               : mode == 3 ? 2.0 * 36 * 128 * 64 * 32 : 2.0 * 28 * 128 * 64 * 32;
        int iters = 2000;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        i8_peak_kernel<<<sms, 128, smem>>>(mode, iters, cyc, 1u);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        // burst: best of 5 short launches; sustained: one launch sized to ~secs
        float best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0); i8_peak_kernel<<<sms, 128, smem>>>(mode, iters, cyc, 2u + r); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double cmax = 0; for (int i = 0; i < sms; i++) if ((double)h[i] > cmax) cmax = (double)h[i];
        const double burst_tops = 2.0 * macs_it * iters * sms / (best * 1e-3) * 1e-12;
        const double mac_per_clk = macs_it * iters / cmax;
        const int iters_long = (int)(iters * (secs * 1e3 / best));
        cudaEventRecord(e0); i8_peak_kernel<<<sms, 128, smem>>>(mode, iters_long, cyc, 9u); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float msl; cudaEventElapsedTime(&msl, e0, e1);
        cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double cl = 0; for (int i = 0; i < sms; i++) if ((double)h[i] > cl) cl = (double)h[i];
        const double sust_tops = 2.0 * macs_it * iters_long * sms / (msl * 1e-3) * 1e-12;
        printf("{\"op\": \"tcgen05_i8_peak\", \"mode\": \"%s\", \"sms\": %d, \"burst_ms\": %.3f, \"burst_TOPs\": %.1f, \"mac_per_clk_per_sm\": %.1f, "
               "\"burst_sm_mhz\": %.0f, \"sustained_s\": %.2f, \"sustained_TOPs\": %.1f, \"sustained_sm_mhz\": %.0f}\n",
               names[mode], sms, best, burst_tops, mac_per_clk, cmax / (best * 1e-3) * 1e-6, msl * 1e-3, sust_tops, cl / (msl * 1e-3) * 1e-6);
        fflush(stdout);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
