// Microbenchmark: L2 -> shared-memory fill rate of cp.async.bulk on B200 -- the second roofline of the int8 (Ozaki) trailing
// update kernel (csrc/ozaki.cu).  That kernel ingests 84 KB of digit planes per 128x64x64 k-block and SM (175 int8 MAC per
// byte); at the tcgen05 kind::i8 peak (8192 MAC/clk/SM) it would need 46.8 B/clk/SM = 6.9 KB/clk chip-wide from L2.
// One CTA per SM, one producer lane issuing bulk copies into a ring of shared-memory stages, nothing consumes the data.
//   mode 0 : the kernel's own pattern -- per stage one 56 KB copy + seven 4 KB copies (84 KB), 2 stages
//   mode 1 : 8 KB copies, 16 stages of 8 KB (latency fully hidden)
//   mode 2 : as mode 0, but every CTA reads its own region (no sharing of lines between SMs)
// The source buffer (default 48 MB) is L2-resident after the warm-up pass; --dram makes it 4 GB (HBM-bound, for contrast).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_fill_peak l2_fill_peak.cu      Output: JSON lines.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; it < (1u << 28); ++it)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int SMEM_BYTES = 2 * 86016;            // 168 KB ring, like the kernel's operand stages

__global__ void __launch_bounds__(32, 1) fill_kernel(const uint8_t* src, size_t src_bytes, int mode, int iters,
                                                     unsigned long long* cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bars[16];
    const int nst = (mode == 1) ? 16 : 2;
    const uint32_t stage_bytes = (mode == 1) ? 8192u : 86016u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; i++) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    // mode 0/1: a group of 16 neighbouring CTAs walks the same 84 KB blocks (like CTAs sharing an A row panel / B column panel);
    // mode 2: private regions
    const size_t nblk = src_bytes / 86016;
    size_t blk = (mode == 2) ? ((size_t)blockIdx.x * (nblk / gridDim.x)) % nblk : ((size_t)(blockIdx.x / 16) * 97) % nblk;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int st = it % nst;
        if (it >= nst) mbar_wait(smem_u32(&bars[st]), ((it / nst) - 1) & 1);     // previous fill of this stage has landed
        const uint32_t bar = smem_u32(&bars[st]);
        const uint32_t dst = smem_u32(smem + (size_t)st * stage_bytes);
        mbar_expect_tx(bar, stage_bytes);
        if (mode == 1) {
            const uint8_t* p = src + (blk * 86016 + (size_t)(it % 10) * 8192) % (src_bytes - 8192);
            bulk_g2s(dst, p, 8192, bar);
            if (it % 10 == 9) blk = (blk + 1) % nblk;
        } else {
            const uint8_t* p = src + blk * 86016;
            bulk_g2s(dst, p, 57344, bar);
#pragma unroll
            for (int s = 0; s < 7; s++) bulk_g2s(dst + 57344 + s * 4096, p + 57344 + s * 4096, 4096, bar);
            blk = (blk + 1) % nblk;
        }
    }
    for (int i = 0; i < nst && i < iters; i++) {
        const int it = iters - 1 - i;
        mbar_wait(smem_u32(&bars[it % nst]), (it / nst) & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
}

int main(int argc, char** argv) {
    bool dram = false;
    for (int i = 1; i < argc; i++) if (!strcmp(argv[i], "--dram")) dram = true;
    int dev = 0, sms = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t src_bytes = dram ? ((size_t)4 << 30) : ((size_t)48 << 20);
    uint8_t* src;
    cudaMalloc(&src, src_bytes);
    cudaMemset(src, 1, src_bytes);
    unsigned long long* cyc;
    cudaMalloc(&cyc, sms * sizeof(unsigned long long));
    cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024);
    for (int mode = 0; mode < 3; mode++) {
        const int iters = (mode == 1) ? 40000 : 4000;
        const double bytes = (double)sms * iters * ((mode == 1) ? 8192.0 : 86016.0);
        fill_kernel<<<sms, 32, SMEM_BYTES + 1024>>>(src, src_bytes, mode, iters / 10, cyc);      // warm-up: L2 residency
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            fill_kernel<<<sms, 32, SMEM_BYTES + 1024>>>(src, src_bytes, mode, iters, cyc);
            cudaEventRecord(e1);
            cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(err)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        unsigned long long* h = (unsigned long long*)malloc(sms * sizeof(unsigned long long));
        cudaMemcpy(h, cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        double csum = 0;
        for (int i = 0; i < sms; i++) csum += (double)h[i];
        const double cyc_avg = csum / sms;
        printf("{\"op\": \"l2_fill_peak\", \"source\": \"%s\", \"mode\": %d, \"sms\": %d, \"ms\": %.3f, \"TBps\": %.3f, "
               "\"bytes_per_clk_per_sm\": %.2f, \"bytes_per_clk_chip\": %.0f, \"sm_mhz_effective\": %.0f}\n",
               dram ? "dram_4GB" : "l2_resident_48MB", mode, sms, best, bytes / best * 1e-9,
               bytes / sms / cyc_avg, bytes / cyc_avg, cyc_avg / best * 1e-3);
        free(h);
    }
    return 0;
}
