// Microbenchmark: FP64 pipe peaks on B200 (DMMA.8x8x4 vs DFMA) -- establishes the FP64 roofline
// denominator that MEASURED_PEAKS.json lacks.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double seed) {
    double c[CH][2];
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double seed) {
    double c[CH];
    double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
#pragma unroll
    for (int i = 0; i < CH; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; cudaMalloc(&out, sizeof(double) * 256 * sms * 8);
    const int iters = 20000;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps;
        {
            float ms = time_it([&] { dmma_kernel<8><<<grid, 256>>>(out, iters, 1.0); });
            double flop = 2.0 * 8 * 8 * 4 * 8.0 * iters * (256 / 32) * grid;
            printf("DMMA.8x8x4 ch8  blocks/SM %d : %.3f ms  %.2f TFLOP/s\n", bps, ms, flop / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma_kernel<16><<<grid, 256>>>(out, iters, 1.0); });
            double flop = 2.0 * 8 * 8 * 4 * 16.0 * iters * (256 / 32) * grid;
            printf("DMMA.8x8x4 ch16 blocks/SM %d : %.3f ms  %.2f TFLOP/s\n", bps, ms, flop / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0); });
            double flop = 2.0 * 8.0 * iters * 256 * grid;
            printf("DFMA       ch8  blocks/SM %d : %.3f ms  %.2f TFLOP/s\n", bps, ms, flop / ms * 1e-9);
        }
    }
    // 128-thread blocks, 1 warp per SMSP
    {
        float ms = time_it([&] { dmma_kernel<16><<<sms, 128>>>(out, iters, 1.0); });
        double flop = 2.0 * 8 * 8 * 4 * 16.0 * iters * 4 * sms;
        printf("DMMA.8x8x4 ch16 4 warps/SM : %.3f ms  %.2f TFLOP/s\n", ms, flop / ms * 1e-9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
