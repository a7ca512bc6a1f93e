// FP64 peak microbenchmark for the roofline denominator of the pruning kernel.
// MEASURED_PEAKS.json (driver-written) holds only HBM GB/s and bf16 TF/s; the internal-node
// contraction runs on the FP64 tensor path (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4), so its
// ceiling is measured here: register-resident DMMA and DFMA loops, no memory traffic.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void dmma_kernel(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters) {
    double c[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
    double* out; cudaMalloc(&out, sizeof(double) * sms * 1024 * 8);
    const int iters = 20000;
    printf(" \"dmma\": [\n");
    int warps_list[] = {4, 8, 16, 32};
    bool first = true;
    double best_dmma = 0;
    for (int wi = 0; wi < 4; wi++) {
        int warps = warps_list[wi];
#define RUN_DMMA(NACC) { float ms = time_ms([&] { dmma_kernel<NACC><<<sms, warps * 32>>>(out, iters); }); \
            double tf = 2.0 * 256 * NACC * (double)iters * warps * sms / (ms * 1e-3) / 1e12; \
            if (tf > best_dmma) best_dmma = tf; \
            printf("%s  {\"warps_per_sm\": %d, \"indep_acc\": %d, \"ms\": %.3f, \"tflops\": %.2f}", first ? "" : ",\n", warps, NACC, ms, tf); first = false; }
        RUN_DMMA(1) RUN_DMMA(4) RUN_DMMA(16)
    }
    printf("\n ],\n \"dfma\": [\n");
    first = true;
    double best_dfma = 0;
    for (int wi = 0; wi < 4; wi++) {
        int warps = warps_list[wi];
#define RUN_DFMA(NACC) { float ms = time_ms([&] { dfma_kernel<NACC><<<sms, warps * 32>>>(out, iters); }); \
            double tf = 2.0 * 32 * NACC * (double)iters * warps * sms / (ms * 1e-3) / 1e12; \
            if (tf > best_dfma) best_dfma = tf; \
            printf("%s  {\"warps_per_sm\": %d, \"indep_acc\": %d, \"ms\": %.3f, \"tflops\": %.2f}", first ? "" : ",\n", warps, NACC, ms, tf); first = false; }
        RUN_DFMA(4) RUN_DFMA(16)
    }
    printf("\n ],\n \"dmma_peak_tflops\": %.2f, \"dfma_peak_tflops\": %.2f}\n", best_dmma, best_dfma);
    // host<->device copy bandwidth (pinned), for the e2e leg
    size_t n = 1ull << 30;
    void *h, *d; cudaMallocHost(&h, n); cudaMalloc(&d, n);
    float ms = time_ms([&] { cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice); });
    fprintf(stderr, "h2d pinned GB/s: %.1f\n", n / (ms * 1e-3) / 1e9);
    ms = time_ms([&] { cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost); });
    fprintf(stderr, "d2h pinned GB/s: %.1f\n", n / (ms * 1e-3) / 1e9);
    return 0;
}
