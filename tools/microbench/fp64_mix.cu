// Microbenchmark: what does a plain FP64 instruction (DMUL) cost the FP64 tensor pipe when it is issued
// in between DMMAs of other warps of the same SM sub-partition? Per SM: `dm` warps run DMMA streams (8
// independent accumulator chains each), `mu` warps run DMUL streams (16 independent chains); both count what
// they retire in a fixed number of clocks. The epilogue of the pruning kernel is 32 DMULs per warp and edge.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(512) mix(int dm_warps, int mu_warps, long long clocks, unsigned long long* out, double* sink) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long n = 0;
    const long long t0 = clock64();
    if (w < dm_warps) {
        double c[8][2];
        for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = 0.0;
        const double a = 1.0 + lane * 1e-9, b = 1e-3;
        while (clock64() - t0 < clocks) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
            n += 64;
        }
        double s = 0;
        for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
        sink[blockIdx.x * 512 + threadIdx.x] = s;
    } else if (w < dm_warps + mu_warps) {
        double x[16];
        for (int i = 0; i < 16; i++) x[i] = 1.0 + i * 1e-3 + lane * 1e-6;
        const double m = 1.0000001;
        while (clock64() - t0 < clocks) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(m));
            n += 64;
        }
        double s = 0;
        for (int i = 0; i < 16; i++) s += x[i];
        sink[blockIdx.x * 512 + threadIdx.x] = s;
    }
    if (lane == 0) out[blockIdx.x * 16 + w] = n;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    unsigned long long* out; double* sink;
    cudaMalloc(&out, sms * 16 * 8); cudaMalloc(&sink, sms * 512 * 8);
    const long long clocks = 4000000;
    const int cfg[][2] = {{8, 0}, {4, 0}, {0, 4}, {0, 8}, {8, 4}, {4, 4}, {8, 8}, {4, 8}, {12, 0}, {12, 4}};
    printf("[\n");
    for (auto& c : cfg) {
        cudaMemset(out, 0, sms * 16 * 8);
        mix<<<sms, 512>>>(c[0], c[1], clocks, out, sink);
        cudaDeviceSynchronize();
        unsigned long long h[16];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);  // SM 0
        double dmma_n = 0, dmul_n = 0;
        for (int w = 0; w < c[0]; w++) dmma_n += h[w];
        for (int w = c[0]; w < c[0] + c[1]; w++) dmul_n += h[w];
        // per sub-partition: pipe clocks per instruction
        printf("{\"dmma_warps\": %d, \"dmul_warps\": %d, \"dmma_per_kclk_per_smsp\": %.2f, \"dmul_per_kclk_per_smsp\": %.2f, \"dmma_pipe_share\": %.3f},\n",
               c[0], c[1], dmma_n / 4 / (clocks / 1000.0), dmul_n / 4 / (clocks / 1000.0), dmma_n / 4 * 16.0 / clocks);
    }
    printf("{}]\n");
    return 0;
}
