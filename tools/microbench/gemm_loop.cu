// Microbenchmark of the pruning kernel's inner contraction in isolation: one warp multiplies its
// register-resident partials (8*T columns x 64 states) by a 64x64 P image in shared memory, over and
// over (acc feeds back as the next input, like walking up a tree). Question it answers: how close to
// the DMMA pipe rate does ONE warp per SM sub-partition get, versus two or four, for T = 1, 2 and for
// B fragments from LDS.64 vs registers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gemm_loop gemm_loop.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int T, int MODE>  // MODE 0: B from LDS.64 per (j,s); 1: B from LDS but j-outer order; 2: B constant register
__global__ void __launch_bounds__(512) gemm_loop(const double* __restrict__ Pimg, double* out, int iters) {
    extern __shared__ double Ps[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) Ps[i] = Pimg[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double cur[T][8][2];
#pragma unroll
    for (int t = 0; t < T; t++)
#pragma unroll
        for (int j = 0; j < 8; j++) cur[t][j][0] = cur[t][j][1] = 1.0 / 64;
    const double* Pb = Ps + lane;
    for (int it = 0; it < iters; it++) {
        double acc[T][8][2];
#pragma unroll
        for (int t = 0; t < T; t++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[t][j][0] = acc[t][j][1] = 0.0;
        if (MODE == 0) {
#pragma unroll
            for (int s = 0; s < 16; s++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const double bf = Pb[(j * 16 + s) * 32];
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][j][0], acc[t][j][1], cur[t][s >> 1][s & 1], bf);
                }
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 8; j++)
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const double bf = Pb[(j * 16 + s) * 32];
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][j][0], acc[t][j][1], cur[t][s >> 1][s & 1], bf);
                }
        } else if (MODE == 3) {  // s outer, T middle, j inner: 8 consecutive DMMAs share the A operand
#pragma unroll
            for (int s = 0; s < 16; s++) {
                double bf[8];
#pragma unroll
                for (int j = 0; j < 8; j++) bf[j] = Pb[(j * 16 + s) * 32];
#pragma unroll
                for (int t = 0; t < T; t++)
#pragma unroll
                    for (int j = 0; j < 8; j++) dmma(acc[t][j][0], acc[t][j][1], cur[t][s >> 1][s & 1], bf[j]);
            }
        } else if (MODE == 4) {  // like 0 but B fragments of k-steps 2s2, 2s2+1 fetched with one LDS.128
            const double2* Pb2 = reinterpret_cast<const double2*>(Ps) + lane;
#pragma unroll
            for (int s2 = 0; s2 < 8; s2++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const double2 bf = Pb2[(j * 8 + s2) * 32];
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][j][0], acc[t][j][1], cur[t][s2][0], bf.x);
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][j][0], acc[t][j][1], cur[t][s2][1], bf.y);
                }
        } else if (MODE == 5) {  // B fragments of n-tiles 2jp, 2jp+1 (same k-step) in one LDS.128: 4 independent accumulator chains per load
            const double2* Pb2 = reinterpret_cast<const double2*>(Ps) + lane;
#pragma unroll
            for (int s = 0; s < 16; s++)
#pragma unroll
                for (int jp = 0; jp < 4; jp++) {
                    const double2 bf = Pb2[(s * 4 + jp) * 32];
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][2 * jp][0], acc[t][2 * jp][1], cur[t][s >> 1][s & 1], bf.x);
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][2 * jp + 1][0], acc[t][2 * jp + 1][1], cur[t][s >> 1][s & 1], bf.y);
                }
        } else if (MODE == 6) {  // like 5 with the loop over n-tile pairs outermost: 4 chains, 16 k-steps each, then the next pair
            const double2* Pb2 = reinterpret_cast<const double2*>(Ps) + lane;
#pragma unroll
            for (int jp = 0; jp < 4; jp++)
#pragma unroll
                for (int s = 0; s < 16; s++) {
                    const double2 bf = Pb2[(jp * 16 + s) * 32];
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][2 * jp][0], acc[t][2 * jp][1], cur[t][s >> 1][s & 1], bf.x);
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][2 * jp + 1][0], acc[t][2 * jp + 1][1], cur[t][s >> 1][s & 1], bf.y);
                }
        } else {
            const double bf = Pb[0];
#pragma unroll
            for (int s = 0; s < 16; s++)
#pragma unroll
                for (int j = 0; j < 8; j++)
#pragma unroll
                    for (int t = 0; t < T; t++) dmma(acc[t][j][0], acc[t][j][1], cur[t][s >> 1][s & 1], bf);
        }
#pragma unroll
        for (int t = 0; t < T; t++)
#pragma unroll
            for (int j = 0; j < 8; j++) { cur[t][j][0] = acc[t][j][0]; cur[t][j][1] = acc[t][j][1]; }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < T; t++)
#pragma unroll
        for (int j = 0; j < 8; j++) s += cur[t][j][0] + cur[t][j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int T, int MODE>
void run(const double* P, double* out, int sms, int warps) {
    const int iters = 2000;
    cudaFuncSetAttribute(gemm_loop<T, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    gemm_loop<T, MODE><<<sms, warps * 32, 32768>>>(P, out, 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        gemm_loop<T, MODE><<<sms, warps * 32, 32768>>>(P, out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 2.0 * 256 * 128 * T * (double)iters * warps * sms;
    printf("{\"T\": %d, \"mode\": %d, \"warps_per_sm\": %d, \"tflops\": %.2f},\n", T, MODE, warps, flops / (best * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *P, *out;
    cudaMalloc(&P, 4096 * 8); cudaMalloc(&out, sms * 512 * 8);
    double h[4096]; for (int i = 0; i < 4096; i++) h[i] = 1.0 / 64;
    cudaMemcpy(P, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("[\n");
    int ws[] = {4, 8};
    for (int wi = 0; wi < 2; wi++) {
        const int w = ws[wi];
        run<2, 0>(P, out, sms, w); run<2, 1>(P, out, sms, w); run<2, 3>(P, out, sms, w); run<2, 4>(P, out, sms, w);
        run<2, 5>(P, out, sms, w); run<2, 6>(P, out, sms, w); run<1, 5>(P, out, sms, w); run<1, 6>(P, out, sms, w);
        run<1, 0>(P, out, sms, w); run<1, 3>(P, out, sms, w); run<1, 4>(P, out, sms, w);
        if (w <= 8) { run<3, 0>(P, out, sms, w); run<3, 3>(P, out, sms, w); }
    }
    printf("{}]\n");
    return 0;
}
