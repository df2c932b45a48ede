// DFMA throughput on one SM partition as a function of resident warps and independent chains per
// warp: tells how much FP64 parallelism the build kernels need to keep the pipe busy.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_latency dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void chains(double *out, int iters, double a, double b)
{
    double x[K];
#pragma unroll
    for (int k = 0; k < K; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int K>
void run(int warps_per_sm, double *out)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    // one CTA per SM, warps_per_sm warps (4 SMSPs share them)
    chains<K><<<148, warps_per_sm * 32>>>(out, 100, 0.999, 1e-3);
    cudaEventRecord(e0);
    chains<K><<<148, warps_per_sm * 32>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma = 148.0 * warps_per_sm * 32.0 * K * iters;
    printf("warps/SM %2d (%.1f/SMSP) chains %d : %.2f TFLOP/s  (%.2f cycles per warp-DFMA per SMSP at 1.965 GHz)\n",
           warps_per_sm, warps_per_sm / 4.0, K, 2 * dfma / ms / 1e9,
           ms * 1e-3 * 1.965e9 / ((double)warps_per_sm / 4.0 * K * iters));
}

int main()
{
    double *out;
    cudaMalloc(&out, 148 * 1024 * 8);
    for (int w : {4, 8, 16, 24, 32}) {
        run<1>(w, out);
        run<2>(w, out);
        run<4>(w, out);
        run<8>(w, out);
    }
    return 0;
}
