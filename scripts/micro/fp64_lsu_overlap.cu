// Do shared-memory wavefronts and FP64 instructions overlap on B200?
// 256 threads x 3 CTAs per SM.  Per iteration every thread issues 32 DFMA (8 independent chains x 4,
// register + constant-bank operands: full pipe rate) and K conflict-free LDS.64 (2 wavefronts per warp
// instruction) whose results only feed integer XORs.  If the two pipes overlap the time stays at
// max(T_dfma, T_lds); if they share a dispatch path it is the sum.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lsu_overlap fp64_lsu_overlap.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ double C[4] = {1e-3, 2e-3, 3e-3, 4e-3};

template <int NDFMA, int K>
__global__ void __launch_bounds__(256, 3) k(double *out, int iters)
{
    __shared__ double sm[256 * 8];
    for (int i = threadIdx.x; i < 256 * 8; i += 256) sm[i] = i;
    __syncthreads();
    double x[8], y[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { x[q] = threadIdx.x * 1e-3 + q + 1.0; y[q] = 0.999 + 1e-6 * (threadIdx.x + q); }
    unsigned long long acc = 0;
    int off = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < NDFMA; ++t)
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = fma(x[q], y[q], C[t & 3]);
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double v = sm[(off + 256 * (j & 7)) & 2047];
            acc ^= (unsigned long long)__double_as_longlong(v);
        }
        off = (off + 32) & 255;
    }
    double s = (double)(acc & 0xff);
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NDFMA, int K>
float run(double *out)
{
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<NDFMA, K><<<148 * 3, 256>>>(out, 10);
    cudaEventRecord(e0);
    k<NDFMA, K><<<148 * 3, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM and iteration: 24 warps x (8 NDFMA DFMA at 2 cycles / 4 partitions) and 24 x K x 2 wavefronts
    const double cyc = ms * 1e-3 * 1.965e9 / iters;
    printf("DFMA/iter %2d  LDS.64/iter %2d : %7.3f ms  %7.1f cycles/iter/SM   (FP64 alone >= %5.0f, LDS wavefronts alone >= %5.0f)\n",
           8 * NDFMA, K, ms, cyc, 24.0 * 8 * NDFMA * 2 / 4, 24.0 * K * 2);
    return ms;
}

int main()
{
    double *out;
    cudaMalloc(&out, 148 * 3 * 256 * 8);
    run<4, 0>(out);
    run<0, 8>(out);
    run<0, 16>(out);
    run<4, 4>(out);
    run<4, 8>(out);
    run<4, 16>(out);
    run<4, 32>(out);
    return 0;
}
