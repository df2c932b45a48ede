// FP64 pipe throughput on B200 as a function of the operand kinds of DFMA (register-file bandwidth):
// 8 independent chains per thread, 256 threads x 3 CTAs per SM (6 warps per SM partition), no other work.
//   K_CONST   x = fma(x, a, b)      a, b kernel parameters (uniform / constant-bank operands)
//   K_REG1    x = fma(x, y, b)      one more per-thread register operand
//   K_REG2    x = fma(x, y, z)      three distinct per-thread register operands
//   K_MUL     x = x * y             DMUL, two register operands
//   K_ADD     x = x + y             DADD
//   K_MIX     poly-like: p = fma(p, w, c_i) with w per-thread register and c_i from constant memory
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { K_CONST, K_REG1, K_REG2, K_MUL, K_ADD, K_MIX };
__constant__ double C[16];

template <int KIND>
__global__ void __launch_bounds__(256, 3) k(double *out, int iters, double a, double b)
{
    double x[8], y[8], z[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        x[q] = threadIdx.x * 1e-3 + q + 1.0;
        y[q] = 0.999 + 1e-6 * (threadIdx.x + q);
        z[q] = 1e-3 * (q + 1) + 1e-7 * threadIdx.x;
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (KIND == K_CONST) x[q] = fma(x[q], a, b);
                if (KIND == K_REG1) x[q] = fma(x[q], y[q], b);
                if (KIND == K_REG2) x[q] = fma(x[q], y[q], z[q]);
                if (KIND == K_MUL) x[q] = x[q] * y[q];
                if (KIND == K_ADD) x[q] = x[q] + z[q];
                if (KIND == K_MIX) x[q] = fma(x[q], y[q], C[t]);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char *name, double *out)
{
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<KIND><<<148 * 3, 256>>>(out, 10, 0.999, 1e-3);
    cudaEventRecord(e0);
    k<KIND><<<148 * 3, 256>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = 148.0 * 3 * 256 * 8 * 4 * iters;       // FP64 instructions (thread level)
    const double per_sm_clk = n / (ms * 1e-3) / 148.0 / 1.965e9;   // lanes per clock per SM at 1965 MHz
    printf("%-44s %8.3f ms  %6.1f FP64 lanes/clk/SM (64 = pipe peak)  %.2f TFLOP/s-equivalent (x2)\n", name, ms,
           per_sm_clk, 2.0 * n / (ms * 1e-3) / 1e12);
}

int main()
{
    double h[16];
    for (int i = 0; i < 16; ++i) h[i] = 1e-3 * (i + 1);
    cudaMemcpyToSymbol(C, h, sizeof(h));
    double *out;
    cudaMalloc(&out, 148 * 3 * 256 * 8);
    run<K_CONST>("DFMA x = fma(x, param, param)", out);
    run<K_REG1>("DFMA x = fma(x, reg, param)", out);
    run<K_REG2>("DFMA x = fma(x, reg, reg)", out);
    run<K_MUL>("DMUL x = x * reg", out);
    run<K_ADD>("DADD x = x + reg", out);
    run<K_MIX>("DFMA x = fma(x, reg, const-bank)", out);
    return 0;
}
