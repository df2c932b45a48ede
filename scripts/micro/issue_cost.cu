// Marginal issue cost of non-FP64 instructions next to a saturated FP64 pipe (B200):
// base loop = 8 chains x 4 DFMA; a variant adds 4 instructions of one kind per chain and iteration.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_cost issue_cost.cu
#include <cstdio>
#include <cuda_runtime.h>

enum { K_NONE, K_LOP3, K_IADD, K_IMAD, K_SHF, K_ISETP_SEL, K_FSEL, K_FADD, K_FFMA, K_MOVPACK, K_LDS, K_I2F64, K_PRMT };

template <int KIND>
__global__ void __launch_bounds__(256, 3) k(double *out, int iters, double a, double b, int s1, float f1)
{
    __shared__ int sm[256];
    sm[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double x[8];
    int v[8];
    float f[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { x[q] = threadIdx.x * 1e-3 + q + 1.0; v[q] = threadIdx.x + q; f[q] = q + 0.5f; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            x[q] = fma(x[q], a, b); x[q] = fma(x[q], a, b); x[q] = fma(x[q], a, b); x[q] = fma(x[q], a, b);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (KIND == K_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[q]) : "r"(s1), "r"(i));
                if (KIND == K_IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(v[q]) : "r"(s1));
                if (KIND == K_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[q]) : "r"(s1), "r"(i));
                if (KIND == K_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(v[q]) : "r"(s1), "r"(i));
                if (KIND == K_ISETP_SEL) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.s32 %0, %2, %0, p;}" : "+r"(v[q]) : "r"(s1), "r"(i));
                if (KIND == K_FSEL) asm volatile("{.reg .pred p; setp.lt.s32 p, %1, %2; selp.f32 %0, %3, %0, p;}" : "+f"(f[q]) : "r"(v[q]), "r"(s1), "f"(f1));
                if (KIND == K_FADD) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[q]) : "f"(f1));
                if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[q]) : "f"(f1));
                if (KIND == K_MOVPACK) { int hi = __double2hiint(x[q]) ^ (t + 1); x[q] = __hiloint2double(hi, __double2loint(x[q])); }
                if (KIND == K_LDS) v[q] = sm[(v[q] + t) & 255];
                if (KIND == K_I2F64) x[q] += (double)v[q];
                if (KIND == K_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(v[q]) : "r"(s1));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += x[q] + v[q] + f[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
double run(const char *name, double *out, double base)
{
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<KIND><<<148 * 3, 256>>>(out, 10, 0.999, 1e-3, 3, 1.5f);
    cudaEventRecord(e0);
    k<KIND><<<148 * 3, 256>>>(out, iters, 0.999, 1e-3, 3, 1.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // cycles per (chain, iteration) group per SM partition: 6 warps per partition, 8 chains
    const double cyc = ms * 1e-3 * 1.965e9 / (6.0 * 8 * iters);
    printf("%-32s %7.2f cycles per group (4 DFMA + 4 x)   marginal %.2f cycles per extra instruction\n", name, cyc,
           base > 0 ? (cyc - base) / 4.0 : 0.0);
    return cyc;
}

int main()
{
    double *out;
    cudaMalloc(&out, 148 * 3 * 256 * 8);
    const double base = run<K_NONE>("4 DFMA only", out, 0);
    run<K_LOP3>("+ LOP3", out, base);
    run<K_IADD>("+ IADD", out, base);
    run<K_IMAD>("+ IMAD", out, base);
    run<K_SHF>("+ SHF", out, base);
    run<K_ISETP_SEL>("+ ISETP+SEL (2 instr)", out, base);
    run<K_FSEL>("+ ISETP+FSEL (2 instr)", out, base);
    run<K_FADD>("+ FADD (fp32)", out, base);
    run<K_FFMA>("+ FFMA (fp32)", out, base);
    run<K_MOVPACK>("+ hi-word xor of a double", out, base);
    run<K_LDS>("+ LDS (dependent)", out, base);
    run<K_I2F64>("+ I2F.F64 + DADD", out, base);
    run<K_PRMT>("+ PRMT", out, base);
    return 0;
}
