// fp64_peak.cu -- measures the B200's FP64 pipe: DFMA, non-fused DMUL+DADD, IEEE division and
// square-root throughput (MEASURED_PEAKS.json has no FP64 figure; the GJK/EPA path is FP64 bound).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_dfma(double *out, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_muladd(double *out, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < ITERS; ++i) {
        x0 = __dadd_rn(__dmul_rn(x0, a), b); x1 = __dadd_rn(__dmul_rn(x1, a), b);
        x2 = __dadd_rn(__dmul_rn(x2, a), b); x3 = __dadd_rn(__dmul_rn(x3, a), b);
        x4 = __dadd_rn(__dmul_rn(x4, a), b); x5 = __dadd_rn(__dmul_rn(x5, a), b);
        x6 = __dadd_rn(__dmul_rn(x6, a), b); x7 = __dadd_rn(__dmul_rn(x7, a), b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_div(double *out, double a, double b)
{
    double x0 = threadIdx.x + 1.0, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    for (int i = 0; i < ITERS / 4; ++i) {
        x0 = a / x0 + b; x1 = a / x1 + b; x2 = a / x2 + b; x3 = a / x3 + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

__global__ void k_sqrt(double *out, double a, double b)
{
    double x0 = threadIdx.x + 1.0, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    for (int i = 0; i < ITERS / 4; ++i) {
        x0 = sqrt(x0) + b; x1 = sqrt(x1) + b; x2 = sqrt(x2) + b; x3 = sqrt(x3) + b;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

template <typename K>
static double time_kernel(K kern, double *out, int blocks, int threads)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<blocks, threads>>>(out, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) kern<<<blocks, threads>>>(out, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / 10.0 * 1e-3;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, threads = 256, blocks = sms * 8;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    double n = (double)blocks * threads;
    double t;
    t = time_kernel(k_dfma, out, blocks, threads);
    double dfma = n * ITERS * 8 / t;
    t = time_kernel(k_muladd, out, blocks, threads);
    double ma = n * ITERS * 8 * 2 / t;
    t = time_kernel(k_div, out, blocks, threads);
    double dv = n * (ITERS / 4) * 4 / t;
    t = time_kernel(k_sqrt, out, blocks, threads);
    double sq = n * (ITERS / 4) * 4 / t;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops_fma\": %.3f, "
           "\"dmul_dadd_instr_per_s\": %.4e, \"ddiv_per_s\": %.4e, \"dsqrt_per_s\": %.4e, "
           "\"dfma_per_clk_per_sm_at_1965mhz\": %.2f}\n",
           p.name, sms, dfma, dfma * 2 / 1e12, ma, dv, sq, dfma / sms / 1.965e9);
    return 0;
}
