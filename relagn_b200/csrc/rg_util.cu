// rg_util.cu -- FP64 FMA throughput micro-benchmark: the measured denominator
// of the FP64-pipe roofline the spectrum kernels are reported against
// (MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only).
#include "rg_kernels.h"

__global__ void __launch_bounds__(256) fp64_fma_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, b = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, b); a1 = fma(a1, m, b); a2 = fma(a2, m, b); a3 = fma(a3, m, b);
        a4 = fma(a4, m, b); a5 = fma(a5, m, b); a6 = fma(a6, m, b); a7 = fma(a7, m, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

int rg_fp64_peak(double *tflops, cudaStream_t st)
{
    int dev = 0;
    RG_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    RG_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
#ifdef RG_HOSTSIM
    const int iters = 8;
#else
    const int iters = 1 << 14;
#endif
    double *d = nullptr;
    RG_CUDA_OK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    RG_CUDA_OK(cudaEventCreate(&e0));
    RG_CUDA_OK(cudaEventCreate(&e1));
    RG_LAUNCH(fp64_fma_kernel, dim3(blocks), dim3(threads), 0, st, d, iters, 1.0); // warm-up
    RG_CUDA_OK(cudaEventRecord(e0, st));
    RG_LAUNCH(fp64_fma_kernel, dim3(blocks), dim3(threads), 0, st, d, iters, 2.0);
    RG_CUDA_OK(cudaEventRecord(e1, st));
    RG_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    RG_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    RG_CUDA_OK(cudaGetLastError());
    double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
    *tflops = ms > 0 ? flops / (ms * 1e-3) / 1e12 : 0.0;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return RG_OK;
}
