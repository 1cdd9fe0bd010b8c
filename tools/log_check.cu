// Bit-exactness check of rg_log_pos (rg_common.cuh) against CUDA's log() -- run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I relagn_b200/csrc -o /tmp/lc tools/log_check.cu && /tmp/lc
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <vector>
#include "rg_common.cuh"
__global__ void k_ref(const double *x, double *y, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = log(x[i]); }
__global__ void k_my(const double *x, double *y, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = rg_log_pos(x[i]); }
int main()
{
    const int N = 1 << 22;
    std::vector<double> hx(N), a(N), b(N);
    for (int i = 0; i < N; i++) { double u = (double)i / N; hx[i] = (i % 3 == 0) ? 0.01 + 3.0 * u : (i % 3 == 1 ? pow(10.0, -300 + 600 * u) : 0.999 + 0.002 * u); }
    hx[0] = 1.0; hx[1] = 0.0; hx[2] = -1.0; hx[3] = 1e-310; hx[4] = INFINITY; hx[5] = NAN; hx[6] = 1.4142135623730951; hx[7] = 0.70710678118654757;
    double *dx, *dy; cudaMalloc(&dx, N * 8); cudaMalloc(&dy, N * 8);
    cudaMemcpy(dx, hx.data(), N * 8, cudaMemcpyHostToDevice);
    k_ref<<<N / 256, 256>>>(dx, dy, N); cudaMemcpy(a.data(), dy, N * 8, cudaMemcpyDeviceToHost);
    k_my<<<N / 256, 256>>>(dx, dy, N); cudaMemcpy(b.data(), dy, N * 8, cudaMemcpyDeviceToHost);
    long ndiff = 0; double maxabs = 0;
    for (int i = 0; i < N; i++) {
        if (std::isnan(a[i]) && std::isnan(b[i])) continue;
        if (a[i] != b[i]) { ndiff++; double r = fabs(a[i] - b[i]); if (r > maxabs) maxabs = r; if (ndiff < 5) printf("diff x=%.17g ref=%.17g my=%.17g\n", hx[i], a[i], b[i]); }
    }
    printf("N=%d bit-different=%ld maxabs=%.3g err=%s\n", N, ndiff, maxabs, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 8; i++) printf("x=%.17g ref=%.17g my=%.17g\n", hx[i], a[i], b[i]);
    return 0;
}
