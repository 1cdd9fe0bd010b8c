// Bit-exactness check of rg_exp_pos (rg_common.cuh) against CUDA's exp() -- run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I relagn_b200/csrc -o /tmp/ec tools/exp_check.cu && /tmp/ec
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <vector>
#include "rg_common.cuh"
__global__ void k_ref(const double *x, double *y, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = exp(x[i]); }
__global__ void k_my(const double *x, double *y, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = rg_exp_pos(x[i]); }
int main()
{
    const int N = 1 << 22;
    std::vector<double> hx(N), a(N), b(N);
    for (int i = 0; i < N; i++) { double u = (double)i / N; hx[i] = (i % 3 == 0) ? u * 720.0 : (i % 3 == 1 ? pow(10.0, -12 + 15 * u) : 700.0 + 12.0 * u); }
    hx[0] = 0; hx[1] = 709.782712893384; hx[2] = 709.7827128933841; hx[3] = 1e5; hx[4] = 1e300; hx[5] = 709.9; hx[6] = 710.0;
    double *dx, *dy; cudaMalloc(&dx, N * 8); cudaMalloc(&dy, N * 8);
    cudaMemcpy(dx, hx.data(), N * 8, cudaMemcpyHostToDevice);
    k_ref<<<N / 256, 256>>>(dx, dy, N); cudaMemcpy(a.data(), dy, N * 8, cudaMemcpyDeviceToHost);
    k_my<<<N / 256, 256>>>(dx, dy, N); cudaMemcpy(b.data(), dy, N * 8, cudaMemcpyDeviceToHost);
    long ndiff = 0; double maxrel = 0; long infmis = 0;
    for (int i = 0; i < N; i++) {
        if (std::isinf(a[i]) != std::isinf(b[i])) { infmis++; if (infmis < 5) printf("inf mismatch x=%.17g ref=%g my=%g\n", hx[i], a[i], b[i]); continue; }
        if (std::isinf(a[i])) continue;
        if (a[i] != b[i]) { ndiff++; double r = fabs(a[i] - b[i]) / a[i]; if (r > maxrel) maxrel = r; }
    }
    printf("N=%d bit-different=%ld maxrel=%.3g inf-mismatch=%ld err=%s\n", N, ndiff, maxrel, infmis, cudaGetErrorString(cudaGetLastError()));
    for (int i = 0; i < 7; i++) printf("x=%.17g ref=%.17g my=%.17g\n", hx[i], a[i], b[i]);
    return 0;
}
