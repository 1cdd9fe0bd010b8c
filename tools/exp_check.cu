#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>
#include <vector>
__constant__ double RG_EXPC[16] = {
    1.4426950408889634074,            // log2(e)
    -6.93147180369123816490e-01,      // -ln2 hi
    -1.90821492927058770002e-10,      // -ln2 lo   (placeholders, replaced below by bit patterns)
};
__device__ __forceinline__ double bits(unsigned hi, unsigned lo) { return __hiloint2double((int)hi, (int)lo); }
__constant__ unsigned long long RG_EXPB[14] = {
    0x3ff71547652b82feULL, // log2 e
    0x3fe62e42fefa39efULL, // ln2 hi
    0x3c7abc9e3b39803fULL, // ln2 lo
    0x3e5ade1569ce2bdfULL, // c11
    0x3e928af3fca213eaULL, // c10
    0x3ec71dee62401315ULL, // c9
    0x3efa01997c89eb71ULL, // c8
    0x3f2a01a014761f65ULL, // c7
    0x3f56c16c1852b7afULL, // c6
    0x3f81111111122322ULL, // c5
    0x3fa55555555502a1ULL, // c4
    0x3fc5555555555511ULL, // c3
    0x3fe000000000000bULL, // c2
    0ULL};
__device__ __forceinline__ double K(int i) { return __longlong_as_double((long long)RG_EXPB[i]); }
__device__ __forceinline__ double exp_pos(double x)
{
    const double MAGIC = 6755399441055744.0;
    const double xc = fmin(x, 709.9);
    const double t = fma(xc, K(0), MAGIC);
    const int n = __double2loint(t);
    const double r = t - MAGIC;
    double f = fma(r, -K(1), xc);
    f = fma(r, -K(2), f);
    double p = fma(f, K(3), K(4));
    p = fma(p, f, K(5)); p = fma(p, f, K(6)); p = fma(p, f, K(7)); p = fma(p, f, K(8)); p = fma(p, f, K(9));
    p = fma(p, f, K(10)); p = fma(p, f, K(11)); p = fma(p, f, K(12));
    p = fma(p, f, 1.0); p = fma(p, f, 1.0);
    const double s = __hiloint2double(__double2hiint(p) + ((n - 1) << 20), __double2loint(p));
    return s * 2.0;
}
__global__ void k_ref(const double *x, double *y, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = exp(x[i]); }
__global__ void k_my(const double *x, double *y, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) y[i] = exp_pos(x[i]); }
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
