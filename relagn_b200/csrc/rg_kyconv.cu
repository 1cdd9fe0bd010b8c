// rg_kyconv.cu -- the stand-alone convolution seam (B1):
//   xspec.callModelFunction('kyconv', energies, params[12], flux)
//   (src/python_version/relagn.py:979-1001; tests/blurr_test.py:51-54,111-119)
// One annulus of arbitrary width and emissivity (alpha, beta, r_break): one CTA of 8 warps builds the shift
// histogram -- ring m goes to warp m mod 8, every warp deposits into its own shared-memory copy with the
// warp-shuffle segmented reduction of rg_transfer.cuh (no atomics), the copies are folded in warp order -- then a
// second kernel applies it to the photon spectrum.  The same device functions build every annulus inside the fused
// spectrum kernel.
#include "rg_kernels.h"
#include "rg_transfer.cuh"

#define KT 256
#define KW (KT / 32)

__global__ void __launch_bounds__(KT)
ky_hist_kernel(TabDesc tab, double a, double theta_o, double r1, double r2, double alpha, double beta, double rb,
               double zs, double dlnE, int nsub, double *__restrict__ H_out, int K_LO, int nH)
{
    RG_DYN_SMEM(smem_raw);
    double *H = reinterpret_cast<double *>(smem_raw); // [KW][nH]
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < KW * nH; i += KT) H[i] = 0.0;
    TabSel ts;
    rg_tab_select(ts, tab, a, theta_o);
    __syncthreads();
    int kmn, kmx;
    warp_build_rings(H + (size_t)warp * nH, K_LO, nH, ts, r1, r2, dlnE, nsub, warp, KW, alpha, beta, rb, zs, 1.0, kmn, kmx);
    __syncthreads();
    for (int i = tid; i < nH; i += KT) {
        double v = 0.0;
        for (int w = 0; w < KW; w++) v += H[(size_t)w * nH + i];
        H_out[i] = v;
    }
}

__global__ void __launch_bounds__(KT)
ky_conv_kernel(const double *__restrict__ H, int kmin, int nk, const double *__restrict__ in, double *__restrict__ out,
               int nE)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nE) return;
    double s = 0;
    for (int i = 0; i < nk; i++) {
        int src = j - (kmin + i);
        if (src >= 0 && src < nE) s += H[i] * in[src];
    }
    out[j] = s;
}

int rg_launch_ky_hist(const TabDesc &tab, double a, double theta_o, double r1, double r2, double alpha, double beta,
                      double rb, double zshift, double dlnE, int nsub, double *d_H, int K_LO, int nH,
                      cudaStream_t st)
{
    if (KT % tab.NPHI != 0 || tab.NPHI < 32) return rg_set_error(RG_E_ARG, "NPHI must divide %d and be >= 32", KT);
    size_t smem = (size_t)KW * nH * sizeof(double);
    if (smem > 220 * 1024) return rg_set_error(RG_E_GRID, "energy grid too fine for the kyconv histogram");
    double zs = 0.0;
    if (zshift != 0) zs = -log(1 + zshift) / dlnE;
    if (nsub <= 0) nsub = rg_auto_nsub(r1, r2, log(r2 / r1), dlnE);
    RG_CUDA_OK(cudaFuncSetAttribute(ky_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RG_LAUNCH(ky_hist_kernel, dim3(1), dim3(KT), smem, st, tab, a, theta_o, r1, r2, alpha, beta, rb, zs, dlnE, nsub,
              d_H, K_LO, nH);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

int rg_launch_ky_conv(const double *d_H, int kmin, int nk, const double *d_in, double *d_out, int nE,
                      cudaStream_t st)
{
    RG_LAUNCH(ky_conv_kernel, dim3((nE + KT - 1) / KT), dim3(KT), 0, st, d_H, kmin, nk, d_in, d_out, nE);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}
