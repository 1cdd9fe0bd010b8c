// rg_kyconv.cu -- the stand-alone convolution seam (B1):
//   xspec.callModelFunction('kyconv', energies, params[12], flux)
//   (src/python_version/relagn.py:979-1001; tests/blurr_test.py:51-54,111-119)
// One annulus of arbitrary width and emissivity (alpha, beta, r_break): the
// shift histogram is built by one CTA (rings in passes of 256/NPHI, shared-
// memory atomics), then a second kernel applies it to the photon spectrum.
// The same device functions (rg_transfer.cuh) build every annulus inside the
// fused spectrum kernel.
#include "rg_kernels.h"
#include "rg_transfer.cuh"

#define KT 256

__global__ void __launch_bounds__(KT)
ky_hist_kernel(TabDesc tab, double a, double theta_o, double r1, double r2, double alpha, double beta, double rb,
               double zs, double dlnE, int nsub, double *__restrict__ H_out, int K_LO, int nH)
{
    RG_DYN_SMEM(smem_raw);
    double *H = reinterpret_cast<double *>(smem_raw);
    __shared__ double sv[KT], wv[KT], rA[8], rFr[8];
    __shared__ int rS0[8], rS1[8];
    const int tid = threadIdx.x;
    for (int i = tid; i < nH; i += KT) H[i] = 0.0;
    TabSel ts;
    rg_tab_select(ts, tab, a, theta_o);
    const int NP = ts.NPHI;
    const int rpp = KT / NP;
    double lnr = log(r2 / r1);
    double dphi = 2 * RGK_PI / NP;
    __syncthreads();
    for (int m0 = 0; m0 < nsub; m0 += rpp) {
        int nr = nsub - m0 < rpp ? nsub - m0 : rpp;
        if (tid < nr) {
            RingGeo g = rg_ring_geo(ts, r1, r2, lnr, m0 + tid, nsub, alpha, beta, rb);
            rA[tid] = g.A; rFr[tid] = g.fr; rS0[tid] = g.s0; rS1[tid] = g.s1;
        }
        __syncthreads();
        int slot = tid / NP, p = tid - slot * NP;
        if (slot < nr) {
            double gg, jj;
            rg_ring_sample(ts, rS0[slot], rS1[slot], rFr[slot], p, gg, jj);
            sv[tid] = log(gg) / dlnE + zs;
            wv[tid] = jj * gg * gg * gg;
        }
        __syncthreads();
        if (slot < nr) {
            int t1 = (p + 1 == NP) ? tid - (NP - 1) : tid + 1;
            double W = 0.5 * (wv[tid] + wv[t1]) * dphi * rA[slot];
            int kf, kl;
            rg_deposit(H, K_LO, nH, sv[tid], sv[t1], W, kf, kl);
        }
        __syncthreads();
    }
    for (int i = tid; i < nH; i += KT) H_out[i] = H[i];
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
    size_t smem = (size_t)nH * sizeof(double);
    if (smem > 200 * 1024) return rg_set_error(RG_E_GRID, "energy grid too fine for the kyconv histogram");
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
