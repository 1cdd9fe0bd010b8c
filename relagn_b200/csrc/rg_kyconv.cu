// rg_kyconv.cu -- the stand-alone convolution seam (B1):
//   xspec.callModelFunction('kyconv', energies, params[12], flux)
//   (src/python_version/relagn.py:979-1001; tests/blurr_test.py:51-54,111-119)
// One annulus of arbitrary width and emissivity (alpha, beta, r_break): one CTA of 8 warps builds the table slice of
// (spin, inclination) in shared memory, then the shift histogram -- ring m goes to warp m mod 8, every warp gathers
// into its own shared-memory copy (rg_transfer.cuh: no atomics), the copies are folded in warp order -- and a
// second kernel applies it to the photon spectrum.  The same device functions build every annulus in hist_kernel.
#include "rg_kernels.h"
#include "rg_transfer.cuh"

#define KT 256
#define KW (KT / 32)

static __host__ __device__ inline int ky_round_up2(int x) { return (x + 1) & ~1; }
static __host__ __device__ inline int ky_dn(int nH) { return ky_round_up2(nH) < 512 ? ky_round_up2(nH) : 512; }

template <int NQ>
__global__ void __launch_bounds__(KT)
ky_hist_kernel(TabDesc tab, double a, double theta_o, double r1, double r2, double alpha, double beta, double rb,
               double zs, double dlnE, int nsub, double *__restrict__ H_out, int K_LO, int nH, int *__restrict__ status,
               double2 *slice)
{
    RG_DYN_SMEM(smem_raw);
    constexpr int NP = 32 * NQ;
    double *H = reinterpret_cast<double *>(smem_raw); // [KW][nH2] | per warp ring scratch
    const int nH2 = ky_round_up2(nH), DN = ky_dn(nH);
    const int tid = threadIdx.x, warp = tid >> 5;
    double *scr_base = H + (size_t)KW * nH2 + (size_t)warp * ky_round_up2(RG_RING_SCR_DOUBLES(NP, DN));
    for (int i = tid; i < KW * nH2; i += KT) H[i] = 0.0;
    __shared__ TabSel s_ts;
    __shared__ int s_hi;
    if (tid == 0) {
        rg_tab_select(s_ts, tab, a, theta_o);
        const double rtop = r2 < RG_TAB_ROUT ? r2 : RG_TAB_ROUT;
        int hi = (int)floor(rg_row_pos(tab.x0, tab.dx, tab.NR, rtop)) + 1;
        if (hi > tab.NR - 1) hi = tab.NR - 1;
        if (hi < s_ts.s_lo + 1) hi = s_ts.s_lo + 1;
        s_hi = hi;
        *status = s_ts.status;
    }
    __syncthreads();
    const TabSel &ts = s_ts;
    if (ts.status == 0) {
        rg_build_slice(ts, s_hi, slice, tid, KT);
        __syncthreads();
        int kmn, kmx;
        warp_build_rings<NQ>(H + (size_t)warp * nH2, K_LO, nH, ts, slice, rg_ring_scr(scr_base, NP, DN), r1, r2, 1.0 / dlnE,
                             nsub, warp, KW, alpha, beta, rb, zs, 1.0, kmn, kmx);
    }
    __syncthreads();
    for (int i = tid; i < nH; i += KT) {
        double v = 0.0;
        for (int w = 0; w < KW; w++) v += H[(size_t)w * nH2 + i];
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

template <int NQ>
static int launch_ky_hist(const TabDesc &tab, double a, double theta_o, double r1, double r2, double alpha, double beta,
                          double rb, double zs, double dlnE, int nsub, double *d_H, int K_LO, int nH, int *d_status,
                          double2 *d_slice, cudaStream_t st)
{
    size_t smem = (size_t)KW * ky_round_up2(nH) * sizeof(double) +
                  (size_t)KW * ky_round_up2(RG_RING_SCR_DOUBLES(32 * NQ, ky_dn(nH))) * sizeof(double);
    if (smem > 220 * 1024) return rg_set_error(RG_E_GRID, "energy grid too fine for the kyconv histogram");
    auto kern = ky_hist_kernel<NQ>;
    RG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RG_LAUNCH(kern, dim3(1), dim3(KT), smem, st, tab, a, theta_o, r1, r2, alpha, beta, rb, zs, dlnE, nsub, d_H, K_LO, nH,
              d_status, d_slice);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

int rg_launch_ky_hist(const TabDesc &tab, double a, double theta_o, double r1, double r2, double alpha, double beta,
                      double rb, double zshift, double dlnE, int nsub, double *d_H, int K_LO, int nH, int *d_status,
                      double2 *d_slice, cudaStream_t st)
{
    double zs = 0.0;
    if (zshift != 0) zs = -log(1 + zshift) / dlnE;
    if (nsub <= 0) nsub = rg_auto_nsub(r1, r2, log(r2 / r1), dlnE);
    switch (tab.NPHI) {
    case 32: return launch_ky_hist<1>(tab, a, theta_o, r1, r2, alpha, beta, rb, zs, dlnE, nsub, d_H, K_LO, nH, d_status, d_slice, st);
    case 64: return launch_ky_hist<2>(tab, a, theta_o, r1, r2, alpha, beta, rb, zs, dlnE, nsub, d_H, K_LO, nH, d_status, d_slice, st);
    case 128: return launch_ky_hist<4>(tab, a, theta_o, r1, r2, alpha, beta, rb, zs, dlnE, nsub, d_H, K_LO, nH, d_status, d_slice, st);
    case 256: return launch_ky_hist<8>(tab, a, theta_o, r1, r2, alpha, beta, rb, zs, dlnE, nsub, d_H, K_LO, nH, d_status, d_slice, st);
    }
    return rg_set_error(RG_E_ARG, "NPHI must be 32, 64, 128 or 256");
}

int rg_launch_ky_conv(const double *d_H, int kmin, int nk, const double *d_in, double *d_out, int nE,
                      cudaStream_t st)
{
    RG_LAUNCH(ky_conv_kernel, dim3((nE + KT - 1) / KT), dim3(KT), 0, st, d_H, kmin, nk, d_in, d_out, nE);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}
