// rg_kernels.h -- internal launch interface between the C-ABI layer (rg_api.cu)
// and the kernel translation units.
#pragma once
#include "rg_common.cuh"

// Kerr transfer tables resident on the device
struct TabDesc {
    const float *data;     // [n_spin][n_inc][2][NR][NPHI]
    const double *spins;   // [n_spin] ascending
    const double *riscos;  // [n_spin]
    const double *thetas;  // [n_inc] ascending, radians
    int n_spin, n_inc, NR, NPHI;
    double dl;
    double lng_min, lng_max; // global range of ln g over all nodes (sizes the shift histogram)
};

// energy grid resident on the device
struct GridDesc {
    const double *Egrid; // [nE] bin centres (keV)
    const double *nu;    // [nE] Hz
    const double *bbpre; // [nE] 2 h nu^3 / c^2
    const double *lgE;   // [nE] log10(Egrid)
    int nE;
    int geometric;
    double dlnE;         // ln(E_{j+1}/E_j) when geometric
};

int rg_launch_setup(const double *d_params, int P, int model, double dr_dex, SetRec *recs, int *sys_counter,
                    int2 *syslist, int sys_cap, double *rq_scratch, int rq_cap, double *d_attrs, cudaStream_t st);

int rg_launch_nthcomp(const SetRec *recs, const int2 *syslist, const int *sys_counter, int sys_cap,
                      const double *p10tab, double *sc_gam, double *sc_g, double *sc_bet, size_t sc_stride,
                      double *sptot, double *header, cudaStream_t st);

int rg_launch_spectrum(const SetRec *recs, int P, const GridDesc &grid, const TabDesc &tab, unsigned flags,
                       const double *p10tab, const double *sptot, const double *header, double *d_out,
                       cudaStream_t st);

// generic kyconv seam: histogram of one (possibly wide) annulus into d_H[K_LO .. K_LO+nH)
int rg_launch_ky_hist(const TabDesc &tab, double a, double theta_o, double r1, double r2, double alpha, double beta,
                      double rb, double zshift, double dlnE, int nsub, double *d_H, int K_LO, int nH,
                      cudaStream_t st);
int rg_launch_ky_conv(const double *d_H, int kmin, int nk, const double *d_in, double *d_out, int nE,
                      cudaStream_t st);

// table generation (device ray tracer)
int rg_launch_raytrace(const double *h_spins, int n_spin, const double *h_thetas, int n_inc, int NR, int NPHI,
                       double dl, float *d_data, long *n_unresolved, cudaStream_t st);

int rg_fp64_peak(double *tflops);
