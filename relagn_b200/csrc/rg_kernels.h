// rg_kernels.h -- internal launch interface between the C-ABI layer (rg_api.cu)
// and the kernel translation units.
#pragma once
#include "rg_common.cuh"

#define RG_SPEC_THREADS 256

// Kerr transfer tables resident on the device (format 2: every node on the same physical radii)
#define RG_TAB_ROUT 1000.0 // outermost table row (the reference leaves kyconv at log r > 3, relagn.py:992)
#define RG_ROW_C 1.0       // rows uniform in x(r) = 1/r + RG_ROW_C / sqrt(r)
struct TabDesc {
    const float *data;     // [n_spin][n_inc][2][NR][NPHI]: g, Jn (the layout of the C ABI)
    const float2 *pairs;   // [n_spin][n_inc][NR][NPHI] of (ln g, Jn): what the kernels read
    const double *spins;   // [n_spin] ascending
    const double *thetas;  // [n_inc] ascending, radians
    const int *first_row;  // [n_spin]: first row with data (rows inside 1.02 x photon orbit are empty)
    int n_spin, n_inc, NR, NPHI;
    double r_in;           // radius of row 0
    double x0, dx;         // row position of radius r: (x(r) - x0) / dx
    double lng_min, lng_max; // range of ln g over the usable rows of all nodes (sizes the shift histogram)
};

// energy grid resident on the device (all [nE] unless noted; built by rg_set_grid)
struct GridDesc {
    const double *Egrid; // bin centres (keV): left edge + dE/2      relagn.py:335-338
    const double *dE;    // bin widths
    const double *nu;    // Hz                                         relagn.py:339
    const double *hnu;   // h * nu   (numerator of relagn.py:48)
    const double *bbpre; // 2 h nu^3 / c^2                              relagn.py:47
    const double *wt;    // trapezoid weights of np.trapz(., nu) on this grid (with the extension joint when n_ext)
    const double *e511;  // Egrid / 511                                pyNTHCOMP.py:70
    const double *ear2;  // Egrid^2                                    pyNTHCOMP.py:75
    const double *lgE;   // log10(Egrid): seeds the Kompaneets-grid cursor search
    const double *rear2; // 1 / Egrid^2
    // low-frequency extension of disc_annuli (relagn.py:868-873), [RG_NEXT] when n_ext
    const double *x_hnu, *x_bbpre, *x_wt;
    int n_ext;
    int nE;
    int geometric;
    double dlnE;         // ln(E_nE/E_0)/nE when geometric
};

struct SpecArgs {
    const SetRec *recs;
    int P;
    GridDesc grid;
    TabDesc tab;
    unsigned flags;
    const double *farr;   // [sys][nE] Comptonised spectra on the energy grid (nthcomp_kernel)
    double *out;          // [P][nE] or [P][4][nE]
    int K_LO, nH;         // widest shift-histogram support over all table nodes: k in [K_LO, K_LO + nH)
    int HL, HH;           // zero halo columns of the transposed local-spectrum buffer
    int NB;               // annulus blocks per parameter set (work items = P * NB)
    double *partial;      // NB > 1: [P][NB][ncomp][nE] partial sums, folded by the reduce kernel
    const double *hist;   // shift histograms built by hist_kernel: [annulus][hstride], slot 0 = (first shift, count)
    int hstride;
    const double2 *edges; // (lo, hi) of every annulus in log10 r: [annulus], rows as in the histogram arena (edges_kernel)
    const double2 *wien;  // [P] (kT*, k*) of every set's reference annulus for the Wien cut (wienref_kernel), or null
    int *queue;           // device work-queue counter (zeroed before the launch)
    double pkupk;         // (4 dE / E) (E / dE) of the geometric grid: the pack-unpack factor of relagn.py:970-974, 1003-1006
    int warps;            // warps per CTA
    unsigned long long *stats; // optional device counters [4]: sum W_ann, conv annuli, sum jmax, systems
};

int rg_spectrum_pick_NC(int nE); // 256-bin chunks per warp pass: 4, 8 or 20; 0 if nE is too large
// shared memory per CTA and the number of warps per CTA that fit (0: does not fit at all)
size_t rg_spectrum_smem_bytes(int NC, int nH, int HL, int HH, int warps, int fp32);
int rg_spectrum_pick_warps(int NC, int nH, int HL, int HH, int fp32);

int rg_launch_setup(const double *d_params, int P, int model, double dr_dex, SetRec *recs, int *sys_counter,
                    int2 *syslist, int sys_cap, double *rq_scratch, int rq_cap, double *d_attrs, int ann_cap,
                    cudaStream_t st);

// n_threads = threads of the solve kernel's grid (systems are dealt round robin); coef = [sys_cap][RG_SPT_STRIDE] double2
int rg_launch_nthcomp(const SetRec *recs, const int2 *syslist, int n_warm_sys, int n_hot_sys, int sys_cap,
                      const double *p10tab, int n_threads, double2 *coef, double *header, int sm_count,
                      const GridDesc &grid, double lgE0, double grid_E1, double *gtab, size_t gtab_doubles, double *farr,
                      unsigned long long *stats, cudaStream_t st);

// latency variant of the Kompaneets kernels for small batches (rg_nthcomp_small.cu): one warp per system
int rg_launch_nthcomp_small(const SetRec *recs, const int2 *syslist, int n_warm_sys, int n_hot_sys, int sys_cap,
                            const double *p10tab, const double2 *edges, double *gscratch, size_t gscratch_doubles,
                            int sm_count, const GridDesc &grid, double *farr, int *queue, unsigned long long *stats,
                            cudaStream_t st);
int rg_launch_spectrum(const SpecArgs &args, int sm_count, cudaStream_t st);
int rg_launch_wienref(const SetRec *recs, int P, const double2 *edges, const double *hist, int hstride, int rel,
                      double2 *wien, cudaStream_t st);

// shift histograms of every annulus of every set (rg_hist.cu); annulus g of set p lives at row recs[p].ann_base + g
struct HistArgs {
    const SetRec *recs;
    int P;
    TabDesc tab;
    double dlnE, inv_dlnE;
    int K_LO, nH;  // widest support over all table nodes
    int NBH;       // CTAs per parameter set
    double *hist;  // [total annuli][hstride]
    int hstride;   // >= nH + 1
    const double2 *edges; // [total annuli] (lo, hi) in log10 r (edges_kernel)
    double2 *slices;      // [slice_slots][NR][NPHI] scratch: the table slice of the set a resident CTA works on
    int slice_slots;
    int *queue;           // device work-queue counter (zeroed by the launcher)
    struct SelRec *sel;   // [P] table selection of every set (tabsel_kernel)
    int HSZ, DN, warp_bytes; // shared-memory layout of a warp (filled by the launcher)
};
int rg_launch_hist(const HistArgs &A, int sm_count, cudaStream_t st);
size_t rg_selrec_bytes();
// radial bin edges of every annulus of every set (relagn.py:731-776), one thread per (set, region): the chain of
// repeated subtractions is walked once here instead of by every consumer lane
int rg_launch_edges(const SetRec *recs, int P, double2 *edges, cudaStream_t st);

// generic kyconv seam: histogram of one (possibly wide) annulus into d_H[0..nH) (shift k = K_LO + index)
int rg_launch_ky_hist(const TabDesc &tab, double a, double theta_o, double r1, double r2, double alpha, double beta,
                      double rb, double zshift, double dlnE, int nsub, double *d_H, int K_LO, int nH, int *d_status,
                      double2 *d_slice, cudaStream_t st);
// out[j] = sum_i H[i] in[j - (kmin + i)]
int rg_launch_ky_conv(const double *d_H, int kmin, int nk, const double *d_in, double *d_out, int nE,
                      cudaStream_t st);

// table generation (device ray tracer): fills d_data [n_spin][n_inc][2][NR][NPHI]
int rg_launch_raytrace(const double *h_spins, int n_spin, const double *h_thetas, int n_inc, int NR, int NPHI,
                       double r_in, int limb, float *d_data, long *n_unresolved, cudaStream_t st);
// host-side geometry of the table format (shared by the ray tracer, the loader and the kernels' row lookup)
double rg_row_x(double r);
double rg_row_radius(double r_in, int NR, int s);
double rg_r_valid(double a);

// layout of the resident 10^(0.02 j) table (rg_ctx::d_p10): five runs of RG_NTH + 4 doubles
#define RG_P10_DPH (3 * (RG_NTH + 4))       // 10^(0.06 j) / (exp(1e-4 10^(0.02 j)) - 1): seed photons x dphdot per unit planck xmin^3
#define RG_P10_RDP (4 * (RG_NTH + 4))       // 1 / (10^(0.02 (j+1)) - 10^(0.02 j)): reciprocal cell widths (interpolation kernel)
#define RG_P10_W1 (RG_NTH + 4)       // sqrt(10^(0.02 j) 10^(0.02 (j+1))): w1 of pyNTHCOMP.py:117 per unit xmin
#define RG_P10_RW1 (2 * (RG_NTH + 4)) // 1 / that
#define RG_P10_LEN (5 * (RG_NTH + 4))
#ifndef RG_NTH_MINB
#define RG_NTH_MINB 4 // resident 128-thread CTAs per SM of the persistent nthcomp grid
#endif
int rg_fp64_peak(double *tflops, cudaStream_t st);
int rg_launch_counters_out(const int *d_counter, int *h_mapped, cudaStream_t st);

// ---- observation epilogue (rg_observe.cu) ----
struct ObsArgs {
    const double *lnu;    // [P][nE] total SED, W/Hz
    const double *params; // [P][RG_NPAR]
    int P, model;
    GridDesc grid;
    const double *ebins;  // [nE+1] model bin edges (device)
    int nE;
    const double *kph;    // [nE] dE / (h E): W/Hz -> photons/s per model bin (rg_set_grid)
    const double *ear;    // [ne+1] instrument bin edges, observer frame (device)
    int ne;
    const double *pos0;   // [ne] ln(ear_i / ebins_0) / dlnE when the model grid is geometric, else null
    double inv_dlnE;
    double *photar;       // [P][ne] or null
    // response in sliced ELL: slice s = channels 32 s .. 32 s + 31, entries sl_off[s] .. sl_off[s+1], laid out [k][lane];
    // col < 0 marks padding
    int n_chan;           // 0: rebin only
    const int *sl_off;
    const int *ell_col;
    const double *ell_val;
    double exposure;
    const double *counts, *sigma; // sigma: 1 / sigma_c (chi^2) or ln D_c (Cash; 0 where D_c = 0)
    int kind;
    double *folded;       // [P][n_chan] or null
    double *stat;         // [P]
};
size_t rg_observe_smem_bytes(int nE, int ne);
int rg_launch_observe(const ObsArgs &A, int sm_count, cudaStream_t st);
int rg_launch_convert(const double *d_lnu, const double *d_params, int P, int model, int ncomp, int unit, int as_flux,
                      const GridDesc &g, double *d_out, int sm_count, cudaStream_t st);
