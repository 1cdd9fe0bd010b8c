/*
 * relagn_b200.h -- C ABI of the B200-native RELAGN spectrum-evaluation path.
 *
 * Plain pointers and sizes only (no torch / C++ types).  The library is built
 * for sm_100a and has NO CPU fallback: every compute entry point fails with
 * RG_E_CUDA when no usable device is present.
 *
 * Seams replaced (paths relative to the reference repo scotthgn/RELAGN):
 *
 *  B1  xspec.callModelFunction('kyconv', energies, params[12], flux)
 *        src/python_version/relagn.py:979-1001, 1095-1099, 1353-1357
 *        (Fortran twin: call kyconv(es, nn, kpar, ifl, ph, pherr),
 *         src/fortran_version/relagn_dir/relagn.f:359,424,547)
 *      -> rg_kyconv()
 *
 *  B2  the model classes' spectrum path: relagn.__init__ / relqso.__init__
 *        (relagn.py:232-357, 1767-1857) + get_totSED / get_*Component
 *        (relagn.py:1412-1577) -> do_rel*Spec / do_nonrel*Spec
 *        (relagn.py:950-1156, 1322-1401)
 *      -> rg_eval_batch() (device buffers) / rg_eval_batch_host() (host buffers)
 *
 *  Kerr transfer tables (KYCONV's FITS tables, not in the reference repo)
 *      -> rg_tables_generate() (device ray tracer) / rg_tables_upload()
 *
 * Conventions: all entry points return RG_OK (0) or a negative RG_E_* code;
 * rg_last_error() returns a static, thread-local message for the last
 * failure.  The caller owns every buffer passed in; the context owns only its
 * scratch, grid and table storage.  One context per device; calls on one
 * context must be serialised by the caller (stream-ordered on `stream`).
 */
#ifndef RELAGN_B200_H
#define RELAGN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_VERSION 100

#define RG_NPAR 15  /* relagn ctor order (relagn.py:232-247):
                       M, dist, log_mdot, a, cos_inc, kTe_hot, kTe_warm,
                       gamma_hot, gamma_warm, r_hot, r_warm, log_rout, fcol,
                       h_max, z.   relqso (relagn.py:1767-1774) uses slots
                       0..6 = M, dist, log_mdot, a, cos_inc, fcol, z. */
#define RG_NATTR 24

/* attribute slots of the per-set record returned next to the spectra
 * (the reference's public attributes, docs/relagn_docs.tex:202-242) */
enum {
    RG_A_RISCO = 0, RG_A_RSG, RG_A_ETA, RG_A_LEDD, RG_A_MDOT_EDD, RG_A_RG,
    RG_A_RH, RG_A_RW, RG_A_ROUT, RG_A_HMAX, RG_A_GAMMA_H, RG_A_KTE_H,
    RG_A_KTE_W, RG_A_GAMMA_W, RG_A_KTSEED, RG_A_LDISS, RG_A_LSEED,
    RG_A_NDISC, RG_A_NWARM, RG_A_NHOT, RG_A_FLAGS, RG_A_STATUS, RG_A_MDOT,
    RG_A_INC
};

/* RG_A_FLAGS bits: the reference's print-and-clamp warnings (relagn.py:388-421) */
#define RG_F_RW_LT_RH 1
#define RG_F_RW_GT_ROUT 2
#define RG_F_RH_LT_RISCO 4
#define RG_F_RW_LT_RISCO 8
#define RG_F_HMAX_GT_RH 16
#define RG_F_HOT_SEED_GE_KTE 32 /* kT_seed >= kTe_hot: hot shape is zero (relagn.py:1212-1221) */

/* RG_A_STATUS per set: the reference's ValueError conditions */
#define RG_S_OK 0
#define RG_S_SPIN 1   /* relagn.py:372-377 */
#define RG_S_INC 2    /* relagn.py:380-386 */
#define RG_S_MDOT 3   /* relagn.py:1861-1874 (relqso) */
#define RG_S_CAP 4    /* internal capacity exceeded (radial bins / nthcomp grid) */

/* model selector */
#define RG_MODEL_RELAGN 0
#define RG_MODEL_RELQSO 1

/* rg_eval_batch flags */
#define RG_FLAG_REL 1u        /* relativistic transfer (get_*(rel=True)) */
#define RG_FLAG_COMPONENTS 2u /* out is [P][4][nE] = disc, warm, hot, total; else [P][nE] total */
#define RG_FLAG_FP32 4u       /* single-precision local spectra / transfer convolution / accumulation (outputs stay
                                 double; setup, Kompaneets solve, histograms and normalisations stay FP64) */

/* error codes */
#define RG_OK 0
#define RG_E_CUDA (-1)     /* CUDA runtime failure / no device */
#define RG_E_ARG (-2)      /* bad argument */
#define RG_E_NOTABLES (-3) /* relativistic call without tables */
#define RG_E_GRID (-4)     /* energy grid unsupported for this call (rel needs a geometric grid) */
#define RG_E_UNSUPPORTED (-5) /* kyconv parameter combination not implemented */
#define RG_E_NOMEM (-6)

typedef struct rg_ctx rg_ctx;

int rg_version(void);
const char *rg_last_error(void);

/* one context per device.  device < 0: current device. */
int rg_ctx_create(int device, rg_ctx **out);
void rg_ctx_destroy(rg_ctx *ctx);

/* ---- energy grid (relagn.py:335-346, new_ear relagn.py:780-805) ----
 * ebins: host, nE+1 ascending bin edges in keV.  Derives dE, bin centres
 * (left + dE/2), nu grid; detects a geometric grid (needed for RG_FLAG_REL). */
int rg_set_grid(rg_ctx *ctx, const double *ebins_host, int nE);

/* ---- Kerr transfer tables (format 2) ----
 * Node grid: spins[n_spin] ascending, thetas[n_inc] ascending (radians).
 * Per node two planes g[NR][NPHI], jn[NR][NPHI] (float32).  Rows are the SAME physical radii for every node,
 * uniform in x(r) = 1/r + 1/sqrt(r) from r_in (row 0) to 1000 r_g (row NR-1: the reference leaves kyconv beyond
 * log r = 3, relagn.py:992); rows inside 1.02 x the photon orbit of the node's spin carry no data (g = jn = 0).
 * Columns phi_t = 2 pi t/NPHI + phi0(r_s), phi0 = disc azimuth of the ray alpha = 0, beta < 0.
 * Layout [n_spin][n_inc][2][NR][NPHI].  A model (a, i) is looked up by Lagrange interpolation over up to 4 spin
 * nodes x 4 inclination nodes of ln g and jn at fixed physical radius; adjacent spin nodes must both have data at
 * the innermost radius of any model between them (upload/generate fail otherwise).
 * generate: ray-traces on the device (*n_unresolved: points the solver lost; they keep jn = 0).
 * upload/download: host <-> device. */
int rg_tables_generate(rg_ctx *ctx, const double *spins, int n_spin, const double *thetas, int n_inc, int NR,
                       int NPHI, double r_in, long *n_unresolved);
int rg_tables_upload(rg_ctx *ctx, const double *spins, int n_spin, const double *thetas, int n_inc, int NR, int NPHI,
                     double r_in, const float *data_host);
/* Limb-darkening law of the local emission (kyconv parameter 10, relagn.py:988; the reference passes 0): 0 isotropic,
 * -1 Laor (1991) 1 + 2.06 mu_e, -2 Haardt (1993) ln(1 + 1/mu_e).  The law is a property of the table set (it multiplies
 * the weight of every sample): call before rg_tables_generate, or after rg_tables_upload to state what the uploaded
 * set was built with.  rg_kyconv refuses a parameter 10 that differs from it. */
int rg_tables_set_limb(rg_ctx *ctx, int law);
int rg_tables_download(rg_ctx *ctx, float *data_host, size_t n_floats);
int rg_tables_info(rg_ctx *ctx, int *n_spin, int *n_inc, int *NR, int *NPHI, double *r_in, size_t *n_floats);

/* ---- B1: the kyconv seam ----
 * energies: host, nE+1 edges (geometric grid); par: the reference's 12
 * parameters [a, inc_deg, r_in, ms, r_out, alpha, beta, r_break, zshift,
 * ntable, nE_internal, norm]; flux: host, nE photons/bin, convolved in place.
 * Supported: ms = 0, ntable = 0 (isotropic), norm = -1 (no renormalisation)
 * -- exactly what the reference passes (relagn.py:997-998). */
int rg_kyconv(rg_ctx *ctx, const double *energies_host, int nE, const double *par12, double *flux_inout_host);

/* shift-histogram of one annulus on a log grid of step dlnE (diagnostic /
 * test entry; the same device code builds every annulus inside
 * rg_eval_batch).  H: host, cap doubles; *kmin, *nk describe the support. */
int rg_ky_kernel(rg_ctx *ctx, double a, double theta_o, double r1, double r2, double alpha, double beta, double rb,
                 double zshift, double dlnE, double *H_host, int cap, int *kmin, int *nk);

/* ---- B2: batched spectrum evaluation ----
 * d_params: DEVICE [P][RG_NPAR]; d_out: DEVICE [P][nE] or [P][4][nE] (W/Hz,
 * the reference's internal Lnu before unit conversion); d_attrs: DEVICE
 * [P][RG_NATTR] or NULL.  Uses the grid of rg_set_grid and, with RG_FLAG_REL,
 * the resident tables.  stream: a cudaStream_t (NULL = default stream).
 * Asynchronous with respect to the host. */
int rg_eval_batch(rg_ctx *ctx, const double *d_params, int P, int model, double dr_dex, unsigned flags,
                  double *d_out, double *d_attrs, void *stream);

/* rg_eval_batch with a completion hook per piece: the spectra come out in
 * pieces of about piece_sets parameter sets (a short remainder joins the last
 * piece; batches small enough to be split over several CTAs per set come as
 * one piece per chunk); on_piece(user, first_set, n_sets) is called on the
 * calling thread right after the kernels that complete rows [first_set,
 * first_set + n_sets) of d_out (and d_attrs) have been enqueued on `stream`
 * -- record an event there and chain the transfer of that range (D2H copy,
 * NCCL gather) on another stream while the next piece is computed.  This is
 * what rg_eval_batch_host and batch.evaluate_sharded do; the reference has
 * no counterpart (its batch loop is the caller's Python loop over objects). */
typedef void (*rg_piece_fn)(void *user, int first_set, int n_sets);
int rg_eval_batch_pieces(rg_ctx *ctx, const double *d_params, int P, int model, double dr_dex, unsigned flags,
                         double *d_out, double *d_attrs, void *stream, int piece_sets, rg_piece_fn on_piece,
                         void *user);

/* same with HOST buffers: stages through pinned memory, copies in and out
 * inside the call, synchronises before returning. */
int rg_eval_batch_host(rg_ctx *ctx, const double *params_host, int P, int model, double dr_dex, unsigned flags,
                       double *out_host, double *attrs_host);

/* ---- f1: observation epilogue on the device -------------------------------
 * Unit conversion of rg_eval_batch's W/Hz spectra, restating the reference's
 * host epilogue _to_newUnit / _to_flux (relagn.py:487-564) for whole batches:
 *   RG_UNIT_SI       W/Hz               (identity)
 *   RG_UNIT_CGS      erg/s/Hz           L * 1e7
 *   RG_UNIT_CGS_WAVE erg/s/Angstrom     L nu^2 / c * 1e-10 * 1e7
 *   RG_UNIT_SI_WAVE  W/Angstrom         L nu^2 / c * 1e-10
 *   RG_UNIT_COUNTS   photons/s/keV      (L / h) / E
 * as_flux != 0 divides by 4 pi d^2 (1 + z) with d in cm (cgs, cgs_wave,
 * counts) or m (SI, SI_wave) -- distance and redshift are read per set from
 * d_params (relagn slots 1 and 14, relqso slots 1 and 6).
 * d_lnu, d_out: DEVICE [P][ncomp][nE]; d_out may alias d_lnu. */
#define RG_UNIT_SI 0
#define RG_UNIT_CGS 1
#define RG_UNIT_CGS_WAVE 2
#define RG_UNIT_SI_WAVE 3
#define RG_UNIT_COUNTS 4
int rg_convert_batch(rg_ctx *ctx, const double *d_lnu, const double *d_params, int P, int model, int ncomp, int unit,
                     int as_flux, double *d_out, void *stream);

/* Instrument grid: the XSPEC additive-model contract of the Fortran twin
 * (relagn.f:17-115: photar(ne) in photons/cm^2/s per bin of the caller's
 * ear(0:ne)).  The model grid is shifted to the observer's frame, edges and
 * photons both divided by (1 + z) (relagn.f:98-102), and rebinned by
 * fractional bin overlap -- XSPEC's inibin/erebin (relagn.f:110-112; HEASOFT,
 * restated from its documented behaviour).
 * ear_host: ne + 1 ascending edges in keV (observer frame). */
int rg_set_instrument(rg_ctx *ctx, const double *ear_host, int ne);
/* d_lnu: DEVICE [P][nE] total SED in W/Hz; d_photar: DEVICE [P][ne]. */
int rg_photar_batch(rg_ctx *ctx, const double *d_lnu, const double *d_params, int P, int model, double *d_photar,
                    void *stream);

/* ---- f2: fit statistic on the device --------------------------------------
 * Folds photar through an instrument response and reduces against observed
 * counts, so that one scalar per parameter set leaves the GPU (the fit / MCMC
 * use the reference documents, docs/relagn_docs.tex:136).
 * Response in CSR over detector channels:
 *   model_c = exposure * sum_{k in [rowptr[c], rowptr[c+1])} val[k] * photar[col[k]]
 * (val = RMF x ARF in cm^2, col = index into the instrument grid).
 * Statistic (XSPEC definitions):
 *   RG_FIT_CHI2  sum_c ((D_c - model_c) / sigma_c)^2
 *   RG_FIT_CSTAT 2 sum_c (model_c - D_c + D_c (ln D_c - ln model_c)), D_c = 0 -> model_c */
#define RG_FIT_CHI2 0
#define RG_FIT_CSTAT 1
int rg_set_response(rg_ctx *ctx, int n_chan, const int *rowptr_host, const int *col_host, const double *val_host,
                    double exposure);
/* counts_host [n_chan]; sigma_host [n_chan] (RG_FIT_CHI2) or NULL (RG_FIT_CSTAT) */
int rg_set_data(rg_ctx *ctx, const double *counts_host, const double *sigma_host, int kind);
/* d_lnu: DEVICE [P][nE]; d_stat: DEVICE [P]; d_folded: DEVICE [P][n_chan] model counts or NULL */
int rg_fitstat_batch(rg_ctx *ctx, const double *d_lnu, const double *d_params, int P, int model, double *d_stat,
                     double *d_folded, void *stream);
/* parameters -> statistic in one call: the spectra stay in a context-owned slab and never leave the device.
 * d_params DEVICE [P][RG_NPAR], d_stat DEVICE [P], d_attrs DEVICE [P][RG_NATTR] or NULL. */
int rg_eval_fitstat_batch(rg_ctx *ctx, const double *d_params, int P, int model, double dr_dex, unsigned flags,
                          double *d_stat, double *d_attrs, void *stream);
/* same with HOST buffers (15 doubles in, 1 double out per set) */
int rg_eval_fitstat_batch_host(rg_ctx *ctx, const double *params_host, int P, int model, double dr_dex,
                               unsigned flags, double *stat_host, double *attrs_host);

/* ---- B3: XSPEC local-model ABI (Fortran twins) ------------------------------
 * relagn_ / relqso_ are the C symbols a gfortran-built XSPEC resolves for
 *   subroutine relagn(ear,ne,param,ifl,photar,photer)   relagn.f:17-115, lmod_relagn.dat (15 parameters)
 *   subroutine relqso(ear,ne,param,ifl,photar,photer)   relqso.f, lmod_relqso.dat (6 parameters + $rel switch)
 * real*4 arrays by reference: ear(0:ne) keV, photar(ne) photons/cm^2/s per bin.  They restate the wrapper of
 * relagn.f:60-112 (2000-bin model grid covering the caller's grid times 1 + z, redshift, inibin/erebin rebin,
 * `save oldpar` memoisation) over rg_eval_batch + rg_photar_batch; the physics is the Python class's.
 * The first call creates a process-wide context on the current device and loads the default Kerr tables
 * (rg_tables_default); rg_xspec_bind substitutes a caller-owned context. */
void relagn_(const float *ear, const int *ne, const float *param, const int *ifl, float *photar, float *photer);
void relqso_(const float *ear, const int *ne, const float *param, const int *ifl, float *photar, float *photer);
/* double-precision core of the two: par[RG_NPAR] in the Python constructor order, flags = RG_FLAG_REL or 0 */
int rg_xspec_eval(int model, unsigned flags, const double *ear_host, int ne, const double *par_host,
                  double *photar_host);
int rg_xspec_bind(rg_ctx *ctx);
/* the package's default Kerr node grid (11 spins x 14 inclinations x 64 x 64): from the .npy cache written by the
 * Python layer (npy_path, or $RELAGN_B200_TABLES when NULL) or ray-traced on the device when there is none */
int rg_tables_default(rg_ctx *ctx, const char *npy_path);

/* number of kernels launched by this context so far (for accounting) */
/* Fnu_seed_hot (relagn.py:1200-1223; an attribute of the reference's class): the hot corona's Comptonised spectral
 * shape of parameter set `set` of the LAST evaluated batch on the energy grid, W/Hz-shaped, nE doubles on the host;
 * zeros when the seed temperature is not below kTe_hot or the set has no hot region.  Valid for sets of the last
 * chunk of the last rg_eval_batch* call (any batch of <= 4096 sets is one chunk). */
int rg_last_hot_shape(rg_ctx *ctx, int set, double *out_host);

long rg_launch_count(rg_ctx *ctx);

/* ---- instrumentation (bench.py's roofline block) ----
 * rg_set_option(RG_OPT_PROFILE, 1): bracket every kernel of rg_eval_batch with
 * CUDA events on the caller's stream and count algorithmic work on the device.
 * rg_get_stat synchronises the recorded events and returns totals accumulated
 * since the last rg_reset_stats. */
#define RG_OPT_PROFILE 1
#define RG_STAT_MS_SETUP 1      /* summed kernel time, ms */
#define RG_STAT_MS_NTHCOMP 2
#define RG_STAT_MS_SPECTRUM 3
#define RG_STAT_N_SETUP 4       /* launches */
#define RG_STAT_N_NTHCOMP 5
#define RG_STAT_N_SPECTRUM 6
#define RG_STAT_SUM_W 7         /* sum over transfer-convolved annuli of the shift-histogram width W_ann */
#define RG_STAT_N_CONV_ANN 8    /* annuli that went through the transfer convolution */
#define RG_STAT_SUM_JMAX 9      /* sum over Kompaneets systems of jmax */
#define RG_STAT_N_SYS 10        /* Kompaneets systems solved */
#define RG_STAT_N_SETS 11       /* parameter sets evaluated */
#define RG_STAT_MS_OBSERVE 12    /* summed time of the observation-epilogue kernels, ms */
#define RG_STAT_N_OBSERVE 13
#define RG_STAT_MS_HIST 14       /* summed time of hist_kernel (shift histograms of the transfer), ms */
#define RG_STAT_N_HIST 15
int rg_set_option(rg_ctx *ctx, int key, double value);
int rg_get_stat(rg_ctx *ctx, int key, double *value);
int rg_reset_stats(rg_ctx *ctx);

/* measured FP64 FMA throughput of the device in TFLOP/s (micro-benchmark,
 * used as the roofline denominator of the FP64-bound kernels) */
int rg_measure_fp64_peak(rg_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
