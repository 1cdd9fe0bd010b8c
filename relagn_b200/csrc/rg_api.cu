// rg_api.cu -- the C ABI (include/relagn_b200.h): context, energy grid, Kerr
// table residency, the kyconv seam and the batched evaluation driver.
// Host-side only; every numeric result comes from the kernels in
// rg_setup.cu / rg_spectrum.cu / rg_kyconv.cu / rg_raytrace.cu.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "rg_kernels.h"

// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

int rg_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int rg_set_cuda_error(int e, const char *msg, const char *what, const char *file, int line)
{
    snprintf(g_err, sizeof g_err, "CUDA error %d (%s) in %s at %s:%d", e, msg, what, file, line);
    return RG_E_CUDA;
}

extern "C" const char *rg_last_error(void) { return g_err; }
extern "C" int rg_version(void) { return RG_VERSION; }

// ---------------------------------------------------------------------------
struct rg_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr; // used by the *_host entry points
    long launches = 0;
    // energy grid
    int nE = 0;
    double *d_grid = nullptr;
    GridDesc gd;
    std::vector<double> h_ebins;
    // tables
    bool have_tab = false;
    TabDesc td;
    float *d_tab = nullptr;
    float2 *d_tab2 = nullptr; // (ln g, jn) pairs: one 8-byte load per table node
    double *d_tabmeta = nullptr;
    int *d_firstrow = nullptr;
    int *d_kystatus = nullptr;
    size_t tab_floats = 0;
    std::vector<double> h_spins, h_thetas;
    std::vector<int> h_firstrow;
    int tab_limb = 0; // limb-darkening law the resident / next generated tables are built with (rg_tables_set_limb)
    double *h_pin = nullptr;        // pinned staging of the small-batch host path: params | spectra | attrs
    size_t h_pin_n = 0;
    int last_first = 0, last_n = 0; // sets of the last evaluated chunk (rg_last_hot_shape)
    cudaStream_t last_stream = nullptr;
    // batch scratch
    double *d_p10 = nullptr;
    SetRec *d_recs = nullptr;
    int chunk_sets = 0;
    int *d_counter = nullptr;
    int *h_counter = nullptr; // pinned
    int2 *d_syslist = nullptr;
    int sys_cap = 0;
    double *d_sptot = nullptr, *d_header = nullptr, *d_farr = nullptr;
    int farr_nE = 0;
    int farr_cap = 0; // systems the farr arena holds at the current grid (<= sys_cap)
    double *d_sc = nullptr;
    int nth_threads = 0;
    size_t sc_doubles = 0;
    double *d_rq = nullptr;
    int rq_cap = 0;
    int *d_queue = nullptr;
    double *d_wien = nullptr;    // [chunk] (kT*, k*) of the Wien cut's reference annulus
    size_t wien_n = 0;
    void *d_sel = nullptr;       // [chunk] table selections (SelRec) of the chunk's sets
    size_t sel_bytes = 0;
    double2 *d_slices = nullptr; // table-slice scratch of the resident hist_kernel CTAs (L2 resident)
    int slice_slots = 0;
    size_t slice_elems = 0;
    double *d_partial = nullptr;
    size_t partial_n = 0;
    double *d_hist = nullptr; // shift-histogram arena [annuli of the chunk][hstride]
    size_t hist_n = 0;
    double *d_edges = nullptr; // radial bin edges [annuli of the chunk] as (lo, hi) pairs
    size_t edges_n = 0;
    cudaStream_t side = nullptr;            // hist_kernel runs here, next to the Kompaneets solves on the caller's stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // kyconv seam scratch
    double *d_kyH = nullptr, *d_kyin = nullptr, *d_kyout = nullptr;
    int ky_cap = 0, kyH_cap = 0;
    // instrumentation
    bool profile = false;
    unsigned long long *d_stats = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[5]; // setup, nthcomp, spectrum, observation epilogue, hist
    double ms_acc[5] = {0, 0, 0, 0, 0};
    long n_acc[5] = {0, 0, 0, 0, 0};
    double n_sets = 0;
    // observation epilogue: model edges, instrument grid, response (sliced ELL), data, spectra slab
    double *d_ebins = nullptr;
    const double *d_kph = nullptr; // inside d_grid
    double *d_ear = nullptr, *d_pos0 = nullptr;
    std::vector<double> h_ear;
    bool pos0_dirty = true;
    int ne = 0;
    int n_chan = 0, fit_kind = RG_FIT_CHI2;
    bool have_data = false;
    int *d_sl_off = nullptr, *d_ell_col = nullptr;
    double *d_ell_val = nullptr, *d_counts = nullptr, *d_sigma = nullptr;
    double exposure = 1.0;
    double *d_slab = nullptr;
    size_t slab_n = 0;
    double *d_stage_stat = nullptr;
    size_t d_stage_stat_n = 0;
    // host staging (pinned) for rg_eval_batch_host
    double *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    double *d_stage_par = nullptr, *d_stage_out = nullptr, *d_stage_attr = nullptr;
    size_t d_stage_par_n = 0, d_stage_out_n = 0, d_stage_attr_n = 0;
};

#define RG_TRY(expr)                 \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != RG_OK) return rc__; \
    } while (0)

static void free_dev(void *p) { if (p) cudaFree(p); }
static int prof_collect(rg_ctx *c);

extern "C" int rg_ctx_create(int device, rg_ctx **out)
{
    if (!out) return rg_set_error(RG_E_ARG, "rg_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return rg_set_error(RG_E_CUDA, "no CUDA device available (%s): this library has no CPU path",
                            e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0) RG_CUDA_OK(cudaGetDevice(&device));
    if (device >= ndev) return rg_set_error(RG_E_ARG, "device %d out of range (%d devices)", device, ndev);
    RG_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp prop;
    RG_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    rg_ctx *c = new rg_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    memset(&c->gd, 0, sizeof c->gd);
    memset(&c->td, 0, sizeof c->td);
    RG_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    RG_CUDA_OK(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    RG_CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    RG_CUDA_OK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    // 10^(0.02 j) with the host libm, exactly as the reference's x grid (pyNTHCOMP.py:112-114)
    // followed by the reciprocal cell widths 1 / (10^(0.02 (j+1)) - 10^(0.02 j)) the interpolation kernel multiplies by
    // and by sqrt(10^(0.02 j) 10^(0.02 (j+1))) and its reciprocal for the Kompaneets coefficients (layout: rg_kernels.h)
    std::vector<double> p10(RG_P10_LEN, 0.0);
    for (int j = 0; j < RG_NTH + 4; j++) p10[j] = pow(10.0, j * 0.02);
    for (int j = 0; j < RG_NTH + 3; j++) {
        p10[RG_P10_RDP + j] = 1.0 / (p10[j + 1] - p10[j]);
        p10[RG_P10_W1 + j] = sqrt(p10[j] * p10[j + 1]);
        p10[RG_P10_RW1 + j] = 1.0 / p10[RG_P10_W1 + j];
        // x[j] dphdot[j] per unit planck xmin^3 (pyNTHCOMP.py:120-122): x[j] / tempbb = 1e-4 10^(0.02 j) for every system
        p10[RG_P10_DPH + j] = (p10[j] * p10[j] * p10[j]) / (exp(1e-4 * p10[j]) - 1);
    }
    RG_CUDA_OK(cudaMalloc(&c->d_p10, p10.size() * sizeof(double)));
    RG_CUDA_OK(cudaMemcpy(c->d_p10, p10.data(), p10.size() * sizeof(double), cudaMemcpyHostToDevice));
    RG_CUDA_OK(cudaMalloc(&c->d_counter, 4 * sizeof(int)));
    RG_CUDA_OK(cudaMallocHost(&c->h_counter, 4 * sizeof(int)));
    *out = c;
    return RG_OK;
}

extern "C" void rg_ctx_destroy(rg_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_dev(c->d_grid); free_dev(c->d_tab); free_dev(c->d_tab2); free_dev(c->d_tabmeta); free_dev(c->d_firstrow); free_dev(c->d_kystatus); free_dev(c->d_p10); free_dev(c->d_recs);
    free_dev(c->d_counter); free_dev(c->d_syslist); free_dev(c->d_sptot); free_dev(c->d_farr); free_dev(c->d_header); free_dev(c->d_sc);
    free_dev(c->d_rq); free_dev(c->d_queue); free_dev(c->d_slices); free_dev(c->d_sel); free_dev(c->d_wien); free_dev(c->d_partial); free_dev(c->d_hist); free_dev(c->d_edges); free_dev(c->d_kyH); free_dev(c->d_kyin); free_dev(c->d_kyout);
    free_dev(c->d_stage_par); free_dev(c->d_stage_out); free_dev(c->d_stage_attr); free_dev(c->d_stats);
    free_dev(c->d_ebins); free_dev(c->d_ear); free_dev(c->d_sl_off); free_dev(c->d_ell_col); free_dev(c->d_ell_val);
    free_dev(c->d_counts); free_dev(c->d_sigma); free_dev(c->d_slab); free_dev(c->d_stage_stat); free_dev(c->d_pos0);
    prof_collect(c);
    if (c->h_counter) cudaFreeHost(c->h_counter);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
}

extern "C" long rg_launch_count(rg_ctx *c) { return c ? c->launches : 0; }

// ---- instrumentation ----
static int prof_begin(rg_ctx *c, int kind, cudaStream_t st)
{
    if (!c->profile) return RG_OK;
    cudaEvent_t a, b;
    RG_CUDA_OK(cudaEventCreate(&a));
    RG_CUDA_OK(cudaEventCreate(&b));
    RG_CUDA_OK(cudaEventRecord(a, st));
    c->ev[kind].push_back(std::make_pair(a, b));
    return RG_OK;
}
static int prof_end(rg_ctx *c, int kind, cudaStream_t st)
{
    if (!c->profile) return RG_OK;
    RG_CUDA_OK(cudaEventRecord(c->ev[kind].back().second, st));
    return RG_OK;
}
static int prof_collect(rg_ctx *c)
{
    for (int k = 0; k < 5; k++) {
        for (auto &pr : c->ev[k]) {
            RG_CUDA_OK(cudaEventSynchronize(pr.second));
            float ms = 0;
            RG_CUDA_OK(cudaEventElapsedTime(&ms, pr.first, pr.second));
            c->ms_acc[k] += ms;
            c->n_acc[k] += 1;
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
        c->ev[k].clear();
    }
    return RG_OK;
}

extern "C" int rg_set_option(rg_ctx *c, int key, double value)
{
    if (!c) return rg_set_error(RG_E_ARG, "rg_set_option: ctx is NULL");
    if (key == RG_OPT_PROFILE) {
        RG_CUDA_OK(cudaSetDevice(c->device));
        c->profile = value != 0;
        if (c->profile && !c->d_stats) {
            RG_CUDA_OK(cudaMalloc(&c->d_stats, 4 * sizeof(unsigned long long)));
            RG_CUDA_OK(cudaMemset(c->d_stats, 0, 4 * sizeof(unsigned long long)));
        }
        return RG_OK;
    }
    return rg_set_error(RG_E_ARG, "rg_set_option: unknown key %d", key);
}

extern "C" int rg_reset_stats(rg_ctx *c)
{
    if (!c) return rg_set_error(RG_E_ARG, "rg_reset_stats: ctx is NULL");
    RG_CUDA_OK(cudaSetDevice(c->device));
    RG_TRY(prof_collect(c));
    for (int k = 0; k < 5; k++) { c->ms_acc[k] = 0; c->n_acc[k] = 0; }
    c->n_sets = 0;
    if (c->d_stats) RG_CUDA_OK(cudaMemset(c->d_stats, 0, 4 * sizeof(unsigned long long)));
    return RG_OK;
}

extern "C" int rg_get_stat(rg_ctx *c, int key, double *value)
{
    if (!c || !value) return rg_set_error(RG_E_ARG, "rg_get_stat: bad arguments");
    RG_CUDA_OK(cudaSetDevice(c->device));
    RG_TRY(prof_collect(c));
    unsigned long long h[4] = {0, 0, 0, 0};
    if (c->d_stats && key >= RG_STAT_SUM_W && key <= RG_STAT_N_SYS) {
        RG_CUDA_OK(cudaDeviceSynchronize());
        RG_CUDA_OK(cudaMemcpy(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost));
    }
    switch (key) {
    case RG_STAT_MS_SETUP: *value = c->ms_acc[0]; break;
    case RG_STAT_MS_NTHCOMP: *value = c->ms_acc[1]; break;
    case RG_STAT_MS_SPECTRUM: *value = c->ms_acc[2]; break;
    case RG_STAT_N_SETUP: *value = (double)c->n_acc[0]; break;
    case RG_STAT_N_NTHCOMP: *value = (double)c->n_acc[1]; break;
    case RG_STAT_N_SPECTRUM: *value = (double)c->n_acc[2]; break;
    case RG_STAT_SUM_W: *value = (double)h[0]; break;
    case RG_STAT_N_CONV_ANN: *value = (double)h[1]; break;
    case RG_STAT_SUM_JMAX: *value = (double)h[2]; break;
    case RG_STAT_N_SYS: *value = (double)h[3]; break;
    case RG_STAT_N_SETS: *value = c->n_sets; break;
    case RG_STAT_MS_OBSERVE: *value = c->ms_acc[3]; break;
    case RG_STAT_N_OBSERVE: *value = (double)c->n_acc[3]; break;
    case RG_STAT_MS_HIST: *value = c->ms_acc[4]; break;
    case RG_STAT_N_HIST: *value = (double)c->n_acc[4]; break;
    default: return rg_set_error(RG_E_ARG, "rg_get_stat: unknown key %d", key);
    }
    return RG_OK;
}

// ---------------------------------------------------------------------------
// energy grid (relagn.py:335-346 / new_ear :780-805)
extern "C" int rg_set_grid(rg_ctx *c, const double *eb, int nE)
{
    if (!c || !eb || nE < 2) return rg_set_error(RG_E_ARG, "rg_set_grid: bad arguments");
    if (rg_spectrum_pick_NC(nE) == 0) return rg_set_error(RG_E_GRID, "nE = %d exceeds the kernel capacity", nE);
    for (int j = 0; j < nE; j++)
        if (!(eb[j + 1] > eb[j]) || !(eb[j] > 0)) return rg_set_error(RG_E_ARG, "bin edges must be positive and ascending");
    RG_CUDA_OK(cudaSetDevice(c->device));
    const int NA = 11; // arrays of nE
    std::vector<double> h((size_t)NA * nE + 3 * RG_NEXT, 0.0);
    double *Egrid = &h[0], *dE = &h[(size_t)nE], *nu = &h[(size_t)2 * nE], *hnu = &h[(size_t)3 * nE],
           *bbpre = &h[(size_t)4 * nE], *wt = &h[(size_t)5 * nE], *e511 = &h[(size_t)6 * nE], *ear2 = &h[(size_t)7 * nE],
           *lgE = &h[(size_t)8 * nE], *kph = &h[(size_t)9 * nE], *rear2 = &h[(size_t)10 * nE];
    double *x_hnu = &h[(size_t)NA * nE], *x_bbpre = x_hnu + RG_NEXT, *x_wt = x_bbpre + RG_NEXT;
    for (int j = 0; j < nE; j++) {
        dE[j] = eb[j + 1] - eb[j];
        Egrid[j] = eb[j] + 0.5 * dE[j];
        nu[j] = Egrid[j] * RGK_KEV / RGK_H;
        hnu[j] = RGK_H * nu[j];
        bbpre[j] = (2 * RGK_H * (nu[j] * nu[j] * nu[j])) / (RGK_C * RGK_C);
        e511[j] = Egrid[j] / 511.0;
        ear2[j] = Egrid[j] * Egrid[j];
        lgE[j] = log10(Egrid[j]);
        kph[j] = dE[j] / (RGK_H * Egrid[j]); // W/Hz -> photons/s per bin (relagn.py:509-516 times dE)
        rear2[j] = 1.0 / ear2[j];
    }
    int n_ext = Egrid[0] < 1e-4 ? RG_NEXT : 0; // relagn.py:868
    std::vector<double> xs; // abscissae of np.trapz: [extension, nu]
    if (n_ext) {
        double numin = nu[0];
        for (int j = 1; j < nE; j++) numin = std::min(numin, nu[j]);
        double start = 1e12, stop = numin - 100; // np.geomspace(1e12, min(nu)-100, 50), relagn.py:869-871
        double ls = log10(start), le = log10(stop), step = (le - ls) / (RG_NEXT - 1);
        for (int i = 0; i < RG_NEXT; i++) xs.push_back(pow(10.0, i == RG_NEXT - 1 ? le : i * step + ls));
        xs[0] = start; xs[RG_NEXT - 1] = stop;
    }
    for (int j = 0; j < nE; j++) xs.push_back(nu[j]);
    int n = (int)xs.size();
    for (int k = 0; k < n; k++) {
        double w = 0;
        if (k > 0) w += 0.5 * (xs[k] - xs[k - 1]);
        if (k + 1 < n) w += 0.5 * (xs[k + 1] - xs[k]);
        if (k < n_ext) {
            x_wt[k] = w;
            x_hnu[k] = RGK_H * xs[k];
            x_bbpre[k] = (2 * RGK_H * (xs[k] * xs[k] * xs[k])) / (RGK_C * RGK_C);
        } else wt[k - n_ext] = w;
    }
    // the hot and warm normalisations integrate over nu_grid only (relagn.py:933,1316): when the
    // extension is on, the disc joint weight differs from theirs in bin 0.  Keep a second copy.
    // (handled by storing the plain weight in x_wt[RG_NEXT .. ] -- not needed: bin 0 of donthcomp is 0.)
    double dlnE = log(eb[nE] / eb[0]) / nE;
    int geometric = 1;
    for (int j = 0; j <= nE; j++)
        if (fabs(log(eb[j] / eb[0]) - j * dlnE) > 1e-4 * dlnE) { geometric = 0; break; }
    free_dev(c->d_grid);
    c->d_grid = nullptr;
    RG_CUDA_OK(cudaMalloc(&c->d_grid, h.size() * sizeof(double)));
    RG_CUDA_OK(cudaMemcpy(c->d_grid, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    GridDesc &g = c->gd;
    double *d = c->d_grid;
    g.Egrid = d; g.dE = d + (size_t)nE; g.nu = d + (size_t)2 * nE; g.hnu = d + (size_t)3 * nE;
    g.bbpre = d + (size_t)4 * nE; g.wt = d + (size_t)5 * nE; g.e511 = d + (size_t)6 * nE; g.ear2 = d + (size_t)7 * nE;
    g.lgE = d + (size_t)8 * nE;
    c->d_kph = d + (size_t)9 * nE;
    g.rear2 = d + (size_t)10 * nE;
    c->pos0_dirty = true;
    g.x_hnu = d + (size_t)NA * nE; g.x_bbpre = g.x_hnu + RG_NEXT; g.x_wt = g.x_bbpre + RG_NEXT;
    g.n_ext = n_ext; g.nE = nE; g.geometric = geometric; g.dlnE = dlnE;
    c->nE = nE;
    c->h_ebins.assign(eb, eb + nE + 1);
    free_dev(c->d_ebins);
    c->d_ebins = nullptr;
    RG_CUDA_OK(cudaMalloc(&c->d_ebins, (size_t)(nE + 1) * sizeof(double)));
    RG_CUDA_OK(cudaMemcpy(c->d_ebins, eb, (size_t)(nE + 1) * sizeof(double), cudaMemcpyHostToDevice));
    return RG_OK;
}

// ---------------------------------------------------------------------------
// tables
static int tables_alloc(rg_ctx *c, const double *spins, int n_spin, const double *thetas, int n_inc, int NR, int NPHI,
                        double r_in)
{
    if (!c || !spins || !thetas || n_spin < 1 || n_inc < 1 || NR < 2 || !(r_in > 1.0) || !(r_in < RG_TAB_ROUT))
        return rg_set_error(RG_E_ARG, "tables: bad arguments");
    if (NPHI < 32 || NPHI > RG_SPEC_THREADS || RG_SPEC_THREADS % NPHI != 0)
        return rg_set_error(RG_E_ARG, "tables: NPHI must be 32, 64, 128 or 256");
    if ((size_t)NR * NPHI * sizeof(double2) > 160 * 1024)
        return rg_set_error(RG_E_ARG, "tables: NR x NPHI = %d x %d exceeds the shared-memory slice (10240 samples)", NR, NPHI);
    for (int i = 1; i < n_spin; i++)
        if (!(spins[i] > spins[i - 1])) return rg_set_error(RG_E_ARG, "tables: spins must ascend");
    for (int i = 1; i < n_inc; i++)
        if (!(thetas[i] > thetas[i - 1])) return rg_set_error(RG_E_ARG, "tables: inclinations must ascend");
    RG_CUDA_OK(cudaSetDevice(c->device));
    free_dev(c->d_tab); free_dev(c->d_tab2); free_dev(c->d_tabmeta); free_dev(c->d_firstrow);
    c->d_tab = nullptr; c->d_tab2 = nullptr; c->d_tabmeta = nullptr; c->d_firstrow = nullptr; c->have_tab = false;
    c->tab_floats = (size_t)n_spin * n_inc * 2 * NR * NPHI;
    RG_CUDA_OK(cudaMalloc(&c->d_tab, c->tab_floats * sizeof(float)));
    std::vector<double> meta(n_spin + n_inc);
    for (int i = 0; i < n_spin; i++) meta[i] = spins[i];
    for (int i = 0; i < n_inc; i++) meta[n_spin + i] = thetas[i];
    RG_CUDA_OK(cudaMalloc(&c->d_tabmeta, meta.size() * sizeof(double)));
    RG_CUDA_OK(cudaMemcpy(c->d_tabmeta, meta.data(), meta.size() * sizeof(double), cudaMemcpyHostToDevice));
    // first row with data per spin node: rows inside 1.02 x photon orbit are empty (kerr_oracle.c: ro_tables_finish)
    std::vector<int> first(n_spin);
    for (int m = 0; m < n_spin; m++) {
        const double rv = rg_r_valid(spins[m]);
        int f = 0;
        while (f < NR && !(rg_row_radius(r_in, NR, f) > rv)) f++;
        first[m] = f;
    }
    RG_CUDA_OK(cudaMalloc(&c->d_firstrow, n_spin * sizeof(int)));
    RG_CUDA_OK(cudaMemcpy(c->d_firstrow, first.data(), n_spin * sizeof(int), cudaMemcpyHostToDevice));
    TabDesc &t = c->td;
    t.data = c->d_tab; t.spins = c->d_tabmeta; t.thetas = c->d_tabmeta + n_spin; t.first_row = c->d_firstrow;
    t.n_spin = n_spin; t.n_inc = n_inc; t.NR = NR; t.NPHI = NPHI; t.r_in = r_in;
    t.x0 = rg_row_x(r_in);
    t.dx = (rg_row_x(RG_TAB_ROUT) - t.x0) / (NR - 1);
    c->h_spins.assign(spins, spins + n_spin);
    c->h_thetas.assign(thetas, thetas + n_inc);
    c->h_firstrow = first;
    if (!c->d_kystatus) RG_CUDA_OK(cudaMalloc(&c->d_kystatus, sizeof(int)));
    return RG_OK;
}

static int ensure_slices(rg_ctx *c, long ctas);

// (dE_0 4 / E_0) (E_0 / dE_0) with the reference's operations (relagn.py:970-974, 1003-1006) on the first bin
static double pack_unpack_const(const std::vector<double> &eb)
{
    const double dE = eb[1] - eb[0], E = eb[0] + dE / 2;
    return ((dE * 4) / E) * (E / dE);
}

static double host_row_pos(const TabDesc &t, double r)
{
    double ell = (rg_row_x(r) - t.x0) / t.dx;
    if (!(ell > 0)) ell = 0;
    if (ell > t.NR - 1) ell = t.NR - 1;
    return ell;
}

// (ln g, Jn) planes for the kernels (ln g through the host's log(), the oracle's: the float rounding of the planes
// must not depend on whose libm computed them), validity of the spin stencils, and the range of ln g over the
// usable rows, which bounds the shift histogram
static int tables_finish(rg_ctx *c, const float *host_data)
{
    std::vector<float> tmp;
    if (!host_data) {
        tmp.resize(c->tab_floats);
        RG_CUDA_OK(cudaMemcpy(tmp.data(), c->d_tab, c->tab_floats * sizeof(float), cudaMemcpyDeviceToHost));
        host_data = tmp.data();
    }
    const TabDesc &t = c->td;
    const int NR = t.NR, NP = t.NPHI, n_spin = t.n_spin, n_inc = t.n_inc;
    size_t plane = (size_t)NR * NP;
    size_t nodes = (size_t)n_spin * n_inc;
    // every spin interval needs its two own nodes down to the innermost row a model of that interval uses
    for (int m = 0; m + 1 < n_spin || m == 0; m++) {
        const int top = m + 1 < n_spin ? m + 1 : m;
        int need = (int)floor(host_row_pos(t, rg_risco(c->h_spins[top])));
        if (need > NR - 2) need = NR - 2;
        if (c->h_firstrow[m] > need || c->h_firstrow[top] > need)
            return rg_set_error(RG_E_ARG, "tables: spin nodes %g and %g are too far apart: node %g has no circular orbit at "
                                          "the innermost radius of a model with spin %g", c->h_spins[m], c->h_spins[top],
                                c->h_spins[m], c->h_spins[top]);
        if (n_spin == 1) break;
    }
    double gmin = 1e300, gmax = 0;
    std::vector<float2> pairs(nodes * plane);
    for (int m = 0; m < n_spin; m++) {
        // rows of node m that can enter a slice: down to the innermost row of the fastest model it is blended into
        const int mt = m + 2 < n_spin ? m + 2 : n_spin - 1;
        int use = (int)floor(host_row_pos(t, rg_risco(c->h_spins[mt])));
        if (use < c->h_firstrow[m]) use = c->h_firstrow[m];
        for (int n = 0; n < n_inc; n++) {
            const size_t node = (size_t)m * n_inc + n;
            const float *g = host_data + node * 2 * plane, *jn = g + plane;
            for (int s = 0; s < NR; s++)
                for (int p = 0; p < NP; p++) {
                    const size_t i = (size_t)s * NP + p;
                    const double v = g[i];
                    if (s >= c->h_firstrow[m] && (!(v > 0) || !(v < 1e30)))
                        return rg_set_error(RG_E_ARG, "tables: non-positive or non-finite g in node %zu row %d", node, s);
                    pairs[node * plane + i].x = v > 0 ? (float)log(v) : 0.0f;
                    pairs[node * plane + i].y = jn[i];
                    if (s >= use) { gmin = std::min(gmin, v); gmax = std::max(gmax, v); }
                }
        }
    }
    RG_CUDA_OK(cudaMalloc(&c->d_tab2, pairs.size() * sizeof(float2)));
    RG_CUDA_OK(cudaMemcpy(c->d_tab2, pairs.data(), pairs.size() * sizeof(float2), cudaMemcpyHostToDevice));
    c->td.pairs = c->d_tab2;
    c->td.lng_min = log(gmin);
    c->td.lng_max = log(gmax);
    c->have_tab = true;
    return RG_OK;
}

extern "C" int rg_tables_upload(rg_ctx *c, const double *spins, int n_spin, const double *thetas, int n_inc, int NR,
                                int NPHI, double r_in, const float *data)
{
    if (!data) return rg_set_error(RG_E_ARG, "rg_tables_upload: data is NULL");
    RG_TRY(tables_alloc(c, spins, n_spin, thetas, n_inc, NR, NPHI, r_in));
    RG_CUDA_OK(cudaMemcpy(c->d_tab, data, c->tab_floats * sizeof(float), cudaMemcpyHostToDevice));
    return tables_finish(c, data);
}

extern "C" int rg_tables_generate(rg_ctx *c, const double *spins, int n_spin, const double *thetas, int n_inc, int NR,
                                  int NPHI, double r_in, long *n_unresolved)
{
    RG_TRY(tables_alloc(c, spins, n_spin, thetas, n_inc, NR, NPHI, r_in));
    long bad = 0;
    RG_TRY(rg_launch_raytrace(spins, n_spin, thetas, n_inc, NR, NPHI, r_in, c->tab_limb, c->d_tab, &bad, c->stream));
    c->launches += 2L * n_spin * n_inc;
    if (n_unresolved) *n_unresolved = bad;
    return tables_finish(c, nullptr);
}

// kyconv parameter 10 (relagn.py:988): the law multiplies the weight of every table sample, so it belongs to the table
// set.  Call before rg_tables_generate (the ray tracer bakes it in) or after rg_tables_upload (states what the
// uploaded set was built with, e.g. by relagn_b200.ky_fits.import_ky_tables(limb=...)).
extern "C" int rg_tables_set_limb(rg_ctx *c, int law)
{
    if (!c || (law != 0 && law != -1 && law != -2))
        return rg_set_error(RG_E_ARG, "rg_tables_set_limb: law must be 0 (isotropic), -1 (Laor) or -2 (Haardt)");
    c->tab_limb = law;
    return RG_OK;
}

extern "C" int rg_tables_download(rg_ctx *c, float *data, size_t n_floats)
{
    if (!c || !c->have_tab) return rg_set_error(RG_E_NOTABLES, "no tables resident");
    if (!data || n_floats != c->tab_floats) return rg_set_error(RG_E_ARG, "rg_tables_download: size mismatch");
    RG_CUDA_OK(cudaSetDevice(c->device));
    RG_CUDA_OK(cudaMemcpy(data, c->d_tab, n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return RG_OK;
}

extern "C" int rg_tables_info(rg_ctx *c, int *n_spin, int *n_inc, int *NR, int *NPHI, double *r_in, size_t *n_floats)
{
    if (!c || !c->have_tab) return rg_set_error(RG_E_NOTABLES, "no tables resident");
    if (n_spin) *n_spin = c->td.n_spin;
    if (n_inc) *n_inc = c->td.n_inc;
    if (NR) *NR = c->td.NR;
    if (NPHI) *NPHI = c->td.NPHI;
    if (r_in) *r_in = c->td.r_in;
    if (n_floats) *n_floats = c->tab_floats;
    return RG_OK;
}

// ---------------------------------------------------------------------------
// B1: kyconv seam
static void ky_range(double dlnE, double zshift, int &K_LO, int &K_HI)
{
    K_LO = (int)floor(log(0.005) / dlnE) - 2;
    K_HI = (int)ceil(log(4.0) / dlnE) + 2;
    double zs = 0;
    if (zshift != 0) zs = -log(1 + zshift) / dlnE;
    K_LO += (int)floor(zs) - 1;
    K_HI += (int)ceil(zs) + 1;
}

static int ky_hist_host(rg_ctx *c, double a, double th, double r1, double r2, double alpha, double beta, double rb,
                        double zshift, double dlnE, int nsub, std::vector<double> &H, int &K_LO)
{
    if (!c->have_tab) return rg_set_error(RG_E_NOTABLES, "kyconv: no Kerr tables resident (rg_tables_generate/upload)");
    if (!(r2 > r1) || !(dlnE > 0)) return rg_set_error(RG_E_ARG, "kyconv: need r_out > r_in and a positive grid step");
    int K_HI;
    ky_range(dlnE, zshift, K_LO, K_HI);
    int nH = K_HI - K_LO + 1;
    RG_CUDA_OK(cudaSetDevice(c->device));
    if (nH > c->kyH_cap) {
        free_dev(c->d_kyH);
        c->d_kyH = nullptr;
        RG_CUDA_OK(cudaMalloc(&c->d_kyH, (size_t)nH * sizeof(double)));
        c->kyH_cap = nH;
    }
    RG_TRY(ensure_slices(c, 1));
    RG_TRY(rg_launch_ky_hist(c->td, a, th, r1, r2, alpha, beta, rb, zshift, dlnE, nsub, c->d_kyH, K_LO, nH, c->d_kystatus,
                             c->d_slices, c->stream));
    c->launches++;
    H.resize(nH);
    int kst = 0;
    RG_CUDA_OK(cudaMemcpyAsync(H.data(), c->d_kyH, (size_t)nH * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RG_CUDA_OK(cudaMemcpyAsync(&kst, c->d_kystatus, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    RG_CUDA_OK(cudaStreamSynchronize(c->stream));
    if (kst) return rg_set_error(RG_E_ARG, "kyconv: the resident tables have no valid spin stencil for a = %g", a);
    return RG_OK;
}

extern "C" int rg_ky_kernel(rg_ctx *c, double a, double theta_o, double r1, double r2, double alpha, double beta,
                            double rb, double zshift, double dlnE, double *H_host, int cap, int *kmin, int *nk)
{
    if (!c || !H_host || !kmin || !nk) return rg_set_error(RG_E_ARG, "rg_ky_kernel: bad arguments");
    std::vector<double> H;
    int K_LO;
    RG_TRY(ky_hist_host(c, a, theta_o, r1, r2, alpha, beta, rb, zshift, dlnE, 0, H, K_LO));
    int lo = 0, hi = (int)H.size() - 1;
    while (lo <= hi && H[lo] == 0) lo++;
    while (hi >= lo && H[hi] == 0) hi--;
    if (lo > hi) { *kmin = 0; *nk = 0; return RG_OK; }
    int n = hi - lo + 1;
    *nk = n;
    *kmin = K_LO + lo;
    if (n > cap) return rg_set_error(RG_E_ARG, "rg_ky_kernel: histogram needs %d entries, cap is %d", n, cap);
    memcpy(H_host, &H[lo], (size_t)n * sizeof(double));
    return RG_OK;
}

extern "C" int rg_kyconv(rg_ctx *c, const double *eb, int nE, const double *par, double *flux)
{
    if (!c || !eb || !par || !flux || nE < 2) return rg_set_error(RG_E_ARG, "rg_kyconv: bad arguments");
    if (par[3] != 0 || par[11] != -1)
        return rg_set_error(RG_E_UNSUPPORTED, "rg_kyconv: only ms=0 and norm=-1 are implemented (what relagn.py:997-998 passes)");
    if (par[9] != (double)c->tab_limb)
        return rg_set_error(RG_E_UNSUPPORTED, "rg_kyconv: limb-darkening law %g requested, the resident tables were built "
                                              "for law %d (rg_tables_set_limb before rg_tables_generate / after upload)",
                            par[9], c->tab_limb);
    double dlnE = log(eb[nE] / eb[0]) / nE;
    for (int j = 0; j <= nE; j++)
        if (fabs(log(eb[j] / eb[0]) - j * dlnE) > 1e-4 * dlnE)
            return rg_set_error(RG_E_GRID, "rg_kyconv: the energy grid must be geometric");
    std::vector<double> H;
    int K_LO;
    double th = par[1] * (RGK_PI / 180.0);
    RG_TRY(ky_hist_host(c, par[0], th, par[2], par[4], par[5], par[6], par[7], par[8], dlnE, 0, H, K_LO));
    int lo = 0, hi = (int)H.size() - 1;
    while (lo <= hi && H[lo] == 0) lo++;
    while (hi >= lo && H[hi] == 0) hi--;
    if (lo > hi) { memset(flux, 0, (size_t)nE * sizeof(double)); return RG_OK; }
    if (nE > c->ky_cap) {
        free_dev(c->d_kyin); free_dev(c->d_kyout);
        c->d_kyin = c->d_kyout = nullptr;
        RG_CUDA_OK(cudaMalloc(&c->d_kyin, (size_t)nE * sizeof(double)));
        RG_CUDA_OK(cudaMalloc(&c->d_kyout, (size_t)nE * sizeof(double)));
        c->ky_cap = nE;
    }
    RG_CUDA_OK(cudaMemcpyAsync(c->d_kyin, flux, (size_t)nE * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RG_TRY(rg_launch_ky_conv(c->d_kyH + lo, K_LO + lo, hi - lo + 1, c->d_kyin, c->d_kyout, nE, c->stream));
    c->launches++;
    RG_CUDA_OK(cudaMemcpyAsync(flux, c->d_kyout, (size_t)nE * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RG_CUDA_OK(cudaStreamSynchronize(c->stream));
    return RG_OK;
}

// ---------------------------------------------------------------------------
// B2: batched evaluation
static int ensure_dev(double **p, size_t *have, size_t need);

// Chunk of sets one setup / nthcomp / hist / spectrum launch sequence covers.  Larger chunks amortise the host
// round trip after the setup kernel and the launch tails (65 536 instead of 16 384: +4.7 % on C4); the arenas scale
// with the chunk (65 536 sets at nE = 1000: 19 GB of Kompaneets solutions, 21 GB of Comptonised spectra, <= 15 GB of
// shift histograms -- sized for the 180 GB of a B200), so small batches get small arenas.
#define RG_CHUNK_MAX 65536
static int pick_chunk_sets(int P)
{
    int cap = RG_CHUNK_MAX;
    if (const char *e = getenv("RG_CHUNK_SETS")) { // tuning knob: upper bound of the chunk
        int v = atoi(e);
        if (v >= 1024 && v <= 131072) cap = v;
    }
    int want = 64; // a single model object (P = 1) gets arenas of tens of MB, not the GBs of a 4096-set chunk
    while (want < P && want < cap) want *= 4;
    return std::min(want, cap);
}

// scratch for the table slices of the resident hist_kernel CTAs (two halves each): 8 per SM at most
static int ensure_slices(rg_ctx *c, long ctas = 1L << 30)
{
    const size_t per = (size_t)c->td.NR * c->td.NPHI;
    const int slots = (int)std::min<long>(c->sm_count * 8L, std::max<long>(2 * ctas, 2));
    if (c->d_slices && c->slice_elems >= per * slots) { c->slice_slots = slots; return RG_OK; }
    free_dev(c->d_slices);
    c->d_slices = nullptr;
    RG_CUDA_OK(cudaMalloc(&c->d_slices, per * slots * sizeof(double2)));
    c->slice_elems = per * slots;
    c->slice_slots = slots;
    return RG_OK;
}

static int ensure_batch_scratch(rg_ctx *c, int model, int P)
{
    if (!c->d_queue) RG_CUDA_OK(cudaMalloc(&c->d_queue, 4 * sizeof(int)));
    const int want = pick_chunk_sets(P);
    {   // threads of the solve kernel's grid: sized by the machine, or by the chunk when that is smaller; and the global
        // scratch of the small-batch Kompaneets kernel / the wide-grid interpolation tables (64 MB at most)
        const int full = c->sm_count * RG_NTH_MINB * 128;
        const int need = (int)std::min<long>(full, (((long)std::max(want, c->chunk_sets) * 40 + 127) / 128) * 128);
        if (need > c->nth_threads) {
            free_dev(c->d_sc);
            c->d_sc = nullptr;
            c->sc_doubles = std::min<size_t>(((size_t)2 * RG_NTH + 1) * need, (size_t)8 << 20);
            RG_CUDA_OK(cudaMalloc(&c->d_sc, c->sc_doubles * sizeof(double)));
            c->nth_threads = need;
        }
    }
    if (want > c->chunk_sets) { // grow: per-set records, arena of Kompaneets systems
        free_dev(c->d_recs); free_dev(c->d_syslist); free_dev(c->d_sptot); free_dev(c->d_header); free_dev(c->d_rq);
        free_dev(c->d_farr);
        c->d_recs = nullptr; c->d_syslist = nullptr; c->d_sptot = c->d_header = c->d_farr = c->d_rq = nullptr;
        c->farr_nE = 0; c->chunk_sets = 0;
        c->sys_cap = want * 40;
        RG_CUDA_OK(cudaMalloc(&c->d_recs, (size_t)want * sizeof(SetRec)));
        RG_CUDA_OK(cudaMalloc(&c->d_syslist, (size_t)c->sys_cap * sizeof(int2)));
        // (the coefficient arena of the large-batch Kompaneets kernels is allocated when they first run: ensure_coef)
        RG_CUDA_OK(cudaMalloc(&c->d_header, (size_t)c->sys_cap * 4 * sizeof(double)));
        c->chunk_sets = want;
    }
    if (c->farr_nE != c->nE) { // Comptonised spectra on the grid: [farr_cap][nE], 6 GB per 16 384 sets of the chunk
        free_dev(c->d_farr);
        c->d_farr = nullptr; c->farr_nE = 0;
        size_t cap = (size_t)(6e9 * (c->chunk_sets / 16384.0)) / ((size_t)c->nE * sizeof(double));
        c->farr_cap = (int)std::min<size_t>((size_t)c->sys_cap, std::max<size_t>(cap, 4096));
        RG_CUDA_OK(cudaMalloc(&c->d_farr, (size_t)c->farr_cap * c->nE * sizeof(double)));
        c->farr_nE = c->nE;
    }
    if (model == RG_MODEL_RELQSO && !c->d_rq) {
        c->rq_cap = 1; // r_hot per set, filled by relqso_rhot_kernel
        RG_CUDA_OK(cudaMalloc(&c->d_rq, (size_t)c->chunk_sets * sizeof(double)));
    }
    return RG_OK;
}

// nosync: the small-batch path.  The arenas are sized from caps (every row of the small system arena is scanned, the
// histogram arena holds RG_SMALL_ANN annuli per set) instead of the setup kernel's counters, so no host round trip
// separates the setup kernel from the rest of the chain; sets that do not fit are flagged RG_S_CAP by the setup kernel
// and the caller re-runs the batch through the counting path (the counters are still written, *overflow reports it).
#define RG_SMALL_SETS 64
static int eval_batch_impl(rg_ctx *c, const double *d_params, int P, int model, double dr_dex, unsigned flags,
                           double *d_out, double *d_attrs, void *stream, int piece_sets, rg_piece_fn on_piece, void *user,
                           bool nosync = false, bool shrink = false)
{
    if (!c || !d_params || !d_out || P < 0) return rg_set_error(RG_E_ARG, "rg_eval_batch: bad arguments");
    if (model != RG_MODEL_RELAGN && model != RG_MODEL_RELQSO) return rg_set_error(RG_E_ARG, "unknown model %d", model);
    if (!(dr_dex > 0)) return rg_set_error(RG_E_ARG, "dr_dex must be positive");
    if (c->nE <= 0) return rg_set_error(RG_E_GRID, "no energy grid set (rg_set_grid)");
    const bool rel = (flags & RG_FLAG_REL) != 0;
    if (rel && !c->have_tab) return rg_set_error(RG_E_NOTABLES, "relativistic evaluation needs Kerr tables");
    if (rel && !c->gd.geometric)
        return rg_set_error(RG_E_GRID, "relativistic evaluation needs a geometric (log-spaced) energy grid");
    if (P == 0) return RG_OK;
    RG_CUDA_OK(cudaSetDevice(c->device));
    RG_TRY(ensure_batch_scratch(c, model, P));
    cudaStream_t st = (cudaStream_t)stream;
    const int nE = c->nE;
    const int NC = rg_spectrum_pick_NC(nE);
    const int ncomp = (flags & RG_FLAG_COMPONENTS) ? 4 : 1;

    SpecArgs A;
    memset(&A, 0, sizeof A);
    A.grid = c->gd;
    A.flags = flags;
    A.farr = c->d_farr;
    A.queue = c->d_queue;
    A.pkupk = c->h_ebins.size() > 1 ? pack_unpack_const(c->h_ebins) : 4.0;
    if (rel) {
        A.tab = c->td;
        int kl, kh;
        ky_range(c->gd.dlnE, 0.0, kl, kh);
        int K_LO = (int)floor(c->td.lng_min / c->gd.dlnE) - 3, K_HI = (int)ceil(c->td.lng_max / c->gd.dlnE) + 3;
        K_LO = std::max(K_LO, kl);
        K_HI = std::min(K_HI, kh);
        A.K_LO = K_LO;
        A.nH = K_HI - K_LO + 1;
        A.HL = std::max(K_HI, 0) / 8 + 3;
        A.HH = std::max(-K_LO, 0) / 8 + 3;
        // row stride HL + columns + HH = 2 (mod 16) doubles: the Planck stage stores bin 32 g + lane at row lane mod 8,
        // column 4 g + lane div 8 -- with this stride the 16 lanes of a half warp hit 16 different bank pairs
        A.HH += ((2 - (A.HL + A.HH)) % 16 + 16) % 16;
        A.hstride = (A.nH + 2) & ~1; // slot 0 = (first shift, count), then up to nH bins
    } else {
        A.K_LO = 0; A.nH = 1; A.HL = 1; A.HH = 1;
    }
    A.warps = rg_spectrum_pick_warps(NC, A.nH, A.HL, A.HH, (flags & RG_FLAG_FP32) ? 1 : 0);
    if (const char *e = getenv("RG_SPEC_WARPS")) { // tuning knob: cap the warps per CTA
        int w = atoi(e);
        if (w >= 1 && w < A.warps) A.warps = w;
    }
    if (A.warps < 1) return rg_set_error(RG_E_GRID, "energy grid too fine for the shared-memory working set");
    // small batches: split every set's annuli over several CTAs (partials are folded in block order)
    {
        const bool comps = ncomp == 4;
        int NB = 1;
        if ((long)P < 2L * c->sm_count) NB = (int)std::min<long>(16, (2L * c->sm_count + P - 1) / P);
        A.NB = NB;
        if (comps || NB > 1) {
            // component mode keeps one partial per (block, warp) and region; total-only mode one per block
            size_t per_set = comps ? (size_t)NB * A.warps * 3 * nE : (size_t)NB * nE;
            size_t need = (size_t)std::min(P, c->chunk_sets) * per_set;
            if (need > c->partial_n) {
                free_dev(c->d_partial);
                c->d_partial = nullptr; c->partial_n = 0;
                RG_CUDA_OK(cudaMalloc(&c->d_partial, need * sizeof(double)));
                c->partial_n = need;
            }
            A.partial = c->d_partial;
        }
    }

    int p0 = 0, chunk = c->chunk_sets;
    while (p0 < P) {
        int n = std::min(chunk, P - p0);
        const double *par = d_params + (size_t)p0 * RG_NPAR;
        double *attrs = d_attrs ? d_attrs + (size_t)p0 * RG_NATTR : nullptr;
        // annuli the chunk may hold: unbounded on the counting path, a cap per set on the small-batch path
        const int ann_per_set = (int)std::min(100000.0, dr_dex * 8.0 + 3.0);
        const int ann_cap = nosync ? n * ann_per_set : 0x7fffffff;
        if (nosync) // unreserved rows of the system arena are recognised by a negative set index
            RG_CUDA_OK(cudaMemsetAsync(c->d_syslist, 0xff, (size_t)c->farr_cap * sizeof(int2), st));
        RG_TRY(prof_begin(c, 0, st));
        RG_TRY(rg_launch_setup(par, n, model, dr_dex, c->d_recs, c->d_counter, c->d_syslist, c->farr_cap, c->d_rq,
                               c->rq_cap, attrs, ann_cap, st));
        RG_TRY(prof_end(c, 0, st));
        c->launches += model == RG_MODEL_RELQSO ? 2 : 1;
        RG_TRY(rg_launch_counters_out(c->d_counter, c->h_counter, st)); // zero-copy store, not a D2H copy
        if (!nosync) RG_CUDA_OK(cudaStreamSynchronize(st));
        // (small-batch path: every row of the arena but the top n is scanned for warm systems, the top n for hot shapes)
        const int n_warm_sys = nosync ? c->farr_cap - n : c->h_counter[0], n_hot_sys = nosync ? n : c->h_counter[2];
        int nsys = nosync ? 1 : n_warm_sys + n_hot_sys;
        const int n_ann_rows = nosync ? ann_cap : c->h_counter[1];
        const size_t hist_need = rel ? (size_t)n_ann_rows * (size_t)A.hstride : 0;
        // (the histogram arena is allocated on demand: 8 GB per 16 384 sets of the chunk at most, 2 GB for the small chunks)
        if (!nosync && (nsys > c->farr_cap || hist_need * sizeof(double) > (size_t)std::max(2e9, 8e9 * (c->chunk_sets / 16384.0)))) {
            if (n == 1) return rg_set_error(RG_E_NOMEM, "one parameter set needs %d Kompaneets systems (cap %d) / %zu histogram doubles", nsys, c->farr_cap, hist_need);
            chunk = std::max(1, n / 2); // arenas too small for this chunk: retry with fewer sets
            continue;
        }
        // The shift histograms (hist_kernel) and the Kompaneets solves (nthcomp kernels) both depend on the setup
        // kernel only and are both latency bound: they run side by side, the histograms on the context's side stream
        // (with profiling on they are serialised on the caller's stream so that the event brackets mean something).
        // radial bin edges of every annulus of the chunk, for hist_kernel and spectrum_kernel
        RG_TRY(ensure_dev(&c->d_edges, &c->edges_n, (size_t)2 * std::max(n_ann_rows, 1)));
        RG_TRY(rg_launch_edges(c->d_recs, n, reinterpret_cast<double2 *>(c->d_edges), st));
        c->launches++;
        A.edges = reinterpret_cast<const double2 *>(c->d_edges);
        bool forked = false;
        unsigned long long *stats = c->profile ? c->d_stats : nullptr;
        int nth_run = c->nth_threads; // threads of the persistent nthcomp grid
        if (const char *e = getenv("RG_NTH_CTAS")) { // tuning knob: resident nthcomp CTAs per SM (<= RG_NTH_MINB)
            int k = atoi(e);
            if (k >= 1 && k <= RG_NTH_MINB) nth_run = c->sm_count * k * 128;
        }
        // Which of the two gets the SMs first is decided by the launch order.  hist_kernel first: its CTAs take every
        // SM and the persistent nthcomp grid moves in as they drain (144.8 ms per 125 000 sets); nthcomp first: the
        // two share the SMs for the whole solve and slow each other down (155.2 ms; one stream, no overlap: 146.8 ms).
        // With the larger shared-memory footprint hist_kernel had before r1r the order was the other way round by
        // 2 % -- measure again when either kernel's resources change (RG_NTH_FIRST=1: tuning knob for the other order)
        const bool nth_first = getenv("RG_NTH_FIRST") != nullptr;
        auto launch_nth = [&]() -> int {
            if (nsys <= 0) return RG_OK;
            if (nosync) { // small batches: the latency variant, one warp per system
                RG_TRY(rg_launch_nthcomp_small(c->d_recs, c->d_syslist, n_warm_sys, n_hot_sys, c->farr_cap, c->d_p10,
                                               reinterpret_cast<const double2 *>(c->d_edges), c->d_sc,
                                               c->sc_doubles, c->sm_count, c->gd, c->d_farr,
                                               c->d_queue + 3, stats, st));
                c->launches++;
                return RG_OK;
            }
            c->launches++; // nthcomp_interp_kernel
            if (!c->d_sptot) // (G', M) of every row of every system: what the forward sweep leaves for the back substitution
                RG_CUDA_OK(cudaMalloc(&c->d_sptot, (size_t)c->sys_cap * RG_SPT_STRIDE * sizeof(double2)));
            RG_TRY(prof_begin(c, 1, st));
            RG_TRY(rg_launch_nthcomp(c->d_recs, c->d_syslist, n_warm_sys, n_hot_sys, c->farr_cap, c->d_p10, nth_run,
                                     reinterpret_cast<double2 *>(c->d_sptot), c->d_header, c->sm_count, c->gd,
                                     log10(c->h_ebins[0] + 0.5 * (c->h_ebins[1] - c->h_ebins[0])),
                                     c->h_ebins[nE - 1] + 0.5 * (c->h_ebins[nE] - c->h_ebins[nE - 1]), c->d_sc,
                                     c->sc_doubles, c->d_farr, stats, st));
            RG_TRY(prof_end(c, 1, st));
            c->launches++;
            return RG_OK;
        };
        if (rel) {
            RG_TRY(ensure_dev(&c->d_hist, &c->hist_n, hist_need));
            HistArgs H;
            memset(&H, 0, sizeof H);
            H.recs = c->d_recs; H.P = n; H.tab = c->td; H.dlnE = c->gd.dlnE; H.inv_dlnE = 1.0 / c->gd.dlnE; H.K_LO = A.K_LO; H.nH = A.nH;
            H.NBH = n >= 2 * c->sm_count ? 1 : (int)std::min<long>(32, (4L * c->sm_count + n - 1) / n);
            H.hist = c->d_hist; H.hstride = A.hstride; H.edges = A.edges;
            RG_TRY(ensure_slices(c, (long)n * H.NBH));
            H.slices = c->d_slices; H.slice_slots = c->slice_slots; H.queue = c->d_queue + 2;
            if (c->sel_bytes < (size_t)n * rg_selrec_bytes()) {
                free_dev(c->d_sel);
                c->d_sel = nullptr;
                c->sel_bytes = (size_t)std::max(n, c->chunk_sets) * rg_selrec_bytes();
                RG_CUDA_OK(cudaMalloc(&c->d_sel, c->sel_bytes));
            }
            H.sel = reinterpret_cast<SelRec *>(c->d_sel);
            cudaStream_t hs = st;
            if (!c->profile && nsys > 0 && !getenv("RG_NO_SIDE_STREAM")) {
                RG_CUDA_OK(cudaEventRecord(c->ev_fork, st));
                RG_CUDA_OK(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
                hs = c->side;
                forked = true;
            }
            if (forked && nth_first) RG_TRY(launch_nth());
            RG_TRY(prof_begin(c, 4, hs));
            RG_TRY(rg_launch_hist(H, c->sm_count, hs));
            RG_TRY(prof_end(c, 4, hs));
            if (forked) RG_CUDA_OK(cudaEventRecord(c->ev_join, c->side));
            c->launches++;
            A.hist = c->d_hist;
        }
        if (!(forked && nth_first)) RG_TRY(launch_nth());
        if (forked) RG_CUDA_OK(cudaStreamWaitEvent(st, c->ev_join, 0));
        // reference annulus of every set for the Wien cut of the spectrum kernel (needs the histograms: after the join)
        RG_TRY(ensure_dev(&c->d_wien, &c->wien_n, (size_t)2 * std::max(n, c->chunk_sets)));
        RG_TRY(rg_launch_wienref(c->d_recs, n, A.edges, rel ? c->d_hist : nullptr, A.hstride, rel ? 1 : 0,
                                 reinterpret_cast<double2 *>(c->d_wien), st));
        c->launches++;
        A.wien = reinterpret_cast<const double2 *>(c->d_wien);
        A.recs = c->d_recs;
        A.P = n;
        A.out = d_out + (size_t)p0 * ncomp * nE;
        A.stats = stats;
        RG_TRY(prof_begin(c, 2, st));
        if (on_piece && piece_sets > 0 && A.NB == 1 && (shrink ? n > piece_sets / 4 : n > piece_sets + piece_sets / 2)) {
            // the chunk's spectrum kernel in sub-launches of piece_sets sets (setup, Kompaneets solves and histograms
            // stay chunk-wide): the caller starts moving a piece out while the next one is computed
            for (int q0 = 0; q0 < n;) {
                int nq = std::min(piece_sets, n - q0);
                if (shrink) {
                    // host pipeline: every piece is half of what is left of the BATCH (between piece_sets / 8 and
                    // piece_sets): few launches (each has a tail) while the batch is long, and a small last piece --
                    // its copy to the host is the only one nothing hides
                    const long left = (long)(n - q0) + (long)(P - p0 - n);
                    nq = (int)std::min<long>(std::min<long>(piece_sets, std::max<long>(piece_sets / 8, left / 2)), n - q0);
                    if (n - q0 - nq < piece_sets / 16) nq = n - q0;
                } else
                if (n - q0 - nq < piece_sets / 2) nq = n - q0; // a short remainder joins the last piece
                A.recs = c->d_recs + q0;
                A.wien = reinterpret_cast<const double2 *>(c->d_wien) + q0; // per-set arrays follow the piece
                A.P = nq;
                A.out = d_out + (size_t)(p0 + q0) * ncomp * nE;
                RG_TRY(rg_launch_spectrum(A, c->sm_count, st));
                c->launches++;
                on_piece(user, p0 + q0, nq);
                q0 += nq;
            }
        } else {
            RG_TRY(rg_launch_spectrum(A, c->sm_count, st));
            c->launches++;
            if (on_piece) on_piece(user, p0, n);
        }
        RG_TRY(prof_end(c, 2, st));
        c->n_sets += n;
        c->last_first = p0; c->last_n = n; c->last_stream = st;
        p0 += n;
    }
    return RG_OK;
}

// relagn.py:1200-1223: Fnu_seed_hot of a set of the last evaluated chunk -- the hot Comptonised shape
// donthcomp(Egrid, [gamma_h, kTe_h, kT_seed, 0, 0]) * h / keV on the grid, zeros when kT_seed >= kTe_h or the set has
// no hot region -- as the Kompaneets kernels left it in the farr arena.
extern "C" int rg_last_hot_shape(rg_ctx *c, int set, double *out_host)
{
    if (!c || !out_host) return rg_set_error(RG_E_ARG, "rg_last_hot_shape: bad arguments");
    if (set < c->last_first || set >= c->last_first + c->last_n)
        return rg_set_error(RG_E_ARG, "rg_last_hot_shape: set %d is not in the last evaluated chunk [%d, %d)", set,
                            c->last_first, c->last_first + c->last_n);
    RG_CUDA_OK(cudaSetDevice(c->device));
    RG_CUDA_OK(cudaStreamSynchronize(c->last_stream));
    SetRec r;
    RG_CUDA_OK(cudaMemcpy(&r, c->d_recs + (set - c->last_first), sizeof r, cudaMemcpyDeviceToHost));
    if (r.status != RG_S_OK || r.n_hot == 0 || !r.has_hotshape) {
        memset(out_host, 0, (size_t)c->nE * sizeof(double));
        return RG_OK;
    }
    RG_CUDA_OK(cudaMemcpy(out_host, c->d_farr + (size_t)r.hot_sys * c->nE, (size_t)c->nE * sizeof(double),
                          cudaMemcpyDeviceToHost));
    return RG_OK;
}

extern "C" int rg_eval_batch(rg_ctx *c, const double *d_params, int P, int model, double dr_dex, unsigned flags,
                             double *d_out, double *d_attrs, void *stream)
{
    return eval_batch_impl(c, d_params, P, model, dr_dex, flags, d_out, d_attrs, stream, 0, nullptr, nullptr);
}

extern "C" int rg_eval_batch_pieces(rg_ctx *c, const double *d_params, int P, int model, double dr_dex, unsigned flags,
                                    double *d_out, double *d_attrs, void *stream, int piece_sets, rg_piece_fn on_piece,
                                    void *user)
{
    if (piece_sets < 1 || !on_piece) return rg_set_error(RG_E_ARG, "rg_eval_batch_pieces: piece_sets >= 1 and a callback are needed");
    return eval_batch_impl(c, d_params, P, model, dr_dex, flags, d_out, d_attrs, stream, piece_sets, on_piece, user);
}

static int ensure_dev(double **p, size_t *have, size_t need)
{
    if (*have >= need) return RG_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; *have = 0;
    RG_CUDA_OK(cudaMalloc(p, need * sizeof(double)));
    *have = need;
    return RG_OK;
}

extern "C" int rg_eval_batch_host(rg_ctx *c, const double *params, int P, int model, double dr_dex, unsigned flags,
                                  double *out, double *attrs)
{
    if (!c || !params || !out || P < 0) return rg_set_error(RG_E_ARG, "rg_eval_batch_host: bad arguments");
    if (c->nE <= 0) return rg_set_error(RG_E_GRID, "no energy grid set (rg_set_grid)");
    if (P == 0) return RG_OK;
    RG_CUDA_OK(cudaSetDevice(c->device));
    const int ncomp = (flags & RG_FLAG_COMPONENTS) ? 4 : 1;
    const size_t per_out = (size_t)ncomp * c->nE;
    // The batch is evaluated in the chunks of rg_eval_batch (up to 65 536 sets: one setup / nthcomp / hist launch
    // sequence each); the spectrum kernel of a chunk goes in sub-launches of PIECE sets and every finished piece is
    // copied out on a second stream behind an event while the next piece is computed -- the only copy nobody hides
    // is the last piece's.  (Until r1r the slabs themselves shrank towards the end of the batch; small chunks run
    // 6-12 % slower per set than large ones.)  Batches beyond SUPER sets go in rounds so that the staging buffer stays
    // bounded.
    // ---- small batches (a single model object above all): one stream, pinned staging owned by the context, no host
    //      round trip between the setup kernel and the rest of the chain (eval_batch_impl nosync), one synchronisation
    if (P <= RG_SMALL_SETS && !c->profile && !getenv("RG_NO_SMALL_PATH")) {
        const size_t n_par = (size_t)P * RG_NPAR, n_out = (size_t)P * per_out, n_at = (size_t)P * RG_NATTR;
        if (c->h_pin_n < n_par + n_out + n_at) {
            if (c->h_pin) cudaFreeHost(c->h_pin);
            c->h_pin = nullptr; c->h_pin_n = 0;
            const size_t want = (size_t)RG_SMALL_SETS * RG_NPAR + (size_t)RG_SMALL_SETS * std::max(per_out, (size_t)4 * c->nE) +
                                (size_t)RG_SMALL_SETS * RG_NATTR;
            RG_CUDA_OK(cudaMallocHost(&c->h_pin, want * sizeof(double)));
            c->h_pin_n = want;
        }
        double *h_par = c->h_pin, *h_out = h_par + n_par, *h_at = h_out + n_out;
        RG_TRY(ensure_dev(&c->d_stage_par, &c->d_stage_par_n, n_par));
        RG_TRY(ensure_dev(&c->d_stage_out, &c->d_stage_out_n, n_out));
        RG_TRY(ensure_dev(&c->d_stage_attr, &c->d_stage_attr_n, n_at));
        memcpy(h_par, params, n_par * sizeof(double));
        RG_CUDA_OK(cudaMemcpyAsync(c->d_stage_par, h_par, n_par * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        int rc = eval_batch_impl(c, c->d_stage_par, P, model, dr_dex, flags, c->d_stage_out, c->d_stage_attr, c->stream, 0,
                                 nullptr, nullptr, true);
        if (rc != RG_OK) return rc;
        RG_CUDA_OK(cudaMemcpyAsync(h_out, c->d_stage_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        RG_CUDA_OK(cudaMemcpyAsync(h_at, c->d_stage_attr, n_at * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        RG_CUDA_OK(cudaStreamSynchronize(c->stream));
        const int ann_cap = P * (int)std::min(100000.0, dr_dex * 8.0 + 3.0);
        bool fits = c->h_counter[1] <= ann_cap && c->h_counter[0] + c->h_counter[2] <= c->farr_cap;
        for (int i = 0; i < P && fits; i++) fits = (int)h_at[(size_t)i * RG_NATTR + RG_A_STATUS] != RG_S_CAP;
        if (fits) {
            memcpy(out, h_out, n_out * sizeof(double));
            if (attrs) memcpy(attrs, h_at, n_at * sizeof(double));
            return RG_OK;
        }
        // a set outgrew the caps (dr_dex x 8 + 3 annuli, 40 Kompaneets systems per set): the counting path below
    }
    int PIECE = 32768; // largest piece; the pieces shrink towards the end of the batch (eval_batch_impl)
    bool shrink = true;
    if (const char *e = getenv("RG_PIECE_SETS")) { int v = atoi(e); if (v >= 1 && v <= 65536) { PIECE = v; shrink = false; } } // tuning / test knob: fixed pieces
    const int SUPER = 262144;
    const int stage_rows = std::min(P, SUPER);
    RG_TRY(ensure_dev(&c->d_stage_par, &c->d_stage_par_n, (size_t)P * RG_NPAR));
    RG_TRY(ensure_dev(&c->d_stage_out, &c->d_stage_out_n, (size_t)stage_rows * per_out));
    RG_TRY(ensure_dev(&c->d_stage_attr, &c->d_stage_attr_n, (size_t)P * RG_NATTR));
    RG_CUDA_OK(cudaMemcpyAsync(c->d_stage_par, params, (size_t)P * RG_NPAR * sizeof(double), cudaMemcpyHostToDevice,
                               c->stream));
    struct Pipe {
        rg_ctx *c; cudaStream_t copy_st; cudaEvent_t ev;
        double *out_host, *attrs_host; size_t per_out; int base; // first set of the round
        bool trace; double t0;
    } pipe;
    pipe.c = c; pipe.copy_st = nullptr; pipe.ev = nullptr; pipe.out_host = out; pipe.attrs_host = attrs;
    pipe.per_out = per_out; pipe.base = 0;
    pipe.trace = getenv("RG_TRACE_HOST") != nullptr; // host-side timeline of the pipeline on stderr
    auto now_ms = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    pipe.t0 = now_ms();
    RG_CUDA_OK(cudaStreamCreateWithFlags(&pipe.copy_st, cudaStreamNonBlocking));
    RG_CUDA_OK(cudaEventCreateWithFlags(&pipe.ev, cudaEventDisableTiming));
    auto on_piece = [](void *u, int first, int n) {
        Pipe *p = static_cast<Pipe *>(u);
        rg_ctx *c = p->c;
        cudaEventRecord(p->ev, c->stream);
        cudaStreamWaitEvent(p->copy_st, p->ev, 0);
        const int g = p->base + first; // set index in the whole batch
        cudaMemcpyAsync(p->out_host + (size_t)g * p->per_out, c->d_stage_out + (size_t)first * p->per_out,
                        (size_t)n * p->per_out * sizeof(double), cudaMemcpyDeviceToHost, p->copy_st);
        if (p->attrs_host)
            cudaMemcpyAsync(p->attrs_host + (size_t)g * RG_NATTR, c->d_stage_attr + (size_t)g * RG_NATTR,
                            (size_t)n * RG_NATTR * sizeof(double), cudaMemcpyDeviceToHost, p->copy_st);
        if (p->trace) {
            auto t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
            fprintf(stderr, "[rg host] piece of %d sets enqueued at %.2f ms\n", n, t - p->t0);
        }
    };
    int rc = RG_OK;
    for (int r0 = 0; r0 < P && rc == RG_OK; r0 += SUPER) {
        const int nr = std::min(SUPER, P - r0);
        pipe.base = r0;
        if (r0 > 0) cudaStreamSynchronize(pipe.copy_st); // the staging buffer is about to be overwritten
        rc = eval_batch_impl(c, c->d_stage_par + (size_t)r0 * RG_NPAR, nr, model, dr_dex, flags, c->d_stage_out,
                             c->d_stage_attr + (size_t)r0 * RG_NATTR, c->stream, PIECE, on_piece, &pipe, false, shrink);
    }
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(pipe.copy_st);
    if (pipe.trace) fprintf(stderr, "[rg host] last D2H done at %.2f ms\n", now_ms() - pipe.t0);
    cudaEventDestroy(pipe.ev);
    cudaStreamDestroy(pipe.copy_st);
    if (rc != RG_OK) return rc;
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

// ---------------------------------------------------------------------------
// f1 / f2: observation epilogue (rg_observe.cu)
extern "C" int rg_convert_batch(rg_ctx *c, const double *d_lnu, const double *d_params, int P, int model, int ncomp,
                                int unit, int as_flux, double *d_out, void *stream)
{
    if (!c || !d_lnu || !d_out || P < 0 || ncomp < 1) return rg_set_error(RG_E_ARG, "rg_convert_batch: bad arguments");
    if (unit < RG_UNIT_SI || unit > RG_UNIT_COUNTS) return rg_set_error(RG_E_ARG, "rg_convert_batch: unknown unit %d", unit);
    if (as_flux && !d_params) return rg_set_error(RG_E_ARG, "rg_convert_batch: flux needs the parameter rows (distance, z)");
    if (model != RG_MODEL_RELAGN && model != RG_MODEL_RELQSO) return rg_set_error(RG_E_ARG, "unknown model %d", model);
    if (c->nE <= 0) return rg_set_error(RG_E_GRID, "no energy grid set (rg_set_grid)");
    if (P == 0) return RG_OK;
    RG_CUDA_OK(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    RG_TRY(prof_begin(c, 3, st));
    RG_TRY(rg_launch_convert(d_lnu, d_params, P, model, ncomp, unit, as_flux, c->gd, d_out, c->sm_count, st));
    RG_TRY(prof_end(c, 3, st));
    c->launches++;
    return RG_OK;
}

extern "C" int rg_set_instrument(rg_ctx *c, const double *ear, int ne)
{
    if (!c || !ear || ne < 1) return rg_set_error(RG_E_ARG, "rg_set_instrument: bad arguments");
    for (int i = 0; i < ne; i++)
        if (!(ear[i + 1] > ear[i]) || !(ear[i] >= 0)) return rg_set_error(RG_E_ARG, "instrument bin edges must ascend");
    RG_CUDA_OK(cudaSetDevice(c->device));
    free_dev(c->d_ear);
    c->d_ear = nullptr; c->ne = 0;
    RG_CUDA_OK(cudaMalloc(&c->d_ear, (size_t)(ne + 1) * sizeof(double)));
    RG_CUDA_OK(cudaMemcpy(c->d_ear, ear, (size_t)(ne + 1) * sizeof(double), cudaMemcpyHostToDevice));
    c->ne = ne;
    c->h_ear.assign(ear, ear + ne + 1);
    c->pos0_dirty = true;
    // a response / data set belongs to one instrument grid
    c->n_chan = 0; c->have_data = false;
    return RG_OK;
}

extern "C" int rg_set_response(rg_ctx *c, int n_chan, const int *rowptr, const int *col, const double *val,
                               double exposure)
{
    if (!c || n_chan < 1 || !rowptr || !col || !val) return rg_set_error(RG_E_ARG, "rg_set_response: bad arguments");
    if (c->ne <= 0) return rg_set_error(RG_E_GRID, "rg_set_response: set the instrument grid first (rg_set_instrument)");
    if (!(exposure > 0)) return rg_set_error(RG_E_ARG, "rg_set_response: exposure must be positive");
    for (int ch = 0; ch < n_chan; ch++) {
        if (rowptr[ch + 1] < rowptr[ch]) return rg_set_error(RG_E_ARG, "rg_set_response: rowptr must not descend");
        for (int k = rowptr[ch]; k < rowptr[ch + 1]; k++)
            if (col[k] < 0 || col[k] >= c->ne) return rg_set_error(RG_E_ARG, "rg_set_response: column %d outside the instrument grid", col[k]);
    }
    // sliced ELL: 32 channels per slice, width = longest row of the slice, entries [k][lane] (coalesced per warp)
    const int ns = (n_chan + 31) / 32;
    std::vector<int> off(ns + 1, 0);
    for (int s = 0; s < ns; s++) {
        int w = 0;
        for (int ch = 32 * s; ch < std::min(n_chan, 32 * s + 32); ch++) w = std::max(w, rowptr[ch + 1] - rowptr[ch]);
        off[s + 1] = off[s] + 32 * w;
    }
    std::vector<int> ecol((size_t)std::max(off[ns], 1), 0); // padding: (column 0, value 0) adds +0
    std::vector<double> eval((size_t)std::max(off[ns], 1), 0.0);
    for (int ch = 0; ch < n_chan; ch++) {
        const int s = ch >> 5, lane = ch & 31;
        for (int k = rowptr[ch]; k < rowptr[ch + 1]; k++) {
            size_t q = (size_t)off[s] + (size_t)32 * (k - rowptr[ch]) + lane;
            ecol[q] = col[k];
            eval[q] = val[k];
        }
    }
    RG_CUDA_OK(cudaSetDevice(c->device));
    free_dev(c->d_sl_off); free_dev(c->d_ell_col); free_dev(c->d_ell_val);
    c->d_sl_off = nullptr; c->d_ell_col = nullptr; c->d_ell_val = nullptr; c->n_chan = 0; c->have_data = false;
    RG_CUDA_OK(cudaMalloc(&c->d_sl_off, off.size() * sizeof(int)));
    RG_CUDA_OK(cudaMalloc(&c->d_ell_col, ecol.size() * sizeof(int)));
    RG_CUDA_OK(cudaMalloc(&c->d_ell_val, eval.size() * sizeof(double)));
    RG_CUDA_OK(cudaMemcpy(c->d_sl_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
    RG_CUDA_OK(cudaMemcpy(c->d_ell_col, ecol.data(), ecol.size() * sizeof(int), cudaMemcpyHostToDevice));
    RG_CUDA_OK(cudaMemcpy(c->d_ell_val, eval.data(), eval.size() * sizeof(double), cudaMemcpyHostToDevice));
    c->n_chan = n_chan;
    c->exposure = exposure;
    return RG_OK;
}

extern "C" int rg_set_data(rg_ctx *c, const double *counts, const double *sigma, int kind)
{
    if (!c || !counts) return rg_set_error(RG_E_ARG, "rg_set_data: bad arguments");
    if (c->n_chan <= 0) return rg_set_error(RG_E_ARG, "rg_set_data: set the response first (rg_set_response)");
    if (kind != RG_FIT_CHI2 && kind != RG_FIT_CSTAT) return rg_set_error(RG_E_ARG, "rg_set_data: unknown statistic %d", kind);
    if (kind == RG_FIT_CHI2) {
        if (!sigma) return rg_set_error(RG_E_ARG, "rg_set_data: chi^2 needs the per-channel sigma");
        for (int i = 0; i < c->n_chan; i++)
            if (!(sigma[i] > 0)) return rg_set_error(RG_E_ARG, "rg_set_data: sigma must be positive (channel %d)", i);
    }
    RG_CUDA_OK(cudaSetDevice(c->device));
    free_dev(c->d_counts); free_dev(c->d_sigma);
    c->d_counts = nullptr; c->d_sigma = nullptr; c->have_data = false;
    RG_CUDA_OK(cudaMalloc(&c->d_counts, (size_t)c->n_chan * sizeof(double)));
    RG_CUDA_OK(cudaMemcpy(c->d_counts, counts, (size_t)c->n_chan * sizeof(double), cudaMemcpyHostToDevice));
    // per-channel constant of the statistic: 1 / sigma_c (chi^2) or ln D_c (Cash)
    std::vector<double> aux((size_t)c->n_chan, 0.0);
    for (int i = 0; i < c->n_chan; i++) aux[i] = kind == RG_FIT_CHI2 ? 1.0 / sigma[i] : (counts[i] > 0 ? log(counts[i]) : 0.0);
    RG_CUDA_OK(cudaMalloc(&c->d_sigma, (size_t)c->n_chan * sizeof(double)));
    RG_CUDA_OK(cudaMemcpy(c->d_sigma, aux.data(), (size_t)c->n_chan * sizeof(double), cudaMemcpyHostToDevice));
    c->fit_kind = kind;
    c->have_data = true;
    return RG_OK;
}

static int observe_launch(rg_ctx *c, const double *d_lnu, const double *d_params, int P, int model, double *d_photar,
                          double *d_stat, double *d_folded, cudaStream_t st)
{
    ObsArgs A;
    memset(&A, 0, sizeof A);
    A.lnu = d_lnu; A.params = d_params; A.P = P; A.model = model;
    if (c->pos0_dirty) { // position of every instrument edge on the (geometric) model grid at z = 0
        free_dev(c->d_pos0);
        c->d_pos0 = nullptr;
        if (c->gd.geometric) {
            std::vector<double> pos((size_t)c->ne);
            for (int i = 0; i < c->ne; i++) pos[i] = c->h_ear[i] > 0 ? log(c->h_ear[i] / c->h_ebins[0]) / c->gd.dlnE : -1e9;
            RG_CUDA_OK(cudaMalloc(&c->d_pos0, pos.size() * sizeof(double)));
            RG_CUDA_OK(cudaMemcpyAsync(c->d_pos0, pos.data(), pos.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            RG_CUDA_OK(cudaStreamSynchronize(st)); // pos is a local
        }
        c->pos0_dirty = false;
    }
    A.grid = c->gd; A.ebins = c->d_ebins; A.nE = c->nE; A.kph = c->d_kph;
    A.ear = c->d_ear; A.ne = c->ne;
    A.pos0 = c->d_pos0; A.inv_dlnE = c->gd.geometric ? 1.0 / c->gd.dlnE : 0.0;
    A.photar = d_photar;
    if (d_stat) {
        A.n_chan = c->n_chan; A.sl_off = c->d_sl_off; A.ell_col = c->d_ell_col; A.ell_val = c->d_ell_val;
        A.exposure = c->exposure; A.counts = c->d_counts; A.sigma = c->d_sigma; A.kind = c->fit_kind;
        A.folded = d_folded; A.stat = d_stat;
    }
    RG_TRY(prof_begin(c, 3, st));
    RG_TRY(rg_launch_observe(A, c->sm_count, st));
    RG_TRY(prof_end(c, 3, st));
    c->launches++;
    return RG_OK;
}

static int observe_check(rg_ctx *c, const void *a, const void *b, const void *o, int P, int model, bool need_fit,
                         const char *who)
{
    if (!c || !a || !b || !o || P < 0) return rg_set_error(RG_E_ARG, "%s: bad arguments", who);
    if (model != RG_MODEL_RELAGN && model != RG_MODEL_RELQSO) return rg_set_error(RG_E_ARG, "unknown model %d", model);
    if (c->nE <= 0) return rg_set_error(RG_E_GRID, "no energy grid set (rg_set_grid)");
    if (c->ne <= 0) return rg_set_error(RG_E_GRID, "%s: no instrument grid set (rg_set_instrument)", who);
    if (need_fit && (c->n_chan <= 0 || !c->have_data))
        return rg_set_error(RG_E_ARG, "%s: needs a response and a data set (rg_set_response, rg_set_data)", who);
    return RG_OK;
}

extern "C" int rg_photar_batch(rg_ctx *c, const double *d_lnu, const double *d_params, int P, int model,
                               double *d_photar, void *stream)
{
    RG_TRY(observe_check(c, d_lnu, d_params, d_photar, P, model, false, "rg_photar_batch"));
    if (P == 0) return RG_OK;
    RG_CUDA_OK(cudaSetDevice(c->device));
    return observe_launch(c, d_lnu, d_params, P, model, d_photar, nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int rg_fitstat_batch(rg_ctx *c, const double *d_lnu, const double *d_params, int P, int model,
                                double *d_stat, double *d_folded, void *stream)
{
    RG_TRY(observe_check(c, d_lnu, d_params, d_stat, P, model, true, "rg_fitstat_batch"));
    if (P == 0) return RG_OK;
    RG_CUDA_OK(cudaSetDevice(c->device));
    return observe_launch(c, d_lnu, d_params, P, model, nullptr, d_stat, d_folded, (cudaStream_t)stream);
}

extern "C" int rg_eval_fitstat_batch(rg_ctx *c, const double *d_params, int P, int model, double dr_dex, unsigned flags,
                                     double *d_stat, double *d_attrs, void *stream)
{
    RG_TRY(observe_check(c, d_params, d_params, d_stat, P, model, true, "rg_eval_fitstat_batch"));
    if (flags & RG_FLAG_COMPONENTS) return rg_set_error(RG_E_ARG, "rg_eval_fitstat_batch: the statistic uses the total SED only");
    if (P == 0) return RG_OK;
    RG_CUDA_OK(cudaSetDevice(c->device));
    const int SLAB = pick_chunk_sets(P); // = the chunk of rg_eval_batch
    RG_TRY(ensure_dev(&c->d_slab, &c->slab_n, (size_t)std::min(P, SLAB) * c->nE));
    for (int p0 = 0; p0 < P; p0 += SLAB) {
        const int n = std::min(SLAB, P - p0);
        const double *par = d_params + (size_t)p0 * RG_NPAR;
        RG_TRY(rg_eval_batch(c, par, n, model, dr_dex, flags, c->d_slab, d_attrs ? d_attrs + (size_t)p0 * RG_NATTR : nullptr,
                             stream));
        RG_TRY(observe_launch(c, c->d_slab, par, n, model, nullptr, d_stat + p0, nullptr, (cudaStream_t)stream));
    }
    return RG_OK;
}

extern "C" int rg_eval_fitstat_batch_host(rg_ctx *c, const double *params, int P, int model, double dr_dex,
                                          unsigned flags, double *stat, double *attrs)
{
    if (!c || !params || !stat || P < 0) return rg_set_error(RG_E_ARG, "rg_eval_fitstat_batch_host: bad arguments");
    if (P == 0) return RG_OK;
    RG_CUDA_OK(cudaSetDevice(c->device));
    RG_TRY(ensure_dev(&c->d_stage_par, &c->d_stage_par_n, (size_t)P * RG_NPAR));
    RG_TRY(ensure_dev(&c->d_stage_stat, &c->d_stage_stat_n, (size_t)P));
    if (attrs) RG_TRY(ensure_dev(&c->d_stage_attr, &c->d_stage_attr_n, (size_t)P * RG_NATTR));
    RG_CUDA_OK(cudaMemcpyAsync(c->d_stage_par, params, (size_t)P * RG_NPAR * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RG_TRY(rg_eval_fitstat_batch(c, c->d_stage_par, P, model, dr_dex, flags, c->d_stage_stat, attrs ? c->d_stage_attr : nullptr,
                                 c->stream));
    RG_CUDA_OK(cudaMemcpyAsync(stat, c->d_stage_stat, (size_t)P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (attrs)
        RG_CUDA_OK(cudaMemcpyAsync(attrs, c->d_stage_attr, (size_t)P * RG_NATTR * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RG_CUDA_OK(cudaStreamSynchronize(c->stream));
    return RG_OK;
}

// ---------------------------------------------------------------------------
extern "C" int rg_measure_fp64_peak(rg_ctx *c, double *tflops)
{
    if (!c || !tflops) return rg_set_error(RG_E_ARG, "rg_measure_fp64_peak: bad arguments");
    RG_CUDA_OK(cudaSetDevice(c->device));
    c->launches += 2;
    return rg_fp64_peak(tflops, c->stream);
}
