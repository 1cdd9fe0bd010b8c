// rg_xspec.cu -- seam B3: the XSPEC local-model ABI of the Fortran twins
//   subroutine relagn(ear, ne, param, ifl, photar, photer)   src/fortran_version/relagn_dir/relagn.f:17-115
//   subroutine relqso(ear, ne, param, ifl, photar, photer)   src/fortran_version/relqso_dir/relqso.f
// as C symbols with gfortran's calling convention (every argument by reference, real*4 arrays), so that
// lmod_relagn.dat / lmod_relqso.dat can load the GPU model in place of the Fortran one.
//
// Host-side driver only: it restates the wrapper logic of relagn.f:60-112 (own model grid of 2000 log bins that
// covers [min(1e-4, ear(0) zfac), max(1e3, ear(ne) zfac)], model evaluation, energies and photons divided by
// 1 + z, inibin/erebin onto the caller's grid, parameter/grid memoisation) on top of the C ABI of this library:
// rg_set_grid + rg_eval_batch (one parameter set) + rg_set_instrument + rg_photar_batch.  The physics behind it
// is the PYTHON class's (relagn.py), not the Fortran twin's, which differs from it in 14 places (SURVEY App. B:
// dr_dex = 30, rounded constants, ...): the numbers are those of get_totSED(rel=True) in photons/cm^2/s per bin.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "rg_kernels.h"

#define XS_NNEW 2000 // relagn.f:28-29
#define RG_TRY_XS(expr)                 \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != RG_OK) return rc__; \
    } while (0)

namespace {
struct XsState {
    rg_ctx *ctx = nullptr;
    bool own_ctx = false;
    double emin = 0, emax = 0;            // model grid currently set
    std::vector<double> ear;              // instrument grid currently set
    std::vector<double> oldpar;           // relagn.f:37 `save oldpar`
    int oldmodel = -1;
    unsigned oldflags = 0;
    std::vector<double> photar;           // memoised result on `ear`
    double *d_par = nullptr, *d_lnu = nullptr, *d_pho = nullptr, *d_attr = nullptr;
    size_t pho_cap = 0;
};
XsState g_xs;

int xs_fail(const char *who)
{
    fprintf(stderr, "relagn_b200 (%s): %s\n", who, rg_last_error());
    return -1;
}
} // namespace

// Default Kerr node grid of the package (relagn_b200/tables.py DEFAULT); the array itself comes from the .npy cache
// the Python layer writes (npy_path or $RELAGN_B200_TABLES) or is ray-traced on the device (about 35 s, once).
extern "C" int rg_tables_default(rg_ctx *ctx, const char *npy_path)
{
    static const double spins[14] = {-0.998, -0.6, -0.2, 0.1, 0.3, 0.5, 0.7, 0.8, 0.9, 0.95, 0.975, 0.9875, 0.995, 0.998};
    static const double deg[14] = {0.0, 10.0, 20.0, 30.0, 40.0, 50.0, 57.5, 65.0, 70.0, 75.0, 77.5, 80.0, 82.5, 85.0};
    double thetas[14];
    for (int i = 0; i < 14; i++) thetas[i] = deg[i] * (RGK_PI / 180.0);
    const int NR = 64, NPHI = 64;
    const double r_in = 1.2369;
    if (!ctx) return rg_set_error(RG_E_ARG, "rg_tables_default: ctx is NULL");
    if (!npy_path) npy_path = getenv("RELAGN_B200_TABLES");
    if (npy_path && *npy_path) {
        FILE *f = fopen(npy_path, "rb");
        if (!f) return rg_set_error(RG_E_ARG, "rg_tables_default: cannot open %s", npy_path);
        // .npy: magic(6) version(2) header length (2 bytes v1, 4 bytes v2+), ASCII dict, then the raw array
        unsigned char head[12];
        size_t n = (size_t)14 * 14 * 2 * NR * NPHI;
        std::vector<float> data(n);
        bool ok = fread(head, 1, 10, f) == 10 && memcmp(head, "\x93NUMPY", 6) == 0;
        size_t hlen = 0;
        if (ok) {
            if (head[6] == 1) hlen = head[8] | (head[9] << 8);
            else { ok = fread(head + 10, 1, 2, f) == 2; hlen = head[8] | (head[9] << 8) | (head[10] << 16) | ((size_t)head[11] << 24); }
        }
        std::vector<char> hdr(hlen + 1, 0);
        ok = ok && fread(hdr.data(), 1, hlen, f) == hlen && strstr(hdr.data(), "'<f4'") && strstr(hdr.data(), "False") &&
             strstr(hdr.data(), "(14, 14, 2, 64, 64)");
        ok = ok && fread(data.data(), sizeof(float), n, f) == n;
        fclose(f);
        if (!ok) return rg_set_error(RG_E_ARG, "rg_tables_default: %s is not the default table array (<f4, C order, 14x14x2x64x64)", npy_path);
        return rg_tables_upload(ctx, spins, 14, thetas, 14, NR, NPHI, r_in, data.data());
    }
    long bad = 0;
    return rg_tables_generate(ctx, spins, 14, thetas, 14, NR, NPHI, r_in, &bad);
}

// Use a caller-owned context (with its resident tables) for the XSPEC entry points; NULL: drop the binding.
extern "C" int rg_xspec_bind(rg_ctx *ctx)
{
    if (g_xs.d_par) cudaFree(g_xs.d_par);
    if (g_xs.d_lnu) cudaFree(g_xs.d_lnu);
    if (g_xs.d_attr) cudaFree(g_xs.d_attr);
    g_xs.d_attr = nullptr;
    if (g_xs.d_pho) cudaFree(g_xs.d_pho);
    if (g_xs.own_ctx && g_xs.ctx) rg_ctx_destroy(g_xs.ctx);
    g_xs = XsState(); // (while bound, the context's model and instrument grids belong to the XSPEC entry points)
    g_xs.ctx = ctx;
    return RG_OK;
}

// double-precision core of relagn_ / relqso_.  par: RG_NPAR slots in the Python constructor order.
extern "C" int rg_xspec_eval(int model, unsigned flags, const double *ear, int ne, const double *par, double *photar)
{
    if (!ear || !par || !photar || ne < 1) return rg_set_error(RG_E_ARG, "rg_xspec_eval: bad arguments");
    XsState &S = g_xs;
    if (!S.ctx) {
        RG_TRY_XS(rg_ctx_create(-1, &S.ctx));
        S.own_ctx = true;
        RG_TRY_XS(rg_tables_default(S.ctx, nullptr));
    }
    const double z = par[model == RG_MODEL_RELQSO ? 6 : 14];
    const double zfac = 1.0 + z;
    // relagn.f:68-90: the model grid
    double newemin = ear[0] == 0.0 ? ear[1] - (ear[2 < ne ? 2 : ne] - ear[1]) / 10.0 : fmin(1.0e-4, ear[0] * zfac);
    double newemax = fmax(1.0e3, ear[ne] * zfac);
    if (!(newemin > 0)) newemin = 1.0e-4;
    bool echange = newemin != S.emin || newemax != S.emax;
    if (echange) {
        std::vector<double> enew(XS_NNEW + 1);
        const double l0 = log(newemin), dl = log(newemax / newemin) / XS_NNEW;
        for (int n = 0; n <= XS_NNEW; n++) enew[n] = exp(l0 + dl * n);
        enew[0] = newemin; enew[XS_NNEW] = newemax;
        RG_TRY_XS(rg_set_grid(S.ctx, enew.data(), XS_NNEW));
        S.emin = newemin; S.emax = newemax;
    }
    bool earchange = (int)S.ear.size() != ne + 1 || memcmp(S.ear.data(), ear, (size_t)(ne + 1) * sizeof(double)) != 0;
    if (earchange) {
        RG_TRY_XS(rg_set_instrument(S.ctx, ear, ne));
        S.ear.assign(ear, ear + ne + 1);
    }
    bool parchange = S.oldmodel != model || S.oldflags != flags || S.oldpar.size() != RG_NPAR ||
                     memcmp(S.oldpar.data(), par, RG_NPAR * sizeof(double)) != 0;
    if (parchange || echange || earchange) { // relagn.f:93-112 (the rebin is redone only when something moved)
        if (!S.d_par) {
            if (cudaMalloc(&S.d_par, RG_NPAR * sizeof(double)) != cudaSuccess ||
                cudaMalloc(&S.d_lnu, XS_NNEW * sizeof(double)) != cudaSuccess)
                return rg_set_error(RG_E_NOMEM, "rg_xspec_eval: device allocation failed");
        }
        if ((size_t)ne > S.pho_cap) {
            if (S.d_pho) cudaFree(S.d_pho);
            S.d_pho = nullptr; S.pho_cap = 0;
            if (cudaMalloc(&S.d_pho, (size_t)ne * sizeof(double)) != cudaSuccess)
                return rg_set_error(RG_E_NOMEM, "rg_xspec_eval: device allocation failed");
            S.pho_cap = ne;
        }
        if (!S.d_attr && cudaMalloc(&S.d_attr, RG_NATTR * sizeof(double)) != cudaSuccess)
            return rg_set_error(RG_E_NOMEM, "rg_xspec_eval: device allocation failed");
        RG_CUDA_OK(cudaMemcpy(S.d_par, par, RG_NPAR * sizeof(double), cudaMemcpyHostToDevice));
        RG_TRY_XS(rg_eval_batch(S.ctx, S.d_par, 1, model, 50.0, flags, S.d_lnu, S.d_attr, nullptr));
        {
            // the reference raises ValueError for these parameter sets (relagn.py:372-386, 1861-1874); XSPEC gets an
            // error return (zeros + a message from the caller) instead of a silently memoised all-zero spectrum
            double at[RG_NATTR];
            RG_CUDA_OK(cudaMemcpy(at, S.d_attr, sizeof at, cudaMemcpyDeviceToHost));
            const int st = (int)at[RG_A_STATUS];
            if (st != RG_S_OK) {
                S.oldpar.clear();
                return rg_set_error(RG_E_ARG, "parameter set rejected: %s",
                                    st == RG_S_SPIN ? "spin outside -0.998 <= a <= 0.998"
                                    : st == RG_S_INC ? "inclination outside 0.09 <= cos(inc) <= 1"
                                    : st == RG_S_MDOT ? "log mdot outside -1.65 ... 0.5" : "arena capacity");
            }
        }
        RG_TRY_XS(rg_photar_batch(S.ctx, S.d_lnu, S.d_par, 1, model, S.d_pho, nullptr));
        S.photar.resize(ne);
        RG_CUDA_OK(cudaMemcpy(S.photar.data(), S.d_pho, (size_t)ne * sizeof(double), cudaMemcpyDeviceToHost));
        S.oldpar.assign(par, par + RG_NPAR);
        S.oldmodel = model; S.oldflags = flags;
    }
    memcpy(photar, S.photar.data(), (size_t)ne * sizeof(double));
    return RG_OK;
}

static void xs_call(int model, unsigned flags, const float *ear, int ne, const double *par, float *photar,
                    float *photer, const char *who)
{
    std::vector<double> e((size_t)ne + 1), ph((size_t)ne, 0.0);
    for (int i = 0; i <= ne; i++) e[i] = (double)ear[i];
    int rc = rg_xspec_eval(model, flags, e.data(), ne, par, ph.data());
    if (rc != RG_OK) xs_fail(who); // XSPEC has no error channel here: zeros + a message, like a failed xwrite
    for (int i = 0; i < ne; i++) {
        photar[i] = rc == RG_OK ? (float)ph[i] : 0.0f;
        if (photer) photer[i] = 0.0f;
    }
}

// lmod_relagn.dat:1-16 -- 15 parameters in the order of the Python constructor (relagn.f:45-59)
extern "C" void relagn_(const float *ear, const int *ne, const float *param, const int *ifl, float *photar,
                        float *photer)
{
    (void)ifl;
    double par[RG_NPAR];
    for (int i = 0; i < RG_NPAR; i++) par[i] = (double)param[i];
    xs_call(RG_MODEL_RELAGN, RG_FLAG_REL, ear, *ne, par, photar, photer, "relagn");
}

// lmod_relqso.dat -- M, Dist, log_mdot, astar, cos_inc, redshift, $rel (relqso.f:49-56); fcol = 1 as in relqso.f
extern "C" void relqso_(const float *ear, const int *ne, const float *param, const int *ifl, float *photar,
                        float *photer)
{
    (void)ifl;
    double par[RG_NPAR];
    memset(par, 0, sizeof par);
    for (int i = 0; i < 5; i++) par[i] = (double)param[i];
    par[5] = 1.0;               // fcol
    par[6] = (double)param[5];  // redshift
    const unsigned flags = param[6] == 0.0f ? 0u : RG_FLAG_REL; // relqso.f:192-198
    xs_call(RG_MODEL_RELQSO, flags, ear, *ne, par, photar, photer, "relqso");
}
