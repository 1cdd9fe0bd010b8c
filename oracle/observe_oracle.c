/*
 * observe_oracle.c -- CPU restatement of the observation epilogue.
 *
 * TEST INFRASTRUCTURE ONLY (see relagn_oracle.h): the checker of
 * relagn_b200/csrc/rg_observe.cu, never part of the product path.
 *
 * What it follows:
 *   ro_convert      src/python_version/relagn.py:487-564 (_to_newUnit, _to_flux), op for op.
 *                   PINNED by tests/golden/golden_units.npz (getter outputs of the unmodified reference).
 *   ro_photar       src/fortran_version/relagn_dir/relagn.f:98-112: model grid and photons divided by (1 + z),
 *                   then XSPEC's inibin/erebin (fractional-overlap rebin).  inibin/erebin live in HEASOFT (not in
 *                   /root/reference, tested by the author with HEASOFT 6.29-6.31): restated from their documented
 *                   behaviour -> PARITY UNPINNED against XSPEC itself; pinned here by photon conservation and the
 *                   identity rebin.
 *   ro_fold_stat    no counterpart in the reference (row f2 is a new fused capability): XSPEC's published
 *                   definitions of the folded model, chi^2 and the Cash statistic (XSPEC manual, App. B).
 */
#include <math.h>
#include <stddef.h>

#include "relagn_oracle.h"

#define K_H 6.62607015e-34
#define K_C 299792458.0
#define K_KEV 1.602176634e-16
#define K_MPC_CM 3.0856775814913674e+24
#define K_PI 3.141592653589793

/* relagn.py:487-564.  L: [nE] W/Hz on the grid ebins[nE+1] (keV). */
void ro_convert(const double *L, int nE, const double *ebins, int unit, int as_flux, double dist_mpc, double z,
                double *out)
{
    double d = dist_mpc * K_MPC_CM; /* cm, relagn.py:285 */
    if (unit == 0 || unit == 3) d = d / 100; /* SI, SI_wave: metres (relagn.py:558-559) */
    double den = 4 * K_PI * (d * d) * (1 + z); /* relagn.py:561 */
    for (int j = 0; j < nE; j++) {
        double dE = ebins[j + 1] - ebins[j];
        double E = ebins[j] + dE / 2;      /* relagn.py:337-338 */
        double nu = (E * K_KEV) / K_H;     /* relagn.py:339 */
        double v = L[j];
        switch (unit) {
        case 1: v = v * 1e7; break;                                      /* cgs        :507 */
        case 2: v = (((v * (nu * nu)) / K_C) * 1e-10) * 1e7; break;      /* cgs_wave   :518-520 */
        case 3: v = ((v * (nu * nu)) / K_C) * 1e-10; break;              /* SI_wave    :522-524 */
        case 4: v = (v / K_H) / E; break;                                /* counts     :509-516 */
        default: break;                                                  /* SI         :526 */
        }
        if (as_flux) v = v / den;
        out[j] = v;
    }
}

/* relagn.f:98-112.  photar[i] = photons/cm^2/s in the observer's bin [ear[i], ear[i+1]]. */
void ro_photar(const double *L, int nE, const double *ebins, double dist_mpc, double z, const double *ear, int ne,
               double *photar)
{
    double d = dist_mpc * K_MPC_CM;
    double inv_den = 1.0 / (4 * K_PI * (d * d) * (1 + z));
    double zf = 1 + z;
    /* photons/cm^2/s per model bin: counts flux (relagn.py:509-516, 561) times the bin width, grouped as
     * L * [dE / (h E)] * [1 / (4 pi d^2 (1 + z))] (the device keeps the first bracket per bin, the second per set;
     * <= 2 ulp from the reference's left-to-right order, checked against ro_convert in tests/test_oracle.py) */
    /* (variable-length arrays keep the oracle allocation-free; nE is at most a few thousand) */
    double ph[nE > 0 ? nE : 1], x[nE + 1];
    for (int j = 0; j <= nE; j++) x[j] = ebins[j] / zf;
    for (int j = 0; j < nE; j++) {
        double dE = ebins[j + 1] - ebins[j];
        double E = ebins[j] + dE / 2;
        ph[j] = (L[j] * (dE / (K_H * E))) * inv_den;
    }
    for (int i = 0; i < ne; i++) {
        double lo = ear[i], hi = ear[i + 1];
        double acc = 0.0;
        if (hi > x[0] && lo < x[nE] && hi > lo) {
            int js = 0, je = nE - 1;
            while (x[js + 1] <= lo) js++; /* first model bin reaching above lo */
            while (x[je] >= hi) je--;     /* last model bin starting below hi */
            if (js == je) {
                double a = lo > x[js] ? lo : x[js], b = hi < x[js + 1] ? hi : x[js + 1];
                acc = ph[js] * ((b - a) / (x[js + 1] - x[js]));
            } else {
                double a = lo > x[js] ? lo : x[js];
                acc = ph[js] * ((x[js + 1] - a) / (x[js + 1] - x[js]));
                for (int k = js + 1; k < je; k++) acc = acc + ph[k];
                double b = hi < x[je + 1] ? hi : x[je + 1];
                acc = acc + ph[je] * ((b - x[je]) / (x[je + 1] - x[je]));
            }
        }
        photar[i] = acc;
    }
}

/* folded model counts and the fit statistic; kind 0 = chi^2, 1 = Cash */
double ro_fold_stat(const double *photar, int n_chan, const int *rowptr, const int *col, const double *val,
                    double exposure, const double *counts, const double *sigma, int kind, double *folded)
{
    double stat = 0.0;
    for (int c = 0; c < n_chan; c++) {
        double m = 0.0;
        for (int k = rowptr[c]; k < rowptr[c + 1]; k++) m = m + val[k] * photar[col[k]];
        m = exposure * m;
        if (folded) folded[c] = m;
        double D = counts[c];
        if (kind == 0) {
            double t = (D - m) * (1.0 / sigma[c]);
            stat = stat + t * t;
        } else {
            double mm = m < 1e-300 ? 1e-300 : m;
            double t = D > 0 ? (mm - D) + D * (log(D) - log(mm)) : mm;
            stat = stat + 2 * t;
        }
    }
    return stat;
}
