/*
 * kerr_oracle.c -- CPU (FP64) restatement of the relativistic transfer seam:
 * Kerr ray tracing -> transfer tables -> KYCONV-style convolution.
 * TEST INFRASTRUCTURE ONLY (see relagn_oracle.h).   PARITY UNPINNED:
 *
 * The reference reaches this arithmetic through
 *   xspec.callModelFunction('kyconv', Ebins, kyparams[12], phs)
 *   (/root/reference/src/python_version/relagn.py:979-1001, 1095-1099, 1353-1357)
 * i.e. HEASOFT/XSPEC's `kyconv` (KY package, Dovciak, Karas & Yaqoob 2004,
 * ApJS 153, 205; HEASOFT 6.29-6.31 per README.md:26) and its FITS tables,
 * which are not in /root/reference and cannot be installed offline.  No test
 * of the reference pins a number at this seam (tests/blurr_test.py only
 * plots).  What is restated here is the published integral
 *
 *   out_j = Int_{r_in}^{r_out} dr Int_0^{2pi} dphi  g^2 l mu_e r
 *           Int_{E_j/g}^{E_{j+1}/g} n_loc(E') dE'
 *
 * with the identity  l mu_e r dr dphi = g dalpha dbeta  (image plane), so
 * G r dr dphi = g^3 dalpha dbeta.  The flat-space limit
 * out = in * pi (r_out^2 - r_in^2) cos(i) is the one the reference itself
 * relies on (relagn.py:968-969, 1010-1013).
 *
 * Discretisation (ours; identical in the CUDA product):
 *  - tables per node (spin a_m, inclination theta_n): g(r,phi) and
 *    Jn(r,phi) = |d(alpha,beta)/d(x,y)| (x,y disc-plane Cartesian), on rows
 *    r_s = r_h + (risco - r_h) 10^(s dl)  (r_h = horizon, both of a_m) and
 *    columns phi_t = 2 pi t / NPHI + phi0(r_s), phi0 = azimuth of the front ray
 *    (removes frame-dragging winding between rows), float32;
 *  - lookup: linear in log10((r - r_h)/(risco - r_h)) of the model's own spin,
 *    in risco(a) between spin nodes and in theta between inclination nodes;
 *  - an annulus is cut into n_sub rings (ro_auto_nsub: radial shift variation per
 *    ring below ~half an energy bin) at log-midpoints with exact area weights; each ring is a closed polyline in phi; every phi segment spreads
 *    its weight uniformly in s = ln(g)/dlnE between its end points and is
 *    deposited on integer shifts with a linear (hat) kernel;
 *  - on a geometric energy grid the convolution is out_j = sum_k Hc[k] in_{j-k}.
 */
#include "relagn_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "par_for.h"

#define PI 3.141592653589793

/* ------------------------------------------------------------------ */
/* Null geodesics in Mino time l with u = 1/r, mu = cos(theta) and the  */
/* equatorial projection (X,Y) = sin(theta)(cos phi_g, sin phi_g) of the */
/* direction vector, phi = phi_g + phi_fd:                              */
/*   u''  = (a^2-lam^2-q) u + 3((a-lam)^2+q) u^2 - 2 a^2 q u^3   (=U'/2) */
/*   mu'' = (a^2-q-lam^2) mu - 2 a^2 mu^3                        (=M'/2) */
/*   X''  = -(q+lam^2+2 a^2 mu^2) X ,  Y'' likewise                      */
/*   phi_fd' = a u (2 - a lam u)/(1 - 2u + a^2 u^2)                      */
/* All right-hand sides are regular through radial/polar turning points  */
/* and through the polar axis (no 1/sin^2 theta term), so theta_o = 0    */
/* and lam = 0 rays need no special case.                                */
/* ------------------------------------------------------------------ */

#define NV 9
typedef struct { double a, lam, q, c1, c2, c3, a2, ql; } geo_t;

static void geo_rhs(const geo_t *G, const double *y, double *f)
{
    double u = y[0], mu = y[2];
    double kap = -(G->ql + 2 * G->a2 * mu * mu);
    f[0] = y[1];
    f[1] = u * (G->c1 + u * (G->c2 + u * G->c3));
    f[2] = y[3];
    f[3] = mu * (G->c1 - 2 * G->a2 * mu * mu);
    f[4] = y[5];
    f[5] = kap * y[4];
    f[6] = y[7];
    f[7] = kap * y[6];
    f[8] = G->a * u * (2 - G->a * G->lam * u) / (1 - 2 * u + G->a2 * u * u);
}

/* same system with mu as the independent variable (Henon's trick) */
static void geo_rhs_mu(const geo_t *G, const double *y, double *f)
{
    double t[NV];
    geo_rhs(G, y, t);
    double iw = 1.0 / t[2]; /* dl/dmu */
    for (int i = 0; i < NV; i++) f[i] = t[i] * iw;
    f[2] = 1.0;
}

typedef void (*rhs_fn)(const geo_t *, const double *, double *);

/* one Dormand-Prince 5(4) step (FSAL: k1 in, k7 out).  returns error norm. */
static double dopri_step(const geo_t *G, rhs_fn rhs, const double *y, const double *k1, double h,
                         double *ynew, double *k7, double rtol, double atol)
{
    double k2[NV], k3[NV], k4[NV], k5[NV], k6[NV], t[NV];
    int i;
    for (i = 0; i < NV; i++) t[i] = y[i] + h * (1.0 / 5) * k1[i];
    rhs(G, t, k2);
    for (i = 0; i < NV; i++) t[i] = y[i] + h * (3.0 / 40 * k1[i] + 9.0 / 40 * k2[i]);
    rhs(G, t, k3);
    for (i = 0; i < NV; i++) t[i] = y[i] + h * (44.0 / 45 * k1[i] - 56.0 / 15 * k2[i] + 32.0 / 9 * k3[i]);
    rhs(G, t, k4);
    for (i = 0; i < NV; i++) t[i] = y[i] + h * (19372.0 / 6561 * k1[i] - 25360.0 / 2187 * k2[i] + 64448.0 / 6561 * k3[i] - 212.0 / 729 * k4[i]);
    rhs(G, t, k5);
    for (i = 0; i < NV; i++) t[i] = y[i] + h * (9017.0 / 3168 * k1[i] - 355.0 / 33 * k2[i] + 46732.0 / 5247 * k3[i] + 49.0 / 176 * k4[i] - 5103.0 / 18656 * k5[i]);
    rhs(G, t, k6);
    for (i = 0; i < NV; i++) ynew[i] = y[i] + h * (35.0 / 384 * k1[i] + 500.0 / 1113 * k3[i] + 125.0 / 192 * k4[i] - 2187.0 / 6784 * k5[i] + 11.0 / 84 * k6[i]);
    rhs(G, ynew, k7);
    double err = 0;
    for (i = 0; i < NV; i++) {
        double e = h * (71.0 / 57600 * k1[i] - 71.0 / 16695 * k3[i] + 71.0 / 1920 * k4[i] - 17253.0 / 339200 * k5[i] + 22.0 / 525 * k6[i] - 1.0 / 40 * k7[i]);
        double sc = atol + rtol * fmax(fabs(y[i]), fabs(ynew[i]));
        double r = e / sc;
        err += r * r;
    }
    return sqrt(err / NV);
}

static double g_rtol = 1e-12, g_atol = 1e-14;
/* limb-darkening law of the local emission (kyconv parameter 10, relagn.py:988; KY: Dovciak et al. 2004):
 * 0 isotropic (what the reference passes), -1 Laor (1991) 1 + 2.06 mu_e, -2 Haardt (1993) ln(1 + 1/mu_e).
 * The law multiplies the weight Jn of every table sample: a table set is built for one law. */
static int g_limb = 0;
void ro_set_limb(int law) { g_limb = law; }
double ro_limb_factor(int law, double mue)
{
    if (law == -1) return 1.0 + 2.06 * mue;
    if (law == -2) return log(1.0 + 1.0 / mue);
    return 1.0;
}
void ro_set_ray_tol(double rtol, double atol) { g_rtol = rtol; g_atol = atol; }

int ro_trace_ray(double a, double theta_o, double alpha, double beta, double *r_hit, double *phi_hit)
{
    const double rtol = g_rtol, atol = g_atol;
    double so = sin(theta_o), co = cos(theta_o);
    geo_t G;
    G.a = a; G.a2 = a * a;
    G.lam = -alpha * so;
    G.q = beta * beta + co * co * (alpha * alpha - a * a);
    G.ql = G.q + G.lam * G.lam;
    G.c1 = G.a2 - G.lam * G.lam - G.q;
    G.c2 = 3 * ((a - G.lam) * (a - G.lam) + G.q);
    G.c3 = -2 * G.a2 * G.q;
    double uh = 1.0 / (1.0 + sqrt(1.0 - a * a));
    /* observer at u=0, phi=0: mu' = beta sin(theta_o); X' = -beta cos(theta_o); Y' = lam/sin(theta_o) = -alpha */
    double y[NV] = {0.0, 1.0, co, beta * so, so, -beta * co, 0.0, -alpha, 0.0};
    double k1[NV], k7[NV], yn[NV];
    geo_rhs(&G, y, k1);
    double h = 1e-3 / (1.0 + hypot(alpha, beta));
    for (int step = 0; step < 200000; step++) {
        double err = dopri_step(&G, geo_rhs, y, k1, h, yn, k7, rtol, atol);
        if (!(err <= 1.0)) { /* reject (also catches NaN) */
            double fac = 0.9 * pow(err, -0.2);
            if (!(fac > 0.1)) fac = 0.1;
            h *= fac;
            if (h < 1e-300) return 0;
            continue;
        }
        if (yn[0] >= uh * (1 - 1e-9)) return 0; /* captured */
        if (yn[2] <= 0.0) {
            /* crossing inside this step: integrate in mu from y[2] to 0 */
            const int NS = 4;
            double ys[NV], kk[NV], kn[NV], yt[NV];
            memcpy(ys, y, sizeof ys);
            double dmu = -y[2] / NS;
            for (int s = 0; s < NS; s++) {
                geo_rhs_mu(&G, ys, kk);
                dopri_step(&G, geo_rhs_mu, ys, kk, dmu, yt, kn, rtol, atol);
                memcpy(ys, yt, sizeof ys);
            }
            if (ys[0] >= uh * (1 - 1e-9) || !(ys[0] > 0)) return 0;
            *r_hit = 1.0 / ys[0];
            *phi_hit = atan2(ys[6], ys[4]) + ys[8];
            return 1;
        }
        memcpy(y, yn, sizeof y);
        memcpy(k1, k7, sizeof k1);
        double fac = err > 0 ? 0.9 * pow(err, -0.2) : 5.0;
        if (fac > 5.0) fac = 5.0;
        if (fac < 0.2) fac = 0.2;
        h *= fac;
    }
    return 0;
}

/* redshift factor of a Keplerian emitter at r for a photon with L_z = lam */
static double kep_g(double a, double r, double lam)
{
    double sr = sqrt(r), r32 = r * sr;
    return sqrt(sr) * sr * sqrt(r32 - 3 * sr + 2 * a) / (r32 + a - lam);
}

static int ray_xy(double a, double th, double al, double be, double *X, double *Y)
{
    double r, ph;
    if (!ro_trace_ray(a, th, al, be, &r, &ph)) return 0;
    *X = r * cos(ph); *Y = r * sin(ph);
    return 1;
}

/* solve for the image-plane point whose ray lands on (r_e, phi_e).
 * al, be: in = guess, out = solution.  returns 1 ok. */
int ro_solve_point(double a, double th, double r_e, double phi_e, double *al, double *be, double *g, double *jn)
{
    double tx = r_e * cos(phi_e), ty = r_e * sin(phi_e);
    double A = *al, B = *be;
    double X0, Y0;
    int ok = ray_xy(a, th, A, B, &X0, &Y0);
    if (!ok) return 0;
    for (int it = 0; it < 40; it++) {
        double rx = X0 - tx, ry = Y0 - ty;
        if (hypot(rx, ry) < 1e-10 * (1 + r_e)) break;
        if (it == 39) return 0;
        double h = 1e-6 * (1 + hypot(A, B));
        double Xa, Ya, Xb, Yb;
        if (!ray_xy(a, th, A + h, B, &Xa, &Ya)) return 0;
        if (!ray_xy(a, th, A, B + h, &Xb, &Yb)) return 0;
        double j11 = (Xa - X0) / h, j21 = (Ya - Y0) / h, j12 = (Xb - X0) / h, j22 = (Yb - Y0) / h;
        double det = j11 * j22 - j12 * j21;
        if (det == 0 || det != det) return 0;
        double dA = -(j22 * rx - j12 * ry) / det;
        double dB = -(-j21 * rx + j11 * ry) / det;
        double lim = 0.5 * (1 + hypot(A, B));
        double sz = hypot(dA, dB);
        if (sz > lim) { dA *= lim / sz; dB *= lim / sz; }
        int tries = 0;
        double Xn, Yn;
        while (!ray_xy(a, th, A + dA, B + dB, &Xn, &Yn)) {
            dA *= 0.5; dB *= 0.5;
            if (++tries > 12) return 0;
        }
        /* simple backtracking on the residual */
        tries = 0;
        while (hypot(Xn - tx, Yn - ty) > hypot(rx, ry) * (1 + 1e-3) && tries < 8) {
            dA *= 0.5; dB *= 0.5; tries++;
            if (!ray_xy(a, th, A + dA, B + dB, &Xn, &Yn)) { tries = 99; break; }
        }
        if (tries == 99) return 0;
        A += dA; B += dB; X0 = Xn; Y0 = Yn;
    }
    /* Jacobian d(r,phi)/d(alpha,beta) by 4th-order central differences.  Polar
     * outputs are used because frame dragging winds phi up quickly with
     * impact parameter near the ISCO of fast spins: in Cartesian (X,Y) the
     * determinant is a small difference of large products. */
    double hc = 1e-4 * (1 + hypot(A, B));
    double dr[2], dp[2];
    for (int attempt = 0;; attempt++) {
        int good = 1;
        for (int dir = 0; dir < 2 && good; dir++) {
            double rr[4], pp[4];
            static const double off[4] = {-2, -1, 1, 2};
            for (int i = 0; i < 4 && good; i++) {
                double aa = A + (dir == 0 ? off[i] * hc : 0), bb = B + (dir == 1 ? off[i] * hc : 0);
                good = ro_trace_ray(a, th, aa, bb, &rr[i], &pp[i]);
            }
            if (!good) break;
            double p0 = pp[1];
            for (int i = 0; i < 4; i++) pp[i] = remainder(pp[i] - p0, 2 * PI);
            dr[dir] = (rr[0] - 8 * rr[1] + 8 * rr[2] - rr[3]) / (12 * hc);
            dp[dir] = (pp[0] - 8 * pp[1] + 8 * pp[2] - pp[3]) / (12 * hc);
        }
        if (good) break;
        if (attempt >= 2) return 0;
        hc *= 0.05; /* near the shadow edge: shrink the stencil */
    }
    double det = r_e * (dr[0] * dp[1] - dr[1] * dp[0]);
    *al = A; *be = B;
    *g = kep_g(a, r_e, -A * sin(th));
    /* cosine of the local emission angle: mu_e = g sqrt(q) / r, q = beta^2 + cos^2(theta_o) (alpha^2 - a^2) */
    double q = B * B + cos(th) * cos(th) * (A * A - a * a);
    double mue = *g * sqrt(q > 0 ? q : 0.0) / r_e;
    *jn = ro_limb_factor(g_limb, mue) / fabs(det);
    return 1;
}

/* continuation from a solved point (r_from; A,B) to r_to at fixed phi:
 * try directly with the scaled guess; on failure bisect the radial step. */
static int march_to(double a, double th, double phi, double r_from, double r_to, double *A, double *B,
                    double *g, double *jn, int depth)
{
    double A0 = *A, B0 = *B;
    double f = r_to / r_from;
    double At = A0 * f, Bt = B0 * f;
    if (ro_solve_point(a, th, r_to, phi, &At, &Bt, g, jn)) { *A = At; *B = Bt; return 1; }
    if (depth >= 10) return 0;
    double r_mid = sqrt(r_from * r_to);
    double Am = A0, Bm = B0, gm, jm;
    if (!march_to(a, th, phi, r_from, r_mid, &Am, &Bm, &gm, &jm, depth + 1)) return 0;
    if (!march_to(a, th, phi, r_mid, r_to, &Am, &Bm, g, jn, depth + 1)) return 0;
    *A = Am; *B = Bm;
    return 1;
}

/* Table rows are the same physical radii for every node: uniform in x(r) = 1/r + C/sqrt(r) from r_in to
 * RO_TAB_ROUT (C = 1).  g and Jn of a Keplerian disc are close to low-order polynomials of r^(-1/2), which keeps
 * the linear row interpolation accurate with few rows; the 1/r term packs the rows closer where the relativistic
 * terms are steep (near the ISCO of fast spins).  Row s: x_s = x(r_in) + s (x(r_out) - x(r_in)) / (NR - 1). */
#define RO_TAB_ROUT 1000.0
#define RO_ROW_C 1.0
static double row_x(double r) { return 1.0 / r + RO_ROW_C / sqrt(r); }
double ro_row_radius(double r_in, int NR, int s)
{
    double x0 = row_x(r_in), x1 = row_x(RO_TAB_ROUT);
    double x = x0 + s * ((x1 - x0) / (NR - 1));
    if (s == 0) return r_in;
    if (s == NR - 1) return RO_TAB_ROUT;
    double u = 0.5 * (-RO_ROW_C + sqrt(RO_ROW_C * RO_ROW_C + 4 * x));
    return 1.0 / (u * u);
}

/* disc azimuth of the "front" ray (alpha = 0, beta < 0) landing on radius r_e:
 * the per-row azimuth origin.  Frame dragging winds the Boyer-Lindquist phi of
 * the whole image pattern with radius; measuring the table azimuth from this
 * ray removes the bulk winding so rows interpolate smoothly at fixed column.
 * (The ring integral is invariant under a per-row azimuth offset.)
 * beta_f: in = guess (<0), out = solution. */
static int front_ray(double a, double th, double r_e, double *beta_f, double *phi0)
{
    double r, p;
    double b0 = *beta_f, b1;
    double f0, f1;
    /* bracket: f(beta) = r_hit(0,beta) - r_e increases with |beta| */
    if (!ro_trace_ray(a, th, 0.0, b0, &r, &p)) f0 = -r_e; else f0 = r - r_e;
    double step = 0.05 * (1 + fabs(b0));
    b1 = b0;
    f1 = f0;
    for (int i = 0; i < 200 && f0 * f1 > 0; i++) {
        b0 = b1; f0 = f1;
        b1 = f1 > 0 ? b1 + step : b1 - step; /* beta<0: larger |beta| -> larger r */
        if (b1 > -1e-9) b1 = -1e-9;
        if (!ro_trace_ray(a, th, 0.0, b1, &r, &p)) f1 = -r_e; else f1 = r - r_e;
        step *= 1.5;
    }
    if (f0 * f1 > 0) return 0;
    for (int i = 0; i < 200; i++) {
        double bm = 0.5 * (b0 + b1), fm;
        if (fabs(f1 - f0) > 0 && f0 != -r_e && f1 != -r_e) { /* regula falsi step, guarded */
            double bs = b1 - f1 * (b1 - b0) / (f1 - f0);
            if ((bs - b0) * (bs - b1) < 0 && (i & 3) != 3) bm = bs;
        }
        if (!ro_trace_ray(a, th, 0.0, bm, &r, &p)) fm = -r_e; else fm = r - r_e;
        if (fm == 0 || fabs(b1 - b0) < 1e-13 * (1 + fabs(bm))) { b0 = b1 = bm; f0 = f1 = fm; break; }
        if (fm * f0 < 0) { b1 = bm; f1 = fm; } else { b0 = bm; f0 = fm; }
    }
    double bm = 0.5 * (b0 + b1);
    if (!ro_trace_ray(a, th, 0.0, bm, &r, &p)) return 0;
    if (fabs(r - r_e) > 1e-8 * r_e) return 0;
    *beta_f = bm; *phi0 = p;
    return 1;
}

typedef struct { double a, theta_o, co; int NR, NPHI; float *g, *jn; int *fail; const double *phi0; const double *rows; double r_valid; double *solA, *solB; } node_ctx;

static void node_body(int t, void *vc)
{
    node_ctx *c = (node_ctx *)vc;
    double a = c->a, theta_o = c->theta_o, co = c->co;
    int NR = c->NR, NPHI = c->NPHI;
    float *g = c->g, *jn = c->jn;
    int nfail = 0;
    double A1 = 0, B1 = 0, A2 = 0, B2 = 0, r1 = 0; /* solutions at rows s+1, s+2 */
    int have = 0;
    for (int s = NR - 1; s >= 0; s--) {
        double r = c->rows[s];
        double phi = 2 * PI * t / NPHI + c->phi0[s];
        double A, B, gg = 1, jj = co;
        int ok = 0;
        c->solA[(size_t)s * NPHI + t] = NAN; c->solB[(size_t)s * NPHI + t] = NAN;
        if (!(r > c->r_valid)) { /* no circular orbit (r <= photon orbit): the row carries no data */
            g[(size_t)s * NPHI + t] = 0.0f;
            jn[(size_t)s * NPHI + t] = 0.0f;
            continue;
        }
        if (have >= 2) { A = 2 * A1 - A2; B = 2 * B1 - B2; ok = ro_solve_point(a, theta_o, r, phi, &A, &B, &gg, &jj); }
        if (!ok && have >= 1) { /* rotate the previous solution by the change of azimuth origin, then march */
            double dph = c->phi0[s] - c->phi0[s + 1];
            (void)dph;
            A = A1; B = B1;
            ok = march_to(a, theta_o, phi, r1, r, &A, &B, &gg, &jj, 0);
        }
        if (!ok) { /* flat-space guess (observer convention: alpha = -r sin phi, beta = -r cos phi cos theta_o) */
            A = -r * sin(phi); B = -r * cos(phi) * co;
            ok = ro_solve_point(a, theta_o, r, phi, &A, &B, &gg, &jj);
        }
        if (!ok) {
            /* unresolved point (direct image folds into the photon-ring structure at the
             * shadow edge: a ~ 0.998, i > ~80 deg, r < ~2): carries no weight. */
            nfail++;
            if (s + 1 < NR) gg = g[(size_t)(s + 1) * NPHI + t];
            g[(size_t)s * NPHI + t] = (float)gg;
            jn[(size_t)s * NPHI + t] = 0.0f;
            continue; /* keep the last good solution for continuation */
        }
        g[(size_t)s * NPHI + t] = (float)gg;
        jn[(size_t)s * NPHI + t] = (float)jj;
        c->solA[(size_t)s * NPHI + t] = A; c->solB[(size_t)s * NPHI + t] = B;
        if (have >= 1) { A2 = A1; B2 = B1; }
        A1 = A; B1 = B; r1 = r;
        have = have < 2 ? have + 1 : 2;
    }
    c->fail[t] = nfail;
}

/* innermost radius a table row may have for spin a: a little outside the photon orbit
 * r_ph = 2 (1 + cos(2/3 acos(-a))), inside which no circular orbit exists (the Keplerian g is imaginary) */
double ro_r_valid(double a)
{
    double rph = 2 * (1 + cos(2.0 / 3.0 * acos(-a)));
    return 1.02 * rph;
}

int ro_node_phi0(double a, double theta_o, int NR, const double *rows, double *phi0)
{
    double bf = 0;
    int bad = 0;
    for (int s = NR - 1; s >= 0; s--) {
        double r = rows[s];
        if (!(r > ro_r_valid(a))) { phi0[s] = s + 1 < NR ? phi0[s + 1] : 0; continue; }
        if (s == NR - 1) bf = -r * cos(theta_o);
        double p = 0;
        if (!front_ray(a, theta_o, r, &bf, &p)) { bad++; p = s + 1 < NR ? phi0[s + 1] : 0; }
        phi0[s] = p;
    }
    return bad;
}

/* continuation along the ring: from the solved point (r, phi_from; A, B) to (r, phi_to), halving the step on failure */
static int march_phi(double a, double th, double r, double phi_from, double phi_to, double *A, double *B,
                     double *g, double *jn, int depth)
{
    double At = *A, Bt = *B;
    if (ro_solve_point(a, th, r, phi_to, &At, &Bt, g, jn)) { *A = At; *B = Bt; return 1; }
    if (depth >= 8) return 0;
    double pm = 0.5 * (phi_from + phi_to);
    double Am = *A, Bm = *B, gm, jm;
    if (!march_phi(a, th, r, phi_from, pm, &Am, &Bm, &gm, &jm, depth + 1)) return 0;
    if (!march_phi(a, th, r, pm, phi_to, &Am, &Bm, g, jn, depth + 1)) return 0;
    *A = Am; *B = Bm;
    return 1;
}

typedef struct { node_ctx *c; int *list; int *fixed; } repair_ctx;

static void repair_body(int i, void *vc)
{
    repair_ctx *rc = (repair_ctx *)vc;
    node_ctx *c = rc->c;
    int NPHI = c->NPHI;
    int s = rc->list[i] / NPHI, t = rc->list[i] % NPHI;
    double r = c->rows[s], phi = 2 * PI * t / NPHI + c->phi0[s];
    for (int side = 0; side < 2 && !rc->fixed[i]; side++) {
        int tn = (t + (side ? 1 : NPHI - 1)) % NPHI;
        double A = c->solA[(size_t)s * NPHI + tn], B = c->solB[(size_t)s * NPHI + tn];
        if (A != A) continue;
        double phin = phi + (side ? 1 : -1) * 2 * PI / NPHI, gg, jj;
        if (march_phi(c->a, c->theta_o, r, phin, phi, &A, &B, &gg, &jj, 0)) {
            /* results are published after the sweep (the neighbours of this sweep must not depend on its order) */
            c->g[(size_t)s * NPHI + t] = (float)gg; c->jn[(size_t)s * NPHI + t] = (float)jj;
            rc->fixed[i] = 1;
            ((double *)c->solA)[(size_t)c->NR * NPHI + i] = A; ((double *)c->solB)[(size_t)c->NR * NPHI + i] = B;
        }
    }
}

/* optional extras of the last ro_make_node_rows call of this thread: image-plane coordinates and emission cosine */
static __thread double *g_extra_alpha = NULL, *g_extra_beta = NULL, *g_extra_mue = NULL;
static __thread int g_phi0_zero = 0;
void ro_node_extras(double *alpha, double *beta, double *mue, int absolute_phi)
{
    g_extra_alpha = alpha; g_extra_beta = beta; g_extra_mue = mue; g_phi0_zero = absolute_phi;
}

int ro_make_node_rows(double a, double theta_o, int NR, const double *rows, int NPHI, float *g, float *jn, int threads)
{
    int *fail = (int *)calloc(NPHI, sizeof(int));
    double *phi0 = (double *)calloc(NR, sizeof(double));
    size_t np = (size_t)NR * NPHI;
    double *solA = (double *)malloc(sizeof(double) * 2 * np), *solB = (double *)malloc(sizeof(double) * 2 * np);
    int nbad0 = ro_node_phi0(a, theta_o, NR, rows, phi0);
    if (g_phi0_zero) memset(phi0, 0, sizeof(double) * NR); /* columns at absolute azimuth (KY's own layout) */
    node_ctx c = {a, theta_o, cos(theta_o), NR, NPHI, g, jn, fail, phi0, rows, ro_r_valid(a), solA, solB};
    ro_par_for(NPHI, threads, node_body, &c);
    int nfail = 0;
    for (int t = 0; t < NPHI; t++) nfail += fail[t];
    /* repair sweeps: points the radial continuation lost are retried along the ring from solved neighbours */
    for (int sweep = 0; sweep < 6 && nfail > 0; sweep++) {
        int *list = (int *)malloc(sizeof(int) * np), *fixed = (int *)calloc(np, sizeof(int));
        int n = 0;
        for (size_t i = 0; i < np; i++)
            if (rows[i / NPHI] > c.r_valid && solA[i] != solA[i]) list[n++] = (int)i;
        repair_ctx rc = {&c, list, fixed};
        ro_par_for(n, threads, repair_body, &rc);
        int nfix = 0;
        for (int i = 0; i < n; i++)
            if (fixed[i]) { solA[list[i]] = solA[np + i]; solB[list[i]] = solB[np + i]; nfix++; }
        free(list); free(fixed);
        nfail = n - nfix;
        if (nfix == 0) break;
    }
    nfail += nbad0;
    if (g_extra_alpha || g_extra_beta || g_extra_mue) {
        for (size_t i = 0; i < np; i++) {
            double A = solA[i], B = solB[i], r = rows[i / NPHI];
            double q = B * B + cos(theta_o) * cos(theta_o) * (A * A - a * a);
            double gg = A == A ? kep_g(a, r, -A * sin(theta_o)) : 0.0;
            if (g_extra_alpha) g_extra_alpha[i] = A;
            if (g_extra_beta) g_extra_beta[i] = B;
            if (g_extra_mue) g_extra_mue[i] = A == A ? gg * sqrt(q > 0 ? q : 0.0) / r : 0.0;
        }
    }
    free(fail); free(phi0); free(solA); free(solB);
    return nfail;
}

int ro_make_node(double a, double theta_o, int NR, int NPHI, double r_in, float *g, float *jn, int threads)
{
    double *rows = (double *)malloc(sizeof(double) * NR);
    for (int s = 0; s < NR; s++) rows[s] = ro_row_radius(r_in, NR, s);
    int nfail = ro_make_node_rows(a, theta_o, NR, rows, NPHI, g, jn, threads);
    free(rows);
    return nfail;
}

/* ------------------------------------------------------------------ */
struct ro_tables {
    int n_spin, n_inc, NR, NPHI;
    double r_in;
    double *spins, *thetas, *rows;
    int *first_row;  /* [n_spin]: first row outside ro_r_valid(spin) */
    float *data;     /* [n_spin][n_inc][2][NR][NPHI]: g, Jn (the layout of the C ABI) */
    float *lg;       /* [n_spin][n_inc][NR][NPHI]: (float) log((double) g), 0 where g = 0 (ro_tables_finish) */
    int ready;
    long serial;
};
static long g_tab_serial = 0;

ro_tables *ro_tables_create(const double *spins, int n_spin, const double *thetas, int n_inc, int NR, int NPHI, double r_in)
{
    ro_tables *t = (ro_tables *)calloc(1, sizeof *t);
    t->n_spin = n_spin; t->n_inc = n_inc; t->NR = NR; t->NPHI = NPHI; t->r_in = r_in;
    t->spins = (double *)malloc(sizeof(double) * n_spin);
    t->thetas = (double *)malloc(sizeof(double) * n_inc);
    t->rows = (double *)malloc(sizeof(double) * NR);
    t->first_row = (int *)calloc(n_spin, sizeof(int));
    for (int i = 0; i < n_spin; i++) t->spins[i] = spins[i];
    for (int i = 0; i < n_inc; i++) t->thetas[i] = thetas[i];
    for (int s = 0; s < NR; s++) t->rows[s] = ro_row_radius(r_in, NR, s);
    t->data = (float *)calloc((size_t)n_spin * n_inc * 2 * NR * NPHI, sizeof(float));
    t->lg = (float *)calloc((size_t)n_spin * n_inc * NR * NPHI, sizeof(float));
    return t;
}

/* derived arrays; call after the data block has been filled or changed */
void ro_tables_finish(ro_tables *t)
{
    size_t plane = (size_t)t->NR * t->NPHI;
    for (int m = 0; m < t->n_spin; m++) {
        double rv = ro_r_valid(t->spins[m]);
        int f = 0;
        while (f < t->NR && !(t->rows[f] > rv)) f++;
        t->first_row[m] = f;
        for (int n = 0; n < t->n_inc; n++) {
            const float *g = t->data + ((size_t)m * t->n_inc + n) * 2 * plane;
            float *lg = t->lg + ((size_t)m * t->n_inc + n) * plane;
            for (size_t i = 0; i < plane; i++) lg[i] = g[i] > 0 ? (float)log((double)g[i]) : 0.0f;
        }
    }
    t->serial = __sync_add_and_fetch(&g_tab_serial, 1);
    t->ready = 1;
}

ro_tables *ro_tables_generate(const double *spins, int n_spin, const double *thetas, int n_inc, int NR, int NPHI, double r_in, int threads)
{
    ro_tables *t = ro_tables_create(spins, n_spin, thetas, n_inc, NR, NPHI, r_in);
    size_t plane = (size_t)NR * NPHI;
    for (int m = 0; m < n_spin; m++)
        for (int n = 0; n < n_inc; n++) {
            float *base = t->data + ((size_t)m * n_inc + n) * 2 * plane;
            ro_make_node(spins[m], thetas[n], NR, NPHI, r_in, base, base + plane, threads);
        }
    ro_tables_finish(t);
    return t;
}

void ro_tables_free(ro_tables *t)
{
    if (!t) return;
    free(t->spins); free(t->thetas); free(t->rows); free(t->first_row); free(t->data); free(t->lg); free(t);
}
float *ro_tables_data(ro_tables *t) { return t->data; }
long ro_tables_nfloats(const ro_tables *t) { return (long)t->n_spin * t->n_inc * 2 * t->NR * t->NPHI; }

/* ------------------------------------------------------------------ */
/* Lookup for a model (a, theta_o): Lagrange interpolation over up to 4 spin nodes (coordinate a) x up to 4
 * inclination nodes (coordinate theta) of ln g and Jn AT THE SAME PHYSICAL RADIUS (the rows are shared by all
 * nodes).  A spin node only enters where it has data: its rows inside ro_r_valid(a_node) are empty, so nodes
 * whose first row lies above the model's innermost row (the one bracketing risco(a) from below) are dropped
 * from the low end of the stencil.                                                                         */
static void lagrange_w(const double *xs, int n, double x, double *w)
{
    for (int i = 0; i < n; i++) {
        double p = 1.0;
        for (int k = 0; k < n; k++)
            if (k != i) p *= (x - xs[k]) / (xs[i] - xs[k]);
        w[i] = p;
    }
}

/* nodes [*i0, *i0 + *k) around x (clamped into the node range), k <= 4 */
static void stencil4(const double *nodes, int n, double x, int *i0, int *k)
{
    int i = 0;
    while (i + 2 < n && x >= nodes[i + 1]) i++;
    int lo = i - 1, hi = i + 2; /* inclusive */
    if (lo < 0) lo = 0;
    if (hi > n - 1) hi = n - 1;
    *i0 = lo; *k = hi - lo + 1;
}

double ro_row_pos(const ro_tables *t, double r)
{
    double x0 = row_x(t->r_in), x1 = row_x(RO_TAB_ROUT);
    double ell = (row_x(r) - x0) / ((x1 - x0) / (t->NR - 1));
    if (!(ell > 0)) ell = 0;
    if (ell > t->NR - 1) ell = t->NR - 1;
    return ell;
}

/* the model's slice: sl_lg[NR][NPHI] = ln g, sl_jn[NR][NPHI] = Jn, rows >= *s_lo filled.  returns 0 ok,
 * 5 when the table has no valid stencil for this spin (node grid too coarse near the top). */
int ro_slice(const ro_tables *t, double a, double theta_o, double *sl_lg, double *sl_jn, int *s_lo)
{
    int NR = t->NR, NP = t->NPHI;
    size_t plane = (size_t)NR * NP;
    double ac = a < t->spins[0] ? t->spins[0] : (a > t->spins[t->n_spin - 1] ? t->spins[t->n_spin - 1] : a);
    double thc = theta_o < t->thetas[0] ? t->thetas[0] : (theta_o > t->thetas[t->n_inc - 1] ? t->thetas[t->n_inc - 1] : theta_o);
    int need = (int)floor(ro_row_pos(t, ro_risco(a)));
    if (need > NR - 2) need = NR - 2;
    if (need < 0) need = 0;
    int m0, km, n0, kn;
    stencil4(t->spins, t->n_spin, ac, &m0, &km);
    while (km > 1 && t->first_row[m0] > need && ac >= t->spins[m0 + 1]) { m0++; km--; } /* drop empty low-spin nodes */
    for (int i = 0; i < km; i++)
        if (t->first_row[m0 + i] > need) return 5;
    stencil4(t->thetas, t->n_inc, thc, &n0, &kn);
    double wa[4], wi[4];
    lagrange_w(t->spins + m0, km, ac, wa);
    lagrange_w(t->thetas + n0, kn, thc, wi);
    for (int s = need; s < NR; s++)
        for (int p = 0; p < NP; p++) {
            size_t i = (size_t)s * NP + p;
            double lg = 0, jn = 0;
            for (int im = 0; im < km; im++)
                for (int in = 0; in < kn; in++) {
                    size_t node = (size_t)(m0 + im) * t->n_inc + (n0 + in);
                    double w = wa[im] * wi[in];
                    lg += w * (double)t->lg[node * plane + i];
                    jn += w * (double)t->data[node * 2 * plane + plane + i];
                }
            sl_lg[i] = lg; sl_jn[i] = jn;
        }
    *s_lo = need;
    return 0;
}

/* hat kernel average over [lo,hi] (in units of bins, relative to bin centre 0) */
static double hat_avg_piece(double lo, double hi, double L, double R, int neg)
{
    /* overlap of [lo,hi] with [L,R]; hat = 1+y on [-1,0], 1-y on [0,1] */
    double l = lo > L ? lo : L, r = hi < R ? hi : R;
    if (r <= l) return 0.0;
    double mid = 0.5 * (l + r);
    return (r - l) * (neg ? 1.0 + mid : 1.0 - mid);
}

static void deposit(double *H, int K_LO, int nH, double sa, double sb, double W)
{
    double lo = sa < sb ? sa : sb, hi = sa < sb ? sb : sa;
    double len = hi - lo;
    if (len < 1e-12) {
        double mid = 0.5 * (lo + hi);
        double kf = floor(mid);
        int k0 = (int)kf;
        double f = mid - kf;
        int i0 = k0 - K_LO;
        if (i0 >= 0 && i0 + 1 < nH) { H[i0] += W * (1 - f); H[i0 + 1] += W * f; }
        return;
    }
    int kfirst = (int)floor(lo), klast = (int)floor(hi) + 1;
    double inv = W / len;
    for (int k = kfirst; k <= klast; k++) {
        double c = hat_avg_piece(lo - k, hi - k, -1.0, 0.0, 1) + hat_avg_piece(lo - k, hi - k, 0.0, 1.0, 0);
        int i = k - K_LO;
        if (c != 0.0 && i >= 0 && i < nH) H[i] += inv * c;
    }
}

/* number of rings an annulus [r1,r2] is cut into: the radial variation of the shift across one ring is kept
 * below about half an energy bin.  |d ln g / d ln r| <= kappa with kappa = 1.5/(r-3) + 0.5/sqrt(r) (gravitational
 * + orbital Doppler term of a Keplerian disc), capped at 0.5 (the value at r = 6) and used as 0.5 for wide annuli
 * (the stand-alone kyconv seam) where one kappa does not describe the whole annulus. */
int ro_auto_nsub(double r1, double r2, double dlnE)
{
    double lnr = log(r2 / r1);
    double kappa = 0.5;
    if (lnr <= 0.2) {
        double rg = sqrt(r1 * r2);
        if (rg > 6.0) {
            kappa = 1.5 / (rg - 3.0) + 0.5 / sqrt(rg);
            if (kappa > 0.5) kappa = 0.5;
        }
    }
    /* + 1e-9: grids commensurate with the radial binning (e.g. 100 bins per decade against 50 annuli per decade)
     * put kappa lnr / dlnE on an integer; the bias keeps the floor independent of last-bit differences of log() */
    int nsub = 1 + (int)floor(kappa * lnr / dlnE + 1e-9);
    return nsub > 4096 ? 4096 : nsub;
}

/* per-thread cache of the last slice (the region drivers call the seam once per annulus of the same model) */
typedef struct { const ro_tables *t; long serial; double a, th; int s_lo, n; double *lg, *w; } slice_cache;
static __thread slice_cache g_sc;

static int get_slice(const ro_tables *t, double a, double theta_o, const double **lg, const double **w, int *s_lo)
{
    size_t n = (size_t)t->NR * t->NPHI;
    if (g_sc.t == t && g_sc.serial == t->serial && g_sc.a == a && g_sc.th == theta_o && g_sc.n == (int)n) {
        *lg = g_sc.lg; *w = g_sc.w; *s_lo = g_sc.s_lo;
        return 0;
    }
    if (g_sc.n != (int)n) {
        free(g_sc.lg); free(g_sc.w);
        g_sc.lg = (double *)malloc(sizeof(double) * n); g_sc.w = (double *)malloc(sizeof(double) * n);
        g_sc.n = (int)n;
    }
    g_sc.t = NULL;
    int st = ro_slice(t, a, theta_o, g_sc.lg, g_sc.w, &g_sc.s_lo);
    if (st) return st;
    /* weight density w = Jn g^3 per slice sample (a Lagrange overshoot below zero carries no weight) */
    for (size_t i = (size_t)g_sc.s_lo * t->NPHI; i < n; i++) {
        double jn = g_sc.w[i] > 0 ? g_sc.w[i] : 0.0;
        g_sc.w[i] = jn * exp(3.0 * g_sc.lg[i]);
    }
    g_sc.t = t; g_sc.serial = t->serial; g_sc.a = a; g_sc.th = theta_o;
    *lg = g_sc.lg; *w = g_sc.w; *s_lo = g_sc.s_lo;
    return 0;
}

int ro_ky_kernel(const ro_tables *t, double a, double theta_o, double r1, double r2,
                 double alpha, double beta, double rb, double zshift, double dlnE, int nsub,
                 double *Hc, int cap, int *kmin, int *nk)
{
    if (!t || !t->ready || !(r2 > r1) || !(dlnE > 0)) return 1;
    int NR = t->NR, NP = t->NPHI;
    const double *LG, *WT;
    int s_lo;
    int st = get_slice(t, a, theta_o, &LG, &WT, &s_lo);
    if (st) return st;
    double inv_dlnE = 1.0 / dlnE;
    double lnr = log(r2 / r1);
    if (nsub <= 0) nsub = ro_auto_nsub(r1, r2, dlnE);
    int K_LO = (int)floor(log(0.005) / dlnE) - 2, K_HI = (int)ceil(log(4.0) / dlnE) + 2;
    double zs = 0;
    if (zshift != 0) zs = -log(1 + zshift) / dlnE;
    K_LO += (int)floor(zs) - 1; K_HI += (int)ceil(zs) + 1;
    int nH = K_HI - K_LO + 1;
    double *H = (double *)calloc(nH, sizeof(double));
    double *sv = (double *)malloc(sizeof(double) * NP), *wv = (double *)malloc(sizeof(double) * NP);
    double dphi = 2 * PI / NP;
    for (int m = 0; m < nsub; m++) {
        double ra = r1 * exp(lnr * m / nsub), rbb = r1 * exp(lnr * (m + 1) / nsub);
        if (m == 0) ra = r1;
        if (m == nsub - 1) rbb = r2;
        double rm = sqrt(ra * rbb);
        double emis = 1.0;
        if (alpha != 0 || beta != 0) emis = rm < rb ? pow(rm, -alpha) : pow(rm, -beta) * pow(rb, beta - alpha);
        double A = 0.5 * (rbb * rbb - ra * ra) * emis;
        double ell = ro_row_pos(t, rm);
        int s0 = (int)floor(ell);
        if (s0 > NR - 2) s0 = NR - 2;
        if (s0 < s_lo) s0 = s_lo; /* (annuli start at risco: only a caller's r_in below the slice gets here) */
        double fr = ell - s0;
        if (fr < 0) fr = 0;
        int s1 = s0 + 1;
        for (int p = 0; p < NP; p++) {
            size_t i0 = (size_t)s0 * NP + p, i1 = (size_t)s1 * NP + p;
            sv[p] = (LG[i0] + fr * (LG[i1] - LG[i0])) * inv_dlnE + zs;
            wv[p] = WT[i0] + fr * (WT[i1] - WT[i0]);
        }
        for (int p = 0; p < NP; p++) {
            int p1 = p + 1 == NP ? 0 : p + 1;
            double W = 0.5 * (wv[p] + wv[p1]) * dphi * A;
            deposit(H, K_LO, nH, sv[p], sv[p1], W);
        }
    }
    free(sv); free(wv);
    int lo = 0, hi = nH - 1;
    while (lo < nH && H[lo] == 0) lo++;
    while (hi >= lo && H[hi] == 0) hi--;
    if (lo > hi) { *kmin = 0; *nk = 0; free(H); return 0; }
    int n = hi - lo + 1;
    if (n > cap) { free(H); *nk = n; return 2; }
    memcpy(Hc, H + lo, sizeof(double) * n);
    *kmin = K_LO + lo; *nk = n;
    free(H);
    return 0;
}

int ro_kyconv(const ro_tables *t, const double *eb, int nE, const double *par, double *flux)
{
    if (!t || nE < 2) return 1;
    if (par[3] != 0 || par[11] != -1) return 3; /* ms, norm: unsupported modes */
    if (par[9] != 0 && par[9] != -1 && par[9] != -2) return 3; /* limb law: a property of the table set (ro_set_limb) */
    double dlnE = log(eb[nE] / eb[0]) / nE;
    for (int j = 0; j <= nE; j++)
        if (fabs(log(eb[j] / eb[0]) - j * dlnE) > 1e-4 * dlnE) return 4; /* not a geometric grid */
    double th = par[1] * (PI / 180.0);
    int cap = (int)(log(4.0 / 0.005) / dlnE) + 64;
    double *Hc = (double *)malloc(sizeof(double) * cap);
    int kmin, nk;
    int st = ro_ky_kernel(t, par[0], th, par[2], par[4], par[5], par[6], par[7], par[8], dlnE, 0, Hc, cap, &kmin, &nk);
    if (st) { free(Hc); return st; }
    double *in = (double *)malloc(sizeof(double) * nE);
    memcpy(in, flux, sizeof(double) * nE);
    for (int j = 0; j < nE; j++) {
        double s = 0;
        for (int i = 0; i < nk; i++) {
            int src = j - (kmin + i);
            if (src >= 0 && src < nE) s += Hc[i] * in[src];
        }
        flux[j] = s;
    }
    free(in); free(Hc);
    return 0;
}
