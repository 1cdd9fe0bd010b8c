// rg_raytrace.cu -- device Kerr ray tracer: generates the transfer tables the
// reference obtains from HEASOFT's KYCONV FITS files (not in the reference
// repository; the seam is relagn.py:979-1001).  Published formulation: Dovciak,
// Karas & Yaqoob 2004, ApJS 153, 205 -- transfer functions g (energy shift) and
// the lensing/area factor of a Keplerian equatorial disc seen by a distant
// observer at inclination theta_o around a Kerr hole of spin a.
//
// Per table node (a, theta_o):  rows = the same physical radii for every node,
// uniform in x(r) = 1/r + 1/sqrt(r) from r_in to 1000 (rows inside 1.02 x the
// photon orbit of the node's spin hold no circular orbit and stay empty),
// columns phi_t = 2 pi t / NPHI + phi0(r_s) where phi0 is the disc azimuth hit
// by the ray alpha = 0, beta < 0 (removes the frame-dragging winding between
// rows).  For every (s, t) the image-plane point (alpha, beta) whose null
// geodesic lands on (r_s, phi_t) is found by Newton iteration on a
// Dormand-Prince 5(4) integration of the geodesic in Mino time; stored are
//   g  = E_obs / E_em                for a Keplerian emitter,
//   Jn = |d(alpha,beta)/d(X,Y)|      image-plane area per disc-plane area.
//
// Parallelisation: one thread per node for the phi0 chain (rows are chained by
// continuation of the root bracket), then one thread per (node, column) that
// walks the rows outside-in with continuation of the Newton start, then repair
// sweeps: points the radial continuation lost (the far side of the innermost
// rows at high inclination, where the image is stretched along the shadow edge)
// are retried along the ring from their solved neighbours, one thread per point.  The work
// is a few seconds once per table set; the result is cached on disk by
// relagn_b200/tables.py.
#include "rg_kernels.h"
#include <vector>

#define NV 9

struct Geo { double a, lam, q, c1, c2, c3, a2, ql; };

__device__ static void geo_rhs(const Geo &G, const double *y, double *f)
{
    // Mino-time system in u = 1/r, mu = cos(theta), (X, Y) = sin(theta)(cos, sin)(phi_geodesic), phi_fd:
    //   u''  = u (c1 + u (c2 + u c3)),  mu'' = mu (c1 - 2 a^2 mu^2),  X'' = -(q + lam^2 + 2 a^2 mu^2) X, same for Y,
    //   phi_fd' = a u (2 - a lam u) / (1 - 2u + a^2 u^2).   Regular through turning points and the polar axis.
    double u = y[0], mu = y[2];
    double kap = -(G.ql + 2 * G.a2 * mu * mu);
    f[0] = y[1];
    f[1] = u * (G.c1 + u * (G.c2 + u * G.c3));
    f[2] = y[3];
    f[3] = mu * (G.c1 - 2 * G.a2 * mu * mu);
    f[4] = y[5];
    f[5] = kap * y[4];
    f[6] = y[7];
    f[7] = kap * y[6];
    f[8] = G.a * u * (2 - G.a * G.lam * u) / (1 - 2 * u + G.a2 * u * u);
}

// rhs with mu as the independent variable (to land exactly on the equatorial plane)
__device__ static void geo_rhs_mu(const Geo &G, const double *y, double *f)
{
    double t[NV];
    geo_rhs(G, y, t);
    double iw = 1.0 / t[2];
    for (int i = 0; i < NV; i++) f[i] = t[i] * iw;
    f[2] = 1.0;
}

template <bool MU>
__device__ static double dopri_step(const Geo &G, const double *y, const double *k1, double h, double *ynew, double *k7,
                                    double rtol, double atol)
{
    double k2[NV], k3[NV], k4[NV], k5[NV], k6[NV], t[NV];
    int i;
    for (i = 0; i < NV; i++) t[i] = y[i] + h * (1.0 / 5) * k1[i];
    if (MU) geo_rhs_mu(G, t, k2); else geo_rhs(G, t, k2);
    for (i = 0; i < NV; i++) t[i] = y[i] + h * (3.0 / 40 * k1[i] + 9.0 / 40 * k2[i]);
    if (MU) geo_rhs_mu(G, t, k3); else geo_rhs(G, t, k3);
    for (i = 0; i < NV; i++) t[i] = y[i] + h * (44.0 / 45 * k1[i] - 56.0 / 15 * k2[i] + 32.0 / 9 * k3[i]);
    if (MU) geo_rhs_mu(G, t, k4); else geo_rhs(G, t, k4);
    for (i = 0; i < NV; i++)
        t[i] = y[i] + h * (19372.0 / 6561 * k1[i] - 25360.0 / 2187 * k2[i] + 64448.0 / 6561 * k3[i] - 212.0 / 729 * k4[i]);
    if (MU) geo_rhs_mu(G, t, k5); else geo_rhs(G, t, k5);
    for (i = 0; i < NV; i++)
        t[i] = y[i] + h * (9017.0 / 3168 * k1[i] - 355.0 / 33 * k2[i] + 46732.0 / 5247 * k3[i] + 49.0 / 176 * k4[i] -
                           5103.0 / 18656 * k5[i]);
    if (MU) geo_rhs_mu(G, t, k6); else geo_rhs(G, t, k6);
    for (i = 0; i < NV; i++)
        ynew[i] = y[i] + h * (35.0 / 384 * k1[i] + 500.0 / 1113 * k3[i] + 125.0 / 192 * k4[i] - 2187.0 / 6784 * k5[i] +
                              11.0 / 84 * k6[i]);
    if (MU) geo_rhs_mu(G, ynew, k7); else geo_rhs(G, ynew, k7);
    double err = 0;
    for (i = 0; i < NV; i++) {
        double e = h * (71.0 / 57600 * k1[i] - 71.0 / 16695 * k3[i] + 71.0 / 1920 * k4[i] - 17253.0 / 339200 * k5[i] +
                        22.0 / 525 * k6[i] - 1.0 / 40 * k7[i]);
        double sc = atol + rtol * fmax(fabs(y[i]), fabs(ynew[i]));
        double r = e / sc;
        err += r * r;
    }
    return sqrt(err / NV);
}

// one ray from the observer's image plane (alpha, beta) to its first crossing of the equatorial plane
__device__ static int trace_ray(double a, double theta_o, double alpha, double beta, double *r_hit, double *phi_hit)
{
    const double rtol = 1e-12, atol = 1e-14;
    double so = sin(theta_o), co = cos(theta_o);
    Geo G;
    G.a = a; G.a2 = a * a;
    G.lam = -alpha * so;
    G.q = beta * beta + co * co * (alpha * alpha - a * a);
    G.ql = G.q + G.lam * G.lam;
    G.c1 = G.a2 - G.lam * G.lam - G.q;
    G.c2 = 3 * ((a - G.lam) * (a - G.lam) + G.q);
    G.c3 = -2 * G.a2 * G.q;
    double uh = 1.0 / (1.0 + sqrt(1.0 - a * a));
    double y[NV] = {0.0, 1.0, co, beta * so, so, -beta * co, 0.0, -alpha, 0.0};
    double k1[NV], k7[NV], yn[NV];
    geo_rhs(G, y, k1);
    double h = 1e-3 / (1.0 + hypot(alpha, beta));
    for (int step = 0; step < 200000; step++) {
        double err = dopri_step<false>(G, y, k1, h, yn, k7, rtol, atol);
        if (!(err <= 1.0)) { // rejected step (also NaN)
            double fac = 0.9 * pow(err, -0.2);
            if (!(fac > 0.1)) fac = 0.1;
            h *= fac;
            if (h < 1e-300) return 0;
            continue;
        }
        if (yn[0] >= uh * (1 - 1e-9)) return 0; // fell through the horizon
        if (yn[2] <= 0.0) {
            // the plane is crossed inside this step: integrate in mu from y[2] down to 0
            const int NS = 4;
            double ys[NV], kk[NV], kn[NV], yt[NV];
            for (int i = 0; i < NV; i++) ys[i] = y[i];
            double dmu = -y[2] / NS;
            for (int s = 0; s < NS; s++) {
                geo_rhs_mu(G, ys, kk);
                dopri_step<true>(G, ys, kk, dmu, yt, kn, rtol, atol);
                for (int i = 0; i < NV; i++) ys[i] = yt[i];
            }
            if (ys[0] >= uh * (1 - 1e-9) || !(ys[0] > 0)) return 0;
            *r_hit = 1.0 / ys[0];
            *phi_hit = atan2(ys[6], ys[4]) + ys[8];
            return 1;
        }
        for (int i = 0; i < NV; i++) { y[i] = yn[i]; k1[i] = k7[i]; }
        double fac = err > 0 ? 0.9 * pow(err, -0.2) : 5.0;
        if (fac > 5.0) fac = 5.0;
        if (fac < 0.2) fac = 0.2;
        h *= fac;
    }
    return 0;
}

// energy shift for a Keplerian emitter at r, photon angular momentum lam
__device__ static double kep_g(double a, double r, double lam)
{
    double sr = sqrt(r), r32 = r * sr;
    return sqrt(sr) * sr * sqrt(r32 - 3 * sr + 2 * a) / (r32 + a - lam);
}

__device__ static int ray_xy(double a, double th, double al, double be, double *X, double *Y)
{
    double r, ph;
    if (!trace_ray(a, th, al, be, &r, &ph)) return 0;
    *X = r * cos(ph); *Y = r * sin(ph);
    return 1;
}

// Newton solve for the image-plane point that lands on (r_e, phi_e); then the Jacobian by 4th-order central
// differences in polar outputs (the Cartesian determinant cancels badly where frame dragging winds phi).
__device__ static int solve_point(double a, double th, double r_e, double phi_e, double *al, double *be, double *g,
                                  double *jn)
{
    double tx = r_e * cos(phi_e), ty = r_e * sin(phi_e);
    double A = *al, B = *be;
    double X0, Y0;
    if (!ray_xy(a, th, A, B, &X0, &Y0)) return 0;
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
        tries = 0;
        while (hypot(Xn - tx, Yn - ty) > hypot(rx, ry) * (1 + 1e-3) && tries < 8) {
            dA *= 0.5; dB *= 0.5; tries++;
            if (!ray_xy(a, th, A + dA, B + dB, &Xn, &Yn)) { tries = 99; break; }
        }
        if (tries == 99) return 0;
        A += dA; B += dB; X0 = Xn; Y0 = Yn;
    }
    double hc = 1e-4 * (1 + hypot(A, B));
    double dr[2], dp[2];
    for (int attempt = 0;; attempt++) {
        int good = 1;
        for (int dir = 0; dir < 2 && good; dir++) {
            double rr[4], pp[4];
            const double off[4] = {-2, -1, 1, 2};
            for (int i = 0; i < 4 && good; i++) {
                double aa = A + (dir == 0 ? off[i] * hc : 0), bb = B + (dir == 1 ? off[i] * hc : 0);
                good = trace_ray(a, th, aa, bb, &rr[i], &pp[i]);
            }
            if (!good) break;
            double p0 = pp[1];
            for (int i = 0; i < 4; i++) pp[i] = remainder(pp[i] - p0, 2 * RGK_PI);
            dr[dir] = (rr[0] - 8 * rr[1] + 8 * rr[2] - rr[3]) / (12 * hc);
            dp[dir] = (pp[0] - 8 * pp[1] + 8 * pp[2] - pp[3]) / (12 * hc);
        }
        if (good) break;
        if (attempt >= 2) return 0;
        hc *= 0.05; // near the shadow edge: shrink the stencil
    }
    double det = r_e * (dr[0] * dp[1] - dr[1] * dp[0]);
    *al = A; *be = B;
    *g = kep_g(a, r_e, -A * sin(th));
    *jn = 1.0 / fabs(det);
    return 1;
}

// continuation from a solved point (r_from; A, B) to r_to at fixed phi: direct with the radially scaled guess,
// bisecting the radial step on failure (explicit stack instead of recursion; depth limit 10)
__device__ static int march_to(double a, double th, double phi, double r_from, double r_to, double *A, double *B,
                               double *g, double *jn)
{
    double tgt[24];
    int dep[24];
    int sp = 0;
    tgt[0] = r_to; dep[0] = 0; sp = 1;
    double cur = r_from, Ac = *A, Bc = *B;
    while (sp > 0) {
        sp--;
        double t = tgt[sp];
        int d = dep[sp];
        double f = t / cur;
        double At = Ac * f, Bt = Bc * f, gg, jj;
        if (solve_point(a, th, t, phi, &At, &Bt, &gg, &jj)) {
            cur = t; Ac = At; Bc = Bt; *g = gg; *jn = jj;
            continue;
        }
        if (d >= 10 || sp + 2 > 24) return 0;
        tgt[sp] = t; dep[sp] = d + 1; sp++;
        tgt[sp] = sqrt(cur * t); dep[sp] = d + 1; sp++;
    }
    *A = Ac; *B = Bc;
    return 1;
}

// limb-darkening law of the local emission (kyconv parameter 10; relagn.py:988 passes 0 = isotropic): -1 Laor (1991)
// 1 + 2.06 mu_e, -2 Haardt (1993) ln(1 + 1/mu_e), with mu_e = g sqrt(q) / r, q = beta^2 + cos^2(theta_o)(alpha^2 - a^2).
// The law multiplies the weight Jn of every table sample: a table set is built for one law.
__device__ static double limb_weight(int law, double a, double th, double r, double A, double B, double g)
{
    if (law == 0) return 1.0;
    const double co = cos(th);
    const double q = B * B + co * co * (A * A - a * a);
    const double mue = g * sqrt(q > 0 ? q : 0.0) / r;
    return law == -1 ? 1.0 + 2.06 * mue : log(1.0 + 1.0 / mue);
}

// continuation along the ring: from the solved point (r, phi_from; A, B) to (r, phi_to), halving the azimuth step on
// failure (explicit stack instead of recursion; depth limit 8) -- kerr_oracle.c: march_phi
__device__ static int march_phi(double a, double th, double r, double phi_from, double phi_to, double *A, double *B,
                                double *g, double *jn)
{
    double tgt[20];
    int dep[20];
    int sp = 0;
    tgt[0] = phi_to; dep[0] = 0; sp = 1;
    double cur = phi_from, Ac = *A, Bc = *B;
    while (sp > 0) {
        sp--;
        double t = tgt[sp];
        int d = dep[sp];
        double At = Ac, Bt = Bc, gg, jj;
        if (solve_point(a, th, r, t, &At, &Bt, &gg, &jj)) {
            cur = t; Ac = At; Bc = Bt; *g = gg; *jn = jj;
            continue;
        }
        if (d >= 8 || sp + 2 > 20) return 0;
        tgt[sp] = t; dep[sp] = d + 1; sp++;
        tgt[sp] = 0.5 * (cur + t); dep[sp] = d + 1; sp++;
    }
    *A = Ac; *B = Bc;
    return 1;
}

// the alpha = 0, beta < 0 ray that lands on radius r_e: beta by bracketing + guarded regula falsi / bisection
__device__ static int front_ray(double a, double th, double r_e, double *beta_f, double *phi0)
{
    double r, p;
    double b0 = *beta_f, b1, f0, f1;
    if (!trace_ray(a, th, 0.0, b0, &r, &p)) f0 = -r_e; else f0 = r - r_e;
    double step = 0.05 * (1 + fabs(b0));
    b1 = b0; f1 = f0;
    for (int i = 0; i < 200 && f0 * f1 > 0; i++) {
        b0 = b1; f0 = f1;
        b1 = f1 > 0 ? b1 + step : b1 - step;
        if (b1 > -1e-9) b1 = -1e-9;
        if (!trace_ray(a, th, 0.0, b1, &r, &p)) f1 = -r_e; else f1 = r - r_e;
        step *= 1.5;
    }
    if (f0 * f1 > 0) return 0;
    for (int i = 0; i < 200; i++) {
        double bm = 0.5 * (b0 + b1), fm;
        if (fabs(f1 - f0) > 0 && f0 != -r_e && f1 != -r_e) {
            double bs = b1 - f1 * (b1 - b0) / (f1 - f0);
            if ((bs - b0) * (bs - b1) < 0 && (i & 3) != 3) bm = bs;
        }
        if (!trace_ray(a, th, 0.0, bm, &r, &p)) fm = -r_e; else fm = r - r_e;
        if (fm == 0 || fabs(b1 - b0) < 1e-13 * (1 + fabs(bm))) { b0 = b1 = bm; f0 = f1 = fm; break; }
        if (fm * f0 < 0) { b1 = bm; f1 = fm; } else { b0 = bm; f0 = fm; }
    }
    double bm = 0.5 * (b0 + b1);
    if (!trace_ray(a, th, 0.0, bm, &r, &p)) return 0;
    if (fabs(r - r_e) > 1e-8 * r_e) return 0;
    *beta_f = bm; *phi0 = p;
    return 1;
}

// node n = m * n_inc + i
__global__ void phi0_kernel(const double *__restrict__ spins, const double *__restrict__ thetas, int n_spin, int n_inc,
                            int NR, const double *__restrict__ rows, const double *__restrict__ rvalid,
                            double *__restrict__ phi0, int *__restrict__ nbad)
{
    int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n_spin * n_inc) return;
    double a = spins[node / n_inc], th = thetas[node % n_inc];
    const double rv = rvalid[node / n_inc];
    double bf = 0;
    int bad = 0;
    double *out = phi0 + (size_t)node * NR;
    for (int s = NR - 1; s >= 0; s--) {
        double r = rows[s];
        if (!(r > rv)) { out[s] = s + 1 < NR ? out[s + 1] : 0; continue; }
        if (s == NR - 1) bf = -r * cos(th);
        double p = 0;
        if (!front_ray(a, th, r, &bf, &p)) { bad++; p = s + 1 < NR ? out[s + 1] : 0; }
        out[s] = p;
    }
    if (bad) atomicAdd(nbad, bad);
}

__global__ void __launch_bounds__(32)
node_kernel(const double *__restrict__ spins, const double *__restrict__ thetas, int n_spin, int n_inc, int NR, int NPHI,
            const double *__restrict__ rows, const double *__restrict__ rvalid, const double *__restrict__ phi0_all,
            float *__restrict__ data, double2 *__restrict__ sol, int limb)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_spin * n_inc * NPHI) return;
    int node = idx / NPHI, t = idx - node * NPHI;
    double a = spins[node / n_inc], theta_o = thetas[node % n_inc];
    const double rv = rvalid[node / n_inc];
    double co = cos(theta_o);
    const double *phi0 = phi0_all + (size_t)node * NR;
    size_t plane = (size_t)NR * NPHI;
    float *g = data + (size_t)node * 2 * plane, *jn = g + plane;
    double2 *so = sol + (size_t)node * plane;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    double A1 = 0, B1 = 0, A2 = 0, B2 = 0, r1 = 0;
    int have = 0;
    for (int s = NR - 1; s >= 0; s--) {
        double r = rows[s];
        double phi = 2 * RGK_PI * t / NPHI + phi0[s];
        double A, B, gg = 1, jj = co;
        int ok = 0;
        so[(size_t)s * NPHI + t] = make_double2(qnan, qnan);
        if (!(r > rv)) { // no circular orbit (r <= photon orbit): the row carries no data
            g[(size_t)s * NPHI + t] = 0.0f;
            jn[(size_t)s * NPHI + t] = 0.0f;
            continue;
        }
        if (have >= 2) { A = 2 * A1 - A2; B = 2 * B1 - B2; ok = solve_point(a, theta_o, r, phi, &A, &B, &gg, &jj); }
        if (!ok && have >= 1) {
            A = A1; B = B1;
            ok = march_to(a, theta_o, phi, r1, r, &A, &B, &gg, &jj);
        }
        if (!ok) { // flat-space guess: alpha = -r sin phi, beta = -r cos phi cos theta_o
            A = -r * sin(phi); B = -r * cos(phi) * co;
            ok = solve_point(a, theta_o, r, phi, &A, &B, &gg, &jj);
        }
        if (!ok) {
            // lost by the radial continuation: left to the repair sweeps; carries no weight if they fail too
            if (s + 1 < NR) gg = g[(size_t)(s + 1) * NPHI + t];
            g[(size_t)s * NPHI + t] = (float)gg;
            jn[(size_t)s * NPHI + t] = 0.0f;
            continue;
        }
        g[(size_t)s * NPHI + t] = (float)gg;
        jn[(size_t)s * NPHI + t] = (float)(jj * limb_weight(limb, a, theta_o, r, A, B, gg));
        so[(size_t)s * NPHI + t] = make_double2(A, B);
        if (have >= 1) { A2 = A1; B2 = B1; }
        A1 = A; B1 = B; r1 = r;
        have = have < 2 ? have + 1 : 2;
    }
}

// one repair sweep: every unresolved point with data (thread = point) is retried from its solved ring neighbours
// t - 1, then t + 1 (kerr_oracle.c: repair_body).  The solutions found are published by publish_kernel after the
// sweep, so the result does not depend on the order the points are visited in.
__global__ void __launch_bounds__(32)
repair_kernel(const double *__restrict__ spins, const double *__restrict__ thetas, int n_spin, int n_inc, int NR, int NPHI,
              const double *__restrict__ rows, const double *__restrict__ rvalid, const double *__restrict__ phi0_all,
              float *__restrict__ data, const double2 *__restrict__ sol, double2 *__restrict__ pending,
              int *__restrict__ counters, int limb)
{
    size_t plane = (size_t)NR * NPHI;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n_spin * n_inc * plane) return;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    pending[idx] = make_double2(qnan, qnan);
    int node = (int)(idx / plane);
    int s = (int)((idx - (size_t)node * plane) / NPHI), t = (int)(idx - (size_t)node * plane - (size_t)s * NPHI);
    const double r = rows[s];
    if (!(r > rvalid[node / n_inc])) return;
    const double2 *so = sol + (size_t)node * plane;
    if (so[(size_t)s * NPHI + t].x == so[(size_t)s * NPHI + t].x) return; // solved
    atomicAdd(counters, 1); // unresolved before this sweep
    double a = spins[node / n_inc], theta_o = thetas[node % n_inc];
    const double phi = 2 * RGK_PI * t / NPHI + phi0_all[(size_t)node * NR + s];
    float *g = data + (size_t)node * 2 * plane, *jn = g + plane;
    for (int side = 0; side < 2; side++) {
        const int tn = (t + (side ? 1 : NPHI - 1)) % NPHI;
        double A = so[(size_t)s * NPHI + tn].x, B = so[(size_t)s * NPHI + tn].y;
        if (A != A) continue;
        const double phin = phi + (side ? 1 : -1) * 2 * RGK_PI / NPHI;
        double gg, jj;
        if (march_phi(a, theta_o, r, phin, phi, &A, &B, &gg, &jj)) {
            g[(size_t)s * NPHI + t] = (float)gg;
            jn[(size_t)s * NPHI + t] = (float)(jj * limb_weight(limb, a, theta_o, r, A, B, gg));
            pending[idx] = make_double2(A, B);
            atomicAdd(counters + 1, 1); // repaired in this sweep
            return;
        }
    }
}

__global__ void publish_kernel(double2 *__restrict__ sol, const double2 *__restrict__ pending, size_t n)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n && pending[idx].x == pending[idx].x) sol[idx] = pending[idx];
}

// geometry of the table format on the host (the same expressions as oracle/kerr_oracle.c, host libm)
double rg_row_x(double r) { return 1.0 / r + RG_ROW_C / sqrt(r); }
double rg_row_radius(double r_in, int NR, int s)
{
    double x0 = rg_row_x(r_in), x1 = rg_row_x(RG_TAB_ROUT);
    double x = x0 + s * ((x1 - x0) / (NR - 1));
    if (s == 0) return r_in;
    if (s == NR - 1) return RG_TAB_ROUT;
    double u = 0.5 * (-RG_ROW_C + sqrt(RG_ROW_C * RG_ROW_C + 4 * x));
    return 1.0 / (u * u);
}
double rg_r_valid(double a)
{
    double rph = 2 * (1 + cos(2.0 / 3.0 * acos(-a)));
    return 1.02 * rph;
}

int rg_launch_raytrace(const double *h_spins, int n_spin, const double *h_thetas, int n_inc, int NR, int NPHI,
                       double r_in, int limb, float *d_data, long *n_unresolved, cudaStream_t st)
{
    int nodes = n_spin * n_inc;
    size_t plane = (size_t)NR * NPHI, npts = (size_t)nodes * plane;
    double *d_meta = nullptr, *d_phi0 = nullptr;
    double2 *d_sol = nullptr, *d_pend = nullptr;
    int *d_cnt = nullptr;
    std::vector<double> meta(2 * n_spin + n_inc + NR);
    for (int i = 0; i < n_spin; i++) { meta[i] = h_spins[i]; meta[n_spin + n_inc + i] = rg_r_valid(h_spins[i]); }
    for (int i = 0; i < n_inc; i++) meta[n_spin + i] = h_thetas[i];
    for (int s = 0; s < NR; s++) meta[2 * n_spin + n_inc + s] = rg_row_radius(r_in, NR, s);
    RG_CUDA_OK(cudaMalloc(&d_meta, meta.size() * sizeof(double)));
    RG_CUDA_OK(cudaMalloc(&d_phi0, (size_t)nodes * NR * sizeof(double)));
    RG_CUDA_OK(cudaMalloc(&d_sol, npts * sizeof(double2)));
    RG_CUDA_OK(cudaMalloc(&d_pend, npts * sizeof(double2)));
    RG_CUDA_OK(cudaMalloc(&d_cnt, 3 * sizeof(int)));
    RG_CUDA_OK(cudaMemcpyAsync(d_meta, meta.data(), meta.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    RG_CUDA_OK(cudaMemsetAsync(d_cnt, 0, 3 * sizeof(int), st));
    const double *d_spins = d_meta, *d_thetas = d_meta + n_spin, *d_rvalid = d_meta + n_spin + n_inc,
                 *d_rows = d_meta + 2 * n_spin + n_inc;
    // one thread per node, spread over the SMs (the row chain is serial)
    RG_LAUNCH(phi0_kernel, dim3(nodes), dim3(1), 0, st, d_spins, d_thetas, n_spin, n_inc, NR, d_rows, d_rvalid, d_phi0,
              d_cnt + 2);
    RG_CUDA_OK(cudaGetLastError());
    int total = nodes * NPHI;
    RG_LAUNCH(node_kernel, dim3((total + 31) / 32), dim3(32), 0, st, d_spins, d_thetas, n_spin, n_inc, NR, NPHI, d_rows,
              d_rvalid, d_phi0, d_data, d_sol, limb);
    RG_CUDA_OK(cudaGetLastError());
    int cnt[3] = {0, 0, 0};
    long left = 0;
    for (int sweep = 0; sweep < 6; sweep++) {
        RG_CUDA_OK(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), st));
        RG_LAUNCH(repair_kernel, dim3((unsigned)((npts + 31) / 32)), dim3(32), 0, st, d_spins, d_thetas, n_spin, n_inc, NR,
                  NPHI, d_rows, d_rvalid, d_phi0, d_data, d_sol, d_pend, d_cnt, limb);
        RG_CUDA_OK(cudaGetLastError());
        RG_LAUNCH(publish_kernel, dim3((unsigned)((npts + 255) / 256)), dim3(256), 0, st, d_sol, d_pend, npts);
        RG_CUDA_OK(cudaGetLastError());
        RG_CUDA_OK(cudaMemcpyAsync(cnt, d_cnt, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
        RG_CUDA_OK(cudaStreamSynchronize(st));
        left = (long)cnt[0] - cnt[1];
        if (cnt[0] == 0 || cnt[1] == 0) break;
    }
    if (n_unresolved) *n_unresolved = left + cnt[2];
    cudaFree(d_meta); cudaFree(d_phi0); cudaFree(d_sol); cudaFree(d_pend); cudaFree(d_cnt);
    return RG_OK;
}
