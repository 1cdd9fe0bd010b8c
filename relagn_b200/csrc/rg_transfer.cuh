// rg_transfer.cuh -- device code of the relativistic transfer (the KYCONV-style
// seam, relagn.py:979-1001): the table slice of one (spin, inclination), ring
// evaluation and the g-shift histogram.  Shared by hist_kernel (every annulus of
// every parameter set) and the stand-alone kyconv entry point.
//
// Discretisation (identical to oracle/kerr_oracle.c):
//   - the tables hold, per (spin, inclination) node, g and Jn = dalpha dbeta / (dX dY) on rows of the SAME physical
//     radii for every node (uniform in x = 1/r + 1/sqrt(r)) and columns phi_t = 2 pi t / NPHI + phi0(r);
//   - the model's slice is a Lagrange blend over up to 4 spin nodes (coordinate a) x up to 4 inclination nodes
//     (coordinate theta) of ln g and Jn -- built ONCE per parameter set into shared memory as (ln g, w = Jn g^3);
//   - an annulus [r1,r2] is cut into n_sub rings at log-midpoints r_m with exact area weights A_m = (r_b^2 - r_a^2)/2;
//     ring m interpolates the slice linearly between its two rows: s_t = ln g_t / dlnE (shift in bins), w_t;
//   - the segment t..t+1 carries W = (w_t + w_{t+1})/2 dphi A_m spread uniformly in s over [s_t, s_{t+1}] and is
//     deposited on integer shifts with a linear (hat) kernel.
#pragma once
#include "rg_kernels.h"

#define RG_FULL 0xffffffffu
#define RG_IMAX 0x7fffffff
#define RG_MAX_STENCIL 16

// the nodes and weights of one (spin, inclination): lives in shared memory, built by one thread
struct TabSel {
    const float2 *plane[RG_MAX_STENCIL]; // node planes [NR][NPHI] of (ln g, Jn)
    double w[RG_MAX_STENCIL];
    int n;      // nodes in the stencil
    int s_lo;   // first row of the slice (the row bracketing risco(a) from below)
    int status; // 0 ok, 1: the node grid has no valid spin stencil here
    int NR, NPHI;
    double x0, dx; // row position (x(r) - x0) / dx
    double dphi;   // 2 pi / NPHI
};

// the selection of one parameter set as tabsel_kernel leaves it for hist_kernel (a multiple of 8 bytes)
struct SelRec { TabSel ts; int s_hi, pad; };

__device__ inline void rg_lagrange_w(const double *xs, int n, double x, double *w)
{
    for (int i = 0; i < n; i++) {
        double p = 1.0;
        for (int k = 0; k < n; k++)
            if (k != i) p *= (x - xs[k]) / (xs[i] - xs[k]);
        w[i] = p;
    }
}

// nodes [i0, i0 + k) around x, k <= 4 (kerr_oracle.c: stencil4)
__device__ inline void rg_stencil4(const double *nodes, int n, double x, int &i0, int &k)
{
    int i = 0;
    while (i + 2 < n && x >= nodes[i + 1]) i++;
    int lo = i - 1, hi = i + 2;
    if (lo < 0) lo = 0;
    if (hi > n - 1) hi = n - 1;
    i0 = lo; k = hi - lo + 1;
}

__device__ inline double rg_row_pos(double x0, double dx, int NR, double r)
{
    double ell = ((1.0 / r + RG_ROW_C / sqrt(r)) - x0) / dx;
    if (!(ell > 0)) ell = 0;
    if (ell > NR - 1) ell = NR - 1;
    return ell;
}

__device__ inline void rg_tab_select(TabSel &ts, const TabDesc &t, double a, double theta_o)
{
    const double ac = a < t.spins[0] ? t.spins[0] : (a > t.spins[t.n_spin - 1] ? t.spins[t.n_spin - 1] : a);
    const double thc = theta_o < t.thetas[0] ? t.thetas[0]
                                             : (theta_o > t.thetas[t.n_inc - 1] ? t.thetas[t.n_inc - 1] : theta_o);
    int need = (int)floor(rg_row_pos(t.x0, t.dx, t.NR, rg_risco(a)));
    if (need > t.NR - 2) need = t.NR - 2;
    if (need < 0) need = 0;
    int m0, km, n0, kn;
    rg_stencil4(t.spins, t.n_spin, ac, m0, km);
    while (km > 1 && t.first_row[m0] > need && ac >= t.spins[m0 + 1]) { m0++; km--; } // drop empty low-spin nodes
    ts.status = 0;
    for (int i = 0; i < km; i++)
        if (t.first_row[m0 + i] > need) ts.status = 1;
    rg_stencil4(t.thetas, t.n_inc, thc, n0, kn);
    double wa[4], wi[4];
    rg_lagrange_w(t.spins + m0, km, ac, wa);
    rg_lagrange_w(t.thetas + n0, kn, thc, wi);
    const size_t plane = (size_t)t.NR * t.NPHI;
    int n = 0;
    for (int im = 0; im < km; im++)
        for (int in = 0; in < kn; in++) {
            ts.plane[n] = t.pairs + ((size_t)(m0 + im) * t.n_inc + (n0 + in)) * plane;
            ts.w[n] = wa[im] * wi[in];
            n++;
        }
    ts.n = n;
    ts.s_lo = need;
    ts.NR = t.NR; ts.NPHI = t.NPHI; ts.x0 = t.x0; ts.dx = t.dx;
    ts.dphi = 2 * RGK_PI / t.NPHI;
}

// The model's slice, rows [ts.s_lo, s_hi], as (ln g, w = max(Jn, 0) g^3) pairs in shared memory: slice[(s - s_lo) NPHI + p].
// Called by all `nthreads` threads of the CTA (tid = its index); the caller synchronises afterwards.
__device__ inline void rg_build_slice(const TabSel &ts, int s_hi, double2 *slice, int tid, int nthreads)
{
    const int NP = ts.NPHI;
    const int first = ts.s_lo * NP, count = (s_hi - ts.s_lo + 1) * NP;
    const int n = ts.n;
    for (int i = tid; i < count; i += nthreads) {
        double lg = 0.0, jn = 0.0;
#pragma unroll 4
        for (int k = 0; k < n; k++) {
            const float2 v = __ldg(ts.plane[k] + first + i);
            const double w = ts.w[k];
            lg += w * (double)v.x;
            jn += w * (double)v.y;
        }
        if (!(jn > 0)) jn = 0.0; // a Lagrange overshoot below zero carries no weight
        slice[i] = make_double2(lg, jn * exp(3.0 * lg));
    }
}

// Number of rings an annulus [r1, r2] is cut into: the radial variation of the shift across one ring stays below
// about half an energy bin.  |d ln g / d ln r| <= kappa = 1.5/(r-3) + 0.5/sqrt(r) (gravitational + orbital Doppler
// term of a Keplerian disc), capped at 0.5 (its value at r = 6); wide annuli (the stand-alone kyconv seam) use 0.5.
__host__ __device__ inline int rg_auto_nsub(double r1, double r2, double lnr, double dlnE)
{
    double kappa = 0.5;
    if (lnr <= 0.2) {
        double rg = sqrt(r1 * r2);
        if (rg > 6.0) {
            kappa = 1.5 / (rg - 3.0) + 0.5 / sqrt(rg);
            if (kappa > 0.5) kappa = 0.5;
        }
    }
    // + 1e-9: grids commensurate with the radial binning put kappa lnr / dlnE on an integer; the bias keeps the floor
    // independent of last-bit differences between CUDA's and the host's log()
    int nsub = 1 + (int)floor(kappa * lnr / dlnE + 1e-9);
    return nsub > 4096 ? 4096 : nsub;
}

// geometry of ring m of an annulus [r1, r2] cut into nsub rings (kerr_oracle.c: ro_ky_kernel)
struct RingGeo { double A, fr; int s0; };

__device__ inline RingGeo rg_ring_geo(const TabSel &ts, double r1, double r2, double lnr, int m, int nsub,
                                      double alpha, double beta, double rb)
{
    // ring edges r1 (r2/r1)^(m/nsub); the outermost edges are r1 and r2 themselves (most annuli are a single ring)
    const double ra = m == 0 ? r1 : r1 * exp(lnr * m / nsub);
    const double rbb = m == nsub - 1 ? r2 : r1 * exp(lnr * (m + 1) / nsub);
    const double rm = sqrt(ra * rbb);
    double emis = 1.0;
    if (alpha != 0 || beta != 0) emis = rm < rb ? pow(rm, -alpha) : pow(rm, -beta) * pow(rb, beta - alpha);
    RingGeo g;
    g.A = 0.5 * (rbb * rbb - ra * ra) * emis;
    const double ell = rg_row_pos(ts.x0, ts.dx, ts.NR, rm);
    int s0 = (int)floor(ell);
    if (s0 > ts.NR - 2) s0 = ts.NR - 2;
    if (s0 < ts.s_lo) s0 = ts.s_lo; // (annuli start at risco: only a caller's r_in below the slice gets here)
    double fr = ell - s0;
    if (fr < 0) fr = 0;
    g.s0 = s0; g.fr = fr;
    return g;
}

// ---------------------------------------------------------------------------
// Generic deposit (any ordering of the segments): one segment per lane into Hs[k - K_LO] (Hs is owned by this
// warp).  The weight W of segment [sa, sb] (in units of shift bins) is spread uniformly and binned with the linear
// (hat) kernel.  Lanes are grouped into runs of equal first bin (one ballot); for every bin offset the run is summed
// by a segmented inclusive scan (shuffles) and written once by the run's last lane; runs that are not adjacent but
// share the first bin are written in successive rounds in lane order, so no two lanes ever update one bin at once
// and the summation order is fixed.  Used for the rare rings whose shift is not a two-branch function of azimuth;
// the usual ring goes through warp_ring_gather below.
struct SegIv { double up, dn; }; // weight handed to bin k+1 / bin k by the part of the segment inside [k, k+1]

__device__ inline SegIv seg_first(bool point, double kf, double lo, double hi, double W, double inv)
{
    const double b = hi < kf + 1.0 ? hi : kf + 1.0;
    const double wl = point ? W : inv * (b - lo);
    SegIv r;
    r.up = wl * (0.5 * (lo + b) - kf);
    r.dn = wl - r.up;
    return r;
}

__device__ inline SegIv seg_next(double km, double hi, double inv)
{
    const double b = hi < km + 1.0 ? hi : km + 1.0;
    const double l = b - km;
    const double wl = l > 0 ? inv * l : 0.0;
    SegIv r;
    r.up = wl * (0.5 * l);
    r.dn = wl - r.up;
    return r;
}

#define RG_SCAN_STEP(d)                                                              \
    {                                                                                \
        const double mk = (lane - (d) >= start) ? 1.0 : 0.0;                         \
        v0 = fma(__shfl_up_sync(RG_FULL, v0, (d)), mk, v0);                          \
        if (nlive > 1) v1 = fma(__shfl_up_sync(RG_FULL, v1, (d)), mk, v1);           \
        if (nlive > 2) v2 = fma(__shfl_up_sync(RG_FULL, v2, (d)), mk, v2);           \
    }

static __device__ __noinline__ void warp_deposit(double *Hs, int K_LO, int nH, double sa, double sb, double W)
{
    const int lane = threadIdx.x & 31;
    double lo = sa < sb ? sa : sb, hi = sa < sb ? sb : sa;
    const double len = hi - lo;
    const bool point = len < 1e-12;
    if (point) lo = hi = 0.5 * (lo + hi);
    const double kf = floor(lo);
    const int kfirst = (int)kf;
    const int nb = point ? 2 : (int)floor(hi) + 1 - kfirst + 1;
    const double inv = point ? 0.0 : W * __drcp_rn(len);
    const int kprev = __shfl_up_sync(RG_FULL, kfirst, 1);
    const unsigned heads = __ballot_sync(RG_FULL, lane == 0 || kfirst != kprev);
    const int start = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));
    const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
    int myround, nrounds;
    if (__all_sync(RG_FULL, lane == 0 || kfirst >= kprev) || __all_sync(RG_FULL, lane == 0 || kfirst <= kprev)) {
        myround = tail ? 0 : -1;
        nrounds = 1;
    } else {
        const unsigned peers = __match_any_sync(RG_FULL, tail ? kfirst : (-0x40000000 - lane));
        myround = tail ? __popc(peers & ((1u << lane) - 1u)) : -1;
        nrounds = __reduce_max_sync(RG_FULL, myround) + 1;
    }
    const int nbmax = __reduce_max_sync(RG_FULL, nb);
    const int maxrun = __reduce_max_sync(RG_FULL, lane - start + 1);
    double carry = 0.0;
    double km = kf;
    for (int c0 = 0; c0 < nbmax; c0 += 3) {
        const int nlive = nbmax - c0 < 3 ? nbmax - c0 : 3;
        double v0, v1 = 0.0, v2 = 0.0;
        {
            const SegIv iv = c0 == 0 ? seg_first(point, kf, lo, hi, W, inv) : seg_next(km, hi, inv);
            v0 = carry + iv.dn; carry = iv.up;
        }
        if (nlive > 1) { const SegIv iv = seg_next(km + 1.0, hi, inv); v1 = carry + iv.dn; carry = iv.up; }
        if (nlive > 2) { const SegIv iv = seg_next(km + 2.0, hi, inv); v2 = carry + iv.dn; carry = iv.up; }
        km += 3.0;
        if (maxrun > 1) RG_SCAN_STEP(1)
        if (maxrun > 2) RG_SCAN_STEP(2)
        if (maxrun > 4) RG_SCAN_STEP(4)
        if (maxrun > 8) RG_SCAN_STEP(8)
        if (maxrun > 16) RG_SCAN_STEP(16)
        const int ib = kfirst + c0 - K_LO;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (c < nlive) {
                const double v = c == 0 ? v0 : (c == 1 ? v1 : v2);
                const unsigned i = (unsigned)(ib + c);
                const bool ok = i < (unsigned)nH;
                for (int rd = 0; rd < nrounds; rd++) {
                    if (myround == rd && ok) Hs[i] += v;
                    __syncwarp();
                }
            }
        }
    }
}
#undef RG_SCAN_STEP

// ---------------------------------------------------------------------------
// Per-warp scratch of the ring routines.
struct RingScr {
    double2 *SW;  // [NP] (shift in bins, weight density) of every ring sample, raw column order
    double *DA;   // [DN] rising-branch part of the ring's D(x) on consecutive integer shifts (see warp_ring_gather)
    double *DB;   // [DN] falling-branch part
    int DN;
};
#define RG_RING_SCR_DOUBLES(NP, DN) (2 * (NP) + 2 * (DN))

__device__ inline RingScr rg_ring_scr(double *base, int NP, int DN)
{
    RingScr sc;
    sc.SW = reinterpret_cast<double2 *>(base); sc.DA = base + 2 * NP; sc.DB = sc.DA + DN; sc.DN = DN;
    return sc;
}

// index (lowest on ties) of the smallest / largest of the per-lane candidates (v, idx): three REDUX instead of a
// five-step shuffle tree.  Doubles are compared through their order-preserving 64-bit integer image.
__device__ inline int warp_arg_extreme(double v, int idx, bool want_max)
{
    long long b = __double_as_longlong(v);
    unsigned long long key = (unsigned long long)(b < 0 ? ~b : (b | (long long)0x8000000000000000ull));
    if (want_max) key = ~key;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(RG_FULL, hi);
    const unsigned mlo = __reduce_min_sync(RG_FULL, hi == mhi ? lo : 0xffffffffu);
    const bool cand = hi == mhi && lo == mlo;
    return (int)__reduce_min_sync(RG_FULL, cand ? (unsigned)idx : 0x7fffffffu);
}

// Shift histogram of ONE ring whose samples (S[t], Wv[t], t < NP) sit in the warp's scratch: Hs[k - K_LO] += ...
// The usual ring is a closed curve with one minimum and one maximum of the shift: two monotone branches.  Then the
// hat deposit is a GATHER: with D(x) = int^x int^x' (weight density) -- per segment 0 below it, W (x - lo)^2 / (2 len)
// inside, W (x - mid) above -- the histogram is the second difference H[k] = D(k+1) - 2 D(k) + D(k-1), and on a
// monotone branch "the segments entirely below x" are a prefix of the branch: D(x) = x P0 - P1 + Q (x - lo)^2 with the
// prefix sums P0 = sum W, P1 = sum W mid over the segments before the one that holds x.  Every lane owns NP / 32
// consecutive segments (rotated so that the minimum comes first), gets its prefix from ONE warp scan, and evaluates D
// at the integer shifts inside its own segments -- each integer belongs to exactly one segment of the rising and one
// of the falling branch, so the two branches are written one after the other without atomics and in a fixed
// order; then one lane per shift takes the second difference.  Shifts are taken relative to floor(s_min) so that D
// stays O(W x support).  Rings with more than two branches (not met on the default tables; checked here) go through
// the generic segmented-scan deposit.
template <int NQ> // NQ = NPHI / 32 segments per lane
__device__ inline void warp_ring_gather(double *Hs, int K_LO, int nH, const RingScr &sc, double cW, int &kf_min, int &kl_max)
{
    const int lane = threadIdx.x & 31;
    constexpr int NP = 32 * NQ, mask = NP - 1;
    // ---- extremes of the shift around the ring
    double vmin = sc.SW[lane].x, vmax = vmin;
    int imin = lane, imax = lane;
#pragma unroll
    for (int q = 1; q < NQ; q++) {
        const int t = lane + 32 * q;
        const double v = sc.SW[t].x;
        if (v < vmin) { vmin = v; imin = t; }
        if (v > vmax) { vmax = v; imax = t; }
    }
    imin = warp_arg_extreme(vmin, imin, false);
    imax = warp_arg_extreme(vmax, imax, true);
    const double smin = sc.SW[imin].x, smax = sc.SW[imax].x;
    const int m = (imax - imin) & mask; // rotated position of the maximum: segments [0, m) rise, [m, NP) fall
    const double xref = floor(smin);
    const int kfirst = (int)xref, klast = (int)floor(smax) + 1; // bins the hat deposit can touch
    kf_min = kfirst < kf_min ? kfirst : kf_min;
    kl_max = klast > kl_max ? klast : kl_max;
    // ---- this lane's segments i = NQ lane + q in rotated order (sample t = (imin + i) mod NP)
    double r[NQ + 1], ws[NQ];
    bool ok = true;
    double l0 = 0.0, l1 = 0.0;
    const int i0 = NQ * lane;
    {
        double2 sw = sc.SW[(imin + i0) & mask];
        r[0] = sw.x - xref;
#pragma unroll
        for (int q = 0; q < NQ; q++) {
            const double2 nx = sc.SW[(imin + i0 + q + 1) & mask];
            r[q + 1] = nx.x - xref;
            ws[q] = (sw.y + nx.y) * cW;
            sw = nx;
            ok = ok && (i0 + q < m ? r[q + 1] >= r[q] : r[q + 1] <= r[q]);
            l0 += ws[q];
            l1 += ws[q] * (0.5 * (r[q] + r[q + 1]));
        }
    }
    ok = __all_sync(RG_FULL, ok);
    if (!ok) {
        // generic path: segments in raw column order, 32 at a time
        for (int q = 0; q < NQ; q++) {
            const int t = lane + 32 * q, t1 = (t + 1) & mask;
            warp_deposit(Hs, K_LO, nH, sc.SW[t].x, sc.SW[t1].x, (sc.SW[t].y + sc.SW[t1].y) * cW);
        }
        __syncwarp();
        return;
    }
    // exclusive prefix over the lanes of (sum W, sum W mid), and the ring totals
    double p0 = l0, p1 = l1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double u0 = __shfl_up_sync(RG_FULL, p0, d), u1 = __shfl_up_sync(RG_FULL, p1, d);
        if (lane >= d) { p0 += u0; p1 += u1; }
    }
    const double tot0 = __shfl_sync(RG_FULL, p0, 31), tot1 = __shfl_sync(RG_FULL, p1, 31);
    p0 -= l0; p1 -= l1;
    const double rmin = smin - xref, rmax = smax - xref;
    // ---- D on the integer shifts x = xb .. xb + DN - 1 (relative to xref), windows of DN - 2 interior points
    const int x_last = klast - kfirst; // last shift with a non-zero histogram entry
    for (int xb = -1; xb + 1 <= x_last; xb += sc.DN - 2) {
        const int xe = xb + sc.DN - 1 < x_last + 1 ? xb + sc.DN - 1 : x_last + 1; // window [xb, xe]
        // every segment writes D_seg(x) = c0 x - c1 + qv (x - lo)^2 at the integer shifts lo <= x < hi it holds: the
        // rising ones into DA (c = the prefix before the segment), the falling ones into DB (c = everything after it).
        // Each interior integer of the ring's range is written exactly once per branch.
        int xn[NQ], xl[NQ];
        double c0[NQ], c1[NQ], qv[NQ], lo[NQ];
        int more = 0;
        {
            double q0 = p0, q1 = p1;
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                const double a = r[q], b = r[q + 1];
                const bool rising = i0 + q < m;
                const double e0 = q0, e1 = q1;
                q0 += ws[q];
                q1 += ws[q] * (0.5 * (a + b));
                lo[q] = rising ? a : b;
                const double hi = rising ? b : a, len = hi - lo[q];
                c0[q] = rising ? e0 : tot0 - q0;
                c1[q] = rising ? e1 : tot1 - q1;
                qv[q] = len > 1e-12 ? (0.5 * ws[q]) * __drcp_rn(len) : 0.0;
                int x0 = (int)ceil(lo[q]), x1 = (int)ceil(hi) - 1;
                if (x0 < xb) x0 = xb;
                if (x1 > xe) x1 = xe;
                xn[q] = x0; xl[q] = x1;
                double *dst = rising ? sc.DA : sc.DB;
                if (x0 <= x1) {
                    const double x = (double)x0, d = x - lo[q];
                    dst[x0 - xb] = fma(qv[q], d * d, fma(x, c0[q], -c1[q]));
                }
                more |= x1 > x0;
            }
        }
        if (__any_sync(RG_FULL, more)) { // segments holding more than one integer shift (steep parts of wide rings)
#pragma unroll
            for (int q = 0; q < NQ; q++) {
                double *dst = (i0 + q < m) ? sc.DA : sc.DB;
                for (int xi = xn[q] + 1; xi <= xl[q]; xi++) {
                    const double x = (double)xi, d = x - lo[q];
                    dst[xi - xb] = fma(qv[q], d * d, fma(x, c0[q], -c1[q]));
                }
            }
        }
        __syncwarp();
        // one lane per shift: D (exterior shifts in closed form), then the second difference on the window's interior
        for (int xp = xb; xp + 1 < xe; xp += 30) {
            const int xi = xp + lane;
            const double x = (double)xi;
            double D = 0.0;
            if (xi <= xe) {
                if (x >= rmax) D = x * tot0 - tot1;
                else if (x > rmin) D = sc.DA[xi - xb] + sc.DB[xi - xb];
            }
            const double Dm = __shfl_up_sync(RG_FULL, D, 1), Dp = __shfl_down_sync(RG_FULL, D, 1);
            if (lane >= 1 && lane <= 30 && xi < xe) {
                const unsigned idx = (unsigned)(kfirst + xi - K_LO);
                if (idx < (unsigned)nH) Hs[idx] += (Dp - D) - (D - Dm);
            }
        }
        __syncwarp();
    }
}

// One warp adds weight x (shift histogram of rings m_begin, m_begin + m_step, ... of annulus [r1, r2] cut into nsub
// rings) to Hs, sampling the slice (shared or global memory).  alpha, beta, rb: kyconv's broken power-law emissivity;
// zs: redshift in units of shift bins.  single: geometry of the annulus' only ring when the caller has it already.
template <int NQ>
__device__ inline void warp_build_rings(double *Hs, int K_LO, int nH, const TabSel &ts, const double2 *slice,
                                        const RingScr &sc, double r1, double r2, double inv_dlnE, int nsub, int m_begin,
                                        int m_step, double alpha, double beta, double rb, double zs, double weight,
                                        int &kmin_out, int &kmax_out, const RingGeo *single = nullptr)
{
    const int lane = threadIdx.x & 31;
    constexpr int NP = 32 * NQ;
    const double lnr = single ? 0.0 : log(r2 / r1);
    int kf_min = RG_IMAX, kl_max = -RG_IMAX;
    const int n_mine = m_begin < nsub ? (nsub - m_begin + m_step - 1) / m_step : 0;
    for (int m0 = 0; m0 < n_mine; m0 += 32) {
        // the geometry of 32 rings at a time, one per lane
        const int nr = n_mine - m0 < 32 ? n_mine - m0 : 32;
        RingGeo mine;
        mine.A = 0; mine.fr = 0; mine.s0 = ts.s_lo;
        if (single) mine = *single;
        else if (lane < nr) mine = rg_ring_geo(ts, r1, r2, lnr, m_begin + (m0 + lane) * m_step, nsub, alpha, beta, rb);
        for (int mi = 0; mi < nr; mi++) {
            RingGeo geo = mine;
            if (!single) {
                geo.A = __shfl_sync(RG_FULL, mine.A, mi); geo.fr = __shfl_sync(RG_FULL, mine.fr, mi);
                geo.s0 = __shfl_sync(RG_FULL, mine.s0, mi);
            }
            const double2 *row0 = slice + (size_t)(geo.s0 - ts.s_lo) * NP, *row1 = row0 + NP;
            double2 a[NQ], b[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++) { a[q] = row0[lane + 32 * q]; b[q] = row1[lane + 32 * q]; }
#pragma unroll
            for (int q = 0; q < NQ; q++)
                sc.SW[lane + 32 * q] = make_double2((a[q].x + geo.fr * (b[q].x - a[q].x)) * inv_dlnE + zs,
                                                    a[q].y + geo.fr * (b[q].y - a[q].y));
            __syncwarp();
            warp_ring_gather<NQ>(Hs, K_LO, nH, sc, 0.5 * ts.dphi * geo.A * weight, kf_min, kl_max);
        }
    }
    kmin_out = kf_min;
    kmax_out = kl_max;
}
