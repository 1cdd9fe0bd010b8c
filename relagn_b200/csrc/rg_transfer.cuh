// rg_transfer.cuh -- device code of the relativistic transfer (the KYCONV-style
// seam, relagn.py:979-1001): table lookup for one (spin, inclination), ring
// evaluation and the g-shift histogram deposit.  Shared by the fused spectrum
// kernel and the stand-alone kyconv entry point.
//
// Discretisation (identical to oracle/kerr_oracle.c):
//   - an annulus [r1,r2] is cut into n_sub rings at log-midpoints r_m with exact
//     area weights A_m = (r_b^2 - r_a^2)/2;
//   - on ring m column t (disc azimuth) the tables give g and Jn = dalpha dbeta /
//     (dX dY); they are blended linearly between the 4 surrounding (spin,
//     inclination) nodes and the 2 surrounding table rows;
//   - s_t = ln g_t / dlnE is the energy shift in bins, w_t = Jn_t g_t^3 the weight
//     density; the segment t..t+1 carries W = (w_t + w_{t+1})/2 dphi A_m spread
//     uniformly in s over [s_t, s_{t+1}] and is deposited on integer shifts with a
//     linear (hat) kernel.
#pragma once
#include "rg_kernels.h"

struct TabSel {
    const double2 *n00, *n01, *n10, *n11; // node planes [NR][NPHI] of (g, jn) pairs
    double w00, w01, w10, w11;
    double risco, rh;  // of the model's own spin
    size_t plane;
    int NR, NPHI;
    double inv_dl;
    double dphi; // 2 pi / NPHI
};

__device__ inline void rg_tab_select(TabSel &ts, const TabDesc &t, double a, double theta_o)
{
    int m0 = 0;
    while (m0 + 2 < t.n_spin && a >= t.spins[m0 + 1]) m0++;
    int m1 = m0 + 1 < t.n_spin ? m0 + 1 : m0;
    double risco = rg_risco(a);
    double fa = 0;
    if (m1 != m0) {
        fa = (t.riscos[m0] - risco) / (t.riscos[m0] - t.riscos[m1]);
        fa = fa < 0 ? 0 : (fa > 1 ? 1 : fa);
    }
    int n0 = 0;
    while (n0 + 2 < t.n_inc && theta_o >= t.thetas[n0 + 1]) n0++;
    int n1 = n0 + 1 < t.n_inc ? n0 + 1 : n0;
    double fi = 0;
    if (n1 != n0) {
        fi = (theta_o - t.thetas[n0]) / (t.thetas[n1] - t.thetas[n0]);
        fi = fi < 0 ? 0 : (fi > 1 ? 1 : fi);
    }
    size_t plane = (size_t)t.NR * t.NPHI;
    ts.n00 = t.pairs + ((size_t)m0 * t.n_inc + n0) * plane;
    ts.n01 = t.pairs + ((size_t)m0 * t.n_inc + n1) * plane;
    ts.n10 = t.pairs + ((size_t)m1 * t.n_inc + n0) * plane;
    ts.n11 = t.pairs + ((size_t)m1 * t.n_inc + n1) * plane;
    ts.w00 = (1 - fa) * (1 - fi); ts.w01 = (1 - fa) * fi; ts.w10 = fa * (1 - fi); ts.w11 = fa * fi;
    ts.risco = risco; ts.rh = 1 + sqrt(1 - a * a);
    ts.plane = plane; ts.NR = t.NR; ts.NPHI = t.NPHI; ts.inv_dl = 1.0 / t.dl;
    ts.dphi = 2 * RGK_PI / t.NPHI;
}

struct RingRow { int s0, s1; double fr; };

__device__ inline RingRow rg_ring_row(const TabSel &ts, double rm)
{
    RingRow rr;
    double ell = log10((rm - ts.rh) / (ts.risco - ts.rh)) * ts.inv_dl;
    if (!(rm > ts.rh)) ell = 0;
    if (ell < 0) ell = 0;
    if (ell > ts.NR - 1) ell = ts.NR - 1;
    int s0 = (int)floor(ell);
    if (s0 > ts.NR - 2) s0 = ts.NR - 2;
    if (s0 < 0) s0 = 0;
    rr.s0 = s0; rr.s1 = s0 + 1 < ts.NR ? s0 + 1 : s0; rr.fr = ell - s0;
    return rr;
}

// g and Jn at column p of the ring (blend of 4 nodes x 2 rows)
__device__ inline void rg_ring_sample(const TabSel &ts, int s0, int s1, double fr, int p, double &g, double &jn)
{
    // one 16-byte load per node and row brings g and Jn together (the tables are float32 widened once at upload)
    const int i0 = s0 * ts.NPHI + p, i1 = s1 * ts.NPHI + p;
    const double2 a0 = __ldg(ts.n00 + i0), b0 = __ldg(ts.n01 + i0), c0 = __ldg(ts.n10 + i0), d0 = __ldg(ts.n11 + i0);
    const double2 a1 = __ldg(ts.n00 + i1), b1 = __ldg(ts.n01 + i1), c1 = __ldg(ts.n10 + i1), d1 = __ldg(ts.n11 + i1);
    double g0 = ts.w00 * a0.x + ts.w01 * b0.x + ts.w10 * c0.x + ts.w11 * d0.x;
    double g1 = ts.w00 * a1.x + ts.w01 * b1.x + ts.w10 * c1.x + ts.w11 * d1.x;
    double j0 = ts.w00 * a0.y + ts.w01 * b0.y + ts.w10 * c0.y + ts.w11 * d0.y;
    double j1 = ts.w00 * a1.y + ts.w01 * b1.y + ts.w10 * c1.y + ts.w11 * d1.y;
    g = g0 + fr * (g1 - g0);
    jn = j0 + fr * (j1 - j0);
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
struct RingGeo { double A, fr; int s0, s1; };

__device__ inline RingGeo rg_ring_geo(const TabSel &ts, double r1, double r2, double lnr, int m, int nsub,
                                      double alpha, double beta, double rb)
{
    // ring edges r1 (r2/r1)^(m/nsub); the outermost edges are r1 and r2 themselves (most annuli are a single ring)
    const double ra = m == 0 ? r1 : r1 * exp(lnr * m / nsub);
    const double rbb = m == nsub - 1 ? r2 : r1 * exp(lnr * (m + 1) / nsub);
    double rm = sqrt(ra * rbb);
    double emis = 1.0;
    if (alpha != 0 || beta != 0) emis = rm < rb ? pow(rm, -alpha) : pow(rm, -beta) * pow(rb, beta - alpha);
    RingGeo g;
    g.A = 0.5 * (rbb * rbb - ra * ra) * emis;
    RingRow rr = rg_ring_row(ts, rm);
    g.s0 = rr.s0; g.s1 = rr.s1; g.fr = rr.fr;
    return g;
}

// ln g of a table sample: rg_log_pos (bit-identical to log(), fewer issue slots); RG_HIST_OWN_LOG=0 for A/B
#ifndef RG_HIST_OWN_LOG
#define RG_HIST_OWN_LOG 1
#endif
#if RG_HIST_OWN_LOG
#define RG_LOG_G(g) rg_log_pos(g)
#else
#define RG_LOG_G(g) log(g)
#endif
#define RG_FULL 0xffffffffu
#define RG_IMAX 0x7fffffff

// ---------------------------------------------------------------------------
// Deterministic warp-cooperative deposit of one segment per lane into Hs[k - K_LO]
// (Hs is owned by this warp).  The weight W of segment [sa, sb] (in units of
// shift bins) is spread uniformly and binned with the linear (hat) kernel onto
// bins kfirst + c.  The part of the segment inside the unit interval
// [k, k+1] carries weight wl = W (overlap / length) centred at offset mid in
// [0, 1]; the hat hands wl mid to bin k+1 and wl - wl mid to bin k -- so one
// interval evaluation feeds two bins (a degenerate segment is the same formula
// with wl = W on its own interval).  Neighbouring lanes mostly start in the
// same bin, so lanes are grouped into runs of equal kfirst (one ballot); for
// every bin offset c the run is summed by a segmented inclusive scan (shuffles,
// as many steps as the longest run needs) and written once by the run's last
// lane.  Runs that are not adjacent but share kfirst are written in successive
// rounds in lane order, so no two lanes ever update one bin at once and the
// summation order is fixed.
struct SegIv { double up, dn; }; // weight handed to bin k+1 / bin k by the part of the segment inside [k, k+1]

// first interval [kf, kf+1], kf = floor(lo): the part starts at lo itself
__device__ inline SegIv seg_first(bool point, double kf, double lo, double hi, double W, double inv)
{
    const double b = hi < kf + 1.0 ? hi : kf + 1.0;
    const double wl = point ? W : inv * (b - lo);
    SegIv r;
    r.up = wl * (0.5 * (lo + b) - kf);
    r.dn = wl - r.up;
    return r;
}

// a later interval [km, km+1]: the part starts at km (lo lies below it), its centre sits at half its length
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

// one step of the segmented inclusive scan over the runs: lanes at least d behind their run head take the value d
// lanes below (the 0/1 factor keeps the addition exact and the code branch-free)
#define RG_SCAN_STEP(d)                                                              \
    {                                                                                \
        const double mk = (lane - (d) >= start) ? 1.0 : 0.0;                         \
        v0 = fma(__shfl_up_sync(RG_FULL, v0, (d)), mk, v0);                          \
        if (nlive > 1) v1 = fma(__shfl_up_sync(RG_FULL, v1, (d)), mk, v1);           \
        if (nlive > 2) v2 = fma(__shfl_up_sync(RG_FULL, v2, (d)), mk, v2);           \
    }

__device__ inline void warp_deposit(double *Hs, int K_LO, int nH, double sa, double sb, double W, int &kf_min,
                                    int &kl_max)
{
    const int lane = threadIdx.x & 31;
    double lo = sa < sb ? sa : sb, hi = sa < sb ? sb : sa;
    const double len = hi - lo;
    const bool point = len < 1e-12;
    if (point) lo = hi = 0.5 * (lo + hi);
    const double kf = floor(lo);
    const int kfirst = (int)kf;
    const int nb = point ? 2 : (int)floor(hi) + 1 - kfirst + 1;
    const double inv = point ? 0.0 : W * __drcp_rn(len); // len >= 1e-12: the plain reciprocal is safe
    kf_min = kfirst < kf_min ? kfirst : kf_min;
    {
        int kl = kfirst + nb - 1;
        kl_max = kl > kl_max ? kl : kl_max;
    }
    // run structure (the same for every bin offset)
    const int kprev = __shfl_up_sync(RG_FULL, kfirst, 1);
    const unsigned heads = __ballot_sync(RG_FULL, lane == 0 || kfirst != kprev);
    const int start = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));
    const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
    // first bins monotone along the lanes (the usual case: a ring's half is one monotone branch of g(phi)): equal
    // first bins are neighbours, every bin has one run and one round does it -- no match needed
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
    double carry = 0.0; // weight the previous interval hands up to the first bin of this group
    double km = kf;     // lower edge of the group's first interval
    for (int c0 = 0; c0 < nbmax; c0 += 3) {
        const int nlive = nbmax - c0 < 3 ? nbmax - c0 : 3; // bins of this group that any lane reaches (warp-uniform)
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
                const bool ok = i < (unsigned)nH; // 0 <= ib + c < nH
                for (int rd = 0; rd < nrounds; rd++) {
                    if (myround == rd && ok) Hs[i] += v;
                    __syncwarp();
                }
            }
        }
    }
}
#undef RG_SCAN_STEP

// One warp adds weight x (shift histogram of rings m_begin, m_begin + m_step, ... of annulus [r1, r2] cut into nsub
// rings) to Hs.  alpha, beta, rb: kyconv's broken power-law emissivity; zs: redshift in units of shift bins.
__device__ inline void warp_build_rings(double *Hs, int K_LO, int nH, const TabSel &ts, double r1, double r2, double dlnE,
                                        int nsub, int m_begin, int m_step, double alpha, double beta, double rb, double zs,
                                        double weight, int &kmin_out, int &kmax_out, const RingGeo *single = nullptr,
                                        double inv_dlnE_in = 0.0)
{
    // single: geometry of the annulus' only ring when the caller has it already (nsub == 1, warp-uniform)
    const int lane = threadIdx.x & 31;
    const int NP = ts.NPHI, nq = NP >> 5;
    const double lnr = single ? 0.0 : log(r2 / r1);
    const double dphi = ts.dphi;
    // shift in bins = ln g / dlnE, as one multiplication per sample (the caller may pass the reciprocal)
    const double inv_dlnE = inv_dlnE_in != 0.0 ? inv_dlnE_in : 1.0 / dlnE;
    int kf_min = RG_IMAX, kl_max = -RG_IMAX;
    const int n_mine = m_begin < nsub ? (nsub - m_begin + m_step - 1) / m_step : 0;
    for (int m0 = 0; m0 < n_mine; m0 += 32) {
        const int nr = n_mine - m0 < 32 ? n_mine - m0 : 32;
        RingGeo geo;
        geo.A = 0; geo.fr = 0; geo.s0 = 0; geo.s1 = 0;
        if (single) geo = *single;
        else if (lane < nr) geo = rg_ring_geo(ts, r1, r2, lnr, m_begin + (m0 + lane) * m_step, nsub, alpha, beta, rb);
        for (int m = 0; m < nr; m++) {
            const double A = single ? geo.A : __shfl_sync(RG_FULL, geo.A, m), fr = single ? geo.fr : __shfl_sync(RG_FULL, geo.fr, m);
            const int s0 = single ? geo.s0 : __shfl_sync(RG_FULL, geo.s0, m), s1 = single ? geo.s1 : __shfl_sync(RG_FULL, geo.s1, m);
            double g, jn;
            rg_ring_sample(ts, s0, s1, fr, lane, g, jn);
            const double cW = 0.5 * dphi * A * weight; // segment weight = (w_t + w_{t+1}) cW
            double s_cur = RG_LOG_G(g) * inv_dlnE + zs, w_cur = jn * g * g * g;
            const double s_first = __shfl_sync(RG_FULL, s_cur, 0), w_first = __shfl_sync(RG_FULL, w_cur, 0);
            for (int q = 0; q < nq; q++) {
                double s_nxt = 0, w_nxt = 0, sn0, wn0;
                if (q + 1 < nq) {
                    rg_ring_sample(ts, s0, s1, fr, lane + 32 * (q + 1), g, jn);
                    s_nxt = RG_LOG_G(g) * inv_dlnE + zs; w_nxt = jn * g * g * g;
                    sn0 = __shfl_sync(RG_FULL, s_nxt, 0); wn0 = __shfl_sync(RG_FULL, w_nxt, 0);
                } else { sn0 = s_first; wn0 = w_first; }
                double sb = __shfl_down_sync(RG_FULL, s_cur, 1), wb = __shfl_down_sync(RG_FULL, w_cur, 1);
                if (lane == 31) { sb = sn0; wb = wn0; }
                const double W = (w_cur + wb) * cW;
                warp_deposit(Hs, K_LO, nH, s_cur, sb, W, kf_min, kl_max);
                s_cur = s_nxt; w_cur = w_nxt;
            }
        }
    }
    kmin_out = __reduce_min_sync(RG_FULL, kf_min);
    kmax_out = __reduce_max_sync(RG_FULL, kl_max);
}

// One warp adds weight x (shift histogram of the whole annulus [r1, r2], ring count by rg_auto_nsub) to Hs.
__device__ inline void warp_build_hist(double *Hs, int K_LO, int nH, const TabSel &ts, double r1, double r2, double dlnE,
                                       double weight, int &kmin_out, int &kmax_out)
{
    const int nsub = rg_auto_nsub(r1, r2, log(r2 / r1), dlnE);
    warp_build_rings(Hs, K_LO, nH, ts, r1, r2, dlnE, nsub, 0, 1, 0.0, 0.0, 1.0, 0.0, weight, kmin_out, kmax_out);
}

