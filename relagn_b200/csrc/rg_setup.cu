// rg_setup.cu -- per-parameter-set setup kernel and the batched Kompaneets
// (nthcomp) solver.
//
//  setup_kernel   : one thread per parameter set.  Restates the reference
//                   constructors relagn.__init__ (relagn.py:232-357) and
//                   relqso.__init__ (relagn.py:1767-1857): derived scalars,
//                   clamps, region bin counts, Lseed (relagn.py:1227-1264),
//                   Ldiss (relagn.py:1285, fixed Gauss-Legendre instead of
//                   QUADPACK), relqso r_hot search (relagn.py:1877-1898) and
//                   gamma_hot (relagn.py:1901-1907); reserves this set's
//                   nthcomp systems.
//  nthcomp_kernel : one thread per system (a warm annulus or a hot shape).
//                   Restates pyNTHCOMP._thcompton/_thermlc
//                   (pyNTHCOMP.py:80-251): the Thomas recurrences are strictly
//                   serial in j, so parallelism is across systems; work arrays
//                   live in a [j][system] scratch so a warp's accesses coalesce.
#include "rg_common.cuh"
#include "rg_kernels.h"

// ---------------------------------------------------------------------------
__device__ static double seed_temp(const SetRec &r, const NtConst &nt)
{   // relagn.py:1164-1197
    double T4 = rg_tnt4(nt, r.r_h);
    double Tedge = pow(T4, 0.25);
    double kT_edge = (RGK_KB * Tedge) / RGK_KEV;
    double kT_seed;
    if (r.r_w != r.r_h) {
        double ysb = pow(r.gamma_w * (4.0 / 9), -4.5);
        kT_seed = exp(ysb) * kT_edge;
    } else {
        double fcol_r = r.fcol < 0 ? rg_fcol(Tedge) : r.fcol;
        kT_seed = kT_edge * fcol_r;
    }
    return kT_seed;
}

// ---------------------------------------------------------------------------
// Corona energetics (relagn.py:1227-1292), one WARP per parameter set: the seed-photon sum over the annuli r_h .. r_out
// and the dissipation integral are a few hundred Novikov-Thorne evaluations each -- run by one thread they were 0.45 ms of
// serial latency in front of every single-object evaluation (setup_kernel was one thread per set).  Lanes take annuli /
// quadrature points in turn; partial sums are folded in a fixed order (lane-strided sums, then a butterfly), so the
// result does not depend on the batch.
#define LUM_BLOCK 256 // edges of the seed-photon grid a warp buffers at a time

__device__ static double lum_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(128) lumin_kernel(SetRec *__restrict__ recs, int P, int model, double *__restrict__ attrs)
{
    __shared__ double s_edges[4][LUM_BLOCK + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int set = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (set >= P) return;
    SetRec &rr = recs[set];
    if (rr.status != RG_S_OK && rr.status != RG_S_CAP) return;
    const SetRec r = rr;
    NtConst nt;
    rg_nt_prepare(nt, r);
    // ---- Ldiss: int_{risco}^{r_h} 2 sigma T^4 2 pi r Rg^2 dr, composite 16-point Gauss-Legendre in ln r (relagn.py:1285-1286)
    double Ldiss = 0.0;
    if (r.r_h > r.risco) {
        const double ta = log(r.risco), tb = log(r.r_h);
        int npan = 2 + (int)ceil((tb - ta) * 2.0);
        if (npan > 16) npan = 16;
        const double GX[8] = {0.0950125098376374401853193, 0.2816035507792589132304605, 0.4580167776572273863424194,
                              0.6178762444026437484466718, 0.7554044083550030338951012, 0.8656312023878317438804679,
                              0.9445750230732325760779884, 0.9894009349916499325961542};
        const double GW[8] = {0.1894506104550684962853967, 0.1826034150449235888667637, 0.1691565193950025381893121,
                              0.1495959888165767320815017, 0.1246289712555338720524763, 0.0951585116824927848099251,
                              0.0622535239386478928628438, 0.0271524594117540948517806};
        double s = 0.0;
        for (int q = lane; q < npan * 8; q += 32) { // (panel, node pair)
            const int p = q >> 3, i = q & 7;
            const double t0 = ta + (tb - ta) * p / npan, t1 = ta + (tb - ta) * (p + 1) / npan;
            const double tm = 0.5 * (t0 + t1), th = 0.5 * (t1 - t0);
            const double rp = exp(tm + th * GX[i]), rm = exp(tm - th * GX[i]);
            s += (GW[i] * (rg_tnt4(nt, rp) * rp * rp + rg_tnt4(nt, rm) * rm * rm)) * th;
        }
        Ldiss = 2 * RGK_SB * 2 * RGK_PI * (r.Rg * r.Rg) * lum_warp_sum(s);
    }
    // ---- Lseed: sum over the annuli of _make_rbins(log r_h, log r_out) (relagn.py:1227-1264), top-down; lane 0 walks the
    //      chain of repeated subtractions LUM_BLOCK edges at a time, the lanes evaluate the annuli
    double Lseed = 0.0;
    {
        BinIter it;
        it.start(log10(r.r_h), log10(r.r_out), r.dlog_r);
        const double hc = r.r_h < r.hmax ? r.r_h : r.hmax;
        const double Rg2 = r.Rg * r.Rg;
        double *E = s_edges[wib];
        double s = 0.0;
        for (;;) {
            int n = 0;
            if (lane == 0) { // edges E[2k] = lo, E[2k+1] = hi of up to LUM_BLOCK / 2 annuli
                double lo, hi;
                while (n < LUM_BLOCK / 2 && it.next(lo, hi)) { E[2 * n] = lo; E[2 * n + 1] = hi; n++; }
            }
            n = __shfl_sync(0xffffffffu, n, 0);
            __syncwarp();
            for (int k = lane; k < n; k += 32) {
                const double lo = E[2 * k], hi = E[2 * k + 1];
                const double dr = pow(10.0, hi) - pow(10.0, lo);
                const double rmid = pow(10.0, lo + r.dlog_r / 2);
                double cov = 0;
                if (hc <= rmid) {
                    const double th0 = asin(hc / rmid);
                    cov = th0 - 0.5 * sin(2 * th0);
                }
                const double Fr = RGK_SB * rg_tnt4(nt, rmid);
                s += 2 * 2 * RGK_PI * rmid * dr * Fr * cov / RGK_PI * Rg2;
            }
            __syncwarp();
            if (n < LUM_BLOCK / 2) break;
        }
        Lseed = lum_warp_sum(s);
    }
    if (lane == 0) {
        rr.Ldiss = Ldiss; rr.Lseed = Lseed;
        double gamma_h = r.gamma_h;
        if (model == RG_MODEL_RELQSO) { gamma_h = (7.0 / 3) * pow(Ldiss / Lseed, -0.1); rr.gamma_h = gamma_h; }
        if (attrs) {
            double *at = attrs + (size_t)set * RG_NATTR;
            at[RG_A_LDISS] = Ldiss; at[RG_A_LSEED] = Lseed; at[RG_A_GAMMA_H] = gamma_h;
        }
    }
}

// relqso._set_rhot (relagn.py:1877-1898): r_hot is the mid-radius of the fine-grid bin (1000 per decade, built
// top-down from log r_sg by repeated subtraction exactly as _make_rbins does) in which the dissipated luminosity,
// summed inside-out, first exceeds 0.02 L_Edd.  One CTA per parameter set: thread 0 runs the subtraction chain
// (serial by construction, ~3000 additions) into shared memory; the expensive part -- the Novikov-Thorne
// temperature of every bin -- is evaluated 128 bins at a time by the whole CTA; thread 0 then adds the 128 terms in
// the reference's order and tests the threshold before every addition, so the chosen bin and the value of r_hot are
// those of the sequential loop, bit for bit.
#define RQ_CAP 4608
__global__ void __launch_bounds__(128)
relqso_rhot_kernel(const double *__restrict__ params, int P, double *__restrict__ rh_out)
{
    __shared__ double edges[RQ_CAP]; // descending: edges[k] = the k-th edge below log r_sg
    __shared__ double terms[128], rmids[128];
    __shared__ int s_n, s_base, s_nedges, s_done;
    __shared__ double s_Ldiss, s_rh;
    const int idx = blockIdx.x, tid = threadIdx.x;
    if (idx >= P) return;
    const double *p = params + (size_t)idx * RG_NPAR;
    const double M = p[0], mdot = pow(10.0, p[2]), a = p[3], cosinc = p[4];
    bool ok = (a >= -0.998 && a <= 0.998) && (cosinc <= 1 && cosinc >= 0.09);
    if (ok) {
        double lmdot = log10(mdot);
        ok = lmdot >= -1.65 && lmdot <= 0.5;
    }
    if (!ok) { // the setup kernel reports the reference's ValueError
        if (tid == 0) rh_out[idx] = 0.0;
        return;
    }
    const double risco = rg_risco(a);
    const double r_sg = 2150 * pow(M / 1e9, -2.0 / 9) * pow(mdot, 4.0 / 9) * pow(0.1, 2.0 / 9);
    const double L_edd = 1.39e31 * M;
    const double eta = 1 - sqrt(1 - 2 / (3 * risco));
    const double Mdot_edd = L_edd / (eta * (RGK_C * RGK_C));
    const double Rg = (RGK_GMSUN * M) / (RGK_C * RGK_C);
    NtConst nt;
    rg_nt_prepare(nt, a, risco, M, mdot, Mdot_edd, Rg);
    const double dlr = 1.0 / 1000;
    const double lg_in = log10(risco), lg_top = log10(r_sg);
    if (tid == 0) {
        int n = 1;
        double t = lg_top;
        edges[0] = t;
        while (t > lg_in && n < RQ_CAP) {
            t = t - dlr;
            edges[n] = t;
            n++;
        }
        int base, n_edges;
        if (t > lg_in) { n = -1; base = 0; n_edges = 0; } // capacity exceeded
        else if (t == lg_in) { base = n - 1; n_edges = n; }
        else if (n > 1) { base = n - 2; n_edges = n - 1; edges[base] = lg_in; }
        else { base = 0; n_edges = 1; edges[0] = lg_in; }
        s_n = n; s_base = base; s_nedges = n_edges; s_done = 0; s_Ldiss = 0.0; s_rh = risco;
    }
    __syncthreads();
    if (s_n < 0) {
        if (tid == 0) rh_out[idx] = nan("");
        return;
    }
    const int base = s_base, nbins = s_nedges - 1;
    const double Rg2 = Rg * Rg, thresh = 0.02 * L_edd;
    for (int i0 = 0; i0 < nbins; i0 += 128) {
        const int i = i0 + tid;
        if (i < nbins) { // ascending edge i lives at chain index base - i
            double e0 = edges[base - i], e1 = edges[base - i - 1];
            double rmid = pow(10.0, e0 + dlr / 2);
            double dr = pow(10.0, e1) - pow(10.0, e0);
            terms[tid] = RGK_SB * rg_tnt4(nt, rmid) * 4 * RGK_PI * rmid * dr * Rg2;
            rmids[tid] = rmid;
        }
        __syncthreads();
        if (tid == 0) {
            double Ldiss = s_Ldiss, rh = s_rh;
            int nq = nbins - i0 < 128 ? nbins - i0 : 128, done = 0;
            for (int q = 0; q < nq; q++) {
                if (!(Ldiss < thresh)) { done = 1; break; }
                Ldiss += terms[q];
                rh = rmids[q];
            }
            s_Ldiss = Ldiss; s_rh = rh; s_done = done;
        }
        __syncthreads();
        if (s_done) break;
    }
    if (tid == 0) rh_out[idx] = s_rh;
}

__global__ void __launch_bounds__(128)
setup_kernel(const double *__restrict__ params, int P, int model, double dr_dex, SetRec *__restrict__ recs,
             int *__restrict__ sys_counter, int2 *__restrict__ syslist, int sys_cap, double *__restrict__ rq_scratch,
             int rq_cap, double *__restrict__ attrs, int ann_cap)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const double *p = params + (size_t)idx * RG_NPAR;
    SetRec r;
    memset(&r, 0, sizeof r);
    r.dlog_r = 1 / dr_dex;
    double log_rout = -1, r_hot_in = 0, r_warm_in = 0;
    r.M = p[0];
    r.mdot = pow(10.0, p[2]);
    r.a = p[3];
    r.cosinc = p[4];
    r.inc = acos(p[4]);
    if (model == RG_MODEL_RELAGN) {
        r.kTe_h = p[5]; r.kTe_w = p[6]; r.gamma_h = p[7]; r.gamma_w = p[8];
        r_hot_in = p[9]; r_warm_in = p[10]; log_rout = p[11];
        r.r_h = r_hot_in; r.r_w = r_warm_in; r.r_out = pow(10.0, log_rout);
        r.fcol = p[12]; r.hmax = p[13];
    } else {
        r.fcol = p[5];
    }
    int status = RG_S_OK;
    if (!(r.a >= -0.998 && r.a <= 0.998)) status = RG_S_SPIN;
    else if (!(r.cosinc <= 1 && r.cosinc >= 0.09)) status = RG_S_INC;
    else if (model == RG_MODEL_RELQSO) {
        double lmdot = log10(r.mdot);
        if (!(lmdot >= -1.65 && lmdot <= 0.5)) status = RG_S_MDOT;
    }
    NtConst nt;
    if (status == RG_S_OK) {
        r.risco = rg_risco(r.a);
        r.r_sg = 2150 * pow(r.M / 1e9, -2.0 / 9) * pow(r.mdot, 4.0 / 9) * pow(0.1, 2.0 / 9);
        r.L_edd = 1.39e31 * r.M;
        r.eta = 1 - sqrt(1 - 2 / (3 * r.risco));
        r.Mdot_edd = r.L_edd / (r.eta * (RGK_C * RGK_C));
        r.Rg = (RGK_GMSUN * r.M) / (RGK_C * RGK_C);
        rg_nt_prepare(nt, r.a, r.risco, r.M, r.mdot, r.Mdot_edd, r.Rg);
        if (model == RG_MODEL_RELAGN) {
            if (log_rout < 0) r.r_out = r.r_sg;
            if (r_warm_in == -1) r.r_w = r.risco;
            if (r_hot_in == -1) r.r_h = r.risco;
            if (!(r.r_w >= r.r_h)) { r.flags |= RG_F_RW_LT_RH; r.r_w = r.r_h; }
            if (!(r.r_w <= r.r_out)) { r.flags |= RG_F_RW_GT_ROUT; r.r_w = r.r_out; }
            if (!(r.r_h >= r.risco)) { r.flags |= RG_F_RH_LT_RISCO; r.r_h = r.risco; }
            if (!(r.r_w >= r.risco)) { r.flags |= RG_F_RW_LT_RISCO; r.r_w = r.risco; }
            if (!(r.hmax <= r.r_h)) { r.flags |= RG_F_HMAX_GT_RH; r.hmax = r.r_h; }
        } else {
            // relqso._set_rhot (relagn.py:1877-1898): done by relqso_rhot_kernel (one CTA per set)
            double rh = rq_scratch[idx];
            if (!(rh == rh)) status = RG_S_CAP; // NaN: fine radial grid longer than the kernel's capacity
            else {
                r.r_h = rh;
                r.r_w = 2 * r.r_h;
                r.r_out = r.r_sg;
                if (r.r_w > r.r_sg) r.r_w = r.r_sg;
                r.kTe_h = 100; r.kTe_w = 0.2; r.gamma_w = 2.5;
                r.hmax = 100.0 < r.r_h ? 100.0 : r.r_h;
            }
        }
    }
    if (status == RG_S_OK) {
        r.lg_risco = log10(r.risco); r.lg_rh = log10(r.r_h); r.lg_rw = log10(r.r_w); r.lg_rout = log10(r.r_out);
        // (Ldiss, Lseed and relqso's gamma_hot: lumin_kernel, one warp per set, right after this kernel)
        r.Ldiss = 0.0; r.Lseed = 0.0;
        r.n_disc = (r.r_w != r.r_out) ? rg_count_bins(r.lg_rw, r.lg_rout, r.dlog_r) : 0;
        r.n_warm = (r.r_w != r.r_h) ? rg_count_bins(r.lg_rh, r.lg_rw, r.dlog_r) : 0;
        r.n_hot = (r.r_h != r.risco) ? rg_count_bins(r.lg_risco, r.lg_rh, r.dlog_r) : 0;
        if (r.r_h != r.risco) {
            r.kTseed = seed_temp(r, nt);
            if (r.kTseed < r.kTe_h) r.has_hotshape = 1;
            else r.flags |= RG_F_HOT_SEED_GE_KTE;
        }
        r.ann_base = atomicAdd(sys_counter + 1, r.n_disc + r.n_warm + r.n_hot); // rows of the shift-histogram arena
        // (the small-batch path sizes the arenas from a cap instead of reading the counters back: a set that does not
        //  fit is flagged and skipped by every later kernel; the host re-runs the batch through the counting path)
        if (r.ann_base + r.n_disc + r.n_warm + r.n_hot > ann_cap) status = RG_S_CAP;
        int need_hot = (r.n_hot > 0 && r.has_hotshape) ? 1 : 0;
        r.sys_base = 0;
        r.hot_sys = 0;
        // warm systems from the bottom of the arena, hot shapes from the top (the host checks that the two did not meet)
        if (r.n_warm > 0) {
            int base = atomicAdd(sys_counter, r.n_warm);
            r.sys_base = base;
            if (base + r.n_warm <= sys_cap) {
                for (int s = 0; s < r.n_warm; s++) syslist[base + s] = make_int2(idx, s);
            } else status = RG_S_CAP;
        }
        if (need_hot) {
            int k = atomicAdd(sys_counter + 2, 1);
            if (k < sys_cap) {
                r.hot_sys = sys_cap - 1 - k;
                syslist[r.hot_sys] = make_int2(idx, -1);
            } else status = RG_S_CAP;
        }
    }
    r.status = status;
    recs[idx] = r;
    if (attrs) {
        double *at = attrs + (size_t)idx * RG_NATTR;
        at[RG_A_RISCO] = r.risco; at[RG_A_RSG] = r.r_sg; at[RG_A_ETA] = r.eta; at[RG_A_LEDD] = r.L_edd;
        at[RG_A_MDOT_EDD] = r.Mdot_edd; at[RG_A_RG] = r.Rg; at[RG_A_RH] = r.r_h; at[RG_A_RW] = r.r_w;
        at[RG_A_ROUT] = r.r_out; at[RG_A_HMAX] = r.hmax; at[RG_A_GAMMA_H] = r.gamma_h; at[RG_A_KTE_H] = r.kTe_h;
        at[RG_A_KTE_W] = r.kTe_w; at[RG_A_GAMMA_W] = r.gamma_w; at[RG_A_KTSEED] = r.kTseed; at[RG_A_LDISS] = r.Ldiss;
        at[RG_A_LSEED] = r.Lseed; at[RG_A_NDISC] = r.n_disc; at[RG_A_NWARM] = r.n_warm; at[RG_A_NHOT] = r.n_hot;
        at[RG_A_FLAGS] = r.flags; at[RG_A_STATUS] = r.status; at[RG_A_MDOT] = r.mdot; at[RG_A_INC] = r.inc;
    }
}

// ---------------------------------------------------------------------------
// nthcomp: one thread per system runs the coefficient phase and the FORWARD sweep of the Thomas algorithm
// (pyNTHCOMP.py:80-242) and leaves, per row, the two numbers the back substitution needs; the back substitution
// itself is a first-order linear recurrence and runs in nthcomp_interp_kernel as a warp scan, fused with the
// interpolation onto the caller's grid.  Nothing is read back here: the kernel has no memory stalls to hide.
//   coef arena: [sys][RG_SPT_STRIDE] double2 = (G'_j, M_j), see below;  header: [sys][4] = xmin, c0, jmax + 1024 jstop,
//   log10(511 xmin);  farr arena: [sys][nE] = the Comptonised spectrum on the caller's energy grid
//
// The reference's back substitution is u[j] = g[j] - gam[j] u[j+1] (pyNTHCOMP.py:243-247) followed by
// sptot[j] = x[j]^4 u[j] bet[j] tautom (pyNTHCOMP.py:249, 170-171).  With s_j = x[j]^4 bet[j] tautom and v_j = s_j u[j]
// (= sptot[j]) this is v_j = G'_j - M_j v_{j+1}, G'_j = s_j g[j], M_j = gam[j] s_j / s_{j+1}: the row scale moves into
// the forward sweep, where bet[j] is at hand, and the back substitution is one fused multiply-add per row.
// s_j / s_{j+1} = 10^-0.08 bet[j] / bet[j+1], and 1 / bet[j+1] is the denominator of pyNTHCOMP.py:141,147 times
// tautom -- no extra division.  (Differences against the reference's operation order: a few 1e-16 per row.)
//
// Rows the caller's grid cannot reach are not stored: the interpolation (pyNTHCOMP.py:57-76) touches the rows from the
// cell of the first grid energy upwards and the two rows around 1 keV (normalisation, pyNTHCOMP.py:49-56); the
// recurrence runs from the top of the grid downwards, so it can stop at `jstop`.
// a thread's stores walk its own system's rows, 16 bytes per row: the two halves of a 32-byte sector arrive one row
// apart, so the line must stay in L2 until then: default cache policy (evict-first stores measured +-0, r2r)
#define RG_COEF_ST(p, v) (*(p) = (v))
#define RG_NTH_Q4 0.83176377110267100 // 10^-0.08 = (x[j] / x[j+1])^4
__device__ __forceinline__ int rg_nth_jstop(double lgx, double lgE0)
{
    // 3 rows of margin below the cell of the lowest energy needed (the cursor's start value is within one row)
    const double lo = lgE0 < 0.0 ? lgE0 : 0.0;
    const double f = floor((lo - lgx) * 50.0) - 3.0;
    return f < 1.0 ? 1 : (f > 1000.0 ? 1000 : (int)f);
}

__device__ static void nthcomp_system(bool valid, const SetRec &r, int which, const double *p10tab, double lgE0,
                                      double2 *__restrict__ coef, double *__restrict__ hd)
{
    double gamma = 2.0, kte = 1.0, ktbb = 0.01;
    if (!valid) {
    } else if (which < 0) { // hot shape: relagn.py:1200-1223
        gamma = r.gamma_h; kte = r.kTe_h; ktbb = r.kTseed;
    } else {         // warm annulus `which` (counted top-down): relagn.py:900-927
        NtConst nt;
        rg_nt_prepare(nt, r);
        BinIter it;
        it.start(r.lg_rh, r.lg_rw, r.dlog_r);
        double lo = 0, hi = 0;
        for (int s = 0; s <= which; s++) it.next(lo, hi);
        double rmid = pow(10.0, lo + r.dlog_r / 2);
        double Tann = pow(rg_tnt4(nt, rmid), 0.25);
        gamma = r.gamma_w; kte = r.kTe_w; ktbb = Tann / 1.16048e7;
    }
    // ---- _thcompton (pyNTHCOMP.py:80-173) ----
    double tempbb = ktbb / 511.0, theta = kte / 511.0;
    double tautom = sqrt(2.250 + 3.0 / (theta * ((gamma + .50) * (gamma + .50) - 2.250))) - 1.50;
    const double delta = 0.02;
    double deltal = delta * log(10.0);
    double xmin = 1e-4 * tempbb, xmax = 40.0 * theta;
    int jmax = (int)(log10(xmax / xmin) / delta) + 1;
    if (jmax > 899) jmax = 899;
    const double lgx = log10(511.0 * xmin);
    if (valid && jmax < 4) { // degenerate grid: the reference would index out of range; flag as empty
        hd[0] = xmin; hd[1] = 0.0; hd[2] = 0.0; hd[3] = lgx;
        valid = false;
    }
    if (!valid) return;
    const int jstop = rg_nth_jstop(lgx, lgE0);
    int jmaxth = (int)(log10(50 * tempbb / xmin) / delta);
    if (jmaxth > 900) jmaxth = 900;
    if (jmaxth > jmax) jmaxth = jmax;
    double pt = RGK_PI * tempbb;
    double planck = 15.0 / (pt * pt * pt * pt);
    int jnr = (int)(log10(0.10 / xmin) / delta + 1);
    if (jnr > jmax - 1) jnr = jmax - 1;
    int jrel = (int)(log10(1 / xmin) / delta + 1);
    if (jrel > jmax) jrel = jmax;
    double xnr = xmin * p10tab[jnr - 1 > 0 ? jnr - 1 : 0];
    double xr = xmin * p10tab[jrel - 1 > 0 ? jrel - 1 : 0];
    double c20 = tautom / deltal;
    double tod = theta / deltal;

    // rolling values: x[j], x[j+1], c2[j-1]
    double xm = xmin * p10tab[0], x0 = xmin * p10tab[1], xp;
    double x32 = sqrt(xm * x0);
    double aa = (tod / x32 + 0.5) / (tod / x32 - 0.5);
    double w1 = x32;
    double c2m = (w1 * w1 * w1 * w1) / (1.0 + 4.60 * w1 + 1.1 * w1 * w1); // c2[0]
    double gam_prev = 0, g_prev = 0;
    // w2[j] = sqrt(x[j-1] x[j]) is w1[j-1], and theta/deltal/w2[j] the previous step's theta/deltal/w1: carried over
    // (same operands, same operations: bit-identical to recomputing them as pyNTHCOMP.py:211-212 does)
    double tow1_prev = tod / x32;
    double b0; // bet[0]
    {
        double w = xm, rel;
        if (w <= 0.05) rel = (1.0 - 2.0 * w + 26.0 * w * w * 0.2);
        else {
            double z1 = (1.0 + w) / (w * w * w), z2 = 1.0 + 2.0 * w, z3 = log(z2), z4 = 2.0 * w * (1.0 + w) / z2;
            double z5 = z3 / 2.0 / w, z6 = (1.0 + 3.0 * w) / z2 / z2;
            rel = (0.75 * (z1 * (z4 - z3) + z5 - z6));
        }
        if (0 < jnr - 1) b0 = 1.0 / tautom / (1.0 + tautom * rel / 3.0);
        else if (0 < jrel) { double flz = 1 - (w - xnr) / (xr - xnr); b0 = 1.0 / tautom / (1.0 + tautom * rel / 3.0 * flz); }
        else b0 = 1.0 / tautom;
    }
    const double itau = 1.0 / tautom, rxr = 1.0 / (xr - xnr), todx = tod / xmin;
    const double third = 1.0 / 3.0;
    // bet[j] of pyNTHCOMP.py:124-148 at x = x[j] (j >= 1) and its denominator (bet = 1 / tautom / den)
    auto bet_at = [&](int j, double w, double &den) -> double {
        den = 1.0;
        if (j >= jrel) return itau;
        double rel;
        if (w <= 0.05) rel = (1.0 - 2.0 * w + 26.0 * w * w * 0.2);
        else {
            // Klein-Nishina factor (pyNTHCOMP.py:128-136) with the six divisions by w and 1 + 2w folded into
            // two reciprocals (last-bit differences only)
            const double rw = 1.0 / w, z2 = 1.0 + 2.0 * w, rz = 1.0 / z2, z3 = log(z2);
            double z1 = (1.0 + w) * (rw * rw * rw), z4 = 2.0 * w * (1.0 + w) * rz;
            double z5 = z3 * 0.5 * rw, z6 = (1.0 + 3.0 * w) * (rz * rz);
            rel = (0.75 * (z1 * (z4 - z3) + z5 - z6));
        }
        const double taukn = tautom * rel;
        if (j < jnr - 1) den = 1.0 + taukn * third;
        else den = 1.0 + taukn * third * (1 - (w - xnr) * rxr);
        return itau * __drcp_rn(den);
    };
    // seed photons (pyNTHCOMP.py:120-122): x[j] / tempbb = 1e-4 10^(0.02 j) for every system, so
    // d[j] = x[j] dphdot[j] = (planck xmin^3) 10^(0.06 j) / (exp(1e-4 10^(0.02 j)) - 1) comes from a resident table
    const double pk3 = planck * (xmin * xmin * xmin);
    double den_cur, den_nxt;
    double bet_cur = bet_at(1, x0, den_cur), bet_nxt;
    // sptot[0] = s_0 aa u[1] = c0 v_1 (pyNTHCOMP.py:248)
    const double c0 = aa * (RG_NTH_Q4 * b0) * (tautom * den_cur);
    hd[0] = xmin; hd[1] = c0; hd[2] = (double)(jmax + 1024 * jstop); hd[3] = lgx;
    // (unrolling by 2 and one reciprocal for the two independent denominators of a row were measured: +-0, r2w)
    for (int j = 1; j < jmax - 1; j++) {
        xp = xmin * p10tab[j + 1];
        // w1 = sqrt(x[j] x[j+1]) = xmin 10^(0.02 j + 0.01) and theta / deltal / w1 from the resident tables: no
        // square root and one division less per row (last-bit differences against pyNTHCOMP.py:117,211)
        const double w1j = xmin * p10tab[RG_P10_W1 + j];
        const double c2den = 1.0 + 4.60 * w1j + 1.1 * w1j * w1j;
        bet_nxt = bet_at(j + 1, xp, den_nxt);
        const double rc2den = __drcp_rn(c2den);
        // (quotients as reciprocal multiplies: rcp.rn + mul instead of div.rn, last-bit differences)
        double c2j = (w1j * w1j * w1j * w1j) * rc2den;
        const double betj = bet_cur;
        // coefficients (pyNTHCOMP.py:210-221)
        const double tow1 = todx * p10tab[RG_P10_RW1 + j], tow2 = tow1_prev;
        double aj = -c20 * c2j * (tow1 + 0.5);
        double t1 = -c20 * c2j * (0.5 - tow1);
        double t2 = c20 * c2m * (tow2 + 0.5);
        double t3 = (x0 * x0 * x0) * (tautom * betj);
        double bj = t1 + t2 + t3;
        double cj = c20 * c2m * (0.5 - tow2);
        double dj = (j < jmaxth) ? pk3 * p10tab[RG_P10_DPH + j] : 0.0;
        // Thomas forward sweep (pyNTHCOMP.py:233-242); one reciprocal of the pivot serves both quotients
        double alp, gj;
        if (j == 1) alp = bj + cj * aa; else alp = bj - cj * gam_prev;
        const double ralp = 1.0 / alp;
        if (j == 1) gj = dj * ralp; else gj = (dj - cj * g_prev) * ralp;
        double gamj = aj * ralp;
        if (j >= jstop) {
            const double sj = x0 * t3;                                      // x^4 bet tautom
            const double ratio = (RG_NTH_Q4 * betj) * (tautom * den_nxt);   // s_j / s_{j+1}
            RG_COEF_ST(&coef[j], make_double2(sj * gj, gamj * ratio));
        }
        gam_prev = gamj; g_prev = gj;
        x0 = xp; c2m = c2j; tow1_prev = tow1; bet_cur = bet_nxt; den_cur = den_nxt;
    }
}

__global__ void __launch_bounds__(128, RG_NTH_MINB)
nthcomp_kernel(const SetRec *__restrict__ recs, const int2 *__restrict__ syslist, int nsys, int n_warm_sys, int n_warm_pad,
               int hot_first, const double *__restrict__ p10tab, int n_threads, double lgE0,
               double2 *__restrict__ coef, double *__restrict__ header, unsigned long long *stats)
{
    // the resident tables are read at every row: keep them in shared memory
    __shared__ double s_p10[RG_P10_RDP]; // 10^(0.02 j) | w1 | 1 / w1 | seed photons
    for (int i = threadIdx.x; i < RG_NTH + 4; i += blockDim.x) {
        s_p10[i] = p10tab[i];
        s_p10[RG_P10_W1 + i] = p10tab[RG_P10_W1 + i];
        s_p10[RG_P10_RW1 + i] = p10tab[RG_P10_RW1 + i];
        s_p10[RG_P10_DPH + i] = p10tab[RG_P10_DPH + i];
    }
    __syncthreads();
    int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long sumj = 0, nsolved = 0;
    const int lane = threadIdx.x & 31;
    for (int s0 = gtid - lane; s0 < nsys; s0 += n_threads) {
        // logical index -> arena row: the warm systems fill rows [0, n_warm_sys), the hot shapes the top of the arena
        const int li = s0 + lane;
        bool valid = li < nsys && (li < n_warm_sys || li >= n_warm_pad);
        const int sys = li < n_warm_pad ? li : hot_first + (li - n_warm_pad);
        int2 ent = valid ? syslist[sys] : make_int2(0, 0);
        if (ent.x < 0) { valid = false; ent.x = 0; } // a row nobody reserved (the small-batch path scans up to a cap)
        const SetRec &r = recs[ent.x];
        valid = valid && r.status == RG_S_OK;
        const size_t so = valid ? (size_t)sys : 0;
        nthcomp_system(valid, r, ent.y, s_p10, lgE0, coef + so * RG_SPT_STRIDE, header + so * 4);
        if (valid) {
            sumj += (unsigned long long)((int)header[(size_t)sys * 4 + 2] & 1023);
            nsolved++;
        }
    }
    if (stats && nsolved) {
        atomicAdd(&stats[2], sumj);
        atomicAdd(&stats[3], nsolved);
    }
}

// Back substitution + donthcomp's interpolation of the Kompaneets solution onto the caller's grid
// (pyNTHCOMP.py:243-249, 49-76), then * h / keV (relagn.py:929-930, 1217-1218).  A CTA of 8 warps takes 8 systems
// at a time:
//  A. one warp per system: v_j = G'_j - M_j v_{j+1} from the top row down to jstop, 32 rows per step -- every lane
//     holds the affine map of its row, five shuffle steps compose them (Kogge-Stone), the value entering from above
//     is applied last.  The solution never exists in HBM, and not even as rows: a lane that holds v_j and v_{j+1} turns
//     the cell between them into the line prim(E) = a + b (E / 511) of pyNTHCOMP.py:66-71 and stores (a, b) into the
//     warp's table, indexed by the cursor value (0: below the x grid, prim = sptot[0]; nth + 1: above it, 0).  The
//     lane that holds the cell around 1 keV computes the normalisation (pyNTHCOMP.py:49-56).
//  B. one warp per eighth of the energy grid, for all 8 systems: a chunk of passes loads the grid values of its bins
//     once and uses them for 8 systems (with one warp per system the four grid arrays were re-read from L2 by every
//     system: the shared-memory tables leave too little L1 to hold them, r2r).  Per bin: cursor value from log10 E,
//     one 16-byte table read, one fused multiply-add.  The reference's monotone cursor (pyNTHCOMP.py:61-62) stops at
//     the first row with 511 x[j] >= E; ceil((log10 E - log10(511 xmin)) / 0.02) is that row unless E lies within
//     rounding of a cell boundary, and only those bins (closer than 1e-9 of a cell; the rounding is 1e-13) run the
//     reference's own comparisons -- every bin lands in exactly the cell the cursor would have reached.  The first
//     lane of a chunk's first pass computes the bin in front of the chunk, for the trapezoid of pyNTHCOMP.py:75 only.
// The table of a warp holds RG_INTERP_RCAP cursor values from the one of the grid's first bin: enough for a grid of
// (RG_INTERP_RCAP - 6) / 50 = 8.8 decades; wider grids take the GLOBAL variant (whole tables in an L2-resident scratch).
#define RG_INTERP_WARPS 8
#ifndef RG_INTERP_GROUP
#define RG_INTERP_GROUP 4 // steps of 32 rows whose loads are in flight together
#endif
#define RG_INTERP_ROWS (RG_NTH + 4)
#ifndef RG_INTERP_PASSES
#define RG_INTERP_PASSES 4 // passes of 32 bins per register-resident chunk of the grid (1000 bins / 8 warps + 1)
#endif
#define RG_INTERP_RCAP 448
template <bool GLOBAL>
__global__ void __launch_bounds__(RG_INTERP_WARPS * 32)
nthcomp_interp_kernel(const SetRec *__restrict__ recs, const int2 *__restrict__ syslist, int nsys, int n_warm_sys,
                      int n_warm_pad, int hot_first, const double *__restrict__ p10tab, const double2 *__restrict__ coef,
                      const double *__restrict__ header, GridDesc grid, double lgE0, double2 *__restrict__ gtab,
                      double *__restrict__ farr)
{
    RG_DYN_SMEM(smem_raw);
    double *s_p10 = reinterpret_cast<double *>(smem_raw);
    double *s_rdp = s_p10 + RG_INTERP_ROWS; // 1 / (10^(0.02 (j+1)) - 10^(0.02 j)): the cell width of the x grid, per unit xmin
    double *s_sc = s_rdp + RG_INTERP_ROWS;  // [warp][4]: xmin, log10(511 xmin), normfac h / keV, -
    int *s_nth = reinterpret_cast<int *>(s_sc + 4 * RG_INTERP_WARPS); // [warp]: rows, -1 = no system here
    int *s_row = s_nth + RG_INTERP_WARPS;                            // [warp]: arena row of the system
    int *s_jb = s_row + RG_INTERP_WARPS;                             // [warp]: cursor value of table entry 0
    double2 *s_tab = reinterpret_cast<double2 *>(s_jb + RG_INTERP_WARPS); // [warp][RG_INTERP_RCAP] (offset 14 816: 16-byte aligned)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tcap = GLOBAL ? RG_INTERP_ROWS : RG_INTERP_RCAP;
    double2 *tabs = GLOBAL ? gtab + (size_t)blockIdx.x * RG_INTERP_WARPS * RG_INTERP_ROWS : s_tab;
    double2 *tab = tabs + (size_t)warp * tcap;
    for (int i = threadIdx.x; i < RG_NTH + 4; i += blockDim.x) s_p10[i] = p10tab[i];
    for (int i = threadIdx.x; i < RG_NTH + 3; i += blockDim.x) s_rdp[i] = p10tab[RG_P10_RDP + i];
    __syncthreads();
    const unsigned FULL = 0xffffffffu;
    const int nE = grid.nE;
    // photon flux -> W/Hz (relagn.py:929-930) folded into the normalisation
    const double outfac = (1.0 / RGK_KEV) * RGK_H;
    // phase B: this warp's bins [b_lo, b_hi)
    const int bpw = (nE + RG_INTERP_WARPS - 1) / RG_INTERP_WARPS;
    const int b_lo = warp * bpw, b_hi = b_lo + bpw < nE ? b_lo + bpw : nE;
    // The header of the next system is requested while the current one is processed: the chain system list -> set
    // record -> header -> coefficient rows is four dependent trips to memory, and a warp has nothing else to do.
    struct Hd { double xmin, c0, lgx; int packed, sys; };
    auto fetch = [&](int li) -> Hd {
        Hd h; h.xmin = 1.0; h.c0 = 0.0; h.lgx = 0.0; h.packed = -1; h.sys = 0;
        if (li >= nsys || (li >= n_warm_sys && li < n_warm_pad)) return h;
        h.sys = li < n_warm_pad ? li : hot_first + (li - n_warm_pad);
        const int set = __ldg(&syslist[h.sys].x);
        if (set < 0 || recs[set].status != RG_S_OK) return h;
        const double *hd = header + (size_t)h.sys * 4;
        h.xmin = __ldg(hd); h.c0 = __ldg(hd + 1); h.packed = (int)__ldg(hd + 2); h.lgx = __ldg(hd + 3);
        return h;
    };
    const int stride = gridDim.x * RG_INTERP_WARPS;
    Hd nxt = fetch(blockIdx.x * RG_INTERP_WARPS + warp);
    for (int base = blockIdx.x * RG_INTERP_WARPS; base < nsys; base += stride) {
        const Hd cur = nxt;
        nxt = fetch(base + warp + stride);
        // ---- A. back substitution of system base + warp into its table of interpolation lines ----
        int nth = cur.packed < 0 ? -1 : (cur.packed & 1023);
        if (nth >= 0 && nth < 4) nth = 0; // degenerate grid (flagged by the solve): the spectrum is empty
        if (nth > 0) {
            const int jstop = cur.packed >> 10;
            const double xmin = cur.xmin, lgx = cur.lgx, rxmin = 1.0 / xmin;
            const double2 *cf = coef + (size_t)cur.sys * RG_SPT_STRIDE;
            const int jtop = nth - 2; // rows jtop down to jstop
            // cursor value of table entry 0: one below the start value of the grid's first bin
            int jb = 0;
            if (!GLOBAL) {
                const double f = floor((lgE0 - lgx) * 50.0) - 1.0;
                jb = f < 0.0 ? 0 : (f > (double)(nth + 1) ? nth + 1 : (int)f);
            }
            // the cell around 1 keV (pyNTHCOMP.py:49-56): rows il = ih - 1 and ih
            const double xx = 1.0 / 511.0;
            int ih = (int)ceil((0.0 - lgx) * 50.0);
            if (ih < 1) ih = 1;
            if (ih > nth) ih = nth;
            while (ih > 1 && !(xx > xmin * s_p10[ih - 1])) ih--;
            while (ih < nth && xx > xmin * s_p10[ih]) ih++;
            const int il = ih - 1;
            double spp_l = 0.0;
            bool have_spp = false;
            // cursor values nth, nth + 1: the cell between the two zero rows, and everything above the x grid
            if (lane < 2) {
                const int k = nth + lane - jb;
                if (k >= 0 && k < tcap) tab[k] = make_double2(0.0, 0.0);
            }
            double carry = 0.0; // v of the row above the current step (u[jmax-1] = 0)
            // the rows of RG_INTERP_GROUP steps are requested together (their addresses do not depend on the recurrence)
            for (int jg = jtop; jg >= jstop; jg -= 32 * RG_INTERP_GROUP) {
                double2 c[RG_INTERP_GROUP];
#pragma unroll
                for (int q = 0; q < RG_INTERP_GROUP; q++) {
                    const int j = jg - 32 * q - lane;
                    c[q] = j >= jstop ? RG_LDCS2(&cf[j]) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int q = 0; q < RG_INTERP_GROUP; q++) {
                    const int j = jg - 32 * q - lane;
                    if (jg - 32 * q < jstop) break; // warp-uniform
                    double A = -c[q].y, B = c[q].x; // v_j = A v_{j+1} + B
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const double Ap = __shfl_up_sync(FULL, A, d), Bp = __shfl_up_sync(FULL, B, d);
                        if (lane >= d) { B = fma(A, Bp, B); A = A * Ap; }
                    }
                    const double v = fma(A, carry, B);
                    double vup = __shfl_up_sync(FULL, v, 1); // v_{j+1}
                    if (lane == 0) vup = carry;
                    carry = __shfl_sync(FULL, v, 31);
                    if (j >= jstop) {
                        // cell [x_j, x_{j+1}], cursor value j + 1: prim = sl + (E/511 - xl) (sh - sl) / (xh - xl)
                        //   = a + b E/511, the division by the cell width as a multiplication by
                        //   (1 / xmin) (1 / (10^(0.02 (j+1)) - 10^(0.02 j)))
                        const double xl = xmin * s_p10[j];
                        const double bb = (vup - v) * (rxmin * s_rdp[j]);
                        const int k = j + 1 - jb;
                        if (k >= 0 && k < tcap) tab[k] = make_double2(fma(-xl, bb, v), bb);
                        if (j == il) { // (il >= 1 here; il = 0 is the lane of row 1, below)
                            const double xh = xmin * s_p10[ih];
                            spp_l = v + (vup - v) * (xx - xl) / (xh - xl);
                            have_spp = true;
                        }
                        if (j == 1) { // sptot[0] = c0 v_1 (pyNTHCOMP.py:248): cursor value 1 (cell [x_0, x_1]) and 0 (below the grid)
                            const double v0 = cur.c0 * v;
                            const double x0 = xmin * s_p10[0];
                            const double b0 = (v - v0) * (rxmin * s_rdp[0]);
                            if (jb <= 1 && 1 - jb < tcap) tab[1 - jb] = make_double2(fma(-x0, b0, v0), b0);
                            if (jb == 0) tab[0] = make_double2(v0, 0.0);
                            if (il == 0) {
                                const double xh = xmin * s_p10[1];
                                spp_l = v0 + (v - v0) * (xx - x0) / (xh - x0);
                                have_spp = true;
                            }
                        }
                    }
                }
            }
            const unsigned m = __ballot_sync(FULL, have_spp);
            const double spp = m ? __shfl_sync(FULL, spp_l, __ffs(m) - 1) : 0.0; // rows above the top are zero
            if (lane == 0) {
                s_sc[4 * warp] = xmin; s_sc[4 * warp + 1] = lgx; s_sc[4 * warp + 2] = (1.0 / spp) * outfac;
                s_jb[warp] = jb;
            }
        }
        if (lane == 0) { s_nth[warp] = nth; s_row[warp] = cur.sys; }
        __syncthreads();
        // ---- B. interpolation onto the grid (pyNTHCOMP.py:57-76): bins b_lo ... b_hi - 1 of all 8 systems ----
        // in chunks of RG_INTERP_PASSES passes whose grid values stay in registers for the 8 systems; every chunk
        // starts one bin early (lane 0 of its first pass supplies the left end of the first trapezoid only)
        for (int c_lo = b_lo; c_lo < b_hi; c_lo += 32 * RG_INTERP_PASSES - 1) {
            double ge[RG_INTERP_PASSES], glg[RG_INTERP_PASSES], ge5[RG_INTERP_PASSES], gr2[RG_INTERP_PASSES], ghd[RG_INTERP_PASSES];
#pragma unroll
            for (int p = 0; p < RG_INTERP_PASSES; p++) {
                const int i = c_lo - 1 + 32 * p + lane;
                const bool in = i >= 0 && i < b_hi;
                ge[p] = in ? __ldg(grid.Egrid + i) : 0.0;
                glg[p] = in ? __ldg(grid.lgE + i) : 0.0;
                ge5[p] = in ? __ldg(grid.e511 + i) : 0.0;
                gr2[p] = in ? __ldg(grid.rear2 + i) : 0.0;
                ghd[p] = in && i > 0 ? 0.5 * (ge[p] - __ldg(grid.Egrid + i - 1)) : 0.0; // half the trapezoid's width; bin 0 is 0
            }
            unsigned inmask = 0; // bit p: this lane's bin of pass p belongs to the warp's range
#pragma unroll
            for (int p = 0; p < RG_INTERP_PASSES; p++) {
                const int i = c_lo - 1 + 32 * p + lane;
                if (i >= 0 && i < b_hi) inmask |= 1u << p;
            }
            const int npass = (b_hi - (c_lo - 1) + 31) >> 5; // passes of this chunk that reach into the range
            for (int q = 0; q < RG_INTERP_WARPS; q++) {
                const int nq = s_nth[q];
                if (nq < 0) continue; // warp-uniform
                double *Fl = farr + (size_t)s_row[q] * nE + (c_lo - 1 + lane); // this lane's bin of pass 0
                if (nq == 0) { // degenerate Kompaneets grid: empty spectrum
#pragma unroll
                    for (int p = 0; p < RG_INTERP_PASSES; p++)
                        if (((inmask >> p) & 1u) && !(p == 0 && lane == 0)) RG_STCG(&Fl[32 * p], 0.0);
                    continue;
                }
                const double xmin = s_sc[4 * q], lgx = s_sc[4 * q + 1], nf = s_sc[4 * q + 2];
                const int jb = s_jb[q];
                const int khi = nq + 1 - jb < tcap - 1 ? nq + 1 - jb : tcap - 1; // entry of "above the x grid"
                const double2 *tq = tabs + (size_t)q * tcap;
                double carry_pr = 0.0;
#pragma unroll
                for (int p = 0; p < RG_INTERP_PASSES; p++) {
                    if (p >= npass) break; // warp-uniform
                    const bool in = (inmask >> p) & 1u;
                    const double t = (glg[p] - lgx) * 50.0, tc = ceil(t), fr = tc - t;
                    // entry = cursor value clamped to [0, nth + 1], minus the table's first (jb >= 0)
                    int k = (int)tc - jb;
                    if (fr < 1e-9 || fr > 1.0 - 1e-9) { // on a cell boundary to rounding: the reference's comparisons decide
                        const double e = ge[p];
                        int j = (int)tc;
                        j = j < 0 ? 0 : (j > nq + 1 ? nq + 1 : j);
                        const int ja = j > 0 ? j - 1 : 0, jc = j <= nq ? j : nq;
                        const bool below = j > 0 && !(511.0 * (xmin * s_p10[ja]) < e);
                        const bool above = j <= nq && 511.0 * (xmin * s_p10[jc]) < e;
                        k = j + (above ? 1 : (below ? -1 : 0)) - jb;
                    }
                    k = k < 0 ? 0 : (k > khi ? khi : k);
                    const double2 ab = tq[in ? k : 0];
                    const double pr = in ? fma(ab.y, ge5[p], ab.x) * gr2[p] : 0.0; // prim / E^2 (pyNTHCOMP.py:75)
                    double pr_prev = __shfl_up_sync(FULL, pr, 1);
                    if (lane == 0) pr_prev = carry_pr;
                    if (in && !(p == 0 && lane == 0)) RG_STCG(&Fl[32 * p], (pr + pr_prev) * (ghd[p] * nf));
                    carry_pr = __shfl_sync(FULL, pr, 31);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
int rg_launch_setup(const double *d_params, int P, int model, double dr_dex, SetRec *recs, int *sys_counter,
                    int2 *syslist, int sys_cap, double *rq_scratch, int rq_cap, double *d_attrs, int ann_cap,
                    cudaStream_t st)
{
    RG_CUDA_OK(cudaMemsetAsync(sys_counter, 0, 3 * sizeof(int), st)); // [0] warm Kompaneets systems, [1] annuli, [2] hot shapes
    if (model == RG_MODEL_RELQSO) {
        RG_LAUNCH(relqso_rhot_kernel, dim3(P), dim3(128), 0, st, d_params, P, rq_scratch);
        RG_CUDA_OK(cudaGetLastError());
    }
    RG_LAUNCH(setup_kernel, dim3((P + 127) / 128), dim3(128), 0, st, d_params, P, model, dr_dex, recs, sys_counter,
              syslist, sys_cap, rq_scratch, rq_cap, d_attrs, ann_cap);
    RG_CUDA_OK(cudaGetLastError());
    RG_LAUNCH(lumin_kernel, dim3((unsigned)(((size_t)P * 32 + 127) / 128)), dim3(128), 0, st, recs, P, model, d_attrs);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

// Radial bin edges (relagn.py:731-776: top-down by repeated subtraction, innermost bin widened) of every annulus,
// written once per (set, region) to edges[ann_base + g] = (lo, hi) in log10 r; g counts disc, warm, hot top-down.
// hist_kernel and spectrum_kernel used to walk this chain per lane up to its own annulus (O(N^2) per set).
__global__ void __launch_bounds__(128) edges_kernel(const SetRec *__restrict__ recs, int P, double2 *__restrict__ edges)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 3 * P) return;
    const int set = t / 3, region = t - 3 * set;
    const SetRec &r = recs[set];
    if (r.status != RG_S_OK) return;
    const double lg_in = region == 0 ? r.lg_rw : (region == 1 ? r.lg_rh : r.lg_risco);
    const double lg_out = region == 0 ? r.lg_rout : (region == 1 ? r.lg_rw : r.lg_rh);
    const int n = region == 0 ? r.n_disc : (region == 1 ? r.n_warm : r.n_hot);
    double2 *e = edges + r.ann_base + (region == 0 ? 0 : (region == 1 ? r.n_disc : r.n_disc + r.n_warm));
    BinIter it;
    it.start(lg_in, lg_out, r.dlog_r);
    double lo = 0, hi = 0;
    for (int a = 0; a < n; a++) {
        it.next(lo, hi);
        e[a] = make_double2(lo, hi);
    }
}

int rg_launch_edges(const SetRec *recs, int P, double2 *edges, cudaStream_t st)
{
    if (P <= 0) return RG_OK;
    RG_LAUNCH(edges_kernel, dim3((3 * P + 127) / 128), dim3(128), 0, st, recs, P, edges);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

// The two counters of the setup kernel reach the host by a store from the device into pinned (mapped) memory: a
// cudaMemcpy D2H would queue on the copy engine behind whatever large copy is in flight there -- the previous slab's
// spectra in rg_eval_batch_host, a caller's own transfer -- and the chunk would wait for all of it.
__global__ void counters_out_kernel(const int *__restrict__ ctr, int *__restrict__ host_mapped)
{
    host_mapped[0] = ctr[0];
    host_mapped[1] = ctr[1];
    host_mapped[2] = ctr[2];
    __threadfence_system();
}

int rg_launch_counters_out(const int *d_counter, int *h_mapped, cudaStream_t st)
{
    RG_LAUNCH(counters_out_kernel, dim3(1), dim3(1), 0, st, d_counter, h_mapped);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

int rg_launch_nthcomp(const SetRec *recs, const int2 *syslist, int n_warm_sys, int n_hot_sys, int sys_cap,
                      const double *p10tab, int n_threads, double2 *coef, double *header, int sm_count,
                      const GridDesc &grid, double lgE0, double grid_E1, double *gtab, size_t gtab_doubles, double *farr,
                      unsigned long long *stats, cudaStream_t st)
{
    const bool gtab_force = getenv("RG_INTERP_GLOBAL") != nullptr; // test knob: the wide-grid variant on any grid
    // logical order: the warm systems, padded to a whole warp, then the hot shapes -- no warp mixes the two kinds
    const int n_warm_pad = (n_warm_sys + 31) & ~31;
    const int nsys = n_hot_sys > 0 ? n_warm_pad + n_hot_sys : n_warm_sys;
    const int hot_first = sys_cap - n_hot_sys;
    if (nsys <= 0) return RG_OK;
    int blocks = (nsys + 127) / 128;
    if (blocks * 128 > n_threads) blocks = n_threads / 128;
    RG_LAUNCH(nthcomp_kernel, dim3(blocks), dim3(128), 0, st, recs, syslist, nsys, n_warm_sys, n_warm_pad, hot_first, p10tab,
              blocks * 128, lgE0, coef, header, stats);
    RG_CUDA_OK(cudaGetLastError());
    // persistent CTAs, 8 systems at a time: resident tables + RG_INTERP_WARPS tables of interpolation lines
    const double lgE1 = log10(grid_E1);
    const bool global = (lgE1 - lgE0) * 50.0 + 6.0 > (double)RG_INTERP_RCAP || gtab_force;
    const size_t smem = (size_t)2 * RG_INTERP_ROWS * sizeof(double) + (size_t)RG_INTERP_WARPS * (4 * sizeof(double) + 3 * sizeof(int)) +
                        (global ? 0 : (size_t)RG_INTERP_WARPS * RG_INTERP_RCAP * sizeof(double2));
#ifndef RG_HOSTSIM
    // (per launch, like the other kernels: the attribute belongs to the current device, and a process may drive several)
    if (!global)
        RG_CUDA_OK(cudaFuncSetAttribute(nthcomp_interp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
    long ictas = ((long)nsys + RG_INTERP_WARPS - 1) / RG_INTERP_WARPS;
    if (ictas > 3L * sm_count) ictas = 3L * sm_count;
    if (global) {
        const long fit = (long)(gtab_doubles / ((size_t)2 * RG_INTERP_WARPS * RG_INTERP_ROWS));
        if (fit < 1) return rg_set_error(RG_E_NOMEM, "nthcomp_interp: no scratch for the interpolation tables");
        if (ictas > fit) ictas = fit;
        RG_LAUNCH(nthcomp_interp_kernel<true>, dim3((unsigned)ictas), dim3(RG_INTERP_WARPS * 32), smem, st, recs, syslist,
                  nsys, n_warm_sys, n_warm_pad, hot_first, p10tab, coef, header, grid, lgE0, reinterpret_cast<double2 *>(gtab), farr);
    } else {
        RG_LAUNCH(nthcomp_interp_kernel<false>, dim3((unsigned)ictas), dim3(RG_INTERP_WARPS * 32), smem, st, recs, syslist,
                  nsys, n_warm_sys, n_warm_pad, hot_first, p10tab, coef, header, grid, lgE0, (double2 *)nullptr, farr);
    }
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}
