// rg_spectrum.cu -- the fused spectrum kernel: one CTA per parameter set walks
// the set's annuli (disc, warm, hot), builds every local spectrum in registers,
// pushes it through the relativistic transfer and accumulates the observed
// spectrum, without touching HBM between the parameter record and the final
// [nE] store.
//
// Reference path restated (src/python_version/relagn.py):
//   disc_annuli :839-896, warmComp_annuli :900-939, hotComp_annuli :1295-1318,
//   do_relDiscSpec :950-1027, do_relWarmCompSpec :1063-1124,
//   do_relHotCompSpec :1322-1384, do_nonrel* :1030-1059,1127-1156,1387-1401,
//   pyNTHCOMP.donthcomp :57-76 (interpolation of the Kompaneets solution on the
//   caller's grid); the kyconv seam (:979-1001) as discretised in
//   rg_transfer.cuh.
//
// Work layout.  RG_SPEC_THREADS threads (8 warps); thread t owns the M
// consecutive energy bins j = M t + c; per-bin grid constants and accumulators
// live in registers for the whole set.  Annuli are handled in groups:
//
//  phase H  one WARP per annulus builds that annulus' g-shift histogram H[k]
//           from the Kerr table slice (4 nodes x 2 rows blended): lanes sample
//           the ring azimuths, neighbouring samples form segments, and every
//           segment's weight is deposited with a warp-shuffle segmented
//           reduction keyed by the shift bin (runs of equal keys are summed by
//           a segmented scan, the run tail owns the bin) -- no atomics, global
//           or shared, and a summation order that does not depend on timing.
//           Histograms of a group sit in a shared-memory arena.
//  stream   per annulus, ONE block barrier: every thread evaluates its M bins
//           of the local spectrum, publishes them (transposed, zero halos) and
//           its share of the nu-integral; after the barrier the normalisation
//           is known and out_j += scale * sum_k H[k] in[j-k] runs as a sliding
//           register window: per tap one conflict-free shared load feeds M
//           FMAs, H[k] is a broadcast load.  Double-buffered so the next
//           annulus can be written while stragglers still read.
//  hot      all hot annuli share one spectral shape, so their histograms are
//           summed (weighted by annulus luminosity) and ONE convolution is done.
#include "rg_kernels.h"
#include "rg_transfer.cuh"

#define T_ RG_SPEC_THREADS
#define NWARP (RG_SPEC_THREADS / 32)
#define FULL 0xffffffffu
#define IMAX 0x7fffffff

static __host__ __device__ inline int round_up2(int x) { return (x + 1) & ~1; }

size_t rg_spectrum_smem_bytes(int M, int nH, int HL, int HH, int arena)
{
    size_t d = 0;
    d += 2 * (size_t)M * (HL + T_ + HH);    // inT, double-buffered
    d += (size_t)round_up2(arena);          // histogram arena
    d += (size_t)round_up2(nH + 3 * M + 4); // hot-region summed histogram
    d += 2 * RG_SPT_STRIDE;                 // staged Kompaneets solutions, double-buffered
    d += 2 * NWARP;                         // reduction scratch, double-buffered
    d += 5 * T_;                            // annulus tile scalars
    d += 8;                                 // set scalars
    return d * sizeof(double) + (4 * T_ + 16) * sizeof(int);
}

int rg_spectrum_pick_M(int nE)
{
    if (nE <= 4 * T_) return 4;
    if (nE <= 8 * T_) return 8;
    if (nE <= 20 * T_) return 20;
    return 0;
}

__device__ inline double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// pyNTHCOMP.py:62-73: value of the Kompaneets solution at photon energy e (keV);
// the reference's monotone cursor restated as a direct search seeded from log10(e).
__device__ inline double nth_prim(const double *spt, const double *__restrict__ p10, double xmin, double lg511xmin,
                                  int nth, double e, double e511, double lge)
{
    int j = (int)ceil((lge - lg511xmin) * 50.0);
    if (j < 0) j = 0;
    if (j > nth + 1) j = nth + 1;
    while (j > 0 && !(511.0 * (xmin * p10[j - 1]) < e)) j--;
    while (j <= nth && 511.0 * (xmin * p10[j]) < e) j++;
    if (j > nth) return 0.0;
    if (j == 0) return spt[0];
    int jl = j - 1;
    double xl = xmin * p10[jl], xh = xmin * p10[jl + 1];
    return spt[jl] + ((e511 - xl) * (spt[jl + 1] - spt[jl]) / (xh - xl));
}

struct Smem {
    double *inT, *arena, *Hsum, *spt, *red, *aR1, *aR2, *aS1, *aS2, *aGeo, *setd;
    int *aFlag, *aKmin, *aKmax, *aJz, *seti;
    int RS, inStride;
};

template <int M>
__device__ inline void carve(Smem &s, unsigned char *raw, int nH, int HL, int HH, int arena)
{
    double *d = reinterpret_cast<double *>(raw);
    s.RS = HL + T_ + HH;
    s.inStride = M * s.RS;
    s.inT = d; d += 2 * (size_t)s.inStride;
    s.arena = d; d += round_up2(arena);
    s.Hsum = d; d += round_up2(nH + 3 * M + 4);
    s.spt = d; d += 2 * RG_SPT_STRIDE;
    s.red = d; d += 2 * NWARP;
    s.aR1 = d; d += T_;
    s.aR2 = d; d += T_;
    s.aS1 = d; d += T_;
    s.aS2 = d; d += T_;
    s.aGeo = d; d += T_;
    s.setd = d; d += 8;
    int *i = reinterpret_cast<int *>(d);
    s.aFlag = i; i += T_;
    s.aKmin = i; i += T_;
    s.aKmax = i; i += T_;
    s.aJz = i; i += T_;
    s.seti = i;
}

// ---------------------------------------------------------------------------
// Deterministic warp-cooperative deposit of one segment per lane into Hs[k - K_LO]
// (Hs is owned by this warp).  The weight W of segment [sa, sb] (in units of
// shift bins) is spread uniformly and binned with the linear (hat) kernel onto
// bins kfirst + c.  Neighbouring lanes mostly start in the same bin, so lanes
// are grouped into runs of equal kfirst (one ballot); for every bin offset c the
// run is summed by a segmented inclusive scan (shuffles) and written once by the
// run's last lane.  Runs that are not adjacent but share kfirst are written in
// successive rounds in lane order, so no two lanes ever update one bin at once
// and the summation order is fixed.
__device__ inline double seg_contrib(bool point, int c, int nb, int kfirst, double lo, double hi, double f, double W,
                                     double inv)
{
    if (c >= nb) return 0.0;
    if (point) return c == 0 ? W * (1 - f) : W * f;
    const int k = kfirst + c;
    return inv * (rg_hat_piece(lo - k, hi - k, -1.0, 0.0, true) + rg_hat_piece(lo - k, hi - k, 0.0, 1.0, false));
}

__device__ inline void warp_deposit(double *Hs, int K_LO, int nH, double sa, double sb, double W, int &kf_min,
                                    int &kl_max)
{
    const int lane = threadIdx.x & 31;
    double lo = sa < sb ? sa : sb, hi = sa < sb ? sb : sa;
    double len = hi - lo;
    const bool point = len < 1e-12;
    int kfirst, nb;
    double f = 0, inv = 0;
    if (point) {
        double mid = 0.5 * (lo + hi);
        double kf = floor(mid);
        f = mid - kf;
        kfirst = (int)kf; nb = 2;
    } else {
        kfirst = (int)floor(lo);
        nb = (int)floor(hi) + 1 - kfirst + 1;
        inv = W / len;
    }
    kf_min = kfirst < kf_min ? kfirst : kf_min;
    {
        int kl = kfirst + nb - 1;
        kl_max = kl > kl_max ? kl : kl_max;
    }
    // run structure (the same for every bin offset)
    const int kprev = __shfl_up_sync(FULL, kfirst, 1);
    const unsigned heads = __ballot_sync(FULL, lane == 0 || kfirst != kprev);
    const int start = 31 - __clz((int)(heads & (0xffffffffu >> (31 - lane))));
    const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
    const unsigned peers = __match_any_sync(FULL, tail ? kfirst : (-0x40000000 - lane));
    const int myround = tail ? __popc(peers & ((1u << lane) - 1u)) : -1;
    const int nrounds = __reduce_max_sync(FULL, myround) + 1;
    const int nbmax = __reduce_max_sync(FULL, nb);
    for (int c0 = 0; c0 < nbmax; c0 += 3) {
        double v0 = seg_contrib(point, c0, nb, kfirst, lo, hi, f, W, inv);
        double v1 = seg_contrib(point, c0 + 1, nb, kfirst, lo, hi, f, W, inv);
        double v2 = seg_contrib(point, c0 + 2, nb, kfirst, lo, hi, f, W, inv);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            double u0 = __shfl_up_sync(FULL, v0, d), u1 = __shfl_up_sync(FULL, v1, d), u2 = __shfl_up_sync(FULL, v2, d);
            if (lane - d >= start) { v0 += u0; v1 += u1; v2 += u2; }
        }
        const int i0 = kfirst + c0 - K_LO;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const double v = c == 0 ? v0 : (c == 1 ? v1 : v2);
            for (int rd = 0; rd < nrounds; rd++) {
                if (myround == rd && c0 + c < nbmax) {
                    int i = i0 + c;
                    if (i >= 0 && i < nH) Hs[i] += v;
                }
                __syncwarp();
            }
        }
    }
}

// One warp builds the shift histogram of annulus [r1, r2] into Hs (zeroed by the caller).
__device__ inline void warp_build_hist(double *Hs, int K_LO, int nH, const TabSel &ts, double r1, double r2, double dlnE,
                                       int &kmin_out, int &kmax_out)
{
    const int lane = threadIdx.x & 31;
    const int NP = ts.NPHI, nq = NP >> 5;
    const double lnr = log(r2 / r1);
    const int nsub = rg_auto_nsub(lnr, dlnE);
    const double dphi = 2 * RGK_PI / NP;
    int kf_min = IMAX, kl_max = -IMAX;
    for (int m0 = 0; m0 < nsub; m0 += 32) {
        const int nr = nsub - m0 < 32 ? nsub - m0 : 32;
        RingGeo geo;
        geo.A = 0; geo.fr = 0; geo.s0 = 0; geo.s1 = 0;
        if (lane < nr) geo = rg_ring_geo(ts, r1, r2, lnr, m0 + lane, nsub, 0.0, 0.0, 1.0);
        for (int m = 0; m < nr; m++) {
            const double A = __shfl_sync(FULL, geo.A, m), fr = __shfl_sync(FULL, geo.fr, m);
            const int s0 = __shfl_sync(FULL, geo.s0, m), s1 = __shfl_sync(FULL, geo.s1, m);
            double g, jn;
            rg_ring_sample(ts, s0, s1, fr, lane, g, jn);
            double s_cur = log(g) / dlnE, w_cur = jn * g * g * g;
            const double s_first = __shfl_sync(FULL, s_cur, 0), w_first = __shfl_sync(FULL, w_cur, 0);
            for (int q = 0; q < nq; q++) {
                double s_nxt = 0, w_nxt = 0, sn0, wn0;
                if (q + 1 < nq) {
                    rg_ring_sample(ts, s0, s1, fr, lane + 32 * (q + 1), g, jn);
                    s_nxt = log(g) / dlnE; w_nxt = jn * g * g * g;
                    sn0 = __shfl_sync(FULL, s_nxt, 0); wn0 = __shfl_sync(FULL, w_nxt, 0);
                } else { sn0 = s_first; wn0 = w_first; }
                double sb = __shfl_down_sync(FULL, s_cur, 1), wb = __shfl_down_sync(FULL, w_cur, 1);
                if (lane == 31) { sb = sn0; wb = wn0; }
                const double W = 0.5 * (w_cur + wb) * dphi * A;
                warp_deposit(Hs, K_LO, nH, s_cur, sb, W, kf_min, kl_max);
                s_cur = s_nxt; w_cur = w_nxt;
            }
        }
    }
    kmin_out = __reduce_min_sync(FULL, kf_min);
    kmax_out = __reduce_max_sync(FULL, kl_max);
}

// tmp[c] = sum_k H[k] in[M t + c - k], k in [kmin, kmax]; Hs0 points at shift K_LO of a slot that is
// zero-padded M in front and 2M behind; base = inT + HL + tid of the buffer to read.
template <int M>
__device__ inline void convolve(const double *Hs0, int K_LO, int nH, int kmin, int kmax, const double *base, int RS,
                                double (&tmp)[M])
{
#pragma unroll
    for (int c = 0; c < M; c++) tmp[c] = 0.0;
    if (kmin < K_LO) kmin = K_LO;
    if (kmax > K_LO + nH - 1) kmax = K_LO + nH - 1;
    if (kmin > kmax) return;
    const int g = (kmin >= 0) ? kmin / M : -((-kmin + M - 1) / M); // floor(kmin / M)
    const int k0 = g * M;
    const double *Hk = Hs0 + (k0 - K_LO);
    double w[M];
#pragma unroll
    for (int c = 0; c < M; c++) w[c] = base[c * RS - g];
    const double *pg = base - g - 1;
    for (int k = k0; k <= kmax; k += M) {
#pragma unroll
        for (int u = 0; u < M; u++) {
            const double h = Hk[u];
#pragma unroll
            for (int c = 0; c < M; c++) tmp[c] = fma(h, w[(c - u + M) % M], tmp[c]);
            // next tap needs in[M t - (k+u+1)]: row (M-u-1) % M, column t - g - 1
            w[(M - 1 - u) % M] = pg[((M - u - 1) % M) * RS];
        }
        Hk += M;
        pg -= 1;
    }
}

#ifndef RG_SPEC_MIN_CTAS
#define RG_SPEC_MIN_CTAS 2
#endif
template <int M>
__global__ void __launch_bounds__(RG_SPEC_THREADS, (M <= 4 ? RG_SPEC_MIN_CTAS : 1)) spectrum_kernel(SpecArgs A)
{
    RG_DYN_SMEM(smem_raw);
    Smem s;
    carve<M>(s, smem_raw, A.nH, A.HL, A.HH, A.arena);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int set = blockIdx.x;
    const int nE = A.grid.nE;
    const bool rel = (A.flags & RG_FLAG_REL) != 0;
    const bool comps = (A.flags & RG_FLAG_COMPONENTS) != 0;
    const SetRec &r = A.recs[set];
    const int j0 = M * tid;
    double *out_set = A.out + (size_t)set * (comps ? 4 : 1) * nE;

    if (r.status != RG_S_OK) { // ValueError sets: zeros (the host class raises)
        for (int c = 0; c < (comps ? 4 : 1); c++)
            for (int j = tid; j < nE; j += T_) out_set[(size_t)c * nE + j] = 0.0;
        return;
    }

    // zero both local-spectrum buffers (halos stay zero for the whole kernel)
    for (int i = tid; i < 2 * s.inStride; i += T_) s.inT[i] = 0.0;
    if (tid == 0) { s.seti[2] = 0; s.seti[3] = 0; }

    // per-bin constants of this thread
    double hnu[M], bbpre[M], wt[M], pk[M];
#pragma unroll
    for (int c = 0; c < M; c++) {
        int j = j0 + c;
        bool ok = j < nE;
        hnu[c] = ok ? A.grid.hnu[j] : 1.0; // padding bins: 0 / (exp(1/kT) - 1) = 0
        bbpre[c] = ok ? A.grid.bbpre[j] : 0.0;
        wt[c] = ok ? A.grid.wt[j] : 0.0;
        // pack factor of relagn.py:970-974 (h and 4 pi d^2 cancel against the unpack at :1003-1006)
        pk[c] = ok ? (A.grid.dE[j] * 4) / A.grid.Egrid[j] : 0.0;
    }
    NtConst nt;
    rg_nt_prepare(nt, r);
    TabSel ts;
    const double dlnE = A.grid.dlnE;
    int K_LO = 0, nH = 1, SS = 8, GH = 1;
    if (rel) {
        rg_tab_select(ts, A.tab, r.a, r.inc);
        // shift range this (spin, inclination) can produce: min/max of ln g over the rows of its 4 nodes
        if (warp == 0) {
            double lo = 1e300, hi = -1e300;
            const double *rr = A.tab.rowrange;
            size_t plane2 = (size_t)A.tab.NR * 2;
            size_t nd[4] = {(size_t)((ts.n00 - A.tab.data) / (2 * ts.plane)), (size_t)((ts.n01 - A.tab.data) / (2 * ts.plane)),
                            (size_t)((ts.n10 - A.tab.data) / (2 * ts.plane)), (size_t)((ts.n11 - A.tab.data) / (2 * ts.plane))};
            for (int q = 0; q < 4; q++)
                for (int row = lane; row < A.tab.NR; row += 32) {
                    double a0 = rr[nd[q] * plane2 + 2 * row], a1 = rr[nd[q] * plane2 + 2 * row + 1];
                    lo = a0 < lo ? a0 : lo; hi = a1 > hi ? a1 : hi;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double a0 = __shfl_xor_sync(FULL, lo, o), a1 = __shfl_xor_sync(FULL, hi, o);
                lo = a0 < lo ? a0 : lo; hi = a1 > hi ? a1 : hi;
            }
            if (lane == 0) {
                int kl = (int)floor(lo / dlnE) - 2, kh = (int)floor(hi / dlnE) + 3;
                if (kl < A.K_LO) kl = A.K_LO;
                if (kh > A.K_LO + A.nH - 1) kh = A.K_LO + A.nH - 1;
                s.seti[0] = kl; s.seti[1] = kh - kl + 1;
            }
        }
    }
    __syncthreads();
    if (rel) {
        K_LO = s.seti[0]; nH = s.seti[1];
        SS = round_up2(nH + 3 * M + 4);
        GH = A.arena / SS;
        GH = GH >= 16 ? 16 : (GH >= 8 ? 8 : (GH < 1 ? 1 : GH)); // whole rounds of the 8 builder warps
    }
    const double Rg2 = r.Rg * r.Rg;
    const double cosfac = r.cosinc / 0.5;
    double tot[M];
#pragma unroll
    for (int c = 0; c < M; c++) tot[c] = 0.0;
    int buf = 0;          // parity of the local-spectrum / reduction buffers
    double *const inBase0 = s.inT + A.HL + tid;

    // ------------------------------------------------------------------
    // regions: 0 disc [r_w, r_out], 1 warm [r_h, r_w], 2 hot [risco, r_h]
    for (int region = 0; region < 3; region++) {
        const int n_ann = region == 0 ? r.n_disc : (region == 1 ? r.n_warm : r.n_hot);
        const double lg_in = region == 0 ? r.lg_rw : (region == 1 ? r.lg_rh : r.lg_risco);
        const double lg_out = region == 0 ? r.lg_rout : (region == 1 ? r.lg_rw : r.lg_rh);
        double accC[M], accF[M]; // through-the-transfer sum (before unpack) and flat / non-relativistic sum
#pragma unroll
        for (int c = 0; c < M; c++) { accC[c] = 0.0; accF[c] = 0.0; }
        double fh[M] = {};       // hot: normalised shape Fnu_seed_hot / trapz (relagn.py:1315-1316)
        bool any_conv = false;
        int hk_min = IMAX, hk_max = -IMAX; // hot: support of the summed histogram
        double Sflat = 0.0;                // hot: sum of L_ann 4 pi r dr over flat annuli

        if (n_ann > 0 && region == 1) { // stage the first warm annulus' Kompaneets solution
            const int sys = r.sys_base;
            const int nth = (int)A.header[(size_t)sys * 4 + 2];
            const double *g_spt = A.sptot + (size_t)sys * RG_SPT_STRIDE;
            __syncthreads();
            for (int q = tid; q <= nth; q += T_) s.spt[q] = g_spt[q];
            __syncthreads();
        }
        if (n_ann > 0 && region == 2) {
            // hot shape (relagn.py:1200-1223): donthcomp(Egrid, [gamma_h, kTe_h, kT_seed, 0, 0]) * h / keV
            double F[M];
#pragma unroll
            for (int c = 0; c < M; c++) F[c] = 0.0;
            __syncthreads();
            if (r.has_hotshape) {
                const int sys = r.sys_base + r.n_warm;
                const double *hd = A.header + (size_t)sys * 4;
                const double xmin = hd[0], normfac = hd[1], lgx = hd[3];
                const int nth = (int)hd[2];
                const double *g_spt = A.sptot + (size_t)sys * RG_SPT_STRIDE;
                for (int q = tid; q <= nth; q += T_) s.spt[q] = g_spt[q];
                __syncthreads();
                double e_prev = 0.0, pr_prev = 0.0;
                if (j0 > 0 && j0 - 1 < nE) {
                    e_prev = A.grid.Egrid[j0 - 1];
                    pr_prev = nth_prim(s.spt, A.p10tab, xmin, lgx, nth, e_prev, A.grid.e511[j0 - 1], A.grid.lgE[j0 - 1]) /
                              A.grid.ear2[j0 - 1];
                }
#pragma unroll
                for (int c = 0; c < M; c++) {
                    int j = j0 + c;
                    if (j < nE) {
                        double e = A.grid.Egrid[j];
                        double pr = nth_prim(s.spt, A.p10tab, xmin, lgx, nth, e, A.grid.e511[j], A.grid.lgE[j]) / A.grid.ear2[j];
                        double ph = j > 0 ? (0.5 * (pr + pr_prev) * (e - e_prev) * normfac) : 0.0;
                        F[c] = ph * (1.0 / RGK_KEV) * RGK_H;
                        e_prev = e; pr_prev = pr;
                    }
                }
            }
            double part = 0;
#pragma unroll
            for (int c = 0; c < M; c++) part += F[c] * wt[c];
            part = warp_sum(part);
            if (lane == 0) s.red[buf * NWARP + warp] = part;
            if (rel) {
                for (int q = tid; q < nH + 3 * M + 4; q += T_) s.Hsum[q] = 0.0;
            }
            __syncthreads();
            double shape_int = 0;
#pragma unroll
            for (int w = 0; w < NWARP; w++) shape_int += s.red[buf * NWARP + w];
#pragma unroll
            for (int c = 0; c < M; c++) fh[c] = F[c] / shape_int; // 0/0 -> NaN as in the reference (relagn.py:1316)
            if (rel) {
#pragma unroll
                for (int c = 0; c < M; c++) inBase0[buf * s.inStride + c * s.RS] = fh[c] * pk[c];
            }
            // (published by the barriers of the annulus loop below)
        }

        for (int tile0 = 0; tile0 < n_ann && !(region == 2 && !rel); tile0 += T_) {
            const int n_tile = n_ann - tile0 < T_ ? n_ann - tile0 : T_;
            // ---- prologue: thread i prepares annulus tile0+i (counted top-down) ----
            __syncthreads();
            if (tid < n_tile) {
                BinIter it;
                it.start(lg_in, lg_out, r.dlog_r);
                double lo = 0, hi = 0;
                for (int q = 0; q <= tile0 + tid; q++) it.next(lo, hi);
                double r1 = pow(10.0, lo), r2 = pow(10.0, hi);
                double rmid = pow(10.0, lo + r.dlog_r / 2);
                double T4 = rg_tnt4(nt, rmid);
                double s1, s2;
                if (region == 0) { // relagn.py:857-883
                    double Tann = pow(T4, 0.25);
                    double fcol_r = r.fcol < 0 ? rg_fcol(Tann) : r.fcol;
                    Tann = Tann * fcol_r;
                    double t = Tann / fcol_r;
                    s1 = RGK_KB * Tann;
                    s2 = RGK_SB * (t * t * t * t) * Rg2;
                } else if (region == 1) { // relagn.py:919-932
                    double Tann = pow(T4, 0.25);
                    s1 = 0.0;
                    s2 = RGK_SB * (Tann * Tann * Tann * Tann) * Rg2;
                } else { // relagn.py:1309-1314
                    double Ldiss_ann = RGK_SB * T4 * Rg2;
                    double Lseed_ann = r.Lseed / (4 * RGK_PI * rmid * (r2 - r1) * n_ann);
                    s1 = Ldiss_ann + Lseed_ann;
                    s2 = 0.0;
                }
                // flat space (relagn.py:1013,1110,1370) / non-relativistic (relagn.py:1046,1143) area factor
                double geo = 2 * RGK_PI * 2 * rmid * (r2 - r1);
                if (region != 2) geo = geo * r.cosinc / 0.5;
                // first bin whose Planck exponent overflows (exp(x) = inf for x > 709.78 -> B = 0 exactly): the
                // local spectrum is identically zero from there on and is neither evaluated nor convolved
                int jz = nE;
                if (region == 0) {
                    const double lim = 710.0 * s1;
                    int a = 0, b = nE; // smallest j in [0, nE] with hnu[j] > lim
                    while (a < b) {
                        int m = (a + b) >> 1;
                        if (A.grid.hnu[m] > lim) b = m; else a = m + 1;
                    }
                    jz = a;
                }
                s.aJz[tid] = jz;
                s.aR1[tid] = r1; s.aR2[tid] = r2; s.aS1[tid] = s1; s.aS2[tid] = s2; s.aGeo[tid] = geo;
                s.aFlag[tid] = (rel && lo <= 3 && hi <= 3) ? 1 : 0; // relagn.py:992
            }
            __syncthreads();

            for (int g0 = 0; g0 < n_tile; g0 += GH) {
                const int ng = n_tile - g0 < GH ? n_tile - g0 : GH;
                // ---- phase H: one warp per annulus ----
                if (rel) {
                    for (int i = g0 + warp; i < g0 + ng; i += NWARP) {
                        if (!s.aFlag[i]) continue;
                        double *slot = s.arena + (size_t)(i - g0) * SS;
                        for (int q = lane; q < SS; q += 32) slot[q] = 0.0;
                        __syncwarp();
                        int kmn, kmx;
                        warp_build_hist(slot + M, K_LO, nH, ts, s.aR1[i], s.aR2[i], dlnE, kmn, kmx);
                        if (lane == 0) {
                            s.aKmin[i] = kmn; s.aKmax[i] = kmx;
                            if (kmn <= kmx) { atomicAdd(&s.seti[2], kmx - kmn + 1); atomicAdd(&s.seti[3], 1); }
                        }
                    }
                    __syncthreads();
                }

                if (region == 2) {
                    // hot: fold the group's histograms into the region histogram, annulus by annulus (fixed order)
                    for (int i = g0; i < g0 + ng; i++) {
                        const double L = s.aS1[i];
                        if (s.aFlag[i]) {
                            const int kmn = s.aKmin[i], kmx = s.aKmax[i];
                            const double *slot = s.arena + (size_t)(i - g0) * SS + M;
                            // bin q is always handled by thread q mod T_: race-free across annuli, fixed order
                            for (int q = tid; q < nH; q += T_) {
                                int k = K_LO + q;
                                if (k >= kmn && k <= kmx) s.Hsum[M + q] += L * slot[q];
                            }
                            hk_min = kmn < hk_min ? kmn : hk_min;
                            hk_max = kmx > hk_max ? kmx : hk_max;
                            any_conv = true;
                        } else {
                            Sflat += L * s.aGeo[i];
                        }
                    }
                    __syncthreads();
                    continue;
                }

                // ---- stream: one barrier per annulus ----
                for (int i = g0; i < g0 + ng; i++) {
                    double L[M];
                    double part = 0;
                    double *inw = inBase0 + buf * s.inStride;
                    if (region == 0) {
                        const double kT = s.aS1[i];
                        if (j0 >= s.aJz[i]) { // Wien tail past the overflow point: all of this thread's bins are 0
#pragma unroll
                            for (int c = 0; c < M; c++) L[c] = 0.0;
                        } else {
#pragma unroll
                            for (int c = 0; c < M; c++) {
                                double ef = exp(hnu[c] / kT) - 1;
                                L[c] = RGK_PI * (bbpre[c] / ef);
                                part += L[c] * wt[c];
                            }
                        }
                        if (A.grid.n_ext && tid < A.grid.n_ext) { // relagn.py:868-873
                            double ef = exp(A.grid.x_hnu[tid] / kT) - 1;
                            part += RGK_PI * (A.grid.x_bbpre[tid] / ef) * A.grid.x_wt[tid];
                        }
                    } else {
                        // warm annulus: Kompaneets solution of system sys_base + (tile0 + i), staged in s.spt[parity]
                        const int gi = tile0 + i;
                        const int sys = r.sys_base + gi;
                        const double *hd = A.header + (size_t)sys * 4;
                        const double xmin = hd[0], normfac = hd[1], lgx = hd[3];
                        const int nth = (int)hd[2];
                        const double *spt = s.spt + (gi & 1) * RG_SPT_STRIDE;
                        if (gi + 1 < n_ann) { // prefetch the next annulus' solution into the other buffer
                            const int nth2 = (int)A.header[(size_t)(sys + 1) * 4 + 2];
                            const double *g_spt = A.sptot + (size_t)(sys + 1) * RG_SPT_STRIDE;
                            double *dst = s.spt + ((gi + 1) & 1) * RG_SPT_STRIDE;
                            for (int q = tid; q <= nth2; q += T_) dst[q] = g_spt[q];
                        }
                        double e_prev = 0.0, pr_prev = 0.0;
                        if (j0 > 0 && j0 - 1 < nE) {
                            e_prev = A.grid.Egrid[j0 - 1];
                            pr_prev = nth_prim(spt, A.p10tab, xmin, lgx, nth, e_prev, A.grid.e511[j0 - 1],
                                               A.grid.lgE[j0 - 1]) / A.grid.ear2[j0 - 1];
                        }
#pragma unroll
                        for (int c = 0; c < M; c++) {
                            int j = j0 + c;
                            L[c] = 0.0;
                            if (j < nE) {
                                double e = A.grid.Egrid[j];
                                double pr = nth_prim(spt, A.p10tab, xmin, lgx, nth, e, A.grid.e511[j], A.grid.lgE[j]) /
                                            A.grid.ear2[j];
                                double ph = j > 0 ? (0.5 * (pr + pr_prev) * (e - e_prev) * normfac) : 0.0;
                                L[c] = ph * (1.0 / RGK_KEV) * RGK_H;
                                e_prev = e; pr_prev = pr;
                            }
                            part += L[c] * wt[c];
                        }
                    }
                    const bool conv = s.aFlag[i] != 0;
                    if (conv) {
#pragma unroll
                        for (int c = 0; c < M; c++) inw[c * s.RS] = L[c] * pk[c];
                    }
                    part = warp_sum(part);
                    if (lane == 0) s.red[buf * NWARP + warp] = part;
                    __syncthreads();
                    double radiance = 0;
#pragma unroll
                    for (int w = 0; w < NWARP; w++) radiance += s.red[buf * NWARP + w];
                    // Lnu = norm * (B / radiance), zero when the annulus emits nothing (relagn.py:885-889, 933-937)
                    const double norm = s.aS2[i];
                    if (conv) {
                        double tmp[M];
                        // taps whose sources all lie in the zero zones (j - k >= jz or j - k < 0) are skipped
                        int k_lo = s.aKmin[i], k_hi = s.aKmax[i];
                        {
                            const int a = j0 - s.aJz[i] + 1, b = j0 + M - 1;
                            k_lo = a > k_lo ? a : k_lo;
                            k_hi = b < k_hi ? b : k_hi;
                        }
                        convolve<M>(s.arena + (size_t)(i - g0) * SS + M, K_LO, nH, k_lo, k_hi, inw, s.RS, tmp);
                        if (radiance != 0) {
                            const double scale = norm / radiance;
#pragma unroll
                            for (int c = 0; c < M; c++) accC[c] = fma(scale, tmp[c], accC[c]);
                        }
                        any_conv = true;
                    } else if (radiance != 0) {
                        const double scale = (norm / radiance) * s.aGeo[i];
#pragma unroll
                        for (int c = 0; c < M; c++) accF[c] = fma(scale, L[c], accF[c]);
                    }
                    buf ^= 1;
                }
                if (rel) __syncthreads(); // the arena is rebuilt by the next group
            }
        }

        // ---- region epilogue ----
        if (n_ann > 0) {
            if (region == 2) {
                if (rel) {
                    if (any_conv) {
                        __syncthreads();
                        double tmp[M];
                        convolve<M>(s.Hsum + M, K_LO, nH, hk_min, hk_max, inBase0 + buf * s.inStride, s.RS, tmp);
#pragma unroll
                        for (int c = 0; c < M; c++) accC[c] = tmp[c];
                        buf ^= 1;
                    }
#pragma unroll
                    for (int c = 0; c < M; c++) {
                        int j = j0 + c;
                        double v = 0.0;
                        if (j < nE) {
                            v = Sflat * fh[c];
                            if (any_conv) v += ((accC[c] / A.grid.dE[j]) * A.grid.Egrid[j]) / cosfac; // relagn.py:1362
                        }
                        accF[c] = v;
                    }
                } else {
                    const double Lum = r.Ldiss + r.Lseed; // relagn.py:1396-1398
#pragma unroll
                    for (int c = 0; c < M; c++) accF[c] = Lum * fh[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < M; c++) {
                    int j = j0 + c;
                    double v = accF[c];
                    if (any_conv && j < nE) v += (accC[c] / A.grid.dE[j]) * A.grid.Egrid[j]; // unpack :1003-1006
                    accF[c] = v;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < M; c++) {
            int j = j0 + c;
            if (j < nE) {
                if (comps) out_set[(size_t)region * nE + j] = accF[c];
                tot[c] += accF[c];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < M; c++) {
        int j = j0 + c;
        if (j < nE) out_set[(size_t)(comps ? 3 : 0) * nE + j] = tot[c];
    }
    if (A.stats) {
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&A.stats[0], (unsigned long long)s.seti[2]);
            atomicAdd(&A.stats[1], (unsigned long long)s.seti[3]);
        }
    }
}

template <int M>
static int launch_spec(const SpecArgs &args, cudaStream_t st)
{
    size_t smem = rg_spectrum_smem_bytes(M, args.nH, args.HL, args.HH, args.arena);
    if (smem > 220 * 1024)
        return rg_set_error(RG_E_GRID, "energy grid too fine for the shared-memory working set (%zu B)", smem);
    RG_CUDA_OK(cudaFuncSetAttribute(spectrum_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RG_LAUNCH(spectrum_kernel<M>, dim3(args.P), dim3(T_), smem, st, args);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

int rg_launch_spectrum(const SpecArgs &args, cudaStream_t st)
{
    int M = rg_spectrum_pick_M(args.grid.nE);
    if (M == 4) return launch_spec<4>(args, st);
    if (M == 8) return launch_spec<8>(args, st);
    if (M == 20) return launch_spec<20>(args, st);
    return rg_set_error(RG_E_GRID, "nE = %d exceeds the kernel capacity (%d)", args.grid.nE, 20 * T_);
}
