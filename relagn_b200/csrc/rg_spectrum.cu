// rg_spectrum.cu -- the fused spectrum kernel: one CTA per parameter set walks
// the set's annuli (disc, warm, hot), builds every local spectrum in registers,
// pushes it through the relativistic transfer and accumulates the observed
// spectrum, all without touching HBM between the parameter record and the
// final [nE] store.
//
// Reference path restated (src/python_version/relagn.py):
//   disc_annuli :839-896, warmComp_annuli :900-939, hotComp_annuli :1295-1318,
//   do_relDiscSpec :950-1027, do_relWarmCompSpec :1063-1124,
//   do_relHotCompSpec :1322-1384, do_nonrel* :1030-1059,1127-1156,1387-1401,
//   pyNTHCOMP.donthcomp :57-76 (interpolation of the Kompaneets solution on the
//   caller's grid); the kyconv seam (:979-1001) as discretised in
//   rg_transfer.cuh.
//
// Work layout.  RG_SPEC_THREADS threads; thread t owns the M consecutive energy
// bins j = M t + c.  Per-bin grid constants and all accumulators stay in
// registers for the whole set.  Per annulus:
//   1. local spectrum B_j in registers, block reduction of its nu-integral;
//   2. normalised spectrum (as photons-per-bin proxy dE/E * 4 Lnu) written to
//      shared memory TRANSPOSED (row c = j mod M, column j div M) with zero halos,
//      so the sliding-window convolution reads conflict-free and branch-free;
//   3. the annulus' g-shift histogram H[k] is built from the Kerr table slice
//      (4 nodes x 2 rows blended) by one thread per (ring, azimuth segment) with
//      shared-memory atomics;
//   4. out_j += sum_k H[k] in[j-k]: per tap one shared load feeds M FMAs
//      (register sliding window), H[k] is a broadcast load.
// The hot region uses linearity: all hot annuli share one spectral shape, so
// their histograms are summed (weighted by the annulus luminosity) and ONE
// convolution is done for the region.
#include "rg_kernels.h"
#include "rg_transfer.cuh"

#define T_ RG_SPEC_THREADS
#define NWARP (RG_SPEC_THREADS / 32)
#define RING_SLOTS 8 // rings sampled per pass = T_/NPHI <= 8 (NPHI >= 32)

size_t rg_spectrum_smem_bytes(int M, int nH, int HL, int HH)
{
    size_t d = 0;
    d += (size_t)M * (HL + T_ + HH); // inT
    d += (size_t)nH + 3 * M + 4;     // H with front pad M and back pad 2M
    d += 2 * T_;                     // sv, wv
    d += 5 * T_;                     // annulus tile scalars
    d += RG_SPT_STRIDE;              // staged Kompaneets solution
    d += 2 * NWARP;                  // reduction scratch (double-buffered)
    d += 2 * RING_SLOTS;             // ring A, fr
    return d * sizeof(double) + (2 * RING_SLOTS + T_ + 8) * sizeof(int);
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
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum, identical in every thread.  `red` must not be rewritten before
// every thread has read it (callers alternate two buffers).
__device__ inline double block_sum(double v, double *red)
{
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < NWARP; w++) s += red[w];
    return s;
}

// pyNTHCOMP.py:62-73: value of the Kompaneets solution at photon energy e (keV)
// with the reference's monotone cursor restated as a direct search.
__device__ inline double nth_prim(const double *spt, const double *__restrict__ p10, double xmin, int nth, double e,
                                  double e511)
{
    int j = (int)ceil(log10(e / (511.0 * xmin)) * 50.0);
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
    double *inT, *H, *sv, *wv, *aR1, *aR2, *aRm, *aS1, *aS2, *spt, *red, *rA, *rFr;
    int *rS0, *rS1, *aConv, *kmm;
    int RS, HP;
};

template <int M>
__device__ inline void carve(Smem &s, unsigned char *raw, int nH, int HL, int HH)
{
    double *d = reinterpret_cast<double *>(raw);
    s.RS = HL + T_ + HH;
    s.HP = M;
    s.inT = d; d += (size_t)M * s.RS;
    s.H = d; d += nH + 3 * M + 4;
    s.sv = d; d += T_;
    s.wv = d; d += T_;
    s.aR1 = d; d += T_;
    s.aR2 = d; d += T_;
    s.aRm = d; d += T_;
    s.aS1 = d; d += T_;
    s.aS2 = d; d += T_;
    s.spt = d; d += RG_SPT_STRIDE;
    s.red = d; d += 2 * NWARP;
    s.rA = d; d += RING_SLOTS;
    s.rFr = d; d += RING_SLOTS;
    int *i = reinterpret_cast<int *>(d);
    s.rS0 = i; i += RING_SLOTS;
    s.rS1 = i; i += RING_SLOTS;
    s.aConv = i; i += T_;
    s.kmm = i;
}

// Build the shift histogram of annulus [r1,r2] into s.H (adds weight*kernel).
// Entered by all threads after a barrier; ends with a barrier.  `pre_synced`
// says ring geometry of the first pass was already published before the last
// barrier (fused with the caller's reduction).
template <int M>
__device__ inline void build_hist(const Smem &s, const TabSel &ts, double r1, double r2, double dlnE, int K_LO,
                                  int nH, double weight)
{
    const int tid = threadIdx.x;
    const int NP = ts.NPHI;
    const int rpp = T_ / NP; // rings per pass
    double lnr = log(r2 / r1);
    int nsub = rg_auto_nsub(lnr, dlnE);
    double dphi = 2 * RGK_PI / NP;
    int kf_min = 0x7fffffff, kl_max = -0x7fffffff;
    for (int m0 = 0; m0 < nsub; m0 += rpp) {
        int nr = nsub - m0 < rpp ? nsub - m0 : rpp;
        if (tid < nr) {
            RingGeo g = rg_ring_geo(ts, r1, r2, lnr, m0 + tid, nsub, 0.0, 0.0, 1.0);
            s.rA[tid] = g.A; s.rFr[tid] = g.fr; s.rS0[tid] = g.s0; s.rS1[tid] = g.s1;
        }
        __syncthreads();
        int slot = tid / NP, p = tid - slot * NP;
        if (slot < nr) {
            double gg, jj;
            rg_ring_sample(ts, s.rS0[slot], s.rS1[slot], s.rFr[slot], p, gg, jj);
            s.sv[tid] = log(gg) / dlnE;
            s.wv[tid] = jj * gg * gg * gg;
        }
        __syncthreads();
        if (slot < nr) {
            int t1 = (p + 1 == NP) ? tid - (NP - 1) : tid + 1;
            double W = 0.5 * (s.wv[tid] + s.wv[t1]) * dphi * s.rA[slot];
            int kf, kl;
            rg_deposit(s.H + s.HP, K_LO, nH, s.sv[tid], s.sv[t1], W * weight, kf, kl);
            kf_min = kf < kf_min ? kf : kf_min;
            kl_max = kl > kl_max ? kl : kl_max;
        }
        if (m0 + rpp < nsub) __syncthreads(); // ring slots are rewritten by the next pass
    }
    // publish the touched shift range
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int a = __shfl_xor_sync(0xffffffffu, kf_min, o), b = __shfl_xor_sync(0xffffffffu, kl_max, o);
        kf_min = a < kf_min ? a : kf_min;
        kl_max = b > kl_max ? b : kl_max;
    }
    if ((tid & 31) == 0 && kf_min <= kl_max) {
        atomicMin(&s.kmm[0], kf_min);
        atomicMax(&s.kmm[1], kl_max);
        atomicMin(&s.kmm[2], kf_min); // this annulus alone (work accounting)
        atomicMax(&s.kmm[3], kl_max);
    }
    __syncthreads();
    if (tid == 0) {
        if (s.kmm[2] <= s.kmm[3]) { s.kmm[4] += s.kmm[3] - s.kmm[2] + 1; s.kmm[5] += 1; }
        s.kmm[2] = 0x7fffffff; s.kmm[3] = -0x7fffffff;
    }
}

// acc[c] += sum_k H[k] in[M t + c - k], k in [kmin, kmax]  (sliding register window)
template <int M>
__device__ inline void convolve(const Smem &s, int K_LO, int nH, int HL, double (&acc)[M])
{
    int kmin = s.kmm[0], kmax = s.kmm[1];
    if (kmin < K_LO) kmin = K_LO;
    if (kmax > K_LO + nH - 1) kmax = K_LO + nH - 1;
    if (kmin > kmax) return;
    // group base aligned down to a multiple of M (H is zero-padded M in front, 2M behind)
    int g = (kmin >= 0) ? kmin / M : -((-kmin + M - 1) / M);
    int k0 = g * M;
    const double *base = s.inT + HL + (int)threadIdx.x;
    const double *Hk = s.H + s.HP + (k0 - K_LO);
    double w[M];
#pragma unroll
    for (int c = 0; c < M; c++) w[c] = base[c * s.RS - g];
    const double *pg = base - g - 1;
    for (int k = k0; k <= kmax; k += M) {
#pragma unroll
        for (int u = 0; u < M; u++) {
            double h = Hk[u];
#pragma unroll
            for (int c = 0; c < M; c++) acc[c] = fma(h, w[(c - u + M) % M], acc[c]);
            // window for the next tap: in[M t - (k+u+1)] lives in row (M-u-1) % M, column t - g - 1
            w[(M - 1 - u) % M] = pg[((M - u - 1) % M) * s.RS];
        }
        Hk += M;
        pg -= 1;
    }
}

template <int M>
__global__ void __launch_bounds__(RG_SPEC_THREADS) spectrum_kernel(SpecArgs A)
{
    RG_DYN_SMEM(smem_raw);
    Smem s;
    carve<M>(s, smem_raw, A.nH, A.HL, A.HH);
    const int tid = threadIdx.x;
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

    // zero the halos of the transposed buffer and the histogram once
    for (int i = tid; i < M * s.RS; i += T_) s.inT[i] = 0.0;
    for (int i = tid; i < A.nH + 3 * M + 4; i += T_) s.H[i] = 0.0;
    if (tid == 0) {
        s.kmm[0] = 0x7fffffff; s.kmm[1] = -0x7fffffff; s.kmm[2] = 0x7fffffff; s.kmm[3] = -0x7fffffff;
        s.kmm[4] = 0; s.kmm[5] = 0;
    }

    // per-bin constants of this thread
    double hnu[M], bbpre[M], wt[M];
#pragma unroll
    for (int c = 0; c < M; c++) {
        int j = j0 + c;
        bool ok = j < nE;
        hnu[c] = ok ? A.grid.hnu[j] : 1.0; // padding bins: 0 / (exp(1/kT) - 1) = 0
        bbpre[c] = ok ? A.grid.bbpre[j] : 0.0;
        wt[c] = ok ? A.grid.wt[j] : 0.0;
    }
    NtConst nt;
    rg_nt_prepare(nt, r);
    TabSel ts;
    if (rel) rg_tab_select(ts, A.tab, r.a, r.inc);
    const double Rg2 = r.Rg * r.Rg;
    const double dlnE = A.grid.dlnE;
    const double cosfac = r.cosinc / 0.5;
    double tot[M];
#pragma unroll
    for (int c = 0; c < M; c++) tot[c] = 0.0;
    int redsel = 0;
    __syncthreads();

    // ------------------------------------------------------------------
    // regions: 0 disc [r_w, r_out], 1 warm [r_h, r_w], 2 hot [risco, r_h]
    for (int region = 0; region < 3; region++) {
        int n_ann = region == 0 ? r.n_disc : (region == 1 ? r.n_warm : r.n_hot);
        double lg_in = region == 0 ? r.lg_rw : (region == 1 ? r.lg_rh : r.lg_risco);
        double lg_out = region == 0 ? r.lg_rout : (region == 1 ? r.lg_rw : r.lg_rh);
        double accC[M], accF[M]; // through-the-transfer sum (before unpack) and flat/non-rel sum
#pragma unroll
        for (int c = 0; c < M; c++) { accC[c] = 0.0; accF[c] = 0.0; }
        double fh[M] = {};       // hot: normalised shape Fnu_seed_hot / trapz (relagn.py:1315-1316)
        double Sflat = 0.0;      // hot: sum of L_ann 4 pi r dr over flat annuli
        bool any_conv = false;

        if (region == 2 && n_ann > 0) {
            // hot shape (relagn.py:1200-1223): donthcomp(Egrid, [gamma_h, kTe_h, kT_seed, 0, 0]) * h/keV
            double F[M];
#pragma unroll
            for (int c = 0; c < M; c++) F[c] = 0.0;
            if (r.has_hotshape) {
                int sys = r.sys_base + r.n_warm;
                const double *hd = A.header + (size_t)sys * 4;
                double xmin = hd[0], normfac = hd[1];
                int nth = (int)hd[2];
                const double *g_spt = A.sptot + (size_t)sys * RG_SPT_STRIDE;
                for (int i = tid; i <= nth; i += T_) s.spt[i] = g_spt[i];
                __syncthreads();
                double e_prev = j0 > 0 && j0 - 1 < nE ? A.grid.Egrid[j0 - 1] : 0.0;
                double pr_prev = j0 > 0 && j0 - 1 < nE
                                     ? nth_prim(s.spt, A.p10tab, xmin, nth, e_prev, A.grid.e511[j0 - 1]) /
                                           A.grid.ear2[j0 - 1]
                                     : 0.0;
#pragma unroll
                for (int c = 0; c < M; c++) {
                    int j = j0 + c;
                    if (j < nE) {
                        double e = A.grid.Egrid[j];
                        double pr = nth_prim(s.spt, A.p10tab, xmin, nth, e, A.grid.e511[j]) / A.grid.ear2[j];
                        double ph = j > 0 ? (0.5 * (pr + pr_prev) * (e - e_prev) * normfac) : 0.0;
                        F[c] = ph * (1.0 / RGK_KEV) * RGK_H;
                        e_prev = e; pr_prev = pr;
                    }
                }
            }
            double part = 0;
#pragma unroll
            for (int c = 0; c < M; c++) part += F[c] * wt[c];
            double shape_int = block_sum(part, s.red + redsel * NWARP);
            redsel ^= 1;
#pragma unroll
            for (int c = 0; c < M; c++) fh[c] = F[c] / shape_int; // 0/0 -> NaN as in the reference (relagn.py:1316)
            if (rel) {
#pragma unroll
                for (int c = 0; c < M; c++) {
                    int j = j0 + c;
                    s.inT[c * s.RS + A.HL + tid] = j < nE ? (A.grid.dE[j] * (fh[c] * 4)) / A.grid.Egrid[j] : 0.0;
                }
                // one histogram for the whole region
                for (int q = tid; q < A.nH; q += T_) s.H[s.HP + q] = 0.0;
                if (tid == 0) { s.kmm[0] = 0x7fffffff; s.kmm[1] = -0x7fffffff; }
            }
            __syncthreads();
        }

        for (int tile0 = 0; tile0 < n_ann && !(region == 2 && !rel); tile0 += T_) {
            int n_tile = n_ann - tile0 < T_ ? n_ann - tile0 : T_;
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
                s.aR1[tid] = r1; s.aR2[tid] = r2; s.aRm[tid] = rmid; s.aS1[tid] = s1; s.aS2[tid] = s2;
                s.aConv[tid] = (rel && lo <= 3 && hi <= 3) ? 1 : 0; // relagn.py:992
            }
            __syncthreads();

            if (region == 2) {
                // hot annuli: accumulate weighted histograms / flat weights only
                if (rel) {
                    for (int i = 0; i < n_tile; i++) {
                        if (s.aConv[i]) {
                            build_hist<M>(s, ts, s.aR1[i], s.aR2[i], dlnE, A.K_LO, A.nH, s.aS1[i]);
                            any_conv = true;
                        } else {
                            Sflat += s.aS1[i] * 2 * RGK_PI * 2 * s.aRm[i] * (s.aR2[i] - s.aR1[i]);
                        }
                    }
                }
                continue;
            }

            for (int i = 0; i < n_tile; i++) {
                double L[M];
                double part = 0;
                const double norm = s.aS2[i];
                if (region == 0) {
                    const double kT = s.aS1[i];
#pragma unroll
                    for (int c = 0; c < M; c++) {
                        double ef = exp(hnu[c] / kT) - 1;
                        L[c] = RGK_PI * (bbpre[c] / ef);
                        part += L[c] * wt[c];
                    }
                    if (A.grid.n_ext && tid < A.grid.n_ext) { // relagn.py:868-873
                        double ef = exp(A.grid.x_hnu[tid] / kT) - 1;
                        part += RGK_PI * (A.grid.x_bbpre[tid] / ef) * A.grid.x_wt[tid];
                    }
                } else {
                    // warm annulus: Kompaneets solution of system sys_base + (tile0+i)
                    int sys = r.sys_base + tile0 + i;
                    const double *hd = A.header + (size_t)sys * 4;
                    double xmin = hd[0], normfac = hd[1];
                    int nth = (int)hd[2];
                    const double *g_spt = A.sptot + (size_t)sys * RG_SPT_STRIDE;
                    for (int q = tid; q <= nth; q += T_) s.spt[q] = g_spt[q];
                    __syncthreads();
                    double e_prev = j0 > 0 && j0 - 1 < nE ? A.grid.Egrid[j0 - 1] : 0.0;
                    double pr_prev = j0 > 0 && j0 - 1 < nE
                                         ? nth_prim(s.spt, A.p10tab, xmin, nth, e_prev, A.grid.e511[j0 - 1]) /
                                               A.grid.ear2[j0 - 1]
                                         : 0.0;
#pragma unroll
                    for (int c = 0; c < M; c++) {
                        int j = j0 + c;
                        L[c] = 0.0;
                        if (j < nE) {
                            double e = A.grid.Egrid[j];
                            double pr = nth_prim(s.spt, A.p10tab, xmin, nth, e, A.grid.e511[j]) / A.grid.ear2[j];
                            double ph = j > 0 ? (0.5 * (pr + pr_prev) * (e - e_prev) * normfac) : 0.0;
                            L[c] = ph * (1.0 / RGK_KEV) * RGK_H;
                            e_prev = e; pr_prev = pr;
                        }
                        part += L[c] * wt[c];
                    }
                }
                double radiance = block_sum(part, s.red + redsel * NWARP);
                redsel ^= 1;
                if (radiance == 0) {
#pragma unroll
                    for (int c = 0; c < M; c++) L[c] = 0.0;
                } else {
#pragma unroll
                    for (int c = 0; c < M; c++) L[c] = norm * (L[c] / radiance);
                }
                if (s.aConv[i]) {
                    // pack (relagn.py:970-974; h and 4 pi d^2 cancel against the unpack at :1003-1006)
#pragma unroll
                    for (int c = 0; c < M; c++) {
                        int j = j0 + c;
                        s.inT[c * s.RS + A.HL + tid] = j < nE ? (A.grid.dE[j] * (L[c] * 4)) / A.grid.Egrid[j] : 0.0;
                    }
                    // fresh histogram
                    for (int q = tid; q < A.nH; q += T_) s.H[s.HP + q] = 0.0;
                    if (tid == 0) { s.kmm[0] = 0x7fffffff; s.kmm[1] = -0x7fffffff; }
                    __syncthreads();
                    build_hist<M>(s, ts, s.aR1[i], s.aR2[i], dlnE, A.K_LO, A.nH, 1.0);
                    convolve<M>(s, A.K_LO, A.nH, A.HL, accC);
                    any_conv = true;
                    __syncthreads(); // inT / H are rewritten by the next annulus
                } else {
                    // flat space (relagn.py:1013,1110) or non-relativistic sum (relagn.py:1046,1143)
                    double geo = 2 * RGK_PI * 2 * s.aRm[i] * (s.aR2[i] - s.aR1[i]);
                    if (rel) geo = geo * r.cosinc / 0.5;
#pragma unroll
                    for (int c = 0; c < M; c++) accF[c] += L[c] * geo;
                    if (region == 1) __syncthreads(); // s.spt is restaged by the next warm annulus
                }
            }
        }

        // ---- region epilogue ----
        if (n_ann > 0) {
            if (region == 2) {
                if (rel) {
                    if (any_conv) {
                        convolve<M>(s, A.K_LO, A.nH, A.HL, accC);
                        __syncthreads();
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
                    // leave the shared histogram clean for nothing else: hot is the last region
                } else {
                    double Lum = r.Ldiss + r.Lseed; // relagn.py:1396-1398
#pragma unroll
                    for (int c = 0; c < M; c++) accF[c] = Lum * fh[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < M; c++) {
                    int j = j0 + c;
                    double v = accF[c];
                    if (!rel) v = v * r.cosinc / 0.5; // relagn.py:1058,1155
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
    if (A.stats && tid == 0) {
        atomicAdd(&A.stats[0], (unsigned long long)s.kmm[4]);
        atomicAdd(&A.stats[1], (unsigned long long)s.kmm[5]);
    }
}

template <int M>
static int launch_spec(const SpecArgs &args, cudaStream_t st)
{
    size_t smem = rg_spectrum_smem_bytes(M, args.nH, args.HL, args.HH);
    if (smem > 220 * 1024) return rg_set_error(RG_E_GRID, "energy grid too fine for the shared-memory histogram (%zu B)", smem);
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
