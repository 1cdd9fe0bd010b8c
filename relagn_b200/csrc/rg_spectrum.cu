// rg_spectrum.cu -- the fused spectrum kernel.  The warps of a persistent CTA share one (parameter set, annulus
// block) item pulled from a device queue; every warp walks its own annuli of that item (disc, warm, hot), builds the
// local spectrum in registers, pushes it through the relativistic transfer and accumulates the observed spectrum
// in registers, without touching HBM between the parameter record and the final [nE] store.
//
// Reference path restated (src/python_version/relagn.py):
//   disc_annuli :839-896, warmComp_annuli :900-939, hotComp_annuli :1295-1318,
//   do_relDiscSpec :950-1027, do_relWarmCompSpec :1063-1124,
//   do_relHotCompSpec :1322-1384, do_nonrel* :1030-1059,1127-1156,1387-1401;
//   the Comptonised spectra arrive on the caller's grid from nthcomp_interp_kernel (pyNTHCOMP.donthcomp :57-76),
//   the transfer kernel of every annulus (the kyconv seam, :979-1001) as a shift histogram from hist_kernel.
//
// Work layout per warp.  The energy grid is cut into NC chunks of 256 bins;
// in chunk ch lane l owns the 8 consecutive bins j = 8 (32 ch + l) + c.  A
// chunk is the unit of skipping (Wien tails past the FP64 overflow point are
// neither evaluated nor convolved) and gives every lane 8 independent
// exp / divide / FMA chains.  Per annulus:
//   1. the g-shift histogram H[k] of the annulus (built by hist_kernel with a warp-shuffle segmented reduction
//      keyed by the shift bin -- no atomics, fixed summation order) is fetched into the warp's shared-memory slot;
//   2. the local spectrum (Planck / Kompaneets interpolation) is evaluated chunk
//      by chunk in registers, written to the warp's shared-memory buffer in
//      transposed form (row j mod 8, column j div 8, zero halos) together with
//      the warp-reduced nu-integral;
//   3. out_j += scale * sum_k H[k] in[j-k] runs per chunk as a sliding register
//      window: one conflict-free shared load + one broadcast load feed 8 FMAs.
// The hot region uses linearity: all hot annuli share one spectral shape, so
// their histograms are summed (weighted by annulus luminosity) and ONE
// convolution is done for the region.
#include "rg_kernels.h"


#define FULL 0xffffffffu
#define IMAX 0x7fffffff
#ifndef RG_LOCKSTEP
#define RG_LOCKSTEP 0 // > 0: block barrier every RG_LOCKSTEP rounds (power of two).  Round 1 gained 3 % from keeping the
                      // warps in one code region; with the round-2 kernel the barrier costs 2.4 % (r2z_ab2), so: off
#endif
#define NGRID 4 // per-bin constant arrays staged in shared memory: hnu, bbpre, wt in bin order + wt transposed
#ifndef RG_SPEC_W64
#define RG_SPEC_W64 6 // warps per CTA of the FP64 kernel on grids of <= 1024 bins (register budget 65536 / (32 W x CTAs per SM))
#endif
#ifndef RG_SPEC_CPS
#define RG_SPEC_CPS 2 // resident CTAs per SM of that kernel (RG_SPEC_W64 x RG_SPEC_CPS warps share the register file).
                      // Two CTAs of 6 warps instead of one of 12 (r2zj: 56.7 against 59.9 ms): an item's annuli go to 6
                      // warps instead of 12, so the warps' shares differ by one annulus in 21 instead of one in 11, and
                      // while one CTA waits for its slowest warp and folds its item the other keeps the SM busy
#endif

static __host__ __device__ inline int round_up2(int x) { return (x + 1) & ~1; }

int rg_spectrum_pick_NC(int nE)
{
    if (nE <= 4 * 256) return 4;
    if (nE <= 8 * 256) return 8;
    if (nE <= 20 * 256) return 20;
    return 0;
}

// shared memory per CTA: [per-bin constants (R, NC <= 8)] [reduction scratch, item slot, hot ranges]
//                        [per warp: transposed spectrum buffer (R) | histogram slot (double) | accumulators (R, NC > 4)]
size_t rg_spectrum_smem_bytes(int NC, int nH, int HL, int HH, int warps, int fp32)
{
    size_t rsz = fp32 ? 4 : 8;
    size_t w = 8 * (size_t)(HL + 32 * NC + HH) * rsz + (size_t)round_up2(nH + 24) * 8;
    if (NC > 4) w += 8 * (size_t)32 * NC * rsz; // accumulators (registers when NC <= 4)
    size_t b = w * warps + (18 + 16) * 8;
    if (NC <= 8) b += NGRID * 8 * (size_t)32 * NC * rsz; // grid constants (global memory for the finest grids)
    return b;
}

int rg_spectrum_pick_warps(int NC, int nH, int HL, int HH, int fp32)
{
    const int per_sm = (NC <= 4 && !fp32) ? RG_SPEC_CPS : 1; // (each resident CTA also pays 1 kB of system shared memory)
    for (int w = (NC <= 4 ? (fp32 ? 16 : RG_SPEC_W64) : 8); w >= 1; w--)
        if (rg_spectrum_smem_bytes(NC, nH, HL, HH, w, fp32) <= (size_t)(228 * 1024) / per_sm - 1024 - 512) return w;
    return 0;
}

#ifndef RG_CONV_HALFSTART
#define RG_CONV_HALFSTART 1
#endif
#ifndef RG_WARM_PREFETCH
#define RG_WARM_PREFETCH 1
#endif
#ifndef RG_WARM_BATCH
#define RG_WARM_BATCH 4 // warm annuli: 4 = one bin per lane and step, coalesced, RG_WARM_STEPS loads in flight; 2 / 1 = 8 bins per lane, loads of 2 chunks / the whole row first; 3 = cp.async into the buffer; 0 = per chunk
#endif
#ifndef RG_WARM_STEPS
#define RG_WARM_STEPS 16 // RG_WARM_BATCH 4: 32-bin steps of a warm row in flight per lane (divides 8 NC)
#endif
#ifndef RG_PLANCK_CHAINS
#define RG_PLANCK_CHAINS 5 // 32-bin steps of the Planck stage in flight per lane (independent exp / reciprocal chains; r2zh: 4 -> 60.6, 5 -> 59.6, 6 -> 61.4, 8 -> 60.1 ms)
#endif
#ifndef RG_WIEN_CUT
#define RG_WIEN_CUT 1 // skip the Wien-tail bins of an annulus whose share of the total is below FP64 resolution
#endif

template <typename T>
__device__ inline T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Planck function of one bin: pi * (2 h nu^3 / c^2) / (exp(x) - 1), x = h nu / kT (relagn.py:45-50).
// FP64: exp(x) - 1 as the reference writes it, then a reciprocal multiply (1/inf = 0 past the overflow point).
// FP32 mode: expm1f keeps the relative accuracy at small x, where exp(x) - 1 would cancel in single precision.
__device__ inline double planck_bin(double hnu, double bbpre, double inv_kT)
{
    double ef = rg_exp_pos(hnu * inv_kT) - 1;
    return (RGK_PI * bbpre) * __drcp_rn(ef);
}
__device__ inline float planck_bin(float hnu, float bbpre, float inv_kT)
{
    float ef = expm1f(hnu * inv_kT);
    return (3.14159265358979f * bbpre) / ef;
}

// tmp[c] = sum_k H[k] in[8 col + c - k] for k in [kmin, kmax]; Hs0 points at shift K_LO of a histogram that is
// zero-padded 8 in front and 16 behind; base = row 0 of the warp's buffer at this lane's column, RS = row stride.
template <typename R>
__device__ inline void convolve8(const double *Hs0, int K_LO, int nH, int kmin, int kmax, const R *base, int RS,
                                 R (&tmp)[8])
{
#pragma unroll
    for (int c = 0; c < 8; c++) tmp[c] = 0;
    if (kmin < K_LO) kmin = K_LO;
    if (kmax > K_LO + nH - 1) kmax = K_LO + nH - 1;
    if (kmin > kmax) return;
    const int g = (kmin >= 0) ? (kmin >> 3) : -((-kmin + 7) >> 3); // floor(kmin / 8)
    const int k0 = g * 8;
    const double *Hk = Hs0 + (k0 - K_LO);
    R w[8];
    const R *pg = base - g - 1;
#if RG_CONV_HALFSTART
    // the first group's lower half holds no tap when kmin sits in its upper half (warp-uniform): enter the window
    // in the state it has after four taps and run taps 4..7 only
    const bool half = kmin - k0 >= 4;
#pragma unroll
    for (int c = 0; c < 4; c++) w[c] = base[c * RS - g];
#pragma unroll
    for (int c = 4; c < 8; c++) w[c] = half ? pg[c * RS] : base[c * RS - g];
    int kb = k0;
    // (the histogram value of tap u + 1 is requested before the 8 FMAs of tap u: the shared-memory latency of the
    //  broadcast load hides behind them instead of in front of every tap -- the slot is zero padded 16 behind)
    R hn = (R)Hk[half ? 4 : 0];
    if (half) {
#pragma unroll
        for (int u = 4; u < 8; u++) {
            const R h = hn;
            hn = (R)Hk[u + 1];
#pragma unroll
            for (int c = 0; c < 8; c++) tmp[c] = fma(h, w[(c - u + 8) % 8], tmp[c]);
            w[(7 - u) % 8] = pg[((7 - u) % 8) * RS];
        }
        Hk += 8;
        pg -= 1;
        kb += 8;
    }
    for (int k = kb; k <= kmax; k += 8) {
#else
#pragma unroll
    for (int c = 0; c < 8; c++) w[c] = base[c * RS - g];
    for (int k = k0; k <= kmax; k += 8) {
#endif
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (u == 4 && k + 4 > kmax) break; // the last group's upper half holds no tap (warp-uniform)
#if RG_CONV_HALFSTART
            const R h = hn;
            hn = (R)Hk[u + 1];
#else
            const R h = (R)Hk[u];
#endif
#pragma unroll
            for (int c = 0; c < 8; c++) tmp[c] = fma(h, w[(c - u + 8) % 8], tmp[c]);
            // next tap needs in[8 col - (k+u+1)]: row 7-u, column col - g - 1
            w[(7 - u) % 8] = pg[((7 - u) % 8) * RS];
        }
        Hk += 8;
        pg -= 1;
    }
}

// per-bin constants: shared memory or global memory for the finest grids.  h nu, 2 h nu^3 / c^2 and the trapezoid
// weights in bin order (the Planck evaluation walks the bins lane by lane) and the weights once more transposed
// ([row c][column], bin 8 col + c: the layout of the 8-bins-per-lane stages)
template <int NC, bool GS, typename R>
struct GridAcc {
    const R *sm;       // [hnu | bbpre | wt][8 * 32 NC] | wt transposed [8][32 NC]
    const GridDesc *g;
    __device__ inline R hnu_n(int j) const { return GS ? sm[j] : (R)(j < g->nE ? __ldg(g->hnu + j) : 1.0); }
    __device__ inline R bbpre_n(int j) const { return GS ? sm[8 * 32 * NC + j] : (R)(j < g->nE ? __ldg(g->bbpre + j) : 0.0); }
    __device__ inline R wt_n(int j) const { return GS ? sm[2 * 8 * 32 * NC + j] : (R)(j < g->nE ? __ldg(g->wt + j) : 0.0); }
    __device__ inline R wt(int c, int col) const
    {
        if (GS) return sm[(3 * 8 + c) * (32 * NC) + col];
        int j = 8 * col + c;
        return (R)(j < g->nE ? __ldg(g->wt + j) : 0.0);
    }
};

// Accumulators: registers (statically indexed through a switch on the chunk, so that the chunk loops with their
// large bodies need not be unrolled) for NC <= 4, the warp's shared memory otherwise.
template <int NC, bool ACC_REG, typename R>
struct Acc {
    R r[ACC_REG ? NC * 8 : 1];
    R *s; // [8][32 NC] when !ACC_REG
    int lane;
    __device__ inline void zero()
    {
        if (ACC_REG) {
#pragma unroll
            for (int i = 0; i < NC * 8; i++) r[i] = 0;
        } else {
            for (int ch = 0; ch < NC; ch++)
#pragma unroll
                for (int c = 0; c < 8; c++) s[c * 32 * NC + 32 * ch + lane] = 0;
        }
    }
    __device__ inline void add(int ch, const R (&v)[8])
    {
        if (ACC_REG) {
            switch (ch) {
            case 0:
#pragma unroll
                for (int c = 0; c < 8; c++) r[c] += v[c];
                break;
            case 1:
#pragma unroll
                for (int c = 0; c < 8; c++) r[(NC > 1 ? 8 : 0) + c] += v[c];
                break;
            case 2:
#pragma unroll
                for (int c = 0; c < 8; c++) r[(NC > 2 ? 16 : 0) + c] += v[c];
                break;
            default:
#pragma unroll
                for (int c = 0; c < 8; c++) r[(NC > 3 ? 24 : 0) + c] += v[c];
                break;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 8; c++) s[c * 32 * NC + 32 * ch + lane] += v[c];
        }
    }
    // write the accumulators into the warp's transposed buffer (row c, column 32 ch + lane)
    __device__ inline void dump(R *inBase, int RS)
    {
        if (ACC_REG) {
#pragma unroll
            for (int ch = 0; ch < NC; ch++)
#pragma unroll
                for (int c = 0; c < 8; c++) inBase[c * RS + 32 * ch] = r[ch * 8 + c];
        } else {
            for (int ch = 0; ch < NC; ch++)
#pragma unroll
                for (int c = 0; c < 8; c++) inBase[c * RS + 32 * ch] = s[c * 32 * NC + 32 * ch + lane];
        }
    }
    // out[j] = acc (and restart) for this lane's bins; fully unrolled for the register variant
    __device__ inline void store(double *out, int nE, bool reset)
    {
        if (ACC_REG) {
#pragma unroll
            for (int ch = 0; ch < NC; ch++) {
                const int j0 = 8 * (32 * ch + lane);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (j0 + c < nE) out[j0 + c] = (double)r[ch * 8 + c];
                    if (reset) r[ch * 8 + c] = 0;
                }
            }
        } else {
            for (int ch = 0; ch < NC; ch++) {
                const int j0 = 8 * (32 * ch + lane);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    R *p = &s[c * 32 * NC + 32 * ch + lane];
                    if (j0 + c < nE) out[j0 + c] = (double)*p;
                    if (reset) *p = 0;
                }
            }
        }
    }
};

// R = double: the FP64 path (reference precision).  R = float (RG_FLAG_FP32): local spectra, the shared-memory
// buffer, the convolution and the accumulators in single precision -- the histogram, every per-annulus scalar, the
// Kompaneets solve and the normalisations stay FP64.
template <int NC, typename R>
__global__ void __launch_bounds__(NC <= 4 ? (sizeof(R) == 4 ? 512 : 32 * RG_SPEC_W64) : 256, (NC <= 4 && sizeof(R) == 8) ? RG_SPEC_CPS : 1) spectrum_kernel(SpecArgs A)
{
    constexpr bool GS = NC <= 8;       // grid constants in shared memory
    constexpr bool ACC_REG = NC <= 4;  // accumulators in registers
    constexpr int NCOL = 32 * NC;
    RG_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int nE = A.grid.nE;
    const bool rel = (A.flags & RG_FLAG_REL) != 0;
    const bool comps = (A.flags & RG_FLAG_COMPONENTS) != 0;
    const int RS = A.HL + NCOL + A.HH;
    const int HSZ = round_up2(A.nH + 24);

    // ---- shared memory carve: [grid constants (R)] [16 doubles + item slot] [per warp: inT (R) | Hslot (double) | acc (R)]
    //      (every piece is a multiple of 8 bytes for both R) ----
    unsigned char *smb = smem_raw;
    R *sgrid = reinterpret_cast<R *>(smb);
    if (GS) smb += (size_t)NGRID * 8 * NCOL * sizeof(R);
    double *s_red = reinterpret_cast<double *>(smb); smb += 16 * sizeof(double);
    int *s_misc = reinterpret_cast<int *>(smb); smb += 2 * sizeof(double);
    int *s_hot = reinterpret_cast<int *>(smb); smb += 16 * sizeof(double); // per warp: shift range of its hot histogram
    const size_t wd = 8 * (size_t)RS * sizeof(R) + (size_t)HSZ * sizeof(double) + (ACC_REG ? 0 : 8 * (size_t)NCOL * sizeof(R));
    unsigned char *const warp0 = smb;
    R *inT = reinterpret_cast<R *>(smb + wd * warp);                   // [8][RS]
    double *Hslot = reinterpret_cast<double *>(inT + 8 * (size_t)RS);  // [HSZ]: 8 zero pad, nH bins, >= 16 zero pad
    R *acc_s = reinterpret_cast<R *>(Hslot + HSZ);                     // [8][NCOL] when !ACC_REG
    (void)acc_s;

    // ---- CTA prologue: stage the per-bin constants (transposed), zero the warp buffers ----
    if (GS) {
        for (int i = tid; i < 8 * NCOL; i += nthreads) {
            int c = i / NCOL, col = i - c * NCOL;
            int j = 8 * col + c;
            bool ok = j < nE;
            sgrid[j] = (R)(ok ? A.grid.hnu[j] : 1.0); // padding bins: 0 / (exp(1/kT) - 1) = 0
            sgrid[8 * NCOL + j] = (R)(ok ? A.grid.bbpre[j] : 0.0);
            sgrid[2 * 8 * NCOL + j] = (R)(ok ? A.grid.wt[j] : 0.0);
            sgrid[(3 * 8 + c) * NCOL + col] = (R)(ok ? A.grid.wt[j] : 0.0);
        }
    }
    for (int i = lane; i < 8 * RS; i += 32) inT[i] = 0;
    for (int i = lane; i < HSZ; i += 32) Hslot[i] = 0.0;
    GridAcc<NC, GS, R> G;
    G.sm = sgrid; G.g = &A.grid;
    R *const inBase = inT + A.HL + lane; // row 0, this lane's column of chunk 0
    unsigned long long st_w = 0, st_n = 0;

    // The warps of a CTA work on ONE (parameter set, annulus block) item at a time, annulus g -> warp g mod nwarps.
    // They exchange no data while walking the annuli; the per-round barrier only keeps them in the same code
    // region (one instruction-cache working set per SM instead of eight).
    for (;;) {
        __syncthreads();
        if (tid == 0) s_misc[0] = atomicAdd(A.queue, 1);
        __syncthreads();
        const int item = s_misc[0];
        if (item >= A.P * A.NB) break;
        const int set = item / A.NB, blk = item - set * A.NB;
        const SetRec &r = A.recs[set];
        // destinations: component mode -> per-warp partial arrays [item][warp][3][nE] (folded by the reduce kernel);
        // total-only mode -> the CTA folds its warps in shared memory and stores once
        double *part_w = comps ? A.partial + (((size_t)item * nwarps + warp) * 3) * nE : nullptr;
        double *out_tot = comps ? nullptr : (A.NB == 1 ? A.out + (size_t)set * nE : A.partial + (size_t)item * nE);
        if (r.status != RG_S_OK) { // ValueError sets: zeros (the host class raises)
            if (comps) { for (int j = lane; j < 3 * nE; j += 32) part_w[j] = 0.0; }
            else { for (int j = tid; j < nE; j += nthreads) out_tot[j] = 0.0; }
            continue;
        }
        const int n_tot = r.n_disc + r.n_warm + r.n_hot;
        const int per = (n_tot + A.NB - 1) / A.NB;
        const int g_begin = blk * per, g_end = (g_begin + per < n_tot) ? g_begin + per : n_tot;
        const int g_warm0 = r.n_disc, g_hot0 = r.n_disc + r.n_warm;
        const bool hot_here = r.n_hot > 0 && g_end > g_hot0 && g_begin < n_tot;       // item holds hot annuli
        const bool owns_hot0 = r.n_hot > 0 && g_begin <= g_hot0 && g_hot0 < g_end;    // ... the first of them
        const bool need_shape = rel ? hot_here : owns_hot0;

        // ---- hot spectral shape (relagn.py:1200-1223): donthcomp(Egrid, [gamma_h, kTe_h, kT_seed, 0, 0]) * h / keV was
        //      put on the grid by nthcomp_kernel; here only shape_int = trapz(Fnu, nu), once per CTA
        double shape_int = 0.0;
        const double *hot_F = nullptr;
        if (need_shape) {
            if (r.has_hotshape) hot_F = A.farr + (size_t)r.hot_sys * nE;
            double part = 0;
            if (hot_F)
                for (int j = tid; j < nE; j += nthreads) part += __ldg(hot_F + j) * __ldg(A.grid.wt + j);
            part = warp_sum(part);
            if (lane == 0) s_red[warp] = part;
            __syncthreads();
            for (int w = 0; w < nwarps; w++) shape_int += s_red[w];
        }

        NtConst nt;
        rg_nt_prepare(nt, r);
        // shift range of the histograms (hist_kernel builds them on the same range): widest support over the table
        const int K_LO = A.K_LO, nH = A.nH;
#if RG_WIEN_CUT
        // reference annulus of the Wien cut (wienref_kernel): the disc annulus with the highest colour temperature kT* and
        // the shift k* of the strongest tap of its transfer histogram -- see the cut in the prologue below
        double kT_ref = 0.0;
        int k_ref = 0;
        if (A.wien && r.n_disc > 0) {
            const double2 wr = __ldg(A.wien + set);
            kT_ref = wr.x; k_ref = (int)wr.y;
        }
#endif
        const double Rg2 = r.Rg * r.Rg;
        const double cosfac = r.cosinc / 0.5;

        Acc<NC, ACC_REG, R> acc;
        acc.s = acc_s; acc.lane = lane;
        acc.zero();
        if (comps) { for (int j = lane; j < 3 * nE; j += 32) part_w[j] = 0.0; }
        __syncwarp();
        int cur_region = 0;                // component mode: region the accumulators currently belong to
        int hz_min = IMAX, hz_max = -IMAX; // touched range of the H slot (for re-zeroing)
        int live_steps = 8 * NC;           // 32-bin steps of the warp's buffer that may hold non-zero values
        bool hot_any_conv = false, hot_mine = false;
        int hk_min = IMAX, hk_max = -IMAX;
        double Sflat = 0.0;

        const int n_rounds = (g_end - g_begin + nwarps - 1) / nwarps;
        for (int m0 = 0; m0 < n_rounds; m0 += 32) {
            // ---- prologue: lane m prepares this warp's annulus of round m0 + m: g = g_begin + warp + nwarps (m0 + m),
            //      counted top-down through disc [r_w, r_out], warm [r_h, r_w], hot [risco, r_h] ----
            double p_s1 = 0, p_s2 = 0, p_geo = 0;
            int p_flag = 0, p_jz = nE, p_region = -1, p_loc = 0, p_kmn = 0, p_nk = 0;
            {
                const int g = g_begin + warp + nwarps * (m0 + lane);
                if (m0 + lane < n_rounds && g < g_end) {
                    const int region = g < g_warm0 ? 0 : (g < g_hot0 ? 1 : 2);
                    const int a = g - (region == 0 ? 0 : (region == 1 ? g_warm0 : g_hot0));
                    const double2 eg = __ldg(A.edges + r.ann_base + g); // (lo, hi) in log10 r (edges_kernel)
                    const double lo = eg.x, hi = eg.y;
                    double r1 = pow(10.0, lo), r2 = pow(10.0, hi);
                    double rmid = pow(10.0, lo + r.dlog_r / 2);
                    double T4 = rg_tnt4(nt, rmid);
                    if (region == 0) { // relagn.py:857-883
                        double Tann = pow(T4, 0.25);
                        double fcol_r = r.fcol < 0 ? rg_fcol(Tann) : r.fcol;
                        Tann = Tann * fcol_r;
                        double t = Tann / fcol_r;
                        p_s1 = RGK_KB * Tann;
                        p_s2 = RGK_SB * (t * t * t * t) * Rg2;
                        // first bin whose Planck exponent overflows (exp(x) = inf for x > 709.78 -> B = 0 exactly)
                        // (single precision: expm1f overflows past x = 88.7)
                        double lim = (sizeof(R) == 4 ? 89.0 : 710.0) * p_s1;
#if RG_WIEN_CUT
                        // Wien cut: an input bin h nu of this annulus reaches the observer through tap k at h nu e^(k dlnE),
                        // where the reference annulus (kT*, tap k*) alone contributes C*; in the Wien regime the ratio of
                        // the two shares is P exp(-h nu [1/kT - e^((k - k*) dlnE) / kT*]) with a prefactor P (ratio of areas,
                        // norms and tap weights) below 1e15.  Past h nu = 80 / bracket(k = k_max of this annulus) the share
                        // is below 1e-18 of the total for every tap: those bins are neither evaluated nor convolved.
                        // (only for annuli whose Planck peak lies on the grid, kT >= h nu_0: the reference normalises B by its
                        //  integral over the GRID, relagn.py:883-889, so an annulus that shows only its Wien tail is scaled
                        //  up by an unbounded factor and the bound on P does not hold)
                        if (kT_ref > 0 && p_s1 >= __ldg(A.grid.hnu)) {
                            int kmax_i = 0;
                            if (rel && lo <= 3 && hi <= 3) {
                                const int2 hd = __ldg(reinterpret_cast<const int2 *>(A.hist + (size_t)(r.ann_base + g) * A.hstride));
                                kmax_i = hd.x + hd.y - 1;
                            }
                            const double br = 1.0 / p_s1 - (rel ? exp((kmax_i - k_ref) * A.grid.dlnE) : 1.0) / kT_ref;
                            if (br > 0 && 80.0 / br < lim) lim = 80.0 / br;
                        }
#endif
                        int lo_j = 0, hi_j = nE;
                        while (lo_j < hi_j) {
                            int m = (lo_j + hi_j) >> 1;
                            if (__ldg(A.grid.hnu + m) > lim) hi_j = m; else lo_j = m + 1;
                        }
                        p_jz = lo_j;
                    } else if (region == 1) { // relagn.py:919-932
                        double Tann = pow(T4, 0.25);
                        p_s2 = RGK_SB * (Tann * Tann * Tann * Tann) * Rg2;
                    } else { // relagn.py:1309-1314
                        double Ldiss_ann = RGK_SB * T4 * Rg2;
                        double Lseed_ann = r.Lseed / (4 * RGK_PI * rmid * (r2 - r1) * r.n_hot);
                        p_s1 = Ldiss_ann + Lseed_ann;
                    }
                    // flat space (relagn.py:1013,1110,1370) / non-relativistic (relagn.py:1046,1143) area factor
                    p_geo = 2 * RGK_PI * 2 * rmid * (r2 - r1);
                    if (region != 2) p_geo = p_geo * r.cosinc / 0.5;
                    p_flag = (rel && lo <= 3 && hi <= 3) ? 1 : 0; // relagn.py:992
                    p_region = region; p_loc = a;
                    if (p_flag) { // support of the annulus' shift histogram (slot 0 of its arena row)
                        const int2 hd = __ldg(reinterpret_cast<const int2 *>(A.hist + (size_t)(r.ann_base + g) * A.hstride));
                        p_kmn = hd.x; p_nk = hd.y;
                    }
                }
            }
            const int mt = n_rounds - m0 < 32 ? n_rounds - m0 : 32;
            for (int mi = 0; mi < mt; mi++) {
#if RG_LOCKSTEP > 0
                if (((m0 + mi) & (RG_LOCKSTEP - 1)) == 0) __syncthreads(); // lockstep only: no data crosses warps here
#endif
                const int region = __shfl_sync(FULL, p_region, mi);
                if (region < 0) continue; // no annulus left for this warp in this round
                if (region == 2 && !rel) continue; // non-relativistic hot component needs no annuli (relagn.py:1387-1401)
                const double s1 = __shfl_sync(FULL, p_s1, mi), s2 = __shfl_sync(FULL, p_s2, mi);
                const double geo = __shfl_sync(FULL, p_geo, mi);
                const bool conv = __shfl_sync(FULL, p_flag, mi) != 0;
                const int jz = __shfl_sync(FULL, p_jz, mi);
                const int loc = __shfl_sync(FULL, p_loc, mi);
#if RG_WARM_PREFETCH
                // the warp's next annulus, if warm: its Comptonised spectrum (8 nE bytes of the farr arena, far larger
                // than L2) is requested into L2 now, a whole annulus ahead of the loads that consume it
                if (mi + 1 < mt) {
                    const int nreg = __shfl_sync(FULL, p_region, mi + 1), nloc = __shfl_sync(FULL, p_loc, mi + 1);
                    if (nreg == 1) {
                        const char *row = reinterpret_cast<const char *>(A.farr + (size_t)(r.sys_base + nloc) * nE);
                        for (int b = lane * 128; b < nE * 8; b += 32 * 128) RG_PREFETCH_L2(row + b);
                    }
                }
#endif
                if (comps && region != cur_region && region != 2) { // (hot contributes after the loop)
                    acc.store(part_w + (size_t)cur_region * nE, nE, true);
                    cur_region = region;
                }
                // The annulus' shift histogram (built by hist_kernel) is requested here and installed in the warp's
                // slot only when the convolution needs it, so that the loads fly during the local-spectrum evaluation.
                // disc / warm: the slot is re-zeroed where the previous annulus touched it.  hot: the slot turns into
                // the region histogram -- every hot annulus adds its own, weighted by the annulus luminosity (the warp
                // meets its hot annuli last, so the slot is free by then).
                int kmn = 0, kmx = -1, nk = 0;
                const double *hrow = nullptr;
                double h0 = 0.0, h1 = 0.0;
                if (conv) {
                    const int g = (region == 0 ? 0 : (region == 1 ? g_warm0 : g_hot0)) + loc;
                    hrow = A.hist + (size_t)(r.ann_base + g) * A.hstride + 1;
                    kmn = __shfl_sync(FULL, p_kmn, mi); nk = __shfl_sync(FULL, p_nk, mi);
                    kmx = kmn + nk - 1;
                    if (lane < nk) h0 = __ldg(hrow + lane);
                    if (lane + 32 < nk) h1 = __ldg(hrow + lane + 32);
                }
                auto install_hist = [&](bool fresh, double wgt) {
                    if (fresh) {
                        for (int k = hz_min + lane; k <= hz_max; k += 32) {
                            int q = k - K_LO;
                            if (q >= 0 && q < nH) Hslot[8 + q] = 0.0;
                        }
                        hz_min = IMAX; hz_max = -IMAX;
                        __syncwarp();
                    }
                    double *Hd = Hslot + 8 + (kmn - K_LO);
                    if (lane < nk) Hd[lane] += wgt * h0;
                    if (lane + 32 < nk) Hd[lane + 32] += wgt * h1;
                    for (int k = lane + 64; k < nk; k += 32) Hd[k] += wgt * __ldg(hrow + k);
                    __syncwarp();
                    if (nk > 0) {
                        hz_min = kmn < hz_min ? kmn : hz_min;
                        hz_max = kmx > hz_max ? kmx : hz_max;
                        st_w += (unsigned long long)nk; st_n += 1;
                    }
                };
                if (conv && region == 2) install_hist(!hot_any_conv, s1);
                if (region == 2) {
                    hot_mine = true;
                    if (conv) {
                        hk_min = kmn < hk_min ? kmn : hk_min;
                        hk_max = kmx > hk_max ? kmx : hk_max;
                        hot_any_conv = true;
                    } else {
                        Sflat += s1 * geo;
                    }
                    continue;
                }

                // ---- local spectrum into the warp's transposed buffer (row j mod 8, column j div 8) ----
                R part = 0;
                if (region == 0) {
                    // Planck function, one bin per lane and step (bins 32 g + lane): the live range [0, jz) is walked in
                    // steps of 32 bins -- in the 8-bins-per-lane layout of the transfer stage the unit of skipping would
                    // be a 256-bin chunk -- four steps at a time for four independent exp / reciprocal chains per lane.
                    // x = h nu / kT as a multiplication by 1/kT (one division per annulus instead of one per bin; the
                    // last-bit difference in x moves exp(x) by < 1e-13 relative)
                    const R inv_kT = (R)(1.0 / s1);
                    R *const inLane = inT + A.HL + (lane & 7) * RS + (lane >> 3); // bin `lane`; bin 32 g + lane: + 4 g
                    const int n_steps = (jz + 31) >> 5;
                    int gq = 0;
                    for (; gq + RG_PLANCK_CHAINS <= n_steps; gq += RG_PLANCK_CHAINS) {
                        R Lq[RG_PLANCK_CHAINS];
#pragma unroll
                        for (int q = 0; q < RG_PLANCK_CHAINS; q++) {
                            const int j = 32 * (gq + q) + lane;
                            Lq[q] = j < jz ? planck_bin(G.hnu_n(j), G.bbpre_n(j), inv_kT) : (R)0;
                        }
#pragma unroll
                        for (int q = 0; q < RG_PLANCK_CHAINS; q++) {
                            part += Lq[q] * G.wt_n(32 * (gq + q) + lane);
                            inLane[4 * (gq + q)] = Lq[q];
                        }
                    }
                    for (; gq < n_steps; gq++) {
                        const int j = 32 * gq + lane;
                        const R Lv = j < jz ? planck_bin(G.hnu_n(j), G.bbpre_n(j), inv_kT) : (R)0;
                        part += Lv * G.wt_n(j);
                        inLane[4 * gq] = Lv;
                    }
                    // zeros where the previous annulus of this warp left values beyond this one's live range
                    for (int gz = n_steps; gz < live_steps; gz++) inLane[4 * gz] = (R)0;
                    live_steps = n_steps;
                } else {
                    // warm annulus: its Comptonised spectrum is already on the grid (nthcomp kernels)
                    const double *w_F = A.farr + (size_t)(r.sys_base + loc) * nE;
#if RG_WARM_BATCH == 3
                    if (sizeof(R) == 8) {
                        // the row goes global -> shared by asynchronous 8-byte copies, every element straight to its
                        // place in the transposed buffer (bin 8 col + c: row c, column col): no register staging, all
                        // 8 NC copies of a lane in flight at once -- one trip to L2 per annulus -- then the
                        // trapezoid sum reads the buffer back in the order of the register variant
                        for (int ch = 0; ch < NC; ch++) {
                            const int j0 = 8 * (32 * ch + lane);
#pragma unroll
                            for (int c = 0; c < 8; c++) {
                                const bool ok = j0 + c < nE;
                                RG_CP_ASYNC8(reinterpret_cast<double *>(&inBase[c * RS + 32 * ch]), ok ? w_F + j0 + c : w_F, ok);
                            }
                        }
                        RG_CP_ASYNC_WAIT();
                        __syncwarp();
                        for (int ch = 0; ch < NC; ch++) {
                            const int col = 32 * ch + lane;
                            if (8 * col < nE) {
#pragma unroll
                                for (int c = 0; c < 8; c++) part += inBase[c * RS + 32 * ch] * G.wt(c, col);
                            }
                        }
                    } else
#endif
#if RG_WARM_BATCH == 4
                    if (true) {
                        // one bin per lane and step, as the Planck stage walks the grid: a warp's load is one contiguous
                        // 256-byte run of the row (with 8 consecutive bins per lane every load touched 32 sectors), the
                        // stores into the transposed buffer are conflict-free (row stride = 2 mod 16), RG_WARM_STEPS
                        // loads in flight per lane
                        R *const inLane = inT + A.HL + (lane & 7) * RS + (lane >> 3); // bin `lane`; bin 32 g + lane: + 4 g
                        constexpr int WS = 8 * NC < RG_WARM_STEPS ? 8 * NC : RG_WARM_STEPS; // (8 NC is a multiple of it)
                        static_assert((8 * NC) % WS == 0, "RG_WARM_STEPS must divide the steps of a row");
#pragma unroll 1
                        for (int g0 = 0; g0 < 8 * NC; g0 += WS) {
                            R L[WS];
#pragma unroll
                            for (int q = 0; q < WS; q++) {
                                const int j = 32 * (g0 + q) + lane;
                                L[q] = (R)(j < nE ? __ldg(w_F + j) : 0.0);
                            }
#pragma unroll
                            for (int q = 0; q < WS; q++) {
                                part += L[q] * G.wt_n(32 * (g0 + q) + lane);
                                inLane[4 * (g0 + q)] = L[q];
                            }
                        }
                    } else
#endif
#if RG_WARM_BATCH
                    if (ACC_REG) {
                        // all loads of the row first (the conv registers are free in this stage), then the sums in
                        // the same order: one trip to L2 per annulus instead of one per 256-bin chunk
                        constexpr int WB = NC <= 4 ? (RG_WARM_BATCH == 1 ? NC : (NC >= 2 ? 2 : 1)) : 1; // chunks per batch
#pragma unroll
                        for (int cb = 0; cb < (NC <= 4 ? NC : 1); cb += WB) {
                            R L[WB][8];
#pragma unroll
                            for (int b = 0; b < WB; b++) {
                                const int j0 = 8 * (32 * (cb + b) + lane);
#pragma unroll
                                for (int c = 0; c < 8; c++) L[b][c] = (R)(j0 + c < nE ? __ldg(w_F + j0 + c) : 0.0);
                            }
#pragma unroll
                            for (int b = 0; b < WB; b++) {
                                const int col = 32 * (cb + b) + lane, j0 = 8 * col;
                                if (j0 < nE) {
#pragma unroll
                                    for (int c = 0; c < 8; c++) part += L[b][c] * G.wt(c, col);
                                }
#pragma unroll
                                for (int c = 0; c < 8; c++) inBase[c * RS + 32 * (cb + b)] = L[b][c];
                            }
                        }
                    } else
#endif
                    for (int ch = 0; ch < NC; ch++) {
                        const int col = 32 * ch + lane, j0 = 8 * col;
                        R L[8];
#pragma unroll
                        for (int c = 0; c < 8; c++) L[c] = 0;
                        if (j0 < nE) {
#pragma unroll
                            for (int c = 0; c < 8; c++) {
                                L[c] = (R)(j0 + c < nE ? __ldg(w_F + j0 + c) : 0.0);
                                part += L[c] * G.wt(c, col);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 8; c++) inBase[c * RS + 32 * ch] = L[c];
                    }
                    live_steps = 8 * NC;
                }
                // (the reference packs W/Hz into photons per bin before kyconv and unpacks afterwards -- x dE 4 / E,
                //  relagn.py:970-974; x E / dE, :1003-1006: on the geometric grid the transfer needs, dE / E is the same
                //  for every bin (to rounding), so the two factors are one constant, folded into the histogram's scale)
                if (region == 0 && A.grid.n_ext) { // relagn.py:868-873
                    for (int q = lane; q < A.grid.n_ext; q += 32) {
                        double ef = exp(A.grid.x_hnu[q] / s1) - 1;
                        part += (R)(RGK_PI * (A.grid.x_bbpre[q] / ef) * A.grid.x_wt[q]);
                    }
                }
                const double radiance = (double)warp_sum(part);
                __syncwarp();
                // Lnu = norm * (B / radiance), zero when the annulus emits nothing (relagn.py:885-889, 933-937)
                // (the histogram is installed already scaled: norm / radiance x the pack-unpack constant)
                // (FP32 mode keeps the histogram unscaled -- the scale alone can exceed the single-precision range -- and
                //  multiplies afterwards in double)
                // norm / radiance can exceed the double range although norm (B / radiance), what the reference computes,
                // does not: a grid that shows only the far Wien tail of a cold annulus (a hard X-ray band) has B ~ 1e-300
                // and the reference scales it up to the annulus' full luminosity.  Then the buffer is multiplied by
                // 2^600 (exact) and the scale carries 2^-600.
                double nr = radiance != 0 ? s2 / radiance : 0.0;
                if (sizeof(R) == 8 && !(fabs(nr) <= 1e290)) { // (the area factor, up to ~1e12, still multiplies it)
                    const double up = 4.149515568880993e180, dn = 2.409919865102884e-181; // 2^600, 2^-600
                    nr = (s2 * dn) / radiance;
                    for (int i = lane; i < 8 * RS; i += 32) inT[i] = (R)((double)inT[i] * up);
                    __syncwarp();
                }
                const double hscale = nr * A.pkupk;
                if (conv) install_hist(true, sizeof(R) == 8 ? hscale : 1.0);
                if (radiance != 0) {
                    if (conv) {
                        for (int ch = 0; ch < NC; ch++) {
                            // taps whose sources all lie in the zero zones (j - k >= jz or j - k < 0) are skipped
                            int k_lo = 256 * ch - (region == 0 ? jz : nE) + 1, k_hi = 256 * ch + 255;
                            k_lo = k_lo > kmn ? k_lo : kmn;
                            k_hi = k_hi < kmx ? k_hi : kmx;
                            if (k_lo > k_hi) continue;
                            const int col = 32 * ch + lane;
                            R tmp[8];
                            convolve8<R>(Hslot + 8, K_LO, nH, k_lo, k_hi, inBase + 32 * ch, RS, tmp);
                            if (sizeof(R) != 8) {
#pragma unroll
                                for (int c = 0; c < 8; c++) tmp[c] = (R)((double)tmp[c] * hscale);
                            }
                            acc.add(ch, tmp);
                        }
                    } else {
                        const double scale = nr * geo;
                        const int nch_live = region == 0 ? ((jz + 255) >> 8) : NC; // chunks holding a non-zero bin
                        for (int ch = 0; ch < nch_live; ch++) {
                            R v[8];
#pragma unroll
                            for (int c = 0; c < 8; c++) v[c] = (R)(scale * (double)inBase[c * RS + 32 * ch]);
                            acc.add(ch, v);
                        }
                    }
                }
                __syncwarp(); // the buffer is rewritten by the next annulus
            }
        }

        // ---- hot region.  Every warp holds the weighted histogram of its own hot annuli; the region is linear in
        //      its one spectral shape, so the CTA sums those histograms (fixed warp order) and does ONE convolution,
        //      shared between the warps by tap range -- instead of one full convolution per warp ----
        const bool nonrel_hot = !rel && owns_hot0 && warp == 0; // (Ldiss + Lseed) * Fnu / trapz(Fnu), relagn.py:1387-1401
        int t_lo = 0, t_hi = -1; // this warp's share of the taps
        bool cta_conv = false;
        // the summed histogram carries the region's luminosity (up to ~1e40 W): it is stored in units of Ldiss + Lseed so
        // that the single-precision variant can hold it, and the convolution is scaled back in double
        const double Lhot = (r.Ldiss + r.Lseed) > 0 ? r.Ldiss + r.Lseed : 1.0;
        double conv_unit = 1.0;
        if (rel && hot_here) {   // (item-level condition: the same for every warp of the CTA)
            if (lane == 0) {
                s_hot[2 * warp] = hot_any_conv ? hk_min : IMAX;
                s_hot[2 * warp + 1] = hot_any_conv ? hk_max : -IMAX;
            }
            __syncthreads();
            int ck_min = IMAX, ck_max = -IMAX;
            for (int w = 0; w < nwarps; w++) {
                ck_min = s_hot[2 * w] < ck_min ? s_hot[2 * w] : ck_min;
                ck_max = s_hot[2 * w + 1] > ck_max ? s_hot[2 * w + 1] : ck_max;
            }
            if (ck_min < K_LO) ck_min = K_LO;
            if (ck_max > K_LO + nH - 1) ck_max = K_LO + nH - 1;
            const int T = (ck_max - ck_min + nwarps) / nwarps; // taps per warp
            // (the sums below sit in 8 registers per lane: wider shares -- the finest grids run 2 warps per CTA over
            //  > 1000 shifts -- keep one convolution per warp)
            cta_conv = ck_min <= ck_max && T <= 256;
            if (!cta_conv && hot_any_conv) { t_lo = hk_min; t_hi = hk_max; }
            if (cta_conv) {
                t_lo = ck_min + warp * T;
                t_hi = t_lo + T - 1 < ck_max ? t_lo + T - 1 : ck_max;
                double hs[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    hs[q] = 0.0;
                    const int k = t_lo + lane + 32 * q;
                    if (k <= t_hi)
                        for (int w = 0; w < nwarps; w++)
                            if (s_hot[2 * w] <= s_hot[2 * w + 1]) // warps without hot annuli hold disc / warm leftovers
                                hs[q] += reinterpret_cast<const double *>(reinterpret_cast<const R *>(warp0 + wd * w) + 8 * (size_t)RS)[8 + k - K_LO];
                }
                __syncthreads(); // every warp has read the others' slots
                for (int k = hz_min + lane; k <= hz_max; k += 32) {
                    int q = k - K_LO;
                    if (q >= 0 && q < nH) Hslot[8 + q] = 0.0;
                }
                hz_min = IMAX; hz_max = -IMAX;
                __syncwarp();
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int k = t_lo + lane + 32 * q;
                    if (k <= t_hi) Hslot[8 + k - K_LO] = hs[q] * (1.0 / Lhot);
                }
                conv_unit = Lhot;
                if (t_lo <= t_hi) { hz_min = t_lo; hz_max = t_hi; }
                __syncwarp();
            }
        }
        if ((rel && (hot_mine || t_lo <= t_hi)) || nonrel_hot) {
            if (comps && cur_region != 2) {
                acc.store(part_w + (size_t)cur_region * nE, nE, true);
                cur_region = 2;
            }
            __syncwarp();
            for (int ch = 0; ch < NC; ch++) {
                const int col = 32 * ch + lane;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int j = 8 * col + c;
                    inBase[c * RS + 32 * ch] = (hot_F && j < nE) ? (R)__ldg(hot_F + j) : (R)0;
                }
            }
            __syncwarp();
            const double Lum = r.Ldiss + r.Lseed;
            const bool conv_here = rel && t_lo <= t_hi;
            // 1 / trapz(Fnu) and the factors of relagn.py:1362 once per item instead of three divisions per bin and warp
            // (0 / 0 -> NaN as the reference's division gives, relagn.py:1316: 1/0 = inf, 0 * inf = NaN)
            const double r_shape = 1.0 / shape_int;
            const double conv_fac = ((conv_unit / shape_int) / cosfac) * A.pkupk; // (pack x unpack: one constant, see above)
            for (int ch = 0; ch < NC; ch++) {
                const int col = 32 * ch + lane;
                R tmp[8];
#pragma unroll
                for (int c = 0; c < 8; c++) tmp[c] = 0;
                if (conv_here) convolve8<R>(Hslot + 8, K_LO, nH, t_lo, t_hi, inBase + 32 * ch, RS, tmp);
                R v[8];
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    v[c] = 0;
                    if (8 * col + c < nE) {
                        // F / shape_int: 0/0 -> NaN as the reference (:1316)
                        double F = (double)inBase[c * RS + 32 * ch];
                        if (rel) {
                            double vv = hot_mine ? Sflat * (F * r_shape) : 0.0;
                            if (conv_here) vv += (double)tmp[c] * conv_fac; // relagn.py:1362
                            v[c] = (R)vv;
                        } else {
                            v[c] = (R)(Lum * (F * r_shape));
                        }
                    }
                }
                acc.add(ch, v);
            }
            __syncwarp();
        }

        // ---- item epilogue ----
        if (comps) {
            acc.store(part_w + (size_t)cur_region * nE, nE, false);
        } else {
            // fold the warps in shared memory (fixed order) and store once
            acc.dump(inBase, RS);
            __syncthreads();
            for (int j = tid; j < nE; j += nthreads) {
                const size_t off = (size_t)(j & 7) * RS + A.HL + (j >> 3);
                R v = 0;
                for (int w = 0; w < nwarps; w++) v += reinterpret_cast<const R *>(warp0 + wd * w)[off];
                out_tot[j] = (double)v;
            }
        }
        // leave the H slot clean for the next item
        for (int k = hz_min + lane; k <= hz_max; k += 32) {
            int q = k - K_LO;
            if (q >= 0 && q < nH) Hslot[8 + q] = 0.0;
        }
    }
    if (A.stats && lane == 0 && st_n) {
        atomicAdd(&A.stats[0], st_w);
        atomicAdd(&A.stats[1], st_n);
    }
}

// Reference annulus of the Wien cut, one warp per parameter set: the disc annulus with the highest colour temperature
// kT* (relagn.py:857-866) and the shift k* of the strongest tap of its transfer histogram (0 when it does not go through
// the transfer).  At an observed energy where annulus i's share is below 1e-18 of the reference annulus' share (itself
// a lower bound of the total there), annulus i cannot move the FP64 sum: spectrum_kernel skips its Planck function and
// its transfer there.  wien[set] = (kT*, k*).
__global__ void __launch_bounds__(256) wienref_kernel(const SetRec *__restrict__ recs, int P, const double2 *__restrict__ edges,
                                                      const double *__restrict__ hist, int hstride, int rel,
                                                      double2 *__restrict__ wien)
{
    const int lane = threadIdx.x & 31;
    const int set = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (set >= P) return;
    const SetRec &r = recs[set];
    if (r.status != RG_S_OK || r.n_disc <= 0) {
        if (lane == 0) wien[set] = make_double2(0.0, 0.0);
        return;
    }
    NtConst nt;
    rg_nt_prepare(nt, r);
    double best = 0.0;
    int bestg = -1;
    for (int g = lane; g < r.n_disc; g += 32) {
        const double lo = __ldg(edges + r.ann_base + g).x;
        const double T4 = rg_tnt4(nt, pow(10.0, lo + r.dlog_r / 2));
        if (T4 > 0) {
            const double Tann = pow(T4, 0.25);
            const double kT = RGK_KB * (Tann * (r.fcol < 0 ? rg_fcol(Tann) : r.fcol));
            if (kT > best) { best = kT; bestg = g; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(FULL, best, o);
        const int og = __shfl_xor_sync(FULL, bestg, o);
        if (ob > best || (ob == best && og > bestg)) { best = ob; bestg = og; }
    }
    int k_ref = 0;
    if (bestg >= 0 && rel) {
        const double2 eg = __ldg(edges + r.ann_base + bestg);
        if (eg.x <= 3 && eg.y <= 3) { // the reference annulus goes through the transfer: its strongest tap
            const double *row = hist + (size_t)(r.ann_base + bestg) * hstride;
            const int2 hd = __ldg(reinterpret_cast<const int2 *>(row));
            double hv = -1.0;
            int hk = 0;
            for (int k = lane; k < hd.y; k += 32) {
                const double v = __ldg(row + 1 + k);
                if (v > hv) { hv = v; hk = hd.x + k; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(FULL, hv, o);
                const int ok = __shfl_xor_sync(FULL, hk, o);
                if (ov > hv || (ov == hv && ok < hk)) { hv = ov; hk = ok; }
            }
            k_ref = hk;
        }
    }
    if (lane == 0) wien[set] = make_double2(best, (double)k_ref);
}

int rg_launch_wienref(const SetRec *recs, int P, const double2 *edges, const double *hist, int hstride, int rel,
                      double2 *wien, cudaStream_t st)
{
    if (P <= 0) return RG_OK;
    RG_LAUNCH(wienref_kernel, dim3((unsigned)(((size_t)P * 32 + 255) / 256)), dim3(256), 0, st, recs, P, edges, hist,
              hstride, rel, wien);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

// Fold partial spectra: out[set][c] = sum_q partial[set][q][c] in fixed order q = 0..nparts-1
// (component mode: q runs over blocks x warps and total = disc + warm + hot; total-only mode: q runs over blocks)
__global__ void __launch_bounds__(256) reduce_parts_kernel(const double *__restrict__ partial, double *__restrict__ out,
                                                           int P, int nparts, int nE, int comps)
{
    const int npc = comps ? 3 : 1;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)P * nE) return;
    int set = (int)(idx / nE), j = (int)(idx - (size_t)set * nE);
    double tot = 0;
    for (int c = 0; c < npc; c++) {
        double s = 0;
        for (int q = 0; q < nparts; q++) s += partial[(((size_t)set * nparts + q) * npc + c) * nE + j];
        if (comps) out[((size_t)set * 4 + c) * nE + j] = s;
        tot += s;
    }
    out[((size_t)set * (comps ? 4 : 1) + (comps ? 3 : 0)) * nE + j] = tot;
}

template <int NC, typename R>
static int launch_spec(const SpecArgs &args, int sm_count, cudaStream_t st)
{
    size_t smem = rg_spectrum_smem_bytes(NC, args.nH, args.HL, args.HH, args.warps, sizeof(R) == 4);
    auto kern = spectrum_kernel<NC, R>;
    RG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long items = (long)args.P * args.NB;
    const long per_sm = (NC <= 4 && sizeof(R) == 8) ? RG_SPEC_CPS : 1;
    long ctas = items < sm_count * per_sm ? items : sm_count * per_sm; // persistent CTAs pull items from the queue
    RG_CUDA_OK(cudaMemsetAsync(args.queue, 0, sizeof(int), st));
    RG_LAUNCH(kern, dim3((unsigned)ctas), dim3(32 * args.warps), smem, st, args);
    RG_CUDA_OK(cudaGetLastError());
    const bool comps = (args.flags & RG_FLAG_COMPONENTS) != 0;
    if (comps || args.NB > 1) {
        size_t n = (size_t)args.P * args.grid.nE;
        RG_LAUNCH(reduce_parts_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, args.partial, args.out, args.P,
                  comps ? args.NB * args.warps : args.NB, args.grid.nE, comps ? 1 : 0);
        RG_CUDA_OK(cudaGetLastError());
    }
    return RG_OK;
}

int rg_launch_spectrum(const SpecArgs &args, int sm_count, cudaStream_t st)
{
    int NC = rg_spectrum_pick_NC(args.grid.nE);
    if (args.warps < 1) return rg_set_error(RG_E_GRID, "energy grid too fine for the shared-memory working set");
    const bool fp32 = (args.flags & RG_FLAG_FP32) != 0;
    if (NC == 4) return fp32 ? launch_spec<4, float>(args, sm_count, st) : launch_spec<4, double>(args, sm_count, st);
    if (NC == 8) return fp32 ? launch_spec<8, float>(args, sm_count, st) : launch_spec<8, double>(args, sm_count, st);
    if (NC == 20) return fp32 ? launch_spec<20, float>(args, sm_count, st) : launch_spec<20, double>(args, sm_count, st);
    return rg_set_error(RG_E_GRID, "nE = %d exceeds the kernel capacity (%d)", args.grid.nE, 20 * 256);
}
