// rg_hist.cu -- the shift histograms of the relativistic transfer: one CTA per parameter set, one WARP per annulus.
//
// The kyconv seam of the reference (relagn.py:979-1001) turns every annulus [r1, r2] into a transfer kernel on the
// log-energy grid; rg_transfer.cuh builds it as a histogram over integer bin shifts.  Per parameter set:
//   phase 1  the table slice of the set's (spin, inclination) -- a Lagrange blend of up to 4 x 4 table nodes, read
//            from L2 with coalesced 8-byte loads -- is built ONCE into shared memory as (ln g, Jn g^3) per (row, column);
//   phase 2  the warps walk the set's annuli; a ring is two shared-memory rows interpolated linearly, its histogram
//            a gather over integer shifts (warp_ring_gather: prefix sums over the two monotone branches of the ring,
//            one lane per shift, no atomics, fixed summation order).
// The consumer (spectrum_kernel: Planck function + banded convolution) is register heavy and runs at 12 warps per
// SM; the histograms (W ~ 20 doubles per annulus) travel through a global arena that stays in L2 / HBM for a few
// hundred microseconds: [annulus][hstride] doubles, slot 0 = (first shift, count) as two ints.
#include "rg_kernels.h"
#include "rg_transfer.cuh"

#ifndef HIST_WARPS
#define HIST_WARPS 16
#endif
#ifndef RG_HIST_MINB
#define RG_HIST_MINB 1 // resident CTAs per SM the register budget is cut for
#endif

static __host__ __device__ inline int hist_round_up2(int x) { return (x + 1) & ~1; }

// one annulus as lane m prepares it for the warp
struct PrepRec { double r1, r2; RingGeo geo; int flag, g, nsub, pad; };
#define HIST_PREP 16 // annuli a warp prepares per pass

// shared memory: [slice: NR NPHI double2] [per warp: H slot | ring scratch | prepared annuli]
static __host__ __device__ inline size_t hist_warp_bytes(int nH, int NP)
{
    return (size_t)hist_round_up2(nH + 24) * sizeof(double) + (size_t)hist_round_up2(RG_RING_SCR_DOUBLES(NP)) * sizeof(double) +
           HIST_PREP * sizeof(PrepRec);
}

__global__ void __launch_bounds__(32 * HIST_WARPS, RG_HIST_MINB) hist_kernel(HistArgs A)
{
    RG_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NP = A.tab.NPHI;
    const int HSZ = hist_round_up2(A.nH + 24);
    double2 *slice = reinterpret_cast<double2 *>(smem_raw);
    unsigned char *wbase = smem_raw + (size_t)A.tab.NR * NP * sizeof(double2) + (size_t)warp * hist_warp_bytes(A.nH, NP);
    double *Hslot = reinterpret_cast<double *>(wbase); // 8 zero pad | nH bins | zero pad
    const RingScr scr = rg_ring_scr(Hslot + HSZ, NP);
    PrepRec *prep = reinterpret_cast<PrepRec *>(Hslot + HSZ + hist_round_up2(RG_RING_SCR_DOUBLES(NP)));
    for (int i = lane; i < HSZ; i += 32) Hslot[i] = 0.0;
    const int set = blockIdx.x / A.NBH, blk = blockIdx.x - set * A.NBH;
    const SetRec &r = A.recs[set];
    if (r.status != RG_S_OK) return;
    const int n_tot = r.n_disc + r.n_warm + r.n_hot;
    const int g_warm0 = r.n_disc, g_hot0 = r.n_disc + r.n_warm;
    // ---- phase 1: the set's table slice
    __shared__ TabSel s_ts;
    __shared__ int s_hi;
    if (threadIdx.x == 0) {
        rg_tab_select(s_ts, A.tab, r.a, r.inc);
        const double rtop = r.r_out < RG_TAB_ROUT ? r.r_out : RG_TAB_ROUT;
        int hi = (int)floor(rg_row_pos(A.tab.x0, A.tab.dx, A.tab.NR, rtop)) + 1;
        if (hi > A.tab.NR - 1) hi = A.tab.NR - 1;
        if (hi < s_ts.s_lo + 1) hi = s_ts.s_lo + 1;
        s_hi = hi;
    }
    __syncthreads();
    const TabSel &ts = s_ts;
    rg_build_slice(ts, s_hi, slice, threadIdx.x, blockDim.x);
    __syncthreads();
    // ---- phase 2: annulus g -> warp (blk HIST_WARPS + warp) mod (HIST_WARPS NBH)
    const int wstride = HIST_WARPS * A.NBH, w0 = blk * HIST_WARPS + warp;
    const int n_mine = w0 < n_tot ? (n_tot - w0 + wstride - 1) / wstride : 0;
    for (int m0 = 0; m0 < n_mine; m0 += HIST_PREP) {
        // lane m prepares this warp's annulus number m0 + m: its edges (relagn.py:731-776, :962-964) and whether the
        // reference sends it through kyconv (:992)
        __syncwarp(); // the previous pass has been read
        if (lane < HIST_PREP && m0 + lane < n_mine) {
            PrepRec pr;
            pr.r1 = 0; pr.r2 = 0; pr.flag = 0; pr.g = 0; pr.nsub = 1; pr.pad = 0;
            pr.geo.A = 0; pr.geo.fr = 0; pr.geo.s0 = ts.s_lo; // of the annulus' only ring when nsub == 1 (most annuli)
            const int g = w0 + wstride * (m0 + lane);
            const double2 eg = __ldg(A.edges + r.ann_base + g); // (lo, hi) in log10 r (edges_kernel)
            const double lo = eg.x, hi = eg.y;
            pr.r1 = pow(10.0, lo); pr.r2 = pow(10.0, hi);
            pr.flag = (lo <= 3 && hi <= 3) ? 1 : 0;
            pr.g = g;
            if (pr.flag) {
                const double lnr = log(pr.r2 / pr.r1);
                pr.nsub = rg_auto_nsub(pr.r1, pr.r2, lnr, A.dlnE);
                if (pr.nsub == 1) pr.geo = rg_ring_geo(ts, pr.r1, pr.r2, lnr, 0, 1, 0.0, 0.0, 1.0);
            }
            prep[lane] = pr;
        }
        __syncwarp();
        const int mt = n_mine - m0 < HIST_PREP ? n_mine - m0 : HIST_PREP;
        for (int mi = 0; mi < mt; mi++) {
            const double r1 = prep[mi].r1, r2 = prep[mi].r2;
            const int flag = prep[mi].flag, g = prep[mi].g, nsub = prep[mi].nsub;
            const RingGeo geo = prep[mi].geo;
            double *row = A.hist + (size_t)(r.ann_base + g) * A.hstride;
            int kmn = 0, kmx = -1;
            if (flag && ts.status == 0) {
                warp_build_rings(Hslot + 8, A.K_LO, A.nH, ts, slice, scr, r1, r2, A.inv_dlnE, nsub, 0, 1, 0.0, 0.0, 1.0, 0.0,
                                 1.0, kmn, kmx, nsub == 1 ? &geo : nullptr);
                if (kmn < A.K_LO) kmn = A.K_LO;
                if (kmx > A.K_LO + A.nH - 1) kmx = A.K_LO + A.nH - 1;
                __syncwarp();
            }
            const int nk = kmx >= kmn ? kmx - kmn + 1 : 0;
            if (lane == 0) *reinterpret_cast<int2 *>(row) = make_int2(kmn, nk);
            for (int k = lane; k < nk; k += 32) {
                const int q = 8 + kmn - A.K_LO + k;
                row[1 + k] = Hslot[q];
                Hslot[q] = 0.0; // leave the slot clean for the next annulus
            }
            __syncwarp();
        }
    }
    (void)g_warm0; (void)g_hot0;
}

int rg_launch_hist(const HistArgs &A, cudaStream_t st)
{
    if (A.P == 0) return RG_OK;
    size_t smem = (size_t)A.tab.NR * A.tab.NPHI * sizeof(double2) + (size_t)HIST_WARPS * hist_warp_bytes(A.nH, A.tab.NPHI);
    if (smem > 227 * 1024 - 1024)
        return rg_set_error(RG_E_GRID, "shift histograms: table slice (%d x %d) and %d shift bins exceed the shared memory",
                            A.tab.NR, A.tab.NPHI, A.nH);
    RG_CUDA_OK(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RG_LAUNCH(hist_kernel, dim3((unsigned)((long)A.P * A.NBH)), dim3(32 * HIST_WARPS), smem, st, A);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}
