// rg_hist.cu -- the shift histograms of the relativistic transfer, one WARP per annulus.
//
// The kyconv seam of the reference (relagn.py:979-1001) turns every annulus [r1, r2] into a transfer kernel on the
// log-energy grid; rg_transfer.cuh builds it as a histogram over integer bin shifts from the Kerr table slice of the
// set's (spin, inclination).  Building it is a chain of dependent warp shuffles and shared-memory updates -- latency
// bound, few registers -- while the consumer (spectrum_kernel: Planck function + banded convolution) is register
// heavy and runs at 12 warps per SM.  Doing the build here, in its own kernel at 3x the occupancy, hides those
// latencies behind other warps; the histograms (W ~ 20 doubles per annulus) travel through a global arena that stays
// in L2 / HBM for a few hundred microseconds: [annulus][hstride] doubles, slot 0 = (first shift, count) as two ints.
#include "rg_kernels.h"
#include "rg_transfer.cuh"

#ifndef HIST_WARPS
#define HIST_WARPS 8
#endif
#ifndef RG_HIST_MINB
#define RG_HIST_MINB 4 // resident CTAs per SM the register budget is cut for
#endif

static __host__ __device__ inline int hist_round_up2(int x) { return (x + 1) & ~1; }

// one annulus as lane m prepares it for the warp
struct PrepRec { double r1, r2; RingGeo geo; int flag, g, nsub, pad; };

__global__ void __launch_bounds__(32 * HIST_WARPS, RG_HIST_MINB) hist_kernel(HistArgs A)
{
    RG_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int HSZ = hist_round_up2(A.nH + 24);
    double *Hslot = reinterpret_cast<double *>(smem_raw) + (size_t)warp * HSZ; // 8 zero pad | nH bins | zero pad
    for (int i = lane; i < HSZ; i += 32) Hslot[i] = 0.0;
    PrepRec *prep = reinterpret_cast<PrepRec *>(reinterpret_cast<double *>(smem_raw) + (size_t)HIST_WARPS * HSZ) + 32 * warp;
    __syncwarp();
    const int set = blockIdx.x / A.NBH, blk = blockIdx.x - set * A.NBH;
    const SetRec &r = A.recs[set];
    if (r.status != RG_S_OK) return;
    const int n_tot = r.n_disc + r.n_warm + r.n_hot;
    const int g_warm0 = r.n_disc, g_hot0 = r.n_disc + r.n_warm;
    // the table selection of the set (4 node planes, blend weights, row geometry: 26 registers' worth, the same for
    // the whole CTA) lives in shared memory: at the 64-register budget of 32 warps per SM the compiler otherwise
    // spills and rematerialises (r1q: 440 bytes of spill loads, S2R / address rebuilds = 10 % of the instructions)
    __shared__ TabSel s_ts;
    if (threadIdx.x == 0) rg_tab_select(s_ts, A.tab, r.a, r.inc);
    __syncthreads();
    const TabSel &ts = s_ts;
    const int wstride = HIST_WARPS * A.NBH, w0 = blk * HIST_WARPS + warp;
    const int n_mine = w0 < n_tot ? (n_tot - w0 + wstride - 1) / wstride : 0;
    for (int m0 = 0; m0 < n_mine; m0 += 32) {
        // lane m prepares this warp's annulus number m0 + m: its edges (relagn.py:731-776, :962-964) and whether the
        // reference sends it through kyconv (:992)
        // (the prepared values go through the warp's shared-memory table, not through registers + shuffles: 13
        //  registers less across the ring loop)
        __syncwarp(); // the previous pass has been read
        if (m0 + lane < n_mine) {
            PrepRec pr;
            pr.r1 = 0; pr.r2 = 0; pr.flag = 0; pr.g = 0; pr.nsub = 1;
            pr.geo.A = 0; pr.geo.fr = 0; pr.geo.s0 = 0; pr.geo.s1 = 0; // of the annulus' only ring when nsub == 1 (most annuli)
            const int g = w0 + wstride * (m0 + lane);
            const int region = g < g_warm0 ? 0 : (g < g_hot0 ? 1 : 2);
            const int a = g - (region == 0 ? 0 : (region == 1 ? g_warm0 : g_hot0));
            (void)a;
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
        const int mt = n_mine - m0 < 32 ? n_mine - m0 : 32;
        for (int mi = 0; mi < mt; mi++) {
            const double r1 = prep[mi].r1, r2 = prep[mi].r2;
            const int flag = prep[mi].flag, g = prep[mi].g, nsub = prep[mi].nsub;
            const RingGeo geo = prep[mi].geo;
            double *row = A.hist + (size_t)(r.ann_base + g) * A.hstride;
            int kmn = 0, kmx = -1;
            if (flag) {
                warp_build_rings(Hslot + 8, A.K_LO, A.nH, ts, r1, r2, A.dlnE, nsub, 0, 1, 0.0, 0.0, 1.0, 0.0, 1.0, kmn, kmx,
                                 nsub == 1 ? &geo : nullptr, A.inv_dlnE);
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
}

int rg_launch_hist(const HistArgs &A, cudaStream_t st)
{
    if (A.P == 0) return RG_OK;
    size_t smem = (size_t)HIST_WARPS * (hist_round_up2(A.nH + 24) * sizeof(double) + 32 * sizeof(PrepRec));
    RG_CUDA_OK(cudaFuncSetAttribute(hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RG_LAUNCH(hist_kernel, dim3((unsigned)((long)A.P * A.NBH)), dim3(32 * HIST_WARPS), smem, st, A);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}
