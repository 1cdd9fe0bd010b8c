// rg_hist.cu -- the shift histograms of the relativistic transfer: one CTA per parameter set, one WARP per annulus.
//
// The kyconv seam of the reference (relagn.py:979-1001) turns every annulus [r1, r2] into a transfer kernel on the
// log-energy grid; rg_transfer.cuh builds it as a histogram over integer bin shifts.  Per parameter set:
//   phase 1  the table slice of the set's (spin, inclination) -- a Lagrange blend of up to 4 x 4 table nodes, read
//            from L2 with coalesced 8-byte loads -- is built ONCE per set as (ln g, Jn g^3) per (row, column) into the
//            CTA's L2-resident scratch region;
//   phase 2  the warps walk the set's annuli; a ring is two slice rows interpolated linearly, its histogram a gather
//            over integer shifts (warp_ring_gather: prefix sums over the two monotone branches of the ring, no
//            atomics, fixed summation order).
// The consumer (spectrum_kernel: Planck function + banded convolution) is register heavy and runs at 12 warps per
// SM; the histograms (W ~ 20 doubles per annulus) travel through a global arena that stays in L2 / HBM for a few
// hundred microseconds: [annulus][hstride] doubles, slot 0 = (first shift, count) as two ints.
#include "rg_kernels.h"
#include "rg_transfer.cuh"

// CTA shape: warps per CTA x resident CTAs per SM the register budget is cut for.  Swept in r2zl-r2zn (ms per 125 000
// sets): 8 x 3 (80 registers, 84 bytes of spills) 23.4; 6 x 4 (80, spills) 23.0; 6 x 3 22.6; 10 x 2 22.5; 4 x 4 (127
// registers) 22.1; 5 x 4 (96 registers, no spills, 20 warps per SM) 21.1 -- the ring routine wants ~96 registers more
// than it wants the last four warps
#ifndef HIST_WARPS
#define HIST_WARPS 5
#endif
#ifndef RG_HIST_MINB
#define RG_HIST_MINB 4
#endif

static __host__ __device__ inline int hist_round_up2(int x) { return (x + 1) & ~1; }

// one annulus as lane m prepares it for the warp
struct PrepRec { double r1, r2; RingGeo geo; int flag, g, nsub, pad; };
#define HIST_PREP 8 // annuli a warp prepares per pass

// shared memory per warp, fixed-size pieces first so that their addresses are the warp's base plus a constant:
//   prepared annuli | ring samples (NP double2) | DA, DB (DN doubles each) | H slot (HSZ doubles)
static __host__ __device__ inline int hist_dn(int HSZ) { return HSZ < 512 ? HSZ : 512; } // entries of the D window
static __host__ __device__ inline size_t hist_warp_bytes(int nH, int NP)
{
    const int HSZ = hist_round_up2(nH + 24);
    return HIST_PREP * sizeof(PrepRec) + (size_t)RG_RING_SCR_DOUBLES(NP, hist_dn(HSZ)) * sizeof(double) + (size_t)HSZ * sizeof(double);
}

// The table selection of every set of the chunk (stencil nodes, Lagrange weights, slice rows): one thread per set
// instead of one thread of each hist_kernel CTA while the other 255 wait (24 FP64 divisions and rg_risco in series).
__global__ void __launch_bounds__(128) tabsel_kernel(const SetRec *__restrict__ recs, int P, TabDesc tab, SelRec *__restrict__ sel)
{
    const int set = blockIdx.x * blockDim.x + threadIdx.x;
    if (set >= P) return;
    const SetRec &r = recs[set];
    if (r.status != RG_S_OK) return;
    SelRec o;
    rg_tab_select(o.ts, tab, r.a, r.inc);
    const double rtop = r.r_out < RG_TAB_ROUT ? r.r_out : RG_TAB_ROUT;
    int hi = (int)floor(rg_row_pos(tab.x0, tab.dx, tab.NR, rtop)) + 1;
    if (hi > tab.NR - 1) hi = tab.NR - 1;
    if (hi < o.ts.s_lo + 1) hi = o.ts.s_lo + 1;
    o.s_hi = hi;
    o.pad = 0;
    sel[set] = o;
}

// Persistent CTAs pull (parameter set, annulus block) items from a device queue.  The set's table slice is built by
// the whole CTA into the CTA's own scratch region of a global arena (2 x 64 kB per resident CTA: it never leaves
// L2) -- in shared memory it would cap the kernel at 16 warps per SM, and the ring routine is a chain of dependent
// shared-memory and shuffle steps that needs the warps to hide its latency.
// Software pipeline, ONE block barrier per item: in iteration k every warp walks its annuli of item k (slice half
// k & 1) and then, without waiting for the other warps, builds its share of the slice of item k + 1 into the other
// half; warp 0 meanwhile fetches item k + 2 from the queue and loads its table selection (three slots).
template <int NQ, int MINB>
__global__ void __launch_bounds__(32 * HIST_WARPS, MINB) hist_kernel(HistArgs A)
{
    RG_DYN_SMEM(smem_raw);
    constexpr int NP = 32 * NQ;
    int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // (the lane index and the warp's offset are used all over the ring routine: opaque to the compiler from here on,
    //  so that it keeps them in registers instead of rebuilding them from the thread index at every use -- 7 % of the
    //  kernel's instructions in profiles/r2c.  The OFFSET, not the pointer: the accesses must stay LDS / STS)
    unsigned woff = (unsigned)warp * (unsigned)A.warp_bytes;
    asm volatile("" : "+r"(lane));
    asm volatile("" : "+r"(woff));
    unsigned char *wbase = smem_raw + woff;
    PrepRec *prep = reinterpret_cast<PrepRec *>(wbase);
    double *scr_base = reinterpret_cast<double *>(wbase + HIST_PREP * sizeof(PrepRec));
    const RingScr scr = rg_ring_scr(scr_base, NP, A.DN);
    double *Hslot = scr_base + RG_RING_SCR_DOUBLES(NP, A.DN); // 8 zero pad | nH bins | zero pad
    for (int i = lane; i < A.HSZ; i += 32) Hslot[i] = 0.0;
    double2 *slice2 = A.slices + (size_t)2 * blockIdx.x * A.tab.NR * NP; // two halves
    const size_t slice_elems = (size_t)A.tab.NR * NP;
    __shared__ SelRec s_sel[3];
    __shared__ int s_items[3];
    const int n_items = A.P * A.NBH;
    // warp 0: next item from the queue and its table selection into slot `slot`
    auto fetch = [&](int slot) {
        int id = 0;
        if (lane == 0) id = atomicAdd(A.queue, 1);
        id = __shfl_sync(RG_FULL, id, 0);
        if (lane == 0) s_items[slot] = id;
        if (id < n_items) {
            const int set = id / A.NBH;
            if (A.recs[set].status == RG_S_OK) {
                const double *src = reinterpret_cast<const double *>(A.sel + set);
                double *dst = reinterpret_cast<double *>(&s_sel[slot]);
                for (int i = lane; i < (int)(sizeof(SelRec) / 8); i += 32) dst[i] = src[i];
            }
        }
    };
    auto build = [&](int slot, int half) {
        const int id = s_items[slot];
        if (id >= n_items) return;
        if (A.recs[id / A.NBH].status != RG_S_OK) return;
        rg_build_slice(s_sel[slot].ts, s_sel[slot].s_hi, slice2 + half * slice_elems, threadIdx.x, blockDim.x);
    };
    if (warp == 0) { fetch(0); fetch(1); }
    __syncthreads();
    build(0, 0);
    for (int k = 0;; k++) {
        __syncthreads(); // half k & 1 is complete, the slots of items k and k + 1 are filled, iteration k - 1 is over
        const int item = s_items[k % 3];
        if (item >= n_items) break;
        if (warp == 0) fetch((k + 2) % 3);
        const int set = item / A.NBH, blk = item - set * A.NBH;
        const SetRec &r = A.recs[set];
        if (r.status == RG_S_OK) {
            const TabSel &ts = s_sel[k % 3].ts;
            const double2 *slice = slice2 + (k & 1) * slice_elems;
            const int n_tot = r.n_disc + r.n_warm + r.n_hot;
            // annulus g -> warp (blk HIST_WARPS + warp) mod (HIST_WARPS NBH)
            const int wstride = HIST_WARPS * A.NBH, w0 = blk * HIST_WARPS + warp;
            const int n_mine = w0 < n_tot ? (n_tot - w0 + wstride - 1) / wstride : 0;
            for (int m0 = 0; m0 < n_mine; m0 += HIST_PREP) {
                // lane m prepares this warp's annulus number m0 + m: its edges (relagn.py:731-776, :962-964) and
                // whether the reference sends it through kyconv (:992)
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
                        warp_build_rings<NQ>(Hslot + 8, A.K_LO, A.nH, ts, slice, scr, r1, r2, A.inv_dlnE, nsub, 0, 1, 0.0,
                                             0.0, 1.0, 0.0, 1.0, kmn, kmx, nsub == 1 ? &geo : nullptr);
                        if (kmn < A.K_LO) kmn = A.K_LO;
                        if (kmx > A.K_LO + A.nH - 1) kmx = A.K_LO + A.nH - 1;
                        __syncwarp();
                    }
                    const int nk = kmx >= kmn ? kmx - kmn + 1 : 0;
                    if (lane == 0) *reinterpret_cast<int2 *>(row) = make_int2(kmn, nk);
                    for (int q = lane; q < nk; q += 32) {
                        const int i = 8 + kmn - A.K_LO + q;
                        row[1 + q] = Hslot[i];
                        Hslot[i] = 0.0; // leave the slot clean for the next annulus
                    }
                    __syncwarp();
                }
            }
        }
        build((k + 1) % 3, (k + 1) & 1);
    }
}

template <int NQ, int MINB>
static int launch_hist(const HistArgs &A_in, int sm_count, cudaStream_t st)
{
    HistArgs A = A_in;
    A.HSZ = hist_round_up2(A.nH + 24);
    A.DN = hist_dn(A.HSZ);
    A.warp_bytes = (int)hist_warp_bytes(A.nH, 32 * NQ);
    const size_t smem = (size_t)HIST_WARPS * A.warp_bytes;
    if (smem > 227 * 1024 - 1024)
        return rg_set_error(RG_E_GRID, "shift histograms: %d shift bins exceed the shared memory", A.nH);
    RG_LAUNCH(tabsel_kernel, dim3((A.P + 127) / 128), dim3(128), 0, st, A.recs, A.P, A.tab, A.sel);
    RG_CUDA_OK(cudaGetLastError());
    auto kern = hist_kernel<NQ, MINB>;
    RG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    RG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * HIST_WARPS, smem));
    if (per_sm < 1) per_sm = 1;
    if (const char *e = getenv("RG_HIST_RES")) { // tuning knob: fewer resident CTAs per SM (room for the Kompaneets solve)
        const int k = atoi(e);
        if (k >= 1 && k < per_sm) per_sm = k;
    }
    long ctas = (long)A.P * A.NBH;
    if (ctas > (long)sm_count * per_sm) ctas = (long)sm_count * per_sm;
    if (ctas > A.slice_slots / 2) ctas = A.slice_slots / 2; // two slice halves per CTA
    RG_CUDA_OK(cudaMemsetAsync(A.queue, 0, sizeof(int), st));
    RG_LAUNCH(kern, dim3((unsigned)ctas), dim3(32 * HIST_WARPS), smem, st, A);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

size_t rg_selrec_bytes() { return sizeof(SelRec); }

int rg_launch_hist(const HistArgs &A, int sm_count, cudaStream_t st)
{
    if (A.P == 0) return RG_OK;
    // resident CTAs per SM the register budget is cut for: 4 (64 registers) or 3 (80); RG_HIST_CTAS: tuning knob
    switch (A.tab.NPHI) {
    case 32: return launch_hist<1, 4>(A, sm_count, st);
    case 64: return launch_hist<2, RG_HIST_MINB>(A, sm_count, st);
    case 128: return launch_hist<4, 3>(A, sm_count, st);
    case 256: return launch_hist<8, 2>(A, sm_count, st);
    }
    return rg_set_error(RG_E_ARG, "NPHI must be 32, 64, 128 or 256");
}
