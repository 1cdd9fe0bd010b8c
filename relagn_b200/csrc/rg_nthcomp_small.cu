// rg_nthcomp_small.cu -- pyNTHCOMP (donthcomp / _thcompton / _thermlc, pyNTHCOMP.py:11-251) for SMALL batches: one WARP
// per Kompaneets system, the tridiagonal system in shared memory.
//
//   P1  coefficients of the system (pyNTHCOMP.py:120-165, 210-221), one row per lane                -> shared memory
//   P2  Thomas forward sweep and back substitution (pyNTHCOMP.py:226-247) by lane 0: the reference's own operations in
//       the reference's order
//   P3  sptot_j = x^4 u_j bet_j tau (pyNTHCOMP.py:170-171, 249), normalisation at 1 keV (:49-56), lanes over rows
//   P4  donthcomp's interpolation onto the caller's grid (pyNTHCOMP.py:57-76) and * h / keV, lanes over energy bins
//
// This is the LATENCY variant.  The throughput kernels (rg_setup.cu: one thread per system + an interpolation kernel)
// take 0.39 ms for the 21 systems of a single model object -- 395 rows of a serial recurrence with ~200 instructions per
// row in one thread; here the rows' coefficients are computed 32 at a time, the serial sweep is 395 x (FMA ->
// reciprocal -> multiply) on shared memory, and the 21 systems run on 21 warps at once.  For large batches it is the
// wrong kernel: the one-lane sweep costs a full warp instruction per operation (139 ms against 33 ms per 125 000 sets,
// profiles/r2e).  Used by the small-batch path of rg_eval_batch_host (P <= 64).
#include "rg_kernels.h"

#ifndef NTH_WARPS
#define NTH_WARPS 7 // warps per CTA; two CTAs per SM: 14 x 16 kB of rows in shared memory
#endif
#define NTH_RCAP 512 // rows a warp's shared-memory arrays hold; longer grids (kTe / kT_seed > ~1e7) use global scratch

struct NthSys { // per-system scalars of _thcompton
    double tempbb, theta, tautom, xmin, c20, tod, todx, aa, itau, rxr, itbb, xnr, planck, x32;
    int jmax, jmaxth, jnr, jrel;
};

// bet[j] of pyNTHCOMP.py:124-165 at x = x[j] (j >= 1)
__device__ __forceinline__ double nth_bet(const NthSys &s, int j, double w)
{
    if (j >= s.jrel) return s.itau;
    double rel;
    if (w <= 0.05) rel = (1.0 - 2.0 * w + 26.0 * w * w * 0.2);
    else {
        // Klein-Nishina factor (pyNTHCOMP.py:128-136) with the six divisions by w and 1 + 2w folded into two
        // reciprocals (last-bit differences only)
        const double rw = 1.0 / w, z2 = 1.0 + 2.0 * w, rz = 1.0 / z2, z3 = log(z2);
        double z1 = (1.0 + w) * (rw * rw * rw), z4 = 2.0 * w * (1.0 + w) * rz;
        double z5 = z3 * 0.5 * rw, z6 = (1.0 + 3.0 * w) * (rz * rz);
        rel = (0.75 * (z1 * (z4 - z3) + z5 - z6));
    }
    const double taukn = s.tautom * rel;
    if (j < s.jnr - 1) return s.itau * __drcp_rn(1.0 + taukn * (1.0 / 3.0));
    const double flz = 1 - (w - s.xnr) * s.rxr;
    return s.itau * __drcp_rn(1.0 + taukn * (1.0 / 3.0) * flz);
}

__global__ void __launch_bounds__(32 * NTH_WARPS, 2)
nthcomp_small_kernel(const SetRec *__restrict__ recs, const int2 *__restrict__ syslist, int nsys, int n_warm_sys, int n_warm_pad,
               int hot_first, const double *__restrict__ p10tab, const double2 *__restrict__ edges,
               double *__restrict__ gscratch, GridDesc grid, double *__restrict__ farr, int *__restrict__ queue,
               unsigned long long *stats)
{
    RG_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *rows_sm = reinterpret_cast<double *>(smem_raw) + (size_t)warp * 4 * NTH_RCAP;
    double *rows_gl = gscratch + ((size_t)blockIdx.x * NTH_WARPS + warp) * 4 * (RG_NTH + 4);
    const double *P10 = p10tab, *W1T = p10tab + RG_P10_W1, *RW1T = p10tab + RG_P10_RW1, *RDP = p10tab + RG_P10_RDP;
    const int nE = grid.nE;
    unsigned long long sumj = 0, nsolved = 0;
    for (;;) {
        int li = 0;
        if (lane == 0) li = atomicAdd(queue, 1);
        li = __shfl_sync(0xffffffffu, li, 0);
        if (li >= nsys) break;
        if (li >= n_warm_sys && li < n_warm_pad) continue;
        // logical index -> arena row: the warm systems fill rows [0, n_warm_sys), the hot shapes the top of the arena
        const int sys = li < n_warm_pad ? li : hot_first + (li - n_warm_pad);
        const int2 ent = syslist[sys];
        if (ent.x < 0) continue; // a row nobody reserved (the small-batch path scans the arena up to a cap)
        const SetRec &r = recs[ent.x];
        if (r.status != RG_S_OK) continue;
        // ---- parameters of the system
        double gamma, kte, ktbb;
        if (ent.y < 0) { // hot shape: relagn.py:1200-1223
            gamma = r.gamma_h; kte = r.kTe_h; ktbb = r.kTseed;
        } else {         // warm annulus ent.y (counted top-down): relagn.py:900-927
            NtConst nt;
            rg_nt_prepare(nt, r);
            const double lo = __ldg(edges + r.ann_base + r.n_disc + ent.y).x;
            const double rmid = pow(10.0, lo + r.dlog_r / 2);
            const double Tann = pow(rg_tnt4(nt, rmid), 0.25);
            gamma = r.gamma_w; kte = r.kTe_w; ktbb = Tann / 1.16048e7;
        }
        // ---- _thcompton (pyNTHCOMP.py:80-173): grid and scalars
        NthSys s;
        s.tempbb = ktbb / 511.0; s.theta = kte / 511.0;
        s.tautom = sqrt(2.250 + 3.0 / (s.theta * ((gamma + .50) * (gamma + .50) - 2.250))) - 1.50;
        const double delta = 0.02;
        const double deltal = delta * log(10.0);
        s.xmin = 1e-4 * s.tempbb;
        const double xmax = 40.0 * s.theta;
        int jmax = (int)(log10(xmax / s.xmin) / delta) + 1;
        if (jmax > 899) jmax = 899;
        double *F = farr + (size_t)sys * nE;
        if (jmax < 4) { // degenerate grid: the reference would index out of range; the spectrum is empty
            for (int i = lane; i < nE; i += 32) F[i] = 0.0;
            continue;
        }
        s.jmax = jmax;
        s.jmaxth = (int)(log10(50 * s.tempbb / s.xmin) / delta);
        if (s.jmaxth > 900) s.jmaxth = 900;
        if (s.jmaxth > jmax) s.jmaxth = jmax;
        const double pt = RGK_PI * s.tempbb;
        s.planck = 15.0 / (pt * pt * pt * pt);
        s.jnr = (int)(log10(0.10 / s.xmin) / delta + 1);
        if (s.jnr > jmax - 1) s.jnr = jmax - 1;
        s.jrel = (int)(log10(1 / s.xmin) / delta + 1);
        if (s.jrel > jmax) s.jrel = jmax;
        s.xnr = s.xmin * P10[s.jnr - 1 > 0 ? s.jnr - 1 : 0];
        const double xr = s.xmin * P10[s.jrel - 1 > 0 ? s.jrel - 1 : 0];
        s.c20 = s.tautom / deltal;
        s.tod = s.theta / deltal;
        s.x32 = sqrt((s.xmin * P10[0]) * (s.xmin * P10[1]));
        s.aa = (s.tod / s.x32 + 0.5) / (s.tod / s.x32 - 0.5);
        s.itau = 1.0 / s.tautom; s.rxr = 1.0 / (xr - s.xnr); s.itbb = 1.0 / s.tempbb; s.todx = s.tod / s.xmin;
        // rows: shared memory, or the warp's global scratch for the rare long grids
        double *RA, *RB, *RC, *RD;
        if (jmax + 1 <= NTH_RCAP) { RA = rows_sm; RB = RA + NTH_RCAP; RC = RB + NTH_RCAP; RD = RC + NTH_RCAP; }
        else { RA = rows_gl; RB = RA + RG_NTH + 4; RC = RB + RG_NTH + 4; RD = RC + RG_NTH + 4; }

        // ---- P1: coefficients a_j u_{j+1} + b_j u_j + c_j u_{j-1} = d_j, j = 1 .. jmax - 2 (pyNTHCOMP.py:210-221)
        for (int j = 1 + lane; j < jmax - 1; j += 32) {
            const double x0 = s.xmin * P10[j];
            // w1 = sqrt(x[j] x[j+1]) = xmin 10^(0.02 j + 0.01) and theta / deltal / w1 from the resident tables: no
            // square root and one division less per row (last-bit differences against pyNTHCOMP.py:117,211)
            const double w1j = s.xmin * W1T[j];
            const double c2j = (w1j * w1j * w1j * w1j) * __drcp_rn(1.0 + 4.60 * w1j + 1.1 * w1j * w1j);
            double c2m, tow2;
            if (j == 1) { // c2[0] and w2[1] from the true square root sqrt(x[0] x[1])
                const double w = s.x32;
                c2m = (w * w * w * w) / (1.0 + 4.60 * w + 1.1 * w * w);
                tow2 = s.tod / s.x32;
            } else {
                const double w = s.xmin * W1T[j - 1];
                c2m = (w * w * w * w) * __drcp_rn(1.0 + 4.60 * w + 1.1 * w * w);
                tow2 = s.todx * RW1T[j - 1];
            }
            const double betj = nth_bet(s, j, x0);
            const double dph = (j < s.jmaxth) ? s.planck * (x0 * x0) * __drcp_rn(rg_exp_pos(x0 * s.itbb) - 1) : 0.0;
            const double tow1 = s.todx * RW1T[j];
            const double aj = -s.c20 * c2j * (tow1 + 0.5);
            const double t1 = -s.c20 * c2j * (0.5 - tow1);
            const double t2 = s.c20 * c2m * (tow2 + 0.5);
            const double t3 = (x0 * x0 * x0) * (s.tautom * betj);
            RA[j] = aj;
            RB[j] = t1 + t2 + t3;
            RC[j] = s.c20 * c2m * (0.5 - tow2);
            RD[j] = x0 * dph;
        }
        __syncwarp();

        // ---- P2: Thomas sweep (pyNTHCOMP.py:226-247), one lane; gamma overwrites a, g then u overwrite d
        if (lane == 0) {
            double gam_prev, g_prev;
            {
                const double alp = RB[1] + RC[1] * s.aa;
                const double ralp = 1.0 / alp;
                g_prev = RD[1] * ralp;
                gam_prev = RA[1] * ralp;
                RA[1] = gam_prev; RD[1] = g_prev;
            }
            // (the loads of row j + 1 do not depend on the recurrence: they are issued ahead of the reciprocal chain)
            double a1 = RA[2], b1 = RB[2], c1 = RC[2], d1 = RD[2];
            for (int j = 2; j < jmax - 1; j++) {
                const double aj = a1, bj = b1, cj = c1, dj = d1;
                if (j + 1 < jmax - 1) { a1 = RA[j + 1]; b1 = RB[j + 1]; c1 = RC[j + 1]; d1 = RD[j + 1]; }
                const double alp = bj - cj * gam_prev;
                const double ralp = 1.0 / alp; // one reciprocal of the pivot serves both quotients (pyNTHCOMP.py:234-241)
                g_prev = (dj - cj * g_prev) * ralp;
                gam_prev = aj * ralp;
                RA[j] = gam_prev; RD[j] = g_prev;
            }
            double u_next = 0.0; // u[jmax - 1]
            for (int j = jmax - 2; j >= 1; j--) {
                const double u = RD[j] - RA[j] * u_next;
                RD[j] = u;
                u_next = u;
            }
            RD[0] = s.aa * u_next; // u[0] = aa u[1]
        }
        __syncwarp();

        // ---- P3: sptot = dphesc x^2 = x^4 u bet tau (pyNTHCOMP.py:170-171, 249) into the b array
        for (int j = lane; j <= jmax; j += 32) {
            double v = 0.0;
            if (j < jmax - 1) {
                const double x = s.xmin * P10[j];
                double bt;
                if (j == 0) { // bet[0] (pyNTHCOMP.py:150-165 at j = 0, with its own divisions)
                    const double w = x;
                    double rel;
                    if (w <= 0.05) rel = (1.0 - 2.0 * w + 26.0 * w * w * 0.2);
                    else {
                        double z1 = (1.0 + w) / (w * w * w), z2 = 1.0 + 2.0 * w, z3 = log(z2), z4 = 2.0 * w * (1.0 + w) / z2;
                        double z5 = z3 / 2.0 / w, z6 = (1.0 + 3.0 * w) / z2 / z2;
                        rel = (0.75 * (z1 * (z4 - z3) + z5 - z6));
                    }
                    if (0 < s.jnr - 1) bt = 1.0 / s.tautom / (1.0 + s.tautom * rel / 3.0);
                    else if (0 < s.jrel) { double flz = 1 - (w - s.xnr) / (xr - s.xnr); bt = 1.0 / s.tautom / (1.0 + s.tautom * rel / 3.0 * flz); }
                    else bt = 1.0 / s.tautom;
                } else bt = nth_bet(s, j, x);
                v = (x * x * RD[j] * bt * s.tautom) * (x * x);
            }
            RB[j] = v;
        }
        __syncwarp();
        // normalisation at 1 keV (pyNTHCOMP.py:49-56)
        const double *spt = RB;
        double normfac;
        {
            const double xx = 1.0 / 511.0;
            int ih = 1;
            {   // the reference's linear cursor `while (ih < jmax and xx > x[ih]) ih++`, seeded with its closed form
                int j0 = (int)(log10(xx / s.xmin) / delta) - 1;
                if (j0 < 1) j0 = 1;
                if (j0 > jmax) j0 = jmax;
                ih = j0;
                while (ih > 1 && !(xx > s.xmin * P10[ih - 1])) ih--;
                while (ih < jmax && xx > s.xmin * P10[ih]) ih++;
            }
            const int il = ih - 1;
            const double xl = s.xmin * P10[il], xh = s.xmin * P10[ih];
            const double spp = spt[il] + (spt[ih] - spt[il]) * (xx - xl) / (xh - xl);
            normfac = 1.0 / spp;
        }

        // ---- P4: interpolation onto the caller's grid (pyNTHCOMP.py:57-76), lanes over consecutive bins.  The
        //      reference's monotone cursor is restated as a direct search seeded from log10(E) and fixed up with the
        //      reference's own comparisons, so every bin lands in exactly the cell the cursor would have reached.
        {
            const double xmin = s.xmin, lgx = log10(511.0 * s.xmin), rxmin = 1.0 / s.xmin;
            const double outfac = (1.0 / RGK_KEV) * RGK_H; // photon flux -> W/Hz (relagn.py:929-930)
            const int nth = jmax;
            double carry_pr = 0.0, carry_e = 0.0; // bin i0 - 1 of the previous pass
            for (int i0 = 0; i0 < nE; i0 += 32) {
                const int i = i0 + lane;
                double e = 0.0, pr = 0.0;
                if (i < nE) {
                    e = __ldg(grid.Egrid + i);
                    int j = (int)ceil((__ldg(grid.lgE + i) - lgx) * 50.0);
                    if (j < 0) j = 0;
                    if (j > nth + 1) j = nth + 1;
                    while (j > 0 && !(511.0 * (xmin * P10[j - 1]) < e)) j--;
                    while (j <= nth && 511.0 * (xmin * P10[j]) < e) j++;
                    double prim = 0.0;
                    if (j <= nth) {
                        if (j > 0) {
                            const int jl = j - 1;
                            // linear interpolation inside the cell [xl, xh] (pyNTHCOMP.py:66-71); the division by the
                            // cell width is a multiplication by (1 / xmin) (1 / (10^(0.02 (jl+1)) - 10^(0.02 jl)))
                            const double xl = xmin * P10[jl];
                            const double sl = spt[jl], sh = spt[jl + 1];
                            prim = sl + ((__ldg(grid.e511 + i) - xl) * (sh - sl)) * (rxmin * RDP[jl]);
                        } else prim = spt[0];
                    }
                    pr = prim * __ldg(grid.rear2 + i); // prim / E^2 (pyNTHCOMP.py:75)
                }
                double pr_prev = __shfl_up_sync(0xffffffffu, pr, 1), e_prev = __shfl_up_sync(0xffffffffu, e, 1);
                if (lane == 0) { pr_prev = carry_pr; e_prev = carry_e; }
                if (i < nE) {
                    const double ph = i > 0 ? (0.5 * (pr + pr_prev) * (e - e_prev) * normfac) : 0.0;
                    F[i] = ph * outfac;
                }
                carry_pr = __shfl_sync(0xffffffffu, pr, 31);
                carry_e = __shfl_sync(0xffffffffu, e, 31);
            }
        }
        __syncwarp(); // the rows are rewritten by the warp's next system
        sumj += (unsigned long long)jmax;
        nsolved++;
    }
    if (stats && lane == 0 && nsolved) {
        atomicAdd(&stats[2], sumj);
        atomicAdd(&stats[3], nsolved);
    }
}

// gscratch: global rows for grids longer than NTH_RCAP, 4 (RG_NTH + 4) doubles per warp of the launch; gscratch_doubles
// bounds the number of CTAs
int rg_launch_nthcomp_small(const SetRec *recs, const int2 *syslist, int n_warm_sys, int n_hot_sys, int sys_cap,
                            const double *p10tab, const double2 *edges, double *gscratch, size_t gscratch_doubles,
                            int sm_count, const GridDesc &grid, double *farr, int *queue, unsigned long long *stats,
                            cudaStream_t st)
{
    // logical order: the warm systems, padded to a multiple of 32, then the hot shapes
    const int n_warm_pad = (n_warm_sys + 31) & ~31;
    const int nsys = n_hot_sys > 0 ? n_warm_pad + n_hot_sys : n_warm_sys;
    const int hot_first = sys_cap - n_hot_sys;
    if (nsys <= 0) return RG_OK;
    const size_t smem = (size_t)NTH_WARPS * 4 * NTH_RCAP * sizeof(double);
    RG_CUDA_OK(cudaFuncSetAttribute(nthcomp_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long ctas = ((long)nsys + NTH_WARPS - 1) / NTH_WARPS;
    if (ctas > 2L * sm_count) ctas = 2L * sm_count; // persistent: two CTAs per SM pull systems from the queue
    const long fit = (long)(gscratch_doubles / ((size_t)NTH_WARPS * 4 * (RG_NTH + 4)));
    if (ctas > fit) ctas = fit;
    if (ctas < 1) return rg_set_error(RG_E_NOMEM, "nthcomp_small: no scratch");
    RG_CUDA_OK(cudaMemsetAsync(queue, 0, sizeof(int), st));
    RG_LAUNCH(nthcomp_small_kernel, dim3((unsigned)ctas), dim3(32 * NTH_WARPS), smem, st, recs, syslist, nsys, n_warm_sys,
              n_warm_pad, hot_first, p10tab, edges, gscratch, grid, farr, queue, stats);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}
