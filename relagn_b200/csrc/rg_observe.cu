// rg_observe.cu -- the observation epilogue on the device:
//   convert_kernel   unit conversion + flux of whole batches           relagn.py:487-564 (_to_newUnit, _to_flux)
//   observe_kernel   per parameter set, one CTA:
//                      photons/cm^2/s per model bin -> observer frame (edges and photons / (1 + z)) -> rebin onto the
//                      instrument grid by fractional overlap          relagn.f:98-112 (XSPEC inibin / erebin)
//                      -> fold through the instrument response (sliced-ELL, shared by every set, L2 resident)
//                      -> chi^2 / Cash against the observed counts: one double per set leaves the GPU.
// Both kernels are HBM-bound streaming passes over the [P][nE] spectra (8 B in + 8 B out per bin for the
// conversion, 8 B in per bin for the statistic).  The arithmetic follows oracle/observe_oracle.c operation for
// operation (explicit _rn intrinsics where nvcc would otherwise contract a*b+c), so photar and the folded counts are
// bit-identical to the oracle; the statistic differs only by the order of the final sum over channels.
#include "rg_kernels.h"

#define OBS_THREADS 256

__device__ inline int par_slot_z(int model) { return model == RG_MODEL_RELQSO ? 6 : 14; }

// ---------------------------------------------------------------------------
// f1: unit conversion.  One warp per spectrum row ([nE] bins of one set / component), grid-stride over rows; the
// per-bin factors E and nu come from the resident grid (L1/L2 hits), the per-set denominator is computed once per row.
template <int UNIT>
__device__ inline double convert_bin(double v, const GridDesc &g, int j)
{
    if (UNIT == RG_UNIT_CGS) return v * 1e7;                                  // relagn.py:507
    if (UNIT == RG_UNIT_COUNTS) return __ddiv_rn(__ddiv_rn(v, RGK_H), __ldg(g.Egrid + j)); // :509-516
    if (UNIT == RG_UNIT_CGS_WAVE || UNIT == RG_UNIT_SI_WAVE) {                // :518-524
        const double nu = __ldg(g.nu + j);
        v = __dmul_rn(__ddiv_rn(__dmul_rn(v, __dmul_rn(nu, nu)), RGK_C), 1e-10);
        return UNIT == RG_UNIT_CGS_WAVE ? __dmul_rn(v, 1e7) : v;
    }
    return v;
}

template <int UNIT, bool FLUX>
__global__ void __launch_bounds__(256) convert_kernel(const double *__restrict__ lnu, const double *__restrict__ params,
                                                      long n_rows, int rows_per_set, int model, GridDesc g,
                                                      double *__restrict__ out)
{
    const int nE = g.nE, lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((long)gridDim.x * blockDim.x) >> 5;
    for (long row = warp0; row < n_rows; row += nwarp) {
        double den = 1.0;
        if (FLUX) { // relagn.py:556-561
            const double *p = params + (row / rows_per_set) * RG_NPAR;
            double d = __dmul_rn(__ldg(p + 1), RGK_MPC_CM);
            if (UNIT == RG_UNIT_SI || UNIT == RG_UNIT_SI_WAVE) d = d / 100;
            den = __dmul_rn(__dmul_rn(4 * RGK_PI, __dmul_rn(d, d)), 1 + __ldg(p + par_slot_z(model)));
        }
        const double *src = lnu + row * nE;
        double *dst = out + row * nE;
        if ((nE & 1) == 0 && (((uintptr_t)lnu | (uintptr_t)out) & 15) == 0) { // 16-byte accesses (rows stay aligned)
            const double2 *src2 = reinterpret_cast<const double2 *>(src);
            double2 *dst2 = reinterpret_cast<double2 *>(dst);
#pragma unroll 4
            for (int j = lane; j < (nE >> 1); j += 32) {
                double2 v = RG_LDCS(src2 + j);
                v.x = convert_bin<UNIT>(v.x, g, 2 * j);
                v.y = convert_bin<UNIT>(v.y, g, 2 * j + 1);
                if (FLUX) { v.x = __ddiv_rn(v.x, den); v.y = __ddiv_rn(v.y, den); }
                RG_STCS(dst2 + j, v);
            }
        } else {
#pragma unroll 4
            for (int j = lane; j < nE; j += 32) {
                double v = convert_bin<UNIT>(RG_LDCS(src + j), g, j);
                if (FLUX) v = __ddiv_rn(v, den);
                RG_STCS(dst + j, v);
            }
        }
    }
}

template <int UNIT>
static int launch_convert_u(const double *d_lnu, const double *d_params, int P, int model, int ncomp, int as_flux,
                            const GridDesc &g, double *d_out, int sm_count, cudaStream_t st)
{
    const long rows = (long)P * ncomp;
    long blocks = (rows + 7) / 8; // 8 warps per CTA
    const long cap = (long)sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (as_flux)
        RG_LAUNCH((convert_kernel<UNIT, true>), dim3((unsigned)blocks), dim3(256), 0, st, d_lnu, d_params, rows, ncomp, model, g, d_out);
    else
        RG_LAUNCH((convert_kernel<UNIT, false>), dim3((unsigned)blocks), dim3(256), 0, st, d_lnu, d_params, rows, ncomp, model, g, d_out);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}

int rg_launch_convert(const double *d_lnu, const double *d_params, int P, int model, int ncomp, int unit, int as_flux,
                      const GridDesc &g, double *d_out, int sm_count, cudaStream_t st)
{
    if ((long)P * ncomp == 0) return RG_OK;
    switch (unit) {
    case RG_UNIT_SI: return launch_convert_u<RG_UNIT_SI>(d_lnu, d_params, P, model, ncomp, as_flux, g, d_out, sm_count, st);
    case RG_UNIT_CGS: return launch_convert_u<RG_UNIT_CGS>(d_lnu, d_params, P, model, ncomp, as_flux, g, d_out, sm_count, st);
    case RG_UNIT_CGS_WAVE: return launch_convert_u<RG_UNIT_CGS_WAVE>(d_lnu, d_params, P, model, ncomp, as_flux, g, d_out, sm_count, st);
    case RG_UNIT_SI_WAVE: return launch_convert_u<RG_UNIT_SI_WAVE>(d_lnu, d_params, P, model, ncomp, as_flux, g, d_out, sm_count, st);
    case RG_UNIT_COUNTS: return launch_convert_u<RG_UNIT_COUNTS>(d_lnu, d_params, P, model, ncomp, as_flux, g, d_out, sm_count, st);
    }
    return rg_set_error(RG_E_ARG, "unknown unit %d", unit);
}

// ---------------------------------------------------------------------------
// f1 (rebin) + f2: one CTA per parameter set, persistent grid-stride over sets.
// shared memory: x[nE+1] observer-frame edges of the model grid | ph[nE] photons/cm^2/s per model bin |
//                pa[ne] photons/cm^2/s per instrument bin | red[OBS_THREADS/32]
__global__ void __launch_bounds__(OBS_THREADS) observe_kernel(ObsArgs A)
{
    RG_DYN_SMEM(smem_raw);
    double *x = reinterpret_cast<double *>(smem_raw);
    double *ph = x + (A.nE + 1);
    double *pa = ph + A.nE;
    double *red = pa + A.ne;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int nE = A.nE, ne = A.ne;
    for (int set = blockIdx.x; set < A.P; set += gridDim.x) {
        const double *p = A.params + (size_t)set * RG_NPAR;
        const double d = __dmul_rn(__ldg(p + 1), RGK_MPC_CM);
        const double zf = 1 + __ldg(p + par_slot_z(A.model));
        const double inv_den = __ddiv_rn(1.0, __dmul_rn(__dmul_rn(4 * RGK_PI, __dmul_rn(d, d)), zf));
        const double *L = A.lnu + (size_t)set * nE;
        // ---- observer-frame edges (relagn.f:98-102) and photons per model bin: counts flux (relagn.py:509-516, 561)
        //      times the bin width, as L * [dE / (h E)] * [1 / (4 pi d^2 (1 + z))] ----
        for (int j = tid; j <= nE; j += nthr) x[j] = __ddiv_rn(__ldg(A.ebins + j), zf);
        for (int j = tid; j < nE; j += nthr) ph[j] = __dmul_rn(__dmul_rn(RG_LDCS(L + j), __ldg(A.kph + j)), inv_den);
        __syncthreads();
        // ---- rebin onto the instrument grid (inibin / erebin): fractional overlap, fixed summation order ----
        const double zshift = A.pos0 ? log(zf) * A.inv_dlnE : 0.0;
        for (int i = tid; i < ne; i += nthr) {
            const double lo = __ldg(A.ear + i), hi = __ldg(A.ear + i + 1);
            double acc = 0.0;
            if (hi > x[0] && lo < x[nE] && hi > lo) {
                // js: first model bin with x[js+1] > lo.  Geometric model grid: start from the analytic position
                // (edge position at z = 0, precomputed per instrument edge, plus the redshift); else bisect.
                int js;
                if (A.pos0) {
                    const double q = floor(__ldg(A.pos0 + i) + zshift);
                    js = q < 0 ? 0 : (q > nE - 1 ? nE - 1 : (int)q);
                } else {
                    int a = 0, b = nE - 1;
                    while (a < b) { int m = (a + b) >> 1; if (x[m + 1] > lo) b = m; else a = m + 1; }
                    js = a;
                }
                while (js < nE - 1 && x[js + 1] <= lo) js++;
                while (js > 0 && x[js] > lo) js--;
                // je: last model bin with x[je] < hi
                int je = js;
                while (je + 1 < nE && x[je + 1] < hi) je++;
                if (js == je) {
                    double l = lo > x[js] ? lo : x[js], r = hi < x[js + 1] ? hi : x[js + 1];
                    acc = __dmul_rn(ph[js], __ddiv_rn(__dsub_rn(r, l), __dsub_rn(x[js + 1], x[js])));
                } else {
                    double l = lo > x[js] ? lo : x[js];
                    acc = __dmul_rn(ph[js], __ddiv_rn(__dsub_rn(x[js + 1], l), __dsub_rn(x[js + 1], x[js])));
                    for (int k = js + 1; k < je; k++) acc = __dadd_rn(acc, ph[k]);
                    double r = hi < x[je + 1] ? hi : x[je + 1];
                    acc = __dadd_rn(acc, __dmul_rn(ph[je], __ddiv_rn(__dsub_rn(r, x[je]), __dsub_rn(x[je + 1], x[je]))));
                }
            }
            pa[i] = acc;
            if (A.photar) A.photar[(size_t)set * ne + i] = acc;
        }
        if (A.n_chan == 0) { __syncthreads(); continue; } // photar only
        __syncthreads();
        // ---- fold through the response (sliced ELL: slice = 32 channels, entries [k][lane]; padding entries are
        //      (column 0, value 0) and add +0) + statistic ----
        double part = 0.0;
        for (int c = tid; c < ((A.n_chan + 31) & ~31); c += nthr) {
            const int s = c >> 5, lane = c & 31;
            const int off = __ldg(A.sl_off + s), w = __ldg(A.sl_off + s + 1) - off;
            const int *cp = A.ell_col + off + lane;
            const double *vp = A.ell_val + off + lane;
            double m = 0.0;
#pragma unroll 4
            for (int k = 0; k < w; k += 32) m = __dadd_rn(m, __dmul_rn(__ldg(vp + k), pa[__ldg(cp + k)]));
            if (c < A.n_chan) {
                m = __dmul_rn(A.exposure, m);
                if (A.folded) A.folded[(size_t)set * A.n_chan + c] = m;
                const double D = __ldg(A.counts + c);
                if (A.kind == RG_FIT_CHI2) {
                    const double t = __dmul_rn(__dsub_rn(D, m), __ldg(A.sigma + c)); // sigma holds 1 / sigma_c
                    part += t * t;
                } else {
                    const double mm = m < 1e-300 ? 1e-300 : m;
                    const double t = D > 0 ? (mm - D) + D * (__ldg(A.sigma + c) - log(mm)) : mm; // sigma holds ln D_c
                    part += 2 * t;
                }
            }
        }
        // fixed-order block reduction
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((tid & 31) == 0) red[tid >> 5] = part;
        __syncthreads();
        if (tid == 0) {
            double s = 0;
            for (int w = 0; w < (nthr >> 5); w++) s += red[w];
            A.stat[set] = s;
        }
        __syncthreads();
    }
}

size_t rg_observe_smem_bytes(int nE, int ne) { return ((size_t)(nE + 1) + nE + ne + OBS_THREADS / 32) * sizeof(double); }

int rg_launch_observe(const ObsArgs &A, int sm_count, cudaStream_t st)
{
    if (A.P == 0) return RG_OK;
    size_t smem = rg_observe_smem_bytes(A.nE, A.ne);
    if (smem > 227 * 1024 - 1024) return rg_set_error(RG_E_GRID, "model + instrument grids exceed the shared-memory working set");
    RG_CUDA_OK(cudaFuncSetAttribute(observe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    long ctas = (long)sm_count * per_sm;
    if (ctas > A.P) ctas = A.P;
    RG_LAUNCH(observe_kernel, dim3((unsigned)ctas), dim3(OBS_THREADS), smem, st, A);
    RG_CUDA_OK(cudaGetLastError());
    return RG_OK;
}
