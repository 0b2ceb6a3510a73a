// kernels_diag.cuh -- the grid-space diagnostics the reference evaluates inside the radial loop on log steps
// (rIter.f90:303-373), fused into one pass over the synthesised grid fields of a level chunk:
//   get_helicity   outMisc.f90:1052-1167      get_hemi      outMisc.f90:991-1050
//   get_visc_heat  power.f90:384-441          get_perpPar   outPar.f90:646-726
//   get_fluxes     outPar.f90:470-582         get_nlBLayers outPar.f90:584-644
// The reference calls them one after the other, each sweeping the (theta, phi) arrays of one level on the host; here one kernel
// reads every field of a (k, phi) point pair once (E/O layout of kernels_grid.cuh: north = E+O, south = E-O) and accumulates
// all requested sums.  Reductions are fixed-shape (per thread -> xor-shuffle tree -> warps in order -> CTAs in order): no
// floating-point atomics, bitwise reproducible.  HBM-bound: 8 n_theta n_phi bytes per field and level, read once.
#pragma once
#include "common.cuh"

namespace magic {

// grid field index of every field the diagnostics read (-1 = absent); same order as the loader below
struct DiagIn { int vr, vt, vp, cvr, dvrdr, dvtdr, dvpdr, dvrdt, dvrdp, dvtdp, dvpdp, s, p, drs, dsdt, dsdp, br, bt, bp, cbt, cbp; };
constexpr int DIAG_NF = 21;
constexpr int DIAG_NSLOT = 32;   // sums accumulated by diag_kernel (registers)
constexpr int DIAG_NOUT = 40;    // = MAGIC_NDIAG: doubles per level of the result (slots 32.. come from diag_phase_kernel)
constexpr int DIAG_NPHASE = 5;   // ekinS, ekinL, volS, min(phi), max(phi)
constexpr int DIAG_NMEAN = 8;    // phi means of vr, cvr, vt, vp, dvrdp, dvpdr, dvtdr, dvrdt (outMisc.f90:1091-1108)
constexpr int DIAG_THREADS = 256;

enum DiagSlot {
    DG_HEL_N = 0, DG_HEL_S, DG_HEL2_N, DG_HEL2_S, DG_HELNA_N, DG_HELNA_S, DG_HELNA2_N, DG_HELNA2_S, DG_HELEA,
    DG_EKIN_N, DG_EKIN_S, DG_VRABS_N, DG_VRABS_S, DG_EMAG_N, DG_EMAG_S, DG_BRABS_N, DG_BRABS_S,
    DG_VISC,
    DG_EPERP, DG_EPAR, DG_EPERPAXI, DG_EPARAXI,
    DG_FKIN, DG_FCONV_S, DG_FCONV_P, DG_FVISC, DG_FRES, DG_FPOYN,
    DG_UH, DG_DUH, DG_GRADT2
};
enum DiagPhaseSlot { DG_PH_EKINS = 32, DG_PH_EKINL, DG_PH_VOLS, DG_PH_MIN, DG_PH_MAX };
enum DiagMask { DM_HEL = 1, DM_HEMI = 2, DM_POWER = 4, DM_PERPPAR = 8, DM_FLUX = 16, DM_VISCBC = 32, DM_PHASE = 64 };

struct DiagArgs {
    DiagIn di;
    const double *gin;
    int n_lev, nh, n_phi, mask;
    int l_mag, l_mag_nl, n_r_max, ktops, kbots;
    double omega_ma, omega_ic, r_cmb, r_icb;
    const LevelInfo *lev;         // the diagnostics' own copy: lDeriv = 1 everywhere, nBc = 0 with lRmsCalc (rIter.f90:190-215)
    const double *sinth, *costh;  // northern values [nh]
    const double *gauss;          // Gauss weight of colatitude pair k [nh]
    int mean_field[DIAG_NMEAN];   // grid field index of each averaged field
    double *means;                // [DIAG_NMEAN][n_lev][2 (E, O)][nh]
    double *partial;              // [n_lev][gridDim.x][DIAG_NSLOT]
};

// phi means of the E and O rows: one warp per row, lane-strided partial sums, xor-shuffle tree
__global__ void __launch_bounds__(DIAG_THREADS) diag_mean_kernel(DiagArgs a) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t nrow = (size_t)DIAG_NMEAN * a.n_lev * 2 * a.nh;
    const size_t row = (size_t)blockIdx.x * (DIAG_THREADS / 32) + warp;
    if (row >= nrow) return;
    const int k = (int)(row % a.nh);
    const size_t t = row / a.nh;
    const int s = (int)(t % 2), lev = (int)((t / 2) % a.n_lev), f = (int)(t / 2 / a.n_lev);
    double sum = 0.0;
    if (a.mean_field[f] >= 0) {
        const double *g = a.gin + ((((size_t)a.mean_field[f] * a.n_lev + lev) * 2 + s) * a.nh + k) * a.n_phi;
        for (int j = lane; j < a.n_phi; j += 32) sum += g[j];
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) a.means[row] = sum * (1.0 / (double)a.n_phi);
}

struct DiagPoint { double vr, vt, vp, cvr, dvrdr, dvtdr, dvpdr, dvrdt, dvrdp, dvtdp, dvpdp, s, p, drs, dsdt, dsdp, br, bt, bp, cbt, cbp; };

// boundary values of transform_to_grid_space with lDeriv = .true. (rIter.f90:555-602, v_rigid_boundary nonlinear_bcs.f90:120-175);
// applies to point values and to phi means alike (the overrides do not depend on phi)
__device__ __forceinline__ void diag_override_raw(const LevelInfo &L, double omega_ma, double omega_ic, double r_cmb, double r_icb, double st,
                                                  double ct, double &vr, double &vt, double &vp, double &cvr) {
    if (L.nBc == 1) vr = 0.0;
    if (L.nBc == 2) {
        const double r2 = (L.nR == 1) ? r_cmb * r_cmb : r_icb * r_icb;
        const double om = (L.nR == 1) ? omega_ma : omega_ic;
        vr = 0.0;
        vt = 0.0;
        vp = r2 * L.rho0 * (st * st) * om;
        cvr = r2 * L.rho0 * 2.0 * ct * om;
    }
}
__device__ __forceinline__ void diag_override(const DiagArgs &a, const LevelInfo &L, double st, double ct, double &vr, double &vt, double &vp,
                                              double &cvr) {
    diag_override_raw(L, a.omega_ma, a.omega_ic, a.r_cmb, a.r_icb, st, ct, vr, vt, vp, cvr);
}

// the sums of one grid point; h = 0 north, 1 south; ct carries the hemisphere sign; m[] = phi means of this hemisphere
__device__ __forceinline__ void diag_point(const DiagArgs &a, const LevelInfo &L, const DiagPoint &p, const double *m, int h, double st, double ct,
                                           double ga, double *acc) {
    const double or1 = L.or1, or2 = L.or2, or4 = L.or4, orho1 = L.orho1, orho2 = L.orho2, beta = L.beta, r = L.r, visc = L.visc;
    const double os2 = 1.0 / (st * st), st2 = st * st, cn2 = ct / st / st;
    const double pn1 = 1.0 / (double)a.n_phi, pn2 = 6.283185307179586476925286766559 / (double)a.n_phi;
    const double w1 = pn1 * ga, w2 = pn2 * ga;
    if (a.mask & DM_HEL) {
        const double vras = m[0], cvras = m[1], vtas = m[2], vpas = m[3], dvrdpas = m[4], dvpdras = m[5], dvtdras = m[6], dvrdtas = m[7];
        const double vrna = p.vr - vras, cvrna = p.cvr - cvras, vtna = p.vt - vtas, vpna = p.vp - vpas;
        const double dvrdpna = p.dvrdp - dvrdpas;
        const double dvpdrna = p.dvpdr - beta * p.vp - dvpdras + beta * vpas;
        const double dvtdrna = p.dvtdr - beta * p.vt - dvtdras + beta * vtas;
        const double dvrdtna = p.dvrdt - dvrdtas;
        const double Hel = or4 * orho2 * p.vr * p.cvr +
                           or2 * orho2 * os2 * (p.vt * (or2 * p.dvrdp - p.dvpdr + beta * p.vp) + p.vp * (p.dvtdr - beta * p.vt - or2 * p.dvrdt));
        const double Helna = or4 * orho2 * vrna * cvrna + or2 * orho2 * os2 * (vtna * (or2 * dvrdpna - dvpdrna) + vpna * (dvtdrna - or2 * dvrdtna));
        acc[DG_HEL_N + h] += w1 * Hel;
        acc[DG_HEL2_N + h] += w1 * Hel * Hel;
        acc[DG_HELNA_N + h] += w1 * Helna;
        acc[DG_HELNA2_N + h] += w1 * Helna * Helna;
        acc[DG_HELEA] += (h ? -1.0 : 1.0) * w1 * Hel;
    }
    if (a.mask & DM_HEMI) {
        acc[DG_EKIN_N + h] += w2 * (0.5 * orho1 * (or2 * p.vr * p.vr + os2 * p.vt * p.vt + os2 * p.vp * p.vp));
        acc[DG_VRABS_N + h] += w2 * (orho1 * fabs(p.vr));
        if (a.l_mag) {
            acc[DG_EMAG_N + h] += w2 * (0.5 * (or2 * p.br * p.br + os2 * p.bt * p.bt + os2 * p.bp * p.bp));
            acc[DG_BRABS_N + h] += w2 * fabs(p.br);
        }
    }
    if (a.mask & DM_POWER) {
        const double t1 = p.dvrdr - (2.0 * or1 + beta) * p.vr;
        const double t2 = cn2 * p.vt + p.dvpdp + p.dvrdr - or1 * p.vr;
        const double t3 = p.dvpdp + cn2 * p.vt + or1 * p.vr;
        const double t6 = 2.0 * p.dvtdp + p.cvr - 2.0 * cn2 * p.vp;
        const double t4 = r * p.dvtdr - (2.0 + beta * r) * p.vt + or1 * p.dvrdt;
        const double t5 = r * p.dvpdr - (2.0 + beta * r) * p.vp + or1 * p.dvrdp;
        const double t7 = beta * p.vr;
        acc[DG_VISC] += w2 * (or2 * orho1 * visc *
                              (2.0 * t1 * t1 + 2.0 * t2 * t2 + 2.0 * t3 * t3 + t6 * t6 + os2 * (t4 * t4 + t5 * t5) - 2.0 * (1.0 / 3.0) * t7 * t7));
    }
    if (a.mask & DM_PERPPAR) {
        const double f = 0.5 * or2 * orho2, vras = m[0], vtas = m[2], vpas = m[3];
        acc[DG_EPERP] += w1 * (f * (or2 * st2 * p.vr * p.vr + (os2 - 1.0) * p.vt * p.vt + 2.0 * or1 * ct * p.vr * p.vt + os2 * p.vp * p.vp));
        acc[DG_EPAR] += w1 * (f * (or2 * (1.0 - st2) * p.vr * p.vr + p.vt * p.vt - 2.0 * or1 * ct * p.vr * p.vt));
        acc[DG_EPERPAXI] += w1 * (f * (or2 * st2 * vras * vras + (os2 - 1.0) * vtas * vtas + 2.0 * or1 * ct * vras * vtas + os2 * vpas * vpas));
        acc[DG_EPARAXI] += w1 * (f * (or2 * (1.0 - st2) * vras * vras + vtas * vtas - 2.0 * or1 * ct * vras * vtas));
    }
    if (a.mask & DM_FLUX) {
        const bool bulk = L.nR != 1 && L.nR != a.n_r_max;
        double fvisc = 0.0;
        if (bulk)
            fvisc = -2.0 * visc * orho1 * p.vr * or2 * (p.dvrdr - (2.0 * or1 + 2.0 * (1.0 / 3.0) * beta) * p.vr) -
                    visc * orho1 * p.vt * os2 * (or2 * p.dvrdt + p.dvtdr - (2.0 * or1 + beta) * p.vt) -
                    visc * orho1 * p.vp * os2 * (or2 * p.dvrdp + p.dvpdr - (2.0 * or1 + beta) * p.vp);
        acc[DG_FKIN] += w2 * (0.5 * or2 * orho2 * (os2 * (p.vt * p.vt + p.vp * p.vp) + or2 * p.vr * p.vr) * p.vr);
        acc[DG_FCONV_S] += w2 * (p.vr * p.s);
        acc[DG_FCONV_P] += w2 * (p.vr * p.p);
        acc[DG_FVISC] += w2 * fvisc;
        if (a.l_mag_nl) {
            acc[DG_FRES] += w2 * (os2 * (p.cbt * p.bp - p.cbp * p.bt));
            acc[DG_FPOYN] += w2 * (-orho1 * or2 * os2 * (p.vp * p.br * p.bp - p.vr * p.bp * p.bp - p.vr * p.bt * p.bt + p.vt * p.br * p.bt));
        }
    }
    if (a.mask & DM_VISCBC) {
        const double uh = or2 * orho2 * os2 * (p.vt * p.vt + p.vp * p.vp);
        const double duh = or2 * orho2 * os2 * (p.dvtdr * p.vt - (or1 + beta) * p.vt * p.vt + p.dvpdr * p.vp - (or1 + beta) * p.vp * p.vp);
        const double grads = p.drs * p.drs + or2 * os2 * (p.dsdt * p.dsdt + p.dsdp * p.dsdp);
        acc[DG_UH] += w1 * sqrt(uh);
        if (uh != 0.0) acc[DG_DUH] += w1 * fabs(duh) / sqrt(uh);
        acc[DG_GRADT2] += w1 * grads;
    }
}

#ifndef MAGIC_DIAG_MINB
#define MAGIC_DIAG_MINB 1   // resident CTAs per SM asked of ptxas (2 caps the kernel at 128 registers)
#endif
__global__ void __launch_bounds__(DIAG_THREADS, MAGIC_DIAG_MINB) diag_kernel(DiagArgs a) {
    const int lev = blockIdx.y;
    const LevelInfo L = a.lev[lev];
    const size_t plane = (size_t)a.nh * a.n_phi;
    const size_t mstride = (size_t)a.n_lev * 2 * a.nh;  // one averaged field
    double acc[DIAG_NSLOT];
#pragma unroll
    for (int i = 0; i < DIAG_NSLOT; i++) acc[i] = 0.0;
    const bool zero_grad_s = (L.nR == 1 && a.ktops == 1) || (L.nR == a.n_r_max && a.kbots == 1);  // rIter.f90:488-495
    for (unsigned pt = blockIdx.x * blockDim.x + threadIdx.x; pt < (unsigned)plane; pt += gridDim.x * blockDim.x) {
        const int k = (int)(pt / (unsigned)a.n_phi);
        const double st = a.sinth[k], ct = a.costh[k], ga = a.gauss[k];
        const int *fidx = &a.di.vr;
        double re[DIAG_NF], ro[DIAG_NF];
#pragma unroll
        for (int f = 0; f < DIAG_NF; f++) {  // every load before the first use (in-order issue)
            re[f] = 0.0;
            ro[f] = 0.0;
            if (fidx[f] >= 0) {
                const double *base = a.gin + (((size_t)fidx[f] * a.n_lev + lev) * 2) * plane + pt;
                re[f] = __ldg(base);
                ro[f] = __ldg(base + plane);
            }
        }
        DiagPoint pn, ps;
        double *pnv = &pn.vr, *psv = &ps.vr;
#pragma unroll
        for (int f = 0; f < DIAG_NF; f++) {
            pnv[f] = re[f] + ro[f];
            psv[f] = re[f] - ro[f];
        }
        const double os2 = 1.0 / (st * st);  // torpol_to_dphspat post-scaling, sht_native.f90:263-270
        pn.dvtdp *= os2; ps.dvtdp *= os2; pn.dvpdp *= os2; ps.dvpdp *= os2;
        if (zero_grad_s) { pn.dsdt = ps.dsdt = 0.0; pn.dsdp = ps.dsdp = 0.0; }
        diag_override(a, L, st, ct, pn.vr, pn.vt, pn.vp, pn.cvr);
        diag_override(a, L, st, -ct, ps.vr, ps.vt, ps.vp, ps.cvr);
        double mn[DIAG_NMEAN], ms[DIAG_NMEAN];
        if (a.mask & (DM_HEL | DM_PERPPAR)) {
#pragma unroll
            for (int f = 0; f < DIAG_NMEAN; f++) {
                const double e = a.means[(size_t)f * mstride + ((size_t)lev * 2 + 0) * a.nh + k];
                const double o = a.means[(size_t)f * mstride + ((size_t)lev * 2 + 1) * a.nh + k];
                mn[f] = e + o;
                ms[f] = e - o;
            }
            diag_override(a, L, st, ct, mn[0], mn[2], mn[3], mn[1]);
            diag_override(a, L, st, -ct, ms[0], ms[2], ms[3], ms[1]);
        }
        diag_point(a, L, pn, mn, 0, st, ct, ga, acc);
        diag_point(a, L, ps, ms, 1, st, -ct, ga, acc);
    }
    __shared__ double red[DIAG_THREADS / 32][DIAG_NSLOT];
#pragma unroll
    for (int i = 0; i < DIAG_NSLOT; i++) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < DIAG_NSLOT) {
        double v = 0.0;
        for (int w = 0; w < DIAG_THREADS / 32; w++) v += red[w][threadIdx.x];
        a.partial[((size_t)lev * gridDim.x + blockIdx.x) * DIAG_NSLOT + threadIdx.x] = v;
    }
}

// CTA partials added in CTA order -> out[lev][slot] (rows of DIAG_NOUT doubles)
__global__ void diag_finish_kernel(const double *partial, int n_part, int n_lev, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lev * DIAG_NSLOT) return;
    const int lev = i / DIAG_NSLOT, slot = i - lev * DIAG_NSLOT;
    double v = 0.0;
    for (int j = 0; j < n_part; j++) v += partial[((size_t)lev * n_part + j) * DIAG_NSLOT + slot];
    out[(size_t)lev * DIAG_NOUT + slot] = v;
}

// get_ekin_solid_liquid (outMisc.f90:1169-1221): kinetic energy of the points with phi >= 1/2 (solid) and of the others (liquid),
// the volume of the solid, and the extrema of phi on the grid (the phase_min / phase_max columns of phase.TAG,
// outMisc.f90:952-953).  Reads vr, vt, vp (a.di) and the phase field (grid field phi_field); same reduction shape as diag_kernel.
__global__ void __launch_bounds__(DIAG_THREADS) diag_phase_kernel(DiagArgs a, int phi_field, double *partial) {
    const int lev = blockIdx.y;
    const LevelInfo L = a.lev[lev];
    const size_t plane = (size_t)a.nh * a.n_phi;
    const double pn2 = 6.283185307179586476925286766559 / (double)a.n_phi;
    double acc[DIAG_NPHASE] = {0.0, 0.0, 0.0, 1e300, -1e300};
    for (unsigned pt = blockIdx.x * blockDim.x + threadIdx.x; pt < (unsigned)plane; pt += gridDim.x * blockDim.x) {
        const int k = (int)(pt / (unsigned)a.n_phi);
        const double st = a.sinth[k], ct = a.costh[k], w2 = pn2 * a.gauss[k], os2 = 1.0 / (st * st);
        const int fi[4] = {a.di.vr, a.di.vt, a.di.vp, phi_field};
        double e[4], o[4];
#pragma unroll
        for (int f = 0; f < 4; f++) {
            const double *base = a.gin + (((size_t)fi[f] * a.n_lev + lev) * 2) * plane + pt;
            e[f] = __ldg(base);
            o[f] = __ldg(base + plane);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            double vr = h ? e[0] - o[0] : e[0] + o[0], vt = h ? e[1] - o[1] : e[1] + o[1], vp = h ? e[2] - o[2] : e[2] + o[2];
            const double phi = h ? e[3] - o[3] : e[3] + o[3];
            double cvr = 0.0;
            diag_override(a, L, st, h ? -ct : ct, vr, vt, vp, cvr);
            const double ekin = 0.5 * L.orho1 * (L.or2 * vr * vr + os2 * vt * vt + os2 * vp * vp);
            if (phi >= 0.5) {
                acc[0] += w2 * ekin;
                acc[2] += w2 * L.r * L.r;
            } else {
                acc[1] += w2 * ekin;
            }
            acc[3] = fmin(acc[3], phi);
            acc[4] = fmax(acc[4], phi);
        }
    }
    __shared__ double red[DIAG_THREADS / 32][DIAG_NPHASE];
#pragma unroll
    for (int i = 0; i < DIAG_NPHASE; i++) {
        double v = acc[i];
        for (int s = 16; s > 0; s >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, s);
            v = i < 3 ? v + u : (i == 3 ? fmin(v, u) : fmax(v, u));
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < DIAG_NPHASE) {
        const int i = threadIdx.x;
        double v = red[0][i];
        for (int w = 1; w < DIAG_THREADS / 32; w++) v = i < 3 ? v + red[w][i] : (i == 3 ? fmin(v, red[w][i]) : fmax(v, red[w][i]));
        partial[((size_t)lev * gridDim.x + blockIdx.x) * DIAG_NPHASE + i] = v;
    }
}

__global__ void diag_phase_finish_kernel(const double *partial, int n_part, int n_lev, double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lev * DIAG_NPHASE) return;
    const int lev = i / DIAG_NPHASE, slot = i - lev * DIAG_NPHASE;
    double v = partial[((size_t)lev * n_part) * DIAG_NPHASE + slot];
    for (int j = 1; j < n_part; j++) {
        const double u = partial[((size_t)lev * n_part + j) * DIAG_NPHASE + slot];
        v = slot < 3 ? v + u : (slot == 3 ? fmin(v, u) : fmax(v, u));
    }
    out[(size_t)lev * DIAG_NOUT + DG_PH_EKINS + slot] = v;
}

}  // namespace magic

// ------------------------------------------------------------------------------------------------------
// get_dtBLM (dtB.f90:144-223): the eleven grid products of the magnetic-field production diagnostics, written as the (N+S, N-S)
// rows the r2c FFT expects (product field k of gout): BtVr, BpVr, BrVt, BrVp, BtVp, BpVt, BpVtBtVpCot, BpVtBtVpSn2, BrVZ, BtVZ,
// BtVZsn2.  gin holds vr, vt, vp, br, bt, bp (fields 0..5); means[0] = phi mean of vp (E/O rows).
namespace magic {

struct DtbArgs {
    const double *gin;
    double *gout;
    int n_lev, nh, n_phi;
    double omega_ma, omega_ic, r_cmb, r_icb;
    const LevelInfo *lev;
    const double *sinth, *costh;
    const double *means;  // [n_lev][2][nh]
};

__global__ void __launch_bounds__(DIAG_THREADS) dtb_product_kernel(DtbArgs a) {
    const int lev = blockIdx.y;
    const LevelInfo L = a.lev[lev];
    const size_t plane = (size_t)a.nh * a.n_phi;
    for (unsigned pt = blockIdx.x * blockDim.x + threadIdx.x; pt < (unsigned)plane; pt += gridDim.x * blockDim.x) {
        const int k = (int)(pt / (unsigned)a.n_phi);
        const double st = a.sinth[k], ct = a.costh[k];
        double n[6], s[6];
#pragma unroll
        for (int f = 0; f < 6; f++) {
            const double *base = a.gin + (((size_t)f * a.n_lev + lev) * 2) * plane + pt;
            const double e = __ldg(base), o = __ldg(base + plane);
            n[f] = e + o;
            s[f] = e - o;
        }
        const double me = a.means[((size_t)lev * 2 + 0) * a.nh + k], mo = a.means[((size_t)lev * 2 + 1) * a.nh + k];
        double vpn = me + mo, vps = me - mo;
        if (L.nBc == 1) { n[0] = 0.0; s[0] = 0.0; }
        if (L.nBc == 2) {  // v_rigid_boundary, nonlinear_bcs.f90:120-175
            const double r2 = (L.nR == 1) ? a.r_cmb * a.r_cmb : a.r_icb * a.r_icb;
            const double om = (L.nR == 1) ? a.omega_ma : a.omega_ic;
            n[0] = s[0] = 0.0; n[1] = s[1] = 0.0;
            n[2] = s[2] = vpn = vps = r2 * L.rho0 * (st * st) * om;
        }
        const double fac = 1.0 / (st * st), cot = ct / st / st / st, orho1 = L.orho1;
        const double vpASn = orho1 * vpn, vpASs = orho1 * vps;
        double pn[11], ps[11];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double *v = h == 0 ? n : s;
            double *p = h == 0 ? pn : ps;
            const double fc = h == 0 ? cot : -cot, vz = h == 0 ? vpASn : vpASs;
            const double vr = v[0], vt = v[1], vp = v[2], br = v[3], bt = v[4], bp = v[5];
            p[0] = orho1 * bt * vr;
            p[1] = orho1 * bp * vr;
            p[2] = orho1 * vt * br;
            p[3] = orho1 * vp * br;
            p[4] = fac * orho1 * bt * vp;
            p[5] = fac * orho1 * bp * vt;
            p[6] = fc * orho1 * (bp * vt + bt * vp);
            p[7] = fac * fac * orho1 * (bp * vt + bt * vp);
            p[8] = fac * br * vz;
            p[9] = fac * bt * vz;
            p[10] = fac * fac * bt * vz;
        }
#pragma unroll
        for (int q = 0; q < 11; q++) {
            double *base = a.gout + (((size_t)q * a.n_lev + lev) * 2) * plane + pt;
            base[0] = pn[q] + ps[q];
            base[plane] = pn[q] - ps[q];
        }
    }
}

}  // namespace magic

// ------------------------------------------------------------------------------------------------------
// Torsional-oscillation sums (rIter.f90:395-404): getTOnext's grid part (TO.f90:330-343) keeps Bs, Bp/s-normalised and Bz of the
// previous step for every level; getTO (TO.f90:141-307) forms, per level and colatitude, the azimuthal means of some twenty
// products of (vr, vt, vp, cvr, dvpdr, br, bt, bp, cbr, cbt, phi) and combines them into fifteen (r, theta) arrays.  One CTA per
// (colatitude pair, level): both hemispheres from one read of the E/O rows, fixed-shape reductions, bitwise repeatable.
namespace magic {

struct ToIn { int vr, vt, vp, cvr, dvpdr, br, bt, bp, cbr, cbt, phi; };
constexpr int TO_NF = 11;
constexpr int TO_NOUT = 15;      // = MAGIC_NTO
constexpr int TO_NSUM = 20;
constexpr int TO_THREADS = 128;

enum ToSlot { TO_V2AS = 0, TO_VAS, TO_DZCOR, TO_DZRSTR, TO_DZASTR, TO_DZLF, TO_BS2, TO_BSP, TO_BPZ, TO_BSZ, TO_BSPD, TO_BPSD, TO_BZPD, TO_BPZD,
              TO_DZPEN };

struct ToArgs {
    ToIn ti;
    const double *gin;
    int n_lev, nh, n_phi, n_theta;
    int l_mag, l_phase_field;
    double omega_ma, omega_ic, r_cmb, r_icb, CorFac, pen, o_dtLast;  // pen = 1 / (epsPhase penaltyFac)^2
    const LevelInfo *lev;
    const double *sinth, *costh;
    double *last;   // [n_lev][3 (Bs, Bp, Bz)][2 (north, south)][nh][n_phi] of the chunk's first level on
    double *out;    // [n_lev][TO_NOUT][n_theta], colatitudes ordered north -> south (n_theta_cal2ord)
};

// cylindrical components the time derivatives are built from (TO.f90:243-246, :332-337)
__device__ __forceinline__ void to_bsbpbz(const LevelInfo &L, double st, double ct, double br, double bt, double bp, double &bs, double &bpl,
                                          double &bz) {
    bs = st * L.or2 * br + ct / st * L.or1 * bt;
    bpl = L.or1 * bp / st;
    bz = ct * L.or2 * br - L.or1 * bt;
}

__global__ void __launch_bounds__(DIAG_THREADS) to_next_kernel(ToArgs a) {
    const int lev = blockIdx.y;
    const LevelInfo L = a.lev[lev];
    const size_t plane = (size_t)a.nh * a.n_phi;
    for (unsigned pt = blockIdx.x * blockDim.x + threadIdx.x; pt < (unsigned)plane; pt += gridDim.x * blockDim.x) {
        const int k = (int)(pt / (unsigned)a.n_phi);
        const double st = a.sinth[k], ct = a.costh[k];
        const int fi[3] = {a.ti.br, a.ti.bt, a.ti.bp};
        double e[3], o[3];
#pragma unroll
        for (int f = 0; f < 3; f++) {
            const double *base = a.gin + (((size_t)fi[f] * a.n_lev + lev) * 2) * plane + pt;
            e[f] = __ldg(base);
            o[f] = __ldg(base + plane);
        }
        double *dst = a.last + (size_t)lev * 6 * plane + pt;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            double bs, bp, bz;
            to_bsbpbz(L, st, h ? -ct : ct, h ? e[0] - o[0] : e[0] + o[0], h ? e[1] - o[1] : e[1] + o[1], h ? e[2] - o[2] : e[2] + o[2], bs, bp, bz);
            dst[(size_t)(0 * 2 + h) * plane] = bs;
            dst[(size_t)(1 * 2 + h) * plane] = bp;
            dst[(size_t)(2 * 2 + h) * plane] = bz;
        }
    }
}

__global__ void __launch_bounds__(TO_THREADS) to_kernel(ToArgs a) {
    const int k = blockIdx.x, lev = blockIdx.y;
    const LevelInfo L = a.lev[lev];
    const size_t plane = (size_t)a.nh * a.n_phi;
    const double st = a.sinth[k], ctn = a.costh[k];
    const double or1 = L.or1, or2 = L.or2, or3 = L.or1 * L.or2, or4 = L.or4, orho1 = L.orho1, beta = L.beta;
    double acc[2][TO_NSUM];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int i = 0; i < TO_NSUM; i++) acc[h][i] = 0.0;
    const int *fidx = &a.ti.vr;
    for (int j = threadIdx.x; j < a.n_phi; j += TO_THREADS) {
        const size_t pt = (size_t)k * a.n_phi + j;
        double e[TO_NF], o[TO_NF];
#pragma unroll
        for (int f = 0; f < TO_NF; f++) {
            e[f] = 0.0;
            o[f] = 0.0;
            if (fidx[f] >= 0) {
                const double *base = a.gin + (((size_t)fidx[f] * a.n_lev + lev) * 2) * plane + pt;
                e[f] = __ldg(base);
                o[f] = __ldg(base + plane);
            }
        }
        double lastv[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (a.l_mag) {
            const double *src = a.last + (size_t)lev * 6 * plane + pt;
#pragma unroll
            for (int q = 0; q < 6; q++) lastv[q] = __ldg(src + (size_t)q * plane);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double ct = h ? -ctn : ctn;
            double v[TO_NF];
#pragma unroll
            for (int f = 0; f < TO_NF; f++) v[f] = h ? e[f] - o[f] : e[f] + o[f];
            double vr = v[0], vt = v[1], vp = v[2], cvr = v[3];
            const double dvpdr = v[4], br = v[5], bt = v[6], bp = v[7], cbr = v[8], cbt = v[9], phi = v[10];
            diag_override_raw(L, a.omega_ma, a.omega_ic, a.r_cmb, a.r_icb, st, ct, vr, vt, vp, cvr);
            double *s = acc[h];
            s[0] += vr; s[1] += vt; s[2] += vp;
            s[3] += vr * vr; s[4] += vt * vt; s[5] += vp * vp;
            s[6] += orho1 * (dvpdr - beta * vp);
            s[7] += cvr;
            s[8] += orho1 * vr * (dvpdr - beta * vp);
            s[9] += orho1 * vt * cvr;
            if (a.l_phase_field) s[10] += phi * vp;
            if (a.l_mag) {
                const double os = 1.0 / st, os2 = os * os;
                const double Bs2F1 = st * st * or4, Bs2F2 = ct * ct * os2 * or2, Bs2F3 = 2.0 * ct * or3, BspF2 = ct * os2 * or2;
                const double BpzF1 = ct * os * or3, BpzF2 = or2 * os, BszF1 = st * ct * or4, BszF2 = (2.0 * ct * ct - 1.0) * os * or3,
                             BszF3 = ct * os * or2;
                s[11] += cbr * bt - cbt * br;
                s[12] += Bs2F1 * br * br + Bs2F2 * bt * bt + Bs2F3 * br * bt;
                s[13] += or3 * br * bp + BspF2 * bt * bp;
                s[14] += BpzF1 * br * bp - BpzF2 * bt * bp;
                s[15] += BszF1 * br * br + BszF2 * br * bt - BszF3 * bt * bt;
                double BsL, BpL, BzL;
                to_bsbpbz(L, st, ct, br, bt, bp, BsL, BpL, BzL);
                const double BsO = lastv[0 * 2 + h], BpO = lastv[1 * 2 + h], BzO = lastv[2 * 2 + h];
                s[16] += BsL * (BpL - BpO);
                s[17] += BpL * (BsL - BsO);
                s[18] += BzL * (BpL - BpO);
                s[19] += BpL * (BzL - BzO);
            }
        }
    }
    __shared__ double red[TO_THREADS / 32][2 * TO_NSUM];
    __shared__ double tot[2 * TO_NSUM];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int i = 0; i < TO_NSUM; i++) {
            double v = acc[h][i];
            for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][h * TO_NSUM + i] = v;
        }
    __syncthreads();
    if (threadIdx.x < 2 * TO_NSUM) {
        double v = 0.0;
        for (int w = 0; w < TO_THREADS / 32; w++) v += red[w][threadIdx.x];
        tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int h = threadIdx.x;
        const double *s = tot + h * TO_NSUM;
        const double ct = h ? -ctn : ctn, os = 1.0 / st, os2 = os * os, pn = 1.0 / (double)a.n_phi;
        const int th = h ? a.n_theta - 1 - k : k;   // ordered colatitude index
        double *o = a.out + (size_t)lev * TO_NOUT * a.n_theta + th;
        const size_t nt = (size_t)a.n_theta;
        const double VrMean = s[0], VtMean = s[1], dVpdrMean = s[6], cVrMean = s[7];
        o[TO_V2AS * nt] = pn * or4 * s[3] + pn * or2 * os2 * s[4] + pn * or2 * os2 * s[5];
        o[TO_VAS * nt] = orho1 * (pn * or1 * os * s[2]);
        o[TO_DZCOR * nt] = -pn * 2.0 * a.CorFac * (or2 * st * VrMean + or1 * ct * os * VtMean);
        o[TO_DZRSTR * nt] = -pn * or3 * os * (s[8] - pn * VrMean * dVpdrMean + s[9] - orho1 * pn * VtMean * cVrMean);
        o[TO_DZASTR * nt] = -pn * or3 * os * pn * (VrMean * dVpdrMean + orho1 * VtMean * cVrMean);
        o[TO_DZPEN * nt] = a.l_phase_field ? -pn * s[10] * or1 * os * a.pen : 0.0;
        o[TO_DZLF * nt] = a.l_mag ? pn * or3 * os * s[11] : 0.0;
        o[TO_BS2 * nt] = pn * s[12];
        o[TO_BSP * nt] = pn * s[13];
        o[TO_BPZ * nt] = pn * s[14];
        o[TO_BSZ * nt] = pn * s[15];
        o[TO_BSPD * nt] = pn * (s[16] * a.o_dtLast);
        o[TO_BPSD * nt] = pn * (s[17] * a.o_dtLast);
        o[TO_BZPD * nt] = pn * (s[18] * a.o_dtLast);
        o[TO_BPZD * nt] = pn * (s[19] * a.o_dtLast);
    }
}

}  // namespace magic

// ------------------------------------------------------------------------------------------------------
// R.m.s. force balance inside the radial loop on lRmsCalc steps (rIter.f90:215-252, 710): get_nl with every level treated as bulk
// (get_nl.f90:242-310), get_nl_RMS (RMS.f90:469-560) and the merge of transform_to_lm_space (rIter.f90:650-667), fused into one
// pass over the synthesised grid fields.  Writes the fourteen grid products transform_to_lm_RMS analyses, as (N+S, N-S) rows:
//   0 Advr (merged)  1 LFr  2 dtVr  3 dpkindr  |  4,5 Advt2, Advp2  6,7 LFt2, LFp2  8,9 CFt2, CFp2  10,11 PFt, PFp  12,13 dtVt, dtVp
namespace magic {

struct RmsIn { int vr, vt, vp, dvrdr, dvtdr, dvpdr, cvr, cvt, cvp, dvrdt, dvrdp, dvtdp, dvpdp, br, bt, bp, cbr, cbt, cbp, dpdt, dpdp, vro, vto, vpo, s, phi; };
constexpr int RMS_NF = 26;
constexpr int RMS_NOUT = 14;   // = MAGIC_NRMS

struct RmsArgs {
    RmsIn ri;
    const double *gin;
    double *gout;
    int n_lev, nh, n_phi;
    int l_conv_nl, l_mag_LF, l_mag_nl, l_adv_curl, n_r_LCR, l_phase_field, l_precession, l_centrifuge, minc;
    double LFfac, CorFac, o_dt;
    double pen, posnalp, oek_time, cafac;   // 1 / (epsPhase penaltyFac)^2;  -2 oek po sin(prec_angle);  oek * time;  dilution_fac ra opr
    const LevelInfo *lev;
    const double *sinth, *costh;
};

__global__ void __launch_bounds__(DIAG_THREADS) rms_kernel(RmsArgs a) {
    const int lev = blockIdx.y;
    const LevelInfo L = a.lev[lev];
    const size_t plane = (size_t)a.nh * a.n_phi;
    const double r = L.r, or1 = L.or1, or2 = L.or2, or3 = L.or1 * L.or2, or4 = L.or4, orho1 = L.orho1, beta = L.beta;
    const bool lf = a.l_mag_LF && L.nR > a.n_r_LCR;
    const int *fidx = &a.ri.vr;
    for (unsigned pt = blockIdx.x * blockDim.x + threadIdx.x; pt < (unsigned)plane; pt += gridDim.x * blockDim.x) {
        const int k = (int)(pt / (unsigned)a.n_phi), j = (int)(pt - (unsigned)k * (unsigned)a.n_phi);
        const double st = a.sinth[k], ctn = a.costh[k], os = 1.0 / st, os2 = os * os;
        double cph = 0.0, sph = 0.0;
        if (a.l_precession) {  // get_nl.f90:346-357: phase oek * time + longitude
            const double ph = a.oek_time + (double)j * (6.283185307179586476925286766559 / (double)(a.n_phi * a.minc));
            cph = cos(ph);
            sph = sin(ph);
        }
        double e[RMS_NF], o[RMS_NF];
#pragma unroll
        for (int f = 0; f < RMS_NF; f++) {
            e[f] = 0.0;
            o[f] = 0.0;
            if (fidx[f] >= 0) {
                const double *base = a.gin + (((size_t)fidx[f] * a.n_lev + lev) * 2) * plane + pt;
                e[f] = __ldg(base);
                o[f] = __ldg(base + plane);
            }
        }
        double res[2][RMS_NOUT];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double ct = h ? -ctn : ctn, cn2 = ct * os2;
            double v[RMS_NF];
#pragma unroll
            for (int f = 0; f < RMS_NF; f++) v[f] = h ? e[f] - o[f] : e[f] + o[f];
            const double vr = v[0], vt = v[1], vp = v[2], dvrdr = v[3], dvtdr = v[4], dvpdr = v[5], cvr = v[6], cvt = v[7], cvp = v[8], dvrdt = v[9],
                         dvrdp = v[10], dvtdp = v[11] * os2, dvpdp = v[12] * os2,  // torpol_to_dphspat post-scaling, sht_native.f90:263-270
                         br = v[13], bt = v[14], bp = v[15], cbr = v[16], cbt = v[17], cbp = v[18];
            double LFr = 0.0, LFt = 0.0, LFp = 0.0, Ar = 0.0, At = 0.0, Ap = 0.0;
            if (lf) {
                LFr = a.LFfac * os2 * (cbt * bp - cbp * bt);
                LFt = a.LFfac * or4 * (cbp * br - cbr * bp);
                LFp = a.LFfac * or4 * (cbr * bt - cbt * br);
            }
            if (a.l_conv_nl) {
                if (a.l_adv_curl) {
                    Ar = -os2 * (cvt * vp - cvp * vt);
                    At = -or4 * (cvp * vr - cvr * vp);
                    Ap = -or4 * (cvr * vt - cvt * vr);
                } else {
                    Ar = -or2 * orho1 * (vr * (dvrdr - (2.0 * or1 + beta) * vr) + os2 * (vt * (dvrdt - r * vt) + vp * (dvrdp - r * vp)));
                    At = or4 * orho1 * (-vr * (dvtdr - beta * vt) + vt * (cn2 * vt + dvpdp + dvrdr) + vp * (cn2 * vp - dvtdp));
                    Ap = or4 * orho1 * (-vr * (dvpdr - beta * vp) - vt * (dvtdp + cvr) - vp * dvpdp);
                }
            }
            if (a.l_phase_field) {  // get_nl.f90:333-339: the penalty is part of Advr, Advt, Advp before get_nl_RMS reads them
                Ar -= v[25] * vr * a.pen;
                At -= or2 * v[25] * vt * a.pen;
                Ap -= or2 * v[25] * vp * a.pen;
            }
            double *q = res[h];
            double PFt = v[19] * or1, PFp = v[20] * or1;
            q[8] = -2.0 * a.CorFac * ct * vp * or1;
            q[9] = 2.0 * a.CorFac * st * (or1 * ct * os * vt + or2 * st * vr);
            double At2 = a.l_conv_nl ? r * At : 0.0, Ap2 = a.l_conv_nl ? r * Ap : 0.0;
            q[6] = (lf && a.l_mag_nl) ? r * LFt : 0.0;
            q[7] = (lf && a.l_mag_nl) ? r * LFp : 0.0;
            q[3] = 0.0;
            if (a.l_adv_curl) {
                const double X = or3 * (or2 * vr * dvrdt - vt * (dvrdr + dvpdp + cn2 * vt) + vp * (cvr + dvtdp - cn2 * vp));
                const double Y = or3 * (or2 * vr * dvrdp + vt * dvtdp + vp * dvpdp);
                PFt -= X;
                PFp -= Y;
                if (a.l_conv_nl) { At2 -= X; Ap2 -= Y; }
                q[3] = or4 * vr * (dvrdr - 2.0 * or1 * vr) + or2 * os2 * (vt * (dvtdr - or1 * vt) + vp * (dvpdr - or1 * vp));
            }
            q[4] = At2; q[5] = Ap2; q[10] = PFt; q[11] = PFp;
            q[2] = a.o_dt * or2 * (vr - v[21]);
            q[12] = a.o_dt * or1 * (vt - v[22]);
            q[13] = a.o_dt * or1 * (vp - v[23]);
            if (a.l_conv_nl && a.l_mag_LF) { if (lf) Ar += LFr; }   // rIter.f90:650-667
            else if (a.l_mag_LF) Ar = lf ? LFr : 0.0;
            if (a.l_precession) Ar += a.posnalp * os * r * (cph * vp * ct + sph * vt);       // PCr, rIter.f90:669-673
            if (a.l_centrifuge) Ar += -a.cafac * r * (st * st * st * st) * v[24];            // CAr, rIter.f90:675-678
            q[0] = Ar;
            q[1] = LFr;
        }
#pragma unroll
        for (int q = 0; q < RMS_NOUT; q++) {
            double *base = a.gout + (((size_t)q * a.n_lev + lev) * 2) * plane + pt;
            base[0] = res[0][q] + res[1][q];
            base[plane] = res[0][q] - res[1][q];
        }
    }
}

}  // namespace magic
