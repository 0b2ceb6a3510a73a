// kernels_grid.cuh -- grid-space kernels: fused get_nl (+ Adv+=LF of rIter.f90:646-667 + Courant maxima of
// courant.f90:209-275) and the layout converters of the per-call API.
//
// Grid fields live as g[field][lev][s][k][phi] (phi fastest): s=0 holds E, s=1 holds O with
// north(k) = E+O, south(k) = E-O.  get_nl is point-wise, so one thread rebuilds both hemispheres' values
// of a (k,phi) pair, forms the products for both, and stores their sum and difference -- exactly the
// f1ES/f1EA combinations the analysis starts from (shtransforms.f90:704-710).
#pragma once
#include "common.cuh"

namespace magic {

// index of each synthesised field inside the synthesis grid array (-1 = absent)
struct GridIn { int vr, vt, vp, cvr, cvt, cvp, s, br, bt, bp, cbr, cbt, cbp, xi, dvrdr, dvtdr, dvpdr, dvrdt, dvrdp, dvtdp, dvpdp, phi; };  // same order as PointIn
// index of each product inside the product grid array (-1 = absent)
struct GridOut { int Advr, Advt, Advp, VSr, VSt, VSp, VxBr, VxBt, VxBp, VXir, VXit, VXip, heat, phiTerms; };

struct NlFlags {
    int l_conv_nl, l_heat_nl, l_mag_nl, l_mag_LF, l_mag, l_mag_kin, l_adv_curl, l_anel, l_chemical_conv, l_precession,
        l_centrifuge, l_cour_alf_damp, l_full_sphere, n_r_LCR, l_phase_field;
    double LFfac, opm, ViscHeatFac, OhmLossFac, oek, po, prec_angle, dilution_fac, ra, opr, omega_ma, omega_ic, r_cmb, r_icb,
        courfac, alffac, time, epsPhase, phaseDiffFac, penaltyFac, tmelt;
};

struct NlArgs {
    NlFlags f;
    GridIn gi;
    GridOut go;
    const double *gin;
    double *gout;
    int n_lev, nh, n_phi, minc;
    const LevelInfo *lev;
    const double *sinth, *costh;  // northern values, [nh]
    unsigned long long *courmax;  // [n_lev][2] bit patterns of the (non-negative) maxima vr2max, vh2max
    const double *ddw, *ddb;      // spectral inputs of the chunk (complex [n_lev][lm_max]) for v_center_sphere, or null
    int lm10, lm11, lm_max;       // st_map indices of (l=1,m=0), (l=1,m=1) (-1 when minc /= 1)
    const double *wgauss;         // 2 pi gauss(k) / n_phi
    double *tq_partial;           // [n_lev][gridDim.x] per-CTA partial sums of gauss * Br * Bp (get_lorentz_torque)
};

__device__ __forceinline__ void atomic_max_pos(unsigned long long *addr, double v) {
    atomicMax(addr, (unsigned long long)__double_as_longlong(v));  // valid for v >= 0
}

struct PointIn { double vr, vt, vp, cvr, cvt, cvp, s, br, bt, bp, cbr, cbt, cbp, xi, dvrdr, dvtdr, dvpdr, dvrdt, dvrdp, dvtdp, dvpdp, phi; };
struct PointOut { double Advr, Advt, Advp, VSr, VSt, VSp, VxBr, VxBt, VxBp, VXir, VXit, VXip, heat, phiTerms; };

// get_nl.f90:213-441 at one grid point.  ct/cn2 carry the hemisphere sign.
// MAG: magnetic fields present.  EXTRA: anything beyond the Boussinesq curl-form set (u.grad u advection, anelastic
// heating, composition, precession, centrifuge); when false those branches and their loads are compiled out.
template <bool MAG, bool EXTRA>
__device__ __forceinline__ void nl_point(const NlFlags &F, const LevelInfo &L, const PointIn &p, double st, double ct, double os2,
                                         double cn2, double phi, PointOut &o) {
    const int nBc = L.nBc, nR = L.nR;
    const double or1 = L.or1, or2 = L.or2, or4 = L.or4, orho1 = L.orho1, beta = L.beta, r = L.r;
    double LFr = 0, LFt = 0, LFp = 0;
    const bool lf_on = MAG && F.l_mag_LF && nBc == 0 && nR > F.n_r_LCR;
    if (lf_on) {
        LFr = F.LFfac * os2 * (p.cbt * p.bp - p.cbp * p.bt);
        LFt = F.LFfac * or4 * (p.cbp * p.br - p.cbr * p.bp);
        LFp = F.LFfac * or4 * (p.cbr * p.bt - p.cbt * p.br);
    }
    double Ar = 0, At = 0, Ap = 0;
    if (F.l_conv_nl && nBc == 0) {
        if (!EXTRA || F.l_adv_curl) {
            Ar = -os2 * (p.cvt * p.vp - p.cvp * p.vt);
            At = -or4 * (p.cvp * p.vr - p.cvr * p.vp);
            Ap = -or4 * (p.cvr * p.vt - p.cvt * p.vr);
        } else {
            Ar = -or2 * orho1 * (p.vr * (p.dvrdr - (2.0 * or1 + beta) * p.vr) +
                                 os2 * (p.vt * (p.dvrdt - r * p.vt) + p.vp * (p.dvrdp - r * p.vp)));
            At = or4 * orho1 * (-p.vr * (p.dvtdr - beta * p.vt) + p.vt * (cn2 * p.vt + p.dvpdp + p.dvrdr) + p.vp * (cn2 * p.vp - p.dvtdp));
            Ap = or4 * orho1 * (-p.vr * (p.dvpdr - beta * p.vp) - p.vt * (p.dvtdp + p.cvr) - p.vp * p.dvpdp);
        }
    }
    o.phiTerms = 0;
    if (EXTRA && F.l_phase_field && nBc == 0) {  // get_nl.f90:333-344: penalty on the velocity in the solid, phase-field source
        const double pen = 1.0 / F.epsPhase / F.epsPhase / F.penaltyFac / F.penaltyFac;
        Ar -= p.phi * p.vr * pen;
        At -= or2 * p.phi * p.vt * pen;
        Ap -= or2 * p.phi * p.vp * pen;
        o.phiTerms = -1.0 / (F.epsPhase * F.epsPhase) * p.phi * (1.0 - p.phi) * (F.phaseDiffFac * (1.0 - 2.0 * p.phi) + p.s - F.tmelt);
    }
    // rIter.f90:646-667
    if (F.l_conv_nl && F.l_mag_LF) {
        if (nR > F.n_r_LCR) { Ar += LFr; At += LFt; Ap += LFp; }
    } else if (F.l_mag_LF) {
        if (nR > F.n_r_LCR) { Ar = LFr; At = LFt; Ap = LFp; }
        else { Ar = 0; At = 0; Ap = 0; }
    }
    if (EXTRA && F.l_precession && nBc == 0) {
        const double posnalp = -2.0 * F.oek * F.po * sin(F.prec_angle);
        const double ph = F.oek * F.time + phi;
        Ar += posnalp * (1.0 / st) * r * (cos(ph) * p.vp * ct + sin(ph) * p.vt);
        At += -posnalp * st * or2 * (cos(ph) * p.vp + sin(ph) * or1 * p.vr);
        Ap += posnalp * st * cos(ph) * or2 * (p.vt - or1 * p.vr * ct);
    }
    if (EXTRA && F.l_centrifuge && nBc == 0) {
        Ar += -F.dilution_fac * r * (st * st * st * st) * F.ra * F.opr * p.s;
        At += -F.dilution_fac * r * (st * st * st) * ct * F.ra * F.opr * p.s;
    }
    o.Advr = Ar; o.Advt = At; o.Advp = Ap;
    o.VSr = o.VSt = o.VSp = 0;
    if (F.l_heat_nl && nBc == 0) {
        o.VSr = p.vr * p.s;
        o.VSt = or2 * p.vt * p.s;
        o.VSp = or2 * p.vp * p.s;
    }
    o.VXir = o.VXit = o.VXip = 0;
    if (EXTRA && F.l_chemical_conv && nBc == 0) {
        o.VXir = p.vr * p.xi;
        o.VXit = or2 * p.vt * p.xi;
        o.VXip = or2 * p.vp * p.xi;
    }
    o.VxBr = o.VxBt = o.VxBp = 0;
    if (MAG && F.l_mag_nl) {
        if (nBc == 0 && nR > F.n_r_LCR) {
            o.VxBr = orho1 * os2 * (p.vt * p.bp - p.vp * p.bt);
            o.VxBt = orho1 * or4 * (p.vp * p.br - p.vr * p.bp);
            o.VxBp = orho1 * or4 * (p.vr * p.bt - p.vt * p.br);
        } else if (nBc == 1 || nR <= F.n_r_LCR) {
            o.VxBt = or4 * orho1 * p.vp * p.br;
            o.VxBp = -or4 * orho1 * p.vt * p.br;
        } else if (nBc == 2) {
            o.VxBt = or4 * orho1 * p.vp * p.br;
            o.VxBp = 0.0;
        }
    }
    o.heat = 0;
    if (EXTRA && F.l_anel && nBc == 0) {
        double t1 = p.dvrdr - (2.0 * or1 + beta) * p.vr;
        double t2 = cn2 * p.vt + p.dvpdp + p.dvrdr - or1 * p.vr;
        double t3 = p.dvpdp + cn2 * p.vt + or1 * p.vr;
        double t6 = 2.0 * p.dvtdp + p.cvr - 2.0 * cn2 * p.vp;
        double t4 = r * p.dvtdr - (2.0 + beta * r) * p.vt + or1 * p.dvrdt;
        double t5 = r * p.dvpdr - (2.0 + beta * r) * p.vp + or1 * p.dvrdp;
        double t7 = beta * p.vr;
        o.heat = F.ViscHeatFac * or4 * orho1 * L.otemp1 * L.visc *
                 (2.0 * t1 * t1 + 2.0 * t2 * t2 + 2.0 * t3 * t3 + t6 * t6 + os2 * (t4 * t4 + t5 * t5) - 2.0 * (1.0 / 3.0) * t7 * t7);
        if (F.l_mag_nl && nR > F.n_r_LCR)
            o.heat += F.OhmLossFac * or2 * L.otemp1 * L.lambda * (or2 * p.cbr * p.cbr + os2 * p.cbt * p.cbt + os2 * p.cbp * p.cbp);
    }
}

constexpr int NL_THREADS = 256;

template <bool MAG, bool EXTRA>
__global__ void __launch_bounds__(NL_THREADS, EXTRA ? 1 : 2) get_nl_kernel(NlArgs a) {
    const int lev = blockIdx.y;
    const LevelInfo L = a.lev[lev];
    const NlFlags &F = a.f;
    const size_t plane = (size_t)a.nh * a.n_phi;  // one (field,lev,s) slab
    double vr2max = 0.0, vh2max = 0.0, tq = 0.0;
    double valri2 = 0.0, valhi2 = 0.0;
    if (F.l_cour_alf_damp) {
        double h = 0.5 * (1.0 + F.opm);
        valri2 = h * h / L.delxr2;
        valhi2 = h * h / L.delxh2;
    }
    const double cf2 = F.courfac * F.courfac, af2 = F.alffac * F.alffac;
    for (unsigned pt = blockIdx.x * blockDim.x + threadIdx.x; pt < (unsigned)plane; pt += gridDim.x * blockDim.x) {
        const int k = (int)(pt / (unsigned)a.n_phi), j = (int)(pt - (unsigned)k * (unsigned)a.n_phi);
        const double st = a.sinth[k], ct = a.costh[k];
        const double os2 = 1.0 / (st * st), cn2 = ct / st / st;
        const double phi = (double)j * (6.283185307179586476925286766559 / (double)(a.n_phi * a.minc));
        // Phase 1: issue every load of this point pair before the first use (the warp scheduler is in-order, so a
        // load->use->load sequence would leave only two requests in flight per warp).  Slots follow PointIn.
        const int *fidx = &a.gi.vr;  // GridIn members are declared in PointIn order
        double re[22], ro[22];
#pragma unroll
        for (int f = 0; f < 22; f++) {
            const bool used = f < 7 || (f < 13 ? MAG : EXTRA);
            re[f] = 0.0;
            ro[f] = 0.0;
            if (used && fidx[f] >= 0) {
                const double *base = a.gin + (((size_t)fidx[f] * a.n_lev + lev) * 2) * plane + pt;
                re[f] = __ldg(base);
                ro[f] = __ldg(base + plane);
            }
        }
        PointIn pn, ps;
        double *pnv = &pn.vr, *psv = &ps.vr;
#pragma unroll
        for (int f = 0; f < 22; f++) {
            pnv[f] = re[f] + ro[f];
            psv[f] = re[f] - ro[f];
        }
        if (EXTRA) {  // torpol_to_dphspat post-scaling by 1/sin^2 (sht_native.f90:263-270)
            pn.dvtdp *= os2; ps.dvtdp *= os2; pn.dvpdp *= os2; ps.dvpdp *= os2;
        }
        // boundary overrides of transform_to_grid_space (rIter.f90:555-602)
        if (L.nBc == 1) { pn.vr = 0.0; ps.vr = 0.0; }
        if (L.nBc == 2) {  // v_rigid_boundary, nonlinear_bcs.f90:120-175 (l_vr_cmb/icb = .false.)
            const double r2 = (L.nR == 1) ? F.r_cmb * F.r_cmb : F.r_icb * F.r_icb;
            const double om = (L.nR == 1) ? F.omega_ma : F.omega_ic;
            pn.vr = ps.vr = 0.0; pn.vt = ps.vt = 0.0;
            pn.vp = ps.vp = r2 * L.rho0 * (st * st) * om;
        }
        if (L.center) {  // v_center_sphere, nonlinear_bcs.f90:177-224 (full sphere, r=0)
            const double y10 = 0.48860251190291992158638462283835, y11 = 0.34549414947133547925878907835150;
            const double cph = cos(phi), sph = sin(phi);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const double *dd = q == 0 ? a.ddw : a.ddb;
                if (dd == nullptr || (q == 1 && !MAG)) continue;
                const double d10 = dd[2 * ((size_t)lev * a.lm_max + a.lm10)];
                double c = 0.0, sn = 0.0;
                if (a.lm11 >= 0) {
                    const double re = dd[2 * ((size_t)lev * a.lm_max + a.lm11)], im = dd[2 * ((size_t)lev * a.lm_max + a.lm11) + 1];
                    c = re * cph - im * sph;
                    sn = re * sph + im * cph;
                }
                const double vrn = y10 * d10 * ct + 2.0 * y11 * st * c, vrs = -y10 * d10 * ct + 2.0 * y11 * st * c;
                const double vtn = st * (-y10 * d10 * st + 2.0 * y11 * ct * c), vts = st * (-y10 * d10 * st - 2.0 * y11 * ct * c);
                const double vpb = -2.0 * y11 * st * sn;
                if (q == 0) { pn.vr = vrn; ps.vr = vrs; pn.vt = vtn; ps.vt = vts; pn.vp = vpb; ps.vp = vpb; }
                else { pn.br = vrn; ps.br = vrs; pn.bt = vtn; ps.bt = vts; pn.bp = vpb; ps.bp = vpb; }
            }
        }
        if (MAG && L.torque) tq += a.wgauss[k] * (pn.br * pn.bp + ps.br * ps.bp);  // outRot.f90:477-478
        PointOut on, os;
        if (L.nl_on) {
            nl_point<MAG, EXTRA>(F, L, pn, st, ct, os2, cn2, phi, on);
            nl_point<MAG, EXTRA>(F, L, ps, st, -ct, os2, -cn2, phi, os);
        } else {
            on = PointOut{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
            os = on;
        }
        auto stf = [&](int fidx, double n, double s) {
            if (fidx < 0) return;
            double *base = a.gout + (((size_t)fidx * a.n_lev + lev) * 2) * plane + pt;
            base[0] = n + s;
            base[plane] = n - s;
        };
        stf(a.go.Advr, on.Advr, os.Advr); stf(a.go.Advt, on.Advt, os.Advt); stf(a.go.Advp, on.Advp, os.Advp);
        stf(a.go.VSr, on.VSr, os.VSr); stf(a.go.VSt, on.VSt, os.VSt); stf(a.go.VSp, on.VSp, os.VSp);
        if (MAG) { stf(a.go.VxBr, on.VxBr, os.VxBr); stf(a.go.VxBt, on.VxBt, os.VxBt); stf(a.go.VxBp, on.VxBp, os.VxBp); }
        if (EXTRA) {
            stf(a.go.VXir, on.VXir, os.VXir); stf(a.go.VXit, on.VXit, os.VXit); stf(a.go.VXip, on.VXip, os.VXip);
            stf(a.go.heat, on.heat, os.heat);
            stf(a.go.phiTerms, on.phiTerms, os.phiTerms);
        }
        // courant.f90:209-275 (XSH_COURANT == 0)
        if (L.cour_on) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const PointIn &p = h == 0 ? pn : ps;
                double vflr2 = L.orho2 * p.vr * p.vr;
                double vflh2 = (p.vt * p.vt + p.vp * p.vp) * os2 * L.orho2;
                if (MAG && F.l_mag && F.l_mag_LF && !F.l_mag_kin) {
                    double valr = p.br * p.br * F.LFfac * L.orho1;
                    double valr2 = valr * valr / (valr + valri2);
                    if (!(valr + valri2 > 0.0)) valr2 = 0.0;
                    vr2max = fmax(vr2max, L.or4 * (cf2 * vflr2 + af2 * valr2));
                    double valh2 = (p.bt * p.bt + p.bp * p.bp) * F.LFfac * os2 * L.orho1;
                    double valh2m = valh2 * valh2 / (valh2 + valhi2);
                    if (!(valh2 + valhi2 > 0.0)) valh2m = 0.0;
                    vh2max = fmax(vh2max, L.or2 * (cf2 * vflh2 + af2 * valh2m));
                } else {
                    vr2max = fmax(vr2max, cf2 * L.or4 * vflr2);
                    vh2max = fmax(vh2max, cf2 * L.or2 * vflh2);
                }
            }
        }
    }
    if (L.cour_on || (MAG && L.torque)) {
        // fixed-shape reductions (xor-shuffle tree, then warps in order): no floating-point atomics, bitwise reproducible
        __shared__ double red[3][NL_THREADS / 32];
        for (int o = 16; o > 0; o >>= 1) {
            vr2max = fmax(vr2max, __shfl_xor_sync(0xffffffffu, vr2max, o));
            vh2max = fmax(vh2max, __shfl_xor_sync(0xffffffffu, vh2max, o));
            tq += __shfl_xor_sync(0xffffffffu, tq, o);
        }
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = vr2max; red[1][threadIdx.x >> 5] = vh2max; red[2][threadIdx.x >> 5] = tq; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < NL_THREADS / 32; w++) { vr2max = fmax(vr2max, red[0][w]); vh2max = fmax(vh2max, red[1][w]); tq += red[2][w]; }
            if (L.cour_on) {
                atomic_max_pos(a.courmax + 2 * lev, vr2max);
                atomic_max_pos(a.courmax + 2 * lev + 1, vh2max);
            }
            if (MAG && L.torque) a.tq_partial[(size_t)lev * gridDim.x + blockIdx.x] = tq;
        }
    }
}

inline void launch_get_nl(const NlArgs &a, bool mag, bool extra, int gx, int n_lev, cudaStream_t st) {
    dim3 g(gx, n_lev);
    if (extra) {
        if (mag) get_nl_kernel<true, true><<<g, NL_THREADS, 0, st>>>(a);
        else get_nl_kernel<false, true><<<g, NL_THREADS, 0, st>>>(a);
    } else {
        if (mag) get_nl_kernel<true, false><<<g, NL_THREADS, 0, st>>>(a);
        else get_nl_kernel<false, false><<<g, NL_THREADS, 0, st>>>(a);
    }
}

// dtrkc/dthkc from the maxima (courant.f90:272-273, rIter.f90:215-216); Lorentz torques from the per-CTA partial sums,
// added in CTA order (outRot.f90:423-483; the mantle torque changes sign, rIter.f90:461)
__global__ void courant_finish_kernel(const unsigned long long *courmax, const LevelInfo *lev, int n_lev, double *dtrkc, double *dthkc,
                                      const double *tq_partial, int n_part, double LFfac, double *torque /* [0]=ic, [1]=ma */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lev) return;
    double vr2max = __longlong_as_double((long long)courmax[2 * i]), vh2max = __longlong_as_double((long long)courmax[2 * i + 1]);
    double a = 1e10, b = 1e10;
    if (lev[i].cour_on) {
        if (vr2max != 0.0) a = fmin(a, sqrt(lev[i].delxr2 / vr2max));
        if (vh2max != 0.0) b = fmin(b, sqrt(lev[i].delxh2 / vh2max));
    }
    dtrkc[i] = a;
    dthkc[i] = b;
    if (lev[i].torque && tq_partial != nullptr) {
        double sum = 0.0;
        for (int j = 0; j < n_part; j++) sum += tq_partial[(size_t)i * n_part + j];
        if (lev[i].torque == 1) torque[0] = LFfac * sum;
        else torque[1] = -(LFfac * sum);
    }
}

// ------------------------------------------------------------------------------------------------------
// Layout converters for the per-call `module sht` API: reference layout f(nlat_padded, n_phi), theta fastest,
// rows N/S interleaved  <->  internal g[s][k][phi].  32x32 shared-memory transposes.
__global__ void grid_export_kernel(const double *__restrict__ g /*[2][nh][nphi]*/, double *__restrict__ f, int nh, int n_phi,
                                   int nlat_padded, const double *__restrict__ kscale /* per-k factor or null */) {
    __shared__ double tn[32][33], ts[32][33];
    int k0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, j = j0 + threadIdx.x;
        if (k < nh && j < n_phi) {
            double e = g[(size_t)k * n_phi + j], o = g[((size_t)nh + k) * n_phi + j];
            double sc = kscale ? kscale[k] : 1.0;
            tn[r][threadIdx.x] = (e + o) * sc;
            ts[r][threadIdx.x] = (e - o) * sc;
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int j = j0 + r, k = k0 + threadIdx.x;
        if (k < nh && j < n_phi) {
            f[(size_t)j * nlat_padded + 2 * k] = tn[threadIdx.x][r];
            f[(size_t)j * nlat_padded + 2 * k + 1] = ts[threadIdx.x][r];
        }
    }
}

// get_br_v_bcs products (nonlinear_bcs.f90:58-66) on one boundary level: gin holds br, vt, vp (fields 0,1,2 of a single-level
// synthesis, E/O layout); gout receives br_vt = fac br vt and br_vp = br (fac vp - omega sin^2 theta) as the (N+S, N-S) rows
// the r2c FFT expects, fields 1 and 2; field 0 (scalar analysis column, unused) is zeroed.
__global__ void br_v_product_kernel(const double *__restrict__ gin, double *__restrict__ gout, int nh, int n_phi,
                                    const double *__restrict__ sinth, double fac, double omega) {
    const size_t plane = (size_t)nh * n_phi;
    for (size_t pt = blockIdx.x * (size_t)blockDim.x + threadIdx.x; pt < plane; pt += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(pt / n_phi);
        const double st2 = sinth[k] * sinth[k];
        const double bre = gin[pt], bro = gin[plane + pt], vte = gin[2 * plane + pt], vto = gin[3 * plane + pt];
        const double vpe = gin[4 * plane + pt], vpo = gin[5 * plane + pt];
        const double brn = bre + bro, brs = bre - bro, vtn = vte + vto, vts = vte - vto, vpn = vpe + vpo, vps = vpe - vpo;
        const double tn = fac * brn * vtn, ts = fac * brs * vts;
        const double pn = brn * (fac * vpn - omega * st2), ps = brs * (fac * vps - omega * st2);
        gout[pt] = 0.0;
        gout[plane + pt] = 0.0;
        gout[2 * plane + pt] = tn + ts;
        gout[3 * plane + pt] = tn - ts;
        gout[4 * plane + pt] = pn + ps;
        gout[5 * plane + pt] = pn - ps;
    }
}

__global__ void grid_import_kernel(const double *__restrict__ f, double *__restrict__ g, int nh, int n_phi, int nlat_padded) {
    __shared__ double te[32][33], to[32][33];
    int k0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int j = j0 + r, k = k0 + threadIdx.x;
        if (k < nh && j < n_phi) {
            double n = f[(size_t)j * nlat_padded + 2 * k], s = f[(size_t)j * nlat_padded + 2 * k + 1];
            te[threadIdx.x][r] = n + s;
            to[threadIdx.x][r] = n - s;
        }
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int k = k0 + r, j = j0 + threadIdx.x;
        if (k < nh && j < n_phi) {
            g[(size_t)k * n_phi + j] = te[r][threadIdx.x];
            g[((size_t)nh + k) * n_phi + j] = to[r][threadIdx.x];
        }
    }
}

}  // namespace magic
