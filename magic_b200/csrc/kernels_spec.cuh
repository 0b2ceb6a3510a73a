// kernels_spec.cuh -- spectral-side kernels: Legendre table build, synthesis operand assembly (the
// sht_native.f90 wrappers), analysis extraction and the get_td epilogue.
#pragma once
#include "common.cuh"

namespace magic {

// ------------------------------------------------------------------------------------------------------
// Table build: plm_theta (plms.f90:14-189, norm=2) evaluated per (order mc, northern colatitude k) for the degrees
// l = m .. l_max+1 (the reference runs the same recurrence to l_max+1 for its derivative table), written in the blocked
// layout of common.cuh.  off[mc*2 + {0,1}] = offsets (doubles) of P_even, P_odd.
__global__ void build_tables_kernel(double *__restrict__ tab, const long long *__restrict__ off,
                                    const double *__restrict__ sinth, const double *__restrict__ costh,
                                    const double *__restrict__ pmm_fac, int nh, int NHP, int l_max, int minc, int n_m) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    int mc = blockIdx.y;
    if (k >= nh) return;
    const int m = mc * minc;
    const double dnorm = 0.28209479177387814347403972578039;  // 1/sqrt(4 pi)
    const double st = sinth[k], ct = costh[k];
    double *Pe = tab + off[mc * 2 + 0], *Po = tab + off[mc * 2 + 1];
    double plm = pmm_fac[mc];
    if (st != 0.0) plm = plm * pow(st, (double)m);
    else if (m != 0) plm = 0.0;
    double plm1 = 0.0, plm2;
    for (int l = m; l <= l_max + 1; l++) {
        const int j = (l - m) >> 1;
        if (((l - m) & 1) == 0) Pe[(size_t)j * NHP + k] = dnorm * plm;
        else Po[(size_t)j * NHP + k] = dnorm * plm;
        // advance the recurrence to degree l+1 (plms.f90:79-113)
        const int ln = l + 1;
        plm2 = plm1;
        plm1 = plm;
        plm = ct * sqrt((double)((2 * ln - 1) * (2 * ln + 1)) / (double)((ln - m) * (ln + m))) * plm1 -
              sqrt(((double)(2 * ln + 1) * (double)(ln + m - 1) * (double)(ln - m - 1)) /
                   ((double)(2 * ln - 3) * (double)(ln - m) * (double)(ln + m))) * plm2;
    }
}

// First colatitude index k (north pole = 0) at which any P entry of order mc (degrees m .. l_max+1) reaches `thr` in
// magnitude.  Rows below it contribute less than thr * l * |coefficient| to any sum -- with thr = 1e-40 that is more than 20
// orders of magnitude below FP64 rounding of the result -- and are skipped by the Legendre GEMMs ("polar optimisation", cf.
// the eps_polar of shtns.f90:59, here with a threshold that cannot change a single result bit that matters).
__global__ void table_kmin_kernel(const double *__restrict__ tab, const long long *__restrict__ off, int nh, int NHP, int l_max,
                                  int minc, double thr, int *__restrict__ kmin) {
    __shared__ int best;
    const int mc = blockIdx.x, m = mc * minc;
    if (threadIdx.x == 0) best = nh;
    __syncthreads();
    const long long rows = (long long)(l_max + 1 - m + 1);  // the two blocks of this order are contiguous
    const double *base = tab + off[mc * 2];
    for (int k = threadIdx.x; k < nh; k += blockDim.x) {
        bool hit = false;
        for (long long r = 0; r < rows && !hit; r++) hit = fabs(base[r * NHP + k]) >= thr;
        if (hit) atomicMin(&best, k);
    }
    __syncthreads();
    if (threadIdx.x == 0) kmin[mc] = best;
}

// Fragment-level refinement of the polar skipping.  The table of order m is negligible for colatitudes polewards of the
// turning point sin(theta) ~ m/l, i.e. in a triangle of the (degree, colatitude) plane, not a rectangle.  For every block
// (mc, b) of the table (rows = degrees j, columns = northern colatitudes k):
//   syn[(mc*2+b)*FS + f]  = first 16-degree tile holding an entry >= thr in colatitudes [8f, 8f+8)   (synthesis: M = theta)
//   an [(mc*2+b)*FA + jf] = first 16-colatitude tile holding an entry >= thr in degrees [8jf, 8jf+8) (analysis:  M = degree)
// 255 = none / padding.  blockIdx = (mc, b); threads stride over fragments.
__global__ void table_fskip_kernel(const double *__restrict__ tab, const long long *__restrict__ off, const int *__restrict__ ne,
                                   const int *__restrict__ no, int nh, int NHP, double thr, int FS, int FA,
                                   unsigned char *__restrict__ syn, unsigned char *__restrict__ an) {
    const int mc = blockIdx.x, b = blockIdx.y;
    const int rows = b == 0 ? ne[mc] : no[mc];
    const double *base = tab + off[mc * 2 + b];
    unsigned char *so = syn + ((size_t)mc * 2 + b) * FS, *ao = an + ((size_t)mc * 2 + b) * FA;
    for (int f = threadIdx.x; f < FS; f += blockDim.x) {
        int first = 255;
        if (8 * f < nh) {
            for (int j = 0; j < rows && first == 255; j++) {
                const double *r = base + (size_t)j * NHP + 8 * f;
                bool hit = false;
#pragma unroll
                for (int c = 0; c < 8; c++) hit |= (8 * f + c < nh) && fabs(r[c]) >= thr;
                if (hit) first = j / BK;
            }
        }
        so[f] = (unsigned char)first;
    }
    for (int jf = threadIdx.x; jf < FA; jf += blockDim.x) {
        int first = 255;
        if (8 * jf < rows) {
            for (int kt = 0; kt < NHP / BK && first == 255; kt++) {
                bool hit = false;
                for (int j = 8 * jf; j < min(rows, 8 * jf + 8); j++) {
                    const double *r = base + (size_t)j * NHP + kt * BK;
#pragma unroll
                    for (int c = 0; c < BK; c++) hit |= (kt * BK + c < nh) && fabs(r[c]) >= thr;
                }
                if (hit) first = kt;
            }
        }
        ao[jf] = (unsigned char)first;
    }
}

// ------------------------------------------------------------------------------------------------------
// Synthesis operand assembly: the pre-scalings of the sht_native.f90 wrappers (l(l+1), or2, i*m, l>lcut masks), the level
// masks of rIter.f90:466-622, and the 3-point combination that turns the reference's sums against dPlm into sums against
// Plm (plms.f90:117-187: dPlm(l) = l c(l+1) Plm(l+1) - (l+1) c(l) Plm(l-1), so
//      sum_l S(l) dPlm(l) = sum_l' Plm(l') [ (l'-1) c(l') S(l'-1) - (l'+2) c(l'+1) S(l'+1) ],   l' = m .. lcut+1).
// One CTA per block of 32 consecutive degrees l' of one order (plus one halo degree on either side): lane i owns degree
// m + 32*jt + i, so the spectral reads are 512-byte contiguous per (source, level); even/odd degrees feed the two parity
// problems.  Warp w of CTA (x, y) takes level y*8 + w.
struct SynthPrepArgs {
    const double *src[MAGIC_MAX_SRC];  // complex [n_lev][lm_max]
    const ScalCol *scal;
    const VecPair *vec;
    int ncol_s, npair_v, n_lev, lm_max;
    int N;
    const LevelInfo *lev;
    const int *lstart;   // lm index of degree l=m for each mc
    const double *clm;   // c(l), l = m .. l_max+2, at lstart[mc] + 2 mc + (l - m)
    int l_max, minc, nsrc;
    double *B;
    const long long *offB;  // per problem (mc*2+s), doubles
    const int2 *blks;       // (mc, jt)
};

__device__ __forceinline__ bool level_enabled(int lmask, const LevelInfo &L) {
    if (lmask == LM_VEL) return L.nBc != 2;
    if (lmask == LM_VELBULK) return L.nBc == 0;
    if (lmask == LM_DERIV) return L.lDeriv != 0;
    return true;
}

constexpr int PREP_WARPS = 8;   // = levels per CTA
#ifndef MAGIC_PREP_BATCH
#define MAGIC_PREP_BATCH 8
#endif
constexpr int PREP_BATCH = MAGIC_PREP_BATCH;  // sources loaded per round of phase 1 (divides MAGIC_MAX_SRC)
constexpr int PREP_LD = 35;     // staging tile: positions 0..33 = degrees l0-1 .. l0+32, odd stride against bank conflicts

__device__ __forceinline__ double2 eval_terms(const Term *t, const double2 *xs /* staging tile of this level */, int pos, int src_stride,
                                              int l, int m, double or2) {
    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        int ft = t[i].ftype;
        if (ft == F_NONE) continue;
        double2 x = xs[t[i].src * src_stride + pos];
        double dlh = (double)(l * (l + 1));
        if (ft == F_ONE) { acc.x += x.x; acc.y += x.y; }
        else if (ft == F_DLH) { acc.x += dlh * x.x; acc.y += dlh * x.y; }
        else if (ft == F_OR2DLH) { double f = or2 * dlh; acc.x += f * x.x; acc.y += f * x.y; }
        else if (ft == F_NEG) { acc.x -= x.x; acc.y -= x.y; }
        else { double dm = (double)m; acc.x += -dm * x.y; acc.y += dm * x.x; }  // F_IM: i*m*x
    }
    return acc;
}

// CTA (x, y): degree block blks[x] = (mc, jt), levels y*8 .. y*8+7.
//   phase 1: warp w reads level y*8+w, lane = degree: every source is one 512-byte request (plus the two halo degrees); the
//            values go to a shared tile xs[src][level][position];
//   phase 2: thread (degree d = tid/8, level lv = tid%8) evaluates all columns of its (degree, level) and stores them: the 8
//            lanes of a degree write 128 contiguous bytes of one operand row.
#ifndef MAGIC_PREP_MINB
#define MAGIC_PREP_MINB 4   // 64 registers: with the dense source numbering of layout_bind (11 staged sources in an MHD run, 49 KB
#endif                      // per CTA) four CTAs fit an SM instead of three: 1.08 -> 0.95 ms per 16-level chunk at l_max = 1023
__global__ void __launch_bounds__(PREP_WARPS * 32, MAGIC_PREP_MINB) synth_prep_kernel(SynthPrepArgs a) {
    extern __shared__ __align__(16) double2 prep_sm[];
    const int2 blk = a.blks[blockIdx.x];
    const int mc = blk.x, jt = blk.y, m = mc * a.minc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lev0 = blockIdx.y * PREP_WARPS;
    const int src_stride = PREP_WARPS * PREP_LD;
    const int l0 = m + 32 * jt;
    {   // ---- phase 1: positions p = 0..33 hold degrees l0-1+p
        const int lev = lev0 + warp;
#pragma unroll
        for (int round = 0; round < 2; round++) {
            const int pos = round == 0 ? lane : 32 + lane;
            if (round == 1 && lane >= 2) break;
            const int l = l0 - 1 + pos;
            const bool ok = lev < a.n_lev && l >= m && l <= a.l_max;
            const size_t lm = (size_t)a.lstart[mc] + (l - m);
            // PREP_BATCH sources at a time, every load issued before the first store: the loads are unconditional (an absent
            // source or an out-of-range degree reads a valid dummy address and is zeroed by a select), so the compiler cannot
            // put a branch between them and each warp keeps PREP_BATCH 512-byte requests in flight (ncu on the branchy form:
            // 48 % of all stall samples on the STS that waited for its own LDG, 2.3 TB/s; measured per 16-level chunk at
            // l_max = 1023: one at a time 1.44 ms, batches of 4 1.18 ms, batches of 8 1.08 ms, 66 registers in each case).
            const size_t off = ok ? 2 * ((size_t)lev * a.lm_max + lm) : 0;
#pragma unroll
            for (int s0 = 0; s0 < MAGIC_MAX_SRC; s0 += PREP_BATCH) {
                if (s0 < a.nsrc) {
                    double2 x[PREP_BATCH];
#pragma unroll
                    for (int q = 0; q < PREP_BATCH; q++) {
                        const double *sp = a.src[s0 + q];
                        const bool have = (s0 + q < a.nsrc) && sp != nullptr && ok;
                        const double *p = have ? sp + off : a.clm;
                        x[q] = __ldg(reinterpret_cast<const double2 *>(p));
                        if (!have) x[q] = make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int q = 0; q < PREP_BATCH; q++)
                        if (s0 + q < a.nsrc) prep_sm[((s0 + q) * PREP_WARPS + warp) * PREP_LD + pos] = x[q];
                }
            }
        }
    }
    __syncthreads();
    // ---- phase 2
    const int lv = threadIdx.x & (PREP_WARPS - 1), d = threadIdx.x >> 3, lev = lev0 + lv;
    if (lev >= a.n_lev) return;
    const int l = l0 + d, par = d & 1, r = d >> 1, L1 = a.l_max + 1;
    const int K_par = par == 0 ? (L1 - m) / 2 + 1 : (L1 - m + 1) / 2;
    if (jt >= (K_par + BK - 1) / BK) return;  // the k-tile this degree block maps to does not exist for my parity
    double *row = a.B + a.offB[mc * 2 + par] + (size_t)(jt * BK + r) * a.N;
    const double dm = (double)m;
    const LevelInfo L = a.lev[lev];
    const double2 *xs = prep_sm + lv * PREP_LD;
    const int pos = d + 1;
    const bool on = l <= a.l_max && l <= L.lcut;
    for (int c = 0; c < a.ncol_s; c++) {
        const ScalCol sc = a.scal[c];
        double2 v = make_double2(0.0, 0.0);
        if (on && level_enabled(sc.lmask, L)) v = eval_terms(sc.t, xs, pos, src_stride, l, m, L.or2);
        *reinterpret_cast<double2 *>(row + 2 * ((size_t)c * a.n_lev + lev)) = v;
    }
    if (a.npair_v == 0) return;
    // degrees l-1, l, l+1 of the (S, T) pair; entry l = 0 is ignored (shtransforms.f90:137-138), degrees above lcut are zero
    const bool on_m = l - 1 >= m && l - 1 >= 1 && l - 1 <= L.lcut;
    const bool on_0 = on && l >= 1;
    const bool on_p = l + 1 <= a.l_max && l + 1 <= L.lcut;
    double cm = 0.0, cp = 0.0;
    if (l <= L1) {
        const double *cl = a.clm + a.lstart[mc] + 2 * mc + (l - m);
        cm = (double)(l - 1) * cl[0];   // dTheta2S-like coefficient of S(l-1)
        cp = (double)(l + 2) * cl[1];   // coefficient of S(l+1)
    }
    for (int pr = 0; pr < a.npair_v; pr++) {
        const VecPair vp = a.vec[pr];
        const double2 z = make_double2(0.0, 0.0);
        double2 S0 = z, T0 = z, Sm = z, Tm = z, Sp = z, Tp = z;
        if (l <= L1 && level_enabled(vp.lmask, L)) {
            if (on_0) { S0 = eval_terms(vp.S, xs, pos, src_stride, l, m, L.or2); T0 = eval_terms(vp.T, xs, pos, src_stride, l, m, L.or2); }
            if (on_m) { Sm = eval_terms(vp.S, xs, pos - 1, src_stride, l - 1, m, L.or2); Tm = eval_terms(vp.T, xs, pos - 1, src_stride, l - 1, m, L.or2); }
            if (on_p) { Sp = eval_terms(vp.S, xs, pos + 1, src_stride, l + 1, m, L.or2); Tp = eval_terms(vp.T, xs, pos + 1, src_stride, l + 1, m, L.or2); }
        }
        // Vtheta = sum_l' P(l') [S'(l') + i m T(l')],  Vphi = sum_l' P(l') [i m S(l') - T'(l')]   (SURVEY.md appendix A with
        // X'(l') = (l'-1) c(l') X(l'-1) - (l'+2) c(l'+1) X(l'+1))
        const double2 Sd = make_double2(cm * Sm.x - cp * Sp.x, cm * Sm.y - cp * Sp.y);
        const double2 Td = make_double2(cm * Tm.x - cp * Tp.x, cm * Tm.y - cp * Tp.y);
        const size_t ct = 2 * ((size_t)(a.ncol_s + 2 * pr) * a.n_lev + lev), cph = 2 * ((size_t)(a.ncol_s + 2 * pr + 1) * a.n_lev + lev);
        *reinterpret_cast<double2 *>(row + ct) = make_double2(Sd.x - dm * T0.y, Sd.y + dm * T0.x);
        *reinterpret_cast<double2 *>(row + cph) = make_double2(-dm * S0.y - Td.x, dm * S0.x - Td.y);
    }
}

// ------------------------------------------------------------------------------------------------------
// Analysis extraction: C matrices of the analysis GEMM -> spectral arrays (the nonlinear_lm_t members of
// get_td.f90:27-45).  Scalar columns are copied.  A vector pair arrives as the projections a(l') = sum_k w A P(l'),
// b(l') = sum_k w B P(l') (A = fft(vp)/sin^2, B = fft(vt)/sin^2; l' = m .. l_max+1) and is combined to the reference's
//     S(l) = [ -i m a(l) + l c(l+1) b(l+1) - (l+1) c(l) b(l-1) ] / l(l+1)
//     T(l) = [ -( l c(l+1) a(l+1) - (l+1) c(l) a(l-1) ) - i m b(l) ] / l(l+1)
// (shtransforms.f90:821-870 with dPlm(l) = l c(l+1) Plm(l+1) - (l+1) c(l) Plm(l-1), plms.f90:117-187).  Degrees above
// lcut are exact zeros (shtransforms.f90:680,693).
struct ExtractArgs {
    const double *C;
    const long long *offC;  // per problem (mc*2+p), doubles
    int N, n_lev, lm_max, nf_s, npair;
    const int *lm2l, *lm2m, *lstart;
    const double *clm;
    int minc;
    const LevelInfo *lev;
    double *out_s;  // [nf_s][n_lev][lm_max] complex
    double *out_v;  // [2*npair][n_lev][lm_max] complex: S, T per pair
};

struct ModeRef {  // where the analysis results of one (l, m) mode live
    const double *own, *up, *dn;  // rows of degrees l, l+1, l-1 (column 0 of the level; dn valid only when has_dn)
    bool has_dn;
    double e, f, dm, ll1;         // l c(l+1), (l+1) c(l), m, l(l+1) (1 for l = 0)
};

__device__ __forceinline__ ModeRef mode_ref(const ExtractArgs &e, int lm, int lev) {
    ModeRef r;
    const int l = e.lm2l[lm], m = e.lm2m[lm], mc = m / e.minc, p = (l - m) & 1, j = (l - m) >> 1;
    const double *cp = e.C + e.offC[mc * 2 + p] + 2 * (size_t)lev, *cq = e.C + e.offC[mc * 2 + 1 - p] + 2 * (size_t)lev;
    r.own = cp + (size_t)j * e.N;
    r.up = cq + (size_t)(j + p) * e.N;
    r.has_dn = l > m;
    r.dn = cq + (size_t)(r.has_dn ? j - 1 + p : 0) * e.N;
    const double *cl = e.clm + e.lstart[mc] + 2 * mc + (l - m);
    r.e = (double)l * cl[1];
    r.f = (double)(l + 1) * cl[0];
    r.dm = (double)m;
    r.ll1 = lm > 0 ? (double)(l * (l + 1)) : 1.0;
    return r;
}

// pair i: column of B (theta-type field) = nf_s + 2 i, column of A (phi-type field) = nf_s + 2 i + 1
__device__ __forceinline__ void pair_combine(const ExtractArgs &e, const ModeRef &r, int i, double2 &S, double2 &T) {
    const size_t cb = 2 * (size_t)(e.nf_s + 2 * i) * e.n_lev, ca = cb + 2 * (size_t)e.n_lev;
    const double2 z = make_double2(0.0, 0.0);
    const double2 a0 = *reinterpret_cast<const double2 *>(r.own + ca), b0 = *reinterpret_cast<const double2 *>(r.own + cb);
    const double2 au = *reinterpret_cast<const double2 *>(r.up + ca), bu = *reinterpret_cast<const double2 *>(r.up + cb);
    const double2 ad = r.has_dn ? *reinterpret_cast<const double2 *>(r.dn + ca) : z;
    const double2 bd = r.has_dn ? *reinterpret_cast<const double2 *>(r.dn + cb) : z;
    S.x = (r.dm * a0.y + (r.e * bu.x - r.f * bd.x)) / r.ll1;
    S.y = (-r.dm * a0.x + (r.e * bu.y - r.f * bd.y)) / r.ll1;
    T.x = (-(r.e * au.x - r.f * ad.x) + r.dm * b0.y) / r.ll1;
    T.y = (-(r.e * au.y - r.f * ad.y) - r.dm * b0.x) / r.ll1;
}

// the same in two halves, so that a caller can issue the loads of several pairs before the first combination
struct PairLoads { double2 a0, b0, au, bu, ad, bd; };
__device__ __forceinline__ PairLoads pair_load(const ExtractArgs &e, const ModeRef &r, int i, bool on) {
    const size_t cb = 2 * (size_t)(e.nf_s + 2 * i) * e.n_lev, ca = cb + 2 * (size_t)e.n_lev;
    const double2 z = make_double2(0.0, 0.0);
    PairLoads p;
    p.a0 = on ? *reinterpret_cast<const double2 *>(r.own + ca) : z;
    p.b0 = on ? *reinterpret_cast<const double2 *>(r.own + cb) : z;
    p.au = on ? *reinterpret_cast<const double2 *>(r.up + ca) : z;
    p.bu = on ? *reinterpret_cast<const double2 *>(r.up + cb) : z;
    p.ad = (on && r.has_dn) ? *reinterpret_cast<const double2 *>(r.dn + ca) : z;
    p.bd = (on && r.has_dn) ? *reinterpret_cast<const double2 *>(r.dn + cb) : z;
    return p;
}
__device__ __forceinline__ void pair_math(const ModeRef &r, const PairLoads &p, double2 &S, double2 &T) {
    S.x = (r.dm * p.a0.y + (r.e * p.bu.x - r.f * p.bd.x)) / r.ll1;
    S.y = (-r.dm * p.a0.x + (r.e * p.bu.y - r.f * p.bd.y)) / r.ll1;
    T.x = (-(r.e * p.au.x - r.f * p.ad.x) + r.dm * p.b0.y) / r.ll1;
    T.y = (-(r.e * p.au.y - r.f * p.ad.y) - r.dm * p.b0.x) / r.ll1;
}

__global__ void __launch_bounds__(256) anal_extract_kernel(ExtractArgs a) {
    int lm = blockIdx.x * blockDim.x + threadIdx.x;
    int lev = blockIdx.y;
    if (lm >= a.lm_max) return;
    const bool on = a.lm2l[lm] <= a.lev[lev].lcut;
    const ModeRef r = mode_ref(a, lm, lev);
    const double2 z = make_double2(0.0, 0.0);
    for (int f = 0; f < a.nf_s; f++) {
        double2 v = on ? *reinterpret_cast<const double2 *>(r.own + 2 * (size_t)f * a.n_lev) : z;
        *reinterpret_cast<double2 *>(a.out_s + 2 * (((size_t)f * a.n_lev + lev) * a.lm_max + lm)) = v;
    }
    for (int i = 0; i < a.npair; i++) {
        double2 S = z, T = z;
        if (on) pair_combine(a, r, i, S, T);
        *reinterpret_cast<double2 *>(a.out_v + 2 * (((size_t)(2 * i) * a.n_lev + lev) * a.lm_max + lm)) = S;
        *reinterpret_cast<double2 *>(a.out_v + 2 * (((size_t)(2 * i + 1) * a.n_lev + lev) * a.lm_max + lm)) = T;
    }
}

// ------------------------------------------------------------------------------------------------------
// get_td epilogue (get_td.f90:133-619) for one chunk of levels; thread per (lm, level).
struct TdFlags {
    int l_conv, l_mag, l_heat, l_conv_nl, l_mag_nl, l_mag_kin, l_anel, l_corr, l_double_curl, l_single_matrix,
        l_chemical_conv, l_anelastic_liquid, l_phase_field;
    double CorFac, epsc, epscXi;
};
struct TdArgs {
    TdFlags f;
    int n_lev, lm_max, l_max, minc;
    const int *lm2l, *lm2m;
    const LevelInfo *lev;
    // nonlinear_lm_t members, each [n_lev][lm_max] complex or null
    const double *AdvrLM, *AdvtLM, *AdvpLM, *VxBrLM, *VxBtLM, *VxBpLM, *VStLM, *VSrLM, *VXitLM, *VXirLM, *heatLM, *phiLM;
    // inputs (level-major, first level of the chunk)
    const double *w, *dw, *ddw, *z, *dz;
    // outputs
    double *dwdt, *dzdt, *dpdt, *dsdt, *dxidt, *dbdt, *djdt, *dVxVhLM, *dVxBhLM, *dVSrLM, *dVXirLM, *dphidt;
};

__device__ __forceinline__ double2 ldc(const double *p, size_t i) { return *reinterpret_cast<const double2 *>(p + 2 * i); }
__device__ __forceinline__ void stc(double *p, size_t i, double2 v) { *reinterpret_cast<double2 *>(p + 2 * i) = v; }
__device__ __forceinline__ double2 c_scale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 c_im(double m, double2 a) { return make_double2(-m * a.y, m * a.x); }  // i*m*a

// inl = index of this (lm, level) inside the nonlinear_lm_t arrays (== i when they live in global memory in the
// [lev][lm] layout, a shared-memory tile index in the fused kernel)
__device__ __forceinline__ void td_compute(const TdArgs &a, int lm, int lev, size_t inl) {
    const LevelInfo L = a.lev[lev];
    const int l = a.lm2l[lm], m = a.lm2m[lm];
    const size_t i = (size_t)lev * a.lm_max + lm;
    const int nBc = L.nBc;
    const double2 zero = make_double2(0.0, 0.0);
    const double dLh = (double)(l * (l + 1)), dm = (double)m;
    // st_map neighbours (blocking.f90:321-335): (l-1,m) or self when l==m ; (l+1,m) or none when l==l_max
    const size_t iS = (l > m) ? i - 1 : i;
    const bool hasA = l < a.l_max;
    const size_t iA = i + 1;
    auto clm = [&](int ll) { return sqrt((double)((ll + m) * (ll - m)) / (double)((2 * ll - 1) * (2 * ll + 1))); };
    const double cl = clm(l), cl1 = clm(l + 1);
    const double dTheta2S = (double)(l - 1) * cl, dTheta2A = (double)(l + 2) * cl1;
    const double dTheta3S = (double)((l - 1) * (l + 1)) * cl, dTheta3A = (double)(l * (l + 2)) * cl1;
    const double dTheta4S = ((double)(l + 1) * cl) * (double)((l - 1) * l);
    const double dTheta4A = ((double)l * cl1) * (double)((l + 1) * (l + 2));
    const TdFlags &F = a.f;
    const bool bulk = (nBc == 0);

    if (F.l_conv) {
        // get_dzdt (get_td.f90:376-452)
        if (bulk) {
            double2 v;
            if (lm == 0) {
                v = zero;
                if (F.l_corr) v = c_scale(2.0 * F.CorFac * L.or2, c_add(c_scale(dTheta3A, ldc(a.dw, iA)), c_scale(L.or1 * dTheta4A, ldc(a.w, iA))));
            } else {
                v = F.l_conv_nl ? c_scale(dLh, ldc(a.AdvpLM, inl)) : zero;
                if (F.l_corr) {
                    double2 cor = zero;
                    if (l < L.lcut) {
                        double2 t = c_im(dm, ldc(a.z, i));
                        t = c_add(t, c_scale(dTheta3A, ldc(a.dw, iA)));
                        t = c_add(t, c_scale(L.or1 * dTheta4A, ldc(a.w, iA)));
                        t = c_add(t, c_scale(dTheta3S, ldc(a.dw, iS)));
                        t = c_sub(t, c_scale(L.or1 * dTheta4S, ldc(a.w, iS)));
                        cor = c_scale(2.0 * F.CorFac * L.or2, t);
                    } else if (l == L.lcut) {
                        double2 t = c_im(dm, ldc(a.z, i));
                        t = c_add(t, c_scale(dTheta3S, ldc(a.dw, iS)));
                        t = c_sub(t, c_scale(L.or1 * dTheta4S, ldc(a.w, iS)));
                        cor = c_scale(2.0 * F.CorFac * L.or2, t);
                    }
                    v = c_add(v, cor);
                }
            }
            stc(a.dzdt, i, v);
        }
        if (F.l_double_curl) {
            // get_dwdt_double_curl (get_td.f90:199-309)
            if (bulk) {
                double2 v, vh = zero;
                if (lm == 0) {
                    v = F.l_conv_nl ? c_scale(L.or2, ldc(a.AdvrLM, inl)) : zero;
                    if (F.l_corr && !F.l_single_matrix) v = c_add(v, c_scale(2.0 * F.CorFac * L.or1 * dTheta2A, ldc(a.z, iA)));
                    stc(a.dwdt, i, v);
                } else {
                    if (F.l_conv_nl) {
                        v = c_scale(dLh * L.or4 * L.orho1, ldc(a.AdvrLM, inl));
                        vh = c_scale(-L.orho1 * L.r * L.r * dLh, ldc(a.AdvtLM, inl));
                    } else v = zero;
                    if (F.l_corr) {
                        double2 cor = zero;
                        if (l <= L.lcut) {
                            double2 q = c_sub(c_scale(L.beta, ldc(a.dw, i)), ldc(a.ddw, i));
                            q = c_add(q, c_scale((L.beta * L.or1 + L.or2) * dLh, ldc(a.w, i)));
                            double2 t = c_im(dm, q);
                            if (l < L.lcut && hasA) {
                                t = c_add(t, c_scale(dTheta3A, c_sub(ldc(a.dz, iA), c_scale(L.beta, ldc(a.z, iA)))));
                            }
                            t = c_add(t, c_scale(dTheta3S, c_sub(ldc(a.dz, iS), c_scale(L.beta, ldc(a.z, iS)))));
                            if (l < L.lcut && hasA)
                                t = c_add(t, c_scale(L.or1, c_sub(c_scale(dTheta4A, ldc(a.z, iA)), c_scale(dTheta4S, ldc(a.z, iS)))));
                            else
                                t = c_sub(t, c_scale(L.or1 * dTheta4S, ldc(a.z, iS)));
                            cor = c_scale(2.0 * F.CorFac * L.or2 * L.orho1, t);
                        }
                        v = c_add(v, cor);
                    }
                    stc(a.dwdt, i, v);
                    stc(a.dVxVhLM, i, vh);
                }
            } else {
                stc(a.dVxVhLM, i, zero);
            }
        } else if (bulk) {
            // get_dwdt (get_td.f90:133-197)
            double2 v = F.l_conv_nl ? c_scale(L.or2, ldc(a.AdvrLM, inl)) : zero;
            if (lm == 0) {
                if (F.l_corr && !F.l_single_matrix) v = c_add(v, c_scale(2.0 * F.CorFac * L.or1 * dTheta2A, ldc(a.z, iA)));
            } else if (F.l_corr) {
                double2 cor = zero;
                if (l < L.lcut) {
                    double2 t = c_im(dm, ldc(a.dw, i));
                    t = c_add(t, c_scale(dTheta2A, ldc(a.z, iA)));
                    t = c_sub(t, c_scale(dTheta2S, ldc(a.z, iS)));
                    cor = c_scale(2.0 * F.CorFac * L.or1, t);
                } else if (l == L.lcut) {
                    double2 t = c_sub(c_im(dm, ldc(a.dw, i)), c_scale(dTheta2S, ldc(a.z, iS)));
                    cor = c_scale(2.0 * F.CorFac * L.or1, t);
                }
                v = c_add(v, cor);
            }
            stc(a.dwdt, i, v);
        }
    }
    if (!F.l_double_curl && bulk && lm > 0) {
        // get_dpdt (get_td.f90:311-374); lm=0 is never written by the reference
        double2 v = F.l_conv_nl ? c_scale(-dLh, ldc(a.AdvtLM, inl)) : zero;
        if (F.l_corr) {
            double2 cor = zero;
            if (l <= L.lcut) {
                double2 q = c_add(ldc(a.dw, i), c_scale(L.or1 * dLh, ldc(a.w, i)));
                double2 t = c_im(-dm, q);
                if (l < L.lcut && hasA) t = c_add(t, c_scale(dTheta3A, ldc(a.z, iA)));
                t = c_add(t, c_scale(dTheta3S, ldc(a.z, iS)));
                cor = c_scale(2.0 * F.CorFac * L.or2, t);
            }
            v = c_add(v, cor);
        }
        stc(a.dpdt, i, v);
    }
    if (F.l_heat) {
        // get_dsdt (get_td.f90:454-521) + dVSrLM from spat_to_qst (rIter.f90:688) + rIter.f90:448-456
        if (bulk) {
            double2 v;
            if (lm == 0) {
                v = make_double2(F.epsc * L.epscProf, 0.0);
                if (F.l_anel) v = c_add(v, F.l_anelastic_liquid ? c_scale(L.temp0, ldc(a.heatLM, inl)) : ldc(a.heatLM, inl));
            } else {
                v = c_scale(dLh, ldc(a.VStLM, inl));
                if (F.l_anel) v = c_add(v, F.l_anelastic_liquid ? c_scale(L.temp0, ldc(a.heatLM, inl)) : ldc(a.heatLM, inl));
            }
            stc(a.dsdt, i, v);
        }
        stc(a.dVSrLM, i, (bulk && !L.l_bound) ? ldc(a.VSrLM, inl) : zero);
    }
    // rIter.f90:698: dphidt = scal_to_SH(phiTerms); get_nl fills phiTerms on bulk levels only (get_nl.f90:333)
    if (F.l_phase_field && a.dphidt) stc(a.dphidt, i, bulk ? ldc(a.phiLM, inl) : zero);
    if (F.l_chemical_conv) {
        if (bulk) stc(a.dxidt, i, lm == 0 ? make_double2(F.epscXi, 0.0) : c_scale(dLh, ldc(a.VXitLM, inl)));
        stc(a.dVXirLM, i, (bulk && !L.l_bound) ? ldc(a.VXirLM, inl) : zero);
    }
    if (F.l_mag) {
        // get_dbdt (get_td.f90:557-619)
        if (bulk) {
            if (F.l_mag_nl || F.l_mag_kin) {
                stc(a.dbdt, i, c_scale(dLh, ldc(a.VxBpLM, inl)));
                stc(a.dVxBhLM, i, c_scale(-dLh * L.r * L.r, ldc(a.VxBtLM, inl)));
                stc(a.djdt, i, c_scale(dLh * L.or4, ldc(a.VxBrLM, inl)));
            } else {
                stc(a.dbdt, i, zero);
                stc(a.djdt, i, zero);
                stc(a.dVxBhLM, i, zero);
            }
        } else {
            if ((F.l_mag_nl || F.l_mag_kin) && lm > 0) stc(a.dVxBhLM, i, c_scale(-dLh * L.r * L.r, ldc(a.VxBtLM, inl)));
            else stc(a.dVxBhLM, i, zero);
        }
    }
}

__global__ void __launch_bounds__(256) get_td_kernel(TdArgs a) {
    int lm = blockIdx.x * blockDim.x + threadIdx.x;
    int lev = blockIdx.y;
    if (lm >= a.lm_max) return;
    td_compute(a, lm, lev, (size_t)lev * a.lm_max + lm);
}

// Fused extraction + get_td: a CTA owns TL consecutive (l,m) modes and all levels of the chunk.  Phase 1 reads the
// analysis GEMM results with the level index fastest (contiguous in the C matrices) into a shared-memory tile; phase 2
// runs get_td with the mode index fastest (contiguous in the spectral inputs/outputs).  The nonlinear_lm_t arrays never
// touch HBM.
struct TdSlots { int s[12]; };  // tile slot of AdvrLM,AdvtLM,AdvpLM,VxBrLM,VxBtLM,VxBpLM,VStLM,VSrLM,VXitLM,VXirLM,heatLM,phiLM (-1 absent)

template <int TL>
__global__ void __launch_bounds__(256) extract_td_kernel(ExtractArgs e, TdArgs t, TdSlots slots) {
    extern __shared__ __align__(16) double2 td_sm[];
    constexpr int TLP = TL + 1;
    const int lm0 = blockIdx.x * TL, n_lev = e.n_lev;
    // a thread owns (mode, level) pairs; tile slot f of a pair: scalar columns first, then (S, T) of every vector pair
    for (int pidx = threadIdx.x; pidx < TL * n_lev; pidx += blockDim.x) {
        const int lev = pidx % n_lev, ll = pidx / n_lev, lm = lm0 + ll;
        const double2 z = make_double2(0.0, 0.0);
        bool on = false;
        ModeRef r;
        if (lm < e.lm_max) {
            on = e.lm2l[lm] <= e.lev[lev].lcut;
            r = mode_ref(e, lm, lev);
        }
        // Loads in batches, every load of a batch issued before its first use: up to four scalar columns, then two vector pairs
        // (12 loads) at a time.  (ncu on the one-at-a-time form: 60 % of all stall samples on the first use of a load that had
        // just been issued -- the trip counts are run-time values, so the compiler did not overlap the iterations.)
        for (int f0 = 0; f0 < e.nf_s; f0 += 4) {
            double2 v[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
                v[q] = (on && f0 + q < e.nf_s) ? *reinterpret_cast<const double2 *>(r.own + 2 * (size_t)(f0 + q) * n_lev) : z;
#pragma unroll
            for (int q = 0; q < 4; q++)
                if (f0 + q < e.nf_s) td_sm[((size_t)(f0 + q) * n_lev + lev) * TLP + ll] = v[q];
        }
        for (int i0 = 0; i0 < e.npair; i0 += 2) {
            const bool two = i0 + 1 < e.npair;
            const PairLoads p0 = pair_load(e, r, i0, on), p1 = pair_load(e, r, two ? i0 + 1 : i0, on && two);
            double2 S = z, T = z;
            if (on) pair_math(r, p0, S, T);
            td_sm[((size_t)(e.nf_s + 2 * i0) * n_lev + lev) * TLP + ll] = S;
            td_sm[((size_t)(e.nf_s + 2 * i0 + 1) * n_lev + lev) * TLP + ll] = T;
            if (two) {
                S = z; T = z;
                if (on) pair_math(r, p1, S, T);
                td_sm[((size_t)(e.nf_s + 2 * i0 + 2) * n_lev + lev) * TLP + ll] = S;
                td_sm[((size_t)(e.nf_s + 2 * i0 + 3) * n_lev + lev) * TLP + ll] = T;
            }
        }
    }
    __syncthreads();
    const double *base = reinterpret_cast<const double *>(td_sm);
    const double **ptrs[12] = {&t.AdvrLM, &t.AdvtLM, &t.AdvpLM, &t.VxBrLM, &t.VxBtLM, &t.VxBpLM, &t.VStLM, &t.VSrLM, &t.VXitLM, &t.VXirLM, &t.heatLM, &t.phiLM};
#pragma unroll
    for (int q = 0; q < 12; q++) *ptrs[q] = slots.s[q] < 0 ? nullptr : base + 2 * (size_t)slots.s[q] * n_lev * TLP;
    for (int idx = threadIdx.x; idx < TL * n_lev; idx += blockDim.x) {
        int ll = idx % TL, lev = idx / TL, lm = lm0 + ll;
        if (lm < e.lm_max) td_compute(t, lm, lev, (size_t)lev * TLP + ll);
    }
}

// ------------------------------------------------------------------------------------------------------
// finish_explicit_assembly (LMLoop.f90:390-453) on LM-distributed device arrays ([n_r_max][nlm] complex, lo order): the radial
// derivatives `work` of dVSrLM / dVXirLM / dVxVhLM / dVxBhLM come from the radial-matrix GEMM; this kernel applies
//   finish_exp_entropy (updateS.f90:543-601):  dsdt  = orho1 (dsdt  - or2 work_s  - l(l+1) or2 dentropy0 w)
//   finish_exp_comp    (updateXI.f90:478-512): dxidt = orho1 (dxidt - or2 work_xi)
//   finish_exp_pol     (updateWP.f90:1002-1031): dwdt += or2 work_v   (l > 0)
//   finish_exp_mag     (updateB.f90:1005-1041):  djdt += or2 work_b   (not the (0,0) mode)
// each only for l <= l_R(n_r).  Null pointers switch a term off.
struct FinishArgs {
    int n_r_max, nlm;
    const int *lo2l;          // degree of local mode i (lo order)
    const int *lo2m;
    const double *or2, *orho1, *dentropy0, *l_R;  // [n_r_max]
    const double *w;          // LM-distributed w (flow container, field 0)
    double *dsdt, *dxidt, *dwdt, *djdt;
    const double *work_s, *work_xi, *work_v, *work_b;
};

__global__ void __launch_bounds__(256) finish_explicit_kernel(FinishArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n_r = blockIdx.y;
    if (i >= a.nlm) return;
    const int l = a.lo2l[i], m = a.lo2m[i];
    if ((double)l > a.l_R[n_r]) return;
    const size_t idx = (size_t)n_r * a.nlm + i;
    const double or2 = a.or2[n_r], orho1 = a.orho1[n_r];
    if (a.dsdt) {
        const double2 d = ldc(a.dsdt, idx), wk = ldc(a.work_s, idx), w = ldc(a.w, idx);
        const double f = (double)(l * (l + 1)) * or2 * a.dentropy0[n_r];
        stc(a.dsdt, idx, make_double2(orho1 * (d.x - or2 * wk.x - f * w.x), orho1 * (d.y - or2 * wk.y - f * w.y)));
    }
    if (a.dxidt) {
        const double2 d = ldc(a.dxidt, idx), wk = ldc(a.work_xi, idx);
        stc(a.dxidt, idx, make_double2(orho1 * (d.x - or2 * wk.x), orho1 * (d.y - or2 * wk.y)));
    }
    if (a.dwdt && l > 0) {
        const double2 d = ldc(a.dwdt, idx), wk = ldc(a.work_v, idx);
        stc(a.dwdt, idx, make_double2(d.x + or2 * wk.x, d.y + or2 * wk.y));
    }
    if (a.djdt && !(l == 0 && m == 0)) {
        const double2 d = ldc(a.djdt, idx), wk = ldc(a.work_b, idx);
        stc(a.djdt, idx, make_double2(d.x + or2 * wk.x, d.y + or2 * wk.y));
    }
}

}  // namespace magic
