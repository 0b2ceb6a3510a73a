/*
 * magic_oracle.c -- CPU restatement of MagIC's radial-loop hot path.  TEST INFRASTRUCTURE ONLY.
 * See magic_oracle.h for scope, conventions and the parity status (pinned / unpinned rows).
 * All file:line citations are relative to /root/reference/src/.
 */
#include "magic_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846264338327950288

struct orc_ctx {
    int l_max, m_max, m_min, minc, n_theta, n_phi, n_m_max, lm_max, nthreads;
    int *lm2l, *lm2m, *lm2, *lm2lmS, *lm2lmA;        /* st_map, 0-based; lm2 is [(l_max+1)*(l_max+1)] */
    double *theta_ord, *gauss;                        /* gauss is scrambled */
    double *sinTheta, *cosTheta, *O_sin_theta, *O_sin_theta_E2, *sinTheta_E2, *cosn_theta_E2, *phi;
    double *dLh, *dTh[8];                             /* dTheta1S,1A,2S,2A,3S,3A,4S,4A */
    double *Plm, *dPlm, *wPlm, *wdPlm;                /* [n_theta/2][lm_max] (Fortran Plm(lm,nThetaNHS)) */
    int *lStart, *lStop;                              /* 0-based inclusive positions per mc */
    double *D_mc2m;
    int nfac, fac[32];
    orc_cplx *trig;                                   /* exp(+2 pi i k / n_phi) */
};

/* ------------------------------------------------------------------------------------------------ */
/* truncation.f90:162-192 prime_decomposition */
static int prime_decomposition(int nlon) {
    double dist_min = 100.0;
    int i0 = 0, j0 = 0, k0 = 0;
    for (int i = 0; i <= 12; i++)
        for (int j = 0; j <= 6; j++)
            for (int k = 0; k <= 6; k++) {
                double res = pow(2.0, i) * pow(3.0, j) * pow(5.0, k);
                double dist = res - nlon;
                if (dist >= 0 && dist < dist_min) { i0 = i; j0 = j; k0 = k; dist_min = dist; }
            }
    return (int)lround(pow(2.0, i0) * pow(3.0, j0) * pow(5.0, k0));
}

/* truncation.f90:55-105 (non-axisymmetric branch) */
void orc_grid_sizes(int l_max_in, int n_phi_tot_in, int minc, int nalias, int out[7]) {
    int l_max = l_max_in, n_phi_tot = n_phi_tot_in, n_theta, n_phi;
    if (l_max == 0) {
        n_phi = n_phi_tot / minc;
        n_theta = n_phi_tot / 2;
        l_max = (nalias * n_theta) / 30;
    } else {
        n_theta = (30 * l_max) / nalias;
        n_phi_tot = prime_decomposition(2 * n_theta);
        n_phi = n_phi_tot / minc;
        n_theta = n_phi_tot / 2;
    }
    int m_max = (l_max / minc) * minc;
    if (m_max > l_max) m_max = l_max;
    int n_m_max = m_max / minc + 1;
    int lm_max = 0;
    for (int m = 0; m <= m_max; m += minc) lm_max += l_max - m + 1;
    out[0] = l_max; out[1] = m_max; out[2] = n_theta; out[3] = n_phi; out[4] = n_m_max; out[5] = lm_max;
    out[6] = n_phi_tot;
}

/* parallel.f90:75-92 getBlocks (1-based inclusive) */
void orc_get_blocks(int n_points, int n_procs, int *start, int *stop) {
    int n_loc = n_points / n_procs, rem = n_points - n_loc * n_procs;
    for (int p = 0; p < n_procs; p++) {
        int a = p + rem - n_procs, b = p + rem + 1 - n_procs;
        start[p] = n_loc * p + (a > 0 ? a : 0) + 1;
        stop[p] = n_loc * (p + 1) + (b > 0 ? b : 0);
        if (p != 0) start[p] = stop[p - 1] + 1;
    }
}

/* horizontal.f90:279-340 gauleg(-1,1,...) */
static void gauleg(int n, double *theta_ord, double *gauss) {
    const double eps = 10.0 * 2.220446049250313e-16;
    int m = (n + 1) / 2;
    for (int i = 1; i <= m; i++) {
        double z = cos(PI * (((double)i - 0.25) / ((double)n + 0.5)));
        double z1 = z + 10.0 * eps, p1 = 0, p2 = 0, p3, pp = 1;
        while (fabs(z - z1) > eps) {
            p1 = 1.0; p2 = 0.0;
            for (int j = 1; j <= n; j++) {
                p3 = p2; p2 = p1;
                p1 = ((double)(2 * j - 1) * z * p2 - (double)(j - 1) * p3) / (double)j;
            }
            pp = (double)n * (z * p1 - p2) / (z * z - 1.0);
            z1 = z;
            z = z1 - p1 / pp;
        }
        theta_ord[i - 1] = acos(z);
        theta_ord[n - i] = acos(-z);
        gauss[i - 1] = 2.0 / ((1.0 - z * z) * pp * pp);
        gauss[n - i] = gauss[i - 1];
    }
}

/* plms.f90:14-189 plm_theta with norm=2; plma/dtheta_plma are [lm_max] in st_map order. */
static void plm_theta(double theta, int max_degree, int min_order, int max_order, int m0, double *plma,
                      double *dtheta_plma) {
    const double dnorm = 1.0 / sqrt(4.0 * PI); /* osq4pi */
    int pos = -1;
    for (int m = min_order; m <= max_order; m += m0) {
        double fac = 1.0;
        for (int j = 3; j <= 2 * m + 1; j += 2) fac = fac * (double)j / (double)(j - 1);
        double plm = sqrt(fac);
        if (sin(theta) != 0.0) plm = plm * pow(sin(theta), m);
        else if (m != 0) plm = 0.0;
        pos++;
        plma[pos] = dnorm * plm;
        double plm1 = 0.0, plm2;
        int l;
        for (l = m + 1; l <= max_degree; l++) {
            plm2 = plm1; plm1 = plm;
            plm = cos(theta) * sqrt((double)((2 * l - 1) * (2 * l + 1)) / (double)((l - m) * (l + m))) * plm1 -
                  sqrt(((double)(2 * l + 1) * (double)(l + m - 1) * (double)(l - m - 1)) /
                       ((double)(2 * l - 3) * (double)(l - m) * (double)(l + m))) * plm2;
            pos++;
            plma[pos] = dnorm * plm;
        }
        l = max_degree + 1;
        plm2 = plm1; plm1 = plm;
        plm = cos(theta) * sqrt((double)((2 * l - 1) * (2 * l + 1)) / (double)((l - m) * (l + m))) * plm1 -
              sqrt((double)((2 * l + 1) * (l + m - 1) * (l - m - 1)) / (double)((2 * l - 3) * (l - m) * (l + m))) * plm2;
        dtheta_plma[pos] = dnorm * plm;
    }
    pos = -1;
    for (int m = min_order; m <= max_order; m += m0) {
        int l = m;
        pos++;
        if (m < max_degree) dtheta_plma[pos] = l / sqrt((double)(2 * l + 3)) * plma[pos + 1];
        else dtheta_plma[pos] = l / sqrt((double)(2 * l + 3)) * dtheta_plma[pos];
        for (l = m + 1; l <= max_degree - 1; l++) {
            pos++;
            dtheta_plma[pos] =
                l * sqrt((double)((l + m + 1) * (l - m + 1)) / (double)((2 * l + 1) * (2 * l + 3))) * plma[pos + 1] -
                (l + 1) * sqrt((double)((l + m) * (l - m)) / (double)((2 * l - 1) * (2 * l + 1))) * plma[pos - 1];
        }
        if (m < max_degree) {
            l = max_degree;
            pos++;
            dtheta_plma[pos] =
                l * sqrt((double)((l + m + 1) * (l - m + 1)) / (double)((2 * l + 1) * (2 * l + 3))) * dtheta_plma[pos] -
                (l + 1) * sqrt((double)((l + m) * (l - m)) / (double)((2 * l - 1) * (2 * l + 1))) * plma[pos - 1];
        }
    }
}

static void factorise(int n, int *fac, int *nfac) {
    int k = 0;
    while (n % 4 == 0) { fac[k++] = 4; n /= 4; }
    while (n % 2 == 0) { fac[k++] = 2; n /= 2; }
    while (n % 3 == 0) { fac[k++] = 3; n /= 3; }
    while (n % 5 == 0) { fac[k++] = 5; n /= 5; }
    for (int p = 7; n > 1; p += 2)
        while (n % p == 0) { fac[k++] = p; n /= p; }
    *nfac = k;
}

orc_ctx *orc_create(int l_max, int m_max, int minc, int n_theta, int n_phi) {
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->l_max = l_max; c->m_max = m_max; c->m_min = 0; c->minc = minc;
    c->n_theta = n_theta; c->n_phi = n_phi; c->n_m_max = m_max / minc + 1; c->nthreads = 1;
    int lm_max = 0;
    for (int m = 0; m <= m_max; m += minc) lm_max += l_max - m + 1;
    c->lm_max = lm_max;
    /* blocking.f90:293-337 get_standard_lm_blocking */
    int L1 = l_max + 1;
    c->lm2l = (int *)malloc(sizeof(int) * lm_max); c->lm2m = (int *)malloc(sizeof(int) * lm_max);
    c->lm2lmS = (int *)malloc(sizeof(int) * lm_max); c->lm2lmA = (int *)malloc(sizeof(int) * lm_max);
    c->lm2 = (int *)malloc(sizeof(int) * L1 * L1);
    for (int i = 0; i < L1 * L1; i++) c->lm2[i] = -1;
    int lm = 0;
    for (int m = 0; m <= m_max; m += minc)
        for (int l = m; l <= l_max; l++) { c->lm2l[lm] = l; c->lm2m[lm] = m; c->lm2[l * L1 + m] = lm; lm++; }
    for (lm = 0; lm < lm_max; lm++) {
        int l = c->lm2l[lm], m = c->lm2m[lm];
        c->lm2lmS[lm] = (l > 0 && l > m) ? c->lm2[(l - 1) * L1 + m] : lm;
        c->lm2lmA[lm] = (l < l_max) ? c->lm2[(l + 1) * L1 + m] : -1;
    }
    /* horizontal.f90:116-229 (l_scramble_theta = .true.) */
    c->theta_ord = (double *)malloc(sizeof(double) * n_theta);
    double *tmp_gauss = (double *)malloc(sizeof(double) * n_theta);
    gauleg(n_theta, c->theta_ord, tmp_gauss);
    double **tv[] = {&c->gauss, &c->sinTheta, &c->cosTheta, &c->O_sin_theta, &c->O_sin_theta_E2,
                     &c->sinTheta_E2, &c->cosn_theta_E2};
    for (int i = 0; i < 7; i++) *tv[i] = (double *)calloc(n_theta, sizeof(double));
    for (int k = 0; k < n_theta / 2; k++) {
        double colat = c->theta_ord[k];
        int n = 2 * k, s = 2 * k + 1;
        c->O_sin_theta[n] = c->O_sin_theta[s] = 1.0 / sin(colat);
        c->O_sin_theta_E2[n] = c->O_sin_theta_E2[s] = 1.0 / (sin(colat) * sin(colat));
        c->sinTheta[n] = c->sinTheta[s] = sin(colat);
        c->sinTheta_E2[n] = c->sinTheta_E2[s] = sin(colat) * sin(colat);
        c->cosTheta[n] = cos(colat); c->cosTheta[s] = -cos(colat);
        c->cosn_theta_E2[n] = cos(colat) / sin(colat) / sin(colat);
        c->cosn_theta_E2[s] = -cos(colat) / sin(colat) / sin(colat);
        c->gauss[n] = tmp_gauss[k]; c->gauss[s] = tmp_gauss[n_theta - 1 - k];
    }
    c->phi = (double *)malloc(sizeof(double) * n_phi);
    for (int j = 0; j < n_phi; j++) c->phi[j] = j * (2.0 * PI / (double)(n_phi * minc));
    /* horizontal.f90:202-229 */
    int L2 = l_max + 2;
    double *clm = (double *)calloc((size_t)L2 * L2, sizeof(double));
    for (int m = 0; m <= m_max; m += minc)
        for (int l = m; l <= l_max + 1; l++)
            clm[l * L2 + m] = sqrt((double)((l + m) * (l - m)) / (double)((2 * l - 1) * (2 * l + 1)));
    c->dLh = (double *)malloc(sizeof(double) * lm_max);
    for (int i = 0; i < 8; i++) c->dTh[i] = (double *)malloc(sizeof(double) * lm_max);
    for (lm = 0; lm < lm_max; lm++) {
        int l = c->lm2l[lm], m = c->lm2m[lm];
        c->dLh[lm] = (double)(l * (l + 1));
        c->dTh[0][lm] = (double)(l + 1) * clm[l * L2 + m];           /* dTheta1S */
        c->dTh[1][lm] = (double)l * clm[(l + 1) * L2 + m];           /* dTheta1A */
        c->dTh[2][lm] = (double)(l - 1) * clm[l * L2 + m];           /* dTheta2S */
        c->dTh[3][lm] = (double)(l + 2) * clm[(l + 1) * L2 + m];     /* dTheta2A */
        c->dTh[4][lm] = (double)((l - 1) * (l + 1)) * clm[l * L2 + m];  /* dTheta3S */
        c->dTh[5][lm] = (double)(l * (l + 2)) * clm[(l + 1) * L2 + m];  /* dTheta3A */
        c->dTh[6][lm] = c->dTh[0][lm] * (double)((l - 1) * l);       /* dTheta4S */
        c->dTh[7][lm] = c->dTh[1][lm] * (double)((l + 1) * (l + 2)); /* dTheta4A */
    }
    free(clm);
    /* shtransforms.f90:38-91 initialize_transforms */
    size_t tsz = (size_t)lm_max * (n_theta / 2);
    c->Plm = (double *)malloc(sizeof(double) * tsz); c->dPlm = (double *)malloc(sizeof(double) * tsz);
    c->wPlm = (double *)malloc(sizeof(double) * tsz); c->wdPlm = (double *)malloc(sizeof(double) * tsz);
    for (int k = 0; k < n_theta / 2; k++) {
        double *pl = c->Plm + (size_t)k * lm_max, *dpl = c->dPlm + (size_t)k * lm_max;
        plm_theta(c->theta_ord[k], l_max, 0, m_max, minc, pl, dpl);
        for (lm = 0; lm < lm_max; lm++) {
            c->wPlm[(size_t)k * lm_max + lm] = 2.0 * PI * tmp_gauss[k] * pl[lm];
            c->wdPlm[(size_t)k * lm_max + lm] = 2.0 * PI * tmp_gauss[k] * dpl[lm];
        }
    }
    free(tmp_gauss);
    c->lStart = (int *)malloc(sizeof(int) * c->n_m_max); c->lStop = (int *)malloc(sizeof(int) * c->n_m_max);
    c->D_mc2m = (double *)malloc(sizeof(double) * c->n_m_max);
    c->lStart[0] = 0; c->lStop[0] = l_max; c->D_mc2m[0] = 0.0;
    for (int mc = 1; mc < c->n_m_max; mc++) {
        int m = mc * minc;
        c->D_mc2m[mc] = (double)m;
        c->lStart[mc] = c->lStop[mc - 1] + 1;
        c->lStop[mc] = c->lStart[mc] + l_max - m;
    }
    /* FFT plan (any exact DFT obeying fft.f90:262-268 will do; this is a mixed-radix Stockham) */
    factorise(n_phi, c->fac, &c->nfac);
    c->trig = (orc_cplx *)malloc(sizeof(orc_cplx) * n_phi);
    for (int k = 0; k < n_phi; k++) {
        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n_phi;
        c->trig[k] = (double)cosl(a) + I * (double)sinl(a);
    }
    return c;
}

void orc_destroy(orc_ctx *c) {
    if (!c) return;
    free(c->lm2l); free(c->lm2m); free(c->lm2); free(c->lm2lmS); free(c->lm2lmA);
    free(c->theta_ord); free(c->gauss); free(c->sinTheta); free(c->cosTheta); free(c->O_sin_theta);
    free(c->O_sin_theta_E2); free(c->sinTheta_E2); free(c->cosn_theta_E2); free(c->phi);
    free(c->dLh); for (int i = 0; i < 8; i++) free(c->dTh[i]);
    free(c->Plm); free(c->dPlm); free(c->wPlm); free(c->wdPlm);
    free(c->lStart); free(c->lStop); free(c->D_mc2m); free(c->trig);
    free(c);
}

void orc_set_threads(orc_ctx *c, int n) { c->nthreads = n > 0 ? n : 1; }
int orc_lm_max(const orc_ctx *c) { return c->lm_max; }
int orc_n_m_max(const orc_ctx *c) { return c->n_m_max; }
const int *orc_lm2l(const orc_ctx *c) { return c->lm2l; }
const int *orc_lm2m(const orc_ctx *c) { return c->lm2m; }
const int *orc_lm2lmS(const orc_ctx *c) { return c->lm2lmS; }
const int *orc_lm2lmA(const orc_ctx *c) { return c->lm2lmA; }
const double *orc_theta_ord(const orc_ctx *c) { return c->theta_ord; }
const double *orc_gauss(const orc_ctx *c) { return c->gauss; }
const double *orc_plm(const orc_ctx *c) { return c->Plm; }
const double *orc_dplm(const orc_ctx *c) { return c->dPlm; }
const double *orc_theta_vec(const orc_ctx *c, int w) {
    const double *v[] = {c->sinTheta, c->cosTheta, c->O_sin_theta, c->O_sin_theta_E2, c->sinTheta_E2, c->cosn_theta_E2};
    return v[w];
}
const double *orc_lm_vec(const orc_ctx *c, int w) { return w == 0 ? c->dLh : c->dTh[w - 1]; }

/* blocking.f90:339-385 (l-major) and :387-544 (snake) */
void orc_lo_map(const orc_ctx *c, int n_procs, int *lo2st, int *lm_start, int *lm_stop) {
    int l_max = c->l_max, L1 = l_max + 1, minc = c->minc, m_max = c->m_max, m_min = c->m_min;
    if (n_procs <= l_max / 2) {
        int nl = l_max + 1 - m_min;
        int *l_list = (int *)calloc((size_t)n_procs * nl, sizeof(int));
        int *l_counter = (int *)calloc(n_procs, sizeof(int));
        int proc = 0, ascending = 1, l0proc = 0;
        for (int l = l_max; l >= m_min; l--) {
            l_list[proc * nl + l_counter[proc]] = l;
            l_counter[proc]++;
            if (l == 0) l0proc = proc;
            if (ascending) {
                if (proc < n_procs - 1) proc++;
                else ascending = 0;
            } else {
                if (proc > 0) proc--;
                else ascending = 1;
            }
        }
        if (l0proc != 0) { /* rotate so that the l=0 owner becomes rank 0 (blocking.f90:454-472) */
            int *tl = (int *)malloc(sizeof(int) * nl);
            memcpy(tl, l_list, sizeof(int) * nl);
            int tc = l_counter[0], pc = 0;
            for (;;) {
                int src = (l0proc + pc) % n_procs;
                if (src != 0) {
                    memcpy(l_list + pc * nl, l_list + src * nl, sizeof(int) * nl);
                    l_counter[pc] = l_counter[src];
                } else {
                    memcpy(l_list + pc * nl, tl, sizeof(int) * nl);
                    l_counter[pc] = tc;
                    break;
                }
                pc = src;
            }
            free(tl);
        }
        for (int i = 0; i < l_counter[0]; i++) /* l=0 first on rank 0 (blocking.f90:476-484) */
            if (l_list[i] == 0) { int t = l_list[0]; l_list[0] = 0; l_list[i] = t; break; }
        int lm = 0;
        for (proc = 0; proc < n_procs; proc++) {
            lm_start[proc] = lm + 1;
            for (int i = 0; i < l_counter[proc]; i++) {
                int l = l_list[proc * nl + i];
                int mm = m_max < l ? m_max : l;
                for (int m = m_min; m <= mm; m += minc) lo2st[lm++] = c->lm2[l * L1 + m];
            }
            lm_stop[proc] = lm;
        }
        free(l_list); free(l_counter);
    } else {
        orc_get_blocks(c->lm_max, n_procs, lm_start, lm_stop);
        int lm = 0;
        for (int l = m_min; l <= l_max; l++) {
            int mm = m_max < l ? m_max : l;
            for (int m = m_min; m <= mm; m += minc) lo2st[lm++] = c->lm2[l * L1 + m];
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* FFT: complex Stockham autosort, mixed radix.  x,y length n; sign=+1 uses trig, -1 its conjugate.  */
static void fft_stockham(const orc_ctx *c, orc_cplx *x, orc_cplx *y, int sign) {
    int n = c->n_phi, len = n, s = 1;
    orc_cplx *src = x, *dst = y;
    for (int f = 0; f < c->nfac; f++) {
        int r = c->fac[f], m = len / r;
        orc_cplx wjk[8][8], tw[8]; /* the radix-r DFT matrix and the twiddles of this p: same table entries as before, hoisted */
        for (int k = 0; k < r; k++)
            for (int j = 0; j < r; j++) {
                orc_cplx w = c->trig[(size_t)((long)(j * k % r) * (n / r)) % n];
                wjk[k][j] = sign > 0 ? w : conj(w);
            }
        for (int p = 0; p < m; p++) {
            for (int k = 0; k < r; k++) {
                orc_cplx t = c->trig[(size_t)((long)p * k * (n / len)) % n];
                tw[k] = sign > 0 ? t : conj(t);
            }
            for (int q = 0; q < s; q++) {
                orc_cplx a[8], b;
                for (int j = 0; j < r; j++) a[j] = src[q + s * (p + m * j)];
                for (int k = 0; k < r; k++) {
                    b = 0;
                    for (int j = 0; j < r; j++) b += a[j] * wjk[k][j];
                    dst[q + s * (r * p + k)] = b * tw[k];
                }
            }
        }
        len = m; s *= r;
        orc_cplx *t = src; src = dst; dst = t;
    }
    if (src != x) memcpy(x, src, sizeof(orc_cplx) * n);
}

/* fft.f90:211-252 ifft_many: x_j = sum_k c_k e^{+2 pi i j k/n}, c_{n-k}=conj(c_k); Im c_0, Im c_{n/2} ignored */
void orc_ifft_many(const orc_ctx *c, const orc_cplx *f, double *g) {
    int n = c->n_phi, nlat = c->n_theta;
#pragma omp parallel num_threads(c->nthreads)
    {
        orc_cplx *x = (orc_cplx *)malloc(sizeof(orc_cplx) * n), *y = (orc_cplx *)malloc(sizeof(orc_cplx) * n);
#pragma omp for
        for (int t = 0; t < nlat; t++) {
            x[0] = creal(f[t]);
            for (int k = 1; k < n / 2; k++) {
                x[k] = f[(size_t)k * nlat + t];
                x[n - k] = conj(x[k]);
            }
            x[n / 2] = creal(f[(size_t)(n / 2) * nlat + t]);
            fft_stockham(c, x, y, +1);
            for (int j = 0; j < n; j++) g[(size_t)j * nlat + t] = creal(x[j]);
        }
        free(x); free(y);
    }
}

/* fft.f90:164-209 fft_many: c_k = (1/n) sum_j x_j e^{-2 pi i j k/n}, k=0..n/2 */
void orc_fft_many(const orc_ctx *c, const double *g, orc_cplx *f) {
    int n = c->n_phi, nlat = c->n_theta;
#pragma omp parallel num_threads(c->nthreads)
    {
        orc_cplx *x = (orc_cplx *)malloc(sizeof(orc_cplx) * n), *y = (orc_cplx *)malloc(sizeof(orc_cplx) * n);
#pragma omp for
        for (int t = 0; t < nlat; t++) {
            for (int j = 0; j < n; j++) x[j] = g[(size_t)j * nlat + t];
            fft_stockham(c, x, y, -1);
            for (int k = 0; k <= n / 2; k++) f[(size_t)k * nlat + t] = x[k] / (double)n;
        }
        free(x); free(y);
    }
}

#include "magic_oracle_sht.inc"
#include "magic_oracle_rloop.inc"
#include "magic_oracle_diag.inc"
