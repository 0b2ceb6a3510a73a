// engine.cu -- SHT handle, batch layouts and pipeline stages of magic_b200.
#include "engine.cuh"

#include <algorithm>

namespace magic {

thread_local std::string g_last_error;

// horizontal.f90:279-340 gauleg(-1,1,...) on the host (n_theta Newton solves; same iteration as the reference)
static void gauleg_host(int n, std::vector<double> &theta_ord, std::vector<double> &gauss) {
    const double pi = 3.14159265358979323846264338327950288;
    const double eps = 10.0 * 2.220446049250313e-16;
    theta_ord.assign(n, 0.0);
    gauss.assign(n, 0.0);
    int m = (n + 1) / 2;
    for (int i = 1; i <= m; i++) {
        double z = cos(pi * (((double)i - 0.25) / ((double)n + 0.5)));
        double z1 = z + 10.0 * eps, p1 = 0, p2 = 0, p3, pp = 1;
        while (fabs(z - z1) > eps) {
            p1 = 1.0;
            p2 = 0.0;
            for (int j = 1; j <= n; j++) {
                p3 = p2;
                p2 = p1;
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

static bool factor_fft(int H, FftPlan &p) {
    p.nfac = 0;
    int n = H;
    while (n % 4 == 0) { p.fac[p.nfac++] = 4; n /= 4; }
    while (n % 2 == 0) { p.fac[p.nfac++] = 2; n /= 2; }
    while (n % 3 == 0) { p.fac[p.nfac++] = 3; n /= 3; }
    while (n % 5 == 0) { p.fac[p.nfac++] = 5; n /= 5; }
    return n == 1 && p.nfac <= 16;
}

int sht_init(magic_sht *h) {
    const int l_max = h->l_max, minc = h->minc, n_m = h->n_m, nh = h->nh;
    MCHECK(cudaSetDevice(h->dev));
    MCHECK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    // st_map (blocking.f90:309-317)
    h->lm2l.clear(); h->lm2m.clear(); h->lstart.assign(n_m, 0); h->ne1.assign(n_m, 0); h->no1.assign(n_m, 0);
    std::vector<double> clm;  // c(l), l = m .. l_max+2 per order (the dTheta coefficients of plms.f90:117-187 / horizontal.f90:202-229)
    for (int mc = 0; mc < n_m; mc++) {
        int m = mc * minc;
        h->lstart[mc] = (int)h->lm2l.size();
        for (int l = m; l <= l_max; l++) { h->lm2l.push_back(l); h->lm2m.push_back(m); }
        h->ne1[mc] = (l_max + 1 - m) / 2 + 1;  // degrees m, m+2, ... <= l_max+1
        h->no1[mc] = (l_max + 1 - m + 1) / 2;  // degrees m+1, m+3, ... <= l_max+1
        for (int l = m; l <= l_max + 2; l++) clm.push_back(sqrt((double)((l + m) * (l - m)) / (double)((2 * l - 1) * (2 * l + 1))));
    }
    if ((int)h->lm2l.size() != h->lm_max) MFAIL("internal: lm_max mismatch");
    gauleg_host(h->n_theta, h->theta_ord, h->gauss);
    // table block offsets
    h->off.assign((size_t)n_m * 2, 0);
    long long pos = 0;
    for (int mc = 0; mc < n_m; mc++) {
        int rows[2] = {h->ne1[mc], h->no1[mc]};
        for (int b = 0; b < 2; b++) { h->off[(size_t)mc * 2 + b] = pos; pos += (long long)rows[b] * h->NHP; }
    }
    long long tab_doubles = pos + (long long)(GEMM_BM + 2 * BK) * std::max(h->NHP, GEMM_BM) + 1024;
    MCHECK(cudaMalloc((void **)&h->d_tab, sizeof(double) * tab_doubles));
    MCHECK(cudaMemsetAsync(h->d_tab, 0, sizeof(double) * tab_doubles, h->stream));
    if (dev_upload_vec(&h->d_off, h->off) || dev_upload_vec(&h->d_clm, clm)) return 1;
    std::vector<double> sinth(nh), costh(nh), wg(nh), os2(nh), pmm(n_m);
    const double pi = 3.14159265358979323846264338327950288;
    for (int k = 0; k < nh; k++) {
        double colat = h->theta_ord[k];
        sinth[k] = sin(colat);
        costh[k] = cos(colat);
        wg[k] = 2.0 * pi * h->gauss[k] / (double)h->n_phi;
        os2[k] = 1.0 / (sin(colat) * sin(colat));
    }
    for (int mc = 0; mc < n_m; mc++) {  // plms.f90:54-59
        int m = mc * minc;
        double fac = 1.0;
        for (int j = 3; j <= 2 * m + 1; j += 2) fac = fac * (double)j / (double)(j - 1);
        pmm[mc] = sqrt(fac);
    }
    double *d_pmm = nullptr;
    if (dev_upload_vec(&h->d_sinth, sinth) || dev_upload_vec(&h->d_costh, costh) || dev_upload_vec(&h->d_wgauss, wg) ||
        dev_upload_vec(&h->d_osin2, os2) || dev_upload_vec(&d_pmm, pmm) || dev_upload_vec(&h->d_lm2l, h->lm2l) ||
        dev_upload_vec(&h->d_lm2m, h->lm2m) || dev_upload_vec(&h->d_lstart, h->lstart))
        return 1;
    {
        dim3 grid((nh + 127) / 128, n_m);
        build_tables_kernel<<<grid, 128, 0, h->stream>>>(h->d_tab, h->d_off, h->d_sinth, h->d_costh, d_pmm, nh, h->NHP, l_max, minc, n_m);
        MCHECK(cudaGetLastError());
    }
    // polar skipping threshold: MAGIC_POLAR_EPS=0 disables it
    if (const char *e = getenv("MAGIC_POLAR_EPS")) h->polar_eps = atof(e);
    h->kmin.assign(n_m, 0);
    if (h->polar_eps > 0.0) {
        int *d_kmin = nullptr;
        MCHECK(cudaMalloc((void **)&d_kmin, sizeof(int) * n_m));
        table_kmin_kernel<<<n_m, 256, 0, h->stream>>>(h->d_tab, h->d_off, nh, h->NHP, l_max, minc, h->polar_eps, d_kmin);
        MCHECK(cudaGetLastError());
        MCHECK(cudaMemcpyAsync(h->kmin.data(), d_kmin, sizeof(int) * n_m, cudaMemcpyDeviceToHost, h->stream));
        MCHECK(cudaStreamSynchronize(h->stream));
        cudaFree(d_kmin);
        // fragment-level tables (MAGIC_POLAR_FRAG=0 keeps the per-order rectangle only)
        const char *fr = getenv("MAGIC_POLAR_FRAG");
        if (!fr || atoi(fr) != 0) {
            h->FS = ((h->NHP / 8 + 15) / 16) * 16 + 16;
            h->FA = ((((l_max + 1) / 2 + 1 + 7) / 8 + 15) / 16) * 16 + 16;
            int *d_ne = nullptr, *d_no = nullptr;
            if (dev_upload_vec(&d_ne, h->ne1) || dev_upload_vec(&d_no, h->no1)) return 1;
            MCHECK(cudaMalloc((void **)&h->d_fskip_syn, (size_t)n_m * 2 * h->FS));
            MCHECK(cudaMalloc((void **)&h->d_fskip_an, (size_t)n_m * 2 * h->FA));
            table_fskip_kernel<<<dim3(n_m, 2), 128, 0, h->stream>>>(h->d_tab, h->d_off, d_ne, d_no, nh, h->NHP, h->polar_eps, h->FS, h->FA,
                                                                    h->d_fskip_syn, h->d_fskip_an);
            MCHECK(cudaGetLastError());
            MCHECK(cudaStreamSynchronize(h->stream));
            cudaFree(d_ne);
            cudaFree(d_no);
        }
    }
    // FFT plan
    h->fft.N = h->n_phi;
    h->fft.H = h->n_phi / 2;
    if (!factor_fft(h->fft.H, h->fft)) MFAIL("n_phi_max/2 must factor into 2,3,5 (fft.f90 supports radices 2,3,4,5)");
    std::vector<double2> tw(h->n_phi);
    for (int k = 0; k < h->n_phi; k++) {
        long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)h->n_phi;
        tw[k] = make_double2((double)cosl(a), (double)sinl(a));
    }
    if (dev_upload_vec(&h->d_tw, tw)) return 1;
    h->fft.tw = h->d_tw;
    MCHECK(gemm_setup_attributes());
    {   // tile shape of the analysis GEMM (common.cuh): wide below MAGIC_GEMM_WIDE_BELOW, MAGIC_GEMM_AN_WIDE=0/1 forces it
        const char *e = getenv("MAGIC_GEMM_AN_WIDE");
        h->an_wide = e ? atoi(e) != 0 : h->l_max < MAGIC_GEMM_WIDE_BELOW;
    }
    MCHECK(cudaFuncSetAttribute(extract_td_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MCHECK(cudaFuncSetAttribute(extract_td_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MCHECK(cudaFuncSetAttribute(extract_td_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MCHECK(cudaFuncSetAttribute(synth_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PREP_WARPS * MAGIC_MAX_SRC * PREP_LD * (int)sizeof(double2)));
    cudaFuncSetAttribute(synth_prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);  // its occupancy is set by the staging tile
    MCHECK(fft_setup_attributes(h->fft.H));
    MCHECK(cudaStreamSynchronize(h->stream));
    cudaFree(d_pmm);
    return 0;
}

void sht_free(magic_sht *h) {
    cudaFree(h->d_fskip_syn); cudaFree(h->d_fskip_an);
    cudaFree(h->d_tab); cudaFree(h->d_off); cudaFree(h->d_clm); cudaFree(h->d_sinth); cudaFree(h->d_costh); cudaFree(h->d_wgauss);
    cudaFree(h->d_osin2); cudaFree(h->d_lm2l); cudaFree(h->d_lm2m); cudaFree(h->d_lstart); cudaFree(h->d_tw);
    if (h->stream) cudaStreamDestroy(h->stream);
}

// ------------------------------------------------------------------------------------------------------
void layout_sizes(const magic_sht *h, const BatchSpec &spec, int n_lev, Layout &L) {
    const int n_m = h->n_m, nh = h->nh, NHP = h->NHP;
    L.n_lev = n_lev;
    L.ncol_s = (int)spec.scal.size();
    L.npair_v = (int)spec.vec.size();
    L.ncol = L.ncol_s + 2 * L.npair_v;
    L.N = L.ncol ? pad_up(2 * L.ncol * n_lev, GEMM_BN) : 0;
    L.nf_s = (int)spec.afield_s.size();
    L.npair_a = (int)spec.afield_vt.size();
    L.nfa = L.nf_s + 2 * L.npair_a;
    L.an_wide = h->an_wide;
    L.Na = L.nfa ? pad_up(2 * L.nfa * n_lev, gemm_bn_an(L.an_wide)) : 0;
    L.offB.assign((size_t)n_m * 2, 0);
    L.offC.assign((size_t)n_m * 2, 0);
    long long pb = 0, pc = 0;
    for (int mc = 0; mc < n_m; mc++)
        for (int s = 0; s < 2; s++) {
            const int prob = mc * 2 + s, K = s == 0 ? h->ne1[mc] : h->no1[mc], kt = (K + BK - 1) / BK;
            L.offB[prob] = pb; pb += (long long)kt * BK * L.N;
            L.offC[prob] = pc; pc += (long long)K * L.Na;
        }
    L.szB = pb;
    L.szF = (long long)n_m * 2 * nh * L.N;
    L.szBa = (long long)n_m * 2 * NHP * L.Na;
    L.szCa = pc + (long long)GEMM_BM * L.Na;
}

int buffers_alloc(magic_sht *h, const BatchSpec &spec, const Layout &L, Buffers &b) {
    size_t plane = (size_t)2 * h->nh * h->n_phi;
    struct { double **p; long long n; } items[] = {
        {&b.B, L.szB}, {&b.F, L.szF},
        {&b.gin, (long long)(plane * spec.nfield_in * L.n_lev)}, {&b.gout, (long long)(plane * spec.nfield_out * L.n_lev)},
        {&b.Ba, L.szBa}, {&b.Ca, L.szCa},
        {&b.nl_s, (long long)2 * L.nf_s * L.n_lev * h->lm_max}, {&b.nl_v, (long long)4 * L.npair_a * L.n_lev * h->lm_max}};
    b.bytes = 0;
    for (auto &it : items) {
        size_t bytes = sizeof(double) * (size_t)std::max<long long>(it.n, 1);
        MCHECK(cudaMalloc((void **)it.p, bytes));
        MCHECK(cudaMemsetAsync(*it.p, 0, bytes, h->stream));
        b.bytes += bytes;
    }
    MCHECK(cudaMalloc((void **)&b.courmax, sizeof(unsigned long long) * 2 * std::max(L.n_lev, 1)));
    MCHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

// The same eight arrays (+ the Courant maxima) carved out of one caller-owned arena: the log-step batches (api_diag.cu) never run
// concurrently, so they share one workspace instead of holding one each.  Every array starts on a 256-byte boundary.
static void buffers_items(const magic_sht *h, const BatchSpec &spec, const Layout &L, long long n[9]) {
    const long long plane = (long long)2 * h->nh * h->n_phi;
    const long long items[9] = {L.szB, L.szF, plane * spec.nfield_in * L.n_lev, plane * spec.nfield_out * L.n_lev, L.szBa, L.szCa,
                                (long long)2 * L.nf_s * L.n_lev * h->lm_max, (long long)4 * L.npair_a * L.n_lev * h->lm_max,
                                (long long)2 * std::max(L.n_lev, 1)};
    for (int i = 0; i < 9; i++) n[i] = std::max<long long>(items[i], 1);
}

size_t buffers_bytes(const magic_sht *h, const BatchSpec &spec, const Layout &L) {
    long long n[9];
    buffers_items(h, spec, L, n);
    size_t tot = 0;
    for (int i = 0; i < 9; i++) tot += ((size_t)n[i] * 8 + 255) / 256 * 256;
    return tot;
}

void buffers_carve(const magic_sht *h, const BatchSpec &spec, const Layout &L, char *base, Buffers &b) {
    long long n[9];
    buffers_items(h, spec, L, n);
    double **ps[8] = {&b.B, &b.F, &b.gin, &b.gout, &b.Ba, &b.Ca, &b.nl_s, &b.nl_v};
    size_t off = 0;
    for (int i = 0; i < 9; i++) {
        if (i < 8) *ps[i] = reinterpret_cast<double *>(base + off);
        else b.courmax = reinterpret_cast<unsigned long long *>(base + off);
        off += ((size_t)n[i] * 8 + 255) / 256 * 256;
    }
    b.bytes = off;
}

void buffers_free(Buffers &b) {
    double *ps[] = {b.B, b.F, b.gin, b.gout, b.Ba, b.Ca, b.nl_s, b.nl_v};
    for (double *p : ps) cudaFree(p);
    cudaFree(b.courmax);
    b = Buffers();
}

void layout_free(Layout &L) {
    cudaFree(L.d_offB); cudaFree(L.d_offC); cudaFree(L.d_prep_blks);
    cudaFree(L.d_probs_syn); cudaFree(L.d_probs_an); cudaFree(L.d_tiles_syn); cudaFree(L.d_tiles_an);
    cudaFree(L.d_colrow); cudaFree(L.d_scal); cudaFree(L.d_vec); cudaFree(L.d_r2c);
    L = Layout();
}

int layout_bind(magic_sht *h, const BatchSpec &spec, Layout &L, const Buffers &buf) {
    const int n_m = h->n_m, nh = h->nh, NHP = h->NHP, n_lev = L.n_lev;
    std::vector<int2> blks;
    std::vector<GemmProb> ps, pa;
    std::vector<int2> ts, ta;
    const int mt_syn = (nh + GEMM_BM - 1) / GEMM_BM;
    for (int mc = 0; mc < n_m; mc++) {
        for (int s = 0; s < 2; s++) {
            const int prob = mc * 2 + s, K = s == 0 ? h->ne1[mc] : h->no1[mc], kt = (K + BK - 1) / BK;
            const double *P = h->d_tab + h->off[(size_t)mc * 2 + s];
            if (L.ncol) {  // ---- synthesis: F[theta_k, n] = sum_l P(l, theta_k) B[l, n]
                GemmProb g{};
                g.A0 = P;
                g.kt0 = kt;
                g.M = nh;
                g.Mlo = (h->kmin[mc] / 8) * 8;
                if (h->d_fskip_syn) g.ks0 = h->d_fskip_syn + ((size_t)mc * 2 + s) * h->FS;
                g.B = buf.B + L.offB[prob];
                g.C = buf.F + (size_t)prob * nh * L.N;
                g.ldb = g.ldc = g.Nstore = L.N;
                g.Nvalid = 2 * L.ncol * n_lev;
                const int pid = (int)ps.size();
                ps.push_back(g);
                for (int mt = 0; mt < mt_syn; mt++)  // tiles entirely below Mlo only store zeros (kernel early-out)
                    for (int nt = 0; nt < L.N / GEMM_BN; nt++) ts.push_back(make_int2(pid, (mt << 16) | nt));
                L.flops_syn += 2.0 * nh * (double)kt * BK * L.N;
            }
            if (L.nfa) {   // ---- analysis: C[l, n] = sum_k P(l, theta_k) Ba[k, n]
                GemmProb g{};
                g.A0 = P;
                g.M = K;
                g.klo = h->kmin[mc] / BK;
                g.kt0 = NHP / BK;  // absolute k-tile count; the kernel starts at max(klo, fragment minimum)
                if (h->d_fskip_an) g.ks0 = h->d_fskip_an + ((size_t)mc * 2 + s) * h->FA;
                g.B = buf.Ba + (size_t)prob * NHP * L.Na;
                g.C = buf.Ca + L.offC[prob];
                g.ldb = g.ldc = g.Nstore = L.Na;
                g.Nvalid = 2 * L.nfa * n_lev;
                const int pid = (int)pa.size();
                pa.push_back(g);
                for (int mt = 0; mt < (K + gemm_bm_an(L.an_wide) - 1) / gemm_bm_an(L.an_wide); mt++)
                    for (int nt = 0; nt < L.Na / gemm_bn_an(L.an_wide); nt++) ta.push_back(make_int2(pid, (mt << 16) | nt));
                L.flops_an += 2.0 * K * (double)(g.kt0 - g.klo) * BK * L.Na;
            }
        }
    }
    for (int mc = 0; mc < n_m; mc++)  // operand assembly: blocks of 32 degrees m .. l_max+1
        for (int jt = 0; jt < (h->l_max + 1 - mc * h->minc + 1 + 31) / 32; jt++) blks.push_back(make_int2(mc, jt));
    L.n_prep_blks = (int)blks.size();
    L.ntiles_syn = (int)ts.size(); L.ntiles_an = (int)ta.size();
    // grid field written by every synthesis column (scalar columns first, then the theta / phi columns of every pair)
    std::vector<int> cr((size_t)std::max(L.ncol, 1) * n_lev, -1);
    for (int c = 0; c < L.ncol; c++) {
        const int fld = c < L.ncol_s ? spec.field_s[c] : spec.field_v[c - L.ncol_s];
        for (int lev = 0; lev < n_lev; lev++) cr[(size_t)c * n_lev + lev] = fld < 0 ? -1 : fld * n_lev + lev;
    }
    // analysis column of every product field: scalar fields first, then (theta-type = B, phi-type = A) of every pair
    std::vector<R2cField> r2c(std::max(spec.nfield_out, 1), R2cField{0, R_NONE});
    for (int i = 0; i < L.nf_s; i++) r2c[spec.afield_s[i]] = R2cField{i, R_W};
    for (int i = 0; i < L.npair_a; i++) {
        r2c[spec.afield_vt[i]] = R2cField{L.nf_s + 2 * i, R_WS};
        r2c[spec.afield_vp[i]] = R2cField{L.nf_s + 2 * i + 1, R_WS};
    }
    // source slots the columns really read, renumbered densely: the operand assembly stages one shared-memory row per source and
    // level, so unused slots (ds, p, xi in an MHD run) would cost occupancy
    std::vector<ScalCol> scal = spec.scal;
    std::vector<VecPair> vec = spec.vec;
    int dense[MAGIC_MAX_SRC];
    bool used[MAGIC_MAX_SRC] = {false};
    auto mark = [&](const Term &t) { if (t.ftype != F_NONE) used[t.src] = true; };
    for (const auto &c : scal) { mark(c.t[0]); mark(c.t[1]); }
    for (const auto &v : vec) { mark(v.S[0]); mark(v.S[1]); mark(v.T[0]); mark(v.T[1]); }
    L.nsrc = 0;
    for (int i = 0; i < MAGIC_MAX_SRC; i++) {
        dense[i] = -1;
        if (used[i]) { dense[i] = L.nsrc; L.src_slot[L.nsrc++] = i; }
    }
    if (L.nsrc == 0) L.nsrc = -1;  // a layout whose terms are patched per call (api_sht.cu): the caller's numbering is kept
    auto remap = [&](Term &t) { if (t.ftype != F_NONE) t.src = dense[t.src]; };
    for (auto &c : scal) { remap(c.t[0]); remap(c.t[1]); }
    for (auto &v : vec) { remap(v.S[0]); remap(v.S[1]); remap(v.T[0]); remap(v.T[1]); }
    if (dev_upload_vec(&L.d_offB, L.offB) || dev_upload_vec(&L.d_offC, L.offC) || dev_upload_vec(&L.d_prep_blks, blks) ||
        dev_upload_vec(&L.d_probs_syn, ps) || dev_upload_vec(&L.d_probs_an, pa) || dev_upload_vec(&L.d_tiles_syn, ts) ||
        dev_upload_vec(&L.d_tiles_an, ta) || dev_upload_vec(&L.d_colrow, cr) ||
        dev_upload_vec(&L.d_scal, scal) || dev_upload_vec(&L.d_vec, vec) || dev_upload_vec(&L.d_r2c, r2c))
        return 1;
    return 0;
}

int layout_build(magic_sht *h, const BatchSpec &spec, int n_lev, Layout &L) {
    layout_sizes(h, spec, n_lev, L);
    return 0;
}

// ------------------------------------------------------------------------------------------------------
int run_synthesis(magic_sht *h, const BatchSpec &spec, const Layout &L, const Buffers &buf, const double *const src[MAGIC_MAX_SRC],
                  const LevelInfo *d_lev, cudaEvent_t *ev) {
    (void)spec;
    SynthPrepArgs a{};
    for (int i = 0; i < MAGIC_MAX_SRC; i++)  // dense numbering of layout_bind, or the caller's own (nsrc < 0)
        a.src[i] = L.nsrc < 0 ? src[i] : (i < L.nsrc ? src[L.src_slot[i]] : nullptr);
    a.scal = L.d_scal; a.vec = L.d_vec;
    a.ncol_s = L.ncol_s; a.npair_v = L.npair_v; a.n_lev = L.n_lev; a.lm_max = h->lm_max;
    a.N = L.N; a.lev = d_lev; a.lstart = h->d_lstart; a.clm = h->d_clm; a.l_max = h->l_max; a.minc = h->minc;
    a.B = buf.B; a.offB = L.d_offB; a.blks = L.d_prep_blks;
    a.nsrc = L.nsrc;
    if (L.nsrc < 0) {
        a.nsrc = 0;
        for (int i = 0; i < MAGIC_MAX_SRC; i++)
            if (src[i]) a.nsrc = i + 1;
    }
    if (ev) cudaEventRecord(ev[0], h->stream);
    if (L.ncol && L.n_prep_blks) {
        size_t smem = (size_t)PREP_WARPS * a.nsrc * PREP_LD * sizeof(double2);
        dim3 grid(L.n_prep_blks, (L.n_lev + PREP_WARPS - 1) / PREP_WARPS);
        synth_prep_kernel<<<grid, PREP_WARPS * 32, smem, h->stream>>>(a);
        h->launches++;
    }
    if (ev) cudaEventRecord(ev[1], h->stream);
    launch_legendre_gemm(false, L.d_probs_syn, L.d_tiles_syn, L.ntiles_syn, h->NHP, h->stream);
    h->launches++;
    if (ev) cudaEventRecord(ev[2], h->stream);
    if (L.ncol) {
        launch_fft_c2r(h->fft, buf.F, L.N, h->n_m, h->nh, L.ncol * L.n_lev, L.d_colrow, buf.gin, h->stream);
        h->launches++;
    }
    if (ev) cudaEventRecord(ev[3], h->stream);
    MCHECK(cudaGetLastError());
    return 0;
}

static ExtractArgs extract_args(const magic_sht *h, const Layout &L, const Buffers &buf, const LevelInfo *d_lev) {
    ExtractArgs e{};
    e.C = buf.Ca; e.offC = L.d_offC; e.N = L.Na; e.n_lev = L.n_lev; e.lm_max = h->lm_max; e.nf_s = L.nf_s; e.npair = L.npair_a;
    e.lm2l = h->d_lm2l; e.lm2m = h->d_lm2m; e.lstart = h->d_lstart; e.clm = h->d_clm; e.minc = h->minc; e.lev = d_lev;
    e.out_s = buf.nl_s; e.out_v = buf.nl_v;
    return e;
}
ExtractArgs make_extract_args(const magic_sht *h, const Layout &L, const Buffers &buf, const LevelInfo *d_lev) {
    return extract_args(h, L, buf, d_lev);
}

int run_analysis(magic_sht *h, const BatchSpec &spec, const Layout &L, const Buffers &buf, const LevelInfo *d_lev, cudaEvent_t *ev,
                 bool extract) {
    if (spec.nfield_out == 0) return 0;
    R2cArgs a{};
    a.grid = buf.gout; a.n_lev = L.n_lev; a.nh = h->nh; a.n_m = h->n_m; a.NHP = h->NHP;
    a.wgauss = h->d_wgauss; a.osin2 = h->d_osin2; a.fields = L.d_r2c;
    a.B = buf.Ba; a.ldB = L.Na; a.minc = h->minc;
    if (ev) cudaEventRecord(ev[0], h->stream);
    launch_fft_r2c(h->fft, a, spec.nfield_out, h->stream);
    h->launches++;
    if (ev) cudaEventRecord(ev[1], h->stream);
    launch_legendre_gemm(true, L.d_probs_an, L.d_tiles_an, L.ntiles_an, h->NHP, h->stream, L.an_wide);
    h->launches++;
    if (ev) cudaEventRecord(ev[2], h->stream);
    if (!extract) { MCHECK(cudaGetLastError()); return 0; }
    const ExtractArgs e = extract_args(h, L, buf, d_lev);
    dim3 g2((h->lm_max + 255) / 256, L.n_lev);
    anal_extract_kernel<<<g2, 256, 0, h->stream>>>(e);
    h->launches++;
    MCHECK(cudaGetLastError());
    return 0;
}

}  // namespace magic
