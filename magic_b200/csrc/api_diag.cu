// api_diag.cu -- C ABI: the in-loop diagnostics of log steps (rIter.f90:303-373) and the grid fields
// graphOut_mpi reads (rIter.f90:303-314).
//
// The reference evaluates get_helicity / get_hemi / get_visc_heat / get_nlBLayers / get_fluxes / get_perpPar level by level on the
// grid arrays the loop has just synthesised, and synthesises extra fields for them (the velocity gradients in curl-form runs,
// pressure, entropy gradient: rIter.f90:483-527).  Here a log step runs a second, synthesis-only column program over the same
// spectral inputs -- exactly the fields the requested diagnostics read, with lDeriv = .true. on the boundary levels as the
// reference sets it (rIter.f90:193-205) -- followed by one fused reduction kernel per level chunk.  The hot path of ordinary
// steps is untouched; nothing but [n_r_loc][MAGIC_NDIAG] doubles crosses PCIe (instead of 20+ grid fields per level).
// Included by lib.cu after api_rloop.cu (it uses magic_rloop).
#include "kernels_diag.cuh"

// ---- the shared workspace of the batches in this file (magic_rloop::aux) ------------------------------------------------------
// level chunk of a batch: bounded by the memory that is free or already held by the shared workspace
static int aux_chunk(const magic_rloop *rl, double bytes_per_level) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
    const double avail = (double)free_b + (double)rl->aux.bytes;
    int chunk = (int)std::min<double>(32.0, 0.5 * avail / bytes_per_level);
    return std::max(1, std::min(chunk, rl->n_r_loc));
}

// binds a batch to the workspace before it runs: grows the arena if this batch needs more (every batch then re-binds its
// descriptors on its next run), carves the batch's arrays, and clears them when another batch used the arena last (the transform
// kernels rely on zero padding, as after buffers_alloc)
static int aux_acquire(magic_rloop *rl, const void *pipe, const BatchSpec &S, int chunk, Layout &lay, Buffers &buf, int &gen) {
    magic_sht *h = rl->h;
    if (lay.n_lev != chunk) layout_sizes(h, S, chunk, lay);
    const size_t need = buffers_bytes(h, S, lay);
    auto &A = rl->aux;
    if (need > A.bytes) {
        MCHECK(cudaStreamSynchronize(h->stream));
        cudaFree(A.p);
        A.p = nullptr; A.bytes = 0; A.gen++; A.owner = nullptr;
        if (cudaMalloc((void **)&A.p, need) != cudaSuccess) {
            cudaGetLastError();
            MFAIL("log-step batch: no device memory for its workspace");
        }
        A.bytes = need;
    }
    if (gen != A.gen) {
        layout_free(lay);
        layout_sizes(h, S, chunk, lay);
        buffers_carve(h, S, lay, A.p, buf);
        if (layout_bind(h, S, lay, buf)) return 1;
        gen = A.gen;
    }
    if (A.owner != pipe) {
        MCHECK(cudaMemsetAsync(A.p, 0, need, h->stream));
        A.owner = pipe;
    }
    return 0;
}

struct DiagPipe {
    int mask = -1, chunk = 0, gx = 0, phi_field = -1;
    BatchSpec spec;
    DiagIn di;
    Layout lay;
    Buffers buf;   // carved from magic_rloop::aux by aux_acquire
    int gen = -1;
    LevelInfo *d_lev = nullptr;
    double *d_gauss = nullptr, *d_means = nullptr, *d_partial = nullptr, *d_partial_ph = nullptr, *d_out = nullptr, *h_out = nullptr;
    double *d_src[S_COUNT] = {nullptr};  // staging of the host-pointer call: complex [chunk][lm_max] per source
    bool need[S_COUNT] = {false};
};

static void diag_free(DiagPipe *d) {
    if (!d) return;
    layout_free(d->lay);
    cudaFree(d->d_lev); cudaFree(d->d_gauss); cudaFree(d->d_means); cudaFree(d->d_partial); cudaFree(d->d_partial_ph); cudaFree(d->d_out);
    if (d->h_out) cudaFreeHost(d->h_out);
    for (int i = 0; i < S_COUNT; i++) cudaFree(d->d_src[i]);
    delete d;
}

// column program of transform_to_grid_space with the output branches taken (rIter.f90:483-622)
static int diag_build(magic_rloop *rl, int mask) {
    magic_sht *h = rl->h;
    const magic_params &P = rl->p;
    if (P.l_full_sphere) MFAIL("magic_rloop_diagnostics: full-sphere runs (v_center_sphere on the r = 0 level) are not supported");
    if (!(P.l_conv || P.l_mag_kin)) MFAIL("magic_rloop_diagnostics: needs a flow (l_conv or l_mag_kin)");
    if (rl->aux.owner == rl->diag) rl->aux.owner = nullptr;  // the next DiagPipe may get the same address: never trust stale contents
    diag_free(rl->diag);
    rl->diag = nullptr;
    DiagPipe *d = new DiagPipe();
    rl->diag = d;
    d->mask = mask;
    BatchSpec &S = d->spec;
    int *dip = (int *)&d->di;
    for (int i = 0; i < DIAG_NF; i++) dip[i] = -1;
    DiagIn &di = d->di;
    const Term N_ = {0, F_NONE};
    const bool flux = mask & DM_FLUX, viscbc = mask & DM_VISCBC;
    const bool grads = mask & (DM_HEL | DM_POWER | DM_FLUX | DM_VISCBC);  // hemi / perpPar / phase read vr, vt, vp only
    if ((mask & DM_PHASE) && !P.l_phase_field) MFAIL("magic_rloop_diagnostics: MAGIC_DIAG_PHASE needs l_phase_field");
    const bool mag = (P.l_mag || P.l_mag_LF) && ((mask & DM_HEMI) || flux);
    int nf = 0;
    auto need = [&](std::initializer_list<int> s) { for (int i : s) d->need[i] = true; };
    if (P.l_heat && (flux || viscbc)) {
        add_scal(S, Term{S_S, F_ONE}, N_, LM_ALL, nf, di.s); need({S_S});
        if (viscbc) {
            add_pair(S, Term{S_S, F_ONE}, N_, N_, N_, LM_ALL, nf, di.dsdt, di.dsdp);   // scal_to_grad_spat, rIter.f90:487
            add_scal(S, Term{S_DS, F_ONE}, N_, LM_ALL, nf, di.drs); need({S_DS});       // rIter.f90:511-513
        }
    }
    if (flux && !P.l_anelastic_liquid) { add_scal(S, Term{S_P, F_ONE}, N_, LM_ALL, nf, di.p); need({S_P}); }  // lPressCalc, rIter.f90:503
    need({S_W, S_DW, S_Z});
    add_scal(S, Term{S_W, F_DLH}, N_, LM_VEL, nf, di.vr);
    add_pair(S, Term{S_DW, F_ONE}, N_, Term{S_Z, F_ONE}, N_, LM_VEL, nf, di.vt, di.vp);
    if (grads) {
        need({S_DDW, S_DZ});
        add_scal(S, Term{S_DW, F_DLH}, N_, LM_ALL, nf, di.dvrdr);
        add_pair(S, Term{S_DDW, F_ONE}, N_, Term{S_DZ, F_ONE}, N_, LM_ALL, nf, di.dvtdr, di.dvpdr);
        add_scal(S, Term{S_Z, F_DLH}, N_, LM_VEL, nf, di.cvr);
        add_pair(S, Term{S_W, F_DLH}, N_, N_, N_, LM_VELBULK, nf, di.dvrdt, di.dvrdp);
        add_pair(S, Term{S_DW, F_IM}, N_, Term{S_Z, F_IM}, N_, LM_VEL, nf, di.dvtdp, di.dvpdp);
    }
    if (mask & DM_PHASE) { add_scal(S, Term{S_PHI, F_ONE}, N_, LM_ALL, nf, d->phi_field); need({S_PHI}); }  // rIter.f90:509
    if (mag) {
        need({S_B, S_DB, S_AJ});
        add_scal(S, Term{S_B, F_DLH}, N_, LM_ALL, nf, di.br);
        add_pair(S, Term{S_DB, F_ONE}, N_, Term{S_AJ, F_ONE}, N_, LM_ALL, nf, di.bt, di.bp);
        if (flux && P.l_mag_nl) {
            need({S_DDB, S_DJ});
            add_pair(S, Term{S_DJ, F_ONE}, N_, Term{S_B, F_OR2DLH}, Term{S_DDB, F_NEG}, LM_ALL, nf, di.cbt, di.cbp);
        }
    }
    S.nfield_in = nf;
    S.nfield_out = 0;
    // level chunk: bounded by free memory (grid + (theta,m) space + operands are about 2.2 grid fields per synthesised field)
    const double per_level = 2.2 * 8.0 * (double)h->n_theta * h->n_phi * nf + 64.0 * h->lm_max * S_COUNT;
    const int chunk = aux_chunk(rl, per_level);
    d->chunk = chunk;   // the workspace itself is bound by aux_acquire at run time
    // the diagnostics' levels: lDeriv = .true. everywhere (rIter.f90:193-205), boundary levels bulk with lRmsCalc (:215)
    std::vector<LevelInfo> lev = rl->lev;
    for (auto &L : lev) {
        L.lDeriv = 1;
        if (mask & 256) L.nBc = 0;
    }
    if (dev_upload_vec(&d->d_lev, lev)) return 1;
    std::vector<double> ga(h->nh);
    for (int k = 0; k < h->nh; k++) ga[k] = h->gauss[k];
    if (dev_upload_vec(&d->d_gauss, ga)) return 1;
    const size_t plane = (size_t)h->nh * h->n_phi;
    d->gx = (int)std::min<size_t>((plane + DIAG_THREADS - 1) / DIAG_THREADS, 4 * 148);
    MCHECK(cudaMalloc((void **)&d->d_means, sizeof(double) * DIAG_NMEAN * chunk * 2 * h->nh));
    MCHECK(cudaMalloc((void **)&d->d_partial, sizeof(double) * (size_t)chunk * d->gx * DIAG_NSLOT));
    MCHECK(cudaMalloc((void **)&d->d_partial_ph, sizeof(double) * (size_t)chunk * d->gx * DIAG_NPHASE));
    MCHECK(cudaMalloc((void **)&d->d_out, sizeof(double) * (size_t)rl->n_r_loc * DIAG_NOUT));
    MCHECK(cudaMemset(d->d_out, 0, sizeof(double) * (size_t)rl->n_r_loc * DIAG_NOUT));
    MCHECK(cudaMallocHost((void **)&d->h_out, sizeof(double) * (size_t)rl->n_r_loc * DIAG_NOUT));
    return 0;
}

static int diag_run(magic_rloop *rl, const magic_fields_in *in, int mask, int ktops, int kbots, double *out, bool host_in) {
    if (!rl || !in || !out) MFAIL("magic_rloop_diagnostics: null argument");
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    if (!rl->diag || rl->diag->mask != mask)
        if (diag_build(rl, mask)) { if (rl->aux.owner == rl->diag) rl->aux.owner = nullptr; diag_free(rl->diag); rl->diag = nullptr; return 1; }
    DiagPipe *d = rl->diag;
    if (aux_acquire(rl, d, d->spec, d->chunk, d->lay, d->buf, d->gen)) return 1;
    const magic_params &P = rl->p;
    const double *ip[S_COUNT];
    in_ptrs(in, ip);
    for (int i = 0; i < S_COUNT; i++)
        if (d->need[i] && !ip[i]) MFAIL("magic_rloop_diagnostics: a required input field is null");
    const size_t lm2 = 2 * (size_t)h->lm_max;
    const int nl = d->chunk, n_r = rl->n_r_loc;
    if (host_in)
        for (int i = 0; i < S_COUNT; i++)
            if (d->need[i] && !d->d_src[i]) MCHECK(cudaMalloc((void **)&d->d_src[i], sizeof(double) * lm2 * nl));
    DiagArgs a{};
    a.di = d->di; a.gin = d->buf.gin; a.n_lev = nl; a.nh = h->nh; a.n_phi = h->n_phi; a.mask = mask & 63;
    a.l_mag = (P.l_mag && d->di.br >= 0) ? 1 : 0; a.l_mag_nl = (P.l_mag_nl && d->di.cbt >= 0) ? 1 : 0;
    a.n_r_max = P.n_r_max; a.ktops = ktops; a.kbots = kbots;
    a.omega_ma = P.omega_ma; a.omega_ic = P.omega_ic; a.r_cmb = P.r_cmb; a.r_icb = P.r_icb;
    a.sinth = h->d_sinth; a.costh = h->d_costh; a.gauss = d->d_gauss;
    const int mf[DIAG_NMEAN] = {d->di.vr, d->di.cvr, d->di.vt, d->di.vp, d->di.dvrdp, d->di.dvpdr, d->di.dvtdr, d->di.dvrdt};
    for (int i = 0; i < DIAG_NMEAN; i++) a.mean_field[i] = mf[i];
    a.means = d->d_means; a.partial = d->d_partial;
    // chunks of exactly `nl` levels; the last one is shifted back so that it ends on the last level (its overlap is recomputed)
    for (int l0 = 0; l0 < n_r; l0 += nl) {
        const int s0 = std::min(l0, n_r - nl);
        const double *src[MAGIC_MAX_SRC];
        for (int i = 0; i < MAGIC_MAX_SRC; i++) src[i] = nullptr;
        for (int i = 0; i < S_COUNT; i++) {
            if (!d->need[i]) continue;
            if (host_in) {
                MCHECK(cudaMemcpyAsync(d->d_src[i], ip[i] + (size_t)s0 * lm2, sizeof(double) * lm2 * nl, cudaMemcpyHostToDevice, h->stream));
                src[i] = d->d_src[i];
            } else {
                src[i] = ip[i] + (size_t)s0 * lm2;
            }
        }
        if (run_synthesis(h, d->spec, d->lay, d->buf, src, d->d_lev + s0, nullptr)) return 1;
        a.lev = d->d_lev + s0;
        if (mask & (DM_HEL | DM_PERPPAR)) {
            const size_t nrow = (size_t)DIAG_NMEAN * nl * 2 * h->nh;
            diag_mean_kernel<<<(unsigned)((nrow + DIAG_THREADS / 32 - 1) / (DIAG_THREADS / 32)), DIAG_THREADS, 0, h->stream>>>(a);
            h->launches++;
        }
        diag_kernel<<<dim3(d->gx, nl), DIAG_THREADS, 0, h->stream>>>(a);
        diag_finish_kernel<<<(nl * DIAG_NSLOT + 127) / 128, 128, 0, h->stream>>>(d->d_partial, d->gx, nl, d->d_out + (size_t)s0 * DIAG_NOUT);
        h->launches += 2;
        if (mask & DM_PHASE) {
            diag_phase_kernel<<<dim3(d->gx, nl), DIAG_THREADS, 0, h->stream>>>(a, d->phi_field, d->d_partial_ph);
            diag_phase_finish_kernel<<<(nl * DIAG_NPHASE + 127) / 128, 128, 0, h->stream>>>(d->d_partial_ph, d->gx, nl, d->d_out + (size_t)s0 * DIAG_NOUT);
            h->launches += 2;
        }
        MCHECK(cudaGetLastError());
    }
    MCHECK(cudaMemcpyAsync(d->h_out, d->d_out, sizeof(double) * (size_t)n_r * DIAG_NOUT, cudaMemcpyDeviceToHost, h->stream));
    MCHECK(cudaStreamSynchronize(h->stream));
    memcpy(out, d->h_out, sizeof(double) * (size_t)n_r * DIAG_NOUT);
    return 0;
}

extern "C" int magic_rloop_diagnostics(magic_rloop *rl, const magic_fields_in *in, int mask, int ktops, int kbots, double *out) {
    return diag_run(rl, in, mask, ktops, kbots, out, true);
}
extern "C" int magic_rloop_diagnostics_dev(magic_rloop *rl, const magic_fields_in *in, int mask, int ktops, int kbots, double *out) {
    return diag_run(rl, in, mask, ktops, kbots, out, false);
}

// ---- graphOut_mpi's inputs (rIter.f90:303-314): vr, vt, vp, [br, bt, bp,] sr, [pr] of one local level on the grid, in the
// reference layout f(nlat_padded, n_phi) with N/S-interleaved rows; the host writes them to the graphic file as it does today
// (out_graph_file.f90:337).  Uses the per-call transforms, so boundary levels get the values transform_to_grid_space produces
// for bulk levels (graphOut is called before the boundary overrides matter: rigid walls hold v = wall motion, which the host
// fills in as before).
extern "C" int magic_rloop_graph_fields(magic_rloop *rl, const magic_fields_in *in, int level, double *vr, double *vt, double *vp, double *br,
                                        double *bt, double *bp, double *sr, double *pr) {
    if (!rl || !in) MFAIL("magic_rloop_graph_fields: null argument");
    if (level < 0 || level >= rl->n_r_loc) MFAIL("magic_rloop_graph_fields: level out of range");
    magic_sht *h = rl->h;
    const size_t off = 2 * (size_t)h->lm_max * level;
    const int lcut = rl->lev[level].lcut;
    if (vr && vt && vp) {
        if (!in->w || !in->dw || !in->z) MFAIL("magic_rloop_graph_fields: w, dw, z needed");
        if (magic_torpol_to_spat(h, in->w + off, in->dw + off, in->z + off, vr, vt, vp, lcut)) return 1;
    }
    if (br && bt && bp) {
        if (!in->b || !in->db || !in->aj) MFAIL("magic_rloop_graph_fields: b, db, aj needed");
        if (magic_torpol_to_spat(h, in->b + off, in->db + off, in->aj + off, br, bt, bp, lcut)) return 1;
    }
    if (sr) {
        if (!in->s) MFAIL("magic_rloop_graph_fields: s needed");
        if (magic_scal_to_spat(h, in->s + off, sr, lcut)) return 1;
    }
    if (pr) {
        if (!in->p) MFAIL("magic_rloop_graph_fields: p needed");
        if (magic_scal_to_spat(h, in->p + off, pr, lcut)) return 1;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// get_dtBLM (rIter.f90:392-395, dtB.f90:144-223): with l_dtB the reference forms eleven grid products of
// (vr, vt, vp, br, bt, bp) on every level and analyses them (2 spat_to_sphertor + 7 scal_to_SH, lcut = l_max).  Here: one more
// column program on the batched pipeline -- 6 synthesised fields, dtb_product_kernel, 11 analysed fields -- so the extra
// transforms ride on the Legendre GEMM / FFT kernels of the hot path.  out: HOST complex [11][n_r_loc][lm_max] in the order
// BtVrLM, BpVrLM, BrVtLM, BrVpLM, BtVpLM, BpVtLM, BpVtBtVpCotLM, BpVtBtVpSn2LM, BrVZLM, BtVZLM, BtVZsn2LM.
struct DtbPipe {
    int chunk = 0, gx = 0;
    BatchSpec spec;
    Layout lay;
    Buffers buf;   // carved from magic_rloop::aux by aux_acquire
    int gen = -1;
    LevelInfo *d_lev_an = nullptr;  // lcut = l_max for the analysis (dtB.f90:210-221)
    double *d_means = nullptr;
    double *d_src[S_COUNT] = {nullptr};
};

static void dtb_free(DtbPipe *d) {
    if (!d) return;
    layout_free(d->lay);
    cudaFree(d->d_lev_an); cudaFree(d->d_means);
    for (int i = 0; i < S_COUNT; i++) cudaFree(d->d_src[i]);
    delete d;
}

static int dtb_build(magic_rloop *rl) {
    magic_sht *h = rl->h;
    const magic_params &P = rl->p;
    if (P.l_full_sphere) MFAIL("magic_rloop_dtb: full-sphere runs are not supported");
    if (!(P.l_mag || P.l_mag_LF)) MFAIL("magic_rloop_dtb: needs a magnetic field");
    DtbPipe *d = new DtbPipe();
    rl->dtb = d;
    BatchSpec &S = d->spec;
    const Term N_ = {0, F_NONE};
    int nf = 0, f0, f1, f2;
    add_scal(S, Term{S_W, F_DLH}, N_, LM_VEL, nf, f0);
    add_pair(S, Term{S_DW, F_ONE}, N_, Term{S_Z, F_ONE}, N_, LM_VEL, nf, f1, f2);
    add_scal(S, Term{S_B, F_DLH}, N_, LM_ALL, nf, f0);
    add_pair(S, Term{S_DB, F_ONE}, N_, Term{S_AJ, F_ONE}, N_, LM_ALL, nf, f1, f2);
    S.nfield_in = nf;  // vr, vt, vp, br, bt, bp = grid fields 0..5
    S.afield_vt = {0, 2};
    S.afield_vp = {1, 3};
    S.afield_s = {4, 5, 6, 7, 8, 9, 10};
    S.nfield_out = 11;
    const double per_level = 2.2 * 8.0 * (double)h->n_theta * h->n_phi * (6 + 11) + 64.0 * h->lm_max * 17;
    const int chunk = aux_chunk(rl, per_level);
    d->chunk = chunk;   // the workspace itself is bound by aux_acquire at run time
    std::vector<LevelInfo> lev = rl->lev;
    for (auto &L : lev) L.lcut = h->l_max;
    if (dev_upload_vec(&d->d_lev_an, lev)) return 1;
    const size_t plane = (size_t)h->nh * h->n_phi;
    d->gx = (int)std::min<size_t>((plane + DIAG_THREADS - 1) / DIAG_THREADS, 8 * 148);
    MCHECK(cudaMalloc((void **)&d->d_means, sizeof(double) * DIAG_NMEAN * chunk * 2 * h->nh));
    return 0;
}

static int dtb_run(magic_rloop *rl, const magic_fields_in *in, double *out, bool host_in) {
    if (!rl || !in || !out) MFAIL("magic_rloop_dtb: null argument");
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    if (!rl->dtb)
        if (dtb_build(rl)) { dtb_free(rl->dtb); rl->dtb = nullptr; return 1; }
    DtbPipe *d = rl->dtb;
    if (aux_acquire(rl, d, d->spec, d->chunk, d->lay, d->buf, d->gen)) return 1;
    const magic_params &P = rl->p;
    const double *ip[S_COUNT];
    in_ptrs(in, ip);
    const int needed[6] = {S_W, S_DW, S_Z, S_B, S_DB, S_AJ};
    for (int i : needed)
        if (!ip[i]) MFAIL("magic_rloop_dtb: w, dw, z, b, db, aj are needed");
    const size_t lm2 = 2 * (size_t)h->lm_max;
    const int nl = d->chunk, n_r = rl->n_r_loc;
    if (host_in)
        for (int i : needed)
            if (!d->d_src[i]) MCHECK(cudaMalloc((void **)&d->d_src[i], sizeof(double) * lm2 * nl));
    DiagArgs ma{};   // phi mean of vp through the diagnostics' mean kernel
    ma.gin = d->buf.gin; ma.n_lev = nl; ma.nh = h->nh; ma.n_phi = h->n_phi;
    for (int i = 0; i < DIAG_NMEAN; i++) ma.mean_field[i] = -1;
    ma.mean_field[0] = 2;
    ma.means = d->d_means;
    DtbArgs a{};
    a.gin = d->buf.gin; a.gout = d->buf.gout; a.n_lev = nl; a.nh = h->nh; a.n_phi = h->n_phi;
    a.omega_ma = P.omega_ma; a.omega_ic = P.omega_ic; a.r_cmb = P.r_cmb; a.r_icb = P.r_icb;
    a.sinth = h->d_sinth; a.costh = h->d_costh; a.means = d->d_means;
    for (int l0 = 0; l0 < n_r; l0 += nl) {
        const int s0 = std::min(l0, n_r - nl);
        const double *src[MAGIC_MAX_SRC];
        for (int i = 0; i < MAGIC_MAX_SRC; i++) src[i] = nullptr;
        for (int i : needed) {
            if (host_in) {
                MCHECK(cudaMemcpyAsync(d->d_src[i], ip[i] + (size_t)s0 * lm2, sizeof(double) * lm2 * nl, cudaMemcpyHostToDevice, h->stream));
                src[i] = d->d_src[i];
            } else {
                src[i] = ip[i] + (size_t)s0 * lm2;
            }
        }
        if (run_synthesis(h, d->spec, d->lay, d->buf, src, rl->d_lev + s0, nullptr)) return 1;
        a.lev = rl->d_lev + s0;
        const size_t nrow = (size_t)DIAG_NMEAN * nl * 2 * h->nh;
        diag_mean_kernel<<<(unsigned)((nrow + DIAG_THREADS / 32 - 1) / (DIAG_THREADS / 32)), DIAG_THREADS, 0, h->stream>>>(ma);
        dtb_product_kernel<<<dim3(d->gx, nl), DIAG_THREADS, 0, h->stream>>>(a);
        h->launches += 2;
        if (run_analysis(h, d->spec, d->lay, d->buf, d->d_lev_an + s0, nullptr, true)) return 1;
        // nl_v: [S0, T0, S1, T1][nl][lm_max] -> out 0..3; nl_s: [7][nl][lm_max] -> out 4..10
        for (int q = 0; q < 11; q++) {
            const double *srcq = q < 4 ? d->buf.nl_v + (size_t)q * nl * lm2 : d->buf.nl_s + (size_t)(q - 4) * nl * lm2;
            MCHECK(cudaMemcpyAsync(out + ((size_t)q * n_r + s0) * lm2, srcq, sizeof(double) * lm2 * nl, cudaMemcpyDeviceToHost, h->stream));
        }
        MCHECK(cudaStreamSynchronize(h->stream));  // the staging buffers are reused by the next chunk
    }
    return 0;
}

extern "C" int magic_rloop_dtb(magic_rloop *rl, const magic_fields_in *in, double *out) { return dtb_run(rl, in, out, true); }
extern "C" int magic_rloop_dtb_dev(magic_rloop *rl, const magic_fields_in *in, double *out) { return dtb_run(rl, in, out, false); }

// ------------------------------------------------------------------------------------------------------
// Torsional-oscillation sums (rIter.f90:395-404; TO.f90:141-352): on lTONext steps the reference keeps three grid fields of every
// level (BsLast, BpLast, BzLast), on lTOCalc steps it sweeps eleven grid fields per level for the azimuthal means of twenty
// products.  Here both are one more synthesis-only column program (vr, vt, vp, cvr, dvpdr, br, bt, bp, cbr, cbt, phi with the
// boundary treatment of the diagnostics) followed by to_next_kernel / to_kernel; the "Last" fields stay on the device between
// the two calls; [n_r_loc][MAGIC_NTO][n_theta] doubles cross PCIe.  The spectral, axisymmetric part of getTOnext / getTOfinish
// (TO.f90:322-329, 344-388: O(l_max) per level) stays with the host, its three get_PAS transforms are magic_toraxi_to_spat calls.
struct ToPipe {
    int chunk = 0;
    BatchSpec spec;
    ToIn ti;
    Layout lay;
    Buffers buf;   // carved from magic_rloop::aux by aux_acquire
    int gen = -1;
    LevelInfo *d_lev = nullptr;
    double *d_last = nullptr, *d_out = nullptr, *h_out = nullptr;
    double *d_src[S_COUNT] = {nullptr};
    bool need[S_COUNT] = {false};
    bool have_last = false;
};

static void to_free(ToPipe *d) {
    if (!d) return;
    layout_free(d->lay);
    cudaFree(d->d_lev); cudaFree(d->d_last); cudaFree(d->d_out);
    if (d->h_out) cudaFreeHost(d->h_out);
    for (int i = 0; i < S_COUNT; i++) cudaFree(d->d_src[i]);
    delete d;
}

static int to_build(magic_rloop *rl) {
    magic_sht *h = rl->h;
    const magic_params &P = rl->p;
    if (P.l_full_sphere) MFAIL("magic_rloop_to: full-sphere runs are not supported");
    if (!(P.l_conv || P.l_mag_kin)) MFAIL("magic_rloop_to: needs a flow (l_conv or l_mag_kin)");
    ToPipe *d = new ToPipe();
    rl->to = d;
    BatchSpec &S = d->spec;
    int *tip = (int *)&d->ti;
    for (int i = 0; i < TO_NF; i++) tip[i] = -1;
    ToIn &ti = d->ti;
    const Term N_ = {0, F_NONE};
    int nf = 0, dvtdr = -1, cbp = -1;
    auto need = [&](std::initializer_list<int> s) { for (int i : s) d->need[i] = true; };
    need({S_W, S_DW, S_DDW, S_Z, S_DZ});
    add_scal(S, Term{S_W, F_DLH}, N_, LM_VEL, nf, ti.vr);
    add_pair(S, Term{S_DW, F_ONE}, N_, Term{S_Z, F_ONE}, N_, LM_VEL, nf, ti.vt, ti.vp);
    add_scal(S, Term{S_Z, F_DLH}, N_, LM_VEL, nf, ti.cvr);
    add_pair(S, Term{S_DDW, F_ONE}, N_, Term{S_DZ, F_ONE}, N_, LM_ALL, nf, dvtdr, ti.dvpdr);
    if (P.l_mag) {
        need({S_B, S_DB, S_DDB, S_AJ, S_DJ});
        add_scal(S, Term{S_B, F_DLH}, N_, LM_ALL, nf, ti.br);
        add_pair(S, Term{S_DB, F_ONE}, N_, Term{S_AJ, F_ONE}, N_, LM_ALL, nf, ti.bt, ti.bp);
        add_scal(S, Term{S_AJ, F_DLH}, N_, LM_ALL, nf, ti.cbr);
        add_pair(S, Term{S_DJ, F_ONE}, N_, Term{S_B, F_OR2DLH}, Term{S_DDB, F_NEG}, LM_ALL, nf, ti.cbt, cbp);
    }
    if (P.l_phase_field) { add_scal(S, Term{S_PHI, F_ONE}, N_, LM_ALL, nf, ti.phi); need({S_PHI}); }
    S.nfield_in = nf;
    S.nfield_out = 0;
    const size_t plane = (size_t)h->nh * h->n_phi;
    if (P.l_mag) {
        cudaError_t e = cudaMalloc((void **)&d->d_last, sizeof(double) * 6 * plane * (size_t)rl->n_r_loc);
        if (e != cudaSuccess) { cudaGetLastError(); MFAIL("magic_rloop_to: no device memory for BsLast / BpLast / BzLast of all local levels"); }
        MCHECK(cudaMemset(d->d_last, 0, sizeof(double) * 6 * plane * (size_t)rl->n_r_loc));
    }
    const double per_level = 2.2 * 8.0 * (double)h->n_theta * h->n_phi * nf + 64.0 * h->lm_max * S_COUNT;
    const int chunk = aux_chunk(rl, per_level);
    d->chunk = chunk;   // the workspace itself is bound by aux_acquire at run time
    std::vector<LevelInfo> lev = rl->lev;   // lDeriv = .true. with lTOCalc (rIter.f90:197-205)
    for (auto &L : lev) L.lDeriv = 1;
    if (dev_upload_vec(&d->d_lev, lev)) return 1;
    const size_t nout = (size_t)rl->n_r_loc * TO_NOUT * h->n_theta;
    MCHECK(cudaMalloc((void **)&d->d_out, sizeof(double) * nout));
    MCHECK(cudaMallocHost((void **)&d->h_out, sizeof(double) * nout));
    return 0;
}

// mode 0: getTOnext (keep the cylindrical field components), mode 1: getTO
static int to_run(magic_rloop *rl, const magic_fields_in *in, int mode, double dtLast, double *out, bool host_in) {
    if (!rl || !in || (mode == 1 && !out)) MFAIL("magic_rloop_to: null argument");
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    if (!rl->to)
        if (to_build(rl)) { to_free(rl->to); rl->to = nullptr; return 1; }
    ToPipe *d = rl->to;
    const magic_params &P = rl->p;
    if (mode == 0 && !P.l_mag) return 0;  // TO.f90:330: only the magnetic terms keep grid fields
    if (aux_acquire(rl, d, d->spec, d->chunk, d->lay, d->buf, d->gen)) return 1;
    if (mode == 1 && !(dtLast > 0.0)) MFAIL("magic_rloop_to: dtLast must be positive");
    const double *ip[S_COUNT];
    in_ptrs(in, ip);
    for (int i = 0; i < S_COUNT; i++)
        if (d->need[i] && !ip[i]) MFAIL("magic_rloop_to: a required input field is null");
    const size_t lm2 = 2 * (size_t)h->lm_max, plane = (size_t)h->nh * h->n_phi;
    const int nl = d->chunk, n_r = rl->n_r_loc;
    if (host_in)
        for (int i = 0; i < S_COUNT; i++)
            if (d->need[i] && !d->d_src[i]) MCHECK(cudaMalloc((void **)&d->d_src[i], sizeof(double) * lm2 * nl));
    ToArgs a{};
    a.ti = d->ti; a.gin = d->buf.gin; a.n_lev = nl; a.nh = h->nh; a.n_phi = h->n_phi; a.n_theta = h->n_theta;
    a.l_mag = P.l_mag ? 1 : 0; a.l_phase_field = P.l_phase_field ? 1 : 0;
    a.omega_ma = P.omega_ma; a.omega_ic = P.omega_ic; a.r_cmb = P.r_cmb; a.r_icb = P.r_icb; a.CorFac = P.CorFac;
    a.pen = P.l_phase_field ? 1.0 / (P.epsPhase * P.epsPhase) / (P.penaltyFac * P.penaltyFac) : 0.0;
    a.o_dtLast = mode == 1 ? 1.0 / dtLast : 0.0;
    a.sinth = h->d_sinth; a.costh = h->d_costh;
    const int gx = (int)std::min<size_t>((plane + DIAG_THREADS - 1) / DIAG_THREADS, 8 * 148);
    for (int l0 = 0; l0 < n_r; l0 += nl) {
        const int s0 = std::min(l0, n_r - nl);
        const double *src[MAGIC_MAX_SRC];
        for (int i = 0; i < MAGIC_MAX_SRC; i++) src[i] = nullptr;
        for (int i = 0; i < S_COUNT; i++) {
            if (!d->need[i]) continue;
            if (host_in) {
                MCHECK(cudaMemcpyAsync(d->d_src[i], ip[i] + (size_t)s0 * lm2, sizeof(double) * lm2 * nl, cudaMemcpyHostToDevice, h->stream));
                src[i] = d->d_src[i];
            } else {
                src[i] = ip[i] + (size_t)s0 * lm2;
            }
        }
        if (run_synthesis(h, d->spec, d->lay, d->buf, src, d->d_lev + s0, nullptr)) return 1;
        a.lev = d->d_lev + s0;
        a.last = d->d_last ? d->d_last + (size_t)s0 * 6 * plane : nullptr;
        a.out = d->d_out + (size_t)s0 * TO_NOUT * h->n_theta;
        if (mode == 0) to_next_kernel<<<dim3(gx, nl), DIAG_THREADS, 0, h->stream>>>(a);
        else to_kernel<<<dim3(h->nh, nl), TO_THREADS, 0, h->stream>>>(a);
        h->launches++;
        MCHECK(cudaGetLastError());
    }
    if (mode == 0) {
        d->have_last = true;
        MCHECK(cudaStreamSynchronize(h->stream));
        return 0;
    }
    const size_t nout = (size_t)n_r * TO_NOUT * h->n_theta;
    MCHECK(cudaMemcpyAsync(d->h_out, d->d_out, sizeof(double) * nout, cudaMemcpyDeviceToHost, h->stream));
    MCHECK(cudaStreamSynchronize(h->stream));
    memcpy(out, d->h_out, sizeof(double) * nout);
    return 0;
}

extern "C" int magic_rloop_to_next(magic_rloop *rl, const magic_fields_in *in) { return to_run(rl, in, 0, 0.0, nullptr, true); }
extern "C" int magic_rloop_to_next_dev(magic_rloop *rl, const magic_fields_in *in) { return to_run(rl, in, 0, 0.0, nullptr, false); }
extern "C" int magic_rloop_to(magic_rloop *rl, const magic_fields_in *in, double dtLast, double *out) { return to_run(rl, in, 1, dtLast, out, true); }
extern "C" int magic_rloop_to_dev(magic_rloop *rl, const magic_fields_in *in, double dtLast, double *out) { return to_run(rl, in, 1, dtLast, out, false); }

// ------------------------------------------------------------------------------------------------------
// R.m.s. force balance (l_RMS; rIter.f90:215-252, 433-435, 710; RMS.f90:469-610).  On lRmsCalc steps the reference
// treats every level as bulk, synthesises the pressure gradient and all velocity gradients, forms fourteen more grid products per
// level (get_nl_RMS) and analyses them (transform_to_lm_RMS: 4 scal_to_SH + 5 spat_to_sphertor).  Here: one more column program
// on the batched pipeline -- up to 24 synthesised fields, rms_kernel, 14 analysed fields.  The velocity of the previous step, which
// get_nl_RMS keeps on the grid (vr_old ..., RMS.f90:545-551), is kept as its three potentials instead (magic_rloop_rms_keep: w, dw,
// z of every local level on the DEVICE) and synthesised with the rest.  out: HOST complex [MAGIC_NRMS][n_r_loc][lm_max] in the order
// AdvrLM, LFrLM, dtVrLM, dpkindrLM, Advt2LM, Advp2LM, LFt2LM, LFp2LM, CFt2LM, CFp2LM, PFt2LM, PFp2LM, dtVtLM, dtVpLM; the spectral
// sums of compute_lm_forces (RMS.f90:612-863) stay with the host.  Phase-field penalty, precession and centrifugal terms are formed as
// get_nl does (the latter two only enter AdvrLM, rIter.f90:669-678); the precession phase uses the time of the last pass of the loop.
static const int S_WOLD = S_XI, S_DWOLD = S_DS, S_ZOLD = S_COUNT;  // source slots this column program does not use otherwise
static_assert(S_COUNT < MAGIC_MAX_SRC, "one spare source slot is needed for the kept toroidal potential");

struct RmsPipe {
    int chunk = 0, gx = 0;
    BatchSpec spec;
    RmsIn ri;
    Layout lay;
    Buffers buf;   // carved from magic_rloop::aux by aux_acquire
    int gen = -1;
    LevelInfo *d_lev = nullptr;          // nBc = 0, lDeriv = 1 everywhere (rIter.f90:215)
    double *d_old[3] = {nullptr};        // w, dw, z of the previous stage-1 call: complex [n_r_loc][lm_max]
    double *d_src[S_COUNT] = {nullptr};
    bool need[S_COUNT] = {false};
    bool have_old = false;
};

static void rms_free(RmsPipe *d) {
    if (!d) return;
    layout_free(d->lay);
    cudaFree(d->d_lev);
    for (int i = 0; i < 3; i++) cudaFree(d->d_old[i]);
    for (int i = 0; i < S_COUNT; i++) cudaFree(d->d_src[i]);
    delete d;
}

static int rms_build(magic_rloop *rl) {
    magic_sht *h = rl->h;
    const magic_params &P = rl->p;
    if (P.l_full_sphere) MFAIL("magic_rloop_rms: full-sphere runs are not supported");
    if (!(P.l_conv || P.l_mag_kin)) MFAIL("magic_rloop_rms: needs a flow (l_conv or l_mag_kin)");
    RmsPipe *d = new RmsPipe();
    rl->rms = d;
    BatchSpec &S = d->spec;
    int *rip = (int *)&d->ri;
    for (int i = 0; i < RMS_NF; i++) rip[i] = -1;
    RmsIn &ri = d->ri;
    const Term N_ = {0, F_NONE};
    int nf = 0;
    auto need = [&](std::initializer_list<int> s) { for (int i : s) d->need[i] = true; };
    need({S_W, S_DW, S_DDW, S_Z, S_DZ, S_P});
    add_scal(S, Term{S_W, F_DLH}, N_, LM_ALL, nf, ri.vr);
    add_pair(S, Term{S_DW, F_ONE}, N_, Term{S_Z, F_ONE}, N_, LM_ALL, nf, ri.vt, ri.vp);
    add_scal(S, Term{S_DW, F_DLH}, N_, LM_ALL, nf, ri.dvrdr);
    add_pair(S, Term{S_DDW, F_ONE}, N_, Term{S_DZ, F_ONE}, N_, LM_ALL, nf, ri.dvtdr, ri.dvpdr);
    add_scal(S, Term{S_Z, F_DLH}, N_, LM_ALL, nf, ri.cvr);
    if (P.l_adv_curl) add_pair(S, Term{S_DZ, F_ONE}, N_, Term{S_W, F_OR2DLH}, Term{S_DDW, F_NEG}, LM_ALL, nf, ri.cvt, ri.cvp);
    add_pair(S, Term{S_W, F_DLH}, N_, N_, N_, LM_ALL, nf, ri.dvrdt, ri.dvrdp);
    add_pair(S, Term{S_DW, F_IM}, N_, Term{S_Z, F_IM}, N_, LM_ALL, nf, ri.dvtdp, ri.dvpdp);
    if (P.l_mag || P.l_mag_LF) {
        need({S_B, S_DB, S_DDB, S_AJ, S_DJ});
        add_scal(S, Term{S_B, F_DLH}, N_, LM_ALL, nf, ri.br);
        add_pair(S, Term{S_DB, F_ONE}, N_, Term{S_AJ, F_ONE}, N_, LM_ALL, nf, ri.bt, ri.bp);
        add_scal(S, Term{S_AJ, F_DLH}, N_, LM_ALL, nf, ri.cbr);
        add_pair(S, Term{S_DJ, F_ONE}, N_, Term{S_B, F_OR2DLH}, Term{S_DDB, F_NEG}, LM_ALL, nf, ri.cbt, ri.cbp);
    }
    add_pair(S, Term{S_P, F_ONE}, N_, N_, N_, LM_ALL, nf, ri.dpdt, ri.dpdp);                              // transform_to_grid_RMS
    add_scal(S, Term{S_WOLD, F_DLH}, N_, LM_ALL, nf, ri.vro);
    add_pair(S, Term{S_DWOLD, F_ONE}, N_, Term{S_ZOLD, F_ONE}, N_, LM_ALL, nf, ri.vto, ri.vpo);
    if (P.l_centrifuge) { add_scal(S, Term{S_S, F_ONE}, N_, LM_ALL, nf, ri.s); need({S_S}); }          // CAr reads the entropy
    if (P.l_phase_field) { add_scal(S, Term{S_PHI, F_ONE}, N_, LM_ALL, nf, ri.phi); need({S_PHI}); }   // the penalty reads phi
    S.nfield_in = nf;
    S.afield_s = {0, 1, 2, 3};
    S.afield_vt = {4, 6, 8, 10, 12};
    S.afield_vp = {5, 7, 9, 11, 13};
    S.nfield_out = RMS_NOUT;
    const size_t lm2 = 2 * (size_t)h->lm_max;
    for (int i = 0; i < 3; i++) MCHECK(cudaMalloc((void **)&d->d_old[i], sizeof(double) * lm2 * (size_t)rl->n_r_loc));
    const double per_level = 2.2 * 8.0 * (double)h->n_theta * h->n_phi * (nf + RMS_NOUT) + 64.0 * h->lm_max * (S_COUNT + RMS_NOUT);
    const int chunk = aux_chunk(rl, per_level);
    d->chunk = chunk;   // the workspace itself is bound by aux_acquire at run time
    std::vector<LevelInfo> lev = rl->lev;
    for (auto &L : lev) { L.lDeriv = 1; L.nBc = 0; }
    if (dev_upload_vec(&d->d_lev, lev)) return 1;
    const size_t plane = (size_t)h->nh * h->n_phi;
    d->gx = (int)std::min<size_t>((plane + DIAG_THREADS - 1) / DIAG_THREADS, 8 * 148);
    return 0;
}

// keeps w, dw, z of all local levels (the reference's vr_old, vt_old, vp_old of RMS.f90:549-551, in spectral form)
static int rms_keep(magic_rloop *rl, const magic_fields_in *in, bool host_in) {
    if (!rl || !in) MFAIL("magic_rloop_rms_keep: null argument");
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    if (!rl->rms)
        if (rms_build(rl)) { rms_free(rl->rms); rl->rms = nullptr; return 1; }
    RmsPipe *d = rl->rms;
    if (!in->w || !in->dw || !in->z) MFAIL("magic_rloop_rms_keep: w, dw, z are needed");
    const size_t bytes = sizeof(double) * 2 * (size_t)h->lm_max * (size_t)rl->n_r_loc;
    const double *src[3] = {in->w, in->dw, in->z};
    for (int i = 0; i < 3; i++)
        MCHECK(cudaMemcpyAsync(d->d_old[i], src[i], bytes, host_in ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, h->stream));
    MCHECK(cudaStreamSynchronize(h->stream));
    d->have_old = true;
    return 0;
}

static int rms_run(magic_rloop *rl, const magic_fields_in *in, double dt, double *out, bool host_in) {
    if (!rl || !in || !out) MFAIL("magic_rloop_rms: null argument");
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    if (!rl->rms)
        if (rms_build(rl)) { rms_free(rl->rms); rl->rms = nullptr; return 1; }
    RmsPipe *d = rl->rms;
    const magic_params &P = rl->p;
    if (aux_acquire(rl, d, d->spec, d->chunk, d->lay, d->buf, d->gen)) return 1;
    if (!d->have_old) MFAIL("magic_rloop_rms: no previous velocity kept (call magic_rloop_rms_keep on every stage-1 step while l_RMS is on)");
    if (!(dt > 0.0)) MFAIL("magic_rloop_rms: dt must be positive");
    const double *ip[S_COUNT];
    in_ptrs(in, ip);
    for (int i = 0; i < S_COUNT; i++)
        if (d->need[i] && !ip[i]) MFAIL("magic_rloop_rms: a required input field is null (w, dw, ddw, z, dz, p and the magnetic potentials)");
    const size_t lm2 = 2 * (size_t)h->lm_max;
    const int nl = d->chunk, n_r = rl->n_r_loc;
    if (host_in)
        for (int i = 0; i < S_COUNT; i++)
            if (d->need[i] && !d->d_src[i]) MCHECK(cudaMalloc((void **)&d->d_src[i], sizeof(double) * lm2 * nl));
    RmsArgs a{};
    a.ri = d->ri; a.gin = d->buf.gin; a.gout = d->buf.gout; a.n_lev = nl; a.nh = h->nh; a.n_phi = h->n_phi;
    a.l_conv_nl = P.l_conv_nl; a.l_mag_LF = P.l_mag_LF; a.l_mag_nl = P.l_mag_nl; a.l_adv_curl = P.l_adv_curl; a.n_r_LCR = P.n_r_LCR;
    a.LFfac = P.LFfac; a.CorFac = P.CorFac; a.o_dt = 1.0 / dt;
    a.l_phase_field = P.l_phase_field; a.l_precession = P.l_precession; a.l_centrifuge = P.l_centrifuge; a.minc = h->minc;
    a.pen = P.l_phase_field ? 1.0 / (P.epsPhase * P.epsPhase) / (P.penaltyFac * P.penaltyFac) : 0.0;
    a.posnalp = -2.0 * P.oek * P.po * sin(P.prec_angle);
    a.oek_time = P.oek * rl->last_time;
    a.cafac = P.dilution_fac * P.ra * P.opr;
    a.sinth = h->d_sinth; a.costh = h->d_costh;
    for (int l0 = 0; l0 < n_r; l0 += nl) {
        const int s0 = std::min(l0, n_r - nl);
        const double *src[MAGIC_MAX_SRC];
        for (int i = 0; i < MAGIC_MAX_SRC; i++) src[i] = nullptr;
        for (int i = 0; i < S_COUNT; i++) {
            if (!d->need[i]) continue;
            if (host_in) {
                MCHECK(cudaMemcpyAsync(d->d_src[i], ip[i] + (size_t)s0 * lm2, sizeof(double) * lm2 * nl, cudaMemcpyHostToDevice, h->stream));
                src[i] = d->d_src[i];
            } else {
                src[i] = ip[i] + (size_t)s0 * lm2;
            }
        }
        src[S_WOLD] = d->d_old[0] + (size_t)s0 * lm2;
        src[S_DWOLD] = d->d_old[1] + (size_t)s0 * lm2;
        src[S_ZOLD] = d->d_old[2] + (size_t)s0 * lm2;
        if (run_synthesis(h, d->spec, d->lay, d->buf, src, d->d_lev + s0, nullptr)) return 1;
        a.lev = d->d_lev + s0;
        rms_kernel<<<dim3(d->gx, nl), DIAG_THREADS, 0, h->stream>>>(a);
        h->launches++;
        MCHECK(cudaGetLastError());
        if (run_analysis(h, d->spec, d->lay, d->buf, d->d_lev + s0, nullptr, true)) return 1;
        // nl_s: [4][nl][lm_max] -> out 0..3; nl_v: [S0, T0, ..., S4, T4][nl][lm_max] -> out 4..13
        for (int q = 0; q < RMS_NOUT; q++) {
            const double *srcq = q < 4 ? d->buf.nl_s + (size_t)q * nl * lm2 : d->buf.nl_v + (size_t)(q - 4) * nl * lm2;
            MCHECK(cudaMemcpyAsync(out + ((size_t)q * n_r + s0) * lm2, srcq, sizeof(double) * lm2 * nl, cudaMemcpyDeviceToHost, h->stream));
        }
        MCHECK(cudaStreamSynchronize(h->stream));  // the result buffers are reused by the next chunk
    }
    return 0;
}

extern "C" int magic_rloop_rms_keep(magic_rloop *rl, const magic_fields_in *in) { return rms_keep(rl, in, true); }
extern "C" int magic_rloop_rms_keep_dev(magic_rloop *rl, const magic_fields_in *in) { return rms_keep(rl, in, false); }
extern "C" int magic_rloop_rms(magic_rloop *rl, const magic_fields_in *in, double dt, double *out) { return rms_run(rl, in, dt, out, true); }
extern "C" int magic_rloop_rms_dev(magic_rloop *rl, const magic_fields_in *in, double dt, double *out) { return rms_run(rl, in, dt, out, false); }
