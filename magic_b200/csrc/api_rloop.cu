// api_rloop.cu -- C ABI: the batched radial loop (rIter.f90:94-464 without the output hooks).
//
// All local radial levels are processed in chunks of `level_chunk` levels; each chunk is one pass of
//   synthesis operand assembly -> Legendre GEMM -> c2r FFT -> get_nl(+Courant) -> r2c FFT -> Legendre GEMM
//   -> extraction -> get_td.
// Levels are independent, so chunking changes nothing but the batch width.
#include "../../include/magic_sht.h"
#include "engine.cuh"

#include <functional>

using namespace magic;

enum Src { S_W = 0, S_DW, S_DDW, S_Z, S_DZ, S_S, S_DS, S_P, S_XI, S_B, S_DB, S_DDB, S_AJ, S_DJ, S_PHI, S_COUNT };
enum OutIdx { O_DWDT = 0, O_DZDT, O_DPDT, O_DSDT, O_DXIDT, O_DBDT, O_DJDT, O_DVXVH, O_DVXBH, O_DVSR, O_DVXIR, O_DPHIDT, O_COUNT };

struct magic_rloop {
    magic_sht *h = nullptr;
    magic_params p;
    int n_r_loc = 0;
    std::vector<LevelInfo> lev;
    LevelInfo *d_lev = nullptr;
    BatchSpec spec;
    GridIn gi;
    GridOut go;
    Buffers buf;
    std::vector<int> chunk_start, chunk_size;
    int level_chunk = 0;      // the chunk length the sizes above were derived from (clamped to this rank's slab)
    int level_chunk_req = 0;  // the same before clamping: a pure function of the truncation and the field set, equal on every rank
    // hooks of the chunk loop (magic_rloop_run_lm_dev): called on the host right before the first / after the last kernel
    // of chunk c is queued; they queue the transposes of that chunk on the communication stream
    std::function<int(int)> hook_before, hook_after;
    struct LmPipe *lmpipe = nullptr;
    // one workspace for the log-step batches below (they never run concurrently): grown to the largest request; `gen` tells a
    // batch that its descriptors point into an arena that no longer exists, `owner` whose data the arena holds
    struct { char *p = nullptr; size_t bytes = 0; int gen = 0; const void *owner = nullptr; } aux;
    struct DiagPipe *diag = nullptr;  // log-step diagnostics (api_diag.cu), built on first use
    struct DtbPipe *dtb = nullptr;    // get_dtBLM batch (api_diag.cu), built on first use
    struct ToPipe *to = nullptr;      // torsional-oscillation sums (api_diag.cu), built on first use
    struct RmsPipe *rms = nullptr;    // r.m.s. force balance batch (api_diag.cu), built on first use
    double last_time = 0.0;           // `time` of the last pass of the loop (the precession terms of the r.m.s. batch use it)
    // LM-side prologue / epilogue: the host's radial scheme as dense matrices, radial functions on all levels
    std::vector<double> D1h, D2h, lmrad_h;  // [n_r][n_r] row-major x2; [4][n_r]: or2, orho1, dentropy0, l_R
    int n_r_mat = 0, lm_derivs = 0, lm_finish = 0;
    std::vector<Layout> lays;    // one layout per distinct chunk size; lays[0] belongs to the largest chunk (it sizes the workspace)
    std::vector<int> lay_sizes;
    int buf_levels = 0;          // number of levels the workspace `buf` was allocated for
    // nl_lm slots
    int a_Advr = -1, a_VSr = -1, a_VxBr = -1, a_VXir = -1, a_heat = -1, a_phi = -1;  // scalar-class analysis columns
    int a_Adv = -1, a_VS = -1, a_VxB = -1, a_VXi = -1;                    // vector pairs
    // resident device copies for the host-pointer entry point
    double *d_in[S_COUNT] = {nullptr};
    double *d_out[O_COUNT] = {nullptr};
    double *d_dtrkc = nullptr, *d_dthkc = nullptr;
    double *d_tq_partial = nullptr, *d_torque = nullptr, *h_torque = nullptr;  // Lorentz torques (ic, ma)
    int tq_parts = 0;
    double torque[2] = {0.0, 0.0};
    // get_br_v_bcs (rIter.f90:267-277): local level index of the CMB / ICB when l_b_nl_cmb / l_b_nl_icb, else -1;
    // results [cmb vt, cmb vp, icb vt, icb vp] x complex [lm_max] on the device and in pinned host memory
    int bc_lev[2] = {-1, -1};
    double *d_brv = nullptr, *h_brv = nullptr;
    bool need_in[S_COUNT] = {false};
    bool need_out[O_COUNT] = {false};
    cudaEvent_t ev[16];
    std::vector<cudaEvent_t> cev;  // 10 timing events per chunk (stage boundaries), read once per run
    double timing[8] = {0};
    double exposed[2] = {0, 0};
    double legendre_flops = 0;
    double units_ref = 0, units_exec = 0;  // scalar-equivalent Legendre passes per bulk level: reference count / executed here
    std::vector<std::pair<const void *, size_t>> registered;  // magic_rloop_pin_host
    // host-pointer path: uploads / downloads of level chunks overlap the compute of neighbouring chunks
    cudaStream_t s_up = nullptr, s_down = nullptr;
    std::vector<cudaEvent_t> up_done, comp_done;
    const double *host_in[S_COUNT] = {nullptr};
    double *host_out[O_COUNT] = {nullptr};
    double *host_dtrkc = nullptr, *host_dthkc = nullptr;  // pinned staging (a D2H copy into pageable memory would block the host)
    bool pipelined = false;
};

// Level chunks of a slab of n_r_loc levels: full chunks of `level_chunk` levels (the GEMM column count 4*npair*n_lev is then
// a multiple of the 64-wide tile for level_chunk % 4 == 0), a remainder of at most level_chunk/4 levels folded into the last
// chunk, a larger one as a short chunk of its own.  A pure function of its arguments: every rank can compute the chunks of
// every other rank (magic_rloop_run_lm_dev).  (Balanced sizes such as 15/16 for 257 levels leave every N-edge GEMM tile 3/4
// full: measured 1.00 vs 0.94 ms per level.)
static void level_chunks(int n_r_loc, int level_chunk, std::vector<int> &start, std::vector<int> &size) {
    start.clear();
    size.clear();
    level_chunk = std::max(1, std::min(level_chunk, n_r_loc));
    const int nfull = n_r_loc / level_chunk, rem = n_r_loc % level_chunk;
    int big = level_chunk, small = 0, nbig = nfull, nsmall = 0;  // nbig chunks of `big`, then nsmall of `small`
    if (rem > 0 && rem <= level_chunk / 4 && nfull >= 1) { small = level_chunk + rem; nbig = nfull - 1; nsmall = 1; }
    else if (rem > 0) { small = rem; nsmall = 1; }
    int pos = 0;
    for (int c = 0; c < nbig + nsmall; c++) {
        const int sz = c < nbig ? big : small;
        start.push_back(pos);
        size.push_back(sz);
        pos += sz;
    }
}

extern "C" int magic_level_chunks(int n_r_loc, int level_chunk, int *n_chunks, int *start, int *size) {
    if (n_r_loc < 1 || level_chunk < 1 || !n_chunks) MFAIL("magic_level_chunks: bad arguments");
    std::vector<int> s, z;
    level_chunks(n_r_loc, level_chunk, s, z);
    *n_chunks = (int)s.size();
    for (size_t i = 0; i < s.size(); i++) {
        if (start) start[i] = s[i];
        if (size) size[i] = z[i];
    }
    return 0;
}

static void add_scal(BatchSpec &s, Term t0, Term t1, int lmask, int &field_counter, int &slot) {
    ScalCol c{};
    c.t[0] = t0; c.t[1] = t1; c.lmask = lmask;
    s.scal.push_back(c);
    slot = field_counter++;
    s.field_s.push_back(slot);
}
static void add_pair(BatchSpec &s, Term S0, Term S1, Term T0, Term T1, int lmask, int &field_counter, int &slot_t, int &slot_p) {
    VecPair v{};
    v.S[0] = S0; v.S[1] = S1; v.T[0] = T0; v.T[1] = T1; v.lmask = lmask;
    s.vec.push_back(v);
    slot_t = field_counter++;
    slot_p = field_counter++;
    s.field_v.push_back(slot_t);
    s.field_v.push_back(slot_p);
}

static void lmpipe_free(struct LmPipe *p);
static void diag_free(struct DiagPipe *d);
static void dtb_free(struct DtbPipe *d);
static void to_free(struct ToPipe *d);
static void rms_free(struct RmsPipe *d);

// Tapered variant for the pipelined multi-rank call: a short first and last chunk (`taper` levels each) so that the inbound
// transpose of the first chunk and the outbound transpose of the last one -- the two that cannot hide under any compute --
// are small; the levels in between are chunked as above.  Pure function as well.
static void level_chunks_tapered(int n_r_loc, int level_chunk, int taper, std::vector<int> &start, std::vector<int> &size) {
    level_chunk = std::max(1, std::min(level_chunk, n_r_loc));
    if (taper <= 0 || taper >= level_chunk || n_r_loc < 2 * taper + std::max(4, level_chunk / 2)) {
        level_chunks(n_r_loc, level_chunk, start, size);
        return;
    }
    std::vector<int> ms, mz;
    level_chunks(n_r_loc - 2 * taper, level_chunk, ms, mz);
    start.assign(1, 0);
    size.assign(1, taper);
    for (size_t i = 0; i < ms.size(); i++) { start.push_back(taper + ms[i]); size.push_back(mz[i]); }
    start.push_back(n_r_loc - taper);
    size.push_back(taper);
}

// (Re)partitions the slab into the given level chunks: one layout per distinct size, the workspace sized for the largest.
static int rloop_set_chunks(magic_rloop *rl, const std::vector<int> &start, const std::vector<int> &size) {
    magic_sht *h = rl->h;
    rl->chunk_start = start;
    rl->chunk_size = size;
    std::vector<int> sizes(size);
    std::sort(sizes.begin(), sizes.end(), std::greater<int>());
    sizes.erase(std::unique(sizes.begin(), sizes.end()), sizes.end());
    for (auto &L : rl->lays) layout_free(L);
    rl->lays.assign(sizes.size(), Layout());
    rl->lay_sizes = sizes;
    layout_sizes(h, rl->spec, sizes[0], rl->lays[0]);
    if (sizes[0] > rl->buf_levels) {
        buffers_free(rl->buf);
        if (buffers_alloc(h, rl->spec, rl->lays[0], rl->buf)) return 1;
        rl->buf_levels = sizes[0];
        cudaFree(rl->d_tq_partial);
        rl->tq_parts = (int)(((size_t)h->nh * h->n_phi + NL_THREADS - 1) / NL_THREADS);
        MCHECK(cudaMalloc((void **)&rl->d_tq_partial, sizeof(double) * (size_t)rl->tq_parts * sizes[0]));
    }
    for (size_t i = 0; i < sizes.size(); i++) {
        if (i) layout_sizes(h, rl->spec, sizes[i], rl->lays[i]);
        if (layout_bind(h, rl->spec, rl->lays[i], rl->buf)) return 1;
    }
    const size_t nchunks = size.size();
    while (rl->up_done.size() < nchunks) {
        cudaEvent_t e;
        MCHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        rl->up_done.push_back(e);
        MCHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        rl->comp_done.push_back(e);
    }
    while (rl->cev.size() < 10 * nchunks) {
        cudaEvent_t e;
        MCHECK(cudaEventCreate(&e));
        rl->cev.push_back(e);
    }
    return 0;
}

extern "C" int magic_rloop_destroy(magic_rloop *rl) {
    if (!rl) return 0;
    cudaSetDevice(rl->h->dev);
    for (const auto &r : rl->registered) cudaHostUnregister((void *)r.first);
    for (auto &L : rl->lays) layout_free(L);
    for (auto e : rl->cev) cudaEventDestroy(e);
    buffers_free(rl->buf);
    for (int i = 0; i < S_COUNT; i++) cudaFree(rl->d_in[i]);
    for (int i = 0; i < O_COUNT; i++) cudaFree(rl->d_out[i]);
    cudaFree(rl->d_dtrkc); cudaFree(rl->d_dthkc); cudaFree(rl->d_lev); cudaFree(rl->d_tq_partial); cudaFree(rl->d_torque);
    cudaFree(rl->d_brv);
    if (rl->h_brv) cudaFreeHost(rl->h_brv);
    lmpipe_free(rl->lmpipe);
    diag_free(rl->diag);
    dtb_free(rl->dtb);
    to_free(rl->to);
    rms_free(rl->rms);
    cudaFree(rl->aux.p);
    if (rl->h_torque) cudaFreeHost(rl->h_torque);
    for (int i = 0; i < 16; i++) cudaEventDestroy(rl->ev[i]);
    for (auto e : rl->up_done) cudaEventDestroy(e);
    for (auto e : rl->comp_done) cudaEventDestroy(e);
    if (rl->host_dtrkc) cudaFreeHost(rl->host_dtrkc);
    if (rl->host_dthkc) cudaFreeHost(rl->host_dthkc);
    if (rl->s_up) cudaStreamDestroy(rl->s_up);
    if (rl->s_down) cudaStreamDestroy(rl->s_down);
    delete rl;
    return 0;
}

extern "C" int magic_rloop_create(magic_sht *h, const magic_params *pp, const magic_radial *rad, int n_r_loc, int level_chunk,
                                  magic_rloop **out) {
    if (!h || !pp || !rad || !out) MFAIL("magic_rloop_create: null argument");
    if (n_r_loc < 1) MFAIL("magic_rloop_create: n_r_loc < 1");
    *out = nullptr;
    MCHECK(cudaSetDevice(h->dev));
    const magic_params &P = *pp;
    // the reference forces l_adv_curl = .false. for anelastic runs (Namelists.f90:540); with both set the viscous heating would
    // read velocity derivatives that the curl-form column program never synthesises
    if (P.l_anel && P.l_adv_curl) MFAIL("magic_rloop_create: l_anel with l_adv_curl is not a configuration of the reference (Namelists.f90:540)");
    magic_rloop *rl = new magic_rloop();
    rl->h = h;
    rl->p = P;
    rl->n_r_loc = n_r_loc;
    for (int i = 0; i < 16; i++) cudaEventCreate(&rl->ev[i]);

    // ---- per-level flags, rIter.f90:181-218
    bool lMagNlBc = false;
    if (((P.l_mag_nl || P.l_mag_kin) && (P.ktopv == 1 || P.l_cond_ma || (P.ktopv == 2 && P.l_rot_ma))) ||
        (P.kbotv == 1 || P.l_cond_ic || (P.kbotv == 2 && P.l_rot_ic)))
        lMagNlBc = true;
    rl->lev.resize(n_r_loc);
    for (int i = 0; i < n_r_loc; i++) {
        LevelInfo &L = rl->lev[i];
        memset(&L, 0, sizeof(L));
        L.nR = rad->nR[i];
        L.lcut = rad->l_R[i];
        if (L.lcut < 0 || L.lcut > h->l_max) { magic_rloop_destroy(rl); MFAIL("magic_rloop_create: l_R out of range"); }
        bool is_cmb = L.nR == 1, is_icb = L.nR == P.n_r_max;
        bool l_bound = is_cmb || is_icb;
        int nBc = 0, lDeriv = 1;
        if (is_cmb) { nBc = P.ktopv; lDeriv = 0; }
        else if (is_icb) { nBc = P.kbotv; lDeriv = 0; }
        bool loop_bound = l_bound;
        if (P.l_parallel_solve || (P.l_single_matrix && P.l_temperature_diff)) { lDeriv = 1; nBc = 0; loop_bound = false; }
        L.nBc = nBc; L.lDeriv = lDeriv; L.l_bound = l_bound ? 1 : 0;
        L.nl_on = (!loop_bound || lMagNlBc) ? 1 : 0;
        L.cour_on = (!P.l_full_sphere || !is_icb) ? 1 : 0;
        L.center = (P.l_full_sphere && is_icb) ? 1 : 0;
        L.torque = 0;  // rIter.f90:279-292
        if (is_icb && P.l_mag_LF && P.l_rot_ic && P.l_cond_ic) L.torque = 1;
        if (is_cmb && P.l_mag_LF && P.l_rot_ma && P.l_cond_ma) L.torque = 2;
        L.r = rad->r[i]; L.or1 = rad->or1[i]; L.or2 = rad->or2[i]; L.or4 = rad->or4[i]; L.orho1 = rad->orho1[i];
        L.orho2 = rad->orho2[i]; L.beta = rad->beta[i]; L.rho0 = rad->rho0[i]; L.otemp1 = rad->otemp1[i]; L.temp0 = rad->temp0[i];
        L.visc = rad->visc[i]; L.lambda = rad->lambda[i]; L.epscProf = rad->epscProf[i]; L.delxr2 = rad->delxr2[i];
        L.delxh2 = rad->delxh2[i];
    }
    if (dev_upload_vec(&rl->d_lev, rl->lev)) { magic_rloop_destroy(rl); return 1; }

    // ---- column program, transform_to_grid_space rIter.f90:466-622
    BatchSpec &S = rl->spec;
    GridIn &gi = rl->gi;
    int *gip = (int *)&gi;
    for (size_t i = 0; i < sizeof(GridIn) / sizeof(int); i++) gip[i] = -1;
    const Term N_ = {0, F_NONE};
    int nf = 0;
    double units_syn = 0, units_an = 0;  // scalar-equivalent Legendre passes per bulk level
    if (P.l_conv || P.l_mag_kin) {
        if (P.l_heat) { add_scal(S, Term{S_S, F_ONE}, N_, LM_ALL, nf, gi.s); rl->need_in[S_S] = true; units_syn += 1; }
        if (P.l_chemical_conv) { add_scal(S, Term{S_XI, F_ONE}, N_, LM_ALL, nf, gi.xi); rl->need_in[S_XI] = true; units_syn += 1; }
        if (P.l_phase_field) { add_scal(S, Term{S_PHI, F_ONE}, N_, LM_ALL, nf, gi.phi); rl->need_in[S_PHI] = true; units_syn += 1; }  // rIter.f90:509
        rl->need_in[S_W] = rl->need_in[S_DW] = rl->need_in[S_Z] = true;
        if (P.l_full_sphere) rl->need_in[S_DDW] = true;
        add_scal(S, Term{S_W, F_DLH}, N_, LM_VEL, nf, gi.vr);
        add_pair(S, Term{S_DW, F_ONE}, N_, Term{S_Z, F_ONE}, N_, LM_VEL, nf, gi.vt, gi.vp);
        units_syn += 5;
        if (P.l_adv_curl) {
            rl->need_in[S_DDW] = rl->need_in[S_DZ] = true;
            add_scal(S, Term{S_Z, F_DLH}, N_, LM_VELBULK, nf, gi.cvr);
            add_pair(S, Term{S_DZ, F_ONE}, N_, Term{S_W, F_OR2DLH}, Term{S_DDW, F_NEG}, LM_VELBULK, nf, gi.cvt, gi.cvp);
            units_syn += 5;
        } else {
            rl->need_in[S_DDW] = rl->need_in[S_DZ] = true;
            add_scal(S, Term{S_DW, F_DLH}, N_, LM_VELBULK, nf, gi.dvrdr);
            add_pair(S, Term{S_DDW, F_ONE}, N_, Term{S_DZ, F_ONE}, N_, LM_VELBULK, nf, gi.dvtdr, gi.dvpdr);
            add_scal(S, Term{S_Z, F_DLH}, N_, LM_VELBULK, nf, gi.cvr);
            add_pair(S, Term{S_W, F_DLH}, N_, N_, N_, LM_VELBULK, nf, gi.dvrdt, gi.dvrdp);
            add_pair(S, Term{S_DW, F_IM}, N_, Term{S_Z, F_IM}, N_, LM_VELBULK, nf, gi.dvtdp, gi.dvpdp);
            units_syn += 5 + 1 + 2 + 4;
        }
    }
    if (P.l_mag || P.l_mag_LF) {
        rl->need_in[S_B] = rl->need_in[S_DB] = rl->need_in[S_AJ] = rl->need_in[S_DDB] = rl->need_in[S_DJ] = true;
        add_scal(S, Term{S_B, F_DLH}, N_, LM_ALL, nf, gi.br);
        add_pair(S, Term{S_DB, F_ONE}, N_, Term{S_AJ, F_ONE}, N_, LM_ALL, nf, gi.bt, gi.bp);
        add_scal(S, Term{S_AJ, F_DLH}, N_, LM_DERIV, nf, gi.cbr);
        add_pair(S, Term{S_DJ, F_ONE}, N_, Term{S_B, F_OR2DLH}, Term{S_DDB, F_NEG}, LM_DERIV, nf, gi.cbt, gi.cbp);
        units_syn += 10;
    }
    S.nfield_in = nf;
    // ---- products and their analysis, transform_to_lm_space rIter.f90:624-712
    GridOut &go = rl->go;
    int *gop = (int *)&go;
    for (size_t i = 0; i < sizeof(GridOut) / sizeof(int); i++) gop[i] = -1;
    int no = 0;
    auto add_qst = [&](int &fr, int &ft, int &fp, int &slot_s, int &slot_v) {
        fr = no++; ft = no++; fp = no++;
        slot_s = (int)S.afield_s.size();
        S.afield_s.push_back(fr);
        slot_v = (int)S.afield_vt.size();
        S.afield_vt.push_back(ft);
        S.afield_vp.push_back(fp);
        units_an += 5;
    };
    if (P.l_conv_nl || P.l_mag_LF) add_qst(go.Advr, go.Advt, go.Advp, rl->a_Advr, rl->a_Adv);
    if (P.l_heat) {
        add_qst(go.VSr, go.VSt, go.VSp, rl->a_VSr, rl->a_VS);
        if (P.l_anel) { go.heat = no++; rl->a_heat = (int)S.afield_s.size(); S.afield_s.push_back(go.heat); units_an += 1; }
    }
    if (P.l_chemical_conv) add_qst(go.VXir, go.VXit, go.VXip, rl->a_VXir, rl->a_VXi);
    if (P.l_phase_field) { go.phiTerms = no++; rl->a_phi = (int)S.afield_s.size(); S.afield_s.push_back(go.phiTerms); units_an += 1; }  // rIter.f90:698
    if (P.l_mag_nl) add_qst(go.VxBr, go.VxBt, go.VxBp, rl->a_VxBr, rl->a_VxB);
    S.nfield_out = no;
    rl->legendre_flops = (units_syn + units_an) * 2.0 * (double)h->n_theta * (double)h->lm_max * (double)n_r_loc;
    rl->units_ref = units_syn + units_an;
    rl->units_exec = (double)(S.scal.size() + 2 * S.vec.size() + S.afield_s.size() + 2 * S.afield_vt.size());

    // ---- outputs needed
    if (P.l_conv) { rl->need_out[O_DZDT] = rl->need_out[O_DWDT] = true; if (P.l_double_curl) rl->need_out[O_DVXVH] = true; }
    if (!P.l_double_curl) rl->need_out[O_DPDT] = true;
    if (P.l_heat) rl->need_out[O_DSDT] = rl->need_out[O_DVSR] = true;
    if (P.l_chemical_conv) rl->need_out[O_DXIDT] = rl->need_out[O_DVXIR] = true;
    if (P.l_phase_field) rl->need_out[O_DPHIDT] = true;
    if (P.l_mag) rl->need_out[O_DBDT] = rl->need_out[O_DJDT] = rl->need_out[O_DVXBH] = true;

    // ---- chunking
    if (level_chunk <= 0) {
        // The automatic choice depends on the truncation and the field set only -- never on this rank's slab or its free
        // memory -- so every rank of a run derives the same value (magic_rloop_run_lm_dev computes the chunks of its peers
        // from it).  A workspace that does not fit fails loudly in buffers_alloc below.
        // measured at l_max=1023 (whole loop, ms per level): 16-level chunks 1.79, 32-level chunks 1.75, 64-level chunks 1.74: with
        // 13 + 9 complex columns per level the GEMM tile counts 26 n_lev / 64 and 18 n_lev / 64 are whole numbers at 32 levels
        level_chunk = 32;
    }
    rl->level_chunk_req = level_chunk;
    level_chunk = std::min(level_chunk, n_r_loc);
    rl->level_chunk = level_chunk;
    {
        std::vector<int> cs, cz;
        level_chunks(n_r_loc, level_chunk, cs, cz);
        if (rloop_set_chunks(rl, cs, cz)) { magic_rloop_destroy(rl); return 1; }
    }
    MCHECK(cudaStreamCreateWithFlags(&rl->s_up, cudaStreamNonBlocking));
    MCHECK(cudaStreamCreateWithFlags(&rl->s_down, cudaStreamNonBlocking));
    MCHECK(cudaMalloc((void **)&rl->d_torque, sizeof(double) * 2));
    {   // l_b_nl_cmb / l_b_nl_icb, Namelists.f90:713-729
        for (int i = 0; i < n_r_loc; i++) {
            if (rl->lev[i].nR == 1 && P.l_mag_nl && P.ktopv == 1 && P.l_cond_ma) rl->bc_lev[0] = i;
            if (rl->lev[i].nR == P.n_r_max && P.l_mag_nl && P.kbotv == 1 && P.l_cond_ic) rl->bc_lev[1] = i;
        }
        if (rl->bc_lev[0] >= 0 || rl->bc_lev[1] >= 0) {
            const size_t nb = sizeof(double) * 8 * (size_t)h->lm_max;
            MCHECK(cudaMalloc((void **)&rl->d_brv, nb));
            MCHECK(cudaMemset(rl->d_brv, 0, nb));
            MCHECK(cudaMallocHost((void **)&rl->h_brv, nb));
            memset(rl->h_brv, 0, nb);
        }
    }
    MCHECK(cudaMallocHost((void **)&rl->h_torque, sizeof(double) * 2));
    MCHECK(cudaMallocHost((void **)&rl->host_dtrkc, sizeof(double) * n_r_loc));
    MCHECK(cudaMallocHost((void **)&rl->host_dthkc, sizeof(double) * n_r_loc));
    MCHECK(cudaMalloc((void **)&rl->d_dtrkc, sizeof(double) * n_r_loc));
    MCHECK(cudaMalloc((void **)&rl->d_dthkc, sizeof(double) * n_r_loc));
    *out = rl;
    return 0;
}

static const double *const *in_ptrs(const magic_fields_in *in, const double *tmp[S_COUNT]) {
    tmp[S_W] = in->w; tmp[S_DW] = in->dw; tmp[S_DDW] = in->ddw; tmp[S_Z] = in->z; tmp[S_DZ] = in->dz; tmp[S_S] = in->s;
    tmp[S_DS] = in->ds; tmp[S_P] = in->p; tmp[S_XI] = in->xi; tmp[S_B] = in->b; tmp[S_DB] = in->db; tmp[S_DDB] = in->ddb;
    tmp[S_AJ] = in->aj; tmp[S_DJ] = in->dj; tmp[S_PHI] = in->phi;
    return tmp;
}
static void out_ptrs(const magic_fields_out *o, double *tmp[O_COUNT]) {
    tmp[O_DWDT] = o->dwdt; tmp[O_DZDT] = o->dzdt; tmp[O_DPDT] = o->dpdt; tmp[O_DSDT] = o->dsdt; tmp[O_DXIDT] = o->dxidt;
    tmp[O_DBDT] = o->dbdt; tmp[O_DJDT] = o->djdt; tmp[O_DVXVH] = o->dVxVhLM; tmp[O_DVXBH] = o->dVxBhLM; tmp[O_DVSR] = o->dVSrLM;
    tmp[O_DVXIR] = o->dVXirLM; tmp[O_DPHIDT] = o->dphidt;
}

// ---- the chunk loop in three pieces: begin (checks, start event), chunk c (all kernels of one level chunk, queued on the
//      handle's stream), end (boundary products, torques, the one host synchronisation of the run, stage timings) --------
struct RunCtx {
    const double *ip[S_COUNT];
    double *op[O_COUNT];
    double *dtrkc, *dthkc;
    double time;
};

static int rloop_begin(magic_rloop *rl, const magic_fields_in *in, const magic_fields_out *out, double time, RunCtx &x) {
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    in_ptrs(in, x.ip);
    out_ptrs(out, x.op);
    for (int i = 0; i < S_COUNT; i++)
        if (rl->need_in[i] && !x.ip[i]) MFAIL("magic_rloop: a required input field is null");
    for (int i = 0; i < O_COUNT; i++)
        if (rl->need_out[i] && !x.op[i]) MFAIL("magic_rloop: a required output field is null");
    if (!out->dtrkc || !out->dthkc) MFAIL("magic_rloop: dtrkc/dthkc are null");
    x.dtrkc = out->dtrkc; x.dthkc = out->dthkc; x.time = time;
    rl->last_time = time;
    cudaEventRecord(rl->ev[15], h->stream);
    MCHECK(cudaMemsetAsync(rl->d_torque, 0, sizeof(double) * 2, h->stream));  // rIter.f90:177-178
    return 0;
}

static int rloop_chunk(magic_rloop *rl, int c, const RunCtx &x) {
    magic_sht *h = rl->h;
    const magic_params &P = rl->p;
    const size_t lm2 = 2 * (size_t)h->lm_max;
    const size_t plane = (size_t)h->nh * h->n_phi;
    const int l0 = rl->chunk_start[c], nl = rl->chunk_size[c];
    const Layout *Lp = nullptr;
    for (size_t i = 0; i < rl->lay_sizes.size(); i++)
        if (rl->lay_sizes[i] == nl) Lp = &rl->lays[i];
    if (!Lp) MFAIL("magic_rloop: internal: no layout for this chunk size");
    const Layout &L = *Lp;
    cudaEvent_t *ev = rl->cev.data() + 10 * (size_t)c;
    const LevelInfo *d_lev = rl->d_lev + l0;
    const double *src[MAGIC_MAX_SRC];
    for (int i = 0; i < MAGIC_MAX_SRC; i++) src[i] = (i < S_COUNT && x.ip[i]) ? x.ip[i] + (size_t)l0 * lm2 : nullptr;
    if (run_synthesis(h, rl->spec, L, rl->buf, src, d_lev, ev)) return 1;  // ev[0..3]
    // ---- get_nl + Courant
    MCHECK(cudaMemsetAsync(rl->buf.courmax, 0, sizeof(unsigned long long) * 2 * nl, h->stream));
    NlArgs a{};
    NlFlags &F = a.f;
    F.l_conv_nl = P.l_conv_nl; F.l_heat_nl = P.l_heat_nl; F.l_mag_nl = P.l_mag_nl; F.l_mag_LF = P.l_mag_LF; F.l_mag = P.l_mag;
    F.l_mag_kin = P.l_mag_kin; F.l_adv_curl = P.l_adv_curl; F.l_anel = P.l_anel; F.l_chemical_conv = P.l_chemical_conv;
    F.l_precession = P.l_precession; F.l_centrifuge = P.l_centrifuge; F.l_cour_alf_damp = P.l_cour_alf_damp;
    F.l_full_sphere = P.l_full_sphere; F.n_r_LCR = P.n_r_LCR; F.l_phase_field = P.l_phase_field;
    F.epsPhase = P.epsPhase; F.phaseDiffFac = P.phaseDiffFac; F.penaltyFac = P.penaltyFac; F.tmelt = P.tmelt;
    F.LFfac = P.LFfac; F.opm = P.opm; F.ViscHeatFac = P.ViscHeatFac; F.OhmLossFac = P.OhmLossFac; F.oek = P.oek; F.po = P.po;
    F.prec_angle = P.prec_angle; F.dilution_fac = P.dilution_fac; F.ra = P.ra; F.opr = P.opr; F.omega_ma = P.omega_ma;
    F.omega_ic = P.omega_ic; F.r_cmb = P.r_cmb; F.r_icb = P.r_icb; F.courfac = P.courfac; F.alffac = P.alffac; F.time = x.time;
    a.gi = rl->gi; a.go = rl->go; a.gin = rl->buf.gin; a.gout = rl->buf.gout; a.n_lev = nl; a.nh = h->nh; a.n_phi = h->n_phi;
    a.minc = h->minc; a.lev = d_lev; a.sinth = h->d_sinth; a.costh = h->d_costh; a.courmax = rl->buf.courmax;
    a.ddw = P.l_full_sphere ? src[S_DDW] : nullptr;
    a.ddb = (P.l_full_sphere && (P.l_mag || P.l_mag_LF)) ? src[S_DDB] : nullptr;
    a.wgauss = h->d_wgauss; a.tq_partial = rl->d_tq_partial;
    a.lm_max = h->lm_max; a.lm10 = 1; a.lm11 = (h->minc == 1 && h->m_max >= 1) ? h->lstart[1] : -1;
    int gx = (int)((plane + NL_THREADS - 1) / NL_THREADS);
    const bool mag = P.l_mag || P.l_mag_LF || P.l_mag_nl;
    const bool extra = !P.l_adv_curl || P.l_anel || P.l_chemical_conv || P.l_precession || P.l_centrifuge || P.l_phase_field;
    launch_get_nl(a, mag, extra, gx, nl, h->stream);
    courant_finish_kernel<<<(nl + 127) / 128, 128, 0, h->stream>>>(rl->buf.courmax, d_lev, nl, x.dtrkc + l0, x.dthkc + l0,
                                                                  rl->d_tq_partial, gx, P.LFfac, rl->d_torque);
    h->launches += 2;
    cudaEventRecord(ev[4], h->stream);
    if (run_analysis(h, rl->spec, L, rl->buf, d_lev, ev + 5, false)) return 1;  // ev[5..7]; extraction is fused below
    // ---- get_td
    TdArgs t{};
    t.f.l_conv = P.l_conv; t.f.l_mag = P.l_mag; t.f.l_heat = P.l_heat; t.f.l_conv_nl = P.l_conv_nl; t.f.l_mag_nl = P.l_mag_nl;
    t.f.l_mag_kin = P.l_mag_kin; t.f.l_anel = P.l_anel; t.f.l_corr = P.l_corr; t.f.l_double_curl = P.l_double_curl;
    t.f.l_single_matrix = P.l_single_matrix; t.f.l_chemical_conv = P.l_chemical_conv; t.f.l_anelastic_liquid = P.l_anelastic_liquid; t.f.l_phase_field = P.l_phase_field;
    t.f.CorFac = P.CorFac; t.f.epsc = P.epsc; t.f.epscXi = P.epscXi;
    t.n_lev = nl; t.lm_max = h->lm_max; t.l_max = h->l_max; t.minc = h->minc; t.lm2l = h->d_lm2l; t.lm2m = h->d_lm2m; t.lev = d_lev;
    // tile slots of the nonlinear_lm_t members inside the fused kernel: scalar-class columns first, then vector columns
    const int nfs = (int)rl->spec.afield_s.size();
    auto ns = [&](int slot) { return slot; };
    auto nv = [&](int pair, int comp) { return pair < 0 ? -1 : nfs + 2 * pair + comp; };
    TdSlots sl;
    sl.s[0] = ns(rl->a_Advr); sl.s[1] = nv(rl->a_Adv, 0); sl.s[2] = nv(rl->a_Adv, 1);
    sl.s[3] = ns(rl->a_VxBr); sl.s[4] = nv(rl->a_VxB, 0); sl.s[5] = nv(rl->a_VxB, 1);
    sl.s[6] = nv(rl->a_VS, 0); sl.s[7] = ns(rl->a_VSr);
    sl.s[8] = nv(rl->a_VXi, 0); sl.s[9] = ns(rl->a_VXir);
    sl.s[10] = ns(rl->a_heat);
    sl.s[11] = ns(rl->a_phi);
    t.w = src[S_W]; t.dw = src[S_DW]; t.ddw = src[S_DDW]; t.z = src[S_Z]; t.dz = src[S_DZ];
    auto o = [&](int i) -> double * { return x.op[i] ? x.op[i] + (size_t)l0 * lm2 : nullptr; };
    t.dwdt = o(O_DWDT); t.dzdt = o(O_DZDT); t.dpdt = o(O_DPDT); t.dsdt = o(O_DSDT); t.dxidt = o(O_DXIDT); t.dbdt = o(O_DBDT);
    t.djdt = o(O_DJDT); t.dVxVhLM = o(O_DVXVH); t.dVxBhLM = o(O_DVXBH); t.dVSrLM = o(O_DVSR); t.dVXirLM = o(O_DVXIR); t.dphidt = o(O_DPHIDT);
    {
        ExtractArgs e = make_extract_args(h, L, rl->buf, d_lev);
        e.out_s = nullptr; e.out_v = nullptr;
        const int nf = e.nf_s + 2 * e.npair;
        // widest tile whose shared-memory footprint still lets several CTAs share an SM
        auto smem = [&](int tl) { return (size_t)nf * nl * (tl + 1) * sizeof(double2); };
        if (smem(32) <= 80 * 1024) extract_td_kernel<32><<<(h->lm_max + 31) / 32, 256, smem(32), h->stream>>>(e, t, sl);
        else if (smem(16) <= 80 * 1024) extract_td_kernel<16><<<(h->lm_max + 15) / 16, 256, smem(16), h->stream>>>(e, t, sl);
        else if (smem(8) <= 200 * 1024) extract_td_kernel<8><<<(h->lm_max + 7) / 8, 256, smem(8), h->stream>>>(e, t, sl);
        else MFAIL("magic_rloop: level chunk too large for the fused get_td tile; lower level_chunk");
        h->launches++;
    }
    cudaEventRecord(ev[9], h->stream);
    MCHECK(cudaGetLastError());
    return 0;
}

static int rloop_end(magic_rloop *rl, const magic_fields_in *in) {
    magic_sht *h = rl->h;
    const size_t lm2 = 2 * (size_t)h->lm_max;
    for (int w = 0; w < 2; w++) {  // get_br_v_bcs on the boundary levels (rIter.f90:267-277)
        const int i = rl->bc_lev[w];
        if (i < 0) continue;
        if (!in->b || !in->dw || !in->z) MFAIL("magic_rloop_run: get_br_v_bcs needs b, dw and z");
        const size_t off = (size_t)i * lm2, o2 = (size_t)w * 2 * lm2;
        const LevelInfo &Lb = rl->lev[i];
        if (br_v_bcs_dev(h, in->b + off, in->dw + off, in->z + off, Lb.lcut, Lb.or2 * Lb.orho1, w == 0 ? rl->p.omega_ma : rl->p.omega_ic,
                         rl->d_brv + o2, rl->d_brv + o2 + lm2))
            return 1;
    }
    if (rl->d_brv) MCHECK(cudaMemcpyAsync(rl->h_brv, rl->d_brv, sizeof(double) * 8 * (size_t)h->lm_max, cudaMemcpyDeviceToHost, h->stream));
    MCHECK(cudaMemcpyAsync(rl->h_torque, rl->d_torque, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stream));
    cudaEventRecord(rl->ev[14], h->stream);
    MCHECK(cudaEventSynchronize(rl->ev[14]));
    rl->torque[0] = rl->h_torque[0];
    rl->torque[1] = rl->h_torque[1];
    // per-stage device times of the run: 0 total, 1 prep, 2 Legendre synthesis, 3 c2r, 4 get_nl, 5 r2c, 6 Legendre analysis, 7 get_td
    double stage[8] = {0};
    float ms;
    const int pairs[7][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {5, 6}, {6, 7}, {7, 9}};
    for (size_t c = 0; c < rl->chunk_start.size(); c++)
        for (int s = 0; s < 7; s++) {
            cudaEventElapsedTime(&ms, rl->cev[10 * c + pairs[s][0]], rl->cev[10 * c + pairs[s][1]]);
            stage[s + 1] += ms;
        }
    cudaEventElapsedTime(&ms, rl->ev[15], rl->ev[14]);
    stage[0] = ms;
    for (int i = 0; i < 8; i++) rl->timing[i] = stage[i];
    // what the pipelined calls could not hide: run start -> first kernel of the first chunk, last kernel of the last chunk -> end
    const size_t nc = rl->chunk_start.size();
    cudaEventElapsedTime(&ms, rl->ev[15], rl->cev[0]);
    rl->exposed[0] = ms;
    cudaEventElapsedTime(&ms, rl->cev[10 * (nc - 1) + 9], rl->ev[14]);
    rl->exposed[1] = ms;
    return 0;
}

extern "C" int magic_rloop_run_dev(magic_rloop *rl, const magic_fields_in *in, const magic_fields_out *out, double time) {
    if (!rl || !in || !out) MFAIL("magic_rloop_run_dev: null argument");
    magic_sht *h = rl->h;
    RunCtx x;
    if (rloop_begin(rl, in, out, time, x)) return 1;
    const size_t lm2 = 2 * (size_t)h->lm_max;
    for (size_t c = 0; c < rl->chunk_start.size(); c++) {
        const int l0 = rl->chunk_start[c], nl = rl->chunk_size[c];
        if (rl->pipelined) MCHECK(cudaStreamWaitEvent(h->stream, rl->up_done[c], 0));
        if (rloop_chunk(rl, (int)c, x)) return 1;
        if (rl->pipelined) {  // results of this chunk go home while the next chunk computes
            MCHECK(cudaEventRecord(rl->comp_done[c], h->stream));
            MCHECK(cudaStreamWaitEvent(rl->s_down, rl->comp_done[c], 0));
            const size_t off = (size_t)l0 * lm2, bytes = sizeof(double) * (size_t)nl * lm2;
            for (int i = 0; i < O_COUNT; i++)
                if (rl->need_out[i]) MCHECK(cudaMemcpyAsync(rl->host_out[i] + off, x.op[i] + off, bytes, cudaMemcpyDeviceToHost, rl->s_down));
            MCHECK(cudaMemcpyAsync(rl->host_dtrkc + l0, x.dtrkc + l0, sizeof(double) * nl, cudaMemcpyDeviceToHost, rl->s_down));
            MCHECK(cudaMemcpyAsync(rl->host_dthkc + l0, x.dthkc + l0, sizeof(double) * nl, cudaMemcpyDeviceToHost, rl->s_down));
        }
    }
    return rloop_end(rl, in);
}

extern "C" int magic_rloop_run(magic_rloop *rl, const magic_fields_in *in, const magic_fields_out *out, double time) {
    if (!rl || !in || !out) MFAIL("magic_rloop_run: null argument");
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    const size_t fbytes = sizeof(double) * 2 * (size_t)h->lm_max * rl->n_r_loc;
    const double *ip[S_COUNT];
    double *op[O_COUNT];
    in_ptrs(in, ip);
    out_ptrs(out, op);
    magic_fields_in din{};
    magic_fields_out dout{};
    const double *dip[S_COUNT] = {nullptr};
    const size_t lm2 = 2 * (size_t)h->lm_max;
    for (int i = 0; i < S_COUNT; i++) {
        if (!rl->need_in[i]) continue;
        if (!ip[i]) MFAIL("magic_rloop_run: a required input field is null");
        if (!rl->d_in[i]) MCHECK(cudaMalloc((void **)&rl->d_in[i], fbytes));
        rl->host_in[i] = ip[i];
        dip[i] = rl->d_in[i];
    }
    // all uploads are queued now, chunk by chunk, on their own stream; the compute stream waits per chunk
    for (size_t c = 0; c < rl->chunk_start.size(); c++) {
        const size_t off = (size_t)rl->chunk_start[c] * lm2, bytes = sizeof(double) * (size_t)rl->chunk_size[c] * lm2;
        for (int i = 0; i < S_COUNT; i++)
            if (rl->need_in[i]) MCHECK(cudaMemcpyAsync(rl->d_in[i] + off, ip[i] + off, bytes, cudaMemcpyHostToDevice, rl->s_up));
        MCHECK(cudaEventRecord(rl->up_done[c], rl->s_up));
    }
    din.w = dip[S_W]; din.dw = dip[S_DW]; din.ddw = dip[S_DDW]; din.z = dip[S_Z]; din.dz = dip[S_DZ]; din.s = dip[S_S]; din.ds = dip[S_DS];
    din.p = dip[S_P]; din.xi = dip[S_XI]; din.b = dip[S_B]; din.db = dip[S_DB]; din.ddb = dip[S_DDB]; din.aj = dip[S_AJ]; din.dj = dip[S_DJ]; din.phi = dip[S_PHI];
    double *dop[O_COUNT] = {nullptr};
    for (int i = 0; i < O_COUNT; i++) {
        if (!rl->need_out[i]) continue;
        if (!op[i]) MFAIL("magic_rloop_run: a required output field is null");
        if (!rl->d_out[i]) {
            MCHECK(cudaMalloc((void **)&rl->d_out[i], fbytes));
            MCHECK(cudaMemsetAsync(rl->d_out[i], 0, fbytes, h->stream));
        }
        rl->host_out[i] = op[i];
        dop[i] = rl->d_out[i];
    }
    dout.dwdt = dop[O_DWDT]; dout.dzdt = dop[O_DZDT]; dout.dpdt = dop[O_DPDT]; dout.dsdt = dop[O_DSDT]; dout.dxidt = dop[O_DXIDT];
    dout.dbdt = dop[O_DBDT]; dout.djdt = dop[O_DJDT]; dout.dVxVhLM = dop[O_DVXVH]; dout.dVxBhLM = dop[O_DVXBH];
    dout.dVSrLM = dop[O_DVSR]; dout.dVXirLM = dop[O_DVXIR]; dout.dphidt = dop[O_DPHIDT];
    dout.dtrkc = rl->d_dtrkc; dout.dthkc = rl->d_dthkc;
    rl->pipelined = true;
    int rc = magic_rloop_run_dev(rl, &din, &dout, time);
    rl->pipelined = false;
    if (rc) return 1;
    MCHECK(cudaStreamSynchronize(rl->s_down));
    MCHECK(cudaStreamSynchronize(h->stream));
    memcpy(out->dtrkc, rl->host_dtrkc, sizeof(double) * rl->n_r_loc);
    memcpy(out->dthkc, rl->host_dthkc, sizeof(double) * rl->n_r_loc);
    return 0;
}

// Page-locking of caller-owned host arrays is explicit (opt in, opt out): the host registers its PERSISTENT containers once
// (the Fortran shim does so for the arrays of fields.f90 / dt_fieldsLast.f90) and must unpin them before freeing them.  The
// run calls never pin anything themselves: a temporary that is freed while still registered would leave a stale
// registration behind, and a later allocation at the same address would be treated as pinned against the old pages.
extern "C" int magic_rloop_pin_host(magic_rloop *rl, const void *ptr, size_t bytes) {
    if (!rl || !ptr || bytes == 0) MFAIL("magic_rloop_pin_host: null argument");
    MCHECK(cudaSetDevice(rl->h->dev));
    for (const auto &r : rl->registered)
        if (r.first == ptr) {
            if (r.second == bytes) return 0;
            MFAIL("magic_rloop_pin_host: this address is already pinned with another size; unpin it first");
        }
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type == cudaMemoryTypeHost) return 0;  // pinned by its owner
    cudaGetLastError();
    MCHECK(cudaHostRegister((void *)ptr, bytes, cudaHostRegisterDefault));
    rl->registered.emplace_back(ptr, bytes);
    return 0;
}
extern "C" int magic_rloop_unpin_host(magic_rloop *rl, const void *ptr) {
    if (!rl || !ptr) MFAIL("magic_rloop_unpin_host: null argument");
    MCHECK(cudaSetDevice(rl->h->dev));
    for (size_t i = 0; i < rl->registered.size(); i++)
        if (rl->registered[i].first == ptr) {
            MCHECK(cudaHostUnregister((void *)ptr));
            rl->registered.erase(rl->registered.begin() + i);
            return 0;
        }
    MFAIL("magic_rloop_unpin_host: this address was not pinned through magic_rloop_pin_host");
}


// ---- LM-distributed containers in, LM-distributed explicit terms out: the whole of step_time.f90:485-612 in one call,
//      the transposes (and, for host containers, the PCIe transfers) pipelined level chunk by level chunk ----------------
// Containers (fields.f90:211-268, dt_fieldsLast.f90:125-214), k = 0..3:
//   in : flow(w,dw,ddw,z,dz)  s(s,ds)  field(b,db,ddb,aj,dj)  xi(xi,dxi)
//   out: dflowdt(dwdt,dzdt,dpdt[,dVxVhLM])  dsdt(dsdt,dVSrLM)  dbdt(dbdt,djdt,dVxBhLM)  dxidt(dxidt,dVXirLM)
struct LmPipe {
    magic_transp *parent = nullptr;
    std::vector<magic_transp *> parts;  // one per global chunk index
    cudaStream_t comm = nullptr, s_up = nullptr, s_down = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_out, ev_up, ev_lmout;
    cudaEvent_t ev_start = nullptr, ev_done = nullptr, ev_ready = nullptr;
    int nf_in[4] = {0, 0, 0, 0}, nf_out[4] = {0, 0, 0, 0};   // container widths as the host declares them (0 = absent)
    int nt_in[4] = {0, 0, 0, 0};                             // leading fields of each inbound container the loop really reads
    double *R_in[4] = {nullptr, nullptr, nullptr, nullptr}, *R_out[4] = {nullptr, nullptr, nullptr, nullptr};
    double *LM_in[4] = {nullptr, nullptr, nullptr, nullptr}, *LM_out[4] = {nullptr, nullptr, nullptr, nullptr};  // host-pointer entry only
    int C = 0, n_procs = 1, rank = 0, n_r_max = 0, nlm = 0;
    std::vector<int> rs;                   // 0-based first level of every rank
    std::vector<std::vector<int>> cs, cz;  // level chunks of every rank (start inside the slab, size)
    // radial matrices of the host's radial scheme (magic_rloop_set_radial_matrices) for the LM-side prologue / epilogue
    double *d_D1 = nullptr, *d_D2 = nullptr;
    int ldD = 0;
    GemmProb *d_dprobs[2] = {nullptr, nullptr};  // cached tile lists of the prologue (0) and epilogue (1) matrix products
    int2 *d_dtiles[2] = {nullptr, nullptr};
    int n_dtiles[2] = {0, 0};
    double *d_work = nullptr;              // [3][n_r_max][nlm] radial derivatives of dVSrLM, dVxBhLM, dVxVhLM
    double *d_lmrad = nullptr;             // [4][n_r_max]: or2, orho1, dentropy0, l_R (as doubles)
    int *d_lo2l = nullptr;                 // degree of every local mode (lo order)
};

static void lmpipe_free(LmPipe *p) {
    if (!p) return;
    for (auto t : p->parts) magic_transp_destroy(t);
    for (auto *v : {&p->ev_in, &p->ev_out, &p->ev_up, &p->ev_lmout})
        for (auto e : *v) cudaEventDestroy(e);
    for (cudaEvent_t e : {p->ev_start, p->ev_done, p->ev_ready})
        if (e) cudaEventDestroy(e);
    for (cudaStream_t s : {p->comm, p->s_up, p->s_down})
        if (s) cudaStreamDestroy(s);
    for (int i = 0; i < 4; i++) { cudaFree(p->R_in[i]); cudaFree(p->R_out[i]); cudaFree(p->LM_in[i]); cudaFree(p->LM_out[i]); }
    cudaFree(p->d_D1); cudaFree(p->d_D2); cudaFree(p->d_work); cudaFree(p->d_lmrad);
    for (int i = 0; i < 2; i++) { cudaFree(p->d_dprobs[i]); cudaFree(p->d_dtiles[i]); }
    cudaFree(p->d_lo2l);
    delete p;
}

static int lmpipe_build(magic_rloop *rl, magic_transp *t) {
    magic_sht *h = rl->h;
    const magic_params &P = rl->p;
    int rank, n_procs, n_r_max, nf;
    if (magic_transp_info(t, &rank, &n_procs, &n_r_max, &nf)) return 1;
    if (nf < 5) MFAIL("magic_rloop_run_lm: the transposer must serve containers of 5 fields");
    int llm, ulm, nRstart, nRstop;
    if (magic_transp_extents(t, &llm, &ulm, &nRstart, &nRstop)) return 1;
    if (nRstop - nRstart + 1 != rl->n_r_loc) MFAIL("magic_rloop_run_lm: the loop and the transposer disagree on the radial slab");
    LmPipe *old = rl->lmpipe;
    LmPipe *p = new LmPipe();
    if (old) lmpipe_free(old);
    rl->lmpipe = p;
    p->parent = t; p->n_procs = n_procs; p->rank = rank; p->n_r_max = n_r_max; p->nlm = ulm - llm + 1;
    const bool flow = P.l_conv || P.l_mag_kin, mag = P.l_mag || P.l_mag_LF;
    if (!flow) MFAIL("magic_rloop_run_lm: a run without the flow containers is not supported");
    p->nf_in[0] = 5; p->nt_in[0] = 5;
    if (P.l_heat) { p->nf_in[1] = 2; p->nt_in[1] = 1; }           // ds is not read by the radial loop
    if (mag) { p->nf_in[2] = 5; p->nt_in[2] = 5; }
    if (P.l_chemical_conv) { p->nf_in[3] = 2; p->nt_in[3] = 1; }
    if (P.l_conv) p->nf_out[0] = P.l_double_curl ? 4 : 3;
    if (P.l_heat) p->nf_out[1] = 2;
    if (P.l_mag) p->nf_out[2] = 3;
    if (P.l_chemical_conv) p->nf_out[3] = 2;
    const size_t fld = 2 * (size_t)h->lm_max * rl->n_r_loc;
    for (int k = 0; k < 4; k++) {
        if (p->nf_in[k]) {
            MCHECK(cudaMalloc((void **)&p->R_in[k], sizeof(double) * fld * p->nf_in[k]));
            MCHECK(cudaMemsetAsync(p->R_in[k], 0, sizeof(double) * fld * p->nf_in[k], h->stream));
        }
        if (p->nf_out[k]) {
            MCHECK(cudaMalloc((void **)&p->R_out[k], sizeof(double) * fld * p->nf_out[k]));
            MCHECK(cudaMemsetAsync(p->R_out[k], 0, sizeof(double) * fld * p->nf_out[k], h->stream));
        }
    }
    // the level chunks of every rank: same rule and same (unclamped) level_chunk everywhere, tapered when the transposes
    // cross NVLink (MAGIC_LM_TAPER, default 4 levels; 0 = off)
    std::vector<int> rs(n_procs), re(n_procs);
    if (magic_get_blocks(n_r_max, n_procs, rs.data(), re.data())) return 1;
    int taper = n_procs > 1 ? 4 : 0;
    if (const char *e = getenv("MAGIC_LM_TAPER")) taper = atoi(e);
    p->rs.resize(n_procs);
    p->cs.assign(n_procs, {});
    p->cz.assign(n_procs, {});
    for (int q = 0; q < n_procs; q++) {
        p->rs[q] = rs[q] - 1;
        level_chunks_tapered(re[q] - rs[q] + 1, rl->level_chunk_req, taper, p->cs[q], p->cz[q]);
        p->C = std::max(p->C, (int)p->cs[q].size());
    }
    if (p->cs[rank] != rl->chunk_start || p->cz[rank] != rl->chunk_size)
        if (rloop_set_chunks(rl, p->cs[rank], p->cz[rank])) return 1;
    {   // highest priority: the compute kernels fill every SM with long grids, so the pack / exchange / unpack CTAs of the
        // communication stream must be picked first whenever a slot frees up or they trail behind the chunk they serve
        int lo = 0, hi = 0;
        MCHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        MCHECK(cudaStreamCreateWithPriority(&p->comm, cudaStreamNonBlocking, hi));
        MCHECK(cudaStreamCreateWithFlags(&p->s_up, cudaStreamNonBlocking));
        MCHECK(cudaStreamCreateWithFlags(&p->s_down, cudaStreamNonBlocking));
    }
    MCHECK(cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming));
    MCHECK(cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
    MCHECK(cudaEventCreateWithFlags(&p->ev_ready, cudaEventDisableTiming));
    for (int c = 0; c < p->C; c++) {
        std::vector<int> off(n_procs), cnt(n_procs);
        for (int q = 0; q < n_procs; q++) {
            const bool has = c < (int)p->cs[q].size();
            off[q] = has ? p->cs[q][c] : re[q] - rs[q] + 1;
            cnt[q] = has ? p->cz[q][c] : 0;
        }
        magic_transp *part = nullptr;
        if (magic_transp_create_part(t, off.data(), cnt.data(), &part)) return 1;
        magic_transp_set_stream(part, (void *)p->comm);
        p->parts.push_back(part);
        for (auto *v : {&p->ev_in, &p->ev_out, &p->ev_up, &p->ev_lmout}) {
            cudaEvent_t e;
            MCHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            v->push_back(e);
        }
    }
    return 0;
}

// host <-> device rows of part c of an LM-distributed container ([nf][n_r_max][nlm] complex): for every rank q the levels of
// q's c-th chunk, for the fields in `mask` (runs of consecutive fields of one rank go in one strided copy)
static int lm_rows_copy(LmPipe *p, int c, int nf, unsigned mask, double *dst, const double *src, cudaMemcpyKind kind, cudaStream_t st) {
    const size_t row = sizeof(double) * 2 * (size_t)p->nlm, pitch = row * p->n_r_max;
    for (int q = 0; q < p->n_procs; q++) {
        if (c >= (int)p->cs[q].size() || p->cz[q][c] == 0) continue;
        const size_t off = (size_t)(p->rs[q] + p->cs[q][c]) * 2 * (size_t)p->nlm;
        for (int f = 0; f < nf;) {
            if (!(mask >> f & 1u)) { f++; continue; }
            int f1 = f;
            while (f1 < nf && (mask >> f1 & 1u)) f1++;
            const size_t fo = off + (size_t)f * 2 * (size_t)p->nlm * p->n_r_max;
            MCHECK(cudaMemcpy2DAsync(dst + fo, pitch, src + fo, pitch, row * p->cz[q][c], f1 - f, kind, st));
            f = f1;
        }
    }
    return 0;
}

// ---- radial-matrix products on LM-distributed device arrays: Y[n_r x 2 nlm] = D[n_r x n_r] X[n_r x 2 nlm] with the Legendre
//      GEMM kernel (D zero-padded to whole tiles; X must be followed by >= 15 readable rows: the containers this file
//      allocates carry that slack).  get_dr / get_ddr of radial_derivatives.f90 for whatever radial scheme the host runs.
struct DerivJob { const double *X; double *Y; int which; };

static int lm_matrix_upload(magic_rloop *rl, LmPipe *p) {
    if (p->d_D1 || rl->n_r_mat == 0) return 0;
    const int n = rl->n_r_mat;
    if (n != p->n_r_max) MFAIL("magic_rloop: the radial matrices were given for another n_r_max than the transposer's");
    p->ldD = pad_up(n, BK);
    const int rows = pad_up(n, GEMM_BM) + GEMM_BM;
    std::vector<double> pad((size_t)rows * p->ldD, 0.0);
    for (int which = 0; which < 2; which++) {
        const std::vector<double> &D = which == 0 ? rl->D1h : rl->D2h;
        std::fill(pad.begin(), pad.end(), 0.0);
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) pad[(size_t)i * p->ldD + j] = D[(size_t)i * n + j];
        if (dev_upload_vec(which == 0 ? &p->d_D1 : &p->d_D2, pad)) return 1;
    }
    if (!rl->lmrad_h.empty() && dev_upload_vec(&p->d_lmrad, rl->lmrad_h)) return 1;
    // degree and order of every local mode (lo order)
    std::vector<int> lo2st(rl->h->lm_max), ls(p->n_procs), le(p->n_procs);
    if (magic_lo_map(rl->h->l_max, rl->h->m_max, rl->h->minc, p->n_procs, lo2st.data(), ls.data(), le.data())) return 1;
    std::vector<int> ll(2 * (size_t)p->nlm);
    for (int i = 0; i < p->nlm; i++) {
        const int st = lo2st[ls[p->rank] - 1 + i];
        ll[i] = rl->h->lm2l[st];
        ll[(size_t)p->nlm + i] = rl->h->lm2m[st];
    }
    if (dev_upload_vec(&p->d_lo2l, ll)) return 1;
    MCHECK(cudaMalloc((void **)&p->d_work, sizeof(double) * (2 * (size_t)p->nlm * p->n_r_max * 4 + 64 * (size_t)p->nlm + 256)));
    MCHECK(cudaMemset(p->d_work, 0, sizeof(double) * (2 * (size_t)p->nlm * p->n_r_max * 4 + 64 * (size_t)p->nlm + 256)));
    return 0;
}

static int lm_matrix_products(magic_rloop *rl, LmPipe *p, int slot, const std::vector<DerivJob> &jobs, cudaStream_t st) {
    if (jobs.empty()) return 0;
    if (p->d_dprobs[slot]) {  // the arrays of a slot never move: its tile list is built once
        launch_legendre_gemm(true, p->d_dprobs[slot], p->d_dtiles[slot], p->n_dtiles[slot], p->ldD, st);
        rl->h->launches++;
        MCHECK(cudaGetLastError());
        return 0;
    }
    const int n = p->n_r_max, N = 2 * p->nlm;
    std::vector<GemmProb> probs;
    std::vector<int2> tiles;
    for (const DerivJob &j : jobs) {
        GemmProb g{};
        g.A0 = j.which == 1 ? p->d_D1 : p->d_D2;
        g.B = j.X; g.C = j.Y;
        g.kt0 = p->ldD / BK; g.M = n; g.ldb = g.ldc = N; g.Nvalid = g.Nstore = N; g.Mlo = 0; g.klo = 0; g.ks0 = nullptr;
        probs.push_back(g);
    }
    for (int mt = 0; mt < (n + GEMM_BM - 1) / GEMM_BM; mt++)
        for (int nt = 0; nt < (N + GEMM_BN - 1) / GEMM_BN; nt++)
            for (size_t pid = 0; pid < probs.size(); pid++) tiles.push_back(make_int2((int)pid, (mt << 16) | nt));
    if ((N + GEMM_BN - 1) / GEMM_BN > 0xffff) MFAIL("magic_rloop: too many local modes for the radial-matrix GEMM tile index");
    if (dev_upload_vec(&p->d_dprobs[slot], probs) || dev_upload_vec(&p->d_dtiles[slot], tiles)) return 1;
    p->n_dtiles[slot] = (int)tiles.size();
    return lm_matrix_products(rl, p, slot, jobs, st);
}

struct LmHostIO {  // host containers of magic_rloop_run_lm (null for the device-pointer call)
    const double *in[4];
    double *out[4];
};

static int lm_run(magic_rloop *rl, magic_transp *t, const double *const lm_in[4], double *const lm_out[4], double *dtrkc, double *dthkc,
                  double time, const LmHostIO *io) {
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    if (!rl->lmpipe || rl->lmpipe->parent != t)
        if (lmpipe_build(rl, t)) return 1;
    LmPipe *p = rl->lmpipe;
    const magic_params &P = rl->p;
    if (P.l_phase_field) MFAIL("magic_rloop_run_lm: the phase field has no LM-distributed container here; use the R-distributed calls");
    for (int k = 0; k < 4; k++) {
        if (p->nf_in[k] && !lm_in[k]) MFAIL("magic_rloop_run_lm: a required inbound container is null");
        if (p->nf_out[k] && !lm_out[k]) MFAIL("magic_rloop_run_lm: a required outbound container is null");
    }
    const size_t fld = 2 * (size_t)h->lm_max * rl->n_r_loc;  // doubles per R-distributed field
    magic_fields_in fin{};
    magic_fields_out fout{};
    fin.w = p->R_in[0]; fin.dw = p->R_in[0] + fld; fin.ddw = p->R_in[0] + 2 * fld; fin.z = p->R_in[0] + 3 * fld; fin.dz = p->R_in[0] + 4 * fld;
    if (p->nf_in[1]) { fin.s = p->R_in[1]; fin.ds = p->R_in[1] + fld; }
    if (p->nf_in[2]) { fin.b = p->R_in[2]; fin.db = p->R_in[2] + fld; fin.ddb = p->R_in[2] + 2 * fld; fin.aj = p->R_in[2] + 3 * fld; fin.dj = p->R_in[2] + 4 * fld; }
    if (p->nf_in[3]) fin.xi = p->R_in[3];
    if (p->nf_out[0]) {
        fout.dwdt = p->R_out[0]; fout.dzdt = p->R_out[0] + fld; fout.dpdt = p->R_out[0] + 2 * fld;
        if (P.l_double_curl) fout.dVxVhLM = p->R_out[0] + 3 * fld;
    }
    if (p->nf_out[1]) { fout.dsdt = p->R_out[1]; fout.dVSrLM = p->R_out[1] + fld; }
    if (p->nf_out[2]) { fout.dbdt = p->R_out[2]; fout.djdt = p->R_out[2] + fld; fout.dVxBhLM = p->R_out[2] + 2 * fld; }
    if (p->nf_out[3]) { fout.dxidt = p->R_out[3]; fout.dVXirLM = p->R_out[3] + fld; }
    fout.dtrkc = dtrkc; fout.dthkc = dthkc;
    RunCtx x;
    if (rloop_begin(rl, &fin, &fout, time, x)) return 1;
    // the transposes start once everything queued on the compute stream so far (the previous step) is done
    MCHECK(cudaEventRecord(p->ev_start, h->stream));
    MCHECK(cudaStreamWaitEvent(p->comm, p->ev_start, 0));
    if (io) {
        MCHECK(cudaStreamWaitEvent(p->s_up, p->ev_start, 0));
        MCHECK(cudaStreamWaitEvent(p->s_down, p->ev_start, 0));
    }
    // ---- options of the host-container call
    const bool derivs = io && rl->lm_derivs, finish = io && rl->lm_finish;
    if ((rl->lm_derivs || rl->lm_finish) && !io) MFAIL("magic_rloop_run_lm_dev: the LM-side prologue / epilogue options belong to the host-container call magic_rloop_run_lm");
    if (derivs || finish) {
        if (rl->n_r_mat == 0) MFAIL("magic_rloop_run_lm: set the radial matrices first (magic_rloop_set_radial_matrices)");
        if (finish && rl->lmrad_h.empty()) MFAIL("magic_rloop_run_lm: set the LM-side radial functions first (magic_rloop_set_lm_radial)");
        if (lm_matrix_upload(rl, p)) return 1;
    }
    const size_t lmf = 2 * (size_t)p->nlm * p->n_r_max;  // doubles per LM-distributed field
    // fields that cross PCIe part by part: up = what the loop reads (ds / dxi never; with the prologue on the device only s / xi),
    // down = every explicit term (with the epilogue on the device not the arrays it consumes or finishes)
    unsigned up_mask[4] = {0x1fu, 0x1u, 0x1fu, 0x1u}, down_mask[4] = {0xfu, 0x3u, 0x7u, 0x3u}, late_mask[4] = {0, 0, 0, 0};
    if (derivs) up_mask[0] = up_mask[2] = 0;
    if (finish) {
        if (P.l_heat) { down_mask[1] = 0; late_mask[1] = 0x1u; }                       // dsdt finished, dVSrLM consumed
        if (P.l_chemical_conv) { down_mask[3] = 0; late_mask[3] = 0x1u; }              // dxidt, dVXirLM
        if (P.l_mag) { down_mask[2] = 0x1u; late_mask[2] = 0x2u; }                     // dbdt as is; djdt finished, dVxBhLM consumed
        if (P.l_double_curl) { down_mask[0] = 0x6u; late_mask[0] = 0x1u; }            // dzdt, dpdt as is; dwdt finished, dVxVhLM consumed
    }
    if (derivs) {
        // prologue: w, z (b, aj) of ALL levels go up first; dw, ddw, dz (db, ddb, dj) are radial-matrix products on the device
        for (int k = 0; k < 4; k += 2) {
            if (!p->nf_in[k]) continue;
            for (int f = 0; f < 4; f += 3)
                MCHECK(cudaMemcpyAsync(p->LM_in[k] + f * lmf, io->in[k] + f * lmf, sizeof(double) * lmf, cudaMemcpyHostToDevice, p->s_up));
        }
        MCHECK(cudaEventRecord(p->ev_ready, p->s_up));
        MCHECK(cudaStreamWaitEvent(h->stream, p->ev_ready, 0));
        std::vector<DerivJob> jobs;
        for (int k = 0; k < 4; k += 2) {
            if (!p->nf_in[k]) continue;
            double *b = p->LM_in[k];
            jobs.push_back({b, b + lmf, 1});
            jobs.push_back({b, b + 2 * lmf, 2});
            jobs.push_back({b + 3 * lmf, b + 4 * lmf, 1});
        }
        if (lm_matrix_products(rl, p, 0, jobs, h->stream)) return 1;
        MCHECK(cudaEventRecord(p->ev_ready, h->stream));
        MCHECK(cudaStreamWaitEvent(p->comm, p->ev_ready, 0));
    }
    auto inbound = [&](int c) -> int {
        if (io) {  // PCIe: the rows of this part, only the fields the loop reads
            for (int k = 0; k < 4; k++)
                if (p->nf_in[k] && up_mask[k] && lm_rows_copy(p, c, p->nf_in[k], up_mask[k], p->LM_in[k], io->in[k], cudaMemcpyHostToDevice, p->s_up)) return 1;
            MCHECK(cudaEventRecord(p->ev_up[c], p->s_up));
            MCHECK(cudaStreamWaitEvent(p->comm, p->ev_up[c], 0));
        }
        for (int k = 0; k < 4; k++)
            if (p->nf_in[k] && magic_transp_lm2r_dev_n(p->parts[c], p->nt_in[k], lm_in[k], p->R_in[k])) return 1;
        MCHECK(cudaEventRecord(p->ev_in[c], p->comm));
        return 0;
    };
    auto outbound = [&](int c, bool computed) -> int {
        if (computed) {
            MCHECK(cudaEventRecord(p->ev_out[c], h->stream));
            MCHECK(cudaStreamWaitEvent(p->comm, p->ev_out[c], 0));
        }
        for (int k = 0; k < 4; k++)
            if (p->nf_out[k] && magic_transp_r2lm_dev_n(p->parts[c], p->nf_out[k], p->R_out[k], lm_out[k])) return 1;
        if (io) {
            MCHECK(cudaEventRecord(p->ev_lmout[c], p->comm));
            MCHECK(cudaStreamWaitEvent(p->s_down, p->ev_lmout[c], 0));
            for (int k = 0; k < 4; k++)
                if (p->nf_out[k] && down_mask[k] && lm_rows_copy(p, c, p->nf_out[k], down_mask[k], io->out[k], p->LM_out[k], cudaMemcpyDeviceToHost, p->s_down)) return 1;
        }
        return 0;
    };
    // Queue order on the communication stream -- identical on every rank, also for ranks with fewer chunks than C (they have
    // no levels in the late parts, their peers do): in(0) .. in(DIST-1), then per chunk c: in(c+DIST), [compute c], out(c).
    // MAGIC_LM_ALIGN=1: the transposes of an iteration (in(c+DIST), out(c-1)) start when chunk c enters its synthesis GEMM --
    // HBM-bound pack / unpack kernels then share the machine with a tensor-bound kernel instead of with the HBM-bound
    // operand assembly and FFTs.
    static int align = -1;
    if (align < 0) { const char *e = getenv("MAGIC_LM_ALIGN"); align = e ? atoi(e) : 0; }
    const int DIST = 2, nloc = (int)rl->chunk_start.size();
    for (int c = 0; c < std::min(DIST, p->C); c++)
        if (inbound(c)) return 1;
    if (!align) {
        for (int c = 0; c < p->C; c++) {
            if (c + DIST < p->C && inbound(c + DIST)) return 1;
            if (c < nloc) {
                MCHECK(cudaStreamWaitEvent(h->stream, p->ev_in[c], 0));
                if (rloop_chunk(rl, c, x)) return 1;
            }
            if (outbound(c, c < nloc)) return 1;
        }
    } else {
        for (int c = 0; c < p->C; c++) {
            if (c < nloc) {
                MCHECK(cudaStreamWaitEvent(h->stream, p->ev_in[c], 0));
                if (rloop_chunk(rl, c, x)) return 1;
                MCHECK(cudaEventRecord(p->ev_out[c], h->stream));
                MCHECK(cudaStreamWaitEvent(p->comm, rl->cev[10 * (size_t)c + 1], 0));  // chunk c has finished its operand assembly
            }
            if (c + DIST < p->C && inbound(c + DIST)) return 1;
            if (c > 0) {
                if (c - 1 < nloc) MCHECK(cudaStreamWaitEvent(p->comm, p->ev_out[c - 1], 0));
                if (outbound(c - 1, false)) return 1;
            }
        }
        if (p->C - 1 < nloc) MCHECK(cudaStreamWaitEvent(p->comm, p->ev_out[p->C - 1], 0));
        if (outbound(p->C - 1, false)) return 1;
    }
    MCHECK(cudaEventRecord(p->ev_done, p->comm));
    MCHECK(cudaStreamWaitEvent(h->stream, p->ev_done, 0));
    if (finish) {
        // epilogue = finish_explicit_assembly (LMLoop.f90:390-453): radial derivatives of dVSrLM, dVxBhLM, dVxVhLM, dVXirLM by
        // the radial matrix, then the point-wise completion of dsdt, djdt, dwdt, dxidt -- all levels of the local modes are here
        std::vector<DerivJob> jobs;
        double *wk = p->d_work;
        FinishArgs fa{};
        fa.n_r_max = p->n_r_max; fa.nlm = p->nlm; fa.lo2l = p->d_lo2l; fa.lo2m = p->d_lo2l + p->nlm;
        fa.or2 = p->d_lmrad; fa.orho1 = p->d_lmrad + p->n_r_max; fa.dentropy0 = p->d_lmrad + 2 * p->n_r_max; fa.l_R = p->d_lmrad + 3 * p->n_r_max;
        fa.w = p->LM_in[0];
        if (P.l_heat) { jobs.push_back({p->LM_out[1] + lmf, wk, 1}); fa.dsdt = p->LM_out[1]; fa.work_s = wk; }
        if (P.l_mag) { jobs.push_back({p->LM_out[2] + 2 * lmf, wk + lmf, 1}); fa.djdt = p->LM_out[2] + lmf; fa.work_b = wk + lmf; }
        if (P.l_double_curl) { jobs.push_back({p->LM_out[0] + 3 * lmf, wk + 2 * lmf, 1}); fa.dwdt = p->LM_out[0]; fa.work_v = wk + 2 * lmf; }
        if (P.l_chemical_conv) { jobs.push_back({p->LM_out[3] + lmf, wk + 3 * lmf, 1}); fa.dxidt = p->LM_out[3]; fa.work_xi = wk + 3 * lmf; }
        if (lm_matrix_products(rl, p, 1, jobs, h->stream)) return 1;
        finish_explicit_kernel<<<dim3((p->nlm + 255) / 256, p->n_r_max), 256, 0, h->stream>>>(fa);
        h->launches++;
        MCHECK(cudaGetLastError());
        MCHECK(cudaEventRecord(p->ev_ready, h->stream));
        MCHECK(cudaStreamWaitEvent(p->s_down, p->ev_ready, 0));
        for (int k = 0; k < 4; k++)
            for (int f = 0; f < p->nf_out[k]; f++)
                if (late_mask[k] >> f & 1u)
                    MCHECK(cudaMemcpyAsync(io->out[k] + f * lmf, p->LM_out[k] + f * lmf, sizeof(double) * lmf, cudaMemcpyDeviceToHost, p->s_down));
    }
    if (rloop_end(rl, &fin)) return 1;
    return 0;
}

extern "C" int magic_rloop_run_lm_dev(magic_rloop *rl, magic_transp *t, const magic_lm_in *in, const magic_lm_out *out, double time) {
    if (!rl || !t || !in || !out) MFAIL("magic_rloop_run_lm_dev: null argument");
    if (!out->dtrkc || !out->dthkc) MFAIL("magic_rloop_run_lm_dev: dtrkc/dthkc are null");
    const double *li[4] = {in->flow, in->s, in->field, in->xi};
    double *lo[4] = {out->dflowdt, out->dsdt, out->dbdt, out->dxidt};
    return lm_run(rl, t, li, lo, out->dtrkc, out->dthkc, time, nullptr);
}

// Host containers in, host containers out: what a type_mpicuda + rIter_cuda_t pair of the Fortran host calls instead of
// transp_LMloc_to_Rloc / radialLoopG / transp_Rloc_to_LMloc (step_time.f90:485-612).  Inside: H2D of the rows of level chunk
// c+2, NVLink transposes of chunk c+1, compute of chunk c, transposes and D2H of chunk c-1 all overlap.
extern "C" int magic_rloop_run_lm(magic_rloop *rl, magic_transp *t, const magic_lm_in *in, const magic_lm_out *out, double time) {
    if (!rl || !t || !in || !out) MFAIL("magic_rloop_run_lm: null argument");
    if (!out->dtrkc || !out->dthkc) MFAIL("magic_rloop_run_lm: dtrkc/dthkc are null");
    magic_sht *h = rl->h;
    MCHECK(cudaSetDevice(h->dev));
    if (!rl->lmpipe || rl->lmpipe->parent != t)
        if (lmpipe_build(rl, t)) return 1;
    LmPipe *p = rl->lmpipe;
    const size_t lmf = 2 * (size_t)p->nlm * p->n_r_max;  // doubles per LM-distributed field
    for (int k = 0; k < 4; k++) {
        if (p->nf_in[k] && !p->LM_in[k]) {
            MCHECK(cudaMalloc((void **)&p->LM_in[k], sizeof(double) * (lmf * p->nf_in[k] + 64 * (size_t)p->nlm + 256)));
            MCHECK(cudaMemsetAsync(p->LM_in[k], 0, sizeof(double) * (lmf * p->nf_in[k] + 64 * (size_t)p->nlm + 256), h->stream));
        }
        if (p->nf_out[k] && !p->LM_out[k]) {
            MCHECK(cudaMalloc((void **)&p->LM_out[k], sizeof(double) * (lmf * p->nf_out[k] + 64 * (size_t)p->nlm + 256)));
            MCHECK(cudaMemsetAsync(p->LM_out[k], 0, sizeof(double) * (lmf * p->nf_out[k] + 64 * (size_t)p->nlm + 256), h->stream));
        }
    }
    LmHostIO io;
    io.in[0] = in->flow; io.in[1] = in->s; io.in[2] = in->field; io.in[3] = in->xi;
    io.out[0] = out->dflowdt; io.out[1] = out->dsdt; io.out[2] = out->dbdt; io.out[3] = out->dxidt;
    for (int k = 0; k < 4; k++) {
        if (p->nf_in[k] && !io.in[k]) MFAIL("magic_rloop_run_lm: a required inbound container is null");
        if (p->nf_out[k] && !io.out[k]) MFAIL("magic_rloop_run_lm: a required outbound container is null");
    }
    const double *li[4] = {p->LM_in[0], p->LM_in[1], p->LM_in[2], p->LM_in[3]};
    double *lo[4] = {p->LM_out[0], p->LM_out[1], p->LM_out[2], p->LM_out[3]};
    if (lm_run(rl, t, li, lo, rl->d_dtrkc, rl->d_dthkc, time, &io)) return 1;
    MCHECK(cudaMemcpyAsync(rl->host_dtrkc, rl->d_dtrkc, sizeof(double) * rl->n_r_loc, cudaMemcpyDeviceToHost, h->stream));
    MCHECK(cudaMemcpyAsync(rl->host_dthkc, rl->d_dthkc, sizeof(double) * rl->n_r_loc, cudaMemcpyDeviceToHost, h->stream));
    MCHECK(cudaStreamSynchronize(p->s_down));
    MCHECK(cudaStreamSynchronize(h->stream));
    memcpy(out->dtrkc, rl->host_dtrkc, sizeof(double) * rl->n_r_loc);
    memcpy(out->dthkc, rl->host_dthkc, sizeof(double) * rl->n_r_loc);
    return 0;
}

extern "C" int magic_rloop_set_radial_matrices(magic_rloop *rl, int n_r_max, const double *D1, const double *D2) {
    if (!rl || !D1 || !D2 || n_r_max < 2) MFAIL("magic_rloop_set_radial_matrices: bad arguments");
    rl->n_r_mat = n_r_max;
    rl->D1h.assign(D1, D1 + (size_t)n_r_max * n_r_max);
    rl->D2h.assign(D2, D2 + (size_t)n_r_max * n_r_max);
    if (rl->lmpipe) { lmpipe_free(rl->lmpipe); rl->lmpipe = nullptr; }  // device copies are rebuilt with the next run
    return 0;
}
extern "C" int magic_rloop_set_lm_radial(magic_rloop *rl, int n_r_max, const double *or2, const double *orho1, const double *dentropy0,
                                         const int *l_R) {
    if (!rl || !or2 || !orho1 || !dentropy0 || !l_R || n_r_max < 2) MFAIL("magic_rloop_set_lm_radial: bad arguments");
    rl->lmrad_h.assign(4 * (size_t)n_r_max, 0.0);
    for (int i = 0; i < n_r_max; i++) {
        rl->lmrad_h[i] = or2[i];
        rl->lmrad_h[(size_t)n_r_max + i] = orho1[i];
        rl->lmrad_h[2 * (size_t)n_r_max + i] = dentropy0[i];
        rl->lmrad_h[3 * (size_t)n_r_max + i] = (double)l_R[i];
    }
    if (rl->lmpipe) { lmpipe_free(rl->lmpipe); rl->lmpipe = nullptr; }
    return 0;
}
extern "C" int magic_rloop_lm_options(magic_rloop *rl, int derivs_on_device, int finish_on_device) {
    if (!rl) MFAIL("null rloop");
    if (finish_on_device && (rl->p.l_anelastic_liquid || rl->p.l_single_matrix))
        MFAIL("magic_rloop_lm_options: the device epilogue covers finish_exp_entropy / _comp / _pol / _mag, not the anelastic-liquid or single-matrix variants");
    rl->lm_derivs = derivs_on_device ? 1 : 0;
    rl->lm_finish = finish_on_device ? 1 : 0;
    return 0;
}

extern "C" int magic_rloop_set_rotation(magic_rloop *rl, double omega_ma, double omega_ic) {
    if (!rl) MFAIL("null rloop");
    rl->p.omega_ma = omega_ma;
    rl->p.omega_ic = omega_ic;
    return 0;
}
extern "C" int magic_rloop_get_torques(const magic_rloop *rl, double *lorentz_torque_ic, double *lorentz_torque_ma) {
    if (!rl) MFAIL("null rloop");
    if (lorentz_torque_ic) *lorentz_torque_ic = rl->torque[0];
    if (lorentz_torque_ma) *lorentz_torque_ma = rl->torque[1];
    return 0;
}
extern "C" int magic_rloop_get_br_v_bcs(const magic_rloop *rl, int boundary, double *br_vt_lm, double *br_vp_lm) {
    if (!rl) MFAIL("null rloop");
    if (boundary < 0 || boundary > 1) MFAIL("magic_rloop_get_br_v_bcs: boundary must be 0 (CMB) or 1 (ICB)");
    if (rl->bc_lev[boundary] < 0) MFAIL("magic_rloop_get_br_v_bcs: this loop has no nonlinear magnetic boundary condition there");
    const size_t n = 2 * (size_t)rl->h->lm_max;
    if (br_vt_lm) memcpy(br_vt_lm, rl->h_brv + (size_t)boundary * 2 * n, sizeof(double) * n);
    if (br_vp_lm) memcpy(br_vp_lm, rl->h_brv + (size_t)boundary * 2 * n + n, sizeof(double) * n);
    return 0;
}
extern "C" int magic_rloop_sync(magic_rloop *rl) {
    if (!rl) MFAIL("null rloop");
    MCHECK(cudaSetDevice(rl->h->dev));
    MCHECK(cudaStreamSynchronize(rl->h->stream));
    return 0;
}
extern "C" long long magic_rloop_launch_count(const magic_rloop *rl) { return rl ? rl->h->launches : 0; }
extern "C" int magic_rloop_level_chunk(const magic_rloop *rl) { return rl ? rl->level_chunk : 0; }
extern "C" int magic_rloop_last_timing(const magic_rloop *rl, double out[8]) {
    if (!rl) MFAIL("null rloop");
    for (int i = 0; i < 8; i++) out[i] = rl->timing[i];
    return 0;
}
extern "C" int magic_rloop_last_exposed(const magic_rloop *rl, double out[2]) {
    if (!rl || !out) MFAIL("null argument");
    out[0] = rl->exposed[0];
    out[1] = rl->exposed[1];
    return 0;
}
extern "C" double magic_rloop_legendre_flops(const magic_rloop *rl) { return rl ? rl->legendre_flops : 0.0; }
extern "C" int magic_rloop_legendre_units(const magic_rloop *rl, double out[2]) {
    if (!rl || !out) MFAIL("magic_rloop_legendre_units: null argument");
    out[0] = rl->units_ref;
    out[1] = rl->units_exec;
    return 0;
}
