// api_sht.cu -- C ABI: handle management and the 17 per-call procedures of `module sht`
// (sht_native.f90:16-20).  Each call runs the batched pipeline with one level and a one-column program.
#include "../../include/magic_sht.h"
#include "engine.cuh"

using namespace magic;

struct CallCtx {
    BatchSpec spec;
    Layout L;
    Buffers buf;
    double *d_src = nullptr;    // [6][lm_max] complex staging
    double *d_stage = nullptr;  // [3][n_phi*nlat_padded] grid staging in the reference layout
    LevelInfo *d_lev = nullptr;
    ScalCol scal_cur;
    VecPair vec_cur;
};

extern "C" const char *magic_last_error(void) { return g_last_error.c_str(); }

extern "C" int magic_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int magic_sht_create(int l_max, int m_max, int minc, int n_theta_max, int n_phi_max, int nlat_padded, int device_id,
                                int *l_scrambled_theta, magic_sht **out) {
    if (!out) MFAIL("magic_sht_create: out is null");
    *out = nullptr;
    if (l_max < 1 || minc < 1 || m_max < 0 || m_max > l_max || m_max % minc != 0) MFAIL("magic_sht_create: bad l_max/m_max/minc");
    if (n_theta_max < 4 || n_theta_max % 4 != 0) MFAIL("magic_sht_create: n_theta_max must be a multiple of 4 (truncation.f90:146)");
    if (n_phi_max < 8 || n_phi_max % 4 != 0) MFAIL("magic_sht_create: n_phi_max must be a multiple of 4 (truncation.f90:137)");
    if (nlat_padded < n_theta_max) MFAIL("magic_sht_create: nlat_padded < n_theta_max");
    if (m_max / minc + 1 > n_phi_max / 2) MFAIL("magic_sht_create: n_m_max exceeds n_phi_max/2");
    int ndev = magic_device_count();
    if (ndev <= 0) MFAIL("magic_sht_create: no CUDA device (magic_b200 has no CPU fallback)");
    if (device_id < 0 || device_id >= ndev) MFAIL("magic_sht_create: bad device_id");
    magic_sht *h = new magic_sht();
    h->l_max = l_max; h->m_max = m_max; h->minc = minc; h->n_theta = n_theta_max; h->n_phi = n_phi_max;
    h->nlat_padded = nlat_padded; h->n_m = m_max / minc + 1; h->nh = n_theta_max / 2; h->NHP = pad_up(h->nh, 16); h->dev = device_id;
    h->lm_max = 0;
    for (int m = 0; m <= m_max; m += minc) h->lm_max += l_max - m + 1;
    if (sht_init(h)) { sht_free(h); delete h; return 1; }
    if (l_scrambled_theta) *l_scrambled_theta = 1;
    *out = h;
    return 0;
}

static void call_free(magic_sht *h) {
    if (!h->call) return;
    CallCtx *c = (CallCtx *)h->call;
    layout_free(c->L);
    buffers_free(c->buf);
    cudaFree(c->d_src); cudaFree(c->d_stage); cudaFree(c->d_lev);
    delete c;
    h->call = nullptr;
}

extern "C" int magic_sht_destroy(magic_sht *h) {
    if (!h) return 0;
    cudaSetDevice(h->dev);
    call_free(h);
    sht_free(h);
    delete h;
    return 0;
}

extern "C" void *magic_sht_stream(const magic_sht *h) { return h ? (void *)h->stream : nullptr; }
extern "C" long long magic_sht_launch_count(const magic_sht *h) { return h ? h->launches : 0; }

extern "C" int magic_sht_get_grid(const magic_sht *h, double *theta_ord, double *gauss) {
    if (!h) MFAIL("null handle");
    for (int i = 0; i < h->n_theta; i++) { theta_ord[i] = h->theta_ord[i]; gauss[i] = h->gauss[i]; }
    return 0;
}

static int call_ctx(magic_sht *h, CallCtx **out) {
    MCHECK(cudaSetDevice(h->dev));
    if (h->call) { *out = (CallCtx *)h->call; return 0; }
    CallCtx *c = new CallCtx();
    h->call = (struct CallCtx *)c;
    ScalCol sc{}; sc.t[0] = Term{0, F_NONE}; sc.t[1] = Term{0, F_NONE}; sc.lmask = LM_ALL;
    VecPair vp{}; vp.S[0] = vp.S[1] = vp.T[0] = vp.T[1] = Term{0, F_NONE}; vp.lmask = LM_ALL;
    c->spec.scal = {sc};
    c->spec.vec = {vp};
    c->spec.field_s = {0};
    c->spec.field_v = {1, 2};
    c->spec.nfield_in = 3;
    c->spec.nfield_out = 3;
    c->spec.afield_s = {0};
    c->spec.afield_vt = {1};
    c->spec.afield_vp = {2};
    layout_sizes(h, c->spec, 1, c->L);
    if (buffers_alloc(h, c->spec, c->L, c->buf)) return 1;
    if (layout_bind(h, c->spec, c->L, c->buf)) return 1;
    MCHECK(cudaMalloc((void **)&c->d_src, sizeof(double) * 2 * 6 * h->lm_max));
    MCHECK(cudaMalloc((void **)&c->d_stage, sizeof(double) * 3 * (size_t)h->n_phi * h->nlat_padded));
    MCHECK(cudaMemset(c->d_stage, 0, sizeof(double) * 3 * (size_t)h->n_phi * h->nlat_padded));
    MCHECK(cudaMalloc((void **)&c->d_lev, sizeof(LevelInfo)));
    *out = c;
    return 0;
}

static Term T_(int src, int ft) { return Term{src, ft}; }
static const Term TNONE = {0, F_NONE};

// Generic synthesis call: up to 6 spectral sources, one scalar column, one (S,T) pair.
// outs[0] <- scalar column, outs[1], outs[2] <- theta / phi components; null pointers are skipped.
static int synth_call(magic_sht *h, int nsrc, const double *const *srcs, const Term q[2], const Term S[2], const Term T[2],
                      double or2, int lcut, double *const outs[3], bool scale_osin2, bool axis_only) {
    if (!h) MFAIL("null handle");
    CallCtx *c;
    if (call_ctx(h, &c)) return 1;
    if (lcut < 0 || lcut > h->l_max) MFAIL("lcut out of range");
    const size_t sb = sizeof(double) * 2 * h->lm_max;
    for (int i = 0; i < nsrc; i++) MCHECK(cudaMemcpyAsync(c->d_src + (size_t)i * 2 * h->lm_max, srcs[i], sb, cudaMemcpyHostToDevice, h->stream));
    ScalCol sc{}; sc.t[0] = q[0]; sc.t[1] = q[1]; sc.lmask = LM_ALL;
    VecPair vp{}; vp.S[0] = S[0]; vp.S[1] = S[1]; vp.T[0] = T[0]; vp.T[1] = T[1]; vp.lmask = LM_ALL;
    MCHECK(cudaMemcpyAsync(c->L.d_scal, &sc, sizeof(sc), cudaMemcpyHostToDevice, h->stream));
    MCHECK(cudaMemcpyAsync(c->L.d_vec, &vp, sizeof(vp), cudaMemcpyHostToDevice, h->stream));
    LevelInfo li{};
    li.nR = 2; li.lcut = lcut; li.nBc = 0; li.lDeriv = 1; li.nl_on = 1; li.or2 = or2;
    MCHECK(cudaMemcpyAsync(c->d_lev, &li, sizeof(li), cudaMemcpyHostToDevice, h->stream));
    const double *src[MAGIC_MAX_SRC];
    for (int i = 0; i < MAGIC_MAX_SRC; i++) src[i] = i < nsrc ? c->d_src + (size_t)i * 2 * h->lm_max : nullptr;
    if (run_synthesis(h, c->spec, c->L, c->buf, src, c->d_lev, nullptr)) return 1;
    const size_t plane = (size_t)2 * h->nh * h->n_phi, gsz = (size_t)h->n_phi * h->nlat_padded;
    dim3 blk(32, 8), grd((h->n_phi + 31) / 32, (h->nh + 31) / 32);
    for (int f = 0; f < 3; f++) {
        if (!outs[f]) continue;
        grid_export_kernel<<<grd, blk, 0, h->stream>>>(c->buf.gin + f * plane, c->d_stage + f * gsz, h->nh, h->n_phi, h->nlat_padded,
                                                       scale_osin2 ? h->d_osin2 : nullptr);
        h->launches++;
        if (axis_only)  // axi_to_spat / toraxi_to_spat return f(theta) only: the phi-independent first column
            MCHECK(cudaMemcpyAsync(outs[f], c->d_stage + f * gsz, sizeof(double) * h->n_theta, cudaMemcpyDeviceToHost, h->stream));
        else
            MCHECK(cudaMemcpyAsync(outs[f], c->d_stage + f * gsz, sizeof(double) * gsz, cudaMemcpyDeviceToHost, h->stream));
    }
    MCHECK(cudaGetLastError());
    MCHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

// Generic analysis call: ins[0] scalar-type grid, ins[1]/ins[2] theta/phi-type grids; outs likewise.
static int anal_call(magic_sht *h, const double *const ins[3], double *const outs[3], int lcut) {
    if (!h) MFAIL("null handle");
    CallCtx *c;
    if (call_ctx(h, &c)) return 1;
    if (lcut < 0 || lcut > h->l_max) MFAIL("lcut out of range");
    const size_t plane = (size_t)2 * h->nh * h->n_phi, gsz = (size_t)h->n_phi * h->nlat_padded;
    dim3 blk(32, 8), grd((h->n_phi + 31) / 32, (h->nh + 31) / 32);
    for (int f = 0; f < 3; f++) {
        if (ins[f]) {
            MCHECK(cudaMemcpyAsync(c->d_stage + f * gsz, ins[f], sizeof(double) * gsz, cudaMemcpyHostToDevice, h->stream));
            grid_import_kernel<<<grd, blk, 0, h->stream>>>(c->d_stage + f * gsz, c->buf.gout + f * plane, h->nh, h->n_phi, h->nlat_padded);
            h->launches++;
        } else {
            MCHECK(cudaMemsetAsync(c->buf.gout + f * plane, 0, sizeof(double) * plane, h->stream));
        }
    }
    LevelInfo li{};
    li.nR = 2; li.lcut = lcut; li.nBc = 0; li.lDeriv = 1; li.nl_on = 1;
    MCHECK(cudaMemcpyAsync(c->d_lev, &li, sizeof(li), cudaMemcpyHostToDevice, h->stream));
    if (run_analysis(h, c->spec, c->L, c->buf, c->d_lev, nullptr)) return 1;
    const size_t sb = sizeof(double) * 2 * h->lm_max;
    if (outs[0]) MCHECK(cudaMemcpyAsync(outs[0], c->buf.nl_s, sb, cudaMemcpyDeviceToHost, h->stream));
    if (outs[1]) MCHECK(cudaMemcpyAsync(outs[1], c->buf.nl_v, sb, cudaMemcpyDeviceToHost, h->stream));
    if (outs[2]) MCHECK(cudaMemcpyAsync(outs[2], c->buf.nl_v + 2 * (size_t)h->lm_max, sb, cudaMemcpyDeviceToHost, h->stream));
    MCHECK(cudaStreamSynchronize(h->stream));
    return 0;
}

int magic::br_v_bcs_dev(magic_sht *h, const double *b, const double *dw, const double *z, int lcut, double fac, double omega,
                        double *br_vt_lm, double *br_vp_lm) {
    CallCtx *c;
    if (call_ctx(h, &c)) return 1;
    // brc = r^2 B_r is the Q part of torpol_to_spat(b, db, aj) (rIter.f90:606); vtc, vpc the horizontal part of
    // torpol_to_spat(w, dw, z) (rIter.f90:555-559: on a stress-free level only vrc is overwritten)
    ScalCol sc{}; sc.t[0] = T_(0, F_DLH); sc.t[1] = TNONE; sc.lmask = LM_ALL;
    VecPair vp{}; vp.S[0] = T_(1, F_ONE); vp.S[1] = TNONE; vp.T[0] = T_(2, F_ONE); vp.T[1] = TNONE; vp.lmask = LM_ALL;
    MCHECK(cudaMemcpyAsync(c->L.d_scal, &sc, sizeof(sc), cudaMemcpyHostToDevice, h->stream));
    MCHECK(cudaMemcpyAsync(c->L.d_vec, &vp, sizeof(vp), cudaMemcpyHostToDevice, h->stream));
    LevelInfo li{};
    li.nR = 2; li.lcut = lcut; li.nBc = 0; li.lDeriv = 1; li.nl_on = 1;
    MCHECK(cudaMemcpyAsync(c->d_lev, &li, sizeof(li), cudaMemcpyHostToDevice, h->stream));
    const double *src[MAGIC_MAX_SRC];
    for (int i = 0; i < MAGIC_MAX_SRC; i++) src[i] = nullptr;
    src[0] = b; src[1] = dw; src[2] = z;
    if (run_synthesis(h, c->spec, c->L, c->buf, src, c->d_lev, nullptr)) return 1;
    br_v_product_kernel<<<296, 256, 0, h->stream>>>(c->buf.gin, c->buf.gout, h->nh, h->n_phi, h->d_sinth, fac, omega);
    h->launches++;
    MCHECK(cudaGetLastError());
    li.lcut = h->l_max;  // spat_to_sphertor(..., l_max), nonlinear_bcs.f90:72
    MCHECK(cudaMemcpyAsync(c->d_lev, &li, sizeof(li), cudaMemcpyHostToDevice, h->stream));
    if (run_analysis(h, c->spec, c->L, c->buf, c->d_lev, nullptr)) return 1;
    const size_t sb = sizeof(double) * 2 * h->lm_max;
    MCHECK(cudaMemcpyAsync(br_vt_lm, c->buf.nl_v, sb, cudaMemcpyDeviceToDevice, h->stream));
    MCHECK(cudaMemcpyAsync(br_vp_lm, c->buf.nl_v + 2 * (size_t)h->lm_max, sb, cudaMemcpyDeviceToDevice, h->stream));
    return 0;
}

extern "C" int magic_scal_to_spat(magic_sht *h, const double *Slm, double *fieldc, int lcut) {
    const double *srcs[] = {Slm};
    Term q[2] = {T_(0, F_ONE), TNONE}, z[2] = {TNONE, TNONE};
    double *outs[3] = {fieldc, nullptr, nullptr};
    return synth_call(h, 1, srcs, q, z, z, 0.0, lcut, outs, false, false);
}

extern "C" int magic_scal_to_grad_spat(magic_sht *h, const double *Slm, double *gradtc, double *gradpc, int lcut) {
    const double *srcs[] = {Slm};
    Term S[2] = {T_(0, F_ONE), TNONE}, z[2] = {TNONE, TNONE};
    double *outs[3] = {nullptr, gradtc, gradpc};
    return synth_call(h, 1, srcs, z, S, z, 0.0, lcut, outs, false, false);
}

extern "C" int magic_pol_to_grad_spat(magic_sht *h, const double *Slm, double *gradtc, double *gradpc, int lcut) {
    const double *srcs[] = {Slm};
    Term S[2] = {T_(0, F_DLH), TNONE}, z[2] = {TNONE, TNONE};
    double *outs[3] = {nullptr, gradtc, gradpc};
    return synth_call(h, 1, srcs, z, S, z, 0.0, lcut, outs, false, false);
}

extern "C" int magic_torpol_to_spat(magic_sht *h, const double *Wlm, const double *dWlm, const double *Zlm, double *vrc, double *vtc,
                                    double *vpc, int lcut) {
    const double *srcs[] = {Wlm, dWlm, Zlm};
    Term q[2] = {T_(0, F_DLH), TNONE}, S[2] = {T_(1, F_ONE), TNONE}, T[2] = {T_(2, F_ONE), TNONE};
    double *outs[3] = {vrc, vtc, vpc};
    return synth_call(h, 3, srcs, q, S, T, 0.0, lcut, outs, false, false);
}

extern "C" int magic_sphtor_to_spat(magic_sht *h, const double *dWlm, const double *Zlm, double *vtc, double *vpc, int lcut) {
    const double *srcs[] = {dWlm, Zlm};
    Term z[2] = {TNONE, TNONE}, S[2] = {T_(0, F_ONE), TNONE}, T[2] = {T_(1, F_ONE), TNONE};
    double *outs[3] = {nullptr, vtc, vpc};
    return synth_call(h, 2, srcs, z, S, T, 0.0, lcut, outs, false, false);
}

extern "C" int magic_torpol_to_dphspat(magic_sht *h, const double *dWlm, const double *Zlm, double *dvtdp, double *dvpdp, int lcut) {
    const double *srcs[] = {dWlm, Zlm};
    Term z[2] = {TNONE, TNONE}, S[2] = {T_(0, F_IM), TNONE}, T[2] = {T_(1, F_IM), TNONE};
    double *outs[3] = {nullptr, dvtdp, dvpdp};
    return synth_call(h, 2, srcs, z, S, T, 0.0, lcut, outs, true, false);
}

extern "C" int magic_pol_to_curlr_spat(magic_sht *h, const double *Qlm, double *cvrc, int lcut) {
    const double *srcs[] = {Qlm};
    Term q[2] = {T_(0, F_DLH), TNONE}, z[2] = {TNONE, TNONE};
    double *outs[3] = {cvrc, nullptr, nullptr};
    return synth_call(h, 1, srcs, q, z, z, 0.0, lcut, outs, false, false);
}

extern "C" int magic_torpol_to_curl_spat(magic_sht *h, double or2, const double *Blm, const double *ddBlm, const double *Jlm,
                                         const double *dJlm, double *cvrc, double *cvtc, double *cvpc, int lcut) {
    const double *srcs[] = {Blm, ddBlm, Jlm, dJlm};
    Term q[2] = {T_(2, F_DLH), TNONE}, S[2] = {T_(3, F_ONE), TNONE}, T[2] = {T_(0, F_OR2DLH), T_(1, F_NEG)};
    double *outs[3] = {cvrc, cvtc, cvpc};
    return synth_call(h, 4, srcs, q, S, T, or2, lcut, outs, false, false);
}

// Inner-core variants: the (r/r_ICB)^l weights are applied on the host exactly as sht_native.f90:143-229
// does before its native_qst_to_spat call (diagnostic path: a few calls per output step).
static int ic_call(magic_sht *h, double r, double r_ICB, const double *a0, const double *a1, const double *a2, const double *a3,
                   bool curl, double *o0, double *o1, double *o2) {
    if (!h) MFAIL("null handle");
    const int lm_max = h->lm_max, l_max = h->l_max;
    std::vector<double> Q(2 * (size_t)lm_max), S(2 * (size_t)lm_max), T(2 * (size_t)lm_max), rDep(l_max + 1), rDep2(l_max + 1);
    double rRatio = r / r_ICB;
    rDep[0] = rRatio;
    rDep2[0] = 1.0 / r_ICB;
    for (int l = 1; l <= l_max; l++) { rDep[l] = rDep[l - 1] * rRatio; rDep2[l] = rDep2[l - 1] * rRatio; }
    for (int lm = 0; lm < lm_max; lm++) {
        int l = h->lm2l[lm];
        double dLh = (double)(l * (l + 1));
        for (int c = 0; c < 2; c++) {
            size_t i = 2 * (size_t)lm + c;
            if (!curl) {  // torpol_to_spat_IC(W=a0, dW=a1, Z=a2), sht_native.f90:213-222
                Q[i] = rDep[l] * dLh * a0[i];
                S[i] = rDep2[l] * ((l + 1) * a0[i] + r * a1[i]);
                T[i] = rDep[l] * a2[i];
            } else {      // torpol_to_curl_spat_IC(dB=a0, ddB=a1, J=a2, dJ=a3), sht_native.f90:170-179
                Q[i] = rDep[l] * dLh * a2[i];
                S[i] = rDep2[l] * ((l + 1) * a2[i] + r * a3[i]);
                T[i] = -rDep2[l] * (2 * (l + 1) * a0[i] + r * a1[i]);
            }
        }
    }
    const double *srcs[] = {Q.data(), S.data(), T.data()};
    Term q[2] = {T_(0, F_ONE), TNONE}, Ss[2] = {T_(1, F_ONE), TNONE}, Ts[2] = {T_(2, F_ONE), TNONE};
    double *outs[3] = {o0, o1, o2};
    return synth_call(h, 3, srcs, q, Ss, Ts, 0.0, l_max, outs, false, false);
}

extern "C" int magic_torpol_to_spat_IC(magic_sht *h, double r, double r_ICB, const double *Wlm, const double *dWlm, const double *Zlm,
                                       double *Br, double *Bt, double *Bp) {
    return ic_call(h, r, r_ICB, Wlm, dWlm, Zlm, nullptr, false, Br, Bt, Bp);
}

extern "C" int magic_torpol_to_curl_spat_IC(magic_sht *h, double r, double r_ICB, const double *dBlm, const double *ddBlm,
                                            const double *Jlm, const double *dJlm, double *cbr, double *cbt, double *cbp) {
    return ic_call(h, r, r_ICB, dBlm, ddBlm, Jlm, dJlm, true, cbr, cbt, cbp);
}

extern "C" int magic_scal_to_SH(magic_sht *h, const double *f, double *fLM, int lcut) {
    const double *ins[3] = {f, nullptr, nullptr};
    double *outs[3] = {fLM, nullptr, nullptr};
    return anal_call(h, ins, outs, lcut);
}

extern "C" int magic_spat_to_qst(magic_sht *h, const double *f, const double *g, const double *hh, double *qLM, double *sLM, double *tLM,
                                 int lcut) {
    const double *ins[3] = {f, g, hh};
    double *outs[3] = {qLM, sLM, tLM};
    return anal_call(h, ins, outs, lcut);
}

extern "C" int magic_spat_to_sphertor(magic_sht *h, const double *f, const double *g, double *fLM, double *gLM, int lcut) {
    const double *ins[3] = {nullptr, f, g};
    double *outs[3] = {nullptr, fLM, gLM};
    return anal_call(h, ins, outs, lcut);
}

// axisymmetric helpers: fl_ax(l_max+1) are the m=0 coefficients (shtransforms.f90:374-491)
static int axi_expand(magic_sht *h, const double *fl_ax, std::vector<double> &full) {
    if (!h) MFAIL("null handle");
    full.assign(2 * (size_t)h->lm_max, 0.0);
    for (int l = 0; l <= h->l_max; l++) { full[2 * l] = fl_ax[2 * l]; full[2 * l + 1] = fl_ax[2 * l + 1]; }
    return 0;
}

extern "C" int magic_axi_to_spat(magic_sht *h, const double *fl_ax, double *f) {
    std::vector<double> full;
    if (axi_expand(h, fl_ax, full)) return 1;
    const double *srcs[] = {full.data()};
    Term q[2] = {T_(0, F_ONE), TNONE}, z[2] = {TNONE, TNONE};
    double *outs[3] = {f, nullptr, nullptr};
    return synth_call(h, 1, srcs, q, z, z, 0.0, h->l_max, outs, false, true);
}

extern "C" int magic_toraxi_to_spat(magic_sht *h, const double *fl_ax, double *ft, double *fp, int lcut) {
    std::vector<double> full;
    if (axi_expand(h, fl_ax, full)) return 1;
    const double *srcs[] = {full.data()};
    Term z[2] = {TNONE, TNONE}, T[2] = {T_(0, F_ONE), TNONE};
    double *outs[3] = {nullptr, ft, fp};
    return synth_call(h, 1, srcs, z, z, T, 0.0, lcut, outs, false, true);
}

extern "C" int magic_dev_malloc(magic_sht *h, size_t bytes, void **ptr) {
    if (!h) MFAIL("null handle");
    MCHECK(cudaSetDevice(h->dev));
    MCHECK(cudaMalloc(ptr, bytes));
    return 0;
}
extern "C" int magic_dev_free(magic_sht *h, void *ptr) {
    if (!h) MFAIL("null handle");
    MCHECK(cudaSetDevice(h->dev));
    MCHECK(cudaFree(ptr));
    return 0;
}
extern "C" int magic_dev_upload(magic_sht *h, void *dst, const void *src, size_t bytes) {
    if (!h) MFAIL("null handle");
    MCHECK(cudaSetDevice(h->dev));
    MCHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    MCHECK(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int magic_dev_download(magic_sht *h, void *dst, const void *src, size_t bytes) {
    if (!h) MFAIL("null handle");
    MCHECK(cudaSetDevice(h->dev));
    MCHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    MCHECK(cudaStreamSynchronize(h->stream));
    return 0;
}
