/*
 * magic_oracle.h -- CPU restatement of MagIC's radial-loop hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for magic_b200.  It restates, loop nest by loop nest, the reference's
 * native (non-SHTns) path:  src/truncation.f90, src/blocking.f90, src/horizontal.f90, src/plms.f90,
 * src/shtransforms.f90, src/sht_native.f90, src/get_nl.f90, src/get_td.f90, src/courant.f90,
 * src/nonlinear_bcs.f90, src/rIter.f90 and the alltoallv flavour of src/mpi_transpose.f90.
 * Every function cites the reference file:line it follows.
 *
 * PARITY STATUS: the reference cannot be compiled in this image (Fortran 2008 + MPI, no Fortran compiler) and holds no
 * transform-level known-answer tests.  The oracle is pinned
 *   (1) against a GOLDEN VECTOR OF THE REFERENCE for the vector synthesis: the kinetic energy of the reference's own
 *       checkpoint fixture samples/boussBenchSat/checkpoint_end.start, evaluated in grid space through
 *       orc_torpol_to_spat, reproduces the e_kin_pol / e_kin_tor / axisymmetric columns of
 *       samples/boussBenchSat/reference.out to 8e-10 (tests/test_reference_energy.py; 9 printed digits);
 *   (2) by analytic known answers and an independent scipy evaluation of the associated Legendre functions
 *       (tests/test_oracle_analytic.py), which also tie the analysis to the synthesis by round trips <= 1e-13.
 *   (3) against the GOLDEN VECTORS OF THE REFERENCE'S OWN END-TO-END TEST for the whole radial loop (syntheses, get_nl,
 *       analyses, get_td, boundary levels, courant): orc_radial_loop, driven by a numpy restatement of the LM-side
 *       host (oracle/lmloop.py), reproduces the first 100 rows of samples/dynamo_benchmark/reference.out (e_kin.TAG,
 *       8 columns) and referenceMag.out (e_mag_oc.TAG, 12 columns) at the autotest tolerance rtol 1e-8
 *       (tests/test_dynamo_benchmark.py; the same test runs the CUDA library through the C ABI under -m gpu).
 *   (4) likewise for the anelastic branch (u.grad u advection, viscous heating, stress-free levels): the first logged row
 *       (10 steps) of samples/hydro_bench_anel/reference.out, whose axisymmetric columns exist only through the quadratic
 *       terms (tests/test_hydro_bench_anel.py; the CUDA library reproduces all 30 rows under -m gpu).
 *   (5) for the double-curl form of get_dwdt, l_R(nR) < l_max and minc = 3 on a grid that ends at r = 0: all 11 rows (100
 *       steps from the shipped checkpoint) of samples/full_sphere/reference.out, finite-difference host restated in
 *       oracle/lmloop_fd.py (tests/test_full_sphere.py; the CUDA library reproduces the same rows under -m gpu);
 *   (6) for the precession branch of get_nl (PCr/PCt/PCp), the `time` argument and m_max < l_max: all 20 rows (200 steps)
 *       of samples/precession/reference.out (tests/test_precession.py);
 *   (7) for rotating conducting walls -- get_nl on the boundary levels (lMagNlBc), v_rigid_boundary with omega_ic, the
 *       Lorentz torque: all 1000 steps of samples/dynamo_benchmark_condICrotIC/reference.out and referenceMag.out, and for
 *       get_br_v_bcs (stress-free walls + conducting inner core) the 100 steps of its restarted stage
 *       (tests/test_condICrotIC.py);
 *   (8) for a magnetic field in an anelastic background (Lorentz force and induction with orho1, Ohmic heating with
 *       lambda(r), both kinds of boundary level under lMagNlBc): all 500 steps of samples/varCond/reference.out and
 *       referenceMag.out (tests/test_varCond.py);
 *   (9) for chemical convection (xi, dxidt, dVXirLM) and for a loop called three times per step on Runge-Kutta stage states:
 *       the 25 BPR353 steps of samples/doubleDiffusion and of samples/boussBenchSat (saturated dynamo, conducting rotating
 *       inner core) from their shipped checkpoints (tests/test_doubleDiffusion.py, tests/test_boussBenchSat.py);
 *   (10) for a radially varying viscosity in the viscous heating: the 250 steps of samples/varProps (tests/test_varProps.py).
 *   (11) for the in-loop diagnostics of log steps (orc_radial_diagnostics: get_helicity, get_hemi, get_visc_heat): helicity.TAG,
 *       hemi.TAG and the viscous column of power.TAG of samples/testOutputs (tests/test_testOutputs.py);
 *   (12) for the phase-field branch of get_nl and get_ekin_solid_liquid: e_kin.TAG and all fifteen columns of phase.TAG of the
 *       Chebyshev stage of samples/phase_field (tests/test_phase_field.py);
 *   (13) for the r.m.s. force-balance batch (orc_radial_RMS) and get_dtBLM (orc_radial_dtB): all sixteen columns of dtVrms.TAG and
 *       all eleven of dtBrms.TAG of samples/testRMSOutputs, host side restated in oracle/rms_host.py (tests/test_testRMSOutputs.py);
 *   (14) for getTO (orc_radial_TO): all seven columns of Tay.TAG of samples/testTOGeosOutputs through the cylindrical averaging of
 *       outTO restated in oracle/rms_host.py (tests/test_testTOGeosOutputs.py).
 * Every one of these tests has a -m gpu leg in which the CUDA library takes the oracle's place in the same time loop.
 * Still "parity unpinned" (no reference vectors reachable here, literal line-cited restatements only): the r = 0 level
 * itself (v_center_sphere: it feeds no output of the loop, tests/test_oracle_analytic.py), get_perpPar / get_fluxes /
 * get_nlBLayers (samples/testRadialOutputs lacks its checkpoint: closed forms only), the arrays of getTO and the eleventh product
 * of get_dtBLM that only movie frames read, and the inner-core (_IC) and axisymmetric syntheses (per-call diagnostics; the
 * toroidal one is exercised by (14)).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call
 * into this library.  The product path (magic_b200/) never links or imports it.
 *
 * Conventions: all indices 0-based in C (Fortran lm=1.. becomes lm=0..).  Grid arrays are
 * f[nphi][nlat] with theta fastest (Fortran f(nlat_padded,n_phi_max)); theta rows are N/S
 * interleaved exactly as the native backend (row 2k = k-th northern node, row 2k+1 its mirror).
 */
#ifndef MAGIC_ORACLE_H
#define MAGIC_ORACLE_H

#include <complex.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double _Complex orc_cplx;

typedef struct orc_ctx orc_ctx;

/* truncation.f90:50-124 -- derive grid sizes. Pass l_max_in>0 (n_phi_tot derived) or l_max_in==0 and
 * n_phi_tot_in>0 (l_max derived).  out[0..6] = l_max, m_max, n_theta_max, n_phi_max, n_m_max, lm_max,
 * n_phi_tot. */
void orc_grid_sizes(int l_max_in, int n_phi_tot_in, int minc, int nalias, int out[7]);

/* getBlocks parallel.f90:75-92 -- start[p], stop[p] are 1-based inclusive like the reference. */
void orc_get_blocks(int n_points, int n_procs, int *start, int *stop);

/* Context = initialize_truncation + initialize_blocking(st_map) + horizontal + initialize_transforms. */
orc_ctx *orc_create(int l_max, int m_max, int minc, int n_theta_max, int n_phi_max);
void orc_destroy(orc_ctx *c);
void orc_set_threads(orc_ctx *c, int nthreads); /* OpenMP threads used inside the transforms */

/* Introspection (pointers into the context; lifetime = context). */
int orc_lm_max(const orc_ctx *c);
int orc_n_m_max(const orc_ctx *c);
const int *orc_lm2l(const orc_ctx *c);
const int *orc_lm2m(const orc_ctx *c);
const int *orc_lm2lmS(const orc_ctx *c);
const int *orc_lm2lmA(const orc_ctx *c);
const double *orc_theta_ord(const orc_ctx *c); /* n_theta, monotone north->south (gauleg output) */
const double *orc_gauss(const orc_ctx *c);     /* n_theta, scrambled like horizontal.f90:180-188 */
const double *orc_plm(const orc_ctx *c);       /* [n_theta/2][lm_max] */
const double *orc_dplm(const orc_ctx *c);
const double *orc_theta_vec(const orc_ctx *c, int which); /* 0 sinTheta 1 cosTheta 2 O_sin_theta
                                                 3 O_sin_theta_E2 4 sinTheta_E2 5 cosn_theta_E2 */
const double *orc_lm_vec(const orc_ctx *c, int which); /* 0 dLh 1..8 dTheta1S,1A,2S,2A,3S,3A,4S,4A */

/* lo_map (blocking.f90:339-544): fills lo2st[lm_lo] = st index (0-based) and lm_balance start/stop
 * (1-based inclusive) for n_procs ranks; snake ordering when n_procs <= l_max/2, else l-major. */
void orc_lo_map(const orc_ctx *c, int n_procs, int *lo2st, int *lm_start, int *lm_stop);

/* fft.f90:164-252 semantics on (nlat, nphi) arrays, theta fastest. */
void orc_ifft_many(const orc_ctx *c, const orc_cplx *f /*[nphi/2+1][nlat]*/, double *g /*[nphi][nlat]*/);
void orc_fft_many(const orc_ctx *c, const double *g, orc_cplx *f);

/* shtransforms.f90 native_* (file:line at each definition in the .c) */
void orc_native_qst_to_spat(const orc_ctx *c, const orc_cplx *Q, const orc_cplx *S, const orc_cplx *T,
                            double *br, double *bt, double *bp, int lcut);
void orc_native_sphtor_to_spat(const orc_ctx *c, const orc_cplx *S, const orc_cplx *T, double *bt,
                               double *bp, int lcut);
void orc_native_sph_to_spat(const orc_ctx *c, const orc_cplx *S, double *sc, int lcut);
void orc_native_sph_to_grad_spat(const orc_ctx *c, const orc_cplx *S, double *gt, double *gp, int lcut);
void orc_native_spat_to_sph(const orc_ctx *c, const double *scal, orc_cplx *fLM, int lcut);
void orc_native_spat_to_sph_tor(const orc_ctx *c, const double *vt, const double *vp, orc_cplx *f1,
                                orc_cplx *f2, int lcut);
void orc_native_axi_to_spat(const orc_ctx *c, const orc_cplx *S, double *sc);
void orc_native_toraxi_to_spat(const orc_ctx *c, const orc_cplx *T, double *bt, double *bp, int lcut);

/* sht_native.f90 wrappers (the `module sht` public list, sht_native.f90:16-20) */
void orc_scal_to_spat(const orc_ctx *c, const orc_cplx *S, double *f, int lcut);
void orc_scal_to_grad_spat(const orc_ctx *c, const orc_cplx *S, double *gt, double *gp, int lcut);
void orc_pol_to_grad_spat(const orc_ctx *c, const orc_cplx *S, double *gt, double *gp, int lcut);
void orc_torpol_to_spat(const orc_ctx *c, const orc_cplx *W, const orc_cplx *dW, const orc_cplx *Z,
                        double *vr, double *vt, double *vp, int lcut);
void orc_sphtor_to_spat(const orc_ctx *c, const orc_cplx *dW, const orc_cplx *Z, double *vt, double *vp,
                        int lcut);
void orc_torpol_to_dphspat(const orc_ctx *c, const orc_cplx *dW, const orc_cplx *Z, double *dvtdp,
                           double *dvpdp, int lcut);
void orc_pol_to_curlr_spat(const orc_ctx *c, const orc_cplx *Q, double *cvr, int lcut);
void orc_torpol_to_curl_spat(const orc_ctx *c, double or2, const orc_cplx *B, const orc_cplx *ddB,
                             const orc_cplx *J, const orc_cplx *dJ, double *cvr, double *cvt, double *cvp,
                             int lcut);
void orc_scal_to_SH(const orc_ctx *c, const double *f, orc_cplx *fLM, int lcut);
void orc_spat_to_qst(const orc_ctx *c, const double *f, const double *g, const double *h, orc_cplx *q,
                     orc_cplx *s, orc_cplx *t, int lcut);
void orc_spat_to_sphertor(const orc_ctx *c, const double *f, const double *g, orc_cplx *fLM, orc_cplx *gLM,
                          int lcut);
void orc_torpol_to_spat_IC(const orc_ctx *c, double r, double r_ICB, const orc_cplx *W, const orc_cplx *dW,
                           const orc_cplx *Z, double *Br, double *Bt, double *Bp);
void orc_torpol_to_curl_spat_IC(const orc_ctx *c, double r, double r_ICB, const orc_cplx *dB,
                                const orc_cplx *ddB, const orc_cplx *J, const orc_cplx *dJ, double *cbr,
                                double *cbt, double *cbp);
void orc_axi_to_spat(const orc_ctx *c, const orc_cplx *fl_ax, double *f);
void orc_toraxi_to_spat(const orc_ctx *c, const orc_cplx *fl_ax, double *ft, double *fp, int lcut);

/* ---- radial loop (rIter.f90:94-712 + get_nl.f90 + get_td.f90 + courant.f90 + nonlinear_bcs.f90) ---- */

/* Run-wide switches and constants (logic.f90 / physical_parameters.f90 values the hot path reads). */
typedef struct {
    int l_conv, l_mag, l_heat, l_conv_nl, l_heat_nl, l_mag_nl, l_mag_LF, l_mag_kin, l_anel, l_adv_curl,
        l_corr, l_double_curl, l_single_matrix, l_chemical_conv, l_precession, l_centrifuge,
        l_anelastic_liquid, l_cour_alf_damp, l_full_sphere, l_parallel_solve, l_temperature_diff;
    int ktopv, kbotv;             /* 1 stress-free, 2 rigid */
    int l_cond_ma, l_cond_ic, l_rot_ma, l_rot_ic;
    int n_r_max, n_r_LCR;         /* radial levels are numbered 1..n_r_max; n_r_cmb=1, n_r_icb=n_r_max */
    double LFfac, CorFac, epsc, epscXi, opm, ViscHeatFac, OhmLossFac;
    double oek, po, prec_angle, dilution_fac, ra, opr;
    double omega_ma, omega_ic, r_cmb, r_icb;
    double courfac, alffac;
    double epsPhase, phaseDiffFac, penaltyFac, tmelt; /* phase field, get_nl.f90:333-344 */
    int l_phase_field;
} orc_params;

/* Radial functions for the n_r levels handed to the loop (all arrays length n_r). */
typedef struct {
    const int *nR;   /* global 1-based level number of each local level */
    const int *l_R;  /* lcut per level (radial.f90:291-307) */
    const double *r, *or1, *or2, *or4, *orho1, *orho2, *beta, *rho0, *otemp1, *temp0, *visc, *lambda,
        *epscProf, *delxr2, *delxh2;
} orc_radial;

/* R-distributed spectral inputs [n_r][lm_max] (Fortran X_Rloc(lm, nR)); any may be NULL if unused. */
typedef struct {
    const orc_cplx *w, *dw, *ddw, *z, *dz, *s, *ds, *p, *xi, *b, *db, *ddb, *aj, *dj;
    const orc_cplx *phi; /* phase field (l_phase_field) */
} orc_fields_in;

/* Outputs of radialLoop (rIter.f90:125-147), [n_r][lm_max]; may be NULL when the switch is off. */
typedef struct {
    orc_cplx *dwdt, *dzdt, *dpdt, *dsdt, *dxidt, *dbdt, *djdt, *dVxVhLM, *dVxBhLM, *dVSrLM, *dVXirLM;
    double *dtrkc, *dthkc; /* [n_r] */
    double *lorentz_torque_ic, *lorentz_torque_ma; /* scalars (rIter.f90:279-292, outRot.f90:423-483); may be NULL */
    /* get_br_v_bcs products [lm_max] (rIter.f90:267-277, nonlinear_bcs.f90:24-74): written when the loop holds the CMB / ICB
     * level and the run has l_b_nl_cmb / l_b_nl_icb (Namelists.f90:713-729); may be NULL */
    orc_cplx *br_vt_lm_cmb, *br_vp_lm_cmb, *br_vt_lm_icb, *br_vp_lm_icb;
    orc_cplx *dphidt; /* l_phase_field: scal_to_SH(phiTerms), rIter.f90:698 */
} orc_fields_out;

/* Executes the body of `do nR=nRstart,nRstop` (rIter.f90:190-444) for n_r levels, with all output
 * flags (lRmsCalc, lTOCalc, ...) false and lPressNext=false.  time is used by precession only. */
void orc_radial_loop(const orc_ctx *c, const orc_params *p, const orc_radial *rad, int n_r,
                     const orc_fields_in *in, const orc_fields_out *out, double time);

/* The grid-space diagnostics of rIter.f90:303-373 (get_helicity, get_hemi, get_visc_heat, get_perpPar, get_fluxes,
 * get_nlBLayers) for n_r levels: out[n_r][40], slots as MAGIC_DG_* of include/magic_sht.h; mask = MAGIC_DIAG_* bits;
 * ktops / kbots: thermal boundary condition types (fixed temperature = 1 zeroes the horizontal entropy gradient on that
 * boundary, rIter.f90:488-495). */
void orc_radial_diagnostics(const orc_ctx *c, const orc_params *p, const orc_radial *rad, int n_r, const orc_fields_in *in,
                            int mask, int ktops, int kbots, double *out);

/* get_dtBLM (dtB.f90:144-223) for n_r levels: out[11][n_r][lm_max] = BtVrLM, BpVrLM, BrVtLM, BrVpLM, BtVpLM, BpVtLM,
 * BpVtBtVpCotLM, BpVtBtVpSn2LM, BrVZLM, BtVZLM, BtVZsn2LM. */
void orc_radial_dtB(const orc_ctx *c, const orc_params *p, const orc_radial *rad, int n_r, const orc_fields_in *in, orc_cplx *out);

/* R.m.s. force balance inside the radial loop on lRmsCalc steps (rIter.f90:215-252, 710; RMS.f90:469-610): out[14][n_r][lm_max] =
 * AdvrLM, LFrLM, dtVrLM, dpkindrLM, Advt2LM, Advp2LM, LFt2LM, LFp2LM, CFt2LM, CFp2LM, PFt2LM, PFp2LM, dtVtLM, dtVpLM; w_old, dw_old,
 * z_old [n_r][lm_max]: the flow potentials of the previous stage-1 call (for vr_old, vt_old, vp_old); dt = tscheme%dt(1); time
 * enters the precession terms. */
void orc_radial_RMS(const orc_ctx *c, const orc_params *p, const orc_radial *rad, int n_r, const orc_fields_in *in, const orc_cplx *w_old,
                    const orc_cplx *dw_old, const orc_cplx *z_old, double dt, double time, orc_cplx *out);

/* Torsional-oscillation sums (rIter.f90:395-404): mode 0 = getTOnext's grid part (TO.f90:330-343; fills last[n_r][3][n_phi][n_theta]
 * = BsLast, BpLast, BzLast), mode 1 = getTO (TO.f90:141-307): out[n_r][15][n_theta], colatitudes unscrambled, arrays V2AS, VAS,
 * dzCorAS, dzRstrAS, dzAstrAS, dzLFAS, Bs2AS, BspAS, BpzAS, BszAS, BspdAS, BpsdAS, BzpdAS, BpzdAS, dzPenAS. */
void orc_radial_TO(const orc_ctx *c, const orc_params *p, const orc_radial *rad, int n_r, const orc_fields_in *in, int mode, double dtLast,
                   double *last, double *out);

/* get_nl only, on caller-provided grids (get_nl.f90:213-441): in/out arrays are [nphi][nlat] each.
 * in: vr vt vp cvr cvt cvp s br bt bp cbr cbt cbp (13) ; out: Advr Advt Advp LFr LFt LFp VSr VSt VSp
 * VxBr VxBt VxBp (12).  Provided so tests can check the grid-space kernel in isolation. */
void orc_get_nl_mhd(const orc_ctx *c, const orc_params *p, int nR, int nBc, double or2, double or4,
                    double orho1, const double *const in[13], double *const out[12]);

/* mpi_transpose.f90:307-359,444-530 (type_mpiatoav) emulated in one process for n_procs ranks:
 * lm2r: arr_LM[p] is rank p's (llm:ulm, n_r_max, n_fields) block -> arr_R[q] (lm_max, nR_loc(q), n_fields)
 * Layouts follow the Fortran column-major declarations. */
void orc_transp_lm2r(const orc_ctx *c, int n_procs, int n_r_max, int n_fields, const orc_cplx *const *arr_LM,
                     orc_cplx *const *arr_R);
void orc_transp_r2lm(const orc_ctx *c, int n_procs, int n_r_max, int n_fields, const orc_cplx *const *arr_R,
                     orc_cplx *const *arr_LM);

#ifdef __cplusplus
}
#endif
#endif
