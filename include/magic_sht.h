/*
 * magic_sht.h -- C ABI of magic_b200, the B200-native backend for MagIC's radial-loop hot path.
 *
 * This is the drop-in boundary.  A Fortran `module sht` shim binds these entry points
 * with iso_c_binding exactly the way src/shtns.f90 binds SHTns' C API (shtns.f90:49-98,112-475): the
 * shim passes the first element of contiguous actuals, the backend owns only its handle, tables and
 * device buffers.  All `file:line` citations are relative to /root/reference/src/.
 *
 * Conventions
 *   - complex spectra are `double[2*lm_max]` (re,im interleaved = Fortran complex(cp)), st_map order
 *     (blocking.f90:309-317: do m=0,m_max,minc; do l=m,l_max).
 *   - grid fields are `double[nlat_padded * n_phi_max]`, theta fastest (Fortran f(nlat_padded,n_phi_max)),
 *     theta rows N/S interleaved like the native backend (l_scrambled_theta=.true., sht_native.f90:29).
 *   - every function returns 0 on success, non-zero on failure (the shim calls abortRun, useful.f90:271);
 *     magic_last_error() gives the message.  There is no CPU fallback: without a CUDA device every
 *     compute entry point fails.
 *   - pointers are HOST pointers unless the function name ends in _dev.
 */
#ifndef MAGIC_SHT_H
#define MAGIC_SHT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct magic_sht magic_sht; /* opaque handle: one per MPI rank <-> one GPU (shtns.f90:28 `sht_l`) */

const char *magic_last_error(void);
int magic_device_count(void);

/* ---------------------------------------------------------------------------------------------- */
/* initialize_sht / finalize_sht  (sht_native.f90:24-38, shtns.f90:32-100)                          */
/* l_max,m_max,minc,n_theta_max,n_phi_max: truncation.f90:55-105.  nlat_padded >= n_theta_max is the  */
/* leading dimension of the caller's grid arrays (truncation::nlat_padded, shtns.f90:65-72).          */
/* *l_scrambled_theta is set to 1 (N/S interleaved rows).                                             */
int magic_sht_create(int l_max, int m_max, int minc, int n_theta_max, int n_phi_max, int nlat_padded,
                     int device_id, int *l_scrambled_theta, magic_sht **out);
int magic_sht_destroy(magic_sht *h);

/* The CUDA stream (cudaStream_t) every kernel, copy and NCCL call of this handle is issued on, so a host can
 * order its own work or record events against it. */
void *magic_sht_stream(const magic_sht *h);
/* Number of kernels launched on behalf of this handle so far. */
long long magic_sht_launch_count(const magic_sht *h);

/* Gauss-Legendre colatitudes (north->south, radians) and weights as horizontal.f90:279-340 computes
 * them; lets the host check that both sides use the same grid. */
int magic_sht_get_grid(const magic_sht *h, double *theta_ord, double *gauss);

/* ---------------------------------------------------------------------------------------------- */
/* The 17 public procedures of `module sht` (sht_native.f90:16-20 == shtns.f90:22-26).               */
/* Same argument order and meaning as the Fortran subroutines; lcut = l_R(nR).                        */
int magic_scal_to_spat(magic_sht *h, const double *Slm, double *fieldc, int lcut);              /* :40  */
int magic_scal_to_grad_spat(magic_sht *h, const double *Slm, double *gradtc, double *gradpc, int lcut); /* :54 */
int magic_pol_to_grad_spat(magic_sht *h, const double *Slm, double *gradtc, double *gradpc, int lcut);  /* :70 */
int magic_torpol_to_spat(magic_sht *h, const double *Wlm, const double *dWlm, const double *Zlm, double *vrc,
                         double *vtc, double *vpc, int lcut);                                    /* :99  */
int magic_sphtor_to_spat(magic_sht *h, const double *dWlm, const double *Zlm, double *vtc, double *vpc,
                         int lcut);                                                               /* :129 */
int magic_torpol_to_curl_spat_IC(magic_sht *h, double r, double r_ICB, const double *dBlm, const double *ddBlm,
                                 const double *Jlm, const double *dJlm, double *cbr, double *cbt, double *cbp); /* :143 */
int magic_torpol_to_spat_IC(magic_sht *h, double r, double r_ICB, const double *Wlm, const double *dWlm,
                            const double *Zlm, double *Br, double *Bt, double *Bp);             /* :188 */
int magic_torpol_to_dphspat(magic_sht *h, const double *dWlm, const double *Zlm, double *dvtdp, double *dvpdp,
                            int lcut);                                                            /* :231 */
int magic_pol_to_curlr_spat(magic_sht *h, const double *Qlm, double *cvrc, int lcut);           /* :274 */
int magic_torpol_to_curl_spat(magic_sht *h, double or2, const double *Blm, const double *ddBlm, const double *Jlm,
                              const double *dJlm, double *cvrc, double *cvtc, double *cvpc, int lcut); /* :302 */
int magic_scal_to_SH(magic_sht *h, const double *f, double *fLM, int lcut);                     /* :338 */
int magic_spat_to_qst(magic_sht *h, const double *f, const double *g, const double *hh, double *qLM, double *sLM,
                      double *tLM, int lcut);                                                     /* :348 */
int magic_spat_to_sphertor(magic_sht *h, const double *f, const double *g, double *fLM, double *gLM, int lcut); /* :366 */
int magic_axi_to_spat(magic_sht *h, const double *fl_ax, double *f);                            /* :381 */
int magic_toraxi_to_spat(magic_sht *h, const double *fl_ax, double *ft, double *fp, int lcut);  /* :393 */

/* ---------------------------------------------------------------------------------------------- */
/* Batched radial loop: replaces `do nR=nRstart,nRstop` of rIter_single_t%radialLoop                 */
/* (rIter.f90:190-444) behind a new rIter_cuda_t extends(rIter_t) (rIteration.f90:15-19).            */

/* logic.f90 / physical_parameters.f90 / num_param.f90 values the loop reads (get_nl.f90:24-34,      */
/* get_td.f90:11-21, courant.f90, rIter.f90:16-40).  Integers are Fortran logicals as 0/1.            */
typedef struct {
    int l_conv, l_mag, l_heat, l_conv_nl, l_heat_nl, l_mag_nl, l_mag_LF, l_mag_kin, l_anel, l_adv_curl, l_corr,
        l_double_curl, l_single_matrix, l_chemical_conv, l_precession, l_centrifuge, l_anelastic_liquid,
        l_cour_alf_damp, l_full_sphere, l_parallel_solve, l_temperature_diff;
    int ktopv, kbotv;
    int l_cond_ma, l_cond_ic, l_rot_ma, l_rot_ic;
    int n_r_max, n_r_LCR;
    double LFfac, CorFac, epsc, epscXi, opm, ViscHeatFac, OhmLossFac;
    double oek, po, prec_angle, dilution_fac, ra, opr;
    double omega_ma, omega_ic, r_cmb, r_icb;
    double courfac, alffac;
    /* phase field (get_nl.f90:333-344; physical_parameters.f90 epsPhase, phaseDiffFac, penaltyFac, tmelt); appended in round 2 */
    double epsPhase, phaseDiffFac, penaltyFac, tmelt;
    int l_phase_field;
} magic_params;

/* radial_functions the loop reads, one entry per LOCAL level (radial.f90:283-307, num_param.f90:30-31). */
typedef struct {
    const int *nR;  /* global 1-based level index (n_r_cmb=1 ... n_r_icb=n_r_max) */
    const int *l_R; /* lcut per level */
    const double *r, *or1, *or2, *or4, *orho1, *orho2, *beta, *rho0, *otemp1, *temp0, *visc, *lambda, *epscProf,
        *delxr2, *delxh2;
} magic_radial;

/* R-distributed inputs X_Rloc(1:lm_max, nRstart:nRstop) = [n_r_loc][lm_max] complex (fields.f90:211-268).
 * NULL where the switch is off. */
typedef struct {
    const double *w, *dw, *ddw, *z, *dz, *s, *ds, *p, *xi, *b, *db, *ddb, *aj, *dj;
    const double *phi; /* phi_Rloc (l_phase_field) */
} magic_fields_in;

/* Outputs of radialLoop (rIter.f90:125-147), same layout; dtrkc/dthkc are [n_r_loc] reals. */
typedef struct {
    double *dwdt, *dzdt, *dpdt, *dsdt, *dxidt, *dbdt, *djdt, *dVxVhLM, *dVxBhLM, *dVSrLM, *dVXirLM;
    double *dtrkc, *dthkc;
    double *dphidt;    /* l_phase_field: scal_to_SH(phiTerms), rIter.f90:698 */
} magic_fields_out;

typedef struct magic_rloop magic_rloop;

/* Plan for n_r_loc local levels; level_chunk = number of levels batched per kernel wave (0 = auto from
 * free HBM).  Copies params/radial to the device. */
int magic_rloop_create(magic_sht *h, const magic_params *p, const magic_radial *rad, int n_r_loc, int level_chunk,
                       magic_rloop **out);
int magic_rloop_destroy(magic_rloop *rl);
/* Host only: the level chunks a loop over n_r_loc levels created with an explicit level_chunk works in (at most
 * n_r_loc entries): full chunks of level_chunk levels, a remainder <= level_chunk/4 folded into the last chunk, a larger one
 * as a short last chunk.  Every rank can compute every other rank's chunks from this (magic_rloop_run_lm_dev relies on it). */
int magic_level_chunks(int n_r_loc, int level_chunk, int *n_chunks, int *start, int *size);
/* Host-pointer call: H2D of the inputs, the loop, D2H of the outputs (what rIter_cuda_t calls). */
int magic_rloop_run(magic_rloop *rl, const magic_fields_in *in, const magic_fields_out *out, double time);
/* Device-pointer call (inputs/outputs already resident in HBM, e.g. produced by magic_transp_*_dev). */
int magic_rloop_run_dev(magic_rloop *rl, const magic_fields_in *in, const magic_fields_out *out, double time);
/* omega_ma / omega_ic change every time step when the mantle / inner core rotate (v_rigid_boundary,
 * nonlinear_bcs.f90:120-175): set them before a run.  After a run, the Lorentz torques of rIter.f90:279-292,461
 * (zero unless l_mag_LF and the wall is rotating and conducting). */
int magic_rloop_set_rotation(magic_rloop *rl, double omega_ma, double omega_ic);
int magic_rloop_get_torques(const magic_rloop *rl, double *lorentz_torque_ic, double *lorentz_torque_ma);
/* Nonlinear magnetic boundary products of get_br_v_bcs (nonlinear_bcs.f90:24-74, called at rIter.f90:267-277): the arguments
 * br_vt_lm_cmb, br_vp_lm_cmb (boundary = 0) / br_vt_lm_icb, br_vp_lm_icb (boundary = 1) of radialLoopG (radialLoop.f90:41),
 * complex [lm_max] HOST arrays, valid after a run on the rank that holds the boundary level.  Only defined for runs with
 * l_b_nl_cmb / l_b_nl_icb (stress-free wall + conducting mantle / inner core, Namelists.f90:713-729); an error otherwise. */
int magic_rloop_get_br_v_bcs(const magic_rloop *rl, int boundary, double *br_vt_lm, double *br_vp_lm);

/* ---- In-loop diagnostics of log steps (rIter.f90:303-373) ------------------------------------------------
 * The reference calls get_helicity (outMisc.f90:1052), get_hemi (:991), get_visc_heat (power.f90:384), get_perpPar
 * (outPar.f90:646), get_fluxes (:470) and get_nlBLayers (:584) on the grid arrays of every level when lHelCalc, lHemiCalc,
 * lPowerCalc, lPerpParCalc, lFluxProfCalc, lViscBcCalc are set (and get_ekin_solid_liquid, outMisc.f90:1169, with lPhaseCalc).
 * One call here returns all requested per-level sums: the fields they read are synthesised on the device (lDeriv = .true. on the boundary levels, as rIter.f90:193-205 sets it when
 * an output flag is on) and reduced by one fused kernel; out is a HOST array [n_r_loc][MAGIC_NDIAG].
 *   mask   MAGIC_DIAG_* bits; MAGIC_DIAG_RMSBULK = lRmsCalc is on, so boundary levels are treated as bulk (rIter.f90:215)
 *   ktops / kbots  thermal boundary types (1 = fixed entropy: horizontal entropy gradient zeroed there, rIter.f90:488-495)
 * Slot s of level i is out[i * MAGIC_NDIAG + s]:
 *   HelASr(nR,1:2) 0,1   Hel2ASr 2,3   HelnaASr 4,5   Helna2ASr 6,7   HelEAASr 8
 *   hemi_ekin_r(nR,1:2) 9,10   hemi_vrabs_r 11,12   hemi_emag_r 13,14   hemi_brabs_r 15,16        viscASr 17
 *   EperpASr 18   EparASr 19   EperpaxiASr 20   EparaxiASr 21
 *   fkinASr 22   sum(vr*sr) 23   sum(vr*pr) 24   fviscASr 25   fresASr 26   fpoynASr 27
 *       (fconvASr = temp0*[23] + ViscHeatFac*ThExpNb*alpha0*temp0*orho1*[24], or [23] alone with l_anelastic_liquid:
 *        outPar.f90:511-517 -- radial functions the host holds)
 *   uhASr 28   duhASr 29   gradT2ASr 30
 *   MAGIC_DIAG_PHASE (get_ekin_solid_liquid, outMisc.f90:1169-1221; needs l_phase_field and in->phi):
 *   ekinSr 32   ekinLr 33   volSr 34   min(phi) 35   max(phi) 36 over the level's grid (phase_min / phase_max of phase.TAG are
 *   the extrema of the last two over the levels, outMisc.f90:952-953)
 * magic_rloop_diagnostics takes HOST field pointers (the *_Rloc arrays rIter_cuda_t already holds), the _dev form device
 * pointers.  Not available for full-sphere runs. */
#define MAGIC_DIAG_HEL 1
#define MAGIC_DIAG_HEMI 2
#define MAGIC_DIAG_POWER 4
#define MAGIC_DIAG_PERPPAR 8
#define MAGIC_DIAG_FLUX 16
#define MAGIC_DIAG_VISCBC 32
#define MAGIC_DIAG_PHASE 64
#define MAGIC_DIAG_RMSBULK 256
#define MAGIC_NDIAG 40
int magic_rloop_diagnostics(magic_rloop *rl, const magic_fields_in *in, int mask, int ktops, int kbots, double *out);
int magic_rloop_diagnostics_dev(magic_rloop *rl, const magic_fields_in *in, int mask, int ktops, int kbots, double *out);
/* Torsional-oscillation sums (rIter.f90:395-404).  On lTONext steps getTOnext (TO.f90:309-352) keeps the
 * cylindrical field components Bs, Bp, Bz of every level on the grid; on lTOCalc steps getTO (TO.f90:141-307) forms the azimuthal
 * means of twenty products of (vr, vt, vp, cvr, dvpdr, br, bt, bp, cbr, cbt, phi) per level and colatitude.  magic_rloop_to_next
 * synthesises br, bt, bp of all local levels and keeps the three components on the DEVICE; magic_rloop_to returns
 * out[n_r_loc][MAGIC_NTO][n_theta_max] (HOST), colatitudes in geographic order north -> south (n_theta_cal2ord), arrays in the order
 *   V2AS 0  VAS 1  dzCorAS 2  dzRstrAS 3  dzAstrAS 4  dzLFAS 5  Bs2AS 6  BspAS 7  BpzAS 8  BszAS 9  BspdAS 10  BpsdAS 11
 *   BzpdAS 12  BpzdAS 13  dzPenAS 14          (the *_Rloc arrays of torsional_oscillations; magnetic ones 0 without l_mag)
 * dtLast is getTO's argument (the previous time step).  Boundary levels as in the diagnostics (lDeriv = .true., rigid-wall
 * values of v_rigid_boundary).  The O(l_max) spectral part of getTOnext / getTOfinish stays with the host (its get_PAS transforms
 * are magic_toraxi_to_spat).  Host field pointers; the _dev forms take device pointers.  Not available for full-sphere runs. */
#define MAGIC_NTO 15
int magic_rloop_to_next(magic_rloop *rl, const magic_fields_in *in);
int magic_rloop_to_next_dev(magic_rloop *rl, const magic_fields_in *in);
int magic_rloop_to(magic_rloop *rl, const magic_fields_in *in, double dtLast, double *out);
int magic_rloop_to_dev(magic_rloop *rl, const magic_fields_in *in, double dtLast, double *out);
/* R.m.s. force balance, the part inside the radial loop (l_RMS; rIter.f90:215-252, 433-435, 710; RMS.f90:469-610).
 * On lRmsCalc steps the reference treats every level as bulk, synthesises grad p and all velocity gradients, forms fourteen more
 * grid products per level (get_nl_RMS) and analyses them (transform_to_lm_RMS).  magic_rloop_rms does that for all local levels as
 * one more batch on the transform kernels: out is a HOST complex array [MAGIC_NRMS][n_r_loc][lm_max] =
 *   AdvrLM 0 (the merged radial advection compute_lm_forces receives, rIter.f90:650-688)   LFrLM 1   dtVrLM 2   dpkindrLM 3
 *   Advt2LM 4   Advp2LM 5   LFt2LM 6   LFp2LM 7   CFt2LM 8   CFp2LM 9   PFt2LM 10   PFp2LM 11   dtVtLM 12   dtVpLM 13
 * (module variables of RMS.f90; the spectral sums of compute_lm_forces, RMS.f90:612-863, stay with the host).  dt = tscheme%dt(1).
 * get_nl_RMS keeps the previous step's velocity on the grid (vr_old, vt_old, vp_old, RMS.f90:545-551, updated at every stage-1
 * call while l_RMS is on); here magic_rloop_rms_keep keeps its potentials w, dw, z on the DEVICE instead -- call it on every
 * stage-1 step, after magic_rloop_rms on lRmsCalc steps.  in->p is needed (transform_to_grid_RMS), in->s with l_centrifuge, in->phi
 * with l_phase_field; the precession terms take the `time` of the last pass of the loop (call magic_rloop_rms after magic_rloop_run
 * of the same stage, as rIter.f90 does).  Host field pointers; the _dev forms take device input pointers (out stays a host array).
 * Not for full-sphere runs. */
#define MAGIC_NRMS 14
int magic_rloop_rms_keep(magic_rloop *rl, const magic_fields_in *in);
int magic_rloop_rms_keep_dev(magic_rloop *rl, const magic_fields_in *in);
int magic_rloop_rms(magic_rloop *rl, const magic_fields_in *in, double dt, double *out);
int magic_rloop_rms_dev(magic_rloop *rl, const magic_fields_in *in, double dt, double *out);
/* get_dtBLM (rIter.f90:392-395, dtB.f90:144-223), what the loop contributes when l_dtB is on: the eleven grid
 * products of (vr, vt, vp, br, bt, bp) and their analyses (2 spat_to_sphertor + 7 scal_to_SH with lcut = l_max) for all local
 * levels, as one more batch on the Legendre GEMM / FFT kernels.  out: HOST complex [11][n_r_loc][lm_max] = BtVrLM, BpVrLM,
 * BrVtLM, BrVpLM, BtVpLM, BpVtLM, BpVtBtVpCotLM, BpVtBtVpSn2LM, BrVZLM, BtVZLM, BtVZsn2LM (the module arrays of dtB.f90:52-56,
 * which get_dH_dtBLM then combines level by level).  in: HOST pointers (w, dw, z, b, db, aj are read); _dev: device pointers. */
int magic_rloop_dtb(magic_rloop *rl, const magic_fields_in *in, double *out);
int magic_rloop_dtb_dev(magic_rloop *rl, const magic_fields_in *in, double *out);
/* The grid fields graphOut_mpi writes (rIter.f90:303-314, out_graph_file.f90:337) for one local level (0-based), HOST arrays
 * f(nlat_padded, n_phi) in the layout of the per-call transforms; NULL outputs are skipped (vr/vt/vp and br/bt/bp as triples). */
int magic_rloop_graph_fields(magic_rloop *rl, const magic_fields_in *in, int level, double *vr, double *vt, double *vp, double *br,
                             double *bt, double *bp, double *sr, double *pr);
/* Page-locks a PERSISTENT host array of the caller (a field container that lives as long as the run) so that magic_rloop_run
 * can overlap its transfers with the compute; unpin before the array is freed.  The run calls never pin on their own: without
 * this call the transfers are staged through pageable memory (correct, slower).  Already page-locked memory is accepted. */
int magic_rloop_pin_host(magic_rloop *rl, const void *ptr, size_t bytes);
int magic_rloop_unpin_host(magic_rloop *rl, const void *ptr);
/* The level chunk in use (magic_rloop_create's level_chunk, or its automatic choice, clamped to n_r_loc). */
int magic_rloop_level_chunk(const magic_rloop *rl);
/* Stream control + kernel accounting for the benchmark. */
int magic_rloop_sync(magic_rloop *rl);
long long magic_rloop_launch_count(const magic_rloop *rl);
/* Device time (ms, CUDA events on the loop's stream) of the last run, split by stage:
 * out[0]=total, [1]=synthesis prep, [2]=Legendre synthesis, [3]=c2r FFT, [4]=get_nl, [5]=r2c FFT,
 * [6]=Legendre analysis, [7]=get_td epilogue. */
int magic_rloop_last_timing(const magic_rloop *rl, double out[8]);
/* Device time (ms) of the last run that lay outside the chunk loop: out[0] = start of the run -> first kernel of the first
 * chunk (the inbound transposes / transfers of the first chunk that nothing could hide), out[1] = last kernel of the last
 * chunk -> end of the run (outbound transposes / transfers of the last chunk, boundary products). */
int magic_rloop_last_exposed(const magic_rloop *rl, double out[2]);
/* Algorithmic FP64 flops of the Legendre stage of one run (U * 2*n_theta*lm_max per level). */
double magic_rloop_legendre_flops(const magic_rloop *rl);
/* Scalar-equivalent Legendre passes per bulk level: out[0] = the reference's count (native_qst_to_spat / native_spat_to_sph_tor
 * sum against Plm AND dPlm: 5 per q/s/t transform, 36 for the MHD set), out[1] = the passes this library
 * executes (dPlm is a 3-point combination of Plm, plms.f90:117-187, so a vector component is ONE pass against Plm: 3 per
 * q/s/t transform, 22 for the MHD set). */
int magic_rloop_legendre_units(const magic_rloop *rl, double out[2]);

/* ---------------------------------------------------------------------------------------------- */
/* r <-> LM redistribution: a 5th type_mpitransp (mpi_transpose.f90:18-54), alltoallv semantics of   */
/* type_mpiatoav (:307-359, :444-530) with the lo<->st permutation fused into pack/unpack.           */
typedef struct magic_transp magic_transp;

/* NCCL bootstrap: rank 0 calls magic_transp_unique_id and broadcasts the 128 bytes (e.g. MPI_Bcast). */
int magic_transp_unique_id(char id[128]);
/* n_r_max levels and lm_max modes over n_procs ranks with getBlocks (parallel.f90:75-92) and
 * lo_map snake ordering (blocking.f90:387-544); n_fields = container width (fields.f90:211-268). */
int magic_transp_create(magic_sht *h, const char id[128], int rank, int n_procs, int n_r_max, int n_fields,
                        magic_transp **out);
int magic_transp_destroy(magic_transp *t);
/* Local extents: llm, ulm (1-based inclusive, lo order), nRstart, nRstop. */
int magic_transp_extents(const magic_transp *t, int *llm, int *ulm, int *nRstart, int *nRstop);
/* A part of a transposer: for every rank q it moves only the levels [lev_off[q], lev_off[q] + lev_cnt[q]) of q's radial
 * slab (offsets relative to the slab; a count may be 0).  arr_LMloc / arr_Rloc keep their full-slab shapes, so the parts
 * of one level partition together do what the parent does, in smaller all-to-alls that can overlap the compute of other
 * level chunks (magic_rloop_run_lm_dev does exactly that).  Parts share the parent's buffers and communicator: use them on
 * one stream and in the same order on every rank; destroy them before the parent. */
int magic_transp_create_part(magic_transp *parent, const int *lev_off, const int *lev_cnt, magic_transp **out);
int magic_transp_info(const magic_transp *t, int *rank, int *n_procs, int *n_r_max, int *n_fields);
/* cudaStream_t on which this object's pack / exchange / unpack work is queued (NULL = the handle's stream). */
int magic_transp_set_stream(magic_transp *t, void *stream);
/* arr_LMloc(llm:ulm, 1:n_r_max, n_fields) -> arr_Rloc(1:lm_max, nRstart:nRstop, n_fields), device pointers. */
int magic_transp_lm2r_dev(magic_transp *t, const double *arr_LMloc, double *arr_Rloc);
int magic_transp_r2lm_dev(magic_transp *t, const double *arr_Rloc, double *arr_LMloc);
/* Same, for a container of n_fields <= the width given at creation (one object and one NCCL communicator can
 * serve all containers of communications.f90:66-69). */
int magic_transp_lm2r_dev_n(magic_transp *t, int n_fields, const double *arr_LMloc, double *arr_Rloc);
int magic_transp_r2lm_dev_n(magic_transp *t, int n_fields, const double *arr_Rloc, double *arr_LMloc);
/* Host-pointer variants (H2D + exchange + D2H). */
int magic_transp_lm2r(magic_transp *t, const double *arr_LMloc, double *arr_Rloc);
int magic_transp_r2lm(magic_transp *t, const double *arr_Rloc, double *arr_LMloc);
/* Pack/unpack halves only (no exchange), for single-process testing of the permutation kernels:
 * lm2r: pack_lm2r fills sendbuf (ordered by destination rank, mpi_transpose.f90:320-333);
 *       unpack_lm2r consumes recvbuf (ordered by source rank, :341-357). */
int magic_transp_pack_lm2r_dev(magic_transp *t, const double *arr_LMloc, double *sendbuf);
int magic_transp_unpack_lm2r_dev(magic_transp *t, const double *recvbuf, double *arr_Rloc);
int magic_transp_pack_r2lm_dev(magic_transp *t, const double *arr_Rloc, double *sendbuf);
int magic_transp_unpack_r2lm_dev(magic_transp *t, const double *recvbuf, double *arr_LMloc);
/* counts/displacements in complex elements, length n_procs each (create_comm_alltoallv :120-152). */
int magic_transp_counts(const magic_transp *t, int dir /*0 lm2r, 1 r2lm*/, long long *scounts, long long *sdisp,
                        long long *rcounts, long long *rdisp);

/* ---- the whole hot path of one time step in one call (step_time.f90:485-612): transp_LMloc_to_Rloc -> radialLoopG ->
 * transp_Rloc_to_LMloc, on LM-distributed containers in the reference's packing (fields.f90:211-268,
 * dt_fieldsLast.f90:125-214).  The work is pipelined level chunk by level chunk: while chunk c computes, the transposes of
 * chunk c+1 (in) and c-1 (out) run on a high-priority second stream and -- for the host-pointer call -- the PCIe transfers
 * of chunks c+2 (up) and c-1 (down) on two more.  With several ranks the first and the last chunk of every slab are short
 * (4 levels; MAGIC_LM_TAPER), because their transposes are the only ones that cannot hide.  Every rank derives every rank's
 * chunks from the same pure function of (slab, level_chunk): create all loops of a run with the same level_chunk (or 0).
 * Field sets: flow (+ heat) (+ magnetic field) (+ composition), pressure or double-curl formulation.
 * Containers of switched-off physics are NULL.  ds / dxi (second field of the s / xi containers) are not read. */
typedef struct {
    const double *flow;  /* complex [5][n_r_max][nlm_loc]: w, dw, ddw, z, dz */
    const double *s;     /* complex [2][n_r_max][nlm_loc]: s, ds */
    const double *field; /* complex [5][n_r_max][nlm_loc]: b, db, ddb, aj, dj (NULL without l_mag) */
    const double *xi;    /* complex [2][n_r_max][nlm_loc]: xi, dxi (NULL without l_chemical_conv) */
} magic_lm_in;
typedef struct {
    double *dflowdt;     /* complex [3 or 4][n_r_max][nlm_loc]: dwdt, dzdt, dpdt (, dVxVhLM with l_double_curl) */
    double *dsdt;        /* complex [2][n_r_max][nlm_loc]: dsdt, dVSrLM */
    double *dbdt;        /* complex [3][n_r_max][nlm_loc]: dbdt, djdt, dVxBhLM (NULL without l_mag) */
    double *dtrkc, *dthkc; /* [n_r_loc]: device for _dev, host otherwise */
    double *dxidt;       /* complex [2][n_r_max][nlm_loc]: dxidt, dVXirLM (NULL without l_chemical_conv) */
} magic_lm_out;
/* DEVICE containers (the LM-side solver lives on the GPU, or a benchmark). */
int magic_rloop_run_lm_dev(magic_rloop *rl, magic_transp *t, const magic_lm_in *in, const magic_lm_out *out, double time);
/* HOST containers: the drop-in call of a Fortran host whose LM loop stays on the CPU -- replaces the three calls
 * transp_LMloc_to_Rloc (step_time.f90:485), radialLoopG (:530) and transp_Rloc_to_LMloc (:612) without intermediate host
 * R-containers.  Page-lock the containers once with magic_rloop_pin_host for asynchronous transfers. */
int magic_rloop_run_lm(magic_rloop *rl, magic_transp *t, const magic_lm_in *in, const magic_lm_out *out, double time);

/* ---- LM-side prologue and epilogue of the host-container call.  In the LM distribution a rank holds ALL
 * radial levels of its modes, so radial derivatives are local: a small FP64 GEMM with the radial scheme's matrix.
 *   magic_rloop_set_radial_matrices: D1, D2 row-major [n_r_max][n_r_max], (D f)(r_i) = sum_j D[i][j] f(r_j) -- what get_dr /
 *     get_ddr (radial_derivatives.f90:714-912) compute on this grid INCLUDING the n_cheb_max truncation; the shim builds them
 *     once by applying get_dr / get_ddr to the unit vectors.
 *   magic_rloop_set_lm_radial: or2, orho1, dentropy0, l_R on all n_r_max levels (finish_exp_entropy, updateS.f90:543-601).
 *   magic_rloop_lm_options(derivs_on_device, finish_on_device):
 *     derivs_on_device: dw, ddw, dz (db, ddb, dj) are computed on the device from w, z (b, aj); only fields 0 and 3 of the flow
 *       and field containers cross PCIe (5 instead of 11 arrays for the MHD set).  The upload of w, z, b, aj then precedes all
 *       compute instead of being pipelined with it.
 *     finish_on_device: finish_explicit_assembly (LMLoop.f90:390-453: finish_exp_entropy, _comp, _pol, _mag) runs on the device
 *       after the outbound transposes; dVSrLM, dVxBhLM, dVxVhLM, dVXirLM stay on the device (6 instead of 8 arrays come down),
 *       and the host must then NOT call finish_explicit_assembly itself (step_time.f90:647). */
int magic_rloop_set_radial_matrices(magic_rloop *rl, int n_r_max, const double *D1, const double *D2);
int magic_rloop_set_lm_radial(magic_rloop *rl, int n_r_max, const double *or2, const double *orho1, const double *dentropy0,
                              const int *l_R);
int magic_rloop_lm_options(magic_rloop *rl, int derivs_on_device, int finish_on_device);

/* Pure host helpers (no CUDA device needed): the decomposition the transposer uses.
 * magic_get_blocks: getBlocks (parallel.f90:75-92), 1-based inclusive start/stop per rank.
 * magic_lo_map: lo2st[lm_lo] = 0-based st_map index of the lm_lo-th entry of lo_map (snake ordering when
 * n_procs <= l_max/2, blocking.f90:387-544, else l-major :339-385) and each rank's 1-based inclusive lm range. */
int magic_get_blocks(int n_points, int n_procs, int *start, int *stop);
int magic_lo_map(int l_max, int m_max, int minc, int n_procs, int *lo2st, int *lm_start, int *lm_stop);

/* Device memory helpers so a host language without a CUDA binding can hold device buffers. */
int magic_dev_malloc(magic_sht *h, size_t bytes, void **ptr);
int magic_dev_free(magic_sht *h, void *ptr);
int magic_dev_upload(magic_sht *h, void *dst_dev, const void *src_host, size_t bytes);
int magic_dev_download(magic_sht *h, void *dst_host, const void *src_dev, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
