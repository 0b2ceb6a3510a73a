module magic_b200_c
   !
   ! iso_c_binding view of include/magic_sht.h (libmagic_b200.so).  One interface per C entry point the three
   ! Fortran shims use:  sht_cuda.f90 (module sht),  rIter_cuda.f90 (rIter_cuda_t),  mpi_transp_cuda.f90 (type_mpicuda).
   ! Modelled on the way src/shtns.f90 includes SHTns' `shtns.f03` interface block (shtns.f90:11,28).
   !
   ! Conventions of the C side: every function returns 0 on success; complex(cp) spectra are passed as the address of
   ! their first element (re,im interleaved); Fortran logicals travel as integer(c_int) 0/1.
   !
   use iso_c_binding
   implicit none
   public

   !-- magic_params (include/magic_sht.h): the run-wide switches and numbers the radial loop reads
   type, bind(C) :: magic_params
      integer(c_int) :: l_conv, l_mag, l_heat, l_conv_nl, l_heat_nl, l_mag_nl, l_mag_LF, l_mag_kin, l_anel, &
      &                 l_adv_curl, l_corr, l_double_curl, l_single_matrix, l_chemical_conv, l_precession,  &
      &                 l_centrifuge, l_anelastic_liquid, l_cour_alf_damp, l_full_sphere, l_parallel_solve, &
      &                 l_temperature_diff
      integer(c_int) :: ktopv, kbotv
      integer(c_int) :: l_cond_ma, l_cond_ic, l_rot_ma, l_rot_ic
      integer(c_int) :: n_r_max, n_r_LCR
      real(c_double) :: LFfac, CorFac, epsc, epscXi, opm, ViscHeatFac, OhmLossFac
      real(c_double) :: oek, po, prec_angle, dilution_fac, ra, opr
      real(c_double) :: omega_ma, omega_ic, r_cmb, r_icb
      real(c_double) :: courfac, alffac
      real(c_double) :: epsPhase, phaseDiffFac, penaltyFac, tmelt
      integer(c_int) :: l_phase_field
   end type magic_params

   !-- magic_radial: radial functions, one entry per LOCAL level
   type, bind(C) :: magic_radial
      type(c_ptr) :: nR, l_R
      type(c_ptr) :: r, or1, or2, or4, orho1, orho2, beta, rho0, otemp1, temp0, visc, lambda, epscProf, &
      &              delxr2, delxh2
   end type magic_radial

   !-- magic_fields_in / magic_fields_out: R-distributed containers (lm_max, nRstart:nRstop); c_null_ptr where unused
   type, bind(C) :: magic_fields_in
      type(c_ptr) :: w, dw, ddw, z, dz, s, ds, p, xi, b, db, ddb, aj, dj
      type(c_ptr) :: phi
   end type magic_fields_in

   type, bind(C) :: magic_fields_out
      type(c_ptr) :: dwdt, dzdt, dpdt, dsdt, dxidt, dbdt, djdt, dVxVhLM, dVxBhLM, dVSrLM, dVXirLM
      type(c_ptr) :: dtrkc, dthkc
      type(c_ptr) :: dphidt
   end type magic_fields_out

   !-- LM-distributed containers of magic_rloop_run_lm (fields.f90:211-268, dt_fieldsLast.f90:125-214, module fieldsLast)
   type, bind(C) :: magic_lm_in
      type(c_ptr) :: flow, s, field, xi
   end type magic_lm_in

   type, bind(C) :: magic_lm_out
      type(c_ptr) :: dflowdt, dsdt, dbdt
      type(c_ptr) :: dtrkc, dthkc
      type(c_ptr) :: dxidt
   end type magic_lm_out

   !-- magic_rloop_diagnostics: mask bits (include/magic_sht.h)
   integer(c_int), parameter :: MAGIC_DIAG_HEL = 1, MAGIC_DIAG_HEMI = 2, MAGIC_DIAG_POWER = 4, MAGIC_DIAG_PERPPAR = 8, &
   &                            MAGIC_DIAG_FLUX = 16, MAGIC_DIAG_VISCBC = 32, MAGIC_DIAG_PHASE = 64, MAGIC_DIAG_RMSBULK = 256, &
   &                            MAGIC_NTO = 15, MAGIC_NRMS = 14

   interface

      !---------------------------------------------------------------- C library
      function c_strlen(s) bind(C, name='strlen') result(n)
         import :: c_ptr, c_size_t
         type(c_ptr), value :: s
         integer(c_size_t) :: n
      end function c_strlen

      !---------------------------------------------------------------- handle
      function magic_last_error() bind(C, name='magic_last_error') result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function magic_last_error

      function magic_device_count() bind(C, name='magic_device_count') result(n)
         import :: c_int
         integer(c_int) :: n
      end function magic_device_count

      function magic_sht_create(l_max, m_max, minc, n_theta_max, n_phi_max, nlat_padded, device_id, &
               &                l_scrambled_theta, h) bind(C, name='magic_sht_create') result(ierr)
         import :: c_int, c_ptr
         integer(c_int), value :: l_max, m_max, minc, n_theta_max, n_phi_max, nlat_padded, device_id
         integer(c_int), intent(out) :: l_scrambled_theta
         type(c_ptr),    intent(out) :: h
         integer(c_int) :: ierr
      end function magic_sht_create

      function magic_sht_destroy(h) bind(C, name='magic_sht_destroy') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: h
         integer(c_int) :: ierr
      end function magic_sht_destroy

      !---------------------------------------------------------------- the 17 procedures of module sht
      function magic_scal_to_spat(h, Slm, fieldc, lcut) bind(C, name='magic_scal_to_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: Slm(*)
         real(c_double), intent(out) :: fieldc(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_scal_to_spat

      function magic_scal_to_grad_spat(h, Slm, gradtc, gradpc, lcut) bind(C, name='magic_scal_to_grad_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: Slm(*)
         real(c_double), intent(out) :: gradtc(*), gradpc(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_scal_to_grad_spat

      function magic_pol_to_grad_spat(h, Slm, gradtc, gradpc, lcut) bind(C, name='magic_pol_to_grad_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: Slm(*)
         real(c_double), intent(out) :: gradtc(*), gradpc(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_pol_to_grad_spat

      function magic_torpol_to_spat(h, Wlm, dWlm, Zlm, vrc, vtc, vpc, lcut) bind(C, name='magic_torpol_to_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: Wlm(*), dWlm(*), Zlm(*)
         real(c_double), intent(out) :: vrc(*), vtc(*), vpc(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_torpol_to_spat

      function magic_sphtor_to_spat(h, dWlm, Zlm, vtc, vpc, lcut) bind(C, name='magic_sphtor_to_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: dWlm(*), Zlm(*)
         real(c_double), intent(out) :: vtc(*), vpc(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_sphtor_to_spat

      function magic_torpol_to_curl_spat_IC(h, r, r_ICB, dBlm, ddBlm, Jlm, dJlm, cbr, cbt, cbp) &
               &   bind(C, name='magic_torpol_to_curl_spat_IC') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         real(c_double), value :: r, r_ICB
         complex(c_double_complex), intent(in) :: dBlm(*), ddBlm(*), Jlm(*), dJlm(*)
         real(c_double), intent(out) :: cbr(*), cbt(*), cbp(*)
         integer(c_int) :: ierr
      end function magic_torpol_to_curl_spat_IC

      function magic_torpol_to_spat_IC(h, r, r_ICB, Wlm, dWlm, Zlm, Br, Bt, Bp) &
               &   bind(C, name='magic_torpol_to_spat_IC') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         real(c_double), value :: r, r_ICB
         complex(c_double_complex), intent(in) :: Wlm(*), dWlm(*), Zlm(*)
         real(c_double), intent(out) :: Br(*), Bt(*), Bp(*)
         integer(c_int) :: ierr
      end function magic_torpol_to_spat_IC

      function magic_torpol_to_dphspat(h, dWlm, Zlm, dvtdp, dvpdp, lcut) bind(C, name='magic_torpol_to_dphspat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: dWlm(*), Zlm(*)
         real(c_double), intent(out) :: dvtdp(*), dvpdp(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_torpol_to_dphspat

      function magic_pol_to_curlr_spat(h, Qlm, cvrc, lcut) bind(C, name='magic_pol_to_curlr_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: Qlm(*)
         real(c_double), intent(out) :: cvrc(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_pol_to_curlr_spat

      function magic_torpol_to_curl_spat(h, or2, Blm, ddBlm, Jlm, dJlm, cvrc, cvtc, cvpc, lcut) &
               &   bind(C, name='magic_torpol_to_curl_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         real(c_double), value :: or2
         complex(c_double_complex), intent(in) :: Blm(*), ddBlm(*), Jlm(*), dJlm(*)
         real(c_double), intent(out) :: cvrc(*), cvtc(*), cvpc(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_torpol_to_curl_spat

      function magic_scal_to_SH(h, f, fLM, lcut) bind(C, name='magic_scal_to_SH') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         real(c_double), intent(in) :: f(*)
         complex(c_double_complex), intent(out) :: fLM(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_scal_to_SH

      function magic_spat_to_qst(h, f, g, hh, qLM, sLM, tLM, lcut) bind(C, name='magic_spat_to_qst') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         real(c_double), intent(in) :: f(*), g(*), hh(*)
         complex(c_double_complex), intent(out) :: qLM(*), sLM(*), tLM(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_spat_to_qst

      function magic_spat_to_sphertor(h, f, g, fLM, gLM, lcut) bind(C, name='magic_spat_to_sphertor') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         real(c_double), intent(in) :: f(*), g(*)
         complex(c_double_complex), intent(out) :: fLM(*), gLM(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_spat_to_sphertor

      function magic_axi_to_spat(h, fl_ax, f) bind(C, name='magic_axi_to_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: fl_ax(*)
         real(c_double), intent(out) :: f(*)
         integer(c_int) :: ierr
      end function magic_axi_to_spat

      function magic_toraxi_to_spat(h, fl_ax, ft, fp, lcut) bind(C, name='magic_toraxi_to_spat') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex
         type(c_ptr), value :: h
         complex(c_double_complex), intent(in) :: fl_ax(*)
         real(c_double), intent(out) :: ft(*), fp(*)
         integer(c_int), value :: lcut
         integer(c_int) :: ierr
      end function magic_toraxi_to_spat

      !---------------------------------------------------------------- batched radial loop
      function magic_rloop_create(h, p, rad, n_r_loc, level_chunk, rl) bind(C, name='magic_rloop_create') result(ierr)
         import :: c_int, c_ptr, magic_params, magic_radial
         type(c_ptr), value :: h
         type(magic_params), intent(in) :: p
         type(magic_radial), intent(in) :: rad
         integer(c_int), value :: n_r_loc, level_chunk
         type(c_ptr), intent(out) :: rl
         integer(c_int) :: ierr
      end function magic_rloop_create

      function magic_rloop_destroy(rl) bind(C, name='magic_rloop_destroy') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: rl
         integer(c_int) :: ierr
      end function magic_rloop_destroy

      function magic_rloop_run(rl, fin, fout, time) bind(C, name='magic_rloop_run') result(ierr)
         import :: c_int, c_ptr, c_double, magic_fields_in, magic_fields_out
         type(c_ptr), value :: rl
         type(magic_fields_in),  intent(in) :: fin
         type(magic_fields_out), intent(in) :: fout
         real(c_double), value :: time
         integer(c_int) :: ierr
      end function magic_rloop_run

      function magic_rloop_run_lm(rl, t, lin, lout, time) bind(C, name='magic_rloop_run_lm') result(ierr)
         import :: c_int, c_ptr, c_double, magic_lm_in, magic_lm_out
         type(c_ptr), value :: rl, t
         type(magic_lm_in),  intent(in) :: lin
         type(magic_lm_out), intent(in) :: lout
         real(c_double), value :: time
         integer(c_int) :: ierr
      end function magic_rloop_run_lm

      function magic_rloop_pin_host(rl, ptr, bytes) bind(C, name='magic_rloop_pin_host') result(ierr)
         import :: c_int, c_ptr, c_size_t
         type(c_ptr), value :: rl, ptr
         integer(c_size_t), value :: bytes
         integer(c_int) :: ierr
      end function magic_rloop_pin_host

      function magic_rloop_unpin_host(rl, ptr) bind(C, name='magic_rloop_unpin_host') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: rl, ptr
         integer(c_int) :: ierr
      end function magic_rloop_unpin_host

      function magic_rloop_set_rotation(rl, omega_ma, omega_ic) bind(C, name='magic_rloop_set_rotation') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: rl
         real(c_double), value :: omega_ma, omega_ic
         integer(c_int) :: ierr
      end function magic_rloop_set_rotation

      function magic_rloop_get_torques(rl, lorentz_torque_ic, lorentz_torque_ma) &
               &   bind(C, name='magic_rloop_get_torques') result(ierr)
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: rl
         real(c_double), intent(out) :: lorentz_torque_ic, lorentz_torque_ma
         integer(c_int) :: ierr
      end function magic_rloop_get_torques

      function magic_rloop_get_br_v_bcs(rl, boundary, br_vt_lm, br_vp_lm) bind(C, name='magic_rloop_get_br_v_bcs') result(ierr)
         import :: c_int, c_ptr, c_double_complex
         type(c_ptr), value :: rl
         integer(c_int), value :: boundary       ! 0 = CMB, 1 = ICB
         complex(c_double_complex), intent(out) :: br_vt_lm(*), br_vp_lm(*)
         integer(c_int) :: ierr
      end function magic_rloop_get_br_v_bcs

      !-- in-loop diagnostics of log steps (rIter.f90:303-373): out(MAGIC_NDIAG, n_r_loc), slots as in include/magic_sht.h
      function magic_rloop_diagnostics(rl, fin, mask, ktops, kbots, out) bind(C, name='magic_rloop_diagnostics') result(ierr)
         import :: c_int, c_ptr, c_double, magic_fields_in
         type(c_ptr), value :: rl
         type(magic_fields_in), intent(in) :: fin
         integer(c_int), value :: mask, ktops, kbots
         real(c_double), intent(out) :: out(40,*)
         integer(c_int) :: ierr
      end function magic_rloop_diagnostics

      !-- get_dtBLM (dtB.f90:144-223) for all local levels: out(lm_max, n_r_loc, 11)
      function magic_rloop_dtb(rl, fin, out) bind(C, name='magic_rloop_dtb') result(ierr)
         import :: c_int, c_ptr, c_double_complex, magic_fields_in
         type(c_ptr), value :: rl
         type(magic_fields_in), intent(in) :: fin
         complex(c_double_complex), intent(out) :: out(*)
         integer(c_int) :: ierr
      end function magic_rloop_dtb

      !-- torsional-oscillation sums (rIter.f90:395-404): getTOnext's grid part keeps Bs, Bp, Bz on the device; getTO returns
      !   out(n_theta_max, MAGIC_NTO, n_r_loc), colatitudes in geographic order, arrays as listed in include/magic_sht.h
      function magic_rloop_to_next(rl, fin) bind(C, name='magic_rloop_to_next') result(ierr)
         import :: c_int, c_ptr, magic_fields_in
         type(c_ptr), value :: rl
         type(magic_fields_in), intent(in) :: fin
         integer(c_int) :: ierr
      end function magic_rloop_to_next

      function magic_rloop_to(rl, fin, dtLast, out) bind(C, name='magic_rloop_to') result(ierr)
         import :: c_int, c_ptr, c_double, magic_fields_in
         type(c_ptr), value :: rl
         type(magic_fields_in), intent(in) :: fin
         real(c_double), value :: dtLast
         real(c_double), intent(out) :: out(*)
         integer(c_int) :: ierr
      end function magic_rloop_to

      !-- r.m.s. force balance (rIter.f90:215-252, 710; RMS.f90:469-610): out(lm_max, n_r_loc, MAGIC_NRMS), arrays as listed in
      !   include/magic_sht.h; magic_rloop_rms_keep keeps the flow potentials of this step for the next call's dtV terms
      function magic_rloop_rms_keep(rl, fin) bind(C, name='magic_rloop_rms_keep') result(ierr)
         import :: c_int, c_ptr, magic_fields_in
         type(c_ptr), value :: rl
         type(magic_fields_in), intent(in) :: fin
         integer(c_int) :: ierr
      end function magic_rloop_rms_keep

      function magic_rloop_rms(rl, fin, dt, out) bind(C, name='magic_rloop_rms') result(ierr)
         import :: c_int, c_ptr, c_double, c_double_complex, magic_fields_in
         type(c_ptr), value :: rl
         type(magic_fields_in), intent(in) :: fin
         real(c_double), value :: dt
         complex(c_double_complex), intent(out) :: out(*)
         integer(c_int) :: ierr
      end function magic_rloop_rms

      !---------------------------------------------------------------- r <-> LM transposer
      function magic_transp_unique_id(id) bind(C, name='magic_transp_unique_id') result(ierr)
         import :: c_int, c_char
         character(kind=c_char), intent(out) :: id(128)
         integer(c_int) :: ierr
      end function magic_transp_unique_id

      function magic_transp_create(h, id, rank, n_procs, n_r_max, n_fields, t) bind(C, name='magic_transp_create') result(ierr)
         import :: c_int, c_ptr, c_char
         type(c_ptr), value :: h
         character(kind=c_char), intent(in) :: id(128)
         integer(c_int), value :: rank, n_procs, n_r_max, n_fields
         type(c_ptr), intent(out) :: t
         integer(c_int) :: ierr
      end function magic_transp_create

      function magic_transp_destroy(t) bind(C, name='magic_transp_destroy') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: t
         integer(c_int) :: ierr
      end function magic_transp_destroy

      function magic_transp_extents(t, llm, ulm, nRstart, nRstop) bind(C, name='magic_transp_extents') result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: t
         integer(c_int), intent(out) :: llm, ulm, nRstart, nRstop
         integer(c_int) :: ierr
      end function magic_transp_extents

      function magic_transp_lm2r(t, arr_LMloc, arr_Rloc) bind(C, name='magic_transp_lm2r') result(ierr)
         import :: c_int, c_ptr, c_double_complex
         type(c_ptr), value :: t
         complex(c_double_complex), intent(in)  :: arr_LMloc(*)
         complex(c_double_complex), intent(out) :: arr_Rloc(*)
         integer(c_int) :: ierr
      end function magic_transp_lm2r

      function magic_transp_r2lm(t, arr_Rloc, arr_LMloc) bind(C, name='magic_transp_r2lm') result(ierr)
         import :: c_int, c_ptr, c_double_complex
         type(c_ptr), value :: t
         complex(c_double_complex), intent(in)  :: arr_Rloc(*)
         complex(c_double_complex), intent(out) :: arr_LMloc(*)
         integer(c_int) :: ierr
      end function magic_transp_r2lm

   end interface

contains

   subroutine magic_check(ierr, where)
      !
      ! The reference has no error codes: failure is abortRun -> MPI_Abort (useful.f90:271-302).
      !
      use useful, only: abortRun
      integer(c_int),   intent(in) :: ierr
      character(len=*), intent(in) :: where

      character(kind=c_char), pointer :: cmsg(:)
      character(len=256) :: msg
      type(c_ptr) :: p
      integer :: i, n

      if ( ierr == 0 ) return
      msg = ''
      p = magic_last_error()
      if ( c_associated(p) ) then
         n = min(int(c_strlen(p)), len(msg))
         if ( n > 0 ) then
            call c_f_pointer(p, cmsg, [n])
            do i=1,n
               msg(i:i) = cmsg(i)
            end do
         end if
      end if
      call abortRun('! '//where//' failed: '//trim(msg))

   end subroutine magic_check
!------------------------------------------------------------------------------
   function addr_z(a) result(p)
      !
      ! Address of the first element of a contiguous complex array.  The deferred interfaces of rIteration.f90 and
      ! mpi_transpose.f90 fix the characteristics of their dummies (no TARGET), so c_loc cannot be applied to them
      ! directly; an explicit-shape or assumed-size actual is passed here by sequence association, i.e. without a copy.
      !
      complex(c_double_complex), target, intent(in) :: a(*)
      type(c_ptr) :: p
      p = c_loc(a)
   end function addr_z
!------------------------------------------------------------------------------
   function addr_r(a) result(p)
      real(c_double), target, intent(in) :: a(*)
      type(c_ptr) :: p
      p = c_loc(a)
   end function addr_r

end module magic_b200_c
