module rIter_cuda_mod
   !
   ! rIter_cuda_t: a second extension of the abstract rIter_t (rIteration.f90:15-19) next to rIter_single_t
   ! (rIter.f90:53-63).  One call hands the whole R-local slab (all levels nRstart:nRstop of every field) to
   ! libmagic_b200.so, which runs transform_to_grid_space, get_nl, transform_to_lm_space, courant and get_td
   ! (rIter.f90:190-444) for all levels at once on the GPU and returns the same output arrays.
   !
   ! Selected in radialLoop.f90:26 instead of `allocate( rIter_single_t :: rIter )`:
   !      allocate( rIter_cuda_t :: rIter )
   ! step_time.f90 and LMLoop are untouched.
   !
   ! On steps that ask for in-loop diagnostics (graphics, movies, TO, helicity, power, RMS, fluxes, ...:
   ! rIter.f90:303-404) the call is delegated to an embedded rIter_single_t, whose transforms go through
   ! `module sht`, i.e. the same library one level at a time.
   !
   use iso_c_binding
   use precision_mod
   use truncation, only: lm_max, lm_maxMag, n_r_max, n_theta_max, n_phi_max, nlat_padded
   use radial_data, only: nRstart, nRstop, nRstartMag, nRstopMag, n_r_cmb, n_r_icb
   use logic, only: l_conv, l_mag, l_heat, l_conv_nl, l_heat_nl, l_mag_nl, l_mag_LF, l_mag_kin, l_anel,    &
       &            l_adv_curl, l_corr, l_double_curl, l_single_matrix, l_chemical_conv, l_precession,      &
       &            l_centrifuge, l_anelastic_liquid, l_cour_alf_damp, l_full_sphere, l_parallel_solve,     &
       &            l_temperature_diff, l_cond_ma, l_cond_ic, l_rot_ma, l_rot_ic, l_b_nl_cmb, l_b_nl_icb,   &
       &            l_phase_field, l_onset, l_dtB, l_dtphaseMovie, l_RMS
   use special, only: lGrenoble
   use physical_parameters, only: ktopv, kbotv, n_r_LCR, LFfac, CorFac, epsc, epscXi, opm, ViscHeatFac,     &
       &                          OhmLossFac, oek, po, prec_angle, dilution_fac, ra, opr, ktops, kbots,     &
       &                          ThExpNb, epsPhase, phaseDiffFac, penaltyFac, tmelt
   use radial_functions, only: r, or1, or2, or4, orho1, orho2, beta, rho0, otemp1, temp0, visc, lambda,     &
       &                       epscProf, l_R, r_cmb, r_icb, alpha0
   !-- per-level sums of the in-loop diagnostics (module variables of the reference, to be made public there)
   use outMisc_mod, only: HelASr, Hel2ASr, HelnaASr, Helna2ASr, HelEAASr, hemi_ekin_r, hemi_vrabs_r,        &
       &                  hemi_emag_r, hemi_brabs_r, ekinSr, ekinLr, volSr, phase_Rloc, temp_Rloc, dtemp_Rloc
   use power, only: viscASr
   !-- r.m.s. force balance: the spectra transform_to_lm_RMS fills (module variables of RMS, to be made public there) and the
   !   routine that consumes them
   use RMS, only: compute_lm_forces, dtVrLM, dtVtLM, dtVpLM, dpkindrLM, Advt2LM, Advp2LM, PFt2LM, PFp2LM, LFt2LM, LFp2LM,     &
       &          CFt2LM, CFp2LM, LFrLM
   !-- torsional oscillations: the (r,theta) arrays getTO fills, and the routines that hold the spectral part
   use torsional_oscillations, only: prep_TO_axi, getTOnext, getTOfinish, V2AS_Rloc, VAS_Rloc, dzCorAS_Rloc, dzRstrAS_Rloc,   &
       &                             dzAstrAS_Rloc, dzLFAS_Rloc, Bs2AS_Rloc, BspAS_Rloc, BpzAS_Rloc, BszAS_Rloc, BspdAS_Rloc, &
       &                             BpsdAS_Rloc, BzpdAS_Rloc, BpzdAS_Rloc, dzPenAS_Rloc
   !-- get_dtBLM's per-level results (module variables of dtB_mod, to be made public there) and the routine that consumes them
   use dtB_mod, only: BtVrLM, BpVrLM, BrVtLM, BrVpLM, BtVpLM, BpVtLM, BpVtBtVpCotLM, BpVtBtVpSn2LM, BrVZLM, BtVZLM,   &
       &              BtVZsn2LM, get_dH_dtBLM
   use outPar_mod, only: EperpASr, EparASr, EperpaxiASr, EparaxiASr, fkinASr, fconvASr, fviscASr, fresASr,  &
       &                 fpoynASr, uhASr, duhASr, gradT2ASr
   use num_param, only: delxr2, delxh2
   use fields, only: s_Rloc, ds_Rloc, z_Rloc, dz_Rloc, p_Rloc, b_Rloc, db_Rloc, ddb_Rloc, aj_Rloc, dj_Rloc, phi_Rloc, &
       &             w_Rloc, dw_Rloc, ddw_Rloc, xi_Rloc, omega_ic, omega_ma,                                &
       &             flow_LMloc_container, s_LMloc_container, field_LMloc_container, xi_LMloc_container
   use fieldsLast, only:    dflowdt_LMloc_container, dsdt_LMloc_container, dbdt_LMloc_container,            &
       &                    dxidt_LMloc_container
   use blocking, only: llm
   use mpi_transp_cuda_mod, only: l_fused_lm, l_outputs_in_lm, n_pending, transp5, run_pending_lm2r
   use time_schemes, only: type_tscheme
   use rIteration, only: rIter_t
   use rIter_mod, only: rIter_single_t
   use useful, only: abortRun
   use constants, only: zero
   use sht, only: sht_h, scal_to_spat
   use magic_b200_c

   implicit none

   private

   type, public, extends(rIter_t) :: rIter_cuda_t
      type(c_ptr) :: rl = c_null_ptr          ! magic_rloop*: plan, device buffers, streams
      type(rIter_single_t) :: single          ! level-at-a-time loop for the diagnostics steps
      integer(c_int), allocatable :: nR_loc(:), l_R_loc(:)
      real(c_double), allocatable :: rad(:,:) ! the 15 radial functions of magic_radial on nRstart:nRstop
   contains
      procedure :: initialize
      procedure :: finalize
      procedure :: radialLoop
      procedure, private :: create_plan
   end type rIter_cuda_t

   !-- results of the r.m.s. and dtB batches: (lm_max, nRstart:nRstop, 14 / 11), allocated on first use and page-locked (a pageable
   !   destination takes these gigabytes at a tenth of the PCIe rate); module variables because c_loc needs a target and the
   !   passed-object dummy of the deferred radialLoop interface (rIteration.f90:43) is not one
   complex(c_double_complex), allocatable, target, save :: rq(:,:,:), dtb(:,:,:)

contains

   subroutine initialize(this)

      class(rIter_cuda_t) :: this

      if ( l_onset ) call abortRun('! rIter_cuda_t: onset mode is not on the GPU path')
      !-- get_lorentz_torque adds the imposed-field term b0r with lGrenoble (outRot.f90:465-478): not on the GPU path
      if ( lGrenoble ) call abortRun('! rIter_cuda_t: lGrenoble (imposed b0 in the Lorentz torque) is not on the GPU path')
      call this%single%initialize()
      !-- the plan needs tscheme%courfac / alffac: it is created by the first radialLoop call

   end subroutine initialize
!------------------------------------------------------------------------------
   subroutine finalize(this)

      class(rIter_cuda_t) :: this

      if ( c_associated(this%rl) ) call magic_check( magic_rloop_destroy(this%rl), 'magic_rloop_destroy' )
      if ( allocated(rq) ) deallocate( rq )      ! after the plan: magic_rloop_destroy releases the page locks it holds
      if ( allocated(dtb) ) deallocate( dtb )
      this%rl = c_null_ptr
      if ( allocated(this%rad) ) deallocate( this%rad, this%nR_loc, this%l_R_loc )
      call this%single%finalize()

   end subroutine finalize
!------------------------------------------------------------------------------
   subroutine create_plan(this, tscheme)
      !
      ! Fills magic_params from logic / physical_parameters and magic_radial from radial_functions(nRstart:nRstop)
      ! (the values get_nl.f90:24-34, get_td.f90:11-21, courant.f90 and rIter.f90:16-40 read) and creates the plan.
      !
      class(rIter_cuda_t), target :: this
      class(type_tscheme), intent(in) :: tscheme

      type(magic_params) :: p
      type(magic_radial) :: rd
      integer :: n_r_loc, n, nR

      n_r_loc = nRstop-nRstart+1
      allocate( this%nR_loc(n_r_loc), this%l_R_loc(n_r_loc), this%rad(n_r_loc,15) )
      do n=1,n_r_loc
         nR = nRstart+n-1
         this%nR_loc(n)  = int(nR,c_int)
         this%l_R_loc(n) = int(l_R(nR),c_int)
         this%rad(n,1)  = r(nR);      this%rad(n,2)  = or1(nR);    this%rad(n,3)  = or2(nR)
         this%rad(n,4)  = or4(nR);    this%rad(n,5)  = orho1(nR);  this%rad(n,6)  = orho2(nR)
         this%rad(n,7)  = beta(nR);   this%rad(n,8)  = rho0(nR);   this%rad(n,9)  = otemp1(nR)
         this%rad(n,10) = temp0(nR);  this%rad(n,11) = visc(nR);   this%rad(n,12) = lambda(nR)
         this%rad(n,13) = epscProf(nR); this%rad(n,14) = delxr2(nR); this%rad(n,15) = delxh2(nR)
      end do
      rd%nR  = c_loc(this%nR_loc);   rd%l_R = c_loc(this%l_R_loc)
      rd%r        = c_loc(this%rad(1,1));  rd%or1    = c_loc(this%rad(1,2));  rd%or2    = c_loc(this%rad(1,3))
      rd%or4      = c_loc(this%rad(1,4));  rd%orho1  = c_loc(this%rad(1,5));  rd%orho2  = c_loc(this%rad(1,6))
      rd%beta     = c_loc(this%rad(1,7));  rd%rho0   = c_loc(this%rad(1,8));  rd%otemp1 = c_loc(this%rad(1,9))
      rd%temp0    = c_loc(this%rad(1,10)); rd%visc   = c_loc(this%rad(1,11)); rd%lambda = c_loc(this%rad(1,12))
      rd%epscProf = c_loc(this%rad(1,13)); rd%delxr2 = c_loc(this%rad(1,14)); rd%delxh2 = c_loc(this%rad(1,15))

      p%l_conv=l2i(l_conv); p%l_mag=l2i(l_mag); p%l_heat=l2i(l_heat); p%l_conv_nl=l2i(l_conv_nl)
      p%l_heat_nl=l2i(l_heat_nl); p%l_mag_nl=l2i(l_mag_nl); p%l_mag_LF=l2i(l_mag_LF); p%l_mag_kin=l2i(l_mag_kin)
      p%l_anel=l2i(l_anel); p%l_adv_curl=l2i(l_adv_curl); p%l_corr=l2i(l_corr); p%l_double_curl=l2i(l_double_curl)
      p%l_single_matrix=l2i(l_single_matrix); p%l_chemical_conv=l2i(l_chemical_conv)
      p%l_precession=l2i(l_precession); p%l_centrifuge=l2i(l_centrifuge)
      p%l_anelastic_liquid=l2i(l_anelastic_liquid); p%l_cour_alf_damp=l2i(l_cour_alf_damp)
      p%l_full_sphere=l2i(l_full_sphere); p%l_parallel_solve=l2i(l_parallel_solve)
      p%l_temperature_diff=l2i(l_temperature_diff)
      p%ktopv=int(ktopv,c_int); p%kbotv=int(kbotv,c_int)
      p%l_cond_ma=l2i(l_cond_ma); p%l_cond_ic=l2i(l_cond_ic); p%l_rot_ma=l2i(l_rot_ma); p%l_rot_ic=l2i(l_rot_ic)
      p%n_r_max=int(n_r_max,c_int); p%n_r_LCR=int(n_r_LCR,c_int)
      p%LFfac=LFfac; p%CorFac=CorFac; p%epsc=epsc; p%epscXi=epscXi; p%opm=opm
      p%ViscHeatFac=ViscHeatFac; p%OhmLossFac=OhmLossFac
      p%oek=oek; p%po=po; p%prec_angle=prec_angle; p%dilution_fac=dilution_fac; p%ra=ra; p%opr=opr
      p%omega_ma=omega_ma; p%omega_ic=omega_ic; p%r_cmb=r_cmb; p%r_icb=r_icb
      p%courfac=tscheme%courfac; p%alffac=tscheme%alffac
      p%epsPhase=epsPhase; p%phaseDiffFac=phaseDiffFac; p%penaltyFac=penaltyFac; p%tmelt=tmelt
      p%l_phase_field=l2i(l_phase_field)

      !-- level_chunk = 0: the library picks its level batch from the truncation and the field set (the same on every rank)
      call magic_check( magic_rloop_create(sht_h, p, rd, int(n_r_loc,c_int), 0_c_int, this%rl), 'magic_rloop_create' )

      !-- page-lock the persistent R-distributed input containers of fields.f90 once, so that the uploads of one level chunk
      !   overlap the compute of another (the library never pins caller memory on its own); finalize unpins them
      if ( l_conv .or. l_mag_kin ) then
         call pin(c_loc(w_Rloc), size(w_Rloc));  call pin(c_loc(dw_Rloc), size(dw_Rloc));  call pin(c_loc(ddw_Rloc), size(ddw_Rloc))
         call pin(c_loc(z_Rloc), size(z_Rloc));  call pin(c_loc(dz_Rloc), size(dz_Rloc))
      end if
      if ( l_heat ) call pin(c_loc(s_Rloc), size(s_Rloc))
      if ( l_chemical_conv ) call pin(c_loc(xi_Rloc), size(xi_Rloc))
      if ( l_mag .or. l_mag_LF ) then
         call pin(c_loc(b_Rloc), size(b_Rloc));   call pin(c_loc(db_Rloc), size(db_Rloc))
         call pin(c_loc(ddb_Rloc), size(ddb_Rloc))
         call pin(c_loc(aj_Rloc), size(aj_Rloc)); call pin(c_loc(dj_Rloc), size(dj_Rloc))
      end if

   contains

      pure integer(c_int) function l2i(l)
         logical, intent(in) :: l
         l2i = merge(1_c_int, 0_c_int, l)
      end function l2i

      subroutine pin(ptr, n_complex)
         type(c_ptr), intent(in) :: ptr
         integer,     intent(in) :: n_complex
         call magic_check( magic_rloop_pin_host(this%rl, ptr, int(16,c_size_t)*int(n_complex,c_size_t)), 'magic_rloop_pin_host' )
      end subroutine pin

   end subroutine create_plan
!------------------------------------------------------------------------------
   subroutine radialLoop(this,l_graph,l_frame,time,timeStage,tscheme,dtLast,         &
              &          lTOCalc,lTONext,lTONext2,lHelCalc,lPowerCalc,lRmsCalc,      &
              &          lPressCalc,lPressNext,lViscBcCalc,lFluxProfCalc,            &
              &          lPerpParCalc,lGeosCalc,lHemiCalc,lPhaseCalc,l_probe_out,    &
              &          dsdt,dwdt,dzdt,dpdt,dxidt,dphidt,dbdt,djdt,dVxVhLM,dVxBhLM, &
              &          dVSrLM,dVXirLM,lorentz_torque_ic,lorentz_torque_ma,         &
              &          br_vt_lm_cmb,br_vp_lm_cmb,br_vt_lm_icb,br_vp_lm_icb,dtrkc,  &
              &          dthkc)

      class(rIter_cuda_t) :: this

      !--- Input of variables (rIteration.f90:34-81):
      logical,             intent(in) :: l_graph,l_frame
      logical,             intent(in) :: lTOcalc,lTONext,lTONext2,lHelCalc
      logical,             intent(in) :: lPowerCalc,lHemiCalc
      logical,             intent(in) :: lViscBcCalc,lFluxProfCalc,lPerpParCalc
      logical,             intent(in) :: lRmsCalc,lGeosCalc,lPhaseCalc
      logical,             intent(in) :: l_probe_out
      logical,             intent(in) :: lPressCalc
      logical,             intent(in) :: lPressNext
      real(cp),            intent(in) :: time,timeStage,dtLast
      class(type_tscheme), intent(in) :: tscheme

      !---- Output of explicit time step:
      complex(cp), intent(out) :: dwdt(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dzdt(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dpdt(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dsdt(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dxidt(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dphidt(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dVSrLM(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dVXirLM(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dbdt(lm_maxMag,nRstartMag:nRstopMag)
      complex(cp), intent(out) :: djdt(lm_maxMag,nRstartMag:nRstopMag)
      complex(cp), intent(out) :: dVxVhLM(lm_max,nRstart:nRstop)
      complex(cp), intent(out) :: dVxBhLM(lm_maxMag,nRstartMag:nRstopMag)
      real(cp),    intent(out) :: lorentz_torque_ma,lorentz_torque_ic
      complex(cp), intent(out) :: br_vt_lm_cmb(:)
      complex(cp), intent(out) :: br_vp_lm_cmb(:)
      complex(cp), intent(out) :: br_vt_lm_icb(:)
      complex(cp), intent(out) :: br_vp_lm_icb(:)
      real(cp),    intent(out) :: dtrkc(nRstart:nRstop),dthkc(nRstart:nRstop)

      !-- Local variables
      type(magic_fields_in)  :: fin
      type(magic_fields_out) :: fout
      type(magic_lm_in)      :: lin
      type(magic_lm_out)     :: lout
      integer :: ist, mask, nR
      logical :: l_diag, l_rms_dev
      real(c_double), allocatable :: dg(:,:)
      real(cp), allocatable :: grd(:,:)
      real(c_double), allocatable :: tq(:,:,:)

      !-- Inputs: the R-distributed containers of fields.f90:211-268, (lm_max, nRstart:nRstop) each; the library
      !   ignores the pointers of switched-off physics
      fin = magic_fields_in(c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, &
      &                     c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, &
      &                     c_null_ptr)
      if ( l_phase_field ) fin%phi = addr_z(phi_Rloc)   ! allocatable without TARGET in fields.f90:47
      if ( l_conv .or. l_mag_kin ) then
         fin%w = c_loc(w_Rloc);  fin%dw = c_loc(dw_Rloc);  fin%ddw = c_loc(ddw_Rloc)
         fin%z = c_loc(z_Rloc);  fin%dz = c_loc(dz_Rloc)
      end if
      if ( l_heat ) then
         fin%s = c_loc(s_Rloc);  fin%ds = c_loc(ds_Rloc)
      end if
      if ( l_chemical_conv ) fin%xi = c_loc(xi_Rloc)
      if ( l_mag .or. l_mag_LF ) then
         fin%b  = c_loc(b_Rloc);   fin%db = c_loc(db_Rloc);  fin%ddb = c_loc(ddb_Rloc)
         fin%aj = c_loc(aj_Rloc);  fin%dj = c_loc(dj_Rloc)
      end if

      !-- l_RMS on the device (the batch has no r = 0 level: full-sphere runs keep the reference's loop)
      l_rms_dev = l_RMS .and. .not. l_full_sphere

      !-- Log steps: get_helicity, get_hemi, get_visc_heat, get_perpPar, get_fluxes, get_nlBLayers and get_ekin_solid_liquid
      !   (rIter.f90:320-367) and the torsional-oscillation sums (getTOnext / getTO, rIter.f90:395-404) are evaluated on the device
      !   after the batched loop (below).  The remaining output hooks keep the reference's level-at-a-time loop (its transforms
      !   still run on the GPU)
      if ( l_graph .or. l_frame .or. ( l_RMS .and. .not. l_rms_dev ) .or.                                    &
      &    ( l_full_sphere .and. ( lTOCalc .or. lTONext .or. lTONext2 ) ) .or.                               &
      &    lGeosCalc .or. l_probe_out .or. ( lPressNext .and. l_double_curl ) .or.                           &
      &    ( l_full_sphere .and. ( lHelCalc .or. lPowerCalc .or. lViscBcCalc .or. lFluxProfCalc .or.        &
      &                            lPerpParCalc .or. lHemiCalc .or. lPhaseCalc ) ) ) then
         !-- ( lPressNext with the double-curl equation: the reference also calls get_dpdt then (rIter.f90:420); the batched
         !   loop only produces dpdt in the pressure formulation, so that step takes the level-at-a-time loop )
         if ( l_fused_lm ) then   ! the recorded transposes become real: the level-at-a-time loop reads the host R arrays
            call run_pending_lm2r()
            l_outputs_in_lm = .false.
         end if
         call this%single%radialLoop(l_graph,l_frame,time,timeStage,tscheme,dtLast,lTOCalc,lTONext,lTONext2,   &
              &                      lHelCalc,lPowerCalc,lRmsCalc .and. .not. l_rms_dev,lPressCalc,lPressNext, &
              &                      lViscBcCalc,                                                              &
              &                      lFluxProfCalc,lPerpParCalc,lGeosCalc,lHemiCalc,lPhaseCalc,l_probe_out,    &
              &                      dsdt,dwdt,dzdt,dpdt,dxidt,dphidt,dbdt,djdt,dVxVhLM,dVxBhLM,dVSrLM,dVXirLM,&
              &                      lorentz_torque_ic,lorentz_torque_ma,br_vt_lm_cmb,br_vp_lm_cmb,            &
              &                      br_vt_lm_icb,br_vp_lm_icb,dtrkc,dthkc)
         !-- the r.m.s. batch of this step still runs on the device (with lRmsCalc off the host loop's get_nl_RMS only refreshes
         !   its own copy of the previous velocity): one owner of vr_old on every step
         if ( l_rms_dev ) then
            if ( .not. c_associated(this%rl) ) call this%create_plan(tscheme)
            call rms_on_device()
         end if
         return
      end if

      if ( .not. c_associated(this%rl) ) call this%create_plan(tscheme)

      !-- Fused mode: LM-distributed host containers in, LM-distributed explicit terms out -- the transposes on either side of
      !   this call (step_time.f90:485, :612) are part of it (mpi_transp_cuda_mod); the explicit terms go into the slice
      !   tscheme%istage of the time-array containers, which is where transp_Rloc_to_LMloc would put them (step_time.f90:1134-1245)
      l_diag = lHelCalc .or. lPowerCalc .or. lViscBcCalc .or. lFluxProfCalc .or. lPerpParCalc .or. lHemiCalc .or. lPhaseCalc &
      &        .or. lTOCalc .or. lTONext .or. lTONext2 .or. l_RMS
      if ( l_fused_lm .and. n_pending > 0 .and.                                                    &
      &    .not. ( l_b_nl_cmb .or. l_b_nl_icb .or. l_diag .or. l_dtB .or. l_phase_field ) ) then
         ist = tscheme%istage
         lin  = magic_lm_in(c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr)
         lout = magic_lm_out(c_null_ptr, c_null_ptr, c_null_ptr, addr_r(dtrkc), addr_r(dthkc), c_null_ptr)
         if ( l_conv .or. l_mag_kin ) lin%flow = c_loc(flow_LMloc_container)
         if ( l_heat ) lin%s = c_loc(s_LMloc_container)
         if ( l_mag .or. l_mag_LF ) lin%field = c_loc(field_LMloc_container)
         if ( l_chemical_conv ) lin%xi = c_loc(xi_LMloc_container)
         if ( l_conv ) lout%dflowdt = c_loc(dflowdt_LMloc_container(llm,1,1,ist))
         if ( l_heat ) lout%dsdt = c_loc(dsdt_LMloc_container(llm,1,1,ist))
         if ( l_mag ) lout%dbdt = c_loc(dbdt_LMloc_container(llm,1,1,ist))
         if ( l_chemical_conv ) lout%dxidt = c_loc(dxidt_LMloc_container(llm,1,1,ist))
         call magic_check( magic_rloop_set_rotation(this%rl, omega_ma, omega_ic), 'magic_rloop_set_rotation' )
         call magic_check( magic_rloop_run_lm(this%rl, transp5, lin, lout, timeStage), 'magic_rloop_run_lm' )
         call magic_check( magic_rloop_get_torques(this%rl, lorentz_torque_ic, lorentz_torque_ma), &
              &            'magic_rloop_get_torques' )
         n_pending = 0
         l_outputs_in_lm = .true.
         dphidt(:,:) = zero
         return
      end if
      if ( l_fused_lm ) then   ! (nonlinear magnetic boundary products need the R-distributed path: rIter.f90:267-277)
         call run_pending_lm2r()
         l_outputs_in_lm = .false.
      end if

      fout = magic_fields_out(c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, &
      &                       c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr)
      if ( l_phase_field ) fout%dphidt = addr_z(dphidt)
      fout%dwdt = addr_z(dwdt);  fout%dzdt = addr_z(dzdt)
      if ( l_double_curl ) then
         fout%dVxVhLM = addr_z(dVxVhLM)
      else
         fout%dpdt = addr_z(dpdt)
      end if
      if ( l_heat ) then
         fout%dsdt = addr_z(dsdt);  fout%dVSrLM = addr_z(dVSrLM)
      end if
      if ( l_chemical_conv ) then
         fout%dxidt = addr_z(dxidt);  fout%dVXirLM = addr_z(dVXirLM)
      end if
      if ( l_mag ) then
         fout%dbdt = addr_z(dbdt);  fout%djdt = addr_z(djdt);  fout%dVxBhLM = addr_z(dVxBhLM)
      end if
      fout%dtrkc = addr_r(dtrkc);  fout%dthkc = addr_r(dthkc)

      !-- omega_ma / omega_ic change from step to step when the walls rotate (v_rigid_boundary)
      call magic_check( magic_rloop_set_rotation(this%rl, omega_ma, omega_ic), 'magic_rloop_set_rotation' )

      !-- The loop: the explicit terms of all local levels (timeStage enters the precession terms, get_nl.f90:346-357)
      call magic_check( magic_rloop_run(this%rl, fin, fout, timeStage), 'magic_rloop_run' )

      !-- rIter.f90:279-292,461: Lorentz torques on a conducting, rotating inner core / mantle
      call magic_check( magic_rloop_get_torques(this%rl, lorentz_torque_ic, lorentz_torque_ma), &
           &            'magic_rloop_get_torques' )

      !-- rIter.f90:267-277: products for the nonlinear magnetic boundary conditions (stress-free + conducting wall)
      if ( l_b_nl_cmb .and. nRstart == n_r_cmb ) then
         call magic_check( magic_rloop_get_br_v_bcs(this%rl, 0_c_int, br_vt_lm_cmb, br_vp_lm_cmb), 'get_br_v_bcs CMB' )
      end if
      if ( l_b_nl_icb .and. nRstop == n_r_icb ) then
         call magic_check( magic_rloop_get_br_v_bcs(this%rl, 1_c_int, br_vt_lm_icb, br_vp_lm_icb), 'get_br_v_bcs ICB' )
      end if

      if ( .not. l_phase_field ) dphidt(:,:) = zero

      !-- rIter.f90:388-395, 442 with l_dtB: the eleven products of get_dtBLM and their analyses for all local levels as one
      !   batch on the device; get_dH_dtBLM then combines them level by level as in the reference
      if ( l_dtB ) then
         if ( .not. allocated(dtb) ) then
            allocate( dtb(lm_max,nRstart:nRstop,11) )
            call magic_check( magic_rloop_pin_host(this%rl, c_loc(dtb), int(16,c_size_t)*size(dtb,kind=c_size_t)), &
                 &            'magic_rloop_pin_host' )
         end if
         call magic_check( magic_rloop_dtb(this%rl, fin, dtb), 'magic_rloop_dtb' )
         do nR=nRstart,nRstop
            BtVrLM(:)=dtb(:,nR,1);  BpVrLM(:)=dtb(:,nR,2);  BrVtLM(:)=dtb(:,nR,3);  BrVpLM(:)=dtb(:,nR,4)
            BtVpLM(:)=dtb(:,nR,5);  BpVtLM(:)=dtb(:,nR,6);  BpVtBtVpCotLM(:)=dtb(:,nR,7)
            BpVtBtVpSn2LM(:)=dtb(:,nR,8)
            BrVZLM(:)=dtb(:,nR,9);  BtVZLM(:)=dtb(:,nR,10);  BtVZsn2LM(:)=dtb(:,nR,11)
            call get_dH_dtBLM(nR)
         end do
      end if

      if ( l_rms_dev ) call rms_on_device()

      !-- rIter.f90:220-222, 395-404, 438 on torsional-oscillation steps: the grid part of getTOnext (Bs, Bp, Bz of the step before
      !   the output) stays on the device, getTO's azimuthal means come back as (theta, array, level); the spectral, axisymmetric
      !   part (prep_TO_axi, getTOnext's dzdVp / dzddVp bookkeeping, getTOfinish) is the reference's own code, level by level --
      !   getTOnext gets zero grids: its BsLast / BpLast / BzLast are only read by getTO, which no longer runs on the host
      if ( lTOCalc .or. lTONext .or. lTONext2 ) then
         if ( lTONext .and. ( .not. lTONext2 ) .and. l_mag ) &
         &  call magic_check( magic_rloop_to_next(this%rl, fin), 'magic_rloop_to_next' )
         if ( lTOCalc ) then
            allocate( tq(n_theta_max,MAGIC_NTO,nRstart:nRstop) )
            call magic_check( magic_rloop_to(this%rl, fin, real(dtLast,c_double), tq), 'magic_rloop_to' )
         end if
         allocate( grd(n_theta_max,n_phi_max) )
         grd(:,:)=0.0_cp
         do nR=nRstart,nRstop
            call prep_TO_axi(z_Rloc(:,nR), dz_Rloc(:,nR))
            if ( lTONext .or. lTONext2 ) call getTOnext(grd,grd,grd,lTONext,lTONext2,tscheme%dt(1),dtLast,nR)
            if ( lTOCalc ) then
               V2AS_Rloc(:,nR)    =tq(:,1,nR);   VAS_Rloc(:,nR)     =tq(:,2,nR);   dzCorAS_Rloc(:,nR)=tq(:,3,nR)
               dzRstrAS_Rloc(:,nR)=tq(:,4,nR);   dzAstrAS_Rloc(:,nR)=tq(:,5,nR)
               if ( l_mag ) then
                  dzLFAS_Rloc(:,nR)=tq(:,6,nR);   Bs2AS_Rloc(:,nR) =tq(:,7,nR);   BspAS_Rloc(:,nR) =tq(:,8,nR)
                  BpzAS_Rloc(:,nR) =tq(:,9,nR);   BszAS_Rloc(:,nR) =tq(:,10,nR);  BspdAS_Rloc(:,nR)=tq(:,11,nR)
                  BpsdAS_Rloc(:,nR)=tq(:,12,nR);  BzpdAS_Rloc(:,nR)=tq(:,13,nR);  BpzdAS_Rloc(:,nR)=tq(:,14,nR)
               end if
               if ( l_phase_field ) dzPenAS_Rloc(:,nR)=tq(:,15,nR)
               call getTOfinish(nR, dtLast)
            end if
         end do
         deallocate( grd )
         if ( lTOCalc ) deallocate( tq )
      end if

      !-- rIter.f90:320-367 on log steps: one call returns the per-level sums of all requested routines; they go where the
      !   reference's routines store them (the arrays below are module variables of outMisc_mod, power and outPar_mod, to be
      !   made public there: HelASr .. HelEAASr, hemi_*_r, viscASr, EperpASr .. EparaxiASr, fkinASr .. fpoynASr, uhASr ..)
      if ( l_diag .and. ( lHelCalc .or. lPowerCalc .or. lViscBcCalc .or. lFluxProfCalc .or. lPerpParCalc .or. lHemiCalc .or. &
      &                   lPhaseCalc ) ) then
         mask = 0
         if ( lHelCalc )      mask = mask + MAGIC_DIAG_HEL
         if ( lHemiCalc )     mask = mask + MAGIC_DIAG_HEMI
         if ( lPowerCalc )    mask = mask + MAGIC_DIAG_POWER
         if ( lPerpParCalc )  mask = mask + MAGIC_DIAG_PERPPAR
         if ( lFluxProfCalc ) mask = mask + MAGIC_DIAG_FLUX
         if ( lRmsCalc )      mask = mask + MAGIC_DIAG_RMSBULK   ! boundary levels are bulk levels on these steps (rIter.f90:215)
         if ( lViscBcCalc )   mask = mask + MAGIC_DIAG_VISCBC
         if ( lPhaseCalc )    mask = mask + MAGIC_DIAG_PHASE
         if ( lFluxProfCalc ) fin%p = c_loc(p_Rloc)    ! lPressCalc is set with lFluxProfCalc (step_time.f90:399)
         allocate( dg(40,nRstart:nRstop) )
         call magic_check( magic_rloop_diagnostics(this%rl, fin, int(mask,c_int), int(ktops,c_int), int(kbots,c_int), dg), &
              &            'magic_rloop_diagnostics' )
         do nR=nRstart,nRstop
            if ( lHelCalc ) then
               HelASr(nR,:)=dg(1:2,nR);  Hel2ASr(nR,:)=dg(3:4,nR);  HelnaASr(nR,:)=dg(5:6,nR)
               Helna2ASr(nR,:)=dg(7:8,nR);  HelEAASr(nR)=dg(9,nR)
            end if
            if ( lHemiCalc ) then
               hemi_ekin_r(nR,:)=dg(10:11,nR);  hemi_vrabs_r(nR,:)=dg(12:13,nR)
               if ( l_mag ) then
                  hemi_emag_r(nR,:)=dg(14:15,nR);  hemi_brabs_r(nR,:)=dg(16:17,nR)
               end if
            end if
            if ( lPowerCalc ) viscASr(nR)=dg(18,nR)
            if ( lPerpParCalc ) then
               EperpASr(nR)=dg(19,nR);  EparASr(nR)=dg(20,nR);  EperpaxiASr(nR)=dg(21,nR);  EparaxiASr(nR)=dg(22,nR)
            end if
            if ( lFluxProfCalc ) then
               fkinASr(nR)=dg(23,nR);  fviscASr(nR)=dg(26,nR)
               if ( l_anelastic_liquid ) then                                       ! outPar.f90:511-517
                  fconvASr(nR)=dg(24,nR)
               else
                  fconvASr(nR)=temp0(nR)*dg(24,nR)+ViscHeatFac*ThExpNb*alpha0(nR)*temp0(nR)*orho1(nR)*dg(25,nR)
               end if
               if ( l_mag_nl ) then
                  fresASr(nR)=dg(27,nR);  fpoynASr(nR)=dg(28,nR)
               end if
            end if
            if ( lViscBcCalc ) then
               uhASr(nR)=dg(29,nR);  duhASr(nR)=dg(30,nR);  gradT2ASr(nR)=dg(31,nR)
            end if
            if ( lPhaseCalc ) then                                                  ! outMisc.f90:1217-1219
               ekinSr(nR)=dg(33,nR);  ekinLr(nR)=dg(34,nR);  volSr(nR)=dg(35,nR)
            end if
         end do
         deallocate( dg )
         !-- outPhase locates the melting radius of every (theta,phi) column from the grid values of phi and s that
         !   get_ekin_solid_liquid keeps (outMisc.f90:1210-1212): two scalar syntheses per level through module sht
         if ( lPhaseCalc ) then
            allocate( grd(nlat_padded,n_phi_max) )
            do nR=nRstart,nRstop
               call scal_to_spat(phi_Rloc(:,nR), grd, l_R(nR))
               phase_Rloc(:,:,nR)=grd(1:n_theta_max,:)
               call scal_to_spat(s_Rloc(:,nR), grd, l_R(nR))
               temp_Rloc(:,:,nR)=grd(1:n_theta_max,:)
               if ( l_dtphaseMovie ) then
                  call scal_to_spat(ds_Rloc(:,nR), grd, l_R(nR))
                  dtemp_Rloc(:,:,nR)=grd(1:n_theta_max,:)
               end if
            end do
            deallocate( grd )
         end if
      end if

   contains

      subroutine rms_on_device()
         !-- l_RMS (rIter.f90:215-252, 433-435, 710): on lRmsCalc steps one more batch returns the fourteen spectra of
         !   transform_to_lm_RMS for all levels (every level treated as bulk, as rIter.f90:215 does); they go into RMS's module
         !   arrays level by level and compute_lm_forces -- the reference's own spectral sums -- runs on them.  get_nl_RMS keeps
         !   the previous step's velocity on the grid at every stage-1 call; here its potentials stay on the device
         !   (magic_rloop_rms_keep)
         if ( l_rms_dev ) then
            if ( lRmsCalc ) then
               if ( .not. allocated(rq) ) then
                  allocate( rq(lm_max,nRstart:nRstop,MAGIC_NRMS) )
                  call magic_check( magic_rloop_pin_host(this%rl, c_loc(rq), int(16,c_size_t)*size(rq,kind=c_size_t)), &
                       &            'magic_rloop_pin_host' )
               end if
               fin%p = c_loc(p_Rloc)
               call magic_check( magic_rloop_rms(this%rl, fin, real(tscheme%dt(1),c_double), rq), 'magic_rloop_rms' )
               do nR=nRstart,nRstop
                  LFrLM(:)  =rq(:,nR,2);   dtVrLM(:) =rq(:,nR,3)
                  if ( l_adv_curl ) dpkindrLM(:)=rq(:,nR,4)
                  Advt2LM(:)=rq(:,nR,5);   Advp2LM(:)=rq(:,nR,6);   LFt2LM(:)=rq(:,nR,7);    LFp2LM(:)=rq(:,nR,8)
                  CFt2LM(:) =rq(:,nR,9);   CFp2LM(:) =rq(:,nR,10);  PFt2LM(:)=rq(:,nR,11);   PFp2LM(:)=rq(:,nR,12)
                  dtVtLM(:) =rq(:,nR,13);  dtVpLM(:) =rq(:,nR,14)
                  call compute_lm_forces(nR, rq(:,nR,1))
               end do
            end if
            if ( tscheme%istage == 1 ) call magic_check( magic_rloop_rms_keep(this%rl, fin), 'magic_rloop_rms_keep' )
         end if
      end subroutine rms_on_device

   end subroutine radialLoop
!------------------------------------------------------------------------------
end module rIter_cuda_mod
