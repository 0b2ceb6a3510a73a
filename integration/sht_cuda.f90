module sht
   !
   ! Third flavour of `module sht` (next to sht_native.f90 and shtns.f90): every transform runs in libmagic_b200.so
   ! on the GPU of this MPI rank.  Same public list and dummy arguments as sht_native.f90:16-20 / shtns.f90:22-26, so
   ! no caller changes (rIter.f90, RMS.f90, TO.f90, dtB.f90, fields_average.f90, init_fields.f90, nonlinear_bcs.f90,
   ! outGeos.f90, out_dtB_frame.f90, out_graph_file.f90, out_movie_file.f90, store_movie_IC.f90).
   !
   ! These per-call entry points move one level's arrays over PCIe per call: right for the diagnostics callers, too
   ! slow for the radial loop, which uses rIter_cuda_t (rIter_cuda.f90) instead.
   !
   ! Assumed-shape actuals are contiguous here as in the reference, and are forwarded as the address of their first
   ! element exactly like shtns.f90:112,128,188.  Grid inputs of the analyses are intent(inout) in the interface
   ! (backends may clobber them, fft_fftw.f90:124-137); this backend leaves them untouched.
   !
   use iso_c_binding
   use precision_mod, only: cp
   use truncation, only: l_max, m_max, m_min, minc, n_theta_max, n_phi_max, nlat_padded
   use useful, only: abortRun
   use parallel_mod, only: rank
   use magic_b200_c

   implicit none

   private

   type(c_ptr), public :: sht_h = c_null_ptr    ! one handle per rank <-> one GPU (shtns.f90:28 `sht_l`)

   public :: initialize_sht, finalize_sht, scal_to_spat, scal_to_grad_spat, pol_to_grad_spat,       &
   &         torpol_to_spat, sphtor_to_spat, torpol_to_curl_spat_IC, torpol_to_spat_IC,             &
   &         torpol_to_dphspat, pol_to_curlr_spat, torpol_to_curl_spat, scal_to_SH, spat_to_qst,    &
   &         spat_to_sphertor, axi_to_spat, toraxi_to_spat

contains

   subroutine initialize_sht(l_scrambled_theta)
      !
      ! sht_native.f90:24-33 / shtns.f90:32-100.  The backend keeps truncation::nlat_padded = n_theta_max and reports
      ! N/S-interleaved theta rows, like the native backend (consumed by horizontal.f90:145-191).
      !
      logical, intent(out) :: l_scrambled_theta

      integer(c_int) :: scr, n_dev

      !-- the library builds lm_max, st_map and lo_map from m = 0 (blocking.f90 loops from m_min): refuse anything else loudly
      if ( m_min /= 0 ) call abortRun('! sht (magic_b200): m_min /= 0 is not supported by the GPU backend')
      n_dev = magic_device_count()
      if ( n_dev < 1 ) call magic_check(1_c_int, 'initialize_sht (no CUDA device; this backend has no CPU path)')
      !-- one rank <-> one GPU of its node
      call magic_check( magic_sht_create(int(l_max,c_int), int(m_max,c_int), int(minc,c_int),            &
           &            int(n_theta_max,c_int), int(n_phi_max,c_int), int(nlat_padded,c_int),             &
           &            int(mod(rank,n_dev),c_int), scr, sht_h), 'magic_sht_create' )
      l_scrambled_theta = ( scr /= 0 )

   end subroutine initialize_sht
!------------------------------------------------------------------------------
   subroutine finalize_sht()

      if ( c_associated(sht_h) ) call magic_check( magic_sht_destroy(sht_h), 'magic_sht_destroy' )
      sht_h = c_null_ptr

   end subroutine finalize_sht
!------------------------------------------------------------------------------
   subroutine scal_to_spat(Slm, fieldc, lcut)

      complex(cp), intent(in) :: Slm(:)
      integer,     intent(in) :: lcut
      real(cp),    intent(out) :: fieldc(:,:)

      call magic_check( magic_scal_to_spat(sht_h, Slm, fieldc, int(lcut,c_int)), 'scal_to_spat' )

   end subroutine scal_to_spat
!------------------------------------------------------------------------------
   subroutine scal_to_grad_spat(Slm, gradtc, gradpc, lcut)

      complex(cp), intent(in) :: Slm(:)
      integer,     intent(in) :: lcut
      real(cp),    intent(out) :: gradtc(:,:)
      real(cp),    intent(out) :: gradpc(:,:)

      call magic_check( magic_scal_to_grad_spat(sht_h, Slm, gradtc, gradpc, int(lcut,c_int)), 'scal_to_grad_spat' )

   end subroutine scal_to_grad_spat
!------------------------------------------------------------------------------
   subroutine pol_to_grad_spat(Slm, gradtc, gradpc, lcut)

      complex(cp), intent(in) :: Slm(:)
      integer,     intent(in) :: lcut
      real(cp),    intent(out) :: gradtc(:,:)
      real(cp),    intent(out) :: gradpc(:,:)

      call magic_check( magic_pol_to_grad_spat(sht_h, Slm, gradtc, gradpc, int(lcut,c_int)), 'pol_to_grad_spat' )

   end subroutine pol_to_grad_spat
!------------------------------------------------------------------------------
   subroutine torpol_to_spat(Wlm, dWlm, Zlm, vrc, vtc, vpc, lcut)

      complex(cp), intent(in) :: Wlm(:), dWlm(:), Zlm(:)
      integer,     intent(in) :: lcut
      real(cp),    intent(out) :: vrc(:,:)
      real(cp),    intent(out) :: vtc(:,:)
      real(cp),    intent(out) :: vpc(:,:)

      call magic_check( magic_torpol_to_spat(sht_h, Wlm, dWlm, Zlm, vrc, vtc, vpc, int(lcut,c_int)), 'torpol_to_spat' )

   end subroutine torpol_to_spat
!------------------------------------------------------------------------------
   subroutine sphtor_to_spat(dWlm, Zlm, vtc, vpc, lcut)

      complex(cp), intent(in) :: dWlm(:), Zlm(:)
      integer,     intent(in) :: lcut
      real(cp),    intent(out) :: vtc(:,:)
      real(cp),    intent(out) :: vpc(:,:)

      call magic_check( magic_sphtor_to_spat(sht_h, dWlm, Zlm, vtc, vpc, int(lcut,c_int)), 'sphtor_to_spat' )

   end subroutine sphtor_to_spat
!------------------------------------------------------------------------------
   subroutine torpol_to_curl_spat_IC(r, r_ICB, dBlm, ddBlm, Jlm, dJlm, cbr, cbt, cbp)

      real(cp),    intent(in) :: r, r_ICB
      complex(cp), intent(in) :: dBlm(:), ddBlm(:)
      complex(cp), intent(in) :: Jlm(:), dJlm(:)
      real(cp),    intent(out) :: cbr(:,:)
      real(cp),    intent(out) :: cbt(:,:)
      real(cp),    intent(out) :: cbp(:,:)

      call magic_check( magic_torpol_to_curl_spat_IC(sht_h, r, r_ICB, dBlm, ddBlm, Jlm, dJlm, cbr, cbt, cbp), &
           &            'torpol_to_curl_spat_IC' )

   end subroutine torpol_to_curl_spat_IC
!------------------------------------------------------------------------------
   subroutine torpol_to_spat_IC(r, r_ICB, Wlm, dWlm, Zlm, Br, Bt, Bp)

      real(cp),    intent(in) :: r, r_ICB
      complex(cp), intent(in) :: Wlm(:), dWlm(:), Zlm(:)
      real(cp),    intent(out) :: Br(:,:)
      real(cp),    intent(out) :: Bt(:,:)
      real(cp),    intent(out) :: Bp(:,:)

      call magic_check( magic_torpol_to_spat_IC(sht_h, r, r_ICB, Wlm, dWlm, Zlm, Br, Bt, Bp), 'torpol_to_spat_IC' )

   end subroutine torpol_to_spat_IC
!------------------------------------------------------------------------------
   subroutine torpol_to_dphspat(dWlm, Zlm, dvtdp, dvpdp, lcut)

      complex(cp), intent(in) :: dWlm(:), Zlm(:)
      integer,     intent(in) :: lcut
      real(cp),    intent(out) :: dvtdp(:,:)
      real(cp),    intent(out) :: dvpdp(:,:)

      call magic_check( magic_torpol_to_dphspat(sht_h, dWlm, Zlm, dvtdp, dvpdp, int(lcut,c_int)), 'torpol_to_dphspat' )

   end subroutine torpol_to_dphspat
!------------------------------------------------------------------------------
   subroutine pol_to_curlr_spat(Qlm, cvrc, lcut)

      complex(cp), intent(in) :: Qlm(:)
      integer,     intent(in) :: lcut
      real(cp),    intent(out) :: cvrc(:,:)

      call magic_check( magic_pol_to_curlr_spat(sht_h, Qlm, cvrc, int(lcut,c_int)), 'pol_to_curlr_spat' )

   end subroutine pol_to_curlr_spat
!------------------------------------------------------------------------------
   subroutine torpol_to_curl_spat(or2, Blm, ddBlm, Jlm, dJlm, cvrc, cvtc, cvpc, lcut)

      complex(cp), intent(in) :: Blm(:), ddBlm(:)
      complex(cp), intent(in) :: Jlm(:), dJlm(:)
      real(cp),    intent(in) :: or2
      integer,     intent(in) :: lcut
      real(cp),    intent(out) :: cvrc(:,:)
      real(cp),    intent(out) :: cvtc(:,:)
      real(cp),    intent(out) :: cvpc(:,:)

      call magic_check( magic_torpol_to_curl_spat(sht_h, or2, Blm, ddBlm, Jlm, dJlm, cvrc, cvtc, cvpc, &
           &            int(lcut,c_int)), 'torpol_to_curl_spat' )

   end subroutine torpol_to_curl_spat
!------------------------------------------------------------------------------
   subroutine scal_to_SH(f, fLM, lcut)

      real(cp),    intent(inout) :: f(:,:)
      integer,     intent(in) :: lcut
      complex(cp), intent(out) :: fLM(:)

      call magic_check( magic_scal_to_SH(sht_h, f, fLM, int(lcut,c_int)), 'scal_to_SH' )

   end subroutine scal_to_SH
!------------------------------------------------------------------------------
   subroutine spat_to_qst(f, g, h, qLM, sLM, tLM, lcut)

      real(cp),    intent(inout) :: f(:,:)
      real(cp),    intent(inout) :: g(:,:)
      real(cp),    intent(inout) :: h(:,:)
      integer,     intent(in) :: lcut
      complex(cp), intent(out) :: qLM(:)
      complex(cp), intent(out) :: sLM(:)
      complex(cp), intent(out) :: tLM(:)

      call magic_check( magic_spat_to_qst(sht_h, f, g, h, qLM, sLM, tLM, int(lcut,c_int)), 'spat_to_qst' )

   end subroutine spat_to_qst
!------------------------------------------------------------------------------
   subroutine spat_to_sphertor(f, g, fLM, gLM, lcut)

      real(cp),    intent(inout) :: f(:,:)
      real(cp),    intent(inout) :: g(:,:)
      integer,     intent(in) :: lcut
      complex(cp), intent(out) :: fLM(:)
      complex(cp), intent(out) :: gLM(:)

      call magic_check( magic_spat_to_sphertor(sht_h, f, g, fLM, gLM, int(lcut,c_int)), 'spat_to_sphertor' )

   end subroutine spat_to_sphertor
!------------------------------------------------------------------------------
   subroutine axi_to_spat(fl_ax, f)

      complex(cp), intent(in) :: fl_ax(l_max+1)
      real(cp),    intent(out) :: f(:)

      call magic_check( magic_axi_to_spat(sht_h, fl_ax, f), 'axi_to_spat' )

   end subroutine axi_to_spat
!------------------------------------------------------------------------------
   subroutine toraxi_to_spat(fl_ax, ft, fp, lcut)

      integer,     intent(in) :: lcut
      complex(cp), intent(in) :: fl_ax(l_max+1)
      real(cp),    intent(out) :: ft(:)
      real(cp),    intent(out) :: fp(:)

      call magic_check( magic_toraxi_to_spat(sht_h, fl_ax, ft, fp, int(lcut,c_int)), 'toraxi_to_spat' )

   end subroutine toraxi_to_spat
!------------------------------------------------------------------------------
end module sht
