module mpi_transp_cuda_mod
   !
   ! type_mpicuda: a fifth extension of type_mpitransp (mpi_transpose.f90:18-54) next to type_mpiatoav / atoaw / atoap
   ! and type_mpiptop.  The redistribution arr_LMloc(llm:ulm,1:n_r_max,1:n_fields) <-> arr_Rloc(1:lm_max,
   ! nRstart:nRstop,1:n_fields) runs on the GPUs: pack with the lo -> st permutation fused in, one grouped NCCL
   ! send/recv over NVLink with the alltoallv counts of create_comm_alltoallv (mpi_transpose.f90:120-152), tiled unpack.
   !
   ! Selected in communications.f90:166-214 by a new value of `mpi_transp`, e.g.
   !      else if ( index(mpi_transp, 'CUDA') /= 0 ) then
   !         allocate( type_mpicuda :: lo2r_s ) ...
   !
   ! The decomposition is the reference's own: getBlocks for the radial levels (parallel.f90:75-92) and the lo_map
   ! ranges llm:ulm (blocking.f90:387-544); create_comm checks that both sides agree.
   !
   use iso_c_binding
   use precision_mod
   use parallel_mod, only: rank, n_procs
   use truncation, only: lm_max, n_r_max
   use radial_data, only: nRstart, nRstop
   use blocking, only: llm, ulm
   use mpi_transp_mod, only: type_mpitransp
   use useful, only: abortRun
   use sht, only: sht_h
   use magic_b200_c
#ifdef WITH_MPI
   use mpi
#endif

   implicit none

   private

   !-- Fused mode (l_fused_lm): MagIC's step is transp_LMloc_to_Rloc -> radialLoopG -> transp_Rloc_to_LMloc
   !   (step_time.f90:485-612).  With the CPU-resident LM loop all three cross PCIe; magic_rloop_run_lm does them in ONE call
   !   on the LM-distributed host containers, chunk-pipelined, without intermediate host R-containers.  step_time.f90 stays
   !   untouched: in fused mode transp_lm2r only RECORDS its request (the loop of the same stage picks the LM containers up
   !   itself), and transp_r2lm returns at once when the loop has already delivered the LM-distributed explicit terms.  On
   !   steps with in-loop diagnostics rIter_cuda_t falls back to the level-at-a-time loop: it first executes the recorded
   !   requests (run_pending_lm2r) so that the R-distributed host arrays exist, and clears l_outputs_in_lm so that the
   !   transposes back are real.
   logical, public :: l_fused_lm = .true.
   logical, public :: l_outputs_in_lm = .false.
   integer, parameter :: n_pending_max = 8
   integer, public :: n_pending = 0
   type(c_ptr) :: pending_t(n_pending_max), pending_lm(n_pending_max), pending_r(n_pending_max)
   type(c_ptr), public :: transp5 = c_null_ptr   ! a transposer that serves containers of up to 5 fields (the loop needs one)

   public :: run_pending_lm2r

   type, public, extends(type_mpitransp) :: type_mpicuda
      type(c_ptr) :: t = c_null_ptr
   contains
      procedure :: create_comm  => create_comm_cuda
      procedure :: destroy_comm => destroy_comm_cuda
      procedure :: transp_lm2r  => transp_lm2r_cuda
      procedure :: transp_r2lm  => transp_r2lm_cuda
   end type type_mpicuda

contains

   subroutine create_comm_cuda(this, n_fields)

      class(type_mpicuda) :: this
      integer, intent(in) :: n_fields

      character(kind=c_char) :: id(128)
      integer(c_int) :: llm_c, ulm_c, nRstart_c, nRstop_c
      integer :: ierr

      this%n_fields = n_fields

      !-- NCCL bootstrap: rank 0 draws the unique id, everybody receives it over MPI
      id(:) = c_null_char
      if ( rank == 0 ) call magic_check( magic_transp_unique_id(id), 'magic_transp_unique_id' )
#ifdef WITH_MPI
      call MPI_Bcast(id, 128, MPI_BYTE, 0, MPI_COMM_WORLD, ierr)
#endif
      call magic_check( magic_transp_create(sht_h, id, int(rank,c_int), int(n_procs,c_int), int(n_r_max,c_int), &
           &            int(n_fields,c_int), this%t), 'magic_transp_create' )

      !-- both sides must cut the problem the same way
      call magic_check( magic_transp_extents(this%t, llm_c, ulm_c, nRstart_c, nRstop_c), 'magic_transp_extents' )
      if ( llm_c /= llm .or. ulm_c /= ulm .or. nRstart_c /= nRstart .or. nRstop_c /= nRstop ) then
         call abortRun('! type_mpicuda: the library and MagIC disagree on llm:ulm / nRstart:nRstop')
      end if
      if ( n_fields >= 5 .and. .not. c_associated(transp5) ) transp5 = this%t

   end subroutine create_comm_cuda
!------------------------------------------------------------------------------
   subroutine destroy_comm_cuda(this)

      class(type_mpicuda) :: this

      if ( c_associated(this%t) ) call magic_check( magic_transp_destroy(this%t), 'magic_transp_destroy' )
      this%t = c_null_ptr

   end subroutine destroy_comm_cuda
!------------------------------------------------------------------------------
   subroutine transp_lm2r_cuda(this, arr_LMloc, arr_Rloc)

      class(type_mpicuda) :: this
      complex(cp), intent(in)  :: arr_LMloc(llm:ulm,1:n_r_max,*)
      complex(cp), intent(out) :: arr_Rloc(1:lm_max,nRstart:nRstop,*)

      if ( l_fused_lm ) then   ! record only: the radial loop of this stage either fuses it or executes it
         if ( n_pending == n_pending_max ) call abortRun('! type_mpicuda: too many pending transposes')
         n_pending = n_pending+1
         pending_t(n_pending)  = this%t
         pending_lm(n_pending) = addr_z(arr_LMloc)   ! (the dummies of the deferred interface are no targets: magic_b200_c)
         pending_r(n_pending)  = addr_z(arr_Rloc)
         return
      end if
      call magic_check( magic_transp_lm2r(this%t, arr_LMloc, arr_Rloc), 'magic_transp_lm2r' )

   end subroutine transp_lm2r_cuda
!------------------------------------------------------------------------------
   subroutine transp_r2lm_cuda(this, arr_Rloc, arr_LMloc)

      class(type_mpicuda) :: this
      complex(cp), intent(in)  :: arr_Rloc(1:lm_max,nRstart:nRstop,*)
      complex(cp), intent(out) :: arr_LMloc(llm:ulm,1:n_r_max,*)

      if ( l_fused_lm .and. l_outputs_in_lm ) return   ! magic_rloop_run_lm has written arr_LMloc already
      call magic_check( magic_transp_r2lm(this%t, arr_Rloc, arr_LMloc), 'magic_transp_r2lm' )

   end subroutine transp_r2lm_cuda
!------------------------------------------------------------------------------
   subroutine run_pending_lm2r()
      !
      ! Executes the transposes recorded in fused mode (host pointers: H2D, exchange, D2H), e.g. before a diagnostics step
      ! that needs the R-distributed host arrays.
      !
      integer :: n
      complex(cp), pointer :: a_lm(:), a_r(:)

      do n=1,n_pending
         call c_f_pointer(pending_lm(n), a_lm, [1])
         call c_f_pointer(pending_r(n), a_r, [1])
         call magic_check( magic_transp_lm2r(pending_t(n), a_lm, a_r), 'magic_transp_lm2r (pending)' )
      end do
      n_pending = 0

   end subroutine run_pending_lm2r
!------------------------------------------------------------------------------
end module mpi_transp_cuda_mod
