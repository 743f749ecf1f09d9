!=============================================================================================
! alps_b200_shim.f90 -- iso_c_binding shim that puts libalps_b200.so under the reference's disp().
!
! Written for danielver02/ALPS (src/): add this file to ALPS_SOURCES in src/Makefile.am, link with
!   -L<repo>/alps_b200 -lalps_b200 -lcudart
! and make the three edits listed in INTEGRATION.md section 3.  It binds the C ABI of
! include/alps_b200.h; every procedure cites the entry point it wraps.
!
! NOT COMPILED IN THIS REPOSITORY'S BUILD: the image has no Fortran compiler (SURVEY.md 8c).  The
! interfaces below are kept in step with include/alps_b200.h by hand; tests/test_abi.py checks the
! C side (every declared symbol exported, no CPU fallback).
!
! Conventions: complex(c_double_complex) values and arrays go over as interleaved (re,im) doubles,
! which is what the C side declares; Fortran arrays go over as they are (column-major, species
! index fastest) -- the library re-tiles them on the device.  Only rank 0 calls the library.
!=============================================================================================
module alps_b200_shim
  use iso_c_binding
  implicit none
  private
  public :: b200_setup, b200_set_k, b200_disp, b200_map, b200_finalize, b200_ngpu, b200_join_ranks

  type, bind(c) :: alps_b200_cfg            ! include/alps_b200.h : alps_b200_cfg
     integer(c_int) :: nspec, nperp, npar, ngamma, npparbar
     real(c_double) :: vA, Bessel_zero, Tlim
     integer(c_int) :: positions_principal, n_resonance_interval, kperp_norm, emulate_nproc
     integer(c_int) :: maxfits, maxorder, device, nmax_cap, batch_max, nmax_force, ngpu
  end type alps_b200_cfg

  interface
     integer(c_int) function alps_b200_init(cfg) bind(c)
       import; type(alps_b200_cfg), intent(in) :: cfg
     end function
     subroutine alps_b200_finalize() bind(c)
     end subroutine
     integer(c_int) function alps_b200_set_species(is, ns, qs, ms, relat, usebM, ACmethod, n_fits, &
          fit_type, perp_correction, logfit, poly_kind, poly_order, poly_log_max) bind(c)
       import
       integer(c_int), value :: is, relat, usebM, ACmethod, n_fits, logfit, poly_kind, poly_order
       real(c_double), value :: ns, qs, ms, poly_log_max
       integer(c_int), intent(in) :: fit_type(*)
       real(c_double), intent(in) :: perp_correction(*)
     end function
     integer(c_int) function alps_b200_set_bm_species(is, bM_nmaxs, bM_Bessel_zeros, bM_betas, bM_alphas, &
          bM_pdrifts) bind(c)
       import; integer(c_int), value :: is, bM_nmaxs
       real(c_double), value :: bM_Bessel_zeros, bM_betas, bM_alphas, bM_pdrifts
     end function
     ! optional dummies of a bind(c) procedure are NULL when absent (Fortran 2018 / TS 29113); an unallocated
     ! allocatable actual argument counts as absent
     integer(c_int) function alps_b200_upload(pp, df0, param_fit, poly_fit_coeffs) bind(c)
       import; real(c_double), intent(in) :: pp(*)
       real(c_double), intent(in), optional :: df0(*), param_fit(*), poly_fit_coeffs(*)
     end function
     integer(c_int) function alps_b200_upload_rel(nspec_rel, f0_rel, df0_rel, gamma_rel, pparbar_rel) bind(c)
       import; integer(c_int), value :: nspec_rel
       real(c_double), intent(in) :: f0_rel(*), df0_rel(*), gamma_rel(*), pparbar_rel(*)
     end function
     integer(c_int) function alps_b200_set_k(kperp, kpar, nmax_out) bind(c)
       import; real(c_double), value :: kperp, kpar; integer(c_int) :: nmax_out(*)
     end function
     integer(c_int) function alps_b200_disp(om, D, chi0, chi0_low, wave) bind(c)
       import; complex(c_double_complex), intent(in) :: om
       complex(c_double_complex) :: D
       complex(c_double_complex), optional :: chi0(*), chi0_low(*), wave(*)
     end function
     integer(c_int) function alps_b200_disp_batch(n, om, D, chi0_opt) bind(c)
       import; integer(c_int), value :: n
       complex(c_double_complex), intent(in) :: om(*); complex(c_double_complex) :: D(*)
       complex(c_double_complex), optional :: chi0_opt(*)
     end function
     integer(c_int) function alps_b200_disp_prefetch(n, om) bind(c)
       import; integer(c_int), value :: n
       complex(c_double_complex), intent(in) :: om(*)
     end function
     integer(c_int) function alps_b200_tps_eval(n, gc, pc, w, npts, gx, px, out) bind(c)
       import; integer(c_int), value :: n, npts
       real(c_double), intent(in) :: gc(*), pc(*), w(*), gx(*), px(*)
       real(c_double) :: out(*)
     end function
     integer(c_int) function alps_b200_comm_unique_id(id) bind(c)
       import; character(kind=c_char) :: id(128)
     end function
     integer(c_int) function alps_b200_comm_init(rank, nranks, id) bind(c)
       import; integer(c_int), value :: rank, nranks; character(kind=c_char), intent(in) :: id(128)
     end function
     integer(c_int) function alps_b200_comm_finalize() bind(c)
       import
     end function
     integer(c_int) function alps_b200_set_partition(kind) bind(c)
       import; integer(c_int), value :: kind
     end function
  end interface

contains

  !> After pass_instructions / derivative_f0 / determine_param_fit on rank 0 (src/ALPS.f90:67-89):
  !> replaces pass_distribution.  alps_b200_init + alps_b200_set_species + alps_b200_upload(_rel).
  subroutine b200_setup(nproc)
    use alps_var, only : nspec, nperp, npar, ngamma, npparbar, vA, Bessel_zero, Tlim, positions_principal, &
         n_resonance_interval, kperp_norm, ns, qs, ms, relativistic, usebM, ACmethod, n_fits, fit_type, &
         perp_correction, logfit, poly_kind, poly_order, poly_log_max, pp, df0, param_fit, poly_fit_coeffs, &
         nspec_rel, f0_rel, df0_rel, gamma_rel, pparbar_rel, bMnmaxs, bMBessel_zeros, bMbetas, bMalphas, bMpdrifts
    use alps_io, only : alps_error
    integer, intent(in) :: nproc           !! MPI size: nmax and the summed harmonic range depend on it
    type(alps_b200_cfg) :: cfg
    integer :: is, ierr

    cfg%nspec = nspec; cfg%nperp = nperp; cfg%npar = npar; cfg%ngamma = ngamma; cfg%npparbar = npparbar
    cfg%vA = vA; cfg%Bessel_zero = Bessel_zero; cfg%Tlim = Tlim
    cfg%positions_principal = positions_principal; cfg%n_resonance_interval = n_resonance_interval
    cfg%kperp_norm = merge(1, 0, kperp_norm); cfg%emulate_nproc = nproc
    cfg%maxfits = maxval(n_fits); cfg%maxorder = maxval(poly_order)
    cfg%device = -1; cfg%nmax_cap = 0; cfg%batch_max = 0; cfg%nmax_force = 0
    cfg%ngpu = b200_ngpu()          ! every GPU of the box: the library partitions b200_map's batch itself
    ierr = alps_b200_init(cfg)
    if (ierr /= 0) call alps_error(ierr)
    do is = 1, nspec
       ierr = alps_b200_set_species(is, ns(is), qs(is), ms(is), merge(1, 0, relativistic(is)), &
            merge(1, 0, usebM(is)), ACmethod(is), n_fits(is), fit_type(is, :), perp_correction(is, :), &
            merge(1, 0, logfit(is)), poly_kind(is), poly_order(is), poly_log_max(is))
       if (ierr /= 0) call alps_error(ierr)
       ! either keep calc_chi on the Fortran side (alps_b200_add_external_chi per omega) or hand the
       ! &bM_spec_j values over and let the library compute the NHDS chi on the device:
       if (usebM(is)) ierr = alps_b200_set_bm_species(is, bMnmaxs(is), bMBessel_zeros(is), bMbetas(is), &
            bMalphas(is), bMpdrifts(is))
    enddo
    ierr = alps_b200_upload(pp, df0, param_fit, poly_fit_coeffs)
    if (ierr /= 0) call alps_error(ierr)
    if (nspec_rel > 0) then
       ierr = alps_b200_upload_rel(nspec_rel, f0_rel, df0_rel, gamma_rel, pparbar_rel)
       if (ierr /= 0) call alps_error(ierr)
    endif
  end subroutine b200_setup

  !> Wherever the reference calls determine_nmax; split_processes; determine_bessel_array
  !> (src/ALPS.f90:93-96, src/ALPS_fns.f90:2465-2472, 3225-3232, 3374-3379).
  subroutine b200_set_k()
    use alps_var, only : kperp, kpar, nmax
    use alps_io, only : alps_error
    integer :: ierr
    ierr = alps_b200_set_k(kperp, kpar, nmax)
    if (ierr /= 0) call alps_error(ierr)
  end subroutine b200_set_k

  !> Body of disp(om) on rank 0 (src/ALPS_fns.f90:333-632): D, and the globals calc_eigen reads.
  !> want_aux = .false. in the root finders (D only: graph replay, memo), .true. where calc_eigen follows.
  double complex function b200_disp(om, want_aux)
    use alps_var, only : chi0, chi0_low, wave
    use alps_io, only : alps_error
    double complex, intent(in) :: om
    logical, intent(in) :: want_aux
    complex(c_double_complex) :: D
    integer :: ierr
    if (want_aux) then
       ierr = alps_b200_disp(om, D, chi0, chi0_low, wave)
    else
       ierr = alps_b200_disp(om, D)
    endif
    if (ierr /= 0) call alps_error(ierr)
    b200_disp = D
  end function b200_disp

  !> The nr x ni loop of map_search (src/ALPS_fns.f90:3697-3757) as one batch: cal(ir,ii) = disp(om(ir,ii)).
  subroutine b200_map(n, om, cal)
    use alps_io, only : alps_error
    integer, intent(in) :: n
    double complex, intent(in) :: om(n)
    double complex, intent(out) :: cal(n)
    integer :: ierr
    ierr = alps_b200_disp_batch(n, om, cal)
    if (ierr /= 0) call alps_error(ierr)
  end subroutine b200_map

  subroutine b200_finalize()
    call alps_b200_finalize()
  end subroutine b200_finalize

  !> One GPU per MPI rank (the reference's own SPMD layout: every rank runs the same driver and calls disp collectively,
  !> src/ALPS.f90:107-133).  Call after b200_setup on EVERY rank, with cfg%device = the rank's local GPU (set
  !> ALPS_B200_DEVICE or edit b200_setup) and cfg%ngpu = 1: rank 0 draws the NCCL id, MPI broadcasts its 128 bytes, every
  !> rank joins the library's communicator.  From then on b200_map (alps_b200_disp_batch) is collective -- every rank
  !> evaluates a slice, one ncclAllGather inside the library, every rank gets all of cal -- while single b200_disp calls
  !> are evaluated by every rank for itself (bitwise the same D everywhere, no exchange).  Replaces split_processes and
  !> the MPI_REDUCE pair of disp (src/ALPS_fns.f90:4079-4207, 519-523).
  subroutine b200_join_ranks()
    use alps_var, only : iproc, nproc
    use alps_io, only : alps_error
    use mpi
    character(kind=c_char) :: id(128)
    integer :: ierr, ierror
    id = c_null_char
    if (iproc == 0) then
       ierr = alps_b200_comm_unique_id(id)
       if (ierr /= 0) call alps_error(ierr)
    endif
    call mpi_bcast(id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierror)
    ierr = alps_b200_comm_init(iproc, nproc, id)
    if (ierr /= 0) call alps_error(ierr)
  end subroutine b200_join_ranks

  !> Devices the library should drive from this (single) calling rank: environment variable ALPS_B200_NGPU,
  !> default 1.  With ngpu > 1 nothing else changes on the Fortran side: b200_map's batch is cut into one slice
  !> per GPU inside alps_b200_disp_batch (OMEGA partition, include/alps_b200.h).
  integer function b200_ngpu()
    character(len=16) :: v
    integer :: stat, n
    b200_ngpu = 1
    call get_environment_variable('ALPS_B200_NGPU', v, status=stat)
    if (stat == 0) then
       read(v, *, iostat=stat) n
       if (stat == 0 .and. n >= 1) b200_ngpu = n
    endif
  end function b200_ngpu

end module alps_b200_shim
