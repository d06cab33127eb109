! ps3d_cuda_mod -- iso_c_binding interface to libps3d_cuda.so (include/ps3d_cuda.h), one interface per entry point.
!
! Drop this file and the replacement modules next to it into the reference source tree (see fortran/README.md).
! The image this repository is built in has no Fortran compiler: these files are reviewed source, the automated
! tests drive the identical C ABI through ctypes (ps3d_b200/lib.py) and a compiled C++ driver
! (examples/ps3d_driver.cpp).
!
! Conventions of the ABI (include/ps3d_cuda.h:7-19): every function returns a status (0 = ok), arrays are passed
! as they are declared in the reference -- f(0:nz, lo2:hi2, lo1:hi1), z fastest -- so an assumed-size dummy
! `f(*)` takes the module arrays of fields.f90 without copies.
module ps3d_cuda_mod
    use, intrinsic :: iso_c_binding
    implicit none

    ! enumerators of include/ps3d_cuda.h
    integer(c_int), parameter :: PS3D_FILTER_HOU_LI = 0, PS3D_FILTER_23_RULE = 1
    integer(c_int), parameter :: PS3D_LSCALE_KOLMOGOROV = 0, PS3D_LSCALE_GEOPHYSICAL = 1
    integer(c_int), parameter :: PS3D_STEPPER_CN2 = 0, PS3D_STEPPER_IMPL_RK4 = 1
    integer(c_int), parameter :: PS3D_PRE_CONSTANT = 0, PS3D_PRE_VORCH = 1, PS3D_PRE_BFMAX = 2,      &
                                 PS3D_PRE_ROLL_MEAN_MAX_STRAIN = 3, PS3D_PRE_MAX_STRAIN = 4,         &
                                 PS3D_PRE_US_MAX_STRAIN = 5, PS3D_PRE_ROLL_MEAN_BFMAX = 6
    integer(c_int), parameter :: PS3D_F_SVOR = 0, PS3D_F_VOR = 1, PS3D_F_VEL = 2, PS3D_F_SVEL = 3,   &
                                 PS3D_F_SVORTS = 4, PS3D_F_PRES = 5, PS3D_F_DELTA = 6,               &
                                 PS3D_F_SBUOY = 7, PS3D_F_BUOY = 8, PS3D_F_SBUOYS = 9
    ! slots of diag(16) of ps3d_cuda_adapt / ps3d_cuda_advance, 1-based here
    integer, parameter :: PS3D_D_VORTMAX = 1, PS3D_D_VORTRMS = 2, PS3D_D_VORCH = 3, PS3D_D_VORMEAN_X = 4,  &
                          PS3D_D_VORMEAN_Y = 5, PS3D_D_VORMEAN_Z = 6, PS3D_D_BFMAX = 7, PS3D_D_GGMAX = 8,  &
                          PS3D_D_UMAX = 9, PS3D_D_VMAX = 10, PS3D_D_WMAX = 11, PS3D_D_USGGMAX = 12,        &
                          PS3D_D_LSGGMAX = 13, PS3D_D_RMV = 14, PS3D_D_DT = 15, PS3D_D_PREFACTOR = 16

    interface
        function ps3d_cuda_last_error() bind(C, name='ps3d_cuda_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function

        ! ---- set-up (mpi_layout.f90:53, parameters.f90:61, sta3dfft.f90:53, inversion_utils.f90:124,222,459) ----
        integer(c_int) function ps3d_cuda_init(nx, ny, nz, lower, extent, rank, nranks, nccl_id) &
                bind(C, name='ps3d_cuda_init')
            import :: c_int, c_double, c_ptr
            integer(c_int), value :: nx, ny, nz, rank, nranks
            real(c_double), intent(in) :: lower(3), extent(3)
            type(c_ptr), value :: nccl_id            ! 128-byte ncclUniqueId; c_null_ptr: one rank or MPI transport
        end function
        integer(c_int) function ps3d_cuda_init_inversion(filtering_id) bind(C, name='ps3d_cuda_init_inversion')
            import :: c_int
            integer(c_int), value :: filtering_id
        end function
        integer(c_int) function ps3d_cuda_init_diffusion(nnu, prediss, lscale, te, en, nu) &
                bind(C, name='ps3d_cuda_init_diffusion')
            import :: c_int, c_double
            integer(c_int), value :: nnu, lscale
            real(c_double), value :: prediss, te, en
            real(c_double), intent(out) :: nu
        end function
        integer(c_int) function ps3d_cuda_finalise() bind(C, name='ps3d_cuda_finalise')
            import :: c_int
        end function

        ! ---- operator mode (sta3dfft.f90:136-377, inversion_utils.f90:549-719) ----
        integer(c_int) function ps3d_cuda_fftxyp2s(fp, fs) bind(C, name='ps3d_cuda_fftxyp2s')
            import :: c_int, c_double
            real(c_double), intent(in)  :: fp(*)
            real(c_double), intent(out) :: fs(*)
        end function
        integer(c_int) function ps3d_cuda_fftxys2p(fs, fp) bind(C, name='ps3d_cuda_fftxys2p')
            import :: c_int, c_double
            real(c_double), intent(in)  :: fs(*)
            real(c_double), intent(out) :: fp(*)
        end function
        integer(c_int) function ps3d_cuda_fftsine(fs) bind(C, name='ps3d_cuda_fftsine')
            import :: c_int, c_double
            real(c_double), intent(inout) :: fs(*)
        end function
        integer(c_int) function ps3d_cuda_fftcosine(fs) bind(C, name='ps3d_cuda_fftcosine')
            import :: c_int, c_double
            real(c_double), intent(inout) :: fs(*)
        end function
        integer(c_int) function ps3d_cuda_diffx(fs, ds) bind(C, name='ps3d_cuda_diffx')
            import :: c_int, c_double
            real(c_double), intent(in)  :: fs(*)
            real(c_double), intent(out) :: ds(*)
        end function
        integer(c_int) function ps3d_cuda_diffy(fs, ds) bind(C, name='ps3d_cuda_diffy')
            import :: c_int, c_double
            real(c_double), intent(in)  :: fs(*)
            real(c_double), intent(out) :: ds(*)
        end function
        integer(c_int) function ps3d_cuda_central_diffz(fs, ds) bind(C, name='ps3d_cuda_central_diffz')
            import :: c_int, c_double
            real(c_double), intent(in)  :: fs(*)
            real(c_double), intent(out) :: ds(*)
        end function
        integer(c_int) function ps3d_cuda_diffz(fs, ds) bind(C, name='ps3d_cuda_diffz')
            import :: c_int, c_double
            real(c_double), intent(in)  :: fs(*)
            real(c_double), intent(out) :: ds(*)
        end function
        integer(c_int) function ps3d_cuda_field_combine_semi_spectral(sf) &
                bind(C, name='ps3d_cuda_field_combine_semi_spectral')
            import :: c_int, c_double
            real(c_double), intent(inout) :: sf(*)
        end function
        integer(c_int) function ps3d_cuda_field_decompose_semi_spectral(sfc) &
                bind(C, name='ps3d_cuda_field_decompose_semi_spectral')
            import :: c_int, c_double
            real(c_double), intent(inout) :: sfc(*)
        end function
        integer(c_int) function ps3d_cuda_field_combine_physical(sf, fc) bind(C, name='ps3d_cuda_field_combine_physical')
            import :: c_int, c_double
            real(c_double), intent(in)  :: sf(*)
            real(c_double), intent(out) :: fc(*)
        end function
        integer(c_int) function ps3d_cuda_field_decompose_physical(fc, sf) &
                bind(C, name='ps3d_cuda_field_decompose_physical')
            import :: c_int, c_double
            real(c_double), intent(in)  :: fc(*)
            real(c_double), intent(out) :: sf(*)
        end function

        ! ---- resident mode: the time loop (utils.f90:136-184, inversion.f90:23,378, advance.f90:77-410,
        !      cn2.f90:40-181, impl_rk4.f90:37-207) ----
        integer(c_int) function ps3d_cuda_upload_vorticity(vor) bind(C, name='ps3d_cuda_upload_vorticity')
            import :: c_int, c_double
            real(c_double), intent(in) :: vor(*)     ! vor(0:nz, lo2:hi2, lo1:hi1, 3) of fields.f90:62
        end function
        integer(c_int) function ps3d_cuda_upload_vorticity_begin(vor) bind(C, name='ps3d_cuda_upload_vorticity_begin')
            import :: c_int, c_double
            real(c_double), intent(in) :: vor(*)     ! asynchronous: keep `vor` untouched until ..._end
        end function
        integer(c_int) function ps3d_cuda_upload_vorticity_end() bind(C, name='ps3d_cuda_upload_vorticity_end')
            import :: c_int
        end function
        integer(c_int) function ps3d_cuda_vor2vel() bind(C, name='ps3d_cuda_vor2vel')
            import :: c_int
        end function
        integer(c_int) function ps3d_cuda_source() bind(C, name='ps3d_cuda_source')
            import :: c_int
        end function
        integer(c_int) function ps3d_cuda_adapt(t, t_limit, alpha, pretype_id, win, dt, diag) &
                bind(C, name='ps3d_cuda_adapt')
            import :: c_int, c_double
            real(c_double), value :: t, t_limit, alpha
            integer(c_int), value :: pretype_id, win
            real(c_double), intent(out) :: dt, diag(16)
        end function
        integer(c_int) function ps3d_cuda_stepper_setup(stepper_id) bind(C, name='ps3d_cuda_stepper_setup')
            import :: c_int
            integer(c_int), value :: stepper_id
        end function
        integer(c_int) function ps3d_cuda_set_diffusion(dt, prefactor) bind(C, name='ps3d_cuda_set_diffusion')
            import :: c_int, c_double
            real(c_double), value :: dt, prefactor
        end function
        integer(c_int) function ps3d_cuda_step(t, dt) bind(C, name='ps3d_cuda_step')
            import :: c_int, c_double
            real(c_double), intent(inout) :: t
            real(c_double), value :: dt
        end function
        integer(c_int) function ps3d_cuda_advance(t, t_limit, alpha, pretype_id, win, dt, diag) &
                bind(C, name='ps3d_cuda_advance')
            import :: c_int, c_double
            real(c_double), intent(inout) :: t
            real(c_double), value :: t_limit, alpha
            integer(c_int), value :: pretype_id, win
            real(c_double), intent(out) :: dt, diag(16)
        end function
        integer(c_int) function ps3d_cuda_download(field_id, comp, host) bind(C, name='ps3d_cuda_download')
            import :: c_int, c_double
            integer(c_int), value :: field_id, comp       ! comp = 0..2
            real(c_double), intent(out) :: host(*)
        end function
        integer(c_int) function ps3d_cuda_upload(field_id, comp, host) bind(C, name='ps3d_cuda_upload')
            import :: c_int, c_double
            integer(c_int), value :: field_id, comp
            real(c_double), intent(in) :: host(*)
        end function
        integer(c_int) function ps3d_cuda_diagnostics(out) bind(C, name='ps3d_cuda_diagnostics')
            import :: c_int, c_double
            real(c_double), intent(out) :: out(8)
        end function
        integer(c_int) function ps3d_cuda_field_stats(out) bind(C, name='ps3d_cuda_field_stats')
            import :: c_int, c_double
            real(c_double), intent(out) :: out(40)   ! out(NC_KE) ... out(NC_ROMAX), field_diagnostics_netcdf.f90:36-75
        end function
        integer(c_int) function ps3d_cuda_genspec(nmax, spec, num, nbins, dk) bind(C, name='ps3d_cuda_genspec')
            import :: c_int, c_double
            integer(c_int), value :: nmax
            real(c_double), intent(out) :: spec(*), num(*)
            integer(c_int), intent(out) :: nbins
            real(c_double), intent(out) :: dk
        end function

        ! ---- physics.f90 and the ENABLE_BUOYANCY build ----
        integer(c_int) function ps3d_cuda_set_physics(f_cor, bfsq) bind(C, name='ps3d_cuda_set_physics')
            import :: c_int, c_double
            real(c_double), intent(in) :: f_cor(3)
            real(c_double), value :: bfsq
        end function
        integer(c_int) function ps3d_cuda_enable_buoyancy() bind(C, name='ps3d_cuda_enable_buoyancy')
            import :: c_int
        end function
        integer(c_int) function ps3d_cuda_upload_buoyancy(buoy) bind(C, name='ps3d_cuda_upload_buoyancy')
            import :: c_int, c_double
            real(c_double), intent(in) :: buoy(*)    ! b'(0:nz, lo2:hi2, lo1:hi1), basic state removed (utils.f90:152-157)
        end function
        integer(c_int) function ps3d_cuda_init_diffusion_buoyancy(nnu, prediss, lscale, te, en, pretype_id, win, nu) &
                bind(C, name='ps3d_cuda_init_diffusion_buoyancy')
            import :: c_int, c_double
            integer(c_int), value :: nnu, lscale, pretype_id, win
            real(c_double), value :: prediss, te, en
            real(c_double), intent(out) :: nu
        end function
        integer(c_int) function ps3d_cuda_set_diffusion_buoyancy(dt, bf) bind(C, name='ps3d_cuda_set_diffusion_buoyancy')
            import :: c_int, c_double
            real(c_double), value :: dt, bf
        end function
        integer(c_int) function ps3d_cuda_buoyancy_diag(out) bind(C, name='ps3d_cuda_buoyancy_diag')
            import :: c_int, c_double
            real(c_double), intent(out) :: out(4)    ! bfmax, rmb, bval, bvisc
        end function

        ! ---- transport (fft_pencil.f90:283-330 / mpi_reverse.f90 replaced by one slab all-to-all per 2-D FFT) ----
        integer(c_int) function ps3d_cuda_set_transport(alltoall, allreduce, user) bind(C, name='ps3d_cuda_set_transport')
            import :: c_int, c_funptr, c_ptr
            type(c_funptr), value :: alltoall, allreduce
            type(c_ptr), value :: user
        end function
        integer(c_int) function ps3d_cuda_comm_stats(n_alltoall, bytes_sent) bind(C, name='ps3d_cuda_comm_stats')
            import :: c_int, c_long_long, c_double
            integer(c_long_long), intent(out) :: n_alltoall
            real(c_double), intent(out) :: bytes_sent
        end function

        ! ---- introspection ----
        integer(c_long_long) function ps3d_cuda_kernel_launches() bind(C, name='ps3d_cuda_kernel_launches')
            import :: c_long_long
        end function
        integer(c_long_long) function ps3d_cuda_tma_launches() bind(C, name='ps3d_cuda_tma_launches')
            import :: c_long_long
        end function
        real(c_double) function ps3d_cuda_last_advance_ms() bind(C, name='ps3d_cuda_last_advance_ms')
            import :: c_double
        end function
        integer(c_int) function ps3d_cuda_time_kernel(which, reps, ms) bind(C, name='ps3d_cuda_time_kernel')
            import :: c_int, c_double
            integer(c_int), value :: which, reps
            real(c_double), intent(out) :: ms
        end function
    end interface

contains

    ! status -> mpi_stop with the library's message (the reference's error path, mpi_utils.f90:15-31)
    subroutine ps3d_cuda_check(ierr, what)
        use mpi_utils, only : mpi_stop
        integer(c_int),   intent(in) :: ierr
        character(len=*), intent(in) :: what
        character(kind=c_char), pointer :: cmsg(:)
        character(len=1024) :: msg
        type(c_ptr) :: p
        integer :: i
        if (ierr == 0) return
        msg = ''
        p = ps3d_cuda_last_error()
        if (c_associated(p)) then
            call c_f_pointer(p, cmsg, [1024])
            do i = 1, 1024
                if (cmsg(i) == c_null_char) exit
                msg(i:i) = cmsg(i)
            enddo
        endif
        call mpi_stop("ps3d_cuda " // what // ": " // trim(msg))
    end subroutine ps3d_cuda_check

    ! option strings of options.f90 -> enumerators
    pure function filtering_id(filtering) result(id)
        character(len=*), intent(in) :: filtering
        integer(c_int) :: id
        id = merge(PS3D_FILTER_23_RULE, PS3D_FILTER_HOU_LI, trim(filtering) == '2/3-rule')
    end function
    pure function length_scale_id(lscale) result(id)
        character(len=*), intent(in) :: lscale
        integer(c_int) :: id
        id = merge(PS3D_LSCALE_GEOPHYSICAL, PS3D_LSCALE_KOLMOGOROV, trim(lscale) == 'geophysical')
    end function
    function pretype_id(pretype) result(id)          ! advance.f90:385-408
        character(len=*), intent(in) :: pretype
        integer(c_int) :: id
        select case (trim(pretype))
            case ('constant');             id = PS3D_PRE_CONSTANT
            case ('vorch');                id = PS3D_PRE_VORCH
            case ('bfmax');                id = PS3D_PRE_BFMAX
            case ('roll-mean-max-strain'); id = PS3D_PRE_ROLL_MEAN_MAX_STRAIN
            case ('roll-mean-bfmax');      id = PS3D_PRE_ROLL_MEAN_BFMAX
            case ('max-strain');           id = PS3D_PRE_MAX_STRAIN
            case ('us-max-strain');        id = PS3D_PRE_US_MAX_STRAIN
            case default;                  id = -1_c_int     ! the library answers with the reference's message
        end select
    end function

end module ps3d_cuda_mod
