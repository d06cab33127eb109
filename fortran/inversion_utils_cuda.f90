! Replacement body of module inversion_utils (reference src/inversion/inversion_utils.f90) on top of libps3d_cuda.
! Public procedures keep their names and argument lists (inversion_utils.f90:80-110); the eight N-sized tables
! (green, filt, phim, phip, thetam, thetap, dthetam, dthetap; :263-369, 484-542) are no longer allocated: the
! library recomputes them per column group on the device.  k2l2 / k2l2i / vhdis stay available on the host for the
! diagnostics that read them.
module inversion_utils
    use, intrinsic :: iso_c_binding
    use constants
    use parameters, only : nx, ny, nz, dx, dxi, extent, ncelli, upper, lower
    use mpi_layout
    use sta3dfft, only : initialise_fft, finalise_fft, rkx, rky, rkz, fftxyp2s, fftxys2p, fftsine, fftcosine
    use options, only : vor_visc, filtering
#ifdef ENABLE_BUOYANCY
    use options, only : buoy_visc
#endif
    use mpi_utils, only : mpi_print, mpi_stop
    use ps3d_cuda_mod
    implicit none

    private

    double precision, allocatable :: k2l2i(:, :), k2l2(:, :)
    double precision, allocatable :: vhdis(:, :)
#ifdef ENABLE_BUOYANCY
    double precision, allocatable :: bhdis(:, :)
    double precision :: bvisc
#endif
    double precision :: dzi, hdzi
    double precision :: vvisc
    logical :: is_initialised = .false.

    public :: init_inversion, finalise_inversion, init_diffusion, vvisc, central_diffz, hdzi, k2l2, k2l2i, vhdis, &
              field_combine_semi_spectral, field_combine_physical,                                                &
              field_decompose_semi_spectral, field_decompose_physical
#ifdef ENABLE_BUOYANCY
    public :: diffz, bvisc, bhdis
#endif

contains

    subroutine init_inversion                                                  ! inversion_utils.f90:222-371
        integer :: kx, ky

        if (is_initialised) return
        is_initialised = .true.

        dzi = dxi(3)
        hdzi = f12 * dxi(3)

        call initialise_fft(extent)                                            ! creates the library context
        call ps3d_cuda_check(ps3d_cuda_init_inversion(filtering_id(filtering)), 'init_inversion')
#ifdef ENABLE_BUOYANCY
        call ps3d_cuda_check(ps3d_cuda_enable_buoyancy(), 'enable_buoyancy')
#endif
        allocate(k2l2i(box%lo(2):box%hi(2), box%lo(1):box%hi(1)))
        allocate(k2l2(box%lo(2):box%hi(2), box%lo(1):box%hi(1)))
        do kx = box%lo(1), box%hi(1)                                           ! :240-258
            do ky = box%lo(2), box%hi(2)
                k2l2(ky, kx) = rkx(kx) ** 2 + rky(ky) ** 2
            enddo
        enddo
        if ((box%lo(1) == 0) .and. (box%lo(2) == 0)) then
            k2l2(0, 0) = one
            k2l2i = one / k2l2
            k2l2(0, 0) = zero
            k2l2i(0, 0) = zero
        else
            k2l2i = one / k2l2
        endif
    end subroutine init_inversion

    subroutine finalise_inversion                                              ! inversion_utils.f90:459-480
        if (.not. is_initialised) return
        if (allocated(k2l2)) deallocate(k2l2, k2l2i)
        if (allocated(vhdis)) deallocate(vhdis)
        call finalise_fft
        is_initialised = .false.
    end subroutine finalise_inversion

    ! inversion_utils.f90:124-153.  te, en come from ps3d_cuda_diagnostics (see setup_fields in fortran/README.md)
    subroutine init_diffusion(te, en)
        double precision, intent(in) :: te, en
        real(c_double) :: nu

        if (.not. is_initialised) call mpi_print("Error: Inversion not initialised!")

        call ps3d_cuda_check(ps3d_cuda_init_diffusion(int(vor_visc%nnu, c_int), vor_visc%prediss,             &
                                                      length_scale_id(vor_visc%length_scale), te, en, nu),    &
                             'init_diffusion')
        vvisc = nu
        allocate(vhdis(box%lo(2):box%hi(2), box%lo(1):box%hi(1)))
        if (vor_visc%nnu == 1) then                                            ! init_dissipation (:190-218)
            vhdis = vvisc * k2l2
        else
            vhdis = vvisc * k2l2 ** vor_visc%nnu
        endif
#ifdef ENABLE_BUOYANCY
        call ps3d_cuda_check(ps3d_cuda_init_diffusion_buoyancy(int(buoy_visc%nnu, c_int), buoy_visc%prediss,          &
                                                               length_scale_id(buoy_visc%length_scale), te, en,       &
                                                               pretype_id(buoy_visc%pretype),                         &
                                                               int(buoy_visc%roll_mean_win_size, c_int), nu),         &
                             'init_diffusion_buoyancy')
        bvisc = nu
        allocate(bhdis(box%lo(2):box%hi(2), box%lo(1):box%hi(1)))
        if (buoy_visc%nnu == 1) then
            bhdis = bvisc * k2l2
        else
            bhdis = bvisc * k2l2 ** buoy_visc%nnu
        endif
#endif
    end subroutine init_diffusion

    subroutine field_decompose_physical(fc, sf)                                ! inversion_utils.f90:549-559
        double precision, intent(in)  :: fc(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        double precision, intent(out) :: sf(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_field_decompose_physical(fc, sf), 'field_decompose_physical')
    end subroutine field_decompose_physical

    subroutine field_decompose_semi_spectral(sfc)                              ! inversion_utils.f90:563-592
        double precision, intent(inout) :: sfc(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_field_decompose_semi_spectral(sfc), 'field_decompose_semi_spectral')
    end subroutine field_decompose_semi_spectral

    subroutine field_combine_physical(sf, fc)                                  ! inversion_utils.f90:599-612
        double precision, intent(in)  :: sf(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        double precision, intent(out) :: fc(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_field_combine_physical(sf, fc), 'field_combine_physical')
    end subroutine field_combine_physical

    subroutine field_combine_semi_spectral(sf)                                 ! inversion_utils.f90:617-645
        double precision, intent(inout) :: sf(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_field_combine_semi_spectral(sf), 'field_combine_semi_spectral')
    end subroutine field_combine_semi_spectral

    subroutine central_diffz(fs, ds)                                           ! inversion_utils.f90:653-680
        double precision, intent(in)  :: fs(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        double precision, intent(out) :: ds(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_central_diffz(fs, ds), 'central_diffz')
    end subroutine central_diffz

#ifdef ENABLE_BUOYANCY
    subroutine diffz(fs, ds)                                                   ! inversion_utils.f90:683-719
        double precision, intent(in)  :: fs(0:nz, box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        double precision, intent(out) :: ds(0:nz, box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_diffz(fs, ds), 'diffz')
    end subroutine diffz
#endif

end module inversion_utils
