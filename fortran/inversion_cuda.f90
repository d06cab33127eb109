! Replacement body of module inversion_mod (reference src/inversion/inversion.f90) on top of libps3d_cuda.
! The prognostic fields are resident in HBM; the module arrays of fields.f90 are host mirrors that are refreshed
! only when somebody needs them (ps3d_cuda_fields_to_host below: output cadence, unit tests).  vor2vel / source
! keep their names and their argument-less interface (inversion.f90:23, :378).
module inversion_mod
    use, intrinsic :: iso_c_binding
    use fields
    use ps3d_cuda_mod
    implicit none

    integer :: vor2vel_timer, vtend_timer

    ! .true.: every call also mirrors its outputs into the module arrays of fields.f90, which makes the
    ! reference's unit tests (unit-tests/test_vor2vel_*.f90, test_vtend.f90) run unchanged; the time loop leaves it
    ! .false. and downloads at output cadence only
    logical :: l_mirror_host = .false.

contains

    ! host svor -> device (for callers that fill svor themselves, e.g. unit-tests/test_vor2vel_1.f90:60-75)
    subroutine ps3d_cuda_svor_to_device
        integer :: nc
        do nc = 1, 3
            call ps3d_cuda_check(ps3d_cuda_upload(PS3D_F_SVOR, int(nc - 1, c_int), svor(:, :, :, nc)), 'upload svor')
        enddo
    end subroutine ps3d_cuda_svor_to_device

    ! device -> the module arrays written by field_netcdf.f90:216-230 and read by the diagnostics
    subroutine ps3d_cuda_fields_to_host
        integer :: nc
        do nc = 1, 3
            call ps3d_cuda_check(ps3d_cuda_download(PS3D_F_SVOR, int(nc - 1, c_int), svor(:, :, :, nc)), 'download svor')
            call ps3d_cuda_check(ps3d_cuda_download(PS3D_F_VOR,  int(nc - 1, c_int), vor(:, :, :, nc)),  'download vor')
            call ps3d_cuda_check(ps3d_cuda_download(PS3D_F_VEL,  int(nc - 1, c_int), vel(:, :, :, nc)),  'download vel')
            call ps3d_cuda_check(ps3d_cuda_download(PS3D_F_SVEL, int(nc - 1, c_int), svel(:, :, :, nc)), 'download svel')
        enddo
#ifdef ENABLE_BUOYANCY
        call ps3d_cuda_check(ps3d_cuda_download(PS3D_F_SBUOY, 0_c_int, sbuoy), 'download sbuoy')
        call ps3d_cuda_check(ps3d_cuda_download(PS3D_F_BUOY,  0_c_int, buoy),  'download buoy')
#endif
    end subroutine ps3d_cuda_fields_to_host

    ! Given the vorticity vector field (svor) in spectral space, returns the associated velocity field (vel)
    ! (inversion.f90:23-226)
    subroutine vor2vel
        call start_timer(vor2vel_timer)
        if (l_mirror_host) call ps3d_cuda_svor_to_device
        call ps3d_cuda_check(ps3d_cuda_vor2vel(), 'vor2vel')
        if (l_mirror_host) call ps3d_cuda_fields_to_host
        call stop_timer(vor2vel_timer)
    end subroutine vor2vel

    ! source terms for vorticity (and buoyancy) in mixed-spectral space (inversion.f90:378-388); the library
    ! evaluates buoyancy_tendency (:232-292) first when ps3d_cuda_enable_buoyancy has been called
    subroutine source
        integer :: nc
        call start_timer(vtend_timer)
        call ps3d_cuda_check(ps3d_cuda_source(), 'source')
        if (l_mirror_host) then
            do nc = 1, 3
                call ps3d_cuda_check(ps3d_cuda_download(PS3D_F_SVORTS, int(nc - 1, c_int), svorts(:, :, :, nc)), &
                                     'download svorts')
            enddo
#ifdef ENABLE_BUOYANCY
            call ps3d_cuda_check(ps3d_cuda_download(PS3D_F_SBUOYS, 0_c_int, sbuoys), 'download sbuoys')
#endif
        endif
        call stop_timer(vtend_timer)
    end subroutine source

end module inversion_mod
