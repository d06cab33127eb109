! Replacement body of module sta3dfft (reference src/fft/sta3dfft.f90) on top of libps3d_cuda.
! Same public list as sta3dfft.f90:30-47 minus the FFT work arrays (xtrig, xfactors, ...: they belong to stafft,
! which is no longer called).  The wavenumber arrays stay on the host because field_diagnostics and the output
! modules use them; their definition is sta3dfft.f90:89-108 / deriv1d.f90:17-22 (pi k / L with the Nyquist entry).
!
! initialise_fft also creates the library context: it is the first module procedure of the path that knows the
! domain (it is reached from init_inversion, inversion_utils.f90:231, after mpi_layout_init).  The library's slab
! decomposition needs the 1-D layout in x: mpi_layout_init must be called with dims = (/world%size, 1/)
! (mpi_layout.f90:72-77 chooses dims with MPI_Dims_create; fix them there, see fortran/README.md).
!
! transpose_to_pencil (fft_pencil.f90:283-330) and reverse_x/reverse_y (mpi_reverse.f90:332-398) have no counterpart:
! the slab exchange lives inside the library (one all-to-all per 2-D FFT, fused into the first sweep).
module sta3dfft
    use, intrinsic :: iso_c_binding
    use mpi_layout
    use mpi_environment, only : world
    use constants, only : zero, pi
    use parameters, only : lower
    use ps3d_cuda_mod
    implicit none

    private

    double precision, protected, allocatable :: rkx(:), hrkx(:), rky(:), hrky(:), rkz(:), rkzi(:)
    integer :: nx, ny, nz
    logical :: is_fft_initialised = .false.

    ! 128-byte ncclUniqueId for runs on several ranks, filled by the driver (rank 0 creates it with
    ! ncclGetUniqueId, MPI_Bcast to all); left unset the library expects ps3d_cuda_set_transport (MPI transport)
    character(kind=c_char), target, public :: ps3d_nccl_id(128)
    logical, public :: ps3d_have_nccl_id = .false.

    public :: initialise_fft, finalise_fft, diffx, diffy, fftxyp2s, fftxys2p, fftsine, fftcosine, &
              rkx, rky, rkz, rkzi

contains

    subroutine initialise_fft(extent)
        double precision, intent(in) :: extent(3)
        integer                      :: k
        type(c_ptr)                  :: id

        if (is_fft_initialised) return

        nx = box%global_size(1)
        ny = box%global_size(2)
        nz = box%global_size(3)

        id = c_null_ptr
        if (world%size > 1 .and. ps3d_have_nccl_id) id = c_loc(ps3d_nccl_id)
        call ps3d_cuda_check(ps3d_cuda_init(int(nx, c_int), int(ny, c_int), int(nz, c_int), lower, extent, &
                                            int(world%rank, c_int), int(world%size, c_int), id), 'init')

        allocate(rkx(0:nx-1), hrkx(nx), rky(0:ny-1), hrky(ny), rkz(0:nz), rkzi(1:nz-1))
        ! sta2dfft.f90:59-66 (init_deriv) and sta3dfft.f90:89-108
        rkx = zero
        rky = zero
        do k = 1, nx / 2 - 1
            rkx(k) = (pi / extent(1)) * dble(2 * k)
            rkx(nx - k) = rkx(k)
        enddo
        rkx(nx / 2) = (pi / extent(1)) * dble(nx)
        do k = 1, ny / 2 - 1
            rky(k) = (pi / extent(2)) * dble(2 * k)
            rky(ny - k) = rky(k)
        enddo
        rky(ny / 2) = (pi / extent(2)) * dble(ny)
        hrkx = zero
        hrky = zero
        do k = 1, nx / 2 - 1
            hrkx(2 * k) = rkx(k)
            hrkx(2 * k + 1) = rkx(k)
        enddo
        do k = 1, ny / 2 - 1
            hrky(2 * k) = rky(k)
            hrky(2 * k + 1) = rky(k)
        enddo
        rkz(0) = zero
        do k = 1, nz
            rkz(k) = (pi / extent(3)) * dble(k)
        enddo
        rkzi(1:nz-1) = 1.0d0 / rkz(1:nz-1)

        is_fft_initialised = .true.
    end subroutine initialise_fft

    subroutine finalise_fft
        if (.not. is_fft_initialised) return
        call ps3d_cuda_check(ps3d_cuda_finalise(), 'finalise')
        deallocate(rkx, hrkx, rky, hrky, rkz, rkzi)
        is_fft_initialised = .false.
    end subroutine finalise_fft

    ! physical -> semi-spectral in x and y (sta3dfft.f90:136-194); fp is NOT destroyed here
    subroutine fftxyp2s(fp, fs)
        double precision, intent(inout) :: fp(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        double precision, intent(out)   :: fs(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_fftxyp2s(fp, fs), 'fftxyp2s')
    end subroutine fftxyp2s

    ! semi-spectral -> physical (sta3dfft.f90:202-260); fs is NOT destroyed here
    subroutine fftxys2p(fs, fp)
        double precision, intent(inout) :: fs(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        double precision, intent(out)   :: fp(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_fftxys2p(fs, fp), 'fftxys2p')
    end subroutine fftxys2p

    subroutine fftsine(fs)                                                    ! sta3dfft.f90:264-278
        double precision, intent(inout) :: fs(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_fftsine(fs), 'fftsine')
    end subroutine fftsine

    subroutine fftcosine(fs)                                                  ! sta3dfft.f90:282-296
        double precision, intent(inout) :: fs(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_fftcosine(fs), 'fftcosine')
    end subroutine fftcosine

    subroutine diffx(fs, ds)                                                  ! sta3dfft.f90:304-337
        double precision, intent(in)  :: fs(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        double precision, intent(out) :: ds(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_diffx(fs, ds), 'diffx')
    end subroutine diffx

    subroutine diffy(fs, ds)                                                  ! sta3dfft.f90:345-377
        double precision, intent(in)  :: fs(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        double precision, intent(out) :: ds(box%lo(3):box%hi(3), box%lo(2):box%hi(2), box%lo(1):box%hi(1))
        call ps3d_cuda_check(ps3d_cuda_diffy(fs, ds), 'diffy')
    end subroutine diffy

end module sta3dfft
