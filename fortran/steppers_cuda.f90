! Replacement bodies of cn2_mod (reference src/stepper/cn2.f90) and impl_rk4_mod (src/stepper/impl_rk4.f90) on top of
! libps3d_cuda.  The derived types keep the reference's deferred-procedure interface (advance.f90:30-60:
! set_diffusion(dt, vorch, bf), setup, step(t, dt)) so that ps3d.f90:90-101 (`bstep = cn2()` / `impl_rk4()`,
! `call bstep%setup`) is unchanged.  The work arrays (vortsm, bsm, epq, emq, svori, svorf, ...; cn2.f90:31,
! impl_rk4.f90:57-72) live on the device.
module cn2_mod
    use, intrinsic :: iso_c_binding
    use advance_mod, only : base_stepper
    use ps3d_cuda_mod
    implicit none

    type, extends(base_stepper) :: cn2
        contains
            procedure :: set_diffusion => cn2_set_diffusion
            procedure :: setup  => cn2_setup
            procedure :: step => cn2_step
    end type

contains

    subroutine cn2_set_diffusion(self, dt, vorch, bf)                          ! cn2.f90:40-79
        class(cn2),       intent(inout) :: self
        double precision, intent(in)    :: dt
        double precision, intent(in)    :: vorch, bf
        call ps3d_cuda_check(ps3d_cuda_set_diffusion(dt, vorch), 'cn2_set_diffusion')
#ifdef ENABLE_BUOYANCY
        call ps3d_cuda_check(ps3d_cuda_set_diffusion_buoyancy(dt, bf), 'cn2_set_diffusion (buoyancy)')
#endif
    end subroutine cn2_set_diffusion

    subroutine cn2_setup(self)                                                 ! cn2.f90:83-88
        class(cn2), intent(inout) :: self
        call ps3d_cuda_check(ps3d_cuda_stepper_setup(PS3D_STEPPER_CN2), 'cn2_setup')
    end subroutine cn2_setup

    subroutine cn2_step(self, t, dt)                                           ! cn2.f90:92-181
        class(cn2),       intent(inout) :: self
        double precision, intent(inout) :: t
        double precision, intent(in)    :: dt
        call ps3d_cuda_check(ps3d_cuda_step(t, dt), 'cn2_step')
    end subroutine cn2_step

end module cn2_mod

module impl_rk4_mod
    use, intrinsic :: iso_c_binding
    use advance_mod, only : base_stepper
    use ps3d_cuda_mod
    implicit none

    type, extends(base_stepper) :: impl_rk4
        contains
            procedure :: set_diffusion => impl_rk4_set_diffusion
            procedure :: setup  => impl_rk4_setup
            procedure :: step => impl_rk4_step
    end type

contains

    subroutine impl_rk4_set_diffusion(self, dt, vorch, bf)                     ! impl_rk4.f90:37-53
        class(impl_rk4),  intent(inout) :: self
        double precision, intent(in)    :: dt
        double precision, intent(in)    :: vorch, bf
        call ps3d_cuda_check(ps3d_cuda_set_diffusion(dt, vorch), 'impl_rk4_set_diffusion')
#ifdef ENABLE_BUOYANCY
        call ps3d_cuda_check(ps3d_cuda_set_diffusion_buoyancy(dt, bf), 'impl_rk4_set_diffusion (buoyancy)')
#endif
    end subroutine impl_rk4_set_diffusion

    subroutine impl_rk4_setup(self)                                            ! impl_rk4.f90:57-72
        class(impl_rk4), intent(inout) :: self
        call ps3d_cuda_check(ps3d_cuda_stepper_setup(PS3D_STEPPER_IMPL_RK4), 'impl_rk4_setup')
    end subroutine impl_rk4_setup

    subroutine impl_rk4_step(self, t, dt)                                      ! impl_rk4.f90:76-207
        class(impl_rk4),  intent(inout) :: self
        double precision, intent(inout) :: t
        double precision, intent(in)    :: dt
        call ps3d_cuda_check(ps3d_cuda_step(t, dt), 'impl_rk4_step')
    end subroutine impl_rk4_step

end module impl_rk4_mod
