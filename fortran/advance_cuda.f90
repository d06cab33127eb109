! Replacement body of module advance_mod (reference src/stepper/advance.f90) on top of libps3d_cuda.
! `advance` keeps its interface (advance.f90:77-104).  Two forms:
!   * the default below makes ONE library call per time step (ps3d_cuda_advance = vor2vel, adapt incl.
!     set_diffusion, source, step, with the cn2 update riding on the source kernel) and hands the diagnostics of
!     adapt to set_netcdf_field_diagnostic exactly as advance.f90:188-193, 315-321, 366 do;
!   * with -DPS3D_CUDA_CALL_BY_CALL the reference's own sequence of calls is kept (vor2vel; adapt; write_step;
!     source; bstep%step), each forwarding to its one-to-one entry point -- same results, four host round trips.
! write_step (utils.f90:64-87) needs the fields of the *beginning* of the step on the host at output cadence: they
! are downloaded before the step when a write is due (write_due mirrors the test inside write_step).
module advance_mod
    use, intrinsic :: iso_c_binding
    use options, only : time, visc_type, vor_visc, stepper, output
#ifdef ENABLE_BUOYANCY
    use options, only : buoy_visc
#endif
    use constants
    use inversion_mod, only : vor2vel, source, ps3d_cuda_fields_to_host
    use utils, only : write_step
    use field_diagnostics_netcdf, only : set_netcdf_field_diagnostic        &
                                       , NC_OMAX, NC_ORMS, NC_OCHAR         &
                                       , NC_OXMEAN, NC_OYMEAN, NC_OZMEAN    &
                                       , NC_GMAX, NC_RGMAX, NC_RBFMAX       &
                                       , NC_BFMAX, NC_UMAX, NC_VMAX         &
                                       , NC_WMAX, NC_USGMAX, NC_LSGMAX
    use ps3d_cuda_mod
    implicit none

    type, abstract :: base_stepper
        contains
            procedure(base_diffusion),  deferred :: set_diffusion
            procedure(base_setup),      deferred :: setup
            procedure(base_step),       deferred :: step
    end type

    abstract interface
        subroutine base_diffusion(self, dt, vorch, bf)
            import base_stepper
            class(base_stepper), intent(inout) :: self
            double precision,    intent(in)    :: dt
            double precision,    intent(in)    :: vorch
            double precision,    intent(in)    :: bf
        end subroutine base_diffusion
        subroutine base_setup(self)
            import base_stepper
            class(base_stepper), intent(inout) :: self
        end subroutine base_setup
        subroutine base_step(self, t, dt)
            import base_stepper
            class(base_stepper), intent(inout) :: self
            double precision,    intent(inout) :: t
            double precision,    intent(in)    :: dt
        end subroutine base_step
    end interface

    integer :: advance_timer

    !Diagnostic quantities (advance.f90:66-73):
    double precision :: bfmax, vortmax, vortrms, ggmax, velmax
    double precision :: usggmax, lsggmax, rmv
#ifdef ENABLE_BUOYANCY
    double precision :: rmb
#endif
    double precision :: vorch

contains

    subroutine hand_over_diagnostics(diag)                 ! advance.f90:188-193, 315-321, 366
        real(c_double), intent(in) :: diag(16)
#ifdef ENABLE_BUOYANCY
        real(c_double) :: bd(4)
#endif
        vortmax = diag(PS3D_D_VORTMAX); vortrms = diag(PS3D_D_VORTRMS); vorch = diag(PS3D_D_VORCH)
        bfmax = diag(PS3D_D_BFMAX); ggmax = diag(PS3D_D_GGMAX)
        usggmax = diag(PS3D_D_USGGMAX); lsggmax = diag(PS3D_D_LSGGMAX); rmv = diag(PS3D_D_RMV)
        call set_netcdf_field_diagnostic(vortmax, NC_OMAX)
        call set_netcdf_field_diagnostic(vortrms, NC_ORMS)
        call set_netcdf_field_diagnostic(vorch, NC_OCHAR)
        call set_netcdf_field_diagnostic(diag(PS3D_D_VORMEAN_X), NC_OXMEAN)
        call set_netcdf_field_diagnostic(diag(PS3D_D_VORMEAN_Y), NC_OYMEAN)
        call set_netcdf_field_diagnostic(diag(PS3D_D_VORMEAN_Z), NC_OZMEAN)
        call set_netcdf_field_diagnostic(bfmax, NC_BFMAX)
        call set_netcdf_field_diagnostic(ggmax, NC_GMAX)
        call set_netcdf_field_diagnostic(diag(PS3D_D_UMAX), NC_UMAX)
        call set_netcdf_field_diagnostic(diag(PS3D_D_VMAX), NC_VMAX)
        call set_netcdf_field_diagnostic(diag(PS3D_D_WMAX), NC_WMAX)
        call set_netcdf_field_diagnostic(usggmax, NC_USGMAX)
        call set_netcdf_field_diagnostic(lsggmax, NC_LSGMAX)
        call set_netcdf_field_diagnostic(rmv, NC_RGMAX)
#ifdef ENABLE_BUOYANCY
        call ps3d_cuda_check(ps3d_cuda_buoyancy_diag(bd), 'buoyancy_diag')
        rmb = bd(2)
        call set_netcdf_field_diagnostic(rmb, NC_RBFMAX)
#endif
    end subroutine hand_over_diagnostics

    ! the test of write_step (utils.f90:77-87), needed here to fetch the fields only when they will be written
    logical function write_due(t)
        use utils, only : nfw_next, nsfw_next          ! add two accessor functions for the private counters nfw, nsfw
        double precision, intent(in) :: t
        write_due = (output%write_fields .and. (t + epsilon(zero) >= nfw_next() * output%field_freq)) .or. &
                    (output%write_field_stats .and. (t + epsilon(zero) >= nsfw_next() * output%field_stats_freq))
    end function write_due

    ! Advances the fields from time t to t + dt (advance.f90:77-104)
    subroutine advance(bstep, t)
        class(base_stepper), intent(inout) :: bstep
        double precision,    intent(inout) :: t
        double precision                   :: dt
        real(c_double)                     :: diag(16)

#ifdef PS3D_CUDA_CALL_BY_CALL
        call vor2vel                                                            ! advance.f90:85
        call adapt(bstep, t, dt)                                                ! :88
        if (write_due(t)) call ps3d_cuda_fields_to_host
        call write_step(t)                                                      ! :91
        call source                                                             ! :95
        call bstep%step(t, dt)                                                  ! :102
#else
        if (write_due(t)) then
            ! the reference writes the state *after* vor2vel and adapt of this step (advance.f90:85-91)
            call vor2vel
            call adapt(bstep, t, dt)
            call ps3d_cuda_fields_to_host
            call write_step(t)
            call source
            call bstep%step(t, dt)
            return
        endif
        call ps3d_cuda_check(ps3d_cuda_advance(t, time%limit, time%alpha, pretype_id(vor_visc%pretype),        &
                                               int(vor_visc%roll_mean_win_size, c_int), dt, diag), 'advance')
        call hand_over_diagnostics(diag)
#endif
    end subroutine advance

    ! Adapts the time step and computes various diagnostics (advance.f90:109-377); includes
    ! bstep%set_diffusion(dt, vval, bval) (:375).  pressure / horizontal_divergence (:278-282) are evaluated by the
    ! library when PS3D_F_PRES / PS3D_F_DELTA are downloaded.
    subroutine adapt(bstep, t, dt)
        class(base_stepper), intent(inout) :: bstep
        double precision,    intent(in)    :: t
        double precision,    intent(inout) :: dt
        real(c_double)                     :: diag(16)
        call ps3d_cuda_check(ps3d_cuda_adapt(t, time%limit, time%alpha, pretype_id(vor_visc%pretype),          &
                                             int(vor_visc%roll_mean_win_size, c_int), dt, diag), 'adapt')
        call hand_over_diagnostics(diag)
    end subroutine adapt

end module advance_mod
