! fsilbm_gpu.f90 -- ISO_C_BINDING shim between the FSILBM3D Fortran driver and libfsilbm_b200.so.
!
! NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT: the build image has no Fortran compiler
! (DESIGN.md "Boundary").  The file is kept small and mechanical: module fsilbm_c holds one
! bind(C) interface per symbol of include/fsilbm.h, module fsilbm_gpu holds the replacement
! bodies for the reference procedures on the hot path.  INTEGRATION.md lists, call point by call
! point, which reference line each routine replaces.  -std=f2003 (reference Makefile:24) allows
! everything used here.
!
! Conventions: scalars by VALUE, arrays by reference (first element), every function returns
! integer(c_int) 0 on success; fsilbm_check turns a non-zero code into the reference's own error
! convention  write(*,*) msg ; stop  (e.g. FluidDomain.f90:704, Solidbody.f90:850).

module fsilbm_c
    use, intrinsic :: iso_c_binding
    implicit none
    public

    ! struct fsilbm_flow (include/fsilbm.h): the slice of FlowCondType (FlowCondition.f90:11-27) the path reads
    type, bind(C) :: fsilbm_flow
        real(c_double) :: nu
        real(c_double) :: denIn
        real(c_double) :: uvwIn(3)
        real(c_double) :: shearRateIn(3)
        integer(c_int) :: velocityKind
        real(c_double) :: volumeForceIn(3)
        real(c_double) :: volumeForceAmp, volumeForceFreq, volumeForcePhi
        real(c_double) :: Uref
    end type fsilbm_flow

    interface
        integer(c_int) function fsilbm_init(device) bind(C, name='fsilbm_init')
            import :: c_int
            integer(c_int), value :: device
        end function
        integer(c_int) function fsilbm_finalize() bind(C, name='fsilbm_finalize')
            import :: c_int
        end function
        type(c_ptr) function fsilbm_last_error() bind(C, name='fsilbm_last_error')
            import :: c_ptr
        end function
        integer(c_int) function fsilbm_block_create(xDim, yDim, zDim, xOffset, xLocal, dh, xmin, ymin, zmin, &
                                                    BndConds, iCollidModel, params, flow, handle) bind(C, name='fsilbm_block_create')
            import :: c_int, c_double, fsilbm_flow
            integer(c_int), value :: xDim, yDim, zDim, xOffset, xLocal, iCollidModel
            real(c_double), value :: dh, xmin, ymin, zmin
            integer(c_int), intent(in) :: BndConds(6)
            real(c_double), intent(in) :: params(10)
            type(fsilbm_flow), intent(in) :: flow
            integer(c_int), intent(out) :: handle
        end function
        integer(c_int) function fsilbm_block_destroy(h) bind(C, name='fsilbm_block_destroy')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_initialise(h, time) bind(C, name='fsilbm_block_initialise')
            import :: c_int, c_double
            integer(c_int), value :: h
            real(c_double), value :: time
        end function
        integer(c_int) function fsilbm_block_get(h, what, val) bind(C, name='fsilbm_block_get')
            import :: c_int, c_double
            integer(c_int), value :: h, what
            real(c_double), intent(out) :: val
        end function
        integer(c_int) function fsilbm_block_upload_fIn(h, fIn) bind(C, name='fsilbm_block_upload_fIn')
            import :: c_int, c_double
            integer(c_int), value :: h
            real(c_double), intent(in) :: fIn(*)
        end function
        integer(c_int) function fsilbm_block_download_fIn(h, fIn) bind(C, name='fsilbm_block_download_fIn')
            import :: c_int, c_double
            integer(c_int), value :: h
            real(c_double), intent(out) :: fIn(*)
        end function
        integer(c_int) function fsilbm_block_set_time(h, blktime) bind(C, name='fsilbm_block_set_time')
            import :: c_int, c_double
            integer(c_int), value :: h
            real(c_double), value :: blktime
        end function
        integer(c_int) function fsilbm_block_update_volume_force(h, volumeForce) bind(C, name='fsilbm_block_update_volume_force')
            import :: c_int, c_double
            integer(c_int), value :: h
            real(c_double), intent(out) :: volumeForce(3)
        end function
        integer(c_int) function fsilbm_block_download_macro(h, den, uuu) bind(C, name='fsilbm_block_download_macro')
            import :: c_int, c_double
            integer(c_int), value :: h
            real(c_double), intent(out) :: den(*), uuu(*)
        end function
        ! the same without waiting (den, uuu must be page-locked); valid after fsilbm_block_download_wait
        integer(c_int) function fsilbm_block_download_macro_async(h, den, uuu) bind(C, name='fsilbm_block_download_macro_async')
            import :: c_int, c_double
            integer(c_int), value :: h
            real(c_double), intent(out) :: den(*), uuu(*)
        end function
        integer(c_int) function fsilbm_block_download_wait(h) bind(C, name='fsilbm_block_download_wait')
            import :: c_int
            integer(c_int), value :: h
        end function
        ! slab runs: 0 body not iterated by this rank, 1 iterated, 2 iterated and led
        integer(c_int) function fsilbm_ibm_body_status(h, nbody, status) bind(C, name='fsilbm_ibm_body_status')
            import :: c_int
            integer(c_int), value :: h, nbody
            integer(c_int), intent(out) :: status(*)
        end function
        integer(c_int) function fsilbm_block_field_stat(h, stat) bind(C, name='fsilbm_block_field_stat')
            import :: c_int, c_double
            integer(c_int), value :: h
            real(c_double), intent(out) :: stat(6)
        end function
        integer(c_int) function fsilbm_block_set_boundary_conditions(h) bind(C, name='fsilbm_block_set_boundary_conditions')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_collide_stream(h) bind(C, name='fsilbm_block_collide_stream')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_sync(h) bind(C, name='fsilbm_block_sync')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_ibm_interaction_force(h, nbody, nelmts, Exyz, Evel, Ea, Eforce, restencil, dt, &
                                                             ntolLBM, dtolLBM, rootBC, iterLBM) bind(C, name='fsilbm_ibm_interaction_force')
            import :: c_int, c_double, c_ptr
            integer(c_int), value :: h, nbody, ntolLBM
            integer(c_int), intent(in) :: nelmts(*), restencil(*), rootBC(6)
            type(c_ptr), intent(in) :: Exyz(*), Evel(*), Ea(*), Eforce(*)   ! one pointer per body
            real(c_double), value :: dt, dtolLBM
            integer(c_int), intent(out) :: iterLBM
        end function
    end interface
end module fsilbm_c

module fsilbm_gpu
    use, intrinsic :: iso_c_binding
    use fsilbm_c
    implicit none
    private
    public :: fsilbm_check, gpu_allocate_block, gpu_initialise_block, gpu_upload_fIn, gpu_refresh_host_fIn, &
              gpu_refresh_host_macro, gpu_update_volume_force, gpu_set_boundary_conditions, gpu_collide_stream, &
              gpu_interaction_force, gpu_field_stat
    integer(c_int), allocatable, public :: gpu_handle(:)    ! one library handle per LBMblks(i)
contains

    ! the reference's error convention: print and stop
    subroutine fsilbm_check(ierr)
        integer(c_int), intent(in) :: ierr
        character(kind=c_char), pointer :: msg(:)
        integer :: n
        if (ierr == 0) return
        call c_f_pointer(fsilbm_last_error(), msg, [512])
        n = 1
        do while (n < 512 .and. msg(n) /= c_null_char)
            n = n + 1
        enddo
        write(*,*) msg(1:n-1)
        stop
    end subroutine

    ! after allocate_fluid_ (FluidDomain.f90:378-408), main.f90:40.  nu etc. come from the global `flow`.
    subroutine gpu_allocate_block(iblock, xDim, yDim, zDim, dh, xmin, ymin, zmin, BndConds, iCollidModel, params, &
                                  nu, denIn, uvwIn, shearRateIn, velocityKind, volumeForceIn, volumeForceAmp, &
                                  volumeForceFreq, volumeForcePhi, Uref)
        integer, intent(in) :: iblock, xDim, yDim, zDim, BndConds(6), iCollidModel, velocityKind
        real(8), intent(in) :: dh, xmin, ymin, zmin, params(10), nu, denIn, uvwIn(3), shearRateIn(3), volumeForceIn(3)
        real(8), intent(in) :: volumeForceAmp, volumeForceFreq, volumeForcePhi, Uref
        type(fsilbm_flow) :: cf
        cf%nu = nu; cf%denIn = denIn; cf%uvwIn = uvwIn; cf%shearRateIn = shearRateIn; cf%velocityKind = velocityKind
        cf%volumeForceIn = volumeForceIn; cf%volumeForceAmp = volumeForceAmp; cf%volumeForceFreq = volumeForceFreq
        cf%volumeForcePhi = volumeForcePhi; cf%Uref = Uref
        ! one GPU: the block is one slab, xOffset = 0, xLocal = xDim
        call fsilbm_check(fsilbm_block_create(xDim, yDim, zDim, 0, xDim, dh, xmin, ymin, zmin, BndConds, iCollidModel, params, cf, gpu_handle(iblock)))
    end subroutine

    ! replaces the body of initialise_ (FluidDomain.f90:433-545); called from initialise_fuild_blocks, main.f90:50
    subroutine gpu_initialise_block(iblock, time, tau, Omega, Omega2)
        integer, intent(in) :: iblock
        real(8), intent(in) :: time
        real(8), intent(out) :: tau, Omega, Omega2
        call fsilbm_check(fsilbm_block_initialise(gpu_handle(iblock), time))
        call fsilbm_check(fsilbm_block_get(gpu_handle(iblock), 0, tau))
        call fsilbm_check(fsilbm_block_get(gpu_handle(iblock), 1, Omega))
        call fsilbm_check(fsilbm_block_get(gpu_handle(iblock), 2, Omega2))
    end subroutine

    ! after read_continue_ filled this%fIn (FluidDomain.f90:187-214), main.f90:58
    subroutine gpu_upload_fIn(iblock, fIn)
        integer, intent(in) :: iblock
        real(8), intent(in) :: fIn(*)      ! fIn(zDim,yDim,xDim,0:lbmDim), passed as its first element
        call fsilbm_check(fsilbm_block_upload_fIn(gpu_handle(iblock), fIn))
    end subroutine

    ! before write_continue_ (FluidDomain.f90:1770-1777), main.f90:118
    subroutine gpu_refresh_host_fIn(iblock, fIn)
        integer, intent(in) :: iblock
        real(8), intent(out) :: fIn(*)
        call fsilbm_check(fsilbm_block_download_fIn(gpu_handle(iblock), fIn))
    end subroutine

    ! replaces calculate_macro_quantities_ where the HOST arrays are needed: before write_flow_ (main.f90:84,130),
    ! write_fluid_flux (:135), write_fluid_information (:137), calculate_turbulent_statistic (:108).
    ! The per-step call at main.f90:107 is dropped: inside the step den/uuu live in registers.
    subroutine gpu_refresh_host_macro(iblock, den, uuu)
        integer, intent(in) :: iblock
        real(8), intent(out) :: den(*), uuu(*)
        call fsilbm_check(fsilbm_block_download_macro(gpu_handle(iblock), den, uuu))
    end subroutine

    ! replaces update_volume_force_ (FluidDomain.f90:1174-1180); LBMBlockComm.f90:283, main.f90:62
    subroutine gpu_update_volume_force(iblock, blktime, volumeForce)
        integer, intent(in) :: iblock
        real(8), intent(in) :: blktime
        real(8), intent(out) :: volumeForce(3)
        call fsilbm_check(fsilbm_block_set_time(gpu_handle(iblock), blktime))
        call fsilbm_check(fsilbm_block_update_volume_force(gpu_handle(iblock), volumeForce))
    end subroutine

    ! replaces set_boundary_conditions_ at start-up (tree_set_boundary_conditions_block, main.f90:63)
    subroutine gpu_set_boundary_conditions(iblock)
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_set_boundary_conditions(gpu_handle(iblock)))
    end subroutine

    ! replaces LBMBlockComm.f90:285-286,288,293-303 in one call (macro, reset, add force, collision, halfwayBCset,
    ! streaming, set_boundary_conditions).  Asynchronous; the next library call on the block orders after it.
    subroutine gpu_collide_stream(iblock)
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_collide_stream(gpu_handle(iblock)))
    end subroutine

    ! replaces calculate_interaction_force (Solidbody.f90:869-918) minus lodFlow assembly.  The caller keeps
    ! UpdatePosVelArea (Solidbody.f90:597-600) before and the nodal-load half of FluidVolumeForce_ (:945-967) after.
    ! markers(b) etc. are c_loc() of VBodies(iFish)%v_Exyz / v_Evel / v_Ea / v_Eforce of the carried bodies.
    subroutine gpu_interaction_force(iblock, nbody, nelmts, Exyz, Evel, Ea, Eforce, restencil, dt, ntolLBM, dtolLBM, rootBC, iterLBM)
        integer, intent(in) :: iblock, nbody, nelmts(nbody), restencil(nbody), ntolLBM, rootBC(6)
        type(c_ptr), intent(in) :: Exyz(nbody), Evel(nbody), Ea(nbody), Eforce(nbody)
        real(8), intent(in) :: dt, dtolLBM
        integer, intent(out) :: iterLBM
        call fsilbm_check(fsilbm_ibm_interaction_force(gpu_handle(iblock), nbody, nelmts, Exyz, Evel, Ea, Eforce, restencil, dt, &
                                                       ntolLBM, dtolLBM, rootBC, iterLBM))
    end subroutine

    ! replaces ComputeFieldStat_ (FluidDomain.f90:1739-1768), main.f90:150
    subroutine gpu_field_stat(iblock, ncell, stat)
        integer, intent(in) :: iblock
        real(8), intent(in) :: ncell
        real(8), intent(out) :: stat(6)
        call fsilbm_check(fsilbm_block_field_stat(gpu_handle(iblock), stat))
        stat(1:3) = dsqrt(stat(1:3) / ncell)
    end subroutine

end module fsilbm_gpu
