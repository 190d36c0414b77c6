! fsilbm_gpu.f90 -- ISO_C_BINDING shim between the FSILBM3D Fortran driver and libfsilbm_b200.so.
!
! NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT: the build image has no Fortran compiler
! (DESIGN.md section 5).  The file is kept mechanical so that it can be checked without one:
!   * module fsilbm_c holds ONE bind(C) interface PER SYMBOL of include/fsilbm.h -- all of them, same names, same argument
!     order and count; tests/test_abi.py::test_fortran_shim_binds_every_symbol compares the two files;
!   * module fsilbm_gpu holds the replacement bodies for the reference procedures on the hot path, one per call point;
!     INTEGRATION.md lists which reference line each routine replaces.
! -std=f2003 (reference Makefile:24) allows everything used here.
!
! Conventions: scalars by VALUE, arrays by reference (first element), C strings as character(kind=c_char) arrays ending in
! c_null_char, every function returns integer(c_int) 0 on success; fsilbm_check turns a non-zero code into the reference's
! own error convention  write(*,*) msg ; stop  (e.g. FluidDomain.f90:704, Solidbody.f90:850).

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
        integer(c_long_long) function fsilbm_launch_count() bind(C, name='fsilbm_launch_count')
            import :: c_long_long
        end function
        integer(c_long_long) function fsilbm_ibm_early_count() bind(C, name='fsilbm_ibm_early_count')
            import :: c_long_long
        end function
        integer(c_int) function fsilbm_trace_dump(path) bind(C, name='fsilbm_trace_dump')
            import :: c_char, c_int
            character(kind=c_char), intent(in) :: path(*)
        end function
        integer(c_int) function fsilbm_set_option(key, val) bind(C, name='fsilbm_set_option')
            import :: c_char, c_int
            character(kind=c_char), intent(in) :: key(*)
            integer(c_int), value :: val
        end function
        integer(c_int) function fsilbm_block_create(xDim, yDim, zDim, xOffset, xLocal, dh, xmin, &
                ymin, zmin, BndConds, iCollidModel, params, flow, out) bind(C, name='fsilbm_block_create')
            import :: c_double, c_int, fsilbm_flow
            integer(c_int), value :: xDim
            integer(c_int), value :: yDim
            integer(c_int), value :: zDim
            integer(c_int), value :: xOffset
            integer(c_int), value :: xLocal
            real(c_double), value :: dh
            real(c_double), value :: xmin
            real(c_double), value :: ymin
            real(c_double), value :: zmin
            integer(c_int), intent(in) :: BndConds(6)
            integer(c_int), value :: iCollidModel
            real(c_double), intent(in) :: params(10)
            type(fsilbm_flow), intent(in) :: flow
            integer(c_int), intent(out) :: out
        end function
        integer(c_int) function fsilbm_block_destroy(h) bind(C, name='fsilbm_block_destroy')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_initialise(h, time) bind(C, name='fsilbm_block_initialise')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), value :: time
        end function
        integer(c_int) function fsilbm_block_get(h, what, val) bind(C, name='fsilbm_block_get')
            import :: c_double, c_int
            integer(c_int), value :: h
            integer(c_int), value :: what
            real(c_double), intent(out) :: val
        end function
        integer(c_int) function fsilbm_block_upload_fIn(h, fIn) bind(C, name='fsilbm_block_upload_fIn')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(in) :: fIn(*)
        end function
        integer(c_int) function fsilbm_block_download_fIn(h, fIn) bind(C, name='fsilbm_block_download_fIn')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(inout) :: fIn(*)
        end function
        integer(c_int) function fsilbm_block_set_time(h, blktime) bind(C, name='fsilbm_block_set_time')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), value :: blktime
        end function
        integer(c_int) function fsilbm_block_update_volume_force(h, &
                volumeForce_out) bind(C, name='fsilbm_block_update_volume_force')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(inout) :: volumeForce_out(3)
        end function
        integer(c_int) function fsilbm_block_download_macro(h, den, uuu) bind(C, name='fsilbm_block_download_macro')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(inout) :: den(*)
            real(c_double), intent(inout) :: uuu(*)
        end function
        integer(c_int) function fsilbm_block_download_macro_async(h, den, uuu) bind(C, name='fsilbm_block_download_macro_async')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(inout) :: den(*)
            real(c_double), intent(inout) :: uuu(*)
        end function
        integer(c_int) function fsilbm_block_download_wait(h) bind(C, name='fsilbm_block_download_wait')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_download_tau_all(h, tau_all) bind(C, name='fsilbm_block_download_tau_all')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(inout) :: tau_all(*)
        end function
        integer(c_int) function fsilbm_block_field_stat(h, out) bind(C, name='fsilbm_block_field_stat')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(inout) :: out(6)
        end function
        integer(c_int) function fsilbm_block_write_flow_window(h, offsetOutput, &
                outputtype, out) bind(C, name='fsilbm_block_write_flow_window')
            import :: c_float, c_int
            integer(c_int), value :: h
            integer(c_int), value :: offsetOutput
            integer(c_int), value :: outputtype
            real(c_float), intent(inout) :: out(*)
        end function
        integer(c_int) function fsilbm_block_write_flow_window_async(h, offsetOutput, &
                outputtype, out) bind(C, name='fsilbm_block_write_flow_window_async')
            import :: c_float, c_int
            integer(c_int), value :: h
            integer(c_int), value :: offsetOutput
            integer(c_int), value :: outputtype
            real(c_float), intent(inout) :: out(*)
        end function
        integer(c_int) function fsilbm_block_turbulent_statistic(h, step, step_s) bind(C, name='fsilbm_block_turbulent_statistic')
            import :: c_int
            integer(c_int), value :: h
            integer(c_int), value :: step
            integer(c_int), value :: step_s
        end function
        integer(c_int) function fsilbm_block_fluid_flux(h, out) bind(C, name='fsilbm_block_fluid_flux')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(inout) :: out(3)
        end function
        integer(c_int) function fsilbm_block_probe_velocity(h, n, coords, velocity) bind(C, name='fsilbm_block_probe_velocity')
            import :: c_double, c_int
            integer(c_int), value :: h
            integer(c_int), value :: n
            real(c_double), intent(in) :: coords(*)
            real(c_double), intent(inout) :: velocity(*)
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
        integer(c_int) function fsilbm_block_stream(h, stream) bind(C, name='fsilbm_block_stream')
            import :: c_int, c_ptr
            integer(c_int), value :: h
            type(c_ptr), intent(out) :: stream
        end function
        integer(c_int) function fsilbm_block_pass_macro(h) bind(C, name='fsilbm_block_pass_macro')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_pass_reset_volume_force(h) bind(C, name='fsilbm_block_pass_reset_volume_force')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_pass_add_volume_force(h) bind(C, name='fsilbm_block_pass_add_volume_force')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_pass_collision(h) bind(C, name='fsilbm_block_pass_collision')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_pass_halfway_bc_set(h) bind(C, name='fsilbm_block_pass_halfway_bc_set')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_pass_streaming(h) bind(C, name='fsilbm_block_pass_streaming')
            import :: c_int
            integer(c_int), value :: h
        end function
        integer(c_int) function fsilbm_block_download_fields(h, den, uuu, force) bind(C, name='fsilbm_block_download_fields')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(inout) :: den(*)
            real(c_double), intent(inout) :: uuu(*)
            real(c_double), intent(inout) :: force(*)
        end function
        integer(c_int) function fsilbm_block_upload_fields(h, den, uuu, force) bind(C, name='fsilbm_block_upload_fields')
            import :: c_double, c_int
            integer(c_int), value :: h
            real(c_double), intent(in) :: den(*)
            real(c_double), intent(in) :: uuu(*)
            real(c_double), intent(in) :: force(*)
        end function
        integer(c_int) function fsilbm_ibm_interaction_force(h, nbody, nelmts, Exyz, Evel, Ea, &
                Eforce, restencil, dt, ntolLBM, dtolLBM, rootBC, iterLBM_out) bind(C, name='fsilbm_ibm_interaction_force')
            import :: c_double, c_int, c_ptr
            integer(c_int), value :: h
            integer(c_int), value :: nbody
            integer(c_int), intent(in) :: nelmts(*)
            type(c_ptr), intent(in) :: Exyz(*)
            type(c_ptr), intent(in) :: Evel(*)
            type(c_ptr), intent(in) :: Ea(*)
            type(c_ptr), intent(in) :: Eforce(*)
            integer(c_int), intent(in) :: restencil(*)
            real(c_double), value :: dt
            integer(c_int), value :: ntolLBM
            real(c_double), value :: dtolLBM
            integer(c_int), intent(in) :: rootBC(6)
            integer(c_int), intent(out) :: iterLBM_out
        end function
        integer(c_int) function fsilbm_ibm_interaction_force_begin(h, nbody, nelmts, Exyz, Evel, &
                Ea, restencil, dt, ntolLBM, dtolLBM, rootBC) bind(C, name='fsilbm_ibm_interaction_force_begin')
            import :: c_double, c_int, c_ptr
            integer(c_int), value :: h
            integer(c_int), value :: nbody
            integer(c_int), intent(in) :: nelmts(*)
            type(c_ptr), intent(in) :: Exyz(*)
            type(c_ptr), intent(in) :: Evel(*)
            type(c_ptr), intent(in) :: Ea(*)
            integer(c_int), intent(in) :: restencil(*)
            real(c_double), value :: dt
            integer(c_int), value :: ntolLBM
            real(c_double), value :: dtolLBM
            integer(c_int), intent(in) :: rootBC(6)
        end function
        integer(c_int) function fsilbm_ibm_interaction_force_wait(h, nbody, &
                Eforce, iterLBM_out) bind(C, name='fsilbm_ibm_interaction_force_wait')
            import :: c_int, c_ptr
            integer(c_int), value :: h
            integer(c_int), value :: nbody
            type(c_ptr), intent(in) :: Eforce(*)
            integer(c_int), intent(out) :: iterLBM_out
        end function
        integer(c_int) function fsilbm_ibm_body_status(h, nbody, status) bind(C, name='fsilbm_ibm_body_status')
            import :: c_int
            integer(c_int), value :: h
            integer(c_int), value :: nbody
            integer(c_int), intent(inout) :: status(*)
        end function
        integer(c_int) function fsilbm_ibm_download_stencil(h, body, Ei, Ew) bind(C, name='fsilbm_ibm_download_stencil')
            import :: c_float, c_int, c_short
            integer(c_int), value :: h
            integer(c_int), value :: body
            integer(c_short), intent(inout) :: Ei(*)
            real(c_float), intent(inout) :: Ew(*)
        end function
        integer(c_int) function fsilbm_pair_create(father, son, interpolateScheme, pair) bind(C, name='fsilbm_pair_create')
            import :: c_int
            integer(c_int), value :: father
            integer(c_int), value :: son
            integer(c_int), value :: interpolateScheme
            integer(c_int), intent(out) :: pair
        end function
        integer(c_int) function fsilbm_pair_create_remote(father, owner_rank, pair) bind(C, name='fsilbm_pair_create_remote')
            import :: c_int
            integer(c_int), value :: father
            integer(c_int), value :: owner_rank
            integer(c_int), intent(out) :: pair
        end function
        integer(c_int) function fsilbm_pair_destroy(pair) bind(C, name='fsilbm_pair_destroy')
            import :: c_int
            integer(c_int), value :: pair
        end function
        integer(c_int) function fsilbm_pair_info(pair, out) bind(C, name='fsilbm_pair_info')
            import :: c_int
            integer(c_int), value :: pair
            integer(c_int), intent(inout) :: out(36)
        end function
        integer(c_int) function fsilbm_pair_extract_layer(pair, time) bind(C, name='fsilbm_pair_extract_layer')
            import :: c_int
            integer(c_int), value :: pair
            integer(c_int), value :: time
        end function
        integer(c_int) function fsilbm_pair_father_to_son(pair, n_timeStep) bind(C, name='fsilbm_pair_father_to_son')
            import :: c_int
            integer(c_int), value :: pair
            integer(c_int), value :: n_timeStep
        end function
        integer(c_int) function fsilbm_pair_son_to_father(pair) bind(C, name='fsilbm_pair_son_to_father')
            import :: c_int
            integer(c_int), value :: pair
        end function
        integer(c_int) function fsilbm_comm_unique_id(id) bind(C, name='fsilbm_comm_unique_id')
            import :: c_char, c_int
            character(kind=c_char), intent(inout) :: id(128)
        end function
        integer(c_int) function fsilbm_comm_init(rank, nranks, id) bind(C, name='fsilbm_comm_init')
            import :: c_char, c_int
            integer(c_int), value :: rank
            integer(c_int), value :: nranks
            character(kind=c_char), intent(in) :: id(128)
        end function
        integer(c_int) function fsilbm_comm_finalize() bind(C, name='fsilbm_comm_finalize')
            import :: c_int
        end function
        integer(c_int) function fsilbm_block_halo_transport(h, mode) bind(C, name='fsilbm_block_halo_transport')
            import :: c_int
            integer(c_int), value :: h
            integer(c_int), intent(out) :: mode
        end function
    end interface
end module fsilbm_c

module fsilbm_gpu
    use, intrinsic :: iso_c_binding
    use fsilbm_c
    implicit none
    private
    public :: fsilbm_check, gpu_init, gpu_finalize, gpu_set_option, gpu_allocate_block, gpu_allocate_slab, gpu_free_block, &
              gpu_initialise_block, gpu_upload_fIn, gpu_refresh_host_fIn, gpu_refresh_host_macro, gpu_refresh_host_macro_async, &
              gpu_refresh_host_wait, gpu_refresh_host_tau_all, gpu_update_volume_force, gpu_set_boundary_conditions, &
              gpu_collide_stream, gpu_sync, gpu_interaction_force, gpu_interaction_force_begin, gpu_interaction_force_wait, &
              gpu_body_status, gpu_download_stencil, gpu_field_stat, gpu_write_flow_window, gpu_write_flow_window_async, gpu_turbulent_statistic, &
              gpu_fluid_flux, gpu_probe_velocity, gpu_pair_create, gpu_pair_create_remote, gpu_pair_free, gpu_pair_info, gpu_extract_interpolate_layer, &
              gpu_interpolation_father_to_son, gpu_deliver_son_to_father, gpu_comm_unique_id, gpu_comm_init, gpu_comm_finalize, &
              gpu_halo_transport, gpu_pass_macro, gpu_pass_reset_volume_force, gpu_pass_add_volume_force, gpu_pass_collision, &
              gpu_pass_halfway_bc_set, gpu_pass_streaming, gpu_download_fields, gpu_upload_fields, gpu_launch_count, &
              gpu_ibm_early_count, gpu_trace_dump, gpu_block_stream
    integer(c_int), allocatable, public :: gpu_handle(:)    ! one library handle per LBMblks(i)
    integer(c_int), allocatable, public :: gpu_pair(:)      ! one pair handle per CommPair (blockTree(i)%comm(j), numbered by the driver)
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

    ! Fortran string -> C string
    function cstr(s) result(c)
        character(len=*), intent(in) :: s
        character(kind=c_char) :: c(len_trim(s)+1)
        integer :: i
        do i = 1, len_trim(s)
            c(i) = s(i:i)
        enddo
        c(len_trim(s)+1) = c_null_char
    end function

    ! ---- process level -------------------------------------------------------------------------------------------------
    ! beside omp_set_num_threads, main.f90:36: one process per GPU
    subroutine gpu_init(device)
        integer, intent(in) :: device
        call fsilbm_check(fsilbm_init(device))
    end subroutine
    subroutine gpu_finalize()
        call fsilbm_check(fsilbm_finalize())
    end subroutine
    ! tuning switches of include/fsilbm.h ("ibm_ordered", "halo", ...)
    subroutine gpu_set_option(key, val)
        character(len=*), intent(in) :: key
        integer, intent(in) :: val
        call fsilbm_check(fsilbm_set_option(cstr(key), val))
    end subroutine
    function gpu_launch_count() result(n)
        integer(c_long_long) :: n
        n = fsilbm_launch_count()
    end function
    function gpu_ibm_early_count() result(n)
        integer(c_long_long) :: n
        n = fsilbm_ibm_early_count()
    end function
    subroutine gpu_trace_dump(path)
        character(len=*), intent(in) :: path
        call fsilbm_check(fsilbm_trace_dump(cstr(path)))
    end subroutine

    ! ---- multi-GPU set-up (no reference counterpart: the reference is one process) ----------------------------------------
    ! rank 0 calls gpu_comm_unique_id and hands id to the other processes (MPI_Bcast, a file, ...); every rank calls gpu_comm_init
    subroutine gpu_comm_unique_id(id)
        character(kind=c_char), intent(out) :: id(128)
        call fsilbm_check(fsilbm_comm_unique_id(id))
    end subroutine
    subroutine gpu_comm_init(rank, nranks, id)
        integer, intent(in) :: rank, nranks
        character(kind=c_char), intent(in) :: id(128)
        call fsilbm_check(fsilbm_comm_init(rank, nranks, id))
    end subroutine
    subroutine gpu_comm_finalize()
        call fsilbm_check(fsilbm_comm_finalize())
    end subroutine
    subroutine gpu_halo_transport(iblock, mode)
        integer, intent(in) :: iblock
        integer, intent(out) :: mode
        call fsilbm_check(fsilbm_block_halo_transport(gpu_handle(iblock), mode))
    end subroutine

    ! ---- blocks -----------------------------------------------------------------------------------------------------------
    ! after allocate_fluid_ (FluidDomain.f90:378-408), main.f90:40.  nu etc. come from the global `flow`.  One GPU: one slab.
    subroutine gpu_allocate_block(iblock, xDim, yDim, zDim, dh, xmin, ymin, zmin, BndConds, iCollidModel, params, &
                                  nu, denIn, uvwIn, shearRateIn, velocityKind, volumeForceIn, volumeForceAmp, &
                                  volumeForceFreq, volumeForcePhi, Uref)
        integer, intent(in) :: iblock, xDim, yDim, zDim, BndConds(6), iCollidModel, velocityKind
        real(8), intent(in) :: dh, xmin, ymin, zmin, params(10), nu, denIn, uvwIn(3), shearRateIn(3), volumeForceIn(3)
        real(8), intent(in) :: volumeForceAmp, volumeForceFreq, volumeForcePhi, Uref
        call gpu_allocate_slab(iblock, xDim, yDim, zDim, 0, xDim, dh, xmin, ymin, zmin, BndConds, iCollidModel, params, &
                               nu, denIn, uvwIn, shearRateIn, velocityKind, volumeForceIn, volumeForceAmp, &
                               volumeForceFreq, volumeForcePhi, Uref)
    end subroutine

    ! the same for an x-slab run: this process owns the global planes xOffset+1 .. xOffset+xLocal of LBMblks(iblock);
    ! every rank creates its blocks in the same order (creation is collective)
    subroutine gpu_allocate_slab(iblock, xDim, yDim, zDim, xOffset, xLocal, dh, xmin, ymin, zmin, BndConds, iCollidModel, params, &
                                 nu, denIn, uvwIn, shearRateIn, velocityKind, volumeForceIn, volumeForceAmp, &
                                 volumeForceFreq, volumeForcePhi, Uref)
        integer, intent(in) :: iblock, xDim, yDim, zDim, xOffset, xLocal, BndConds(6), iCollidModel, velocityKind
        real(8), intent(in) :: dh, xmin, ymin, zmin, params(10), nu, denIn, uvwIn(3), shearRateIn(3), volumeForceIn(3)
        real(8), intent(in) :: volumeForceAmp, volumeForceFreq, volumeForcePhi, Uref
        type(fsilbm_flow) :: cf
        cf%nu = nu; cf%denIn = denIn; cf%uvwIn = uvwIn; cf%shearRateIn = shearRateIn; cf%velocityKind = velocityKind
        cf%volumeForceIn = volumeForceIn; cf%volumeForceAmp = volumeForceAmp; cf%volumeForceFreq = volumeForceFreq
        cf%volumeForcePhi = volumeForcePhi; cf%Uref = Uref
        call fsilbm_check(fsilbm_block_create(xDim, yDim, zDim, xOffset, xLocal, dh, xmin, ymin, zmin, BndConds, iCollidModel, &
                                              params, cf, gpu_handle(iblock)))
    end subroutine

    subroutine gpu_free_block(iblock)
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_destroy(gpu_handle(iblock)))
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
    ! the same without waiting (den, uuu page-locked); what the reference gets from fork()ing its writer (FluidDomain.f90:1702)
    subroutine gpu_refresh_host_macro_async(iblock, den, uuu)
        integer, intent(in) :: iblock
        real(8), intent(out) :: den(*), uuu(*)
        call fsilbm_check(fsilbm_block_download_macro_async(gpu_handle(iblock), den, uuu))
    end subroutine
    subroutine gpu_refresh_host_wait(iblock)
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_download_wait(gpu_handle(iblock)))
    end subroutine
    ! tau_all(z,y,x) of the LES models (FluidDomain.f90:51,1279,1422,1505)
    subroutine gpu_refresh_host_tau_all(iblock, tau_all)
        integer, intent(in) :: iblock
        real(8), intent(out) :: tau_all(*)
        call fsilbm_check(fsilbm_block_download_tau_all(gpu_handle(iblock), tau_all))
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
    subroutine gpu_sync(iblock)
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_sync(gpu_handle(iblock)))
    end subroutine
    subroutine gpu_block_stream(iblock, stream)
        integer, intent(in) :: iblock
        type(c_ptr), intent(out) :: stream
        call fsilbm_check(fsilbm_block_stream(gpu_handle(iblock), stream))
    end subroutine

    ! ---- immersed boundary -----------------------------------------------------------------------------------------------
    ! replaces calculate_interaction_force (Solidbody.f90:869-918) minus lodFlow assembly.  The caller keeps
    ! UpdatePosVelArea (Solidbody.f90:597-600) before and the nodal-load half of FluidVolumeForce_ (:945-967) after.
    ! Exyz(b) etc. are c_loc() of VBodies(iFish)%v_Exyz / v_Evel / v_Ea / v_Eforce of the carried bodies.
    subroutine gpu_interaction_force(iblock, nbody, nelmts, Exyz, Evel, Ea, Eforce, restencil, dt, ntolLBM, dtolLBM, rootBC, iterLBM)
        integer, intent(in) :: iblock, nbody, nelmts(nbody), restencil(nbody), ntolLBM, rootBC(6)
        type(c_ptr), intent(in) :: Exyz(nbody), Evel(nbody), Ea(nbody), Eforce(nbody)
        real(8), intent(in) :: dt, dtolLBM
        integer, intent(out) :: iterLBM
        call fsilbm_check(fsilbm_ibm_interaction_force(gpu_handle(iblock), nbody, nelmts, Exyz, Evel, Ea, Eforce, restencil, dt, &
                                                       ntolLBM, dtolLBM, rootBC, iterLBM))
    end subroutine
    ! the same in two halves: _begin at Solidbody.f90:601, gpu_collide_stream right after it, _wait before the nodal loads (:911)
    subroutine gpu_interaction_force_begin(iblock, nbody, nelmts, Exyz, Evel, Ea, restencil, dt, ntolLBM, dtolLBM, rootBC)
        integer, intent(in) :: iblock, nbody, nelmts(nbody), restencil(nbody), ntolLBM, rootBC(6)
        type(c_ptr), intent(in) :: Exyz(nbody), Evel(nbody), Ea(nbody)
        real(8), intent(in) :: dt, dtolLBM
        call fsilbm_check(fsilbm_ibm_interaction_force_begin(gpu_handle(iblock), nbody, nelmts, Exyz, Evel, Ea, restencil, dt, &
                                                             ntolLBM, dtolLBM, rootBC))
    end subroutine
    subroutine gpu_interaction_force_wait(iblock, nbody, Eforce, iterLBM)
        integer, intent(in) :: iblock, nbody
        type(c_ptr), intent(in) :: Eforce(nbody)
        integer, intent(out) :: iterLBM
        call fsilbm_check(fsilbm_ibm_interaction_force_wait(gpu_handle(iblock), nbody, Eforce, iterLBM))
    end subroutine
    subroutine gpu_body_status(iblock, nbody, status)
        integer, intent(in) :: iblock, nbody
        integer, intent(out) :: status(nbody)
        call fsilbm_check(fsilbm_ibm_body_status(gpu_handle(iblock), nbody, status))
    end subroutine
    ! v_Ei / v_Ew of body ibody (0-based position in the last call's list) as the library holds them (int16 / real(4))
    subroutine gpu_download_stencil(iblock, ibody, Ei, Ew)
        integer, intent(in) :: iblock, ibody
        integer(c_short), intent(out) :: Ei(*)
        real(c_float), intent(out) :: Ew(*)
        call fsilbm_check(fsilbm_ibm_download_stencil(gpu_handle(iblock), ibody, Ei, Ew))
    end subroutine

    ! ---- output and diagnostics from the device state --------------------------------------------------------------------------
    ! replaces ComputeFieldStat_ (FluidDomain.f90:1739-1768), main.f90:150
    subroutine gpu_field_stat(iblock, ncell, stat)
        integer, intent(in) :: iblock
        real(8), intent(in) :: ncell
        real(8), intent(out) :: stat(6)
        call fsilbm_check(fsilbm_block_field_stat(gpu_handle(iblock), stat))
        stat(1:3) = dsqrt(stat(1:3) / ncell)
    end subroutine
    ! replaces the staging loops of write_flow_ (FluidDomain.f90:1640-1699): OUTtmp as real(4), then the caller writes the file
    subroutine gpu_write_flow_window(iblock, offsetOutput, outputtype, outtmp)
        integer, intent(in) :: iblock, offsetOutput, outputtype
        real(c_float), intent(out) :: outtmp(*)
        call fsilbm_check(fsilbm_block_write_flow_window(gpu_handle(iblock), offsetOutput, outputtype, outtmp))
    end subroutine
    ! the same without waiting (outtmp page-locked): filled while the following steps run, valid after gpu_refresh_host_wait --
    ! the overlap the reference gets from fork()ing its writer (FluidDomain.f90:1702)
    subroutine gpu_write_flow_window_async(iblock, offsetOutput, outputtype, outtmp)
        integer, intent(in) :: iblock, offsetOutput, outputtype
        real(c_float), intent(out) :: outtmp(*)
        call fsilbm_check(fsilbm_block_write_flow_window_async(gpu_handle(iblock), offsetOutput, outputtype, outtmp))
    end subroutine
    ! replaces calculate_turbulent_statistic_ (FluidDomain.f90:1147-1172), main.f90:108
    subroutine gpu_turbulent_statistic(iblock, step, step_s)
        integer, intent(in) :: iblock, step, step_s
        call fsilbm_check(fsilbm_block_turbulent_statistic(gpu_handle(iblock), step, step_s))
    end subroutine
    ! replaces the plane sums of write_fluid_flux (FluidDomain.f90:2019-2046), main.f90:135
    subroutine gpu_fluid_flux(iblock, flux)
        integer, intent(in) :: iblock
        real(8), intent(out) :: flux(3)
        call fsilbm_check(fsilbm_block_fluid_flux(gpu_handle(iblock), flux))
    end subroutine
    ! replaces grid_value_interpolation for the probes of write_fluid_information (FlowCondition.f90:195-222, Util.f90:123-157)
    subroutine gpu_probe_velocity(iblock, n, coords, velocity)
        integer, intent(in) :: iblock, n
        real(8), intent(in) :: coords(3,n)
        real(8), intent(out) :: velocity(3,n)
        call fsilbm_check(fsilbm_block_probe_velocity(gpu_handle(iblock), n, coords, velocity))
    end subroutine

    ! ---- grid refinement: the call points of LBMBlockComm.f90 ------------------------------------------------------------------
    ! build_blocks_comunication (LBMBlockComm.f90:32-96) for blockTree(father)%comm(j): ipair is the driver's number of that pair
    subroutine gpu_pair_create(ipair, ifather, ison, interpolateScheme)
        integer, intent(in) :: ipair, ifather, ison, interpolateScheme
        call fsilbm_check(fsilbm_pair_create(gpu_handle(ifather), gpu_handle(ison), interpolateScheme, gpu_pair(ipair)))
    end subroutine
    ! slab runs, a son across a slab interface: the rank next to the son's owner registers the pair (same ipair numbering); its
    ! father block then follows the owner's transfers through device-side flags.  The transfer wrappers below are no-ops there.
    subroutine gpu_pair_create_remote(ipair, ifather, owner_rank)
        integer, intent(in) :: ipair, ifather, owner_rank
        call fsilbm_check(fsilbm_pair_create_remote(gpu_handle(ifather), owner_rank, gpu_pair(ipair)))
    end subroutine
    subroutine gpu_pair_free(ipair)
        integer, intent(in) :: ipair
        call fsilbm_check(fsilbm_pair_destroy(gpu_pair(ipair)))
    end subroutine
    ! sds, s, f, si, fi (CommPair, LBMBlockComm.f90:11-18) and the extents, for check_blocks_params' messages
    subroutine gpu_pair_info(ipair, info)
        integer, intent(in) :: ipair
        integer, intent(out) :: info(36)
        call fsilbm_check(fsilbm_pair_info(gpu_pair(ipair), info))
    end subroutine
    ! replaces extract_interpolate_layer (LBMBlockComm.f90:340-505) at :290 (time = 1) and :305 (time = 2), once per son
    subroutine gpu_extract_interpolate_layer(ipair, time)
        integer, intent(in) :: ipair, time
        call fsilbm_check(fsilbm_pair_extract_layer(gpu_pair(ipair), time))
    end subroutine
    ! replaces interpolation_father_to_son (LBMBlockComm.f90:655-806) at :313
    subroutine gpu_interpolation_father_to_son(ipair, n_timeStep)
        integer, intent(in) :: ipair, n_timeStep
        call fsilbm_check(fsilbm_pair_father_to_son(gpu_pair(ipair), n_timeStep))
    end subroutine
    ! replaces deliver_son_to_father (LBMBlockComm.f90:546-653) at :315
    subroutine gpu_deliver_son_to_father(ipair)
        integer, intent(in) :: ipair
        call fsilbm_check(fsilbm_pair_son_to_father(gpu_pair(ipair)))
    end subroutine

    ! ---- the reference's call granularity, pass by pass (parity runs; 3.5x the traffic of gpu_collide_stream) -------------------
    subroutine gpu_pass_macro(iblock)                 ! calculate_macro_quantities_, LBMBlockComm.f90:285
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_pass_macro(gpu_handle(iblock)))
    end subroutine
    subroutine gpu_pass_reset_volume_force(iblock)    ! ResetVolumeForce_, :286
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_pass_reset_volume_force(gpu_handle(iblock)))
    end subroutine
    subroutine gpu_pass_add_volume_force(iblock)      ! add_volume_force_, :288
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_pass_add_volume_force(gpu_handle(iblock)))
    end subroutine
    subroutine gpu_pass_collision(iblock)             ! collision_, :293
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_pass_collision(gpu_handle(iblock)))
    end subroutine
    subroutine gpu_pass_halfway_bc_set(iblock)        ! halfwayBCset_, :296
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_pass_halfway_bc_set(gpu_handle(iblock)))
    end subroutine
    subroutine gpu_pass_streaming(iblock)             ! streaming_, :299
        integer, intent(in) :: iblock
        call fsilbm_check(fsilbm_block_pass_streaming(gpu_handle(iblock)))
    end subroutine
    subroutine gpu_download_fields(iblock, den, uuu, force)
        integer, intent(in) :: iblock
        real(8), intent(out) :: den(*), uuu(*), force(*)
        call fsilbm_check(fsilbm_block_download_fields(gpu_handle(iblock), den, uuu, force))
    end subroutine
    subroutine gpu_upload_fields(iblock, den, uuu, force)
        integer, intent(in) :: iblock
        real(8), intent(in) :: den(*), uuu(*), force(*)
        call fsilbm_check(fsilbm_block_upload_fields(gpu_handle(iblock), den, uuu, force))
    end subroutine

end module fsilbm_gpu
