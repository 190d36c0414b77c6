! A small Fortran driver over the ISO_C_BINDING shim (fortran/fsilbm_gpu.f90), written the way INTEGRATION.md edits the reference's
! main loop: block set-up, start-up sequence of main.f90:50-64, time loop with update_volume_force / interaction force / fused update,
! read-back.  tests/test_gpu_fortran_shim.py EXECUTES it with the Fortran interpreter of oracle/ftn against libfsilbm_b200.so (this
! image has no Fortran compiler) and compares the populations and marker forces it writes with the CPU oracle.
program drive_shim
    use, intrinsic :: iso_c_binding
    use fsilbm_c
    use fsilbm_gpu
    implicit none
    integer :: xDim, yDim, zDim, nsteps, nmark, ntolLBM, n, iterLBM, model, i
    integer :: BndConds(6), nelmts(1), restencil(1), iters_total
    real(8) :: dh, nu, denIn, uvwIn(3), shear(3), vforce(3), Uref, params(10), dtolLBM
    real(8) :: tau, Omega, Omega2, vf(3), stat(6)
    real(8), allocatable, target :: fIn(:,:,:,:), Exyz(:,:), Evel(:,:), Ea(:), Eforce(:,:)
    type(c_ptr) :: pExyz(1), pEvel(1), pEa(1), pEforce(1)
    open(11, file='case.txt', status='old', action='read')
    read(11,*) xDim, yDim, zDim, nsteps, model
    read(11,*) BndConds(1:6)
    read(11,*) nu, denIn, Uref
    read(11,*) uvwIn(1:3)
    read(11,*) shear(1:3)
    read(11,*) vforce(1:3)
    read(11,*) nmark, ntolLBM, dtolLBM
    close(11)
    dh = 1.0d0
    params = 0.0d0
    allocate(fIn(zDim,yDim,xDim,0:18))
    open(13, file='f0.bin', form='unformatted', status='old', access='stream')
    read(13) fIn
    close(13)
    if (nmark > 0) then
        allocate(Exyz(3,nmark), Evel(3,nmark), Ea(nmark), Eforce(3,nmark))
        open(13, file='markers.bin', form='unformatted', status='old', access='stream')
        read(13) Exyz, Evel, Ea
        close(13)
        pExyz(1) = c_loc(Exyz); pEvel(1) = c_loc(Evel); pEa(1) = c_loc(Ea); pEforce(1) = c_loc(Eforce)
        nelmts(1) = nmark
    endif

    call gpu_init(0)
    call gpu_set_option('ibm_early', 1)
    allocate(gpu_handle(1), gpu_pair(1))
    call gpu_allocate_block(1, xDim, yDim, zDim, dh, 0.0d0, 0.0d0, 0.0d0, BndConds, model, params, nu, denIn, uvwIn, shear, 0, &
                            vforce, 0.0d0, 0.0d0, 0.0d0, Uref)
    call gpu_initialise_block(1, 0.0d0, tau, Omega, Omega2)               ! main.f90:50
    call gpu_upload_fIn(1, fIn)                                             ! main.f90:58 (check_is_continue)
    call gpu_update_volume_force(1, 0.0d0, vf)                              ! main.f90:62
    call gpu_set_boundary_conditions(1)                                     ! main.f90:63
    iters_total = 0
    do n = 1, nsteps                                                        ! main.f90:93-108
        call gpu_update_volume_force(1, dble(n)*dh, vf)                     ! LBMBlockComm.f90:283
        if (nmark > 0) then
            restencil(1) = 0
            if (n == 1) restencil(1) = 1
            call gpu_interaction_force(1, 1, nelmts, pExyz, pEvel, pEa, pEforce, restencil, dh, ntolLBM, dtolLBM, BndConds, iterLBM)
            iters_total = iters_total + iterLBM
        endif
        call gpu_collide_stream(1)                                          ! LBMBlockComm.f90:285-303
    enddo
    call gpu_refresh_host_fIn(1, fIn)                                       ! main.f90:118
    call gpu_field_stat(1, dble(xDim)*dble(yDim)*dble(zDim), stat)          ! main.f90:150
    open(14, file='f1.bin', form='unformatted', access='stream')
    write(14) fIn
    close(14)
    if (nmark > 0) then
        open(14, file='force.bin', form='unformatted', access='stream')
        write(14) Eforce
        close(14)
    endif
    write(*,'(A,F10.6,A,I6)') ' tau ', tau, ' IBM iterations ', iters_total
    write(*,'(A,F18.12)') ' FIELDSTAT L2 u ', stat(1)
    call gpu_free_block(1)
    call gpu_finalize()
end program drive_shim
