/*
 * fsilbm.h -- C ABI of libfsilbm_b200.so: the B200 (sm_100a) implementation of the FSILBM3D
 * hot path (fp64 D3Q19 collide-stream-boundary update + immersed-boundary coupling).
 *
 * The reference has no FFI for this path: the Fortran driver reaches it through type-bound
 * procedures on the module globals LBMblks(:) (FluidDomain.f90:57) and VBodies(:)
 * (Solidbody.f90:69).  Each entry point below names the reference procedure (file:line,
 * relative to /root/reference/src) whose call it replaces.  fortran/fsilbm_gpu.f90 holds the
 * ISO_C_BINDING interfaces for exactly these symbols; INTEGRATION.md shows where the driver
 * calls them.
 *
 * Conventions
 *  - plain C: ints, doubles, pointers; no C++/torch types.  Fortran passes scalars BY VALUE
 *    through the bind(C) interfaces.
 *  - every function returns 0 on success and a non-zero code otherwise; the message is in
 *    fsilbm_last_error().  The reference's convention is "write(*,*) msg; stop"
 *    (e.g. FluidDomain.f90:704, Solidbody.f90:850); the shim turns a non-zero return into that.
 *  - host arrays use the reference's Fortran layout: fIn(z,y,x,0:18) = C [19][X][Y][Z], z fastest
 *    (FluidDomain.f90:384); uuu/force(z,y,x,1:3) = C [3][X][Y][Z]; den(z,y,x) = C [X][Y][Z];
 *    marker arrays v_Exyz(3,n) = C [n][3] (Solidbody.f90:1068-1069).
 *  - callers are single-threaded (all reference call points are in serial regions).
 *  - there is no CPU fallback: every call fails with FSILBM_ERR_CUDA if no sm_100-class
 *    device is usable.
 */
#ifndef FSILBM_H
#define FSILBM_H

#ifdef __cplusplus
extern "C" {
#endif

#define FSILBM_OK 0
#define FSILBM_ERR_ARG 1       /* bad argument / unknown handle                                  */
#define FSILBM_ERR_CUDA 2      /* CUDA runtime failure (incl. no device)                         */
#define FSILBM_ERR_BC 3        /* 'has no such boundary condition' (FluidDomain.f90:704)         */
#define FSILBM_ERR_STENCIL 4   /* 'index out of xmin/xmax bound'   (Solidbody.f90:850,861)       */
#define FSILBM_ERR_NAN 5       /* 'Nan found in PenaltyForce'      (Solidbody.f90:1028-1031)     */
#define FSILBM_ERR_MODEL 6     /* collision model not provided (12/13 are broken upstream)       */
#define FSILBM_ERR_COMM 7      /* NCCL failure                                                   */
#define FSILBM_ERR_PERIODIC 8  /* 'Periodic boundaries must apper in pairs' (FluidDomain.f90:120) */

typedef int fsilbm_handle;

/* The slice of FlowCondType the hot path reads (FlowCondition.f90:11-27). */
typedef struct fsilbm_flow {
    double nu;                /* flow%nu, set at Solidbody.f90:282                 */
    double denIn;             /* flow%denIn                                        */
    double uvwIn[3];          /* flow%uvwIn                                        */
    double shearRateIn[3];    /* flow%shearRateIn                                  */
    int    velocityKind;      /* 0 shear (FluidDomain.f90:1807), 2 oscillatory (:1821) */
    double volumeForceIn[3];  /* flow%volumeForceIn                                */
    double volumeForceAmp, volumeForceFreq, volumeForcePhi; /* FluidDomain.f90:1177 */
    double Uref;              /* flow%Uref (IBM tolerance, FIELDSTAT)              */
} fsilbm_flow;

/* ---- process-level ------------------------------------------------------------------------ */

/* Select the CUDA device of this process (one process per GPU).  No reference counterpart
 * (the reference's only set-up call is omp_set_num_threads, main.f90:36). */
int fsilbm_init(int device);
int fsilbm_finalize(void);
const char *fsilbm_last_error(void);
/* Number of kernels this library has launched since fsilbm_init (bench.py's gpu_launches). */
long long fsilbm_launch_count(void);
/* Number of fsilbm_ibm_interaction_force calls that ran beside the tail of the previous collide-stream update ("early IBM":
 * fsilbm_block_collide_stream updates the x-planes around the bodies first, so the next interaction-force call needs to wait
 * for those planes only; option "ibm_early" = 0 turns it off).  Diagnostics/tests. */
long long fsilbm_ibm_early_count(void);
/* Diagnostics: with option "trace" = 1 a timing event is recorded after every launch of a step; this writes the marks
 * (name, stream, device completion time, host issue time; microseconds) as CSV and clears them. */
int fsilbm_trace_dump(const char *path);
/* Tuning/testing switches, no reference counterpart.
 *   "trace"                   1 = record the launch marks that fsilbm_trace_dump writes
 *   "force_ghost"             1 = stream through the ghost planes even on one rank (tests the slab path)
 *   "halo"                    1 (default) peer-memory halo over NVLink, 0 ncclSend/ncclRecv (see fsilbm_block_halo_transport)
 *   "halo_timeout_s"          how long a rank waits for a neighbour (halo flags, IBM loop-control mailbox) before it reports
 *                             FSILBM_ERR_COMM (default 120)
 *   "ibm_ordered"             1 (default) interpolation and spreading keep the reference's serial summation order (bit-identical to the
 *                             serial reference, reproducible), 0 warp shuffles + fp64 atomics (round-off differences;
 *                             single-GPU comparison arm, refused on slab runs)
 *   "ibm_single_launch"       1 (default) calculate_interaction_force is one cooperative kernel (on slab runs with the loop control
 *                             exchanged through peer memory), 0 one kernel per phase (slab runs: ncclAllReduce of the loop control)
 *   "ibm_early"               1 (default) fsilbm_block_collide_stream updates the x-planes around the bodies first so that the next
 *                             fsilbm_ibm_interaction_force runs beside the rest of the update, 0 strictly one after the other
 *   "update_split"            1 (default) on one GPU the launch over those planes goes to a high-priority stream of its own behind the
 *                             interaction-force call while the rest of the update starts at once on the compute stream (the device never
 *                             waits for an iteration that outlasts the previous update), 0 both launches queued on the compute stream
 *   "ibm_early_blocks_per_sm" 0 (default): the cooperative IBM grid that shares the SMs with that update is sized per call -- one block per
 *                             SM unless the iteration would outlast the rest of the update (a large body in a small block), then up to 4;
 *                             1..4 fixes it
 *   "ibm_early_blocks"        > 0: that grid as an absolute number of blocks instead (0 = use the per-SM figure)
 *   "ibm_force_exchange"      slab runs: 1 (default) every rank passes the same body list and gets every force back, 0 per-rank lists
 *                             (see fsilbm_ibm_body_status below) */
int fsilbm_set_option(const char *key, int value);

/* ---- fluid block: replaces type LBMBlock's procedures --------------------------------------- */

/* read_fuild_blocks (FluidDomain.f90:76-105) + allocate_fluid_ (:378-408).
 * xDim is the GLOBAL x extent; this process owns global planes [xOffset, xOffset+xLocal)
 * (x-slab decomposition; pass xOffset=0, xLocal=xDim for one GPU). */
int fsilbm_block_create(int xDim, int yDim, int zDim, int xOffset, int xLocal,
                        double dh, double xmin, double ymin, double zmin,
                        const int BndConds[6], int iCollidModel, const double params[10],
                        const fsilbm_flow *flow, fsilbm_handle *out);
int fsilbm_block_destroy(fsilbm_handle h);

/* initialise_ (FluidDomain.f90:433-545): tau/Omega/Omega2/MRT matrices, f = f_eq(denIn, U(x)). */
int fsilbm_block_initialise(fsilbm_handle h, double time);
/* scalars derived at initialise: what = 0 tau, 1 Omega, 2 Omega2 */
int fsilbm_block_get(fsilbm_handle h, int what, double *value);

/* check_is_continue / read_continue_ (FluidDomain.f90:128-237,1779-1789) hand fIn over with
 * these; fIn is the LOCAL slab [19][xLocal][Y][Z]. */
int fsilbm_block_upload_fIn(fsilbm_handle h, const double *fIn);
/* write_continue_ (FluidDomain.f90:1770-1777). */
int fsilbm_block_download_fIn(fsilbm_handle h, double *fIn);

/* LBMblks(:)%blktime = time (main.f90:97). */
int fsilbm_block_set_time(fsilbm_handle h, double blktime);
/* update_volume_force_ (FluidDomain.f90:1174-1180); F is evaluated on the host with libm's sin,
 * as the reference does.  volumeForce_out may be NULL. */
int fsilbm_block_update_volume_force(fsilbm_handle h, double volumeForce_out[3]);

/* calculate_macro_quantities_ (FluidDomain.f90:1128-1145) for the writers/probes/FIELDSTAT:
 * computes den, uuu of the local slab from the current fIn and copies them to the host.
 * Either pointer may be NULL.  (Inside a step the library derives den/uuu in registers.) */
int fsilbm_block_download_macro(fsilbm_handle h, double *den, double *uuu);
/* The same without waiting: den/uuu are computed into device staging fields behind the work already queued and copied to the
 * host (PINNED memory, or the copy is not asynchronous) on the library's copy stream while later steps run -- what the
 * reference gets from fork()ing its writer (FluidDomain.f90:1702).  The host arrays are valid after
 * fsilbm_block_download_wait (or fsilbm_block_sync); a second read-back waits for the first. */
int fsilbm_block_download_macro_async(fsilbm_handle h, double *den, double *uuu);
int fsilbm_block_download_wait(fsilbm_handle h);
/* tau_all(z,y,x) (FluidDomain.f90:51): the local relaxation time the LES models 11/14/15 write during collision
 * (:1279,1422,1505); the block's tau everywhere for the other models.  C [X][Y][Z] of the local slab. */
int fsilbm_block_download_tau_all(fsilbm_handle h, double *tau_all);
/* ComputeFieldStat_ (FluidDomain.f90:1739-1768) on the local slab: out = sum(u^2)/Uref^2 for
 * u,v,w then max|u|/Uref for u,v,w (the caller finishes sqrt(sum/N) after reducing over ranks). */
int fsilbm_block_field_stat(fsilbm_handle h, double out[6]);

/* Output and diagnostics computed from the device state, so the host never needs full den/uuu arrays:
 *  - write_flow_ staging (FluidDomain.f90:1640-1699): fills `out` = OUTtmp as real(4), C [nfields][nx][ny][nz] over the
 *    window [offsetOutput, dim-offsetOutput) of the local slab; fields p,u,v,w (outputtype 1) or those + <u>,<v>,<w>,
 *    <uu>,<vv>,<ww>,<uv>,<uw>,<vw> (13 fields, outputtype >= 2).  The caller writes the file header and these bytes
 *    (and forks if it wants to: the buffer is plain host memory).  `_async` returns once the work is queued: `out` (PINNED
 *    memory) is filled on the copy stream while later steps run and is valid after fsilbm_block_download_wait -- the
 *    reference's own overlap of output and computation (its fork()ed writer, :1702), at half the bytes of den/uuu in fp64.
 *  - calculate_turbulent_statistic_ (:1147-1172): running means kept on the device (real(4) 1/n quirk of :1153 kept).
 *  - write_fluid_flux (:2019-2046): out = un-normalised fluxIn, fluxMid, fluxOut of the planes this rank owns
 *    (sum over ranks, then divide by denIn*Uref*Zref*Yref as :2051-2054 does).
 *  - write_fluid_information / grid_value_interpolation (FlowCondition.f90:195-222, Util.f90:123-157): trilinear
 *    velocity at n probe points, C [n][3]; contributions of corners this rank owns (sum over ranks). */
int fsilbm_block_write_flow_window(fsilbm_handle h, int offsetOutput, int outputtype, float *out);
int fsilbm_block_write_flow_window_async(fsilbm_handle h, int offsetOutput, int outputtype, float *out);
int fsilbm_block_turbulent_statistic(fsilbm_handle h, int step, int step_s);
int fsilbm_block_fluid_flux(fsilbm_handle h, double out[3]);
int fsilbm_block_probe_velocity(fsilbm_handle h, int n, const double *coords, double *velocity);

/* set_boundary_conditions_ (FluidDomain.f90:616-1126) on the current fIn; the start-up call of
 * tree_set_boundary_conditions_block (main.f90:63).  Reproduces the first-call skip of the
 * half-way codes 203/204 (:660-661). */
int fsilbm_block_set_boundary_conditions(fsilbm_handle h);

/* The fused step: calculate_macro_quantities + ResetVolumeForce + add_volume_force + collision
 * + halfwayBCset + streaming + set_boundary_conditions of one block
 * (LBMBlockComm.f90:285-303 minus IBM_FEM), one read and one write of fIn.  Uses the velocity
 * correction and force left by the last fsilbm_ibm_interaction_force call of this step, if any.
 * Asynchronous: returns after enqueueing. */
int fsilbm_block_collide_stream(fsilbm_handle h);
/* Wait for all work enqueued on the block. */
int fsilbm_block_sync(fsilbm_handle h);
/* The cudaStream_t the block's kernels are launched on (so a host can time them with CUDA events). */
int fsilbm_block_stream(fsilbm_handle h, void **stream);

/* Un-fused single passes, for per-procedure parity tests and for drivers that keep the
 * reference's call granularity.  They operate on device-resident den/uuu/force fields that are
 * allocated on first use (3.5x the memory traffic of the fused step).
 *   macro     : calculate_macro_quantities_ (FluidDomain.f90:1128)
 *   reset/add : ResetVolumeForce_ (:1195), add_volume_force_ (:1182)
 *   collision : collision_ (:1208), halfwayBCset_ (:567), streaming_ (:1514)          */
int fsilbm_block_pass_macro(fsilbm_handle h);
int fsilbm_block_pass_reset_volume_force(fsilbm_handle h);
int fsilbm_block_pass_add_volume_force(fsilbm_handle h);
int fsilbm_block_pass_collision(fsilbm_handle h);
int fsilbm_block_pass_halfway_bc_set(fsilbm_handle h);
int fsilbm_block_pass_streaming(fsilbm_handle h);
/* copy the device den/uuu/force fields of the un-fused passes to/from the host (NULL = skip) */
int fsilbm_block_download_fields(fsilbm_handle h, double *den, double *uuu, double *force);
int fsilbm_block_upload_fields(fsilbm_handle h, const double *den, const double *uuu, const double *force);

/* ---- immersed boundary: replaces calculate_interaction_force (Solidbody.f90:869-918) -------- */

/* Device part of FSInteraction_force (Solidbody.f90:589-602) for the bodies carried by block h:
 * UpdateElmtInterp_ (:760), the PenaltyForce_ loop (:895-906, :981) and the Eulerian half of
 * FluidVolumeForce_ (:968-976).  The host keeps UpdatePosVelArea_ (:599) and the nodal-load half
 * of FluidVolumeForce_ (:945-967), which need FEM internals.
 *   Exyz[b], Evel[b]  : v_Exyz(3,n), v_Evel(3,n) of body b     (host, in)
 *   Ea[b]             : v_Ea(n), already times IBPenaltyBeta (:613,624)  (host, in)
 *   Eforce[b]         : v_Eforce(3,n)                           (host, out)
 *   restencil[b]      : the test at :885 (v_move==1 .or. iBodyModel==2 .or. count_Interp==0)
 *   rootBC            : m_boundaryConditions, the ROOT block's codes (main.f90:44, Solidbody.f90:337)
 *   dt                : LBMBlockComm.f90:328 passes dh for dt
 * Synchronous: Eforce and iterLBM_out are valid on return.  The velocity correction and the force
 * field stay on the device for the next fsilbm_block_collide_stream. */
int fsilbm_ibm_interaction_force(fsilbm_handle h, int nbody, const int *nelmts,
                                 const double *const *Exyz, const double *const *Evel,
                                 const double *const *Ea, double *const *Eforce,
                                 const int *restencil, double dt, int ntolLBM, double dtolLBM,
                                 const int rootBC[6], int *iterLBM_out);
/* The same call in two halves, so that the host never stands between the interaction force and the update that consumes it:
 *   _begin  enqueues everything on the device (marker upload, stencils, penalty iteration, spreading, asynchronous read-back of
 *           v_Eforce into pinned memory) and returns at once;
 *   fsilbm_block_collide_stream may then be called immediately -- it waits for the box fields ON THE DEVICE;
 *   _wait   blocks until the read-back has landed, copies v_Eforce out and reports iterLBM and the reference's fatal
 *           conditions (stencil out of the domain, NaN).  Every _begin is followed by exactly one _wait before the next _begin.
 * fsilbm_ibm_interaction_force above is _begin followed by _wait. */
int fsilbm_ibm_interaction_force_begin(fsilbm_handle h, int nbody, const int *nelmts,
                                       const double *const *Exyz, const double *const *Evel,
                                       const double *const *Ea, const int *restencil,
                                       double dt, int ntolLBM, double dtolLBM, const int rootBC[6]);
int fsilbm_ibm_interaction_force_wait(fsilbm_handle h, int nbody, double *const *Eforce, int *iterLBM_out);
/* Slab runs (x-slab decomposition over several GPUs).  By default a body is iterated only by the ranks whose planes its
 * stencil box touches: a box inside one slab costs no communication at all, the two (rarely more) ranks sharing a box send
 * one another the box planes they own and then iterate it redundantly, bit-identically; the loop control of :895-906 (sum of
 * |dU| over ALL bodies) is all-reduced, two numbers per iteration, so the call is COLLECTIVE over the ranks of the block.
 *   option "ibm_force_exchange" = 1 (default): every rank passes the SAME body list and gets every body's v_Eforce back
 *     (one all-reduce of the leaders' values); nbody = 0 returns at once.
 *   option "ibm_force_exchange" = 0: every rank passes only the bodies it holds -- at least every body whose stencil box
 *     (and every body sharing cells with it) touches its slab, in the same relative order on every rank; bodies whose box
 *     is elsewhere are ignored and get zero force.  Every rank calls every step, also with nbody = 0.
 * fsilbm_ibm_body_status: per body of the last call, 0 not iterated by this rank, 1 iterated, 2 iterated and led (this rank
 * owns the first plane of its box and reported its residual and forces). */
int fsilbm_ibm_body_status(fsilbm_handle h, int nbody, int *status);
/* v_Ei(12,n) as int16 and v_Ew(12,n) as float of body b after the last call (parity checks). */
int fsilbm_ibm_download_stencil(fsilbm_handle h, int body, short *Ei, float *Ew);

/* ---- grid refinement: type CommPair and the father<->son transfers (LBMBlockComm.f90) ------------------- */

/* build_blocks_comunication (LBMBlockComm.f90:32-96) + allocate_fIn_tau (:213-264) + check_blocks_params (:508-544)
 * for one father/son pair: dh_father = 2*dh_son, son extents odd (even where the son is periodic), son corners on
 * father nodes, else FSILBM_ERR_ARG with the reference's message.  Son faces with BndConds = 0 (BCfluid) are coupled.
 * interpolateScheme is flow%interpolateScheme (2 = cubic, otherwise linear; FlowCondition.f90:71).
 * The tree itself (build_block_tree :195, CompareBlocks) stays with the host. */
int fsilbm_pair_create(fsilbm_handle father, fsilbm_handle son, int interpolateScheme, int *pair);
/* Slab runs.  A son lives whole on ONE rank.  If its footprint lies inside that rank's father slab nothing else is needed.  If it
 * reaches into the slab of the rank directly to the left or right (a son across a slab interface), the owner's kernels read and
 * write the neighbour's father planes through the peer-mapped population buffers of the halo (NVLink), and the NEIGHBOUR
 * registers the pair with fsilbm_pair_create_remote(father, owner_rank, &pair): its father block then announces every finished
 * step to the owner and waits for the owner's son->father delivery before it starts the next one (device-side flags, no host
 * in the loop).  Both calls are made at the same point of the run (before the first step); the other pair functions are no-ops
 * on a remote registration.  Needs option "halo" = 1 with CUDA IPC available; the father must not be an LES block. */
int fsilbm_pair_create_remote(fsilbm_handle father, int owner_rank, int *pair);
int fsilbm_pair_destroy(int pair);
/* out[0:6] sds, [6:12] s, [12:18] f, [18:24] si, [24:30] fi (1-based plane indices as in CommPair, :11-18),
 * [30:33] xDimS,yDimS,zDimS, [33:36] xDimF,yDimF,zDimF */
int fsilbm_pair_info(int pair, int out[36]);
/* extract_interpolate_layer (:340-505) for this son: time 1 before the father's collision, 2 after its boundary step. */
int fsilbm_pair_extract_layer(int pair, int time);
/* interpolation_father_to_son (:655-806) with interpolate_fIn (:808), interpolate_tau (:907), fIn_GridTransform (:958). */
int fsilbm_pair_father_to_son(int pair, int n_timeStep);
/* deliver_son_to_father (:546-653). */
int fsilbm_pair_son_to_father(int pair);

/* ---- multi-GPU: x-slab halo exchange (no reference counterpart; the reference is one process) */

/* rank 0 fills id[128] (an ncclUniqueId); the host distributes it; every rank calls comm_init. */
int fsilbm_comm_unique_id(char id[128]);
int fsilbm_comm_init(int rank, int nranks, const char id[128]);
int fsilbm_comm_finalize(void);
/* How block h exchanges its x-slab halo: 0 = single rank (x wraps inside the kernel), 1 = ncclSend/ncclRecv on a
 * high-priority stream overlapped with the interior update, 2 = the edge-plane kernels store the outgoing
 * populations straight into the neighbour GPUs' memory over NVLink (CUDA-IPC-mapped) and raise arrival flags there.
 * Mode 2 is the default (fsilbm_set_option("halo", 1)); the library drops to mode 1 when IPC mapping is unavailable.
 * Blocks of a multi-rank run must be created in the same order on every rank (creation is collective). */
int fsilbm_block_halo_transport(fsilbm_handle h, int *mode);

#ifdef __cplusplus
}
#endif
#endif /* FSILBM_H */
