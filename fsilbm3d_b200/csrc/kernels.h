// kernels.h -- host-visible parameter blocks and launchers of the CUDA kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "d3q19.cuh"

namespace fsilbm {

// Device layout of the populations of one x-slab: f[q][xp][y][z], z fastest, xp = x_local + 1 so
// that planes xp = 0 and xp = X + 1 are ghost planes that receive what streams out of the slab.
// This is the reference's fIn(z,y,x,q) (FluidDomain.f90:384) plus the two ghost planes.
struct Geom {
    int X, Y, Z;       // local extents (X = planes owned by this rank)
    int XG, xOffset;   // global x extent and global index of local plane 0
    size_t plane;      // Y*Z
    size_t pstride;    // (X+2)*Y*Z, distance between populations
    double dh, xmin, ymin, zmin;
};

constexpr int MAX_BOXES = 16;

// Sparse region(s) around the immersed bodies in which the IBM-corrected velocity uuu and the
// Eulerian IBM force live (the reference keeps both as full fields, FluidDomain.f90:386-387).
// Box coordinates are GLOBAL cell indices; lo is normalised to [0,N), a box may wrap periodically.
struct IbmBoxes {
    int n;
    int lo[MAX_BOXES][3];    // x,y,z start
    int ext[MAX_BOXES][3];   // x,y,z extent
    long long off[MAX_BOXES];
    long long ncell;         // total cells; component k of a field starts at k*ncell
    double *u;               // [3][ncell]
    double *force;           // [3][ncell]
};

// One launch of the fused step covers a list of x-plane ranges, taken in list order: blockIdx.z walks the ranges one after the
// other.  CTAs are dispatched in blockIdx order, so the planes listed first finish first -- the slab's edge planes (whose
// stores ARE the halo transfer) and the planes around the immersed bodies (which the next interaction-force call waits for)
// go to the front of the list and the whole update is still one or two launches.
constexpr int MAX_SEG = 2 * MAX_BOXES + 4;

struct StepParams {
    Geom g;
    const double *fA;
    double *fB;
    int nseg;                 // ranges of local planes processed by this launch, in this order
    int seg_begin[MAX_SEG], seg_count[MAX_SEG];
    int wrap_x;             // 1: streaming wraps in x inside the slab (single rank, FluidDomain.f90:1603-1604,1618-1619)
    CollideConsts cc;
    double hF[3];           // 0.5d0*volumeForce(k)*dh, FluidDomain.f90:1137-1139
    double Fvol[3];         // volumeForce(k), :1188-1190
    IbmBoxes boxes;
    // Slab halo over NVLink peer memory (the edge-plane launch of a multi-GPU run only; all null otherwise).
    // halo_hi: where population q with ex=+1 leaving local plane X-1 goes = the RIGHT NEIGHBOUR'S streamed buffer at its plane 0,
    // address halo_hi + q*halo_hi_ps + y*Z + z; halo_lo likewise for ex=-1 leaving plane 0 into the left neighbour's last plane.
    // The last CTA of an edge plane publishes `step` to sig_hi / sig_lo (flags in the neighbours' memory) after a system
    // fence; cta_counter[0] counts the CTAs of plane 0, cta_counter[1] those of plane X-1.
    double *halo_hi, *halo_lo;
    size_t halo_hi_ps, halo_lo_ps;
    unsigned long long *sig_hi, *sig_lo;
    unsigned int *cta_counter;
    unsigned long long step;
    // LES models (11 Smagorinsky, 14 WALE, 15 Vreman)
    double *tau_all;          // [X][Y][Z]
    const double *uuu;        // velocity field of this step (models 14, 15), written by macro_full_kernel: local plane 0 of component 0
    size_t uuu_ncomp;         // distance between its components: X*Y*Z on one GPU, (X+2)*Y*Z on an x-slab (ghost planes, see LesCtx)
};

// Receiving side of the peer-memory halo: one thread waits until both neighbours have published this step (their edge planes
// have stored into this rank's streamed buffer).
struct HaloWaitParams {
    const unsigned long long *flag_lo, *flag_hi;
    unsigned long long step;
    int *err;                              // set to 1 if a flag did not arrive within the time limit
    unsigned long long timeout_ns;
};

struct FaceParams {
    Geom g;
    double *f;              // populations the face rule is applied to (post-stream)
    const double *fA;       // pre-collision populations (stash / layer-2 kernels)
    int face, code;         // 0..5 = xmin,xmax,ymin,ymax,zmin,zmax
    int na, nb;             // face-local extents: x faces (z,y); y faces (z,x); z faces (y,x)
    VelocityField vel;
    double denIn;
    double wallc;           // coordinate of the wall along the face normal
    double *stash;          // [19][nb][na] post-collision boundary layer (fIn_hw*, FluidDomain.f90:567-614)
    double *l2den;          // [nb][na]    den of the first interior layer (code 102, :643)
    double *l2u;            // [3][nb][na] uuu of the first interior layer
    // for the stash / layer-2 kernels, which redo macro + collision of single layers
    CollideConsts cc;
    double hF[3], Fvol[3];
    IbmBoxes boxes;
    int model;
    double *tau_all;          // LES blocks
    const double *uuu;
    size_t uuu_ncomp;
};

struct FieldParams {
    Geom g;
    double *f;              // in place
    double *den, *uuu, *force;   // [X][Y][Z], [3][X][Y][Z], [3][X][Y][Z] (no ghost planes)
    CollideConsts cc;
    double hF[3], Fvol[3];
    double *tau_all;        // LES blocks
};

// ---- launchers (fluid_kernels.cu) ---------------------------------------------------------------
void upload_mrt(int slot, const double *M_COLLID, const double *M_FORCE, cudaStream_t s);
int launch_collide_push(const StepParams &p, int model, cudaStream_t s);
void step_add_planes(StepParams &p, int begin, int count);   // appends a range of local planes to the launch list
void launch_initialise(const Geom &g, double *f, const VelocityField &vel, double denIn, cudaStream_t s);
// ncomp: distance between the components of uuu (0 = X*Y*Z, the ghost-free layout)
void launch_macro_full(const Geom &g, const double *f, const double hF[3], double *den, double *uuu, cudaStream_t s, const IbmBoxes *boxes = nullptr,
                       size_t ncomp = 0);
void launch_bc_face(const FaceParams &p, cudaStream_t s);
void launch_bc_face_pair(const FaceParams &lo, const FaceParams &hi, cudaStream_t s);   // the two faces of one axis in one launch
void launch_stash_face(const FaceParams &p, cudaStream_t s);
void launch_layer2_face(const FaceParams &p, cudaStream_t s);
void launch_init_layer2(const FaceParams &p, cudaStream_t s);
void launch_field_stat(const Geom &g, const double *f, const double hF[3], double invUref, double *out6, cudaStream_t s);
void launch_wrap_x(const Geom &g, double *f, cudaStream_t s);
void launch_halo_wait(const HaloWaitParams &p, cudaStream_t s);
// un-fused passes
void launch_pass_fill(double *p, size_t n, double v, cudaStream_t s);
void launch_pass_add_force(const FieldParams &p, cudaStream_t s);
int launch_pass_collision(const FieldParams &p, int model, cudaStream_t s);
void launch_pass_halfway(const FaceParams &p, cudaStream_t s);
void launch_pass_streaming(const Geom &g, const double *fA, double *fB, cudaStream_t s);

// ---- IBM (ibm_kernels.cu) ------------------------------------------------------------------------
struct IbmBody {
    int n;                    // v_nelmts
    const double *Exyz;       // [n][3]
    const double *Evel;       // [n][3]
    const double *Ea;         // [n]
    double *Eforce;           // [n][3]
    short *Ei;                // [n][12]  integer(2), Solidbody.f90:45
    float *Ew;                // [n][12]  real(4),    Solidbody.f90:46
    int *cell;                // [n][12]  box-local premultiplied offsets (x*ny*nz, y*nz, z) of the stencil
    long long *boff;          // [n] offset of the marker's box in the box arrays
    unsigned char *owned;     // [n][4] 1 if stencil plane ix(m) lies in this rank's slab
    double *felt;             // [n][3] forceElemTemp, Solidbody.f90:1025
    double *tol;              // [n]   |dU1|+|dU2|+|dU3| of the marker, :1023
};

struct IbmCtl {               // device-resident control block of the penalty iteration (Solidbody.f90:893-906)
    int iter;                 // iterLBM
    int done;                 // loop condition false
    int err;                  // bit 0: stencil out of domain (:850,861); bit 1: NaN (:1028); bit 2: stencil outside box
    double dmax;              // dmaxLBM
    double tol_acc;           // running sum of the markers' |dU| of the current iteration (single-launch path)
};

// Cell-centric view of all stencils of a block: for every box cell the (body, marker, stencil node) triples that touch
// it, sorted ascending -- i.e. in the order the reference's serial spreading loops visit the cell (bodies in order,
// markers 1..n, x,y,z loops; Solidbody.f90:898-903,1034-1048,938-978).  Spreading as an ordered per-cell gather makes
// the corrected velocity and the force field bit-identical to the serial reference and independent of scheduling.
struct IbmCsr {
    int *count;                   // [ncell+1] scratch of the build (zero outside it)
    int *off;                     // [ncell+1] exclusive prefix of the per-cell entry counts
    unsigned long long *entry;    // [off[ncell]]  body << 40 | marker << 8 | node (node = 16*a + 4*b + c)
};

constexpr int MAX_IBM_PHASE_BODIES = 64;

// Loop control of the penalty iteration across the ranks of a slab run, exchanged through peer memory (NVLink) from inside
// ibm_loop_kernel: every rank stores {sum of |dU| of the bodies it leads, their marker count, sequence number} into slot
// [seq % IBM_CTL_SLOTS][own rank] of EVERY rank's mailbox and then reads the R entries of its own mailbox; all ranks add
// them in rank order and so take the same decision (Solidbody.f90:895-906) without a collective call.
constexpr int MAX_PEERS = 16;
constexpr int IBM_CTL_SLOTS = 4;
struct CtlSlot { double tol, cnt; unsigned long long seq, pad; };
struct IbmCtlExchange {
    int nranks, rank;                 // nranks <= 1: no exchange
    unsigned char *mailbox[MAX_PEERS];  // every rank's mailbox (own included)
    unsigned long long seq_base;      // sequence number of iteration 0 of this call
    unsigned long long timeout_ns;
    double cnt_local;                 // markers of the bodies this rank leads
};
struct IbmLoopParams {        // the single-launch form of calculate_interaction_force (ibm_loop_kernel)
    Geom g;
    const IbmBody *bodies;    // device array [nbody]
    int nbody;
    IbmBoxes boxes;
    int rootBC[6];
    IbmCtl *ctl;
    const double *fA;
    double hF[3];
    int ntol;
    double dtol, Uref, dsum;  // dsum = total marker count (:902)
    double invh3_pen, invh3;  // 0.5*dt/dh^3/denIn (:996) and 1/dh^3 (:936)
    int nphase;               // bodies sharing a stencil box are taken one per phase, in body order
    int phase_start[MAX_IBM_PHASE_BODIES + 1];
    int phase_body[MAX_IBM_PHASE_BODIES];
    unsigned int *barrier;    // grid barrier counter (zero between launches)
    int ordered;              // 1: ordered (bit-reproducible) gather / spreading through csr; 0: warp shuffles + fp64 atomics
    int do_stencil;           // 1: phase 0 computes the stencils (atomic mode); 0: they were computed before the launch
    int do_macro;             // 1: phase 0 fills the box cells from fA; 0: done before the launch (multi-rank: all-reduced box velocities)
    IbmCsr csr;
    int phase_of_body[MAX_IBM_PHASE_BODIES];
    double *tol_partial;      // [2][gridDim.x] per-block sums of the markers' |dU| (ordered mode), double-buffered over iterations
    unsigned long long *prof; // optional [64] globaltimer stamps of block 0 at the phase boundaries (FSILBM_IBM_PROFILE=1)
    unsigned char lead[MAX_IBM_PHASE_BODIES];   // slab runs: 1 = this rank reports the body's residual (it owns the first plane of its box)
    IbmCtlExchange xc;
};
// a rank that iterates no body still takes part in the loop-control exchange (one small block)
void launch_ibm_ctl_only(const IbmCtlExchange &xc, int ntol, double dtol, double Uref, IbmCtl *ctl, cudaStream_t s);
int launch_ibm_loop(const IbmLoopParams &p, int max_markers, int blocks_per_sm, int blocks_total, cudaStream_t s);
int ibm_loop_max_blocks();

// build of IbmCsr after the stencils are known: count -> scan -> fill -> per-cell sort
size_t ibm_csr_scan_bytes(long long ncell);
int launch_ibm_csr_build(const IbmBody *bodies_dev, int nbody, int max_n, const IbmBoxes &boxes, const IbmCsr &csr, void *scan_tmp, size_t scan_bytes, cudaStream_t s);
void launch_ibm_stencil_all(const Geom &g, const IbmBody *bodies_dev, int nbody, int max_n, const IbmBoxes &boxes, const int rootBC[6], IbmCtl *ctl, cudaStream_t s);
void launch_ibm_gather_ordered(const IbmBody &b, const IbmBoxes &boxes, double *partialU, const IbmCtl *ctl, int fused, double invh3, cudaStream_t s);
void launch_ibm_scatter_ordered(const IbmBody *bodies_dev, int body, const IbmBoxes &boxes, const IbmCsr &csr, const IbmCtl *ctl, cudaStream_t s);
void launch_ibm_spread_ordered(const IbmBody *bodies_dev, const IbmBoxes &boxes, const IbmCsr &csr, double invh3, cudaStream_t s);

void launch_ibm_stencil(const Geom &g, const IbmBody &b, const IbmBoxes &boxes, const int rootBC[6], IbmCtl *ctl, cudaStream_t s);
void launch_ibm_macro_box(const Geom &g, const double *fA, const double hF[3], const IbmBoxes &boxes, cudaStream_t s);
void launch_ibm_gather(const IbmBody &b, const IbmBoxes &boxes, double *partialU, const IbmCtl *ctl, int fused, double invh3, cudaStream_t s);
void launch_ibm_scatter(const IbmBody &b, const IbmBoxes &boxes, const IbmCtl *ctl, cudaStream_t s);
void launch_ibm_check(const IbmBody *bodies_dev, int nbody, double Uref, int ntol, double dtol, IbmCtl *ctl, cudaStream_t s);
void launch_ibm_spread(const IbmBody &b, const IbmBoxes &boxes, double invh3, cudaStream_t s);
// loop control of slab runs: local sum over the bodies this rank leads -> (all-reduce) -> decision, see ibm_kernels.cu
void launch_ibm_tol_sum(const IbmBody *bodies_dev, const int *lead_dev, int nbody, const IbmCtl *ctl, double *out2, cudaStream_t s);
void launch_ibm_decide(const double *in2, double Uref, int ntol, double dtol, IbmCtl *ctl, cudaStream_t s);

// ---- grid refinement (refine_kernels.cu): one son face coupled to its father, LBMBlockComm.f90:340-979 ----------
// The father's populations as the rank that owns the son sees them: its own slab and, for a son across a slab interface, the
// neighbouring slabs through their peer-mapped buffers (NVLink loads / stores from inside the face kernels).
struct FatherView {
    int nseg;                  // 1 (own slab only) .. 3
    double *f[3];              // current population buffer of each segment ([19][X+2][Y][Z])
    int x0[3], X[3];           // first global plane and thickness of each segment
    size_t pstride[3];         // (X+2)*Y*Z of each segment
};

struct PairFaceParams {
    Geom gF, gS;
    FatherView fv;             // father populations (current buffers): read by extract, written by son->father
    double *fS;                // son populations (current buffer)
    int axis;                  // face normal: 0 x, 1 y, 2 z; in-plane axes b (faster) and a: x faces (z,y), y faces (z,x), z faces (y,x)
    int scheme;                // flow%interpolateScheme: 2 = cubic, otherwise linear
    int bF, aF, bS, aS;        // coarse footprint and son face extents
    int fplane, fb0, fa0;      // father plane f(j) and footprint origin, 0-based
    int splane;                // son boundary plane s(j), 0-based
    int siplane, sib0, sia0;   // son inner plane si(j) and in-plane start, 0-based
    int fiplane, fib0, fia0;   // father inner plane fi(j) and in-plane start, 0-based
    int nb, na;                // father nodes overwritten by son->father
    double *buf[2];            // fIn_F?t1 / t2, planar [q][aF][bF]
    double *tbuf[2];           // tau_F?t1 / t2 [aF][bF]; null when neither block has a tau_all field
    double tauF, tauS;         // block tau (constant-tau models)
    const double *tauF_all, *tauS_all;   // tau_all fields [X][Y][Z] of LES blocks, else null
    double hF[3];              // 0.5*volumeForce*dh of the block whose force enters fIn_GridTransform
};
// the coupled faces of one pair, in the reference's face order (j = 1..6 of LBMBlockComm.f90:354, 669)
struct PairFaces { int n; PairFaceParams face[6]; };
void launch_pair_extract(const PairFaces &ps, int time, cudaStream_t s);
void launch_pair_f2s(const PairFaces &ps, int t, cudaStream_t s);
void launch_pair_s2f(const PairFaces &ps, cudaStream_t s);
// one-thread kernel: release-store `value` to a (possibly peer-mapped) 64-bit flag at system scope
void launch_flag_signal(unsigned long long *flag, unsigned long long value, cudaStream_t s);

// ---- output / diagnostics on device state (io_kernels.cu) ---------------------------------------------------------
struct FlowWindowParams {
    Geom g;
    const double *f;
    const double *uuu_ave;    // [9][X][Y][Z], outputtype >= 2 only
    float *out;               // [nfields][nx][ny][nz]
    int x0, off;              // first local x plane of the window; offsetOutput (y and z windows start at `off`)
    int nx, ny, nz;
    int outputtype;
    double hF[3], denIn, invUref, invUrefs;
};
void launch_flow_window(const FlowWindowParams &p, cudaStream_t s);
void launch_turbulent_statistic(const Geom &g, const double *f, const double hF[3], double *ave, double invStep, cudaStream_t s);
void launch_fluid_flux(const Geom &g, const double *f, const double hF[3], const int xl[3], double *out3, cudaStream_t s);
void launch_probe(const Geom &g, const double *f, const double hF[3], int n, const double *coords, double *out, cudaStream_t s);

long long kernel_launch_count();
void count_launch(int n = 1);

}  // namespace fsilbm
