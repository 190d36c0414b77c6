// fsilbm_api.cu -- the C ABI of include/fsilbm.h: block state, step orchestration, IBM driver,
// slab halo exchange.  Host code only; kernels live in fluid_kernels.cu / ibm_kernels.cu.
#include "../../include/fsilbm.h"
#include "kernels.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <dlfcn.h>
#include <functional>
#include <memory>
#include <string>
#include <vector>

using namespace fsilbm;

// ---- error plumbing ---------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess) return fail(FSILBM_ERR_CUDA, "CUDA: %s at %s:%d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static double wall_seconds()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- launch trace (diagnostics; option "trace") -----------------------------------------------------------
// A timing event is recorded on the launching stream after every launch of a step, together with the host clock; the dump
// lists, per mark, when the device finished it and when the host issued it.  This is how the step time-lines in profiles/
// were taken (nsys is not available on the GPU boxes).
namespace {
struct TraceMark { cudaEvent_t ev; const char *name; int stream; double host_s; };
std::vector<TraceMark> g_trace;
int g_trace_on = 0;
cudaEvent_t g_trace_origin = nullptr;
double g_trace_origin_host = 0.0;
void trace_mark(cudaStream_t s, int stream_id, const char *name)
{
    if (!g_trace_on || g_trace.size() >= 100000) return;
    TraceMark m{nullptr, name, stream_id, wall_seconds()};
    if (cudaEventCreate(&m.ev) != cudaSuccess || cudaEventRecord(m.ev, s) != cudaSuccess) return;
    g_trace.push_back(m);
}
}  // namespace
#define TRACE(stream, id, name) do { if (g_trace_on) trace_mark(stream, id, name); } while (0)

// ---- NCCL through dlopen (only multi-rank runs need it) ------------------------------------------------
namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_p;
struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
    int (*CommInitRank)(ncclComm_p *, int, ncclUniqueId_t, int) = nullptr;
    int (*CommDestroy)(ncclComm_p) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_p, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    ncclComm_p comm = nullptr;
    int rank = 0, nranks = 1;
} g_nccl;
constexpr int kNcclFloat64 = 8;   // ncclDouble
constexpr int kNcclSum = 0;       // ncclSum
constexpr int kNcclChar = 0;      // ncclChar

int nccl_load()
{
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail(FSILBM_ERR_COMM, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                  \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.lib, name);                                  \
    if (!g_nccl.field) return fail(FSILBM_ERR_COMM, "NCCL symbol %s missing", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return 0;
}
#define NCK(call)                                                                                          \
    do {                                                                                                   \
        int r_ = (call);                                                                                   \
        if (r_ != 0) return fail(FSILBM_ERR_COMM, "NCCL: %s at %s:%d", g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// ---- block state ------------------------------------------------------------------------------------
struct Interval { int s, l; };   // start in [0,N), length <= N, on a circle of N (or a line if not periodic)

inline int imod(int a, int n) { int r = a % n; return r < 0 ? r + n : r; }

// union of two overlapping intervals on a circle of N
inline bool overlap(const Interval &a, const Interval &b, int N) { return imod(b.s - a.s, N) < a.l || imod(a.s - b.s, N) < b.l; }
inline Interval merge(const Interval &a, const Interval &b, int N)
{
    Interval r;
    if (imod(b.s - a.s, N) < a.l) { r.s = a.s; r.l = std::max(a.l, imod(b.s - a.s, N) + b.l); }
    else { r.s = b.s; r.l = std::max(b.l, imod(a.s - b.s, N) + a.l); }
    if (r.l >= N) { r.s = 0; r.l = N; }
    return r;
}

struct HostBox { Interval ax[3]; };

// bounding interval of the base indices i (1-based, Solidbody.f90:817-820) of one body along one axis,
// widened to the 4-point stencil i-1..i+2 plus one guard cell each side
Interval axis_interval(int imin, int imax, int N, bool periodic)
{
    int lo = imin - 1 - 1 - 1, hi = imax + 2 - 1 + 1;   // 0-based, guard of 1
    Interval r;
    if (periodic) {
        int len = hi - lo + 1;
        if (len >= N) { r.s = 0; r.l = N; }
        else { r.s = imod(lo, N); r.l = len; }
    } else {
        lo = std::max(lo, 0); hi = std::min(hi, N - 1);
        if (hi < lo) { lo = 0; hi = 0; }
        r.s = lo; r.l = hi - lo + 1;
    }
    return r;
}


struct BodyDev {
    int n = 0;
    double *Exyz = nullptr, *ExyzStencil = nullptr, *Evel = nullptr, *Ea = nullptr, *Eforce = nullptr, *felt = nullptr, *tol = nullptr;
    double *partialU = nullptr;
    short *Ei = nullptr;
    float *Ew = nullptr;
    int *cell = nullptr;
    long long *boff = nullptr;
    unsigned char *owned = nullptr;
    // Exyz, Evel, Ea and Eforce point into the block's packed marker / force buffers (Block::mk_dev, force_dev)
    HostBox hbox{};             // box of the stencils last built (host index arithmetic of UpdateElmtInterp_)
    int cidx[3] = {0, 0, 0};    // cell of the first marker at that time
    bool have_box = false;
    bool stencil_valid = false;   // ExyzStencil / v_Ei / v_Ew on THIS rank match the body's current markers (slab runs: a body is
                                  // only stencilled by the ranks that iterate it)
    int status = 0;             // last call: 0 not iterated by this rank (no plane of its box here), 1 iterated, 2 iterated and led
    void release()
    {
        cudaFree(ExyzStencil); cudaFree(felt); cudaFree(tol);
        cudaFree(partialU); cudaFree(Ei); cudaFree(Ew); cudaFree(cell); cudaFree(boff); cudaFree(owned);
        *this = BodyDev();
    }
    IbmBody view() const
    {
        IbmBody b;
        b.n = n; b.Exyz = ExyzStencil; b.Evel = Evel; b.Ea = Ea; b.Eforce = Eforce; b.Ei = Ei; b.Ew = Ew;
        b.cell = cell; b.boff = boff; b.owned = owned; b.felt = felt; b.tol = tol;
        return b;
    }
};

struct Block {
    Geom g{};
    int bc[6]{}, periodic[3]{};
    int model = 1;
    double params[10]{};
    fsilbm_flow flow{};
    double tau = 0, Omega = 0, Omega2 = 0;
    double Mc[Q * Q]{}, Mf[Q * Q]{};
    int mrt_slot = -1;        // entry of the MRT matrix table (c_MRT) this block collides with; -1 = none held
    bool initialised = false;
    double *f[2] = {nullptr, nullptr};
    int cur = 0;
    double volumeForce[3] = {0, 0, 0};
    double blktime = 0;
    double *stash[6]{}, *l2den[6]{}, *l2u[6]{};
    bool hw_alloc[6]{};
    double *den = nullptr, *uuu = nullptr, *force = nullptr;   // un-fused fields / download staging
    double *tau_all = nullptr;                                 // [X][Y][Z], LES models only (FluidDomain.f90:1279,1422,1505)
    double *uuu_les = nullptr;                                 // x-slab of a WALE / Vreman block: [3][X+2][Y][Z], neighbours' edge planes in the ghosts
    double *uuu_ave = nullptr;                                 // [9][X][Y][Z], running means of calculate_turbulent_statistic_ (:1147)
    float *outtmp = nullptr; size_t outtmp_cap = 0;            // OUTtmp staging of write_flow_ (:30,399-403)
    double *scratch = nullptr; size_t scratch_cap = 0;         // probes / flux
    double *stat = nullptr;
    // IBM
    IbmBoxes boxes{};
    long long box_capacity = 0;
    std::vector<BodyDev> bodies;
    double *mk_dev = nullptr, *force_dev = nullptr;     // packed [per body: Exyz 3n | Evel 3n | Ea n] and [per body: Eforce 3n]
    double *mk_pin = nullptr, *force_pin = nullptr;     // their pinned host mirrors
    size_t marker_cap = 0, marker_total = 0;
    std::vector<size_t> mk_off, f_off;
    std::vector<int> active_prev;                       // bodies iterated by this rank at the last call
    std::vector<int> plane_owner;                       // slab runs: rank owning each global x-plane
    IbmBody *bodies_dev = nullptr;
    int *lead_dev = nullptr;
    double *tol2 = nullptr;
    int bodies_dev_cap = 0;
    cudaStream_t ibm_stream = nullptr;                  // marker upload, stencils and cell lists run beside the compute stream
    cudaEvent_t ev_ibm = nullptr;
    // Early IBM: collide_stream updates the x-planes around the bodies first (ev_early), so the next interaction-force call
    // can run on its own stream beside the rest of that update instead of behind it.  early_x0/x1: GLOBAL x-planes [x0, x1)
    // whose streamed populations are final once ev_early has passed.
    cudaEvent_t ev_macro = nullptr, ev_io = nullptr;    // asynchronous read-backs: den/uuu or OUTtmp staged -> copied out on io_stream
    bool io_pending = false;
    cudaStream_t ibm_main_stream = nullptr;
    cudaStream_t body_stream = nullptr;                 // single-rank early IBM: the planes around the bodies (high priority), beside the rest on `stream`
    cudaEvent_t ev_early = nullptr;
    bool early_ok = false, is_father = false;
    double update_clock_us = 0.0;                       // g_update_clock_us right after this block's last update was issued
    int early_n = 0, early_x0[MAX_BOXES]{}, early_x1[MAX_BOXES]{};
    IbmCtl *ctl = nullptr;
    IbmCtl *ctl_pin = nullptr;                          // pinned host copy of the control block, filled by the asynchronous read-back
    // fsilbm_ibm_interaction_force_begin has enqueued a call whose results (forces, iteration count, error bits) have not been
    // collected by fsilbm_ibm_interaction_force_wait yet; ev_ibm_done follows its last device operation on `stream`
    struct IbmPending { bool active = false; cudaStream_t stream = nullptr; int nbody = 0, nact = 0; bool mailbox = false; double t0 = 0, t1 = 0, t2 = 0; } ibm_pending;
    cudaEvent_t ev_ibm_done = nullptr;
    unsigned int *ibm_barrier = nullptr;
    bool ibm_active = false;
    IbmCsr csr{};                                              // cell-centric stencil lists of the ordered IBM mode
    long long csr_cell_cap = 0, csr_entry_cap = 0;
    void *csr_scan_tmp = nullptr; size_t csr_scan_bytes = 0;
    bool csr_valid = false;
    double *tol_partial = nullptr;
    unsigned long long *ibm_prof = nullptr; int ibm_prof_calls = 0;
    double ibm_host_t[3] = {0.0, 0.0, 0.0};   // FSILBM_IBM_PROFILE: host seconds in box set-up / enqueue / wait, summed over 100 calls
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    cudaStream_t io_stream = nullptr;                   // device-to-host copies of the asynchronous read-backs (not the comm stream: on a slab that one carries the edge planes)
    cudaEvent_t ev_edge = nullptr, ev_comm = nullptr;
    // peer-memory halo (multi-GPU): see halo_setup
    struct Halo {
        bool enabled = false;
        int left = -1, right = -1;                 // neighbour ranks (-1: domain end)
        unsigned char *region = nullptr;           // [4 x u64 arrival flags | pad to 256 B][IBM loop-control mailbox]
        unsigned char *peer_left = nullptr, *peer_right = nullptr;   // the neighbours' regions, peer-mapped through CUDA IPC
        double *peer_f[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [left/right][buffer]: the neighbours' POPULATION buffers, peer-mapped:
                                                   // the edge planes store what leaves the slab straight into the neighbour's streamed buffer
        int peer_X[2] = {0, 0};                    // the neighbours' slab thickness (their buffers have X + 2 planes per population)
        std::vector<void *> opened;                // every IPC mapping of this block, for the teardown
        std::vector<unsigned char *> peer;         // every rank's region (own pointer at the own rank): the IBM loop-control mailbox sits in its header
        bool mailbox = false;                      // all ranks mapped and nranks <= MAX_PEERS
        unsigned long long ctl_seq = 0;            // loop-control exchanges completed so far (identical on every rank)
        unsigned int *counters = nullptr;          // last-CTA counters of the two edge launches
        int *err = nullptr;
        unsigned long long step = 0;
        // sons across this block's slab interfaces (see Pair): pairs registered by the neighbour on either side, and the
        // deliveries this rank has completed towards either side
        int remote_pairs[2] = {0, 0};
        unsigned long long remote_base[2] = {0, 0};     // halo.step when the first registration of that side was made
        int cross_pairs[2] = {0, 0};
        unsigned long long delivered[2] = {0, 0};
    } halo;
    cudaEvent_t ev_pre = nullptr;                  // slab runs: the pre-collision face kernels of a step are done (the edge stream starts behind it)
};

std::vector<std::unique_ptr<Block>> g_blocks;
int g_device = -1;
cudaStream_t g_stream = nullptr;   // every block of the process computes on this one stream: the reference walks its blocks sequentially too
// type CommPair, LBMBlockComm.f90:11-18 (indices kept 1-based as in the reference)
struct Pair {
    int father = -1, son = -1, scheme = 1;
    int sds[6]{}, s[6]{}, f[6]{}, si[6]{}, fi[6]{}, dimS[3]{}, dimF[3]{};
    double *buf[6][2]{}, *tbuf[6][2]{};
    // A son across a slab interface of its father.  On the rank that owns the son: cross[side] = the footprint reaches into the
    // left / right neighbour's slab (its planes are read and written through the peer-mapped buffers).  On that neighbour the
    // pair exists as a registration only (remote = true, son = -1): its father block signals every finished step to the owner
    // and waits for the owner's delivery before it starts the next one.
    bool cross[2] = {false, false};
    bool remote = false;
    int owner_side = -1;       // remote registration: where the owner is (0 left, 1 right)
};
std::vector<std::unique_ptr<Pair>> g_pairs;
int g_force_ghost = 0;
int g_halo_timeout_s = 120;   // a neighbour this late is treated as lost (the wait kernel gives up, fsilbm_block_sync reports it)
int g_ibm_single_launch = 1;   // 1: calculate_interaction_force as one cooperative kernel on single-rank blocks; 0: one kernel per phase
int g_ibm_force_exchange = 1;  // slab runs: 1 every rank passes the same bodies and gets every body's forces back (one all-reduce);
                               // 0 every rank passes only the bodies near its slab (distributed lists; the call is then collective even with none)
int g_ibm_early_blocks = 0;    // blocks per SM of the cooperative IBM kernel when it runs beside a collide-stream update; 0 = chosen per call
int g_ibm_early_total = 0;     // > 0: that grid as an absolute block count
double g_update_clock_us = 0.0;   // estimated device time of every collide-stream update issued so far (all blocks share the compute stream)
int g_update_split = 1;        // one GPU, early IBM: 1 the planes around the bodies on the body stream beside the rest; 0 queued before the rest
int g_ibm_early = 1;           // 1: the planes around the bodies are updated first and the next IBM call overlaps the rest of the update
int g_ibm_ordered = 1;         // 1: interpolation and spreading keep the reference's serial summation order (bit-reproducible); 0: shuffles + fp64 atomics
int g_halo_mode = 1;   // 1: edge kernels store into the neighbours' memory over NVLink (default); 0: ncclSend/ncclRecv

Block *get(fsilbm_handle h)
{
    if (h < 0 || h >= (int)g_blocks.size() || !g_blocks[h]) return nullptr;
    return g_blocks[h].get();
}

inline void face_dims(const Geom &g, int face, int &na, int &nb)
{
    const int axis = face / 2;
    if (axis == 0) { na = g.Z; nb = g.Y; }
    else if (axis == 1) { na = g.Z; nb = g.X; }
    else { na = g.Y; nb = g.X; }
}

// does this rank hold the global boundary plane of `face`?
inline bool owns_face(const Block &b, int face)
{
    if (face == 0) return b.g.xOffset == 0;
    if (face == 1) return b.g.xOffset + b.g.X == b.g.XG;
    return true;
}

int velocity_field(const Block &b, double time, VelocityField &v)
{
    v.kind = b.flow.velocityKind;
    for (int k = 0; k < 3; k++) { v.uvwIn[k] = b.flow.uvwIn[k]; v.shear[k] = b.flow.shearRateIn[k]; v.uniform[k] = b.flow.uvwIn[k]; }
    if (v.kind == 0) return 0;
    if (v.kind == 2) {   // evaluate_oscillatory_velocity, FluidDomain.f90:1813-1824 (host libm cos, as the reference)
        const double Pi = 3.141592653589793;
        const double velocityAmp = b.flow.shearRateIn[0], velocityFreq = b.flow.shearRateIn[1], velocityPhi = b.flow.shearRateIn[2];
        v.uniform[0] = b.flow.uvwIn[0] + velocityAmp * cos(2 * Pi * velocityFreq * time + velocityPhi / 180.0 * Pi);
        return 0;
    }
    return fail(FSILBM_ERR_ARG, "velocityKind %d: the reference defines only 0 and 2 (FluidDomain.f90:1795-1799)", v.kind);
}

CollideConsts collide_consts(const Block &b)
{
    CollideConsts c;
    c.Omega = b.Omega; c.Omega2 = b.Omega2;
    c.dt3 = 3.0 * b.g.dh;             // FluidDomain.f90:1213
    c.cF = 1.0 - 0.5 * b.Omega;       // :1227
    c.mrt_slot = b.mrt_slot;
    c.tau = b.tau; c.nu = b.flow.nu; c.dh = b.g.dh;
    return c;
}

void half_force(const Block &b, double hF[3])
{
    for (int k = 0; k < 3; k++) hF[k] = 0.5 * b.volumeForce[k] * b.g.dh;   // FluidDomain.f90:1137
}

// The velocity field the WALE / Vreman closures difference: the ghost-free staging field on one GPU, the ghosted slab field otherwise.
static inline void les_field(const Block &b, const double *&uuu, size_t &ncomp)
{
    if (b.uuu_les) { uuu = b.uuu_les + b.g.plane; ncomp = (size_t)(b.g.X + 2) * b.g.plane; }
    else { uuu = b.uuu; ncomp = (size_t)b.g.X * b.g.plane; }
}

// The MRT matrix table in constant memory has MRT_SLOTS entries.  An entry is keyed by its contents: blocks with the same
// relaxation time (same dh and nu) share one, an entry is freed when its last block is destroyed or re-initialised, and a block
// whose matrices find no free entry is refused (FSILBM_ERR_MODEL) instead of overwriting another block's.
struct MrtEntry { int refs = 0; std::vector<double> key; };
static MrtEntry g_mrt[MRT_SLOTS];
static void mrt_release(int &slot)
{
    if (slot >= 0 && slot < MRT_SLOTS && g_mrt[slot].refs > 0) g_mrt[slot].refs--;
    slot = -1;
}
static int mrt_acquire(const double *Mc, const double *Mf, cudaStream_t s, bool *fresh)
{
    std::vector<double> key(Mc, Mc + Q * Q);
    key.insert(key.end(), Mf, Mf + Q * Q);
    *fresh = false;
    for (int i = 0; i < MRT_SLOTS; i++)
        if (g_mrt[i].refs > 0 && g_mrt[i].key.size() == key.size() && std::memcmp(g_mrt[i].key.data(), key.data(), sizeof(double) * key.size()) == 0) {
            g_mrt[i].refs++;
            return i;
        }
    for (int i = 0; i < MRT_SLOTS; i++)
        if (g_mrt[i].refs == 0) {
            g_mrt[i].refs = 1;
            g_mrt[i].key.swap(key);
            *fresh = true;
            return i;
        }
    return -1;
}

// calculate_MRT_params, FluidDomain.f90:466-522; matmul sums run over the inner index ascending from zero
void mrt_matrices(double Omega, double *Mc, double *Mf)
{
    static const double s0 = 0.0, s1 = 1.19, s2 = 1.4, s4 = 1.2, s10 = 1.4, s16 = 1.98;   // ConstParams.f90:28
    double M[Q][Q], MI[Q][Q], MM[Q][Q], T[Q][Q];
    for (int I = 0; I < Q; I++) {
        const double e1 = EX(I), e2 = EY(I), e3 = EZ(I);
        const double sq = (double)(EX(I) * EX(I) + EY(I) * EY(I) + EZ(I) * EZ(I));
        M[0][I] = 1.0;
        M[1][I] = 19.0 * sq - 30.0;
        M[2][I] = (21.0 * (sq * sq) - 53.0 * sq + 24.0) / 2.0;
        M[3][I] = e1; M[5][I] = e2; M[7][I] = e3;
        M[4][I] = (5.0 * sq - 9.0) * e1; M[6][I] = (5.0 * sq - 9.0) * e2; M[8][I] = (5.0 * sq - 9.0) * e3;
        M[9][I] = 3.0 * (e1 * e1) - sq;
        M[10][I] = (3.0 * sq - 5.0) * (3.0 * (e1 * e1) - sq);
        M[11][I] = e2 * e2 - e3 * e3;
        M[12][I] = (3.0 * sq - 5.0) * (e2 * e2 - e3 * e3);
        M[13][I] = e1 * e2; M[14][I] = e2 * e3; M[15][I] = e3 * e1;
        M[16][I] = (e2 * e2 - e3 * e3) * e1;
        M[17][I] = (e3 * e3 - e1 * e1) * e2;
        M[18][I] = (e1 * e1 - e2 * e2) * e3;
    }
    for (int i = 0; i < Q; i++) for (int j = 0; j < Q; j++) MI[i][j] = M[j][i];
    for (int i = 0; i < Q; i++) for (int j = 0; j < Q; j++) { double s = 0.0; for (int k = 0; k < Q; k++) s = s + M[i][k] * MI[k][j]; MM[i][j] = s; }
    for (int I = 0; I < Q; I++) for (int r = 0; r < Q; r++) MI[r][I] = MI[r][I] / MM[I][I];
    const double S[Q] = {s0, s1, s2, s0, s4, s0, s4, s0, s4, Omega, s10, Omega, s10, Omega, Omega, Omega, s16, s16, s16};
    for (int i = 0; i < Q; i++) for (int j = 0; j < Q; j++) { double s = 0.0; for (int k = 0; k < Q; k++) s = s + MI[i][k] * (k == j ? S[k] : 0.0); T[i][j] = s; }
    for (int i = 0; i < Q; i++) for (int j = 0; j < Q; j++) { double s = 0.0; for (int k = 0; k < Q; k++) s = s + T[i][k] * M[k][j]; Mc[i * Q + j] = s; }
    for (int i = 0; i < Q; i++) for (int j = 0; j < Q; j++) Mf[i * Q + j] = ((i == j) ? 1.0 : 0.0) - 0.5 * Mc[i * Q + j];
}

bool valid_bc(int c)
{
    switch (c) {
    case 101: case 102: case 103: case 104: case 201: case 202: case 203: case 204: case 301: case 302: case 0: case 1: return true;
    default: return false;
    }
}

FaceParams face_params(Block &b, int face, double *f, const double *fA, const VelocityField &vel)
{
    FaceParams p{};
    p.g = b.g; p.f = f; p.fA = fA; p.face = face; p.code = b.bc[face];
    face_dims(b.g, face, p.na, p.nb);
    p.vel = vel; p.denIn = b.flow.denIn;
    const int axis = face / 2, hi = face & 1;
    const double lo = axis == 0 ? b.g.xmin : axis == 1 ? b.g.ymin : b.g.zmin;
    const int N = axis == 0 ? b.g.XG : axis == 1 ? b.g.Y : b.g.Z;
    double hic = lo + b.g.dh * (N - 1);                 // xmax, FluidDomain.f90:94 (non-periodic faces only reach here)
    p.wallc = hi ? hic : lo;
    if (p.code == BCmoving_Wall_halfway) p.wallc = hi ? hic + b.g.dh * 0.5 : lo - b.g.dh * 0.5;   // :688,772
    p.stash = b.stash[face]; p.l2den = b.l2den[face]; p.l2u = b.l2u[face];
    p.cc = collide_consts(b);
    half_force(b, p.hF);
    for (int k = 0; k < 3; k++) p.Fvol[k] = b.volumeForce[k];
    p.boxes = b.boxes;
    if (!b.ibm_active) p.boxes.n = 0;
    p.model = b.model;
    p.tau_all = b.tau_all;
    les_field(b, p.uuu, p.uuu_ncomp);
    return p;
}

// set_boundary_conditions_ on buffer f, faces in the reference's order (FluidDomain.f90:622-1125).  The two faces of one axis
// touch disjoint cells (their own three outermost layers) whenever the extent is >= 6, so they share one launch; the axes
// follow one another as in the reference, because on shared edge lines a later face reads what an earlier one wrote.
int apply_boundary_conditions(Block &b, double *f)
{
    VelocityField vel;
    if (int rc = velocity_field(b, b.blktime, vel)) return rc;
    const int Ns[3] = {b.g.XG, b.g.Y, b.g.Z};
    for (int axis = 0; axis < 3; axis++) {
        FaceParams fp[2];
        int nfp = 0;
        for (int face = 2 * axis; face < 2 * axis + 2; face++) {
            const int code = b.bc[face];
            if (code == BCPeriodic || code == BCfluid || code == BCfluid_father) continue;   // :701-702
            if (!owns_face(b, face)) continue;
            int na, nb;
            face_dims(b.g, face, na, nb);
            if (code == BCstationary_Wall_halfway || code == BCmoving_Wall_halfway) {
                if (!b.hw_alloc[face]) {   // :660-661: the first call allocates the stash and skips the rule
                    CK(cudaMalloc(&b.stash[face], sizeof(double) * (size_t)Q * na * nb));
                    CK(cudaMemsetAsync(b.stash[face], 0, sizeof(double) * (size_t)Q * na * nb, b.stream));
                    b.hw_alloc[face] = true;
                    continue;
                }
            }
            fp[nfp++] = face_params(b, face, f, nullptr, vel);
        }
        if (nfp == 2 && Ns[axis] >= 6 && (axis != 0 || b.g.X == b.g.XG)) launch_bc_face_pair(fp[0], fp[1], b.stream);
        else for (int k = 0; k < nfp; k++) launch_bc_face(fp[k], b.stream);
    }
    CK(cudaGetLastError());
    return 0;
}

int ensure_fields(Block &b, bool need_force)
{
    if (b.io_pending) { CK(cudaEventSynchronize(b.ev_io)); b.io_pending = false; }   // an asynchronous read-back still copies out of den/uuu
    const size_t n = (size_t)b.g.X * b.g.plane;
    if (!b.den) CK(cudaMalloc(&b.den, sizeof(double) * n));
    if (!b.uuu) CK(cudaMalloc(&b.uuu, sizeof(double) * 3 * n));
    if (need_force && !b.force) { CK(cudaMalloc(&b.force, sizeof(double) * 3 * n)); CK(cudaMemsetAsync(b.force, 0, sizeof(double) * 3 * n, b.stream)); }
    return 0;
}

// one-plane halo exchange of the outgoing populations between x-neighbours:
// ghost plane X+1 (pops with ex=+1) -> right neighbour's plane 1; ghost plane 0 (ex=-1) -> left neighbour's plane X.
int halo_exchange(Block &b, double *fB, cudaStream_t s)
{
    const Geom &g = b.g;
    const int R = g_nccl.nranks, r = g_nccl.rank;
    const bool per = b.periodic[0] == 1;
    const int right = (r + 1 < R) ? r + 1 : (per ? 0 : -1);
    const int left = (r > 0) ? r - 1 : (per ? R - 1 : -1);
    static const int up[5] = {1, 7, 9, 11, 13}, dn[5] = {2, 8, 10, 12, 14};
    NCK(g_nccl.GroupStart());
    // NCCL pairs the sends and receives between two ranks in posting order.  With two ranks and a periodic x
    // axis the left and the right neighbour are the same rank, so post sends as (up->right, dn->left) and
    // receives as (up<-left, dn<-right): the peer's first send (its "up") then meets my first receive.
    for (int k = 0; k < 5; k++) {
        if (right >= 0) NCK(g_nccl.Send(fB + up[k] * g.pstride + (size_t)(g.X + 1) * g.plane, g.plane, kNcclFloat64, right, g_nccl.comm, s));
        if (left >= 0) NCK(g_nccl.Send(fB + dn[k] * g.pstride, g.plane, kNcclFloat64, left, g_nccl.comm, s));
        if (left >= 0) NCK(g_nccl.Recv(fB + up[k] * g.pstride + (size_t)1 * g.plane, g.plane, kNcclFloat64, left, g_nccl.comm, s));
        if (right >= 0) NCK(g_nccl.Recv(fB + dn[k] * g.pstride + (size_t)g.X * g.plane, g.plane, kNcclFloat64, right, g_nccl.comm, s));
    }
    NCK(g_nccl.GroupEnd());
    return 0;
}


constexpr size_t kHaloFlagBytes = 256;                                                   // 4 x u64 arrival flags
constexpr size_t kHaloMailboxBytes = sizeof(CtlSlot) * IBM_CTL_SLOTS * MAX_PEERS;        // loop-control mailbox [slot][rank]
constexpr size_t kHaloHeaderBytes = ((kHaloFlagBytes + kHaloMailboxBytes + 4095) / 4096) * 4096;

inline unsigned long long *halo_flag(unsigned char *region, int side, int parity) { return (unsigned long long *)region + (side * 2 + parity); }
// flags of the sons across an interface, in the same region: kind 0 = "the neighbour on `side` has finished father step n",
// kind 1 = "the neighbour on `side` (owner of sons reaching into this slab) has delivered n son->father transfers"
inline unsigned long long *refine_flag(unsigned char *region, int kind, int side) { return (unsigned long long *)region + (8 + kind * 2 + side); }

// Peer-memory halo set-up (collective over the communicator; every rank creates its blocks in the same order).
// Each rank exports its two population buffers and a small flag/mailbox region with cudaIpcGetMemHandle, the handles are
// all-gathered with NCCL, and each rank maps its neighbours' buffers (and every rank's region: the IBM loop-control mailbox
// lives there).  From then on the edge planes of collide_push_kernel store the populations that leave the slab straight into
// the NEIGHBOUR'S streamed buffer over NVLink and raise a flag in its region: no staging copy, no NCCL call, no host on the
// per-step path.  Returns 0 and leaves halo.enabled = false when IPC is not available (the step then uses ncclSend/ncclRecv).
//
// Why writing into the neighbour's buffer is safe: a rank starts the edge planes of step k+1 only after its step k is complete,
// which includes having seen the neighbour's flags of step k -- and those are raised by the neighbour's edge planes of step k,
// the only readers of the cells (plane 0 / plane X-1 of the neighbour's step-k source buffer) that step k+1's stores overwrite.
int halo_setup(Block &b)
{
    Block::Halo &h = b.halo;
    const int R = g_nccl.nranks, r = g_nccl.rank;
    const bool per = b.periodic[0] == 1;
    h.right = (r + 1 < R) ? r + 1 : (per ? 0 : -1);
    h.left = (r > 0) ? r - 1 : (per ? R - 1 : -1);
    const size_t bytes = kHaloHeaderBytes;
    CK(cudaMalloc(&h.region, bytes));
    CK(cudaMemset(h.region, 0, bytes));
    CK(cudaMalloc(&h.counters, 2 * sizeof(unsigned int)));
    CK(cudaMemset(h.counters, 0, 2 * sizeof(unsigned int)));
    CK(cudaMalloc(&h.err, sizeof(int)));
    CK(cudaMemset(h.err, 0, sizeof(int)));
    cudaIpcMemHandle_t mine[3];
    void *exported[3] = {h.region, b.f[0], b.f[1]};
    int ok = 1;
    for (int k = 0; k < 3; k++)
        if (cudaIpcGetMemHandle(&mine[k], exported[k]) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    // all-gather {ok flag, slab thickness, three handles}
    constexpr size_t rec = 256;
    static_assert(8 + 3 * sizeof(cudaIpcMemHandle_t) <= rec, "record too small");
    unsigned char hostrec[rec] = {0};
    hostrec[0] = (unsigned char)ok;
    memcpy(hostrec + 4, &b.g.X, sizeof(int));
    memcpy(hostrec + 8, mine, sizeof(mine));
    unsigned char *dsend = nullptr, *drecv = nullptr;
    CK(cudaMalloc(&dsend, rec));
    CK(cudaMalloc(&drecv, rec * R));
    CK(cudaMemcpy(dsend, hostrec, rec, cudaMemcpyHostToDevice));
    NCK(g_nccl.AllGather(dsend, drecv, rec, kNcclChar, g_nccl.comm, b.stream));
    CK(cudaStreamSynchronize(b.stream));
    std::vector<unsigned char> all(rec * R);
    CK(cudaMemcpy(all.data(), drecv, rec * R, cudaMemcpyDeviceToHost));
    cudaFree(dsend); cudaFree(drecv);
    int all_ok = 1;
    for (int i = 0; i < R; i++) all_ok &= all[rec * i];
    int opened = 1;
    if (all_ok) {
        auto open = [&](int peer, int which) -> void * {
            cudaIpcMemHandle_t hh;
            memcpy(&hh, all.data() + rec * peer + 8 + sizeof(hh) * which, sizeof(hh));
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, hh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; return nullptr; }
            h.opened.push_back(ptr);
            return ptr;
        };
        // every rank's region is mapped (not only the two neighbours'): the header also carries the mailbox through which the
        // IBM penalty iteration exchanges its loop control (ibm_loop_kernel)
        h.peer.assign(R, nullptr);
        h.peer[r] = h.region;
        for (int p = 0; p < R && opened; p++) if (p != r) h.peer[p] = (unsigned char *)open(p, 0);
        const int nb[2] = {h.left, h.right};
        for (int sd = 0; sd < 2 && opened; sd++) {
            if (nb[sd] < 0) continue;
            memcpy(&h.peer_X[sd], all.data() + rec * nb[sd] + 4, sizeof(int));
            if (sd == 1 && nb[1] == nb[0]) { h.peer_f[1][0] = h.peer_f[0][0]; h.peer_f[1][1] = h.peer_f[0][1]; continue; }   // two ranks, periodic: one peer
            for (int k = 0; k < 2 && opened; k++) h.peer_f[sd][k] = (double *)open(nb[sd], 1 + k);
        }
        if (opened) {
            if (h.left >= 0) h.peer_left = h.peer[h.left];
            if (h.right >= 0) h.peer_right = h.peer[h.right];
        }
    }
    // agree on the outcome (a rank that could not map its neighbour forces everyone onto the NCCL path)
    int *dflag = nullptr;
    CK(cudaMalloc(&dflag, sizeof(int)));
    int mineok = (all_ok && opened) ? 0 : 1;   // sum of failures
    CK(cudaMemcpy(dflag, &mineok, sizeof(int), cudaMemcpyHostToDevice));
    NCK(g_nccl.AllReduce(dflag, dflag, 1, 2 /* ncclInt32 */, kNcclSum, g_nccl.comm, b.stream));
    CK(cudaStreamSynchronize(b.stream));
    int failures = 0;
    CK(cudaMemcpy(&failures, dflag, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dflag);
    h.enabled = failures == 0;
    h.mailbox = h.enabled && R <= MAX_PEERS;
    return 0;
}

void halo_teardown(Block &b)
{
    Block::Halo &h = b.halo;
    if (!h.region) return;
    // (the peer mappings are closed after the barrier below: a neighbour may still be storing into this rank's buffers)
    if (g_nccl.comm) {   // nobody frees a region a neighbour may still be writing into
        int *dflag = nullptr;
        if (cudaMalloc(&dflag, sizeof(int)) == cudaSuccess) {
            cudaMemset(dflag, 0, sizeof(int));
            g_nccl.AllReduce(dflag, dflag, 1, 2, kNcclSum, g_nccl.comm, b.stream);
            cudaStreamSynchronize(b.stream);
            cudaFree(dflag);
        }
    }
    for (void *ptr : h.opened) cudaIpcCloseMemHandle(ptr);
    if (g_nccl.comm) {   // ... and nobody frees buffers a neighbour still has mapped
        int *dflag = nullptr;
        if (cudaMalloc(&dflag, sizeof(int)) == cudaSuccess) {
            cudaMemset(dflag, 0, sizeof(int));
            g_nccl.AllReduce(dflag, dflag, 1, 2, kNcclSum, g_nccl.comm, b.stream);
            cudaStreamSynchronize(b.stream);
            cudaFree(dflag);
        }
    }
    cudaFree(h.region); cudaFree(h.counters); cudaFree(h.err);
    h = Block::Halo();
}

// an asynchronous read-back still copies out of the den/uuu staging fields: wait before they are overwritten
int io_wait(Block &b)
{
    if (b.io_pending) { CK(cudaEventSynchronize(b.ev_io)); b.io_pending = false; }
    return 0;
}

}  // namespace

// =====================================================================================================
extern "C" {

const char *fsilbm_last_error(void) { return g_err; }
long long fsilbm_launch_count(void) { return kernel_launch_count(); }
static long long g_ibm_early_calls = 0;
long long fsilbm_ibm_early_count(void) { return g_ibm_early_calls; }

int fsilbm_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(FSILBM_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(FSILBM_ERR_ARG, "device %d out of range [0,%d)", device, n);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(FSILBM_ERR_CUDA, "device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
    g_device = device;
    if (!g_stream) CK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    return 0;
}

int fsilbm_finalize(void)
{
    for (size_t i = 0; i < g_pairs.size(); i++)
        if (g_pairs[i]) fsilbm_pair_destroy((int)i);
    g_pairs.clear();
    for (size_t i = 0; i < g_blocks.size(); i++)
        if (g_blocks[i]) fsilbm_block_destroy((int)i);
    g_blocks.clear();
    fsilbm_comm_finalize();
    if (g_stream) { cudaStreamDestroy(g_stream); g_stream = nullptr; }
    g_device = -1;
    return 0;
}

int fsilbm_trace_dump(const char *path)
{
    if (!path) return fail(FSILBM_ERR_ARG, "null path");
    CK(cudaDeviceSynchronize());
    FILE *fp = fopen(path, "w");
    if (!fp) return fail(FSILBM_ERR_ARG, "cannot write %s", path);
    fprintf(fp, "mark,stream,device_done_us,host_issued_us\n");
    for (const TraceMark &m : g_trace) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_trace_origin, m.ev);
        fprintf(fp, "%s,%d,%.3f,%.3f\n", m.name, m.stream, (double)ms * 1e3, (m.host_s - g_trace_origin_host) * 1e6);
        cudaEventDestroy(m.ev);
    }
    fclose(fp);
    g_trace.clear();
    return 0;
}

int fsilbm_set_option(const char *key, int value)
{
    if (!key) return fail(FSILBM_ERR_ARG, "null key");
    if (!strcmp(key, "trace")) {
        if (value && !g_trace_on) {
            CK(cudaDeviceSynchronize());
            if (!g_trace_origin) CK(cudaEventCreate(&g_trace_origin));
            CK(cudaEventRecord(g_trace_origin, g_stream));
            CK(cudaEventSynchronize(g_trace_origin));
            g_trace_origin_host = wall_seconds();
        }
        g_trace_on = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(key, "force_ghost")) { g_force_ghost = value ? 1 : 0; return 0; }
    if (!strcmp(key, "ibm_single_launch")) { g_ibm_single_launch = value ? 1 : 0; return 0; }
    if (!strcmp(key, "ibm_early_blocks_per_sm")) { if (value < 0 || value > 4) return fail(FSILBM_ERR_ARG, "ibm_early_blocks_per_sm must be 0 (automatic) or 1..4"); g_ibm_early_blocks = value; return 0; }
    if (!strcmp(key, "update_split")) { g_update_split = value ? 1 : 0; return 0; }
    if (!strcmp(key, "ibm_early_blocks")) { if (value < 0) return fail(FSILBM_ERR_ARG, "ibm_early_blocks must be >= 0"); g_ibm_early_total = value; return 0; }
    if (!strcmp(key, "ibm_early")) { g_ibm_early = value ? 1 : 0; for (auto &bp : g_blocks) if (bp) bp->early_ok = false; return 0; }
    if (!strcmp(key, "ibm_force_exchange")) { g_ibm_force_exchange = value ? 1 : 0; return 0; }
    if (!strcmp(key, "ibm_ordered")) { g_ibm_ordered = value ? 1 : 0; for (auto &bp : g_blocks) if (bp) bp->csr_valid = false; return 0; }
    if (!strcmp(key, "halo_timeout_s")) { if (value < 1) return fail(FSILBM_ERR_ARG, "halo_timeout_s must be >= 1"); g_halo_timeout_s = value; return 0; }
    if (!strcmp(key, "halo")) { if (value < 0 || value > 1) return fail(FSILBM_ERR_ARG, "halo must be 0 (NCCL) or 1 (peer stores)"); g_halo_mode = value; return 0; }
    return fail(FSILBM_ERR_ARG, "unknown option %s", key);
}

int fsilbm_block_create(int xDim, int yDim, int zDim, int xOffset, int xLocal, double dh, double xmin, double ymin, double zmin,
                        const int BndConds[6], int iCollidModel, const double params[10], const fsilbm_flow *flow, fsilbm_handle *out)
{
    if (g_device < 0) return fail(FSILBM_ERR_CUDA, "fsilbm_init has not succeeded (no CPU fallback)");
    if (!BndConds || !params || !flow || !out) return fail(FSILBM_ERR_ARG, "null argument");
    if (xDim < 1 || yDim < 1 || zDim < 1) return fail(FSILBM_ERR_ARG, "non-positive grid size");
    if (xDim > 32767 || yDim > 32767 || zDim > 32767)   // FluidDomain.f90:88-91
        return fail(FSILBM_ERR_ARG, "Grid number exceeds 32767, please try to reduced the grid size.");
    if (xOffset < 0 || xLocal < 1 || xOffset + xLocal > xDim) return fail(FSILBM_ERR_ARG, "bad slab [%d,%d) of %d", xOffset, xOffset + xLocal, xDim);
    for (int i = 0; i < 6; i++) if (!valid_bc(BndConds[i])) return fail(FSILBM_ERR_BC, "face %d has no such boundary condition: %d", i, BndConds[i]);
    if (!(iCollidModel == 1 || iCollidModel == 2 || iCollidModel == 3 || iCollidModel == 11 || iCollidModel == 14 || iCollidModel == 15))
        return fail(FSILBM_ERR_MODEL, "iCollidModel %d not provided (1 SRT, 2 TRT, 3 MRT, 11 Smagorinsky, 14 WALE, 15 Vreman; 12 and 13 read "
                                      "uninitialised variables upstream, FluidDomain.f90:1245-1246,1292,1297-1304)", iCollidModel);
    if ((iCollidModel == 14 || iCollidModel == 15) && xLocal != xDim && xLocal < 3)
        return fail(FSILBM_ERR_MODEL, "iCollidModel %d differences velocity over three x-planes at the domain faces: a slab needs at least 3 planes", iCollidModel);
    auto b = std::make_unique<Block>();
    for (int i = 0; i < 3; i++) {   // check_periodic_boundary_, FluidDomain.f90:110-125
        b->periodic[i] = 0;
        if (BndConds[2 * i] == BCPeriodic || BndConds[2 * i + 1] == BCPeriodic) {
            if (BndConds[2 * i] == BndConds[2 * i + 1]) b->periodic[i] = 1;
            else return fail(FSILBM_ERR_PERIODIC, "Stop! Periodic boundaries must apper in pairs: %d %d", BndConds[2 * i], BndConds[2 * i + 1]);
        }
    }
    Geom &g = b->g;
    g.X = xLocal; g.Y = yDim; g.Z = zDim; g.XG = xDim; g.xOffset = xOffset;
    g.plane = (size_t)yDim * zDim; g.pstride = (size_t)(xLocal + 2) * g.plane;
    g.dh = dh; g.xmin = xmin; g.ymin = ymin; g.zmin = zmin;
    memcpy(b->bc, BndConds, sizeof(int) * 6);
    memcpy(b->params, params, sizeof(double) * 10);
    b->flow = *flow;
    b->model = iCollidModel;
    const size_t bytes = sizeof(double) * Q * g.pstride;
    for (int i = 0; i < 2; i++) {
        CK(cudaMalloc(&b->f[i], bytes));
        CK(cudaMemset(b->f[i], 0, bytes));
    }
    b->stream = g_stream;
    {   // the NCCL transport's stream outranks the compute stream so its CTAs are scheduled as soon as SM slots free up
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&b->comm_stream, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithFlags(&b->io_stream, cudaStreamNonBlocking));
    }
    CK(cudaEventCreateWithFlags(&b->ev_pre, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_edge, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_comm, cudaEventDisableTiming));
    {   // marker upload, stencils and cell lists slip in beside the running collide-stream kernel: their CTAs go first when SM slots free up
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&b->ibm_stream, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&b->ibm_main_stream, cudaStreamNonBlocking, hi));
        CK(cudaStreamCreateWithPriority(&b->body_stream, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&b->ev_early, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_macro, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_io, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&b->ev_ibm, cudaEventDisableTiming));
    CK(cudaMalloc(&b->ctl, sizeof(IbmCtl)));
    CK(cudaMallocHost(&b->ctl_pin, sizeof(IbmCtl)));
    CK(cudaEventCreateWithFlags(&b->ev_ibm_done, cudaEventDisableTiming));
    CK(cudaMalloc(&b->ibm_barrier, sizeof(unsigned int)));
    CK(cudaMemset(b->ibm_barrier, 0, sizeof(unsigned int)));
    CK(cudaMalloc(&b->stat, sizeof(double) * 6));
    if (g_nccl.nranks > 1 && g_nccl.comm && g_halo_mode == 1 && xLocal != xDim)   // collective over the ranks that share the block
        if (int rc = halo_setup(*b)) return rc;
    int slot = -1;
    for (size_t i = 0; i < g_blocks.size(); i++) if (!g_blocks[i]) { slot = (int)i; break; }
    if (slot < 0) { g_blocks.emplace_back(); slot = (int)g_blocks.size() - 1; }
    g_blocks[slot] = std::move(b);
    *out = slot;
    return 0;
}

int fsilbm_block_destroy(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    cudaStreamSynchronize(b->stream);
    cudaStreamSynchronize(b->comm_stream);
    cudaStreamSynchronize(b->io_stream);
    halo_teardown(*b);
    mrt_release(b->mrt_slot);
    for (int i = 0; i < 2; i++) cudaFree(b->f[i]);
    for (int i = 0; i < 6; i++) { cudaFree(b->stash[i]); cudaFree(b->l2den[i]); cudaFree(b->l2u[i]); }
    cudaFree(b->den); cudaFree(b->uuu); cudaFree(b->force); cudaFree(b->stat); cudaFree(b->tau_all);
    cudaFree(b->uuu_ave); cudaFree(b->uuu_les); cudaFree(b->outtmp); cudaFree(b->scratch);
    cudaFree(b->boxes.u); cudaFree(b->boxes.force);
    for (auto &bd : b->bodies) bd.release();
    cudaStreamSynchronize(b->ibm_stream); cudaStreamSynchronize(b->ibm_main_stream); cudaStreamSynchronize(b->body_stream);
    cudaEventDestroy(b->ev_early); cudaStreamDestroy(b->ibm_main_stream); cudaStreamDestroy(b->body_stream);
    cudaEventDestroy(b->ev_macro); cudaEventDestroy(b->ev_io);
    cudaFree(b->bodies_dev); cudaFree(b->lead_dev); cudaFree(b->tol2); cudaFree(b->ctl); cudaFree(b->ibm_barrier);
    cudaFreeHost(b->ctl_pin); cudaEventDestroy(b->ev_ibm_done);
    cudaFree(b->mk_dev); cudaFree(b->force_dev); cudaFreeHost(b->mk_pin); cudaFreeHost(b->force_pin);
    cudaEventDestroy(b->ev_ibm); cudaStreamDestroy(b->ibm_stream);
    cudaFree(b->csr.count); cudaFree(b->csr.off); cudaFree(b->csr.entry); cudaFree(b->csr_scan_tmp); cudaFree(b->tol_partial);
    cudaEventDestroy(b->ev_edge); cudaEventDestroy(b->ev_comm); cudaEventDestroy(b->ev_pre);
    cudaStreamDestroy(b->comm_stream); cudaStreamDestroy(b->io_stream);
    g_blocks[h].reset();
    return 0;
}

int fsilbm_block_initialise(fsilbm_handle h, double time)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    const double Cs2 = 1.0 / 3.0;                            // ConstParams.f90:39
    b->tau = b->flow.nu / (b->g.dh * Cs2) + 0.5;             // calculate_SRT_params, FluidDomain.f90:452
    b->Omega = 1.0 / b->tau;
    if (b->model == 2) {                                     // calculate_TRT_params, :458-464
        const double lambda = b->params[0];
        const double tmp = (lambda * 4.0 - 1.0) * b->Omega + 2.0;
        b->Omega2 = 2.0 * (2.0 - b->Omega) / tmp;
    } else if (b->model == 3) {
        mrt_matrices(b->Omega, b->Mc, b->Mf);
        mrt_release(b->mrt_slot);
        bool fresh = false;
        b->mrt_slot = mrt_acquire(b->Mc, b->Mf, b->stream, &fresh);
        if (b->mrt_slot < 0)
            return fail(FSILBM_ERR_MODEL, "MRT: %d blocks with distinct relaxation matrices are alive in this process; the matrix table holds %d "
                                          "(blocks of equal dh and nu share an entry)", MRT_SLOTS + 1, MRT_SLOTS);
        if (fresh) {
            upload_mrt(b->mrt_slot, b->Mc, b->Mf, b->stream);
            CK(cudaStreamSynchronize(b->stream));   // Mc/Mf are pageable host memory
        }
    }
    if (b->model >= 11) {                                    // tau_all = tau, FluidDomain.f90:454-455
        const size_t n = (size_t)b->g.X * b->g.plane;
        if (!b->tau_all) CK(cudaMalloc(&b->tau_all, sizeof(double) * n));
        launch_pass_fill(b->tau_all, n, b->tau, b->stream);
        if (b->model != 11) if (int rc = ensure_fields(*b, false)) return rc;
    }
    b->blktime = time;
    VelocityField vel;
    if (int rc = velocity_field(*b, time, vel)) return rc;
    launch_initialise(b->g, b->f[b->cur], vel, b->flow.denIn, b->stream);
    // den/uuu of the first interior layers as initialise_ leaves them, for code 102 at the start-up BC call
    for (int face = 0; face < 6; face++) {
        if (b->bc[face] != BCnEq_DirecletU || !owns_face(*b, face)) continue;
        int na, nb;
        face_dims(b->g, face, na, nb);
        if (!b->l2den[face]) {
            CK(cudaMalloc(&b->l2den[face], sizeof(double) * (size_t)na * nb));
            CK(cudaMalloc(&b->l2u[face], sizeof(double) * 3 * (size_t)na * nb));
        }
        FaceParams p = face_params(*b, face, b->f[b->cur], b->f[b->cur], vel);
        launch_init_layer2(p, b->stream);
    }
    CK(cudaGetLastError());
    b->initialised = true;
    b->ibm_active = false;
    return 0;
}

int fsilbm_block_get(fsilbm_handle h, int what, double *value)
{
    Block *b = get(h);
    if (!b || !value) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    switch (what) {
    case 0: *value = b->tau; break;
    case 1: *value = b->Omega; break;
    case 2: *value = b->Omega2; break;
    default: return fail(FSILBM_ERR_ARG, "unknown scalar %d", what);
    }
    return 0;
}

static int await_deliveries(Block &b);   // sons across a slab interface: see fsilbm_block_collide_stream

int fsilbm_block_upload_fIn(fsilbm_handle h, const double *fIn)
{
    Block *b = get(h);
    if (!b || !fIn) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    b->early_ok = false;
    const Geom &g = b->g;
    const size_t n = (size_t)g.X * g.plane;
    // one strided copy (19 rows of X planes into rows of X + 2 planes), ordered after the work already queued on the block's
    // stream; from page-locked memory it runs at the link's rate without a host round trip per population
    if (sizeof(double) * g.pstride < ((size_t)1 << 31)) {   // cudaMemcpy2D pitches are limited to 2 GiB
        CK(cudaMemcpy2DAsync(b->f[b->cur] + g.plane, sizeof(double) * g.pstride, fIn, sizeof(double) * n, sizeof(double) * n, Q,
                             cudaMemcpyHostToDevice, b->stream));
    } else {
        for (int q = 0; q < Q; q++)
            CK(cudaMemcpyAsync(b->f[b->cur] + q * g.pstride + g.plane, fIn + q * n, sizeof(double) * n, cudaMemcpyHostToDevice, b->stream));
    }
    CK(cudaStreamSynchronize(b->stream));   // the caller may reuse fIn on return
    return 0;
}

int fsilbm_block_download_fIn(fsilbm_handle h, double *fIn)
{
    Block *b = get(h);
    if (!b || !fIn) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    if (int rc = await_deliveries(*b)) return rc;
    const Geom &g = b->g;
    const size_t n = (size_t)g.X * g.plane;
    if (sizeof(double) * g.pstride < ((size_t)1 << 31)) {
        CK(cudaMemcpy2DAsync(fIn, sizeof(double) * n, b->f[b->cur] + g.plane, sizeof(double) * g.pstride, sizeof(double) * n, Q,
                             cudaMemcpyDeviceToHost, b->stream));
    } else {
        for (int q = 0; q < Q; q++)
            CK(cudaMemcpyAsync(fIn + q * n, b->f[b->cur] + q * g.pstride + g.plane, sizeof(double) * n, cudaMemcpyDeviceToHost, b->stream));
    }
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

int fsilbm_block_set_time(fsilbm_handle h, double blktime)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->blktime = blktime;
    return 0;
}

int fsilbm_block_update_volume_force(fsilbm_handle h, double out[3])
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    const double Pi = 3.141592653589793;   // ConstParams.f90:36
    const fsilbm_flow &fl = b->flow;
    b->volumeForce[0] = fl.volumeForceIn[0] + fl.volumeForceAmp * sin(2.0 * Pi * fl.volumeForceFreq * b->blktime + fl.volumeForcePhi / 180.0 * Pi);
    b->volumeForce[1] = fl.volumeForceIn[1];
    b->volumeForce[2] = fl.volumeForceIn[2];
    if (out) for (int k = 0; k < 3; k++) out[k] = b->volumeForce[k];
    return 0;
}

int fsilbm_block_download_macro_async(fsilbm_handle h, double *den, double *uuu)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    if (int rc = io_wait(*b)) return rc;
    if (int rc = ensure_fields(*b, false)) return rc;
    if (int rc = await_deliveries(*b)) return rc;
    double hF[3];
    half_force(*b, hF);
    launch_macro_full(b->g, b->f[b->cur], hF, b->den, b->uuu, b->stream);
    CK(cudaGetLastError());
    CK(cudaEventRecord(b->ev_macro, b->stream));
    CK(cudaStreamWaitEvent(b->io_stream, b->ev_macro, 0));
    const size_t n = (size_t)b->g.X * b->g.plane;
    if (den) CK(cudaMemcpyAsync(den, b->den, sizeof(double) * n, cudaMemcpyDeviceToHost, b->io_stream));
    if (uuu) CK(cudaMemcpyAsync(uuu, b->uuu, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, b->io_stream));
    CK(cudaEventRecord(b->ev_io, b->io_stream));
    b->io_pending = true;
    return 0;
}

int fsilbm_block_download_wait(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    return io_wait(*b);
}

int fsilbm_block_download_macro(fsilbm_handle h, double *den, double *uuu)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    if (int rc = io_wait(*b)) return rc;
    if (int rc = ensure_fields(*b, false)) return rc;
    if (int rc = await_deliveries(*b)) return rc;
    double hF[3];
    half_force(*b, hF);
    launch_macro_full(b->g, b->f[b->cur], hF, b->den, b->uuu, b->stream);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(b->stream));
    const size_t n = (size_t)b->g.X * b->g.plane;
    if (den) CK(cudaMemcpy(den, b->den, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (uuu) CK(cudaMemcpy(uuu, b->uuu, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    return 0;
}

int fsilbm_block_download_tau_all(fsilbm_handle h, double *tau_all)
{
    Block *b = get(h);
    if (!b || !tau_all) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    CK(cudaStreamSynchronize(b->stream));
    const size_t n = (size_t)b->g.X * b->g.plane;
    if (b->tau_all) { CK(cudaMemcpy(tau_all, b->tau_all, sizeof(double) * n, cudaMemcpyDeviceToHost)); }
    else for (size_t i = 0; i < n; i++) tau_all[i] = b->tau;   // constant-tau models, FluidDomain.f90:454-455
    return 0;
}

int fsilbm_block_field_stat(fsilbm_handle h, double out[6])
{
    Block *b = get(h);
    if (!b || !out) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    double hF[3];
    half_force(*b, hF);
    CK(cudaMemsetAsync(b->stat, 0, sizeof(double) * 6, b->stream));
    launch_field_stat(b->g, b->f[b->cur], hF, 1.0 / b->flow.Uref, b->stat, b->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, b->stat, sizeof(double) * 6, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

int fsilbm_block_set_boundary_conditions(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    return apply_boundary_conditions(*b, b->f[b->cur]);
}

// Sons across a slab interface (see Pair).  A rank into whose slab a neighbour's son reaches must not touch the planes the son
// rewrites before the neighbour's son->father delivery of the last finished step has landed: one-thread wait on the block's stream.
static int await_deliveries(Block &b)
{
    Block::Halo &h = b.halo;
    if (!h.enabled) return 0;
    for (int sd = 0; sd < 2; sd++) {
        if (h.remote_pairs[sd] <= 0 || h.step <= h.remote_base[sd]) continue;
        HaloWaitParams u{};
        u.flag_lo = refine_flag(h.region, 1, sd);
        u.step = (h.step - h.remote_base[sd]) * (unsigned long long)h.remote_pairs[sd];
        u.err = h.err; u.timeout_ns = (unsigned long long)g_halo_timeout_s * 1000000000ull;
        launch_halo_wait(u, b.stream);
    }
    return 0;
}
// ... and tells the owner of such a son when its father step is complete (collide-stream, halo, face kernels)
static void announce_father_step(Block &b)
{
    Block::Halo &h = b.halo;
    if (!h.enabled) return;
    for (int sd = 0; sd < 2; sd++) {
        if (h.remote_pairs[sd] <= 0) continue;
        unsigned char *owner_region = sd == 0 ? h.peer_left : h.peer_right;
        launch_flag_signal(refine_flag(owner_region, 0, sd ^ 1), h.step, b.stream);   // the owner sees this rank on its other side
    }
}
// The owner's side: wait until the neighbours whose planes a cross pair touches have finished the father step this rank is at
static void await_father_step(Block &F, const Pair &p)
{
    Block::Halo &h = F.halo;
    for (int sd = 0; sd < 2; sd++) {
        if (!p.cross[sd] || h.step == 0) continue;
        HaloWaitParams u{};
        u.flag_lo = refine_flag(h.region, 0, sd);
        u.step = h.step;
        u.err = h.err; u.timeout_ns = (unsigned long long)g_halo_timeout_s * 1000000000ull;
        launch_halo_wait(u, g_stream);
    }
}

int fsilbm_block_collide_stream(fsilbm_handle h)
{
    Block *bp = get(h);
    if (!bp) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    Block &b = *bp;
    if (!b.initialised) return fail(FSILBM_ERR_ARG, "block %d not initialised", h);
    if (int rc = await_deliveries(b)) return rc;
    const Geom &g = b.g;
    const double *fA = b.f[b.cur];
    double *fB = b.f[b.cur ^ 1];
    VelocityField vel;
    if (int rc = velocity_field(b, b.blktime, vel)) return rc;
    const bool multi = g_nccl.nranks > 1 && g_nccl.comm && g.X != g.XG;   // a block cut into x-slabs (sons stay whole on one rank)
    const bool ghost = multi || g_force_ghost;
    // Early IBM (see Block::ev_early): planes A = [box - 2, box + 2) of every stencil box are updated apart from the rest.  After A
    // the streamed populations of the planes [box - 1, box + 1) are final -- provided no face kernel, halo or x-wrap touches them,
    // hence the conditions below -- and the next interaction-force call may start while the rest is still being updated.
    int nA = 0, A0[MAX_BOXES], A1[MAX_BOXES], V0[MAX_BOXES], V1[MAX_BOXES];
    const int lower = multi ? 1 : 0, upper = multi ? g.X - 1 : g.X;   // planes between the edge planes (multi: those go first anyway)
    const IbmBoxes &bxs = b.boxes;
    bool early = g_ibm_early && b.ibm_active && bxs.n > 0 && b.model < 11 && !b.is_father && (multi ? b.halo.enabled : !ghost);
    bool edge_dep = false;   // slab runs: a box reaches the slab's edge planes, whose final values also need the neighbour's halo
    if (early) {
        const int Ns[3] = {g.XG, g.Y, g.Z};
        for (int i = 0; i < bxs.n && early; i++) {
            for (int k = 0; k < 3; k++) {
                const int lo = bxs.lo[i][k], hi = lo + bxs.ext[i][k];
                if (k == 0 && hi > Ns[0]) early = false;                              // wraps in x
                if (b.periodic[k] != 1 && (lo < 3 || hi > Ns[k] - 3)) early = false;    // within reach of a face kernel
            }
            int a0 = bxs.lo[i][0] - g.xOffset - 2, a1 = bxs.lo[i][0] + bxs.ext[i][0] - g.xOffset + 2;
            bool cut_lo = false, cut_hi = false;
            if (multi) {   // a box across (or next to) a slab interface: this rank's share of it, up to the edge plane
                if (a0 < lower) { a0 = lower; cut_lo = true; }
                if (a1 > upper) { a1 = upper; cut_hi = true; }
                if (a1 <= a0) early = false;
            } else if (a0 < lower || a1 > upper) early = false;
            A0[nA] = a0; A1[nA] = a1;
            V0[nA] = cut_lo ? 0 : a0 + 1; V1[nA] = cut_hi ? g.X : a1 - 1;
            edge_dep = edge_dep || cut_lo || cut_hi;
            nA++;
        }
        if (early) {   // sort and merge
            for (int i = 1; i < nA; i++)
                for (int j = i; j > 0 && A0[j] < A0[j - 1]; j--) {
                    std::swap(A0[j], A0[j - 1]); std::swap(A1[j], A1[j - 1]); std::swap(V0[j], V0[j - 1]); std::swap(V1[j], V1[j - 1]);
                }
            int m = 0;
            for (int i = 1; i < nA; i++) {
                if (A0[i] <= A1[m]) { A1[m] = std::max(A1[m], A1[i]); V0[m] = std::min(V0[m], V0[i]); V1[m] = std::max(V1[m], V1[i]); }
                else { m++; A0[m] = A0[i]; A1[m] = A1[i]; V0[m] = V0[i]; V1[m] = V1[i]; }
            }
            nA = m + 1;
        }
    }
    b.early_ok = false;
    // An interaction-force call still in flight on another stream: whatever reads its box fields follows it ON THE DEVICE.  On one
    // GPU with early IBM that is only the launch over the planes A: it goes to the block's high-priority body stream behind that
    // call, while the rest of the update -- which holds no box cell -- starts on the compute stream at once.  If the iteration is
    // over when the update begins, the body stream's CTAs are scheduled first and A is finished first, as if the two launches were
    // queued one after the other; if it is not (a large body in a small block), the device updates the rest meanwhile instead of
    // idling, and A still finishes as early as its input allows, which is what the next interaction-force call waits for.
    const bool ibm_inflight = b.ibm_pending.active && b.ibm_pending.stream != b.stream;
    // (Slab runs keep the two launches on the compute stream: there the planes A of a body across an interface also wait for the
    // neighbour's edge planes, and with the rest already filling the SMs those arrived 0.4 ms later -- heave1024 x2 1.71 -> 1.89 ms.)
    const bool split = early && !multi && g_update_split;
    cudaStream_t sA = split ? b.body_stream : b.stream;
    if (split) {
        CK(cudaEventRecord(b.ev_pre, b.stream));          // the previous step (and whatever else was queued on the compute stream)
        CK(cudaStreamWaitEvent(sA, b.ev_pre, 0));
    }
    if (ibm_inflight) CK(cudaStreamWaitEvent(sA, b.ev_ibm_done, 0));

    if (b.model == 14 || b.model == 15) {
        // WALE / Vreman difference the velocity of this step across neighbouring cells (FluidDomain.f90:1343-1385,
        // 1445-1484): materialise uuu (with the IBM correction where a body is near) before the fused kernel.
        double hF0[3];
        half_force(b, hF0);
        if (int rc = io_wait(b)) return rc;
        if (g.X != g.XG) {
            // x-slab: the differences across a slab interface need the neighbour's edge plane of THIS step's velocity (the
            // reference switches to one-sided differences only at the domain faces, FluidDomain.f90:1343-1385): one plane each way
            if (!(g_nccl.nranks > 1 && g_nccl.comm)) return fail(FSILBM_ERR_COMM, "a slab of a WALE/Vreman block needs the communicator (fsilbm_comm_init)");
            const size_t nc = (size_t)(g.X + 2) * g.plane;
            if (!b.uuu_les) { CK(cudaMalloc(&b.uuu_les, sizeof(double) * 3 * nc)); CK(cudaMemsetAsync(b.uuu_les, 0, sizeof(double) * 3 * nc, b.stream)); }
            launch_macro_full(g, fA, hF0, nullptr, b.uuu_les + g.plane, b.stream, b.ibm_active ? &b.boxes : nullptr, nc);
            const int R = g_nccl.nranks, r = g_nccl.rank;
            const int right = r + 1 < R ? r + 1 : -1, left = r > 0 ? r - 1 : -1;   // no wrap: the reference never differences across a periodic x face
            NCK(g_nccl.GroupStart());
            for (int k = 0; k < 3; k++) {
                double *u = b.uuu_les + (size_t)k * nc;
                if (right >= 0) NCK(g_nccl.Send(u + (size_t)g.X * g.plane, g.plane, kNcclFloat64, right, g_nccl.comm, b.stream));        // my plane X-1
                if (left >= 0) NCK(g_nccl.Send(u + g.plane, g.plane, kNcclFloat64, left, g_nccl.comm, b.stream));                        // my plane 0
                if (left >= 0) NCK(g_nccl.Recv(u, g.plane, kNcclFloat64, left, g_nccl.comm, b.stream));                                  // ghost -1
                if (right >= 0) NCK(g_nccl.Recv(u + (size_t)(g.X + 1) * g.plane, g.plane, kNcclFloat64, right, g_nccl.comm, b.stream));  // ghost X
            }
            NCK(g_nccl.GroupEnd());
        } else {
            launch_macro_full(g, fA, hF0, nullptr, b.uuu, b.stream, b.ibm_active ? &b.boxes : nullptr);
        }
    }
    // per-face side buffers taken from the pre-collision state (see kernels.h FaceParams)
    for (int face = 0; face < 6; face++) {
        if (!owns_face(b, face)) continue;
        const int code = b.bc[face];
        if (code == BCnEq_DirecletU) {
            int na, nb;
            face_dims(g, face, na, nb);
            if (!b.l2den[face]) {
                CK(cudaMalloc(&b.l2den[face], sizeof(double) * (size_t)na * nb));
                CK(cudaMalloc(&b.l2u[face], sizeof(double) * 3 * (size_t)na * nb));
            }
            FaceParams p = face_params(b, face, fB, fA, vel);
            launch_layer2_face(p, sA);
        } else if ((code == BCstationary_Wall_halfway || code == BCmoving_Wall_halfway) && b.hw_alloc[face]) {
            FaceParams p = face_params(b, face, fB, fA, vel);
            launch_stash_face(p, sA);   // halfwayBCset_, LBMBlockComm.f90:296
        }
    }

    StepParams p{};
    p.g = g; p.fA = fA; p.fB = fB;
    p.cc = collide_consts(b);
    half_force(b, p.hF);
    for (int k = 0; k < 3; k++) p.Fvol[k] = b.volumeForce[k];
    p.boxes = b.boxes;
    if (!b.ibm_active) p.boxes.n = 0;
    p.tau_all = b.tau_all;
    les_field(b, p.uuu, p.uuu_ncomp);
    p.wrap_x = ghost ? 0 : 1;
    // The launch list (StepParams::seg_*): [edge planes of a slab] [planes A around the bodies] | [the rest].  Without early IBM
    // the whole update is ONE launch; with it, two: everything up to A, then -- event ev_early recorded in between -- the rest,
    // which holds no box cell (A covers every box with two planes to spare) and takes the IBM-free instantiation.
    auto planes_first = [&](StepParams &q) {
        if (multi) { step_add_planes(q, 0, 1); if (g.X > 1) step_add_planes(q, g.X - 1, 1); }
        if (early) for (int i = 0; i < nA; i++) step_add_planes(q, A0[i], A1[i] - A0[i]);
        else step_add_planes(q, lower, upper - lower);
    };
    auto planes_rest = [&](StepParams &q) {
        int at = lower;
        for (int i = 0; i <= nA; i++) {
            const int end = i < nA ? A0[i] : upper;
            step_add_planes(q, at, end - at);
            if (i < nA) at = A1[i];
        }
    };
    auto note_early = [&]() {
        b.early_n = nA;
        for (int i = 0; i < nA; i++) { b.early_x0[i] = V0[i] + g.xOffset; b.early_x1[i] = V1[i] + g.xOffset; }
        b.early_ok = true;
    };
    auto refuse = [&]() { return fail(FSILBM_ERR_MODEL, "collision model %d", b.model); };
    if (!multi) {
        StepParams q = p;
        planes_first(q);
        TRACE(b.stream, 0, "step_begin");
        if (launch_collide_push(q, b.model, sA)) return refuse();
        TRACE(sA, split ? 4 : 0, early ? "collide_A" : "collide_all");
        if (early) {
            CK(cudaEventRecord(b.ev_early, sA));
            StepParams rest = p;
            rest.boxes.n = 0;
            planes_rest(rest);
            if (launch_collide_push(rest, b.model, b.stream)) return refuse();
            TRACE(b.stream, 0, "collide_rest");
            if (split) CK(cudaStreamWaitEvent(b.stream, b.ev_early, 0));   // the two streams meet before the face kernels
            note_early();
        }
        if (ghost) launch_wrap_x(g, fB, b.stream);
    } else if (b.halo.enabled) {
        // Edge stream (high priority) beside the compute stream.  Edge stream: the two edge planes -- their CTAs ARE the
        // transfer: what leaves the slab is stored straight into the neighbour's streamed buffer over NVLink, the last CTA of a
        // plane raises the neighbour's arrival flag -- then a one-thread kernel that waits for the neighbours' flags of this
        // step.  Compute stream: every other plane, in one or two launches of the plain kernel.  The two streams meet before the
        // face kernels.  (Folding the edge planes into the main launch made every CTA of it pay for the edge instantiation:
        // 0.84 ms per step on channel256 x2 against 0.78 on one GPU.)
        Block::Halo &h = b.halo;
        h.step++;
        const int par = (int)(h.step & 1);
        cudaStream_t es = b.comm_stream;
        const int nxt = b.cur ^ 1;
        StepParams e = p;
        e.step = h.step;
        if (h.left >= 0) {    // ex = -1 populations of plane 0 -> the left neighbour's plane X_left - 1 (xp = X_left)
            e.halo_lo = h.peer_f[0][nxt] + (size_t)h.peer_X[0] * g.plane; e.halo_lo_ps = (size_t)(h.peer_X[0] + 2) * g.plane;
            e.sig_lo = halo_flag(h.peer_left, 1, par);
        }
        if (h.right >= 0) {   // ex = +1 populations of plane X-1 -> the right neighbour's plane 0 (xp = 1)
            e.halo_hi = h.peer_f[1][nxt] + g.plane; e.halo_hi_ps = (size_t)(h.peer_X[1] + 2) * g.plane;
            e.sig_hi = halo_flag(h.peer_right, 0, par);
        }
        e.cta_counter = h.counters;
        step_add_planes(e, 0, 1);
        if (g.X > 1) step_add_planes(e, g.X - 1, 1);
        CK(cudaEventRecord(b.ev_pre, b.stream));          // the previous step and this step's pre-collision face kernels
        CK(cudaStreamWaitEvent(es, b.ev_pre, 0));
        if (launch_collide_push(e, b.model, es)) return refuse();
        TRACE(es, 3, "collide_edges");
        HaloWaitParams u{};
        u.step = h.step; u.err = h.err; u.timeout_ns = (unsigned long long)g_halo_timeout_s * 1000000000ull;
        if (h.left >= 0) u.flag_lo = halo_flag(h.region, 0, par);
        if (h.right >= 0) u.flag_hi = halo_flag(h.region, 1, par);
        launch_halo_wait(u, es);
        TRACE(es, 3, "halo_arrived");
        CK(cudaEventRecord(b.ev_edge, es));
        TRACE(b.stream, 0, "step_begin");
        StepParams q = p;
        if (early) for (int i = 0; i < nA; i++) step_add_planes(q, A0[i], A1[i] - A0[i]);
        else step_add_planes(q, lower, upper - lower);
        if (launch_collide_push(q, b.model, b.stream)) return refuse();
        TRACE(b.stream, 0, early ? "collide_A" : "collide_inner");
        if (early) {
            // a box across the interface: its planes are final only once the neighbour's populations have arrived
            if (edge_dep) CK(cudaStreamWaitEvent(b.stream, b.ev_edge, 0));
            CK(cudaEventRecord(b.ev_early, b.stream));
            StepParams rest = p;
            rest.boxes.n = 0;
            planes_rest(rest);
            if (launch_collide_push(rest, b.model, b.stream)) return refuse();
            TRACE(b.stream, 0, "collide_rest");
            note_early();
        }
        CK(cudaStreamWaitEvent(b.stream, b.ev_edge, 0));
    } else {
        // NCCL transport: edge planes first, then the exchange on its own stream overlapped with the interior update
        StepParams e = p;
        step_add_planes(e, 0, 1);
        if (g.X > 1) step_add_planes(e, g.X - 1, 1);
        if (launch_collide_push(e, b.model, b.stream)) return refuse();
        CK(cudaEventRecord(b.ev_edge, b.stream));
        CK(cudaStreamWaitEvent(b.comm_stream, b.ev_edge, 0));
        if (int rc = halo_exchange(b, fB, b.comm_stream)) return rc;
        CK(cudaEventRecord(b.ev_comm, b.comm_stream));
        StepParams in = p;
        step_add_planes(in, 1, g.X - 2);
        if (launch_collide_push(in, b.model, b.stream)) return refuse();
        CK(cudaStreamWaitEvent(b.stream, b.ev_comm, 0));
    }
    CK(cudaGetLastError());
    if (int rc = apply_boundary_conditions(b, fB)) return rc;   // LBMBlockComm.f90:303
    TRACE(b.stream, 0, "faces");
    announce_father_step(b);
    b.cur ^= 1;
    b.ibm_active = false;
    g_update_clock_us += (double)g.X * (double)g.plane * 304.0 / 6.2e6;
    b.update_clock_us = g_update_clock_us;
    return 0;
}

int fsilbm_block_sync(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    if (int rc = await_deliveries(*b)) return rc;   // a neighbour's son may still be rewriting father nodes of this slab
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaStreamSynchronize(b->comm_stream));
    CK(cudaStreamSynchronize(b->io_stream));
    b->io_pending = false;
    if (b->halo.enabled) {
        int e = 0;
        CK(cudaMemcpy(&e, b->halo.err, sizeof(int), cudaMemcpyDeviceToHost));
        if (e) return fail(FSILBM_ERR_COMM, "halo: a neighbour's arrival flag did not come within %d s (rank %d)", g_halo_timeout_s, g_nccl.rank);
    }
    return 0;
}

int fsilbm_block_halo_transport(fsilbm_handle h, int *mode)
{
    Block *b = get(h);
    if (!b || !mode) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    *mode = (g_nccl.nranks > 1 && g_nccl.comm && b->g.X != b->g.XG) ? (b->halo.enabled ? 2 : 1) : 0;
    return 0;
}

int fsilbm_block_stream(fsilbm_handle h, void **stream)
{
    Block *b = get(h);
    if (!b || !stream) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    *stream = (void *)b->stream;
    return 0;
}

// ---- output / diagnostics computed on the device state ------------------------------------
static int flow_window(fsilbm_handle h, int offsetOutput, int outputtype, float *out, bool async)
{
    Block *b = get(h);
    if (!b || !out) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    if (outputtype < 1) return 0;   // FluidDomain.f90:1639
    if (int rc = io_wait(*b)) return rc;   // an earlier asynchronous read-back may still be copying out of the staging buffer
    const Geom &g = b->g;
    // window in global x: [off, XG-off), intersected with the local slab
    const int gx0 = std::max(offsetOutput, g.xOffset), gx1 = std::min(g.XG - offsetOutput, g.xOffset + g.X);
    FlowWindowParams p{};
    p.g = g; p.f = b->f[b->cur];
    p.x0 = gx0 - g.xOffset; p.off = offsetOutput;
    p.nx = gx1 - gx0; p.ny = g.Y - 2 * offsetOutput; p.nz = g.Z - 2 * offsetOutput;
    if (p.nx <= 0 || p.ny <= 0 || p.nz <= 0) return 0;
    p.outputtype = outputtype;
    if (outputtype >= 2 && !b->uuu_ave) {
        // the reference allocates uuu_ave with the block (FluidDomain.f90:390) and zeroes it in initialise_ (:539): its first write_flow_blocks call
        // (main.f90:84, before any step) writes those zeros
        const size_t nave = (size_t)g.X * g.plane;
        CK(cudaMalloc(&b->uuu_ave, sizeof(double) * 9 * nave));
        CK(cudaMemsetAsync(b->uuu_ave, 0, sizeof(double) * 9 * nave, b->stream));
    }
    p.uuu_ave = b->uuu_ave;
    half_force(*b, p.hF);
    p.denIn = b->flow.denIn; p.invUref = 1.0 / b->flow.Uref; p.invUrefs = 1.0 / b->flow.Uref / b->flow.Uref;   // :1650-1651
    const int nfields = outputtype >= 2 ? 13 : 4;
    const size_t n = (size_t)p.nx * p.ny * p.nz * nfields;
    if (n > b->outtmp_cap) {
        CK(cudaStreamSynchronize(b->stream));
        cudaFree(b->outtmp);
        CK(cudaMalloc(&b->outtmp, sizeof(float) * n));
        b->outtmp_cap = n;
    }
    p.out = b->outtmp;
    if (outputtype == 2) CK(cudaMemsetAsync(b->outtmp, 0, sizeof(float) * 4 * (n / nfields), b->stream));   // fields 0:3 are not refreshed (:1653)
    launch_flow_window(p, b->stream);
    CK(cudaGetLastError());
    if (async) {   // the copy leaves on the copy stream while later steps run; fsilbm_block_download_wait collects it
        CK(cudaEventRecord(b->ev_macro, b->stream));
        CK(cudaStreamWaitEvent(b->io_stream, b->ev_macro, 0));
        CK(cudaMemcpyAsync(out, b->outtmp, sizeof(float) * n, cudaMemcpyDeviceToHost, b->io_stream));
        CK(cudaEventRecord(b->ev_io, b->io_stream));
        b->io_pending = true;
        return 0;
    }
    CK(cudaMemcpyAsync(out, b->outtmp, sizeof(float) * n, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

int fsilbm_block_write_flow_window(fsilbm_handle h, int offsetOutput, int outputtype, float *out) { return flow_window(h, offsetOutput, outputtype, out, false); }
int fsilbm_block_write_flow_window_async(fsilbm_handle h, int offsetOutput, int outputtype, float *out) { return flow_window(h, offsetOutput, outputtype, out, true); }

int fsilbm_block_turbulent_statistic(fsilbm_handle h, int step, int step_s)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    if (step < step_s) return 0;   // :1152 (the outputtype >= 2 half of the test is the caller's)
    const size_t n = (size_t)b->g.X * b->g.plane;
    if (!b->uuu_ave) {
        CK(cudaMalloc(&b->uuu_ave, sizeof(double) * 9 * n));
        CK(cudaMemsetAsync(b->uuu_ave, 0, sizeof(double) * 9 * n, b->stream));
    }
    const float invStepF = 1 / (float)(step - step_s + 1);   // 'invStep = 1 / real(step - step_s + 1)': default real is real(4), :1153
    double hF[3];
    half_force(*b, hF);
    launch_turbulent_statistic(b->g, b->f[b->cur], hF, b->uuu_ave, (double)invStepF, b->stream);
    CK(cudaGetLastError());
    return 0;
}

static int ensure_scratch(Block &b, size_t n)
{
    if (n <= b.scratch_cap) return 0;
    CK(cudaStreamSynchronize(b.stream));
    cudaFree(b.scratch);
    CK(cudaMalloc(&b.scratch, sizeof(double) * n));
    b.scratch_cap = n;
    return 0;
}

int fsilbm_block_fluid_flux(fsilbm_handle h, double out[3])
{
    Block *b = get(h);
    if (!b || !out) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    const Geom &g = b->g;
    const int gx[3] = {0, (g.XG + 1) / 2 - 1, g.XG - 1};   // x = 1, ixMid = (xDim+1)/2, xDim (1-based), :2032
    int xl[3];
    for (int k = 0; k < 3; k++) { xl[k] = gx[k] - g.xOffset; if (xl[k] < 0 || xl[k] >= g.X) xl[k] = -1; }
    if (int rc = ensure_scratch(*b, 64)) return rc;
    double hF[3];
    half_force(*b, hF);
    launch_fluid_flux(g, b->f[b->cur], hF, xl, b->scratch, b->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, b->scratch, sizeof(double) * 3, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

int fsilbm_block_probe_velocity(fsilbm_handle h, int n, const double *coords, double *velocity)
{
    Block *b = get(h);
    if (!b || n < 0 || (n > 0 && (!coords || !velocity))) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    if (n == 0) return 0;
    const Geom &g = b->g;
    const double mx[3] = {g.xmin + g.dh * (g.XG - 1), g.ymin + g.dh * (g.Y - 1), g.zmin + g.dh * (g.Z - 1)}, mn[3] = {g.xmin, g.ymin, g.zmin};
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++)
            if (coords[3 * i + k] < mn[k] || coords[3 * i + k] > mx[k])
                return fail(FSILBM_ERR_ARG, "fluid probe %d is not in selected block (FlowCondition.f90:212)", i + 1);
    if (int rc = ensure_scratch(*b, (size_t)6 * n)) return rc;
    CK(cudaMemcpyAsync(b->scratch, coords, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, b->stream));
    double hF[3];
    half_force(*b, hF);
    launch_probe(g, b->f[b->cur], hF, n, b->scratch, b->scratch + 3 * n, b->stream);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(velocity, b->scratch + 3 * n, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    return 0;
}

// ---- un-fused passes ------------------------------------------------------------------------------
static FieldParams field_params(Block &b)
{
    FieldParams p{};
    p.g = b.g; p.f = b.f[b.cur]; p.den = b.den; p.uuu = b.uuu; p.force = b.force;
    p.cc = collide_consts(b);
    half_force(b, p.hF);
    for (int k = 0; k < 3; k++) p.Fvol[k] = b.volumeForce[k];
    p.tau_all = b.tau_all;
    return p;
}

int fsilbm_block_pass_macro(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    if (int rc = ensure_fields(*b, true)) return rc;
    double hF[3];
    half_force(*b, hF);
    launch_macro_full(b->g, b->f[b->cur], hF, b->den, b->uuu, b->stream);
    CK(cudaGetLastError());
    return 0;
}

int fsilbm_block_pass_reset_volume_force(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    if (int rc = ensure_fields(*b, true)) return rc;
    launch_pass_fill(b->force, 3 * (size_t)b->g.X * b->g.plane, 0.0, b->stream);
    CK(cudaGetLastError());
    return 0;
}

int fsilbm_block_pass_add_volume_force(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    if (int rc = ensure_fields(*b, true)) return rc;
    launch_pass_add_force(field_params(*b), b->stream);
    CK(cudaGetLastError());
    return 0;
}

int fsilbm_block_pass_collision(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    if (!b->initialised) return fail(FSILBM_ERR_ARG, "block %d not initialised", h);
    if (int rc = ensure_fields(*b, true)) return rc;
    if (launch_pass_collision(field_params(*b), b->model, b->stream)) return fail(FSILBM_ERR_MODEL, "collision model %d", b->model);
    CK(cudaGetLastError());
    return 0;
}

int fsilbm_block_pass_halfway_bc_set(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    VelocityField vel;
    if (int rc = velocity_field(*b, b->blktime, vel)) return rc;
    for (int face = 0; face < 6; face++) {
        const int code = b->bc[face];
        if (!(code == BCstationary_Wall_halfway || code == BCmoving_Wall_halfway) || !b->hw_alloc[face] || !owns_face(*b, face)) continue;
        FaceParams p = face_params(*b, face, b->f[b->cur], b->f[b->cur], vel);
        launch_pass_halfway(p, b->stream);
    }
    CK(cudaGetLastError());
    return 0;
}

int fsilbm_block_pass_streaming(fsilbm_handle h)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    if (b->g.X != b->g.XG) return fail(FSILBM_ERR_ARG, "the un-fused streaming pass is single-slab only");
    launch_pass_streaming(b->g, b->f[b->cur], b->f[b->cur ^ 1], b->stream);
    CK(cudaGetLastError());
    b->cur ^= 1;
    return 0;
}

int fsilbm_block_download_fields(fsilbm_handle h, double *den, double *uuu, double *force)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    if (int rc = ensure_fields(*b, true)) return rc;
    CK(cudaStreamSynchronize(b->stream));
    const size_t n = (size_t)b->g.X * b->g.plane;
    if (den) CK(cudaMemcpy(den, b->den, sizeof(double) * n, cudaMemcpyDeviceToHost));
    if (uuu) CK(cudaMemcpy(uuu, b->uuu, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    if (force) CK(cudaMemcpy(force, b->force, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
    return 0;
}

int fsilbm_block_upload_fields(fsilbm_handle h, const double *den, const double *uuu, const double *force)
{
    Block *b = get(h);
    if (!b) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    b->early_ok = false;
    if (int rc = ensure_fields(*b, true)) return rc;
    CK(cudaStreamSynchronize(b->stream));
    const size_t n = (size_t)b->g.X * b->g.plane;
    if (den) CK(cudaMemcpy(b->den, den, sizeof(double) * n, cudaMemcpyHostToDevice));
    if (uuu) CK(cudaMemcpy(b->uuu, uuu, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    if (force) CK(cudaMemcpy(b->force, force, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
    return 0;
}

// ---- IBM ------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// Device storage of the bodies of a block: per-body scratch + one packed marker buffer and one packed force buffer (a single
// H2D / D2H each, staged through pinned host memory so that the copies are truly asynchronous).
int ibm_body_storage(Block &b, int nbody, const int *nelmts)
{
    cudaStream_t s = b.stream, s2 = b.ibm_stream;
    bool relayout = (int)b.bodies.size() != nbody;
    if (relayout) {
        for (auto &bd : b.bodies) bd.release();
        b.bodies.assign(nbody, BodyDev());
        b.csr_valid = false;
    }
    for (int ib = 0; ib < nbody; ib++) {
        BodyDev &bd = b.bodies[ib];
        const int n = nelmts[ib];
        if (n < 1) return fail(FSILBM_ERR_ARG, "body %d has no markers", ib);
        if (bd.n != n) {
            bd.release();
            b.csr_valid = false;
            relayout = true;
            bd.n = n;
            CK(cudaMalloc(&bd.ExyzStencil, sizeof(double) * 3 * n));
            CK(cudaMalloc(&bd.felt, sizeof(double) * 3 * n));
            CK(cudaMalloc(&bd.tol, sizeof(double) * n)); CK(cudaMalloc(&bd.partialU, sizeof(double) * 3 * n));
            CK(cudaMalloc(&bd.Ei, sizeof(short) * 12 * n)); CK(cudaMalloc(&bd.Ew, sizeof(float) * 12 * n));
            CK(cudaMalloc(&bd.cell, sizeof(int) * 12 * n)); CK(cudaMalloc(&bd.boff, sizeof(long long) * n));
            CK(cudaMalloc(&bd.owned, 4 * n));
        }
    }
    if (!relayout) return 0;
    size_t ntot = 0;
    b.mk_off.assign(nbody, 0); b.f_off.assign(nbody, 0);
    for (int ib = 0; ib < nbody; ib++) { b.mk_off[ib] = 7 * ntot; b.f_off[ib] = 3 * ntot; ntot += (size_t)nelmts[ib]; }
    b.marker_total = ntot;
    if (ntot > b.marker_cap) {
        CK(cudaStreamSynchronize(s)); CK(cudaStreamSynchronize(s2));
        cudaFree(b.mk_dev); cudaFree(b.force_dev); cudaFreeHost(b.mk_pin); cudaFreeHost(b.force_pin);
        b.mk_dev = b.force_dev = b.mk_pin = b.force_pin = nullptr;
        CK(cudaMalloc(&b.mk_dev, sizeof(double) * 7 * ntot)); CK(cudaMalloc(&b.force_dev, sizeof(double) * 3 * ntot));
        CK(cudaMallocHost(&b.mk_pin, sizeof(double) * 7 * ntot)); CK(cudaMallocHost(&b.force_pin, sizeof(double) * 3 * ntot));
        b.marker_cap = ntot;
    }
    for (int ib = 0; ib < nbody; ib++) {
        BodyDev &bd = b.bodies[ib];
        const size_t n = bd.n;
        bd.Exyz = b.mk_dev + b.mk_off[ib]; bd.Evel = bd.Exyz + 3 * n; bd.Ea = bd.Evel + 3 * n;
        bd.Eforce = b.force_dev + b.f_off[ib];
    }
    return 0;
}

// Box around the stencils of one body, from the marker positions P(3,n) the stencils are (re)built from: the index arithmetic
// of UpdateElmtInterp_ (Solidbody.f90:771-780,817-820).  floor((x - x0)*invdh) does not decrease with x, so the extreme
// indices are those of the extreme coordinates.
void ibm_body_box(const Geom &g, BodyDev &bd, const double *P, const int rootBC[6])
{
    const double invdh = 1.0 / g.dh;
    const double mins[3] = {g.xmin, g.ymin, g.zmin};
    const int Ns[3] = {g.XG, g.Y, g.Z};
    double lo[3] = {P[0], P[1], P[2]}, hi[3] = {P[0], P[1], P[2]};
    for (int e = 1; e < bd.n; e++) {
        const double v0 = P[3 * e], v1 = P[3 * e + 1], v2 = P[3 * e + 2];
        lo[0] = std::min(lo[0], v0); hi[0] = std::max(hi[0], v0);
        lo[1] = std::min(lo[1], v1); hi[1] = std::max(hi[1], v1);
        lo[2] = std::min(lo[2], v2); hi[2] = std::max(hi[2], v2);
    }
    for (int a = 0; a < 3; a++) {
        int i0 = (int)floor((P[a] - mins[a]) * invdh);
        const double x0 = mins[a] + (double)i0 * g.dh;
        bd.cidx[a] = imod(i0, Ns[a]);
        i0 = i0 + 1;
        const int imin = (int)floor((lo[a] - x0) * invdh) + i0, imax = (int)floor((hi[a] - x0) * invdh) + i0;
        bd.hbox.ax[a] = axis_interval(imin, imax, Ns[a], rootBC[2 * a] == BCPeriodic);
    }
    bd.have_box = true;
}

// Boxes that overlap are merged so that bodies sharing cells share storage (Gauss-Seidel coupling, Solidbody.f90:898-903).
// Box order = order of the first member; members in body order.
struct MBox { HostBox hb; std::vector<int> members; };
std::vector<MBox> ibm_merge_boxes(const Block &b, int nbody)
{
    const int Ns[3] = {b.g.XG, b.g.Y, b.g.Z};
    std::vector<MBox> mb(nbody);
    for (int ib = 0; ib < nbody; ib++) { mb[ib].hb = b.bodies[ib].hbox; mb[ib].members.assign(1, ib); }
    for (bool changed = true; changed;) {
        changed = false;
        for (size_t i = 0; i < mb.size() && !changed; i++)
            for (size_t j = i + 1; j < mb.size() && !changed; j++) {
                bool ov = true;
                for (int a = 0; a < 3; a++) ov = ov && overlap(mb[i].hb.ax[a], mb[j].hb.ax[a], Ns[a]);
                if (ov) {
                    for (int a = 0; a < 3; a++) mb[i].hb.ax[a] = merge(mb[i].hb.ax[a], mb[j].hb.ax[a], Ns[a]);
                    mb[i].members.insert(mb[i].members.end(), mb[j].members.begin(), mb[j].members.end());
                    mb.erase(mb.begin() + j);
                    changed = true;
                }
            }
    }
    return mb;
}

// Slab runs: the rank that owns each global x-plane of the block (collective over the ranks, done once per block).
int ibm_plane_owners(Block &b)
{
    if (!b.plane_owner.empty()) return 0;
    const Geom &g = b.g;
    int mine[2] = {g.xOffset, g.X}, *dsend = nullptr, *drecv = nullptr;
    CK(cudaMalloc(&dsend, sizeof(mine))); CK(cudaMalloc(&drecv, sizeof(mine) * g_nccl.nranks));
    CK(cudaMemcpy(dsend, mine, sizeof(mine), cudaMemcpyHostToDevice));
    NCK(g_nccl.AllGather(dsend, drecv, sizeof(mine), kNcclChar, g_nccl.comm, b.stream));
    CK(cudaStreamSynchronize(b.stream));
    std::vector<int> all(2 * g_nccl.nranks);
    CK(cudaMemcpy(all.data(), drecv, sizeof(mine) * g_nccl.nranks, cudaMemcpyDeviceToHost));
    cudaFree(dsend); cudaFree(drecv);
    b.plane_owner.assign(g.XG, -1);
    for (int r = 0; r < g_nccl.nranks; r++)
        for (int x = all[2 * r]; x < all[2 * r] + all[2 * r + 1] && x < g.XG; x++) b.plane_owner[x] = r;
    for (int x = 0; x < g.XG; x++) if (b.plane_owner[x] < 0) { b.plane_owner.clear(); return fail(FSILBM_ERR_COMM, "x-plane %d belongs to no rank's slab", x); }
    return 0;
}

// A stretch of box planes dx in [dx0, dx1) owned by one rank (see block_comm.ibm_box_participants for the rule in Python)
struct Run { int rank, dx0, dx1; };

// The participants of a shared box (a body across a slab interface) send one another the box planes they own: ncclSend/ncclRecv
// between the two or three ranks concerned, one group, no collective.  Every rank walks (box, component, run) in the same
// order, so the messages of a pair of ranks meet in posting order.
int ibm_exchange_shared_boxes(const IbmBoxes &bx, const std::vector<int> &kept, const std::vector<char> &shared,
                              const std::vector<std::vector<Run>> &runs, int me, cudaStream_t s)
{
    NCK(g_nccl.GroupStart());
    for (int i = 0; i < bx.n; i++) {
        const int k = kept[i];
        if (!shared[k]) continue;
        const size_t slab = (size_t)bx.ext[i][1] * bx.ext[i][2];
        std::vector<int> parts;
        for (const Run &r : runs[k]) if (std::find(parts.begin(), parts.end(), r.rank) == parts.end()) parts.push_back(r.rank);
        for (int c = 0; c < 3; c++)
            for (const Run &r : runs[k]) {
                double *ptr = bx.u + (size_t)c * bx.ncell + bx.off[i] + (size_t)r.dx0 * slab;
                const size_t cnt = (size_t)(r.dx1 - r.dx0) * slab;
                if (r.rank == me) { for (int p : parts) if (p != me) NCK(g_nccl.Send(ptr, cnt, kNcclFloat64, p, g_nccl.comm, s)); }
                else NCK(g_nccl.Recv(ptr, cnt, kNcclFloat64, r.rank, g_nccl.comm, s));
            }
    }
    NCK(g_nccl.GroupEnd());
    return 0;
}

}  // namespace

extern "C" {

int fsilbm_ibm_interaction_force_begin(fsilbm_handle h, int nbody, const int *nelmts, const double *const *Exyz, const double *const *Evel,
                                       const double *const *Ea, const int *restencil, double dt, int ntolLBM,
                                       double dtolLBM, const int rootBC[6])
{
    Block *bp = get(h);
    if (!bp) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    Block &b = *bp;
    if (nbody < 0 || (nbody > 0 && (!nelmts || !Exyz || !Evel || !Ea || !restencil || !rootBC)))
        return fail(FSILBM_ERR_ARG, "null argument");
    if (b.ibm_pending.active)
        return fail(FSILBM_ERR_ARG, "fsilbm_ibm_interaction_force_begin: the previous call on block %d has not been collected (fsilbm_ibm_interaction_force_wait)", h);
    const Geom &g = b.g;
    cudaStream_t s = b.stream, s2 = b.ibm_stream;
    const bool multi = g_nccl.nranks > 1 && g_nccl.comm && g.X != g.XG;   // this block is split into x-slabs
    const bool ordered = g_ibm_ordered != 0;
    // Slab runs, default form ("local"): a body is iterated only by the ranks whose planes its stencil box touches.
    // Shared boxes (a body across a slab interface) exchange their owned box planes pairwise and are then iterated
    // redundantly -- bit-identically -- by their participants; the loop control is all-reduced (two numbers per iteration).
    if (multi && !ordered)
        return fail(FSILBM_ERR_ARG, "slab runs iterate a body redundantly on the ranks that share it, which needs the ordered (bit-reproducible) IBM form: "
                                    "option ibm_ordered = 0 is a single-GPU comparison arm");
    const bool local = multi;
    const bool lists_replicated = g_ibm_force_exchange != 0;   // every rank passes the same bodies and wants every force back
    b.ibm_pending = Block::IbmPending();
    if (nbody == 0 && !(local && !lists_replicated)) { b.ibm_active = false; return 0; }   // Solidbody.f90:891
    static const bool want_prof = getenv("FSILBM_IBM_PROFILE") != nullptr;
    const double tp0 = want_prof ? wall_seconds() : 0.0;
    const int me = g_nccl.rank;

    if (int rc = ibm_body_storage(b, nbody, nelmts)) return rc;

    // -- host: the box around the stencils of each body, merged where they overlap
    const int Ns[3] = {g.XG, g.Y, g.Z};
    std::vector<char> re(nbody, 0);
    for (int ib = 0; ib < nbody; ib++) {
        BodyDev &bd = b.bodies[ib];
        re[ib] = (restencil[ib] != 0 || !bd.have_box) ? 1 : 0;
        if (!re[ib]) continue;
        b.csr_valid = false;
        ibm_body_box(g, bd, Exyz[ib], rootBC);
    }
    std::vector<MBox> mb = ibm_merge_boxes(b, nbody);
    // -- slab runs: which ranks own planes of each box
    std::vector<std::vector<Run>> runs(mb.size());
    std::vector<char> keep(mb.size(), 1), shared(mb.size(), 0);
    if (local) {
        if (int rc = ibm_plane_owners(b)) return rc;
        for (size_t k = 0; k < mb.size(); k++) {
            const Interval &ix = mb[k].hb.ax[0];
            bool mine_in = false;
            for (int dx = 0; dx < ix.l; dx++) {
                const int r = b.plane_owner[(ix.s + dx) % g.XG];
                if (runs[k].empty() || runs[k].back().rank != r) runs[k].push_back(Run{r, dx, dx + 1}); else runs[k].back().dx1 = dx + 1;
                mine_in = mine_in || r == me;
            }
            keep[k] = mine_in ? 1 : 0;
            for (const Run &r : runs[k]) if (r.rank != runs[k][0].rank) shared[k] = 1;
        }
    }
    std::vector<int> kept;
    for (size_t k = 0; k < mb.size(); k++) if (keep[k]) kept.push_back((int)k);
    if (local && (int)kept.size() > MAX_BOXES) return fail(FSILBM_ERR_ARG, "more than %d separate stencil boxes touch this slab", MAX_BOXES);
    // More separate stencil boxes than the box table holds: the two whose covering box adds the fewest cells are joined (their
    // bodies then share one box; results do not change, a box is only storage) until the table fits.  Said once on stderr.
    auto cover = [&](const HostBox &x0, const HostBox &y0, HostBox &out) {
        long long vol = 1;
        for (int a = 0; a < 3; a++) {
            const Interval &x = x0.ax[a], &y = y0.ax[a];
            Interval r;
            const int d1 = imod(y.s - x.s, Ns[a]) + y.l, d2 = imod(x.s - y.s, Ns[a]) + x.l;
            if (std::max(d1, x.l) <= std::max(d2, y.l)) { r.s = x.s; r.l = std::max(d1, x.l); } else { r.s = y.s; r.l = std::max(d2, y.l); }
            if (r.l >= Ns[a]) { r.s = 0; r.l = Ns[a]; }
            out.ax[a] = r;
            vol *= r.l;
        }
        return vol;
    };
    auto volume = [&](const HostBox &x) { return (long long)x.ax[0].l * x.ax[1].l * x.ax[2].l; };
    if (!local && (int)kept.size() > MAX_BOXES) {
        static bool said = false;
        if (!said) {
            said = true;
            fprintf(stderr, "fsilbm: %d separate stencil boxes in one block, the box table holds %d: nearest boxes are joined\n", (int)kept.size(), MAX_BOXES);
        }
    }
    while (!local && (int)kept.size() > MAX_BOXES) {
        size_t bi = 0, bj = 1;
        long long best = -1;
        HostBox tmp;
        for (size_t i = 0; i < kept.size(); i++)
            for (size_t j = i + 1; j < kept.size(); j++) {
                const long long extra = cover(mb[kept[i]].hb, mb[kept[j]].hb, tmp) - volume(mb[kept[i]].hb) - volume(mb[kept[j]].hb);
                if (best < 0 || extra < best) { best = extra; bi = i; bj = j; }
            }
        MBox &a0 = mb[kept[bi]], &b0 = mb[kept[bj]];
        cover(a0.hb, b0.hb, tmp);
        a0.hb = tmp;
        a0.members.insert(a0.members.end(), b0.members.begin(), b0.members.end());
        std::sort(a0.members.begin(), a0.members.end());
        kept.erase(kept.begin() + bj);
        // the covering box may now reach a third box: no cell may lie in two boxes (the kernels look a cell up in the first box
        // that holds it), so whatever it overlaps joins it as well
        for (bool changed = true; changed;) {
            changed = false;
            for (size_t j = 0; j < kept.size() && !changed; j++) {
                if (j == bi) continue;
                bool ov = true;
                for (int a = 0; a < 3; a++) ov = ov && overlap(mb[kept[bi]].hb.ax[a], mb[kept[j]].hb.ax[a], Ns[a]);
                if (!ov) continue;
                const size_t into = std::min(bi, j), from = std::max(bi, j);   // the earlier box keeps its place (box order = first member)
                MBox &x = mb[kept[into]], &y = mb[kept[from]];
                cover(x.hb, y.hb, tmp);
                x.hb = tmp;
                x.members.insert(x.members.end(), y.members.begin(), y.members.end());
                std::sort(x.members.begin(), x.members.end());
                kept.erase(kept.begin() + from);
                bi = into;
                changed = true;
            }
        }
    }
    // active bodies (caller order = device order, so the cell lists keep the reference's body order), their box group and
    // whether this rank leads them (owns the first plane of their box: it reports their residual and their forces)
    std::vector<int> group_of(nbody, -1), lead_of(nbody, 0);
    for (size_t kk = 0; kk < kept.size(); kk++)
        for (int ib : mb[kept[kk]].members) {
            group_of[ib] = (int)kk;
            lead_of[ib] = (!local || runs[kept[kk]][0].rank == me) ? 1 : 0;
        }
    std::vector<int> act;
    for (int ib = 0; ib < nbody; ib++) {
        if (group_of[ib] >= 0) act.push_back(ib);
        b.bodies[ib].status = group_of[ib] < 0 ? 0 : (lead_of[ib] ? 2 : 1);
    }
    const int nact = (int)act.size();
    // a body that was never stencilled on this rank (its box did not touch this slab until a merge with a moving neighbour brought
    // it here) must be, whatever the caller's restencil flag says: the flag describes the body, not this rank's copy of it
    for (int ib : act)
        if (!re[ib] && !b.bodies[ib].stencil_valid) { re[ib] = 1; b.csr_valid = false; }
    for (int ib = 0; ib < nbody; ib++)
        if (group_of[ib] < 0 && re[ib]) b.bodies[ib].stencil_valid = false;   // moved while inactive here: this rank's copy is stale
    {   // the set of active bodies decides the device body table and the cell lists
        if (b.active_prev != act) { b.csr_valid = false; b.active_prev = act; }
    }
    IbmBoxes &bx = b.boxes;
    bx.n = (int)kept.size();
    long long total = 0;
    for (int i = 0; i < bx.n; i++) {
        const HostBox &hb = mb[kept[i]].hb;
        for (int a = 0; a < 3; a++) { bx.lo[i][a] = hb.ax[a].s; bx.ext[i][a] = hb.ax[a].l; }
        bx.off[i] = total;
        total += (long long)bx.ext[i][0] * bx.ext[i][1] * bx.ext[i][2];
    }
    bx.ncell = total;
    if (total > b.box_capacity) {
        CK(cudaStreamSynchronize(s)); CK(cudaStreamSynchronize(s2));
        cudaFree(bx.u); cudaFree(bx.force);
        const long long cap = total + total / 4 + 1024;
        CK(cudaMalloc(&bx.u, sizeof(double) * 3 * cap));
        CK(cudaMalloc(&bx.force, sizeof(double) * 3 * cap));
        b.box_capacity = cap;
    }
    const double tp1 = want_prof ? wall_seconds() : 0.0;

    // -- side stream: marker upload, UpdateElmtInterp_ and the cell lists need nothing of the fluid state, so they run
    //    beside whatever the compute stream is still doing (the collide-stream kernel of the previous step)
    if (b.marker_total) CK(cudaMemsetAsync(b.force_dev, 0, sizeof(double) * 3 * b.marker_total, s2));   // Solidbody.f90:889
    for (int k = 0; k < nact; k++) {
        const int ib = act[k];
        BodyDev &bd = b.bodies[ib];
        const size_t n = bd.n;
        double *pin = b.mk_pin + b.mk_off[ib];
        memcpy(pin, Exyz[ib], sizeof(double) * 3 * n);
        memcpy(pin + 3 * n, Evel[ib], sizeof(double) * 3 * n);
        memcpy(pin + 6 * n, Ea[ib], sizeof(double) * n);
        CK(cudaMemcpyAsync(bd.Exyz, pin, sizeof(double) * 7 * n, cudaMemcpyHostToDevice, s2));
        if (re[ib]) {
            CK(cudaMemcpyAsync(bd.ExyzStencil, bd.Exyz, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, s2));
            bd.stencil_valid = true;
        }
    }
    IbmCtl ctl0;
    ctl0.iter = 0; ctl0.done = (ntolLBM <= 0) ? 1 : 0; ctl0.err = 0; ctl0.dmax = 1e10; ctl0.tol_acc = 0.0;   // :893-894
    CK(cudaMemcpyAsync(b.ctl, &ctl0, sizeof(IbmCtl), cudaMemcpyHostToDevice, s2));
    if (b.bodies_dev_cap < nact) {
        CK(cudaStreamSynchronize(s)); CK(cudaStreamSynchronize(s2));
        cudaFree(b.bodies_dev); cudaFree(b.lead_dev);
        CK(cudaMalloc(&b.bodies_dev, sizeof(IbmBody) * nact));
        CK(cudaMalloc(&b.lead_dev, sizeof(int) * nact));
        b.bodies_dev_cap = nact;
    }
    std::vector<IbmBody> views(nact);
    std::vector<int> lead(nact);
    for (int k = 0; k < nact; k++) { views[k] = b.bodies[act[k]].view(); lead[k] = lead_of[act[k]]; }
    if (nact) {
        CK(cudaMemcpyAsync(b.bodies_dev, views.data(), sizeof(IbmBody) * nact, cudaMemcpyHostToDevice, s2));
        CK(cudaMemcpyAsync(b.lead_dev, lead.data(), sizeof(int) * nact, cudaMemcpyHostToDevice, s2));
    }

    double hF[3];
    half_force(b, hF);
    const double invh3_pen = 0.5 * dt * ((1.0 / g.dh) * (1.0 / g.dh) * (1.0 / g.dh)) / b.flow.denIn;   // :996
    const double invh3 = (1.0 / g.dh) * (1.0 / g.dh) * (1.0 / g.dh);                                     // :936
    // slab runs: with every rank's mailbox mapped (peer memory) the loop control is exchanged from inside the single cooperative
    // kernel; otherwise one kernel per phase with an ncclAllReduce of the two numbers per iteration
    const bool mailbox = local && b.halo.mailbox && g_ibm_single_launch;
    IbmCtlExchange xc{};
    if (mailbox) {
        xc.nranks = g_nccl.nranks; xc.rank = me;
        for (int r = 0; r < g_nccl.nranks; r++) xc.mailbox[r] = b.halo.peer[r] + kHaloFlagBytes;
        xc.seq_base = b.halo.ctl_seq + 1;
        xc.timeout_ns = (unsigned long long)g_halo_timeout_s * 1000000000ull;
        xc.cnt_local = 0.0;
        for (int ib : act) if (lead_of[ib]) xc.cnt_local = xc.cnt_local + (double)nelmts[ib];
    }
    if (mailbox && nact > MAX_IBM_PHASE_BODIES)   // the other ranks are in the mailbox protocol: no silent switch to another path
        return fail(FSILBM_ERR_ARG, "more than %d bodies touch this slab (ibm_single_launch = 0 lifts the limit)", MAX_IBM_PHASE_BODIES);
    bool single = (!multi || mailbox) && g_ibm_single_launch && nact <= MAX_IBM_PHASE_BODIES && nact > 0;
    Geom gsten = g;
    if (local) { gsten.xOffset = 0; gsten.X = g.XG; }   // stencil_marker: every stencil plane counts as owned
    if (ordered && nact) {
        // stencils first (the cell lists are built from them), then the lists; both survive while no body restencils
        long long entries = 0;
        for (int k = 0; k < nact; k++) entries += (long long)views[k].n * 64;
        if (bx.ncell + 1 > b.csr_cell_cap) {
            CK(cudaStreamSynchronize(s)); CK(cudaStreamSynchronize(s2));
            cudaFree(b.csr.count); cudaFree(b.csr.off); cudaFree(b.csr_scan_tmp);
            const long long cap = bx.ncell + bx.ncell / 4 + 1024;
            CK(cudaMalloc(&b.csr.count, sizeof(int) * (size_t)cap));
            CK(cudaMalloc(&b.csr.off, sizeof(int) * (size_t)cap));
            b.csr_scan_bytes = ibm_csr_scan_bytes(cap);
            CK(cudaMalloc(&b.csr_scan_tmp, b.csr_scan_bytes ? b.csr_scan_bytes : 1));
            b.csr_cell_cap = cap;
            b.csr_valid = false;
        }
        if (entries > b.csr_entry_cap) {
            CK(cudaStreamSynchronize(s)); CK(cudaStreamSynchronize(s2));
            cudaFree(b.csr.entry);
            CK(cudaMalloc(&b.csr.entry, sizeof(unsigned long long) * (size_t)entries));
            b.csr_entry_cap = entries;
            b.csr_valid = false;
        }
        if (!b.tol_partial) CK(cudaMalloc(&b.tol_partial, sizeof(double) * 2 * (size_t)ibm_loop_max_blocks()));
        if (!b.csr_valid) {
            int max_n = 0;
            for (int k = 0; k < nact; k++) max_n = std::max(max_n, views[k].n);
            launch_ibm_stencil_all(gsten, b.bodies_dev, nact, max_n, bx, rootBC, b.ctl, s2);
            if (launch_ibm_csr_build(b.bodies_dev, nact, max_n, bx, b.csr, b.csr_scan_tmp, b.csr_scan_bytes, s2)) return fail(FSILBM_ERR_CUDA, "IBM cell-list build failed");
            b.csr_valid = true;
        }
    }
    // Early IBM: when every box lies inside the planes the last collide_stream finished first (and clear of the faces), the rest
    // of this call runs on its own stream behind ev_early, beside the remainder of that update, instead of behind all of it.
    bool use_early = b.early_ok && g_ibm_early && bx.n > 0 && ordered;
    for (int i = 0; i < bx.n && use_early; i++) {
        bool inside = false;
        int lo = bx.lo[i][0], hi = lo + bx.ext[i][0];
        const bool wraps = hi > g.XG;
        if (multi) { lo = std::max(lo, g.xOffset); hi = std::min(hi, g.xOffset + g.X); }   // this rank's share of a box across an interface
        for (int k = 0; k < b.early_n; k++) inside = inside || (lo >= b.early_x0[k] && hi <= b.early_x1[k]);
        for (int k = 1; k < 3; k++)
            if (b.periodic[k] != 1 && (bx.lo[i][k] < 3 || bx.lo[i][k] + bx.ext[i][k] > Ns[k] - 3)) inside = false;
        if (b.periodic[0] != 1 && (bx.lo[i][0] < 3 || bx.lo[i][0] + bx.ext[i][0] > Ns[0] - 3)) inside = false;
        if (wraps) inside = false;
        use_early = inside;
    }
    b.early_ok = false;   // one use per update
    if (use_early) {
        g_ibm_early_calls++;
        s = b.ibm_main_stream;
        CK(cudaStreamWaitEvent(s, b.ev_early, 0));
    } else if (local && bx.n == 0 && b.halo.enabled && g_ibm_early) {
        // no body in this slab: what is left (loop-control exchange, force exchange) reads nothing of the fluid state, so it need
        // not queue behind the update either -- the ranks that do iterate bodies early are then not held up by this one
        s = b.ibm_main_stream;
    }
    TRACE(s2, 1, "ibm_upload_stencils");
    CK(cudaEventRecord(b.ev_ibm, s2));
    CK(cudaStreamWaitEvent(s, b.ev_ibm, 0));

    // -- compute stream: everything that reads the populations
    bool macro_done = false;
    if (local) {
        // every kept box: this rank's planes from its populations, zero elsewhere; then the participants of a shared box send
        // one another the planes they own (ncclSend/ncclRecv between the two or three ranks concerned, no collective)
        bool any_shared = false;
        for (int i = 0; i < bx.n; i++) any_shared = any_shared || shared[kept[i]];
        if (any_shared || !single) { launch_ibm_macro_box(g, b.f[b.cur], hF, bx, s); macro_done = true; }
        if (any_shared)
            if (int rc = ibm_exchange_shared_boxes(bx, kept, shared, runs, me, s)) return rc;
    }
    if (single) {
        // one cooperative launch for UpdateElmtInterp_, the box macro, the whole penalty iteration and the force spreading
        IbmLoopParams lp{};
        lp.g = g; lp.bodies = b.bodies_dev; lp.nbody = nact; lp.boxes = bx;
        for (int k = 0; k < 6; k++) lp.rootBC[k] = rootBC[k];
        lp.ctl = b.ctl; lp.fA = b.f[b.cur];
        for (int k = 0; k < 3; k++) lp.hF[k] = hF[k];
        lp.ntol = ntolLBM; lp.dtol = dtolLBM; lp.Uref = b.flow.Uref;
        lp.dsum = 0.0;
        for (int k = 0; k < nact; k++) lp.dsum = lp.dsum + (double)views[k].n;   // :902
        lp.invh3_pen = invh3_pen; lp.invh3 = invh3; lp.barrier = b.ibm_barrier;
        if (want_prof && !b.ibm_prof) { CK(cudaMalloc(&b.ibm_prof, 64 * sizeof(unsigned long long))); CK(cudaMemset(b.ibm_prof, 0, 64 * sizeof(unsigned long long))); }
        lp.prof = b.ibm_prof;
        lp.ordered = ordered ? 1 : 0; lp.do_stencil = ordered ? 0 : 1; lp.do_macro = macro_done ? 0 : 1; lp.csr = b.csr; lp.tol_partial = b.tol_partial;
        for (int k = 0; k < nact; k++) lp.lead[k] = (unsigned char)lead[k];
        lp.xc = xc;
        // phases: the k-th body (in body order) of every box group
        std::vector<int> rank_in_group(nact, 0), seen(bx.n > 0 ? bx.n : 1, 0);
        int nphase = 0;
        for (int k = 0; k < nact; k++) { rank_in_group[k] = seen[group_of[act[k]]]++; nphase = std::max(nphase, rank_in_group[k] + 1); }
        lp.nphase = nphase;
        int pos = 0, max_markers = 0;
        for (int ph = 0; ph < nphase; ph++) {
            lp.phase_start[ph] = pos;
            int markers = 0;
            for (int k = 0; k < nact; k++) if (rank_in_group[k] == ph) { lp.phase_body[pos++] = k; markers += views[k].n; }
            max_markers = std::max(max_markers, markers);
        }
        lp.phase_start[nphase] = pos;
        for (int k = 0; k < nact; k++) lp.phase_of_body[k] = rank_in_group[k];
        // Grid of the cooperative kernel when it shares the SMs with a running update: every one of its blocks takes a CTA slot
        // from that update for as long as the iteration lasts, and the iteration is latency-bound -- its length grows far slower
        // than 1/blocks -- so the smallest grid that still finishes in time costs the update least.  Bodies that do not move
        // (no re-stencil) leave the host nothing to do between two calls: the iteration may take most of the update, one block
        // per ~110 markers (74 blocks for the 8192-marker plate: 1.683 -> 1.636 ms per step on plate512).  Moving/flexible bodies
        // have the host's structural solve waiting for the forces: they keep ibm_early_blocks_per_sm blocks per SM.
        int early_total = 0, early_bps = g_ibm_early_blocks;
        if (use_early && early_bps == 0) {
            // One block per SM costs the update beside it least, but the iteration must not outlast that update: what is left of it
            // (the planes outside the ranges it finished first) against the iteration's length at one block per SM (measured:
            // ntol x markers x 7.5 ns -- 0.3 ms for 8 192 markers x 5, 1.25 ms for 32 768 x 5).  A large body in a small block (a plate
            // meshed at a refined son's spacing) gets up to four blocks per SM; a root-block body keeps one.
            long long total_markers = 0;
            for (int k = 0; k < nact; k++) total_markers += views[k].n;
            double planes_rest = (double)g.X;
            for (int k = 0; k < b.early_n; k++) {
                const int lo = std::max(b.early_x0[k], g.xOffset), hi = std::min(b.early_x1[k], g.xOffset + g.X);
                if (hi > lo) planes_rest -= (double)(hi - lo);
            }
            // ... plus the updates of OTHER blocks issued since (a son's first call of a root step is issued behind its father's
            // update, which is many times longer than the son's own: the iteration then runs beside that, and one block per SM --
            // which costs the update beside it least -- is early enough)
            const double t_rest_us = std::max(planes_rest, 1.0) * (double)g.plane * 304.0 / 6.2e6 + (g_update_clock_us - b.update_clock_us);
            const double t_ibm_us = (double)std::max(ntolLBM, 1) * (double)total_markers * 0.0075;
            early_bps = (int)std::ceil(t_ibm_us / t_rest_us);
            early_bps = std::max(1, std::min(4, early_bps));
        }
        if (use_early) {
            early_total = g_ibm_early_total;
            bool any_re = false;
            for (int ib : act) any_re = any_re || re[ib];
            if (early_total == 0 && !any_re) {
                static int sms = 0;
                if (!sms) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g_device);
                early_total = std::max(sms / 4, std::min(2 * sms, (max_markers + 109) / 110));
            }
        }
        if (launch_ibm_loop(lp, max_markers, use_early ? early_bps : 0, early_total, s)) {
            cudaGetLastError();
            if (mailbox) return fail(FSILBM_ERR_CUDA, "cooperative launch of the IBM iteration failed");   // the other ranks are in the mailbox protocol
            single = false;   // no cooperative launch: take the phase-by-phase path
        }
    }
    if (mailbox && nact == 0) {
        launch_ibm_ctl_only(xc, ntolLBM, dtolLBM, b.flow.Uref, b.ctl, s);   // no body here: report zeros, follow the others' decision
        single = true;
    }
    if (!single) {
        // -- UpdateElmtInterp_ (:883-888); the box-relative offsets are rebuilt every call because the boxes move with the
        //    bodies (the ordered mode did it above, together with its cell lists)
        if (!ordered) for (int k = 0; k < nact; k++) launch_ibm_stencil(g, views[k], bx, rootBC, b.ctl, s);
        // -- calculate_macro_quantities + ResetVolumeForce restricted to the boxes (LBMBlockComm.f90:285-286)
        if (!macro_done) launch_ibm_macro_box(g, b.f[b.cur], hF, bx, s);
        if (local && !b.tol2) CK(cudaMalloc(&b.tol2, 2 * sizeof(double)));
        // -- penalty iteration (:895-906); launches beyond convergence return at once on the device flag
        for (int it = 0; it < ntolLBM; it++) {
            for (int k = 0; k < nact; k++) {
                auto gather = ordered ? launch_ibm_gather_ordered : launch_ibm_gather;
                BodyDev &bd = b.bodies[act[k]];
                gather(views[k], bx, bd.partialU, b.ctl, 1, invh3_pen, s);
                if (ordered) launch_ibm_scatter_ordered(b.bodies_dev, k, bx, b.csr, b.ctl, s);
                else launch_ibm_scatter(views[k], bx, b.ctl, s);
            }
            if (local) {   // residual and marker count of the bodies this rank leads -> totals over all ranks -> the same decision everywhere
                launch_ibm_tol_sum(b.bodies_dev, b.lead_dev, nact, b.ctl, b.tol2, s);
                NCK(g_nccl.AllReduce(b.tol2, b.tol2, 2, kNcclFloat64, kNcclSum, g_nccl.comm, s));
                launch_ibm_decide(b.tol2, b.flow.Uref, ntolLBM, dtolLBM, b.ctl, s);
            } else {
                launch_ibm_check(b.bodies_dev, nact, b.flow.Uref, ntolLBM, dtolLBM, b.ctl, s);
            }
        }
        // -- FluidVolumeForce_, Eulerian half (:968-976)
        if (ordered) { if (nact) launch_ibm_spread_ordered(b.bodies_dev, bx, b.csr, invh3, s); }
        else for (int k = 0; k < nact; k++) launch_ibm_spread(views[k], bx, invh3, s);
    }
    CK(cudaGetLastError());
    TRACE(s, use_early ? 2 : 0, "ibm_iteration");
    if (local && lists_replicated && b.marker_total) {
        // every rank wants every body's forces: the leader's values plus zeros from everybody else (exact)
        for (int k = 0; k < nact; k++)
            if (!lead[k]) CK(cudaMemsetAsync(b.bodies[act[k]].Eforce, 0, sizeof(double) * 3 * (size_t)views[k].n, s));
        NCK(g_nccl.AllReduce(b.force_dev, b.force_dev, 3 * b.marker_total, kNcclFloat64, kNcclSum, g_nccl.comm, s));
    }

    // -- results towards the host: control block and marker forces into pinned memory, asynchronously; the event marks the end of
    //    everything this call put on the device.  Nothing here waits: fsilbm_block_collide_stream may be enqueued right away (it
    //    waits for the event on the device), the host collects the results with fsilbm_ibm_interaction_force_wait.
    CK(cudaMemcpyAsync(b.ctl_pin, b.ctl, sizeof(IbmCtl), cudaMemcpyDeviceToHost, s));
    if (b.marker_total) CK(cudaMemcpyAsync(b.force_pin, b.force_dev, sizeof(double) * 3 * b.marker_total, cudaMemcpyDeviceToHost, s));
    TRACE(s, use_early ? 2 : 0, "ibm_forces_d2h");
    CK(cudaEventRecord(b.ev_ibm_done, s));
    b.ibm_pending.active = true; b.ibm_pending.stream = s; b.ibm_pending.nbody = nbody; b.ibm_pending.nact = nact; b.ibm_pending.mailbox = mailbox;
    b.ibm_pending.t0 = tp0; b.ibm_pending.t1 = tp1; b.ibm_pending.t2 = want_prof ? wall_seconds() : 0.0;
    b.ibm_active = bx.n > 0;
    return 0;
}

int fsilbm_ibm_interaction_force_wait(fsilbm_handle h, int nbody, double *const *Eforce, int *iterLBM_out)
{
    Block *bp = get(h);
    if (!bp) return fail(FSILBM_ERR_ARG, "bad handle %d", h);
    Block &b = *bp;
    if (iterLBM_out) *iterLBM_out = 0;
    if (!b.ibm_pending.active) return 0;   // nothing was enqueued (no body: Solidbody.f90:891)
    if (nbody != b.ibm_pending.nbody || (nbody > 0 && !Eforce)) return fail(FSILBM_ERR_ARG, "fsilbm_ibm_interaction_force_wait: %d bodies were passed to _begin", b.ibm_pending.nbody);
    static const bool want_prof = getenv("FSILBM_IBM_PROFILE") != nullptr;
    const int me = g_nccl.rank;
    CK(cudaEventSynchronize(b.ev_ibm_done));
    b.ibm_pending.active = false;
    const IbmCtl ctl1 = *b.ctl_pin;
    for (int ib = 0; ib < nbody; ib++) memcpy(Eforce[ib], b.force_pin + b.f_off[ib], sizeof(double) * 3 * (size_t)b.bodies[ib].n);
    if (want_prof) {
        const double tp3 = wall_seconds();
        b.ibm_host_t[0] += b.ibm_pending.t1 - b.ibm_pending.t0; b.ibm_host_t[1] += b.ibm_pending.t2 - b.ibm_pending.t1; b.ibm_host_t[2] += tp3 - b.ibm_pending.t2;
        if ((b.ibm_prof_calls % 100) == 20) {
            fprintf(stderr, "[ibm host, us per call] boxes %.1f  enqueue %.1f  begin->collected %.1f  (%d of %d bodies active)\n", b.ibm_host_t[0] * 1e4, b.ibm_host_t[1] * 1e4,
                    b.ibm_host_t[2] * 1e4, b.ibm_pending.nact, nbody);
            b.ibm_host_t[0] = b.ibm_host_t[1] = b.ibm_host_t[2] = 0.0;
        }
        if (!b.ibm_prof) b.ibm_prof_calls++;
    }
    if (b.ibm_prof && (b.ibm_prof_calls++ % 100) == 20) {
        unsigned long long hp[64];
        cudaMemcpy(hp, b.ibm_prof, sizeof(hp), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[ibm_loop phases, us]");
        for (unsigned long long k = 1; k < hp[0] && k < 63; k++) fprintf(stderr, " %.1f", (double)(hp[1 + k] - hp[k]) * 1e-3);
        fprintf(stderr, "  (ncell %lld)\n", (long long)b.boxes.ncell);
    }
    if (b.ibm_pending.mailbox) b.halo.ctl_seq += (unsigned long long)ctl1.iter;   // the same count on every rank
    if (ctl1.err) b.csr_valid = false;
    if (ctl1.err & 8) return fail(FSILBM_ERR_COMM, "IBM loop control: a rank did not report within %d s (rank %d)", g_halo_timeout_s, me);
    if (ctl1.err & 1) return fail(FSILBM_ERR_STENCIL, "index out of xmin/xmax bound (Solidbody.f90:850,861)");
    if (ctl1.err & 4) return fail(FSILBM_ERR_STENCIL, "internal: marker stencil outside its IBM box");
    if (ctl1.err & 2) return fail(FSILBM_ERR_NAN, "Nan found in PenaltyForce (Solidbody.f90:1029)");
    if (iterLBM_out) *iterLBM_out = ctl1.iter;
    return 0;
}

// calculate_interaction_force as one blocking call (Solidbody.f90:869-918): _begin + _wait
int fsilbm_ibm_interaction_force(fsilbm_handle h, int nbody, const int *nelmts, const double *const *Exyz, const double *const *Evel,
                                 const double *const *Ea, double *const *Eforce, const int *restencil, double dt, int ntolLBM,
                                 double dtolLBM, const int rootBC[6], int *iterLBM_out)
{
    if (iterLBM_out) *iterLBM_out = 0;
    if (nbody > 0 && !Eforce) return fail(FSILBM_ERR_ARG, "null argument");
    if (int rc = fsilbm_ibm_interaction_force_begin(h, nbody, nelmts, Exyz, Evel, Ea, restencil, dt, ntolLBM, dtolLBM, rootBC)) return rc;
    return fsilbm_ibm_interaction_force_wait(h, nbody, Eforce, iterLBM_out);
}

int fsilbm_ibm_body_status(fsilbm_handle h, int nbody, int *status)
{
    Block *b = get(h);
    if (!b || !status || nbody != (int)b->bodies.size()) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    for (int ib = 0; ib < nbody; ib++) status[ib] = b->bodies[ib].status;
    return 0;
}

int fsilbm_ibm_download_stencil(fsilbm_handle h, int body, short *Ei, float *Ew)
{
    Block *b = get(h);
    if (!b || body < 0 || body >= (int)b->bodies.size()) return fail(FSILBM_ERR_ARG, "bad handle/body");
    CK(cudaStreamSynchronize(b->stream));
    const size_t n = b->bodies[body].n;
    if (Ei) CK(cudaMemcpy(Ei, b->bodies[body].Ei, sizeof(short) * 12 * n, cudaMemcpyDeviceToHost));
    if (Ew) CK(cudaMemcpy(Ew, b->bodies[body].Ew, sizeof(float) * 12 * n, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- grid refinement: CommPair and the father<->son transfers (LBMBlockComm.f90) -----------------------
namespace {
Pair *get_pair(int h)
{
    if (h < 0 || h >= (int)g_pairs.size() || !g_pairs[h]) return nullptr;
    return g_pairs[h].get();
}
inline void pair_axes(int j, int &axis, int &bAx, int &aAx)
{
    axis = j / 2;
    if (axis == 0) { bAx = 2; aAx = 1; } else if (axis == 1) { bAx = 2; aAx = 0; } else { bAx = 1; aAx = 0; }
}
// everything one son face needs; `force_of` selects whose volumeForce/dh enter fIn_GridTransform (0 father, 1 son)
PairFaceParams pair_face(const Pair &p, Block &F, Block &S, int j, int force_of)
{
    PairFaceParams q{};
    int axis, bAx, aAx;
    pair_axes(j, axis, bAx, aAx);
    q.gF = F.g; q.gS = S.g;
    q.fS = S.f[S.cur];
    q.fv.nseg = 1;
    q.fv.f[0] = F.f[F.cur]; q.fv.x0[0] = F.g.xOffset; q.fv.X[0] = F.g.X; q.fv.pstride[0] = F.g.pstride;
    for (int sd = 0; sd < 2; sd++) {
        if (!p.cross[sd]) continue;
        // the neighbours flip their buffers in step with this rank (one flip per collide-stream of the block on every rank)
        const int k = q.fv.nseg++;
        q.fv.f[k] = F.halo.peer_f[sd][F.cur];
        q.fv.X[k] = F.halo.peer_X[sd];
        q.fv.x0[k] = sd == 0 ? F.g.xOffset - F.halo.peer_X[0] : F.g.xOffset + F.g.X;
        q.fv.pstride[k] = (size_t)(F.halo.peer_X[sd] + 2) * F.g.plane;
    }
    q.axis = axis; q.scheme = p.scheme;
    q.bF = p.dimF[bAx]; q.aF = p.dimF[aAx]; q.bS = p.dimS[bAx]; q.aS = p.dimS[aAx];
    q.fplane = p.f[j] - 1; q.fb0 = p.f[2 * bAx] - 1; q.fa0 = p.f[2 * aAx] - 1;
    q.splane = p.s[j] - 1;
    q.siplane = p.si[j] - 1; q.sib0 = p.si[2 * bAx] - 1; q.sia0 = p.si[2 * aAx] - 1;
    q.fiplane = p.fi[j] - 1; q.fib0 = p.fi[2 * bAx] - 1; q.fia0 = p.fi[2 * aAx] - 1;
    q.nb = (p.si[2 * bAx + 1] - p.si[2 * bAx]) / 2 + 1;
    q.na = (p.si[2 * aAx + 1] - p.si[2 * aAx]) / 2 + 1;
    for (int t = 0; t < 2; t++) { q.buf[t] = p.buf[j][t]; q.tbuf[t] = p.tbuf[j][t]; }
    q.tauF = F.tau; q.tauS = S.tau; q.tauF_all = F.tau_all; q.tauS_all = S.tau_all;
    half_force(force_of ? S : F, q.hF);
    return q;
}
}  // namespace

int fsilbm_pair_create(fsilbm_handle father, fsilbm_handle son, int interpolateScheme, int *pair)
{
    Block *F = get(father), *S = get(son);
    if (!F || !S || !pair || father == son) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    // Slab runs: a son lives whole on the rank whose father slab contains its footprint (checked below); the transfers are
    // then local to that rank, which creates the son and the pair alone.  A son itself cut into slabs is not provided.
    if (S->g.X != S->g.XG)
        return fail(FSILBM_ERR_ARG, "the son block of a refinement pair must be whole on one GPU (only root blocks are cut into x-slabs)");
    const int m_gridDelta = 2;
    const Geom &gf = F->g, &gs = S->g;
    // check_blocks_params, LBMBlockComm.f90:508-544
    int sdim[3] = {gs.X, gs.Y, gs.Z};
    const double smin[3] = {gs.xmin, gs.ymin, gs.zmin}, fmin[3] = {gf.xmin, gf.ymin, gf.zmin};
    double res1 = 0.0, res2 = 0.0;
    bool flag = fabs(gf.dh - gs.dh * (double)m_gridDelta) > 1e-8;
    for (int k = 0; k < 3; k++) {
        const int r = S->periodic[k] == 1 ? 0 : 1;
        flag = flag || (sdim[k] % m_gridDelta) != r;
        double smax = smin[k] + gs.dh * (sdim[k] - 1);            // FluidDomain.f90:94-105
        if (S->periodic[k] == 1) smax = smax + gs.dh;
        res1 = res1 + (smin[k] - fmin[k]) / gf.dh;
        res2 = res2 + (smax - fmin[k]) / gf.dh;
    }
    res1 = fabs(res1 - (double)lround(res1));
    res2 = fabs(res2 - (double)lround(res2));
    if (flag || res1 + res2 > 1e-8)
        return fail(FSILBM_ERR_ARG, "grid points do not match between fluid blocks (LBMBlockComm.f90:537): if son block have periodic boundarys, "
                                    "an even number of grid points is needed. Otherwise an odd number is needed.");
    auto p = std::make_unique<Pair>();
    p->father = father; p->son = son; p->scheme = interpolateScheme;
    // deliver_son_to_father rewrites father nodes inside the son's footprint between the steps: no early IBM on a father.  A son
    // only has its outermost planes rewritten (interpolation_father_to_son), and early IBM keeps 3 cells away from such faces.
    // (set once the pair is known to be valid, at the end)
    // build_blocks_comunication, :32-96
    for (int j = 0; j < 6; j++) p->sds[j] = S->bc[j] == BCfluid ? ((j % 2 == 0) ? 1 : -1) : 0;
    int sD[3];
    for (int k = 0; k < 3; k++) sD[k] = sdim[k] - (S->periodic[k] == 1 ? 1 : 0);
    const int ratio = (int)floor(gf.dh / gs.dh + 0.5);
    for (int k = 0; k < 3; k++) {
        p->s[2 * k] = 1; p->s[2 * k + 1] = sD[k];
        p->f[2 * k] = (int)floor((smin[k] - fmin[k]) / gf.dh + 1.5);
        p->f[2 * k + 1] = p->f[2 * k] + (sD[k] - 1) / ratio;
        p->dimS[k] = sdim[k];
        p->dimF[k] = p->f[2 * k + 1] - p->f[2 * k] + 1;
    }
    const int fdim[3] = {gf.XG, gf.Y, gf.Z};
    for (int k = 0; k < 3; k++)
        if (p->f[2 * k] < 1 || p->f[2 * k + 1] > fdim[k]) return fail(FSILBM_ERR_ARG, "son block is not inside its father along axis %d", k);
    if (p->f[0] - 1 < gf.xOffset || p->f[1] - 1 >= gf.xOffset + gf.X) {
        // the footprint reaches into a neighbouring slab: possible when those planes can be addressed through the peer-mapped
        // population buffers of the halo (one neighbour deep), and the neighbour registers the pair (fsilbm_pair_create_remote)
        Block::Halo &h = F->halo;
        const int lo = p->f[0] - 1, hi = p->f[1] - 1;   // 0-based global father planes touched by the transfers
        const bool want_l = lo < gf.xOffset, want_r = hi >= gf.xOffset + gf.X;
        const bool ok_l = !want_l || (h.enabled && h.left >= 0 && h.left == g_nccl.rank - 1 && h.peer_f[0][0] && lo >= gf.xOffset - h.peer_X[0]);
        const bool ok_r = !want_r || (h.enabled && h.right >= 0 && h.right == g_nccl.rank + 1 && h.peer_f[1][0] && hi < gf.xOffset + gf.X + h.peer_X[1]);
        if (!ok_l || !ok_r)
            return fail(FSILBM_ERR_ARG, "son block spans father planes %d..%d but this rank's father slab is %d..%d: a son may reach into the directly "
                                        "neighbouring slabs only, and only with the peer-memory halo (option \"halo\" = 1, CUDA IPC available)",
                        p->f[0], p->f[1], gf.xOffset + 1, gf.xOffset + gf.X);
        if (F->model >= 11) return fail(FSILBM_ERR_MODEL, "a son across a slab interface of an LES father (tau_all field) is not provided");
        p->cross[0] = want_l; p->cross[1] = want_r;
    }
    for (int j = 0; j < 6; j++) { p->si[j] = p->s[j] + p->sds[j] * ratio; p->fi[j] = p->f[j] + p->sds[j]; }
    // allocate_fIn_tau, :213-264
    // (by collision model, not by whether tau_all exists yet: the reference builds its block tree before initialise_, main.f90:34,50)
    const bool need_tau = F->model >= 11 || S->model >= 11;
    for (int j = 0; j < 6; j++) {
        if (S->bc[j] != BCfluid) continue;
        int axis, bAx, aAx;
        pair_axes(j, axis, bAx, aAx);
        const size_t n = (size_t)p->dimF[bAx] * p->dimF[aAx];
        for (int t = 0; t < 2; t++) {
            CK(cudaMalloc(&p->buf[j][t], sizeof(double) * Q * n));
            CK(cudaMemsetAsync(p->buf[j][t], 0, sizeof(double) * Q * n, g_stream));
            if (need_tau) { CK(cudaMalloc(&p->tbuf[j][t], sizeof(double) * n)); CK(cudaMemsetAsync(p->tbuf[j][t], 0, sizeof(double) * n, g_stream)); }
        }
    }
    int slot = -1;
    for (size_t i = 0; i < g_pairs.size(); i++) if (!g_pairs[i]) { slot = (int)i; break; }
    if (slot < 0) { g_pairs.emplace_back(); slot = (int)g_pairs.size() - 1; }
    for (int sd = 0; sd < 2; sd++) if (p->cross[sd]) F->halo.cross_pairs[sd]++;
    g_pairs[slot] = std::move(p);
    F->is_father = true;
    F->early_ok = S->early_ok = false;
    *pair = slot;
    return 0;
}

// The neighbour's half of a son across a slab interface: rank `owner_rank` (the rank directly left or right of this one) holds a
// son block whose footprint reaches into THIS rank's slab of `father`.  From now on this rank's father block tells the owner
// when each of its steps is complete and does not start the next step before the owner's son->father delivery has landed.
int fsilbm_pair_create_remote(fsilbm_handle father, int owner_rank, int *pair)
{
    Block *F = get(father);
    if (!F || !pair) return fail(FSILBM_ERR_ARG, "bad handle/argument");
    Block::Halo &h = F->halo;
    if (!h.enabled) return fail(FSILBM_ERR_ARG, "a son across a slab interface needs the peer-memory halo (option \"halo\" = 1, CUDA IPC available)");
    int side = -1;
    if (owner_rank == g_nccl.rank - 1 && h.left == owner_rank) side = 0;
    if (owner_rank == g_nccl.rank + 1 && h.right == owner_rank) side = 1;
    if (side < 0) return fail(FSILBM_ERR_ARG, "rank %d is not a direct neighbour of rank %d in the slab decomposition", owner_rank, g_nccl.rank);
    auto p = std::make_unique<Pair>();
    p->father = father; p->remote = true; p->owner_side = side;
    if (h.remote_pairs[side] == 0) h.remote_base[side] = h.step;
    else if (h.remote_base[side] != h.step) return fail(FSILBM_ERR_ARG, "register every son across an interface before the first step after the previous registration");
    h.remote_pairs[side]++;
    int slot = -1;
    for (size_t i = 0; i < g_pairs.size(); i++) if (!g_pairs[i]) { slot = (int)i; break; }
    if (slot < 0) { g_pairs.emplace_back(); slot = (int)g_pairs.size() - 1; }
    g_pairs[slot] = std::move(p);
    F->is_father = true;
    F->early_ok = false;
    *pair = slot;
    return 0;
}

int fsilbm_pair_destroy(int pair)
{
    Pair *p = get_pair(pair);
    if (!p) return fail(FSILBM_ERR_ARG, "bad pair %d", pair);
    cudaStreamSynchronize(g_stream);
    for (int j = 0; j < 6; j++) for (int t = 0; t < 2; t++) { cudaFree(p->buf[j][t]); cudaFree(p->tbuf[j][t]); }
    if (Block *F = get(p->father)) {
        if (p->remote && F->halo.remote_pairs[p->owner_side] > 0) F->halo.remote_pairs[p->owner_side]--;
        for (int sd = 0; sd < 2; sd++) if (p->cross[sd] && F->halo.cross_pairs[sd] > 0) F->halo.cross_pairs[sd]--;
    }
    g_pairs[pair].reset();
    return 0;
}

int fsilbm_pair_info(int pair, int out[36])
{
    Pair *p = get_pair(pair);
    if (!p || !out) return fail(FSILBM_ERR_ARG, "bad pair/argument");
    for (int j = 0; j < 6; j++) { out[j] = p->sds[j]; out[6 + j] = p->s[j]; out[12 + j] = p->f[j]; out[18 + j] = p->si[j]; out[24 + j] = p->fi[j]; }
    for (int k = 0; k < 3; k++) { out[30 + k] = p->dimS[k]; out[33 + k] = p->dimF[k]; }
    return 0;
}

int fsilbm_pair_extract_layer(int pair, int time)
{
    Pair *p = get_pair(pair);
    if (!p || (time != 1 && time != 2)) return fail(FSILBM_ERR_ARG, "bad pair/argument");
    if (p->remote) return 0;   // the neighbour's registration of a son across the interface: the owner does the transfers
    Block *F = get(p->father), *S = get(p->son);
    if (!F || !S) return fail(FSILBM_ERR_ARG, "pair %d refers to a destroyed block", pair);
    await_father_step(*F, *p);
    PairFaces ps{};
    for (int j = 0; j < 6; j++) {
        if (S->bc[j] != BCfluid) continue;   // :354
        ps.face[ps.n++] = pair_face(*p, *F, *S, j, 0);
    }
    launch_pair_extract(ps, time, g_stream);
    TRACE(g_stream, 0, "pair_extract");
    CK(cudaGetLastError());
    return 0;
}

int fsilbm_pair_father_to_son(int pair, int n_timeStep)
{
    Pair *p = get_pair(pair);
    if (!p) return fail(FSILBM_ERR_ARG, "bad pair %d", pair);
    if (p->remote) return 0;
    Block *F = get(p->father), *S = get(p->son);
    if (!F || !S) return fail(FSILBM_ERR_ARG, "pair %d refers to a destroyed block", pair);
    PairFaces ps{};
    for (int j = 0; j < 6; j++) {
        if (p->sds[j] == 0) continue;
        ps.face[ps.n++] = pair_face(*p, *F, *S, j, 0);   // father's volumeForce, dh: :663-664
    }
    launch_pair_f2s(ps, n_timeStep == 0 ? 0 : 1, g_stream);
    TRACE(g_stream, 0, "pair_f2s");
    CK(cudaGetLastError());
    return 0;
}

int fsilbm_pair_son_to_father(int pair)
{
    Pair *p = get_pair(pair);
    if (!p) return fail(FSILBM_ERR_ARG, "bad pair %d", pair);
    if (p->remote) return 0;
    Block *F = get(p->father), *S = get(p->son);
    if (!F || !S) return fail(FSILBM_ERR_ARG, "pair %d refers to a destroyed block", pair);
    await_father_step(*F, *p);
    PairFaces ps{};
    for (int j = 0; j < 6; j++) {
        if (p->sds[j] == 0) continue;
        ps.face[ps.n++] = pair_face(*p, *F, *S, j, 1);   // son's volumeForce, dh: :552-553
    }
    launch_pair_s2f(ps, g_stream);
    for (int sd = 0; sd < 2; sd++) {
        if (!p->cross[sd]) continue;   // tell the neighbour that the father nodes of its slab are rewritten: it may start its next step
        Block::Halo &h = F->halo;
        h.delivered[sd]++;
        launch_flag_signal(refine_flag(sd == 0 ? h.peer_left : h.peer_right, 1, sd ^ 1), h.delivered[sd], g_stream);
    }
    TRACE(g_stream, 0, "pair_s2f");
    CK(cudaGetLastError());
    return 0;
}

// ---- comm -----------------------------------------------------------------------------------------
int fsilbm_comm_unique_id(char id[128])
{
    if (int rc = nccl_load()) return rc;
    ncclUniqueId_t u;
    NCK(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return 0;
}

int fsilbm_comm_init(int rank, int nranks, const char id[128])
{
    if (g_device < 0) return fail(FSILBM_ERR_CUDA, "fsilbm_init has not succeeded");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(FSILBM_ERR_ARG, "bad rank %d of %d", rank, nranks);
    g_nccl.rank = rank; g_nccl.nranks = nranks;
    if (nranks == 1) return 0;
    if (int rc = nccl_load()) return rc;
    ncclUniqueId_t u;
    memcpy(u.internal, id, 128);
    NCK(g_nccl.CommInitRank(&g_nccl.comm, nranks, u, rank));
    return 0;
}

int fsilbm_comm_finalize(void)
{
    if (g_nccl.comm) { g_nccl.CommDestroy(g_nccl.comm); g_nccl.comm = nullptr; }
    g_nccl.rank = 0; g_nccl.nranks = 1;
    return 0;
}

}  // extern "C"
