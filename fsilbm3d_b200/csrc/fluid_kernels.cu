// fluid_kernels.cu -- sm_100a kernels for the D3Q19 fluid block (reference: FluidDomain.f90).
//
// Hot kernel: collide_push_kernel.  One thread per lattice cell, threads contiguous along z (the
// fastest index of fIn(z,y,x,q), FluidDomain.f90:384), so each of the 19 population loads of a warp
// is one contiguous 256-byte segment.  The kernel derives den/uuu in registers
// (calculate_macro_quantities_, :1128), takes the IBM-corrected velocity and IBM force from the
// sparse box arrays where a body is near, applies collision_ (:1208) and writes every population
// to its streamed position in the second buffer (streaming_, :1514, periodic wrap on every axis).
// Per cell: 19 fp64 reads + 19 fp64 writes = 304 bytes, the algorithmic minimum.
// Not a contraction: no tensor cores.  HBM-bound.
#define FSILBM_DEFINE_CONSTANTS
#include "kernels.h"
#include <atomic>

namespace fsilbm {

static std::atomic<long long> g_launches{0};
long long kernel_launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add(n); }

void upload_mrt(int slot, const double *M_COLLID, const double *M_FORCE, cudaStream_t s)
{
    cudaMemcpyToSymbolAsync(c_MRT, M_COLLID, sizeof(double) * Q * Q, sizeof(double) * (size_t)(slot * 2 + 0) * Q * Q,
                            cudaMemcpyHostToDevice, s);
    cudaMemcpyToSymbolAsync(c_MRT, M_FORCE, sizeof(double) * Q * Q, sizeof(double) * (size_t)(slot * 2 + 1) * Q * Q,
                            cudaMemcpyHostToDevice, s);
}

// ---- IBM box lookup -----------------------------------------------------------------------------
__device__ __forceinline__ long long box_lookup(const IbmBoxes &B, int gx, int y, int z, int XG, int Y, int Z)
{
    for (int b = 0; b < B.n; b++) {
        int dx = gx - B.lo[b][0]; if (dx < 0) dx += XG;
        if (dx >= B.ext[b][0]) continue;
        int dy = y - B.lo[b][1]; if (dy < 0) dy += Y;
        if (dy >= B.ext[b][1]) continue;
        int dz = z - B.lo[b][2]; if (dz < 0) dz += Z;
        if (dz >= B.ext[b][2]) continue;
        return B.off[b] + ((long long)dx * B.ext[b][1] + dy) * B.ext[b][2] + dz;
    }
    return -1;
}

// den, uuu, force of one cell as collision_ sees them (LBMBlockComm.f90:285-288 then :293)
__device__ __forceinline__ void cell_state(const double (&f)[Q], const double (&hF)[3], const double (&Fvol)[3],
                                           const IbmBoxes &boxes, bool ibm, int gx, int y, int z, int XG, int Y, int Z,
                                           double &den, double &u1, double &u2, double &u3, double &F1, double &F2, double &F3)
{
    macro_from_f(f, hF[0], hF[1], hF[2], den, u1, u2, u3);
    F1 = Fvol[0]; F2 = Fvol[1]; F3 = Fvol[2];   // 0.d0 + volumeForce, FluidDomain.f90:1201,1188
    if (ibm) {
        const long long c = box_lookup(boxes, gx, y, z, XG, Y, Z);
        if (c >= 0) {
            u1 = boxes.u[c]; u2 = boxes.u[boxes.ncell + c]; u3 = boxes.u[2 * boxes.ncell + c];
            F1 = boxes.force[c] + Fvol[0];
            F2 = boxes.force[boxes.ncell + c] + Fvol[1];
            F3 = boxes.force[2 * boxes.ncell + c] + Fvol[2];
        }
    }
}

// ---- the fused step -----------------------------------------------------------------------------
// Push form: aligned loads, z-shifted stores.
// EDGE (launches of a multi-GPU run that contain the slab's edge planes): populations that leave the slab in x are stored
// straight into the neighbour GPU's receive planes over NVLink (peer-mapped pointers p.halo_hi / p.halo_lo) instead of the
// local ghost plane, and the last CTA of an edge plane publishes the step number to the neighbour's arrival flag, so the
// transfer is part of the compute kernel and needs no copy engine, no NCCL kernel and no host involvement.
// does any stencil box hold cells of global plane gx?  (uniform over a CTA: the box test per cell is skipped elsewhere)
__device__ __forceinline__ bool plane_in_boxes(const IbmBoxes &B, int gx, int XG)
{
    for (int b = 0; b < B.n; b++) {
        int dx = gx - B.lo[b][0]; if (dx < 0) dx += XG;
        if (dx < B.ext[b][0]) return true;
    }
    return false;
}

// Registers: SRT and TRT fit 96 registers without spilling, i.e. FIVE CTAs of 128 threads per SM.  Memory-bound as the kernel
// is, four would do -- the fifth slot is what the cooperative IBM kernel takes when it runs beside this one (early IBM), so
// that it no longer pushes the update below the occupancy it needs.  MRT and the LES closures spill at 96 and keep four.
template <int MODEL, bool IBM, bool EDGE>
__global__ void __launch_bounds__(128, (MODEL <= 2 ? 5 : 4)) collide_push_kernel(const __grid_constant__ StepParams p)
{
    const int Z = p.g.Z, Y = p.g.Y, X = p.g.X;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    int x = 0;
    {   // blockIdx.z walks the plane ranges of this launch in list order
        int bz = blockIdx.z;
#pragma unroll 1
        for (int sgm = 0; sgm < p.nseg; sgm++) {
            const int cnt = p.seg_count[sgm];
            if (bz < cnt) { x = p.seg_begin[sgm] + bz; break; }
            bz -= cnt;
        }
    }
    const bool edge_plane = EDGE && (x == 0 || x == X - 1);
    const bool active = (z < Z && y < Y);
    if (!edge_plane && !active) return;
    const size_t plane = p.g.plane, ps = p.g.pstride;
    if (active) {
    const size_t base = (size_t)(x + 1) * plane + (size_t)y * Z + z;

    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = __ldg(p.fA + q * ps + base);

    double den, u1, u2, u3, F1, F2, F3;
    const bool ibm_here = IBM && plane_in_boxes(p.boxes, p.g.xOffset + x, p.g.XG);
    cell_state(f, p.hF, p.Fvol, p.boxes, ibm_here, p.g.xOffset + x, y, z, p.g.XG, Y, Z, den, u1, u2, u3, F1, F2, F3);
    if (MODEL >= 11) {
        const LesCtx les{p.uuu, p.tau_all, p.uuu_ncomp, X, Y, Z, p.g.xOffset + x, p.g.XG, x, y, z, true};
        collide<MODEL>(f, den, u1, u2, u3, F1, F2, F3, p.cc, &les);
    } else {
        collide<MODEL>(f, den, u1, u2, u3, F1, F2, F3, p.cc);
    }

    // streaming_: population q moves to (x+ex, y+ey, z+ez), periodic wrap on y and z always, on x
    // inside the slab only when this rank holds the whole x extent; otherwise into the ghost planes.
    int xp = x + 1, xm = x - 1;
    if (p.wrap_x) { if (xp == X) xp = 0; if (xm < 0) xm = X - 1; }
    const int yp = (y + 1 == Y) ? 0 : y + 1, ym = (y == 0) ? Y - 1 : y - 1;
    const int zp = (z + 1 == Z) ? 0 : z + 1, zm = (z == 0) ? Z - 1 : z - 1;
    const size_t ox[3] = {(size_t)(x + 1) * plane, (size_t)(xp + 1) * plane, (size_t)(xm + 1) * plane};
    const size_t oy[3] = {(size_t)y * Z, (size_t)yp * Z, (size_t)ym * Z};
    const size_t oz[3] = {(size_t)z, (size_t)zp, (size_t)zm};
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const int ix = EX(q) == 0 ? 0 : (EX(q) > 0 ? 1 : 2);
        const int iy = EY(q) == 0 ? 0 : (EY(q) > 0 ? 1 : 2);
        const int iz = EZ(q) == 0 ? 0 : (EZ(q) > 0 ? 1 : 2);
        double *dst = p.fB + q * ps + ox[ix] + oy[iy] + oz[iz];
        if (EDGE) {
            if (EX(q) > 0 && x == X - 1 && p.halo_hi) dst = p.halo_hi + (size_t)q * p.halo_hi_ps + oy[iy] + oz[iz];
            if (EX(q) < 0 && x == 0 && p.halo_lo) dst = p.halo_lo + (size_t)q * p.halo_lo_ps + oy[iy] + oz[iz];
        }
        *dst = f[q];
    }
    }
    if (edge_plane) {
        // publish: the CTA's stores (peer stores included) are ordered before thread 0 by the barrier, thread 0 alone fences them
        // system-wide (fences are cumulative; one membar.sys per CTA instead of one per thread, which held the whole launch up
        // by ~40 us), the CTA counts in, the last CTA of the plane raises the flag
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            __threadfence_system();
            const unsigned int total = gridDim.x * gridDim.y;
            unsigned int *counter = p.cta_counter + ((x == 0) ? 0 : 1);
            const unsigned int done = atomicAdd(counter, 1u);
            if (done == total - 1) {
                *counter = 0;
                __threadfence_system();
                if (p.sig_hi && x == X - 1) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.sig_hi), "l"(p.step) : "memory"); }
                if (p.sig_lo && x == 0) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.sig_lo), "l"(p.step) : "memory"); }
            }
        }
    }
}

static inline void line_block(int Z, dim3 &block, int &bz, int &by)
{
    bz = Z >= 128 ? 128 : ((Z + 31) / 32) * 32;
    by = 128 / bz; if (by < 1) by = 1;
    block = dim3(bz, by, 1);
}

void step_add_planes(StepParams &p, int begin, int count)
{
    if (count <= 0) return;
    if (p.nseg > 0 && p.seg_begin[p.nseg - 1] + p.seg_count[p.nseg - 1] == begin) { p.seg_count[p.nseg - 1] += count; return; }
    if (p.nseg >= MAX_SEG) return;   // cannot happen: at most 2 edge planes + MAX_BOXES ranges + MAX_BOXES + 1 gaps
    p.seg_begin[p.nseg] = begin; p.seg_count[p.nseg] = count; p.nseg++;
}

int launch_collide_push(const StepParams &p, int model, cudaStream_t s)
{
    int planes = 0;
    for (int i = 0; i < p.nseg; i++) planes += p.seg_count[i];
    if (planes <= 0) return 0;
    dim3 block; int bz, by;
    line_block(p.g.Z, block, bz, by);
    dim3 grid((p.g.Z + bz - 1) / bz, (p.g.Y + by - 1) / by, planes);
    const bool ibm = p.boxes.n > 0;
    const bool edge = p.cta_counter != nullptr;
#define FSILBM_LAUNCH(M)                                                                                     \
    do {                                                                                                     \
        if (edge) { if (ibm) collide_push_kernel<M, true, true><<<grid, block, 0, s>>>(p); else collide_push_kernel<M, false, true><<<grid, block, 0, s>>>(p); } \
        else { if (ibm) collide_push_kernel<M, true, false><<<grid, block, 0, s>>>(p); else collide_push_kernel<M, false, false><<<grid, block, 0, s>>>(p); } \
    } while (0)
    if (model == 1) FSILBM_LAUNCH(1);
    else if (model == 2) FSILBM_LAUNCH(2);
    else if (model == 3) FSILBM_LAUNCH(3);
    else if (model == 11) FSILBM_LAUNCH(11);
    else if (model == 14) FSILBM_LAUNCH(14);
    else if (model == 15) FSILBM_LAUNCH(15);
    else return 1;
#undef FSILBM_LAUNCH
    count_launch();
    return 0;
}

// ---- initialise_flow, FluidDomain.f90:524-544 -----------------------------------------------------
__global__ void initialise_kernel(Geom g, double *f, VelocityField vel, double denIn)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.z;
    if (z >= g.Z || y >= g.Y) return;
    const double xC = g.xmin + g.dh * (double)(g.xOffset + x);
    const double yC = g.ymin + g.dh * (double)y;
    const double zC = g.zmin + g.dh * (double)z;
    double v1, v2, v3, d[Q];
    evaluate_velocity(vel, zC, yC, xC, v1, v2, v3);
    equilibrium(denIn, v1, v2, v3, d);
    const size_t base = (size_t)(x + 1) * g.plane + (size_t)y * g.Z + z;
#pragma unroll
    for (int q = 0; q < Q; q++) f[q * g.pstride + base] = d[q];
}

void launch_initialise(const Geom &g, double *f, const VelocityField &vel, double denIn, cudaStream_t s)
{
    dim3 block; int bz, by;
    line_block(g.Z, block, bz, by);
    dim3 grid((g.Z + bz - 1) / bz, (g.Y + by - 1) / by, g.X);
    initialise_kernel<<<grid, block, 0, s>>>(g, f, vel, denIn);
    count_launch();
}

// ---- calculate_macro_quantities_ over the slab (writers, probes, un-fused pass) --------------------
__global__ void macro_full_kernel(Geom g, const double *f, double hF1, double hF2, double hF3, double *den, double *uuu, size_t ncomp,
                                  const __grid_constant__ IbmBoxes boxes)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.z;
    if (z >= g.Z || y >= g.Y) return;
    const size_t base = (size_t)(x + 1) * g.plane + (size_t)y * g.Z + z;
    double fl[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) fl[q] = f[q * g.pstride + base];
    double d, u1, u2, u3;
    macro_from_f(fl, hF1, hF2, hF3, d, u1, u2, u3);
    if (boxes.n > 0) {   // the IBM-corrected velocity where a body is near (what collision_ and the LES differences see)
        const long long bc = box_lookup(boxes, g.xOffset + x, y, z, g.XG, g.Y, g.Z);
        if (bc >= 0) { u1 = boxes.u[bc]; u2 = boxes.u[boxes.ncell + bc]; u3 = boxes.u[2 * boxes.ncell + bc]; }
    }
    const size_t c = (size_t)x * g.plane + (size_t)y * g.Z + z;
    if (den) den[c] = d;
    if (uuu) { uuu[c] = u1; uuu[ncomp + c] = u2; uuu[2 * ncomp + c] = u3; }
}

void launch_macro_full(const Geom &g, const double *f, const double hF[3], double *den, double *uuu, cudaStream_t s, const IbmBoxes *boxes, size_t ncomp)
{
    if (!ncomp) ncomp = (size_t)g.X * g.plane;
    dim3 block; int bz, by;
    line_block(g.Z, block, bz, by);
    dim3 grid((g.Z + bz - 1) / bz, (g.Y + by - 1) / by, g.X);
    IbmBoxes none{};
    macro_full_kernel<<<grid, block, 0, s>>>(g, f, hF[0], hF[1], hF[2], den, uuu, ncomp, boxes ? *boxes : none);
    count_launch();
}

// ---- ComputeFieldStat_, FluidDomain.f90:1739-1768 (partial sums; the host finishes) -----------------
__global__ void field_stat_kernel(Geom g, const double *f, double hF1, double hF2, double hF3, double invUref, double *out6)
{
    // grid-stride over cells; block reduce; atomics on 6 doubles (max via CAS on the bit pattern of non-negative doubles)
    double s[3] = {0.0, 0.0, 0.0}, m[3] = {-1.0, -1.0, -1.0};
    const size_t ncell = (size_t)g.X * g.plane;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (size_t)gridDim.x * blockDim.x) {
        const size_t x = c / g.plane, r = c - x * g.plane;
        const size_t base = (x + 1) * g.plane + r;
        double fl[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) fl[q] = f[q * g.pstride + base];
        double d, u[3];
        macro_from_f(fl, hF1, hF2, hF3, d, u[0], u[1], u[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double t = fabs(u[k] * invUref);
            s[k] += t * t;
            if (t > m[k]) m[k] = t;
        }
    }
    __shared__ double sh[6][256];
    for (int k = 0; k < 3; k++) { sh[k][threadIdx.x] = s[k]; sh[3 + k][threadIdx.x] = m[k]; }
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off)
            for (int k = 0; k < 3; k++) {
                sh[k][threadIdx.x] += sh[k][threadIdx.x + off];
                sh[3 + k][threadIdx.x] = fmax(sh[3 + k][threadIdx.x], sh[3 + k][threadIdx.x + off]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int k = 0; k < 3; k++) {
            atomicAdd(out6 + k, sh[k][0]);
            // max of non-negative doubles == max of their bit patterns as signed 64-bit integers
            if (sh[3 + k][0] >= 0.0) atomicMax((long long *)(out6 + 3 + k), __double_as_longlong(sh[3 + k][0]));
        }
}

void launch_field_stat(const Geom &g, const double *f, const double hF[3], double invUref, double *out6, cudaStream_t s)
{
    field_stat_kernel<<<148 * 4, 256, 0, s>>>(g, f, hF[0], hF[1], hF[2], invUref, out6);
    count_launch();
}

// ---- face helpers -----------------------------------------------------------------------------------
// face-local (a,b) and layer (0 = boundary layer) -> local cell (x,y,z)
__device__ __forceinline__ void face_cell(const Geom &g, int face, int a, int b, int layer, int &x, int &y, int &z)
{
    const int axis = face >> 1, hi = face & 1;
    if (axis == 0) { z = a; y = b; x = hi ? g.X - 1 - layer : layer; }
    else if (axis == 1) { z = a; x = b; y = hi ? g.Y - 1 - layer : layer; }
    else { y = a; x = b; z = hi ? g.Z - 1 - layer : layer; }
}
__device__ __forceinline__ size_t cell_index(const Geom &g, int x, int y, int z) { return (size_t)(x + 1) * g.plane + (size_t)y * g.Z + z; }

// ---- set_boundary_conditions_, FluidDomain.f90:616-1126: one face, one thread per face node ---------
__device__ __forceinline__ void bc_face_node(const FaceParams &p, const int a, const int b)
{
    const Geom &g = p.g;
    const int face = p.face, axis = face >> 1;
    int x, y, z, x2, y2, z2, x3, y3, z3;
    face_cell(g, face, a, b, 0, x, y, z);
    face_cell(g, face, a, b, 1, x2, y2, z2);
    face_cell(g, face, a, b, 2, x3, y3, z3);
    const size_t c1 = cell_index(g, x, y, z), c2 = cell_index(g, x2, y2, z2), c3 = cell_index(g, x3, y3, z3);
    const size_t ps = g.pstride;
    double *f = p.f;
    double xC = g.xmin + g.dh * (double)(g.xOffset + x), yC = g.ymin + g.dh * (double)y, zC = g.zmin + g.dh * (double)z;
    if (axis == 0) xC = p.wallc; else if (axis == 1) yC = p.wallc; else zC = p.wallc;
    const size_t sidx = (size_t)b * p.na + a, sstride = (size_t)p.na * p.nb;

    // the 5 incoming populations of this face, selected at run time from the constexpr tables
    int I[5], Mi[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        I[k] = face == 0 ? FACE_IN(0, k) : face == 1 ? FACE_IN(1, k) : face == 2 ? FACE_IN(2, k) : face == 3 ? FACE_IN(3, k) : face == 4 ? FACE_IN(4, k) : FACE_IN(5, k);
        Mi[k] = face == 0 ? FACE_MIRROR(0, k) : face == 1 ? FACE_MIRROR(1, k) : face == 2 ? FACE_MIRROR(2, k) : face == 3 ? FACE_MIRROR(3, k) : face == 4 ? FACE_MIRROR(4, k) : FACE_MIRROR(5, k);
    }
    constexpr int oppoT[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};

    double v1, v2, v3;
    switch (p.code) {
    case BCEq_DirecletU: {   // :623-632
        double fe[Q];
        evaluate_velocity(p.vel, zC, yC, xC, v1, v2, v3);
        equilibrium(p.denIn, v1, v2, v3, fe);
#pragma unroll
        for (int q = 0; q < Q; q++) f[q * ps + c1] = fe[q];
        break;
    }
    case BCnEq_DirecletU: {  // :633-647
        double fe[Q], fei[Q];
        evaluate_velocity(p.vel, zC, yC, xC, v1, v2, v3);
        equilibrium(p.denIn, v1, v2, v3, fe);
        equilibrium(p.l2den[sidx], p.l2u[sidx], p.l2u[sstride + sidx], p.l2u[2 * sstride + sidx], fei);
        for (int k = 0; k < 5; k++) {
            const int q = I[k];
            double feq = 0.0, feiq = 0.0;
#pragma unroll
            for (int j = 0; j < Q; j++) if (j == q) { feq = fe[j]; feiq = fei[j]; }
            f[q * ps + c1] = feq + (f[q * ps + c2] - feiq);
        }
        break;
    }
    case BCorder1_Extrapolate:  // :648-649
        for (int k = 0; k < 5; k++) f[I[k] * ps + c1] = f[I[k] * ps + c2];
        break;
    case BCorder2_Extrapolate:  // :650-651
        for (int k = 0; k < 5; k++) f[I[k] * ps + c1] = 2.0 * f[I[k] * ps + c2] - f[I[k] * ps + c3];
        break;
    case BCstationary_Wall: {   // :652-658
        double t[5];
        for (int k = 0; k < 5; k++) t[k] = f[oppoT[I[k]] * ps + c1];
        for (int k = 0; k < 5; k++) f[I[k] * ps + c1] = t[k];
        break;
    }
    case BCstationary_Wall_halfway: {  // :659-669
        for (int k = 0; k < 5; k++) f[I[k] * ps + c1] = p.stash[oppoT[I[k]] * sstride + sidx];
        break;
    }
    case BCmoving_Wall:            // :670-679
    case BCmoving_Wall_halfway: {  // :680-693
        double in[Q], out[Q];
        evaluate_velocity(p.vel, zC, yC, xC, v1, v2, v3);
        if (p.code == BCmoving_Wall) {
#pragma unroll
            for (int q = 0; q < Q; q++) in[q] = f[q * ps + c1];
        } else {
#pragma unroll
            for (int q = 0; q < Q; q++) in[q] = p.stash[q * sstride + sidx];
        }
        moving_wall(p.denIn, v1, v2, v3, in, out);
        for (int k = 0; k < 5; k++) {
            const int q = I[k];
            double o = 0.0;
#pragma unroll
            for (int j = 0; j < Q; j++) if (j == q) o = out[j];
            f[q * ps + c1] = o;
        }
        break;
    }
    case BCSymmetric: {  // :694-700
        double t[5];
        for (int k = 0; k < 5; k++) t[k] = f[Mi[k] * ps + c1];
        for (int k = 0; k < 5; k++) f[I[k] * ps + c1] = t[k];
        break;
    }
    default: break;
    }
}

__global__ void bc_face_kernel(const __grid_constant__ FaceParams p)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < p.na) bc_face_node(p, a, blockIdx.y);
}

// The two faces of one axis in one launch (blockIdx.z = 0: low face, 1: high face).  Exact as long as the two rules touch
// disjoint cells -- each reads and writes its own three outermost layers only -- which the caller guarantees (extent >= 6).
// Faces of different axes stay separate launches in the reference's order: they share edge lines, where a later face
// reads what an earlier one wrote (FluidDomain.f90:622-1125).
__global__ void bc_face_pair_kernel(const __grid_constant__ FaceParams lo, const __grid_constant__ FaceParams hi)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= lo.na) return;
    if (blockIdx.z == 0) bc_face_node(lo, a, blockIdx.y); else bc_face_node(hi, a, blockIdx.y);
}

static inline void face_grid(const FaceParams &p, dim3 &grid, dim3 &block)
{
    block = dim3(128, 1, 1);
    grid = dim3((p.na + 127) / 128, p.nb, 1);
}

void launch_bc_face(const FaceParams &p, cudaStream_t s)
{
    dim3 grid, block;
    face_grid(p, grid, block);
    bc_face_kernel<<<grid, block, 0, s>>>(p);
    count_launch();
}

void launch_bc_face_pair(const FaceParams &lo, const FaceParams &hi, cudaStream_t s)
{
    dim3 grid, block;
    face_grid(lo, grid, block);
    grid.z = 2;
    bc_face_pair_kernel<<<grid, block, 0, s>>>(lo, hi);
    count_launch();
}

// ---- halfwayBCset_ inside the fused step: post-collision boundary layer -> stash ---------------------
// (FluidDomain.f90:567-614 copies fIn of the boundary layer after collision_; here the boundary layer is
//  collided once more from the pre-collision buffer with the identical arithmetic.)
template <int MODEL>
__global__ void stash_face_kernel(const __grid_constant__ FaceParams p)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= p.na) return;
    const Geom &g = p.g;
    int x, y, z;
    face_cell(g, p.face, a, b, 0, x, y, z);
    const size_t c1 = cell_index(g, x, y, z);
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = p.fA[q * g.pstride + c1];
    double den, u1, u2, u3, F1, F2, F3;
    cell_state(f, p.hF, p.Fvol, p.boxes, p.boxes.n > 0, g.xOffset + x, y, z, g.XG, g.Y, g.Z, den, u1, u2, u3, F1, F2, F3);
    if (MODEL >= 11) {
        const LesCtx les{p.uuu, p.tau_all, p.uuu_ncomp, g.X, g.Y, g.Z, g.xOffset + x, g.XG, x, y, z, false};   // the main kernel writes tau_all, not this re-collision
        collide<MODEL>(f, den, u1, u2, u3, F1, F2, F3, p.cc, &les);
    } else {
        collide<MODEL>(f, den, u1, u2, u3, F1, F2, F3, p.cc);
    }
    const size_t sidx = (size_t)b * p.na + a, sstride = (size_t)p.na * p.nb;
#pragma unroll
    for (int q = 0; q < Q; q++) p.stash[q * sstride + sidx] = f[q];
}

void launch_stash_face(const FaceParams &p, cudaStream_t s)
{
    dim3 grid, block;
    face_grid(p, grid, block);
    if (p.model == 1) stash_face_kernel<1><<<grid, block, 0, s>>>(p);
    else if (p.model == 2) stash_face_kernel<2><<<grid, block, 0, s>>>(p);
    else if (p.model == 3) stash_face_kernel<3><<<grid, block, 0, s>>>(p);
    else if (p.model == 11) stash_face_kernel<11><<<grid, block, 0, s>>>(p);
    else if (p.model == 14) stash_face_kernel<14><<<grid, block, 0, s>>>(p);
    else stash_face_kernel<15><<<grid, block, 0, s>>>(p);
    count_launch();
}

// ---- den/uuu of the first interior layer for BCnEq_DirecletU (FluidDomain.f90:643) ---------------------
__global__ void layer2_face_kernel(const __grid_constant__ FaceParams p)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= p.na) return;
    const Geom &g = p.g;
    int x, y, z;
    face_cell(g, p.face, a, b, 1, x, y, z);
    const size_t c = cell_index(g, x, y, z);
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = p.fA[q * g.pstride + c];
    double den, u1, u2, u3, F1, F2, F3;
    cell_state(f, p.hF, p.Fvol, p.boxes, p.boxes.n > 0, g.xOffset + x, y, z, g.XG, g.Y, g.Z, den, u1, u2, u3, F1, F2, F3);
    const size_t sidx = (size_t)b * p.na + a, sstride = (size_t)p.na * p.nb;
    p.l2den[sidx] = den;
    p.l2u[sidx] = u1; p.l2u[sstride + sidx] = u2; p.l2u[2 * sstride + sidx] = u3;
}

void launch_layer2_face(const FaceParams &p, cudaStream_t s)
{
    dim3 grid, block;
    face_grid(p, grid, block);
    layer2_face_kernel<<<grid, block, 0, s>>>(p);
    count_launch();
}

// den = denIn, uuu = evaluate_velocity as initialise_ leaves them (FluidDomain.f90:535-536), for the
// start-up boundary call of main.f90:63
__global__ void init_layer2_kernel(const __grid_constant__ FaceParams p)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= p.na) return;
    const Geom &g = p.g;
    int x, y, z;
    face_cell(g, p.face, a, b, 1, x, y, z);
    const double xC = g.xmin + g.dh * (double)(g.xOffset + x), yC = g.ymin + g.dh * (double)y, zC = g.zmin + g.dh * (double)z;
    double v1, v2, v3;
    evaluate_velocity(p.vel, zC, yC, xC, v1, v2, v3);
    const size_t sidx = (size_t)b * p.na + a, sstride = (size_t)p.na * p.nb;
    p.l2den[sidx] = p.denIn;
    p.l2u[sidx] = v1; p.l2u[sstride + sidx] = v2; p.l2u[2 * sstride + sidx] = v3;
}

void launch_init_layer2(const FaceParams &p, cudaStream_t s)
{
    dim3 grid, block;
    face_grid(p, grid, block);
    init_layer2_kernel<<<grid, block, 0, s>>>(p);
    count_launch();
}

// ---- un-fused passes ------------------------------------------------------------------------------
__global__ void fill_kernel(double *p, size_t n, double v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
void launch_pass_fill(double *p, size_t n, double v, cudaStream_t s)
{
    fill_kernel<<<148 * 8, 256, 0, s>>>(p, n, v);
    count_launch();
}

// add_volume_force_, FluidDomain.f90:1182-1193
__global__ void add_force_kernel(double *force, size_t n, double F1, double F2, double F3)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        force[i] = force[i] + F1;
        force[n + i] = force[n + i] + F2;
        force[2 * n + i] = force[2 * n + i] + F3;
    }
}
void launch_pass_add_force(const FieldParams &p, cudaStream_t s)
{
    add_force_kernel<<<148 * 8, 256, 0, s>>>(p.force, (size_t)p.g.X * p.g.plane, p.Fvol[0], p.Fvol[1], p.Fvol[2]);
    count_launch();
}

// collision_ as its own pass on stored den/uuu/force, FluidDomain.f90:1208-1263
template <int MODEL>
__global__ void collision_fields_kernel(const __grid_constant__ FieldParams p)
{
    const Geom &g = p.g;
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.z;
    if (z >= g.Z || y >= g.Y) return;
    const size_t base = cell_index(g, x, y, z);
    const size_t n = (size_t)g.X * g.plane, c = (size_t)x * g.plane + (size_t)y * g.Z + z;
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = p.f[q * g.pstride + base];
    if (MODEL >= 11) {
        const LesCtx les{p.uuu, p.tau_all, n, g.X, g.Y, g.Z, g.xOffset + x, g.XG, x, y, z, true};
        collide<MODEL>(f, p.den[c], p.uuu[c], p.uuu[n + c], p.uuu[2 * n + c], p.force[c], p.force[n + c], p.force[2 * n + c], p.cc, &les);
    } else {
        collide<MODEL>(f, p.den[c], p.uuu[c], p.uuu[n + c], p.uuu[2 * n + c], p.force[c], p.force[n + c], p.force[2 * n + c], p.cc);
    }
#pragma unroll
    for (int q = 0; q < Q; q++) p.f[q * g.pstride + base] = f[q];
}
int launch_pass_collision(const FieldParams &p, int model, cudaStream_t s)
{
    dim3 block; int bz, by;
    line_block(p.g.Z, block, bz, by);
    dim3 grid((p.g.Z + bz - 1) / bz, (p.g.Y + by - 1) / by, p.g.X);
    if (model == 1) collision_fields_kernel<1><<<grid, block, 0, s>>>(p);
    else if (model == 2) collision_fields_kernel<2><<<grid, block, 0, s>>>(p);
    else if (model == 3) collision_fields_kernel<3><<<grid, block, 0, s>>>(p);
    else if (model == 11) collision_fields_kernel<11><<<grid, block, 0, s>>>(p);
    else if (model == 14) collision_fields_kernel<14><<<grid, block, 0, s>>>(p);
    else if (model == 15) collision_fields_kernel<15><<<grid, block, 0, s>>>(p);
    else return 1;
    count_launch();
    return 0;
}

// halfwayBCset_ as its own pass: copy the boundary layer of f to the stash, FluidDomain.f90:567-614
__global__ void halfway_copy_kernel(const __grid_constant__ FaceParams p)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (a >= p.na) return;
    int x, y, z;
    face_cell(p.g, p.face, a, b, 0, x, y, z);
    const size_t c1 = cell_index(p.g, x, y, z);
    const size_t sidx = (size_t)b * p.na + a, sstride = (size_t)p.na * p.nb;
#pragma unroll
    for (int q = 0; q < Q; q++) p.stash[q * sstride + sidx] = p.f[q * p.g.pstride + c1];
}
void launch_pass_halfway(const FaceParams &p, cudaStream_t s)
{
    dim3 grid, block;
    face_grid(p, grid, block);
    halfway_copy_kernel<<<grid, block, 0, s>>>(p);
    count_launch();
}

// streaming_ as its own pass (periodic shift of every population into the second buffer), :1514-1625
__global__ void streaming_kernel(Geom g, const double *fA, double *fB)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.z;
    if (z >= g.Z || y >= g.Y) return;
    const size_t base = cell_index(g, x, y, z);
    int xp = x + 1, xm = x - 1;
    if (xp == g.X) xp = 0;
    if (xm < 0) xm = g.X - 1;
    const int yp = (y + 1 == g.Y) ? 0 : y + 1, ym = (y == 0) ? g.Y - 1 : y - 1;
    const int zp = (z + 1 == g.Z) ? 0 : z + 1, zm = (z == 0) ? g.Z - 1 : z - 1;
    const size_t ox[3] = {(size_t)(x + 1) * g.plane, (size_t)(xp + 1) * g.plane, (size_t)(xm + 1) * g.plane};
    const size_t oy[3] = {(size_t)y * g.Z, (size_t)yp * g.Z, (size_t)ym * g.Z};
    const size_t oz[3] = {(size_t)z, (size_t)zp, (size_t)zm};
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const int ix = EX(q) == 0 ? 0 : (EX(q) > 0 ? 1 : 2);
        const int iy = EY(q) == 0 ? 0 : (EY(q) > 0 ? 1 : 2);
        const int iz = EZ(q) == 0 ? 0 : (EZ(q) > 0 ? 1 : 2);
        fB[q * g.pstride + ox[ix] + oy[iy] + oz[iz]] = fA[q * g.pstride + base];
    }
}
void launch_pass_streaming(const Geom &g, const double *fA, double *fB, cudaStream_t s)
{
    dim3 block; int bz, by;
    line_block(g.Z, block, bz, by);
    dim3 grid((g.Z + bz - 1) / bz, (g.Y + by - 1) / by, g.X);
    streaming_kernel<<<grid, block, 0, s>>>(g, fA, fB);
    count_launch();
}

// Single-rank fold of the ghost planes back into the slab (only used when a caller forces ghost-plane
// streaming on one rank, e.g. to test the multi-rank path without a second GPU).
__global__ void wrap_x_kernel(Geom g, double *f)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.plane) return;
    constexpr int up[5] = {1, 7, 9, 11, 13}, dn[5] = {2, 8, 10, 12, 14};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        f[up[k] * g.pstride + (size_t)1 * g.plane + i] = f[up[k] * g.pstride + (size_t)(g.X + 1) * g.plane + i];
        f[dn[k] * g.pstride + (size_t)g.X * g.plane + i] = f[dn[k] * g.pstride + i];
    }
}
void launch_wrap_x(const Geom &g, double *f, cudaStream_t s)
{
    wrap_x_kernel<<<(unsigned)((g.plane + 255) / 256), 256, 0, s>>>(g, f);
    count_launch();
}

// Receiving side of the peer-memory halo: the neighbours' edge planes store straight into this rank's streamed buffer, so all
// that is left to do is to wait -- one thread, acquire loads at system scope -- until both have published this step.
__global__ void halo_wait_kernel(const __grid_constant__ HaloWaitParams p)
{
    if (threadIdx.x != 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const unsigned long long *flags[2] = {p.flag_lo, p.flag_hi};
    for (int sd = 0; sd < 2; sd++) {
        if (!flags[sd]) continue;
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags[sd]) : "memory");
            if (v >= p.step) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > p.timeout_ns) { atomicExch(p.err, 1); return; }
            __nanosleep(200);
        }
    }
}
void launch_halo_wait(const HaloWaitParams &p, cudaStream_t s)
{
    halo_wait_kernel<<<1, 32, 0, s>>>(p);
    count_launch();
}

}  // namespace fsilbm
