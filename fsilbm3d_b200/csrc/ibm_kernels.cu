// ibm_kernels.cu -- sm_100a kernels of the immersed-boundary coupling (reference: Solidbody.f90).
//
// The reference keeps uuu and force as full Eulerian fields and loops serially over markers when it
// spreads (Solidbody.f90:1034-1048, :938-978).  Here the corrected velocity and the IBM force exist
// only inside small boxes around the bodies (IbmBoxes); interpolation and spreading are
// warp-cooperative: one warp per Lagrangian marker, two of the 64 stencil nodes per lane, shuffle
// reduction for the gather, fp64 atomics for the scatter.  Stencil weights are rounded to fp32 and
// indices held as int16 exactly as the reference stores them (Solidbody.f90:45-46,790-804).
#include "kernels.h"

#include <cub/device/device_scan.cuh>

namespace fsilbm {

// Phi, Solidbody.f90:822-833
__device__ __forceinline__ double Phi(double x_)
{
    const double r = fabs(x_);
    if (r < 1.0) return (3.0 - 2.0 * r + sqrt(1.0 + 4.0 * r * (1.0 - r))) * 0.125;
    else if (r < 2.0) return (5.0 - 2.0 * r - sqrt(-7.0 + 4.0 * r * (3.0 - r))) * 0.125;
    return 0.0;
}

// trimedindex, Solidbody.f90:834-866 (1-based). Returns false where the reference stops.
__device__ __forceinline__ bool trimedindex(int i_, int xDim_, int (&ix_)[4], int bcLo, int bcHi)
{
#pragma unroll
    for (int k_ = -1; k_ <= 2; k_++) {
        int v = i_ + k_;
        if (v < 1) {
            if (bcLo == BCPeriodic) v = v + xDim_;
            else if ((bcLo == BCSymmetric || bcLo == BCstationary_Wall) && v == 0) v = 2;
            else if (bcLo == BCstationary_Wall_halfway && v == 0) v = 1;
            else return false;
        } else if (v > xDim_) {
            if (bcHi == BCPeriodic) v = v - xDim_;
            else if ((bcHi == BCSymmetric || bcHi == BCstationary_Wall) && v == xDim_ + 1) v = xDim_ - 1;
            else if (bcHi == BCstationary_Wall_halfway && v == xDim_ + 1) v = xDim_;
            else return false;
        }
        if (v < 1 || v > xDim_) return false;   // periodic image still outside: the reference would index out of bounds
        ix_[k_ + 1] = v;
    }
    return true;
}

struct RootBC { int c[6]; };

// UpdateElmtInterp_, Solidbody.f90:760-806, one marker
__device__ __forceinline__ void stencil_marker(const Geom &g, const IbmBody &b, const IbmBoxes &boxes, const RootBC &bc, IbmCtl *ctl, int iEL)
{
    const double dh = g.dh;
    const double invdh = 1.0 / dh;
    // anchor on the body's first marker, :772-780
    int i0 = (int)floor((b.Exyz[0] - g.xmin) * invdh);
    const double x0 = g.xmin + (double)i0 * dh; i0 = i0 + 1;
    int j0 = (int)floor((b.Exyz[1] - g.ymin) * invdh);
    const double y0 = g.ymin + (double)j0 * dh; j0 = j0 + 1;
    int k0 = (int)floor((b.Exyz[2] - g.zmin) * invdh);
    const double z0 = g.zmin + (double)k0 * dh; k0 = k0 + 1;
    // minloc_fast, :811-821
    double detx = (b.Exyz[3 * iEL + 0] - x0) * invdh; int i = (int)floor(detx); detx = detx - (double)i; i = i + i0;
    double dety = (b.Exyz[3 * iEL + 1] - y0) * invdh; int j = (int)floor(dety); dety = dety - (double)j; j = j + j0;
    double detz = (b.Exyz[3 * iEL + 2] - z0) * invdh; int k = (int)floor(detz); detz = detz - (double)k; k = k + k0;
    int ix[4], jy[4], kz[4];
    const bool ok = trimedindex(i, g.XG, ix, bc.c[0], bc.c[1]) && trimedindex(j, g.Y, jy, bc.c[2], bc.c[3]) &&
                    trimedindex(k, g.Z, kz, bc.c[4], bc.c[5]);
    if (!ok) {
        atomicOr(&ctl->err, 1);
        for (int m = 0; m < 12; m++) { b.cell[12 * iEL + m] = 0; b.Ei[12 * iEL + m] = 0; b.Ew[12 * iEL + m] = 0.f; }
        for (int m = 0; m < 4; m++) b.owned[4 * iEL + m] = 0;
        b.boff[iEL] = 0;
        return;
    }
#pragma unroll
    for (int m = 0; m < 4; m++) {
        b.Ei[12 * iEL + m] = (short)ix[m];
        b.Ei[12 * iEL + 4 + m] = (short)jy[m];
        b.Ei[12 * iEL + 8 + m] = (short)kz[m];
        b.Ew[12 * iEL + m] = (float)Phi((double)(m - 1) - detx);
        b.Ew[12 * iEL + 4 + m] = (float)Phi((double)(m - 1) - dety);
        b.Ew[12 * iEL + 8 + m] = (float)Phi((double)(m - 1) - detz);
    }
    // locate the box of this marker (the one holding the stencil's base node) and express all 12
    // indices relative to it
    int bsel = -1;
    for (int bb = 0; bb < boxes.n; bb++) {
        int dx = (ix[1] - 1) - boxes.lo[bb][0]; if (dx < 0) dx += g.XG;
        int dy = (jy[1] - 1) - boxes.lo[bb][1]; if (dy < 0) dy += g.Y;
        int dz = (kz[1] - 1) - boxes.lo[bb][2]; if (dz < 0) dz += g.Z;
        if (dx < boxes.ext[bb][0] && dy < boxes.ext[bb][1] && dz < boxes.ext[bb][2]) { bsel = bb; break; }
    }
    bool inbox = bsel >= 0;
    if (inbox) {
        const int ey = boxes.ext[bsel][1], ez = boxes.ext[bsel][2];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            int dx = (ix[m] - 1) - boxes.lo[bsel][0]; if (dx < 0) dx += g.XG;
            int dy = (jy[m] - 1) - boxes.lo[bsel][1]; if (dy < 0) dy += g.Y;
            int dz = (kz[m] - 1) - boxes.lo[bsel][2]; if (dz < 0) dz += g.Z;
            if (dx >= boxes.ext[bsel][0] || dy >= ey || dz >= ez) inbox = false;
            b.cell[12 * iEL + m] = dx * ey * ez;
            b.cell[12 * iEL + 4 + m] = dy * ez;
            b.cell[12 * iEL + 8 + m] = dz;
            const int lx = (ix[m] - 1) - g.xOffset;
            b.owned[4 * iEL + m] = (lx >= 0 && lx < g.X) ? 1 : 0;
        }
        b.boff[iEL] = boxes.off[bsel];
    }
    if (!inbox) {
        atomicOr(&ctl->err, 4);
        for (int m = 0; m < 12; m++) b.cell[12 * iEL + m] = 0;
        for (int m = 0; m < 4; m++) b.owned[4 * iEL + m] = 0;
        b.boff[iEL] = 0;
    }
}

__global__ void ibm_stencil_kernel(Geom g, IbmBody b, const __grid_constant__ IbmBoxes boxes, RootBC bc, IbmCtl *ctl)
{
    const int iEL = blockIdx.x * blockDim.x + threadIdx.x;
    if (iEL >= b.n) return;
    stencil_marker(g, b, boxes, bc, ctl, iEL);
}

// all bodies of a block in one launch: blockIdx.y = body
__global__ void ibm_stencil_all_kernel(Geom g, const IbmBody *bodies, const __grid_constant__ IbmBoxes boxes, RootBC bc, IbmCtl *ctl)
{
    const IbmBody b = bodies[blockIdx.y];
    const int iEL = blockIdx.x * blockDim.x + threadIdx.x;
    if (iEL >= b.n) return;
    stencil_marker(g, b, boxes, bc, ctl, iEL);
}

void launch_ibm_stencil_all(const Geom &g, const IbmBody *bodies_dev, int nbody, int max_n, const IbmBoxes &boxes, const int rootBC[6], IbmCtl *ctl, cudaStream_t s)
{
    if (nbody < 1) return;
    RootBC bc;
    for (int i = 0; i < 6; i++) bc.c[i] = rootBC[i];
    ibm_stencil_all_kernel<<<dim3((max_n + 127) / 128, nbody), 128, 0, s>>>(g, bodies_dev, boxes, bc, ctl);
    count_launch();
}

void launch_ibm_stencil(const Geom &g, const IbmBody &b, const IbmBoxes &boxes, const int rootBC[6], IbmCtl *ctl, cudaStream_t s)
{
    RootBC bc;
    for (int i = 0; i < 6; i++) bc.c[i] = rootBC[i];
    ibm_stencil_kernel<<<(b.n + 127) / 128, 128, 0, s>>>(g, b, boxes, bc, ctl);
    count_launch();
}

// calculate_macro_quantities_ (FluidDomain.f90:1136-1139) on the box cells + ResetVolumeForce_ (:1201-1203)
__device__ __forceinline__ void macro_box_cell(const Geom &g, const double *fA, double hF1, double hF2, double hF3, const IbmBoxes &boxes, long long i)
{
    int bb = 0;
    while (bb + 1 < boxes.n && i >= boxes.off[bb + 1]) bb++;
    const long long r = i - boxes.off[bb];
    const int ez = boxes.ext[bb][2], ey = boxes.ext[bb][1];
    const int dz = (int)(r % ez), dy = (int)((r / ez) % ey), dx = (int)(r / ((long long)ez * ey));
    int gx = boxes.lo[bb][0] + dx; if (gx >= g.XG) gx -= g.XG;
    int y = boxes.lo[bb][1] + dy; if (y >= g.Y) y -= g.Y;
    int z = boxes.lo[bb][2] + dz; if (z >= g.Z) z -= g.Z;
    const int x = gx - g.xOffset;
    double u1 = 0.0, u2 = 0.0, u3 = 0.0;
    if (x >= 0 && x < g.X) {
        const size_t base = (size_t)(x + 1) * g.plane + (size_t)y * g.Z + z;
        double f[Q], den;
#pragma unroll
        for (int q = 0; q < Q; q++) f[q] = fA[q * g.pstride + base];
        macro_from_f(f, hF1, hF2, hF3, den, u1, u2, u3);
    }
    boxes.u[i] = u1; boxes.u[boxes.ncell + i] = u2; boxes.u[2 * boxes.ncell + i] = u3;
    boxes.force[i] = 0.0; boxes.force[boxes.ncell + i] = 0.0; boxes.force[2 * boxes.ncell + i] = 0.0;
}

__global__ void ibm_macro_box_kernel(Geom g, const double *fA, double hF1, double hF2, double hF3, const __grid_constant__ IbmBoxes boxes)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= boxes.ncell) return;
    macro_box_cell(g, fA, hF1, hF2, hF3, boxes, i);
}

void launch_ibm_macro_box(const Geom &g, const double *fA, const double hF[3], const IbmBoxes &boxes, cudaStream_t s)
{
    if (boxes.ncell <= 0) return;
    ibm_macro_box_kernel<<<(unsigned)((boxes.ncell + 255) / 256), 256, 0, s>>>(g, fA, hF[0], hF[1], hF[2], boxes);
    count_launch();
}

// per-marker finish of PenaltyForce_'s first loop, Solidbody.f90:1016-1025
__device__ __forceinline__ void marker_force(const IbmBody &b, int iEL, double U1, double U2, double U3, double invh3)
{
    const double d1 = b.Evel[3 * iEL + 0] - U1, d2 = b.Evel[3 * iEL + 1] - U2, d3 = b.Evel[3 * iEL + 2] - U3;
    const double Ea = b.Ea[iEL];
    const double f1 = d1 * Ea, f2 = d2 * Ea, f3 = d3 * Ea;
    b.tol[iEL] = fabs(d1) + fabs(d2) + fabs(d3);
    b.Eforce[3 * iEL + 0] = b.Eforce[3 * iEL + 0] + f1;
    b.Eforce[3 * iEL + 1] = b.Eforce[3 * iEL + 1] + f2;
    b.Eforce[3 * iEL + 2] = b.Eforce[3 * iEL + 2] + f3;
    b.felt[3 * iEL + 0] = f1 * invh3;
    b.felt[3 * iEL + 1] = f2 * invh3;
    b.felt[3 * iEL + 2] = f3 * invh3;
}

// PenaltyForce_ interpolation, Solidbody.f90:1000-1015: one warp per marker.
// Lane 0 also finishes the marker (:1016-1025).
// the 64-node gather of one marker by one warp; the sums are valid in lane 0
__device__ __forceinline__ void gather_marker(const IbmBody &b, const IbmBoxes &boxes, int iEL, int lane, double &s1, double &s2, double &s3)
{
    const long long boff = b.boff[iEL];
    s1 = 0.0; s2 = 0.0; s3 = 0.0;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int pnt = lane + 32 * h;
        const int a = pnt >> 4, bb = (pnt >> 2) & 3, c = pnt & 3;
        if (b.owned[4 * iEL + a]) {
            const long long idx = boff + b.cell[12 * iEL + a] + b.cell[12 * iEL + 4 + bb] + b.cell[12 * iEL + 8 + c];
            const double rx = (double)b.Ew[12 * iEL + a], ry = (double)b.Ew[12 * iEL + 4 + bb], rz = (double)b.Ew[12 * iEL + 8 + c];
            s1 = s1 + boxes.u[idx] * rx * ry * rz;   // :1012
            s2 = s2 + boxes.u[boxes.ncell + idx] * rx * ry * rz;
            s3 = s3 + boxes.u[2 * boxes.ncell + idx] * rx * ry * rz;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, off);
        s2 += __shfl_down_sync(0xffffffffu, s2, off);
        s3 += __shfl_down_sync(0xffffffffu, s3, off);
    }
}

__global__ void ibm_gather_kernel(IbmBody b, const __grid_constant__ IbmBoxes boxes, double *partialU, const IbmCtl *ctl, int fused, double invh3)
{
    if (ctl->done) return;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= b.n) return;
    const int iEL = warp;
    double s1, s2, s3;
    gather_marker(b, boxes, iEL, lane, s1, s2, s3);
    if (lane == 0) marker_force(b, iEL, s1, s2, s3, invh3);
}

void launch_ibm_gather(const IbmBody &b, const IbmBoxes &boxes, double *partialU, const IbmCtl *ctl, int fused, double invh3, cudaStream_t s)
{
    const int threads = 128, warps_per_block = threads / 32;
    ibm_gather_kernel<<<(b.n + warps_per_block - 1) / warps_per_block, threads, 0, s>>>(b, boxes, partialU, ctl, fused, invh3);
    count_launch();
}

// PenaltyForce_ velocity correction, Solidbody.f90:1034-1048: uuu -= forceElemTemp*rx*ry*rz
__device__ __forceinline__ void scatter_marker(const IbmBody &b, const IbmBoxes &boxes, int iEL, int lane)
{
    const long long boff = b.boff[iEL];
    const double f1 = b.felt[3 * iEL + 0], f2 = b.felt[3 * iEL + 1], f3 = b.felt[3 * iEL + 2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int pnt = lane + 32 * h;
        const int a = pnt >> 4, bb = (pnt >> 2) & 3, c = pnt & 3;
        if (!b.owned[4 * iEL + a]) continue;
        const long long idx = boff + b.cell[12 * iEL + a] + b.cell[12 * iEL + 4 + bb] + b.cell[12 * iEL + 8 + c];
        const double rx = (double)b.Ew[12 * iEL + a], ry = (double)b.Ew[12 * iEL + 4 + bb], rz = (double)b.Ew[12 * iEL + 8 + c];
        atomicAdd(&boxes.u[idx], -(f1 * rx * ry * rz));
        atomicAdd(&boxes.u[boxes.ncell + idx], -(f2 * rx * ry * rz));
        atomicAdd(&boxes.u[2 * boxes.ncell + idx], -(f3 * rx * ry * rz));
    }
}

__global__ void ibm_scatter_kernel(IbmBody b, const __grid_constant__ IbmBoxes boxes, const IbmCtl *ctl)
{
    if (ctl->done) return;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= b.n) return;
    scatter_marker(b, boxes, warp, lane);
}

void launch_ibm_scatter(const IbmBody &b, const IbmBoxes &boxes, const IbmCtl *ctl, cudaStream_t s)
{
    const int threads = 128, warps_per_block = threads / 32;
    ibm_scatter_kernel<<<(b.n + warps_per_block - 1) / warps_per_block, threads, 0, s>>>(b, boxes, ctl);
    count_launch();
}

// loop control of calculate_interaction_force, Solidbody.f90:895-906: one block, deterministic sum
__global__ void ibm_check_kernel(const IbmBody *bodies, int nbody, double Uref, int ntol, double dtol, IbmCtl *ctl)
{
    if (ctl->done) return;
    __shared__ double sh[256];
    double dmax = 0.0, dsum = 0.0;
    for (int ib = 0; ib < nbody; ib++) {
        const IbmBody b = bodies[ib];
        double t = 0.0;
        for (int i = threadIdx.x; i < b.n; i += blockDim.x) t += b.tol[i];
        sh[threadIdx.x] = t;
        __syncthreads();
        for (int off = blockDim.x / 2; off > 0; off >>= 1) {
            if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
            __syncthreads();
        }
        dmax = dmax + sh[0];          // :901
        dsum = dsum + (double)b.n;    // :902
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (!isfinite(dmax)) atomicOr(&ctl->err, 2);   // :1028-1031
        dmax = dmax / (dsum * Uref);                   // :904
        const int iter = ctl->iter + 1;                // :905
        ctl->iter = iter;
        ctl->dmax = dmax;
        ctl->done = !(iter < ntol && dmax > dtol);     // :895
    }
}

void launch_ibm_check(const IbmBody *bodies_dev, int nbody, double Uref, int ntol, double dtol, IbmCtl *ctl, cudaStream_t s)
{
    ibm_check_kernel<<<1, 256, 0, s>>>(bodies_dev, nbody, Uref, ntol, dtol, ctl);
    count_launch();
}

// Loop control across ranks (slab runs in which every body is iterated only by the ranks whose planes its stencils touch):
// each rank sums |dU| over the bodies it LEADS (fixed-shape tree) together with their marker count, the two numbers are
// all-reduced, and every rank takes the same decision from the totals (Solidbody.f90:901-906).
__global__ void __launch_bounds__(1024) ibm_tol_sum_kernel(const IbmBody *bodies, const int *lead, int nbody, const IbmCtl *ctl, double *out2)
{
    __shared__ double sh[1024];
    double t = 0.0, cnt = 0.0;
    if (!ctl->done)
        for (int ib = 0; ib < nbody; ib++) {
            if (!lead[ib]) continue;
            const IbmBody b = bodies[ib];
            for (int i = threadIdx.x; i < b.n; i += blockDim.x) t += b.tol[i];
            cnt = cnt + (double)b.n;          // :902
        }
    sh[threadIdx.x] = t;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = sh[0]; out2[1] = cnt; }
}

__global__ void ibm_decide_kernel(const double *in2, double Uref, int ntol, double dtol, IbmCtl *ctl)
{
    if (ctl->done) return;
    double dmax = in2[0];
    if (!isfinite(dmax)) atomicOr(&ctl->err, 2);   // :1028-1031
    dmax = dmax / (in2[1] * Uref);                 // :904
    const int iter = ctl->iter + 1;                // :905
    ctl->iter = iter;
    ctl->dmax = dmax;
    ctl->done = !(iter < ntol && dmax > dtol);     // :895
}

void launch_ibm_tol_sum(const IbmBody *bodies_dev, const int *lead_dev, int nbody, const IbmCtl *ctl, double *out2, cudaStream_t s)
{
    ibm_tol_sum_kernel<<<1, 1024, 0, s>>>(bodies_dev, lead_dev, nbody, ctl, out2);
    count_launch();
}

void launch_ibm_decide(const double *in2, double Uref, int ntol, double dtol, IbmCtl *ctl, cudaStream_t s)
{
    ibm_decide_kernel<<<1, 1, 0, s>>>(in2, Uref, ntol, dtol, ctl);
    count_launch();
}

// Eulerian half of FluidVolumeForce_, Solidbody.f90:968-976: force += -(v_Eforce*invh3)*rx*ry*rz
__device__ __forceinline__ void spread_marker(const IbmBody &b, const IbmBoxes &boxes, double invh3, int iEL, int lane)
{
    const long long boff = b.boff[iEL];
    const double f1 = b.Eforce[3 * iEL + 0] * invh3, f2 = b.Eforce[3 * iEL + 1] * invh3, f3 = b.Eforce[3 * iEL + 2] * invh3;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int pnt = lane + 32 * h;
        const int a = pnt >> 4, bb = (pnt >> 2) & 3, c = pnt & 3;
        if (!b.owned[4 * iEL + a]) continue;
        const long long idx = boff + b.cell[12 * iEL + a] + b.cell[12 * iEL + 4 + bb] + b.cell[12 * iEL + 8 + c];
        const double rx = (double)b.Ew[12 * iEL + a], ry = (double)b.Ew[12 * iEL + 4 + bb], rz = (double)b.Ew[12 * iEL + 8 + c];
        atomicAdd(&boxes.force[idx], -f1 * rx * ry * rz);
        atomicAdd(&boxes.force[boxes.ncell + idx], -f2 * rx * ry * rz);
        atomicAdd(&boxes.force[2 * boxes.ncell + idx], -f3 * rx * ry * rz);
    }
}

__global__ void ibm_spread_kernel(IbmBody b, const __grid_constant__ IbmBoxes boxes, double invh3)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= b.n) return;
    spread_marker(b, boxes, invh3, warp, lane);
}

void launch_ibm_spread(const IbmBody &b, const IbmBoxes &boxes, double invh3, cudaStream_t s)
{
    const int threads = 128, warps_per_block = threads / 32;
    ibm_spread_kernel<<<(b.n + warps_per_block - 1) / warps_per_block, threads, 0, s>>>(b, boxes, invh3);
    count_launch();
}


// ---- ordered (bit-reproducible) interpolation and spreading ---------------------------------------------------------
// The reference interpolates with three nested serial loops (x outer, z inner; Solidbody.f90:1009-1015) and spreads
// with a serial loop over the markers (:1034-1048, :938-978).  Floating-point addition does not associate, so a warp
// shuffle tree or fp64 atomics give the same numbers only to round-off -- and inside a closed fluid-structure loop the
// beam solver's ill-conditioned Newmark operator amplifies that round-off.  These variants keep the reference's
// summation ORDER: one thread walks a marker's 64 nodes in loop order, and spreading is turned into a per-cell gather
// over IbmCsr, whose entries are sorted in exactly the order the serial loops reach the cell.

// PenaltyForce_ interpolation, ordered.  A block of 256 threads takes MARKERS_PER_BLOCK = 16 markers at a time: 16 lanes
// per marker form the 64 products of its nodes in parallel -- lane l those of (a, b, c) = (h, l/4, l%4), h = 0..3, each as
// the reference forms it, ((u*rx)*ry)*rz -- and park them in shared memory; then one thread per marker adds them in the
// order of the loops at :1009-1015 (x outer, z inner) and finishes the marker (:1016-1025).  The 64-term chain of additions
// is the only serial part and runs out of shared memory.
constexpr int MARKER_LANES = 16;
constexpr int MARKERS_PER_BLOCK = 16;   // 256 threads
struct GatherSmem { double p[3][64][MARKERS_PER_BLOCK]; };   // [component][node][marker]: conflict-free both ways

// m0: first marker of this block's batch.  Must be called by all 256 threads of the block.  Returns |dU| of the marker
// finished by this thread (threads 0..15), else 0.
// Every load of a batch is issued before anything waits for one: the stencil offsets, weights and ownership flags of a marker
// come first, then all twelve velocity loads of a lane; the finishing threads fetch their marker's velocity, area and force
// sum before the products are formed.  (With the loads inside the `owned` branches each of the four x-planes cost its own
// three dependent round trips to L2 / HBM -- beside a running collide-stream update the box fields are mostly NOT in L2 --
// and the interpolation of 32 768 markers took 160 us per sweep whatever the grid.)  A plane that is not owned contributes
// nothing to the sum, exactly as the skipped iteration of the reference loop: its products are replaced by +0, and
// s + (+0) = s for every s this chain can hold (it starts from +0).
__device__ __forceinline__ double gather_batch_ordered(const IbmBody &b, const IbmBoxes &boxes, int m0, GatherSmem &sm, double *partialU, int fused, double invh3)
{
    const int mi = threadIdx.x / MARKER_LANES, gl = threadIdx.x & (MARKER_LANES - 1);
    const int iEL = m0 + mi;
    const bool fin = threadIdx.x < MARKERS_PER_BLOCK && m0 + (int)threadIdx.x < b.n;
    const int m = m0 + threadIdx.x;
    double ev1 = 0.0, ev2 = 0.0, ev3 = 0.0, ea = 0.0, ef1 = 0.0, ef2 = 0.0, ef3 = 0.0;
    unsigned int own_m = 0;
    if (fin) {
        own_m = *(const unsigned int *)(b.owned + 4 * m);
        if (fused) {
            ev1 = b.Evel[3 * m + 0]; ev2 = b.Evel[3 * m + 1]; ev3 = b.Evel[3 * m + 2];
            ea = b.Ea[m];
            ef1 = b.Eforce[3 * m + 0]; ef2 = b.Eforce[3 * m + 1]; ef3 = b.Eforce[3 * m + 2];
        }
    }
    if (iEL < b.n) {
        const int bb = gl >> 2, c = gl & 3;
        const int *cl = b.cell + 12 * iEL;
        const float *ew = b.Ew + 12 * iEL;
        const unsigned int own = *(const unsigned int *)(b.owned + 4 * iEL);
        const long long base = b.boff[iEL] + cl[4 + bb] + cl[8 + c];
        const double ry = (double)ew[4 + bb], rz = (double)ew[8 + c];
        const double *u1 = boxes.u, *u2 = boxes.u + boxes.ncell, *u3 = boxes.u + 2 * boxes.ncell;
        long long idx[4];
        double rx[4], v1[4], v2[4], v3[4];
#pragma unroll
        for (int a = 0; a < 4; a++) { idx[a] = base + cl[a]; rx[a] = (double)ew[a]; }   // offsets of a plane that is not owned are valid too (0 at worst)
#pragma unroll
        for (int a = 0; a < 4; a++) { v1[a] = u1[idx[a]]; v2[a] = u2[idx[a]]; v3[a] = u3[idx[a]]; }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const bool o = ((own >> (8 * a)) & 0xffu) != 0;
            const double p1 = v1[a] * rx[a] * ry * rz;   // :1012
            const double p2 = v2[a] * rx[a] * ry * rz;
            const double p3 = v3[a] * rx[a] * ry * rz;
            sm.p[0][16 * a + gl][mi] = o ? p1 : 0.0; sm.p[1][16 * a + gl][mi] = o ? p2 : 0.0; sm.p[2][16 * a + gl][mi] = o ? p3 : 0.0;
        }
    }
    __syncthreads();
    double tol = 0.0;
    if (fin) {
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            if (((own_m >> (8 * a)) & 0xffu) == 0) continue;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                s1 = s1 + sm.p[0][16 * a + k][threadIdx.x];
                s2 = s2 + sm.p[1][16 * a + k][threadIdx.x];
                s3 = s3 + sm.p[2][16 * a + k][threadIdx.x];
            }
        }
        if (fused) {   // marker_force (:1016-1025) on the values fetched above
            const double d1 = ev1 - s1, d2 = ev2 - s2, d3 = ev3 - s3;
            const double f1 = d1 * ea, f2 = d2 * ea, f3 = d3 * ea;
            tol = fabs(d1) + fabs(d2) + fabs(d3);
            b.tol[m] = tol;
            b.Eforce[3 * m + 0] = ef1 + f1;
            b.Eforce[3 * m + 1] = ef2 + f2;
            b.Eforce[3 * m + 2] = ef3 + f3;
            b.felt[3 * m + 0] = f1 * invh3;
            b.felt[3 * m + 1] = f2 * invh3;
            b.felt[3 * m + 2] = f3 * invh3;
        } else { partialU[3 * m + 0] = s1; partialU[3 * m + 1] = s2; partialU[3 * m + 2] = s3; }
    }
    __syncthreads();
    return tol;
}

__global__ void __launch_bounds__(256) ibm_gather_ordered_kernel(IbmBody b, const __grid_constant__ IbmBoxes boxes, double *partialU, const IbmCtl *ctl, int fused, double invh3)
{
    if (ctl->done) return;
    __shared__ GatherSmem sm;
    gather_batch_ordered(b, boxes, blockIdx.x * MARKERS_PER_BLOCK, sm, partialU, fused, invh3);
}

void launch_ibm_gather_ordered(const IbmBody &b, const IbmBoxes &boxes, double *partialU, const IbmCtl *ctl, int fused, double invh3, cudaStream_t s)
{
    ibm_gather_ordered_kernel<<<(b.n + MARKERS_PER_BLOCK - 1) / MARKERS_PER_BLOCK, 256, 0, s>>>(b, boxes, partialU, ctl, fused, invh3);
    count_launch();
}

__device__ __forceinline__ unsigned long long csr_key(int body, int marker, int node)
{
    return ((unsigned long long)body << 40) | ((unsigned long long)(unsigned int)marker << 8) | (unsigned long long)node;
}

// velocity correction of one box cell by one thread (:1034-1048): the entries of bodies whose phase is `ph` (or, with
// phase_of_body null, of body `only_body`), in (body, marker, node) order.  One thread per cell keeps tens of thousands of
// cells in flight, which hides the dependent loads of an entry (key -> marker -> weights, force) better than giving a
// cell several lanes does: measured 5 us against 11 us per phase at 55k cells / 524k entries.
__device__ __forceinline__ void scatter_cell(const IbmBody *bodies, const IbmBoxes &boxes, const IbmCsr &csr, long long c, const int *phase_of_body,
                                             int ph, int only_body)
{
    const int beg = csr.off[c], end = csr.off[c + 1];
    if (beg == end) return;
    double u1 = boxes.u[c], u2 = boxes.u[boxes.ncell + c], u3 = boxes.u[2 * boxes.ncell + c];
    bool any = false;
    // four entries at a time: all weights and forces are fetched before the ordered subtraction, so the dependent loads of
    // the entries overlap instead of queueing behind one another; the keys of the NEXT four are requested before those loads
    // are waited for, so a group costs one round trip to memory, not two
    unsigned long long key[4];
#pragma unroll
    for (int j = 0; j < 4; j++) key[j] = beg + j < end ? csr.entry[beg + j] : ~0ull;
    for (int e0 = beg; e0 < end; e0 += 4) {
        unsigned long long nkey[4];
#pragma unroll
        for (int j = 0; j < 4; j++) nkey[j] = e0 + 4 + j < end ? csr.entry[e0 + 4 + j] : ~0ull;
        double q1[4], q2[4], q3[4];
        bool act[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int body = (int)(key[j] >> 40);
            act[j] = e0 + j < end && (phase_of_body ? phase_of_body[body] == ph : body == only_body);
            q1[j] = 0.0; q2[j] = 0.0; q3[j] = 0.0;
            if (act[j]) {
                const int m = (int)((key[j] >> 8) & 0xffffffffull), node = (int)(key[j] & 63ull);
                const IbmBody &b = bodies[body];
                const double rx = (double)b.Ew[12 * m + (node >> 4)], ry = (double)b.Ew[12 * m + 4 + ((node >> 2) & 3)], rz = (double)b.Ew[12 * m + 8 + (node & 3)];
                q1[j] = b.felt[3 * m + 0] * rx * ry * rz;   // :1044
                q2[j] = b.felt[3 * m + 1] * rx * ry * rz;
                q3[j] = b.felt[3 * m + 2] * rx * ry * rz;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (act[j]) { u1 = u1 - q1[j]; u2 = u2 - q2[j]; u3 = u3 - q3[j]; any = true; }
#pragma unroll
        for (int j = 0; j < 4; j++) key[j] = nkey[j];
    }
    if (any) { boxes.u[c] = u1; boxes.u[boxes.ncell + c] = u2; boxes.u[2 * boxes.ncell + c] = u3; }
}

// Eulerian half of FluidVolumeForce_ for one box cell (:968-976), all bodies in order
__device__ __forceinline__ void spread_cell(const IbmBody *bodies, const IbmBoxes &boxes, const IbmCsr &csr, long long c, double invh3)
{
    const int beg = csr.off[c], end = csr.off[c + 1];
    if (beg == end) return;
    double f1 = boxes.force[c], f2 = boxes.force[boxes.ncell + c], f3 = boxes.force[2 * boxes.ncell + c];
    unsigned long long key[4];
#pragma unroll
    for (int j = 0; j < 4; j++) key[j] = beg + j < end ? csr.entry[beg + j] : ~0ull;
    for (int e0 = beg; e0 < end; e0 += 4) {
        unsigned long long nkey[4];   // requested before this group's weights and forces are waited for (see scatter_cell)
#pragma unroll
        for (int j = 0; j < 4; j++) nkey[j] = e0 + 4 + j < end ? csr.entry[e0 + 4 + j] : ~0ull;
        double q1[4], q2[4], q3[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            q1[j] = 0.0; q2[j] = 0.0; q3[j] = 0.0;
            if (e0 + j < end) {
                const int body = (int)(key[j] >> 40), m = (int)((key[j] >> 8) & 0xffffffffull), node = (int)(key[j] & 63ull);
                const IbmBody &b = bodies[body];
                const double rx = (double)b.Ew[12 * m + (node >> 4)], ry = (double)b.Ew[12 * m + 4 + ((node >> 2) & 3)], rz = (double)b.Ew[12 * m + 8 + (node & 3)];
                const double e1 = b.Eforce[3 * m + 0] * invh3, e2 = b.Eforce[3 * m + 1] * invh3, e3 = b.Eforce[3 * m + 2] * invh3;   // :968
                q1[j] = -e1 * rx * ry * rz;   // :972
                q2[j] = -e2 * rx * ry * rz;
                q3[j] = -e3 * rx * ry * rz;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (e0 + j < end) { f1 = f1 + q1[j]; f2 = f2 + q2[j]; f3 = f3 + q3[j]; }   // :974
#pragma unroll
        for (int j = 0; j < 4; j++) key[j] = nkey[j];
    }
    boxes.force[c] = f1; boxes.force[boxes.ncell + c] = f2; boxes.force[2 * boxes.ncell + c] = f3;
}

__global__ void ibm_scatter_cells_kernel(const IbmBody *bodies, int body, const __grid_constant__ IbmBoxes boxes, IbmCsr csr, const IbmCtl *ctl)
{
    if (ctl->done) return;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < boxes.ncell) scatter_cell(bodies, boxes, csr, c, nullptr, 0, body);
}

void launch_ibm_scatter_ordered(const IbmBody *bodies_dev, int body, const IbmBoxes &boxes, const IbmCsr &csr, const IbmCtl *ctl, cudaStream_t s)
{
    ibm_scatter_cells_kernel<<<(unsigned)((boxes.ncell + 127) / 128), 128, 0, s>>>(bodies_dev, body, boxes, csr, ctl);
    count_launch();
}

__global__ void ibm_spread_cells_kernel(const IbmBody *bodies, const __grid_constant__ IbmBoxes boxes, IbmCsr csr, double invh3)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < boxes.ncell) spread_cell(bodies, boxes, csr, c, invh3);
}

void launch_ibm_spread_ordered(const IbmBody *bodies_dev, const IbmBoxes &boxes, const IbmCsr &csr, double invh3, cudaStream_t s)
{
    ibm_spread_cells_kernel<<<(unsigned)((boxes.ncell + 127) / 128), 128, 0, s>>>(bodies_dev, boxes, csr, invh3);
    count_launch();
}

// -- IbmCsr build: one thread per (marker, node)
template <bool FILL>
__global__ void ibm_csr_nodes_kernel(const IbmBody *bodies, IbmCsr csr)
{
    const int body = blockIdx.y;
    const IbmBody b = bodies[body];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)b.n * 64) return;
    const int iEL = (int)(t >> 6), node = (int)(t & 63);
    const int a = node >> 4, bb = (node >> 2) & 3, c = node & 3;
    if (!b.owned[4 * iEL + a]) return;
    const long long idx = b.boff[iEL] + b.cell[12 * iEL + a] + b.cell[12 * iEL + 4 + bb] + b.cell[12 * iEL + 8 + c];
    if (!FILL) atomicAdd(&csr.count[idx], 1);
    else {
        const int slot = atomicSub(&csr.count[idx], 1) - 1;   // counts run back down to zero
        csr.entry[csr.off[idx] + slot] = csr_key(body, iEL, node);
    }
}

__global__ void ibm_csr_sort_kernel(IbmCsr csr, long long ncell)
{
    // A warp looks at 32 consecutive cells (one coalesced read of their offsets) and then sorts, one after the other, only those
    // that hold more than one entry -- five in six box cells hold none: rank sort of up to 32 keys (all distinct) across the
    // lanes, longer lists by insertion.  (One warp per cell spent most of its 109 us on empty cells.)
    const long long c0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) << 5;
    if (c0 >= ncell) return;
    const int lane = threadIdx.x & 31;
    const long long cl = c0 + lane;
    const int mybeg = cl < ncell ? csr.off[cl] : 0, myend = cl < ncell ? csr.off[cl + 1] : 0;
    unsigned int todo = __ballot_sync(0xffffffffu, myend - mybeg > 1);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int beg = __shfl_sync(0xffffffffu, mybeg, src), end = __shfl_sync(0xffffffffu, myend, src), cnt = end - beg;
        if (cnt <= 32) {
            const unsigned long long k = lane < cnt ? csr.entry[beg + lane] : ~0ull;
            int rank = 0;
            for (int j = 0; j < cnt; j++) rank += __shfl_sync(0xffffffffu, k, j) < k ? 1 : 0;
            __syncwarp();
            if (lane < cnt) csr.entry[beg + rank] = k;
        } else if (lane == 0) {
            for (int i = beg + 1; i < end; i++) {
                const unsigned long long k = csr.entry[i];
                int j = i - 1;
                while (j >= beg && csr.entry[j] > k) { csr.entry[j + 1] = csr.entry[j]; j--; }
                csr.entry[j + 1] = k;
            }
        }
        __syncwarp();
    }
}

size_t ibm_csr_scan_bytes(long long ncell)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int *)nullptr, (int *)nullptr, (int)(ncell + 1));
    return bytes;
}

int launch_ibm_csr_build(const IbmBody *bodies_dev, int nbody, int max_n, const IbmBoxes &boxes, const IbmCsr &csr, void *scan_tmp, size_t scan_bytes, cudaStream_t s)
{
    if (cudaMemsetAsync(csr.count, 0, sizeof(int) * (size_t)(boxes.ncell + 1), s) != cudaSuccess) return 1;
    const dim3 grid((unsigned)(((long long)max_n * 64 + 255) / 256), nbody);   // blockIdx.y = body
    ibm_csr_nodes_kernel<false><<<grid, 256, 0, s>>>(bodies_dev, csr);
    count_launch();
    if (cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, csr.count, csr.off, (int)(boxes.ncell + 1), s) != cudaSuccess) return 1;
    count_launch();
    ibm_csr_nodes_kernel<true><<<grid, 256, 0, s>>>(bodies_dev, csr);
    count_launch();
    ibm_csr_sort_kernel<<<(unsigned)((boxes.ncell + 127) / 128), 128, 0, s>>>(csr, boxes.ncell);   // 32 cells per warp
    count_launch();
    return 0;
}


// ---- the whole of calculate_interaction_force in ONE cooperative launch (single-rank blocks) -----------------------
// Phases are separated by grid-wide barriers instead of kernel boundaries: stencils + box macro | per iteration and per
// body group: gather+force | velocity correction | loop control | ... | force spreading.  Bodies whose stencil boxes
// are disjoint cannot see each other's correction, so each phase processes one body of every box group at once; bodies
// that share a box keep the reference's sequential (Gauss-Seidel) order, Solidbody.f90:898-903.
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        while (*((volatile unsigned int *)bar) < epoch) __nanosleep(40);   // back off: hundreds of pollers on one line delay the arrivals
        __threadfence();
    }
    __syncthreads();
}

// Loop-control exchange of one iteration (see IbmCtlExchange in kernels.h).  Called by every thread of the block; `publish` is
// true in exactly one block per rank.  Thread r < nranks stores this rank's numbers into rank r's mailbox, then waits for rank
// r's entry in the own mailbox.  The totals are formed in rank order, so every block of every rank gets the same two numbers.
// Returns false if a rank did not report within the time limit.
struct CtlShared { double tol[MAX_PEERS], cnt[MAX_PEERS]; int bad; };
__device__ __forceinline__ bool ctl_exchange(const IbmCtlExchange &xc, int it, double local_tol, bool publish, CtlShared &sh, double &tol_sum, double &cnt_sum)
{
    const unsigned long long seq = xc.seq_base + (unsigned long long)it;
    const int slot = (int)(seq % IBM_CTL_SLOTS);
    if (threadIdx.x == 0) sh.bad = 0;
    __syncthreads();
    if ((int)threadIdx.x < xc.nranks) {
        const int r = threadIdx.x;
        if (publish) {
            CtlSlot *dst = (CtlSlot *)xc.mailbox[r] + slot * xc.nranks + xc.rank;
            asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(&dst->tol), "d"(local_tol) : "memory");
            asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(&dst->cnt), "d"(xc.cnt_local) : "memory");
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&dst->seq), "l"(seq) : "memory");
        }
        const CtlSlot *src = (const CtlSlot *)xc.mailbox[xc.rank] + slot * xc.nranks + r;
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(&src->seq) : "memory");
            if (v == seq) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > xc.timeout_ns) { sh.bad = 1; __trap(); }   // a rank is lost: abort the launch (the blocks of a cooperative grid must not part ways)
            __nanosleep(100);
        }
        double a, c;
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(a) : "l"(&src->tol) : "memory");
        asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(c) : "l"(&src->cnt) : "memory");
        sh.tol[r] = a; sh.cnt[r] = c;
    }
    __syncthreads();
    double t = 0.0, n = 0.0;
    for (int r = 0; r < xc.nranks; r++) { t = t + sh.tol[r]; n = n + sh.cnt[r]; }
    tol_sum = t; cnt_sum = n;
    const bool ok = sh.bad == 0;
    __syncthreads();
    return ok;
}

// a rank whose slab no body touches: it reports zeros and follows the others' decision
__global__ void ibm_ctl_only_kernel(const __grid_constant__ IbmCtlExchange xc, int ntol, double dtol, double Uref, IbmCtl *ctl)
{
    __shared__ CtlShared sh;
    bool done = ctl->done != 0;
    for (int it = 0; it < ntol && !done; it++) {
        double dmax, cnt;
        const bool ok = ctl_exchange(xc, it, 0.0, true, sh, dmax, cnt);
        const bool bad = !isfinite(dmax);
        dmax = dmax / (cnt * Uref);
        done = !ok || !(it + 1 < ntol && dmax > dtol);
        if (threadIdx.x == 0) {
            if (bad) atomicOr(&ctl->err, 2);
            if (!ok) atomicOr(&ctl->err, 8);
            ctl->iter = it + 1; ctl->dmax = dmax; ctl->done = done ? 1 : 0;
        }
    }
}

void launch_ibm_ctl_only(const IbmCtlExchange &xc, int ntol, double dtol, double Uref, IbmCtl *ctl, cudaStream_t s)
{
    ibm_ctl_only_kernel<<<1, 32, 0, s>>>(xc, ntol, dtol, Uref, ctl);
    count_launch();
}

__device__ __forceinline__ void prof_stamp(const IbmLoopParams &p, int &k)
{
    if (p.prof && blockIdx.x == 0 && threadIdx.x == 0 && k < 63) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.prof[1 + k] = t;
        p.prof[0] = (unsigned long long)(k + 1);
    }
    k++;
}

// MINB = 4: 64 registers per thread, no spills.  (A 48-register build, MINB = 5 -- a 256-thread block then takes exactly the register
// space of one 96-register CTA of the collide kernel it runs beside -- was measured slower on every workload and is gone.)
template <int MINB>
__global__ void __launch_bounds__(256, MINB) ibm_loop_kernel(const __grid_constant__ IbmLoopParams p)
{
    unsigned int epoch = 0;
    int pk = 0;
    prof_stamp(p, pk);
    if (p.prof) {   // profiling only: ten empty grid barriers, so that the first interval / 10 is the cost of one
        for (int i = 0; i < 10; i++) grid_barrier(p.barrier, epoch);
        prof_stamp(p, pk);
    }
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = (gridDim.x * blockDim.x) >> 5;
    const long long gthread = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthread = (long long)gridDim.x * blockDim.x;
    RootBC bc;
    for (int i = 0; i < 6; i++) bc.c[i] = p.rootBC[i];
    // phase 0: UpdateElmtInterp_ for every marker (unless done before the launch, which the ordered mode needs for its
    // cell lists), calculate_macro_quantities + ResetVolumeForce on the box cells
    if (p.do_stencil)
        for (int ib = 0; ib < p.nbody; ib++) {
            const IbmBody b = p.bodies[ib];
            for (long long i = gthread; i < b.n; i += nthread) stencil_marker(p.g, b, p.boxes, bc, p.ctl, (int)i);
        }
    if (p.do_macro)
        for (long long i = gthread; i < p.boxes.ncell; i += nthread) macro_box_cell(p.g, p.fA, p.hF[0], p.hF[1], p.hF[2], p.boxes, i);
    grid_barrier(p.barrier, epoch);
    prof_stamp(p, pk);
    __shared__ double sh_tol[256];
    __shared__ CtlShared sh_ctl;
    __shared__ GatherSmem sh_gather;
    __shared__ IbmBody sh_bodies[MAX_IBM_PHASE_BODIES];   // the body table next to the SM: the cell loops dereference it per entry
    for (int i = threadIdx.x; i < p.nbody; i += blockDim.x) sh_bodies[i] = p.bodies[i];
    __syncthreads();
    bool done = ((volatile IbmCtl *)p.ctl)->done != 0;   // ntolLBM <= 0; nothing writes it before the second barrier
    for (int it = 0; it < p.ntol; it++) {
        if (p.ordered ? done : (((volatile IbmCtl *)p.ctl)->done != 0)) break;   // uniform over the grid
        double tol_block = 0.0;                        // ordered mode, thread 0: this block's share of the iteration's sum of |dU|
        double *tol_partial = p.tol_partial + (size_t)(it & 1) * gridDim.x;   // double-buffered: blocks may be one barrier apart
        for (int ph = 0; ph < p.nphase; ph++) {
            // PenaltyForce_ first loop (Solidbody.f90:1000-1027) for the ph-th body of every group
            double tol = 0.0;
            if (p.ordered) {
                for (int k = p.phase_start[ph]; k < p.phase_start[ph + 1]; k++) {
                    const IbmBody &b = sh_bodies[p.phase_body[k]];
                    const bool lead = p.lead[p.phase_body[k]] != 0;   // a body across a slab interface is reported by one rank only
                    for (int m0 = blockIdx.x * MARKERS_PER_BLOCK; m0 < b.n; m0 += gridDim.x * MARKERS_PER_BLOCK) {   // uniform over the block
                        const double t = gather_batch_ordered(b, p.boxes, m0, sh_gather, nullptr, 1, p.invh3_pen);
                        if (lead) tol += t;
                    }
                }
                sh_tol[threadIdx.x] = tol;             // fixed-shape tree: the same sum on every run
                __syncthreads();
                for (int off = 128; off > 0; off >>= 1) {
                    if ((int)threadIdx.x < off) sh_tol[threadIdx.x] += sh_tol[threadIdx.x + off];
                    __syncthreads();
                }
                if (threadIdx.x == 0) {
                    tol_block += sh_tol[0];
                    if (ph == p.nphase - 1) tol_partial[blockIdx.x] = tol_block;
                }
            } else {
                for (int k = p.phase_start[ph]; k < p.phase_start[ph + 1]; k++) {
                    const IbmBody b = p.bodies[p.phase_body[k]];
                    for (int m = gwarp; m < b.n; m += nwarp) {
                        double s1, s2, s3;
                        gather_marker(b, p.boxes, m, lane, s1, s2, s3);
                        if (lane == 0) { marker_force(b, m, s1, s2, s3, p.invh3_pen); tol += b.tol[m]; }
                    }
                }
                if (lane == 0) sh_tol[threadIdx.x >> 5] = tol;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double t = 0.0;
                    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += sh_tol[w];
                    if (t != 0.0 || !(t == t)) atomicAdd(&p.ctl->tol_acc, t);
                }
            }
            grid_barrier(p.barrier, epoch);
            prof_stamp(p, pk);
            // velocity correction (:1034-1048)
            if (p.ordered) {
                for (long long c = gthread; c < p.boxes.ncell; c += nthread) scatter_cell(sh_bodies, p.boxes, p.csr, c, p.phase_of_body, ph, 0);
            } else {
                for (int k = p.phase_start[ph]; k < p.phase_start[ph + 1]; k++) {
                    const IbmBody b = p.bodies[p.phase_body[k]];
                    for (int m = gwarp; m < b.n; m += nwarp) scatter_marker(b, p.boxes, m, lane);
                }
            }
            grid_barrier(p.barrier, epoch);
            prof_stamp(p, pk);
        }
        // loop control, :895-906
        if (p.ordered) {
            // every block forms the same fixed-shape sum of the per-block partials (written before the barrier that ended the
            // interpolation phase) and so reaches the same decision without a further barrier; block 0 records it for the host
            double t = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) t += ((volatile double *)tol_partial)[i];
            sh_tol[threadIdx.x] = t;
            __syncthreads();
            for (int off = 128; off > 0; off >>= 1) {
                if ((int)threadIdx.x < off) sh_tol[threadIdx.x] += sh_tol[threadIdx.x + off];
                __syncthreads();
            }
            double dmax = sh_tol[0], dsum = p.dsum;
            __syncthreads();
            bool heard = true;
            if (p.xc.nranks > 1) heard = ctl_exchange(p.xc, it, dmax, blockIdx.x == 0, sh_ctl, dmax, dsum);   // totals over the ranks of the slab run
            const bool bad = !isfinite(dmax);               // :1028-1031
            dmax = dmax / (dsum * p.Uref);
            done = !heard || !(it + 1 < p.ntol && dmax > p.dtol);
            if (gthread == 0) {
                IbmCtl *c = p.ctl;
                if (!heard) atomicOr(&c->err, 8);
                if (bad) atomicOr(&c->err, 2);
                c->iter = it + 1;
                c->dmax = dmax;
                c->done = done ? 1 : 0;
            }
        } else {
            if (gthread == 0) {
                IbmCtl *c = p.ctl;
                double dmax = c->tol_acc;
                if (!isfinite(dmax)) atomicOr(&c->err, 2);   // :1028-1031
                dmax = dmax / (p.dsum * p.Uref);
                c->iter = c->iter + 1;
                c->dmax = dmax;
                c->tol_acc = 0.0;
                c->done = !(c->iter < p.ntol && dmax > p.dtol);
                __threadfence();
            }
            grid_barrier(p.barrier, epoch);
        }
    }
    // FluidVolumeForce_, Eulerian half (:968-976)
    if (p.ordered) {
        for (long long c = gthread; c < p.boxes.ncell; c += nthread) spread_cell(sh_bodies, p.boxes, p.csr, c, p.invh3);
    } else {
        for (int ib = 0; ib < p.nbody; ib++) {
            const IbmBody b = p.bodies[ib];
            for (int m = gwarp; m < b.n; m += nwarp) spread_marker(b, p.boxes, p.invh3, m, lane);
        }
    }
    prof_stamp(p, pk);
}

int ibm_loop_max_blocks()
{
    static int max_blocks = 0;
    if (!max_blocks) {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ibm_loop_kernel<4>, 256, 0);
        max_blocks = sms * (per_sm > 0 ? per_sm : 1);
    }
    return max_blocks;
}

// blocks_per_sm > 0: the launch shares the SMs with a running collide-stream kernel (early IBM); one block per SM leaves that
// kernel three of its four CTA slots (its 125 registers per thread fill the register file with four)
int launch_ibm_loop(const IbmLoopParams &p, int max_markers, int blocks_per_sm, int blocks_total, cudaStream_t s)
{
    int max_blocks = ibm_loop_max_blocks();
    if (blocks_per_sm > 0) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms > 0 && sms * blocks_per_sm < max_blocks) max_blocks = sms * blocks_per_sm;
    }
    if (blocks_total > 0 && blocks_total < max_blocks) max_blocks = blocks_total;
    int want = (max_markers + 7) / 8;           // one warp per marker of the largest phase
    long long cells = (p.boxes.ncell + 255) / 256;
    if (p.ordered) want = (max_markers + MARKERS_PER_BLOCK - 1) / MARKERS_PER_BLOCK;   // 16 lanes per marker, one thread per box cell
    if (cells > want) want = (int)(cells > max_blocks ? max_blocks : cells);
    int blocks = want < 1 ? 1 : (want > max_blocks ? max_blocks : want);
    if (cudaMemsetAsync(p.barrier, 0, sizeof(unsigned int), s) != cudaSuccess) return 1;   // the barrier counts up from zero in every launch
    void *args[] = {(void *)&p};
    cudaError_t e = cudaLaunchCooperativeKernel((void *)ibm_loop_kernel<4>, dim3(blocks), dim3(256), args, 0, s);
    if (e != cudaSuccess) return 1;
    count_launch();
    return 0;
}

}  // namespace fsilbm
