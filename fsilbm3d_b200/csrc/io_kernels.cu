// io_kernels.cu -- device side of the reference's output / diagnostics procedures, so that den and uuu never have
// to exist as full host arrays between outputs:
//   write_flow_ staging (FluidDomain.f90:1650-1699)       -> flow_window_kernel   (real(4) p,u,v,w [+ averages])
//   calculate_turbulent_statistic_ (:1147-1172)            -> turbulent_statistic_kernel
//   write_fluid_flux (:2019-2056)                          -> fluid_flux_kernel
//   grid_value_interpolation (Util.f90:123-166) for probes -> probe_kernel
// All of them derive den/uuu from the current populations exactly as calculate_macro_quantities_ (:1136-1139) does
// at main.f90:107 (i.e. WITHOUT the IBM velocity correction: reference quirk).
#include "kernels.h"

namespace fsilbm {

__device__ __forceinline__ void macro_at(const Geom &g, const double *f, const double (&hF)[3], int x, int y, int z, double &den, double &u1,
                                         double &u2, double &u3)
{
    const size_t base = (size_t)(x + 1) * g.plane + (size_t)y * g.Z + z;
    double fl[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) fl[q] = f[q * g.pstride + base];
    macro_from_f(fl, hF[0], hF[1], hF[2], den, u1, u2, u3);
}

// OUTtmp(z,y,x,0:3) [and 4:12] of write_flow_ over the output window [off, dim-off) of the local slab
__global__ void flow_window_kernel(const __grid_constant__ FlowWindowParams p)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, x = blockIdx.z;
    if (z >= p.nz) return;
    const int gx = p.x0 + x, gy = p.off + y, gz = p.off + z;
    const size_t n = (size_t)p.nx * p.ny * p.nz, c = ((size_t)x * p.ny + y) * p.nz + z;
    if (p.outputtype != 2) {
        double den, u1, u2, u3;
        macro_at(p.g, p.f, p.hF, gx, gy, gz, den, u1, u2, u3);
        p.out[c] = (float)((1.0 / 3.0) * (den - p.denIn));   // Cs2*(den-denIn), :1658
        p.out[n + c] = (float)(u1 * p.invUref);
        p.out[2 * n + c] = (float)(u2 * p.invUref);
        p.out[3 * n + c] = (float)(u3 * p.invUref);
    }
    if (p.outputtype >= 2) {
        const size_t nc = (size_t)p.g.X * p.g.plane, cc = (size_t)gx * p.g.plane + (size_t)gy * p.g.Z + gz;
#pragma unroll
        for (int k = 0; k < 3; k++) p.out[(4 + k) * n + c] = (float)(p.uuu_ave[k * nc + cc] * p.invUref);         // :1672-1674
#pragma unroll
        for (int k = 3; k < 9; k++) p.out[(4 + k) * n + c] = (float)(p.uuu_ave[k * nc + cc] * p.invUrefs);        // :1683-1696
    }
}
void launch_flow_window(const FlowWindowParams &p, cudaStream_t s)
{
    if (p.nx <= 0 || p.ny <= 0 || p.nz <= 0) return;
    dim3 block(128), grid((p.nz + 127) / 128, p.ny, p.nx);
    flow_window_kernel<<<grid, block, 0, s>>>(p);
    count_launch();
}

// calculate_turbulent_statistic_, FluidDomain.f90:1153-1168; the later lines use the already-updated means, as the
// reference's statement order does
__global__ void turbulent_statistic_kernel(Geom g, const double *f, double hF1, double hF2, double hF3, double *ave, double invStep)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, x = blockIdx.z;
    if (z >= g.Z) return;
    const double hF[3] = {hF1, hF2, hF3};
    double den, u[3];
    macro_at(g, f, hF, x, y, z, den, u[0], u[1], u[2]);
    const size_t n = (size_t)g.X * g.plane, c = (size_t)x * g.plane + (size_t)y * g.Z + z;
    const double w = 1.0 - invStep;
    double m[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { m[k] = ave[k * n + c] * w + invStep * u[k]; ave[k * n + c] = m[k]; }
#pragma unroll
    for (int k = 0; k < 3; k++) ave[(3 + k) * n + c] = ave[(3 + k) * n + c] * w + invStep * (u[k] - m[k]) * (u[k] - m[k]);
    ave[6 * n + c] = ave[6 * n + c] * w + invStep * (u[0] - m[0]) * (u[1] - m[1]);
    ave[7 * n + c] = ave[7 * n + c] * w + invStep * (u[0] - m[0]) * (u[2] - m[2]);
    ave[8 * n + c] = ave[8 * n + c] * w + invStep * (u[1] - m[1]) * (u[2] - m[2]);
}
void launch_turbulent_statistic(const Geom &g, const double *f, const double hF[3], double *ave, double invStep, cudaStream_t s)
{
    dim3 block(128), grid((g.Z + 127) / 128, g.Y, g.X);
    turbulent_statistic_kernel<<<grid, block, 0, s>>>(g, f, hF[0], hF[1], hF[2], ave, invStep);
    count_launch();
}

// write_fluid_flux, FluidDomain.f90:2019-2046: sum over the plane local x = xl of uuu(1)*den*dh*dh*wy*wz (one block per plane)
__global__ void fluid_flux_kernel(Geom g, const double *f, double hF1, double hF2, double hF3, int xl0, int xl1, int xl2, double *out3)
{
    const int xl = blockIdx.x == 0 ? xl0 : (blockIdx.x == 1 ? xl1 : xl2);
    __shared__ double sh[256];
    double s = 0.0;
    if (xl >= 0) {
        const double hF[3] = {hF1, hF2, hF3};
        for (size_t i = threadIdx.x; i < g.plane; i += blockDim.x) {
            const int j = (int)(i / g.Z), k = (int)(i % g.Z);
            const double wz = (k == 0 || k == g.Z - 1) ? 0.5 : 1.0, wy = (j == 0 || j == g.Y - 1) ? 0.5 : 1.0;
            double den, u1, u2, u3;
            macro_at(g, f, hF, xl, j, k, den, u1, u2, u3);
            s = s + u1 * den * g.dh * g.dh * wy * wz;
        }
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = blockDim.x / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out3[blockIdx.x] = sh[0];
}
void launch_fluid_flux(const Geom &g, const double *f, const double hF[3], const int xl[3], double *out3, cudaStream_t s)
{
    fluid_flux_kernel<<<3, 256, 0, s>>>(g, f, hF[0], hF[1], hF[2], xl[0], xl[1], xl[2], out3);
    count_launch();
}

// grid_value_interpolation, Util.f90:123-157, for the three velocity components of each probe
__global__ void probe_kernel(Geom g, const double *f, double hF1, double hF2, double hF3, int n, const double *coords, double *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double hF[3] = {hF1, hF2, hF3};
    const double dh = g.dh, X = coords[3 * i], Y = coords[3 * i + 1], Zc = coords[3 * i + 2];
    const int x1 = (int)floor((X - g.xmin) / dh + 1), y1 = (int)floor((Y - g.ymin) / dh + 1), z1 = (int)floor((Zc - g.zmin) / dh + 1);
    const double dx1 = X - (g.xmin + dh * (x1 - 1)), dy1 = Y - (g.ymin + dh * (y1 - 1)), dz1 = Zc - (g.zmin + dh * (z1 - 1));
    const double dx2 = dh - dx1, dy2 = dh - dy1, dz2 = dh - dz1;
    const double coffe = 1.0 / (dh * dh * dh);
    const double c[8] = {dx2 * dy2 * dz2 * coffe, dx2 * dy2 * dz1 * coffe, dx2 * dy1 * dz2 * coffe, dx1 * dy2 * dz2 * coffe,
                         dx2 * dy1 * dz1 * coffe, dx1 * dy2 * dz1 * coffe, dx1 * dy1 * dz2 * coffe, dx1 * dy1 * dz1 * coffe};
    // corner order of :154-155: (z1,y1,x1) (z2,y1,x1) (z1,y2,x1) (z1,y1,x2) (z2,y2,x1) (z2,y1,x2) (z1,y2,x2) (z2,y2,x2)
    const int ox[8] = {0, 0, 0, 1, 0, 1, 1, 1}, oy[8] = {0, 0, 1, 0, 1, 0, 1, 1}, oz[8] = {0, 1, 0, 0, 1, 1, 0, 1};
    double v[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < 8; k++) {
        // clamp only to stay inside the allocation when a probe sits exactly on the last plane (its weight is then zero)
        const int xx = min(x1 - 1 + ox[k], g.XG - 1) - g.xOffset, yy = min(y1 - 1 + oy[k], g.Y - 1), zz = min(z1 - 1 + oz[k], g.Z - 1);
        double den, u[3] = {0.0, 0.0, 0.0};
        if (xx >= 0 && xx < g.X) macro_at(g, f, hF, xx, yy, zz, den, u[0], u[1], u[2]);
        for (int j = 0; j < 3; j++) v[j] = k == 0 ? c[0] * u[j] : v[j] + c[k] * u[j];
    }
    for (int j = 0; j < 3; j++) out[3 * i + j] = v[j];
}
void launch_probe(const Geom &g, const double *f, const double hF[3], int n, const double *coords, double *out, cudaStream_t s)
{
    if (n <= 0) return;
    probe_kernel<<<(n + 63) / 64, 64, 0, s>>>(g, f, hF[0], hF[1], hF[2], n, coords, out);
    count_launch();
}

}  // namespace fsilbm
