// d3q19.cuh -- D3Q19 lattice constants and per-cell device arithmetic (fp64).
//
// The arithmetic follows the reference expression by expression, in the reference's
// left-to-right evaluation order, with terms whose lattice coefficient is zero dropped
// (adding an exact zero never changes a partial sum).  The library is compiled with
// -fmad=false so no multiply-add is contracted: an x86-64 gfortran build of the reference
// (Makefile:19,24, no -march) has no FMA either.  Citations: /root/reference/src.
#pragma once
#include <cuda_runtime.h>

namespace fsilbm {

constexpr int Q = 19;

// ConstParams.f90:11-20
__host__ __device__ constexpr int EX(int q) { constexpr int e[Q] = {0, 1, -1, 0, 0, 0, 0, 1, -1, 1, -1, 1, -1, 1, -1, 0, 0, 0, 0}; return e[q]; }
__host__ __device__ constexpr int EY(int q) { constexpr int e[Q] = {0, 0, 0, 1, -1, 0, 0, 1, 1, -1, -1, 0, 0, 0, 0, 1, -1, 1, -1}; return e[q]; }
__host__ __device__ constexpr int EZ(int q) { constexpr int e[Q] = {0, 0, 0, 0, 0, 1, -1, 0, 0, 0, 0, 1, 1, -1, -1, 1, 1, -1, -1}; return e[q]; }
__host__ __device__ constexpr int OPPO(int q) { constexpr int o[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15}; return o[q]; }
// ConstParams.f90:22-25
__host__ __device__ constexpr double WT(int q) { return q == 0 ? 1.0 / 3.0 : (q <= 6 ? 1.0 / 18.0 : 1.0 / 36.0); }
// ConstParams.f90:19-20
__host__ __device__ constexpr int POSDIR(int p) { constexpr int d[9] = {1, 3, 5, 7, 8, 11, 12, 15, 16}; return d[p]; }
__host__ __device__ constexpr int NEGDIR(int p) { constexpr int d[9] = {2, 4, 6, 10, 9, 14, 13, 18, 17}; return d[p]; }

// ConstParams.f90:31-34
enum BC : int {
    BCEq_DirecletU = 101, BCnEq_DirecletU = 102, BCorder1_Extrapolate = 103, BCorder2_Extrapolate = 104,
    BCstationary_Wall = 201, BCmoving_Wall = 202, BCstationary_Wall_halfway = 203, BCmoving_Wall_halfway = 204,
    BCPeriodic = 301, BCSymmetric = 302, BCfluid = 0, BCfluid_father = 1
};

// incoming populations per face (FluidDomain.f90:645,729,813,897,981,1065) and their mirror
// sources for BCSymmetric (:697,781,865,949,1033,1117)
__host__ __device__ constexpr int FACE_IN(int face, int k)
{
    constexpr int t[6][5] = {{1, 7, 9, 11, 13}, {2, 8, 10, 12, 14}, {3, 7, 8, 15, 17},
                             {4, 9, 10, 16, 18}, {5, 11, 12, 15, 16}, {6, 13, 14, 17, 18}};
    return t[face][k];
}
__host__ __device__ constexpr int FACE_MIRROR(int face, int k)
{
    constexpr int t[6][5] = {{2, 8, 10, 12, 14}, {1, 7, 9, 11, 13}, {4, 9, 10, 16, 18},
                             {3, 7, 8, 15, 17}, {6, 13, 14, 17, 18}, {5, 11, 12, 15, 16}};
    return t[face][k];
}

// e*v for a lattice coefficient e in {-1,0,1}: exact, no multiply issued
__device__ __forceinline__ double emul(int e, double v) { return e == 0 ? 0.0 : (e > 0 ? v : -v); }

// u . e_q  evaluated as (u1*e1 + u2*e2) + u3*e3 (FluidDomain.f90:1219,1832) with zero terms dropped
__device__ __forceinline__ double udote(int q, double u1, double u2, double u3)
{
    const int e1 = EX(q), e2 = EY(q), e3 = EZ(q);
    if (e1 == 0 && e2 == 0 && e3 == 0) return 0.0;
    if (e2 == 0 && e3 == 0) return emul(e1, u1);
    if (e1 == 0 && e3 == 0) return emul(e2, u2);
    if (e1 == 0 && e2 == 0) return emul(e3, u3);
    if (e3 == 0) return emul(e1, u1) + emul(e2, u2);
    if (e2 == 0) return emul(e1, u1) + emul(e3, u3);
    if (e1 == 0) return emul(e2, u2) + emul(e3, u3);
    return (emul(e1, u1) + emul(e2, u2)) + emul(e3, u3);
}

// calculate_macro_quantities_, FluidDomain.f90:1136-1139.  hF[k] = 0.5d0*volumeForce(k)*dh.
__device__ __forceinline__ void macro_from_f(const double (&f)[Q], const double hF1, const double hF2, const double hF3,
                                             double &den, double &u1, double &u2, double &u3)
{
    double d = f[0];
#pragma unroll
    for (int q = 1; q < Q; q++) d = d + f[q];
    double m1 = f[1]; m1 = m1 - f[2]; m1 = m1 + f[7]; m1 = m1 - f[8]; m1 = m1 + f[9];
    m1 = m1 - f[10]; m1 = m1 + f[11]; m1 = m1 - f[12]; m1 = m1 + f[13]; m1 = m1 - f[14];
    double m2 = f[3]; m2 = m2 - f[4]; m2 = m2 + f[7]; m2 = m2 + f[8]; m2 = m2 - f[9];
    m2 = m2 - f[10]; m2 = m2 + f[15]; m2 = m2 - f[16]; m2 = m2 + f[17]; m2 = m2 - f[18];
    double m3 = f[5]; m3 = m3 - f[6]; m3 = m3 + f[11]; m3 = m3 + f[12]; m3 = m3 - f[13];
    m3 = m3 - f[14]; m3 = m3 + f[15]; m3 = m3 + f[16]; m3 = m3 - f[17]; m3 = m3 - f[18];
    den = d;
    u1 = (m1 + hF1) / d;
    u2 = (m2 + hF2) / d;
    u3 = (m3 + hF3) / d;
}

// calculate_distribution_funcion, FluidDomain.f90:1827-1834
__device__ __forceinline__ void equilibrium(double density, double v1, double v2, double v3, double (&dist)[Q])
{
    double uSqr = v1 * v1;
    uSqr = uSqr + v2 * v2;
    uSqr = uSqr + v3 * v3;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const double uxyz = udote(q, v1, v2, v3);
        dist[q] = WT(q) * density * (1.0 + 3.0 * uxyz + 4.5 * uxyz * uxyz - 1.5 * uSqr);
    }
}

// evaluate_moving_wall, FluidDomain.f90:1837-1843 : out(q) = in(oppo(q)) + 2.0*wt(q)*density*uxyz(q)*3.0
__device__ __forceinline__ void moving_wall(double density, double v1, double v2, double v3, const double (&in)[Q], double (&out)[Q])
{
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const double uxyz = udote(q, v1, v2, v3);
        out[q] = in[OPPO(q)] + 2.0 * WT(q) * density * uxyz * 3.0;
    }
}

struct CollideConsts {
    double Omega, Omega2;     // FluidDomain.f90:453,463
    double dt3;               // 3.d0*dh, :1213
    double cF;                // 1.d0-0.5d0*Omega, :1227
    int mrt_slot;             // which c_MRT entry holds M_COLLID (:514) / M_FORCE (:521) of this block
    double tau, nu, dh;       // LES closures (:1278, :1420, :1503)
};

// What the LES closures contained in collision_ need beyond the cell itself (models 11, 14, 15).
struct LesCtx {
    const double *uuu;        // velocity field of this step (WALE, Vreman: neighbour differences), pointing at local plane 0 of component
                              // 0; null for model 11.  One GPU: [3][X][Y][Z].  x-slab: [3][X+2][Y][Z] with the neighbours' edge planes
                              // in the ghost planes -1 and X (exchanged before the update)
    double *tau_all;          // [X][Y][Z] (FluidDomain.f90:1279,1422,1505)
    size_t ncomp;             // distance between the components of uuu
    int X, Y, Z;              // local extents
    int gx, XG;               // global x of this cell and global x extent: where the reference switches to one-sided differences
    int x, y, z;              // this cell (local)
    bool write_tau;           // false when a boundary layer is collided a second time for the half-way stash
};

// ConstParams.f90:40-45
__device__ __forceinline__ double CsmagConst() { return 2.0 * 0.17 * 0.17 * sqrt(2.0) * 9.0; }
__device__ __forceinline__ double CWALEConst() { return 0.50 * 0.50; }
__device__ __forceinline__ double CvremConst() { return 2.5 * 0.17 * 0.17; }

// center_diff / onesid_diff (FluidDomain.f90:1425-1434) of uuu(.,.,.,k) along `axis`, branching as the reference does
__device__ __forceinline__ double les_grad(const LesCtx &c, int k, int axis, double invdh)
{
    const size_t stride = axis == 0 ? (size_t)c.Y * c.Z : (axis == 1 ? (size_t)c.Z : 1);
    const int pos = axis == 0 ? c.gx : (axis == 1 ? c.y : c.z), dim = axis == 0 ? c.XG : (axis == 1 ? c.Y : c.Z);
    const double *u = c.uuu + (size_t)k * c.ncomp + ((size_t)c.x * c.Y + c.y) * c.Z + c.z;
    if (pos > 0 && pos < dim - 1) return (u[stride] - *(u - stride)) * invdh;
    if (pos == 0) return (-3.0 * u[0] + 4.0 * u[stride] - u[2 * stride]) * invdh;
    return (-3.0 * u[0] + 4.0 * *(u - stride) - *(u - 2 * stride)) * invdh;
}

// smag (:1265-1281), WALE (:1311-1424), vrem (:1435-1507).  fneq = -(f_eq - f).  Returns omega0.
template <int MODEL>
__device__ __forceinline__ double les_omega(const double (&fEqmf)[Q], double rho, const CollideConsts &cc, const LesCtx &c)
{
    const size_t cell = ((size_t)c.x * c.Y + c.y) * c.Z + c.z;
    if (MODEL == 15) {
        const double invdh = cc.dh;   // sic, :1444
        double a[3][3];
#pragma unroll
        for (int j = 0; j < 3; j++)
#pragma unroll
            for (int i = 0; i < 3; i++) a[i][j] = 0.5 * les_grad(c, i, j, invdh);
        const double b11 = a[0][0] * a[0][0], b12 = a[0][1] * a[0][1], b13 = a[0][2] * a[0][2];
        const double b21 = a[1][0] * a[1][0], b22 = a[1][1] * a[1][1], b23 = a[1][2] * a[1][2];
        const double b31 = a[2][0] * a[2][0], b32 = a[2][1] * a[2][1], b33 = a[2][2] * a[2][2];
        const double aa = b11 + b12 + b13 + b21 + b22 + b23 + b31 + b32 + b33;
        const double d12 = a[0][0] * a[1][0] + a[0][1] * a[1][1] + a[0][2] * a[1][2];
        const double d13 = a[0][0] * a[2][0] + a[0][1] * a[2][1] + a[0][2] * a[2][2];
        const double d23 = a[1][0] * a[2][0] + a[1][1] * a[2][1] + a[1][2] * a[2][2];
        const double bb = (b11 + b12 + b13) * (b21 + b22 + b23) - d12 * d12 + (b11 + b12 + b13) * (b31 + b32 + b33) - d13 * d13 +
                          (b21 + b22 + b23) * (b31 + b32 + b33) - d23 * d23;
        double OP = sqrt(bb / aa);
        if (!isfinite(OP)) OP = 0.0;
        const double tau__ = (cc.nu + CvremConst() * OP * cc.dh * cc.dh) / (cc.dh * (1.0 / 3.0)) + 0.5;
        if (c.write_tau) c.tau_all[cell] = tau__;
        return 1.0 / tau__;
    }
    double fneq[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) fneq[q] = -fEqmf[q];
    const double Q11 = fneq[1] + fneq[2] + fneq[7] + fneq[8] + fneq[9] + fneq[10] + fneq[11] + fneq[12] + fneq[13] + fneq[14];
    const double Q22 = fneq[3] + fneq[4] + fneq[7] + fneq[8] + fneq[9] + fneq[10] + fneq[15] + fneq[16] + fneq[17] + fneq[18];
    const double Q33 = fneq[5] + fneq[6] + fneq[11] + fneq[12] + fneq[13] + fneq[14] + fneq[15] + fneq[16] + fneq[17] + fneq[18];
    const double Q12 = fneq[7] - fneq[8] - fneq[9] + fneq[10];
    const double Q13 = fneq[11] - fneq[12] - fneq[13] + fneq[14];
    const double Q23 = fneq[15] - fneq[16] - fneq[17] + fneq[18];
    if (MODEL == 11) {
        const double Qq = Q11 * Q11 + Q22 * Q22 + Q33 * Q33 + 2.0 * (Q12 * Q12 + Q13 * Q13 + Q23 * Q23);
        const double tau_t = sqrt(cc.tau * cc.tau + CsmagConst() * sqrt(Qq) / rho);
        if (c.write_tau) c.tau_all[cell] = 0.5 * (cc.tau + tau_t);
        return 2.0 / (cc.tau + tau_t);
    }
    // MODEL == 14
    const double invdh = cc.dh;   // sic, :1322
    double tau__ = c.tau_all[cell];
    const double S11 = -1.5 * invdh * Q11 / (rho * tau__), S22 = -1.5 * invdh * Q22 / (rho * tau__), S33 = -1.5 * invdh * Q33 / (rho * tau__);
    const double S12 = -1.5 * invdh * Q12 / (rho * tau__), S13 = -1.5 * invdh * Q13 / (rho * tau__), S23 = -1.5 * invdh * Q23 / (rho * tau__);
    const double S = S11 * S11 + S22 * S22 + S33 * S33 + 2.0 * (S12 * S12 + S13 * S13 + S23 * S23);
    double ox = les_grad(c, 2, 1, invdh);
    ox = 0.5 * (ox - les_grad(c, 1, 2, invdh));
    double oy = les_grad(c, 0, 2, invdh);
    oy = 0.5 * (oy - les_grad(c, 2, 0, invdh));
    double oz = les_grad(c, 1, 0, invdh);
    oz = 0.5 * (oz - les_grad(c, 0, 1, invdh));
    const double O12 = -0.5 * oz, O13 = 0.5 * oy, O23 = -0.5 * ox;
    const double O = 2.0 * (O12 * O12 + O23 * O23 + O13 * O13);
    const double SO11 = -(0.0 + S11 * S11 * O12 * O12 + S11 * S11 * O13 * O13 + 0.0 + S12 * S12 * O12 * O12 + S12 * S12 * O13 * O13 + 0.0 +
                          S13 * S13 * O12 * O12 + S13 * S13 * O13 * O13);
    const double SO22 = -(S12 * S12 * O12 * O12 + 0.0 + S12 * S12 * O23 * O23 + S22 * S22 * O12 * O12 + 0.0 + S22 * S22 * O23 * O23 +
                          S23 * S23 * O12 * O12 + 0.0 + S23 * S23 * O23 * O23);
    const double SO33 = -(S13 * S13 * O13 * O13 + S13 * S13 * O23 * O23 + 0.0 + S23 * S23 * O13 * O13 + S23 * S23 * O23 * O23 + 0.0 +
                          S33 * S33 * O13 * O13 + S33 * S33 * O23 * O23 + 0.0);
    const double SO12 = -(0.0 + 0.0 + S11 * S12 * O13 * O23 + 0.0 + 0.0 + S12 * S22 * O13 * O23 + 0.0 + 0.0 + S13 * S23 * O13 * O23);
    const double SO13 = (0.0 + S11 * S13 * O12 * O23 + 0.0 + 0.0 + S12 * S23 * O12 * O23 + 0.0 + 0.0 + S13 * S33 * O12 * O23 + 0.0);
    const double SO23 = -(S12 * S13 * O12 * O13 + 0.0 + 0.0 + S22 * S23 * O12 * O13 + 0.0 + 0.0 + S23 * S33 * O12 * O13 + 0.0 + 0.0);
    const double SO = SO11 + SO22 + SO33 + 2.0 * (SO12 + SO13 + SO23);
    const double SdSd = (S * S + O * O) / 6.0 + 2.0 * S * O / 3.0 + 2.0 * SO;
    double OP = pow(SdSd, 1.5) / (pow(S, 2.5) + pow(SdSd, 1.25));
    if (!isfinite(OP) || OP < 0.0) OP = 0.0;
    tau__ = (cc.nu + CWALEConst() * OP * cc.dh * cc.dh) / (cc.dh * (1.0 / 3.0)) + 0.5;
    if (c.write_tau) c.tau_all[cell] = tau__;
    return 1.0 / tau__;
}

// MRT matrices of up to MRT_SLOTS blocks, row-major [19][19]: [slot][0] = M_COLLID, [slot][1] = M_FORCE
constexpr int MRT_SLOTS = 8;
#ifdef FSILBM_DEFINE_CONSTANTS
__constant__ double c_MRT[MRT_SLOTS][2][Q * Q];
#endif

// The first half of collision_ (FluidDomain.f90:1218-1224) for one population:
// fEq = f_eq - f (:1220), Flb = Guo forcing (:1221-1224).
struct CellPre { double a, den, u1, u2, u3, F1, F2, F3, dt3; };
__device__ __forceinline__ CellPre cell_pre(double den, double u1, double u2, double u3, double F1, double F2, double F3, double dt3)
{
    double uSqr = u1 * u1;
    uSqr = uSqr + u2 * u2;
    uSqr = uSqr + u3 * u3;
    CellPre c;
    c.a = 1.0 - 1.5 * uSqr;
    c.den = den; c.u1 = u1; c.u2 = u2; c.u3 = u3; c.F1 = F1; c.F2 = F2; c.F3 = F3; c.dt3 = dt3;
    return c;
}
__device__ __forceinline__ void collide_term(int q, const CellPre &c, double fq, double &fEq, double &Flb)
{
    const double uxyz = udote(q, c.u1, c.u2, c.u3);
    fEq = WT(q) * c.den * (c.a + uxyz * (3.0 + 4.5 * uxyz)) - fq;
    const double c3 = 3.0 * uxyz;
    const int e1 = EX(q), e2 = EY(q), e3 = EZ(q);
    const double t1 = e1 == 0 ? -c.u1 : (e1 > 0 ? (1.0 - c.u1) + c3 : (-1.0 - c.u1) - c3);
    const double t2 = e2 == 0 ? -c.u2 : (e2 > 0 ? (1.0 - c.u2) + c3 : (-1.0 - c.u2) - c3);
    const double t3 = e3 == 0 ? -c.u3 : (e3 > 0 ? (1.0 - c.u3) + c3 : (-1.0 - c.u3) - c3);
    Flb = c.dt3 * WT(q) * (t1 * c.F1 + t2 * c.F2 + t3 * c.F3);
}

// collision_, FluidDomain.f90:1225-1238.  MODEL: 1 SRT, 2 TRT, 3 MRT.  f is updated in place.
#ifdef FSILBM_DEFINE_CONSTANTS
template <int MODEL>
__device__ __forceinline__ void collide(double (&f)[Q], double den, double u1, double u2, double u3, double F1, double F2,
                                        double F3, const CollideConsts &c, const LesCtx *les = nullptr)
{
    const CellPre pre = cell_pre(den, u1, u2, u3, F1, F2, F3, c.dt3);
    if (MODEL == 11 || MODEL == 14 || MODEL == 15) {   // :1239-1258
        double fEq[Q], Flb[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) collide_term(q, pre, f[q], fEq[q], Flb[q]);
        const double omega = les_omega<MODEL>(fEq, den, c, *les);
        const double cF = 1.0 - 0.5 * omega;
#pragma unroll
        for (int q = 0; q < Q; q++) f[q] = f[q] + omega * fEq[q] + cF * Flb[q];
    } else if (MODEL == 1) {   // :1227
#pragma unroll
        for (int q = 0; q < Q; q++) {
            double fEq, Flb;
            collide_term(q, pre, f[q], fEq, Flb);
            f[q] = f[q] + c.Omega * fEq + c.cF * Flb;
        }
    } else if (MODEL == 2) {   // :1230-1235
        const double hO = 0.5 * c.Omega, hO2 = 0.5 * c.Omega2;
        const double cS = 0.5 - 0.25 * c.Omega, cA = 0.5 - 0.25 * c.Omega2;
        {
            double fEq, Flb;
            collide_term(0, pre, f[0], fEq, Flb);
            f[0] = f[0] + (c.Omega * fEq + c.cF * Flb);
        }
#pragma unroll
        for (int p = 0; p < 9; p++) {
            const int ip = POSDIR(p), in = NEGDIR(p);
            double ep, Fp, en, Fn;
            collide_term(ip, pre, f[ip], ep, Fp);
            collide_term(in, pre, f[in], en, Fn);
            const double S = hO * (ep + en) + cS * (Fp + Fn);
            const double A = hO2 * (ep - en) + cA * (Fp - Fn);
            f[ip] = f[ip] + (S + A);
            f[in] = f[in] + (S - A);
        }
    } else {   // :1238
        double fEq[Q], Flb[Q], mc[Q], mf[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) collide_term(q, pre, f[q], fEq[q], Flb[q]);
#pragma unroll
        for (int i = 0; i < Q; i++) { mc[i] = 0.0; mf[i] = 0.0; }
#pragma unroll
        for (int k = 0; k < Q; k++) {
#pragma unroll
            for (int i = 0; i < Q; i++) mc[i] = mc[i] + c_MRT[c.mrt_slot][0][i * Q + k] * fEq[k];
        }
#pragma unroll
        for (int k = 0; k < Q; k++) {
#pragma unroll
            for (int i = 0; i < Q; i++) mf[i] = mf[i] + c_MRT[c.mrt_slot][1][i * Q + k] * Flb[k];
        }
#pragma unroll
        for (int q = 0; q < Q; q++) f[q] = f[q] + mc[q] + mf[q];
    }
}
#endif

// evaluate_velocity for velocityKind 0 (evaluate_shear_velocity, FluidDomain.f90:1803-1810).
// Kind 2 (oscillatory, :1813-1824) is uniform in space: the host evaluates it with libm's cos and
// passes it in `uniform`.
struct VelocityField {
    int kind;
    double uvwIn[3], shear[3];
    double uniform[3];
};
__device__ __forceinline__ void evaluate_velocity(const VelocityField &v, double zC, double yC, double xC, double &o1, double &o2, double &o3)
{
    if (v.kind == 0) {
        o1 = v.uvwIn[0] + 0.0 * v.shear[0] + yC * v.shear[1] + zC * v.shear[2];
        o2 = v.uvwIn[1] + xC * v.shear[0] + 0.0 * v.shear[1] + zC * v.shear[2];
        o3 = v.uvwIn[2] + xC * v.shear[0] + yC * v.shear[1] + 0.0 * v.shear[2];
    } else {
        o1 = v.uniform[0]; o2 = v.uniform[1]; o3 = v.uniform[2];
    }
}

}  // namespace fsilbm
