// refine_kernels.cu -- sm_100a kernels of the father<->son grid-refinement transfers (reference:
// LBMBlockComm.f90:340-979).  All of them are face-sized (a few 10^4 nodes), one thread per face node with
// the node's 19 populations in registers; buffers are planar [q][a][b] with b (the reference's first, faster
// face index) fastest so that loads and stores coalesce.
//
// The reference interpolates in passes over a temporary plane (first along b on the coincident rows, then
// along a, then the periodic closures).  Here every son node evaluates its own value directly through the same
// chain of expressions (interp_* below), so one launch per face replaces the passes; the floating-point
// expression tree per node is identical to the reference's, hence bit-identical results.
#include "kernels.h"

namespace fsilbm {

// value of the coarse buffer at 1-based (b1, a1)
struct CoarseView {
    const double *p;   // one population's plane [aF][bF] (or the tau plane)
    int bF;
    __device__ __forceinline__ double operator()(int b1, int a1) const { return p[(size_t)(a1 - 1) * bF + (b1 - 1)]; }
};

// interpolate_fIn pass 1 (LBMBlockComm.f90:828-841 cubic, :876-884 linear): rows with odd a, 1 <= b <= bStmp
__device__ __forceinline__ double interp_b(const CoarseView &F, int scheme, int bStmp, int b, int a)
{
    const int a1 = a / 2 + 1;
    if (b & 1) return F(b / 2 + 1, a1);
    const int bo = b - 1, b1 = bo / 2 + 1;   // the odd b of the loop iteration that writes fS(:,b,a)
    if (scheme == 2) {
        if (bo == 1) return 0.375 * F(b1, a1) + 0.75 * F(b1 + 1, a1) - 0.125 * F(b1 + 2, a1);
        if (bo == bStmp - 2) return 0.375 * F(b1 + 1, a1) + 0.75 * F(b1, a1) - 0.125 * F(b1 - 1, a1);
        return -0.0625 * F(b1 - 1, a1) + 0.5625 * F(b1, a1) + 0.5625 * F(b1 + 1, a1) - 0.0625 * F(b1 + 2, a1);
    }
    return (F(b1, a1) + F(b1 + 1, a1)) * 0.5;
}

// pass 2 (:844-856 cubic, :887-891 linear): 1 <= b <= bStmp, 1 <= a <= aStmp
__device__ __forceinline__ double interp_a(const CoarseView &F, int scheme, int bStmp, int aStmp, int b, int a)
{
    if (a & 1) return interp_b(F, scheme, bStmp, b, a);
    if (scheme == 2) {
        if (a == 2) return 0.375 * interp_b(F, 2, bStmp, b, a - 1) + 0.75 * interp_b(F, 2, bStmp, b, a + 1) - 0.125 * interp_b(F, 2, bStmp, b, a + 3);
        if (a == aStmp - 1) return 0.375 * interp_b(F, 2, bStmp, b, a + 1) + 0.75 * interp_b(F, 2, bStmp, b, a - 1) - 0.125 * interp_b(F, 2, bStmp, b, a - 3);
        return -0.0625 * interp_b(F, 2, bStmp, b, a - 3) + 0.5625 * interp_b(F, 2, bStmp, b, a - 1) + 0.5625 * interp_b(F, 2, bStmp, b, a + 1) -
               0.0625 * interp_b(F, 2, bStmp, b, a + 3);
    }
    return (interp_b(F, scheme, bStmp, b, a - 1) + interp_b(F, scheme, bStmp, b, a + 1)) * 0.5;
}

// periodic closure in b (:857-863 cubic, :892-898 linear): 1 <= b <= bS, 1 <= a <= aStmp
__device__ __forceinline__ double interp_r2(const CoarseView &F, int scheme, int bStmp, int aStmp, int b, int a)
{
    if (b <= bStmp) return interp_a(F, scheme, bStmp, aStmp, b, a);
    if (scheme == 2)
        return -0.0625 * interp_a(F, 2, bStmp, aStmp, bStmp - 2, a) + 0.5625 * interp_a(F, 2, bStmp, aStmp, bStmp, a) +
               0.5625 * interp_a(F, 2, bStmp, aStmp, 1, a) - 0.0625 * interp_a(F, 2, bStmp, aStmp, 3, a);
    return (interp_a(F, scheme, bStmp, aStmp, bStmp, a) + interp_a(F, scheme, bStmp, aStmp, 1, a)) * 0.5;
}

// periodic closure in a incl. the corner (:864-871 cubic, :899-905 linear): any son node
__device__ __forceinline__ double interp_node(const CoarseView &F, int scheme, int bS, int aS, int b, int a)
{
    const int bStmp = (bS & 1) ? bS : bS - 1, aStmp = (aS & 1) ? aS : aS - 1;
    if (a <= aStmp) return interp_r2(F, scheme, bStmp, aStmp, b, a);
    if (scheme == 2)
        return -0.0625 * interp_r2(F, 2, bStmp, aStmp, b, aStmp - 2) + 0.5625 * interp_r2(F, 2, bStmp, aStmp, b, aStmp) +
               0.5625 * interp_r2(F, 2, bStmp, aStmp, b, 1) - 0.0625 * interp_r2(F, 2, bStmp, aStmp, b, 3);
    return (interp_r2(F, scheme, bStmp, aStmp, b, aStmp) + interp_r2(F, scheme, bStmp, aStmp, b, 1)) * 0.5;
}

// fIn_GridTransform, LBMBlockComm.f90:958-979
__device__ __forceinline__ void grid_transform(double (&f)[Q], double coeff, double hF1, double hF2, double hF3)
{
    double den, u1, u2, u3;
    macro_from_f(f, hF1, hF2, hF3, den, u1, u2, u3);   // cpt_macro :971-977, same sums as calculate_macro_quantities_
    double uSqr = u1 * u1;
    uSqr = uSqr + u2 * u2;
    uSqr = uSqr + u3 * u3;
    const double a = 1.0 - 1.5 * uSqr;
#pragma unroll
    for (int q = 0; q < Q; q++) {
        const double uxyz = udote(q, u1, u2, u3);
        const double fEq = WT(q) * den * (a + uxyz * (3.0 + 4.5 * uxyz));
        f[q] = fEq + coeff * (f[q] - fEq);
    }
}

// local (x,y,z), 0-based, of face node (b, a) (0-based) on the plane `pl` (0-based) normal to `axis`
__device__ __forceinline__ void face_xyz(int axis, int pl, int b, int a, int &x, int &y, int &z)
{
    if (axis == 0) { x = pl; y = a; z = b; }
    else if (axis == 1) { y = pl; x = a; z = b; }
    else { z = pl; x = a; y = b; }
}

// segment of the father view that holds global plane x (segment 0 = the own slab)
__device__ __forceinline__ int father_segment(const FatherView &v, int x)
{
    int sg = 0;
    for (int i = 1; i < v.nseg; i++) if (x >= v.x0[i] && x < v.x0[i] + v.X[i]) sg = i;
    return sg;
}

// extract_interpolate_layer, LBMBlockComm.f90:340-505: every coupled son face in one launch (blockIdx.z = face; the faces fill
// separate buffers).  time 1: t1 <- father plane; time 2: t2 <- father plane, t1 <- 0.5*(t1+t2).
__global__ void pair_extract_kernel(const __grid_constant__ PairFaces ps, int time)
{
    const PairFaceParams &p = ps.face[blockIdx.z];
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= p.bF || a >= p.aF) return;
    int x, y, z;
    face_xyz(p.axis, p.fplane, p.fb0 + b, p.fa0 + a, x, y, z);
    const Geom &g = p.gF;
    // father indices are global; the plane lies in this rank's slab or, for a son across an interface, in a neighbour's
    const int sg = father_segment(p.fv, x);
    const double *fF = p.fv.f[sg];
    const size_t ps_ = p.fv.pstride[sg];
    x -= p.fv.x0[sg];
    const size_t cell = (size_t)(x + 1) * g.plane + (size_t)y * g.Z + z;
    const size_t n = (size_t)a * p.bF + b, nn = (size_t)p.aF * p.bF;
    double v[Q], o[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) v[q] = fF[q * ps_ + cell];
    if (time == 1) {
#pragma unroll
        for (int q = 0; q < Q; q++) p.buf[0][q * nn + n] = v[q];
    } else {
#pragma unroll
        for (int q = 0; q < Q; q++) o[q] = p.buf[0][q * nn + n];
#pragma unroll
        for (int q = 0; q < Q; q++) { p.buf[1][q * nn + n] = v[q]; p.buf[0][q * nn + n] = 0.5 * (o[q] + v[q]); }
    }
    if (p.tbuf[0]) {   // tau_F?t1 / t2 only exist when a block carries a tau_all field (LES models)
        const double t = p.tauF_all ? p.tauF_all[(size_t)x * g.plane + (size_t)y * g.Z + z] : p.tauF;
        if (time == 1) p.tbuf[0][n] = t;
        else { p.tbuf[1][n] = t; p.tbuf[0][n] = 0.5 * (p.tbuf[0][n] + t); }
    }
}

// interpolation_father_to_son, LBMBlockComm.f90:655-806: every coupled son face in one launch (blockIdx.z = face), one thread per
// son node.  The reference takes the faces one after the other, so a node on an edge or corner of the son keeps the value of the
// LAST face that holds it: a thread whose node also lies on the boundary plane of a later face of the list leaves it to that face
// (every face covers its whole plane, and the values come from the layer buffers only, never from the son: no other coupling
// between the faces).
// Linear scheme, node outside the periodic closures (all but one row / column of a periodic son face): which coarse values a node
// combines depends on the parity of (b, a) only, not on the population -- so the two row offsets are worked out once and the 19
// populations become 19 x (1..4) INDEPENDENT loads, all in flight together, combined exactly as interp_b / interp_a combine them
// ((F + F) * 0.5 along b, then (row + row) * 0.5 along a).  Walking interp_node population by population (the general path, kept
// for the cubic scheme and the closures) is one dependent round trip to memory per population: 187 us per call on the six faces of
// a 321 x 129 x 385 son, whose whole update takes 830.
__global__ void __launch_bounds__(128) pair_f2s_kernel(const __grid_constant__ PairFaces ps, int t)
{
    const PairFaceParams &p = ps.face[blockIdx.z];
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= p.bS || a >= p.aS) return;
    int x, y, z;
    face_xyz(p.axis, p.splane, b, a, x, y, z);
    for (int k = blockIdx.z + 1; k < ps.n; k++) {
        const int ax = ps.face[k].axis, pl = ps.face[k].splane;
        if ((ax == 0 ? x : (ax == 1 ? y : z)) == pl) return;
    }
    const size_t nn = (size_t)p.aF * p.bF;
    const int bStmp = (p.bS & 1) ? p.bS : p.bS - 1, aStmp = (p.aS & 1) ? p.aS : p.aS - 1;
    const int B = b + 1, A = a + 1;   // 1-based, as the reference counts
    double f[Q];
    if (p.scheme != 2 && B <= bStmp && A <= aStmp) {
        const bool be = !(B & 1), ae = !(A & 1);
        const int c = (be ? B - 1 : B) / 2 + 1;                              // interp_b: F(b/2+1, .) or F(b1, .) + F(b1+1, .)
        const int r0 = (ae ? A - 1 : A) / 2 + 1, r1 = (A + 1) / 2 + 1;        // interp_a: the row itself, or rows a-1 and a+1
        const size_t o0 = (size_t)(r0 - 1) * p.bF + (c - 1), o1 = (size_t)(r1 - 1) * p.bF + (c - 1);
        const double *base = p.buf[t];
        double v00[Q], v01[Q], v10[Q], v11[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            v00[q] = base[q * nn + o0];
            v01[q] = be ? base[q * nn + o0 + 1] : 0.0;
            v10[q] = ae ? base[q * nn + o1] : 0.0;
            v11[q] = (ae && be) ? base[q * nn + o1 + 1] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const double lo = be ? (v00[q] + v01[q]) * 0.5 : v00[q];
            const double hi = be ? (v10[q] + v11[q]) * 0.5 : v10[q];
            f[q] = ae ? (lo + hi) * 0.5 : lo;
        }
    } else {
#pragma unroll
        for (int k = 0; k < Q; k++) f[k] = 0.0;
#pragma unroll 1
        for (int q = 0; q < Q; q++) {
            CoarseView F{p.buf[t] + q * nn, p.bF};
            const double v = interp_node(F, p.scheme, p.bS, p.aS, B, A);
#pragma unroll
            for (int k = 0; k < Q; k++) f[k] = k == q ? v : f[k];   // keeps f in registers (no dynamic indexing)
        }
    }
    const Geom &g = p.gS;
    double tauF = p.tauF;   // constant tau: the linear interpolation of a constant is that constant, bit for bit
    if (p.tbuf[0]) { CoarseView T{p.tbuf[t], p.bF}; tauF = interp_node(T, 1, p.bS, p.aS, B, A); }
    const double tauS = p.tauS_all ? p.tauS_all[(size_t)x * g.plane + (size_t)y * g.Z + z] : p.tauS;
    const double coeff = (tauS / tauF) / 2.0;                    // :688
    grid_transform(f, coeff, p.hF[0], p.hF[1], p.hF[2]);          // father's volumeForce and dh, :663-664
    const size_t cell = (size_t)(x + 1) * g.plane + (size_t)y * g.Z + z;
#pragma unroll
    for (int q = 0; q < Q; q++) p.fS[q * g.pstride + cell] = f[q];
}

// deliver_son_to_father, LBMBlockComm.f90:546-653: father plane fi(j) <- son plane si(j), stride 2; every coupled face in one launch
// (blockIdx.z = face).  Where the rectangles of two faces meet, both take the father node from the son node that coincides with it
// -- the same node, the same expressions -- so the faces may run side by side.
__global__ void pair_s2f_kernel(const __grid_constant__ PairFaces ps)
{
    const PairFaceParams &p = ps.face[blockIdx.z];
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= p.nb || a >= p.na) return;
    int xs, ys, zs, xf, yf, zf;
    face_xyz(p.axis, p.siplane, p.sib0 + 2 * b, p.sia0 + 2 * a, xs, ys, zs);
    face_xyz(p.axis, p.fiplane, p.fib0 + b, p.fia0 + a, xf, yf, zf);
    const Geom &gs = p.gS, &gf = p.gF;
    const int sg = father_segment(p.fv, xf);
    double *fF = p.fv.f[sg];
    const size_t psF = p.fv.pstride[sg];
    xf -= p.fv.x0[sg];
    const size_t cs = (size_t)(xs + 1) * gs.plane + (size_t)ys * gs.Z + zs;
    const size_t cf = (size_t)(xf + 1) * gf.plane + (size_t)yf * gf.Z + zf;
    double f[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) f[q] = p.fS[q * gs.pstride + cs];
    const double tauF = p.tauF_all ? p.tauF_all[(size_t)xf * gf.plane + (size_t)yf * gf.Z + zf] : p.tauF;
    const double tauS = p.tauS_all ? p.tauS_all[(size_t)xs * gs.plane + (size_t)ys * gs.Z + zs] : p.tauS;
    const double coeff = (tauF / tauS) * 2.0;                    // :563
    grid_transform(f, coeff, p.hF[0], p.hF[1], p.hF[2]);          // son's volumeForce and dh, :552-553
#pragma unroll
    for (int q = 0; q < Q; q++) fF[q * psF + cf] = f[q];
}

__global__ void flag_signal_kernel(unsigned long long *flag, unsigned long long value)
{
    if (threadIdx.x != 0) return;
    __threadfence_system();   // everything this stream wrote before (also into peer memory) is visible before the flag is
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(value) : "memory");
}

void launch_pair_extract(const PairFaces &ps, int time, cudaStream_t s)
{
    if (ps.n <= 0) return;
    int mb = 0, ma = 0;
    for (int i = 0; i < ps.n; i++) { mb = ps.face[i].bF > mb ? ps.face[i].bF : mb; ma = ps.face[i].aF > ma ? ps.face[i].aF : ma; }
    dim3 block(128), grid((mb + 127) / 128, ma, ps.n);
    pair_extract_kernel<<<grid, block, 0, s>>>(ps, time);
    count_launch();
}
void launch_flag_signal(unsigned long long *flag, unsigned long long value, cudaStream_t s)
{
    flag_signal_kernel<<<1, 32, 0, s>>>(flag, value);
    count_launch();
}
void launch_pair_f2s(const PairFaces &ps, int t, cudaStream_t s)
{
    if (ps.n <= 0) return;
    int mb = 0, ma = 0;
    for (int i = 0; i < ps.n; i++) { mb = ps.face[i].bS > mb ? ps.face[i].bS : mb; ma = ps.face[i].aS > ma ? ps.face[i].aS : ma; }
    dim3 block(128), grid((mb + 127) / 128, ma, ps.n);
    pair_f2s_kernel<<<grid, block, 0, s>>>(ps, t);
    count_launch();
}
void launch_pair_s2f(const PairFaces &ps, cudaStream_t s)
{
    int mb = 0, ma = 0;
    for (int i = 0; i < ps.n; i++) { mb = ps.face[i].nb > mb ? ps.face[i].nb : mb; ma = ps.face[i].na > ma ? ps.face[i].na : ma; }
    if (ps.n <= 0 || mb <= 0 || ma <= 0) return;
    dim3 block(128), grid((mb + 127) / 128, ma, ps.n);
    pair_s2f_kernel<<<grid, block, 0, s>>>(ps);
    count_launch();
}

}  // namespace fsilbm
