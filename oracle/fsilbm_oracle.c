/*
 * fsilbm_oracle.c -- CPU ORACLE for the FSILBM3D hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement of the reference's Fortran algorithm for the
 * D3Q19 collide/stream/boundary update (FluidDomain.f90) and the immersed-boundary
 * coupling (Solidbody.f90), in the reference's own pass structure, array layout
 * (fIn(z,y,x,q), z fastest) and left-to-right evaluation order.  Every function
 * cites the reference file:line it follows (paths relative to /root/reference/src).
 *
 * PINNED AGAINST THE REFERENCE: the reference ships no tests or golden vectors and is
 * Fortran; no Fortran compiler exists in this environment.  Its unmodified sources are
 * therefore executed by the interpreter in oracle/ftn/ (main.f90 from start to end on 32
 * cases, tests/reference_cases.py); what the reference leaves is committed as
 * tests/golden/ref_*.npz and this oracle reproduces it bit for bit
 * (tests/test_reference_golden.py).  Also held by the analytic known-answer tests in
 * tests/test_oracle_kat.py.  oracle/pin_with_reference.py repeats the comparison on a
 * gfortran build where one exists.  See DESIGN.md section 5.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (fsilbm3d_b200/, libfsilbm_b200.so) never
 * links, imports or calls it.
 *
 * Build: see oracle/Makefile  (gcc -O3 -fopenmp -ffp-contract=off, no -ffast-math,
 * mirroring the reference Makefile:19,24  "-O3 -fopenmp" without fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define Q 19
#define LBMDIM 18

/* ---- ConstParams.f90:11-25 ------------------------------------------------------ */
static const int ee[Q][3] = {
    {0, 0, 0},  {1, 0, 0},  {-1, 0, 0}, {0, 1, 0},  {0, -1, 0}, {0, 0, 1},  {0, 0, -1},
    {1, 1, 0},  {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1},
    {-1, 0, -1}, {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}};
static const int oppo[Q] = {0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15};
static const int positivedirs[9] = {1, 3, 5, 7, 8, 11, 12, 15, 16};
static const int negativedirs[9] = {2, 4, 6, 10, 9, 14, 13, 18, 17};
static const double wt[Q] = {1.0 / 3.0,  1.0 / 18.0, 1.0 / 18.0, 1.0 / 18.0, 1.0 / 18.0,
                             1.0 / 18.0, 1.0 / 18.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0,
                             1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0,
                             1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
/* ConstParams.f90:28 */
static const double s0 = 0.0, s1 = 1.19, s2 = 1.4, s4 = 1.2, s10 = 1.4, s16 = 1.98;
/* ConstParams.f90:31-34 */
enum {
    BCEq_DirecletU = 101, BCnEq_DirecletU = 102, BCorder1_Extrapolate = 103, BCorder2_Extrapolate = 104,
    BCstationary_Wall = 201, BCmoving_Wall = 202, BCstationary_Wall_halfway = 203, BCmoving_Wall_halfway = 204,
    BCPeriodic = 301, BCSymmetric = 302, BCfluid = 0, BCfluid_father = 1
};
static const double Pi = 3.141592653589793; /* ConstParams.f90:36 */
static const double Cs2 = 1.0 / 3.0;        /* ConstParams.f90:39 */
static const double Csmag = 0.17;           /* ConstParams.f90:40 */
#define CsmagConst (2.0 * Csmag * Csmag * sqrt(2.0) * 9.0) /* :42 */
static const double CWALE = 0.50;           /* :43 */
#define CWALEConst (CWALE * CWALE)          /* :44 */
#define CvremConst (2.5 * Csmag * Csmag)    /* :45 */

/* incoming population sets per face, FluidDomain.f90:645,729,813,897,981,1065 */
static const int face_in[6][5] = {{1, 7, 9, 11, 13}, {2, 8, 10, 12, 14}, {3, 7, 8, 15, 17},
                                  {4, 9, 10, 16, 18}, {5, 11, 12, 15, 16}, {6, 13, 14, 17, 18}};
/* symmetric (mirror) source sets, FluidDomain.f90:697,781,865,949,1033,1117 */
static const int face_mirror[6][5] = {{2, 8, 10, 12, 14}, {1, 7, 9, 11, 13}, {4, 9, 10, 16, 18},
                                      {3, 7, 8, 15, 17}, {6, 13, 14, 17, 18}, {5, 11, 12, 15, 16}};

/* ---- the slice of FlowCondType the hot path reads (FlowCondition.f90:11-27) ------ */
typedef struct {
    double nu, denIn;
    double uvwIn[3], shearRateIn[3];
    int velocityKind;
    double volumeForceIn[3], volumeForceAmp, volumeForceFreq, volumeForcePhi;
    double Uref;
} orc_flow;

/* ---- type LBMBlock, FluidDomain.f90:17-56 (hot-path members only) ---------------- */
typedef struct {
    int iCollidModel;
    int xDim, yDim, zDim;
    double dh, xmin, ymin, zmin, xmax, ymax, zmax;
    int BndConds[6], periodic_bc[3];
    double params[10], tau, Omega, Omega2;
    double M_COLLID[Q][Q], M_FORCE[Q][Q];
    int npsize;
    int *OMPpartition, *OMPparindex, *OMPeid;
    double *OMPedge;
    double *fIn, *uuu, *force, *den, *tau_all;
    double *fIn_hw[6]; /* (0:18, a, b), q fastest; NULL until the first BC call (FluidDomain.f90:660) */
    double volumeForce[3];
    double blktime;
    orc_flow flow;
} orc_block;

#define NZ (b->zDim)
#define NY (b->yDim)
#define NX (b->xDim)
/* 0-based linear index of Fortran fIn(z+1,y+1,x+1,q), FluidDomain.f90:384 */
#define F4(b, z, y, x, q) ((size_t)(z) + (size_t)(b)->zDim * ((size_t)(y) + (size_t)(b)->yDim * ((size_t)(x) + (size_t)(b)->xDim * (size_t)(q))))
#define F3(b, z, y, x) ((size_t)(z) + (size_t)(b)->zDim * ((size_t)(y) + (size_t)(b)->yDim * (size_t)(x)))

/* ================================================================================== */
/* helpers: FluidDomain.f90:1791-1843                                                 */
/* ================================================================================== */

/* evaluate_velocity + evaluate_shear_velocity + evaluate_oscillatory_velocity, :1791-1824 */
static int evaluate_velocity(const orc_flow *fl, double time, double zCoord, double yCoord, double xCoord,
                             const double velocityIn[3], double velocityOut[3], const double shearRate[3])
{
    if (fl->velocityKind == 0) {
        /* :1807-1809, terms kept in source order (the 0*shearRate terms included) */
        velocityOut[0] = velocityIn[0] + 0 * shearRate[0] + yCoord * shearRate[1] + zCoord * shearRate[2];
        velocityOut[1] = velocityIn[1] + xCoord * shearRate[0] + 0 * shearRate[1] + zCoord * shearRate[2];
        velocityOut[2] = velocityIn[2] + xCoord * shearRate[0] + yCoord * shearRate[1] + 0 * shearRate[2];
        return 0;
    } else if (fl->velocityKind == 2) {
        /* :1818-1823 */
        double velocityAmp = shearRate[0], velocityFreq = shearRate[1], velocityPhi = shearRate[2];
        velocityOut[0] = velocityIn[0] + velocityAmp * cos(2 * Pi * velocityFreq * time + velocityPhi / 180.0 * Pi);
        velocityOut[1] = velocityIn[1];
        velocityOut[2] = velocityIn[2];
        return 0;
    }
    /* the reference leaves velocityOut undefined for any other kind (:1795-1799) */
    velocityOut[0] = velocityOut[1] = velocityOut[2] = NAN;
    return 1;
}

/* calculate_distribution_funcion, :1827-1834 */
static void calculate_distribution_funcion(double density, const double velocity[3], double distribution[Q])
{
    double uSqr = 0.0, uxyz[Q];
    for (int k = 0; k < 3; k++) uSqr = uSqr + velocity[k] * velocity[k];
    for (int q = 0; q < Q; q++)
        uxyz[q] = velocity[0] * ee[q][0] + velocity[1] * ee[q][1] + velocity[2] * ee[q][2];
    for (int q = 0; q < Q; q++)
        distribution[q] = wt[q] * density * (1.0 + 3.0 * uxyz[q] + 4.5 * uxyz[q] * uxyz[q] - 1.5 * uSqr);
}

/* evaluate_moving_wall, :1837-1843 */
static void evaluate_moving_wall(double density, const double velocity[3], const double distributionIn[Q],
                                 double distributionOut[Q])
{
    double uxyz[Q];
    for (int q = 0; q < Q; q++)
        uxyz[q] = velocity[0] * ee[q][0] + velocity[1] * ee[q][1] + velocity[2] * ee[q][2];
    for (int q = 0; q < Q; q++)
        distributionOut[q] = distributionIn[oppo[q]] + 2.0 * wt[q] * density * uxyz[q] * 3.0;
}

/* ================================================================================== */
/* block life cycle                                                                   */
/* ================================================================================== */

/* check_periodic_boundary_, FluidDomain.f90:110-125 */
static int check_periodic_boundary(orc_block *b)
{
    for (int i = 0; i < 3; i++) {
        b->periodic_bc[i] = 0;
        if (b->BndConds[2 * i] == BCPeriodic || b->BndConds[2 * i + 1] == BCPeriodic) {
            if (b->BndConds[2 * i] == b->BndConds[2 * i + 1]) b->periodic_bc[i] = 1;
            else return 1; /* 'Periodic boundaries must apper in pairs' -> stop */
        }
    }
    return 0;
}

/* OMPPrePartition, FluidDomain.f90:411-430 (1-based parindex kept) */
static void OMPPrePartition(int xDim, int np, int *partition, int *parindex)
{
    int psize = xDim / np;
    int residual = xDim - psize * np;
    parindex[0] = 1;
    parindex[np] = xDim + 1;
    for (int p = 1; p <= np; p++) {
        if (p > np - residual) partition[p - 1] = psize + 1;
        else partition[p - 1] = psize;
        if (p > 1) parindex[p - 1] = parindex[p - 2] + partition[p - 2];
    }
}

/* read_fuild_blocks (:76-105) + allocate_fluid_ (:378-408): parameters arrive as arguments
 * instead of from inFlow.dat. */
orc_block *orc_block_create(int xDim, int yDim, int zDim, double dh, double xmin, double ymin, double zmin,
                            const int BndConds[6], int iCollidModel, const double params[10], int npsize,
                            const orc_flow *flow)
{
    orc_block *b = (orc_block *)calloc(1, sizeof(orc_block));
    if (!b) return NULL;
    b->iCollidModel = iCollidModel;
    b->xDim = xDim; b->yDim = yDim; b->zDim = zDim;
    b->dh = dh; b->xmin = xmin; b->ymin = ymin; b->zmin = zmin;
    memcpy(b->BndConds, BndConds, sizeof(int) * 6);
    memcpy(b->params, params, sizeof(double) * 10);
    b->flow = *flow;
    if (xDim > 32767 || yDim > 32767 || zDim > 32767) { free(b); return NULL; } /* :88-91 */
    if (check_periodic_boundary(b)) { free(b); return NULL; }
    b->xmax = xmin + dh * (xDim - 1); /* :94-105 */
    b->ymax = ymin + dh * (yDim - 1);
    b->zmax = zmin + dh * (zDim - 1);
    if (b->periodic_bc[0] == 1) b->xmax = b->xmax + dh;
    if (b->periodic_bc[1] == 1) b->ymax = b->ymax + dh;
    if (b->periodic_bc[2] == 1) b->zmax = b->zmax + dh;
    size_t n = (size_t)xDim * yDim * zDim;
    b->fIn = (double *)malloc(sizeof(double) * n * Q);
    b->uuu = (double *)malloc(sizeof(double) * n * 3);
    b->force = (double *)malloc(sizeof(double) * n * 3);
    b->den = (double *)malloc(sizeof(double) * n);
    b->tau_all = (double *)malloc(sizeof(double) * n);
    if (npsize < 1) npsize = 1;
    if (npsize > xDim) npsize = xDim;
    b->npsize = npsize;
    b->OMPpartition = (int *)malloc(sizeof(int) * npsize);
    b->OMPparindex = (int *)malloc(sizeof(int) * (npsize + 1));
    b->OMPeid = (int *)malloc(sizeof(int) * npsize);
    b->OMPedge = (double *)malloc(sizeof(double) * (size_t)zDim * yDim * npsize);
    if (!b->fIn || !b->uuu || !b->force || !b->den || !b->tau_all || !b->OMPedge) return NULL;
    OMPPrePartition(xDim, npsize, b->OMPpartition, b->OMPparindex);
    /* first touch in the OpenMP distribution the sweeps use */
#pragma omp parallel for schedule(static) num_threads(b->npsize)
    for (int x = 0; x < xDim; x++) {
        size_t plane = (size_t)yDim * zDim;
        for (int q = 0; q < Q; q++) memset(b->fIn + ((size_t)q * xDim + x) * plane, 0, sizeof(double) * plane);
        for (int k = 0; k < 3; k++) {
            memset(b->uuu + ((size_t)k * xDim + x) * plane, 0, sizeof(double) * plane);
            memset(b->force + ((size_t)k * xDim + x) * plane, 0, sizeof(double) * plane);
        }
        memset(b->den + (size_t)x * plane, 0, sizeof(double) * plane);
        memset(b->tau_all + (size_t)x * plane, 0, sizeof(double) * plane);
    }
    return b;
}

void orc_block_destroy(orc_block *b)
{
    if (!b) return;
    free(b->fIn); free(b->uuu); free(b->force); free(b->den); free(b->tau_all);
    free(b->OMPpartition); free(b->OMPparindex); free(b->OMPeid); free(b->OMPedge);
    for (int i = 0; i < 6; i++) free(b->fIn_hw[i]);
    free(b);
}

double *orc_block_fIn(orc_block *b) { return b->fIn; }
double *orc_block_uuu(orc_block *b) { return b->uuu; }
double *orc_block_force(orc_block *b) { return b->force; }
double *orc_block_den(orc_block *b) { return b->den; }
double *orc_block_tau_all(orc_block *b) { return b->tau_all; }
double *orc_block_volumeForce(orc_block *b) { return b->volumeForce; }
void orc_block_set_blktime(orc_block *b, double t) { b->blktime = t; }
double orc_block_get(orc_block *b, int what)
{
    switch (what) {
    case 0: return b->tau;
    case 1: return b->Omega;
    case 2: return b->Omega2;
    case 3: return b->xmax;
    case 4: return b->ymax;
    case 5: return b->zmax;
    default: return NAN;
    }
}
double *orc_block_M_COLLID(orc_block *b) { return &b->M_COLLID[0][0]; }
double *orc_block_M_FORCE(orc_block *b) { return &b->M_FORCE[0][0]; }

/* calculate_MRT_params, FluidDomain.f90:466-522.  MATMUL sums run over the inner index in
 * ascending order from zero (gfortran's inlined matmul loop nest). */
static void calculate_MRT_params(orc_block *b)
{
    double M_MRT[Q][Q], M_MRTI[Q][Q], M[Q][Q], S_D[Q][Q], S[Q], T[Q][Q];
    for (int I = 0; I < Q; I++) {
        double e1 = ee[I][0], e2 = ee[I][1], e3 = ee[I][2];
        double sq = (double)(ee[I][0] * ee[I][0] + ee[I][1] * ee[I][1] + ee[I][2] * ee[I][2]);
        M_MRT[0][I] = 1.0;
        M_MRT[1][I] = 19.0 * sq - 30.0;
        M_MRT[2][I] = (21.0 * (sq * sq) - 53.0 * sq + 24.0) / 2.0;
        M_MRT[3][I] = e1;
        M_MRT[5][I] = e2;
        M_MRT[7][I] = e3;
        M_MRT[4][I] = (5.0 * sq - 9.0) * e1;
        M_MRT[6][I] = (5.0 * sq - 9.0) * e2;
        M_MRT[8][I] = (5.0 * sq - 9.0) * e3;
        M_MRT[9][I] = 3.0 * (e1 * e1) - sq;
        M_MRT[10][I] = (3.0 * sq - 5.0) * (3.0 * (e1 * e1) - sq);
        M_MRT[11][I] = e2 * e2 - e3 * e3;
        M_MRT[12][I] = (3.0 * sq - 5.0) * (e2 * e2 - e3 * e3);
        M_MRT[13][I] = e1 * e2;
        M_MRT[14][I] = e2 * e3;
        M_MRT[15][I] = e3 * e1;
        M_MRT[16][I] = (e2 * e2 - e3 * e3) * e1;
        M_MRT[17][I] = (e3 * e3 - e1 * e1) * e2;
        M_MRT[18][I] = (e1 * e1 - e2 * e2) * e3;
    }
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < Q; j++) M_MRTI[i][j] = M_MRT[j][i]; /* :500 */
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < Q; j++) { /* :501 */
            double s = 0.0;
            for (int k = 0; k < Q; k++) s = s + M_MRT[i][k] * M_MRTI[k][j];
            M[i][j] = s;
        }
    for (int I = 0; I < Q; I++) /* :502-504 */
        for (int r = 0; r < Q; r++) M_MRTI[r][I] = M_MRTI[r][I] / M[I][I];
    const double Om = b->Omega;
    const double Sv[Q] = {s0, s1, s2, s0, s4, s0, s4, s0, s4, Om, s10, Om, s10, Om, Om, Om, s16, s16, s16}; /* :507 */
    memcpy(S, Sv, sizeof(S));
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < Q; j++) S_D[i][j] = (i == j) ? S[i] : 0.0;
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < Q; j++) {
            double s = 0.0;
            for (int k = 0; k < Q; k++) s = s + M_MRTI[i][k] * S_D[k][j];
            T[i][j] = s;
        }
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < Q; j++) { /* :514 */
            double s = 0.0;
            for (int k = 0; k < Q; k++) s = s + T[i][k] * M_MRT[k][j];
            b->M_COLLID[i][j] = s;
        }
    for (int i = 0; i < Q; i++)
        for (int j = 0; j < Q; j++) b->M_FORCE[i][j] = ((i == j) ? 1.0 : 0.0) - 0.5 * b->M_COLLID[i][j]; /* :521 */
}

/* initialise_, FluidDomain.f90:433-545 */
int orc_block_initialise(orc_block *b, double time)
{
    /* calculate_SRT_params :450-456 */
    b->tau = b->flow.nu / (b->dh * Cs2) + 0.5;
    b->Omega = 1.0 / b->tau;
    size_t n = (size_t)NX * NY * NZ;
    for (size_t i = 0; i < n; i++) b->tau_all[i] = b->tau;
    if (b->iCollidModel == 2) { /* calculate_TRT_params :458-464 */
        double lambda = b->params[0];
        double tmp = (lambda * 4.0 - 1.0) * b->Omega + 2.0;
        b->Omega2 = 2.0 * (2.0 - b->Omega) / tmp;
    } else if (b->iCollidModel == 3) {
        calculate_MRT_params(b);
    }
    b->blktime = time;
    /* initialise_flow :524-544 */
    int bad = 0;
    for (int x = 0; x < NX; x++) {
        double xCoord = b->xmin + b->dh * x;
        for (int y = 0; y < NY; y++) {
            double yCoord = b->ymin + b->dh * y;
            for (int z = 0; z < NZ; z++) {
                double zCoord = b->zmin + b->dh * z;
                double u[3], d[Q];
                b->den[F3(b, z, y, x)] = b->flow.denIn;
                bad |= evaluate_velocity(&b->flow, b->blktime, zCoord, yCoord, xCoord, b->flow.uvwIn, u, b->flow.shearRateIn);
                for (int k = 0; k < 3; k++) b->uuu[F4(b, z, y, x, k)] = u[k];
                calculate_distribution_funcion(b->den[F3(b, z, y, x)], u, d);
                for (int q = 0; q < Q; q++) b->fIn[F4(b, z, y, x, q)] = d[q];
            }
        }
    }
    return bad;
}

/* ================================================================================== */
/* macro quantities and forces                                                        */
/* ================================================================================== */

/* calculate_macro_quantities_, FluidDomain.f90:1128-1145 */
void orc_calculate_macro_quantities(orc_block *b)
{
#pragma omp parallel for schedule(static) num_threads(b->npsize)
    for (int x = 0; x < NX; x++)
        for (int y = 0; y < NY; y++)
            for (int z = 0; z < NZ; z++) {
                double den = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
                for (int q = 0; q < Q; q++) {
                    double f = b->fIn[F4(b, z, y, x, q)];
                    den = den + f;
                    m1 = m1 + f * ee[q][0];
                    m2 = m2 + f * ee[q][1];
                    m3 = m3 + f * ee[q][2];
                }
                b->den[F3(b, z, y, x)] = den;
                b->uuu[F4(b, z, y, x, 0)] = (m1 + 0.5 * b->volumeForce[0] * b->dh) / den;
                b->uuu[F4(b, z, y, x, 1)] = (m2 + 0.5 * b->volumeForce[1] * b->dh) / den;
                b->uuu[F4(b, z, y, x, 2)] = (m3 + 0.5 * b->volumeForce[2] * b->dh) / den;
            }
}

/* update_volume_force_, FluidDomain.f90:1174-1180 */
void orc_update_volume_force(orc_block *b)
{
    const orc_flow *fl = &b->flow;
    b->volumeForce[0] = fl->volumeForceIn[0] +
                        fl->volumeForceAmp * sin(2.0 * Pi * fl->volumeForceFreq * b->blktime + fl->volumeForcePhi / 180.0 * Pi);
    b->volumeForce[1] = fl->volumeForceIn[1];
    b->volumeForce[2] = fl->volumeForceIn[2];
}

/* add_volume_force_, FluidDomain.f90:1182-1193 */
void orc_add_volume_force(orc_block *b)
{
    size_t plane = (size_t)NY * NZ;
#pragma omp parallel for schedule(static) num_threads(b->npsize)
    for (int x = 0; x < NX; x++)
        for (int k = 0; k < 3; k++) {
            double *p = b->force + ((size_t)k * NX + x) * plane;
            for (size_t i = 0; i < plane; i++) p[i] = p[i] + b->volumeForce[k];
        }
}

/* ResetVolumeForce_, FluidDomain.f90:1195-1206 */
void orc_reset_volume_force(orc_block *b)
{
    size_t plane = (size_t)NY * NZ;
#pragma omp parallel for schedule(static) num_threads(b->npsize)
    for (int x = 0; x < NX; x++)
        for (int k = 0; k < 3; k++) {
            double *p = b->force + ((size_t)k * NX + x) * plane;
            for (size_t i = 0; i < plane; i++) p[i] = 0.0;
        }
}

/* ================================================================================== */
/* collision_, FluidDomain.f90:1208-1263 (SRT :1227, TRT :1230-1235, MRT :1238)        */
/* ================================================================================== */
/* ---- LES closures contained in collision_: smag :1265-1281, WALE :1311-1424, vrem :1435-1507 ------ */
static void les_Q(const double fneq[Q], double *Q11, double *Q22, double *Q33, double *Q12, double *Q13, double *Q23)
{
    *Q11 = fneq[1] + fneq[2] + fneq[7] + fneq[8] + fneq[9] + fneq[10] + fneq[11] + fneq[12] + fneq[13] + fneq[14];
    *Q22 = fneq[3] + fneq[4] + fneq[7] + fneq[8] + fneq[9] + fneq[10] + fneq[15] + fneq[16] + fneq[17] + fneq[18];
    *Q33 = fneq[5] + fneq[6] + fneq[11] + fneq[12] + fneq[13] + fneq[14] + fneq[15] + fneq[16] + fneq[17] + fneq[18];
    *Q12 = fneq[7] - fneq[8] - fneq[9] + fneq[10];
    *Q13 = fneq[11] - fneq[12] - fneq[13] + fneq[14];
    *Q23 = fneq[15] - fneq[16] - fneq[17] + fneq[18];
}
static double center_diff(double g1, double g2, double invdx) { return (g1 - g2) * invdx; }                  /* :1425-1429 */
static double onesid_diff(double g1, double g2, double g3, double invdx) { return (-3.0 * g1 + 4.0 * g2 - g3) * invdx; } /* :1430-1434 */
/* d(uuu(.,.,.,k))/d(axis) as the reference branches it: central inside, one-sided on the first/last plane.
 * c = 0-based (x,y,z); the reference's "invdh" is dh (:1322,1444). */
static double les_grad(const orc_block *b, int k, int axis, const int c[3], double invdh)
{
    const int dim[3] = {NX, NY, NZ};
    int p[3] = {c[0], c[1], c[2]}, m[3] = {c[0], c[1], c[2]}, p2[3] = {c[0], c[1], c[2]};
    if (c[axis] > 0 && c[axis] < dim[axis] - 1) {
        p[axis] += 1; m[axis] -= 1;
        return center_diff(b->uuu[F4(b, p[2], p[1], p[0], k)], b->uuu[F4(b, m[2], m[1], m[0], k)], invdh);
    } else if (c[axis] == 0) {
        p[axis] += 1; p2[axis] += 2;
    } else {
        p[axis] -= 1; p2[axis] -= 2;
    }
    return onesid_diff(b->uuu[F4(b, c[2], c[1], c[0], k)], b->uuu[F4(b, p[2], p[1], p[0], k)], b->uuu[F4(b, p2[2], p2[1], p2[0], k)], invdh);
}
static double les_smag(orc_block *b, const double fneq[Q], double rho, int x, int y, int z)
{
    double Q11, Q22, Q33, Q12, Q13, Q23;
    les_Q(fneq, &Q11, &Q22, &Q33, &Q12, &Q13, &Q23);
    double Qq = Q11 * Q11 + Q22 * Q22 + Q33 * Q33 + 2.0 * (Q12 * Q12 + Q13 * Q13 + Q23 * Q23);
    double tau_t = sqrt(b->tau * b->tau + CsmagConst * sqrt(Qq) / rho);
    b->tau_all[F3(b, z, y, x)] = 0.5 * (b->tau + tau_t);
    return 2.0 / (b->tau + tau_t);
}
static double les_wale(orc_block *b, const double fneq[Q], double rho, int x, int y, int z)
{
    const double invdh = b->dh; /* sic, :1322 */
    const int c[3] = {x, y, z};
    double Q11, Q22, Q33, Q12, Q13, Q23;
    les_Q(fneq, &Q11, &Q22, &Q33, &Q12, &Q13, &Q23);
    double tau__ = b->tau_all[F3(b, z, y, x)];
    double S11 = -1.5 * invdh * Q11 / (rho * tau__), S22 = -1.5 * invdh * Q22 / (rho * tau__), S33 = -1.5 * invdh * Q33 / (rho * tau__);
    double S12 = -1.5 * invdh * Q12 / (rho * tau__), S13 = -1.5 * invdh * Q13 / (rho * tau__), S23 = -1.5 * invdh * Q23 / (rho * tau__);
    double S = S11 * S11 + S22 * S22 + S33 * S33 + 2.0 * (S12 * S12 + S13 * S13 + S23 * S23);
    double ox = les_grad(b, 2, 1, c, invdh);            /* dw/dy */
    ox = 0.5 * (ox - les_grad(b, 1, 2, c, invdh));      /* - dv/dz */
    double oy = les_grad(b, 0, 2, c, invdh);            /* du/dz */
    oy = 0.5 * (oy - les_grad(b, 2, 0, c, invdh));      /* - dw/dx */
    double oz = les_grad(b, 1, 0, c, invdh);            /* dv/dx */
    oz = 0.5 * (oz - les_grad(b, 0, 1, c, invdh));      /* - du/dy */
    double O12 = -0.5 * oz, O13 = 0.5 * oy, O23 = -0.5 * ox;
    double O = 2.0 * (O12 * O12 + O23 * O23 + O13 * O13);
    double SO11 = -(0.0 + S11 * S11 * O12 * O12 + S11 * S11 * O13 * O13 + 0.0 + S12 * S12 * O12 * O12 + S12 * S12 * O13 * O13 + 0.0 +
                    S13 * S13 * O12 * O12 + S13 * S13 * O13 * O13);
    double SO22 = -(S12 * S12 * O12 * O12 + 0.0 + S12 * S12 * O23 * O23 + S22 * S22 * O12 * O12 + 0.0 + S22 * S22 * O23 * O23 +
                    S23 * S23 * O12 * O12 + 0.0 + S23 * S23 * O23 * O23);
    double SO33 = -(S13 * S13 * O13 * O13 + S13 * S13 * O23 * O23 + 0.0 + S23 * S23 * O13 * O13 + S23 * S23 * O23 * O23 + 0.0 +
                    S33 * S33 * O13 * O13 + S33 * S33 * O23 * O23 + 0.0);
    double SO12 = -(0.0 + 0.0 + S11 * S12 * O13 * O23 + 0.0 + 0.0 + S12 * S22 * O13 * O23 + 0.0 + 0.0 + S13 * S23 * O13 * O23);
    double SO13 = (0.0 + S11 * S13 * O12 * O23 + 0.0 + 0.0 + S12 * S23 * O12 * O23 + 0.0 + 0.0 + S13 * S33 * O12 * O23 + 0.0);
    double SO23 = -(S12 * S13 * O12 * O13 + 0.0 + 0.0 + S22 * S23 * O12 * O13 + 0.0 + 0.0 + S23 * S33 * O12 * O13 + 0.0 + 0.0);
    double SO = SO11 + SO22 + SO33 + 2.0 * (SO12 + SO13 + SO23);
    double SdSd = (S * S + O * O) / 6.0 + 2.0 * S * O / 3.0 + 2.0 * SO;
    double OP = pow(SdSd, 1.5) / (pow(S, 2.5) + pow(SdSd, 1.25));
    if (!isfinite(OP) || OP < 0.0) OP = 0.0;
    tau__ = (b->flow.nu + CWALEConst * OP * b->dh * b->dh) / (b->dh * Cs2) + 0.5;
    b->tau_all[F3(b, z, y, x)] = tau__;
    return 1.0 / tau__;
}
static double les_vrem(orc_block *b, int x, int y, int z)
{
    const double invdh = b->dh; /* sic, :1444 */
    const int c[3] = {x, y, z};
    double a[3][3], bb_[3][3]; /* a(i,j) = 0.5 * d u_i / d x_j */
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) a[i][j] = 0.5 * les_grad(b, i, j, c, invdh);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) bb_[i][j] = a[i][j] * a[i][j];
    double b11 = bb_[0][0], b12 = bb_[0][1], b13 = bb_[0][2], b21 = bb_[1][0], b22 = bb_[1][1], b23 = bb_[1][2], b31 = bb_[2][0], b32 = bb_[2][1], b33 = bb_[2][2];
    double a11 = a[0][0], a12 = a[0][1], a13 = a[0][2], a21 = a[1][0], a22 = a[1][1], a23 = a[1][2], a31 = a[2][0], a32 = a[2][1], a33 = a[2][2];
    double aa = b11 + b12 + b13 + b21 + b22 + b23 + b31 + b32 + b33;
    double bb = (b11 + b12 + b13) * (b21 + b22 + b23) - (a11 * a21 + a12 * a22 + a13 * a23) * (a11 * a21 + a12 * a22 + a13 * a23) +
                (b11 + b12 + b13) * (b31 + b32 + b33) - (a11 * a31 + a12 * a32 + a13 * a33) * (a11 * a31 + a12 * a32 + a13 * a33) +
                (b21 + b22 + b23) * (b31 + b32 + b33) - (a21 * a31 + a22 * a32 + a23 * a33) * (a21 * a31 + a22 * a32 + a23 * a33);
    double OP = sqrt(bb / aa);
    if (!isfinite(OP)) OP = 0.0;
    double tau__ = (b->flow.nu + CvremConst * OP * b->dh * b->dh) / (b->dh * Cs2) + 0.5;
    b->tau_all[F3(b, z, y, x)] = tau__;
    return 1.0 / tau__;
}

int orc_collision(orc_block *b)
{
    const double dt3 = 3.0 * b->dh; /* :1213 */
    const int model = b->iCollidModel;
    if (model != 1 && model != 2 && model != 3 && model != 11 && model != 14 && model != 15) return 1; /* 12, 13 are broken upstream */
#pragma omp parallel for schedule(static) num_threads(b->npsize)
    for (int x = 0; x < NX; x++)
        for (int y = 0; y < NY; y++)
            for (int z = 0; z < NZ; z++) {
                double uSqr, uxyz[Q], fEq[Q], Flb[Q];
                const double u1 = b->uuu[F4(b, z, y, x, 0)], u2 = b->uuu[F4(b, z, y, x, 1)], u3 = b->uuu[F4(b, z, y, x, 2)];
                const double F1 = b->force[F4(b, z, y, x, 0)], F2 = b->force[F4(b, z, y, x, 1)], F3_ = b->force[F4(b, z, y, x, 2)];
                const double den = b->den[F3(b, z, y, x)];
                uSqr = 0.0; /* :1218 */
                uSqr = uSqr + u1 * u1; uSqr = uSqr + u2 * u2; uSqr = uSqr + u3 * u3;
                for (int q = 0; q < Q; q++) uxyz[q] = u1 * ee[q][0] + u2 * ee[q][1] + u3 * ee[q][2]; /* :1219 */
                for (int q = 0; q < Q; q++) /* :1220 -- note: this is f_eq - f */
                    fEq[q] = wt[q] * den * ((1.0 - 1.5 * uSqr) + uxyz[q] * (3.0 + 4.5 * uxyz[q])) - b->fIn[F4(b, z, y, x, q)];
                for (int q = 0; q < Q; q++) /* :1221-1224 */
                    Flb[q] = dt3 * wt[q] *
                             ((ee[q][0] - u1 + 3.0 * uxyz[q] * ee[q][0]) * F1 + (ee[q][1] - u2 + 3.0 * uxyz[q] * ee[q][1]) * F2 +
                              (ee[q][2] - u3 + 3.0 * uxyz[q] * ee[q][2]) * F3_);
                if (model == 1) { /* :1227 */
                    for (int q = 0; q < Q; q++) {
                        size_t i = F4(b, z, y, x, q);
                        b->fIn[i] = b->fIn[i] + b->Omega * fEq[q] + (1.0 - 0.5 * b->Omega) * Flb[q];
                    }
                } else if (model == 2) { /* :1230-1235 */
                    fEq[0] = b->Omega * fEq[0] + (1.0 - 0.5 * b->Omega) * Flb[0];
                    for (int p = 0; p < 9; p++) {
                        int ip = positivedirs[p], in = negativedirs[p];
                        uxyz[ip] = 0.5 * b->Omega * (fEq[ip] + fEq[in]) + (0.5 - 0.25 * b->Omega) * (Flb[ip] + Flb[in]);
                    }
                    for (int p = 0; p < 9; p++) {
                        int ip = positivedirs[p], in = negativedirs[p];
                        uxyz[in] = 0.5 * b->Omega2 * (fEq[ip] - fEq[in]) + (0.5 - 0.25 * b->Omega2) * (Flb[ip] - Flb[in]);
                    }
                    for (int p = 0; p < 9; p++) {
                        int ip = positivedirs[p], in = negativedirs[p];
                        fEq[ip] = uxyz[ip] + uxyz[in];
                    }
                    for (int p = 0; p < 9; p++) {
                        int ip = positivedirs[p], in = negativedirs[p];
                        fEq[in] = uxyz[ip] - uxyz[in];
                    }
                    for (int q = 0; q < Q; q++) {
                        size_t i = F4(b, z, y, x, q);
                        b->fIn[i] = b->fIn[i] + fEq[q];
                    }
                } else if (model == 11 || model == 14 || model == 15) { /* :1239-1258 */
                    double fneq[Q], omega;
                    for (int q = 0; q < Q; q++) fneq[q] = -fEq[q];
                    if (model == 11) omega = les_smag(b, fneq, den, x, y, z);
                    else if (model == 14) omega = les_wale(b, fneq, den, x, y, z);
                    else omega = les_vrem(b, x, y, z);
                    for (int q = 0; q < Q; q++) {
                        size_t i = F4(b, z, y, x, q);
                        b->fIn[i] = b->fIn[i] + omega * fEq[q] + (1.0 - 0.5 * omega) * Flb[q];
                    }
                } else { /* :1238 */
                    double mc[Q], mf[Q];
                    for (int i = 0; i < Q; i++) { mc[i] = 0.0; mf[i] = 0.0; }
                    for (int k = 0; k < Q; k++)
                        for (int i = 0; i < Q; i++) mc[i] = mc[i] + b->M_COLLID[i][k] * fEq[k];
                    for (int k = 0; k < Q; k++)
                        for (int i = 0; i < Q; i++) mf[i] = mf[i] + b->M_FORCE[i][k] * Flb[k];
                    for (int q = 0; q < Q; q++) {
                        size_t i = F4(b, z, y, x, q);
                        b->fIn[i] = b->fIn[i] + mc[q] + mf[q];
                    }
                }
            }
    return 0;
}

/* ================================================================================== */
/* streaming_, FluidDomain.f90:1514-1625: in-place periodic shifts, same pass structure */
/* ================================================================================== */

/* swapzy :1525-1568 */
static void swapzy(orc_block *b, int dz, int dy, int i)
{
    if (dz == 0 && dy == 0) return;
    const int zDim = NZ, yDim = NY, xDim = NX;
#pragma omp parallel num_threads(b->npsize)
    {
        double *tmpz = (double *)malloc(sizeof(double) * zDim);
#pragma omp for schedule(static)
        for (int x = 0; x < xDim; x++) {
            double *f = b->fIn + ((size_t)i * xDim + x) * (size_t)yDim * zDim; /* f(:,:,x,i) */
            if (dz == 1) {
                for (int y = 0; y < yDim; y++) {
                    double *l = f + (size_t)y * zDim;
                    double temp = l[zDim - 1];
                    for (int z = zDim - 1; z >= 1; z--) l[z] = l[z - 1];
                    l[0] = temp;
                }
            } else if (dz == -1) {
                for (int y = 0; y < yDim; y++) {
                    double *l = f + (size_t)y * zDim;
                    double temp = l[0];
                    for (int z = 0; z < zDim - 1; z++) l[z] = l[z + 1];
                    l[zDim - 1] = temp;
                }
            }
            if (dy == 1) {
                memcpy(tmpz, f + (size_t)(yDim - 1) * zDim, sizeof(double) * zDim);
                for (int y = yDim - 1; y >= 1; y--) memcpy(f + (size_t)y * zDim, f + (size_t)(y - 1) * zDim, sizeof(double) * zDim);
                memcpy(f, tmpz, sizeof(double) * zDim);
            } else if (dy == -1) {
                memcpy(tmpz, f, sizeof(double) * zDim);
                for (int y = 0; y < yDim - 1; y++) memcpy(f + (size_t)y * zDim, f + (size_t)(y + 1) * zDim, sizeof(double) * zDim);
                memcpy(f + (size_t)(yDim - 1) * zDim, tmpz, sizeof(double) * zDim);
            }
        }
        free(tmpz);
    }
}

/* swapx :1570-1594 with swapxeAtom :1596-1609 and swapxwAtom :1611-1624 (1-based x kept) */
static void swapx(orc_block *b, int dx, int i)
{
    if (dx == 0) return;
    const int xDim = NX;
    const size_t plane = (size_t)NY * NZ;
    double *f = b->fIn + (size_t)i * xDim * plane; /* f(:,:,1..xDim,i), plane x at (x-1)*plane */
    const int np = b->npsize;
#pragma omp parallel for schedule(static) num_threads(np)
    for (int p = 1; p <= np; p++) {
        int xbgn = b->OMPparindex[p - 1], xend = b->OMPparindex[p] - 1;
        double *edge = b->OMPedge + (size_t)(p - 1) * plane;
        if (dx == -1) { /* swapxwAtom */
            int eid = xbgn - 1;
            if (eid == 0) eid = xDim;
            b->OMPeid[p - 1] = eid;
            memcpy(edge, f + (size_t)(xbgn - 1) * plane, sizeof(double) * plane);
            for (int x = xbgn; x <= xend - 1; x++) memcpy(f + (size_t)(x - 1) * plane, f + (size_t)x * plane, sizeof(double) * plane);
        } else if (dx == 1) { /* swapxeAtom */
            int eid = xend + 1;
            if (eid == xDim + 1) eid = 1;
            b->OMPeid[p - 1] = eid;
            memcpy(edge, f + (size_t)(xend - 1) * plane, sizeof(double) * plane);
            for (int x = xend; x >= xbgn + 1; x--) memcpy(f + (size_t)(x - 1) * plane, f + (size_t)(x - 2) * plane, sizeof(double) * plane);
        }
    }
#pragma omp parallel for schedule(static) num_threads(np)
    for (int p = 1; p <= np; p++)
        memcpy(f + (size_t)(b->OMPeid[p - 1] - 1) * plane, b->OMPedge + (size_t)(p - 1) * plane, sizeof(double) * plane);
}

/* streaming_ :1514-1521 */
void orc_streaming(orc_block *b)
{
    for (int i = 0; i <= LBMDIM; i++) {
        swapzy(b, ee[i][2], ee[i][1], i);
        swapx(b, ee[i][0], i);
    }
}

/* ================================================================================== */
/* halfwayBCset_, FluidDomain.f90:567-614                                             */
/* ================================================================================== */
static void face_dims(const orc_block *b, int face, int *na, int *nb)
{
    /* stash shapes: x faces (0:18,zDim,yDim) :661; y faces (0:18,zDim,xDim) :829; z faces (0:18,yDim,xDim) :997 */
    int axis = face / 2;
    if (axis == 0) { *na = b->zDim; *nb = b->yDim; }
    else if (axis == 1) { *na = b->zDim; *nb = b->xDim; }
    else { *na = b->yDim; *nb = b->xDim; }
}

/* map face-local (a,bb) + layer offset (0 = boundary layer, 1 = next interior, ...) to 0-based (z,y,x) */
static inline void face_cell(const orc_block *b, int face, int a, int bb, int layer, int *z, int *y, int *x)
{
    int axis = face / 2, hi = face & 1;
    if (axis == 0) { *z = a; *y = bb; *x = hi ? b->xDim - 1 - layer : layer; }
    else if (axis == 1) { *z = a; *x = bb; *y = hi ? b->yDim - 1 - layer : layer; }
    else { *y = a; *x = bb; *z = hi ? b->zDim - 1 - layer : layer; }
}

void orc_halfway_bc_set(orc_block *b)
{
    for (int face = 0; face < 6; face++) {
        int code = b->BndConds[face];
        if (code != BCstationary_Wall_halfway && code != BCmoving_Wall_halfway) continue;
        if (!b->fIn_hw[face]) continue; /* the reference would fault here; cannot happen after main.f90:63 */
        int na, nb;
        face_dims(b, face, &na, &nb);
        for (int bb = 0; bb < nb; bb++)
            for (int a = 0; a < na; a++) {
                int z, y, x;
                face_cell(b, face, a, bb, 0, &z, &y, &x);
                double *st = b->fIn_hw[face] + ((size_t)bb * na + a) * Q;
                for (int q = 0; q < Q; q++) st[q] = b->fIn[F4(b, z, y, x, q)];
            }
    }
}

/* ================================================================================== */
/* set_boundary_conditions_, FluidDomain.f90:616-1126.  The six copies in the reference */
/* (x-min :623-705, x-max :707-789, y-min :791-873, y-max :875-957, z-min :959-1041,   */
/* z-max :1043-1125) differ only in the face geometry; this is one parametrised copy,  */
/* executed in the same face order so later faces read what earlier faces wrote.       */
/* ================================================================================== */
int orc_set_boundary_conditions(orc_block *b)
{
    int err = 0;
    for (int face = 0; face < 6; face++) {
        const int code = b->BndConds[face];
        const int axis = face / 2, hi = face & 1;
        const int *I = face_in[face];
        int na, nb;
        face_dims(b, face, &na, &nb);
        if (code == BCPeriodic || code == BCfluid || code == BCfluid_father) continue; /* :701-702 */
        if (code == BCstationary_Wall_halfway || code == BCmoving_Wall_halfway) {
            if (!b->fIn_hw[face]) { /* :660-661, first call allocates and skips */
                b->fIn_hw[face] = (double *)calloc((size_t)na * nb * Q, sizeof(double));
                continue;
            }
        }
        if (!(code == BCEq_DirecletU || code == BCnEq_DirecletU || code == BCorder1_Extrapolate ||
              code == BCorder2_Extrapolate || code == BCstationary_Wall || code == BCmoving_Wall ||
              code == BCstationary_Wall_halfway || code == BCmoving_Wall_halfway || code == BCSymmetric))
            return 100 + face; /* 'has no such boundary condition' -> stop (:704) */
        /* wall coordinate along the face normal */
        double wallc;
        {
            double lo = axis == 0 ? b->xmin : axis == 1 ? b->ymin : b->zmin;
            double hic = axis == 0 ? b->xmax : axis == 1 ? b->ymax : b->zmax;
            wallc = hi ? hic : lo;
            if (code == BCmoving_Wall_halfway) wallc = hi ? hic + b->dh * 0.5 : lo - b->dh * 0.5; /* :688,772 */
        }
        for (int bb = 0; bb < nb; bb++)
            for (int a = 0; a < na; a++) {
                int z, y, x, z2, y2, x2, z3, y3, x3;
                face_cell(b, face, a, bb, 0, &z, &y, &x);
                face_cell(b, face, a, bb, 1, &z2, &y2, &x2);
                face_cell(b, face, a, bb, 2, &z3, &y3, &x3);
                double xCoord = b->xmin + b->dh * x, yCoord = b->ymin + b->dh * y, zCoord = b->zmin + b->dh * z;
                if (axis == 0) xCoord = wallc; else if (axis == 1) yCoord = wallc; else zCoord = wallc;
                double velocity[3], fEq[Q], fEqi[Q], fTmp[Q], cur[Q];
                switch (code) {
                case BCEq_DirecletU: /* :623-632 */
                    err |= evaluate_velocity(&b->flow, b->blktime, zCoord, yCoord, xCoord, b->flow.uvwIn, velocity, b->flow.shearRateIn);
                    calculate_distribution_funcion(b->flow.denIn, velocity, fEq);
                    for (int q = 0; q < Q; q++) b->fIn[F4(b, z, y, x, q)] = fEq[q];
                    break;
                case BCnEq_DirecletU: { /* :633-647 */
                    err |= evaluate_velocity(&b->flow, b->blktime, zCoord, yCoord, xCoord, b->flow.uvwIn, velocity, b->flow.shearRateIn);
                    calculate_distribution_funcion(b->flow.denIn, velocity, fEq);
                    double u2[3] = {b->uuu[F4(b, z2, y2, x2, 0)], b->uuu[F4(b, z2, y2, x2, 1)], b->uuu[F4(b, z2, y2, x2, 2)]};
                    calculate_distribution_funcion(b->den[F3(b, z2, y2, x2)], u2, fEqi);
                    for (int k = 0; k < 5; k++)
                        b->fIn[F4(b, z, y, x, I[k])] = fEq[I[k]] + (b->fIn[F4(b, z2, y2, x2, I[k])] - fEqi[I[k]]);
                    break;
                }
                case BCorder1_Extrapolate: /* :648-649 */
                    for (int k = 0; k < 5; k++) b->fIn[F4(b, z, y, x, I[k])] = b->fIn[F4(b, z2, y2, x2, I[k])];
                    break;
                case BCorder2_Extrapolate: /* :650-651 */
                    for (int k = 0; k < 5; k++)
                        b->fIn[F4(b, z, y, x, I[k])] = 2.0 * b->fIn[F4(b, z2, y2, x2, I[k])] - b->fIn[F4(b, z3, y3, x3, I[k])];
                    break;
                case BCstationary_Wall: /* :652-658 */
                    for (int k = 0; k < 5; k++) fTmp[I[k]] = b->fIn[F4(b, z, y, x, oppo[I[k]])];
                    for (int k = 0; k < 5; k++) b->fIn[F4(b, z, y, x, I[k])] = fTmp[I[k]];
                    break;
                case BCstationary_Wall_halfway: { /* :659-669 */
                    const double *st = b->fIn_hw[face] + ((size_t)bb * na + a) * Q;
                    for (int k = 0; k < 5; k++) fTmp[I[k]] = st[oppo[I[k]]];
                    for (int k = 0; k < 5; k++) b->fIn[F4(b, z, y, x, I[k])] = fTmp[I[k]];
                    break;
                }
                case BCmoving_Wall: /* :670-679 */
                    err |= evaluate_velocity(&b->flow, b->blktime, zCoord, yCoord, xCoord, b->flow.uvwIn, velocity, b->flow.shearRateIn);
                    for (int q = 0; q < Q; q++) cur[q] = b->fIn[F4(b, z, y, x, q)];
                    evaluate_moving_wall(b->flow.denIn, velocity, cur, fTmp);
                    for (int k = 0; k < 5; k++) b->fIn[F4(b, z, y, x, I[k])] = fTmp[I[k]];
                    break;
                case BCmoving_Wall_halfway: { /* :680-693 */
                    const double *st = b->fIn_hw[face] + ((size_t)bb * na + a) * Q;
                    err |= evaluate_velocity(&b->flow, b->blktime, zCoord, yCoord, xCoord, b->flow.uvwIn, velocity, b->flow.shearRateIn);
                    evaluate_moving_wall(b->flow.denIn, velocity, st, fTmp);
                    for (int k = 0; k < 5; k++) b->fIn[F4(b, z, y, x, I[k])] = fTmp[I[k]];
                    break;
                }
                case BCSymmetric: /* :694-700 */
                    for (int k = 0; k < 5; k++) fTmp[I[k]] = b->fIn[F4(b, z, y, x, face_mirror[face][k])];
                    for (int k = 0; k < 5; k++) b->fIn[F4(b, z, y, x, I[k])] = fTmp[I[k]];
                    break;
                default: break;
                }
            }
    }
    return err;
}

/* ComputeFieldStat_, FluidDomain.f90:1739-1768: out = L2 u,v,w then Linf u,v,w (serial sum order) */
void orc_compute_field_stat(orc_block *b, double out[6])
{
    double invUref = 1.0 / b->flow.Uref;
    for (int i = 0; i < 3; i++) {
        double uL2 = 0.0, uLinf = -1.0;
        for (int x = 0; x < NX; x++)
            for (int y = 0; y < NY; y++)
                for (int z = 0; z < NZ; z++) {
                    double temp = fabs(b->uuu[F4(b, z, y, x, i)] * invUref);
                    uL2 = uL2 + temp * temp;
                    if (temp > uLinf) uLinf = temp;
                }
        out[i] = sqrt(uL2 / ((double)NX * (double)NY * (double)NZ));
        out[3 + i] = uLinf;
    }
}

/* ================================================================================== */
/* IBM: type VirtualBody, Solidbody.f90:25-68 (marker state only)                      */
/* ================================================================================== */
typedef struct {
    int v_nelmts;
    int v_move, iBodyModel, count_Interp;
    double *v_Exyz, *v_Evel, *v_Ea, *v_Eforce; /* (3,n) (3,n) (n) (3,n) column-major */
    int16_t *v_Ei;                             /* (12,n) integer(2) */
    float *v_Ew;                               /* (12,n) real(4)    */
} orc_body;

orc_body *orc_body_create(int nelmts, int v_move, int iBodyModel)
{
    orc_body *v = (orc_body *)calloc(1, sizeof(orc_body));
    v->v_nelmts = nelmts; v->v_move = v_move; v->iBodyModel = iBodyModel; v->count_Interp = 0;
    v->v_Exyz = (double *)calloc((size_t)3 * nelmts, sizeof(double));
    v->v_Evel = (double *)calloc((size_t)3 * nelmts, sizeof(double));
    v->v_Ea = (double *)calloc((size_t)nelmts, sizeof(double));
    v->v_Eforce = (double *)calloc((size_t)3 * nelmts, sizeof(double));
    v->v_Ei = (int16_t *)calloc((size_t)12 * nelmts, sizeof(int16_t));
    v->v_Ew = (float *)calloc((size_t)12 * nelmts, sizeof(float));
    return v;
}
void orc_body_destroy(orc_body *v)
{
    if (!v) return;
    free(v->v_Exyz); free(v->v_Evel); free(v->v_Ea); free(v->v_Eforce); free(v->v_Ei); free(v->v_Ew); free(v);
}
double *orc_body_Exyz(orc_body *v) { return v->v_Exyz; }
double *orc_body_Evel(orc_body *v) { return v->v_Evel; }
double *orc_body_Ea(orc_body *v) { return v->v_Ea; }
double *orc_body_Eforce(orc_body *v) { return v->v_Eforce; }
int16_t *orc_body_Ei(orc_body *v) { return v->v_Ei; }
float *orc_body_Ew(orc_body *v) { return v->v_Ew; }

/* Phi, Solidbody.f90:822-833 */
double orc_Phi(double x_)
{
    double r = fabs(x_);
    if (r < 1.0) return (3.0 - 2.0 * r + sqrt(1.0 + 4.0 * r * (1.0 - r))) * 0.125;
    else if (r < 2.0) return (5.0 - 2.0 * r - sqrt(-7.0 + 4.0 * r * (3.0 - r))) * 0.125;
    return 0.0;
}

/* minloc_fast, Solidbody.f90:811-821 */
static void minloc_fast(double x_, double x0_, int i0_, double invdh_, int *index_, double *offset_)
{
    *offset_ = (x_ - x0_) * invdh_;
    *index_ = (int)floor(*offset_);
    *offset_ = *offset_ - (double)(*index_);
    *index_ = *index_ + i0_;
}

/* trimedindex, Solidbody.f90:834-866 (1-based indices; returns nonzero where the reference stops) */
static int trimedindex(int i_, int xDim_, int ix_[4], const int bc[2])
{
    for (int k_ = -1; k_ <= 2; k_++) {
        int v = i_ + k_;
        if (v < 1) {
            if (bc[0] == BCPeriodic) v = v + xDim_;
            else if ((bc[0] == BCSymmetric || bc[0] == BCstationary_Wall) && v == 0) v = 2;
            else if (bc[0] == BCstationary_Wall_halfway && v == 0) v = 1;
            else return 1; /* 'index out of xmin bound' */
        } else if (v > xDim_) {
            if (bc[1] == BCPeriodic) v = v - xDim_;
            else if ((bc[1] == BCSymmetric || bc[1] == BCstationary_Wall) && v == xDim_ + 1) v = xDim_ - 1;
            else if (bc[1] == BCstationary_Wall_halfway && v == xDim_ + 1) v = xDim_;
            else return 2; /* 'index out of xmax bound' */
        }
        ix_[k_ + 1] = v;
    }
    return 0;
}

/* UpdateElmtInterp_, Solidbody.f90:760-806; bc = m_boundaryConditions (root block's, main.f90:44) */
int orc_update_elmt_interp(orc_body *v, double dh, double xmin, double ymin, double zmin, int xDim, int yDim, int zDim,
                           const int bc[6])
{
    double invdh = 1.0 / dh;
    int i0 = (int)floor((v->v_Exyz[0] - xmin) * invdh);
    double x0 = xmin + (double)i0 * dh;
    i0 = i0 + 1;
    int j0 = (int)floor((v->v_Exyz[1] - ymin) * invdh);
    double y0 = ymin + (double)j0 * dh;
    j0 = j0 + 1;
    int k0 = (int)floor((v->v_Exyz[2] - zmin) * invdh);
    double z0 = zmin + (double)k0 * dh;
    k0 = k0 + 1;
    int err = 0;
#pragma omp parallel for schedule(static) reduction(| : err)
    for (int iEL = 0; iEL < v->v_nelmts; iEL++) {
        int i, j, k, ix[4], jy[4], kz[4];
        double detx, dety, detz;
        minloc_fast(v->v_Exyz[3 * iEL + 0], x0, i0, invdh, &i, &detx);
        minloc_fast(v->v_Exyz[3 * iEL + 1], y0, j0, invdh, &j, &dety);
        minloc_fast(v->v_Exyz[3 * iEL + 2], z0, k0, invdh, &k, &detz);
        err |= trimedindex(i, xDim, ix, bc + 0);
        err |= trimedindex(j, yDim, jy, bc + 2);
        err |= trimedindex(k, zDim, kz, bc + 4);
        for (int m = 0; m < 4; m++) {
            v->v_Ei[12 * iEL + m] = (int16_t)ix[m];
            v->v_Ei[12 * iEL + 4 + m] = (int16_t)jy[m];
            v->v_Ei[12 * iEL + 8 + m] = (int16_t)kz[m];
            v->v_Ew[12 * iEL + m] = (float)orc_Phi((double)(m - 1) - detx);
            v->v_Ew[12 * iEL + 4 + m] = (float)orc_Phi((double)(m - 1) - dety);
            v->v_Ew[12 * iEL + 8 + m] = (float)orc_Phi((double)(m - 1) - detz);
        }
    }
    return err;
}

#define U4(z, y, x, k) ((size_t)(z) + (size_t)zDim * ((size_t)(y) + (size_t)yDim * ((size_t)(x) + (size_t)xDim * (size_t)(k))))

/* PenaltyForce_, Solidbody.f90:981-1049 */
int orc_penalty_force(orc_body *v, double dt, double dh, int xDim, int yDim, int zDim, double denIn, double *tolerance,
                      double *ntolsum, double *uuu)
{
    const int n = v->v_nelmts;
    double *forceElemTemp = (double *)malloc(sizeof(double) * 3 * (size_t)n);
    double invh3 = 0.5 * dt * ((1.0 / dh) * (1.0 / dh) * (1.0 / dh)) / denIn; /* :996, (1/dh)**3 = ((1/dh)*(1/dh))*(1/dh) */
    double tol = 0.0;
    *ntolsum = (double)n;
#pragma omp parallel for schedule(static) reduction(+ : tol)
    for (int iEL = 0; iEL < n; iEL++) {
        int ix[4], jy[4], kz[4];
        double rx[4], ry[4], rz[4], velElemIB[3] = {0.0, 0.0, 0.0}, forceTemp[3];
        for (int m = 0; m < 4; m++) {
            ix[m] = v->v_Ei[12 * iEL + m]; jy[m] = v->v_Ei[12 * iEL + 4 + m]; kz[m] = v->v_Ei[12 * iEL + 8 + m];
            rx[m] = v->v_Ew[12 * iEL + m]; ry[m] = v->v_Ew[12 * iEL + 4 + m]; rz[m] = v->v_Ew[12 * iEL + 8 + m];
        }
        for (int x = 0; x < 4; x++)
            for (int y = 0; y < 4; y++)
                for (int z = 0; z < 4; z++)
                    for (int k = 0; k < 3; k++) /* :1012 */
                        velElemIB[k] = velElemIB[k] + uuu[U4(kz[z] - 1, jy[y] - 1, ix[x] - 1, k)] * rx[x] * ry[y] * rz[z];
        for (int k = 0; k < 3; k++) {
            velElemIB[k] = v->v_Evel[3 * iEL + k] - velElemIB[k]; /* :1016 */
            forceTemp[k] = velElemIB[k] * v->v_Ea[iEL];           /* :1017 */
        }
        tol = tol + fabs(velElemIB[0]) + fabs(velElemIB[1]) + fabs(velElemIB[2]); /* :1023 */
        for (int k = 0; k < 3; k++) {
            v->v_Eforce[3 * iEL + k] = v->v_Eforce[3 * iEL + k] + forceTemp[k]; /* :1024 */
            forceElemTemp[(size_t)k * n + iEL] = forceTemp[k] * invh3;          /* :1025 */
        }
    }
    *tolerance = tol;
    if (!isfinite(tol)) { free(forceElemTemp); return 1; } /* :1028-1031 'Nan found in PenaltyForce' */
    /* correct velocity, serial, :1034-1048 */
    for (int iEL = 0; iEL < n; iEL++) {
        int ix[4], jy[4], kz[4];
        double rx[4], ry[4], rz[4];
        for (int m = 0; m < 4; m++) {
            ix[m] = v->v_Ei[12 * iEL + m]; jy[m] = v->v_Ei[12 * iEL + 4 + m]; kz[m] = v->v_Ei[12 * iEL + 8 + m];
            rx[m] = v->v_Ew[12 * iEL + m]; ry[m] = v->v_Ew[12 * iEL + 4 + m]; rz[m] = v->v_Ew[12 * iEL + 8 + m];
        }
        for (int x = 0; x < 4; x++)
            for (int y = 0; y < 4; y++)
                for (int z = 0; z < 4; z++)
                    for (int k = 0; k < 3; k++) {
                        size_t c = U4(kz[z] - 1, jy[y] - 1, ix[x] - 1, k);
                        uuu[c] = uuu[c] - forceElemTemp[(size_t)k * n + iEL] * rx[x] * ry[y] * rz[z];
                    }
    }
    free(forceElemTemp);
    return 0;
}

/* FluidVolumeForce_, Solidbody.f90:920-979: the Eulerian scatter half (:968-976).  The nodal-load
 * half (:945-967) needs FEM internals and stays host-side; see orc_plate_nodal_loads. */
void orc_fluid_volume_force(orc_body *v, double dh, int xDim, int yDim, int zDim, double *force)
{
    double invh3 = (1.0 / dh) * (1.0 / dh) * (1.0 / dh); /* :936 */
    for (int iEL = 0; iEL < v->v_nelmts; iEL++) {
        int ix[4], jy[4], kz[4];
        double rx[4], ry[4], rz[4], forceElemTemp[3];
        for (int m = 0; m < 4; m++) {
            ix[m] = v->v_Ei[12 * iEL + m]; jy[m] = v->v_Ei[12 * iEL + 4 + m]; kz[m] = v->v_Ei[12 * iEL + 8 + m];
            rx[m] = v->v_Ew[12 * iEL + m]; ry[m] = v->v_Ew[12 * iEL + 4 + m]; rz[m] = v->v_Ew[12 * iEL + 8 + m];
        }
        for (int k = 0; k < 3; k++) forceElemTemp[k] = v->v_Eforce[3 * iEL + k] * invh3; /* :968 */
        for (int x = 0; x < 4; x++)
            for (int y = 0; y < 4; y++)
                for (int z = 0; z < 4; z++)
                    for (int k = 0; k < 3; k++) { /* :972-974 */
                        double forceTemp = -forceElemTemp[k] * rx[x] * ry[y] * rz[z];
                        size_t c = U4(kz[z] - 1, jy[y] - 1, ix[x] - 1, k);
                        force[c] = force[c] + forceTemp;
                    }
    }
}

/* nodal-load half of FluidVolumeForce_, Solidbody.f90:945-967, for one marker set whose structural
 * element of marker iEL is vtor[iEL] (1-based), element axis centre xc(3,nEL); lodFlow is (6,nND)
 * with localToGlobal being node0/node1 (1-based) per element. */
void orc_plate_nodal_loads(const orc_body *v, const int *vtor, const int *node0, const int *node1, const double *xc,
                           double *lodFlow)
{
    for (int iEL = 0; iEL < v->v_nelmts; iEL++) {
        int iElem = vtor[iEL] - 1;
        const double *F = v->v_Eforce + 3 * iEL;
        double rr[3], M[3];
        for (int k = 0; k < 3; k++) rr[k] = v->v_Exyz[3 * iEL + k] - xc[3 * iElem + k];
        M[0] = rr[1] * F[2] - rr[2] * F[1];
        M[1] = rr[2] * F[0] - rr[0] * F[2];
        M[2] = rr[0] * F[1] - rr[1] * F[0];
        double *l0 = lodFlow + 6 * (node0[iElem] - 1), *l1 = lodFlow + 6 * (node1[iElem] - 1);
        for (int k = 0; k < 3; k++) {
            l0[k] = l0[k] + 0.5 * F[k];
            l1[k] = l1[k] + 0.5 * F[k];
        }
        for (int k = 0; k < 3; k++) {
            l0[3 + k] = l0[3 + k] + 0.5 * M[k];
            l1[3 + k] = l1[3 + k] + 0.5 * M[k];
        }
    }
}

/* calculate_interaction_force, Solidbody.f90:869-918.  Returns iterLBM (>=0) or -1 on a fatal
 * condition of the reference (stencil out of domain, NaN). */
int orc_calculate_interaction_force(orc_body **bodies, int nbodies, double dt, double dh, double xmin, double ymin,
                                    double zmin, int xDim, int yDim, int zDim, double *uuu, double *force,
                                    const int rootBC[6], double denIn, double Uref, int ntolLBM, double dtolLBM)
{
    for (int i = 0; i < nbodies; i++) {
        orc_body *v = bodies[i];
        if (v->v_move == 1 || v->iBodyModel == 2 || v->count_Interp == 0) { /* :885 */
            if (orc_update_elmt_interp(v, dh, xmin, ymin, zmin, xDim, yDim, zDim, rootBC)) return -1;
            v->count_Interp = 1;
        }
        memset(v->v_Eforce, 0, sizeof(double) * 3 * (size_t)v->v_nelmts); /* :889 */
    }
    int iterLBM = 0;
    if (nbodies > 0) {
        double dmaxLBM = 1e10;
        while (iterLBM < ntolLBM && dmaxLBM > dtolLBM) { /* :895 */
            dmaxLBM = 0.0;
            double dsum = 0.0;
            for (int i = 0; i < nbodies; i++) {
                double tol, ntol;
                if (orc_penalty_force(bodies[i], dt, dh, xDim, yDim, zDim, denIn, &tol, &ntol, uuu)) return -1;
                dmaxLBM = dmaxLBM + tol;
                dsum = dsum + ntol;
            }
            dmaxLBM = dmaxLBM / (dsum * Uref); /* :904 */
            iterLBM = iterLBM + 1;
        }
    }
    for (int i = 0; i < nbodies; i++) orc_fluid_volume_force(bodies[i], dh, xDim, yDim, zDim, force); /* :914-917 */
    return iterLBM;
}

/* ================================================================================== */
/* one time step of a block without sons: LBMBlockComm.f90:279-305 (FEM Solver excluded; */
/* the caller advances the bodies between steps, main.f90:93-107 order otherwise).       */
/* ================================================================================== */
int orc_step(orc_block *b, orc_body **bodies, int nbodies, const int rootBC[6], int ntolLBM, double dtolLBM,
             int *iterLBM_out)
{
    orc_update_volume_force(b);        /* :283 */
    orc_calculate_macro_quantities(b); /* :285 */
    orc_reset_volume_force(b);         /* :286 */
    int it = orc_calculate_interaction_force(bodies, nbodies, b->dh, b->dh, b->xmin, b->ymin, b->zmin, b->xDim, b->yDim,
                                             b->zDim, b->uuu, b->force, rootBC, b->flow.denIn, b->flow.Uref, ntolLBM,
                                             dtolLBM); /* :287 -> :328 */
    if (it < 0) return 2;
    if (iterLBM_out) *iterLBM_out = it;
    orc_add_volume_force(b); /* :288 */
    if (orc_collision(b)) return 3; /* :293 */
    orc_halfway_bc_set(b);          /* :296 */
    orc_streaming(b);               /* :299 */
    if (orc_set_boundary_conditions(b)) return 4; /* :303 */
    return 0;
}

/* the two halves of orc_step, for the block-tree recursion (LBMBlockComm.f90:283-288 and :293-303) */
int orc_step_pre(orc_block *b, orc_body **bodies, int nbodies, const int rootBC[6], int ntolLBM, double dtolLBM, int *iterLBM_out)
{
    orc_update_volume_force(b);
    orc_calculate_macro_quantities(b);
    orc_reset_volume_force(b);
    int it = orc_calculate_interaction_force(bodies, nbodies, b->dh, b->dh, b->xmin, b->ymin, b->zmin, b->xDim, b->yDim,
                                             b->zDim, b->uuu, b->force, rootBC, b->flow.denIn, b->flow.Uref, ntolLBM, dtolLBM);
    if (it < 0) return 2;
    if (iterLBM_out) *iterLBM_out = it;
    orc_add_volume_force(b);
    return 0;
}
int orc_step_post(orc_block *b)
{
    if (orc_collision(b)) return 3;
    orc_halfway_bc_set(b);
    orc_streaming(b);
    if (orc_set_boundary_conditions(b)) return 4;
    return 0;
}

/* ================================================================================== */
/* Grid refinement: type CommPair and the father<->son transfers, LBMBlockComm.f90     */
/* ================================================================================== */
typedef struct {
    orc_block *F, *S;
    int sds[6], s[6], f[6], si[6], fi[6]; /* 1-based plane indices as in the reference (:14-15) */
    int dimS[3], dimF[3];                 /* xDimS.. / xDimF.. (:16) */
    int interpolateScheme;                /* flow%interpolateScheme (:825) */
    double *fIn_F[6][2];                  /* fIn_F?t1 / t2: (0:18, b, a), q fastest (:220-260) */
    double *tau_F[6][2];                  /* tau_F?t1 / t2: (b, a) */
} orc_pair;

/* in-plane axes of face j: b = the faster one, a = the slower one (x faces: z,y; y faces: z,x; z faces: y,x) */
static void pair_axes(int j, int *axis, int *bAx, int *aAx)
{
    *axis = j / 2;
    if (*axis == 0) { *bAx = 2; *aAx = 1; }
    else if (*axis == 1) { *bAx = 2; *aAx = 0; }
    else { *bAx = 1; *aAx = 0; }
}
/* linear index into a block's fIn from 1-based (x,y,z) given per axis */
static size_t blk_idx(const orc_block *b, const int c[3], int q) { return F4(b, c[2] - 1, c[1] - 1, c[0] - 1, q); }
static size_t blk_idx3(const orc_block *b, const int c[3]) { return F3(b, c[2] - 1, c[1] - 1, c[0] - 1); }

/* build_blocks_comunication :32-96 + check_blocks_params :508-544 + allocate_fIn_tau :213-264 */
orc_pair *orc_pair_create(orc_block *F, orc_block *S, int interpolateScheme)
{
    const int m_gridDelta = 2;
    int r[3] = {1, 1, 1};
    for (int k = 0; k < 3; k++) if (S->periodic_bc[k] == 1) r[k] = 0;
    int flag = fabs(F->dh - S->dh * (double)m_gridDelta) > 1e-8 || (S->xDim % m_gridDelta) != r[0] ||
               (S->yDim % m_gridDelta) != r[1] || (S->zDim % m_gridDelta) != r[2];
    double res1 = (S->xmin - F->xmin) / F->dh + (S->ymin - F->ymin) / F->dh + (S->zmin - F->zmin) / F->dh;
    double res2 = (S->xmax - F->xmin) / F->dh + (S->ymax - F->ymin) / F->dh + (S->zmax - F->zmin) / F->dh;
    res1 = fabs(res1 - (double)lround(res1));
    res2 = fabs(res2 - (double)lround(res2));
    if (flag || res1 + res2 > 1e-8) return NULL; /* 'grid points do not match between fluid blocks' */
    orc_pair *p = (orc_pair *)calloc(1, sizeof(orc_pair));
    p->F = F; p->S = S; p->interpolateScheme = interpolateScheme;
    for (int j = 0; j < 6; j++) {
        if (S->BndConds[j] == BCfluid) p->sds[j] = (j % 2 == 0) ? 1 : -1;
        else p->sds[j] = 0;
    }
    int sD[3] = {S->xDim, S->yDim, S->zDim};
    for (int k = 0; k < 3; k++) if (S->periodic_bc[k] == 1) sD[k] = sD[k] - 1;
    const double smin[3] = {S->xmin, S->ymin, S->zmin}, fmin[3] = {F->xmin, F->ymin, F->zmin};
    int ratio = (int)floor(F->dh / S->dh + 0.5);
    for (int k = 0; k < 3; k++) {
        p->s[2 * k] = 1;
        p->s[2 * k + 1] = sD[k];
        p->f[2 * k] = (int)floor((smin[k] - fmin[k]) / F->dh + 1.5);
        p->f[2 * k + 1] = p->f[2 * k] + (sD[k] - 1) / ratio;
    }
    for (int j = 0; j < 6; j++) {
        p->si[j] = p->s[j] + p->sds[j] * ratio;
        p->fi[j] = p->f[j] + p->sds[j];
    }
    p->dimS[0] = S->xDim; p->dimS[1] = S->yDim; p->dimS[2] = S->zDim;
    for (int k = 0; k < 3; k++) p->dimF[k] = p->f[2 * k + 1] - p->f[2 * k] + 1;
    for (int j = 0; j < 6; j++) {
        if (S->BndConds[j] != BCfluid) continue;
        int axis, bAx, aAx;
        pair_axes(j, &axis, &bAx, &aAx);
        size_t n = (size_t)p->dimF[bAx] * p->dimF[aAx];
        for (int t = 0; t < 2; t++) {
            p->fIn_F[j][t] = (double *)calloc(n * Q, sizeof(double));
            p->tau_F[j][t] = (double *)calloc(n, sizeof(double));
        }
    }
    return p;
}
void orc_pair_destroy(orc_pair *p)
{
    if (!p) return;
    for (int j = 0; j < 6; j++) for (int t = 0; t < 2; t++) { free(p->fIn_F[j][t]); free(p->tau_F[j][t]); }
    free(p);
}
/* out[0:6]=sds, [6:12]=s, [12:18]=f, [18:24]=si, [24:30]=fi, [30:33]=dimS, [33:36]=dimF */
void orc_pair_get(const orc_pair *p, int out[36])
{
    for (int j = 0; j < 6; j++) { out[j] = p->sds[j]; out[6 + j] = p->s[j]; out[12 + j] = p->f[j]; out[18 + j] = p->si[j]; out[24 + j] = p->fi[j]; }
    for (int k = 0; k < 3; k++) { out[30 + k] = p->dimS[k]; out[33 + k] = p->dimF[k]; }
}

/* extract_interpolate_layer :340-505 for one son; time = 1 or 2 */
void orc_pair_extract_layer(orc_pair *p, int time)
{
    orc_block *F = p->F;
    for (int j = 0; j < 6; j++) {
        if (p->S->BndConds[j] != BCfluid) continue;
        int axis, bAx, aAx;
        pair_axes(j, &axis, &bAx, &aAx);
        const int bF = p->dimF[bAx], aF = p->dimF[aAx];
        double *ft = p->fIn_F[j][time - 1], *tt = p->tau_F[j][time - 1];
        for (int a = 1; a <= aF; a++)
            for (int b = 1; b <= bF; b++) {
                int c[3];
                c[axis] = p->f[j];
                c[bAx] = b + p->f[2 * bAx] - 1;
                c[aAx] = a + p->f[2 * aAx] - 1;
                const size_t n = (size_t)(b - 1) + (size_t)bF * (a - 1);
                for (int e = 0; e < Q; e++) ft[e + Q * n] = F->fIn[blk_idx(F, c, e)];
                tt[n] = F->tau_all[blk_idx3(F, c)];
            }
        if (time == 2) { /* :370-377 */
            const size_t n = (size_t)bF * aF;
            double *f1 = p->fIn_F[j][0], *f2 = p->fIn_F[j][1], *t1 = p->tau_F[j][0], *t2 = p->tau_F[j][1];
            for (size_t i = 0; i < n * Q; i++) f1[i] = 0.5 * (f1[i] + f2[i]);
            for (size_t i = 0; i < n; i++) t1[i] = 0.5 * (t1[i] + t2[i]);
        }
    }
}

/* fIn_GridTransform :958-979 (Dupuis-Chopard rescale of the non-equilibrium part) */
static void fIn_GridTransform(double fIn[Q], double coeff, const double volumeForce[3], double dh)
{
    double den = 0.0, m[3] = {0.0, 0.0, 0.0}, uuu[3];
    for (int q = 0; q < Q; q++) den = den + fIn[q];
    for (int k = 0; k < 3; k++) {
        for (int q = 0; q < Q; q++) m[k] = m[k] + fIn[q] * ee[q][k];
        uuu[k] = (m[k] + 0.5 * volumeForce[k] * dh) / den;
    }
    double uSqr = 0.0;
    for (int k = 0; k < 3; k++) uSqr = uSqr + uuu[k] * uuu[k];
    for (int q = 0; q < Q; q++) {
        double uxyz = uuu[0] * ee[q][0] + uuu[1] * ee[q][1] + uuu[2] * ee[q][2];
        double fEq = wt[q] * den * ((1.0 - 1.5 * uSqr) + uxyz * (3.0 + 4.5 * uxyz));
        fIn[q] = fEq + coeff * (fIn[q] - fEq);
    }
}

/* interpolate_fIn :808-905 (nq = 19) and interpolate_tau :907-956 (nq = 1, always linear).
 * fF(0:nq-1, bF, aF) -> fS(0:nq-1, bS, aS), 1-based b,a as in the reference. */
static void interpolate_plane(int nq, int scheme, int bF, int aF, const double *fF, int bS, int aS, double *fS)
{
#define FF(e, b, a) fF[(e) + (size_t)nq * ((size_t)((b) - 1) + (size_t)bF * ((a) - 1))]
#define FS(e, b, a) fS[(e) + (size_t)nq * ((size_t)((b) - 1) + (size_t)bS * ((a) - 1))]
    int r1 = 0, r2 = 0, bStmp = bS, aStmp = aS;
    if (bS % 2 == 0) { bStmp = bS - 1; r2 = 1; }
    if (aS % 2 == 0) { aStmp = aS - 1; r1 = 1; }
    (void)aF;
    for (int e = 0; e < nq; e++) {
        if (scheme == 2) {
            for (int b = 1; b <= bStmp; b += 2) {
                int b1 = b / 2 + 1;
                for (int a = 1; a <= aStmp; a += 2) {
                    int a1 = a / 2 + 1;
                    FS(e, b, a) = FF(e, b1, a1);
                    if (1 == b) FS(e, b + 1, a) = 0.375 * FF(e, b1, a1) + 0.75 * FF(e, b1 + 1, a1) - 0.125 * FF(e, b1 + 2, a1);
                    else if (b == bStmp - 2) FS(e, b + 1, a) = 0.375 * FF(e, b1 + 1, a1) + 0.75 * FF(e, b1, a1) - 0.125 * FF(e, b1 - 1, a1);
                    else if (b != bStmp) FS(e, b + 1, a) = -0.0625 * FF(e, b1 - 1, a1) + 0.5625 * FF(e, b1, a1) + 0.5625 * FF(e, b1 + 1, a1) - 0.0625 * FF(e, b1 + 2, a1);
                }
            }
            for (int b = 1; b <= bStmp; b++)
                for (int a = 2; a <= aStmp; a += 2) {
                    if (2 == a) FS(e, b, a) = 0.375 * FS(e, b, a - 1) + 0.75 * FS(e, b, a + 1) - 0.125 * FS(e, b, a + 3);
                    else if (a == aStmp - 1) FS(e, b, a) = 0.375 * FS(e, b, a + 1) + 0.75 * FS(e, b, a - 1) - 0.125 * FS(e, b, a - 3);
                    else FS(e, b, a) = -0.0625 * FS(e, b, a - 3) + 0.5625 * FS(e, b, a - 1) + 0.5625 * FS(e, b, a + 1) - 0.0625 * FS(e, b, a + 3);
                }
            if (r2 == 1)
                for (int a = 1; a <= aStmp; a++)
                    FS(e, bStmp + 1, a) = -0.0625 * FS(e, bStmp - 2, a) + 0.5625 * FS(e, bStmp, a) + 0.5625 * FS(e, 1, a) - 0.0625 * FS(e, 3, a);
            if (r1 == 1) {
                for (int b = 1; b <= bStmp; b++)
                    FS(e, b, aStmp + 1) = -0.0625 * FS(e, b, aStmp - 2) + 0.5625 * FS(e, b, aStmp) + 0.5625 * FS(e, b, 1) - 0.0625 * FS(e, b, 3);
                if (r2 == 1)
                    FS(e, bStmp + 1, aStmp + 1) = -0.0625 * FS(e, bStmp + 1, aStmp - 2) + 0.5625 * FS(e, bStmp + 1, aStmp) + 0.5625 * FS(e, bStmp + 1, 1) - 0.0625 * FS(e, bStmp + 1, 3);
            }
        } else {
            for (int b = 1; b <= bStmp; b += 2) {
                int b1 = b / 2 + 1;
                for (int a = 1; a <= aStmp; a += 2) {
                    int a1 = a / 2 + 1;
                    FS(e, b, a) = FF(e, b1, a1);
                    if (b < bStmp) FS(e, b + 1, a) = (FF(e, b1, a1) + FF(e, b1 + 1, a1)) * 0.5;
                }
            }
            for (int b = 1; b <= bStmp; b++)
                for (int a = 2; a <= aStmp; a += 2) FS(e, b, a) = (FS(e, b, a - 1) + FS(e, b, a + 1)) * 0.5;
            if (r2 == 1)
                for (int a = 1; a <= aStmp; a++) FS(e, bStmp + 1, a) = (FS(e, bStmp, a) + FS(e, 1, a)) * 0.5;
            if (r1 == 1) {
                for (int b = 1; b <= bStmp; b++) FS(e, b, aStmp + 1) = (FS(e, b, aStmp) + FS(e, b, 1)) * 0.5;
                if (r2 == 1) FS(e, bStmp + 1, aStmp + 1) = (FS(e, bStmp + 1, aStmp) + FS(e, bStmp + 1, 1)) * 0.5;
            }
        }
    }
#undef FF
#undef FS
}

/* interpolation_father_to_son :655-806 */
void orc_pair_father_to_son(orc_pair *p, int n_timeStep)
{
    const int m_gridDelta = 2;
    orc_block *S = p->S, *F = p->F;
    const double *VF = F->volumeForce;
    const double dh = F->dh;
    for (int j = 0; j < 6; j++) {
        if (!((j % 2 == 0 && p->sds[j] == 1) || (j % 2 == 1 && p->sds[j] == -1))) continue;
        int axis, bAx, aAx;
        pair_axes(j, &axis, &bAx, &aAx);
        const int bS = p->dimS[bAx], aS = p->dimS[aAx], bF = p->dimF[bAx], aF = p->dimF[aAx];
        double *tmpf = (double *)calloc((size_t)Q * bS * aS, sizeof(double));
        double *tmptau = (double *)calloc((size_t)bS * aS, sizeof(double));
        const int t = n_timeStep == 0 ? 0 : 1;
        interpolate_plane(Q, p->interpolateScheme, bF, aF, p->fIn_F[j][t], bS, aS, tmpf);
        interpolate_plane(1, 1, bF, aF, p->tau_F[j][t], bS, aS, tmptau);
        for (int a = 1; a <= aS; a++)
            for (int b = 1; b <= bS; b++) {
                int c[3];
                c[axis] = p->s[j]; c[bAx] = b; c[aAx] = a;
                const size_t n = (size_t)(b - 1) + (size_t)bS * (a - 1);
                double coeff = (S->tau_all[blk_idx3(S, c)] / tmptau[n]) / (double)m_gridDelta;
                fIn_GridTransform(tmpf + Q * n, coeff, VF, dh);
                for (int e = 0; e < Q; e++) S->fIn[blk_idx(S, c, e)] = tmpf[e + Q * n];
            }
        free(tmpf); free(tmptau);
    }
}

/* deliver_son_to_father :546-653 */
void orc_pair_son_to_father(orc_pair *p)
{
    const int m_gridDelta = 2;
    orc_block *S = p->S, *F = p->F;
    const double *VF = S->volumeForce;
    const double dh = S->dh;
    for (int j = 0; j < 6; j++) {
        if (!((j % 2 == 0 && p->sds[j] == 1) || (j % 2 == 1 && p->sds[j] == -1))) continue;
        int axis, bAx, aAx;
        pair_axes(j, &axis, &bAx, &aAx);
        for (int aSn = p->si[2 * aAx]; aSn <= p->si[2 * aAx + 1]; aSn += m_gridDelta) {
            int aFn = (aSn - p->si[2 * aAx]) / 2 + p->fi[2 * aAx];
            for (int bSn = p->si[2 * bAx]; bSn <= p->si[2 * bAx + 1]; bSn += m_gridDelta) {
                int bFn = (bSn - p->si[2 * bAx]) / 2 + p->fi[2 * bAx];
                int cS[3], cF[3];
                cS[axis] = p->si[j]; cS[bAx] = bSn; cS[aAx] = aSn;
                cF[axis] = p->fi[j]; cF[bAx] = bFn; cF[aAx] = aFn;
                double coeff = (F->tau_all[blk_idx3(F, cF)] / S->tau_all[blk_idx3(S, cS)]) * (double)m_gridDelta;
                double tmpf[Q];
                for (int e = 0; e < Q; e++) tmpf[e] = S->fIn[blk_idx(S, cS, e)];
                fIn_GridTransform(tmpf, coeff, VF, dh);
                for (int e = 0; e < Q; e++) F->fIn[blk_idx(F, cF, e)] = tmpf[e];
            }
        }
    }
}

void orc_omp_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_omp_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
