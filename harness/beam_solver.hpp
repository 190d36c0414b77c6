// beam_solver.hpp -- C++ restatement of the reference's structural solver (modules SegmentStructure and SolidSolver,
// /root/reference/src/SolidSolver.f90) for the stand-in driver.  SURVEY section 8, row f1.
//
// In a drop-in build this code does not exist: the Fortran driver keeps its own SolidSolver.f90 and only the marker
// arrays cross the C ABI (include/fsilbm.h, fsilbm_ibm_interaction_force).  No Fortran compiler exists in this image,
// so flexible-plate cases (BASELINE configs[0], [3], [4]) can only be run end to end if the driver side is restated;
// this file is that restatement: 2-node 12-dof Timoshenko frame elements with nodal triads (Doyle), lumped mass,
// geometric stiffness from the axial force, Newmark-beta in time, full Newton-Raphson, matrix-free block-Jacobi CG.
// All arrays keep the Fortran index meaning (m[i][j] == m(i+1,j+1)); node and dof numbers are 0-based here.
#pragma once
#include <array>
#include <string>
#include <vector>

namespace harness {

constexpr int nElmtDofs = 12;
using Mat12 = std::array<std::array<double, 12>, 12>;
using Mat3 = std::array<std::array<double, 3>, 3>;
using Vec3 = std::array<double, 3>;

// module-level parameters of SolidSolver (SolidSolver.f90:1154-1158, set by Set_SolidSolver_Params :1249-1264)
struct SolidSolverParams {
    double dampK = 0, dampM = 0, GeoGamma = 0, NewmarkGamma = 0.5, NewmarkBeta = 0.25, dtolFEM = 1e-6;
    double pi = 3.141592653589793;
    int ntolFEM = 20, isKB = 0;
    Vec3 g{0, 0, 0};
};

void Segment_get_angle_triad(const Mat3 &triad_11, const Mat3 &triad_22, double &tx, double &ty, double &tz);   // :909
void Segment_FiniteRot(double t1, double t2, double t3, Mat3 &rr);                                               // :1066
void AoAtoTTT(const Vec3 &AoA, Mat3 &TTT);                                                                       // :2404
bool Invert6x6(const double A[6][6], double Ainv[6][6]);                                                         // :2234

struct Segment {   // type Segment, SolidSolver.f90:13-56
    int node0 = 0, node1 = 0, itype = 2, Nspan = 1;
    int m_localToGlobal[12]{}, bc[12]{};
    double x00[12]{}, x0[12]{}, x1[12]{}, xnxt[12]{};
    double dx0 = 0, dy0 = 0, dz0 = 0, dx1 = 0, dy1 = 0, dz1 = 0, xll0 = 0, xmm0 = 0, xnn0 = 0, xll1 = 0, xmm1 = 0, xnn1 = 0, len0 = 0, len1 = 0;
    double Lspan = 0, spanlen = 0;
    Vec3 dirc00{}, dirc0{}, dirc1{}, dircnxt{};
    double geoFRM = 0, areaElem00 = 0, strainEnergy[2]{};
    Mat3 triad_ee{}, triad_n1{}, triad_n2{}, m_rotMat{};
    double m_property[8]{};
    Mat12 m_coefMat{}, m_tanMat{}, m_stfMat{}, m_masMat{}, m_geoMat{};
    Mat12 m_coefT{};     // transpose of m_coefMat (kept by UpdateMatrix): Multiply walks it column by column

    void Build(int p0Id, int p1Id, int itype_, int Nspan_, const std::vector<std::array<double, 8>> &xyz, const std::array<double, 8> &material,
               const std::vector<std::array<int, 6>> &boundary);                                                  // :60
    void Init();                                                                                                 // :102
    void cptdxyz1();                                                                                             // :119
    void UpdateMatrix(const double coeffs[8], double gamma, double dampM, double dampK);                         // :131
    void UpdateLoad(const double coeffs[8], double dampM, double dampK, const std::vector<double> &dspO, const std::vector<double> &dsp,
                    const std::vector<double> &vel, const std::vector<double> &acc, std::vector<double> &lodEffe) const;   // :159
    void MassMultiply(const double q[12], double mq[12]) const;                                                  // :195
    void BoundaryCond(int iter, std::vector<double> &x, std::vector<char> &fixed, const std::vector<double> &vBC) const;   // :236
    void Multiply(const std::vector<double> &x, std::vector<double> &b) const;                                   // :280
    void LocToGlobal(const double lx[12], std::vector<double> &x) const;                                         // :306
    void FormMassMatrix();                                                                                       // :318
    void FormStiffMatrix();                                                                                      // :365
    void FormGeomMatrix();                                                                                       // :480
    void InitTriad_D();                                                                                          // :602
    void RigidUpdateTriad_D();                                                                                   // :612
    static void BuildAxisDirTriad(double l, double m, double n, const Vec3 &d, Mat3 &triad);                     // :624
    void RotateMatrix();                                                                                         // :699
    void RKR(Mat12 &ek) const;                                                                                   // :709
    void BodyStress_D(std::vector<double> &lodInte);                                                             // :750
    void StrainEnergy_D();                                                                                       // :841
    void UpdateTriad_D(const std::vector<double> &dspnn);                                                        // :966
    void MakeTriad_ee();                                                                                         // :999
    void MapReferencePosToCurrent(double coordsOut[12], const Mat3 &TTT, const Vec3 &XYZ, const Vec3 &AoA) const;   // :1126
    void MapReferenceDirToCurrent(Vec3 &dirc, const Mat3 &TTT) const;                                            // :1139

  private:
    void local_end_rotations(double ub[12]) const;   // the shared first half of BodyStress_D and StrainEnergy_D
};

struct BeamSolver {   // type BeamSolver, SolidSolver.f90:1160-1217
    const SolidSolverParams *P = nullptr;
    std::vector<Segment> m_elements;
    int nND = 0, nEL = 0, nMT = 0, gEQ = 0;
    double FishInfo[3]{};
    std::vector<double> pos, dsp, vel, acc;   // (1:6, 1:nND): [6*node + i]
    std::vector<double> mss;                  // (1:3, 1:nND)
    std::vector<double> lodInte, lodExte, lodEffe, lodFlow, lodRepl, lodGrav, vBC;
    double coeffs[8]{};
    std::string FEmeshName;
    int iBodyModel = 1;
    double Freq = 0, denR = 0, KB = 0, KS = 0, EmR = 0, psR = 0, tcR = 0, St = 0, elmax = 0, elmin = 0;
    Vec3 XYZ{}, XYZo{}, initXYZVel{}, XYZAmpl{}, XYZPhi{}, UVW{};
    Vec3 AoA{}, AoAo{}, AoAAmpl{}, AoAPhi{}, AoAd{}, WWW1{}, WWW2{}, WWW3{};
    Mat3 TTT00{}, TTT0{}, TTTnxt{};
    int isMotionGiven[6]{};
    long long cg_iterations = 0;   // not in the reference: total CG iterations, for the tests

    void ReadBuild(double &nAsfac, double &nLchod);                                                              // :1270
    void Initialise(double time);                                                                                // :1401
    void calculate_angle_material(double Lref, double Uref, double denIn, double &uMax, const Vec3 &uuuIn, double &nLthck);   // :1533
    void structure(int iFish, double time, int isubstep, double deltat, double subdeltat);                       // :1820
    void UpdateStrainEnergy();                                                                                   // :2340
    // writers
    void write_solid(double Lref, double Uref, double Aref, double Fref, int iFish, FILE *fh) const;             // :1621
    void write_solid_params(FILE *fh) const;                                                                     // :1694
    void write_solid_materials(FILE *fh) const;                                                                  // :1712
    void write_solid_info(const std::string &groupNum, const Vec3 &XYZo_, double Lref, double Uref, double Aref, double Fref, double Pref, double Eref);   // :1729
    void write_solid_probes(const std::string &groupNum, const Vec3 &XYZo_, const std::vector<int> &solidProbingNode, double Lref, double Uref, double Aref) const;   // :1801

  private:
    void UpdateVelFromPosAngular(const Vec3 &WWW, const Vec3 &UVW_);                                             // :1478
    void UpdateNewmarkCoeffs(double dt);                                                                         // :1879
    void Solver(int iFish);                                                                                      // :1894
    void UpdateMatrixANDLoad(const std::vector<double> &dspO);                                                   // :1932
    void CG_Solve(std::vector<double> &x, const std::vector<double> &b, int iterNR);                             // :1949
    void MatrixMultipy(const std::vector<double> &x, std::vector<double> &b) const;                              // :2035
    void UpdateDspANDTride(int iter, const std::vector<double> &dspn, double &dnorm);                            // :2304
};

}  // namespace harness
