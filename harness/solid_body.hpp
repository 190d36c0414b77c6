// solid_body.hpp -- C++ restatement of the HOST half of module SolidBody (/root/reference/src/Solidbody.f90) for
// the stand-in driver: the per-body marker bookkeeping around the beam solver (PlateBuild_, PlateUpdatePosVelArea_,
// the nodal-load half of FluidVolumeForce_), the reference quantities and the DatBody / DatBodySpan / DatInfo
// writers.  The DEVICE half (UpdateElmtInterp_, the PenaltyForce_ loop, the force spread) is the library's
// fsilbm_ibm_interaction_force (include/fsilbm.h); nothing here touches the GPU.  SURVEY section 8, row f1.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "beam_solver.hpp"
#include "inflow.hpp"

namespace harness {

struct VirtualBody {   // type VirtualBody, Solidbody.f90:25-68
    BeamSolver rbm;
    int v_carrierFluidId = -1;   // 0-based block index
    int v_nelmts = 0, v_type = 1, v_move = 0, count_Interp = 0;
    std::vector<double> v_Exyz, v_Evel, v_Eforce;   // (3, n): [3*i + k]
    std::vector<double> v_Ea;
    std::vector<int> vtor, rtov;                    // 0-based element / first-marker indices

    void PlateBuild();                                                       // :1051
    void PlateUpdatePosVelArea(double IBPenaltyAlpha, double denIn);         // :604
    void UpdatePosVelArea(double IBPenaltyAlpha, double denIn);              // :729
    void NodalLoads();                                                       // :945-967, the host half of FluidVolumeForce_
    void PlateWrite_body(int iFish, FILE *fh, double Lref) const;            // :1186
    void Write_force(int iFish, FILE *fh, double Lref, double Fref) const;   // :1272
};

struct SolidBodies {   // the module variables and procedures of SolidBody
    SolidSolverParams P;
    int m_nFish = 0, m_nGroup = 0, m_ntolLBM = 1;
    double m_dtolLBM = 0, m_IBPenaltyAlpha = 1, m_denIn = 1;
    Vec3 m_uvwIn{};
    double m_Aref = 0, m_Eref = 0, m_Fref = 0, m_Lref = 0, m_Pref = 0, m_Tref = 0, m_Uref = 0;
    int m_boundaryConditions[6]{};
    std::vector<int> m_numX, m_numY, m_numZ, m_fishNum;
    std::vector<Vec3> m_XYZo;
    std::vector<VirtualBody> VBodies;

    void read_solid_files(const InFlow &in, const Vec3 &g);                                           // :73
    void allocate_solid_memory(FlowCond &flow);                                                        // :286
    void calculate_reference_params(FlowCond &flow) const;                                             // :219
    void set_solidbody_parameters(const FlowCond &flow, const int BndConds[6]);                        // :328
    void Initialise_solid_bodies(double time);                                                         // :350
    void Solver(const std::vector<int> &bodies, double time, int isubstep, double deltat, double subdeltat);   // :386
    void Advance(const std::vector<int> &bodies, double time, int numsubstep, double deltat);          // loads + sub-steps + markers, threads over bodies
    void write_solid_field(double time) const;                                                         // :477
    void Write_solid_v_bodies(double time) const;                                                      // :403
    void Write_solid_v_forces(double time) const;                                                      // :428
    void Write_solid_Check(const std::string &filename) const;                                         // :501
    void write_solid_Information(double time, const std::vector<int> &solidProbingNode);               // :523
    static void write_information_titles(int nGroup, const FlowCond &flow);                            // FlowCondition.f90:129
};

}  // namespace harness
