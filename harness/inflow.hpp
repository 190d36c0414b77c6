// inflow.hpp -- reader of the reference's keyword-sectioned parameter file inFlow.dat, restated in C++ for the
// harness (the real driver keeps its Fortran readers).  Citations: /root/reference/src.
//   found_keyword  Util.f90:55-75    readNextData  Util.f90:89-104    readequal  Util.f90:106-121
//   read_flow_conditions  FlowCondition.f90:31-77     read_probe_params  FlowCondition.f90:79-112
//   read_fuild_blocks     FluidDomain.f90:61-108      read_solid_files   Solidbody.f90:73-203 (the lines of the SolidBody section)
#pragma once
#include <array>
#include <string>
#include <vector>

namespace harness {

struct FlowCond {                     // type FlowCondType, FlowCondition.f90:11-27
    int npsize = 1;
    int isConCmpt = 0, numsubstep = 1;
    double timeSimTotal = 0, timeContiDelta = 0, timeWriteBegin = 0, timeWriteEnd = 0;
    double timeFlowDelta = 0, timeBodyDelta = 0, timeInfoDelta = 0;
    double Re = 0, denIn = 1;
    std::array<double, 3> uvwIn{}, shearRateIn{}, volumeForceIn{};
    int velocityKind = 0;
    double volumeForceAmp = 0, volumeForceFreq = 0, volumeForcePhi = 0;
    int LrefType = 0, TrefType = 0, UrefType = 0;
    double Lref = 1, Tref = 1, Uref = 1;
    int ntolLBM = 1;
    double dtolLBM = 1e-10;
    int interpolateScheme = 1;
    // derived (calculate_reference_params, Solidbody.f90:219-284; allocate_solid_memory, :286-326)
    double nu = 0, Mu = 0, Aref = 0, Fref = 0, Eref = 0, Pref = 0;
    double Asfac = 0, Lchod = 0, Lspan = 0, AR = 0;
    // probes
    int fluidProbingNum = 0, inWhichBlock = 0, solidProbingNum = 0;
    std::vector<std::array<double, 3>> fluidProbingCoords;
    std::vector<int> solidProbingNode;
};

struct BlockSpec {                    // the per-block lines of the FluidBlocks section, FluidDomain.f90:76-86
    int ID = 1, iCollidModel = 1, offsetOutput = 0, outputtype = 1;
    int xDim = 0, yDim = 0, zDim = 0;
    double dh = 1, xmin = 0, ymin = 0, zmin = 0;
    std::array<int, 6> BndConds{};
    std::array<double, 10> params{};
};

struct SolidHeader {                  // first five data lines of the SolidBody section, Solidbody.f90:103-113
    double IBPenaltyAlpha = 1, GeoGamma = 0, NewmarkGamma = 0.5, NewmarkBeta = 0.25, dampK = 0, dampM = 0, dtolFEM = 1e-6;
    int ntolFEM = 20, nFish = 0, nGroup = 0, isKB = 0;
};

struct SolidGroup {                   // the per-group lines of the SolidBody section, Solidbody.f90:124-161
    int fishNum = 0, numX = 1, numY = 1, numZ = 1;
    std::string FEmeshName;
    int iBodyModel = 1, iBodyType = 1;
    std::array<int, 6> isMotionGiven{};
    double denR = 0, psR = 0, EmR = 0, tcR = 0, KB = 0, KS = 0, freq = 0, St = 0;
    std::array<double, 3> firstXYZ{}, deltaXYZ{}, initXYZVel{}, XYZAmpl{}, XYZPhi{}, AoAo{}, AoAAmpl{}, AoAPhi{};
};

struct InFlow {
    FlowCond flow;
    std::vector<BlockSpec> blocks;
    SolidHeader solid;
    std::vector<SolidGroup> groups;
};

// Throws std::runtime_error with the reference's message ("<keyword> is not found in inFlow.dat", "end of file
// encounter in readNextData", ...) on malformed input.
InFlow read_inflow(const std::string &filename);

// calculate_reference_params (Solidbody.f90:219-284) for a run without bodies (m_nFish = 0); with bodies the
// structural side supplies the kinematics (SolidBodies::calculate_reference_params in solid_body.hpp).
void calculate_reference_params(FlowCond &flow, int nFish);

}  // namespace harness
