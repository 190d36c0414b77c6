// inflow.cpp -- see inflow.hpp.
#include "inflow.hpp"

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace harness {
namespace {

class Reader {
  public:
    explicit Reader(const std::string &filename) : name_(filename)
    {
        std::ifstream in(filename);
        if (!in) throw std::runtime_error("cannot open " + filename);
        std::string line;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            lines_.push_back(line);
        }
    }
    void rewind() { pos_ = 0; }

    // found_keyword, Util.f90:55-75: list-directed read of the first token of each following line; both strings have
    // only their FIRST character lower-cased (to_lowercase declares `character:: string`, i.e. length 1, :78-87);
    // the section is found when the token contains the keyword.
    void found_keyword(std::string keyword)
    {
        first_lower(keyword);
        while (pos_ < lines_.size()) {
            std::string tok = first_token(lines_[pos_++]);
            if (tok.empty()) continue;   // a list-directed read skips blank lines
            first_lower(tok);
            if (tok.find(keyword) != std::string::npos) return;
        }
        throw std::runtime_error(keyword + " is not found in inFlow.dat");
    }

    // readNextData, Util.f90:89-104: next line that does not start with '#' after adjustl
    std::string next_data()
    {
        while (pos_ < lines_.size()) {
            std::string b = lines_[pos_++];
            const size_t i = b.find_first_not_of(" \t");
            b = i == std::string::npos ? std::string() : b.substr(i);
            if (b.empty() || b[0] != '#') return b;
        }
        throw std::runtime_error("end of file encounter in readNextData");
    }

    // readequal, Util.f90:106-121: skip to the next line whose first token starts with '='
    void readequal()
    {
        while (pos_ < lines_.size()) {
            const std::string tok = first_token(lines_[pos_++]);
            if (!tok.empty() && tok[0] == '=') return;
        }
        throw std::runtime_error("end of file encounter in readequal");
    }

  private:
    static void first_lower(std::string &s)
    {
        if (!s.empty() && s[0] >= 'A' && s[0] <= 'Z') s[0] = (char)(s[0] - 'A' + 'a');
    }
    static std::string first_token(const std::string &line)
    {
        const size_t i = line.find_first_not_of(" \t");
        if (i == std::string::npos) return std::string();
        const size_t j = line.find_first_of(" \t,/", i);
        return line.substr(i, j == std::string::npos ? std::string::npos : j - i);
    }
    std::string name_;
    std::vector<std::string> lines_;
    size_t pos_ = 0;
};

// list-directed read of `n` numeric items from a buffer: blanks or commas separate, Fortran d-exponents allowed,
// anything after the n-th item is ignored (trailing comments in the parameter file)
std::vector<double> items(const std::string &buffer, size_t n)
{
    std::string b = buffer;
    for (char &c : b) if (c == ',') c = ' ';
    std::istringstream is(b);
    std::vector<double> out;
    std::string tok;
    while (out.size() < n && (is >> tok)) {
        for (char &c : tok) if (c == 'd' || c == 'D') c = 'e';
        char *end = nullptr;
        const double v = std::strtod(tok.c_str(), &end);
        if (end == tok.c_str()) throw std::runtime_error("bad numeric item '" + tok + "' in line: " + buffer);
        out.push_back(v);
    }
    if (out.size() < n) throw std::runtime_error("too few items in line: " + buffer);
    return out;
}

}  // namespace

InFlow read_inflow(const std::string &filename)
{
    InFlow in;
    Reader r(filename);
    FlowCond &f = in.flow;
    // read_flow_conditions, FlowCondition.f90:31-77
    r.found_keyword("Parallel");
    f.npsize = (int)items(r.next_data(), 1)[0];
    r.rewind();
    r.found_keyword("FlowCondition");
    { auto v = items(r.next_data(), 2); f.isConCmpt = (int)v[0]; f.numsubstep = (int)v[1]; }
    { auto v = items(r.next_data(), 2); f.timeSimTotal = v[0]; f.timeContiDelta = v[1]; }
    { auto v = items(r.next_data(), 2); f.timeWriteBegin = v[0]; f.timeWriteEnd = v[1]; }
    { auto v = items(r.next_data(), 3); f.timeFlowDelta = v[0]; f.timeBodyDelta = v[1]; f.timeInfoDelta = v[2]; }
    { auto v = items(r.next_data(), 2); f.Re = v[0]; f.denIn = v[1]; }
    { auto v = items(r.next_data(), 3); f.uvwIn = {v[0], v[1], v[2]}; }
    { auto v = items(r.next_data(), 4); f.shearRateIn = {v[0], v[1], v[2]}; f.velocityKind = (int)v[3]; }
    { auto v = items(r.next_data(), 3); f.volumeForceIn = {v[0], v[1], v[2]}; }
    { auto v = items(r.next_data(), 3); f.volumeForceAmp = v[0]; f.volumeForceFreq = v[1]; f.volumeForcePhi = v[2]; }
    { auto v = items(r.next_data(), 2); f.LrefType = (int)v[0]; f.Lref = v[1]; }
    { auto v = items(r.next_data(), 2); f.TrefType = (int)v[0]; f.Tref = v[1]; }
    { auto v = items(r.next_data(), 2); f.UrefType = (int)v[0]; f.Uref = v[1]; }
    { auto v = items(r.next_data(), 2); f.ntolLBM = (int)v[0]; f.dtolLBM = v[1]; }
    f.interpolateScheme = (int)items(r.next_data(), 1)[0];

    // read_solid_files, Solidbody.f90:99-113 (global lines only; per-group lines belong to the structural solver)
    r.rewind();
    r.found_keyword("SolidBody");
    SolidHeader &s = in.solid;
    { auto v = items(r.next_data(), 2); s.IBPenaltyAlpha = v[0]; s.GeoGamma = v[1]; }
    { auto v = items(r.next_data(), 2); s.NewmarkGamma = v[0]; s.NewmarkBeta = v[1]; }
    { auto v = items(r.next_data(), 2); s.dampK = v[0]; s.dampM = v[1]; }
    { auto v = items(r.next_data(), 2); s.dtolFEM = v[0]; s.ntolFEM = (int)v[1]; }
    { auto v = items(r.next_data(), 3); s.nFish = (int)v[0]; s.nGroup = (int)v[1]; s.isKB = (int)v[2]; }
    if (s.IBPenaltyAlpha <= 1e-6) throw std::runtime_error("ERROR: IBPenaltyalpha should be positive (default 1)");
    for (int ig = 0; ig < s.nGroup; ig++) {   // Solidbody.f90:124-161
        SolidGroup g;
        { auto v = items(r.next_data(), 4); g.fishNum = (int)v[0]; g.numX = (int)v[1]; g.numY = (int)v[2]; g.numZ = (int)v[3]; }
        {
            std::string b = r.next_data();   // read(buffer,*) t_FEmeshName: first blank/comma-delimited item, quotes allowed
            const size_t i0 = b.find_first_not_of(" \t");
            if (i0 == std::string::npos) throw std::runtime_error("empty FE mesh name in the SolidBody section");
            if (b[i0] == '\'' || b[i0] == '"') { const size_t j = b.find(b[i0], i0 + 1); g.FEmeshName = b.substr(i0 + 1, j == std::string::npos ? std::string::npos : j - i0 - 1); }
            else { const size_t j = b.find_first_of(" \t,/", i0); g.FEmeshName = b.substr(i0, j == std::string::npos ? std::string::npos : j - i0); }
        }
        { auto v = items(r.next_data(), 2); g.iBodyModel = (int)v[0]; g.iBodyType = (int)v[1]; }
        { auto v = items(r.next_data(), 3); for (int k = 0; k < 3; k++) g.isMotionGiven[k] = (int)v[k]; }
        { auto v = items(r.next_data(), 3); for (int k = 0; k < 3; k++) g.isMotionGiven[3 + k] = (int)v[k]; }
        { auto v = items(r.next_data(), 2); g.denR = v[0]; g.psR = v[1]; }
        { auto v = items(r.next_data(), 2); if (s.isKB == 0) { g.EmR = v[0]; g.tcR = v[1]; } else { g.KB = v[0]; g.KS = v[1]; } }
        { auto v = items(r.next_data(), 2); g.freq = v[0]; g.St = v[1]; }
        auto vec3 = [&](std::array<double, 3> &a) { auto v = items(r.next_data(), 3); a = {v[0], v[1], v[2]}; };
        vec3(g.firstXYZ); vec3(g.deltaXYZ); vec3(g.initXYZVel); vec3(g.XYZAmpl); vec3(g.XYZPhi); vec3(g.AoAo); vec3(g.AoAAmpl); vec3(g.AoAPhi);
        if (ig < s.nGroup - 1) r.readequal();
        if (s.isKB != 0 && s.isKB != 1) { g.EmR = 0.0; g.tcR = 0.0; g.KB = 0.0; g.KS = 0.0; }   // :178-183
        in.groups.push_back(g);
    }

    // read_fuild_blocks, FluidDomain.f90:61-108
    r.rewind();
    r.found_keyword("FluidBlocks");
    const int nblock = (int)items(r.next_data(), 1)[0];
    for (int ib = 0; ib < nblock; ib++) {
        BlockSpec b;
        { auto v = items(r.next_data(), 4); b.ID = (int)v[0]; b.iCollidModel = (int)v[1]; b.offsetOutput = (int)v[2]; b.outputtype = (int)v[3]; }
        { auto v = items(r.next_data(), 3); b.xDim = (int)v[0]; b.yDim = (int)v[1]; b.zDim = (int)v[2]; }
        { auto v = items(r.next_data(), 4); b.dh = v[0]; b.xmin = v[1]; b.ymin = v[2]; b.zmin = v[3]; }
        { auto v = items(r.next_data(), 6); for (int k = 0; k < 6; k++) b.BndConds[k] = (int)v[k]; }
        { auto v = items(r.next_data(), 10); for (int k = 0; k < 10; k++) b.params[k] = v[k]; }
        if (ib < nblock - 1) r.readequal();
        if (b.xDim > 32767 || b.yDim > 32767 || b.zDim > 32767)
            throw std::runtime_error("Grid number exceeds 32767, please try to reduced the grid size.");
        in.blocks.push_back(b);
    }

    // read_probe_params, FlowCondition.f90:79-112
    r.rewind();
    r.found_keyword("ProbingFluid");
    { auto v = items(r.next_data(), 2); f.fluidProbingNum = (int)v[0]; f.inWhichBlock = (int)v[1]; }
    for (int i = 0; i < f.fluidProbingNum; i++) { auto v = items(r.next_data(), 3); f.fluidProbingCoords.push_back({v[0], v[1], v[2]}); }
    r.rewind();
    r.found_keyword("ProbingSolid");
    f.solidProbingNum = (int)items(r.next_data(), 1)[0];
    for (int i = 0; i < f.solidProbingNum; i++) f.solidProbingNode.push_back((int)items(r.next_data(), 1)[0]);
    return in;
}

void calculate_reference_params(FlowCond &flow, int nFish)
{
    if (flow.LrefType == 0) {
        if (nFish == 0) flow.Lref = 1.0;   // 'LrefType and nFish is 0, Lref is adjusted to 1', Solidbody.f90:228-230
        else throw std::runtime_error("LrefType 0 with bodies needs the chord length from the structural solver");
    }
    switch (flow.UrefType) {
    case 0: flow.Uref = std::fabs(flow.uvwIn[0]); break;
    case 1: flow.Uref = std::fabs(flow.uvwIn[1]); break;
    case 2: flow.Uref = std::fabs(flow.uvwIn[2]); break;
    case 3: flow.Uref = std::sqrt(flow.uvwIn[0] * flow.uvwIn[0] + flow.uvwIn[1] * flow.uvwIn[1] + flow.uvwIn[2] * flow.uvwIn[2]); break;
    case 4:
        if (flow.velocityKind == 2) flow.Uref = std::fabs(flow.shearRateIn[0]);
        else throw std::runtime_error("oscillatory flow must set velocityKind to 2");
        break;
    case 5: case 6: case 7: throw std::runtime_error("UrefType 5..7 take the reference velocity from body kinematics; no bodies here");
    default: break;   // 'Use input reference velocity'
    }
    if (flow.TrefType == 0) flow.Tref = flow.Lref / flow.Uref;
    else if (flow.TrefType == 1) throw std::runtime_error("TrefType 1 takes the reference time from body kinematics; no bodies here");
    flow.Aref = flow.Uref / flow.Tref;
    flow.nu = flow.Uref * flow.Lref / flow.Re;   // Solidbody.f90:282
    flow.Mu = flow.nu * flow.denIn;
}

}  // namespace harness
