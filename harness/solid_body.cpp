// solid_body.cpp -- see solid_body.hpp.  Citations: /root/reference/src/Solidbody.f90 unless noted.
#include "solid_body.hpp"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "fortran_io.hpp"

namespace harness {

// ---------------------------------------------------------------------------------------------------------------
// type VirtualBody
// ---------------------------------------------------------------------------------------------------------------

void VirtualBody::PlateBuild()
{
    rtov.assign(rbm.nEL + 1, 0);
    v_nelmts = 0;
    for (int i = 0; i < rbm.nEL; i++) {
        rtov[i] = v_nelmts;
        v_nelmts = v_nelmts + rbm.m_elements[i].Nspan;
    }
    rtov[rbm.nEL] = v_nelmts;
    vtor.assign(v_nelmts, 0);
    for (int i = 0; i < rbm.nEL; i++)
        for (int j = rtov[i]; j < rtov[i + 1]; j++) vtor[j] = i;
    v_Exyz.assign(3 * (size_t)v_nelmts, 0.0);
    v_Evel.assign(3 * (size_t)v_nelmts, 0.0);
    v_Eforce.assign(3 * (size_t)v_nelmts, 0.0);
    v_Ea.assign(v_nelmts, 0.0);
}

void VirtualBody::PlateUpdatePosVelArea(double IBPenaltyAlpha, double denIn)
{
    const double IBPenaltyBeta = -IBPenaltyAlpha * 2.0 * denIn;
    for (int i = 0; i < rbm.nEL; i++) {
        const Segment &e = rbm.m_elements[i];
        const int i1 = e.node0, i2 = e.node1;
        double tmpxyz[3], tmpvel[3], omega[3];
        for (int k = 0; k < 3; k++) {
            tmpxyz[k] = 0.5 * (rbm.pos[6 * i1 + k] + rbm.pos[6 * i2 + k]);
            tmpvel[k] = 0.5 * (rbm.vel[6 * i1 + k] + rbm.vel[6 * i2 + k]);
            omega[k] = 0.5 * (rbm.vel[6 * i1 + 3 + k] + rbm.vel[6 * i2 + 3 + k]);
        }
        const double left = e.Lspan, len = e.spanlen;
        const double dl = len / (double)e.Nspan;
        const double dh = e.len1;
        const double area = dl * dh * IBPenaltyBeta;
        double dirc[3];
        for (int k = 0; k < 3; k++) dirc[k] = 0.5 * (e.triad_n1[k][1] + e.triad_n2[k][1]);
        const double dir_norm = std::sqrt(dirc[0] * dirc[0] + dirc[1] * dirc[1] + dirc[2] * dirc[2]);
        if (dir_norm > 1e-12) for (int k = 0; k < 3; k++) dirc[k] = dirc[k] / dir_norm;
        else for (int k = 0; k < 3; k++) dirc[k] = e.triad_ee[k][1];
        const int cnt = rtov[i];
        for (int s = 1; s <= e.Nspan; s++) {
            const double ls = dl * (0.5 + (double)(s - 1)) - left;
            const double rspan[3] = {dirc[0] * ls, dirc[1] * ls, dirc[2] * ls};
            const double wspin[3] = {omega[1] * rspan[2] - omega[2] * rspan[1], omega[2] * rspan[0] - omega[0] * rspan[2],
                                     omega[0] * rspan[1] - omega[1] * rspan[0]};
            const size_t m = (size_t)(cnt + s - 1);
            for (int k = 0; k < 3; k++) {
                v_Exyz[3 * m + k] = tmpxyz[k] + rspan[k];
                v_Evel[3 * m + k] = tmpvel[k] + wspin[k];
            }
            v_Ea[m] = area;
        }
    }
}

void VirtualBody::UpdatePosVelArea(double IBPenaltyAlpha, double denIn)
{
    if (v_type == 1 && (v_move == 1 || rbm.iBodyModel == 2)) PlateUpdatePosVelArea(IBPenaltyAlpha, denIn);
}

void VirtualBody::NodalLoads()
{
    for (int iEL = 0; iEL < v_nelmts; iEL++) {
        const double *F = &v_Eforce[3 * (size_t)iEL];
        const Segment &e = rbm.m_elements[vtor[iEL]];
        double rr[3];
        for (int k = 0; k < 3; k++) rr[k] = v_Exyz[3 * (size_t)iEL + k] - 0.5 * (e.x1[k] + e.x1[6 + k]);
        const double M[3] = {rr[1] * F[2] - rr[2] * F[1], rr[2] * F[0] - rr[0] * F[2], rr[0] * F[1] - rr[1] * F[0]};
        for (int k = 0; k < 3; k++) {
            rbm.lodFlow[e.m_localToGlobal[k]] = rbm.lodFlow[e.m_localToGlobal[k]] + 0.5 * F[k];
            rbm.lodFlow[e.m_localToGlobal[6 + k]] = rbm.lodFlow[e.m_localToGlobal[6 + k]] + 0.5 * F[k];
            rbm.lodFlow[e.m_localToGlobal[3 + k]] = rbm.lodFlow[e.m_localToGlobal[3 + k]] + 0.5 * M[k];
            rbm.lodFlow[e.m_localToGlobal[9 + k]] = rbm.lodFlow[e.m_localToGlobal[9 + k]] + 0.5 * M[k];
        }
    }
}

void VirtualBody::PlateWrite_body(int iFish, FILE *fh, double Lref) const
{
    const int nEL = rbm.nEL, nSta = nEL + 2;
    std::fprintf(fh, "ZONE    T = \"fish%s\" N = %s, E = %s, DATAPACKING=POINT, ZONETYPE=FEQUADRILATERAL\n", fmtI(iFish, 4, 4).c_str(),
                 fmtI(2 * nSta, 7).c_str(), fmtI(nSta - 1, 7).c_str());
    const auto &el = rbm.m_elements;
    auto centre = [&](int i, double xc[3]) { for (int k = 0; k < 3; k++) xc[k] = 0.5 * (el[i].x1[k] + el[i].x1[6 + k]); };
    auto middir = [&](int i, double d[3]) { for (int k = 0; k < 3; k++) d[k] = 0.5 * (el[i].triad_n1[k][1] + el[i].triad_n2[k][1]); };
    auto write_section = [&](const double xc[3], double Ls, double Rs, const double dir[3]) {
        double d[3] = {dir[0], dir[1], dir[2]};
        const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (nd > 1e-14) for (double &v : d) v = v / nd;
        std::fprintf(fh, "%s%s%s\n", fmtE((xc[0] - Ls * d[0]) / Lref, 20, 10).c_str(), fmtE((xc[1] - Ls * d[1]) / Lref, 20, 10).c_str(),
                     fmtE((xc[2] - Ls * d[2]) / Lref, 20, 10).c_str());
        std::fprintf(fh, "%s%s%s\n", fmtE((xc[0] + Rs * d[0]) / Lref, 20, 10).c_str(), fmtE((xc[1] + Rs * d[1]) / Lref, 20, 10).c_str(),
                     fmtE((xc[2] + Rs * d[2]) / Lref, 20, 10).c_str());
    };
    auto extrapolated = [&](int a, int b, bool first) {   // a: the end element, b: its neighbour
        double xc[3], xc1[3], xc2[3], d1[3], d2[3], dir[3];
        if (nEL >= 2) {
            centre(a, xc1); centre(b, xc2);
            for (int k = 0; k < 3; k++) xc[k] = first ? xc1[k] - 0.5 * (xc2[k] - xc1[k]) : xc1[k] + 0.5 * (xc1[k] - xc2[k]);
            const double Ls = 1.5 * el[a].Lspan - 0.5 * el[b].Lspan;
            const double Rs = 1.5 * (el[a].spanlen - el[a].Lspan) - 0.5 * (el[b].spanlen - el[b].Lspan);
            middir(a, d1); middir(b, d2);
            for (int k = 0; k < 3; k++) dir[k] = 1.5 * d1[k] - 0.5 * d2[k];
            write_section(xc, Ls, Rs, dir);
        } else {
            for (int k = 0; k < 3; k++) { xc[k] = first ? el[0].x1[k] : el[0].x1[6 + k]; dir[k] = first ? el[0].triad_n1[k][1] : el[0].triad_n2[k][1]; }
            write_section(xc, el[0].Lspan, el[0].spanlen - el[0].Lspan, dir);
        }
    };
    extrapolated(0, nEL >= 2 ? 1 : 0, true);
    for (int i = 0; i < nEL; i++) {
        double xc[3], dir[3];
        centre(i, xc); middir(i, dir);
        write_section(xc, el[i].Lspan, el[i].spanlen - el[i].Lspan, dir);
    }
    extrapolated(nEL - 1, nEL >= 2 ? nEL - 2 : 0, false);
    for (int i = 1; i <= nSta - 1; i++)
        std::fprintf(fh, "%s%s%s%s\n", fmtI(2 * i - 1, 8).c_str(), fmtI(2 * i, 8).c_str(), fmtI(2 * (i + 1), 8).c_str(), fmtI(2 * (i + 1) - 1, 8).c_str());
}

void VirtualBody::Write_force(int iFish, FILE *fh, double Lref, double Fref) const
{
    std::fprintf(fh, "ZONE    T = \"fish%s\" I = %s, DATAPACKING=POINT\n", fmtI(iFish, 4, 4).c_str(), fmtI(v_nelmts, 8).c_str());
    const double invLFref = 1.0 / (Lref * Fref);
    const double xyz0[3] = {rbm.pos[0], rbm.pos[1], rbm.pos[2]};
    for (int i = 0; i < v_nelmts; i++) {
        const double *xyz = &v_Exyz[3 * (size_t)i], *F = &v_Eforce[3 * (size_t)i];
        const double r[3] = {xyz[0] - xyz0[0], xyz[1] - xyz0[1], xyz[2] - xyz0[2]};
        const double M[3] = {r[1] * F[2] - r[2] * F[1], r[2] * F[0] - r[0] * F[2], r[0] * F[1] - r[1] * F[0]};
        std::string row;
        for (int k = 0; k < 3; k++) row += fmtE(xyz[k] / Lref, 20, 10);
        for (int k = 0; k < 3; k++) row += fmtE(F[k] / Fref, 20, 10);
        for (int k = 0; k < 3; k++) row += fmtE(M[k] * invLFref, 20, 10);
        std::fprintf(fh, "%s\n", row.c_str());
    }
}

// ---------------------------------------------------------------------------------------------------------------
// module SolidBody
// ---------------------------------------------------------------------------------------------------------------

void SolidBodies::read_solid_files(const InFlow &in, const Vec3 &g)
{
    const SolidHeader &h = in.solid;
    m_IBPenaltyAlpha = h.IBPenaltyAlpha;
    m_nFish = h.nFish;
    m_nGroup = h.nGroup;
    P.dampK = h.dampK; P.dampM = h.dampM; P.GeoGamma = h.GeoGamma; P.NewmarkGamma = h.NewmarkGamma; P.NewmarkBeta = h.NewmarkBeta;
    P.dtolFEM = h.dtolFEM; P.ntolFEM = h.ntolFEM; P.isKB = h.isKB; P.g = g;   // Set_SolidSolver_Params, SolidSolver.f90:1249
    m_fishNum.assign(m_nGroup + 1, 0);
    m_numX.assign(m_nGroup, 1); m_numY.assign(m_nGroup, 1); m_numZ.assign(m_nGroup, 1);
    m_XYZo.assign(m_nFish, Vec3{});
    VBodies.assign(m_nFish, VirtualBody());
    m_fishNum[0] = 1;
    int order1 = 0, order2 = 0;
    for (int ig = 0; ig < m_nGroup; ig++) {
        const SolidGroup &G = in.groups[ig];
        m_fishNum[ig + 1] = G.fishNum; m_numX[ig] = G.numX; m_numY[ig] = G.numY; m_numZ[ig] = G.numZ;
        order1 = order1 + m_fishNum[ig];
        order2 = order2 + m_fishNum[ig + 1];
        if (order2 > m_nFish) throw std::runtime_error("SolidBody section: the group sizes add up to more bodies than nFish");
        for (int iFish = order1; iFish <= order2; iFish++) {   // 1-based, as :165
            VirtualBody &B = VBodies[iFish - 1];
            B.v_type = G.iBodyType;
            if (G.iBodyType == -1) throw std::runtime_error("iBodyType = -1 (gmsh surface body) is not provided by this stand-in driver");
            BeamSolver &r = B.rbm;   // Beam_SetSolver, SolidSolver.f90:1219
            r.P = &P;
            r.FEmeshName = G.FEmeshName; r.iBodyModel = G.iBodyModel;
            for (int k = 0; k < 6; k++) r.isMotionGiven[k] = G.isMotionGiven[k];
            r.denR = G.denR; r.psR = G.psR; r.EmR = G.EmR; r.tcR = G.tcR; r.KB = G.KB; r.KS = G.KS; r.Freq = G.freq; r.St = G.St;
            const int order3 = iFish - order1;
            const int lineX = order3 % G.numX, lineY = (order3 / G.numX) % G.numY, lineZ = order3 / (G.numX * G.numY);
            m_XYZo[iFish - 1] = {G.firstXYZ[0] + G.deltaXYZ[0] * (double)lineX, G.firstXYZ[1] + G.deltaXYZ[1] * (double)lineY,
                                 G.firstXYZ[2] + G.deltaXYZ[2] * (double)lineZ};
            r.XYZo = m_XYZo[iFish - 1];
            r.initXYZVel = G.initXYZVel; r.XYZAmpl = G.XYZAmpl; r.XYZPhi = G.XYZPhi; r.AoAo = G.AoAo; r.AoAAmpl = G.AoAAmpl; r.AoAPhi = G.AoAPhi;
        }
    }
}

void SolidBodies::allocate_solid_memory(FlowCond &flow)
{
    std::vector<double> nAsfac(m_nFish, 0.0), nLchod(m_nFish, 0.0);
    std::printf("=========================================================\n");
    for (int iFish = 0; iFish < m_nFish; iFish++) {
        VirtualBody &B = VBodies[iFish];
        const BeamSolver &r = B.rbm;
        auto asum = [](const Vec3 &v) { return std::fabs(v[0]) + std::fabs(v[1]) + std::fabs(v[2]); };
        int given = 0;
        for (int k = 0; k < 6; k++) given += r.isMotionGiven[k];
        if (asum(r.initXYZVel) > 1e-5 || asum(r.XYZAmpl) > 1e-5 || asum(r.AoAAmpl) > 1e-5 || given < 6) B.v_move = 1;
        B.rbm.P = &P;   // VBodies may have been reallocated since read_solid_files
        B.rbm.ReadBuild(nAsfac[iFish], nLchod[iFish]);
        std::printf(" read FEMeshFile %11d end,  isMoving: %11d\n", iFish + 1, B.v_move);
    }
    std::printf("=========================================================\n");
    if (m_nFish > 0) {
        const int maxN = (int)(std::max_element(nAsfac.begin(), nAsfac.end()) - nAsfac.begin());
        flow.Asfac = nAsfac[maxN];
        flow.Lchod = nLchod[maxN];
        const BeamSolver &r = VBodies[maxN].rbm;
        if (VBodies[maxN].v_type == 1) {
            double s = 0.0;
            for (const Segment &e : r.m_elements) s += e.spanlen;
            flow.Lspan = s / (double)r.nEL;
        } else {
            double lo = 1e300, hi = -1e300;
            for (const Segment &e : r.m_elements) { lo = std::min({lo, e.x00[2], e.x00[8]}); hi = std::max({hi, e.x00[2], e.x00[8]}); }
            flow.Lspan = hi - lo;
        }
        if ((flow.Lchod - 1.0) <= 1e-2) flow.Lchod = 1.0;   // as written at :313-314 (no abs)
        if ((flow.Lspan - 1.0) <= 1e-2) flow.Lspan = 1.0;
        flow.AR = VBodies[maxN].v_type == 1 ? flow.Lspan * flow.Lspan / flow.Asfac : 1.0;
    } else {
        flow.Asfac = 0.0; flow.Lchod = 0.0; flow.Lspan = 0.0; flow.AR = 0.0;
    }
}

void SolidBodies::calculate_reference_params(FlowCond &flow) const
{
    const double pi = 3.141592653589793;   // ConstParams.f90:36
    if (flow.LrefType == 0) {
        if (m_nFish == 0) { flow.Lref = 1.0; std::printf(" LrefType and nFish is 0, Lref is adjusted to 1\n"); }
        else flow.Lref = flow.Lchod;
    } else std::printf(" Use input reference length\n");
    auto maxFreq = [&]() { double f = -1e300; for (const VirtualBody &B : VBodies) f = std::max(f, B.rbm.Freq); return f; };
    auto nUref = [&](double factor) {
        double u = -1e300;
        for (const VirtualBody &B : VBodies) {
            const Vec3 &a = B.rbm.XYZAmpl;
            u = std::max(u, 2.0 * pi * B.rbm.Freq * std::max({std::fabs(a[0]), std::fabs(a[1]), std::fabs(a[2])}) * factor);
        }
        return u;
    };
    switch (flow.UrefType) {
    case 0: flow.Uref = std::fabs(flow.uvwIn[0]); break;
    case 1: flow.Uref = std::fabs(flow.uvwIn[1]); break;
    case 2: flow.Uref = std::fabs(flow.uvwIn[2]); break;
    case 3: flow.Uref = std::sqrt(flow.uvwIn[0] * flow.uvwIn[0] + flow.uvwIn[1] * flow.uvwIn[1] + flow.uvwIn[2] * flow.uvwIn[2]); break;
    case 4:
        if (flow.velocityKind == 2) flow.Uref = std::fabs(flow.shearRateIn[0]);
        else throw std::runtime_error("oscillatory flow must set velocityKind to 2");
        break;
    case 5: flow.Uref = flow.Lref * maxFreq(); break;
    case 6: flow.Uref = nUref(1.0); break;
    case 7: flow.Uref = nUref(2.0); break;   // Park 2017 pof
    default: std::printf(" Use input reference velocity\n"); break;
    }
    if (flow.TrefType == 0) flow.Tref = flow.Lref / flow.Uref;
    else if (flow.TrefType == 1) flow.Tref = 1 / maxFreq();
    else std::printf(" Use input reference time\n");
    flow.Aref = flow.Uref / flow.Tref;
    // Uref**2 is a primary of the product chain in the Fortran (Solidbody.f90:278-280): (0.5*denIn)*(Uref*Uref)*..., not ((0.5*denIn)*Uref)*Uref
    flow.Fref = 0.5 * flow.denIn * (flow.Uref * flow.Uref) * flow.Asfac;
    flow.Eref = 0.5 * flow.denIn * (flow.Uref * flow.Uref) * flow.Asfac * flow.Lref;
    flow.Pref = 0.5 * flow.denIn * (flow.Uref * flow.Uref) * flow.Asfac * flow.Uref;
    flow.nu = flow.Uref * flow.Lref / flow.Re;
    flow.Mu = flow.nu * flow.denIn;
}

void SolidBodies::set_solidbody_parameters(const FlowCond &flow, const int BndConds[6])
{
    m_denIn = flow.denIn;
    m_uvwIn = {flow.uvwIn[0], flow.uvwIn[1], flow.uvwIn[2]};
    for (int k = 0; k < 6; k++) m_boundaryConditions[k] = BndConds[k];
    m_Aref = flow.Aref; m_Eref = flow.Eref; m_Fref = flow.Fref; m_Lref = flow.Lref; m_Pref = flow.Pref; m_Tref = flow.Tref; m_Uref = flow.Uref;
    m_ntolLBM = flow.ntolLBM;
    m_dtolLBM = flow.dtolLBM;
    // Calculate_Solid_params, :374-384
    double uMax = 0.0, nLthck = 0.0;
    for (VirtualBody &B : VBodies) B.rbm.calculate_angle_material(m_Lref, m_Uref, m_denIn, uMax, m_uvwIn, nLthck);
}

void SolidBodies::Initialise_solid_bodies(double time)
{
    for (VirtualBody &B : VBodies) {
        B.rbm.P = &P;
        B.rbm.Initialise(time);
        if (B.v_type == 1) {   // Initialise_, :360-372
            B.PlateBuild();
            B.PlateUpdatePosVelArea(m_IBPenaltyAlpha, m_denIn);
        } else throw std::runtime_error("not implemented body type");
    }
}

namespace {
// Persistent workers for the loops over bodies (the reference's !$OMP PARALLEL DO SCHEDULE(DYNAMIC), Solidbody.f90:392):
// the beams are independent of one another, and a step has several such loops, so the threads are kept between calls.
class BodyPool {
public:
    ~BodyPool()
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : workers_) t.join();
    }
    // fn(i) for i in [0, n), dynamically scheduled over at most `nthreads` threads (the caller is one of them)
    void run(size_t n, size_t nthreads, const std::function<void(size_t)> &fn)
    {
        nthreads = std::min(nthreads, n);
        if (nthreads <= 1) {
            for (size_t i = 0; i < n; i++) fn(i);
            return;
        }
        std::lock_guard<std::mutex> serial(run_mutex_);
        {
            std::lock_guard<std::mutex> lock(m_);
            while (workers_.size() + 1 < nthreads) workers_.emplace_back([this] { loop(); });
            fn_ = &fn; n_ = n; next_.store(0);
            wanted_ = nthreads - 1; busy_ = wanted_;
            gen_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lock(m_);
        done_.wait(lock, [this] { return busy_ == 0; });
        fn_ = nullptr;
    }

private:
    void work()
    {
        for (size_t i = next_++; i < n_; i = next_++) (*fn_)(i);
    }
    void loop()
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [&] { return stop_ || (gen_ != seen && wanted_ > 0); });
                if (stop_) return;
                seen = gen_;
                wanted_--;
            }
            work();
            {
                std::lock_guard<std::mutex> lock(m_);
                busy_--;
            }
            done_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_, run_mutex_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t n_ = 0, wanted_ = 0, busy_ = 0;
    std::atomic<size_t> next_{0};
    unsigned long long gen_ = 0;
    bool stop_ = false;
};
BodyPool g_pool;

size_t body_threads()
{
    size_t nt = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("FSILBM_SOLID_THREADS")) nt = (size_t)std::max(1, std::atoi(e));
    return nt < 1 ? 1 : nt;
}

void over_bodies(const std::vector<int> &bodies, const std::function<void(int)> &per_body)
{
    std::mutex err_mutex;
    std::string err;
    g_pool.run(bodies.size(), body_threads(), [&](size_t i) {
        try {
            per_body(bodies[i]);
        } catch (const std::exception &ex) {
            std::lock_guard<std::mutex> lock(err_mutex);
            if (err.empty()) err = ex.what();
        }
    });
    if (!err.empty()) throw std::runtime_error(err);
}
}  // namespace

void SolidBodies::Solver(const std::vector<int> &bodies, double time, int isubstep, double deltat, double subdeltat)
{
    over_bodies(bodies, [&](int iFish) { VBodies[iFish].rbm.structure(iFish + 1, time, isubstep, deltat, subdeltat); });
}

// The host work of one step for the listed bodies, body by body in parallel: the nodal-load half of FluidVolumeForce_
// (lodFlow = 0, Solidbody.f90:911; loads :945-967) from the v_Eforce just received, the numsubstep structural sub-steps of
// FSInteraction_force's caller (LBMBlockComm.f90:333-335) and UpdatePosVelArea_ (:729) for the next step's markers.
// Per body this is exactly the sequence the driver issues one call at a time.
void SolidBodies::Advance(const std::vector<int> &bodies, double time, int numsubstep, double deltat)
{
    const double subdeltat = deltat / (double)numsubstep;
    over_bodies(bodies, [&](int iFish) {
        VirtualBody &B = VBodies[iFish];
        B.count_Interp = 1;
        std::fill(B.rbm.lodFlow.begin(), B.rbm.lodFlow.end(), 0.0);
        B.NodalLoads();
        for (int isub = 1; isub <= numsubstep; isub++) B.rbm.structure(iFish + 1, time, isub, deltat, subdeltat);
        B.UpdatePosVelArea(m_IBPenaltyAlpha, m_denIn);
    });
}

void SolidBodies::write_solid_field(double time) const
{
    const std::string name = "./DatBody/Bodies_" + time_stamp10(time, m_Tref) + ".dat";
    FILE *fh = std::fopen(name.c_str(), "w");
    if (!fh) return;
    std::fprintf(fh, "TITLE = \"ASCII File.\"\n");
    std::fprintf(fh, "VARIABLES = \"x\" \"y\" \"z\" \"u\" \"v\" \"w\" \"ax\" \"ay\" \"az\" \"fxi\" \"fyi\" \"fzi\" \"fxr\" \"fyr\" \"fzr\"\n");
    for (int iFish = 0; iFish < m_nFish; iFish++) VBodies[iFish].rbm.write_solid(m_Lref, m_Uref, m_Aref, m_Fref, iFish + 1, fh);
    std::fclose(fh);
}

void SolidBodies::Write_solid_v_bodies(double time) const
{
    const std::string name = "./DatBodySpan/BodiesVirtual_" + time_stamp10(time, m_Tref) + ".dat";
    FILE *fh = std::fopen(name.c_str(), "w");
    if (!fh) return;
    std::fprintf(fh, "TITLE = \"ASCII File.\"\n");
    std::fprintf(fh, "VARIABLES = \"x\" \"y\" \"z\"\n");
    for (int iFish = 0; iFish < m_nFish; iFish++) VBodies[iFish].PlateWrite_body(iFish + 1, fh, m_Lref);
    std::fclose(fh);
}

void SolidBodies::Write_solid_v_forces(double time) const
{
    const std::string name = "./DatBodySpan/ForcesVirtual_" + time_stamp10(time, m_Tref) + ".dat";
    FILE *fh = std::fopen(name.c_str(), "w");
    if (!fh) return;
    std::fprintf(fh, "TITLE = \"ASCII File.\"\n");
    std::fprintf(fh, "VARIABLES = \"x\" \"y\" \"z\" \"fx\" \"fy\" \"fz\" \"Mx\" \"My\" \"Mz\"\n");
    for (int iFish = 0; iFish < m_nFish; iFish++) VBodies[iFish].Write_force(iFish + 1, fh, m_Lref, m_Fref);
    std::fclose(fh);
}

void SolidBodies::Write_solid_Check(const std::string &filename) const
{
    FILE *fh = std::fopen(filename.c_str(), "a");
    if (!fh) return;
    for (int iFish = 0; iFish < m_nFish; iFish++) {
        std::fprintf(fh, "============================= nFish = %s ==============================\n", fmtI(iFish + 1, 4, 4).c_str());
        std::fprintf(fh, "inWhichBlock : %s\n", fmtI(VBodies[iFish].v_carrierFluidId + 1, 4, 4).c_str());
        std::fprintf(fh, "---------------------------------------------------------------------------\n");
        VBodies[iFish].rbm.write_solid_params(fh);
        VBodies[iFish].rbm.write_solid_materials(fh);
    }
    std::fprintf(fh, "====================================================================\n");
    std::fclose(fh);
}

namespace {
std::string group_name(int iGroup) { return fmtI(iGroup, 3, 3); }
}  // namespace

void SolidBodies::write_solid_Information(double time, const std::vector<int> &solidProbingNode)
{
    const std::string timeName = time_stamp10(time, m_Tref);
    int order1 = 0, order2 = 0;
    for (int ig = 0; ig < m_nGroup; ig++) {
        const std::string groupNum = group_name(ig + 1), base = "./DatInfo/Group" + groupNum;
        std::vector<std::string> files = {"_forces.dat", "_firstNode.dat", "_lastNode.dat", "_centerNode.dat", "_nodeAverage.dat", "_power.dat", "_energy.dat"};
        for (size_t j = 0; j < solidProbingNode.size(); j++) files.push_back("_solidProbes_" + fmtI((long long)j + 1, 4, 4) + ".dat");
        for (const std::string &f : files) {
            FILE *fh = std::fopen((base + f).c_str(), "a");
            if (!fh) continue;
            std::fprintf(fh, " ZONE T = \"time%s\", I = %d, J = %d, K = %d, f = point\n", timeName.c_str(), m_numX[ig], m_numY[ig], m_numZ[ig]);
            std::fclose(fh);
        }
        order1 = order1 + m_fishNum[ig];
        order2 = order2 + m_fishNum[ig + 1];
        for (int iFish = order1; iFish <= order2; iFish++) {
            BeamSolver &r = VBodies[iFish - 1].rbm;
            r.write_solid_info(groupNum, m_XYZo[iFish - 1], m_Lref, m_Uref, m_Aref, m_Fref, m_Pref, m_Eref);
            r.write_solid_probes(groupNum, m_XYZo[iFish - 1], solidProbingNode, m_Lref, m_Uref, m_Aref);
        }
    }
}

void SolidBodies::write_information_titles(int nGroup, const FlowCond &flow)
{
    auto title = [](const std::string &file, const char *vars) {
        FILE *fh = std::fopen(file.c_str(), "w");
        if (!fh) return;
        std::fprintf(fh, " %s\n", vars);   // list-directed write: one leading blank
        std::fclose(fh);
    };
    const char *node15 = "VARIABLES = \"x\"  \"y\"  \"z\"  \"dx\"  \"dy\"  \"dz\"  \"rx\"  \"ry\"  \"rz\"  \"u\"  \"v\"  \"w\"  \"ax\"  \"ay\"  \"az\"";
    for (int ig = 1; ig <= nGroup; ig++) {
        const std::string base = "./DatInfo/Group" + group_name(ig);
        title(base + "_firstNode.dat", node15);
        title(base + "_lastNode.dat", node15);
        title(base + "_centerNode.dat", node15);
        title(base + "_nodeAverage.dat", "VARIABLES = \"x\"  \"y\"  \"z\"  \"dx\"  \"dy\"  \"dz\"  \"u\"  \"v\"  \"w\"  \"ax\"  \"ay\"  \"az\"");
        title(base + "_forces.dat", "VARIABLES = \"x\"  \"y\"  \"z\"  \"Fx\"  \"Fy\"  \"Fz\"");
        title(base + "_power.dat", "VARIABLES = \"x\"  \"y\"  \"z\"  \"Ptot\"  \"Px\"  \"Py\"  \"Pz\"");
        title(base + "_energy.dat", "VARIABLES = \"x\"  \"y\"  \"z\"  \"Etot\"  \"Evel\"  \"Ep\"  \"Es\"  \"Eb\"");
        for (int j = 1; j <= flow.solidProbingNum; j++) title(base + "_solidProbes_" + fmtI(j, 4, 4) + ".dat", node15);
    }
}

}  // namespace harness
