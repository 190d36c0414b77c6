// solid_capi.cpp -- see solid_capi.h.
#include "solid_capi.h"

#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

#include "solid_body.hpp"

namespace {
struct Ctx {
    harness::InFlow in;
    harness::SolidBodies solid;
};
std::map<int, std::unique_ptr<Ctx>> g_ctx;
int g_next = 1;
std::string g_err;

template <class F> int guarded(F &&f)
{
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}
Ctx &ctx(int h)
{
    auto it = g_ctx.find(h);
    if (it == g_ctx.end()) throw std::runtime_error("unknown structural handle");
    return *it->second;
}
harness::VirtualBody &body(int h, int b)
{
    Ctx &c = ctx(h);
    if (b < 0 || b >= c.solid.m_nFish) throw std::runtime_error("body index out of range");
    return c.solid.VBodies[b];
}
}  // namespace

extern "C" {

const char *fsolid_last_error(void) { return g_err.c_str(); }

int fsolid_open(const char *inflow_path, const int rootBC[6], int *handle)
{
    return guarded([&] {
        auto c = std::make_unique<Ctx>();
        c->in = harness::read_inflow(inflow_path);
        c->solid.read_solid_files(c->in, harness::Vec3{0.0, 0.0, 0.0});
        c->solid.allocate_solid_memory(c->in.flow);
        c->solid.calculate_reference_params(c->in.flow);
        c->solid.set_solidbody_parameters(c->in.flow, rootBC);
        c->solid.Initialise_solid_bodies(0.0);
        *handle = g_next++;
        g_ctx[*handle] = std::move(c);
    });
}

int fsolid_close(int handle) { return guarded([&] { ctx(handle); g_ctx.erase(handle); }); }

int fsolid_flow(int handle, double out[16])
{
    return guarded([&] {
        const harness::FlowCond &f = ctx(handle).in.flow;
        const double v[16] = {f.Lref, f.Uref, f.Tref, f.Aref, f.Fref, f.Eref, f.Pref, f.nu, f.Asfac, f.Lchod, f.Lspan, f.AR, f.denIn,
                              (double)f.ntolLBM, f.dtolLBM, (double)f.numsubstep};
        std::memcpy(out, v, sizeof(v));
    });
}

int fsolid_nfish(int handle)
{
    auto it = g_ctx.find(handle);
    return it == g_ctx.end() ? -1 : it->second->solid.m_nFish;
}

int fsolid_body_info(int handle, int b, int out[8])
{
    return guarded([&] {
        const harness::VirtualBody &B = body(handle, b);
        const int v[8] = {B.rbm.nND, B.rbm.nEL, B.v_nelmts, B.v_move, B.rbm.iBodyModel, B.rbm.gEQ, B.count_Interp, B.v_type};
        std::memcpy(out, v, sizeof(v));
    });
}

int fsolid_update_pos_vel_area(int handle, int b)
{
    return guarded([&] { Ctx &c = ctx(handle); body(handle, b).UpdatePosVelArea(c.solid.m_IBPenaltyAlpha, c.solid.m_denIn); });
}

int fsolid_markers(int handle, int b, double *Exyz, double *Evel, double *Ea)
{
    return guarded([&] {
        const harness::VirtualBody &B = body(handle, b);
        if (Exyz) std::memcpy(Exyz, B.v_Exyz.data(), B.v_Exyz.size() * sizeof(double));
        if (Evel) std::memcpy(Evel, B.v_Evel.data(), B.v_Evel.size() * sizeof(double));
        if (Ea) std::memcpy(Ea, B.v_Ea.data(), B.v_Ea.size() * sizeof(double));
    });
}

int fsolid_set_eforce(int handle, int b, const double *Eforce)
{
    return guarded([&] { harness::VirtualBody &B = body(handle, b); std::memcpy(B.v_Eforce.data(), Eforce, B.v_Eforce.size() * sizeof(double)); });
}

int fsolid_fluid_loads(int handle, int b)
{
    return guarded([&] {
        harness::VirtualBody &B = body(handle, b);
        B.count_Interp = 1;
        std::fill(B.rbm.lodFlow.begin(), B.rbm.lodFlow.end(), 0.0);
        B.NodalLoads();
    });
}

int fsolid_set_lodflow(int handle, int b, const double *lodFlow)
{
    return guarded([&] { harness::VirtualBody &B = body(handle, b); std::memcpy(B.rbm.lodFlow.data(), lodFlow, B.rbm.lodFlow.size() * sizeof(double)); });
}

int fsolid_structure(int handle, int b, double time, int isubstep, double deltat, double subdeltat)
{
    return guarded([&] { body(handle, b).rbm.structure(b + 1, time, isubstep, deltat, subdeltat); });
}

int fsolid_solver(int handle, double time, int isubstep, double deltat, double subdeltat)
{
    return guarded([&] {
        Ctx &c = ctx(handle);
        std::vector<int> all(c.solid.m_nFish);
        for (int i = 0; i < c.solid.m_nFish; i++) all[i] = i;
        c.solid.Solver(all, time, isubstep, deltat, subdeltat);
    });
}

int fsolid_marker_ptrs(int handle, int b, double **Exyz, double **Evel, double **Ea, double **Eforce)
{
    return guarded([&] {
        harness::VirtualBody &B = body(handle, b);
        if (Exyz) *Exyz = B.v_Exyz.data();
        if (Evel) *Evel = B.v_Evel.data();
        if (Ea) *Ea = B.v_Ea.data();
        if (Eforce) *Eforce = B.v_Eforce.data();
    });
}

int fsolid_advance(int handle, int nbodies, const int *bodies, double time, int numsubstep, double deltat)
{
    return guarded([&] {
        Ctx &c = ctx(handle);
        std::vector<int> list(bodies, bodies + nbodies);
        for (int b : list) body(handle, b);
        if (numsubstep < 1) throw std::runtime_error("fsolid_advance: numsubstep < 1");
        c.solid.Advance(list, time, numsubstep, deltat);
    });
}

int fsolid_get(int handle, int b, int what, double *out)
{
    return guarded([&] {
        harness::VirtualBody &B = body(handle, b);
        harness::BeamSolver &r = B.rbm;
        auto copy = [&](const std::vector<double> &v) { std::memcpy(out, v.data(), v.size() * sizeof(double)); };
        switch (what) {
        case 0: copy(r.pos); break;
        case 1: copy(r.dsp); break;
        case 2: copy(r.vel); break;
        case 3: copy(r.acc); break;
        case 4: copy(r.lodFlow); break;
        case 5: copy(r.lodInte); break;
        case 6: out[0] = r.FishInfo[0]; out[1] = r.FishInfo[1]; out[2] = r.FishInfo[2]; out[3] = (double)r.cg_iterations; break;
        case 7:
            for (int e = 0; e < r.nEL; e++)
                for (int i = 0; i < 3; i++)
                    for (int j = 0; j < 3; j++) {
                        out[27 * e + 3 * i + j] = r.m_elements[e].triad_ee[i][j];
                        out[27 * e + 9 + 3 * i + j] = r.m_elements[e].triad_n1[i][j];
                        out[27 * e + 18 + 3 * i + j] = r.m_elements[e].triad_n2[i][j];
                    }
            break;
        case 8: copy(r.mss); break;
        case 9:
            r.UpdateStrainEnergy();
            for (int e = 0; e < r.nEL; e++) { out[2 * e] = r.m_elements[e].strainEnergy[0]; out[2 * e + 1] = r.m_elements[e].strainEnergy[1]; }
            break;
        case 10: for (int e = 0; e < r.nEL; e++) for (int k = 0; k < 8; k++) out[8 * e + k] = r.m_elements[e].m_property[k]; break;
        case 11: for (int e = 0; e < r.nEL; e++) for (int k = 0; k < 12; k++) out[12 * e + k] = r.m_elements[e].x1[k]; break;
        case 12: for (int k = 0; k < 3; k++) { out[k] = r.XYZ[k]; out[3 + k] = r.AoA[k]; out[6 + k] = r.UVW[k]; out[9 + k] = r.WWW3[k]; } break;
        case 13: copy(r.lodGrav); break;
        case 14: for (int e = 0; e < r.nEL; e++) { out[3 * e] = r.m_elements[e].len0; out[3 * e + 1] = r.m_elements[e].len1; out[3 * e + 2] = r.m_elements[e].geoFRM; } break;
        default: throw std::runtime_error("fsolid_get: unknown selector");
        }
    });
}

int fsolid_write(int handle, int what, double time)
{
    return guarded([&] {
        Ctx &c = ctx(handle);
        switch (what) {
        case 0: c.solid.write_solid_field(time); break;
        case 1: c.solid.Write_solid_v_bodies(time); break;
        case 2: c.solid.Write_solid_v_forces(time); break;
        case 3: c.solid.write_solid_Information(time, c.in.flow.solidProbingNode); break;
        case 4: harness::SolidBodies::write_information_titles(c.solid.m_nGroup, c.in.flow); break;
        case 5: c.solid.Write_solid_Check("Check.dat"); break;
        default: throw std::runtime_error("fsolid_write: unknown selector");
        }
    });
}

}  // extern "C"
