// fsilbm_harness.cpp -- C++ stand-in for the reference's Fortran driver (PROGRAM main, main.f90:13-151), linked
// against libfsilbm_b200.so through include/fsilbm.h only.
//
// Why it exists: the drop-in boundary is a C ABI meant for the Fortran driver, and no Fortran compiler exists in
// the build image.  This program replays the driver's call sequence call for call -- same inFlow.dat,
// same block tree (LBMBlockComm.f90:98-211), same output cadence (main.f90:115-141), same DatFlow / DatContinue /
// DatInfo byte formats, same FIELDSTAT lines -- so the whole path from parameter file to result files can be
// exercised.  Structural bodies (SolidBody section with nFish > 0, plates given by a line mesh) run on the C++
// restatement of the beam solver in beam_solver.cpp / solid_body.cpp, which stands where SolidSolver.f90 and the host
// half of Solidbody.f90 stand in the real driver; only marker arrays cross the C ABI.  Citations: /root/reference/src.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "../include/fsilbm.h"
#include "inflow.hpp"
#include "solid_body.hpp"

using harness::BlockSpec;
using harness::FlowCond;

namespace {

void ck(int rc)
{
    if (rc != 0) {   // the reference's convention: write(*,*) msg ; stop
        std::printf(" %s\n", fsilbm_last_error());
        std::exit(1);
    }
}

struct Blk {
    BlockSpec spec;
    fsilbm_handle h = -1;
    double xmax = 0, ymax = 0, zmax = 0;   // FluidDomain.f90:94-105
    int periodic[3] = {0, 0, 0};
    double blktime = 0;
};

struct Node {   // blockTreeNode, LBMBlockComm.f90:19-25
    int fatherId = -1;
    std::vector<int> sons;
    std::vector<int> comm;   // pair handles
};

std::vector<Blk> g_blk;
std::vector<Node> g_tree;
int g_root = -1;
harness::SolidBodies g_solid;                 // module SolidBody (VBodies and its m_* variables)
std::vector<std::vector<int>> g_carried;      // LBMblks(:)%carriedBodies, 0-based body indices per block
int g_numsubstep = 1;
long long g_iterLBM_total = 0;
const double MachineTolerace = 1.0e-12;   // ConstParams.f90:36

// CompareBlocks, FluidDomain.f90:1845-1972
int CompareBlocks(int i, int j)
{
    const Blk &bi = g_blk[i], &bj = g_blk[j];
    double vi[6] = {bi.spec.xmin, bi.xmax, bi.spec.ymin, bi.ymax, bi.spec.zmin, bi.zmax};
    double vj[6] = {bj.spec.xmin, bj.xmax, bj.spec.ymin, bj.ymax, bj.spec.zmin, bj.zmax};
    auto adjust = [](const Blk &a, double *va, const Blk &b, double *vb) {   // :1868-1902
        int nfluid = 0;
        for (int p = 0; p < 6; p++) nfluid += a.spec.BndConds[p] == 0;
        if (nfluid != 1) return;
        for (int p = 0; p < 6; p++) {
            if (a.spec.BndConds[p] != 0) continue;
            if (p % 2 == 0) { if (b.spec.BndConds[p + 1] == 1 && std::fabs(vb[p + 1] - va[p] - b.spec.dh) < MachineTolerace) va[p + 1] = va[p]; }
            else { if (b.spec.BndConds[p - 1] == 1 && std::fabs(va[p] - vb[p - 1] - b.spec.dh) < MachineTolerace) va[p - 1] = va[p]; }
        }
    };
    adjust(bi, vi, bj, vj);
    adjust(bj, vj, bi, vi);
    int cnt = 0, align = 0;
    for (int k = 0; k < 3; k++) {
        const int lo = 2 * k, hi = 2 * k + 1;
        const bool d1 = vi[lo] < vj[lo] || std::fabs(vi[lo] - vj[lo]) < MachineTolerace;
        const bool d2 = vj[hi] < vi[hi] || std::fabs(vj[hi] - vi[hi]) < MachineTolerace;
        const bool d3 = vj[lo] < vi[lo] || std::fabs(vj[lo] - vi[lo]) < MachineTolerace;
        const bool d4 = vi[hi] < vj[hi] || std::fabs(vi[hi] - vj[hi]) < MachineTolerace;
        const bool d5 = vi[hi] < vj[lo], d6 = vj[hi] < vi[lo];
        if (d1 && d2 && !(d3 && d4)) cnt++;
        else if (d3 && d4 && !(d1 && d2)) cnt--;
        else if (d1 && d2 && d3 && d4) align++;
        else if (d5 || d6) { std::printf(" Notice, blocks no overlaps %d %d\n", bi.spec.ID, bj.spec.ID); return 0; }
    }
    if (align > 0) { if (cnt < 0) cnt -= align; if (cnt > 0) cnt += align; }
    if (cnt == 3) return 1;
    if (cnt == -3) return -1;
    std::printf(" Warning, blocks partial overlaps %d %d\n", bi.spec.ID, bj.spec.ID);
    return 0;
}

// array_to_tree, LBMBlockComm.f90:135-193
void array_to_tree(const std::vector<int> &iblocks, int rootnode, int interpolateScheme)
{
    const int nb = (int)iblocks.size();
    if (nb == 0) return;
    std::vector<int> fa(nb, -1);
    for (int i = 0; i < nb - 1; i++)
        for (int j = i + 1; j < nb; j++) {
            const int t = CompareBlocks(iblocks[i], iblocks[j]);
            if (t == 1) fa[j] = i; else if (t == -1) fa[i] = j;
        }
    for (bool changed = true; changed;) {
        changed = false;
        for (int i = 0; i < nb; i++)
            if (fa[i] >= 0 && fa[fa[i]] >= 0 && fa[i] != fa[fa[i]]) { fa[i] = fa[fa[i]]; changed = true; }
    }
    for (int i = 0; i < nb; i++) {
        if (fa[i] >= 0) continue;
        const int r = iblocks[i];
        g_tree[rootnode].sons.push_back(r);
        g_tree[r].fatherId = rootnode;
        int pair = -1;
        ck(fsilbm_pair_create(g_blk[rootnode].h, g_blk[r].h, interpolateScheme, &pair));   // build_blocks_comunication :32-96
        g_tree[rootnode].comm.push_back(pair);
        std::vector<int> sub;
        for (int j = 0; j < nb; j++) if (fa[j] >= 0 && iblocks[fa[j]] == r) sub.push_back(iblocks[j]);
        array_to_tree(sub, r, interpolateScheme);
    }
}

// build_block_tree :195-211 with findremove_blockTreeRoot :98-133
void build_block_tree(int interpolateScheme)
{
    const int nb = (int)g_blk.size();
    g_tree.assign(nb, Node());
    std::vector<int> fa(nb, -1);
    for (int i = 0; i < nb - 1; i++)
        for (int j = i + 1; j < nb; j++) {
            const int t = CompareBlocks(i, j);
            if (t == 1) fa[j] = i; else if (t == -1) fa[i] = j;
        }
    int cr = 0;
    std::vector<int> rest;
    for (int i = 0; i < nb; i++) { if (fa[i] < 0) { cr++; g_root = i; } else rest.push_back(i); }
    if (cr > 1) { std::printf(" Error: there exist more than one block tree root\n"); std::exit(1); }
    array_to_tree(rest, g_root, interpolateScheme);
}

void tree_set_boundary_conditions_block(int node)   // LBMBlockComm.f90:266-277
{
    ck(fsilbm_block_set_boundary_conditions(g_blk[node].h));
    for (int s : g_tree[node].sons) tree_set_boundary_conditions_block(s);
}

// find_carrier_fluidblock + FindCarrierFluidBlock, FluidDomain.f90:1974-2017
void FindCarrierFluidBlock()
{
    g_carried.assign(g_blk.size(), std::vector<int>());
    for (int iFish = 0; iFish < g_solid.m_nFish; iFish++) {
        harness::VirtualBody &B = g_solid.VBodies[iFish];
        const double *x = B.v_Exyz.data();   // first marker
        double dh = 1e10;
        int n = -1;
        for (size_t i = 0; i < g_blk.size(); i++) {
            const Blk &b = g_blk[i];
            if (b.spec.xmin <= x[0] && x[0] <= b.xmax && b.spec.ymin <= x[1] && x[1] <= b.ymax && b.spec.zmin <= x[2] && x[2] <= b.zmax && b.spec.dh < dh) {
                dh = b.spec.dh;
                n = (int)i;
            }
        }
        if (n == -1) throw std::runtime_error("Error: carrier fluid block not found");
        B.v_carrierFluidId = n;
        g_carried[n].push_back(iFish);
    }
}

// IBM_FEM, LBMBlockComm.f90:320-338, in two halves: FSInteraction_force (Solidbody.f90:589-602; the device half is one
// library call) and the numsubstep structural sub-steps.  The reference runs them back to back before the collision;
// the sub-steps only advance the beams with the loads just computed and touch no fluid state, so the driver issues them
// after the (asynchronous) collide-stream launch and the host structural solve overlaps the device work of the step.
void IBM_FEM_solver(int node, double time)
{
    const std::vector<int> &bodies = g_carried[node];
    if (bodies.empty()) return;
    const double dh = g_blk[node].spec.dh, dt_solid = dh / (double)g_numsubstep;
    for (int isubstep = 1; isubstep <= g_numsubstep; isubstep++) g_solid.Solver(bodies, time, isubstep, dh, dt_solid);   // LBMBlockComm.f90:333-335
}

void IBM_FEM_force(int node)
{
    const std::vector<int> &bodies = g_carried[node];
    if (bodies.empty()) return;
    Blk &b = g_blk[node];
    const double dh = b.spec.dh;
    std::vector<int> nelmts, restencil;
    std::vector<const double *> Exyz, Evel, Ea;
    std::vector<double *> Eforce;
    for (int iFish : bodies) {
        harness::VirtualBody &B = g_solid.VBodies[iFish];
        B.UpdatePosVelArea(g_solid.m_IBPenaltyAlpha, g_solid.m_denIn);                                  // Solidbody.f90:599
        nelmts.push_back(B.v_nelmts);
        restencil.push_back(B.v_move == 1 || B.rbm.iBodyModel == 2 || B.count_Interp == 0 ? 1 : 0);     // :885
        Exyz.push_back(B.v_Exyz.data()); Evel.push_back(B.v_Evel.data()); Ea.push_back(B.v_Ea.data()); Eforce.push_back(B.v_Eforce.data());
    }
    int iterLBM = 0;
    ck(fsilbm_ibm_interaction_force(b.h, (int)bodies.size(), nelmts.data(), Exyz.data(), Evel.data(), Ea.data(), Eforce.data(), restencil.data(),
                                    dh, g_solid.m_ntolLBM, g_solid.m_dtolLBM, g_solid.m_boundaryConditions, &iterLBM));   // :869-918 on the device
    g_iterLBM_total += iterLBM;
    for (int iFish : bodies) {
        harness::VirtualBody &B = g_solid.VBodies[iFish];
        B.count_Interp = 1;
        std::fill(B.rbm.lodFlow.begin(), B.rbm.lodFlow.end(), 0.0);                                     // :911
    }
    for (int iFish : bodies) g_solid.VBodies[iFish].NodalLoads();                                       // :945-967
}

// tree_collision_streaming_IBM_FEM, LBMBlockComm.f90:279-318
void tree_collision_streaming_IBM_FEM(int node)
{
    Blk &b = g_blk[node];
    ck(fsilbm_block_set_time(b.h, b.blktime));
    ck(fsilbm_block_update_volume_force(b.h, nullptr));                         // :283
    IBM_FEM_force(node);                                                        // :287 (macro :285 and reset :286 happen on the device)
    for (int p : g_tree[node].comm) ck(fsilbm_pair_extract_layer(p, 1));        // :290
    ck(fsilbm_block_collide_stream(b.h));                                       // :285-303
    IBM_FEM_solver(node, b.blktime);                                            // :333-335, overlapping the launch above
    for (int p : g_tree[node].comm) ck(fsilbm_pair_extract_layer(p, 2));        // :305
    for (size_t i = 0; i < g_tree[node].sons.size(); i++) {                     // :307-317
        const int s = g_tree[node].sons[i];
        for (int n = 0; n < 2; n++) {
            g_blk[s].blktime = g_blk[s].blktime + (double)n * g_blk[s].spec.dh;
            tree_collision_streaming_IBM_FEM(s);
            ck(fsilbm_pair_father_to_son(g_tree[node].comm[i], n));
        }
        ck(fsilbm_pair_son_to_father(g_tree[node].comm[i]));
    }
}

long nint(double v) { return (long)(v < 0 ? -std::floor(-v + 0.5) : std::floor(v + 0.5)); }
std::string name10(double v)   // write(fileName,'(I10)') nint(v*1d5), blanks -> '0'
{
    char buf[32];
    std::snprintf(buf, sizeof(buf), "%10ld", nint(v * 1e5));
    for (char *c = buf; *c; c++) if (*c == ' ') *c = '0';
    return buf;
}

// write_flow_blocks, FluidDomain.f90:349-366 -> write_flow_ :1628-1737 (written in-process; no fork needed: the
// staging buffer is plain host memory)
void write_flow_blocks(double time, const FlowCond &flow)
{
    for (Blk &b : g_blk) {
        const BlockSpec &s = b.spec;
        if (s.outputtype < 1) continue;
        const int o = s.offsetOutput;
        const int nx = s.xDim - 2 * o, ny = s.yDim - 2 * o, nz = s.zDim - 2 * o;
        const int nf = s.outputtype >= 2 ? 13 : 4;
        std::vector<float> out((size_t)nf * nx * ny * nz);
        ck(fsilbm_block_write_flow_window(b.h, o, s.outputtype, out.data()));
        const int head_i[4] = {nx, ny, nz, s.ID};
        const double head_d[4] = {s.xmin + o * s.dh, s.ymin + o * s.dh, s.zmin + o * s.dh, s.dh};
        char bname[8];
        std::snprintf(bname, sizeof(bname), "%03d", s.ID);
        const size_t n = (size_t)nx * ny * nz;
        if (s.outputtype != 2) {
            const std::string path = "./DatFlow/Flow" + name10(time / flow.Tref) + "_b" + bname;
            FILE *fh = std::fopen(path.c_str(), "wb");
            if (!fh) throw std::runtime_error("cannot write " + path);
            std::fwrite(head_i, sizeof(int), 4, fh);
            std::fwrite(head_d, sizeof(double), 4, fh);
            std::fwrite(out.data(), sizeof(float), 4 * n, fh);
            std::fclose(fh);
        }
        if (s.outputtype >= 2) {
            const std::string path = std::string("./DatFlow/MeanFlow_b") + bname;
            FILE *fh = std::fopen(path.c_str(), "wb");
            if (!fh) throw std::runtime_error("cannot write " + path);
            std::fwrite(head_i, sizeof(int), 4, fh);
            std::fwrite(head_d, sizeof(double), 4, fh);
            std::fwrite(out.data(), sizeof(float), n, fh);
            std::fwrite(out.data() + 4 * n, sizeof(float), 9 * n, fh);
            std::fclose(fh);
        }
    }
}

// write_continue_blocks, FluidDomain.f90:268-285
void write_continue_blocks(int step, double time)
{
    const std::string path = "./DatContinue/continue" + name10(time);
    FILE *fh = std::fopen(path.c_str(), "wb");
    if (!fh) throw std::runtime_error("cannot write " + path);
    const int nblocks = (int)g_blk.size();
    std::fwrite(&nblocks, sizeof(int), 1, fh);
    std::fwrite(&step, sizeof(int), 1, fh);
    std::fwrite(&time, sizeof(double), 1, fh);
    for (Blk &b : g_blk) {
        const BlockSpec &s = b.spec;
        const double geo[4] = {s.xmin, s.ymin, s.zmin, s.dh};
        const int dims[3] = {s.xDim, s.yDim, s.zDim};
        std::vector<double> f((size_t)19 * s.xDim * s.yDim * s.zDim);
        ck(fsilbm_block_download_fIn(b.h, f.data()));
        std::fwrite(geo, sizeof(double), 4, fh);
        std::fwrite(dims, sizeof(int), 3, fh);
        std::fwrite(f.data(), sizeof(double), f.size(), fh);
    }
    std::fclose(fh);
}

// check_is_continue, FluidDomain.f90:128-237
bool check_is_continue(int &step, double &time, int isContinue)
{
    FILE *fh = isContinue >= 1 ? std::fopen("./DatContinue/continue", "rb") : nullptr;
    if (!fh) {
        if (isContinue >= 1) std::printf(" Warning: the continue file is not found in DatContinue!\n");
        std::printf("====================== New computing ====================\n");
        return false;
    }
    std::printf("=================== Continue computing ==================\n");
    struct Saved { double xmin, ymin, zmin, dh; int X, Y, Z; std::vector<double> f; };
    int nblocks = 0;
    bool ok = std::fread(&nblocks, sizeof(int), 1, fh) == 1 && std::fread(&step, sizeof(int), 1, fh) == 1 && std::fread(&time, sizeof(double), 1, fh) == 1;
    std::vector<Saved> sv(ok ? nblocks : 0);
    for (Saved &s : sv) {
        double geo[4]; int dims[3];
        ok = ok && std::fread(geo, sizeof(double), 4, fh) == 4 && std::fread(dims, sizeof(int), 3, fh) == 3;
        if (!ok) break;
        s.xmin = geo[0]; s.ymin = geo[1]; s.zmin = geo[2]; s.dh = geo[3]; s.X = dims[0]; s.Y = dims[1]; s.Z = dims[2];
        s.f.resize((size_t)19 * s.X * s.Y * s.Z);
        ok = ok && std::fread(s.f.data(), sizeof(double), s.f.size(), fh) == s.f.size();
    }
    std::fclose(fh);
    if (!ok) throw std::runtime_error("./DatContinue/continue is truncated");
    std::vector<int> sortdh(nblocks);
    for (int i = 0; i < nblocks; i++) sortdh[i] = i;
    for (int i = 0; i < nblocks - 1; i++)   // :156-165 (as written there: compares the UNSORTED dh of slots i and j)
        for (int j = i + 1; j < nblocks; j++)
            if (sv[i].dh > sv[j].dh) std::swap(sortdh[i], sortdh[j]);
    for (Blk &b : g_blk) {
        const BlockSpec &s = b.spec;
        const size_t ncell = (size_t)s.xDim * s.yDim * s.zDim;
        std::vector<double> f(19 * ncell);
        ck(fsilbm_block_download_fIn(b.h, f.data()));
        for (int x = 0; x < s.xDim; x++) {
            const double xC = s.xmin + x * s.dh;
            for (int y = 0; y < s.yDim; y++) {
                const double yC = s.ymin + y * s.dh;
                for (int z = 0; z < s.zDim; z++) {
                    const double zC = s.zmin + z * s.dh;
                    for (int j = 0; j < nblocks; j++) {
                        const Saved &t = sv[sortdh[j]];
                        const double mx = t.xmin + (t.X - 1) * t.dh, my = t.ymin + (t.Y - 1) * t.dh, mz = t.zmin + (t.Z - 1) * t.dh;
                        if (!(zC >= t.zmin && zC <= mz && yC >= t.ymin && yC <= my && xC >= t.xmin && xC <= mx)) continue;
                        double c1 = (xC - t.xmin) / t.dh, c2 = (yC - t.ymin) / t.dh, c3 = (zC - t.zmin) / t.dh;
                        int i1 = (int)std::floor(c1), i3 = (int)std::floor(c2), i5 = (int)std::floor(c3);
                        if (i1 == t.X - 1) i1--;
                        if (i3 == t.Y - 1) i3--;
                        if (i5 == t.Z - 1) i5--;
                        c1 -= i1; c2 -= i3; c3 -= i5;
                        auto at = [&](int q, int xx, int yy, int zz) { return t.f[((size_t)q * t.X + xx) * t.Y * t.Z + (size_t)yy * t.Z + zz]; };
                        for (int q = 0; q < 19; q++)
                            f[q * ncell + ((size_t)x * s.yDim + y) * s.zDim + z] =
                                at(q, i1, i3, i5) * (1 - c3) * (1 - c2) * (1 - c1) + at(q, i1 + 1, i3, i5) * (1 - c3) * (1 - c2) * c1 +
                                at(q, i1, i3 + 1, i5) * (1 - c3) * c2 * (1 - c1) + at(q, i1 + 1, i3 + 1, i5) * (1 - c3) * c2 * c1 +
                                at(q, i1, i3, i5 + 1) * c3 * (1 - c2) * (1 - c1) + at(q, i1 + 1, i3, i5 + 1) * c3 * (1 - c2) * c1 +
                                at(q, i1, i3 + 1, i5 + 1) * c3 * c2 * (1 - c1) + at(q, i1 + 1, i3 + 1, i5 + 1) * c3 * c2 * c1;
                        break;
                    }
                }
            }
        }
        ck(fsilbm_block_upload_fIn(b.h, f.data()));
    }
    return true;
}

std::string e20_10(double v)   // Fortran E20.10: 0.dddddddddd E+xx, ten significant digits correctly rounded from the binary value
{
    if (v == 0.0) return "    0.0000000000E+00";
    char m[64];
    std::snprintf(m, sizeof(m), "%.9E", std::fabs(v));          // d.dddddddddE+xx, exact decimal rounding by the C library
    std::string digits;
    digits += m[0];
    digits.append(m + 2, 9);
    const int ex = std::atoi(std::strchr(m, 'E') + 1) + 1;
    char out[96];
    std::snprintf(out, sizeof(out), "%s0.%sE%+03d", v < 0 ? "-" : "", digits.c_str(), ex);
    std::string s(out);
    return std::string(s.size() < 20 ? 20 - s.size() : 0, ' ') + s;
}

void write_fluid_flux(int root, double time, const FlowCond &flow)   // FluidDomain.f90:2019-2056
{
    double raw[3];
    ck(fsilbm_block_fluid_flux(g_blk[root].h, raw));
    const Blk &b = g_blk[root];
    const double Yref = b.ymax - b.spec.ymin, Zref = b.zmax - b.spec.zmin, d = flow.denIn * flow.Uref * Zref * Yref;
    FILE *fh = std::fopen("./DatInfo/FluidFlux.dat", "a");
    if (!fh) return;
    std::fprintf(fh, "%s%s%s%s\n", e20_10(time / flow.Tref).c_str(), e20_10(raw[0] / d).c_str(), e20_10(raw[1] / d).c_str(), e20_10(raw[2] / d).c_str());
    std::fclose(fh);
}

void write_fluid_information(double time, const FlowCond &flow)   // FlowCondition.f90:195-222
{
    const int n = flow.fluidProbingNum;
    if (n <= 0) return;
    std::vector<double> co(3 * n), vel(3 * n);
    for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) co[3 * i + k] = flow.fluidProbingCoords[i][k];
    ck(fsilbm_block_probe_velocity(g_blk[flow.inWhichBlock - 1].h, n, co.data(), vel.data()));
    for (int i = 0; i < n; i++) {
        char path[64];
        std::snprintf(path, sizeof(path), "./DatInfo/FluidProbes_%04d.dat", i + 1);
        FILE *fh = std::fopen(path, "a");
        if (!fh) continue;
        std::fprintf(fh, "%s%s%s%s\n", e20_10(time / flow.Tref).c_str(), e20_10(vel[3 * i] / flow.Uref).c_str(), e20_10(vel[3 * i + 1] / flow.Uref).c_str(),
                     e20_10(vel[3 * i + 2] / flow.Uref).c_str());
        std::fclose(fh);
    }
}

void computeFieldStat_blocks()   // FluidDomain.f90:368-374 -> :1739-1768
{
    static const char *names[6] = {"L2 u", "L2 v", "L2 w", "Linfinity u", "Linfinity v", "Linfinity w"};
    for (Blk &b : g_blk) {
        double st[6];
        ck(fsilbm_block_field_stat(b.h, st));
        const double n = (double)b.spec.xDim * (double)b.spec.yDim * (double)b.spec.zDim;
        for (int k = 0; k < 3; k++) st[k] = std::sqrt(st[k] / n);
        for (int k = 0; k < 6; k++) std::printf(" FIELDSTAT %s %18.12f\n", names[k], st[k]);
    }
}

bool on_cadence(double tT, double delta, double half)   // DABS(t - delta*NINT(t/delta)) <= 0.5*dt/Tref, main.f90:117,124,129,134
{
    return std::fabs(tT - delta * (double)nint(tT / delta)) <= half;
}

}  // namespace

int main(int argc, char **argv)
{
    std::string parameterFile = "inFlow.dat";
    bool parse_only = false;
    int device = 0;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--parse-only")) parse_only = true;
        else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
        else parameterFile = argv[i];
    }
    try {
        harness::InFlow in = harness::read_inflow(parameterFile);
        FlowCond &flow = in.flow;
        g_solid.read_solid_files(in, harness::Vec3{0.0, 0.0, 0.0});               // main.f90:30 (g = 0, :25)
        g_solid.allocate_solid_memory(flow);                                       // main.f90:39 (reads the structural mesh files)
        g_solid.calculate_reference_params(flow);                                  // main.f90:43
        g_numsubstep = flow.numsubstep;
        if (parse_only) {
            std::printf("{\"npsize\": %d, \"isConCmpt\": %d, \"numsubstep\": %d, \"timeSimTotal\": %.17g, \"Re\": %.17g, \"denIn\": %.17g, "
                        "\"uvwIn\": [%.17g, %.17g, %.17g], \"velocityKind\": %d, \"Lref\": %.17g, \"Tref\": %.17g, \"Uref\": %.17g, \"nu\": %.17g, "
                        "\"ntolLBM\": %d, \"dtolLBM\": %.17g, \"interpolateScheme\": %d, \"nFish\": %d, \"nblocks\": %zu, \"fluidProbingNum\": %d, \"blocks\": [",
                        flow.npsize, flow.isConCmpt, flow.numsubstep, flow.timeSimTotal, flow.Re, flow.denIn, flow.uvwIn[0], flow.uvwIn[1], flow.uvwIn[2],
                        flow.velocityKind, flow.Lref, flow.Tref, flow.Uref, flow.nu, flow.ntolLBM, flow.dtolLBM, flow.interpolateScheme, in.solid.nFish,
                        in.blocks.size(), flow.fluidProbingNum);
            for (size_t i = 0; i < in.blocks.size(); i++) {
                const BlockSpec &b = in.blocks[i];
                std::printf("%s{\"ID\": %d, \"iCollidModel\": %d, \"dims\": [%d, %d, %d], \"dh\": %.17g, \"BndConds\": [%d, %d, %d, %d, %d, %d], \"params1\": %.17g}",
                            i ? ", " : "", b.ID, b.iCollidModel, b.xDim, b.yDim, b.zDim, b.dh, b.BndConds[0], b.BndConds[1], b.BndConds[2], b.BndConds[3],
                            b.BndConds[4], b.BndConds[5], b.params[0]);
            }
            std::printf("]}\n");
            return 0;
        }
        mkdir("./DatFlow", 0755); mkdir("./DatContinue", 0755); mkdir("./DatInfo", 0755);
        if (in.solid.nFish > 0) { mkdir("./DatBody", 0755); mkdir("./DatBodySpan", 0755); }

        ck(fsilbm_init(device));                                                   // main.f90:36 (omp_set_num_threads)
        fsilbm_flow cf{};
        cf.nu = flow.nu; cf.denIn = flow.denIn; cf.velocityKind = flow.velocityKind;
        for (int k = 0; k < 3; k++) { cf.uvwIn[k] = flow.uvwIn[k]; cf.shearRateIn[k] = flow.shearRateIn[k]; cf.volumeForceIn[k] = flow.volumeForceIn[k]; }
        cf.volumeForceAmp = flow.volumeForceAmp; cf.volumeForceFreq = flow.volumeForceFreq; cf.volumeForcePhi = flow.volumeForcePhi; cf.Uref = flow.Uref;
        for (const BlockSpec &s : in.blocks) {                                     // allocate_fuild_memory_blocks, main.f90:40
            Blk b;
            b.spec = s;
            for (int i = 0; i < 3; i++) b.periodic[i] = (s.BndConds[2 * i] == 301 && s.BndConds[2 * i + 1] == 301) ? 1 : 0;
            b.xmax = s.xmin + s.dh * (s.xDim - 1) + (b.periodic[0] ? s.dh : 0.0);
            b.ymax = s.ymin + s.dh * (s.yDim - 1) + (b.periodic[1] ? s.dh : 0.0);
            b.zmax = s.zmin + s.dh * (s.zDim - 1) + (b.periodic[2] ? s.dh : 0.0);
            ck(fsilbm_block_create(s.xDim, s.yDim, s.zDim, 0, s.xDim, s.dh, s.xmin, s.ymin, s.zmin, s.BndConds.data(), s.iCollidModel, s.params.data(), &cf, &b.h));
            g_blk.push_back(b);
        }
        double time = 0.0, start_time = 0.0;
        int step = 0;
        build_block_tree(flow.interpolateScheme);                                  // main.f90:33 (after the blocks exist here)
        g_solid.set_solidbody_parameters(flow, g_blk[g_root].spec.BndConds.data()); // main.f90:44-45
        g_solid.Initialise_solid_bodies(0.0);                                      // main.f90:48
        FindCarrierFluidBlock();                                                   // main.f90:49
        for (Blk &b : g_blk) ck(fsilbm_block_initialise(b.h, time));               // main.f90:50
        g_solid.Write_solid_Check("Check.dat");                                    // main.f90:55 (the fluid part of Check.dat is not reproduced)
        check_is_continue(step, start_time, flow.isConCmpt);                       // main.f90:58
        harness::SolidBodies::write_information_titles(g_solid.m_nGroup, flow);    // main.f90:59
        {
            FILE *fh = std::fopen("./DatInfo/FluidFlux.dat", "w");
            if (fh) { std::fprintf(fh, " VARIABLES = \"t\"  \"inlet\"  \"middle\"  \"outlet\"\n"); std::fclose(fh); }
            for (int i = 1; i <= flow.fluidProbingNum; i++) {
                char name[64];
                std::snprintf(name, sizeof(name), "./DatInfo/FluidProbes_%04d.dat", i);
                fh = std::fopen(name, "w");
                if (fh) { std::fprintf(fh, " VARIABLES = \"t\"  \"u\"  \"v\"  \"w\" \n"); std::fclose(fh); }
            }
        }
        for (Blk &b : g_blk) ck(fsilbm_block_update_volume_force(b.h, nullptr));   // main.f90:62
        tree_set_boundary_conditions_block(g_root);                                // main.f90:63
        const double dt_fluid = g_blk[g_root].spec.dh;                             // main.f90:67
        if (flow.isConCmpt == 1) time = start_time * flow.Tref; else { start_time = 0.0; step = 0; }
        int start_ave = step;
        if (flow.timeWriteBegin >= start_time) start_ave = step + (int)nint((flow.timeWriteBegin - start_time) * flow.Tref / dt_fluid);
        std::printf(" the start step for fluid averaging(if used): %d\n", start_ave);
        write_flow_blocks(time, flow);                                             // main.f90:84
        if (g_solid.m_nFish > 0) {                                                 // main.f90:85-87
            g_solid.write_solid_field(time);
            g_solid.Write_solid_v_bodies(time);
            g_solid.Write_solid_v_forces(time);
        }
        std::printf(" Time loop beginning\n");
        const double half = 0.5 * dt_fluid / flow.Tref;
        while (time / flow.Tref < flow.timeSimTotal) {                             // main.f90:93
            time = time + dt_fluid;
            step = step + 1;
            for (Blk &b : g_blk) b.blktime = time;                                 // :97
            tree_collision_streaming_IBM_FEM(g_root);                              // :106
            for (Blk &b : g_blk)                                                   // :108 (macro of :107 is implicit on the device)
                if (b.spec.outputtype >= 2) ck(fsilbm_block_turbulent_statistic(b.h, step, start_ave));
            const double tT = time / flow.Tref;
            if (on_cadence(tT, flow.timeContiDelta, half)) write_continue_blocks(step, tT);          // :117-119
            if (tT - flow.timeWriteBegin >= -half && tT - flow.timeWriteEnd <= half) {               // :121
                if (g_solid.m_nFish > 0 && on_cadence(tT, flow.timeBodyDelta, half)) {               // :124-128
                    g_solid.write_solid_field(time);
                    g_solid.Write_solid_v_bodies(time);
                    g_solid.Write_solid_v_forces(time);
                }
                if (on_cadence(tT, flow.timeFlowDelta, half)) write_flow_blocks(time, flow);         // :129-131
            }
            if (on_cadence(tT, flow.timeInfoDelta, half)) {                                          // :134
                write_fluid_flux(g_root, time, flow);
                if (flow.inWhichBlock >= 1 && flow.inWhichBlock <= (int)g_blk.size()) write_fluid_information(time, flow);
                else std::printf("Warning: invalid flow%%inWhichBlock = %d; it must be in [1, %zu].\n", flow.inWhichBlock, g_blk.size());
                g_solid.write_solid_Information(time, flow.solidProbingNode);                        // :143
            }
        }
        std::printf(" Steps:%8d  Time/Tref:%14.8f\n", step, time / flow.Tref);
        std::printf("=========================================================\n");
        computeFieldStat_blocks();                                                 // main.f90:150
        if (g_solid.m_nFish > 0) {   // not in the reference: a machine-readable end state of every body for the parity tests
            FILE *fh = std::fopen("./DatInfo/BodiesFinal.txt", "w");
            if (fh) {
                for (int iFish = 0; iFish < g_solid.m_nFish; iFish++) {
                    const harness::BeamSolver &r = g_solid.VBodies[iFish].rbm;
                    std::fprintf(fh, "BODY %d nND %d iterFEM %.0f dnorm %.17g cg_iterations %lld\n", iFish + 1, r.nND, r.FishInfo[1], r.FishInfo[2], r.cg_iterations);
                    for (int n = 0; n < r.nND; n++) {
                        std::fprintf(fh, "NODE %d", n + 1);
                        for (int k = 0; k < 6; k++) std::fprintf(fh, " %.17g", r.pos[6 * n + k]);
                        for (int k = 0; k < 6; k++) std::fprintf(fh, " %.17g", r.vel[6 * n + k]);
                        for (int k = 0; k < 6; k++) std::fprintf(fh, " %.17g", r.lodFlow[6 * n + k]);
                        std::fprintf(fh, "\n");
                    }
                }
                std::fclose(fh);
            }
            std::printf(" IBM iterations (total): %lld\n", g_iterLBM_total);
        }
        std::printf(" kernel launches: %lld\n", fsilbm_launch_count());
        fsilbm_finalize();
    } catch (const std::exception &e) {
        std::printf(" %s\n", e.what());
        return 1;
    }
    return 0;
}
