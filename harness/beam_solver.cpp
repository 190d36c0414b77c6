// beam_solver.cpp -- see beam_solver.hpp.  Citations: /root/reference/src/SolidSolver.f90.
#include "beam_solver.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "fortran_io.hpp"

namespace harness {
namespace {

inline double dot3(const Vec3 &a, const Vec3 &b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

Mat3 matmul3(const Mat3 &a, const Mat3 &b)
{
    Mat3 c{};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += a[i][k] * b[k][j];
            c[i][j] = s;
        }
    return c;
}

Vec3 matvec3(const Mat3 &a, const double *x)
{
    Vec3 y{};
    for (int i = 0; i < 3; i++) y[i] = a[i][0] * x[0] + a[i][1] * x[1] + a[i][2] * x[2];
    return y;
}

// Segment_global_to_local, :954-964
void global_to_local(const Mat3 &triad, double tx, double ty, double tz, double &tx2, double &ty2, double &tz2)
{
    tx2 = triad[0][0] * tx + triad[1][0] * ty + triad[2][0] * tz;
    ty2 = triad[0][1] * tx + triad[1][1] * ty + triad[2][1] * tz;
    tz2 = triad[0][2] * tx + triad[1][2] * ty + triad[2][2] * tz;
}

// the structural data file: list-directed reads (blank / comma separated, Fortran D exponents)
struct DatFile {
    std::vector<std::vector<std::string>> rows;   // tokens of every non-empty line
    size_t pos = 0;
    explicit DatFile(const std::string &name)
    {
        std::ifstream in(name);
        if (!in) throw std::runtime_error("cannot open structural mesh file " + name);
        std::string line;
        while (std::getline(in, line)) {
            for (char &c : line) if (c == ',' || c == '\r' || c == '\t') c = ' ';
            std::istringstream is(line);
            std::vector<std::string> t;
            std::string tok;
            while (is >> tok) t.push_back(tok);
            if (!t.empty()) rows.push_back(t);
        }
    }
    const std::vector<std::string> &next()
    {
        if (pos >= rows.size()) throw std::runtime_error("unexpected end of the structural mesh file");
        return rows[pos++];
    }
    // FindSection, :1352-1369: rewind, then the first line whose first token equals the section name
    void find_section(const std::string &name, const std::string &file)
    {
        for (pos = 0; pos < rows.size();)
            if (rows[pos++][0] == name) return;
        throw std::runtime_error("ERROR: cannot find section " + name + " in " + file);
    }
    static double num(const std::string &tok)
    {
        std::string t = tok;
        for (char &c : t) if (c == 'd' || c == 'D') c = 'e';
        char *end = nullptr;
        const double v = std::strtod(t.c_str(), &end);
        if (end == t.c_str()) throw std::runtime_error("bad numeric item '" + tok + "' in the structural mesh file");
        return v;
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// module SegmentStructure
// ---------------------------------------------------------------------------------------------------------------

void Segment::Build(int p0Id, int p1Id, int itype_, int Nspan_, const std::vector<std::array<double, 8>> &xyz, const std::array<double, 8> &material,
                    const std::vector<std::array<int, 6>> &boundary)
{
    node0 = p0Id;
    node1 = p1Id;
    for (int i = 0; i < 8; i++) m_property[i] = material[i];
    for (int i = 0; i < 6; i++) {
        m_localToGlobal[i] = p0Id * 6 + i;
        m_localToGlobal[i + 6] = p1Id * 6 + i;
    }
    for (int i = 0; i < 12; i++) x00[i] = 0.0;
    for (int i = 0; i < 3; i++) { x00[i] = xyz[p0Id][i]; x00[6 + i] = xyz[p1Id][i]; }
    areaElem00 = std::sqrt((x00[6] - x00[0]) * (x00[6] - x00[0]) + (x00[7] - x00[1]) * (x00[7] - x00[1]) + (x00[8] - x00[2]) * (x00[8] - x00[2]));
    for (int i = 0; i < 6; i++) { bc[i] = boundary[p0Id][i]; bc[i + 6] = boundary[p1Id][i]; }
    itype = itype_;
    Nspan = Nspan_;
    Lspan = 0.5 * (xyz[p0Id][3] + xyz[p1Id][3]);
    spanlen = 0.5 * (xyz[p0Id][4] + xyz[p1Id][4]) + Lspan;
    for (int i = 0; i < 3; i++) dirc00[i] = 0.5 * (xyz[p0Id][5 + i] + xyz[p1Id][5 + i]);
    const double dirc_norm = std::sqrt(dirc00[0] * dirc00[0] + dirc00[1] * dirc00[1] + dirc00[2] * dirc00[2]);
    if (dirc_norm > 1e-10) for (int i = 0; i < 3; i++) dirc00[i] = dirc00[i] / dirc_norm;
    else throw std::runtime_error("the span directions of the two nodes are opposite directions; no unique bisector exists.");
}

void Segment::Init()
{
    for (int i = 0; i < 12; i++) { x1[i] = x0[i]; xnxt[i] = x1[i]; }
    dirc1 = dirc0;
    dircnxt = dirc1;
    dx0 = x0[6] - x0[0];
    dy0 = x0[7] - x0[1];
    dz0 = x0[8] - x0[2];
    len0 = std::sqrt(dx0 * dx0 + dy0 * dy0 + dz0 * dz0);
    xll0 = dx0 / len0;
    xmm0 = dy0 / len0;
    xnn0 = dz0 / len0;
    cptdxyz1();
}

void Segment::cptdxyz1()
{
    dx1 = x1[6] - x1[0];
    dy1 = x1[7] - x1[1];
    dz1 = x1[8] - x1[2];
    len1 = std::sqrt(dx1 * dx1 + dy1 * dy1 + dz1 * dz1);
    xll1 = dx1 / len1;
    xmm1 = dy1 / len1;
    xnn1 = dz1 / len1;
}

void Segment::UpdateMatrix(const double coeffs[8], double gamma, double dampM, double dampK)
{
    FormGeomMatrix();
    RotateMatrix();
    RKR(m_stfMat);
    RKR(m_geoMat);
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 12; j++) {
            m_tanMat[i][j] = m_stfMat[i][j] + gamma * m_geoMat[i][j];
            m_coefMat[i][j] = m_tanMat[i][j] + coeffs[0] * m_masMat[i][j] + coeffs[1] * dampM * m_masMat[i][j];
        }
    if (dampK > 0.0)
        for (int i = 0; i < 12; i++)
            for (int j = 0; j < 12; j++) m_coefMat[i][j] = m_coefMat[i][j] + coeffs[1] * dampK * m_stfMat[i][j];
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 12; j++) m_coefT[j][i] = m_coefMat[i][j];
}

void Segment::UpdateLoad(const double coeffs[8], double dampM, double dampK, const std::vector<double> &dspO, const std::vector<double> &dsp,
                         const std::vector<double> &vel, const std::vector<double> &acc, std::vector<double> &lodEffe) const
{
    double qM[12], qC[12], qMC[12], massLoad[12];
    for (int i = 0; i < 6; i++) {
        const int a = node0 * 6 + i, b = node1 * 6 + i;
        qM[i] = coeffs[0] * (dspO[a] - dsp[a]) + coeffs[2] * vel[a] + coeffs[3] * acc[a];
        qM[i + 6] = coeffs[0] * (dspO[b] - dsp[b]) + coeffs[2] * vel[b] + coeffs[3] * acc[b];
        qC[i] = coeffs[1] * (dspO[a] - dsp[a]) + coeffs[4] * vel[a] + coeffs[5] * acc[a];
        qC[i + 6] = coeffs[1] * (dspO[b] - dsp[b]) + coeffs[4] * vel[b] + coeffs[5] * acc[b];
    }
    for (int i = 0; i < 12; i++) qMC[i] = qM[i] + dampM * qC[i];
    MassMultiply(qMC, massLoad);
    LocToGlobal(massLoad, lodEffe);
    if (dampK > 0.0) {
        double dampKLoad[12];
        for (int i = 0; i < 12; i++) {
            double s = 0.0;
            for (int j = 0; j < 12; j++) s += m_stfMat[i][j] * qC[j];
            dampKLoad[i] = dampK * s;
        }
        LocToGlobal(dampKLoad, lodEffe);
    }
}

void Segment::MassMultiply(const double q[12], double mq[12]) const
{
    for (int i = 0; i < 12; i++) mq[i] = 0.0;
    for (int i = 0; i < 3; i++) {
        mq[i] = m_masMat[i][i] * q[i];
        mq[i + 6] = m_masMat[i + 6][i + 6] * q[i + 6];
    }
    for (int i = 0; i < 3; i++) {
        double s = 0.0, t = 0.0;
        for (int j = 0; j < 3; j++) { s += m_masMat[3 + i][3 + j] * q[3 + j]; t += m_masMat[9 + i][9 + j] * q[9 + j]; }
        mq[3 + i] = s;
        mq[9 + i] = t;
    }
}

void Segment::BoundaryCond(int iter, std::vector<double> &x, std::vector<char> &fixed, const std::vector<double> &vBC) const
{
    for (int i = 0; i < nElmtDofs; i++)
        if (bc[i] > 0) {
            const int gid = m_localToGlobal[i];
            fixed[gid] = 1;
            x[gid] = iter == 1 ? vBC[gid] : 0.0;
        }
}

void Segment::Multiply(const std::vector<double> &x, std::vector<double> &b) const
{
    double lx[12], lb[12];
    for (int i = 0; i < 12; i++) lx[i] = x[m_localToGlobal[i]];
    // row i: ((0 + K(i,1) x(1)) + K(i,2) x(2)) + ... as MATMUL accumulates it; the loops are turned inside out (column by column,
    // all rows at once) so that the twelve independent sums advance together in vector registers -- same additions, same order per row
    for (int i = 0; i < 12; i++) lb[i] = 0.0;
    for (int j = 0; j < 12; j++) {
        const double xj = lx[j];
        const double *col = m_coefT[j].data();
        for (int i = 0; i < 12; i++) lb[i] += col[i] * xj;
    }
    LocToGlobal(lb, b);
}

void Segment::LocToGlobal(const double lx[12], std::vector<double> &x) const
{
    for (int i = 0; i < nElmtDofs; i++) x[m_localToGlobal[i]] = x[m_localToGlobal[i]] + lx[i];
}

void Segment::FormMassMatrix()
{
    const double area = m_property[2], rho = m_property[3], ziy = m_property[6], ziz = m_property[7], length = len0;
    for (auto &r : m_masMat) r.fill(0.0);
    const double roal = rho * area * length / 2.0;
    m_masMat[0][0] = roal;
    m_masMat[1][1] = roal;
    m_masMat[2][2] = roal;
    m_masMat[3][3] = roal * (ziy + ziz) / area;
    m_masMat[4][4] = roal * ziy / area;
    m_masMat[5][5] = roal * ziz / area;
    for (int i = 0; i < 6; i++) m_masMat[6 + i][6 + i] = m_masMat[i][i];
}

namespace {
inline void sym(Mat12 &m, int i, int j, double v) { m[i - 1][j - 1] = v; m[j - 1][i - 1] = v; }   // 1-based, both triangles
}

void Segment::FormStiffMatrix()
{
    const double emod = m_property[0], gmod = m_property[1], area = m_property[2], zix = m_property[5], ziy = m_property[6], ziz = m_property[7];
    const double length = len0;
    for (auto &r : m_stfMat) r.fill(0.0);
    const double Invlength = 1.0 / length;
    const double ksy = 5.0 / 6.0, ksz = 5.0 / 6.0;
    const double phiy = 12.0 * emod * ziz / (ksy * gmod * area * length * length);
    const double phiz = 12.0 * emod * ziy / (ksz * gmod * area * length * length);
    const double ky1 = 12.0 * emod * ziz / (length * length * length * (1.0 + phiy));
    const double ky2 = 6.0 * emod * ziz / (length * length * (1.0 + phiy));
    const double ky3 = (4.0 + phiy) * emod * ziz / (length * (1.0 + phiy));
    const double ky4 = (2.0 - phiy) * emod * ziz / (length * (1.0 + phiy));
    const double kz1 = 12.0 * emod * ziy / (length * length * length * (1.0 + phiz));
    const double kz2 = 6.0 * emod * ziy / (length * length * (1.0 + phiz));
    const double kz3 = (4.0 + phiz) * emod * ziy / (length * (1.0 + phiz));
    const double kz4 = (2.0 - phiz) * emod * ziy / (length * (1.0 + phiz));
    Mat12 &k = m_stfMat;
    k[0][0] = area * emod * Invlength;
    k[1][1] = ky1;
    k[2][2] = kz1;
    k[3][3] = gmod * zix * Invlength;
    k[4][4] = kz3;
    k[5][5] = ky3;
    for (int i = 0; i < 6; i++) k[6 + i][6 + i] = k[i][i];
    sym(k, 1, 7, -k[0][0]);
    sym(k, 2, 6, ky2);
    sym(k, 2, 8, -ky1);
    sym(k, 2, 12, ky2);
    sym(k, 6, 8, -ky2);
    sym(k, 6, 12, ky4);
    sym(k, 8, 12, -ky2);
    sym(k, 3, 5, -kz2);
    sym(k, 3, 9, -kz1);
    sym(k, 3, 11, -kz2);
    sym(k, 5, 9, kz2);
    sym(k, 5, 11, kz4);
    sym(k, 9, 11, kz2);
    sym(k, 4, 10, -k[3][3]);
}

void Segment::FormGeomMatrix()
{
    const double s = geoFRM;
    const double emod = m_property[0], gmod = m_property[1], area = m_property[2], zix = m_property[5], ziy = m_property[6], ziz = m_property[7];
    const double length = len0;
    for (auto &r : m_geoMat) r.fill(0.0);
    const double ksy = 5.0 / 6.0, ksz = 5.0 / 6.0;
    const double phiy = 12.0 * emod * ziz / (ksy * gmod * area * length * length);
    const double phiz = 12.0 * emod * ziy / (ksz * gmod * area * length * length);
    const double L2 = length * length;
    const double dy = (1.0 + phiy) * (1.0 + phiy), dz = (1.0 + phiz) * (1.0 + phiz);
    const double gy1 = s / length * (6.0 / 5.0 + 2.0 * phiy + phiy * phiy) / dy;
    const double gy2 = s / length * (length / 10.0) / dy;
    const double gy3 = s / length * (2.0 * L2 / 15.0 + phiy * L2 / 6.0 + phiy * phiy * L2 / 12.0) / dy;
    const double gy4 = s / length * (-L2 / 30.0 - phiy * L2 / 6.0 - phiy * phiy * L2 / 12.0) / dy;
    const double gz1 = s / length * (6.0 / 5.0 + 2.0 * phiz + phiz * phiz) / dz;
    const double gz2 = s / length * (length / 10.0) / dz;
    const double gz3 = s / length * (2.0 * L2 / 15.0 + phiz * L2 / 6.0 + phiz * phiz * L2 / 12.0) / dz;
    const double gz4 = s / length * (-L2 / 30.0 - phiz * L2 / 6.0 - phiz * phiz * L2 / 12.0) / dz;
    Mat12 &g = m_geoMat;
    g[1][1] = gy1; g[5][5] = gy3; g[7][7] = gy1; g[11][11] = gy3;
    sym(g, 2, 6, gy2);
    sym(g, 2, 8, -gy1);
    sym(g, 2, 12, gy2);
    sym(g, 6, 8, -gy2);
    sym(g, 6, 12, gy4);
    sym(g, 8, 12, -gy2);
    g[2][2] = gz1; g[4][4] = gz3; g[8][8] = gz1; g[10][10] = gz3;
    sym(g, 3, 5, -gz2);
    sym(g, 3, 9, -gz1);
    sym(g, 3, 11, -gz2);
    sym(g, 5, 9, gz2);
    sym(g, 5, 11, gz4);
    sym(g, 9, 11, gz2);
    const double gt = s * zix / (area * length);
    g[3][3] = gt;
    g[9][9] = gt;
    sym(g, 4, 10, -gt);
}

void Segment::InitTriad_D()
{
    BuildAxisDirTriad(xll0, xmm0, xnn0, dirc0, triad_n1);
    triad_n2 = triad_n1;
    triad_ee = triad_n1;
}

void Segment::RigidUpdateTriad_D()
{
    BuildAxisDirTriad(xll1, xmm1, xnn1, dirc1, triad_n1);
    triad_n2 = triad_n1;
    triad_ee = triad_n1;
}

void Segment::BuildAxisDirTriad(double l, double m, double n, const Vec3 &d, Mat3 &triad)
{
    Vec3 ex{l, m, n}, ey{}, ez{}, dir = d;
    double dd = std::sqrt(dot3(ex, ex));
    if (dd > 1e-14) for (double &v : ex) v = v / dd;
    dd = std::sqrt(dot3(dir, dir));
    if (dd > 1e-14) for (double &v : dir) v = v / dd;
    const double proj = dot3(dir, ex);
    for (int i = 0; i < 3; i++) ey[i] = dir[i] - proj * ex[i];
    dd = std::sqrt(dot3(ey, ey));
    if (dd <= 1e-10) {   // no span direction: Doyle's default triad
        if (std::fabs(ex[2]) > 0.995) {
            ey = {0.0, 1.0, 0.0};
            ez = {-ex[2], 0.0, 0.0};
        } else {
            dd = std::sqrt(ex[0] * ex[0] + ex[1] * ex[1]);
            ey = {-ex[1] / dd, ex[0] / dd, 0.0};
            ez = {-ex[0] * ex[2] / dd, -ex[1] * ex[2] / dd, dd};
        }
    } else {
        for (double &v : ey) v = v / dd;
        ez = {ex[1] * ey[2] - ex[2] * ey[1], ex[2] * ey[0] - ex[0] * ey[2], ex[0] * ey[1] - ex[1] * ey[0]};
        dd = std::sqrt(dot3(ez, ez));
        if (dd > 1e-14) for (double &v : ez) v = v / dd;
    }
    for (int i = 0; i < 3; i++) { triad[i][0] = ex[i]; triad[i][1] = ey[i]; triad[i][2] = ez[i]; }
}

void Segment::RotateMatrix()
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) m_rotMat[i][j] = triad_ee[j][i];
}

void Segment::RKR(Mat12 &ek) const
{
    // [R]^T [K] [R] block by block, R = m_rotMat (:717-746)
    double ktemp[12][12];
    for (int i = 0; i <= 3; i++)
        for (int j = 0; j <= 3; j++) {
            const int j1 = i * 3, j2 = j * 3;
            for (int k = 0; k < 3; k++)
                for (int ii = 0; ii < 3; ii++) {
                    double s = 0.0;
                    for (int jj = 0; jj < 3; jj++) s = s + ek[j1 + k][j2 + jj] * m_rotMat[jj][ii];
                    ktemp[j1 + k][j2 + ii] = s;
                }
            for (int k = 0; k < 3; k++)
                for (int ii = 0; ii < 3; ii++) {
                    double s = 0.0;
                    for (int jj = 0; jj < 3; jj++) s = s + m_rotMat[jj][k] * ktemp[j1 + jj][j2 + ii];   // rt(k,jj) = rotMat(jj,k)
                    ek[j1 + k][j2 + ii] = s;
                }
        }
}

void Segment::local_end_rotations(double ub[12]) const
{
    const double du = dx1 - dx0, dv = dy1 - dy0, dw = dz1 - dz0;
    const double dl = ((dx0 + dx1) * du + (dy0 + dy1) * dv + (dz0 + dz1) * dw) / (len0 + len1);
    double tx, ty, tz, tx1, ty1, tz1, tx2, ty2, tz2;
    Segment_get_angle_triad(triad_ee, triad_n1, tx, ty, tz);
    global_to_local(triad_ee, tx, ty, tz, tx1, ty1, tz1);
    Segment_get_angle_triad(triad_ee, triad_n2, tx, ty, tz);
    global_to_local(triad_ee, tx, ty, tz, tx2, ty2, tz2);
    for (int i = 0; i < 12; i++) ub[i] = 0.0;
    ub[3] = tx1; ub[4] = ty1; ub[5] = tz1;
    ub[6] = dl;
    ub[9] = tx2; ub[10] = ty2; ub[11] = tz2;
}

void Segment::BodyStress_D(std::vector<double> &lodInte)
{
    double ub[12], forceb[12], force[12];
    local_end_rotations(ub);
    const double emod = m_property[0], area = m_property[2];
    const double fxx = ub[6] * emod * area / len0;
    geoFRM = fxx;
    for (int i = 0; i < 12; i++) {
        double s = 0.0;
        for (int j = 0; j < 12; j++) s += m_stfMat[i][j] * ub[j];
        forceb[i] = s;
    }
    for (int i = 0; i < 12; i++) force[i] = 0.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            force[0 + i] = force[0 + i] + triad_ee[i][j] * forceb[0 + j];
            force[3 + i] = force[3 + i] + triad_ee[i][j] * forceb[3 + j];
            force[6 + i] = force[6 + i] + triad_ee[i][j] * forceb[6 + j];
            force[9 + i] = force[9 + i] + triad_ee[i][j] * forceb[9 + j];
        }
    LocToGlobal(force, lodInte);
}

void Segment::StrainEnergy_D()
{
    double ub[12];
    local_end_rotations(ub);
    Mat12 Strech{}, BendTor = m_stfMat;
    Strech[0][0] = m_stfMat[0][0]; Strech[0][6] = m_stfMat[0][6]; Strech[6][6] = m_stfMat[6][6]; Strech[6][0] = m_stfMat[6][0];
    BendTor[0][0] = 0.0; BendTor[0][6] = 0.0; BendTor[6][6] = 0.0; BendTor[6][0] = 0.0;
    double es = 0.0, eb = 0.0;
    for (int i = 0; i < 12; i++) {
        double s = 0.0, b = 0.0;
        for (int j = 0; j < 12; j++) { s += Strech[i][j] * ub[j]; b += BendTor[i][j] * ub[j]; }
        es += s * ub[i];
        eb += b * ub[i];
    }
    strainEnergy[0] = 0.5 * es;
    strainEnergy[1] = 0.5 * eb;
}

void Segment_get_angle_triad(const Mat3 &triad_11, const Mat3 &triad_22, double &tx, double &ty, double &tz)
{
    Mat3 rr{};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s = s + triad_22[i][k] * triad_11[j][k];
            rr[i][j] = s;
        }
    const double dtx = (rr[2][1] - rr[1][2]) / 2.0;
    const double dty = (rr[0][2] - rr[2][0]) / 2.0;
    const double dtz = (rr[1][0] - rr[0][1]) / 2.0;
    double trace_rr = rr[0][0] + rr[1][1] + rr[2][2];
    trace_rr = (trace_rr - 1.0) / 2.0;
    if (trace_rr > 1.0) trace_rr = 1.0;
    if (trace_rr < -1.0) trace_rr = -1.0;
    const double sint = std::sqrt(dtx * dtx + dty * dty + dtz * dtz);
    const double theta = std::acos(trace_rr);
    if (sint < 1e-10 || theta < 1e-10) {
        tx = dtx; ty = dty; tz = dtz;
    } else {
        const double factor = theta / sint;
        tx = factor * dtx; ty = factor * dty; tz = factor * dtz;
    }
}

void Segment::UpdateTriad_D(const std::vector<double> &dspnn)
{
    Mat3 rr;
    Segment_FiniteRot(dspnn[node0 * 6 + 3], dspnn[node0 * 6 + 4], dspnn[node0 * 6 + 5], rr);
    triad_n1 = matmul3(rr, triad_n1);
    Segment_FiniteRot(dspnn[node1 * 6 + 3], dspnn[node1 * 6 + 4], dspnn[node1 * 6 + 5], rr);
    triad_n2 = matmul3(rr, triad_n2);
}

void Segment::MakeTriad_ee()
{
    for (auto &r : triad_ee) r.fill(0.0);
    triad_ee[0][0] = xll1;
    triad_ee[1][0] = xmm1;
    triad_ee[2][0] = xnn1;
    double tx, ty, tz;
    Segment_get_angle_triad(triad_n1, triad_n2, tx, ty, tz);
    tx = tx / 2.0; ty = ty / 2.0; tz = tz / 2.0;
    Mat3 rr;
    Segment_FiniteRot(tx, ty, tz, rr);
    const Mat3 triad_aa = matmul3(rr, triad_n1);
    double r2e1 = 0.0, r3e1 = 0.0;
    for (int k = 0; k < 3; k++) {
        r2e1 = r2e1 + triad_aa[k][1] * triad_ee[k][0];
        r3e1 = r3e1 + triad_aa[k][2] * triad_ee[k][0];
    }
    for (int j = 0; j < 3; j++) {
        triad_ee[j][1] = triad_aa[j][1] - r2e1 * (triad_aa[j][0] + triad_ee[j][0]) / 2.0;
        triad_ee[j][2] = triad_aa[j][2] - r3e1 * (triad_aa[j][0] + triad_ee[j][0]) / 2.0;
    }
    double d21 = 0.0;
    for (int k = 0; k < 3; k++) d21 += triad_ee[k][1] * triad_ee[k][0];
    for (int k = 0; k < 3; k++) triad_ee[k][1] = triad_ee[k][1] - d21 * triad_ee[k][0];
    double dd = 0.0;
    for (int k = 0; k < 3; k++) dd += triad_ee[k][1] * triad_ee[k][1];
    dd = std::sqrt(dd);
    if (dd > 1e-14) for (int k = 0; k < 3; k++) triad_ee[k][1] = triad_ee[k][1] / dd;
    triad_ee[0][2] = triad_ee[1][0] * triad_ee[2][1] - triad_ee[2][0] * triad_ee[1][1];
    triad_ee[1][2] = triad_ee[2][0] * triad_ee[0][1] - triad_ee[0][0] * triad_ee[2][1];
    triad_ee[2][2] = triad_ee[0][0] * triad_ee[1][1] - triad_ee[1][0] * triad_ee[0][1];
}

void Segment_FiniteRot(double t1, double t2, double t3, Mat3 &rr)
{
    const double tt = std::sqrt(t1 * t1 + t2 * t2 + t3 * t3);
    const double ss = std::sin(tt), cc = std::cos(tt);
    const Mat3 rr1 = {{{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}}};
    const Mat3 rr2 = {{{0.0, -t3, t2}, {t3, 0.0, -t1}, {-t2, t1, 0.0}}};
    const Mat3 rr3 = {{{-t3 * t3 - t2 * t2, t2 * t1, t3 * t1}, {t1 * t2, -t3 * t3 - t1 * t1, t3 * t2}, {t1 * t3, t2 * t3, -t2 * t2 - t1 * t1}}};
    double c1, c2;
    if (tt < 1e-10) { c1 = 1.0; c2 = 0.5; }
    else { c1 = ss / tt; c2 = (1.0 - cc) / (tt * tt); }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) rr[i][j] = rr1[i][j] + rr2[i][j] * c1 + rr3[i][j] * c2;
}

void Segment::MapReferencePosToCurrent(double coordsOut[12], const Mat3 &TTT, const Vec3 &XYZ, const Vec3 &AoA) const
{
    const Vec3 a = matvec3(TTT, x00), b = matvec3(TTT, x00 + 6);
    for (int i = 0; i < 3; i++) {
        coordsOut[i] = a[i] + XYZ[i];
        coordsOut[3 + i] = AoA[i];
        coordsOut[6 + i] = b[i] + XYZ[i];
        coordsOut[9 + i] = AoA[i];
    }
}

void Segment::MapReferenceDirToCurrent(Vec3 &dirc, const Mat3 &TTT) const { dirc = matvec3(TTT, dirc00.data()); }

// ---------------------------------------------------------------------------------------------------------------
// module SolidSolver
// ---------------------------------------------------------------------------------------------------------------

void BeamSolver::ReadBuild(double &nAsfac, double &nLchod)
{
    DatFile f(FEmeshName);
    f.next();   // title line
    {
        const auto &r = f.next();
        if (r.size() < 3) throw std::runtime_error("structural mesh file: second line must hold nND nEL nMT");
        nND = (int)DatFile::num(r[0]); nEL = (int)DatFile::num(r[1]); nMT = (int)DatFile::num(r[2]);
    }
    gEQ = nND * 6;
    vBC.assign(gEQ, 0.0);
    std::vector<std::array<double, 8>> xyz(nND), material(nMT);
    std::vector<std::array<int, 6>> boundary(nND);
    f.find_section("POINT", FEmeshName);
    if ((int)DatFile::num(f.next()[0]) != nND) throw std::runtime_error("ERROR: POINT number inconsistent");
    for (int i = 0; i < nND; i++) {
        const auto &r = f.next();
        if (r.size() < 9) throw std::runtime_error("POINT rows need: id X Y Z Lspan Rspan dirX dirY dirZ");
        for (int k = 0; k < 8; k++) xyz[i][k] = DatFile::num(r[1 + k]);
    }
    f.find_section("MATERIAL", FEmeshName);
    if ((int)DatFile::num(f.next()[0]) != nMT) throw std::runtime_error("ERROR: MATERIAL number inconsistent");
    for (int i = 0; i < nMT; i++) {
        const auto &r = f.next();
        if (r.size() < 9) throw std::runtime_error("MATERIAL rows need: id E G A RHO GAMMA JT IY IZ");
        for (int k = 0; k < 8; k++) material[i][k] = DatFile::num(r[1 + k]);
    }
    f.find_section("CONSTRAINT", FEmeshName);
    if ((int)DatFile::num(f.next()[0]) != nND) throw std::runtime_error("ERROR: CONSTRAINT number inconsistent");
    for (int i = 0; i < nND; i++) {
        const auto &r = f.next();
        if (r.size() < 7) throw std::runtime_error("CONSTRAINT rows need: id and six flags");
        for (int k = 0; k < 6; k++) boundary[i][k] = (int)DatFile::num(r[1 + k]);
    }
    m_elements.assign(nEL, Segment());
    f.find_section("ELEMENT", FEmeshName);
    if ((int)DatFile::num(f.next()[0]) != nEL) throw std::runtime_error("ERROR: ELEMENT number inconsistent");
    for (int n = 0; n < nEL; n++) {
        const auto &r = f.next();
        if (r.size() < 7) throw std::runtime_error("ELEMENT rows need: id I J K TYPE MAT Nspan");
        const int tmpid = (int)DatFile::num(r[0]), i = (int)DatFile::num(r[1]), j = (int)DatFile::num(r[2]);
        const int itype = (int)DatFile::num(r[4]), imat = (int)DatFile::num(r[5]), Nspan = (int)DatFile::num(r[6]);
        if (i < 1 || i > nND || j < 1 || j > nND || imat < 1 || imat > nMT) throw std::runtime_error("ELEMENT row refers to an unknown node or material");
        if (1 <= tmpid && tmpid <= nEL) m_elements[tmpid - 1].Build(i - 1, j - 1, itype, Nspan, xyz, material[imat - 1], boundary);
    }
    // Beam_cptAsfac, :1371-1384
    nAsfac = 0.0;
    elmax = -1e300; elmin = 1e300;
    double xlo = 1e300, xhi = -1e300, ylo = 1e300, yhi = -1e300;
    for (const Segment &e : m_elements) {
        nAsfac += e.areaElem00;
        elmax = std::max(elmax, e.areaElem00);
        elmin = std::min(elmin, e.areaElem00);
        xlo = std::min({xlo, e.x00[0], e.x00[6]}); xhi = std::max({xhi, e.x00[0], e.x00[6]});
        ylo = std::min({ylo, e.x00[1], e.x00[7]}); yhi = std::max({yhi, e.x00[1], e.x00[7]});
    }
    nLchod = xhi - xlo;
    const double lentemp = yhi - ylo;
    if (lentemp > nLchod) nLchod = lentemp;
    // Beam_adjustBC, :1386-1397
    for (Segment &e : m_elements) {
        if (e.bc[0] == 1) for (int k = 0; k < 6; k++) e.bc[k] = isMotionGiven[k];
        if (e.bc[6] == 1) for (int k = 0; k < 6; k++) e.bc[6 + k] = isMotionGiven[k];
    }
}

void BeamSolver::Initialise(double time)
{
    const double m_pi = P->pi;
    for (auto &r : TTT00) r.fill(0.0);
    TTT00[0][0] = 1.0; TTT00[1][1] = 1.0; TTT00[2][2] = 1.0;
    for (int k = 0; k < 3; k++) {
        XYZ[k] = XYZo[k] + XYZAmpl[k] * std::cos(2.0 * m_pi * Freq * time + XYZPhi[k]) + initXYZVel[k] * time;
        AoA[k] = AoAo[k] + AoAAmpl[k] * std::cos(2.0 * m_pi * Freq * time + AoAPhi[k]);
    }
    AoAtoTTT(AoA, TTT0);
    AoAtoTTT(AoA, TTTnxt);
    Segment_get_angle_triad(TTT0, TTTnxt, AoAd[0], AoAd[1], AoAd[2]);
    for (Segment &e : m_elements) {
        e.MapReferencePosToCurrent(e.x0, TTT0, XYZ, AoAd);
        e.MapReferenceDirToCurrent(e.dirc0, TTT0);
        e.Init();
    }
    // InitLoad :1448, InitPosDspVelAcc :1462
    for (auto *v : {&lodInte, &lodExte, &lodEffe, &lodFlow, &lodRepl, &lodGrav}) v->assign(gEQ, 0.0);
    pos.assign(6 * nND, 0.0);
    for (const Segment &e : m_elements)
        for (int i = 0; i < 6; i++) { pos[e.node0 * 6 + i] = e.x1[i]; pos[e.node1 * 6 + i] = e.x1[6 + i]; }
    dsp.assign(6 * nND, 0.0); vel.assign(6 * nND, 0.0); acc.assign(6 * nND, 0.0); mss.assign(3 * nND, 0.0);
    if (iBodyModel == 1) {
        for (int k = 0; k < 3; k++) {
            UVW[k] = -2.0 * m_pi * Freq * XYZAmpl[k] * std::sin(2.0 * m_pi * Freq * time + XYZPhi[k]) + initXYZVel[k];
            WWW1[k] = -2.0 * m_pi * Freq * AoAAmpl[k] * std::sin(2.0 * m_pi * Freq * time + AoAPhi[k]);
        }
        WWW2 = {WWW1[0] * std::cos(AoA[1]) + WWW1[2],
                WWW1[0] * std::sin(AoA[1]) * std::sin(AoA[2]) + WWW1[1] * std::cos(AoA[2]),
                WWW1[0] * std::sin(AoA[1]) * std::cos(AoA[2]) - WWW1[1] * std::sin(AoA[2])};
        WWW3 = matvec3(TTT0, WWW2.data());
        UpdateVelFromPosAngular(WWW3, UVW);
    }
    // InitTriadANDFormMass :1496
    for (Segment &e : m_elements) {
        e.InitTriad_D();
        e.FormMassMatrix();
        e.RotateMatrix();
        e.RKR(e.m_masMat);
    }
    // InitGrav :1513
    double gvec[12] = {0}, grav[12];
    for (int k = 0; k < 3; k++) { gvec[k] = P->g[k]; gvec[6 + k] = P->g[k]; }
    for (const Segment &e : m_elements) {
        e.MassMultiply(gvec, grav);
        e.LocToGlobal(grav, lodGrav);
        for (int j = 0; j < 3; j++) {
            mss[e.node0 * 3 + j] = mss[e.node0 * 3 + j] + e.m_masMat[j][j];
            mss[e.node1 * 3 + j] = mss[e.node1 * 3 + j] + e.m_masMat[j + 6][j + 6];
        }
    }
}

void BeamSolver::UpdateVelFromPosAngular(const Vec3 &WWW, const Vec3 &UVW_)
{
    for (int n = 0; n < nND; n++) {
        const double rel[3] = {pos[6 * n] - XYZ[0], pos[6 * n + 1] - XYZ[1], pos[6 * n + 2] - XYZ[2]};
        vel[6 * n + 0] = (WWW[1] * rel[2] - WWW[2] * rel[1]) + UVW_[0];
        vel[6 * n + 1] = (WWW[2] * rel[0] - WWW[0] * rel[2]) + UVW_[1];
        vel[6 * n + 2] = (WWW[0] * rel[1] - WWW[1] * rel[0]) + UVW_[2];
        for (int k = 0; k < 3; k++) vel[6 * n + 3 + k] = WWW[k];
    }
}

void BeamSolver::calculate_angle_material(double Lref, double Uref, double denIn, double &uMax, const Vec3 &uuuIn, double &nLthck)
{
    const double m_pi = P->pi;
    St = Lref * Freq / Uref;
    for (int k = 0; k < 3; k++) {
        AoAo[k] = AoAo[k] / 180.0 * m_pi;
        AoAAmpl[k] = AoAAmpl[k] / 180.0 * m_pi;
        AoAPhi[k] = AoAPhi[k] / 180.0 * m_pi;
        XYZPhi[k] = XYZPhi[k] / 180.0 * m_pi;
    }
    double xmax = 0, ymax = 0, zmax = 0;
    for (const Segment &e : m_elements) {
        xmax = std::max({xmax, std::fabs(e.x00[0]), std::fabs(e.x00[6])});
        ymax = std::max({ymax, std::fabs(e.x00[1]), std::fabs(e.x00[7])});
        zmax = std::max({zmax, std::fabs(e.x00[2]), std::fabs(e.x00[8])});
    }
    const double rRot[3] = {std::max(ymax, zmax), std::max(xmax, zmax), std::max(xmax, ymax)};
    double amax = 0, rmax = 0, umaxIn = 0;
    for (int k = 0; k < 3; k++) {
        amax = std::max(amax, std::fabs(XYZAmpl[k]));
        rmax = std::max(rmax, std::fabs(AoAAmpl[k]) * rRot[k]);
        umaxIn = std::max(umaxIn, std::fabs(uuuIn[k]));
    }
    uMax = std::max({uMax, umaxIn, 2.0 * m_pi * amax * Freq, 2.0 * m_pi * rmax * Freq});
    // Fortran's x**n with an integer n is a primary of the product chain it stands in: Uref**2 = Uref*Uref, so KS*denIn*Uref**2 is
    // (KS*denIn)*(Uref*Uref), not ((KS*denIn)*Uref)*Uref.  Beyond x**2 gfortran (without fast-math) calls libgcc's __powidf2, which
    // squares and multiplies from the low bit: x**5 = x*((x*x)*(x*x)), x**3 = x*(x*x)
    const double U2 = Uref * Uref;
    auto pow5 = [](double x) { const double x2 = x * x; return x * (x2 * x2); };
    nLthck = 0.0;   // left undefined by the reference when isKB is neither 0 nor 1
    if (P->isKB == 0) {
        for (Segment &e : m_elements) {
            const double len = e.spanlen;
            e.m_property[0] = EmR * denIn * U2;
            e.m_property[1] = e.m_property[0] / (2.0 * (1.0 + psR));
            nLthck = tcR * Lref;
            e.m_property[2] = len * nLthck;
            e.m_property[3] = denR * len * Lref * denIn / e.m_property[2];
            const double ratio = nLthck / len;
            e.m_property[5] = len * (nLthck * nLthck * nLthck) / 3.0 * (1.0 - 0.63 * ratio + 0.052 * pow5(ratio));
            e.m_property[6] = len * (nLthck * nLthck * nLthck) / 12.0;
            e.m_property[7] = nLthck * (len * len * len) / 12.0;
        }
        const double len = m_elements[0].spanlen;
        KB = m_elements[0].m_property[0] * m_elements[0].m_property[6] / (denIn * U2 * (Lref * Lref * Lref) * len);
        KS = m_elements[0].m_property[0] * m_elements[0].m_property[2] / (denIn * U2 * Lref * len);
    }
    if (P->isKB == 1) {
        for (Segment &e : m_elements) {
            const double len = e.spanlen;
            nLthck = std::sqrt(KB / KS * 12.0) * Lref;
            e.m_property[2] = len * nLthck;
            e.m_property[0] = KS * denIn * U2 * Lref * len / e.m_property[2];
            e.m_property[1] = e.m_property[0] / (2.0 * (1.0 + psR));
            e.m_property[3] = denR * len * Lref * denIn / e.m_property[2];
            const double ratio = nLthck / len;
            e.m_property[5] = len * (nLthck * nLthck * nLthck) / 3.0 * (1.0 - 0.63 * ratio + 0.052 * pow5(ratio));
            e.m_property[6] = len * (nLthck * nLthck * nLthck) / 12.0;
            e.m_property[7] = nLthck * (len * len * len) / 12.0;
        }
        const double len = m_elements[0].spanlen;
        nLthck = m_elements[0].m_property[2] / len;
        EmR = m_elements[0].m_property[0] / (denIn * U2);
        tcR = nLthck / Lref;
    }
}

void BeamSolver::structure(int iFish, double time, int isubstep, double deltat, double subdeltat)
{
    const double m_pi = P->pi;
    const double t = time - deltat + (double)isubstep * subdeltat;
    if (iBodyModel != 1 && iBodyModel != 2) throw std::runtime_error("no define body model");
    for (int k = 0; k < 3; k++) {
        XYZ[k] = XYZo[k] + XYZAmpl[k] * std::cos(2.0 * m_pi * Freq * t + XYZPhi[k]) + initXYZVel[k] * t;
        AoA[k] = AoAo[k] + AoAAmpl[k] * std::cos(2.0 * m_pi * Freq * t + AoAPhi[k]);
    }
    AoAtoTTT(AoA, TTTnxt);
    Segment_get_angle_triad(TTT0, TTTnxt, AoAd[0], AoAd[1], AoAd[2]);
    if (iBodyModel == 1) {   // rigid body, prescribed motion (:1826-1857)
        for (Segment &e : m_elements) {
            e.MapReferencePosToCurrent(e.xnxt, TTTnxt, XYZ, AoAd);
            e.MapReferenceDirToCurrent(e.dircnxt, TTTnxt);
            for (int i = 0; i < 12; i++) e.x1[i] = e.xnxt[i];
            e.dirc1 = e.dircnxt;
            for (int i = 0; i < 6; i++) { pos[e.node0 * 6 + i] = e.x1[i]; pos[e.node1 * 6 + i] = e.x1[6 + i]; }
            e.cptdxyz1();
            e.RigidUpdateTriad_D();
        }
        for (int k = 0; k < 3; k++) {
            UVW[k] = -2.0 * m_pi * Freq * XYZAmpl[k] * std::sin(2.0 * m_pi * Freq * t + XYZPhi[k]) + initXYZVel[k];
            WWW1[k] = -2.0 * m_pi * Freq * AoAAmpl[k] * std::sin(2.0 * m_pi * Freq * t + AoAPhi[k]);
        }
        WWW2 = {WWW1[0] * std::cos(AoA[1]) + WWW1[2],
                WWW1[0] * std::sin(AoA[1]) * std::sin(AoA[2]) + WWW1[1] * std::cos(AoA[2]),
                WWW1[0] * std::sin(AoA[1]) * std::cos(AoA[2]) - WWW1[1] * std::sin(AoA[2])};
        WWW3 = matvec3(TTTnxt, WWW2.data());
        UpdateVelFromPosAngular(WWW3, UVW);
    } else {                 // elastic model (:1859-1872)
        for (Segment &e : m_elements) e.MapReferencePosToCurrent(e.xnxt, TTTnxt, XYZ, AoAd);
        UpdateNewmarkCoeffs(subdeltat);
        Solver(iFish);
    }
}

void BeamSolver::UpdateNewmarkCoeffs(double dt)
{
    const double beta = P->NewmarkBeta, gamma = P->NewmarkGamma;
    coeffs[2] = 1.0 / (beta * dt);
    coeffs[1] = gamma * coeffs[2];
    coeffs[0] = coeffs[2] / dt;
    coeffs[3] = 0.5 / beta - 1.0;
    coeffs[4] = gamma / beta - 1.0;
    coeffs[5] = dt * (0.5 * gamma / beta - 1.0);
    coeffs[6] = dt * (1.0 - gamma);
    coeffs[7] = dt * gamma;
}

void BeamSolver::Solver(int iFish)
{
    // InitDspVelAccATTimeT, :1916-1930
    std::fill(vBC.begin(), vBC.end(), 0.0);
    for (const Segment &e : m_elements)
        for (int i = 0; i < 12; i++) vBC[e.m_localToGlobal[i]] = e.xnxt[i] - e.x1[i];
    for (int i = 0; i < gEQ; i++) lodExte[i] = lodFlow[i] + lodGrav[i] + lodRepl[i];
    const std::vector<double> dspO = dsp, velO = vel, accO = acc;
    std::vector<double> dspn(gEQ, 0.0);
    double dnorm = 1.0;
    int iter;
    for (iter = 1; iter <= P->ntolFEM; iter++) {
        UpdateMatrixANDLoad(dspO);
        CG_Solve(dspn, lodEffe, iter);
        UpdateDspANDTride(iter, dspn, dnorm);
        if (dnorm <= P->dtolFEM) break;
    }
    // UpdateVelAcc, :2350-2359
    for (int i = 0; i < 6 * nND; i++) {
        acc[i] = coeffs[0] * (dsp[i] - dspO[i]) - coeffs[2] * velO[i] - coeffs[3] * accO[i];
        vel[i] = velO[i] + coeffs[6] * accO[i] + coeffs[7] * acc[i];
    }
    FishInfo[0] = (double)iFish;
    FishInfo[1] = (double)iter;   // a DO loop that runs out leaves iter = ntolFEM + 1, as here
    FishInfo[2] = dnorm;
}

void BeamSolver::UpdateMatrixANDLoad(const std::vector<double> &dspO)
{
    std::fill(lodInte.begin(), lodInte.end(), 0.0);
    for (Segment &e : m_elements) {
        e.FormStiffMatrix();
        e.BodyStress_D(lodInte);
        e.UpdateMatrix(coeffs, P->GeoGamma, P->dampM, P->dampK);
    }
    for (int i = 0; i < gEQ; i++) lodEffe[i] = lodExte[i] - lodInte[i];
    for (const Segment &e : m_elements) e.UpdateLoad(coeffs, P->dampM, P->dampK, dspO, dsp, vel, acc, lodEffe);
}

void BeamSolver::MatrixMultipy(const std::vector<double> &x, std::vector<double> &b) const
{
    std::fill(b.begin(), b.end(), 0.0);
    for (const Segment &e : m_elements) e.Multiply(x, b);
}

void BeamSolver::CG_Solve(std::vector<double> &x, const std::vector<double> &b, int iterNR)
{
    const int n = gEQ, max_iter = 10000;
    const double err = 1e-6;
    std::vector<double> r(n), p(n), Ap(n), z(n), xFixed(n, 0.0);
    std::vector<char> fixed(n, 0);
    auto dot = [n](const std::vector<double> &a, const std::vector<double> &c) { double s = 0.0; for (int i = 0; i < n; i++) s += a[i] * c[i]; return s; };
    auto zero_fixed = [&](std::vector<double> &v) { for (int i = 0; i < n; i++) if (fixed[i]) v[i] = 0.0; };
    auto breakdown = [&](double value, double residual_norm, const char *name, const char *where) {   // Beam_CheckCGBreakdown, :2281
        if (std::fabs(value) <= 1e-30) {
            char msg[256];
            std::snprintf(msg, sizeof(msg), "ERROR: CG solver breakdown. Location: %s Quantity: %s Value: %g residual_norm: %g", where, name, value, residual_norm);
            throw std::runtime_error(msg);
        }
    };
    for (const Segment &e : m_elements) e.BoundaryCond(iterNR, xFixed, fixed, vBC);   // Beam_BoundaryCond, :2022
    x = xFixed;
    MatrixMultipy(x, Ap);
    for (int i = 0; i < n; i++) r[i] = b[i] - Ap[i];
    zero_fixed(r);
    // node-wise 6x6 block-Jacobi preconditioner (precondType = 2): :2127-2136, :2174-2205
    std::vector<double> blockM(36 * nND, 0.0);
    for (const Segment &e : m_elements)
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < 6; j++) {
                blockM[36 * e.node0 + 6 * i + j] += e.m_coefMat[i][j];
                blockM[36 * e.node1 + 6 * i + j] += e.m_coefMat[6 + i][6 + j];
            }
    for (int node = 0; node < nND; node++) {
        double B[6][6], inv[6][6];
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) B[i][j] = blockM[36 * node + 6 * i + j];
        for (int i = 0; i < 6; i++)
            if (fixed[node * 6 + i]) {
                for (int j = 0; j < 6; j++) { B[i][j] = 0.0; B[j][i] = 0.0; }
                B[i][i] = 1.0;
            }
        if (!Invert6x6(B, inv)) {
            std::printf(" WARNING: block Jacobi inverse failed at node %d\n Use identity block for this node.\n", node + 1);
            for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) inv[i][j] = i == j ? 1.0 : 0.0;
        }
        for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) blockM[36 * node + 6 * i + j] = inv[i][j];
    }
    auto precond = [&](const std::vector<double> &rr, std::vector<double> &zz) {   // :2207-2232
        for (int node = 0; node < nND; node++)
            for (int i = 0; i < 6; i++) {
                double s = 0.0;
                for (int j = 0; j < 6; j++) s += blockM[36 * node + 6 * i + j] * rr[node * 6 + j];
                zz[node * 6 + i] = fixed[node * 6 + i] ? 0.0 : s;
            }
    };
    precond(r, z);
    double residual_norm = std::sqrt(dot(r, r));
    if (residual_norm <= err) return;
    p = z;
    double rsold = dot(r, z);
    breakdown(rsold, residual_norm, "rsold", "initial dot_product(r,z)");
    for (int iter = 1; iter <= max_iter; iter++) {
        cg_iterations++;
        MatrixMultipy(p, Ap);
        zero_fixed(Ap);
        const double pAp = dot(p, Ap);
        breakdown(pAp, residual_norm, "pAp", "dot_product(p,Ap) before alpha");
        const double alpha = rsold / pAp;
        for (int i = 0; i < n; i++) x[i] = x[i] + alpha * p[i];
        for (int i = 0; i < n; i++) if (fixed[i]) x[i] = xFixed[i];
        for (int i = 0; i < n; i++) r[i] = r[i] - alpha * Ap[i];
        zero_fixed(r);
        residual_norm = std::sqrt(dot(r, r));
        if (residual_norm <= err) break;
        precond(r, z);
        const double rsnew = dot(r, z);
        breakdown(rsold, residual_norm, "rsold", "rsold before beta=rsnew/rsold");
        const double beta = rsnew / rsold;
        for (int i = 0; i < n; i++) p[i] = z[i] + beta * p[i];
        zero_fixed(p);
        rsold = rsnew;
    }
    if (residual_norm > err)
        std::printf(" WARNING: CG did not fully converge.\n residual_norm = %g err = %g\n max_iter = %d\n", residual_norm, err, max_iter);
}

void BeamSolver::UpdateDspANDTride(int iter, const std::vector<double> &dspn, double &dnorm)
{
    const double beta0 = 1.0;
    const int maxramp = 1;
    double beta;
    if (iter <= maxramp) beta = std::pow(2.0, iter) / std::pow(2.0, maxramp) * beta0;
    else beta = 1.0 * beta0;
    std::vector<double> dspnn(6 * nND);
    for (int i = 0; i < 6 * nND; i++) dspnn[i] = beta * dspn[i];
    for (int i = 0; i < 6 * nND; i++) dsp[i] = dsp[i] + dspnn[i];
    for (Segment &e : m_elements) {
        for (int i = 0; i < 6; i++) {
            e.x1[i] = e.x0[i] + dsp[e.node0 * 6 + i];
            e.x1[6 + i] = e.x0[6 + i] + dsp[e.node1 * 6 + i];
        }
        for (int i = 0; i < 6; i++) { pos[e.node0 * 6 + i] = e.x1[i]; pos[e.node1 * 6 + i] = e.x1[6 + i]; }
        e.cptdxyz1();
        e.UpdateTriad_D(dspnn);
        e.MakeTriad_ee();
    }
    double m = 0.0;
    for (int i = 0; i < gEQ; i++) m = std::max(m, (beta * dspn[i]) * (beta * dspn[i]));
    dnorm = std::fabs(m);
}

void BeamSolver::UpdateStrainEnergy()
{
    for (Segment &e : m_elements) {
        e.FormStiffMatrix();
        e.StrainEnergy_D();
    }
}

bool Invert6x6(const double A[6][6], double Ainv[6][6])
{
    double aug[6][12];
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) { aug[i][j] = A[i][j]; aug[i][6 + j] = i == j ? 1.0 : 0.0; }
    for (int i = 0; i < 6; i++) {
        int p = i;
        double pivot = std::fabs(aug[i][i]);
        for (int k = i + 1; k < 6; k++)
            if (std::fabs(aug[k][i]) > pivot) { pivot = std::fabs(aug[k][i]); p = k; }
        if (pivot <= 1e-30) {
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) Ainv[a][b] = 0.0;
            return false;
        }
        if (p != i) for (int j = 0; j < 12; j++) std::swap(aug[i][j], aug[p][j]);
        pivot = aug[i][i];
        for (int j = 0; j < 12; j++) aug[i][j] = aug[i][j] / pivot;
        for (int k = 0; k < 6; k++)
            if (k != i) {
                const double factor = aug[k][i];
                for (int j = 0; j < 12; j++) aug[k][j] = aug[k][j] - factor * aug[i][j];
            }
    }
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) Ainv[i][j] = aug[i][6 + j];
    return true;
}

void AoAtoTTT(const Vec3 &AoA, Mat3 &TTT)
{
    // the private "pi" below is the reference's (a typo of pi, :2406); it only enters the snap-to-axis tests
    const double pi = 3.141562653589793, eps = 1e-5;
    auto cs = [&](double a, double &vcos, double &vsin) {
        vcos = std::cos(a); vsin = std::sin(a);
        if (std::fabs(a) < eps) { vcos = 1.0; vsin = 0.0; }
        if (std::fabs(a - 0.5 * pi) < eps) { vcos = 0.0; vsin = 1.0; }
        if (std::fabs(a + 0.5 * pi) < eps) { vcos = 0.0; vsin = -1.0; }
    };
    double c, s;
    cs(AoA[0], c, s);
    const Mat3 rrx = {{{1.0, 0.0, 0.0}, {0.0, c, -s}, {0.0, s, c}}};
    cs(AoA[1], c, s);
    const Mat3 rry = {{{c, 0.0, s}, {0.0, 1.0, 0.0}, {-s, 0.0, c}}};
    cs(AoA[2], c, s);
    const Mat3 rrz = {{{c, -s, 0.0}, {s, c, 0.0}, {0.0, 0.0, 1.0}}};
    Mat3 I = {{{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}}};
    TTT = matmul3(rrz, I);
    TTT = matmul3(rry, TTT);
    TTT = matmul3(rrx, TTT);
}

// ---------------------------------------------------------------------------------------------------------------
// writers
// ---------------------------------------------------------------------------------------------------------------

void BeamSolver::write_solid(double Lref, double Uref, double Aref, double Fref, int iFish, FILE *fh) const
{
    const int numVar = 15;
    std::fprintf(fh, " ZONE T = \"fish%s\"\n", fmtI(iFish, 4, 4).c_str());
    std::fprintf(fh, " STRANDID=0, SOLUTIONTIME=0\n");
    std::fprintf(fh, " Nodes=%s, Elements=%s, ZONETYPE=", fmtI(nND, 8).c_str(), fmtI(nEL, 8).c_str());
    const int ElmType = m_elements[0].itype;
    if (ElmType == 2) std::fprintf(fh, "FELINESEG\n");
    std::fprintf(fh, " DATAPACKING=POINT\n");
    std::fprintf(fh, " DT=(");
    for (int i = 1; i <= numVar - 1; i++) std::fprintf(fh, "SINGLE ");
    std::fprintf(fh, "SINGLE )\n");
    for (int i = 0; i < nND; i++) {
        std::string row;
        for (int k = 0; k < 3; k++) row += fmtE(pos[6 * i + k] / Lref, 28, 18);
        for (int k = 0; k < 3; k++) row += fmtE(vel[6 * i + k] / Uref, 28, 18);
        for (int k = 0; k < 3; k++) row += fmtE(acc[6 * i + k] / Aref, 28, 18);
        for (int k = 0; k < 3; k++) row += fmtE(lodFlow[6 * i + k] / Fref, 28, 18);
        for (int k = 0; k < 3; k++) row += fmtE(lodRepl[6 * i + k] / Fref, 28, 18);
        std::fprintf(fh, "%s\n", row.c_str());
    }
    if (ElmType == 2)
        for (const Segment &e : m_elements) std::fprintf(fh, " %s %s\n", fmtI(e.node0 + 1, 11).c_str(), fmtI(e.node1 + 1, 11).c_str());   // list-directed integers
}

void BeamSolver::write_solid_params(FILE *fh) const
{
    std::fprintf(fh, "EmR, tcR    =%s%s\n", fmtF(EmR, 20, 10).c_str(), fmtF(tcR, 20, 10).c_str());
    std::fprintf(fh, "nND = %s  nEL = %s  \n", fmtI(nND, 8).c_str(), fmtI(nEL, 8).c_str());
    std::fprintf(fh, "nMT = %s  gEQ = %s  \n", fmtI(nMT, 8).c_str(), fmtI(gEQ, 8).c_str());
}

void BeamSolver::write_solid_materials(FILE *fh) const
{
    std::fprintf(fh, "------------------------------- m_element( %s ) ------------------------------\n", fmtI(1, 4, 4).c_str());
    const char *names[8] = {"E    =", "G    =", "A    =", "rho  =", "gamma=", "Jt   =", "Iy   =", "Iz   ="};
    for (int k = 0; k < 8; k++) std::fprintf(fh, "%s%s\n", names[k], fmtE(m_elements[0].m_property[k], 20, 10).c_str());
}

namespace {
void append_row(const std::string &file, const std::vector<double> &v)
{
    FILE *fh = std::fopen(file.c_str(), "a");
    if (!fh) return;
    std::string row;
    for (double x : v) row += fmtE(x, 20, 10);
    std::fprintf(fh, "%s\n", row.c_str());
    std::fclose(fh);
}
}  // namespace

void BeamSolver::write_solid_info(const std::string &groupNum, const Vec3 &XYZo_, double Lref, double Uref, double Aref, double Fref, double Pref, double Eref)
{
    const std::string base = "./DatInfo/Group" + groupNum;
    auto node_row = [&](int nd) {
        std::vector<double> v;
        for (int k = 0; k < 3; k++) v.push_back(XYZo_[k] / Lref);
        for (int k = 0; k < 3; k++) v.push_back((pos[6 * nd + k] - XYZo_[k]) / Lref);
        for (int k = 3; k < 6; k++) v.push_back(pos[6 * nd + k]);
        for (int k = 0; k < 3; k++) v.push_back(vel[6 * nd + k] / Uref);
        for (int k = 0; k < 3; k++) v.push_back(acc[6 * nd + k] / Aref);
        return v;
    };
    append_row(base + "_firstNode.dat", node_row(0));
    append_row(base + "_lastNode.dat", node_row(nND - 1));
    append_row(base + "_centerNode.dat", node_row((nND + 1) / 2 - 1));
    double msum[3] = {0, 0, 0}, xcm[3] = {0, 0, 0}, vcm[3] = {0, 0, 0}, acm[3] = {0, 0, 0};
    for (int n = 0; n < nND; n++)
        for (int k = 0; k < 3; k++) {
            msum[k] += mss[3 * n + k];
            xcm[k] += pos[6 * n + k] * mss[3 * n + k];
            vcm[k] += vel[6 * n + k] * mss[3 * n + k];
            acm[k] += acc[6 * n + k] * mss[3 * n + k];
        }
    {
        std::vector<double> v;
        for (int k = 0; k < 3; k++) v.push_back(XYZo_[k] / Lref);
        for (int k = 0; k < 3; k++) v.push_back((xcm[k] / msum[k] - XYZo_[k]) / Lref);
        for (int k = 0; k < 3; k++) v.push_back(vcm[k] / msum[k] / Uref);
        for (int k = 0; k < 3; k++) v.push_back(acm[k] / msum[k] / Aref);
        append_row(base + "_nodeAverage.dat", v);
    }
    double F[3] = {0, 0, 0}, Pa[3] = {0, 0, 0};
    for (int n = 0; n < nND; n++)
        for (int k = 0; k < 3; k++) {
            F[k] += lodFlow[6 * n + k];
            Pa[k] += lodFlow[6 * n + k] * vel[6 * n + k];
        }
    append_row(base + "_forces.dat", {XYZo_[0] / Lref, XYZo_[1] / Lref, XYZo_[2] / Lref, F[0] / Fref, F[1] / Fref, F[2] / Fref});
    const double Pax = Pa[0] / Pref, Pay = Pa[1] / Pref, Paz = Pa[2] / Pref, Ptot = Pax + Pay + Paz;
    append_row(base + "_power.dat", {XYZo_[0] / Lref, XYZo_[1] / Lref, XYZo_[2] / Lref, Ptot, Pax, Pay, Paz});
    UpdateStrainEnergy();
    double Es = 0.0, Eb = 0.0, Ev = 0.0;
    for (const Segment &e : m_elements) { Es = Es + e.strainEnergy[0]; Eb = Eb + e.strainEnergy[1]; }
    Es = Es / Eref;
    Eb = Eb / Eref;
    const double Ep = Es + Eb;
    for (const Segment &e : m_elements) {
        double ve[12];
        for (int i = 0; i < 6; i++) { ve[i] = vel[e.node0 * 6 + i]; ve[6 + i] = vel[e.node1 * 6 + i]; }
        double s = 0.0;
        for (int i = 0; i < 12; i++) {
            double t = 0.0;
            for (int j = 0; j < 12; j++) t += e.m_masMat[i][j] * ve[j];
            s += ve[i] * t;
        }
        Ev = Ev + 0.5 * s;
    }
    Ev = Ev / Eref;
    const double Etot = Ev + Ep;
    append_row(base + "_energy.dat", {XYZo_[0] / Lref, XYZo_[1] / Lref, XYZo_[2] / Lref, Etot, Ev, Ep, Es, Eb});
}

void BeamSolver::write_solid_probes(const std::string &groupNum, const Vec3 &XYZo_, const std::vector<int> &solidProbingNode, double Lref, double Uref, double Aref) const
{
    for (size_t i = 0; i < solidProbingNode.size(); i++) {
        const int nd = solidProbingNode[i] - 1;
        if (nd < 0 || nd >= nND) continue;
        std::vector<double> v;
        for (int k = 0; k < 3; k++) v.push_back(XYZo_[k] / Lref);
        for (int k = 0; k < 3; k++) v.push_back((pos[6 * nd + k] - XYZo_[k]) / Lref);
        for (int k = 3; k < 6; k++) v.push_back(pos[6 * nd + k]);
        for (int k = 0; k < 3; k++) v.push_back(vel[6 * nd + k] / Uref);
        for (int k = 0; k < 3; k++) v.push_back(acc[6 * nd + k] / Aref);
        append_row("./DatInfo/Group" + groupNum + "_solidProbes_" + fmtI((long long)i + 1, 4, 4) + ".dat", v);
    }
}

}  // namespace harness
