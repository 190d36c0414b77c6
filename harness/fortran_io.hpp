// fortran_io.hpp -- the few Fortran formatted-output edit descriptors the reference's writers use, so the C++
// stand-in driver emits the same text files (Ew.d, zero-padded Iw.w, the 10-digit time stamp of the file names).
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace harness {

// Fortran Ew.d: [-]0.ddd...dE+xx (d digits after the point, mantissa in [0.1,1)); right-justified in w columns.
inline std::string fmtE(double v, int w, int d)
{
    char buf[96];
    std::string body;
    if (std::isnan(v)) body = "NaN";
    else if (std::isinf(v)) body = v < 0 ? "-Infinity" : "Infinity";
    else {
        std::snprintf(buf, sizeof(buf), "%.*e", d - 1, std::fabs(v));   // D.DDDDe+XX with d significant digits
        std::string s(buf);
        const size_t e = s.find('e');
        std::string digits = s.substr(0, 1) + (d > 1 ? s.substr(2, e - 2) : std::string());
        int ex = std::atoi(s.c_str() + e + 1);
        if (v != 0.0) ex += 1;
        char ebuf[16];
        if (std::abs(ex) < 100) std::snprintf(ebuf, sizeof(ebuf), "E%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
        else std::snprintf(ebuf, sizeof(ebuf), "%c%03d", ex < 0 ? '-' : '+', std::abs(ex));
        body = std::string(std::signbit(v) && v != 0.0 ? "-" : "") + "0." + digits + ebuf;
        if ((int)body.size() > w && body.compare(std::signbit(v) && v != 0.0 ? 1 : 0, 2, "0.") == 0)
            body.erase(std::signbit(v) && v != 0.0 ? 1 : 0, 1);           // the optional leading zero is dropped first
    }
    if ((int)body.size() > w) return std::string(w, '*');
    return std::string(w - body.size(), ' ') + body;
}

// Fortran Fw.d
inline std::string fmtF(double v, int w, int d)
{
    char buf[128];
    std::snprintf(buf, sizeof(buf), "%.*f", d, v);
    std::string body(buf);
    if ((int)body.size() > w && body.compare(0, 2, "0.") == 0) body.erase(0, 1);
    if ((int)body.size() > w && body.compare(0, 3, "-0.") == 0) body.erase(1, 1);
    if ((int)body.size() > w) return std::string(w, '*');
    return std::string(w - body.size(), ' ') + body;
}

// Fortran Iw (right-justified) and Iw.m (zero-padded to m digits)
inline std::string fmtI(long long v, int w, int m = 0)
{
    char buf[64];
    if (m > 0) std::snprintf(buf, sizeof(buf), "%0*lld", m, v); else std::snprintf(buf, sizeof(buf), "%lld", v);
    std::string body(buf);
    if (w <= 0) return body;   // I0
    if ((int)body.size() > w) return std::string(w, '*');
    return std::string(w - body.size(), ' ') + body;
}

// write(fileName,'(I10)') nint(time/Tref*1d5), blanks replaced by '0' (FluidDomain.f90:1640-1646, Solidbody.f90:410-414)
inline std::string time_stamp10(double time, double Tref)
{
    return fmtI((long long)std::llround(time / Tref * 1e5), 10, 10);
}

}  // namespace harness
