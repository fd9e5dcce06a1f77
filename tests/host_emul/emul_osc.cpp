// CPU emulation of the on-the-fly oscillator (welle.io_b200/csrc/osc_factors.h: the same table builders and the same formula the
// device code instantiates, with IEEE double operations from <cmath>): counts the indices whose value differs from the
// reference's oscillator table.  Must be 0 for the device path to be bit-exact.
#include "../../welle.io_b200/csrc/osc_factors.h"
#include <cstring>
#include <vector>

struct F2 { float x, y; };
struct D2 { double x, y; };
struct HostOps {
    static double mul(double a, double b) { return a * b; }             // compiled with -ffp-contract=off
    static double fma(double a, double b, double c) { return std::fma(a, b, c); }
    static float to_float(double a) { return (float)a; }
};

extern "C" int emul_osc_mismatches(int* patched_out)
{
    std::vector<F2> tab(dabb::OSC_RATE);
    dabb::build_osc_table_t(tab.data());
    std::vector<D2> hi(dabb::OSC_HI);
    double theta; int patched = 0;
    dabb::build_osc_factors_t(tab.data(), hi.data(), &theta, &patched);
    int bad = 0;
    for (int m = 0; m < dabb::OSC_RATE; m++) {
        F2 o;
        dabb::osc_formula<HostOps>(hi[m >> dabb::OSC_LO_BITS].x, hi[m >> dabb::OSC_LO_BITS].y, (double)(m & ((1 << dabb::OSC_LO_BITS) - 1)), theta, o.x, o.y);
        if (memcmp(&o, &tab[m], sizeof o)) bad++;
    }
    if (patched_out) *patched_out = patched;
    return bad;
}
