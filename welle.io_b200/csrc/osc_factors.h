// osc_factors.h — the oscillator of OFDMProcessor::getSamples (backend/ofdm-processor.cpp:92-94,211-214) without its table.
//
// The reference multiplies sample k by oscillatorTable[m], m = the running phase index in [0, 2 048 000), where
//   oscillatorTable[i] = ((float)cos(2.0 * M_PI * i / 2048000), (float)sin(2.0 * M_PI * i / 2048000)).
// Reading that 16 MB table on the GPU is a scattered 8-byte gather per lane.  Instead:
//   osc(m) = float( H[m >> 10] * exp(j theta (m & 1023)) ),  theta = 2 pi / 2 048 000,
// in double precision with explicitly rounded operations: H = 2000 correctly rounded factors, the small rotation from
// its Taylor polynomial (theta * 1023 < 3.2e-3, the neglected terms are below 1e-18).  The result is the table entry for
// ALL 2 048 000 indices once the three quarter-turn factors take the table's own value (there the reference's double
// angle is an ulp off and its cos/sin returns 6e-17 instead of 0) — checked exhaustively on the CPU by
// tests/host_emul/emul_osc.cpp and on the device at every dabb_create (osc_verify_kernel).
//
// Plain C++ (no CUDA types) so that the table builders and the formula are shared by tables.cpp, the device code and the
// CPU emulation.  `Ops` supplies the individually rounded double operations of the platform.
#pragma once
#include <cmath>
#if defined(__CUDACC__)
#define OSC_HD __host__ __device__ __forceinline__
#else
#define OSC_HD inline
#endif

namespace dabb {

constexpr int OSC_RATE = 2048000, OSC_HI = 2000, OSC_LO_BITS = 10;

template <class F2> inline void build_osc_table_t(F2* osc)
{
    for (int i = 0; i < OSC_RATE; i++) {
        osc[i].x = (float)cos(2.0 * M_PI * i / OSC_RATE);
        osc[i].y = (float)sin(2.0 * M_PI * i / OSC_RATE);
    }
}

// hi[a] = correctly rounded double of exp(j 2 pi 1024 a / 2 048 000) (80-bit evaluation, then rounded); where float(hi[a])
// is not the table's entry for m = 1024 a (the on-the-fly value for r = 0 is float(H[a]) itself), hi[a] becomes that entry
template <class F2, class D2> inline void build_osc_factors_t(const F2* osc_table, D2* hi, double* theta, int* patched)
{
    const long double two_pi = 2.0L * 3.141592653589793238462643383279502884L;
    int n = 0;
    for (int a = 0; a < OSC_HI; a++) {
        const long double x = two_pi * (long double)(a << OSC_LO_BITS) / OSC_RATE;
        hi[a].x = (double)cosl(x); hi[a].y = (double)sinl(x);
        const F2 t = osc_table[a << OSC_LO_BITS];
        if ((float)hi[a].x != t.x || (float)hi[a].y != t.y) { hi[a].x = (double)t.x; hi[a].y = (double)t.y; n++; }
    }
    *theta = (double)(two_pi / OSC_RATE);
    if (patched) *patched = n;
}

// the formula; r = m & 1023 given as an exactly converted double
template <class Ops> OSC_HD void osc_formula(double hx, double hy, double r, double theta, float& ox, float& oy)
{
    const double y = Ops::mul(r, theta), y2 = Ops::mul(y, y);
    const double c = Ops::fma(y2, Ops::fma(y2, 1.0 / 24, -0.5), 1.0);
    const double sn = Ops::mul(y, Ops::fma(y2, Ops::fma(y2, 1.0 / 120, -1.0 / 6), 1.0));
    const double wr = Ops::fma(hx, c, -Ops::mul(hy, sn)), wi = Ops::fma(hx, sn, Ops::mul(hy, c));
    ox = Ops::to_float(wr); oy = Ops::to_float(wi);
}

} // namespace dabb
