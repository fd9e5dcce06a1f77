// CPU emulation of the CTA-level FFT passes of welle.io_b200/csrc/ofdm_core.cuh (same header, compiled as plain C++):
// loops over the 128 "threads" between the barriers.  Exposes emul_fft2048() so tests can compare with the oracle.
#include "../../welle.io_b200/csrc/ofdm_core.cuh"
#include <cmath>
#include <cstring>
#include <vector>
using namespace dabb;

static void fill(float2* lay, bool inverse)
{
    std::vector<float2> tw(TU);
    const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
    for (int i = 0; i < TU; i++) { double ph = -2 * pi * i / TU; if (inverse) ph *= -1; tw[i].x = (float)cos(ph); tw[i].y = (float)sin(ph); }
    for (int j = 0; j < 3; j++) lay[TwLayout::A3 + j] = tw[256 * (j + 1)];
    for (int j = 0; j < 3; j++) for (int k = 0; k < 8; k++) lay[TwLayout::B2 + 8 * j + k] = tw[64 * k * (j + 1)];
    for (int j = 0; j < 3; j++) for (int k = 0; k < 32; k++) lay[TwLayout::B3 + 32 * j + k] = tw[16 * k * (j + 1)];
    for (int j = 0; j < 3; j++) for (int k = 0; k < 128; k++) lay[TwLayout::C4 + 128 * j + k] = tw[4 * k * (j + 1)];
    for (int j = 0; j < 3; j++) for (int k = 0; k < 512; k++) lay[TwLayout::C5 + 512 * j + k] = tw[k * (j + 1)];
}

template <bool INV> static void run(const float2* in, float2* out, int* conflicts)
{
    std::vector<float2> tw(TwLayout::TOTAL), xbuf(TU);
    fill(tw.data(), INV);
    auto bank_check = [&](const std::vector<int>& addrs) {   // addrs of one half-warp (8-byte words)
        int cnt[16] = {0}; for (int a : addrs) cnt[a & 15]++;
        int mx = 0; for (int c : cnt) mx = c > mx ? c : mx; if (conflicts && mx > *conflicts) *conflicts = mx;
    };
    // pass A
    for (int t = 0; t < 128; t++)
        for (int h = 0; h < 2; h++) {
            float2 x[8], y[8];
            for (int c = 0; c < 8; c++) x[c] = in[t + 128 * h + 256 * c];
            passA_block<true, INV>(x, y, tw.data());
            int q = rev4x4(t + 128 * h);
            for (int e = 0; e < 8; e++) xbuf[swz(8 * q + e)] = y[e];
        }
    for (int w = 0; w < 8; w++) for (int h = 0; h < 2; h++) for (int e = 0; e < 8; e++) {
        std::vector<int> a; for (int l = 0; l < 16; l++) { int t = 16 * w + l; a.push_back(swz(8 * rev4x4(t + 128 * h) + e)); } bank_check(a); }
    // pass B
    for (int t = 0; t < 128; t++) {
        float2 v[16]; int kk = t & 7, base = 128 * (t >> 3) + kk;
        for (int b = 0; b < 4; b++) for (int a = 0; a < 4; a++) v[a + 4 * b] = xbuf[swz(base + 8 * a + 32 * b)];
        passB<true, INV>(v, kk, tw.data());
        for (int b = 0; b < 4; b++) for (int a = 0; a < 4; a++) xbuf[swz(base + 8 * a + 32 * b)] = v[a + 4 * b];
    }
    for (int w = 0; w < 8; w++) for (int ab = 0; ab < 16; ab++) {
        std::vector<int> a; for (int l = 0; l < 16; l++) { int t = 16 * w + l; a.push_back(swz(128 * (t >> 3) + (t & 7) + 8 * (ab & 3) + 32 * (ab >> 2))); } bank_check(a); }
    // pass C
    for (int t = 0; t < 128; t++) {
        float2 v[16];
        for (int c = 0; c < 16; c++) v[c] = xbuf[swz(t + 128 * c)];
        passC<true, INV>(v, t, tw.data());
        for (int c = 0; c < 16; c++) out[t + 128 * c] = v[c];
    }
    for (int w = 0; w < 8; w++) for (int c = 0; c < 16; c++) {
        std::vector<int> a; for (int l = 0; l < 16; l++) a.push_back(swz(16 * w + l + 128 * c)); bank_check(a); }
}

extern "C" int emul_fft2048(const float* in, float* out, int inverse)
{
    int conflicts = 0;
    if (inverse) run<true>((const float2*)in, (float2*)out, &conflicts); else run<false>((const float2*)in, (float2*)out, &conflicts);
    return conflicts;   // worst half-warp bank multiplicity of the exchange-buffer accesses (1 = conflict free)
}

extern "C" void emul_demap(const float* X, const float* P, int8_t* sre, int8_t* sim, float* r1)
{
    float2 r; demap_one<true>(make_float2(X[0], X[1]), make_float2(P[0], P[1]), *sre, *sim, r); r1[0] = r.x; r1[1] = r.y;
}
