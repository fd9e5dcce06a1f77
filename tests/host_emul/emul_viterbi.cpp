// CPU emulation of the per-thread packed-16-bit Viterbi of welle.io_b200/csrc/viterbi_core.cuh (same header).
#include "../../welle.io_b200/csrc/viterbi_core.cuh"
#include <vector>
using namespace dabb;
extern "C" void emul_viterbi(int nbits, const int8_t* soft, uint8_t* out)
{
    const int nsteps = nbits + 6;
    std::vector<uint32_t> dl(nsteps), dh(nsteps);
    uint32_t Q[32]; vit_init(Q);
    for (int g = 0; g < nsteps / 6; g++) {
        if (g % 4 == 0) vit_normalize(Q);
        uint32_t w[6], dec[12];
        for (int s = 0; s < 6; s++) { const int8_t* p = soft + 4 * (6 * g + s); w[s] = vit_sym(p[0]) | (vit_sym(p[1]) << 8) | (vit_sym(p[2]) << 16) | (vit_sym(p[3]) << 24); }
        vit_six_steps(Q, w, dec);
        for (int s = 0; s < 6; s++) { dl[6 * g + s] = dec[2 * s]; dh[6 * g + s] = dec[2 * s + 1]; }
    }
    unsigned state = 0;
    for (int t = nbits - 1; t >= 0; t--) {
        const uint64_t d = (uint64_t)dl[t + 6] | ((uint64_t)dh[t + 6] << 32);
        unsigned pos = 0;
        switch ((t + 6) % 6) {       // the decision layout of step s is that of vit_acs<s % 6>
            case 0: pos = vit_dec_pos<0>(state); break; case 1: pos = vit_dec_pos<1>(state); break; case 2: pos = vit_dec_pos<2>(state); break;
            case 3: pos = vit_dec_pos<3>(state); break; case 4: pos = vit_dec_pos<4>(state); break; default: pos = vit_dec_pos<5>(state); break;
        }
        const unsigned k = (unsigned)(d >> pos) & 1u;
        state = (state >> 1) | (k << 5);
        out[t] = (uint8_t)k;
    }
}
