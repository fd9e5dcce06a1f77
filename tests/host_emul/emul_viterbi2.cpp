// CPU emulation of the two-threads-per-codeword Viterbi (welle.io_b200/csrc/viterbi_core2.cuh, same header): the two threads of a pair
// are two register sets; the SHFL exchange of phase 0 is a copy.  De-puncturing tables, window expansion, traceback and packing as in
// emul_vitdec.cpp.
#include "../../welle.io_b200/csrc/viterbi_core2.cuh"
#include <cstring>
#include <vector>
using namespace dabb;
extern "C" int emul_vitdec2(const int8_t* frag, const int16_t* map, int nbits, const uint32_t* prbs_words, uint8_t* out_bytes)
{
    const int nsteps = nbits + 6, groups = nsteps / 6, nstages = (groups + 3) / 4;
    std::vector<vit_u2> steps; std::vector<uint32_t> soff;
    build_vit_tables(map, nsteps, steps, soff);
    std::vector<vit_u2> dec(nsteps);
    uint32_t Q[2][16];
    vit2_init(Q[0], 0); vit2_init(Q[1], 1);
    for (int s = 0; s < nstages; s++) {
        { const uint32_t a = vit2_local_min(Q[0]), b = vit2_local_min(Q[1]); const uint32_t m = vminu16x2(a, b); for (int t = 0; t < 2; t++) for (int r = 0; r < 16; r++) Q[t][r] -= m; }
        uint32_t row[32];
        memcpy(row, frag + (soff[s] & ~15u), 128);
        int nx = (int)((soff[s] & 15u) >> 2);
        uint32_t w0 = row[nx], w1 = row[nx + 1];
        nx += 2;
        for (int gq = 0; gq < 4; gq++) {
            const int g = 4 * s + gq;
            if (g >= groups) break;
            uint32_t w[6];
            for (int k = 0; k < 6; k++) {
                const vit_u2 e = steps[6 * g + k];
                w[k] = vit_expand_step(w0, w1, e.x, e.y);
                if (e.x & 0x10000u) { if (nx >= 32) return -1; w0 = w1; w1 = row[nx]; nx++; }
            }
            // phase 0 with the exchange, then phases 1..5
            uint32_t X[16], Y[16];
            memcpy(X, Q[0], sizeof X); memcpy(Y, Q[1], sizeof Y);
            dec[6 * g + 0].x = vit2_acs<0>(Q[0], X, Y, w[0], 0); dec[6 * g + 0].y = vit2_acs<0>(Q[1], X, Y, w[0], 1);
            for (int t = 0; t < 2; t++) {
                uint32_t* o = t ? &dec[6 * g + 1].y : &dec[6 * g + 1].x;
                o[0] = vit2_acs<1>(Q[t], X, Y, w[1], t);
                (t ? dec[6 * g + 2].y : dec[6 * g + 2].x) = vit2_acs<2>(Q[t], X, Y, w[2], t);
                (t ? dec[6 * g + 3].y : dec[6 * g + 3].x) = vit2_acs<3>(Q[t], X, Y, w[3], t);
                (t ? dec[6 * g + 4].y : dec[6 * g + 4].x) = vit2_acs<4>(Q[t], X, Y, w[4], t);
                (t ? dec[6 * g + 5].y : dec[6 * g + 5].x) = vit2_acs<5>(Q[t], X, Y, w[5], t);
            }
        }
    }
    uint32_t state = 0;
    uint32_t* out = reinterpret_cast<uint32_t*>(out_bytes);
    for (int tb = nbits - 96; tb >= 0; tb -= 96) {
        uint32_t acc[3] = {0, 0, 0};
        vit_u2 d[24];
        for (int k = 0; k < 24; k++) d[k] = dec[tb + 72 + k + 6];
        vit2_traceback24<3>(state, d, acc);
        for (int k = 0; k < 24; k++) d[k] = dec[tb + 48 + k + 6];
        vit2_traceback24<2>(state, d, acc);
        for (int k = 0; k < 24; k++) d[k] = dec[tb + 24 + k + 6];
        vit2_traceback24<1>(state, d, acc);
        for (int k = 0; k < 24; k++) d[k] = dec[tb + k + 6];
        vit2_traceback24<0>(state, d, acc);
        for (int j = 0; j < 3; j++) {
            const uint32_t v = vit_pack_be(acc[2 - j]);
            const int wi = (tb >> 5) + j;
            out[wi] = prbs_words ? v ^ prbs_words[wi] : v;
        }
    }
    return 0;
}
