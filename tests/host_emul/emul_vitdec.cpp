// CPU emulation of ONE THREAD of welle.io_b200/csrc/viterbi.cu's decoder kernel, built from the same shared header code
// (viterbi_core.cuh: expansion tables, window expansion, packed ACS, 24-step traceback, output packing): punctured softbits + map
// -> packed output bytes.  The staging (cp.async.bulk of the aligned 128 bytes per stage) is emulated by indexing the fragment from
// the aligned base.
#include "../../welle.io_b200/csrc/viterbi_core.cuh"
#include <cstring>
#include <vector>
using namespace dabb;
// frag: punctured softbits (>= -127), readable up to 160 bytes past the end; map: 4 entries per step (>= 0 present);
// out: nbits / 8 bytes; prbs_words may be null
extern "C" int emul_vitdec(const int8_t* frag, const int16_t* map, int nbits, const uint32_t* prbs_words, uint8_t* out_bytes)
{
    const int nsteps = nbits + 6, groups = nsteps / 6, nstages = (groups + 3) / 4;
    std::vector<vit_u2> steps; std::vector<uint32_t> soff;
    build_vit_tables(map, nsteps, steps, soff);
    std::vector<vit_u2> dec(nsteps);
    uint32_t Q[32]; vit_init(Q);
    for (int s = 0; s < nstages; s++) {
        vit_normalize(Q);
        uint32_t row[32];
        memcpy(row, frag + (soff[s] & ~15u), 128);                 // what the bulk copy brings in
        int nx = (int)((soff[s] & 15u) >> 2);
        uint32_t w0 = row[nx], w1 = row[nx + 1];
        nx += 2;
        for (int gq = 0; gq < 4; gq++) {
            const int g = 4 * s + gq;
            if (g >= groups) break;
            uint32_t w[6], d[12];
            for (int k = 0; k < 6; k++) {
                const vit_u2 e = steps[6 * g + k];
                w[k] = vit_expand_step(w0, w1, e.x, e.y);
                if (e.x & 0x10000u) { if (nx >= 32) return -1; w0 = w1; w1 = row[nx]; nx++; }
            }
            vit_six_steps(Q, w, d);
            for (int k = 0; k < 6; k++) { dec[6 * g + k].x = d[2 * k]; dec[6 * g + k].y = d[2 * k + 1]; }
        }
    }
    uint32_t state = 0;
    uint32_t* out = reinterpret_cast<uint32_t*>(out_bytes);
    for (int tb = nbits - 96; tb >= 0; tb -= 96) {
        uint32_t acc[3] = {0, 0, 0};
        vit_u2 d[24];
        for (int k = 0; k < 24; k++) d[k] = dec[tb + 72 + k + 6];
        vit_traceback24<3>(state, d, acc);
        for (int k = 0; k < 24; k++) d[k] = dec[tb + 48 + k + 6];
        vit_traceback24<2>(state, d, acc);
        for (int k = 0; k < 24; k++) d[k] = dec[tb + 24 + k + 6];
        vit_traceback24<1>(state, d, acc);
        for (int k = 0; k < 24; k++) d[k] = dec[tb + k + 6];
        vit_traceback24<0>(state, d, acc);
        for (int j = 0; j < 3; j++) {
            const uint32_t v = vit_pack_be(acc[2 - j]);
            const int wi = (tb >> 5) + j;
            out[wi] = prbs_words ? v ^ prbs_words[wi] : v;
        }
    }
    return 0;
}
