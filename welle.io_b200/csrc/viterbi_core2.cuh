// viterbi_core2.cuh — the K=7 rate-1/4 add-compare-select of viterbi_core.cuh with one codeword on TWO threads (a lane pair of a
// warp): 32 states = 16 packed registers per thread.  Same semantics (backend/viterbi.cpp, see viterbi_core.cuh), same packed-16-bit
// arithmetic, same predicate-free decision words; twice the warps per codeword batch and a register footprint small enough for six
// 128-thread CTAs per SM.
//
// Layout.  A state is six bits.  At phase PH = (step mod 6) one bit position is the PAIR bit (the two halves of a 32-bit register:
// position pp = PH, as in the one-thread kernel's layout L_PH), one is the THREAD bit (which of the two threads owns the state:
// position tp = (PH + 5) mod 6) and the other four index the thread's 16 registers (in increasing position order).  A trellis step
// moves every bit up one position (new state = 2 i + input bit), the role that sat at position 5 re-enters at position 0:
//   PH = 1..4  position 5 holds a register bit: both butterfly inputs old[i], old[i+32] are this thread's registers - plain SIMD
//   PH = 5     position 5 holds the pair bit: a register is (old[i], old[i+32]); halves are duplicated with PRMT (as L_5 -> L_0)
//   PH = 0     position 5 holds the thread bit: thread 0 owns old[i], thread 1 old[i+32]; the partners exchange their 16 registers
//              with SHFL, thread 0 then computes the even new states 2 i, thread 1 the odd ones 2 i + 1
// so five of six steps need no communication.  A butterfly's branch pattern depends on the thread bit (it is one of the butterfly
// index bits, or - at PH = 0 - selects the complemented metrics): every thread permutes its eight metric words once per step
// (vit2_mc) and the butterfly code itself is identical for both threads.
#pragma once
#include "viterbi_core.cuh"

namespace dabb {

VIT_HD constexpr int vit2_pp(int ph) { return ph; }
VIT_HD constexpr int vit2_tp(int ph) { return (ph + 5) % 6; }
// register index (0..15) of state s in the layout of phase ph: its bits except the pair and the thread bit, in increasing position order
VIT_HD constexpr int vit2_reg(int s, int ph)
{
    int r = 0, k = 0;
    for (int b = 0; b < 6; b++) if (b != vit2_pp(ph) && b != vit2_tp(ph)) { r |= ((s >> b) & 1) << k; k++; }
    return r;
}
// spread the bits of j over the positions 0..4 that are not in `skip_mask`
VIT_HD constexpr int vit2_deposit(int j, int skip_mask)
{
    int out = 0, k = 0;
    for (int b = 0; b < 5; b++) if (!((skip_mask >> b) & 1)) { out |= ((j >> k) & 1) << b; k++; }
    return out;
}
VIT_HD constexpr int vit2_delta(int pos) { return pos == 5 ? 7 : vit_pat(1 << pos); }      // pattern change when butterfly index bit `pos` flips

// the eight packed metric words of a step (vit_mc) as seen by thread tau: index q of the result = word q ^ (tau ? MASK : 0)
template <int DELTA, int MASK> VIT_HD void vit2_mc(uint32_t w, uint32_t tau, uint32_t (&MC)[8])
{
#if defined(__CUDA_ARCH__)
    const uint32_t A = __dp4a(w, 0x01000001u, 0u), B = __dp4a(w, 0x00000100u, 0u), C = __dp4a(w, 0x00010000u, 0u);
#else
    const uint32_t A = (w & 0xFF) + (w >> 24), B = (w >> 8) & 0xFF, C = (w >> 16) & 0xFF;
#endif
    const uint32_t PA0 = (DELTA & 1) ? A * 0xFFFF0001u + (510u << 16) : A * 0x00010001u;
    const uint32_t PB0 = (DELTA & 2) ? B * 0xFFFF0001u + (255u << 16) : B * 0x00010001u;
    const uint32_t PC0 = (DELTA & 4) ? C * 0xFFFF0001u + (255u << 16) : C * 0x00010001u;
    const uint32_t PA1 = 0x01FE01FEu - PA0, PB1 = 0x00FF00FFu - PB0, PC1 = 0x00FF00FFu - PC0;
    const bool t = tau != 0;
    // index bit k of the word is flipped for thread 1 where MASK has bit k
    const uint32_t PA[2] = {(MASK & 1) && t ? PA1 : PA0, (MASK & 1) && t ? PA0 : PA1};
    const uint32_t PB[2] = {(MASK & 2) && t ? PB1 : PB0, (MASK & 2) && t ? PB0 : PB1};
    const uint32_t PC[2] = {(MASK & 4) && t ? PC1 : PC0, (MASK & 4) && t ? PC0 : PC1};
    uint32_t AB[4];
#pragma unroll
    for (int q = 0; q < 4; q++) AB[q] = PA[q & 1] + PB[q >> 1];
#pragma unroll
    for (int p = 0; p < 8; p++) MC[p] = AB[p & 3] + PC[p >> 2];
}

#define VIT2_T(a, b) ((a) + 0x7FFF7FFFu - (b))

// One step at phase PH for the thread with thread bit tau.  Q: its 16 registers in the layout of PH; on return in the layout of PH + 1.
// X / Y (PH == 0 only): the 16 registers of thread 0 / thread 1 of the pair (one of them is Q itself).  Returns the thread's 32
// decision bits of the step (vit2_dec_pos<PH> tells where a new state's bit is).
template <int PH> VIT_HD uint32_t vit2_acs(uint32_t (&Q)[16], const uint32_t (&X)[16], const uint32_t (&Y)[16], const uint32_t w, const uint32_t tau)
{
    uint32_t N[16];
    uint32_t W = 0u;
    constexpr int pp = vit2_pp(PH), tp = vit2_tp(PH);
    if constexpr (PH >= 1 && PH <= 4) {
        uint32_t MC[8];
        vit2_mc<vit2_delta(pp), vit2_delta(tp)>(w, tau, MC);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int ilo = vit2_deposit(j, (1 << pp) | (1 << tp));          // butterfly index with pair bit 0 (and thread bit 0: folded into MC)
            const int ra = vit2_reg(ilo, PH), rb = vit2_reg(ilo + 32, PH);
            const int p = vit_pat(ilo);
            const uint32_t m0 = Q[ra] + MC[p], m1 = Q[rb] + MC[p ^ 7], m2 = Q[ra] + MC[p ^ 7], m3 = Q[rb] + MC[p];
            const int ne = 2 * ilo;
            N[vit2_reg(ne, (PH + 1) % 6)] = vmin16(m0, m1);
            N[vit2_reg(ne + 1, (PH + 1) % 6)] = vmin16(m2, m3);
            W |= vit_signbytes(VIT2_T(m0, m1), VIT2_T(m2, m3)) & (0x01010101u << j);      // lanes: ne, ne + 1, ne + (2 << pp), ne + 1 + (2 << pp)
        }
    } else if constexpr (PH == 5) {
        uint32_t XC[8];
        vit2_mc<7, vit2_delta(tp)>(w, tau, XC);
#pragma unroll
        for (int g = 0; g < 8; g++) {
            uint32_t t[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = 2 * g + h, p = vit_pat(r);           // butterfly i = r (+ 16 tau, folded into XC); register r = (old[i], old[i+32])
                const uint32_t x = dup_lo(Q[r]) + XC[p], y = dup_hi(Q[r]) + XC[p ^ 7];
                N[r] = vmin16(x, y);                               // (new[2i], new[2i+1]): layout of phase 0, same register index
                t[h] = VIT2_T(x, y);
            }
            W |= vit_signbytes(t[0], t[1]) & (0x01010101u << g);   // lanes: 4g, 4g + 2, 4g + 1, 4g + 3 (+ 32 tau)
        }
    } else {
        // PH == 0: X[r] = (old[2r], old[2r+1]), Y[r] = (old[2r+32], old[2r+33]); this thread computes new[2i + tau]:
        // tau = 0: min(old[i] + m, old[i+32] + M - m); tau = 1: min(old[i] + M - m, old[i+32] + m)  -> the metric words complemented
        uint32_t MC[8];
        vit2_mc<vit2_delta(0), 7>(w, tau, MC);
#pragma unroll
        for (int g = 0; g < 8; g++) {
            uint32_t t[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int r = 2 * g + h, p = vit_pat(2 * r);
                const uint32_t x = X[r] + MC[p], y = Y[r] + MC[p ^ 7];
                N[r] = vmin16(x, y);                               // (new[4r + tau], new[4r + 2 + tau]): layout of phase 1, same register index
                t[h] = VIT2_T(x, y);
            }
            W |= vit_signbytes(t[0], t[1]) & (0x01010101u << g);   // lanes: 8g + tau, 8g + 4 + tau, 8g + 2 + tau, 8g + 6 + tau
        }
    }
#pragma unroll
    for (int r = 0; r < 16; r++) Q[r] = N[r];
    return W;
}

// which thread owns the decision of NEW state n produced by the step of phase PH, and at which bit of its word
template <int PH> VIT_HD void vit2_dec_pos(uint32_t n, uint32_t& thread, uint32_t& bit)
{
    if constexpr (PH >= 1 && PH <= 4) {
        thread = (n >> PH) & 1u;
        // group: the bits of n at the positions 1..5 other than PH and PH + 1, in increasing order
        uint32_t j = 0; int k = 0;
#pragma unroll
        for (int b = 1; b < 6; b++) if (b != PH && b != PH + 1) { j |= ((n >> b) & 1u) << k; k++; }
        bit = 8u * ((n & 1u) | (((n >> (PH + 1)) & 1u) << 1)) + j;
    } else if constexpr (PH == 5) {
        thread = (n >> 5) & 1u;
        bit = 8u * (((n >> 1) & 1u) | ((n & 1u) << 1)) + ((n >> 2) & 7u);
    } else {
        thread = n & 1u;
        bit = 8u * (((n >> 2) & 1u) | (((n >> 1) & 1u) << 1)) + (n >> 3);
    }
}

// start metrics in the layout of phase 0 (pair bit 0, thread bit 5): thread tau, register r = states (2r + 32 tau, 2r + 1 + 32 tau)
VIT_HD void vit2_init(uint32_t (&Q)[16], uint32_t tau)
{
#pragma unroll
    for (int r = 0; r < 16; r++) Q[r] = 63u | (63u << 16);
    if (tau == 0) Q[0] = 0u | (63u << 16);
}
// minimum over the thread's 32 states as (min, min) packed; the caller combines it with the partner's and subtracts
VIT_HD uint32_t vit2_local_min(const uint32_t (&Q)[16])
{
    uint32_t m = Q[0];
#pragma unroll
    for (int r = 1; r < 16; r++) m = vminu16x2(m, Q[r]);
    const uint32_t lo = m & 0xFFFF, hi = m >> 16;
    const uint32_t mn = lo < hi ? lo : hi;
    return mn | (mn << 16);
}

// traceback step: d = {word of thread 0, word of thread 1} of the trellis step whose phase is PH
template <int PH> VIT_HD void vit2_tb_step(uint32_t& state, const vit_u2& d)
{
    uint32_t th, pos;
    vit2_dec_pos<PH>(state, th, pos);
    const uint32_t word = th ? d.y : d.x;
#if defined(__CUDA_ARCH__)
    const uint32_t rot = __funnelshift_r(word, word, pos - 5u);
#else
    const uint32_t sh = (pos - 5u) & 31u;
    const uint32_t rot = sh ? ((word >> sh) | (word << (32 - sh))) : word;
#endif
    state = (rot & 32u) | (state >> 1);
}
template <int Q24> VIT_HD void vit2_traceback24(uint32_t& state, const vit_u2 (&d)[24], uint32_t (&acc)[3])
{
#pragma unroll
    for (int h = 3; h >= 0; h--) {
        vit2_tb_step<5>(state, d[6 * h + 5]); vit2_tb_step<4>(state, d[6 * h + 4]); vit2_tb_step<3>(state, d[6 * h + 3]);
        vit2_tb_step<2>(state, d[6 * h + 2]); vit2_tb_step<1>(state, d[6 * h + 1]); vit2_tb_step<0>(state, d[6 * h + 0]);
        const int lo = 96 - (24 * Q24 + 6 * h + 6);
        acc[lo >> 5] |= state << (lo & 31);
        if ((lo & 31) > 26) acc[(lo >> 5) + 1] |= state >> (32 - (lo & 31));
    }
}

} // namespace dabb
