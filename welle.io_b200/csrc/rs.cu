// rs.cu — DAB+ superframe handling on sm_100a:
//   superframe_kernel   SuperframeFilter::Feed sliding 5-frame window (dabplus_decoder.cpp:49-105), RSDecoder::DecodeSuperframe
//                       (:326-359), CheckSync Fire code + AU start table (:171-215), AU CRCs (:122-131)
//   rs_superframes_kernel  the same without window state, for the stage-level API
// RS(120,110) over GF(2^8) (field polynomial 0x11D, first root alpha^0, 10 roots, shortened by 135): one thread per
// interleaved codeword column; syndromes by Horner, then — only when a syndrome is non-zero — Berlekamp-Massey, Chien
// search over all 255 locations and Forney.  rs_decode_column is transliterated from libs/fec/decode_rs.h:159-298 (KA9Q libfec, index
// form -> polynomial form, same step sequence and early exits) because the behaviour on uncorrectable words - which positions get
// touched, what count is reported - is defined by that sequence and must match.  Latency/issue bound with negligible volume (12 codewords per 96 kbit/s
// superframe every 120 ms of signal).
#include "common.cuh"
#include "viterbi.cuh"

namespace dabb {

namespace {

struct Gf {
    const uint8_t* ex; const uint8_t* lg;
    const uint16_t* lz;      // log with lz[0] = 512; ex[512 ..] holds zeros, so ex[lz[v] + r] is v * alpha^r for every v incl. 0 (no test in the Horner loop)
    __device__ __forceinline__ uint8_t mul(uint8_t a, uint8_t b) const { return (a && b) ? ex[lg[a] + lg[b]] : 0; }
    __device__ __forceinline__ uint8_t div(uint8_t a, uint8_t b) const { return a ? ex[lg[a] + 255 - lg[b]] : 0; }   // b == 0 acts like b == 1 (index A0 = 255 in the reference)
    __device__ __forceinline__ uint8_t apow(int e) const { e %= 255; if (e < 0) e += 255; return ex[e]; }
};

constexpr int RS_MAX_COLS = 64;          // interleaved codewords per superframe = bitrate / 8

// syndromes of all S columns, one thread per (column, root): Horner over the 120 rows (decode_rs.h:89-101)
__device__ void rs_syndromes(const uint8_t* sf, int S, const Gf& gf, uint8_t* synd, int t, int nthreads)
{
    for (int idx = t; idx < S * 10; idx += nthreads) {
        const int i = idx / 10, r = idx - 10 * i;
        uint8_t v = sf[i];
        for (int j = 1; j < 120; j++) {
            const uint8_t d = sf[j * S + i];
            v = d ^ gf.ex[gf.lz[v] + r];
        }
        synd[idx] = v;
    }
}

// decode column i of the interleaved superframe sf (S columns, 120 rows) in place, given its 10 syndromes.
// returns corrected count or -1
__device__ int rs_decode_column(uint8_t* sf, int S, int i, const Gf& gf, const uint8_t* syn)
{
    constexpr int NR = 10, NN = 255, PAD = 135;
    uint8_t s[NR];
    int any = 0;
#pragma unroll
    for (int r = 0; r < NR; r++) { s[r] = syn[r]; any |= s[r]; }
    if (!any) return 0;

    uint8_t lambda[NR + 1], b[NR + 1], t[NR + 1];
    for (int k = 0; k <= NR; k++) { lambda[k] = 0; b[k] = 0; }
    lambda[0] = 1; b[0] = 1;
    int el = 0;
    for (int r = 1; r <= NR; r++) {
        uint8_t discr = 0;
        for (int k = 0; k < r; k++) discr ^= gf.mul(lambda[k], s[r - k - 1]);
        if (discr == 0) {
            for (int k = NR; k > 0; k--) b[k] = b[k - 1];
            b[0] = 0;
        } else {
            t[0] = lambda[0];
            for (int k = 0; k < NR; k++) t[k + 1] = lambda[k + 1] ^ gf.mul(discr, b[k]);
            if (2 * el <= r - 1) {
                el = r - el;
                for (int k = 0; k <= NR; k++) b[k] = gf.div(lambda[k], discr);
            } else {
                for (int k = NR; k > 0; k--) b[k] = b[k - 1];
                b[0] = 0;
            }
            for (int k = 0; k <= NR; k++) lambda[k] = t[k];
        }
    }
    int deg = 0;
    for (int k = 0; k <= NR; k++) if (lambda[k]) deg = k;
    int root[NR], loc[NR], count = 0;
    for (int q = 1, k = 0; q <= NN; q++, k = (k + 1) % NN) {
        uint8_t v = 1;
        for (int j = deg; j > 0; j--) v ^= gf.mul(lambda[j], gf.apow(j * q));
        if (v) continue;
        root[count] = q; loc[count] = k;
        if (++count == deg) break;
    }
    if (deg != count) return -1;
    uint8_t omega[NR + 1];
    const int dego = deg - 1;
    for (int k = 0; k <= dego; k++) {
        uint8_t tmp = 0;
        for (int j = k; j >= 0; j--) tmp ^= gf.mul(s[k - j], lambda[j]);
        omega[k] = tmp;
    }
    for (int j = count - 1; j >= 0; j--) {
        uint8_t num1 = 0;
        for (int k = dego; k >= 0; k--) num1 ^= gf.mul(omega[k], gf.apow(k * root[j]));
        const uint8_t num2 = gf.apow(root[j] * (0 - 1) + NN);
        uint8_t den = 0;
        const int top = (deg < NR - 1 ? deg : NR - 1) & ~1;
        for (int k = top; k >= 0; k -= 2) den ^= gf.mul(lambda[k + 1], gf.apow(k * root[j]));
        if (num1 != 0 && loc[j] >= PAD) sf[(loc[j] - PAD) * S + i] ^= gf.div(gf.mul(num1, num2), den);
    }
    return count;
}

__device__ unsigned crc16_msb(const uint8_t* d, int n, unsigned poly, unsigned crc)
{
    for (int i = 0; i < n; i++) {
        crc ^= (unsigned)d[i] << 8;
#pragma unroll
        for (int b = 0; b < 8; b++) crc = (crc & 0x8000) ? ((crc << 1) ^ poly) & 0xFFFF : (crc << 1) & 0xFFFF;
    }
    return crc;
}

// a * b mod x^16 + x^12 + x^5 + 1 over GF(2)
__device__ __forceinline__ unsigned ccitt_mulmod(unsigned a, unsigned b)
{
    unsigned r = 0;
#pragma unroll
    for (int bit = 15; bit >= 0; bit--) {
        r = ((r << 1) ^ ((r & 0x8000u) ? 0x1021u : 0u)) & 0xFFFFu;
        if ((b >> bit) & 1u) r ^= a;
    }
    return r;
}

struct SfShared {
    int corr; int uncorr; int sync; int num_aus; int au_start[7]; int au_mask;
    unsigned au_crc[6]; int chunk_base[7];
    uint16_t crc_tab[256];         // CRC-16-CCITT byte table
    uint16_t y_pow[512];           // x^(128 k) mod P: weight of a 16-byte chunk that has 16 k bytes behind it
    uint8_t synd[10 * RS_MAX_COLS];
};

// compile-time tables for the chunked CRC
constexpr unsigned ccitt_step(unsigned r) { return ((r << 1) ^ ((r & 0x8000u) ? 0x1021u : 0u)) & 0xFFFFu; }
constexpr unsigned ccitt_mulmod_c(unsigned a, unsigned b)
{
    unsigned r = 0;
    for (int bit = 15; bit >= 0; bit--) { r = ccitt_step(r); if ((b >> bit) & 1u) r ^= a; }
    return r;
}
struct CrcTabs {
    uint16_t tab[256]; uint16_t ypow[512];
    constexpr CrcTabs() : tab{}, ypow{}
    {
        for (int k = 0; k < 256; k++) { unsigned c = (unsigned)k << 8; for (int b = 0; b < 8; b++) c = ccitt_step(c); tab[k] = (uint16_t)c; }
        unsigned y = 1;
        for (int q = 0; q < 128; q++) y = ccitt_step(y);        // x^128 mod P
        unsigned acc = 1;
        for (int k = 0; k < 512; k++) { ypow[k] = (uint16_t)acc; acc = ccitt_mulmod_c(acc, y); }
    }
};
__device__ const CrcTabs g_crc_tabs{};

__device__ void sf_tables_init(SfShared& sh, int t, int nthreads)
{
    for (int k = t; k < 256; k += nthreads) sh.crc_tab[k] = g_crc_tabs.tab[k];
    for (int k = t; k < 512; k += nthreads) sh.y_pow[k] = g_crc_tabs.ypow[k];
}

// RS + CheckSync + AU CRCs on the superframe in sf (shared or global), collective over the CTA.
// The AU CRCs (CRC-16-CCITT, init 0xFFFF, inverted; dabplus_decoder.cpp:122-131) are computed in 16-byte chunks counted
// from the end of each AU: reg = sum_k rawcrc(chunk_k) * x^(128 k) mod P, the initial value folded into the first two bytes.
__device__ void process_superframe(uint8_t* sf, int sf_len, const Gf& gf, SfShared& sh, int t, int nthreads)
{
    const int S = sf_len / 120;
    if (t == 0) { sh.corr = 0; sh.uncorr = 0; sh.sync = 0; sh.num_aus = 0; sh.au_mask = 0; }
    if (t < 6) sh.au_crc[t] = 0;
    rs_syndromes(sf, S, gf, sh.synd, t, nthreads);
    __syncthreads();
    for (int i = t; i < S; i += nthreads) {
        const int c = rs_decode_column(sf, S, i, gf, sh.synd + 10 * i);
        if (c < 0) atomicOr(&sh.uncorr, 1); else if (c > 0) atomicAdd(&sh.corr, c);
    }
    __syncthreads();
    if (t == 0) {
        int ok = !(sf[3] == 0 && sf[4] == 0);
        if (ok) ok = ((unsigned)(sf[0] << 8 | sf[1]) == crc16_msb(sf + 2, 9, 0x782F, 0));
        if (ok) {
            const int dac = sf[2] & 0x40, sbr = sf[2] & 0x20;
            const int na = dac ? (sbr ? 3 : 6) : (sbr ? 2 : 4);
            sh.num_aus = na;
            sh.au_start[0] = dac ? (sbr ? 6 : 11) : (sbr ? 5 : 8);
            sh.au_start[na] = sf_len / 120 * 110;
            sh.au_start[1] = sf[3] << 4 | sf[4] >> 4;
            if (na >= 3) sh.au_start[2] = (sf[4] & 0x0F) << 8 | sf[5];
            if (na >= 4) sh.au_start[3] = sf[6] << 4 | sf[7] >> 4;
            if (na == 6) { sh.au_start[4] = (sf[7] & 0x0F) << 8 | sf[8]; sh.au_start[5] = sf[9] << 4 | sf[10] >> 4; }
            for (int k = 0; k < na; k++) if (sh.au_start[k] >= sh.au_start[k + 1]) ok = 0;
            // chunk directory for the CRC pass; an AU shorter than its own CRC cannot be checked (the reference would
            // read out of bounds here), one that ends beyond the superframe neither
            int base = 0;
            for (int k = 0; k < na; k++) {
                sh.chunk_base[k] = base;
                const int alen = sh.au_start[k + 1] - sh.au_start[k];
                if (ok && alen >= 2 && sh.au_start[k + 1] <= sf_len) base += (alen - 2 + 15) / 16;
            }
            sh.chunk_base[na] = base;
        }
        sh.sync = ok;
    }
    __syncthreads();
    if (sh.sync) {
        const int na = sh.num_aus, total = sh.chunk_base[na];
        for (int idx = t; idx < total; idx += nthreads) {
            int a = 0;
            while (a + 1 < na && idx >= sh.chunk_base[a + 1]) a++;
            const int k = idx - sh.chunk_base[a];                    // chunk k has 16 k bytes behind it
            const uint8_t* au = sf + sh.au_start[a];
            const int L = sh.au_start[a + 1] - sh.au_start[a] - 2;   // bytes covered by the CRC
            const int hi = L - 16 * k, lo = hi - 16 > 0 ? hi - 16 : 0;
            unsigned crc = 0;
            for (int q = lo; q < hi; q++) {
                unsigned byte = au[q];
                if (q < 2 && L >= 2) byte ^= 0xFFu;                  // initial value 0xFFFF
                crc = ((crc << 8) ^ sh.crc_tab[((crc >> 8) ^ byte) & 0xFFu]) & 0xFFFFu;
            }
            if (L < 2) crc = crc16_msb(au, L, 0x1021, 0xFFFF);       // degenerate AU: single chunk, plain loop
            else if (k) crc = ccitt_mulmod(crc, sh.y_pow[k & 511]);
            atomicXor(&sh.au_crc[a], crc);
        }
    }
    __syncthreads();
    if (sh.sync && t < sh.num_aus) {
        const uint8_t* au = sf + sh.au_start[t];
        const int alen = sh.au_start[t + 1] - sh.au_start[t];
        if (alen >= 2 && sh.au_start[t + 1] <= sf_len) {
            const unsigned stored = au[alen - 2] << 8 | au[alen - 1];
            unsigned reg = sh.au_crc[t];
            if (alen == 2) reg = 0xFFFF;                             // empty message: the register is still the initial value
            if (stored == ((~reg) & 0xFFFF)) atomicOr(&sh.au_mask, 1 << t);
        }
    }
    __syncthreads();
}

constexpr int SF_THREADS = 128;

// info layout per (stream, slot): [0] n_logical [1] n_events [2] uncorr_mask [3..6] corr [7] sf_ready [8] au_count [9] au_mask
__global__ void __launch_bounds__(SF_THREADS)
superframe_kernel(SuperframeParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ SfShared sh;
    __shared__ __align__(16) uint8_t gfe[512 + 16], gfl[256];
    __shared__ uint16_t gfz[256];
    const int s = blockIdx.x, t = threadIdx.x;
    int32_t* info = p.info + (int64_t)s * 16;
    if (t < 16) info[t] = 0;
    if (p.active && !p.active[s]) return;
    MscSlotState& st = p.slots[s * p.n_slots + p.slot];
    if (!st.enabled) return;
    const int flen = 3 * st.bitrate, sf_len = 5 * flen;      // multiples of 24 / 120 bytes: all copies below go by words
    int count = st.sf_frame_count;
    {   // the tables are only needed when this step can complete a 5-frame window
        int nv = 0;
        for (int c = 0; c < 4; c++) nv += p.valid[s * 4 + c] ? 1 : 0;
        if (st.dabplus && count + nv >= 5) {
            for (int i = t; i < 128; i += SF_THREADS) reinterpret_cast<uint32_t*>(gfe)[i] = reinterpret_cast<const uint32_t*>(p.gf_exp)[i];
            for (int i = t; i < 64; i += SF_THREADS) reinterpret_cast<uint32_t*>(gfl)[i] = reinterpret_cast<const uint32_t*>(p.gf_log)[i];
            for (int i = t; i < 256; i += SF_THREADS) gfz[i] = i ? (uint16_t)p.gf_log[i] : (uint16_t)512;
            if (t < 16) gfe[512 + t] = 0;
            sf_tables_init(sh, t, SF_THREADS);
        }
    }
    const Gf gf{gfe, gfl, gfz};
    uint8_t* win = smem;             // raw window, sf_len
    uint8_t* sf = smem + sf_len;     // working copy
    uint8_t* gwin = p.window + (int64_t)s * p.window_pitch;
    for (int i = t; i < sf_len / 4; i += SF_THREADS) reinterpret_cast<uint32_t*>(win)[i] = reinterpret_cast<const uint32_t*>(gwin)[i];
    __syncthreads();
    int n_logical = 0, n_events = 0, uncorr_mask = 0, ready = 0;
    for (int c = 0; c < 4; c++) {
        if (!p.valid[s * 4 + c]) continue;
        n_logical++;
        if (!st.dabplus) continue;
        const uint8_t* fr = p.logical + (int64_t)(s * 4 + c) * p.logical_stride;
        if (count == 5) {
            // slide by one frame (forward copy: safe because each thread strides in increasing order after a barrier)
            for (int base = 0; base < 4 * flen; base += SF_THREADS) {
                const int i = base + t;
                uint8_t v = 0;
                if (i < 4 * flen) v = win[i + flen];
                __syncthreads();
                if (i < 4 * flen) win[i] = v;
                __syncthreads();
            }
        } else count++;
        for (int i = t; i < flen / 4; i += SF_THREADS) reinterpret_cast<uint32_t*>(win + (count - 1) * flen)[i] = reinterpret_cast<const uint32_t*>(fr)[i];
        __syncthreads();
        if (count < 5) continue;
        for (int i = t; i < sf_len / 4; i += SF_THREADS) reinterpret_cast<uint32_t*>(sf)[i] = reinterpret_cast<const uint32_t*>(win)[i];
        __syncthreads();
        process_superframe(sf, sf_len, gf, sh, t, SF_THREADS);
        if (t == 0) {
            info[3 + n_events] = sh.corr;
        }
        if (sh.uncorr) uncorr_mask |= 1 << n_events;
        n_events++;
        if (sh.sync) {
            uint8_t* out = p.sf_out + (int64_t)s * p.sf_pitch;
            for (int i = t; i < sf_len / 4; i += SF_THREADS) reinterpret_cast<uint32_t*>(out)[i] = reinterpret_cast<const uint32_t*>(sf)[i];
            if (t == 0) { info[7] = 1; info[8] = sh.num_aus; info[9] = sh.au_mask; }
            ready = 1;
            count = 0;
        }
        __syncthreads();
    }
    (void)ready;
    for (int i = t; i < sf_len / 4; i += SF_THREADS) reinterpret_cast<uint32_t*>(gwin)[i] = reinterpret_cast<const uint32_t*>(win)[i];
    if (t == 0) { info[0] = n_logical; info[1] = n_events; info[2] = uncorr_mask; st.sf_frame_count = count; }
}

__global__ void __launch_bounds__(SF_THREADS)
rs_superframes_kernel(uint8_t* sfs, int sf_len, int32_t* info, const uint8_t* gf_exp, const uint8_t* gf_log)
{
    __shared__ SfShared sh;
    __shared__ __align__(16) uint8_t gfe[512 + 16], gfl[256];
    __shared__ uint16_t gfz[256];
    const int t = threadIdx.x;
    for (int i = t; i < 512; i += SF_THREADS) gfe[i] = gf_exp[i];
    if (t < 16) gfe[512 + t] = 0;
    for (int i = t; i < 256; i += SF_THREADS) { gfl[i] = gf_log[i]; gfz[i] = i ? (uint16_t)gf_log[i] : (uint16_t)512; }
    sf_tables_init(sh, t, SF_THREADS);
    __syncthreads();
    const Gf gf{gfe, gfl, gfz};
    uint8_t* sf = sfs + (int64_t)blockIdx.x * sf_len;
    process_superframe(sf, sf_len, gf, sh, t, SF_THREADS);
    if (t == 0) {
        int32_t* o = info + 4 * (int64_t)blockIdx.x;
        o[0] = sh.corr; o[1] = sh.uncorr; o[2] = sh.sync; o[3] = sh.au_mask | (sh.num_aus << 8);
    }
}

} // namespace

void launch_superframe(const SuperframeParams& p, cudaStream_t st)
{
    const size_t smem = 2 * (size_t)p.window_pitch;
    if (smem > 48 * 1024) cudaFuncSetAttribute(superframe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     // per device: no caching across contexts
    superframe_kernel<<<p.n_streams, SF_THREADS, smem, st>>>(p);
}

void launch_rs_superframes(uint8_t* sf, int n, int sf_len, int32_t* info, const uint8_t* gf_exp, const uint8_t* gf_log, cudaStream_t st)
{
    rs_superframes_kernel<<<n, SF_THREADS, 0, st>>>(sf, sf_len, info, gf_exp, gf_log);
}

} // namespace dabb
