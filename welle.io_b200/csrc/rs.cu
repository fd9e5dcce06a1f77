// rs.cu — DAB+ superframe handling on sm_100a:
//   superframe_kernel   SuperframeFilter::Feed sliding 5-frame window (dabplus_decoder.cpp:49-105), RSDecoder::DecodeSuperframe
//                       (:326-359), CheckSync Fire code + AU start table (:171-215), AU CRCs (:122-131)
//   rs_superframes_kernel  the same without window state, for the stage-level API
// RS(120,110) over GF(2^8) (field polynomial 0x11D, first root alpha^0, 10 roots, shortened by 135): one thread per
// interleaved codeword column; syndromes by Horner, then — only when a syndrome is non-zero — Berlekamp-Massey, Chien
// search over all 255 locations and Forney, following the same algebra as libs/fec/decode_rs.h:71-298 so that the
// behaviour on uncorrectable words matches.  Latency/issue bound with negligible volume (12 codewords per 96 kbit/s
// superframe every 120 ms of signal).
#include "common.cuh"
#include "viterbi.cuh"

namespace dabb {

namespace {

struct Gf {
    const uint8_t* ex; const uint8_t* lg;
    __device__ __forceinline__ uint8_t mul(uint8_t a, uint8_t b) const { return (a && b) ? ex[lg[a] + lg[b]] : 0; }
    __device__ __forceinline__ uint8_t div(uint8_t a, uint8_t b) const { return a ? ex[lg[a] + 255 - lg[b]] : 0; }   // b == 0 acts like b == 1 (index A0 = 255 in the reference)
    __device__ __forceinline__ uint8_t apow(int e) const { e %= 255; if (e < 0) e += 255; return ex[e]; }
};

// decode column i of the interleaved superframe sf (S columns, 120 rows) in place. returns corrected count or -1
__device__ int rs_decode_column(uint8_t* sf, int S, int i, const Gf& gf)
{
    constexpr int NR = 10, NN = 255, PAD = 135, LEN = 120;
    uint8_t s[NR];
    int any = 0;
#pragma unroll
    for (int r = 0; r < NR; r++) s[r] = sf[i];
    for (int j = 1; j < LEN; j++) {
        const uint8_t d = sf[j * S + i];
#pragma unroll
        for (int r = 0; r < NR; r++) s[r] = d ^ (s[r] ? gf.ex[gf.lg[s[r]] + r] : 0);
    }
#pragma unroll
    for (int r = 0; r < NR; r++) any |= s[r];
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

struct SfShared { int corr; int uncorr; int sync; int num_aus; int au_start[7]; int au_mask; };

// RS + CheckSync + AU CRCs on the superframe in sf (shared or global), collective over the CTA
__device__ void process_superframe(uint8_t* sf, int sf_len, const Gf& gf, SfShared& sh, int t, int nthreads)
{
    const int S = sf_len / 120;
    if (t == 0) { sh.corr = 0; sh.uncorr = 0; sh.sync = 0; sh.num_aus = 0; sh.au_mask = 0; }
    __syncthreads();
    for (int i = t; i < S; i += nthreads) {
        const int c = rs_decode_column(sf, S, i, gf);
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
        }
        sh.sync = ok;
    }
    __syncthreads();
    if (sh.sync && t < sh.num_aus) {
        const uint8_t* au = sf + sh.au_start[t];
        const int alen = sh.au_start[t + 1] - sh.au_start[t];
        // an AU shorter than its own CRC cannot be checked; the reference would read out of bounds here
        if (alen >= 2 && sh.au_start[t + 1] <= sf_len) {
            const unsigned stored = au[alen - 2] << 8 | au[alen - 1];
            const unsigned calc = (~crc16_msb(au, alen - 2, 0x1021, 0xFFFF)) & 0xFFFF;
            if (stored == calc) atomicOr(&sh.au_mask, 1 << t);
        }
    }
    __syncthreads();
}

constexpr int SF_THREADS = 64;

// info layout per (stream, slot): [0] n_logical [1] n_events [2] uncorr_mask [3..6] corr [7] sf_ready [8] au_count [9] au_mask
__global__ void __launch_bounds__(SF_THREADS)
superframe_kernel(SuperframeParams p)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ SfShared sh;
    __shared__ uint8_t gfe[512], gfl[256];
    const int s = blockIdx.x, t = threadIdx.x;
    int32_t* info = p.info + (int64_t)s * 16;
    if (t < 16) info[t] = 0;
    if (p.active && !p.active[s]) return;
    MscSlotState& st = p.slots[s * p.n_slots + p.slot];
    if (!st.enabled) return;
    const int flen = 3 * st.bitrate, sf_len = 5 * flen;
    for (int i = t; i < 512; i += SF_THREADS) gfe[i] = p.gf_exp[i];
    for (int i = t; i < 256; i += SF_THREADS) gfl[i] = p.gf_log[i];
    const Gf gf{gfe, gfl};
    uint8_t* win = smem;             // raw window, sf_len
    uint8_t* sf = smem + sf_len;     // working copy
    uint8_t* gwin = p.window + (int64_t)s * p.window_pitch;
    int count = st.sf_frame_count;
    for (int i = t; i < sf_len; i += SF_THREADS) win[i] = gwin[i];
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
        for (int i = t; i < flen; i += SF_THREADS) win[(count - 1) * flen + i] = fr[i];
        __syncthreads();
        if (count < 5) continue;
        for (int i = t; i < sf_len; i += SF_THREADS) sf[i] = win[i];
        __syncthreads();
        process_superframe(sf, sf_len, gf, sh, t, SF_THREADS);
        if (t == 0) {
            info[3 + n_events] = sh.corr;
        }
        if (sh.uncorr) uncorr_mask |= 1 << n_events;
        n_events++;
        if (sh.sync) {
            uint8_t* out = p.sf_out + (int64_t)s * p.sf_pitch;
            for (int i = t; i < sf_len; i += SF_THREADS) out[i] = sf[i];
            if (t == 0) { info[7] = 1; info[8] = sh.num_aus; info[9] = sh.au_mask; }
            ready = 1;
            count = 0;
        }
        __syncthreads();
    }
    (void)ready;
    for (int i = t; i < sf_len; i += SF_THREADS) gwin[i] = win[i];
    if (t == 0) { info[0] = n_logical; info[1] = n_events; info[2] = uncorr_mask; st.sf_frame_count = count; }
}

__global__ void __launch_bounds__(SF_THREADS)
rs_superframes_kernel(uint8_t* sfs, int sf_len, int32_t* info, const uint8_t* gf_exp, const uint8_t* gf_log)
{
    __shared__ SfShared sh;
    __shared__ uint8_t gfe[512], gfl[256];
    const int t = threadIdx.x;
    for (int i = t; i < 512; i += SF_THREADS) gfe[i] = gf_exp[i];
    for (int i = t; i < 256; i += SF_THREADS) gfl[i] = gf_log[i];
    __syncthreads();
    const Gf gf{gfe, gfl};
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
    static size_t configured = 0;
    if (smem > configured) { cudaFuncSetAttribute(superframe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); configured = smem; }
    superframe_kernel<<<p.n_streams, SF_THREADS, smem, st>>>(p);
}

void launch_rs_superframes(uint8_t* sf, int n, int sf_len, int32_t* info, const uint8_t* gf_exp, const uint8_t* gf_log, cudaStream_t st)
{
    rs_superframes_kernel<<<n, SF_THREADS, 0, st>>>(sf, sf_len, info, gf_exp, gf_log);
}

} // namespace dabb
