/*
 * dab_oracle.h — CPU restatement of the welle.io DAB/DAB+ physical-layer hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped library: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load liboracle.so.
 * The product (welle.io_b200/csrc -> libdab_b200.so) never links, loads or calls it and has no CPU fallback.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit (integer stages) or bit-for-bit on the
 * float stages too (same operation order, no FMA contraction) against the unmodified reference compiled from
 * /root/reference into oracle/_ref/libwelle_ref.so (KISS-FFT option) by tests/test_oracle_vs_ref.py, and
 * against the committed golden vectors under tests/golden/ that were generated from that reference build
 * (tests/golden/make_golden.py).  The reference's libfftw3f option is NOT pinned (library absent here).
 *
 * Each function cites the reference file:line (relative to /root/reference/src) it follows.
 */
#ifndef DAB_ORACLE_H
#define DAB_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mode I constants, backend/dab-constants.cpp:351-361 */
enum { ORC_L = 76, ORC_K = 1536, ORC_TU = 2048, ORC_TS = 2552, ORC_TG = 504, ORC_TNULL = 2656, ORC_TF = 196608,
       ORC_INPUT_RATE = 2048000, ORC_CARRIER_DIFF = 1000 };

/* ---- tables ---- */
void orc_perm_table(int16_t out[ORC_K]);              /* backend/freq-interleaver.cpp:35-59 */
void orc_prs_table(float out[2 * ORC_TU]);            /* backend/phasereference.cpp:33-52, phasetable.cpp:24-183 */
void orc_pcodes(int8_t out[24 * 32]);                 /* backend/protTables.cpp:25-51 */
void orc_prbs(uint8_t* out, int n);                   /* backend/fic-handler.cpp:62-71, energy_dispersal.h:40-49 */

/* ---- FFT: restatement of libs/kiss_fft/kiss_fft.c (radix 4,4,4,4,4,2 for N=2048) ---- */
void orc_fft(int n, const float* in, float* out, int inverse);       /* unscaled, like kiss_fft() */
void orc_ifft_scaled(int n, const float* in, float* out);            /* various/fft.cpp:146-158 (1/N scaling) */

/* ---- time sync: PhaseReference::findIndex, ThresholdBeforePeak, phasereference.cpp:73-97,212-253 ---- */
int orc_find_index(const float* v /* T_u complex */, float* cir /* T_u, may be NULL */);
/* placement: 0 ThresholdBeforePeak (default), 1 StrongestPeak, 2 EarliestPeakWithBinning (phasereference.cpp:93-256) */
int orc_find_index_m(const float* v, float* cir, int placement);
/* coarse AFC: OFDMProcessor::processPRS, PatternOfZeros, ofdm-processor.cpp:537-616 */
int orc_coarse_pattern_of_zeros(const float* prs /* T_u complex */);
/* method: 0 PatternOfZeros (default), 1 GetMiddle, 2 CorrelatePRS (ofdm-processor.cpp:537-644) */
int orc_coarse(const float* prs, int method);

/* ---- OFDM demod of one aligned frame: ofdm-decoder.cpp:144-230 ---- */
void orc_ofdm_demod_frame(const float* prs /* T_u cpx */, const float* syms /* 75*T_s cpx */,
                          int8_t* soft /* 75*3072 */, float* r1s /* 75*K cpx or NULL */);
int orc_snr(const float* prs_spectrum);   /* OfdmDecoder::get_snr method 1, ofdm-decoder.cpp:240-286 */

/* ---- Viterbi K=7 rate 1/4: backend/viterbi.cpp:227-354 ---- */
void orc_viterbi(int nbits, const int8_t* in /* (nbits+6)*4 */, uint8_t* out /* nbits, one bit per byte */);
void orc_conv_encode(const uint8_t* bits, int nbits, uint8_t* out /* (nbits+6)*4 */);   /* viterbi.cpp:213-221 (commented encoder) */

/* ---- FIC: fic-handler.cpp:111-230 ---- */
int orc_check_crc_bits(const uint8_t* bits, int n);   /* various/MathHelper.h:53-80 */
void orc_fic_decode(const int8_t* soft /* 3*3072 */, uint8_t* fib_bits /* 12*256 */, uint8_t* crc_ok /* 12 */);

/* ---- MSC ---- */
typedef struct {
    int bitrate;       /* kbit/s */
    int nblk;          /* number of (L,PI) pairs (2 for EEP, 3 or 4 for UEP) */
    int L[4];          /* number of 128-position blocks */
    int PI[4];         /* puncture vector index 1..24 */
    int in_bits;       /* punctured bits per CIF = CUs*64 */
} orc_prot_t;
/* eep-protection.cpp:32-113 / uep-protection.cpp:38-167. returns 0 ok, -1 invalid */
int orc_prot_eep(int bitrate, int profile_a, int level, orc_prot_t* p);
int orc_prot_uep(int bitrate, int level, orc_prot_t* p);
/* Subchannel::bitrate(), dab-constants.cpp:404-440 (EEP only; UEP via table index handled by caller) */
int orc_eep_bitrate(int length_cu, int profile_a, int level);
/* eep-protection.cpp:115-152 / uep-protection.cpp:169-239: depuncture + Viterbi -> 24*bitrate bits */
void orc_msc_deconvolve(const orc_prot_t* p, const int8_t* in, uint8_t* outbits);
void orc_dedisperse(uint8_t* bits, int n);            /* energy_dispersal.h:35-54 */
void orc_pack_bits(const uint8_t* bits, int nbytes, uint8_t* out);   /* decoder_adapter.cpp:57-67 */

/* time de-interleaver, dab-audio.cpp:113-149 */
typedef struct {
    int fragment; int index; int count;
    int8_t* hist;    /* 16 * fragment */
} orc_deint_t;
orc_deint_t* orc_deint_new(int fragment);
void orc_deint_free(orc_deint_t*);
/* returns 1 when out is valid (after the 16-CIF warm-up) */
int orc_deint_push(orc_deint_t*, const int8_t* in, int8_t* out);

/* ---- RS(120,110) / CRC: libs/fec/{init_rs,decode_rs,encode_rs}.h, dabplus_decoder.cpp:316-359, tools.cpp:35-73 ---- */
void orc_rs_encode(const uint8_t data[110], uint8_t parity[10]);
int orc_rs_decode_codeword(uint8_t cw[120], int corr_pos[10]);   /* returns count or -1; positions include the 135 pad offset */
void orc_rs_decode_superframe(uint8_t* sf, int sf_len, int* corr, int* uncorr);
unsigned orc_crc16(const uint8_t* d, int n, unsigned poly, int init_invert, int final_invert);
unsigned orc_crc_fire(const uint8_t* d, int n);       /* poly 0x782F, no inversions */
unsigned orc_crc_ccitt(const uint8_t* d, int n);      /* poly 0x1021, init FFFF, inverted */

/* superframe filter, dabplus_decoder.cpp:49-142,171-215 */
typedef struct {
    int frame_len, frame_count, sf_len;
    uint8_t* sf_raw; uint8_t* sf;
    int num_aus; int au_start[7];
} orc_sff_t;
typedef struct {
    int attempted;     /* RS decode ran (window full) */
    int corr, uncorr;  /* FECInfo */
    int sync_ok;       /* CheckSync passed */
    int num_aus; int au_crc_ok_mask;
} orc_sff_result_t;
orc_sff_t* orc_sff_new(void);
void orc_sff_free(orc_sff_t*);
/* feeds one logical frame; if res->sync_ok, sf_out (if non-NULL) gets the post-RS superframe (sf_len bytes) */
void orc_sff_feed(orc_sff_t*, const uint8_t* frame, int len, orc_sff_result_t* res, uint8_t* sf_out);

/* ---- closed-loop receiver: OFDMProcessor::run (ofdm-processor.cpp:235-501) + everything downstream ---- */
typedef struct orc_rx orc_rx_t;
typedef struct {
    int disable_coarse;
    int fft_placement;        /* see orc_find_index_m */
    int freqsync_method;      /* see orc_coarse */
    /* one selected sub-channel (optional: set subch_len_cu = 0 for FIC only) */
    int subch_start_cu, subch_len_cu;
    orc_prot_t prot;
    int dabplus;
    int select_after_frames;  /* the sub-channel becomes active when this many frames have been decoded (reference: selection from a callback) */
    int select_after_symbol;  /* ...and from this symbol index on within that frame (processMscBlock granularity) */
} orc_rx_cfg_t;
typedef struct {
    int start_index; int fine; int coarse; int snr_raw; long frame_pos; /* sample offset of the T_u window read at SyncOnPhase */
} orc_frame_info_t;
orc_rx_t* orc_rx_new(const orc_rx_cfg_t* cfg);
void orc_rx_free(orc_rx_t*);
/* Run over iq[0..n). Outputs are appended to caller buffers (capacity in records / bytes); returns frames decoded. */
long orc_rx_run(orc_rx_t*, const float* iq, long nsamples,
                uint8_t* fibs /* 33 B per FIB: crc flag + 32 packed bytes */, long fib_cap, long* n_fibs,
                uint8_t* msc /* logical frame bytes */, long msc_cap, long* n_msc,
                int* rs_events /* (uncorr, corr) pairs */, long rs_cap, long* n_rs,
                orc_frame_info_t* finfo, long finfo_cap,
                int8_t* soft_tap /* optional: 75*3072 per frame */, long soft_cap_frames);

#ifdef __cplusplus
}
#endif
#endif
