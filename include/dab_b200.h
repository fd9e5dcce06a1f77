/*
 * dab_b200.h — C ABI of the B200-native DAB/DAB+ physical-layer decode path (libdab_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ / torch types.  It is what the host glue
 * (welle.io_b200/host/radio-receiver.{h,cpp}, which mirrors the reference's RadioReceiver surface) and any other
 * FFI (ctypes in welle.io_b200/dabb200.py, or a cgo/JNI stub) binds.  Every entry point names the reference
 * interface it replaces (paths relative to /root/reference/src).
 *
 * Conventions: every function returns 0 on success or a negative DABB_E_* code; dabb_last_error() gives the text.
 * No CPU fallback exists: without a CUDA device dabb_create fails with DABB_E_NODEVICE.
 * One submitting thread per context.  All device work of a context runs on the context's own CUDA stream.
 *
 * Streams: a context decodes `n_streams` independent ensembles ("streams") in lock-step, one transmission
 * frame (96 ms, 196 608 samples) per stream per dabb_process() call.  Per-stream receiver state (sample position,
 * NCO phase, fine/coarse frequency correctors, FIC success counter, 16-CIF time de-interleaver history,
 * 5-frame superframe window) lives on the device between calls.
 */
#ifndef DAB_B200_H
#define DAB_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DABB_ABI_VERSION 3

enum {
    DABB_OK = 0,
    DABB_E_NODEVICE = -1,   /* no CUDA device / driver */
    DABB_E_CUDA = -2,       /* CUDA runtime error (see dabb_last_error) */
    DABB_E_ARG = -3,        /* invalid argument */
    DABB_E_STATE = -4,      /* call not valid in this state */
    DABB_E_NOMEM = -5,
    DABB_E_UNSUPPORTED = -6 /* e.g. transmission mode other than I */
};

/* Mode I constants (backend/dab-constants.cpp:351-361) */
enum { DABB_L = 76, DABB_K = 1536, DABB_TU = 2048, DABB_TS = 2552, DABB_TG = 504, DABB_TNULL = 2656, DABB_TF = 196608,
       DABB_SOFT_PER_FRAME = 75 * 3072, DABB_FIB_BYTES = 32, DABB_FIBS_PER_FRAME = 12, DABB_MAX_SUBCH = 4 };

typedef struct dabb_ctx dabb_ctx;

/* FFT arithmetic: DABB_FFT_EXACT reproduces the reference's KISS-FFT build bit for bit (same radix-4/2
 * factorisation, separate multiply/add roundings); DABB_FFT_FMA lets the compiler contract to FMA
 * (<= 1e-6 relative difference, faster). */
enum { DABB_FFT_EXACT = 0, DABB_FFT_FMA = 1 };
/* Oscillator arithmetic: DABB_NCO_EXACT reproduces every entry of the reference's 2 048 000-entry table bit for bit (verified
 * exhaustively at dabb_create); DABB_NCO_FAST evaluates the same phase with fp32 sincospi (each oscillator sample within ~1e-6 of
 * the table value: soft-decision intermediates within 1e-4 relative, decoded FIC/MSC bytes unchanged in the tests). */
enum { DABB_NCO_EXACT = 0, DABB_NCO_FAST = 1 };

/* RadioReceiverOptions::fftPlacementMethod / freqsyncMethod (backend/radio-receiver-options.h:35-64).  0 is the reference's
 * default in both enums (the reference's own enumerator values differ: see INTEGRATION.md for the mapping). */
enum { DABB_PLACEMENT_THRESHOLD_BEFORE_PEAK = 0, DABB_PLACEMENT_STRONGEST_PEAK = 1, DABB_PLACEMENT_EARLIEST_PEAK_WITH_BINNING = 2 };
enum { DABB_FREQSYNC_PATTERN_OF_ZEROS = 0, DABB_FREQSYNC_GET_MIDDLE = 1, DABB_FREQSYNC_CORRELATE_PRS = 2 };

typedef struct {
    int32_t abi_version;        /* DABB_ABI_VERSION */
    int32_t device;             /* CUDA device ordinal */
    int32_t n_streams;          /* streams decoded per dabb_process call */
    int32_t transmission_mode;  /* only 1 (RadioReceiver ctor arg, backend/radio-receiver.cpp:66-80) */
    int32_t fft_mode;           /* DABB_FFT_* */
    int32_t disable_coarse;     /* RadioReceiverOptions::disableCoarseCorrector (radio-receiver-options.h:66-85) */
    int32_t keep_taps;          /* 1: keep CIR / softbits / constellation / null symbol of the last frame readable via dabb_read_tap */
    int32_t n_subch_slots;      /* sub-channels decodable per stream at once (1..DABB_MAX_SUBCH), 0 -> 1 */
    int32_t max_subch_cu;       /* largest selectable sub-channel in capacity units, 0 -> 144 */
    int32_t ofdm_groups;        /* CTAs per frame in the OFDM kernel (divisor of 75), 0 -> chosen from n_streams */
    int32_t fft_placement;      /* DABB_PLACEMENT_* (PhaseReference::findIndex variant) */
    int32_t freqsync_method;    /* DABB_FREQSYNC_* (OFDMProcessor::processPRS variant) */
    int32_t nco_mode;           /* DABB_NCO_*: arithmetic of the frequency-correcting oscillator (ofdm-processor.cpp:92-94,211-214) */
    int32_t ofdm_tail_split;    /* 0: automatic (large batches cut their last frames into short CTAs), -1: off */
    int32_t reserved[2];
} dabb_config;

/* replaces: RadioReceiver::setReceiverOptions (backend/radio-receiver.cpp:119-124): takes effect from the next frame */
typedef struct {
    int32_t disable_coarse;
    int32_t fft_placement;      /* DABB_PLACEMENT_* */
    int32_t freqsync_method;    /* DABB_FREQSYNC_* */
    int32_t decode_tii;         /* RadioReceiverOptions::decodeTII: with keep_taps, also produce tap 4 (the two spectra TIIDecoder starts from) */
    int32_t reserved[4];
} dabb_options;

/* replaces: RadioReceiver::RadioReceiver / ~RadioReceiver (backend/radio-receiver.h:52-116) */
int dabb_create(const dabb_config* cfg, dabb_ctx** out);
void dabb_destroy(dabb_ctx* ctx);
const char* dabb_last_error(const dabb_ctx* ctx);   /* ctx may be NULL for create failures */
int dabb_abi_version(void);

/* replaces: OFDMProcessor::restart (backend/ofdm-processor.cpp:115-132): stream goes back to acquisition at
 * logical sample position `pos` with zeroed correctors, cleared FIC counter, de-interleaver and superframe window */
int dabb_stream_reset(dabb_ctx* ctx, int32_t first_stream, int32_t count, int64_t pos);
int dabb_set_options(dabb_ctx* ctx, const dabb_options* opt);
/* introspection.  DABB_INFO_OSC_MODE: 1 = the oscillator of OFDMProcessor::getSamples (ofdm-processor.cpp:92-94,211-214) is
 * evaluated on the fly, which dabb_create allows only after comparing it on the device with the 2 048 000-entry table for every
 * index (DABB_INFO_OSC_MISMATCHES must be 0), 0 = table lookups; DABB_INFO_OSC_PATCHED: how many of the 2000 double-precision
 * factors had to be replaced by the table's own value for that (3: the quarter turns) */
enum { DABB_INFO_OSC_MODE = 0, DABB_INFO_OSC_MISMATCHES = 1, DABB_INFO_OSC_PATCHED = 2 };
int dabb_get_info(dabb_ctx* ctx, int32_t what, int64_t* out);

/* replaces: MscHandler::addSubchannel / removeSubchannel (backend/msc-handler.cpp:61-127) + DabAudio ctor
 * (backend/dab-audio.cpp:46-85).  slot in [0, DABB_MAX_SUBCH).  Protection exactly as ProtectionSettings
 * (backend/dab-constants.h:152-166): short_form -> UEP (bitrate + uep_level), else EEP profile/level. */
typedef struct {
    int32_t subch_id;     /* informational */
    int32_t start_cu;     /* Subchannel::startAddr */
    int32_t length_cu;    /* Subchannel::length */
    int32_t bitrate;      /* kbit/s = Subchannel::bitrate() */
    int32_t short_form;   /* 1 = UEP */
    int32_t uep_level;    /* ProtectionSettings::uepLevel (1..5) */
    int32_t eep_profile_a;/* 1 = EEP-A, 0 = EEP-B */
    int32_t eep_level;    /* 1..4 */
    int32_t dabplus;      /* 1 = DAB+ (RS + Fire code + AU CRC), 0 = DAB MP2 (logical frames only) */
} dabb_subchannel;
int dabb_select_subchannel(dabb_ctx* ctx, int32_t first_stream, int32_t count, int32_t slot, const dabb_subchannel* sc);
int dabb_remove_subchannel(dabb_ctx* ctx, int32_t first_stream, int32_t count, int32_t slot);

/* Per-stream, per-call result record (POD). */
typedef struct {
    int32_t status;         /* DABB_FRAME_* */
    int32_t start_index;    /* PhaseReference::findIndex result (phasereference.cpp:73) */
    int32_t fine_corr;      /* OFDMProcessor::fineCorrector after the frame (ofdm-processor.cpp:450) */
    int32_t coarse_corr;    /* OFDMProcessor::coarseCorrector */
    int32_t snr_raw;        /* OfdmDecoder::get_snr value for this PRS (ofdm-decoder.cpp:240) */
    int32_t fib_crc_mask;   /* bit f set = FIB f CRC ok (fic-handler.cpp:214-229) */
    int32_t fic_ratio;      /* FicHandler::getFicDecodeRatioPercent()/10 */
    int32_t n_logical[DABB_MAX_SUBCH];     /* logical frames produced this call per slot (0..4) */
    int32_t n_rs_events[DABB_MAX_SUBCH];   /* RS decode attempts this call per slot (0..4) */
    int32_t rs_uncorr_mask[DABB_MAX_SUBCH];/* bit e = attempt e had uncorrectable codewords */
    int32_t rs_corr[DABB_MAX_SUBCH][4];    /* corrected symbols per attempt (RSDecoder::DecodeSuperframe) */
    int32_t sf_ready[DABB_MAX_SUBCH];      /* 1 = a superframe passed the Fire-code sync check this call */
    int32_t sf_au_count[DABB_MAX_SUBCH];
    int32_t sf_au_crc_mask[DABB_MAX_SUBCH];/* bit i = AU i CRC ok (dabplus_decoder.cpp:122-131) */
    int64_t next_pos;       /* logical sample position where the next frame's T_u read starts */
    float   freq_corr_re, freq_corr_im;    /* FreqCorr accumulator (ofdm-processor.cpp:435-442) */
    float   slevel;
    int32_t acq_failed;     /* null searches that failed inside this call (each one re-enters OFDMProcessor::run's notSynced state,
                               ofdm-processor.cpp:253-323: what RadioReceiver::restart(doScan) counts for onSignalPresence) */
    int32_t reserved[2];
} dabb_frame_result;

enum { DABB_FRAME_DECODED = 0, DABB_FRAME_NEED_SAMPLES = 1, DABB_FRAME_NO_SYNC = 2, DABB_FRAME_ACQUIRING = 3 };

/* One decode step for all streams.
 *   iq            interleaved (re,im) samples in `iq_format` (default cf32), stream s at iq + s*stride_samples complex samples
 *   iq_is_host    0: device pointer; 1: host pointer (pinned or pageable) -> copied H2D inside the call
 *   buf_start[s]  logical sample index of the first sample in stream s's buffer (host array, n_streams entries)
 *   buf_len       complex samples available per stream buffer
 * The stream consumes [pos, pos + T_u + startIndex + 75*T_s + T_null) of its logical sample axis; if the buffer
 * does not cover what is needed the stream reports DABB_FRAME_NEED_SAMPLES and its state is unchanged.
 * Outputs are written to host memory (any of them may be NULL):
 *   results [n_streams]
 *   fibs    [n_streams][12][32]   FIB bytes, MSB first (RadioControllerInterface::onFIBDecodeSuccess payload packed)
 *   msc     [n_streams][DABB_MAX_SUBCH][4][msc_stride] logical frames (3*bitrate bytes each; DecoderAdapter::addtoFrame data)
 *   sf      [n_streams][DABB_MAX_SUBCH][sf_stride]     post-RS superframe (15*bitrate bytes) when sf_ready
 * replaces: one iteration of OFDMProcessor::run (ofdm-processor.cpp:324-490) + OfdmDecoder::workerthread
 * (ofdm-decoder.cpp:93-130) + FicHandler::processFicBlock + MscHandler::processMscBlock + DabAudio::run +
 * DecoderAdapter::addtoFrame + SuperframeFilter::Feed, for n_streams receivers at once.
 */
typedef struct {
    const float* iq; int32_t iq_is_host; int64_t stride_samples; const int64_t* buf_start; int64_t buf_len;
    dabb_frame_result* results; uint8_t* fibs; uint8_t* msc; int32_t msc_stride; uint8_t* sf; int32_t sf_stride;
    int32_t iq_format;   /* DABB_IQ_*: raw sample format of `iq` (strides and lengths stay in complex samples); converted on the
                            device exactly like CRAWFile::convertSamples (input/raw_file.cpp:324-366) */
    int32_t carry_samples;   /* dabb_submit only: the first `carry_samples` complex samples of every stream's window are the LAST
                                `carry_samples` of the window given to the previous dabb_submit (a receiver reads its input once; the overlap
                                a frame needs is kept on the device): they are taken from the previous staging slot instead of being copied
                                host->device again.  0 = copy the whole window.  Ignored by dabb_process / dabb_process_async. */
} dabb_io;
enum { DABB_IQ_CF32 = 0, DABB_IQ_U8 = 1, DABB_IQ_S8 = 2, DABB_IQ_S16LE = 3, DABB_IQ_S16BE = 4 };
int dabb_process(dabb_ctx* ctx, const dabb_io* io);
/* asynchronous form for benchmarking with device-resident inputs: enqueue only, no host copies */
int dabb_process_async(dabb_ctx* ctx, const dabb_io* io);
int dabb_sync(dabb_ctx* ctx);
/* Pipelined form of dabb_process for HOST buffers (io->iq_is_host must be 1; pinned memory for full overlap): dabb_submit enqueues the
 * host->device copy of the step's samples on a copy stream (two device staging slots), the kernels, and the device->host copy of the
 * results into pinned staging, and returns without waiting; at most two steps may be outstanding (DABB_E_STATE otherwise).
 * dabb_collect waits for the OLDEST outstanding step and writes its results to the output pointers given at that dabb_submit (they
 * must stay valid until then; io->iq must stay valid until the step has been collected).  Results come back in submit order; the
 * copy of step n+1 overlaps the kernels of step n and the read-back of step n-1. */
int dabb_submit(dabb_ctx* ctx, const dabb_io* io);
int dabb_collect(dabb_ctx* ctx);
/* makes the context's main stream (dabb_cuda_stream) wait for everything enqueued so far on the library's other streams, so that an
 * event recorded on it afterwards marks the completion of all work of the preceding dabb_process_async calls (timing) */
int dabb_join_lanes(dabb_ctx* ctx);
void* dabb_cuda_stream(dabb_ctx* ctx);              /* cudaStream_t of the context */
int64_t dabb_kernel_launches(const dabb_ctx* ctx);  /* number of CUDA kernels this context has launched */

/* per-kernel device timing: while enabled, one CUDA event is recorded on the context's stream after every kernel launch;
 * dabb_profile_read returns {"kernel": {"ms": total, "n": launches}, ...} as JSON text */
int dabb_profile(dabb_ctx* ctx, int32_t enable);
int dabb_profile_read(dabb_ctx* ctx, char* json_out, size_t cap);

/* taps (valid after a dabb_process when keep_taps=1):
 *   0 = softbits int8 [n_streams][75*3072]
 *   1 = CIR float [n_streams][2048]                     RadioControllerInterface::onNewImpulseResponse (phasereference.cpp:93)
 *   2 = constellation cf32 [n_streams][75][16]          onConstellationPoints: r1 of every 96th carrier (ofdm-decoder.cpp:216-218)
 *   3 = null symbol cf32 [n_streams][2656], NCO applied  onNewNullSymbol (ofdm-processor.cpp:462-469)
 *   4 = TII spectra cf32 [n_streams][2][2048] (only while dabb_options.decode_tii is set): fft::Forward of the frame's phase
 *       reference symbol and of the last T_u samples of the null symbol that follows it, what TIIDecoder::run computes first
 *       (tii-decoder.cpp:197-214); the pattern analysis itself is host work (welle.io_b200/host/tii.h) */
int dabb_read_tap(dabb_ctx* ctx, int32_t what, void* host_out, size_t bytes);

/* ---------------------------------------------------------------------------------------------------------
 * Stage-level entry points (stateless).  Pointers are DEVICE pointers unless named host_*.
 * ------------------------------------------------------------------------------------------------------- */

/* OfdmDecoder::processPRS + decodeDataSymbol for n frames (ofdm-decoder.cpp:144-230).
 * Frame f: PRS useful part at iq + f*stride + prs_start[f] (2048 samples) followed by 75 symbols of 2552 samples.
 * soft_out: int8 [n][75][3072]; r1_out (may be NULL): float2 [n][75][1536] pre-quantisation products;
 * freqcorr_out (may be NULL): float2 [n] CP correlation sums (ofdm-processor.cpp:436-442).
 * nco (may be NULL): int32 [n][2] = {localPhase at the first PRS sample, phase increment Hz} (ofdm-processor.cpp:211-214) */
int dabb_ofdm_demod(dabb_ctx* ctx, const float* iq, int64_t stride_samples, const int64_t* prs_start, int32_t n_frames,
                    const int32_t* nco, int8_t* soft_out, float* r1_out, float* freqcorr_out);

/* PhaseReference::findIndex, ThresholdBeforePeak (phasereference.cpp:73-97,212-253) for n windows of 2048 samples
 * at iq + f*stride + win_start[f]; index_out int32[n]; cir_out (may be NULL) float [n][2048] */
int dabb_find_index(dabb_ctx* ctx, const float* iq, int64_t stride_samples, const int64_t* win_start, int32_t n,
                    int32_t* index_out, float* cir_out);
/* the same with any of the three placements (phasereference.cpp:93-256); negative index_out = no synchronisation, with the
 * reference's own negative value for StrongestPeak */
int dabb_find_index_ex(dabb_ctx* ctx, const float* iq, int64_t stride_samples, const int64_t* win_start, int32_t n,
                       int32_t placement, int32_t* index_out, float* cir_out);
/* OFDMProcessor::processPRS (ofdm-processor.cpp:537-644): coarse frequency estimate, in carriers, of n aligned phase reference
 * symbols (2048 samples at iq + f*stride + prs_start[f]); offset_out int32[n]; 100 = no estimate */
int dabb_coarse_estimate(dabb_ctx* ctx, const float* iq, int64_t stride_samples, const int64_t* prs_start, int32_t n,
                         int32_t freqsync_method, int32_t* offset_out);

/* Viterbi::deconvolve (viterbi.cpp:227-245) for n codewords: soft int8 [n][(nbits+6)*4] (already de-punctured,
 * 0 = punctured), bits_out uint8 [n][nbits] (one bit per byte) */
int dabb_viterbi(dabb_ctx* ctx, const int8_t* soft, int32_t n_codewords, int32_t nbits, uint8_t* bits_out);

/* FicHandler::processFicBlock x3 (fic-handler.cpp:111-230): soft int8 [n][9216] (symbols 1..3),
 * fib_out uint8 [n][12][32], crc_mask_out int32 [n] */
int dabb_fic_decode(dabb_ctx* ctx, const int8_t* soft, int32_t n_frames, uint8_t* fib_out, int32_t* crc_mask_out);

/* EEPProtection/UEPProtection::deconvolve + EnergyDispersal + byte pack (eep-protection.cpp:115-152,
 * energy_dispersal.h:35-54, decoder_adapter.cpp:57-67): soft int8 [n][length_cu*64] (time-de-interleaved),
 * bytes_out uint8 [n][3*bitrate] */
int dabb_msc_decode(dabb_ctx* ctx, const dabb_subchannel* sc, const int8_t* soft, int32_t n_cifs, uint8_t* bytes_out);

/* RSDecoder::DecodeSuperframe + SuperframeFilter::CheckSync + AU CRCs (dabplus_decoder.cpp:326-359,171-215,122-131)
 * on n superframes in place: sf uint8 [n][sf_len]; info int32 [n][4] = {corr, uncorr, sync_ok, au_crc_mask|num_aus<<8} */
int dabb_rs_superframes(dabb_ctx* ctx, uint8_t* sf, int32_t n, int32_t sf_len, int32_t* info);

/* device memory helpers so that non-CUDA hosts (ctypes) can stage buffers */
int dabb_dev_alloc(dabb_ctx* ctx, size_t bytes, void** out);
int dabb_dev_free(dabb_ctx* ctx, void* p);
int dabb_memcpy_h2d(dabb_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int dabb_memcpy_d2h(dabb_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
