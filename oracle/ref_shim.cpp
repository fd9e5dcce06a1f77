/*
 * ref_shim.cpp — TEST INFRASTRUCTURE ONLY (never linked into the shipped library).
 *
 * extern "C" entry points that drive the UNMODIFIED reference backend classes (compiled from
 * /root/reference by oracle/Makefile into oracle/_ref/libwelle_ref.so) so that the Python parity tests
 * and the C restatement in dab_oracle.c can be pinned against the real reference.
 * All code in this file is ours; it only #includes the reference's public headers.
 *
 * Two groups:
 *   ref_*      stage-level calls (tables, FFT, findIndex, demap, FIC, Viterbi, EEP/UEP, DabAudio chain, RS)
 *   ref_e2e_*  the reference RadioReceiver run end-to-end on an in-memory cf32 stream through a
 *              flow-controlled CVirtualInput (SURVEY.md §8c option B: an un-paced input makes
 *              OfdmDecoder::pushAllSymbols overwrite unconsumed frames, ofdm-decoder.cpp:132-139).
 */
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <sys/stat.h>
#include <algorithm>
#include <complex>
#include <condition_variable>
#include <deque>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>

/* SURVEY 8c option (iii): OFDMProcessor::processPRS (the coarse frequency estimators) is a private member; this
 * harness translation unit alone sees the reference's classes with private members accessible.  The reference objects
 * themselves are compiled unmodified. */
#define private public
#include "radio-receiver.h"
#undef private
#include "virtual_input.h"
#include "freq-interleaver.h"
#include "phasereference.h"
#include "protTables.h"
#include "viterbi.h"
#include "eep-protection.h"
#include "uep-protection.h"
#include "energy_dispersal.h"
#include "dab-audio.h"
#include "dabplus_decoder.h"
#include "tools.h"
#include "fft.h"
#include "MathHelper.h"

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void segv_bt(int sig) { void* a[64]; int n = backtrace(a, 64); backtrace_symbols_fd(a, n, 2); _exit(139); }
extern "C" void ref_debug_segv(void) { signal(SIGSEGV, segv_bt); signal(SIGABRT, segv_bt); }
extern "C" {
#include "fec.h"
}

namespace {

struct NullRadioController : RadioControllerInterface {
    void onSNR(float) override {}
    void onFrequencyCorrectorChange(int, int) override {}
    void onSyncChange(char) override {}
    void onSignalPresence(bool) override {}
    void onServiceDetected(uint32_t) override {}
    void onNewEnsemble(uint16_t) override {}
    void onSetEnsembleLabel(DabLabel&) override {}
    void onDateTimeUpdate(const dab_date_time_t&) override {}
    void onFIBDecodeSuccess(bool, const uint8_t*) override {}
    void onNewImpulseResponse(std::vector<float>&&) override {}
    void onConstellationPoints(std::vector<DSPCOMPLEX>&&) override {}
    void onNewNullSymbol(std::vector<DSPCOMPLEX>&&) override {}
    void onTIIMeasurement(tii_measurement_t&&) override {}
    void onMessage(message_level_t, const std::string&, const std::string&) override {}
};

struct NullProgramme : ProgrammeHandlerInterface {
    std::mutex m;
    std::vector<int> rs_events;   // pairs: uncorr, corrected
    std::atomic<int> frame_cb{0};
    void onFrameErrors(int) override { frame_cb++; }
    void onNewAudio(std::vector<int16_t>&&, int, const std::string&) override {}
    void onRsErrors(bool u, int c) override {
        std::lock_guard<std::mutex> l(m);
        rs_events.push_back(u ? 1 : 0);
        rs_events.push_back(c);
    }
    void onAacErrors(int) override {}
    void onNewDynamicLabel(const std::string&) override {}
    void onMOT(const mot_file_t&) override {}
    void onPADLengthError(size_t, size_t) override {}
};

long file_size(const char* p)
{
    struct stat st;
    if (stat(p, &st) != 0) return -1;
    return (long)st.st_size;
}

} // namespace

extern "C" {

/* ---------------- tables ---------------- */
void ref_perm_table(int mode, int16_t* out)
{
    DABParams p(mode);
    FrequencyInterleaver fi(p);
    for (int i = 0; i < p.K; i++) out[i] = fi.mapIn(i);
}

void ref_prs_table(int mode, float* out /* 2*T_u */)
{
    DABParams p(mode);
    PhaseReference pr(p, FFTPlacementMethod::ThresholdBeforePeak);
    for (int i = 0; i < p.T_u; i++) {
        DSPCOMPLEX c = pr[i];
        out[2 * i] = c.real();
        out[2 * i + 1] = c.imag();
    }
}

void ref_pcodes(int8_t* out /* 24*32 */)
{
    for (int r = 0; r < 24; r++) memcpy(out + 32 * r, getPCodes(r), 32);
}

void ref_dab_params(int mode, int32_t* out /* L,K,T_null,T_F,T_s,T_u,guard,carrierDiff */)
{
    DABParams p(mode);
    out[0] = p.L; out[1] = p.K; out[2] = p.T_null; out[3] = p.T_F;
    out[4] = p.T_s; out[5] = p.T_u; out[6] = p.guardLength; out[7] = p.carrierDiff;
}

/* ---------------- FFT (fft::Forward / fft::Backward, KISS build) ---------------- */
void ref_fft_forward(int n, float* inout)
{
    fft::Forward f(n);
    memcpy(f.getVector(), inout, sizeof(DSPCOMPLEX) * n);
    f.do_FFT();
    memcpy(inout, f.getVector(), sizeof(DSPCOMPLEX) * n);
}

void ref_fft_backward(int n, float* inout)
{
    fft::Backward f(n);
    memcpy(f.getVector(), inout, sizeof(DSPCOMPLEX) * n);
    f.do_IFFT();
    memcpy(inout, f.getVector(), sizeof(DSPCOMPLEX) * n);
}

/* ---------------- PhaseReference::findIndex ---------------- */
int ref_find_index(int mode, int method /*0 strongest,1 binning,2 threshold*/, const float* v, float* cir_out)
{
    DABParams p(mode);
    FFTPlacementMethod m = method == 0 ? FFTPlacementMethod::StrongestPeak
                         : method == 1 ? FFTPlacementMethod::EarliestPeakWithBinning
                                       : FFTPlacementMethod::ThresholdBeforePeak;
    PhaseReference pr(p, m);
    std::vector<DSPCOMPLEX> buf(p.T_u);
    memcpy(buf.data(), v, sizeof(DSPCOMPLEX) * p.T_u);
    std::vector<float> cir;
    int idx = pr.findIndex(buf.data(), cir);
    if (cir_out && (int)cir.size() == p.T_u) memcpy(cir_out, cir.data(), sizeof(float) * p.T_u);
    return idx;
}

/* ---------------- OFDM demap of one frame ----------------
 * The 15-line loop of OfdmDecoder::decodeDataSymbol (ofdm-decoder.cpp:198-219) sits in a private method
 * of a class that owns a thread, so this is the one place the shim restates reference statements: it
 * calls the reference's own fft::Forward, FrequencyInterleaver::mapIn and l1_norm and performs the same
 * float expressions in the same order.  prs = the 2048 useful PRS samples, syms = 75 symbols of T_s
 * samples (guard first).  soft: 75*2K int8.  r1s (optional): 75*K complex, pre-quantisation products.
 */
void ref_ofdm_demod_frame(int mode, const float* prs, const float* syms, int8_t* soft, float* r1s, float* spectra)
{
    DABParams p(mode);
    FrequencyInterleaver il(p);
    fft::Forward f(p.T_u);
    DSPCOMPLEX* buf = f.getVector();
    std::vector<DSPCOMPLEX> phaseReference(p.T_u);
    const int T_g = p.T_s - p.T_u;

    memcpy(buf, prs, sizeof(DSPCOMPLEX) * p.T_u);
    f.do_FFT();
    memcpy(phaseReference.data(), buf, sizeof(DSPCOMPLEX) * p.T_u);
    if (spectra) memcpy(spectra, buf, sizeof(DSPCOMPLEX) * p.T_u);

    for (int l = 1; l < p.L; l++) {
        const DSPCOMPLEX* s = reinterpret_cast<const DSPCOMPLEX*>(syms) + (size_t)(l - 1) * p.T_s;
        memcpy(buf, s + T_g, sizeof(DSPCOMPLEX) * p.T_u);
        f.do_FFT();
        if (spectra) memcpy(spectra + (size_t)2 * l * p.T_u, buf, sizeof(DSPCOMPLEX) * p.T_u);
        softbit_t* ibits = soft + (size_t)(l - 1) * 2 * p.K;
        for (int16_t i = 0; i < p.K; i++) {
            int16_t index = il.mapIn(i);
            if (index < 0) index += p.T_u;
            const DSPCOMPLEX r1 = buf[index] * conj(phaseReference[index]);
            phaseReference[index] = buf[index];
            const DSPFLOAT ab1 = 127.0f / l1_norm(r1);
            ibits[i] = -real(r1) * ab1;
            ibits[p.K + i] = -imag(r1) * ab1;
            if (r1s) {
                r1s[((size_t)(l - 1) * p.K + i) * 2] = real(r1);
                r1s[((size_t)(l - 1) * p.K + i) * 2 + 1] = imag(r1);
            }
        }
    }
}

/* ---------------- Viterbi::deconvolve ---------------- */
void ref_viterbi(int nbits, const int8_t* in /* (nbits+6)*4 */, uint8_t* out /* nbits */)
{
    Viterbi v(nbits);
    std::vector<softbit_t> tmp(in, in + (size_t)(nbits + 6) * 4);
    v.deconvolve(tmp.data(), out);
}

/* ---------------- FicHandler: 3 FIC symbols -> 12 FIBs ---------------- */
namespace {
struct FibCapture : NullRadioController {
    uint8_t* bits; uint8_t* ok; int n = 0;
    void onFIBDecodeSuccess(bool crc, const uint8_t* fib) override {
        memcpy(bits + 256 * n, fib, 256);
        ok[n] = crc ? 1 : 0;
        n++;
    }
};
}
int ref_fic_decode(const int8_t* soft /* 3*3072 */, uint8_t* fib_bits /* 12*256 */, uint8_t* crc_ok /* 12 */)
{
    FibCapture cap; cap.bits = fib_bits; cap.ok = crc_ok;
    FicHandler fic(cap);
    for (int blk = 1; blk <= 3; blk++) fic.processFicBlock(soft + (blk - 1) * 3072, blk);
    return cap.n;
}

/* ---------------- EEP / UEP deconvolve (+ optional energy dispersal) ---------------- */
void ref_eep_deconvolve(int bitrate, int profile_a, int level, const int8_t* in, uint8_t* outbits, int dedisperse)
{
    EEPProtection e(bitrate, profile_a != 0, level);
    std::vector<uint8_t> out(24 * bitrate);
    e.deconvolve(in, 0, out.data());
    if (dedisperse) { EnergyDispersal d; d.dedisperse(out); }
    memcpy(outbits, out.data(), out.size());
}

void ref_uep_deconvolve(int bitrate, int level, const int8_t* in, uint8_t* outbits, int dedisperse)
{
    UEPProtection e(bitrate, level);
    std::vector<uint8_t> out(24 * bitrate);
    e.deconvolve(in, 0, out.data());
    if (dedisperse) { EnergyDispersal d; d.dedisperse(out); }
    memcpy(outbits, out.data(), out.size());
}

int ref_subchannel_bitrate(int length, int short_form, int uep_index, int eep_profile_a, int eep_level)
{
    Subchannel s;
    s.length = length;
    s.protectionSettings.shortForm = short_form != 0;
    s.protectionSettings.uepTableIndex = uep_index;
    s.protectionSettings.eepProfile = eep_profile_a ? EEPProtectionProfile::EEP_A : EEPProtectionProfile::EEP_B;
    s.protectionSettings.eepLevel = (EEPProtectionLevel)eep_level;
    try { return s.bitrate(); } catch (...) { return -1; }
}

/* ---------------- DabAudio chain: CIF slices -> logical-frame bytes (dump file) ----------------
 * Drives the real DabAudio (time de-interleaver, EEP/UEP, Viterbi, energy dispersal, DecoderAdapter
 * byte pack + dump, SuperframeFilter RS/Fire).  DabAudio::run only consumes a fragment when strictly more
 * than one fragment is buffered (dab-audio.cpp:127), so one extra zero fragment is pushed at the end.
 * Returns number of bytes in the dump file; rs_events gets (uncorr, corrected) pairs.
 */
long ref_dabaudio_chain(const int8_t* cifs, int ncif, int fragment, int bitrate,
                        int short_form, int uep_level, int eep_profile_a, int eep_level, int dabplus,
                        const char* dump_path, int* rs_events, int rs_cap, int* n_rs)
{
    NullProgramme ph;
    ProtectionSettings ps;
    ps.shortForm = short_form != 0;
    ps.uepLevel = uep_level;
    ps.eepProfile = eep_profile_a ? EEPProtectionProfile::EEP_A : EEPProtectionProfile::EEP_B;
    ps.eepLevel = (EEPProtectionLevel)eep_level;
    const long expect = (long)(ncif > 16 ? ncif - 16 : 0) * 3 * bitrate;
    {
        DabAudio da(dabplus ? AudioServiceComponentType::DABPlus : AudioServiceComponentType::DAB,
                    fragment, bitrate, ps, ph, dump_path);
        for (int c = 0; c < ncif; c++) da.process(cifs + (size_t)c * fragment, fragment);
        std::vector<softbit_t> zero(fragment, 0);
        da.process(zero.data(), fragment);
        /* wait for the worker: one onFrameErrors callback per produced logical frame */
        const int want = ncif > 16 ? ncif - 16 : 0;
        for (int spin = 0; spin < 20000 && ph.frame_cb.load() < want; spin++)
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }   /* dtor joins the thread and closes the dump file */
    if (n_rs) {
        int n = (int)ph.rs_events.size() / 2;
        if (n > rs_cap) n = rs_cap;
        for (int i = 0; i < 2 * n; i++) rs_events[i] = ph.rs_events[i];
        *n_rs = n;
    }
    (void)expect;
    return file_size(dump_path);
}

/* ---------------- RS(120,110) superframe + CRCs ---------------- */
void ref_rs_decode_superframe(uint8_t* sf, int sf_len, int* corr, int* uncorr)
{
    RSDecoder d;
    int c = 0; bool u = false;
    d.DecodeSuperframe(sf, sf_len, c, u);
    *corr = c; *uncorr = u ? 1 : 0;
}

void ref_rs_encode(const uint8_t* data110, uint8_t* parity10)
{
    void* rs = init_rs_char(8, 0x11D, 0, 1, 10, 135);
    encode_rs_char(rs, const_cast<uint8_t*>(data110), parity10);
    free_rs_char(rs);
}

int ref_rs_decode_codeword(uint8_t* cw120, int* corr_pos10)
{
    void* rs = init_rs_char(8, 0x11D, 0, 1, 10, 135);
    int r = decode_rs_char(rs, cw120, corr_pos10, 0);
    free_rs_char(rs);
    return r;
}

unsigned ref_crc_fire(const uint8_t* d, int n) { return CalcCRC::CalcCRC_FIRE_CODE.Calc(d, n); }
unsigned ref_crc_ccitt(const uint8_t* d, int n) { return CalcCRC::CalcCRC_CRC16_CCITT.Calc(d, n); }
int ref_check_crc_bits(const uint8_t* bits, int n) { return check_CRC_bits(bits, n) ? 1 : 0; }

/* SuperframeFilter driven directly: logical frames in, per-attempt (uncorr,corr) and AU-CRC failures out */
namespace {
struct SfObserver : SubchannelSinkObserver {
    std::vector<int> fec; int au_err = 0; int pad_calls = 0;
    void FECInfo(int c, bool u) override { fec.push_back(u ? 1 : 0); fec.push_back(c); }
    void AudioError(const std::string&) override { au_err++; }
    void ProcessPAD(const uint8_t*, size_t, bool, const uint8_t*) override { pad_calls++; }
};
}
int ref_superframe_filter(const uint8_t* frames, int nframes, int frame_len, int* fec_out, int fec_cap,
                          int* au_err, int* good_aus)
{
    SfObserver ob;
    SuperframeFilter sf(&ob, false, false);
    for (int i = 0; i < nframes; i++) sf.Feed(frames + (size_t)i * frame_len, frame_len);
    int n = (int)ob.fec.size() / 2;
    if (n > fec_cap) n = fec_cap;
    for (int i = 0; i < 2 * n; i++) fec_out[i] = ob.fec[i];
    *au_err = ob.au_err;
    *good_aus = ob.pad_calls;   /* CheckForPAD runs once per AU whose CRC matched (dabplus_decoder.cpp:122-137) */
    return n;
}

/* ---------------- OFDMProcessor::processPRS (coarse frequency estimate) on one aligned PRS ----------------
 * freqsync in the reference's enum order: 0 GetMiddle, 1 CorrelatePRS, 2 PatternOfZeros */
namespace { struct NullInput; }
extern "C" int ref_process_prs(const float* prs, int freqsync);

/* ---------------- end-to-end: RadioReceiver on an in-memory stream ---------------- */
namespace {

struct GatedMemInput : CVirtualInput {
    const DSPCOMPLEX* d; size_t n; std::atomic<size_t> pos{0};
    std::atomic<long> frames_done{0};
    std::atomic<bool> stopped{false};
    long slack = 0;          /* frames consumed without a decoded frame coming out (acquisition, sync losses): each 250 ms stall adds one, for good */
    long T_F;
    GatedMemInput(const float* iq, size_t nsamples, long tf) : d(reinterpret_cast<const DSPCOMPLEX*>(iq)), n(nsamples), T_F(tf) {}
    CDeviceID getID() override { return CDeviceID::RAWFILE; }
    void setFrequency(int) override {}
    int getFrequency() const override { return 0; }
    /* fails once fewer samples remain than the largest single read (T_null): OFDMProcessor::getSamples then
     * throws InputFailure instead of spinning (ofdm-processor.cpp:193-201) */
    bool is_ok() override { size_t p = pos.load(); return p < n && n - p >= 2656; }
    bool restart() override { return true; }
    void stop() override { stopped = true; }
    void reset() override {}
    int32_t getSamples(DSPCOMPLEX* b, int32_t cnt) override {
        /* gate: never run more than two frames ahead of the decoder worker; a 250 ms stall opens the gate
         * by one frame - permanently - so that (re-)acquisition, which consumes samples without producing frames, cannot deadlock */
        auto t0 = std::chrono::steady_clock::now();
        while (!stopped && (long)pos.load() + cnt > (frames_done.load() + 2 + slack) * T_F) {
            std::this_thread::sleep_for(std::chrono::microseconds(50));
            if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(250)) { slack++; t0 = std::chrono::steady_clock::now(); }
        }
        size_t p = pos.load();
        size_t m = p < n ? std::min<size_t>(cnt, n - p) : 0;
        memcpy(b, d + p, m * sizeof(DSPCOMPLEX));
        if (m < (size_t)cnt) memset(b + m, 0, (cnt - m) * sizeof(DSPCOMPLEX));
        pos = p + cnt;
        return cnt;
    }
    std::vector<DSPCOMPLEX> getSpectrumSamples(int sz) override { return std::vector<DSPCOMPLEX>(sz); }
    int32_t getSamplesToRead() override { size_t p = pos.load(); return p < n ? (int32_t)std::min<size_t>(n - p, 1 << 20) : 0; }
    float setGain(int) override { return 0; }
    float getGain() const override { return 0; }
    int getGainCount() override { return 0; }
    void setAgc(bool) override {}
    std::string getDescription() override { return "gated-mem"; }
};

struct E2EController : NullRadioController {
    RadioReceiver* rx = nullptr; GatedMemInput* in = nullptr; NullProgramme* ph = nullptr;
    std::string dump; int select_at_fib = 24; bool selected = false; bool select_ok = false;
    std::vector<uint8_t> fibs;      /* 33 bytes per FIB: crc flag + 32 packed bytes */
    std::vector<int> corr;          /* (fine, coarse) events */
    std::vector<float> snr;
    std::vector<float> cirs; int keep_cir = 0;
    int sync_true = 0, sync_false = 0; std::atomic<bool> failed{false};
    int nfib = 0;
    void onFIBDecodeSuccess(bool ok, const uint8_t* fib) override {
        uint8_t rec[33]; rec[0] = ok ? 1 : 0;
        for (int i = 0; i < 32; i++) { uint8_t b = 0; for (int j = 0; j < 8; j++) b = (b << 1) | (fib[8 * i + j] & 1); rec[1 + i] = b; }
        fibs.insert(fibs.end(), rec, rec + 33);
        nfib++;
        if (!selected && select_at_fib >= 0 && nfib >= select_at_fib) {
            auto list = rx->getServiceList();
            if (!list.empty()) { selected = true; select_ok = rx->playSingleProgramme(*ph, dump, list.front()); }
        }
    }
    void onConstellationPoints(std::vector<DSPCOMPLEX>&&) override { in->frames_done++; }
    void onFrequencyCorrectorChange(int f, int c) override { corr.push_back(f); corr.push_back(c); }
    void onSNR(float s) override { snr.push_back(s); }
    void onSyncChange(char s) override { if (s) sync_true++; else sync_false++; }
    void onNewImpulseResponse(std::vector<float>&& d) override { if (keep_cir) cirs.insert(cirs.end(), d.begin(), d.end()); }
    void onInputFailure() override { failed = true; }
};

struct E2EResult {
    std::vector<uint8_t> fibs; std::vector<int> rs; std::vector<int> corr; std::vector<float> snr; std::vector<float> cirs;
    int sync_true = 0, sync_false = 0, select_ok = 0; long frames_done = 0; double seconds = 0; long msc_bytes = 0;
};
E2EResult g_last;

} // namespace

/* Runs the reference receiver over iq[0..nsamples). Returns number of FIB callbacks. */
int ref_e2e_run2(const float* iq, long nsamples, int disable_coarse, int select_at_fib, const char* msc_dump_path, int keep_cir, int fft_placement, int freqsync);
int ref_e2e_run(const float* iq, long nsamples, int disable_coarse, int select_at_fib, const char* msc_dump_path, int keep_cir)
{
    return ref_e2e_run2(iq, nsamples, disable_coarse, select_at_fib, msc_dump_path, keep_cir, 2, 2);
}
/* fft_placement / freqsync in the reference's enum order (radio-receiver-options.h:35-64) */
int ref_e2e_run2(const float* iq, long nsamples, int disable_coarse, int select_at_fib, const char* msc_dump_path, int keep_cir, int fft_placement, int freqsync)
{
    g_last = E2EResult();
    DABParams p(1);
    /* The session is deliberately leaked: tearing the reference receiver down can dead-lock (DabAudio's destructor
     * clears `running` and notifies without holding the mutex its worker checks `running` under, dab-audio.cpp:87-93 vs
     * :123-128, so the notification can be lost).  The OFDM thread has already exited by then; the remaining worker
     * threads idle on their condition variables. */
    struct Session {
        GatedMemInput in; E2EController ri; NullProgramme ph; RadioReceiver rx;
        Session(const float* iq, long n, long tf, RadioReceiverOptions rro) : in(iq, (size_t)n, tf), rx(ri, in, rro) {}
    };
    RadioReceiverOptions rro; rro.disableCoarseCorrector = disable_coarse != 0; rro.decodeTII = false;
    rro.fftPlacementMethod = fft_placement == 0 ? FFTPlacementMethod::StrongestPeak : fft_placement == 1 ? FFTPlacementMethod::EarliestPeakWithBinning : FFTPlacementMethod::ThresholdBeforePeak;
    rro.freqsyncMethod = freqsync == 0 ? FreqsyncMethod::GetMiddle : freqsync == 1 ? FreqsyncMethod::CorrelatePRS : FreqsyncMethod::PatternOfZeros;
    auto t0 = std::chrono::steady_clock::now();
    Session* S = new Session(iq, nsamples, p.T_F, rro);
    GatedMemInput& in = S->in; E2EController& ri = S->ri; NullProgramme& ph = S->ph; RadioReceiver& rx = S->rx;
    {
        ri.rx = &rx; ri.in = &in; ri.ph = &ph; ri.dump = msc_dump_path ? msc_dump_path : ""; ri.select_at_fib = select_at_fib; ri.keep_cir = keep_cir;
        rx.restart(false);
        while (!ri.failed.load() && in.is_ok()) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        /* input exhausted: OFDMProcessor::getSamples throws InputFailure once fewer samples remain than it needs */
        for (int i = 0; i < 3000 && !ri.failed.load(); i++) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        g_last.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        /* let the OfdmDecoder worker and the DabAudio thread drain */
        long last = -1; int stable = 0;
        for (int i = 0; i < 5000 && stable < 300; i++) {
            long cur = in.frames_done.load() * 1000003L + ph.frame_cb.load();
            if (cur == last) stable++; else { stable = 0; last = cur; }
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        in.stopped = true;
    }
    g_last.fibs = ri.fibs; g_last.rs = ph.rs_events; g_last.corr = ri.corr; g_last.snr = ri.snr; g_last.cirs = ri.cirs;
    g_last.sync_true = ri.sync_true; g_last.sync_false = ri.sync_false; g_last.select_ok = ri.select_ok ? 1 : 0;
    g_last.frames_done = in.frames_done.load();
    g_last.msc_bytes = (msc_dump_path && msc_dump_path[0]) ? file_size(msc_dump_path) : 0;
    return (int)(g_last.fibs.size() / 33);
}

long ref_e2e_get(int what, void* out, long cap_bytes)
{
    const void* src = nullptr; long n = 0;
    switch (what) {
        case 0: src = g_last.fibs.data(); n = (long)g_last.fibs.size(); break;
        case 1: src = g_last.rs.data(); n = (long)(g_last.rs.size() * sizeof(int)); break;
        case 2: src = g_last.corr.data(); n = (long)(g_last.corr.size() * sizeof(int)); break;
        case 3: src = g_last.snr.data(); n = (long)(g_last.snr.size() * sizeof(float)); break;
        case 4: src = g_last.cirs.data(); n = (long)(g_last.cirs.size() * sizeof(float)); break;
        default: return -1;
    }
    if (out && n > 0) memcpy(out, src, n < cap_bytes ? n : cap_bytes);
    return n;
}

void ref_e2e_stats(double* out /* sync_true, sync_false, select_ok, frames_done, seconds, msc_bytes */)
{
    out[0] = g_last.sync_true; out[1] = g_last.sync_false; out[2] = g_last.select_ok;
    out[3] = (double)g_last.frames_done; out[4] = g_last.seconds; out[5] = (double)g_last.msc_bytes;
}

} // extern "C"

extern "C" int ref_process_prs(const float* prs, int freqsync)
{
    static const float dummy[16] = {0};
    DABParams p(1);
    GatedMemInput in(dummy, 0, p.T_F);
    E2EController ri;
    RadioReceiverOptions rro;
    RadioReceiver rx(ri, in, rro);          /* never started: only the OFDMProcessor member's tables are used */
    std::vector<DSPCOMPLEX> v(p.T_u);
    memcpy(v.data(), prs, sizeof(DSPCOMPLEX) * p.T_u);
    const FreqsyncMethod m = freqsync == 0 ? FreqsyncMethod::GetMiddle : freqsync == 1 ? FreqsyncMethod::CorrelatePRS : FreqsyncMethod::PatternOfZeros;
    return rx.ofdmProcessor.processPRS(v.data(), m);
}

/* ---------------- FIBProcessor: service database after a sequence of FIBs, as text (same format as welle.io_b200/host/fig-db.h) ---- */
namespace {
struct FibRecorder : RadioControllerInterface {
    std::string log;
    void onSNR(float) override {} void onFrequencyCorrectorChange(int, int) override {} void onSyncChange(char) override {} void onSignalPresence(bool) override {}
    void onServiceDetected(uint32_t s) override { char t[64]; snprintf(t, sizeof t, "cb serviceDetected %u\n", s); log += t; }
    void onNewEnsemble(uint16_t e) override { char t[64]; snprintf(t, sizeof t, "cb newEnsemble %u\n", e); log += t; }
    void onSetEnsembleLabel(DabLabel&) override { char t[64]; snprintf(t, sizeof t, "cb ensembleLabel %u\n", eid ? *eid : 0u); log += t; }
    void onDateTimeUpdate(const dab_date_time_t&) override { log += "cb dateTime 0\n"; }
    void onFIBDecodeSuccess(bool, const uint8_t*) override {} void onNewImpulseResponse(std::vector<float>&&) override {}
    void onConstellationPoints(std::vector<DSPCOMPLEX>&&) override {} void onNewNullSymbol(std::vector<DSPCOMPLEX>&&) override {}
    void onTIIMeasurement(tii_measurement_t&&) override {} void onMessage(message_level_t, const std::string&, const std::string&) override {}
    void onRestartService() override { log += "cb restartService 0\n"; }
    const uint16_t* eid = nullptr;
};
std::string hexs(const std::string& s) { std::string o; char t[4]; for (unsigned char c : s) { snprintf(t, sizeof t, "%02x", c); o += t; } return o; }
std::string short_hex(const DabLabel& l) { for (unsigned char c : l.fig1_label) if (c >= 0x7B || c < 0x20 || c == 0x24 || c == 0x5C || c == 0x5E || c == 0x60) return "-"; return hexs(l.fig1_shortlabel_utf8()); }
std::string ext_dump(const char* head, const DabLabel& l)
{
    if (l.segments.empty() && l.segment_count == 0) return std::string();
    std::string o = head; char t[64];
    snprintf(t, sizeof t, " %d %d %d %d", l.toggle_flag ? 1 : 0, (int)l.segment_count, l.fig2_rfu ? 1 : 0, (int)l.extended_label_charset); o += t;
    for (const auto& kv : l.segments) { snprintf(t, sizeof t, " %d:", kv.first); o += t; o += hexs(std::string(kv.second.begin(), kv.second.end())); }
    o += " utf8="; o += hexs(l.fig2_label());
    return o + "\n";
}
}

extern "C" int ref_fib_dump(const uint8_t* fibs, int n, char* out, int cap)
{
    FibRecorder rec;
    FIBProcessor fp(rec);
    rec.eid = &fp.ensembleId;
    for (int i = 0; i < n; i++) {
        uint8_t bits[256 + 2048] = {0};           /* zero bits behind the FIB: fields of malformed FIGs that run past it read zeros */
        for (int k = 0; k < 256; k++) bits[k] = (fibs[32 * i + (k >> 3)] >> (7 - (k & 7))) & 1;
        try { fp.processFIB(bits, 0); } catch (const std::exception& e) { rec.log += std::string("exception ") + e.what() + "\n"; }
    }
    std::string o = rec.log; char t[256];
    const DabLabel el = fp.getEnsembleLabel();
    snprintf(t, sizeof t, "E %u %u %d %u [%s] [%s]\n", fp.getEnsembleId(), fp.getEnsembleEcc(), (int)el.charset, el.fig1_flag, hexs(el.fig1_label).c_str(), short_hex(el).c_str()); o += t;
    o += ext_dump("XE 0 0", el);
    const auto services = fp.getServiceList();
    for (const auto& s : services) {
        snprintf(t, sizeof t, "S %u %d %d %d %u [%s] [%s]\n", s.serviceId, s.language, s.programType, (int)s.serviceLabel.charset, s.serviceLabel.fig1_flag, hexs(s.serviceLabel.fig1_label).c_str(),
                 short_hex(s.serviceLabel).c_str()); o += t;
        snprintf(t, sizeof t, "XS %u 0", s.serviceId); o += ext_dump(t, s.serviceLabel);
    }
    for (const auto& s : services) for (const auto& c : fp.getComponents(s)) {
        snprintf(t, sizeof t, "C %u %d %d %d %d %d %u %d %d %d %d %d %u [%s]\n", c.SId, c.componentNr, c.TMid, c.ASCTy, c.DSCTy, c.subchannelId, c.SCId, c.PS_flag, c.CAflag, c.DGflag,
                 c.packetAddress, (int)c.componentLabel.charset, c.componentLabel.fig1_flag, hexs(c.componentLabel.fig1_label).c_str()); o += t;
        snprintf(t, sizeof t, "XC %u %d", c.SId, c.componentNr); o += ext_dump(t, c.componentLabel);
    }
    for (const auto& u : fp.subChannels) {
        if (u.subChId == -1) continue;
        snprintf(t, sizeof t, "U %d %d %d %d %d %d %d %d %d %d %d\n", u.subChId, u.startAddr, u.length, u.programmeNotData ? 1 : 0, u.protectionSettings.shortForm ? 1 : 0,
                 u.protectionSettings.uepTableIndex, u.protectionSettings.uepLevel, (int)u.protectionSettings.eepProfile, (int)u.protectionSettings.eepLevel, u.language, u.fecScheme); o += t;
    }
    const auto& dt = fp.dateTime;
    snprintf(t, sizeof t, "T %d %d %d %d %d %d %d %d\n", dt.year, dt.month, dt.day, dt.hour, dt.minutes, dt.seconds, dt.hourOffset, dt.minuteOffset); o += t;
    if ((int)o.size() + 1 > cap) return -(int)o.size() - 1;
    memcpy(out, o.c_str(), o.size() + 1);
    return (int)o.size();
}


/* ---- TIIDecoder (backend/tii-decoder.cpp): the unmodified class fed frame by frame; the harness waits until its thread is idle again
 * before the next frame so that every frame is analysed (the receiver itself drops frames while the decoder is busy).
 * nulls: n x 2656 complex floats (the null symbol, oscillator applied), prss: n x 2048; out: 4 floats per measurement */
extern "C" int ref_tii_run(const float* nulls, const float* prss, int n, float* out, int cap)
{
    struct Ri : NullRadioController {
        std::vector<tii_measurement_t> got; std::mutex m;
        void onTIIMeasurement(tii_measurement_t&& t) override { std::lock_guard<std::mutex> l(m); got.push_back(t); }
    } ri;
    DABParams params(1);
    int k = 0;
    {
        TIIDecoder dec(params, ri);
        for (int f = 0; f < n; f++) {
            std::vector<complexf> nul(reinterpret_cast<const complexf*>(nulls) + (size_t)f * 2656, reinterpret_cast<const complexf*>(nulls) + (size_t)(f + 1) * 2656);
            std::vector<complexf> prs(reinterpret_cast<const complexf*>(prss) + (size_t)f * 2048, reinterpret_cast<const complexf*>(prss) + (size_t)(f + 1) * 2048);
            dec.pushSymbols(nul, prs);
            for (;;) {
                { std::unique_lock<std::mutex> l(dec.m_state_mutex); if (dec.m_state == TIIDecoder::State::Idle) break; }
                std::this_thread::sleep_for(std::chrono::microseconds(200));
            }
        }
    }
    for (const auto& t : ri.got) { if (k < cap) { out[4 * k] = (float)t.comb; out[4 * k + 1] = (float)t.pattern; out[4 * k + 2] = (float)t.delay_samples; out[4 * k + 3] = t.error; } k++; }
    return k;
}
