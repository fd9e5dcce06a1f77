/*
 * radio-receiver.cpp — host glue: the reference's RadioReceiver surface (backend/radio-receiver.h:52-116) implemented over
 * the C ABI of libdab_b200.so (include/dab_b200.h).  One worker thread replaces the reference's OFDMProcessor, OfdmDecoder
 * and DabAudio threads: it pulls samples from the InputInterface exactly like OFDMProcessor::getSamples
 * (backend/ofdm-processor.cpp:186-224), hands one frame's worth to dabb_process() (n_streams = 1) and turns the POD
 * results back into the reference's callbacks.  No signal processing happens here.
 *
 * Also contains the small FIG 0/0, 0/1, 0/2, 1/0, 1/1 reader the glue needs to map a Service to its sub-channel
 * (ETSI EN 300 401 §6.2-6.4, §8.1); the reference's full FIBProcessor (backend/fib-processor.cpp) is out of scope
 *  — in particular its two-sightings / time-decay acceptance rule is not reproduced.
 */
#include "dab_api.h"
#include "../../include/dab_b200.h"

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <thread>

/* ---------------------------------------------------------------------------------------------------------------- */
DABParams::DABParams(int mode) { setMode(mode); }
void DABParams::setMode(int mode)
{
    /* backend/dab-constants.cpp:347-402 */
    switch (mode) {
        case 1: dabMode = 1; L = 76; K = 1536; T_F = 196608; T_null = 2656; T_s = 2552; T_u = 2048; guardLength = 504; carrierDiff = 1000; break;
        case 2: dabMode = 2; L = 76; K = 384; T_null = 664; T_F = 49152; T_s = 638; T_u = 512; guardLength = 126; carrierDiff = 4000; break;
        case 3: dabMode = 3; L = 153; K = 192; T_F = 49152; T_null = 345; T_s = 319; T_u = 256; guardLength = 63; carrierDiff = 2000; break;
        case 4: dabMode = 4; L = 76; K = 768; T_F = 98304; T_null = 1328; T_s = 1276; T_u = 1024; guardLength = 252; carrierDiff = 2000; break;
        default: throw std::out_of_range("Unknown mode " + std::to_string(mode));
    }
}

/* ETSI EN 300 401 Table 8 (short-form sub-channel sizes): size in CU, protection level, bit rate */
static const int16_t kUepSize[64] = {16,21,24,29,35, 24,29,35,42,52, 29,35,42,52, 32,42,48,58,70, 40,52,58,70,84, 48,58,70,84,104, 58,70,84,104,
                                     64,84,96,116,140, 80,104,116,140,168, 96,116,140,168,208, 116,140,168,208,232, 128,168,192,232,280, 160,208,280, 192,280,416};
static const int8_t kUepLevel[64] = {5,4,3,2,1, 5,4,3,2,1, 5,4,3,2, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2,1, 5,4,2, 5,3,1};
static const int16_t kUepRate[64] = {32,32,32,32,32, 48,48,48,48,48, 56,56,56,56, 64,64,64,64,64, 80,80,80,80,80, 96,96,96,96,96, 112,112,112,112,
                                     128,128,128,128,128, 160,160,160,160,160, 192,192,192,192,192, 224,224,224,224,224, 256,256,256,256,256, 320,320,320, 384,384,384};

int Subchannel::bitrate() const
{
    const auto& ps = protectionSettings;
    if (ps.shortForm) return kUepRate[ps.uepTableIndex & 63];
    static const int da[5] = {0, 12, 8, 6, 4}, db[5] = {0, 27, 21, 18, 15};
    const int lv = (int)ps.eepLevel;
    if (lv < 1 || lv > 4) throw std::runtime_error("Unsupported protection");
    return ps.eepProfile == EEPProtectionProfile::EEP_A ? length / da[lv] * 8 : length / db[lv] * 32;   /* dab-constants.cpp:404-440 */
}
int Subchannel::numCU() const
{
    const auto& ps = protectionSettings;
    if (ps.shortForm) return kUepSize[ps.uepTableIndex & 63];
    const int b = bitrate();
    switch ((int)ps.eepLevel) {       /* dab-constants.cpp:442-477 */
        case 1: return ps.eepProfile == EEPProtectionProfile::EEP_A ? (b * 12) >> 3 : (b * 27) >> 5;
        case 2: return ps.eepProfile == EEPProtectionProfile::EEP_A ? b : (b * 21) >> 5;
        case 3: return ps.eepProfile == EEPProtectionProfile::EEP_A ? (b * 6) >> 3 : (b * 18) >> 5;
        case 4: return ps.eepProfile == EEPProtectionProfile::EEP_A ? b >> 1 : (b * 15) >> 5;
    }
    return -1;
}

const char* fftPlacementMethodToString(FFTPlacementMethod m)
{
    switch (m) { case FFTPlacementMethod::StrongestPeak: return "StrongestPeak"; case FFTPlacementMethod::EarliestPeakWithBinning: return "EarliestPeakWithBinning"; default: return "ThresholdBeforePeak"; }
}
const char* freqSyncMethodToString(FreqsyncMethod m)
{
    switch (m) { case FreqsyncMethod::GetMiddle: return "GetMiddle"; case FreqsyncMethod::CorrelatePRS: return "CorrelatePRS"; default: return "PatternOfZeros"; }
}

/* ---------------------------------------------------------------------------------------------------------------- */
namespace {

struct FigDb {
    std::mutex m;
    uint16_t eid = 0; uint8_t ecc = 0; DabLabel ensLabel; bool haveEns = false;
    std::map<uint32_t, Service> services;
    std::map<uint32_t, std::vector<ServiceComponent>> comps;
    std::map<int, Subchannel> subch;
    /* FIBProcessor's acceptance rule (fib-processor.cpp:285-327): a saturating sighting counter per SId (max 4), all counters
     * decremented once per second of wall clock, a service is listed from its second sighting on.  When a counter reaches zero
     * the reference calls dropService() with the counter value instead of the SId (:302), i.e. it drops service 0: mirrored. */
    std::map<uint32_t, int8_t> repeatCount;
    std::chrono::steady_clock::time_point lastDecrement = std::chrono::steady_clock::now();

    void clear() { std::lock_guard<std::mutex> l(m); eid = 0; haveEns = false; ensLabel = DabLabel(); services.clear(); comps.clear(); subch.clear(); repeatCount.clear(); }

    bool sighting(uint32_t sid)      /* true when the service becomes listed now */
    {
        const auto now = std::chrono::steady_clock::now();
        if (lastDecrement + std::chrono::seconds(1) < now) {
            for (auto it = repeatCount.begin(); it != repeatCount.end();) {
                if (it->second > 0) { it->second--; ++it; }
                else if (it->second == 0) { services.erase((uint32_t)it->second); comps.erase((uint32_t)it->second); it = repeatCount.erase(it); }
                else ++it;
            }
            lastDecrement = now;
        }
        int8_t& c = repeatCount[sid];
        if (c < 4) c++;
        if (!services.count(sid) && c >= 2) { services.emplace(sid, Service(sid)); return true; }
        return false;
    }

    /* returns the list of newly detected service ids */
    std::vector<uint32_t> parseFib(const uint8_t* b /* 30 data bytes */, bool& newEnsemble, bool& newEnsLabel)
    {
        std::vector<uint32_t> fresh;
        std::lock_guard<std::mutex> l(m);
        int p = 0;
        while (p < 30) {
            const int type = b[p] >> 5, len = b[p] & 0x1F;
            if (b[p] == 0xFF || len == 0 || p + 1 + len > 30) break;
            const uint8_t* d = b + p + 1;
            if (type == 0 && len >= 1) {
                const int pd = (d[0] >> 5) & 1, ext = d[0] & 0x1F;
                const uint8_t* q = d + 1; int n = len - 1;
                if (ext == 0 && n >= 4) { uint16_t e = q[0] << 8 | q[1]; if (!haveEns || e != eid) { eid = e; haveEns = true; newEnsemble = true; } }
                else if (ext == 1) {
                    int i = 0;
                    while (i + 3 <= n) {
                        Subchannel s; s.subChId = q[i] >> 2; s.startAddr = ((q[i] & 3) << 8) | q[i + 1];
                        if (q[i + 2] & 0x80) {      /* long form */
                            if (i + 4 > n) break;
                            const int opt = (q[i + 2] >> 4) & 7;
                            s.protectionSettings.shortForm = false;
                            s.protectionSettings.eepProfile = opt == 0 ? EEPProtectionProfile::EEP_A : EEPProtectionProfile::EEP_B;
                            s.protectionSettings.eepLevel = (EEPProtectionLevel)(((q[i + 2] >> 2) & 3) + 1);
                            s.length = ((q[i + 2] & 3) << 8) | q[i + 3];
                            i += 4;
                        } else {
                            const int idx = q[i + 2] & 0x3F;
                            s.protectionSettings.shortForm = true; s.protectionSettings.uepTableIndex = idx;
                            s.protectionSettings.uepLevel = kUepLevel[idx]; s.length = kUepSize[idx];
                            i += 3;
                        }
                        subch[s.subChId] = s;
                    }
                } else if (ext == 2) {
                    int i = 0;
                    while (i < n) {
                        uint32_t sid;
                        if (pd) { if (i + 5 > n) break; sid = (uint32_t)q[i] << 24 | q[i + 1] << 16 | q[i + 2] << 8 | q[i + 3]; i += 4; }
                        else { if (i + 3 > n) break; sid = q[i] << 8 | q[i + 1]; i += 2; }
                        const int nc = q[i] & 0x0F; i++;
                        if (i + 2 * nc > n) break;
                        std::vector<ServiceComponent> v;
                        for (int c = 0; c < nc; c++, i += 2) {
                            ServiceComponent sc; sc.SId = sid; sc.componentNr = c; sc.TMid = q[i] >> 6;
                            if (sc.TMid == 0) { sc.ASCTy = q[i] & 0x3F; sc.subchannelId = q[i + 1] >> 2; }
                            else if (sc.TMid == 1) { sc.DSCTy = q[i] & 0x3F; sc.subchannelId = q[i + 1] >> 2; }
                            else if (sc.TMid == 3) { sc.SCId = ((q[i] & 0x3F) << 6) | (q[i + 1] >> 2); }
                            sc.PS_flag = (q[i + 1] >> 1) & 1; sc.CAflag = q[i + 1] & 1;
                            v.push_back(sc);
                        }
                        if (sighting(sid)) fresh.push_back(sid);
                        comps[sid] = v;
                    }
                }
            } else if (type == 1 && len >= 19) {
                const int ext = d[0] & 7;
                std::string label((const char*)d + 3, 16);
                const uint16_t flag = d[19] << 8 | d[20 > len ? len : 20];
                if (ext == 0) { ensLabel.fig1_label = label; ensLabel.fig1_flag = flag; newEnsLabel = true; }
                else if (ext == 1) {
                    const uint32_t sid = d[1] << 8 | d[2];
                    auto it = services.find(sid);
                    if (it != services.end()) { it->second.serviceLabel.fig1_label = label; it->second.serviceLabel.fig1_flag = flag; }
                }
            }
            p += 1 + len;
        }
        return fresh;
    }
};

} // namespace

struct RadioReceiver::Impl {
    RadioControllerInterface& rci; InputInterface& input; RadioReceiverOptions rro; DABParams params;
    dabb_ctx* ctx = nullptr;
    std::thread worker; std::atomic<bool> running{false};
    FigDb db;
    struct Slot { bool used = false; uint32_t sid = 0; ProgrammeHandlerInterface* handler = nullptr; FILE* dump = nullptr; int bitrate = 0; bool dabplus = false; };
    Slot slots[DABB_MAX_SUBCH];
    std::mutex slotMutex;
    std::mutex ctxMutex;                    /* serialises dabb_set_options with dabb_process (one submitting thread per handle) */
    bool synced = false; float snr = 0; int snrCount = 0; long sampleCnt = 0;

    /* the reference's enumerators (radio-receiver-options.h:35-64) -> DABB_PLACEMENT_* / DABB_FREQSYNC_* (0 = the reference's default) */
    static int placementOf(FFTPlacementMethod m) { return m == FFTPlacementMethod::StrongestPeak ? DABB_PLACEMENT_STRONGEST_PEAK : m == FFTPlacementMethod::EarliestPeakWithBinning ? DABB_PLACEMENT_EARLIEST_PEAK_WITH_BINNING : DABB_PLACEMENT_THRESHOLD_BEFORE_PEAK; }
    static int freqsyncOf(FreqsyncMethod m) { return m == FreqsyncMethod::GetMiddle ? DABB_FREQSYNC_GET_MIDDLE : m == FreqsyncMethod::CorrelatePRS ? DABB_FREQSYNC_CORRELATE_PRS : DABB_FREQSYNC_PATTERN_OF_ZEROS; }
    void applyOptions(const RadioReceiverOptions& o)
    {
        dabb_options op; memset(&op, 0, sizeof op);
        op.disable_coarse = o.disableCoarseCorrector ? 1 : 0; op.fft_placement = placementOf(o.fftPlacementMethod); op.freqsync_method = freqsyncOf(o.freqsyncMethod);
        std::lock_guard<std::mutex> l(ctxMutex);
        dabb_set_options(ctx, &op);
    }

    Impl(RadioControllerInterface& r, InputInterface& i, RadioReceiverOptions o, int mode) : rci(r), input(i), rro(o), params(mode)
    {
        if (mode != 1) throw std::runtime_error("B200 backend: only transmission mode I is implemented");
        dabb_config cfg; memset(&cfg, 0, sizeof cfg);
        cfg.abi_version = DABB_ABI_VERSION; cfg.device = 0; cfg.n_streams = 1; cfg.transmission_mode = 1; cfg.fft_mode = DABB_FFT_EXACT;
        cfg.disable_coarse = o.disableCoarseCorrector ? 1 : 0; cfg.fft_placement = placementOf(o.fftPlacementMethod); cfg.freqsync_method = freqsyncOf(o.freqsyncMethod);
        cfg.keep_taps = 1; cfg.n_subch_slots = DABB_MAX_SUBCH; cfg.max_subch_cu = 416; cfg.ofdm_groups = 25;
        if (dabb_create(&cfg, &ctx) != DABB_OK) throw std::runtime_error(std::string("B200 backend: ") + dabb_last_error(nullptr));
    }
    ~Impl() { stopWorker(); closeSlots(); if (ctx) dabb_destroy(ctx); }

    void closeSlots() { std::lock_guard<std::mutex> l(slotMutex); for (auto& s : slots) { if (s.dump) fclose(s.dump); s = Slot(); } }
    void stopWorker() { running = false; if (worker.joinable()) worker.join(); }

    /* pull exactly n samples like OFDMProcessor::getSamples (ofdm-processor.cpp:186-210); false = input failure / stop */
    bool pull(DSPCOMPLEX* dst, int64_t n)
    {
        while (n > 0 && running) {
            int32_t avail = input.getSamplesToRead();
            while (avail == 0 && running) {
                if (!input.is_ok()) return false;
                std::this_thread::sleep_for(std::chrono::microseconds(10));
                avail = input.getSamplesToRead();
            }
            if (!running) return false;
            const int32_t want = (int32_t)std::min<int64_t>(n, std::min<int64_t>(avail, 32768));
            const int32_t got = input.getSamples(dst, want);
            if (got <= 0) { if (!input.is_ok()) return false; continue; }
            dst += got; n -= got; sampleCnt += got;
        }
        return n == 0;
    }

    void run()
    {
        const int64_t TF = DABB_TF;
        const int64_t need_track = DABB_TU + (DABB_TU - 1) + 75LL * DABB_TS + DABB_TNULL;
        std::vector<DSPCOMPLEX> buf((size_t)(8 * TF));
        int64_t acq_need = 3 * TF;
        int64_t buf_start = 0, have = 0, pos = 0;     /* logical index of buf[0], samples held, receiver position */
        bool tracking = false;
        dabb_frame_result res; uint8_t fibs[12 * 32];
        std::vector<uint8_t> msc((size_t)DABB_MAX_SUBCH * 4 * 1152), sf((size_t)DABB_MAX_SUBCH * 5760);
        dabb_stream_reset(ctx, 0, 1, 0);
        bool failed = false;
        while (running) {
            /* acquisition needs the sLevel warm-up, the null search and a whole frame: 3 frames; tracking one frame */
            const int64_t need_end = pos + (tracking ? need_track : acq_need);
            if (pos > buf_start + TF) {            /* drop consumed samples */
                const int64_t drop = pos - buf_start;
                memmove(buf.data(), buf.data() + drop, (size_t)(have - drop) * sizeof(DSPCOMPLEX));
                buf_start += drop; have -= drop;
            }
            const int64_t missing = need_end - (buf_start + have);
            if (missing > 0) { if (!pull(buf.data() + have, missing)) { failed = running.load(); break; } have += missing; }
            dabb_io io; memset(&io, 0, sizeof io);
            io.iq = reinterpret_cast<const float*>(buf.data()); io.iq_is_host = 1; io.stride_samples = (int64_t)buf.size(); io.buf_start = &buf_start; io.buf_len = have;
            io.results = &res; io.fibs = fibs; io.msc = msc.data(); io.msc_stride = 1152; io.sf = sf.data(); io.sf_stride = 5760;
            int prc; { std::lock_guard<std::mutex> l(ctxMutex); prc = dabb_process(ctx, &io); }
            if (prc != DABB_OK) { rci.onMessage(message_level_t::Error, "B200 backend", dabb_last_error(ctx)); failed = true; break; }
            pos = res.next_pos;
            if (res.status == DABB_FRAME_DECODED) {
                if (!synced) { synced = true; rci.onSyncChange(true); }
                tracking = true; acq_need = 3 * TF;
                dispatch(res, fibs, msc.data(), sf.data());
            } else if (res.status == DABB_FRAME_NO_SYNC) {
                if (synced) { synced = false; rci.onSyncChange(false); }
                tracking = false;
            } else if (res.status == DABB_FRAME_ACQUIRING) {
                if (synced) { synced = false; rci.onSyncChange(false); }
                tracking = false;
                /* the null search did not finish inside the samples at hand: give it one more frame, and after five frames
                 * without a null start over two frames further on */
                if (acq_need < 5 * TF) acq_need += TF;
                else { pos += 2 * TF; acq_need = 3 * TF; dabb_stream_reset(ctx, 0, 1, pos); }
            }
        }
        if (failed) { running = false; rci.onInputFailure(); }
    }

    void dispatch(const dabb_frame_result& r, const uint8_t* fibs, const uint8_t* msc, const uint8_t* sf)
    {
        /* OfdmDecoder::processPRS: snr = 0.7 snr + 0.3 get_snr, reported every 10th frame (ofdm-decoder.cpp:156-160) */
        snr = 0.7f * snr + 0.3f * (float)r.snr_raw;
        if (++snrCount > 10) { rci.onSNR(snr); snrCount = 0; }
        if (sampleCnt > INPUT_RATE / 5) { rci.onFrequencyCorrectorChange(r.fine_corr, r.coarse_corr); sampleCnt = 0; }   /* ofdm-processor.cpp:218-223 */
        {
            std::vector<float> cir(DABB_TU);
            if (dabb_read_tap(ctx, 1, cir.data(), cir.size() * sizeof(float)) == DABB_OK) rci.onNewImpulseResponse(std::move(cir));
        }
        {   /* OfdmDecoder hands over r1 of every 96th carrier once per frame (ofdm-decoder.cpp:119-125,216-218) */
            std::vector<DSPCOMPLEX> pts((size_t)(DABB_L - 1) * DABB_K / 96);
            if (dabb_read_tap(ctx, 2, pts.data(), pts.size() * sizeof(DSPCOMPLEX)) == DABB_OK) rci.onConstellationPoints(std::move(pts));
        }
        {   /* the null symbol that follows the frame, as read with the corrected oscillator (ofdm-processor.cpp:462-469) */
            std::vector<DSPCOMPLEX> nul(DABB_TNULL);
            if (dabb_read_tap(ctx, 3, nul.data(), nul.size() * sizeof(DSPCOMPLEX)) == DABB_OK) rci.onNewNullSymbol(std::move(nul));
        }
        for (int f = 0; f < 12; f++) {
            uint8_t bits[256];
            for (int i = 0; i < 256; i++) bits[i] = (fibs[32 * f + (i >> 3)] >> (7 - (i & 7))) & 1;
            const bool ok = (r.fib_crc_mask >> f) & 1;
            rci.onFIBDecodeSuccess(ok, bits);
            if (ok) {
                bool newEns = false, newLabel = false;
                const auto fresh = db.parseFib(fibs + 32 * f, newEns, newLabel);
                if (newEns) rci.onNewEnsemble(db.eid);
                if (newLabel) { DabLabel l = db.ensLabel; rci.onSetEnsembleLabel(l); }
                for (uint32_t sid : fresh) rci.onServiceDetected(sid);
            }
        }
        std::lock_guard<std::mutex> l(slotMutex);
        for (int k = 0; k < DABB_MAX_SUBCH; k++) {
            Slot& s = slots[k];
            if (!s.used) continue;
            const int flen = 3 * s.bitrate;
            for (int c = 4 - r.n_logical[k]; c < 4; c++) {
                const uint8_t* fr = msc + ((size_t)k * 4 + c) * 1152;
                if (s.dump) fwrite(fr, flen, 1, s.dump);                 /* DecoderAdapter::addtoFrame dump (decoder_adapter.cpp:71-73) */
                if (s.handler) s.handler->onFrameErrors(0);
            }
            for (int e = 0; e < r.n_rs_events[k]; e++)
                if (s.handler) s.handler->onRsErrors((r.rs_uncorr_mask[k] >> e) & 1, r.rs_corr[k][e]);   /* DecoderAdapter::FECInfo */
            if (r.sf_ready[k] && s.handler) s.handler->onSuperframe(sf + (size_t)k * 5760, 5 * flen, r.sf_au_count[k], r.sf_au_crc_mask[k]);
        }
    }

    bool play(ProgrammeHandlerInterface& handler, const std::string& dumpFileName, const Service& srv, bool unique)
    {
        Subchannel sub; AudioServiceComponentType at = AudioServiceComponentType::Unknown;
        {
            std::lock_guard<std::mutex> l(db.m);
            auto it = db.comps.find(srv.serviceId);
            if (it == db.comps.end()) return false;
            for (const auto& sc : it->second) {
                if (sc.transportMode() != TransportMode::Audio) continue;
                auto st = db.subch.find(sc.subchannelId);
                if (st == db.subch.end()) continue;
                if (sc.audioType() == AudioServiceComponentType::Unknown) continue;
                sub = st->second; at = sc.audioType();
                break;
            }
        }
        if (!sub.valid()) return false;
        std::lock_guard<std::mutex> l(slotMutex);
        if (unique) for (int k = 0; k < DABB_MAX_SUBCH; k++) if (slots[k].used) { dabb_remove_subchannel(ctx, 0, 1, k); if (slots[k].dump) fclose(slots[k].dump); slots[k] = Slot(); }
        for (int k = 0; k < DABB_MAX_SUBCH; k++) if (slots[k].used && slots[k].sid == srv.serviceId) return true;   /* already decoding (msc-handler.cpp:69-74) */
        int k = 0;
        while (k < DABB_MAX_SUBCH && slots[k].used) k++;
        if (k == DABB_MAX_SUBCH) return false;
        dabb_subchannel sc; memset(&sc, 0, sizeof sc);
        sc.subch_id = sub.subChId; sc.start_cu = sub.startAddr; sc.length_cu = sub.length; sc.bitrate = sub.bitrate();
        sc.short_form = sub.protectionSettings.shortForm; sc.uep_level = sub.protectionSettings.uepLevel;
        sc.eep_profile_a = sub.protectionSettings.eepProfile == EEPProtectionProfile::EEP_A; sc.eep_level = (int)sub.protectionSettings.eepLevel;
        sc.dabplus = at == AudioServiceComponentType::DABPlus;
        if (dabb_select_subchannel(ctx, 0, 1, k, &sc) != DABB_OK) { rci.onMessage(message_level_t::Error, "B200 backend", dabb_last_error(ctx)); return false; }
        Slot& s = slots[k]; s.used = true; s.sid = srv.serviceId; s.handler = &handler; s.bitrate = sc.bitrate; s.dabplus = sc.dabplus;
        if (!dumpFileName.empty()) s.dump = fopen(dumpFileName.c_str(), "wb");
        return true;
    }
};

RadioReceiver::RadioReceiver(RadioControllerInterface& rci, InputInterface& input, RadioReceiverOptions rro, int transmission_mode)
{
    DABParams check(transmission_mode);   /* throws std::out_of_range like the reference for a bad mode */
    (void)check;
    d.reset(new Impl(rci, input, rro, transmission_mode));
}
RadioReceiver::~RadioReceiver() {}

void RadioReceiver::restart(bool doScan)
{
    (void)doScan;
    d->stopWorker(); d->closeSlots(); d->db.clear();
    for (int k = 0; k < DABB_MAX_SUBCH; k++) dabb_remove_subchannel(d->ctx, 0, 1, k);
    d->input.restart();
    d->synced = false; d->running = true;
    d->worker = std::thread(&Impl::run, d.get());
}
void RadioReceiver::restart_decoder()
{
    d->closeSlots(); d->db.clear();
    for (int k = 0; k < DABB_MAX_SUBCH; k++) dabb_remove_subchannel(d->ctx, 0, 1, k);
}
void RadioReceiver::stop()
{
    d->stopWorker(); d->closeSlots(); d->db.clear();
}
void RadioReceiver::setReceiverOptions(const RadioReceiverOptions rro) { d->rro = rro; d->applyOptions(rro); }
bool RadioReceiver::playSingleProgramme(ProgrammeHandlerInterface& h, const std::string& dump, const Service& s) { return d->play(h, dump, s, true); }
bool RadioReceiver::addServiceToDecode(ProgrammeHandlerInterface& h, const std::string& dump, const Service& s) { return d->play(h, dump, s, false); }
bool RadioReceiver::removeServiceToDecode(const Service& s)
{
    std::lock_guard<std::mutex> l(d->slotMutex);
    for (int k = 0; k < DABB_MAX_SUBCH; k++)
        if (d->slots[k].used && d->slots[k].sid == s.serviceId) {
            dabb_remove_subchannel(d->ctx, 0, 1, k);
            if (d->slots[k].dump) fclose(d->slots[k].dump);
            d->slots[k] = Impl::Slot();
            return true;
        }
    return false;
}
uint16_t RadioReceiver::getEnsembleId(void) const { std::lock_guard<std::mutex> l(d->db.m); return d->db.eid; }
uint8_t RadioReceiver::getEnsembleEcc(void) const { std::lock_guard<std::mutex> l(d->db.m); return d->db.ecc; }
DabLabel RadioReceiver::getEnsembleLabel(void) const { std::lock_guard<std::mutex> l(d->db.m); return d->db.ensLabel; }
std::vector<Service> RadioReceiver::getServiceList(void) const
{
    std::lock_guard<std::mutex> l(d->db.m);
    std::vector<Service> v;
    for (const auto& kv : d->db.services) v.push_back(kv.second);
    return v;
}
Service RadioReceiver::getService(uint32_t sId) const
{
    std::lock_guard<std::mutex> l(d->db.m);
    auto it = d->db.services.find(sId);
    return it == d->db.services.end() ? Service(0) : it->second;
}
std::list<ServiceComponent> RadioReceiver::getComponents(const Service& s) const
{
    std::lock_guard<std::mutex> l(d->db.m);
    std::list<ServiceComponent> out;
    auto it = d->db.comps.find(s.serviceId);
    if (it != d->db.comps.end()) out.assign(it->second.begin(), it->second.end());
    return out;
}
bool RadioReceiver::serviceHasAudioComponent(const Service& s) const
{
    for (const auto& sc : getComponents(s))
        if (sc.transportMode() == TransportMode::Audio && sc.audioType() != AudioServiceComponentType::Unknown) return true;
    return false;
}
Subchannel RadioReceiver::getSubchannel(const ServiceComponent& sc) const
{
    std::lock_guard<std::mutex> l(d->db.m);
    auto it = d->db.subch.find(sc.subchannelId);
    return it == d->db.subch.end() ? Subchannel() : it->second;
}
DABParams& RadioReceiver::getParams() { return d->params; }
RadioReceiverStats RadioReceiver::getReceiverStats() const { return RadioReceiverStats(); }
