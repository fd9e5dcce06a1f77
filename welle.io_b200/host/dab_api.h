/*
 * dab_api.h — host-side mirror of the reference receiver's public surface for the PHY decode path, so that callers
 * written against the reference (welle-cli, the Qt test harness, input devices) compile and run against the B200
 * backend unchanged.  Names, argument meaning, ownership and threading follow the reference:
 *   InputInterface               backend/radio-controller.h:188-215   (pull interface implemented by CVirtualInput / CRAWFile)
 *   RadioControllerInterface     backend/radio-controller.h:83-138    (callbacks from the backend worker thread)
 *   ProgrammeHandlerInterface    backend/radio-controller.h:142-178   (per selected sub-channel)
 *   RadioReceiverOptions         backend/radio-receiver-options.h:66-85
 *   RadioReceiver                backend/radio-receiver.h:52-116
 *   Service / ServiceComponent / Subchannel / DabLabel / DABParams    backend/dab-constants.h:71-198
 * Only declarations live here; the implementation (radio-receiver.cpp) drives libdab_b200.so through include/dab_b200.h.
 * Parts of the reference surface that belong to out-of-scope subsystems (audio decoding, PAD/MOT, TII, FIG 2 labels,
 * date/time) are declared so that user code compiles, and are simply never invoked by this backend.
 */
#ifndef DABB_HOST_API_H
#define DABB_HOST_API_H
#include <chrono>
#include <complex>
#include <cstdint>
#include <list>
#include <map>
#include <memory>
#include <string>
#include <vector>

typedef float DSPFLOAT;
typedef std::complex<DSPFLOAT> DSPCOMPLEX;
typedef std::complex<float> complexf;
typedef int8_t softbit_t;
#define INPUT_RATE 2048000

enum class CharacterSet : uint8_t { EbuLatin = 0x00, UnicodeUcs2 = 0x06, UnicodeUtf8 = 0x0F, Undefined };     /* backend/charsets.h:34-39 */
enum class TransportMode { Audio = 0, StreamData = 1, FIDC = 2, PacketData = 3 };
enum class AudioServiceComponentType { DAB, DABPlus, Unknown };

class DABParams {
public:
    explicit DABParams(int mode = 1);
    void setMode(int mode);          /* throws std::out_of_range for anything but mode 1..4; only mode 1 decodes */
    uint8_t dabMode;
    int16_t L, K, T_null;
    int32_t T_F;
    int16_t T_s, T_u, guardLength, carrierDiff;
};

struct DabLabel {
    /* FIG 1 label, encoded according to charset (usually EBU Latin, ETSI TS 101 756 Annex C) */
    CharacterSet charset = CharacterSet::EbuLatin;
    std::string fig1_label;
    uint16_t fig1_flag = 0;
    void setCharset(uint8_t charset_id) { charset = static_cast<CharacterSet>(charset_id); }
    /* the EBU Latin -> UTF-8 table is a UI concern and not part of this backend: bytes >= 0x80 (and the few EBU code points below
     * 0x80 that differ from ASCII) are returned as they are */
    std::string fig1_label_utf8() const { return fig1_label; }
    /* the abbreviated label: the characters whose bit (MSB = first character) is set in the 16-bit flag field (EN 300 401 5.2.2.3) */
    std::string fig1_shortlabel_utf8() const
    {
        std::string o;
        for (size_t i = 0; i < fig1_label.size() && i < 16; i++) if (fig1_flag & (0x8000u >> i)) o += fig1_label[i];
        return o;
    }

    /* extended label from FIG 2 segments (UTF-8 or UCS-2), same fields as the reference (backend/dab-constants.h:88-118) */
    std::map<int, std::vector<uint8_t>> segments;
    size_t segment_count = 0;
    CharacterSet extended_label_charset = CharacterSet::Undefined;
    uint8_t toggle_flag = 0;
    bool fig2_rfu = false;
    /* all segments concatenated as UTF-8; empty until every segment has arrived */
    std::string fig2_label() const
    {
        std::vector<uint8_t> cat;
        for (size_t i = 0; i < segment_count; i++) {
            auto it = segments.find((int)i);
            if (it == segments.end()) return std::string();
            cat.insert(cat.end(), it->second.begin(), it->second.end());
        }
        if (extended_label_charset == CharacterSet::UnicodeUtf8) return std::string(cat.begin(), cat.end());
        if (extended_label_charset != CharacterSet::UnicodeUcs2) return std::string();
        /* UCS-2 / UTF-16 -> UTF-8.  The reference reinterprets the segment bytes as host-order char16_t (charsets.cpp:111-121), i.e.
         * little endian on the machines it runs on although DAB transmits UCS-2 big endian; kept so that labels come out the same */
        std::string o;
        auto put = [&o](unsigned c) {
            if (c < 0x80) o += (char)c;
            else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 0x3F)); }
            else if (c < 0x10000) { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 0x3F)); o += (char)(0x80 | (c & 0x3F)); }
            else { o += (char)(0xF0 | (c >> 18)); o += (char)(0x80 | ((c >> 12) & 0x3F)); o += (char)(0x80 | ((c >> 6) & 0x3F)); o += (char)(0x80 | (c & 0x3F)); }
        };
        for (size_t i = 0; i + 1 < cat.size(); i += 2) {
            unsigned c = (unsigned)cat[i] | (unsigned)cat[i + 1] << 8;
            if (c >= 0xD800 && c < 0xDC00 && i + 3 < cat.size()) {
                const unsigned lo = (unsigned)cat[i + 2] | (unsigned)cat[i + 3] << 8;
                if (lo >= 0xDC00 && lo < 0xE000) { c = 0x10000 + ((c - 0xD800) << 10) + (lo - 0xDC00); i += 2; }
            }
            if (c >= 0xD800 && c < 0xE000) c = 0xFFFD;        /* unpaired surrogate: the reference's converter throws here */
            put(c);
        }
        return o;
    }
    std::string utf8_label() const { const std::string f2 = fig2_label(); return f2.empty() ? fig1_label_utf8() : f2; }
};

struct Service {
    explicit Service(uint32_t sid = 0) : serviceId(sid) {}
    uint32_t serviceId = 0;
    DabLabel serviceLabel;
    int16_t language = 0;
    int16_t programType = 0;
};

struct ServiceComponent {
    int8_t TMid = 0;
    uint32_t SId = 0;
    int16_t componentNr = 0;
    DabLabel componentLabel;
    int16_t ASCTy = 0;
    int16_t PS_flag = 0;
    int16_t subchannelId = 0;
    uint16_t SCId = 0;
    uint8_t CAflag = 0;
    int16_t DSCTy = 0;
    uint8_t DGflag = 0;
    int16_t packetAddress = 0;
    TransportMode transportMode() const { return (TransportMode)TMid; }
    AudioServiceComponentType audioType() const { return ASCTy == 63 ? AudioServiceComponentType::DABPlus : (ASCTy == 0 ? AudioServiceComponentType::DAB : AudioServiceComponentType::Unknown); }
};

enum class EEPProtectionProfile { EEP_A, EEP_B };
enum class EEPProtectionLevel { EEP_1 = 1, EEP_2 = 2, EEP_3 = 3, EEP_4 = 4 };

struct ProtectionSettings {
    bool shortForm = false;
    int16_t uepTableIndex = 0;
    int16_t uepLevel = 0;
    EEPProtectionProfile eepProfile = EEPProtectionProfile::EEP_A;
    EEPProtectionLevel eepLevel = EEPProtectionLevel::EEP_3;
};

struct Subchannel {
    int32_t subChId = -1;
    int32_t startAddr = 0;
    int32_t length = 0;
    bool programmeNotData = true;
    ProtectionSettings protectionSettings;
    int16_t language = 0;
    int16_t fecScheme = 0;
    int bitrate() const;             /* throws std::runtime_error("Unsupported protection") like the reference */
    int numCU() const;
    bool valid() const { return subChId != -1; }
};

struct dab_date_time_t { int year = 0, month = 0, day = 0, hour = 0, minutes = 0, seconds = 0, hourOffset = 0, minuteOffset = 0; };
struct tii_measurement_t {
    int comb = 0, pattern = 0; float error = 0; int delay_samples = 0;
    float getDelayKm(void) const { return delay_samples * (3e8f / 1000.0f / 2048000.0f); }      /* tii-decoder.cpp:133-137 */
};
struct mot_file_t { std::vector<uint8_t> data; int content_sub_type = 0; std::string content_name, click_through_url; uint8_t category = 0, slide_id = 0; std::string category_title; };
enum class message_level_t { Information, Error };

class RadioControllerInterface {
public:
    virtual ~RadioControllerInterface() {}
    virtual void onSNR(float snr) = 0;
    virtual void onFrequencyCorrectorChange(int fine, int coarse) = 0;
    virtual void onSyncChange(char isSync) = 0;
    virtual void onSignalPresence(bool isSignal) = 0;
    virtual void onServiceDetected(uint32_t sId) = 0;
    virtual void onNewEnsemble(uint16_t eId) = 0;
    virtual void onSetEnsembleLabel(DabLabel& label) = 0;
    virtual void onDateTimeUpdate(const dab_date_time_t& dateTime) = 0;
    /* fib: 256 bytes holding one bit each, valid only during the call */
    virtual void onFIBDecodeSuccess(bool crcCheckOk, const uint8_t* fib) = 0;
    virtual void onNewImpulseResponse(std::vector<float>&& data) = 0;
    virtual void onConstellationPoints(std::vector<DSPCOMPLEX>&& data) = 0;
    virtual void onNewNullSymbol(std::vector<DSPCOMPLEX>&& data) = 0;
    virtual void onTIIMeasurement(tii_measurement_t&& m) = 0;
    virtual void onMessage(message_level_t level, const std::string& text, const std::string& text2 = std::string()) = 0;
    virtual void onInputFailure(void) {}
    virtual void onRestartService(void) {}
};

class ProgrammeHandlerInterface {
public:
    virtual ~ProgrammeHandlerInterface() {}
    virtual void onFrameErrors(int frameErrors) = 0;
    virtual void onNewAudio(std::vector<int16_t>&& audioData, int sampleRate, const std::string& mode) = 0;
    virtual void onRsErrors(bool uncorrectedErrors, int numCorrectedErrors) = 0;
    virtual void onAacErrors(int aacErrors) = 0;
    virtual void onNewDynamicLabel(const std::string& label) = 0;
    virtual void onMOT(const mot_file_t& mot_file) = 0;
    virtual void onPADLengthError(size_t announced_xpad_len, size_t xpad_len) = 0;
    /* B200 backend extension (defaulted, so reference handlers keep compiling): the post-RS DAB+ superframe whose Fire
     * code matched, with the per-AU CRC mask - what the reference hands to its AAC decoder */
    virtual void onSuperframe(const uint8_t* /*sf*/, size_t /*len*/, int /*num_aus*/, int /*au_crc_ok_mask*/) {}
};

enum class DeviceParam { BiasTee, SoapySDRAntenna, SoapySDRDriverArgs, SoapySDRClockSource };

class InputInterface {
public:
    virtual ~InputInterface() {}
    virtual void setFrequency(int frequency) = 0;
    virtual int getFrequency(void) const = 0;
    virtual bool is_ok(void) = 0;
    virtual bool restart(void) = 0;
    virtual void stop(void) = 0;
    virtual void reset(void) = 0;
    virtual int32_t getSamples(DSPCOMPLEX* buffer, int32_t size) = 0;
    virtual std::vector<DSPCOMPLEX> getSpectrumSamples(int size) = 0;
    virtual int32_t getSamplesToRead(void) = 0;
    virtual float setGain(int gain) = 0;
    virtual float getGain(void) const = 0;
    virtual int getGainCount(void) = 0;
    virtual void setAgc(bool agc) = 0;
    virtual std::string getDescription(void) = 0;
    virtual bool setDeviceParam(DeviceParam, int) { return false; }
    virtual bool setDeviceParam(DeviceParam, const std::string&) { return false; }
};

enum class FreqsyncMethod { GetMiddle = 0, CorrelatePRS = 1, PatternOfZeros = 2 };
enum class FFTPlacementMethod { StrongestPeak, EarliestPeakWithBinning, ThresholdBeforePeak };
struct RadioReceiverOptions {
    FFTPlacementMethod fftPlacementMethod = FFTPlacementMethod::ThresholdBeforePeak;
    bool decodeTII = false;
    bool disableCoarseCorrector = false;
    FreqsyncMethod freqsyncMethod = FreqsyncMethod::PatternOfZeros;
};
const char* fftPlacementMethodToString(FFTPlacementMethod fft_placement);
const char* freqSyncMethodToString(FreqsyncMethod method);

struct RadioReceiverStats { std::chrono::system_clock::time_point timeLastFCT0Frame; };

class RadioReceiver {
public:
    /* throws std::out_of_range for an unknown mode, std::runtime_error if the GPU backend cannot be created */
    RadioReceiver(RadioControllerInterface& rci, InputInterface& input, RadioReceiverOptions rro, int transmission_mode = 1);
    ~RadioReceiver();
    RadioReceiver(const RadioReceiver&) = delete;
    RadioReceiver& operator=(const RadioReceiver&) = delete;

    void restart(bool doScan);
    void restart_decoder();
    void stop();
    void setReceiverOptions(const RadioReceiverOptions rro);
    bool playSingleProgramme(ProgrammeHandlerInterface& handler, const std::string& dumpFileName, const Service& s);
    bool addServiceToDecode(ProgrammeHandlerInterface& handler, const std::string& dumpFileName, const Service& s);
    bool removeServiceToDecode(const Service& s);
    uint16_t getEnsembleId(void) const;
    uint8_t getEnsembleEcc(void) const;
    DabLabel getEnsembleLabel(void) const;
    std::vector<Service> getServiceList(void) const;
    Service getService(uint32_t sId) const;
    std::list<ServiceComponent> getComponents(const Service& s) const;
    bool serviceHasAudioComponent(const Service& s) const;
    Subchannel getSubchannel(const ServiceComponent& sc) const;
    DABParams& getParams();
    RadioReceiverStats getReceiverStats() const;

private:
    struct Impl;
    std::unique_ptr<Impl> d;
};

#endif
