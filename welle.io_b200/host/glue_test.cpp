// glue_test — drives the host glue (RadioReceiver over libdab_b200.so) the way the reference's harnesses do
// (src/tests/backend_tests.cpp:103-155, welle-cli -D): file-backed InputInterface, FIB + .msc dumps, RS statistics.
//   glue_test <in.cf32> <out-prefix> [select_at_fib=12] [disable_coarse=1] [decode_tii=0] [controller_thread=0]
// controller_thread=1: the service is selected, removed and selected again from the main thread while the receiver's worker thread
// is decoding (what a GUI does; RadioReceiver serialises the calls into the context), instead of from the FIB callback
#include "radio-receiver.h"
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>

struct FileInput : InputInterface {
    std::vector<DSPCOMPLEX> d; std::atomic<size_t> pos{0};
    explicit FileInput(const char* fn) { FILE* f = fopen(fn, "rb"); if (!f) { perror(fn); exit(2); } fseek(f, 0, SEEK_END); size_t n = ftell(f) / 8; fseek(f, 0, SEEK_SET); d.resize(n); if (fread(d.data(), 8, n, f) != n) exit(2); fclose(f); }
    void setFrequency(int) override {} int getFrequency() const override { return 0; }
    bool is_ok() override { return pos.load() < d.size(); } bool restart() override { return true; } void stop() override {} void reset() override {}
    int32_t getSamples(DSPCOMPLEX* b, int32_t n) override { size_t p = pos.load(); size_t m = std::min<size_t>(n, d.size() - p); memcpy(b, d.data() + p, m * 8); pos = p + m; return (int32_t)m; }
    std::vector<DSPCOMPLEX> getSpectrumSamples(int n) override { return std::vector<DSPCOMPLEX>(n); }
    int32_t getSamplesToRead() override { size_t p = pos.load(); return (int32_t)std::min<size_t>(d.size() - p, 1 << 20); }
    float setGain(int) override { return 0; } float getGain() const override { return 0; } int getGainCount() override { return 0; } void setAgc(bool) override {}
    std::string getDescription() override { return "file"; }
};
struct Prog : ProgrammeHandlerInterface {
    FILE* rs; int frames = 0, sfs = 0;
    void onFrameErrors(int) override { frames++; } void onNewAudio(std::vector<int16_t>&&, int, const std::string&) override {}
    void onRsErrors(bool u, int c) override { fprintf(rs, "%d %d\n", u ? 1 : 0, c); } void onAacErrors(int) override {} void onNewDynamicLabel(const std::string&) override {}
    void onMOT(const mot_file_t&) override {} void onPADLengthError(size_t, size_t) override {}
    void onSuperframe(const uint8_t*, size_t, int, int) override { sfs++; }
};
struct Ctl : RadioControllerInterface {
    RadioReceiver* rx = nullptr; Prog* ph = nullptr; FILE* fibs; std::string dump; int select_at = 12, nfib = 0, ok = 0; bool sel = false, selok = false; std::atomic<bool> failed{false};
    int syncs = 0, services = 0, cirs = 0, consts = 0, nulls = 0; size_t tapsz_ok = 1;
    void onSNR(float) override {} void onFrequencyCorrectorChange(int, int) override {} void onSyncChange(char s) override { if (s) syncs++; } void onSignalPresence(bool) override {}
    void onServiceDetected(uint32_t) override { services++; } void onNewEnsemble(uint16_t) override {} void onSetEnsembleLabel(DabLabel&) override {} void onDateTimeUpdate(const dab_date_time_t&) override {}
    void onFIBDecodeSuccess(bool o, const uint8_t* fib) override {
        uint8_t rec[33]; rec[0] = o; for (int i = 0; i < 32; i++) { uint8_t b = 0; for (int j = 0; j < 8; j++) b = (b << 1) | (fib[8 * i + j] & 1); rec[1 + i] = b; }
        fwrite(rec, 33, 1, fibs); nfib++; ok += o;
        if (!from_controller && !sel && nfib >= select_at) { auto l = rx->getServiceList(); if (!l.empty()) { sel = true; selok = rx->playSingleProgramme(*ph, dump, l.front()); } }
    }
    void onNewImpulseResponse(std::vector<float>&& v) override { cirs++; tapsz_ok &= v.size() == 2048; }
    void onConstellationPoints(std::vector<DSPCOMPLEX>&& v) override { consts++; tapsz_ok &= v.size() == 1200; }   /* (L-1) K / 96 */
    void onNewNullSymbol(std::vector<DSPCOMPLEX>&& v) override { nulls++; tapsz_ok &= v.size() == 2656; }
    FILE* tii = nullptr; int tiis = 0; bool from_controller = false; std::atomic<int> zaps{0};
    void onTIIMeasurement(tii_measurement_t&& m) override { tiis++; if (tii) fprintf(tii, "%d %d %d %.1f\n", m.comb, m.pattern, m.delay_samples, m.error); } void onMessage(message_level_t, const std::string& a, const std::string& b) override { fprintf(stderr, "msg: %s %s\n", a.c_str(), b.c_str()); }
    void onInputFailure() override { failed = true; }
};
int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: glue_test in.cf32 out-prefix [select_at_fib]\n"); return 2; }
    FileInput in(argv[1]); Ctl ri; Prog ph; std::string pre = argv[2];
    ri.fibs = fopen((pre + ".fibs").c_str(), "wb"); ph.rs = fopen((pre + ".rs").c_str(), "w"); ri.dump = pre + ".msc"; ri.ph = &ph;
    if (argc > 3) ri.select_at = atoi(argv[3]);
    RadioReceiverOptions rro; rro.disableCoarseCorrector = argc > 4 ? atoi(argv[4]) != 0 : true;    /* default like the parity harness (welle-cli -u) */
    rro.decodeTII = argc > 5 && atoi(argv[5]) != 0;                                                /* welle-cli -T */
    if (rro.decodeTII) ri.tii = fopen((pre + ".tii").c_str(), "w");
    ri.from_controller = argc > 6 && atoi(argv[6]) != 0;
    double secs = 0;
    {
        RadioReceiver rx(ri, in, rro);
        ri.rx = &rx;
        const auto t0 = std::chrono::steady_clock::now();
        rx.restart(false);
        if (ri.from_controller) {
            /* select - remove - select ... from this thread, a few milliseconds apart, while the worker decodes */
            while (!ri.failed.load() && rx.getServiceList().empty()) std::this_thread::sleep_for(std::chrono::microseconds(200));
            for (int z = 0; z < 4 && !ri.failed.load(); z++) {
                auto l = rx.getServiceList();
                if (l.empty()) break;
                ri.selok = rx.playSingleProgramme(ph, ri.dump, l.front()); ri.sel = true; ri.zaps++;
                if (z == 3) break;
                std::this_thread::sleep_for(std::chrono::milliseconds(4));
                rx.removeServiceToDecode(l.front());
                std::this_thread::sleep_for(std::chrono::milliseconds(2));
            }
        }
        while (!ri.failed.load()) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();      /* restart .. input exhausted */
        rx.stop();
    }
    fclose(ri.fibs); fclose(ph.rs); if (ri.tii) fclose(ri.tii);
    printf("tii=%d zaps=%d ", ri.tiis, ri.zaps.load());
    printf("fibs=%d ok=%d services=%d selected=%d logical_frames=%d superframes=%d syncs=%d cirs=%d consts=%d nulls=%d tapsizes=%d seconds=%.4f\n", ri.nfib, ri.ok, ri.services, ri.selok ? 1 : 0, ph.frames, ph.sfs, ri.syncs,
           ri.cirs, ri.consts, ri.nulls, (int)ri.tapsz_ok, secs);
    return 0;
}
