// rawfile_test — the drop-in boundary exercised with the reference's OWN input class: this file is compiled together with the
// UNMODIFIED /root/reference/src/input/raw_file.cpp (+ its raw_file.h, virtual_input.h and various/ringbuffer.h where they lie)
// against the B200 host glue's headers (this directory: radio-controller.h, dab-constants.h, radio-receiver.h are the include
// names the reference uses) and linked with libwelle_b200_host.so - recipe: oracle/Makefile, target `rawfile`.
// The calls are the ones welle-cli makes (welle-cli.cpp:514-516,612-664) / tests/backend_tests.cpp:103-155 makes:
//   CRAWFile in(rci, /*throttle*/false, /*rewind*/false); in.setFileName(path, "auto"); RadioReceiver rx(rci, in, rro);
//   rx.restart(false); ... rx.playSingleProgramme(handler, dump, service); ... in.endWasReached(); rx.stop();
// Output like glue_test: <prefix>.fibs (33 bytes per FIB: CRC flag + 32 bytes), <prefix>.msc (logical frames), <prefix>.rs.
//   rawfile_test <recording.{cf32,u8,s8,s16le,s16be}.iq> <out-prefix> [select_at_fib=12] [disable_coarse=1] [scan=0]
#include "raw_file.h"
#include "radio-receiver.h"
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>

struct Prog : ProgrammeHandlerInterface {
    FILE* rs = nullptr; int frames = 0, sfs = 0;
    void onFrameErrors(int) override { frames++; } void onNewAudio(std::vector<int16_t>&&, int, const std::string&) override {}
    void onRsErrors(bool u, int c) override { fprintf(rs, "%d %d\n", u ? 1 : 0, c); } void onAacErrors(int) override {} void onNewDynamicLabel(const std::string&) override {}
    void onMOT(const mot_file_t&) override {} void onPADLengthError(size_t, size_t) override {}
    void onSuperframe(const uint8_t*, size_t, int, int) override { sfs++; }
};
struct Ctl : RadioControllerInterface {
    RadioReceiver* rx = nullptr; Prog* ph = nullptr; FILE* fibs = nullptr; std::string dump; int select_at = 12, nfib = 0, ok = 0; bool sel = false, selok = false;
    std::atomic<bool> failed{false}; std::atomic<int> syncs{0}; int services = 0, presence_true = 0, presence_false = 0, snrs = 0;
    void onSNR(float) override { snrs++; } void onFrequencyCorrectorChange(int, int) override {} void onSyncChange(char s) override { if (s) syncs++; }
    void onSignalPresence(bool p) override { if (p) presence_true++; else presence_false++; }
    void onServiceDetected(uint32_t) override { services++; } void onNewEnsemble(uint16_t) override {} void onSetEnsembleLabel(DabLabel&) override {} void onDateTimeUpdate(const dab_date_time_t&) override {}
    void onFIBDecodeSuccess(bool o, const uint8_t* fib) override {
        uint8_t rec[33]; rec[0] = o; for (int i = 0; i < 32; i++) { uint8_t b = 0; for (int j = 0; j < 8; j++) b = (b << 1) | (fib[8 * i + j] & 1); rec[1 + i] = b; }
        fwrite(rec, 33, 1, fibs); nfib++; ok += o;
        if (!sel && nfib >= select_at) { auto l = rx->getServiceList(); if (!l.empty()) { sel = true; selok = rx->playSingleProgramme(*ph, dump, l.front()); } }
    }
    void onNewImpulseResponse(std::vector<float>&&) override {} void onConstellationPoints(std::vector<DSPCOMPLEX>&&) override {} void onNewNullSymbol(std::vector<DSPCOMPLEX>&&) override {}
    void onTIIMeasurement(tii_measurement_t&&) override {}
    void onMessage(message_level_t, const std::string& a, const std::string& b) override { fprintf(stderr, "msg: %s %s\n", a.c_str(), b.c_str()); }
    void onInputFailure() override { failed = true; }
};

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: rawfile_test recording.iq out-prefix [select_at_fib] [disable_coarse] [scan]\n"); return 2; }
    Ctl ri; Prog ph; const std::string pre = argv[2];
    ri.fibs = fopen((pre + ".fibs").c_str(), "wb"); ph.rs = fopen((pre + ".rs").c_str(), "w"); ri.dump = pre + ".msc"; ri.ph = &ph;
    if (argc > 3) ri.select_at = atoi(argv[3]);
    RadioReceiverOptions rro; rro.disableCoarseCorrector = argc > 4 ? atoi(argv[4]) != 0 : true;
    const bool scan = argc > 5 && atoi(argv[5]) != 0;
    CRAWFile in(ri, false, false);                       /* the reference's own file input, un-throttled, no rewind */
    in.setFileName(argv[1], "auto");
    const auto t0 = std::chrono::steady_clock::now();
    {
        RadioReceiver rx(ri, in, rro);
        ri.rx = &rx;
        rx.restart(scan);
        /* after the end of the file CRAWFile delivers zeros for ever (raw_file.cpp:276-280): stop a little after endWasReached() */
        while (!in.endWasReached() && !ri.failed.load()) std::this_thread::sleep_for(std::chrono::milliseconds(5));
        std::this_thread::sleep_for(std::chrono::milliseconds(300));
        const auto stats = rx.getReceiverStats();
        const double age = std::chrono::duration<double>(std::chrono::system_clock::now() - stats.timeLastFCT0Frame).count();
        rx.stop();
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fclose(ri.fibs); fclose(ph.rs);
        printf("fibs=%d ok=%d services=%d selected=%d logical_frames=%d superframes=%d syncs=%d presence_true=%d presence_false=%d snr_reports=%d fct0_age_s=%.3f seconds=%.3f\n",
               ri.nfib, ri.ok, ri.services, ri.selok ? 1 : 0, ph.frames, ph.sfs, ri.syncs.load(), ri.presence_true, ri.presence_false, ri.snrs, age, secs);
    }
    return 0;
}
