/*
 * batch_decode.cpp — native batch driver over the C ABI: decodes many RAW IQ recordings at once, one stream per file, on one GPU.
 *
 *   batch_decode [--format cf32|u8|s8|s16le|s16be] [--device N] [--coarse] --out DIR file1 [file2 ...]
 *
 * What `welle-cli -f FILE -D` does for one recording (welle-cli.cpp:215-235,612-707: dump the FIC, tune to the first audio
 * service, dump its logical frames) done for all files in lock-step: every iteration hands each stream's next window of samples
 * to one dabb_process() call (host buffers, copied to the device inside the call), then walks the per-stream result records:
 * FIBs (33-byte records: CRC flag + 32 bytes) go to DIR/<name>.fic, the FIGs feed the stream's service database (fig-db.h), the
 * first audio service is selected as soon as its sub-channel is known (streams with the same protection profile share a slot of
 * the library), logical frames go to DIR/<name>.msc, Reed-Solomon statistics to DIR/<name>.rs.  Streams progress independently
 * (per-stream sample position, acquisition state, end of file).  No signal processing happens here.
 *
 * Sample formats are the RAW-file formats of CRAWFile (input/raw_file.cpp:324-366); they are shipped as they are and converted on
 * the device.
 */
#include "dab_api.h"
#include "fig-db.h"
#include "../../include/dab_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

struct Stream {
    std::string path, name; FILE* in = nullptr; FILE* fic = nullptr; FILE* msc = nullptr; FILE* rs = nullptr;
    int64_t buf_start = 0, have = 0, pos = 0, acq_need = 0;      /* logical sample indices; `have` samples are buffered from buf_start */
    bool tracking = false, eof = false, done = false;
    std::unique_ptr<dabb_host::FigDb> db{new dabb_host::FigDb()};
    int slot = -1; int bitrate = 0; uint32_t sid = 0;
    long frames = 0, fibs_ok = 0, fibs = 0, logical = 0;
};

struct Profile { dabb_subchannel sc; };

int bytes_per_sample(int fmt) { return fmt == DABB_IQ_CF32 ? 8 : (fmt == DABB_IQ_U8 || fmt == DABB_IQ_S8) ? 2 : 4; }

} // namespace

int main(int argc, char** argv)
{
    int fmt = DABB_IQ_CF32, device = 0; bool coarse = false; std::string out;
    std::vector<std::string> files;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--format" && i + 1 < argc) {
            const std::string f = argv[++i];
            fmt = f == "cf32" ? DABB_IQ_CF32 : f == "u8" ? DABB_IQ_U8 : f == "s8" ? DABB_IQ_S8 : f == "s16le" ? DABB_IQ_S16LE : f == "s16be" ? DABB_IQ_S16BE : -1;
            if (fmt < 0) { fprintf(stderr, "unknown format %s\n", f.c_str()); return 2; }
        }
        else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--out" && i + 1 < argc) out = argv[++i];
        else if (a == "--coarse") coarse = true;
        else files.push_back(a);
    }
    if (out.empty() || files.empty()) { fprintf(stderr, "usage: batch_decode [--format cf32|u8|s8|s16le|s16be] [--device N] [--coarse] --out DIR file...\n"); return 2; }

    const int S = (int)files.size();
    const int64_t TF = DABB_TF, W = 6 * TF;                       /* per-stream window capacity in samples */
    const int64_t need_track = DABB_TU + (DABB_TU - 1) + 75LL * DABB_TS + DABB_TNULL;
    const int bps = bytes_per_sample(fmt);

    dabb_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = DABB_ABI_VERSION; cfg.device = device; cfg.n_streams = S; cfg.transmission_mode = 1; cfg.fft_mode = DABB_FFT_EXACT;
    cfg.disable_coarse = coarse ? 0 : 1; cfg.n_subch_slots = DABB_MAX_SUBCH; cfg.max_subch_cu = 416;
    dabb_ctx* ctx = nullptr;
    if (dabb_create(&cfg, &ctx) != DABB_OK) { fprintf(stderr, "batch_decode: %s\n", dabb_last_error(nullptr)); return 1; }

    std::vector<Stream> st((size_t)S);
    for (int s = 0; s < S; s++) {
        Stream& z = st[(size_t)s];
        z.path = files[(size_t)s];
        const size_t slash = z.path.find_last_of('/');
        z.name = slash == std::string::npos ? z.path : z.path.substr(slash + 1);
        z.in = fopen(z.path.c_str(), "rb");
        if (!z.in) { fprintf(stderr, "batch_decode: cannot open %s\n", z.path.c_str()); return 1; }
        z.fic = fopen((out + "/" + z.name + ".fic").c_str(), "wb"); z.msc = fopen((out + "/" + z.name + ".msc").c_str(), "wb"); z.rs = fopen((out + "/" + z.name + ".rs").c_str(), "w");
        if (!z.fic || !z.msc || !z.rs) { fprintf(stderr, "batch_decode: cannot write into %s\n", out.c_str()); return 1; }
        z.acq_need = 3 * TF;
    }
    dabb_stream_reset(ctx, 0, S, 0);

    std::vector<uint8_t> iq((size_t)S * (size_t)W * (size_t)bps);
    std::vector<int64_t> buf_start((size_t)S);
    std::vector<dabb_frame_result> res((size_t)S);
    std::vector<uint8_t> fibs((size_t)S * 12 * 32), msc((size_t)S * DABB_MAX_SUBCH * 4 * 1152), sf((size_t)S * DABB_MAX_SUBCH * 5760);
    std::vector<Profile> profiles;                               /* slot k decodes profiles[k] */

    int active = S;
    while (active > 0) {
        /* 1. every live stream gets the samples its next step needs: [pos, pos + need) inside its window */
        for (int s = 0; s < S; s++) {
            Stream& z = st[(size_t)s];
            uint8_t* w = iq.data() + (size_t)s * (size_t)W * (size_t)bps;
            if (z.done) { buf_start[(size_t)s] = z.buf_start; continue; }
            if (z.pos > z.buf_start) {                            /* drop consumed samples: the window always starts at the receiver position */
                const int64_t drop = std::min(z.pos - z.buf_start, z.have);
                memmove(w, w + drop * bps, (size_t)(z.have - drop) * (size_t)bps);
                z.buf_start += drop; z.have -= drop;
                if (z.buf_start < z.pos && z.have == 0) {         /* the receiver skipped ahead of everything buffered */
                    if (fseek(z.in, (long)((z.pos - z.buf_start) * bps), SEEK_CUR) != 0) z.eof = true;
                    z.buf_start = z.pos;
                }
            }
            const int64_t need_end = z.pos + (z.tracking ? need_track : z.acq_need);
            int64_t missing = std::min(need_end - (z.buf_start + z.have), W - z.have);
            if (missing > 0 && !z.eof) {
                const size_t got = fread(w + z.have * bps, (size_t)bps, (size_t)missing, z.in);
                z.have += (int64_t)got;
                if ((int64_t)got < missing) z.eof = true;
            }
            if (z.buf_start + z.have < need_end && z.eof) { z.done = true; active--; }       /* not enough samples left for another frame */
            if (z.have < W) memset(w + z.have * bps, fmt == DABB_IQ_U8 ? 0x80 : 0, (size_t)(W - z.have) * (size_t)bps);   /* silence behind the data */
            buf_start[(size_t)s] = z.buf_start;
        }
        if (active <= 0) break;
        /* the ABI takes one window length for all streams.  Whether a stream really holds the samples its next step needs was decided
         * above from its own fill level (the rest of its window is silence); a finished stream is given a window far behind its position,
         * so the library reports DABB_FRAME_NEED_SAMPLES for it and touches nothing */
        int64_t buf_len = 0;                                      /* = what is copied to the device per stream: no more than the neediest stream asks for */
        for (int s = 0; s < S; s++) {
            const Stream& z = st[(size_t)s];
            if (z.done) buf_start[(size_t)s] = -(1LL << 40);
            else buf_len = std::max(buf_len, std::min<int64_t>(W, (z.pos - z.buf_start) + (z.tracking ? need_track : z.acq_need)));
        }

        dabb_io io; memset(&io, 0, sizeof io);
        io.iq = reinterpret_cast<const float*>(iq.data()); io.iq_is_host = 1; io.stride_samples = W; io.buf_start = buf_start.data(); io.buf_len = buf_len;
        io.results = res.data(); io.fibs = fibs.data(); io.msc = msc.data(); io.msc_stride = 1152; io.sf = sf.data(); io.sf_stride = 5760; io.iq_format = fmt;
        if (dabb_process(ctx, &io) != DABB_OK) { fprintf(stderr, "batch_decode: %s\n", dabb_last_error(ctx)); return 1; }

        /* 2. per-stream results */
        for (int s = 0; s < S; s++) {
            Stream& z = st[(size_t)s];
            if (z.done) continue;
            const dabb_frame_result& r = res[(size_t)s];
            if (r.status == DABB_FRAME_NEED_SAMPLES) { z.done = true; active--; continue; }      /* cannot happen for a live stream: stop it rather than spin */
            z.pos = r.next_pos;
            if (r.status == DABB_FRAME_NO_SYNC) { z.tracking = false; continue; }
            if (r.status == DABB_FRAME_ACQUIRING) {
                z.tracking = false;
                if (z.acq_need < 5 * TF) z.acq_need += TF;
                else { z.pos += 2 * TF; z.acq_need = 3 * TF; dabb_stream_reset(ctx, s, 1, z.pos); }
                if (z.eof) { z.done = true; active--; }
                continue;
            }
            z.tracking = true; z.acq_need = 3 * TF; z.frames++;
            for (int f = 0; f < 12; f++) {
                const uint8_t* fb = fibs.data() + ((size_t)s * 12 + (size_t)f) * 32;
                const uint8_t ok = (r.fib_crc_mask >> f) & 1;
                fputc(ok, z.fic); fwrite(fb, 32, 1, z.fic);
                z.fibs++; z.fibs_ok += ok;
                if (ok) { dabb_host::FigEvents ev; z.db->parseFib(fb, ev); }
            }
            if (z.slot >= 0) {
                const int flen = 3 * z.bitrate, k = z.slot;
                for (int c = 4 - r.n_logical[k]; c < 4; c++) { fwrite(msc.data() + (((size_t)s * DABB_MAX_SUBCH + (size_t)k) * 4 + (size_t)c) * 1152, (size_t)flen, 1, z.msc); z.logical++; }
                for (int e = 0; e < r.n_rs_events[k]; e++) fprintf(z.rs, "%d %d\n", (r.rs_uncorr_mask[k] >> e) & 1, r.rs_corr[k][e]);
            } else {
                /* tune to the first audio service whose sub-channel organisation is known */
                for (const auto& c : z.db->components) {
                    if (c.transportMode() != TransportMode::Audio || c.audioType() == AudioServiceComponentType::Unknown) continue;
                    if (c.subchannelId < 0 || c.subchannelId >= 64 || !z.db->subch[(size_t)c.subchannelId].valid()) continue;
                    const Subchannel& u = z.db->subch[(size_t)c.subchannelId];
                    dabb_subchannel sc; memset(&sc, 0, sizeof sc);
                    sc.subch_id = u.subChId; sc.start_cu = u.startAddr; sc.length_cu = u.length; sc.bitrate = u.bitrate();
                    sc.short_form = u.protectionSettings.shortForm; sc.uep_level = u.protectionSettings.uepLevel;
                    sc.eep_profile_a = u.protectionSettings.eepProfile == EEPProtectionProfile::EEP_A; sc.eep_level = (int)u.protectionSettings.eepLevel;
                    sc.dabplus = c.audioType() == AudioServiceComponentType::DABPlus;
                    /* streams with the same code geometry share a slot */
                    int k = -1;
                    for (size_t q = 0; q < profiles.size(); q++) {
                        const dabb_subchannel& p = profiles[q].sc;
                        if (p.bitrate == sc.bitrate && p.short_form == sc.short_form && p.uep_level == sc.uep_level && p.eep_profile_a == sc.eep_profile_a && p.eep_level == sc.eep_level && p.dabplus == sc.dabplus) k = (int)q;
                    }
                    if (k < 0 && profiles.size() < DABB_MAX_SUBCH) { profiles.push_back(Profile{sc}); k = (int)profiles.size() - 1; }
                    if (k < 0) { fprintf(stderr, "batch_decode: %s: more than %d different protection profiles in one batch, not decoding its audio\n", z.name.c_str(), DABB_MAX_SUBCH); z.slot = -2; break; }
                    if (dabb_select_subchannel(ctx, s, 1, k, &sc) != DABB_OK) { fprintf(stderr, "batch_decode: %s: %s\n", z.name.c_str(), dabb_last_error(ctx)); z.slot = -2; break; }
                    z.slot = k; z.bitrate = sc.bitrate; z.sid = c.SId;
                    break;
                }
            }
        }
    }
    for (auto& z : st) {
        printf("%s frames=%ld fibs=%ld fib_crc_ok=%ld service=0x%X bitrate=%d logical_frames=%ld\n", z.name.c_str(), z.frames, z.fibs, z.fibs_ok, z.sid, z.bitrate, z.logical);
        fclose(z.in); fclose(z.fic); fclose(z.msc); fclose(z.rs);
    }
    dabb_destroy(ctx);
    return 0;
}
