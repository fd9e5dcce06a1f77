/*
 * mock_dab_b200.c — TEST DOUBLE of the C ABI (include/dab_b200.h) for exercising the host glue on a machine without a GPU.
 *
 * NOT part of the product and never built into the tree: tests/test_glue_mock.py compiles it into a temporary directory as
 * "libdab_b200.so" and runs welle.io_b200/glue_test with LD_LIBRARY_PATH pointing there.  It does no signal processing of its
 * own: at dabb_create it loads the IQ file named by $DABB_MOCK_IQ and lets the oracle (oracle/liboracle.so, test infrastructure)
 * decode the whole stream; dabb_process() then hands out one frame record after the other (FIBs, CRC mask, correctors, logical
 * frames) in the layout of the ABI.  RS statistics, superframes and the diagnostic taps are not served (the GPU tests cover them).
 * Only the entry points the glue calls are implemented.
 */
#include "../../include/dab_b200.h"
#include "../../oracle/dab_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct dabb_ctx {
    float* iq; long n;
    long frames; uint8_t* fibs; orc_frame_info_t* info;
    int k;                                   /* next frame to hand out */
    int sel_frame; int flen; uint8_t* msc; long n_msc;
    char err[128];
};

static long run_oracle(dabb_ctx* c, const orc_rx_cfg_t* cfg, uint8_t* fibs, long fib_cap, uint8_t* msc, long msc_cap, long* n_msc, orc_frame_info_t* info, long info_cap)
{
    orc_rx_t* rx = orc_rx_new(cfg);
    long nf = 0, nm = 0, nr = 0;
    int* rs = (int*)malloc(sizeof(int) * 2 * 4 * (size_t)info_cap);
    const long frames = orc_rx_run(rx, c->iq, c->n, fibs, fib_cap, &nf, msc, msc_cap, &nm, rs, 4 * info_cap, &nr, info, info_cap, NULL, 0);
    free(rs); orc_rx_free(rx);
    if (n_msc) *n_msc = nm;
    return frames;
}

int dabb_abi_version(void) { return DABB_ABI_VERSION; }
const char* dabb_last_error(const dabb_ctx* c) { return c ? c->err : "mock backend"; }

int dabb_create(const dabb_config* cfg, dabb_ctx** out)
{
    if (!cfg || !out || cfg->abi_version != DABB_ABI_VERSION) return DABB_E_ARG;
    const char* path = getenv("DABB_MOCK_IQ");
    if (!path) return DABB_E_NODEVICE;
    FILE* f = fopen(path, "rb");
    if (!f) return DABB_E_ARG;
    fseek(f, 0, SEEK_END); const long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    dabb_ctx* c = (dabb_ctx*)calloc(1, sizeof *c);
    c->n = bytes / 8; c->iq = (float*)malloc((size_t)bytes);
    if (fread(c->iq, 1, (size_t)bytes, f) != (size_t)bytes) { fclose(f); return DABB_E_ARG; }
    fclose(f);
    const long cap = c->n / ORC_TF + 2;
    c->fibs = (uint8_t*)calloc((size_t)cap * 12, 33); c->info = (orc_frame_info_t*)calloc((size_t)cap, sizeof(orc_frame_info_t));
    orc_rx_cfg_t oc; memset(&oc, 0, sizeof oc);
    oc.disable_coarse = cfg->disable_coarse; oc.fft_placement = cfg->fft_placement; oc.freqsync_method = cfg->freqsync_method;
    c->frames = run_oracle(c, &oc, c->fibs, cap * 12, NULL, 0, NULL, c->info, cap);
    c->sel_frame = -1;
    *out = c;
    return DABB_OK;
}
void dabb_destroy(dabb_ctx* c) { if (!c) return; free(c->iq); free(c->fibs); free(c->info); free(c->msc); free(c); }
int dabb_stream_reset(dabb_ctx* c, int32_t first, int32_t count, int64_t pos) { (void)first; (void)count; (void)pos; return c ? DABB_OK : DABB_E_ARG; }
int dabb_set_options(dabb_ctx* c, const dabb_options* o) { return c && o ? DABB_OK : DABB_E_ARG; }

int dabb_select_subchannel(dabb_ctx* c, int32_t first, int32_t count, int32_t slot, const dabb_subchannel* sc)
{
    (void)first; (void)count;
    if (!c || !sc || slot != 0) return DABB_E_ARG;
    orc_rx_cfg_t oc; memset(&oc, 0, sizeof oc);
    oc.disable_coarse = 1; oc.subch_start_cu = sc->start_cu; oc.subch_len_cu = sc->length_cu; oc.dabplus = sc->dabplus; oc.select_after_frames = c->k;
    const int rc = sc->short_form ? orc_prot_uep(sc->bitrate, sc->uep_level, &oc.prot) : orc_prot_eep(sc->bitrate, sc->eep_profile_a, sc->eep_level, &oc.prot);
    if (rc) { snprintf(c->err, sizeof c->err, "unsupported protection"); return DABB_E_UNSUPPORTED; }
    const long cap = c->n / ORC_TF + 2;
    free(c->msc);
    c->flen = 3 * sc->bitrate;
    c->msc = (uint8_t*)calloc((size_t)cap * 4, (size_t)c->flen);
    uint8_t* fibs = (uint8_t*)calloc((size_t)cap * 12, 33); orc_frame_info_t* info = (orc_frame_info_t*)calloc((size_t)cap, sizeof *info);
    run_oracle(c, &oc, fibs, cap * 12, c->msc, cap * 4 * c->flen, &c->n_msc, info, cap);
    free(fibs); free(info);
    c->sel_frame = c->k;
    return DABB_OK;
}
int dabb_remove_subchannel(dabb_ctx* c, int32_t first, int32_t count, int32_t slot) { (void)first; (void)count; (void)slot; if (!c) return DABB_E_ARG; c->sel_frame = -1; return DABB_OK; }

int dabb_process(dabb_ctx* c, const dabb_io* io)
{
    if (!c || !io || !io->results) return DABB_E_ARG;
    dabb_frame_result* r = io->results;
    memset(r, 0, sizeof *r);
    if (c->k >= c->frames) { r->status = DABB_FRAME_ACQUIRING; return DABB_OK; }
    const orc_frame_info_t* fi = &c->info[c->k];
    r->status = DABB_FRAME_DECODED; r->start_index = fi->start_index; r->fine_corr = fi->fine; r->coarse_corr = fi->coarse; r->snr_raw = fi->snr_raw;
    r->next_pos = fi->frame_pos + ORC_TU + fi->start_index + 75L * ORC_TS + ORC_TNULL;
    for (int f = 0; f < 12; f++) {
        const uint8_t* rec = c->fibs + 33 * ((size_t)c->k * 12 + f);
        if (rec[0]) r->fib_crc_mask |= 1 << f;
        if (io->fibs) memcpy(io->fibs + 32 * f, rec + 1, 32);
    }
    /* the time de-interleaver delivers its first logical frame 16 CIFs (4 frames) after the selection took effect */
    if (c->sel_frame >= 0 && c->k >= c->sel_frame + 4 && io->msc) {
        const long first = ((long)(c->k - c->sel_frame - 4)) * 4;
        int n = 0;
        for (int q = 0; q < 4; q++) if ((first + q + 1) * c->flen <= c->n_msc) n++;
        r->n_logical[0] = n;
        for (int q = 0; q < n; q++) memcpy(io->msc + (size_t)(4 - n + q) * io->msc_stride, c->msc + (size_t)(first + q) * c->flen, (size_t)c->flen);
    }
    c->k++;
    return DABB_OK;
}
int dabb_read_tap(dabb_ctx* c, int32_t what, void* out, size_t bytes) { (void)c; (void)what; (void)out; (void)bytes; return DABB_E_STATE; }
