// api.cu — the C ABI of include/dab_b200.h: context, per-stream receiver state, the per-frame pipeline and the
// stage-level entry points.  Host code here only sequences kernels on the context's stream; all signal processing
// is in ofdm.cu / viterbi.cu / rs.cu and the small state-machine kernels below.  There is no CPU fallback.
#include "../../include/dab_b200.h"
#include "common.cuh"
#include "viterbi.cuh"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

using namespace dabb;

namespace {

// per-stream receiver state (device resident), the variables OFDMProcessor carries between frames
// (ofdm-processor.h: coarseCorrector, fineCorrector, localPhase, sLevel) + FicHandler::fic_decode_success_ratio
struct StreamState {
    int64_t pos;            // logical sample index of the next sample to read
    int32_t coarse, fine;   // Hz; fine has int16 semantics
    int32_t local_phase;    // localPhase after the last sample read
    int32_t acq_failed;     // failed null searches since the last result record (acquire_kernel counts, advance_kernel hands over)
    int32_t acquired;       // 0: null search needed, 1: tracking
    int32_t started;        // 0: sLevel warm-up (T_F/2 samples) not done yet
    float slevel;
    int32_t start_index;    // last findIndex
    int64_t nframes;
};

struct StepScratch {        // per stream, rewritten every step
    int64_t win_start;      // T_u window for findIndex, relative to the stream buffer
    int64_t prs_start;
    int32_t nco_sync[2];
    int32_t nco_frame[4];
    int32_t active;         // this stream decodes a frame in this step
    int32_t status;
    // receiver state after this frame, filled by advance_kernel for the result record
    int32_t r_start_index, r_fine, r_coarse;
    float r_fx, r_fy, r_slevel;
    int64_t r_next_pos;
    int64_t null_pos;       // logical index of the first sample of the null symbol that follows this frame
    int32_t null_lp, null_ph;   // NCO phase before that sample, and the increment the null is read with
    int32_t r_acq_failed, pad1;
};

std::string g_create_error;

} // namespace

struct dabb_ctx {
    int device = 0; cudaStream_t stream = nullptr; cudaStream_t streamB = nullptr; cudaEvent_t evA[2] = {nullptr, nullptr}, evB[2] = {nullptr, nullptr}; bool evB_valid[2] = {false, false};
    int64_t step = 0; int last_parity = 0; int32_t* d_fic_ratio = nullptr; int32_t* d_coarse = nullptr; int ofdm_smem_floor = 0; int vit_stages_now = 3;
    int prio_b = 0; cudaStream_t streamP = nullptr; cudaEvent_t evPre = nullptr, evPost = nullptr;      // time-sync kernels of lane A at the highest priority (see dabb_create)
    cudaStream_t stream2 = nullptr; cudaEvent_t ev_ofdm = nullptr, ev_fic = nullptr; uint2* d_dec_fic = nullptr; int S = 0; int fft_mode = 0; int disable_coarse = 0; int keep_taps = 0; int placement = 0; int freqsync = 0;
    int decode_tii = 0; float2* d_tii = nullptr; bool persistent = false; unsigned int* d_work = nullptr; TraceBuf trace{nullptr, nullptr, 0}; std::string trace_path; int n_slots = 1; int max_cu = 144; int ring_pitch = 0; int flen_max = 0;
    std::string err; int64_t launches = 0; int osc_mismatches = -1; int osc_patched = 0; std::vector<float2> h_osc;
    HostTables* host = nullptr; DevTables dev{};
    std::vector<void*> allocs;
    StreamState* d_state = nullptr; StepScratch* d_scr = nullptr; MscSlotState* d_slots = nullptr;
    int64_t* d_buf_start = nullptr; int64_t* d_win = nullptr; int64_t* d_prs = nullptr; int32_t* d_nco_sync = nullptr; int32_t* d_nco_frame = nullptr;
    int32_t* d_active = nullptr; int32_t* d_index = nullptr; float* d_cir = nullptr; float* d_cir_work = nullptr; int search_generic = 0; int vit_split = 0; float2* d_r1 = nullptr; float2* d_null = nullptr; int32_t* d_snr = nullptr; float2* d_fc = nullptr; float* d_lvl = nullptr;
    int8_t* d_soft = nullptr; uint2* d_fic_steptab = nullptr; uint32_t* d_fic_stage_off = nullptr; uint2* d_dec = nullptr; size_t dec_bytes = 0; uint8_t* d_fibs = nullptr; int32_t* d_crc = nullptr;
    dabb_frame_result* d_results = nullptr;
    // per slot
    struct Slot {
        bool configured = false; ProtProfile prof{}; int nsteps = 0, nbits = 0, frag_pitch = 0, flen = 0;
        uint2* d_steptab = nullptr; uint32_t* d_stage_off = nullptr; uint32_t* d_prbs_words = nullptr; int8_t* d_frag = nullptr; int32_t* d_valid = nullptr;
        uint8_t* d_logical = nullptr; uint2* d_dec = nullptr; int8_t* d_ring = nullptr; uint8_t* d_window = nullptr; uint8_t* d_sf = nullptr; int32_t* d_info = nullptr;
    } slot[DABB_MAX_SUBCH];
    uint32_t* d_fic_prbs_words = nullptr;
    std::vector<MscSlotState> h_slots;
    float2* d_iq_stage = nullptr; size_t iq_stage_samples = 0; uint8_t* d_raw_stage = nullptr; size_t raw_stage_bytes = 0;
    // pipelined host-buffer path (dabb_submit / dabb_collect): copy stream, two device staging slots, two result slots in pinned memory
    cudaStream_t streamC = nullptr; cudaEvent_t evH2D[2] = {nullptr, nullptr}, evStageFree[2] = {nullptr, nullptr}; bool stageFreeValid[2] = {false, false};
    void* d_stage2[2] = {nullptr, nullptr}; size_t stage2_bytes[2] = {0, 0};
    int stage_last_fmt = -1; int64_t stage_last_len = 0; std::vector<int64_t> stage_last_start;     // the previous dabb_submit's window (carry_samples)
    struct Pending {
        bool used = false; cudaEvent_t done = nullptr; dabb_io io{};
        dabb_frame_result* h_res = nullptr; uint8_t* h_fibs = nullptr; uint8_t* h_msc = nullptr; uint8_t* h_sf = nullptr; size_t msc_bytes = 0, sf_bytes = 0;
        int flen[DABB_MAX_SUBCH] = {0, 0, 0, 0}; size_t msc_off[DABB_MAX_SUBCH] = {0, 0, 0, 0}, sf_off[DABB_MAX_SUBCH] = {0, 0, 0, 0};
    } pend[2];
    int pend_head = 0, pend_count = 0; int64_t submits = 0;
    // pinned host staging for results
    dabb_frame_result* h_results = nullptr; uint8_t* h_fibs = nullptr; uint8_t* h_msc = nullptr; uint8_t* h_sf = nullptr;
    int groups = 1; int fc_pitch = 1; int tail_frames = 0; int tail_groups = 5; int nco_fast = 0;
    const int32_t** d_info_tab = nullptr;
    // optional per-kernel timing: one event after every launch, durations = differences of consecutive events
    bool prof = false; std::vector<cudaEvent_t> prof_ev; std::vector<const char*> prof_name; size_t prof_used = 0;
    std::vector<std::pair<std::string, std::pair<double, long>>> prof_acc;
};

namespace {

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return DABB_E_CUDA; } } while (0)

template <typename T> int dalloc(dabb_ctx* ctx, T** p, size_t n, bool zero = true)
{
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T) > 0 ? n * sizeof(T) : 16);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return DABB_E_NOMEM; }
    if (zero) cudaMemsetAsync(q, 0, n * sizeof(T), ctx->stream);
    ctx->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

// frees a device allocation made through dalloc / tracked in ctx->allocs
template <typename T> void dfree(dabb_ctx* ctx, T*& p)
{
    if (!p) return;
    for (size_t i = 0; i < ctx->allocs.size(); i++) if (ctx->allocs[i] == (void*)p) { ctx->allocs.erase(ctx->allocs.begin() + i); break; }
    cudaFree((void*)p);
    p = nullptr;
}

void prof_mark(dabb_ctx* ctx, const char* what)
{
    if (!ctx->prof) return;
    if (ctx->prof_used == ctx->prof_ev.size()) {
        if (ctx->prof_ev.size() >= 16384) return;
        cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return;
        ctx->prof_ev.push_back(e); ctx->prof_name.push_back(what);
    }
    ctx->prof_name[ctx->prof_used] = what;
    cudaEventRecord(ctx->prof_ev[ctx->prof_used++], ctx->stream);
}

int check_launch(dabb_ctx* ctx, const char* what)
{
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = std::string(what) + ": " + cudaGetErrorString(e); return DABB_E_CUDA; }
    prof_mark(ctx, what);
    return 0;
}

void prof_collect(dabb_ctx* ctx)
{
    if (ctx->prof_used < 2) { ctx->prof_used = 0; return; }
    cudaStreamSynchronize(ctx->stream);
    for (size_t i = 1; i < ctx->prof_used; i++) {
        const char* nm = ctx->prof_name[i];
        if (!strcmp(nm, "__step_begin")) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof_ev[i - 1], ctx->prof_ev[i]) != cudaSuccess) continue;
        bool found = false;
        for (auto& a : ctx->prof_acc) if (a.first == nm) { a.second.first += ms; a.second.second++; found = true; break; }
        if (!found) ctx->prof_acc.push_back({nm, {ms, 1}});
    }
    ctx->prof_used = 0;
}

void pack_prbs_words(const uint8_t* bits, int nbits, std::vector<uint32_t>& w)
{
    // same packing as the decoder output: bit t lands in byte t/8 (MSB first), bytes little-endian in the word
    w.assign((nbits + 31) / 32, 0);
    for (int t = 0; t < nbits; t++) if (bits[t]) w[t >> 5] |= 1u << (8 * ((t >> 3) & 3) + 7 - (t & 7));
}

// ---------------------------------------------------------------------------------------------------------------
// state-machine kernels (one thread per stream)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int32_t mod_rate(int64_t v) { v %= INPUT_RATE; if (v < 0) v += INPUT_RATE; return (int32_t)v; }

// Initial / re-acquisition: OFDMProcessor::run up to SyncOnPhase (ofdm-processor.cpp:248-323): sLevel warm-up over
// T_F/2 samples, then the 50-sample envelope dip search.  Strictly sequential per stream, exact arithmetic order
// (double IIR), only runs while a stream is not tracking.
__global__ void acquire_kernel(StreamState* st, const float2* iq, int64_t stride, const int64_t* buf_start, int64_t buf_len, const float2* osc, int S)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    StreamState z = st[s];
    if (z.acquired) return;
    const float2* src = iq + (int64_t)s * stride;
    int64_t pos = z.pos - buf_start[s];            // physical index
    const int64_t end = buf_len;
    int32_t lp = z.local_phase; float sl = z.slevel;
    if (pos < 0) return;
    auto get = [&](int32_t phase, float& l1) -> bool {
        if (pos >= end) return false;
        float2 v = src[pos++];
        lp -= phase; lp = (lp + INPUT_RATE) % INPUT_RATE;
        const float2 o = osc[lp];
        const float re = __fsub_rn(__fmul_rn(v.x, o.x), __fmul_rn(v.y, o.y)), im = __fadd_rn(__fmul_rn(v.x, o.y), __fmul_rn(v.y, o.x));
        l1 = __fadd_rn(fabsf(re), fabsf(im));
        sl = (float)__dadd_rn(__dmul_rn(0.00001, (double)l1), __dmul_rn(1 - 0.00001, (double)sl));
        return true;
    };
    float l1;
    // Progress is kept: the state is written back at every point from which the search can be resumed exactly (after the warm-up and
    // at the start of every attempt), so a buffer that ends in the middle of an attempt costs only that attempt, and eight failed
    // attempts (no signal) end the call with the position advanced - the reference's notSynced loop keeps running the same way
    // (ofdm-processor.cpp:253-323).
    auto checkpoint = [&]() { z.started = 1; z.pos = pos + buf_start[s]; z.local_phase = lp; z.slevel = sl; st[s] = z; };
    if (!z.started) {
        sl = 0.f;
        for (int i = 0; i < TF / 2; i++) if (!get(0, l1)) return;   // not enough samples for the warm-up: state untouched
        checkpoint();
    }
    const int32_t phase = z.coarse + z.fine;
    // local ring of the last 64 envelope values is enough for the 50-tap moving sum
    float env[64];
    for (int attempt = 0; attempt < 8; attempt++) {
        checkpoint();
        int idx = 0; float cur = 0.f;
        for (int i = 0; i < 50; i++) { if (!get(0, l1)) return; env[idx & 63] = l1; cur = __fadd_rn(cur, l1); idx++; }
        int counter = 0; bool fail = false;
        while ((double)__fdiv_rn(cur, 50.f) > __dmul_rn(0.50, (double)sl)) {
            if (!get(phase, l1)) return;
            env[idx & 63] = l1; cur = __fadd_rn(cur, __fsub_rn(l1, env[(idx - 50) & 63])); idx++;
            if (++counter > TF) { fail = true; break; }
        }
        if (fail) { z.acq_failed++; continue; }
        counter = 0;
        while ((double)__fdiv_rn(cur, 50.f) < __dmul_rn(0.75, (double)sl)) {
            if (!get(phase, l1)) return;
            env[idx & 63] = l1; cur = __fadd_rn(cur, __fsub_rn(l1, env[(idx - 50) & 63])); idx++;
            if (++counter > TNULL + 50) { fail = true; break; }
        }
        if (fail) { z.acq_failed++; continue; }
        z.acquired = 1;
        break;
    }
    checkpoint();      // acquired: tracking starts here; not acquired: the next call continues the search from here
}

__global__ void plan_kernel(StreamState* st, StepScratch* scr, const int64_t* buf_start, int64_t buf_len, int S,
                            int64_t* win, int32_t* nco_sync, int32_t* active)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const StreamState z = st[s];
    StepScratch c{};
    const int64_t rel = z.pos - buf_start[s];
    const int64_t need = (int64_t)TU + (TU - 1) + 75 * (int64_t)TS + TNULL;
    if (!z.acquired) { c.active = 0; c.status = DABB_FRAME_ACQUIRING; }
    else if (rel < 0 || rel + need > buf_len) { c.active = 0; c.status = DABB_FRAME_NEED_SAMPLES; }
    else {
        c.active = 1; c.status = DABB_FRAME_DECODED; c.win_start = rel;
        const int32_t p1 = z.coarse + z.fine;
        c.nco_sync[0] = mod_rate((int64_t)z.local_phase - p1); c.nco_sync[1] = p1;
    }
    scr[s] = c;
    win[s] = c.win_start; nco_sync[2 * s] = c.nco_sync[0]; nco_sync[2 * s + 1] = c.nco_sync[1]; active[s] = c.active;
}

__global__ void post_sync_kernel(StreamState* st, StepScratch* scr, const int32_t* index, const int32_t* coarse_corr, int S, int64_t* prs, int32_t* nco_frame, int32_t* active)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    StepScratch c = scr[s];
    if (!c.active) return;
    StreamState z = st[s];
    const int32_t p1 = z.coarse + z.fine;
    const int idx = index[s];
    z.start_index = idx;
    if (idx < 0) {
        // SyncOnPhase failed: the T_u samples are consumed and the null search starts over (ofdm-processor.cpp:347-350)
        z.pos += TU; z.local_phase = mod_rate((int64_t)z.local_phase - (int64_t)TU * p1); z.acquired = 0;
        c.active = 0; c.status = DABB_FRAME_NO_SYNC;
        st[s] = z; scr[s] = c; active[s] = 0;
        return;
    }
    c.prs_start = c.win_start + idx;
    // phase applied to PRS sample 0 (the (idx+1)-th sample read in this frame)
    const int32_t lp_prs0 = mod_rate((int64_t)z.local_phase - (int64_t)(idx + 1) * p1);
    // coarse corrector (ofdm-processor.cpp:397-409): find_index_kernel evaluated processPRS where the FIC ratio asked for it
    if (coarse_corr) {
        const int corr = coarse_corr[s];
        if (corr != 100) { z.coarse += corr * 1000; if (abs(z.coarse) > 35000) z.coarse = 0; }     // 100: not evaluated / no estimate (:404-408)
    }
    const int32_t p2 = z.coarse + z.fine;
    // phase applied to the first data-symbol sample = lp after 2048+idx samples, minus p2; expressed at index 2048
    const int32_t lp_after_prs = mod_rate((int64_t)z.local_phase - (int64_t)(TU + idx) * p1);
    const int32_t lp_sym0 = mod_rate((int64_t)lp_after_prs - p2 + (int64_t)TU * p2);   // so that lp(i) = lp_sym0 - i*p2 for i >= 2048
    c.nco_frame[0] = lp_prs0; c.nco_frame[1] = p1; c.nco_frame[2] = lp_sym0; c.nco_frame[3] = p2;
    scr[s] = c; st[s] = z;
    prs[s] = c.prs_start;
    nco_frame[4 * s] = lp_prs0; nco_frame[4 * s + 1] = p1; nco_frame[4 * s + 2] = lp_sym0; nco_frame[4 * s + 3] = p2;
}

// lane A, right after the OFDM kernel: the part of OFDMProcessor::run that feeds the next frame's synchronisation
// (fine corrector update :450-451, consumed samples, NCO phase, corrector wrap :478-486)
__global__ void advance_kernel(StreamState* st, StepScratch* scr, int S, int groups /* partial sums per frame */, const float2* fc_part, const float* lvl_part)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    StepScratch c = scr[s];
    StreamState z = st[s];
    c.r_start_index = z.start_index; c.r_fx = 0.f; c.r_fy = 0.f;
    c.r_acq_failed = z.acq_failed;
    if (z.acq_failed) { z.acq_failed = 0; if (!c.active) st[s].acq_failed = 0; }
    if (c.active) {
        float fx = 0.f, fy = 0.f;      // FreqCorr: partial sums in group order
        for (int g = 0; g < groups; g++) { fx += fc_part[(int64_t)s * groups + g].x; fy += fc_part[(int64_t)s * groups + g].y; }
        const int idx = z.start_index;
        const int32_t p1 = c.nco_frame[1], p2 = c.nco_frame[3];
        // fineCorrector += 0.1 * arg(FreqCorr) / M_PI * (carrierDiff / 2), int16 store (ofdm-processor.cpp:450-451)
        const double upd = 0.1 * (double)(float)atan2((double)fy, (double)fx) / 3.14159265358979323846 * 500;
        z.fine = (int32_t)(int16_t)((double)z.fine + upd);
        const int32_t p3 = z.coarse + z.fine;
        const int64_t lpn = (int64_t)z.local_phase - (int64_t)(TU + idx) * p1 - (int64_t)75 * TS * p2;
        c.null_pos = z.pos + (int64_t)TU + idx + 75 * (int64_t)TS; c.null_lp = mod_rate(lpn); c.null_ph = p3;
        const int64_t lp = lpn - (int64_t)TNULL * p3;
        z.local_phase = mod_rate(lp);
        z.pos += (int64_t)TU + idx + 75 * (int64_t)TS + TNULL;
        if (z.fine > 500) { z.coarse += 1000; z.fine -= 1000; }
        else if (z.fine < -500) { z.coarse -= 1000; z.fine += 1000; }
        // sLevel (ofdm-processor.cpp:166,215) is an IIR over EVERY sample read: exact while a stream searches (acquire_kernel), and
        // while it tracks it follows the same recursion on a sub-sample (one magnitude per symbol and thread of the OFDM kernel,
        // decayed per symbol) - only a re-acquisition after a sync loss ever reads it
        if (lvl_part) {
            float est = 0.f;
            for (int g = 0; g < groups; g++) est += lvl_part[(int64_t)s * groups + g];
            const float ln1ma = -1.000005e-5f;          // ln(1 - 1e-5)
            const float n_frame = (float)(TU + idx + 75 * TS + TNULL);
            z.slevel = z.slevel * expf(n_frame * ln1ma) + est * (1.0f - expf((float)TS * ln1ma)) * expf((float)TNULL * ln1ma);
        }
        z.nframes++;
        c.r_fx = fx; c.r_fy = fy;
        st[s] = z;
    }
    c.r_fine = z.fine; c.r_coarse = z.coarse; c.r_next_pos = z.pos; c.r_slevel = z.slevel;
    scr[s] = c;
}

// diagnostics tap: the null symbol as OFDMProcessor::run hands it to onNewNullSymbol (ofdm-processor.cpp:462-469): T_null
// samples read with the NCO at coarse + fine (after this frame's fine update)
__global__ void null_tap_kernel(const StepScratch* scr, const float2* iq, int64_t stride, const int64_t* buf_start, int64_t buf_len, const float2* osc, float2* out)
{
    const int s = blockIdx.x;
    const StepScratch c = scr[s];
    float2* dst = out + (int64_t)s * TNULL;
    const int64_t p0 = c.null_pos - buf_start[s];
    const bool mix = c.null_lp != 0 || c.null_ph != 0;
    for (int i = threadIdx.x; i < TNULL; i += blockDim.x) {
        float2 v = make_float2(0.f, 0.f);
        if (c.active && p0 >= 0 && p0 + i < buf_len) {
            v = iq[(int64_t)s * stride + p0 + i];
            if (mix) {
                const float2 o = osc[mod_rate((int64_t)c.null_lp - (int64_t)(i + 1) * c.null_ph)];
                v = make_float2(__fsub_rn(__fmul_rn(v.x, o.x), __fmul_rn(v.y, o.y)), __fadd_rn(__fmul_rn(v.x, o.y), __fmul_rn(v.y, o.x)));
            }
        }
        dst[i] = v;
    }
}

// lane B, after FIC / MSC / RS of the frame: decoder-side state and the result record
__global__ void finalize_kernel(const StepScratch* scr, MscSlotState* slots, int n_slots, int S, int32_t* fic_ratio,
                                const int32_t* snr, const int32_t* crc, const int32_t* const* slot_info, dabb_frame_result* res)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const StepScratch c = scr[s];
    dabb_frame_result r;
    memset(&r, 0, sizeof r);
    r.status = c.status;
    int ratio = fic_ratio[s];
    if (c.active) {
        // FIC success counter, one saturating update per FIB in order (fic-handler.cpp:214-229)
        const int mask = crc[s];
        for (int f = 0; f < 12; f++) { if ((mask >> f) & 1) { if (ratio < 10) ratio++; } else if (ratio > 0) ratio--; }
        fic_ratio[s] = ratio;
        r.snr_raw = snr[s]; r.fib_crc_mask = mask;
        for (int k = 0; k < n_slots; k++) {
            MscSlotState& m = slots[s * n_slots + k];
            if (!m.enabled) continue;
            m.cif_count += 4;
            const int32_t* inf = slot_info[k] + (int64_t)s * 16;
            r.n_logical[k] = inf[0]; r.n_rs_events[k] = inf[1]; r.rs_uncorr_mask[k] = inf[2];
            for (int e = 0; e < 4; e++) r.rs_corr[k][e] = inf[3 + e];
            r.sf_ready[k] = inf[7]; r.sf_au_count[k] = inf[8]; r.sf_au_crc_mask[k] = inf[9];
        }
    }
    r.start_index = c.r_start_index; r.freq_corr_re = c.r_fx; r.freq_corr_im = c.r_fy;
    r.fine_corr = c.r_fine; r.coarse_corr = c.r_coarse; r.fic_ratio = ratio; r.next_pos = c.r_next_pos; r.slevel = c.r_slevel; r.acq_failed = c.r_acq_failed;
    res[s] = r;
}

__global__ void reset_kernel(StreamState* st, MscSlotState* slots, int n_slots, int first, int count, int64_t pos, int32_t* fic_ratio)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    StreamState z{}; z.pos = pos; z.start_index = -1;
    st[first + i] = z; fic_ratio[first + i] = 0;
    for (int k = 0; k < n_slots; k++) { MscSlotState& m = slots[(first + i) * n_slots + k]; m.cif_count = 0; m.sf_frame_count = 0; }
}

__global__ void set_slot_kernel(MscSlotState* slots, int n_slots, int slot, int first, int count, MscSlotState v)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    slots[(first + i) * n_slots + slot] = v;
}

void sync_all(dabb_ctx* ctx)
{
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->streamP) cudaStreamSynchronize(ctx->streamP);
    if (ctx->streamB) cudaStreamSynchronize(ctx->streamB);
    if (ctx->stream2) cudaStreamSynchronize(ctx->stream2);
}

// raw file formats -> cf32, value for value like CRAWFile::convertSamples (input/raw_file.cpp:336-363): u8 (b-128)/128,
// s8 b/128, "s16le"/"s16be" unscaled with the byte order the reference applies to each name
__global__ void convert_iq_kernel(const uint8_t* __restrict__ raw, int64_t raw_stride_samples, int fmt, float2* __restrict__ out, int64_t out_stride, int64_t len, int S)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= len || s >= S) return;
    float re, im;
    if (fmt == DABB_IQ_U8) { const uchar2 b = reinterpret_cast<const uchar2*>(raw)[(int64_t)s * raw_stride_samples + i]; re = (float)((int)b.x - 128) / 128.0f; im = (float)((int)b.y - 128) / 128.0f; }
    else if (fmt == DABB_IQ_S8) { const char2 b = reinterpret_cast<const char2*>(raw)[(int64_t)s * raw_stride_samples + i]; re = (float)b.x / 128.0f; im = (float)b.y / 128.0f; }
    else {
        const uchar4 b = reinterpret_cast<const uchar4*>(raw)[(int64_t)s * raw_stride_samples + i];
        int16_t I, Q;
        if (fmt == DABB_IQ_S16LE) { I = (int16_t)((b.x << 8) | b.y); Q = (int16_t)((b.z << 8) | b.w); }
        else { I = (int16_t)((b.y << 8) | b.x); Q = (int16_t)((b.w << 8) | b.z); }
        re = (float)I; im = (float)Q;
    }
    out[(int64_t)s * out_stride + i] = make_float2(re, im);
}

int ensure_dec(dabb_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->dec_bytes) return 0;
    if (ctx->d_dec) { cudaStreamSynchronize(ctx->stream); dfree(ctx, ctx->d_dec); ctx->dec_bytes = 0; }
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc(decisions): ") + cudaGetErrorString(e); return DABB_E_NOMEM; }
    ctx->allocs.push_back(q);
    ctx->d_dec = (uint2*)q; ctx->dec_bytes = bytes;
    return 0;
}

} // namespace

// =====================================================================================================================
extern "C" {

int dabb_abi_version(void) { return DABB_ABI_VERSION; }

const char* dabb_last_error(const dabb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dabb_create(const dabb_config* cfg, dabb_ctx** out)
{
    if (!cfg || !out) { g_create_error = "null argument"; return DABB_E_ARG; }
    if (cfg->abi_version != DABB_ABI_VERSION) { g_create_error = "ABI version mismatch"; return DABB_E_ARG; }
    if (cfg->transmission_mode != 1 && cfg->transmission_mode != 0) { g_create_error = "only transmission mode I is supported"; return DABB_E_UNSUPPORTED; }
    if (cfg->n_streams < 1) { g_create_error = "n_streams < 1"; return DABB_E_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_create_error = "no CUDA device: this library has no CPU fallback"; return DABB_E_NODEVICE; }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "bad device ordinal"; return DABB_E_ARG; }
    dabb_ctx* ctx = new dabb_ctx();
    ctx->device = cfg->device; ctx->S = cfg->n_streams; ctx->fft_mode = cfg->fft_mode; ctx->disable_coarse = cfg->disable_coarse; ctx->keep_taps = cfg->keep_taps; ctx->placement = cfg->fft_placement; ctx->freqsync = cfg->freqsync_method;
    ctx->n_slots = cfg->n_subch_slots > 0 ? (cfg->n_subch_slots > DABB_MAX_SUBCH ? DABB_MAX_SUBCH : cfg->n_subch_slots) : 1;
    ctx->max_cu = cfg->max_subch_cu > 0 ? cfg->max_subch_cu : 144;
    ctx->groups = cfg->ofdm_groups > 0 ? cfg->ofdm_groups : (ctx->S >= 1024 ? 1 : (ctx->S >= 64 ? 5 : 25));
    if (75 % ctx->groups) ctx->groups = 1;
    ctx->nco_fast = cfg->nco_mode == DABB_NCO_FAST;
    auto fail = [&](int code) { g_create_error = ctx->err; dabb_destroy(ctx); return code; };
    if (cudaSetDevice(ctx->device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return fail(DABB_E_CUDA); }
    // Stream priorities.  Lane A (time sync + OFDM of frame n+1) and lane B (FIC / MSC Viterbi + RS of frame n) become runnable at the
    // same moment; lane B gets the higher priority and runs first, in a burst (its CTAs take every slot the OFDM launch frees).
    // DABB_LANES=partition is the measured alternative (CTA timelines with DABB_TRACE): lane A higher, the OFDM kernel
    // persistent and capped at four CTAs per SM by a 45 KB shared-memory request, so that one two-buffer Viterbi CTA (16 384 registers,
    // 40 KB) runs beside them for the whole launch.  The two kernels then do share every SM - and slow each other down by as much as
    // they overlap (OFDM 3.2 -> 4.3 ms with the Viterbi CTAs beside it; step 4.94 vs 4.70 ms), so the burst order stays the default.
    const char* lanes_env = getenv("DABB_LANES");
    const bool partition = lanes_env ? !strcmp(lanes_env, "partition") : false;
    int prio_lo = 0, prio_hi = 0; cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, partition ? prio_hi : prio_lo) != cudaSuccess) { ctx->err = "cudaStreamCreate failed"; return fail(DABB_E_CUDA); }
    if (ofdm_init_constants() != 0) { ctx->err = "cudaMemcpyToSymbol(ofdm constants) failed"; return fail(DABB_E_CUDA); }
    // lane B (FIC/MSC/RS of frame n) runs on its own stream so that it overlaps lane A (time sync + OFDM of frame n+1)
    {
        int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lane B gets the higher priority: its CTAs take the SM resources lane A leaves free
        // Three levels where the device has them: the short time-sync kernels that precede the OFDM kernel (plan, findIndex, post_sync:
        // 0.2 ms) > lane B > the OFDM kernel.  With only two levels the block scheduler serves lane B's grids first and the time sync of
        // frame n+1 crawls in the slots the Viterbi CTAs leave - the CTA timelines showed 0.6 ms per step with Viterbi CTAs alone on the
        // GPU because the OFDM launch behind that time sync had not started.  Measured: no gain (the Viterbi kernel runs at
        // full speed in that phase, and beside OFDM CTAs it does not), so it is an experiment switch: DABB_PREFIX_LANE=1.
        const bool prefix = !partition && getenv("DABB_PREFIX_LANE") && atoi(getenv("DABB_PREFIX_LANE")) == 1 && hi < lo - 1;
        const int prio_b = prefix ? hi + 1 : hi;
        if (prefix) {
            if (cudaStreamCreateWithPriority(&ctx->streamP, cudaStreamNonBlocking, hi) != cudaSuccess ||
                cudaEventCreateWithFlags(&ctx->evPre, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->evPost, cudaEventDisableTiming) != cudaSuccess) { ctx->err = "cudaStreamCreate failed"; return fail(DABB_E_CUDA); }
        }
        ctx->prio_b = prio_b;
        if (cudaStreamCreateWithPriority(&ctx->streamB, cudaStreamNonBlocking, partition ? lo : prio_b) != cudaSuccess) { ctx->err = "cudaStreamCreate failed"; return fail(DABB_E_CUDA); }
        // co-residency experiment: 46 KB per OFDM CTA -> four of them per SM (49 152 registers, 188 KB), leaving exactly the 16 384 registers
        // and 40 KB one two-stage Viterbi CTA needs, so that the integer ACS work runs in the issue slots the shared-memory-bound OFDM
        // kernel leaves free instead of taking turns with it
        ctx->ofdm_smem_floor = (getenv("DABB_CORESIDENT") || partition) ? 45 * 1024 : 0;
        ctx->persistent = partition;
    }
    for (int i = 0; i < 2; i++) if (cudaEventCreateWithFlags(&ctx->evA[i], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&ctx->evB[i], cudaEventDisableTiming) != cudaSuccess) { ctx->err = "event creation failed"; return fail(DABB_E_CUDA); }
    // second stream: the FIC chain (de-puncture, Viterbi, CRC) overlaps the MSC chain; both only depend on the OFDM kernel
    { int lo = 0, hi = 0; cudaDeviceGetStreamPriorityRange(&lo, &hi); if (cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, partition ? lo : ctx->prio_b) != cudaSuccess) { ctx->err = "cudaStreamCreate failed"; return fail(DABB_E_CUDA); } }
    if ( cudaEventCreateWithFlags(&ctx->ev_ofdm, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fic, cudaEventDisableTiming) != cudaSuccess) { ctx->err = "stream/event creation failed"; return fail(DABB_E_CUDA); }
    ctx->host = new HostTables();
    build_host_tables(*ctx->host);
    const int S = ctx->S;
    int rc = 0;
    ctx->tail_frames = (cfg->ofdm_tail_split == 0 && ctx->groups == 1) ? ofdm_tail_frames(S) : 0;
    if (ctx->tail_frames && getenv("DABB_TAIL_FRAMES")) { const int v = atoi(getenv("DABB_TAIL_FRAMES")); if (v >= 0 && v <= S) ctx->tail_frames = v; }   // experiments
    ctx->tail_groups = 5;            // 16-symbol CTAs: measured best of 5 / 15 / 25 (3.207 -> 3.199 ms per 8192 frames)
    if (getenv("DABB_TAIL_GROUPS")) { const int v = atoi(getenv("DABB_TAIL_GROUPS")); if (v > 0 && 75 % v == 0) ctx->tail_groups = v; }
    ctx->fc_pitch = ctx->tail_frames ? ctx->tail_groups : ctx->groups;
    // tables
    float2 *tf, *ti, *pr, *osc; int16_t *ip, *fm; uint8_t *ge, *gl, *pb;
    if ((rc = dalloc(ctx, &tf, TwLayout::TOTAL)) || (rc = dalloc(ctx, &ti, TwLayout::TOTAL)) || (rc = dalloc(ctx, &pr, TU)) || (rc = dalloc(ctx, &osc, INPUT_RATE, false)) ||
        (rc = dalloc(ctx, &ip, sizeof ctx->host->invperm / sizeof(int16_t))) || (rc = dalloc(ctx, &fm, 3096)) || (rc = dalloc(ctx, &ge, 512)) || (rc = dalloc(ctx, &gl, 256)) || (rc = dalloc(ctx, &pb, sizeof ctx->host->prbs)))
        return fail(rc);
    {
        ctx->h_osc.resize(INPUT_RATE);
        build_osc_table(ctx->h_osc.data());
        cudaMemcpyAsync(osc, ctx->h_osc.data(), sizeof(float2) * INPUT_RATE, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(tf, ctx->host->tw_fwd, sizeof(float2) * TwLayout::TOTAL, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(ti, ctx->host->tw_inv, sizeof(float2) * TwLayout::TOTAL, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(pr, ctx->host->prs_ref, sizeof(float2) * TU, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(ip, ctx->host->invperm, sizeof ctx->host->invperm, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(fm, ctx->host->fic_map, sizeof(int16_t) * 3096, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(ge, ctx->host->gf_exp, 512, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(gl, ctx->host->gf_log, 256, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(pb, ctx->host->prbs, sizeof ctx->host->prbs, cudaMemcpyHostToDevice, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = "table upload failed"; return fail(DABB_E_CUDA); }
    }
    ctx->dev.tw_fwd = tf; ctx->dev.tw_inv = ti; ctx->dev.prs_ref = pr; ctx->dev.osc = osc; ctx->dev.invperm = ip; ctx->dev.fic_map = fm;
    ctx->dev.gf_exp = ge; ctx->dev.gf_log = gl; ctx->dev.prbs = pb;
    {
        // on-the-fly oscillator: upload the factors, compare against the table for every index on the device, adopt it only when
        // nothing differs; DABB_OSC_TABLE=1 in the environment forces the table lookups (A/B measurements)
        double2* hi;
        if ((rc = dalloc(ctx, &hi, 2000))) return fail(rc);
        std::vector<double2> h_hi(2000);
        build_osc_factors(ctx->h_osc.data(), h_hi.data(), &ctx->dev.osc_theta, &ctx->osc_patched);
        cudaMemcpyAsync(hi, h_hi.data(), sizeof(double2) * 2000, cudaMemcpyHostToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        ctx->dev.osc_hi = hi; ctx->dev.osc_mode = 0;
        ctx->osc_mismatches = launch_osc_verify(ctx->dev, ctx->stream);
        if (ctx->osc_mismatches == 0 && !getenv("DABB_OSC_TABLE")) ctx->dev.osc_mode = 1;
        std::vector<float2>().swap(ctx->h_osc);
    }
    {
        std::vector<uint32_t> w; pack_prbs_words(ctx->host->prbs, 768, w);
        if ((rc = dalloc(ctx, &ctx->d_fic_prbs_words, w.size()))) return fail(rc);
        cudaMemcpy(ctx->d_fic_prbs_words, w.data(), w.size() * 4, cudaMemcpyHostToDevice);
        // expansion tables of the decoder kernel for the FIC puncturing (fic-handler.cpp:144-191)
        std::vector<uint2> steps; std::vector<uint32_t> soff;
        build_vit_tables_u2(ctx->host->fic_map, 774, steps, soff);
        if ((rc = dalloc(ctx, &ctx->d_fic_steptab, steps.size())) || (rc = dalloc(ctx, &ctx->d_fic_stage_off, soff.size()))) return fail(rc);
        cudaMemcpy(ctx->d_fic_steptab, steps.data(), steps.size() * sizeof(uint2), cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_fic_stage_off, soff.data(), soff.size() * 4, cudaMemcpyHostToDevice);
    }
    ctx->vit_split = getenv("DABB_VIT_SPLIT") ? atoi(getenv("DABB_VIT_SPLIT")) : 0;       // traceback as its own launch (A/B)
    ctx->search_generic = getenv("DABB_SEARCH_GENERIC") != nullptr;     // A/B: ThresholdBeforePeak through the literal sliding maximum
    if (getenv("DABB_TRACE")) {
        ctx->trace_path = getenv("DABB_TRACE"); ctx->trace.cap = 1u << 20;
        if ((rc = dalloc(ctx, &ctx->trace.rec, (size_t)3 * ctx->trace.cap)) || (rc = dalloc(ctx, &ctx->trace.count, 1))) return fail(rc);
    }
    ctx->ring_pitch = ctx->max_cu * 64;
    if ((rc = dalloc(ctx, &ctx->d_state, S)) || (rc = dalloc(ctx, &ctx->d_scr, 2 * (size_t)S)) || (rc = dalloc(ctx, &ctx->d_fic_ratio, S)) || (rc = dalloc(ctx, &ctx->d_coarse, S)) || (rc = dalloc(ctx, &ctx->d_slots, (size_t)S * ctx->n_slots)) ||
        (rc = dalloc(ctx, &ctx->d_buf_start, S)) || (rc = dalloc(ctx, &ctx->d_win, 2 * (size_t)S)) || (rc = dalloc(ctx, &ctx->d_prs, 2 * (size_t)S)) ||
        (rc = dalloc(ctx, &ctx->d_nco_sync, 4 * (size_t)S)) || (rc = dalloc(ctx, &ctx->d_nco_frame, 8 * (size_t)S)) || (rc = dalloc(ctx, &ctx->d_active, 2 * (size_t)S)) ||
        (rc = dalloc(ctx, &ctx->d_index, 2 * (size_t)S)) || (rc = dalloc(ctx, &ctx->d_snr, 2 * (size_t)S)) || (rc = dalloc(ctx, &ctx->d_fc, 2 * (size_t)S * ctx->fc_pitch)) || (rc = dalloc(ctx, &ctx->d_lvl, 2 * (size_t)S * ctx->fc_pitch)) ||
        (rc = dalloc(ctx, &ctx->d_soft, 2 * (size_t)S * DABB_SOFT_PER_FRAME + VIT_FRAG_SLACK, false)) ||
        (rc = dalloc(ctx, &ctx->d_fibs, (size_t)S * 12 * 32)) || (rc = dalloc(ctx, &ctx->d_crc, S)) || (rc = dalloc(ctx, &ctx->d_results, S)) ||
        (rc = dalloc(ctx, &ctx->d_info_tab, DABB_MAX_SUBCH)) || (rc = dalloc(ctx, &ctx->d_cir_work, (size_t)S * TU, false)))
        return fail(rc);
    if (ctx->keep_taps && ((rc = dalloc(ctx, &ctx->d_cir, (size_t)S * TU)) || (rc = dalloc(ctx, &ctx->d_r1, (size_t)S * 75 * 1536)) || (rc = dalloc(ctx, &ctx->d_null, (size_t)S * TNULL)))) return fail(rc);
    if ((rc = ensure_dec(ctx, vit_dec_bytes(S * 4, 774)))) return fail(rc);
    { void* q = nullptr; if (cudaMalloc(&q, vit_dec_bytes(S * 4, 774)) != cudaSuccess) { ctx->err = "cudaMalloc(FIC decisions)"; return fail(DABB_E_NOMEM); } ctx->allocs.push_back(q); ctx->d_dec_fic = (uint2*)q; }
    ctx->h_slots.assign((size_t)S * ctx->n_slots, MscSlotState{});
    // info arrays of unconfigured slots are never read (finalize skips disabled slots)
    cudaHostAlloc((void**)&ctx->h_results, sizeof(dabb_frame_result) * S, cudaHostAllocDefault);
    cudaHostAlloc((void**)&ctx->h_fibs, (size_t)S * 12 * 32, cudaHostAllocDefault);
    {
        reset_kernel<<<(S + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_state, ctx->d_slots, ctx->n_slots, 0, S, 0, ctx->d_fic_ratio);
        ctx->launches++;
    }
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { ctx->err = std::string("init failed: ") + cudaGetErrorString(cudaGetLastError()); return fail(DABB_E_CUDA); }
    *out = ctx;
    return DABB_OK;
}

void dabb_destroy(dabb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    if (ctx->trace.rec && !ctx->trace_path.empty()) {       // CTA timeline: kind, sm, start ns, end ns per line
        unsigned int n = 0; cudaMemcpy(&n, ctx->trace.count, 4, cudaMemcpyDeviceToHost);
        if (n > ctx->trace.cap) n = ctx->trace.cap;
        std::vector<unsigned long long> h((size_t)3 * n);
        cudaMemcpy(h.data(), ctx->trace.rec, h.size() * 8, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(ctx->trace_path.c_str(), "w")) { for (unsigned int i = 0; i < n; i++) fprintf(f, "%llu %llu %llu %llu\n", h[3 * i] >> 32, h[3 * i] & 0xFFFFFFFFull, h[3 * i + 1], h[3 * i + 2]); fclose(f); }
    }
    for (void* p : ctx->allocs) cudaFree(p);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    if (ctx->d_iq_stage) cudaFree(ctx->d_iq_stage);
    if (ctx->d_raw_stage) cudaFree(ctx->d_raw_stage);
    for (int i = 0; i < 2; i++) {
        if (ctx->d_stage2[i]) cudaFree(ctx->d_stage2[i]);
        if (ctx->evH2D[i]) cudaEventDestroy(ctx->evH2D[i]);
        if (ctx->evStageFree[i]) cudaEventDestroy(ctx->evStageFree[i]);
        auto& p = ctx->pend[i];
        if (p.done) cudaEventDestroy(p.done);
        if (p.h_res) cudaFreeHost(p.h_res); if (p.h_fibs) cudaFreeHost(p.h_fibs); if (p.h_msc) cudaFreeHost(p.h_msc); if (p.h_sf) cudaFreeHost(p.h_sf);
    }
    if (ctx->streamC) cudaStreamDestroy(ctx->streamC);
    if (ctx->h_results) cudaFreeHost(ctx->h_results);
    if (ctx->h_fibs) cudaFreeHost(ctx->h_fibs);
    if (ctx->h_msc) cudaFreeHost(ctx->h_msc);
    if (ctx->h_sf) cudaFreeHost(ctx->h_sf);
    if (ctx->ev_ofdm) cudaEventDestroy(ctx->ev_ofdm);
    if (ctx->ev_fic) cudaEventDestroy(ctx->ev_fic);
    for (int i = 0; i < 2; i++) { if (ctx->evA[i]) cudaEventDestroy(ctx->evA[i]); if (ctx->evB[i]) cudaEventDestroy(ctx->evB[i]); }
    if (ctx->streamP) cudaStreamDestroy(ctx->streamP);
    if (ctx->evPre) cudaEventDestroy(ctx->evPre);
    if (ctx->evPost) cudaEventDestroy(ctx->evPost);
    if (ctx->streamB) cudaStreamDestroy(ctx->streamB);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx->host;
    delete ctx;
}

void* dabb_cuda_stream(dabb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t dabb_kernel_launches(const dabb_ctx* ctx) { return ctx ? ctx->launches : 0; }
int dabb_sync(dabb_ctx* ctx) { if (!ctx) return DABB_E_ARG; cudaSetDevice(ctx->device); CK(cudaStreamSynchronize(ctx->stream)); CK(cudaStreamSynchronize(ctx->streamB)); CK(cudaStreamSynchronize(ctx->stream2)); if (ctx->prof) prof_collect(ctx); return 0; }

int dabb_join_lanes(dabb_ctx* ctx)
{
    if (!ctx) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->prof || !ctx->evB_valid[ctx->last_parity]) return DABB_OK;       // serial mode: everything is on the main stream already
    CK(cudaStreamWaitEvent(ctx->stream, ctx->evB[ctx->last_parity], 0));       // lane B's last record follows the FIC stream's join
    return DABB_OK;
}

int dabb_stream_reset(dabb_ctx* ctx, int32_t first, int32_t count, int64_t pos)
{
    if (!ctx || first < 0 || count < 0 || first + count > ctx->S) return DABB_E_ARG;
    if (!count) return 0;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    reset_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_state, ctx->d_slots, ctx->n_slots, first, count, pos, ctx->d_fic_ratio);
    for (int i = first; i < first + count; i++) for (int k = 0; k < ctx->n_slots; k++) { ctx->h_slots[(size_t)i * ctx->n_slots + k].cif_count = 0; ctx->h_slots[(size_t)i * ctx->n_slots + k].sf_frame_count = 0; }
    { int rc_ = check_launch(ctx, "reset_kernel"); cudaStreamSynchronize(ctx->stream); ctx->evB_valid[0] = ctx->evB_valid[1] = false; return rc_; }
}

int dabb_select_subchannel(dabb_ctx* ctx, int32_t first, int32_t count, int32_t slot, const dabb_subchannel* sc)
{
    if (!ctx || !sc || first < 0 || count < 1 || first + count > ctx->S || slot < 0 || slot >= ctx->n_slots) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    ProtProfile prof;
    if (make_prot_profile(sc->bitrate, sc->short_form, sc->uep_level, sc->eep_profile_a, sc->eep_level, prof) < 0) { ctx->err = "unsupported protection profile"; return DABB_E_UNSUPPORTED; }
    if (sc->length_cu < 1 || sc->length_cu > ctx->max_cu || sc->start_cu < 0 || sc->start_cu + sc->length_cu > 864) { ctx->err = "sub-channel does not fit (raise dabb_config.max_subch_cu)"; return DABB_E_ARG; }
    if (prof.in_bits > sc->length_cu * 64) { ctx->err = "protection profile needs more bits than the sub-channel holds"; return DABB_E_ARG; }
    if (sc->dabplus && sc->bitrate / 8 > 64) { ctx->err = "DAB+ sub-channels above 512 kbit/s are not supported"; return DABB_E_UNSUPPORTED; }
    auto& sl = ctx->slot[slot];
    const int S = ctx->S;
    int rc = 0;
    if (sl.configured && (memcmp(&sl.prof, &prof, sizeof prof) != 0)) {
        // a slot decodes one code geometry for all its streams in one launch
        bool others = false;
        for (int i = 0; i < S; i++) if ((i < first || i >= first + count) && ctx->h_slots[(size_t)i * ctx->n_slots + slot].enabled) others = true;
        if (others) { ctx->err = "streams sharing a slot must use the same bitrate/protection; use another slot"; return DABB_E_UNSUPPORTED; }
        // the slot changes its code geometry: give back the buffers sized for the old one (all streams are idle: sync_all above)
        dfree(ctx, sl.d_steptab); dfree(ctx, sl.d_stage_off); dfree(ctx, sl.d_prbs_words); dfree(ctx, sl.d_frag); dfree(ctx, sl.d_valid);
        dfree(ctx, sl.d_logical); dfree(ctx, sl.d_window); dfree(ctx, sl.d_sf); dfree(ctx, sl.d_info); dfree(ctx, sl.d_dec);
        sl.configured = false;
    }
    if (!sl.configured) {
        sl.prof = prof; sl.nbits = 24 * prof.bitrate; sl.nsteps = sl.nbits + 6; sl.frag_pitch = ctx->ring_pitch; sl.flen = 3 * prof.bitrate;
        std::vector<int16_t> map((size_t)sl.nsteps * 4);
        build_msc_map(*ctx->host, prof, map.data());
        std::vector<uint2> steps; std::vector<uint32_t> soff;
        build_vit_tables_u2(map.data(), sl.nsteps, steps, soff);
        std::vector<uint32_t> w; pack_prbs_words(ctx->host->prbs, sl.nbits, w);
        const int flen_pad = (sl.flen + 15) & ~15;
        if ((rc = dalloc(ctx, &sl.d_steptab, steps.size())) || (rc = dalloc(ctx, &sl.d_stage_off, soff.size())) || (rc = dalloc(ctx, &sl.d_prbs_words, w.size())) ||
            (rc = dalloc(ctx, &sl.d_frag, (size_t)S * 4 * sl.frag_pitch + VIT_FRAG_SLACK)) ||
            (rc = dalloc(ctx, &sl.d_valid, (size_t)S * 4)) || (rc = dalloc(ctx, &sl.d_logical, (size_t)S * 4 * flen_pad)) ||
            (rc = dalloc(ctx, &sl.d_window, (size_t)S * 5 * flen_pad)) || (rc = dalloc(ctx, &sl.d_sf, (size_t)S * 5 * flen_pad)) ||
            (rc = dalloc(ctx, &sl.d_info, (size_t)S * 16)))
            return rc;
        if (!sl.d_ring && (rc = dalloc(ctx, &sl.d_ring, (size_t)S * MSC_RING * ctx->ring_pitch))) return rc;
        CK(cudaMemcpyAsync(sl.d_steptab, steps.data(), steps.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(sl.d_stage_off, soff.data(), soff.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(sl.d_prbs_words, w.data(), w.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        { void* q = nullptr; if (cudaMalloc(&q, vit_dec_bytes(S * 4, sl.nsteps)) != cudaSuccess) { ctx->err = "cudaMalloc(slot decisions)"; return DABB_E_NOMEM; } ctx->allocs.push_back(q); sl.d_dec = (uint2*)q; }
        sl.configured = true;
        if (sl.flen > ctx->flen_max) ctx->flen_max = sl.flen;
    }
    MscSlotState v{}; v.enabled = 1; v.start_cu = sc->start_cu; v.frag = sc->length_cu * 64; v.bitrate = sc->bitrate; v.dabplus = sc->dabplus; v.cif_count = 0; v.sf_frame_count = 0;
    set_slot_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_slots, ctx->n_slots, slot, first, count, v);
    for (int i = first; i < first + count; i++) ctx->h_slots[(size_t)i * ctx->n_slots + slot] = v;
    { int rc_ = check_launch(ctx, "set_slot_kernel"); cudaStreamSynchronize(ctx->stream); return rc_; }
}

int dabb_remove_subchannel(dabb_ctx* ctx, int32_t first, int32_t count, int32_t slot)
{
    if (!ctx || first < 0 || count < 1 || first + count > ctx->S || slot < 0 || slot >= ctx->n_slots) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    MscSlotState v{};
    set_slot_kernel<<<(count + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_slots, ctx->n_slots, slot, first, count, v);
    for (int i = first; i < first + count; i++) ctx->h_slots[(size_t)i * ctx->n_slots + slot] = v;
    { int rc_ = check_launch(ctx, "set_slot_kernel"); cudaStreamSynchronize(ctx->stream); return rc_; }
}


static int run_fic(dabb_ctx* ctx, const int8_t* soft, int64_t soft_stride, const int32_t* active, int n_frames, uint8_t* fibs, int32_t* crc, cudaStream_t st, uint2* dec)
{
    int rc;
    // FIC codeword (frame f, block b) = softbits [2304 b, 2304 b + 2304) of the frame's first three symbols (fic-handler.cpp:111-127):
    // the decoder kernel reads them where the OFDM kernel wrote them and de-punctures on the fly
    ViterbiParams vp{};
    vp.frag = soft; vp.cw_div = 4; vp.outer_stride = soft_stride; vp.inner_stride = 2304; vp.steptab = ctx->d_fic_steptab; vp.stage_off = ctx->d_fic_stage_off;
    vp.n_cw = n_frames * 4; vp.nsteps = 774; vp.nbits = 768;
    vp.dec = dec; vp.out = fibs; vp.out_stride = 96; vp.prbs_words = ctx->d_fic_prbs_words; vp.valid = nullptr;
    vp.trace = ctx->trace; vp.trace_kind = 2; vp.split = ctx->vit_split;
    launch_viterbi(vp, st, ctx->vit_stages_now);
    if ((rc = check_launch(ctx, "viterbi_kernel(FIC)"))) return rc;
    launch_fic_crc(fibs, active, n_frames, crc, st);
    return check_launch(ctx, "fic_crc_kernel");
}

static int enqueue_step(dabb_ctx* ctx, const dabb_io* io, const float2* iq, int64_t stride);

int dabb_process_async(dabb_ctx* ctx, const dabb_io* io)
{
    if (!ctx || !io || !io->iq || !io->buf_start) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    const int S = ctx->S;
    int rc;
    cudaStream_t A = ctx->stream;
    const float2* iq = reinterpret_cast<const float2*>(io->iq);
    int64_t stride = io->stride_samples;
    const int fmt = io->iq_format;
    if (fmt < 0 || fmt > DABB_IQ_S16BE) { ctx->err = "unknown iq_format"; return DABB_E_ARG; }
    if (io->iq_is_host || fmt != DABB_IQ_CF32) {
        const size_t need = (size_t)S * io->buf_len;
        if (need > ctx->iq_stage_samples) {
            sync_all(ctx);
            if (ctx->d_iq_stage) cudaFree(ctx->d_iq_stage);
            ctx->d_iq_stage = nullptr; ctx->iq_stage_samples = 0;
            cudaError_t e = cudaMalloc((void**)&ctx->d_iq_stage, need * sizeof(float2));
            if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc(iq staging): ") + cudaGetErrorString(e); return DABB_E_NOMEM; }
            ctx->iq_stage_samples = need;
        }
        if (fmt == DABB_IQ_CF32) {
            CK(cudaMemcpy2DAsync(ctx->d_iq_stage, (size_t)io->buf_len * sizeof(float2), io->iq, (size_t)stride * sizeof(float2), (size_t)io->buf_len * sizeof(float2), S,
                                 cudaMemcpyHostToDevice, A));
        } else {
            const size_t bps = (fmt == DABB_IQ_U8 || fmt == DABB_IQ_S8) ? 2 : 4;
            const uint8_t* raw = reinterpret_cast<const uint8_t*>(io->iq);
            int64_t raw_stride = stride;
            if (io->iq_is_host) {
                if (need * bps > ctx->raw_stage_bytes) {
                    sync_all(ctx);
                    if (ctx->d_raw_stage) cudaFree(ctx->d_raw_stage);
                    ctx->d_raw_stage = nullptr; ctx->raw_stage_bytes = 0;
                    cudaError_t e = cudaMalloc((void**)&ctx->d_raw_stage, need * bps);
                    if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc(raw staging): ") + cudaGetErrorString(e); return DABB_E_NOMEM; }
                    ctx->raw_stage_bytes = need * bps;
                }
                CK(cudaMemcpy2DAsync(ctx->d_raw_stage, (size_t)io->buf_len * bps, io->iq, (size_t)stride * bps, (size_t)io->buf_len * bps, S, cudaMemcpyHostToDevice, A));
                raw = ctx->d_raw_stage; raw_stride = io->buf_len;
            }
            const dim3 grid((unsigned)((io->buf_len + 255) / 256), (unsigned)S);
            convert_iq_kernel<<<grid, 256, 0, A>>>(raw, raw_stride, fmt, ctx->d_iq_stage, io->buf_len, io->buf_len, S);
            if ((rc = check_launch(ctx, "convert_iq_kernel"))) return rc;
        }
        iq = ctx->d_iq_stage; stride = io->buf_len;
    }
    return enqueue_step(ctx, io, iq, stride);
}

// one decode step on device-resident cf32 samples (stream s at iq + s * stride)
static int enqueue_step(dabb_ctx* ctx, const dabb_io* io, const float2* iq, int64_t stride)
{
    const int S = ctx->S;
    int rc;
    // Two lanes: A = acquisition, time sync, OFDM demod and the sync-state update of frame n; B = FIC, MSC, RS and the
    // result record of frame n.  Lane A of frame n+1 only needs lane A of frame n, so B(n) overlaps A(n+1); the per-frame
    // scratch and the softbits are double buffered by frame parity.  With per-kernel profiling on everything runs
    // serially on the main stream.
    const int par = (int)(ctx->step & 1);
    const bool serial = ctx->prof;
    // three staging buffers by default; DABB_LANES=partition uses two (39 KB) so that one Viterbi CTA fits beside four OFDM CTAs
    ctx->vit_stages_now = getenv("DABB_VIT_STAGES") ? atoi(getenv("DABB_VIT_STAGES")) : (ctx->persistent && !serial ? 2 : 3);
    cudaStream_t A = ctx->stream, B = serial ? ctx->stream : ctx->streamB;
    StepScratch* scr = ctx->d_scr + (size_t)par * S;
    int64_t* d_win = ctx->d_win + (size_t)par * S; int64_t* d_prs = ctx->d_prs + (size_t)par * S;
    int32_t* d_nco_sync = ctx->d_nco_sync + (size_t)par * 2 * S; int32_t* d_nco_frame = ctx->d_nco_frame + (size_t)par * 4 * S;
    int32_t* d_active = ctx->d_active + (size_t)par * S; int32_t* d_index = ctx->d_index + (size_t)par * S; int32_t* d_snr = ctx->d_snr + (size_t)par * S;
    float2* d_fc = ctx->d_fc + (size_t)par * S * ctx->fc_pitch; float* d_lvl = ctx->d_lvl + (size_t)par * S * ctx->fc_pitch;
    int8_t* d_soft = ctx->d_soft + (size_t)par * S * DABB_SOFT_PER_FRAME;

    // the buffers of this parity were last used by lane B two frames ago
    if (!serial && ctx->evB_valid[par]) CK(cudaStreamWaitEvent(A, ctx->evB[par], 0));
    // the time-sync kernels run on the highest-priority stream: it inherits everything lane A has waited for so far (previous step, input
    // copies, buffer hand-back) through one event, and lane A continues behind its last kernel
    cudaStream_t A_low = A;
    if (!serial && ctx->streamP) { CK(cudaEventRecord(ctx->evPre, A)); CK(cudaStreamWaitEvent(ctx->streamP, ctx->evPre, 0)); A = ctx->streamP; }
    CK(cudaMemcpyAsync(ctx->d_buf_start, io->buf_start, sizeof(int64_t) * S, cudaMemcpyHostToDevice, A));
    prof_mark(ctx, "__step_begin");
    const int tb = 128, gb = (S + tb - 1) / tb;
    // ---------------- lane A
    acquire_kernel<<<(S + 31) / 32, 32, 0, A>>>(ctx->d_state, iq, stride, ctx->d_buf_start, io->buf_len, ctx->dev.osc, S);
    if ((rc = check_launch(ctx, "acquire_kernel"))) return rc;
    plan_kernel<<<gb, tb, 0, A>>>(ctx->d_state, scr, ctx->d_buf_start, io->buf_len, S, d_win, d_nco_sync, d_active);
    if ((rc = check_launch(ctx, "plan_kernel"))) return rc;
    SyncParams sp{}; sp.iq = iq; sp.stride = stride; sp.win_start = d_win; sp.nco = d_nco_sync; sp.active = d_active; sp.index_out = d_index; sp.cir_out = ctx->d_cir; sp.n = S;
    sp.fic_ratio = ctx->d_fic_ratio; sp.coarse_out = ctx->disable_coarse ? nullptr : ctx->d_coarse; sp.placement = ctx->placement; sp.freqsync = ctx->freqsync;
    sp.cir_work = ctx->d_cir_work; sp.search_generic = ctx->search_generic;
    launch_find_index(ctx->dev, sp, ctx->fft_mode, A, 0);
    if ((rc = check_launch(ctx, "find_index_kernel"))) return rc;
    launch_find_index(ctx->dev, sp, ctx->fft_mode, A, 1);
    if ((rc = check_launch(ctx, "find_search_kernel"))) return rc;
    if (sp.coarse_out) {
        launch_find_index(ctx->dev, sp, ctx->fft_mode, A, 2);
        if ((rc = check_launch(ctx, "find_coarse_kernel"))) return rc;
    }
    post_sync_kernel<<<gb, tb, 0, A>>>(ctx->d_state, scr, d_index, ctx->disable_coarse ? nullptr : ctx->d_coarse, S, d_prs, d_nco_frame, d_active);
    if ((rc = check_launch(ctx, "post_sync_kernel"))) return rc;
    if (A != A_low) { CK(cudaEventRecord(ctx->evPost, A)); A = A_low; CK(cudaStreamWaitEvent(A, ctx->evPost, 0)); }
    OfdmParams op{}; op.iq = iq; op.stride = stride; op.prs_start = d_prs; op.nco = d_nco_frame; op.active = d_active; op.soft = d_soft; op.soft_stride = DABB_SOFT_PER_FRAME;
    op.r1 = ctx->d_r1; op.freqcorr = d_fc; op.level = d_lvl; op.snr = d_snr; op.n_frames = S; op.groups = ctx->groups; op.sym_per_cta = 75 / ctx->groups;
    op.trace = ctx->trace;
    if (ctx->persistent && !serial) {
        if (!ctx->d_work && (rc = dalloc(ctx, &ctx->d_work, 1))) return rc;
        CK(cudaMemsetAsync(ctx->d_work, 0, sizeof(unsigned int), A));
        op.work = ctx->d_work;
    }
    op.n_full = S - ctx->tail_frames; op.tail_groups = ctx->tail_frames ? ctx->tail_groups : 1; op.fc_pitch = ctx->fc_pitch; op.nco_fast = ctx->nco_fast;
    // pipelined mode: 50 KB per CTA -> four OFDM CTAs per SM, leaving registers and 23 KB of shared memory for one lane-B CTA
    op.smem_floor = (serial && !getenv("DABB_CORESIDENT_SERIAL")) ? 0 : ctx->ofdm_smem_floor;     // DABB_CORESIDENT_SERIAL: time the capped kernel alone
    launch_ofdm_demod(ctx->dev, op, ctx->fft_mode, A);
    if ((rc = check_launch(ctx, "ofdm_demod_kernel"))) return rc;
    advance_kernel<<<gb, tb, 0, A>>>(ctx->d_state, scr, S, ctx->fc_pitch, d_fc, d_lvl);
    if ((rc = check_launch(ctx, "advance_kernel"))) return rc;
    if (ctx->d_null) {
        null_tap_kernel<<<S, 256, 0, A>>>(scr, iq, stride, ctx->d_buf_start, io->buf_len, ctx->dev.osc, ctx->d_null);
        if ((rc = check_launch(ctx, "null_tap_kernel"))) return rc;
        if (ctx->decode_tii) {
            if (!ctx->d_tii && (rc = dalloc(ctx, &ctx->d_tii, (size_t)S * 2 * TU))) return rc;
            launch_tii_spectra(ctx->dev, iq, stride, d_prs, d_nco_frame, d_active, ctx->d_null, ctx->d_tii, S, A);
            if ((rc = check_launch(ctx, "tii_spectra_kernel"))) return rc;
        }
    }
    if (!serial) { CK(cudaEventRecord(ctx->evA[par], A)); CK(cudaStreamWaitEvent(B, ctx->evA[par], 0)); }
    // ---------------- lane B: FIC chain forked onto its own stream, MSC chain per slot (longest first), then the result record.
    // (One fused Viterbi launch over FIC + MSC codewords was measured slower: 1.44 ms vs 0.34 + 1.02 ms.)
    cudaStream_t F = serial ? ctx->stream : ctx->stream2;
    if (!serial) { CK(cudaEventRecord(ctx->ev_ofdm, B)); CK(cudaStreamWaitEvent(F, ctx->ev_ofdm, 0)); }
    if ((rc = run_fic(ctx, d_soft, DABB_SOFT_PER_FRAME, d_active, S, ctx->d_fibs, ctx->d_crc, F, ctx->d_dec_fic))) return rc;
    if (!serial) CK(cudaEventRecord(ctx->ev_fic, F));
    const int32_t* h_info[DABB_MAX_SUBCH] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < ctx->n_slots; k++) {
        auto& sl = ctx->slot[k];
        h_info[k] = sl.d_info;
        if (!sl.configured) continue;
        const int flen_pad = (sl.flen + 15) & ~15;
        CK(cudaMemsetAsync(sl.d_valid, 0, sizeof(int32_t) * S * 4, B));
        MscCollectParams cp{}; cp.soft = d_soft; cp.soft_stride = DABB_SOFT_PER_FRAME; cp.active = d_active; cp.slots = ctx->d_slots; cp.n_slots = ctx->n_slots; cp.slot = k; cp.ring = sl.d_ring; cp.ring_pitch = ctx->ring_pitch;
        launch_msc_collect(cp, S, B);
        if ((rc = check_launch(ctx, "msc_collect_kernel"))) return rc;
        MscPrepParams pp{}; pp.active = d_active; pp.slots = ctx->d_slots; pp.n_slots = ctx->n_slots; pp.slot = k; pp.ring = sl.d_ring; pp.ring_pitch = ctx->ring_pitch;
        pp.frag_out = sl.d_frag; pp.frag_pitch = sl.frag_pitch; pp.valid = sl.d_valid;
        launch_msc_gather(pp, S, B);
        if ((rc = check_launch(ctx, "msc_gather_kernel"))) return rc;
        ViterbiParams vp{}; vp.frag = sl.d_frag; vp.cw_div = 1; vp.outer_stride = sl.frag_pitch; vp.inner_stride = 0; vp.steptab = sl.d_steptab; vp.stage_off = sl.d_stage_off;
        vp.n_cw = S * 4; vp.nsteps = sl.nsteps; vp.nbits = sl.nbits; vp.dec = sl.d_dec;
        vp.out = sl.d_logical; vp.out_stride = flen_pad; vp.prbs_words = sl.d_prbs_words; vp.valid = sl.d_valid;
        vp.trace = ctx->trace; vp.trace_kind = 3; vp.split = ctx->vit_split;
        launch_viterbi(vp, B, ctx->vit_stages_now);
        if ((rc = check_launch(ctx, "viterbi_kernel(MSC)"))) return rc;
        SuperframeParams fp{}; fp.active = d_active; fp.slots = ctx->d_slots; fp.n_slots = ctx->n_slots; fp.slot = k; fp.n_streams = S; fp.logical = sl.d_logical; fp.logical_stride = flen_pad;
        fp.valid = sl.d_valid; fp.window = sl.d_window; fp.window_pitch = 5 * flen_pad; fp.sf_out = sl.d_sf; fp.sf_pitch = 5 * flen_pad; fp.info = sl.d_info; fp.gf_exp = ctx->dev.gf_exp; fp.gf_log = ctx->dev.gf_log;
        launch_superframe(fp, B);
        if ((rc = check_launch(ctx, "superframe_kernel"))) return rc;
    }
    if (!serial) CK(cudaStreamWaitEvent(B, ctx->ev_fic, 0));
    CK(cudaMemcpyAsync((void*)ctx->d_info_tab, h_info, sizeof(void*) * DABB_MAX_SUBCH, cudaMemcpyHostToDevice, B));
    finalize_kernel<<<gb, tb, 0, B>>>(scr, ctx->d_slots, ctx->n_slots, S, ctx->d_fic_ratio, d_snr, ctx->d_crc, ctx->d_info_tab, ctx->d_results);
    if ((rc = check_launch(ctx, "finalize_kernel"))) return rc;
    if (!serial) { CK(cudaEventRecord(ctx->evB[par], B)); ctx->evB_valid[par] = true; }
    ctx->last_parity = par; ctx->step++;
    return DABB_OK;
}

int dabb_process(dabb_ctx* ctx, const dabb_io* io)
{
    int rc = dabb_process_async(ctx, io);
    if (rc) return rc;
    const int S = ctx->S;
    cudaStream_t B = ctx->prof ? ctx->stream : ctx->streamB;
    if (io->results) CK(cudaMemcpyAsync(ctx->h_results, ctx->d_results, sizeof(dabb_frame_result) * S, cudaMemcpyDeviceToHost, B));
    if (io->fibs) CK(cudaMemcpyAsync(ctx->h_fibs, ctx->d_fibs, (size_t)S * 12 * 32, cudaMemcpyDeviceToHost, B));
    CK(cudaStreamSynchronize(B));
    CK(cudaStreamSynchronize(ctx->stream));
    if (io->results) memcpy(io->results, ctx->h_results, sizeof(dabb_frame_result) * S);
    if (io->fibs) memcpy(io->fibs, ctx->h_fibs, (size_t)S * 12 * 32);
    // logical frames and superframes: strided device -> host copies (small)
    for (int k = 0; k < ctx->n_slots; k++) {
        auto& sl = ctx->slot[k];
        if (!sl.configured) continue;
        const int flen_pad = (sl.flen + 15) & ~15;
        if (io->msc) {
            if (io->msc_stride < sl.flen) { ctx->err = "msc_stride smaller than the logical frame"; return DABB_E_ARG; }
            // device rows: [S*4][flen_pad] -> host [S][MAX_SUBCH][4][msc_stride]
            for (int c = 0; c < 4; c++)
                CK(cudaMemcpy2D(io->msc + ((size_t)k * 4 + c) * io->msc_stride, (size_t)DABB_MAX_SUBCH * 4 * io->msc_stride,
                                sl.d_logical + (size_t)c * flen_pad, (size_t)4 * flen_pad, sl.flen, S, cudaMemcpyDeviceToHost));
        }
        if (io->sf) {
            if (io->sf_stride < 5 * sl.flen) { ctx->err = "sf_stride smaller than the superframe"; return DABB_E_ARG; }
            CK(cudaMemcpy2D(io->sf + (size_t)k * io->sf_stride, (size_t)DABB_MAX_SUBCH * io->sf_stride,
                            sl.d_sf, (size_t)5 * flen_pad, 5 * sl.flen, S, cudaMemcpyDeviceToHost));
        }
    }
    return DABB_OK;
}

// ---- pipelined host-buffer path -------------------------------------------------------------------------------------------------
// dabb_submit: H2D of the step's samples on a copy stream into one of two device staging slots, the kernels on the lanes, D2H of the
// results into one of two pinned result slots on lane B - and return.  dabb_collect: wait for the OLDEST outstanding step and hand
// its results to the pointers given at submit.  With two steps outstanding the copy of step n+1 overlaps the kernels of step n and
// the result read-back of step n-1.
static int ensure_pending(dabb_ctx* ctx, dabb_ctx::Pending& p)
{
    const int S = ctx->S;
    if (!p.done) CK(cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming));
    if (!p.h_res) CK(cudaHostAlloc((void**)&p.h_res, sizeof(dabb_frame_result) * S, cudaHostAllocDefault));
    if (!p.h_fibs) CK(cudaHostAlloc((void**)&p.h_fibs, (size_t)S * 12 * 32, cudaHostAllocDefault));
    size_t msc = 0, sf = 0;
    for (int k = 0; k < ctx->n_slots; k++) {
        const auto& sl = ctx->slot[k];
        p.flen[k] = sl.configured ? sl.flen : 0;
        if (!sl.configured) continue;
        const size_t fp = (size_t)((sl.flen + 15) & ~15);
        p.msc_off[k] = msc; p.sf_off[k] = sf;
        msc += (size_t)S * 4 * fp; sf += (size_t)S * 5 * fp;
    }
    if (msc > p.msc_bytes) { if (p.h_msc) cudaFreeHost(p.h_msc); p.h_msc = nullptr; CK(cudaHostAlloc((void**)&p.h_msc, msc, cudaHostAllocDefault)); p.msc_bytes = msc; }
    if (sf > p.sf_bytes) { if (p.h_sf) cudaFreeHost(p.h_sf); p.h_sf = nullptr; CK(cudaHostAlloc((void**)&p.h_sf, sf, cudaHostAllocDefault)); p.sf_bytes = sf; }
    return DABB_OK;
}

int dabb_submit(dabb_ctx* ctx, const dabb_io* io)
{
    if (!ctx || !io || !io->iq || !io->buf_start || !io->iq_is_host) return DABB_E_ARG;
    if (ctx->prof) { ctx->err = "dabb_submit is not available while per-kernel profiling is on"; return DABB_E_STATE; }
    if (ctx->pend_count == 2) { ctx->err = "two steps are outstanding: call dabb_collect first"; return DABB_E_STATE; }
    cudaSetDevice(ctx->device);
    const int S = ctx->S, fmt = io->iq_format;
    if (fmt < 0 || fmt > DABB_IQ_S16BE) { ctx->err = "unknown iq_format"; return DABB_E_ARG; }
    int rc;
    if (!ctx->streamC) {
        CK(cudaStreamCreateWithFlags(&ctx->streamC, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) { CK(cudaEventCreateWithFlags(&ctx->evH2D[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->evStageFree[i], cudaEventDisableTiming)); }
    }
    const int slot = (int)(ctx->submits & 1);
    dabb_ctx::Pending& p = ctx->pend[(ctx->pend_head + ctx->pend_count) & 1];
    if ((rc = ensure_pending(ctx, p))) return rc;
    cudaStream_t A = ctx->stream, B = ctx->streamB, Cs = ctx->streamC;
    const size_t bps = fmt == DABB_IQ_CF32 ? 8 : ((fmt == DABB_IQ_U8 || fmt == DABB_IQ_S8) ? 2 : 4);
    const size_t need = (size_t)S * io->buf_len * bps;
    if (need > ctx->stage2_bytes[slot]) {
        sync_all(ctx); cudaStreamSynchronize(Cs);
        if (ctx->d_stage2[slot]) cudaFree(ctx->d_stage2[slot]);
        ctx->d_stage2[slot] = nullptr; ctx->stage2_bytes[slot] = 0;
        cudaError_t e = cudaMalloc(&ctx->d_stage2[slot], need);
        if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc(pipelined staging): ") + cudaGetErrorString(e); return DABB_E_NOMEM; }
        ctx->stage2_bytes[slot] = need; ctx->stageFreeValid[slot] = false;
    }
    // the staging slot was last read by the step submitted two calls ago
    if (ctx->stageFreeValid[slot]) CK(cudaStreamWaitEvent(Cs, ctx->evStageFree[slot], 0));
    // carry_samples: the head of every window is the tail of the previous one, which still lies in the other staging slot
    const int64_t carry = io->carry_samples;
    if (carry) {
        if (carry < 0 || carry >= io->buf_len || ctx->submits == 0 || ctx->stage_last_fmt != fmt || carry > ctx->stage_last_len || !ctx->d_stage2[slot ^ 1]) {
            ctx->err = "carry_samples needs a previous dabb_submit of the same format with a window at least that long"; return DABB_E_ARG;
        }
        for (int s = 0; s < S; s++)
            if (io->buf_start[s] != ctx->stage_last_start[s] + ctx->stage_last_len - carry) { ctx->err = "carry_samples: a window does not start where the previous one ended minus the carry"; return DABB_E_ARG; }
        const unsigned char* prev = reinterpret_cast<const unsigned char*>(ctx->d_stage2[slot ^ 1]);
        CK(cudaMemcpy2DAsync(ctx->d_stage2[slot], (size_t)io->buf_len * bps, prev + (size_t)(ctx->stage_last_len - carry) * bps, (size_t)ctx->stage_last_len * bps,
                             (size_t)carry * bps, S, cudaMemcpyDeviceToDevice, Cs));          // same stream as the copy that filled the other slot
    }
    CK(cudaMemcpy2DAsync(reinterpret_cast<unsigned char*>(ctx->d_stage2[slot]) + (size_t)carry * bps, (size_t)io->buf_len * bps,
                         reinterpret_cast<const unsigned char*>(io->iq) + (size_t)carry * bps, (size_t)io->stride_samples * bps, (size_t)(io->buf_len - carry) * bps, S, cudaMemcpyHostToDevice, Cs));
    ctx->stage_last_fmt = fmt; ctx->stage_last_len = io->buf_len; ctx->stage_last_start.assign(io->buf_start, io->buf_start + S);
    CK(cudaEventRecord(ctx->evH2D[slot], Cs));
    CK(cudaStreamWaitEvent(A, ctx->evH2D[slot], 0));
    const float2* iq = reinterpret_cast<const float2*>(ctx->d_stage2[slot]);
    if (fmt != DABB_IQ_CF32) {
        const size_t ns = (size_t)S * io->buf_len;
        if (ns > ctx->iq_stage_samples) {
            sync_all(ctx);
            if (ctx->d_iq_stage) cudaFree(ctx->d_iq_stage);
            ctx->d_iq_stage = nullptr; ctx->iq_stage_samples = 0;
            cudaError_t e = cudaMalloc((void**)&ctx->d_iq_stage, ns * sizeof(float2));
            if (e != cudaSuccess) { ctx->err = std::string("cudaMalloc(iq staging): ") + cudaGetErrorString(e); return DABB_E_NOMEM; }
            ctx->iq_stage_samples = ns;
        }
        const dim3 grid((unsigned)((io->buf_len + 255) / 256), (unsigned)S);
        convert_iq_kernel<<<grid, 256, 0, A>>>(reinterpret_cast<const uint8_t*>(ctx->d_stage2[slot]), io->buf_len, fmt, ctx->d_iq_stage, io->buf_len, io->buf_len, S);
        if ((rc = check_launch(ctx, "convert_iq_kernel"))) return rc;
        CK(cudaEventRecord(ctx->evStageFree[slot], A)); ctx->stageFreeValid[slot] = true;       // the raw bytes are consumed
        iq = ctx->d_iq_stage;
    }
    if ((rc = enqueue_step(ctx, io, iq, io->buf_len))) return rc;
    if (fmt == DABB_IQ_CF32) { CK(cudaEventRecord(ctx->evStageFree[slot], A)); ctx->stageFreeValid[slot] = true; }   // lane A has read the samples
    // results -> pinned slot, on lane B behind the step's result record
    if (io->results) CK(cudaMemcpyAsync(p.h_res, ctx->d_results, sizeof(dabb_frame_result) * S, cudaMemcpyDeviceToHost, B));
    if (io->fibs) CK(cudaMemcpyAsync(p.h_fibs, ctx->d_fibs, (size_t)S * 12 * 32, cudaMemcpyDeviceToHost, B));
    for (int k = 0; k < ctx->n_slots; k++) {
        const auto& sl = ctx->slot[k];
        if (!sl.configured) continue;
        const size_t fp = (size_t)((sl.flen + 15) & ~15);
        if (io->msc) {
            if (io->msc_stride < sl.flen) { ctx->err = "msc_stride smaller than the logical frame"; return DABB_E_ARG; }
            CK(cudaMemcpyAsync(p.h_msc + p.msc_off[k], sl.d_logical, (size_t)S * 4 * fp, cudaMemcpyDeviceToHost, B));
        }
        if (io->sf) {
            if (io->sf_stride < 5 * sl.flen) { ctx->err = "sf_stride smaller than the superframe"; return DABB_E_ARG; }
            CK(cudaMemcpyAsync(p.h_sf + p.sf_off[k], sl.d_sf, (size_t)S * 5 * fp, cudaMemcpyDeviceToHost, B));
        }
    }
    CK(cudaEventRecord(p.done, B));
    p.io = *io; p.used = true;
    ctx->pend_count++; ctx->submits++;
    return DABB_OK;
}

int dabb_collect(dabb_ctx* ctx)
{
    if (!ctx) return DABB_E_ARG;
    if (ctx->pend_count == 0) { ctx->err = "nothing outstanding"; return DABB_E_STATE; }
    cudaSetDevice(ctx->device);
    dabb_ctx::Pending& p = ctx->pend[ctx->pend_head];
    CK(cudaEventSynchronize(p.done));
    const int S = ctx->S;
    const dabb_io& io = p.io;
    if (io.results) memcpy(io.results, p.h_res, sizeof(dabb_frame_result) * S);
    if (io.fibs) memcpy(io.fibs, p.h_fibs, (size_t)S * 12 * 32);
    for (int k = 0; k < ctx->n_slots; k++) {
        const int flen = p.flen[k];
        if (!flen) continue;
        const size_t fp = (size_t)((flen + 15) & ~15);
        if (io.msc)       // pinned rows [S*4][fp] -> caller [S][MAX_SUBCH][4][msc_stride]
            for (int s = 0; s < S; s++) for (int c = 0; c < 4; c++)
                memcpy(io.msc + (((size_t)s * DABB_MAX_SUBCH + k) * 4 + c) * io.msc_stride, p.h_msc + p.msc_off[k] + ((size_t)s * 4 + c) * fp, (size_t)flen);
        if (io.sf)
            for (int s = 0; s < S; s++) memcpy(io.sf + ((size_t)s * DABB_MAX_SUBCH + k) * io.sf_stride, p.h_sf + p.sf_off[k] + (size_t)s * 5 * fp, (size_t)5 * flen);
    }
    p.used = false;
    ctx->pend_head ^= 1; ctx->pend_count--;
    return DABB_OK;
}

int dabb_profile(dabb_ctx* ctx, int32_t enable)
{
    if (!ctx) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    if (!enable && ctx->prof) prof_collect(ctx);
    if (enable && !ctx->prof) { ctx->prof_acc.clear(); ctx->prof_used = 0; }
    ctx->prof = enable != 0;
    return 0;
}

int dabb_profile_read(dabb_ctx* ctx, char* out, size_t cap)
{
    if (!ctx || !out || cap < 3) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    prof_collect(ctx);
    std::string j = "{";
    for (size_t i = 0; i < ctx->prof_acc.size(); i++) {
        char b[256];
        snprintf(b, sizeof b, "%s\"%s\": {\"ms\": %.6f, \"n\": %ld}", i ? ", " : "", ctx->prof_acc[i].first.c_str(), ctx->prof_acc[i].second.first, ctx->prof_acc[i].second.second);
        j += b;
    }
    j += "}";
    if (j.size() + 1 > cap) return DABB_E_ARG;
    memcpy(out, j.c_str(), j.size() + 1);
    return 0;
}

int dabb_read_tap(dabb_ctx* ctx, int32_t what, void* host_out, size_t bytes)
{
    if (!ctx || !host_out) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    if (what == 0) { const size_t n = (size_t)ctx->S * DABB_SOFT_PER_FRAME; CK(cudaMemcpy(host_out, ctx->d_soft + (size_t)ctx->last_parity * n, bytes < n ? bytes : n, cudaMemcpyDeviceToHost)); return 0; }
    if (what == 1 && ctx->d_cir) { const size_t n = (size_t)ctx->S * TU * 4; CK(cudaMemcpy(host_out, ctx->d_cir, bytes < n ? bytes : n, cudaMemcpyDeviceToHost)); return 0; }
    if (what == 2 && ctx->d_r1) {      // every 96th logical carrier of every data symbol (constellationDecimation, ofdm-decoder.h:87)
        const size_t n = (size_t)ctx->S * 75 * 16;
        if (bytes < n * 8) { ctx->err = "constellation tap needs n_streams * 1200 * 8 bytes"; return DABB_E_ARG; }
        CK(cudaMemcpy2D(host_out, 8, ctx->d_r1, 96 * 8, 8, n, cudaMemcpyDeviceToHost));
        return 0;
    }
    if (what == 4 && ctx->d_tii && ctx->decode_tii) { const size_t n = (size_t)ctx->S * 2 * TU * 8; CK(cudaMemcpy(host_out, ctx->d_tii, bytes < n ? bytes : n, cudaMemcpyDeviceToHost)); return 0; }
    if (what == 3 && ctx->d_null) { const size_t n = (size_t)ctx->S * TNULL * 8; CK(cudaMemcpy(host_out, ctx->d_null, bytes < n ? bytes : n, cudaMemcpyDeviceToHost)); return 0; }
    ctx->err = "tap not available (keep_taps = 0?)";
    return DABB_E_STATE;
}

// ------------------------------------------------------------------------------------------------ stage-level API
int dabb_ofdm_demod(dabb_ctx* ctx, const float* iq, int64_t stride, const int64_t* prs_start, int32_t n, const int32_t* nco, int8_t* soft, float* r1, float* fc)
{
    if (!ctx || !iq || !prs_start || !soft || n < 1) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    OfdmParams op{}; op.iq = reinterpret_cast<const float2*>(iq); op.stride = stride; op.prs_start = prs_start; op.active = nullptr; op.soft = soft; op.soft_stride = DABB_SOFT_PER_FRAME;
    op.r1 = reinterpret_cast<float2*>(r1); op.freqcorr = reinterpret_cast<float2*>(fc); op.snr = nullptr; op.n_frames = n;
    op.groups = (fc || n >= 1024) ? 1 : (n >= 64 ? 5 : 25); op.sym_per_cta = 75 / op.groups; op.n_full = n; op.tail_groups = 1; op.fc_pitch = op.groups; op.nco_fast = ctx->nco_fast;
    int32_t* nco4 = nullptr;
    if (nco) {
        // expand {lp applied to PRS sample 0, Hz} to the 4-entry form of the pipeline (same increment for PRS and symbols)
        std::vector<int32_t> h(2 * (size_t)n), h4(4 * (size_t)n);
        CK(cudaMemcpy(h.data(), nco, sizeof(int32_t) * 2 * n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) { h4[4 * i] = h[2 * i]; h4[4 * i + 1] = h[2 * i + 1]; h4[4 * i + 2] = h[2 * i]; h4[4 * i + 3] = h[2 * i + 1]; }
        CK(cudaMalloc((void**)&nco4, sizeof(int32_t) * 4 * n));
        CK(cudaMemcpy(nco4, h4.data(), sizeof(int32_t) * 4 * n, cudaMemcpyHostToDevice));
        op.nco = nco4;
    }
    launch_ofdm_demod(ctx->dev, op, ctx->fft_mode, ctx->stream);
    int rc = check_launch(ctx, "ofdm_demod_kernel");
    if (nco4) { cudaStreamSynchronize(ctx->stream); cudaFree(nco4); }
    return rc;
}

int dabb_find_index_ex(dabb_ctx* ctx, const float* iq, int64_t stride, const int64_t* win_start, int32_t n, int32_t placement, int32_t* index_out, float* cir_out)
{
    if (!ctx || !iq || !win_start || !index_out || n < 1 || placement < 0 || placement > 2) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    SyncParams sp{}; sp.iq = reinterpret_cast<const float2*>(iq); sp.stride = stride; sp.win_start = win_start; sp.nco = nullptr; sp.active = nullptr; sp.index_out = index_out; sp.cir_out = cir_out; sp.n = n; sp.fic_ratio = nullptr; sp.coarse_out = nullptr; sp.placement = placement; sp.freqsync = 0;
    // magnitudes between the transform and the search kernel: the context's buffer when the batch fits, else a temporary
    float* work = n <= ctx->S ? ctx->d_cir_work : nullptr;
    bool temp = false;
    if (!work) { if (cudaMalloc((void**)&work, (size_t)n * TU * sizeof(float)) != cudaSuccess) { ctx->err = "cudaMalloc(findIndex work buffer)"; return DABB_E_NOMEM; } temp = true; }
    sp.cir_work = work; sp.search_generic = ctx->search_generic;
    launch_find_index(ctx->dev, sp, ctx->fft_mode, ctx->stream);
    int rc = check_launch(ctx, "find_index_kernel");
    if (temp) { cudaStreamSynchronize(ctx->stream); cudaFree(work); }
    return rc;
}

int dabb_find_index(dabb_ctx* ctx, const float* iq, int64_t stride, const int64_t* win_start, int32_t n, int32_t* index_out, float* cir_out)
{
    return dabb_find_index_ex(ctx, iq, stride, win_start, n, DABB_PLACEMENT_THRESHOLD_BEFORE_PEAK, index_out, cir_out);
}

int dabb_coarse_estimate(dabb_ctx* ctx, const float* iq, int64_t stride, const int64_t* prs_start, int32_t n, int32_t method, int32_t* offset_out)
{
    if (!ctx || !iq || !prs_start || !offset_out || n < 1 || method < 0 || method > 2) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    launch_coarse(ctx->dev, reinterpret_cast<const float2*>(iq), stride, prs_start, n, method, offset_out, ctx->stream);
    return check_launch(ctx, "coarse_kernel");
}

int dabb_get_info(dabb_ctx* ctx, int32_t what, int64_t* out)
{
    if (!ctx || !out) return DABB_E_ARG;
    switch (what) {
        case DABB_INFO_OSC_MODE: *out = ctx->dev.osc_mode; return DABB_OK;
        case DABB_INFO_OSC_MISMATCHES: *out = ctx->osc_mismatches; return DABB_OK;
        case DABB_INFO_OSC_PATCHED: *out = ctx->osc_patched; return DABB_OK;
        default: return DABB_E_ARG;
    }
}

int dabb_set_options(dabb_ctx* ctx, const dabb_options* opt)
{
    if (!ctx || !opt || opt->fft_placement < 0 || opt->fft_placement > 2 || opt->freqsync_method < 0 || opt->freqsync_method > 2) return DABB_E_ARG;
    ctx->disable_coarse = opt->disable_coarse != 0; ctx->placement = opt->fft_placement; ctx->freqsync = opt->freqsync_method; ctx->decode_tii = opt->decode_tii != 0;
    return DABB_OK;
}

// stage-level helper: the decoder kernel on n codewords of `frag` punctured softbits each (device, any alignment / values): the input
// is copied into a padded 16-byte pitched buffer with -128 mapped to -127 (same symbol after the reference's clamp)
static int run_stage_viterbi(dabb_ctx* ctx, const int8_t* soft, int n_cw, int64_t in_stride, int frag, const int16_t* map, int nsteps, int nbits,
                             const uint32_t* d_prbs_words, uint8_t* out, int64_t out_stride)
{
    std::vector<uint2> steps; std::vector<uint32_t> soff;
    build_vit_tables_u2(map, nsteps, steps, soff);
    if ((int)soff.back() > frag) { ctx->err = "de-puncturing map consumes more softbits than provided"; return DABB_E_ARG; }
    const int pitch = (frag + 15) & ~15;
    uint2* d_steps = nullptr; uint32_t* d_soff = nullptr; int8_t* tmp = nullptr;
    CK(cudaMalloc((void**)&d_steps, steps.size() * sizeof(uint2))); CK(cudaMalloc((void**)&d_soff, soff.size() * 4));
    CK(cudaMalloc((void**)&tmp, (size_t)n_cw * pitch + VIT_FRAG_SLACK));
    CK(cudaMemcpyAsync(d_steps, steps.data(), steps.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_soff, soff.data(), soff.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(tmp, 0, (size_t)n_cw * pitch + VIT_FRAG_SLACK, ctx->stream));
    int rc = 0;
    if (pitch == frag && in_stride == frag) { launch_clamp_copy(soft, tmp, (int64_t)n_cw * frag, ctx->stream); rc = check_launch(ctx, "clamp_copy_kernel"); }
    else {
        CK(cudaMemcpy2DAsync(tmp, pitch, soft, (size_t)in_stride, frag, n_cw, cudaMemcpyDeviceToDevice, ctx->stream));
        launch_clamp_copy(tmp, tmp, (int64_t)n_cw * pitch, ctx->stream); rc = check_launch(ctx, "clamp_copy_kernel");
    }
    if (!rc) rc = ensure_dec(ctx, vit_dec_bytes(n_cw, nsteps));
    if (!rc) {
        ViterbiParams vp{}; vp.frag = tmp; vp.cw_div = 1; vp.outer_stride = pitch; vp.inner_stride = 0; vp.steptab = d_steps; vp.stage_off = d_soff;
        vp.n_cw = n_cw; vp.nsteps = nsteps; vp.nbits = nbits; vp.dec = ctx->d_dec; vp.out = out; vp.out_stride = out_stride; vp.prbs_words = d_prbs_words; vp.valid = nullptr;
        launch_viterbi(vp, ctx->stream);
        rc = check_launch(ctx, "viterbi_kernel");
    }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_steps); cudaFree(d_soff); cudaFree(tmp);
    return rc;
}

int dabb_viterbi(dabb_ctx* ctx, const int8_t* soft, int32_t n_cw, int32_t nbits, uint8_t* bits_out)
{
    if (!ctx || !soft || !bits_out || n_cw < 1 || nbits < 32 || (nbits % 32) || ((nbits + 6) % 6)) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    const int nsteps = nbits + 6, ob = nbits / 8;
    uint8_t* bytes = nullptr;
    CK(cudaMalloc((void**)&bytes, (size_t)n_cw * ob));
    std::vector<int16_t> map((size_t)nsteps * 4);
    std::fill(map.begin(), map.end(), (int16_t)0);                      // >= 0 = present: nothing punctured, the caller's zeros are softbit 0
    int rc = run_stage_viterbi(ctx, soft, n_cw, (int64_t)nsteps * 4, nsteps * 4, map.data(), nsteps, nbits, nullptr, bytes, ob);
    if (!rc) { launch_unpack_bits(bytes, ob, n_cw, nbits, bits_out, ctx->stream); rc = check_launch(ctx, "unpack_bits_kernel"); }
    cudaStreamSynchronize(ctx->stream);
    cudaFree(bytes);
    return rc;
}

int dabb_fic_decode(dabb_ctx* ctx, const int8_t* soft, int32_t n_frames, uint8_t* fib_out, int32_t* crc_mask_out)
{
    if (!ctx || !soft || !fib_out || !crc_mask_out || n_frames < 1) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    // codeword (f, b) = softbits [9216 f + 2304 b, +2304)
    int rc = run_stage_viterbi(ctx, soft, n_frames * 4, 2304, 2304, ctx->host->fic_map, 774, 768, ctx->d_fic_prbs_words, fib_out, 96);
    if (!rc) { launch_fic_crc(fib_out, nullptr, n_frames, crc_mask_out, ctx->stream); rc = check_launch(ctx, "fic_crc_kernel"); }
    cudaStreamSynchronize(ctx->stream);
    return rc;
}

int dabb_msc_decode(dabb_ctx* ctx, const dabb_subchannel* sc, const int8_t* soft, int32_t n, uint8_t* bytes_out)
{
    if (!ctx || !sc || !soft || !bytes_out || n < 1) return DABB_E_ARG;
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    ProtProfile prof;
    if (make_prot_profile(sc->bitrate, sc->short_form, sc->uep_level, sc->eep_profile_a, sc->eep_level, prof) < 0) { ctx->err = "unsupported protection profile"; return DABB_E_UNSUPPORTED; }
    const int frag = sc->length_cu * 64, nbits = 24 * prof.bitrate, nsteps = nbits + 6, flen = 3 * prof.bitrate;
    if (prof.in_bits > frag) { ctx->err = "protection profile needs more bits than the sub-channel holds"; return DABB_E_ARG; }
    std::vector<int16_t> map((size_t)nsteps * 4);
    build_msc_map(*ctx->host, prof, map.data());
    std::vector<uint32_t> w; pack_prbs_words(ctx->host->prbs, nbits, w);
    uint32_t* d_w = nullptr;
    CK(cudaMalloc((void**)&d_w, w.size() * 4));
    CK(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
    // output rows are 3*bitrate bytes (a multiple of 4)
    int rc = run_stage_viterbi(ctx, soft, n, frag, frag, map.data(), nsteps, nbits, d_w, bytes_out, flen);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_w);
    return rc;
}

int dabb_rs_superframes(dabb_ctx* ctx, uint8_t* sf, int32_t n, int32_t sf_len, int32_t* info)
{
    if (!ctx || !sf || !info || n < 1 || sf_len < 120 || (sf_len % 120)) return DABB_E_ARG;
    if (sf_len / 120 > 64) { ctx->err = "superframes above 512 kbit/s (64 interleaved codewords) are not supported"; return DABB_E_UNSUPPORTED; }
    cudaSetDevice(ctx->device);
    sync_all(ctx);
    launch_rs_superframes(sf, n, sf_len, info, ctx->dev.gf_exp, ctx->dev.gf_log, ctx->stream);
    return check_launch(ctx, "rs_superframes_kernel");
}

int dabb_dev_alloc(dabb_ctx* ctx, size_t bytes, void** out) { if (!ctx || !out) return DABB_E_ARG; cudaSetDevice(ctx->device); CK(cudaMalloc(out, bytes)); return 0; }
int dabb_dev_free(dabb_ctx* ctx, void* p) { if (!ctx) return DABB_E_ARG; cudaSetDevice(ctx->device); CK(cudaStreamSynchronize(ctx->stream)); CK(cudaFree(p)); return 0; }
int dabb_memcpy_h2d(dabb_ctx* ctx, void* d, const void* s, size_t n) { if (!ctx) return DABB_E_ARG; cudaSetDevice(ctx->device); CK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, ctx->stream)); CK(cudaStreamSynchronize(ctx->stream)); return 0; }
int dabb_memcpy_d2h(dabb_ctx* ctx, void* d, const void* s, size_t n) { if (!ctx) return DABB_E_ARG; cudaSetDevice(ctx->device); CK(cudaStreamSynchronize(ctx->stream)); CK(cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost)); return 0; }

} // extern "C"
