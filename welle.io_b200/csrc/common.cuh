// common.cuh — shared declarations of the libdab_b200 CUDA sources
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "ofdm_core.cuh"

namespace dabb {

// ---- device tables (built on the host by tables.cpp, uploaded once per context) ----
struct DevTables {
    const float2* tw_fwd;      // TwLayout::TOTAL entries, forward transform
    const float2* tw_inv;      // inverse transform
    const int16_t* invperm;    // [2048] bin -> logical carrier index 0..1535, -1 for unused bins (freq-interleaver.cpp:35-91);
                               // [2048 + bin] its index in ofdm_demod_kernel's softbit staging area, [4096 + c] the slot of logical chunk c there
    const float2* prs_ref;     // [2048] PhaseReference::refTable (phasereference.cpp:45-51)
    const float2* osc;         // [2 048 000] oscillator table (ofdm-processor.cpp:92-94)
    // the same values computed on the fly (osc_mode = 1): osc[m] == float(H[m >> 10] * exp(j theta (m & 1023))) in double, the small
    // rotation from its Taylor polynomial; verified for all 2 048 000 m at context creation (else osc_mode stays 0)
    const double2* osc_hi;     // [2000] exp(j 2 pi 1024 a / 2 048 000), correctly rounded doubles - except at the quarter turns, where
                               // the reference's own double angle is an ulp off and the factor is the table value itself
    double osc_theta;          // 2 pi / 2 048 000, correctly rounded
    int32_t osc_mode;          // 0: table lookups (a scattered 8-byte gather per sample), 1: on the fly
    const uint8_t* prbs;       // [9216+] energy-dispersal PRBS bits (fic-handler.cpp:62-71)
    const int16_t* fic_map;    // [3096] mother-code position -> index into the 2304 punctured softbits or -1
    const uint8_t* gf_exp;     // [512]
    const uint8_t* gf_log;     // [256]
};

// ---- optional CTA timeline (environment DABB_TRACE=<file>): every CTA of the two hot kernels records {kind, SM, start, end} (global
// nanosecond timer) so that the overlap of the two lanes can be drawn; off (nullptr) in normal operation
struct TraceBuf { unsigned long long* rec; unsigned int* count; unsigned int cap; };
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long trace_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int trace_smid() { unsigned int s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
__device__ __forceinline__ void trace_put(const TraceBuf& tb, unsigned int kind, unsigned long long t0)
{
    const unsigned int i = atomicAdd(tb.count, 1u);
    if (i < tb.cap) { tb.rec[3 * i] = ((unsigned long long)kind << 32) | trace_smid(); tb.rec[3 * i + 1] = t0; tb.rec[3 * i + 2] = trace_now(); }
}
#endif

// ---- OFDM demod launch parameters ----
struct OfdmParams {
    const float2* iq; int64_t stride;            // complex samples
    const int64_t* prs_start;                    // [n] first useful PRS sample, relative to iq + f*stride
    const int32_t* nco;                          // [n][4] or nullptr: {phase applied to PRS sample 0, Hz (PRS), phase extrapolated to index 0 for the data symbols, Hz}
    const int32_t* active;                       // [n] or nullptr: 0 = skip frame
    int8_t* soft; int64_t soft_stride;           // per frame stride in bytes (>= 75*3072)
    float2* r1;                                  // optional tap [n][75][1536]
    float2* freqcorr;                            // optional [n][fc_pitch] partial CP correlation sums
    float* level;                                // optional [n][fc_pitch] partial signal-level estimates (see advance_kernel in api.cu)
    int32_t* snr;                                // optional [n] get_snr value of the PRS
    int n_frames; int groups; int sym_per_cta;   // groups * sym_per_cta == 75
    int n_full; int tail_groups;                 // frames >= n_full are cut into tail_groups CTAs each (tail_groups divides 75; 1 / 0 = off)
    int fc_pitch;                                // freqcorr entries per frame (>= groups, tail_groups); 0 -> the larger of the two
    int nco_fast;                                // DABB_NCO_FAST: fp32 oscillator (tolerance mode)
    TraceBuf trace;                              // rec == nullptr: off
    unsigned int* work;                          // persistent launch: item counter (zeroed by the caller before the launch); nullptr: one CTA per item
    int smem_floor;                              // request at least this much dynamic shared memory (caps CTAs/SM so that other kernels fit beside it)
};

struct SyncParams {
    const float2* iq; int64_t stride;
    const int64_t* win_start;                    // [n]
    const int32_t* nco;                          // [n][2] or nullptr
    const int32_t* active;
    int32_t* index_out; float* cir_out; int n;
    float* cir_work;                             // [n][T_u] magnitudes between the transform kernel and the search kernel
    // coarse frequency corrector (OFDMProcessor::processPRS, PatternOfZeros): evaluated for streams whose FIC success
    // counter is below 5 (ofdm-processor.cpp:397); result = carrier offset, or 100 when not evaluated / no estimate
    const int32_t* fic_ratio; int32_t* coarse_out;
    int placement;      // DABB_PLACEMENT_*: 0 ThresholdBeforePeak, 1 StrongestPeak, 2 EarliestPeakWithBinning
    int freqsync;       // DABB_FREQSYNC_*: 0 PatternOfZeros, 1 GetMiddle, 2 CorrelatePRS
    int search_generic; // 1: ThresholdBeforePeak through the literal sliding-maximum search (find_search_kernel) instead of the warp-per-window form (A/B tests)
};

int ofdm_init_constants();     // per-device constants of ofdm.cu (call once after cudaSetDevice)
void launch_ofdm_demod(const DevTables& tb, const OfdmParams& p, int fft_mode, cudaStream_t st);
int ofdm_tail_frames(int n_frames);
void launch_find_index(const DevTables& tb, const SyncParams& p, int fft_mode, cudaStream_t st, int part = -1);
void launch_tii_spectra(const DevTables& tb, const float2* iq, int64_t stride, const int64_t* prs_start, const int32_t* nco_frame, const int32_t* active,
                        const float2* nulls, float2* out, int n, cudaStream_t st);
void launch_coarse(const DevTables& tb, const float2* iq, int64_t stride, const int64_t* prs_start, int n, int freqsync, int32_t* out, cudaStream_t st);

// ---- host-side table builders (tables.cpp) ----
struct HostTables {
    float2 tw_fwd[TwLayout::TOTAL], tw_inv[TwLayout::TOTAL];
    int16_t perm[KC]; int16_t invperm[TU + TU + 192];   // [0, T_u) bin -> logical carrier; [T_u, 2 T_u) bin -> softbit staging index; then the staging slot of each 16-byte chunk
    float2 prs_ref[TU];
    uint8_t prbs[16384];
    int16_t fic_map[3096];
    uint8_t gf_exp[512], gf_log[256];
    int8_t pcodes[24][32];
};
void build_host_tables(HostTables& t);
void build_osc_table(float2* osc /* INPUT_RATE entries */);
void build_osc_factors(const float2* osc_table, double2* hi /* 2000 */, double* theta, int* patched);
// compares the on-the-fly oscillator with the table for every index on the device; returns the number of differing entries
int launch_osc_verify(const DevTables& tb, cudaStream_t st);

// protection profile -> (L, PI) blocks; returns punctured length or -1
struct ProtProfile { int bitrate; int nblk; int L[4]; int PI[4]; int in_bits; };
int make_prot_profile(int bitrate, int short_form, int uep_level, int eep_profile_a, int eep_level, ProtProfile& out);
// mother-code map: for each of 4*(24*bitrate+6) positions the index into the punctured input or -1
void build_msc_map(const HostTables& t, const ProtProfile& p, int16_t* map);

} // namespace dabb
