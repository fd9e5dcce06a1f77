/*
 * tii.h — transmitter identification (TII) analysis of the host glue, after TIIDecoder (backend/tii-decoder.cpp:196-384).
 *
 * The two 2048-point transforms TIIDecoder::run starts from (phase reference symbol, last T_u samples of the null symbol) are made on
 * the GPU (tii_spectra_kernel, tap 4 of dabb_read_tap) with the reference's own FFT arithmetic; everything after them is a few hundred
 * scalar operations per frame with data-dependent containers and lives here, on the host, like in the reference:
 *   - pairs of adjacent null-symbol carriers are multiplied (first x conj(second)) and summed over the four carrier blocks (:236-265);
 *     a pair counts as active when |sum| > 0.4 |PRS carrier|^2 (:274-283)
 *   - every active pair votes for the (comb, pattern) combinations that contain it (table of EN 300 401 clause 14.8.1, :30-101,
 *     148-158); combinations with >= 4 votes are analysed unless there are 10 or more of them (:285-308)
 *   - analyse_phase (:318-383): for every delay hypothesis err = -4 .. 499 samples the absolute phase error of the combination's 32
 *     carriers against the PRS is accumulated into a table (truncated to an integer at every addition, like the reference's
 *     uint64_t += float); after five measurements the smallest entry is reported through onTIIMeasurement and the table is cleared
 * The float arithmetic (std::complex product, arg, polar, abs) and the containers (std::unordered_map with the reference's key types
 * and hash, filled in the reference's order) are the reference's, so that on the same spectra the same measurements come out in the
 * same order, ties included.  Unlike the reference, whose decoder thread skips frames while it is busy, every frame is analysed.
 */
#ifndef DABB_HOST_TII_H
#define DABB_HOST_TII_H
#include "dab_api.h"
#include <algorithm>
#include <cmath>
#include <functional>
#include <unordered_map>
#include <unordered_set>

namespace dabb_host {

struct TiiCombPattern {
    int comb = 0, pattern = 0;
    bool operator==(const TiiCombPattern& o) const { return comb == o.comb && pattern == o.pattern; }
};
struct TiiCombPatternHash { std::size_t operator()(const TiiCombPattern& cp) const noexcept { return (std::size_t)(cp.comb * 100 + cp.pattern); } };

class TiiAnalyzer {
public:
    TiiAnalyzer()
    {
        /* the 70 patterns: the 8-bit words with exactly four bits set, in ascending order, first position = most significant bit */
        int n = 0;
        for (int w = 0; w < 256 && n < 70; w++) {
            int ones = 0;
            for (int b = 0; b < 8; b++) ones += (w >> b) & 1;
            if (ones != 4) continue;
            for (int b = 0; b < 8; b++) pat_[n][b] = (w >> (7 - b)) & 1;
            n++;
        }
        for (int c = 0; c < 24; c++)
            for (int p = 0; p < 70; p++)
                for (int k = 0; k < 384; k++)
                    for (int b = 0; b < 8; b++)
                        if (k == 1 + 2 * c + 48 * b && pat_[p][b]) cpPerCarrier_[k].insert(TiiCombPattern{c, p});
    }

    /* carriers of one combination, sorted (CombPattern::generateCarriers, :108-131) */
    std::vector<int> carriersOf(const TiiCombPattern& cp) const
    {
        std::vector<int> v;
        v.reserve(32);
        for (int k = 0; k < 384; k++)
            for (int b = 0; b < 8; b++)
                if (k == 1 + 2 * cp.comb + 48 * b && pat_[cp.pattern][b]) {
                    v.push_back(k - 769); v.push_back(k - 769 + 1); v.push_back(k - 385); v.push_back(k - 385 + 1);
                    v.push_back(k); v.push_back(k + 1); v.push_back(k + 384); v.push_back(k + 384 + 1);
                }
        std::sort(v.begin(), v.end());
        return v;
    }

    /* one frame: spectra in natural FFT bin order (2048 bins each) */
    void process(const complexf* nullSpec, const complexf* prsSpec, const std::function<void(tii_measurement_t&&)>& emit)
    {
        std::vector<complexf> mult(192);
        std::vector<float> prsPow(192);
        for (size_t i = 0; i < 192; i++) prsPow[i] = std::norm(prsSpec[1 + 2 * i]);
        const size_t kStart[] = {2048 - 768, 2048 - 384, 1, 385};
        for (size_t k : kStart)
            for (size_t i = 0; i < 192; i++) mult[i] += nullSpec[k + 2 * i] * std::conj(nullSpec[k + 2 * i + 1]);
        std::vector<int> carriers;
        for (size_t i = 0; i < 192; i++)
            if (std::abs(mult[i]) > prsPow[i] * 0.4f) carriers.push_back((int)(i * 2 + 1));
        std::unordered_map<TiiCombPattern, int, TiiCombPatternHash> votes;
        for (const int k : carriers)
            if (cpPerCarrier_.count(k))
                for (const auto& cp : cpPerCarrier_[k]) votes[cp]++;
        size_t likely = 0;
        for (const auto& v : votes) if (v.second >= 4) likely++;
        if (likely >= 10) return;            /* the threshold was off: skip the frame (:298-300) */
        for (const auto& v : votes) if (v.second >= 4) analysePhase(v.first, nullSpec, prsSpec, emit);
    }

    void clear() { meas_.clear(); }

private:
    struct Meas { std::unordered_map<float, uint64_t> errorPerCorrection; size_t n = 0; };

    void analysePhase(const TiiCombPattern& cp, const complexf* n, const complexf* p, const std::function<void(tii_measurement_t&&)>& emit)
    {
        const std::vector<int> carriers = carriersOf(cp);
        auto ix = [](int k) { return k < 0 ? 2048 + k : k; };
        std::vector<float> phasesPrs(carriers.size());
        for (size_t i = 0; i + 1 < carriers.size(); i += 2) { phasesPrs[i] = std::arg(p[ix(carriers[i])]); phasesPrs[i + 1] = phasesPrs[i]; }
        Meas& m = meas_[cp];
        for (int err = -4; err < 500; err++) {
            float absErr = 0;
            for (size_t j = 0; j < carriers.size(); j++) {
                constexpr float pi = (float)M_PI;
                const complexf rot = std::polar(1.0f, 2.0f * pi * err * carriers[j] / 2048.0f);
                const float delta = std::arg(n[ix(carriers[j])] * rot) - phasesPrs[j];
                absErr += std::abs(delta);
            }
            m.errorPerCorrection[(float)err] += absErr;       /* uint64_t += float: truncates at every addition, like the reference */
        }
        if (++m.n >= 5) {
            auto best = std::min_element(m.errorPerCorrection.begin(), m.errorPerCorrection.end(),
                                         [](const std::pair<const float, uint64_t>& a, const std::pair<const float, uint64_t>& b) { return a.second < b.second; });
            if (best != m.errorPerCorrection.end()) {
                tii_measurement_t r;
                r.error = (float)best->second; r.delay_samples = (int)best->first; r.comb = cp.comb; r.pattern = cp.pattern;
                emit(std::move(r));
            }
            m.errorPerCorrection.clear();
            m.n = 0;
        }
    }

    int pat_[70][8];
    std::unordered_map<int, std::unordered_set<TiiCombPattern, TiiCombPatternHash>> cpPerCarrier_;
    std::unordered_map<TiiCombPattern, Meas, TiiCombPatternHash> meas_;
};

} // namespace dabb_host
#endif
