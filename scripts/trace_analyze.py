"""CTA timeline written with DABB_TRACE=<file> (kind sm start_ns end_ns per line; kind 1 = ofdm_demod CTA, 2 / 3 = FIC / MSC Viterbi forward
pass of a CTA's thread 0, 12 / 13 = the same CTA at the end of its traceback): how much the two lanes really overlap."""
import sys
import numpy as np
d = np.loadtxt(sys.argv[1], dtype=np.int64)
kind, sm, t0, t1 = d[:, 0], d[:, 1], d[:, 2], d[:, 3]
base = t0.min()
t0 = (t0 - base) / 1e6; t1 = (t1 - base) / 1e6          # ms
print(f"{len(d)} records over {t1.max():.2f} ms, SMs {sm.min()}..{sm.max()}")
# steps = bursts of OFDM CTAs: split the OFDM start times at gaps > 0.2 ms
o = np.sort(t0[kind == 1])
cuts = np.nonzero(np.diff(o) > 0.15)[0]
starts = np.concatenate([[o[0]], o[cuts + 1]]); ends = np.concatenate([o[cuts], [o[-1]]])
print(f"{len(starts)} OFDM launches")
for name, k in (("ofdm", 1), ("fic fwd", 2), ("msc fwd", 3), ("fic all", 12), ("msc all", 13)):
    m = kind == k
    if m.any():
        dur = t1[m] - t0[m]
        print(f"  {name:8s} n={m.sum():7d} CTA duration ms: median {np.median(dur):.3f} p10 {np.percentile(dur, 10):.3f} p90 {np.percentile(dur, 90):.3f}")
# steady-state window: the pipelined steps are the ones whose OFDM launches are < 6 ms apart; take launches 9..12 (after the 8 warm-up steps)
sel = range(9, min(13, len(starts)))
for i in sel:
    w0 = starts[i]; w1 = starts[i + 1] if i + 1 < len(starts) else ends[i] + 1.0
    grid = np.linspace(w0, w1, 400, endpoint=False)
    def resident(k):
        m = np.isin(kind, k); a, b = t0[m], t1[m]
        return np.array([((a <= g) & (b > g)).sum() for g in grid])
    ro, rv = resident([1]), resident([12, 13])
    both = ((ro > 0) & (rv > 0)).mean()
    print(f"step from OFDM launch {i}: period {w1 - w0:.3f} ms | OFDM CTAs resident avg {ro.mean():.0f} (max {ro.max()}) | Viterbi CTAs resident avg {rv.mean():.0f} (max {rv.max()}) | "
          f"time with both kinds resident {100 * both:.0f} % | OFDM-only {100 * ((ro > 0) & (rv == 0)).mean():.0f} % | Viterbi-only {100 * ((ro == 0) & (rv > 0)).mean():.0f} % | neither {100 * ((ro == 0) & (rv == 0)).mean():.0f} %")
    # coarse timeline: 20 bins
    bins = np.linspace(w0, w1, 21)
    line_o = [int(ro[(grid >= bins[j]) & (grid < bins[j + 1])].mean()) for j in range(20)]
    line_v = [int(rv[(grid >= bins[j]) & (grid < bins[j + 1])].mean()) for j in range(20)]
    print("    OFDM CTAs   :", " ".join(f"{x:4d}" for x in line_o))
    print("    Viterbi CTAs:", " ".join(f"{x:4d}" for x in line_v))
