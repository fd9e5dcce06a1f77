"""GPU parity tests, closed loop: dabb_process() (time sync -> OFDM -> FIC -> MSC -> RS) frame by frame for several
streams against the CPU oracle's closed-loop receiver (oracle.rx_run, itself pinned bit-exact against the reference's
RadioReceiver in tests/test_oracle_vs_ref.py).  FIBs, logical-frame bytes, RS statistics, startIndex and the fine
corrector must be identical; the softbits of every frame too (DABB_FFT_EXACT)."""
import numpy as np
import pytest

import dabtx
from conftest import load_pkg

pytestmark = pytest.mark.gpu
TF = 196608


def run_gpu(pkg, sigs, bitrate=96, cu=72, level=3, n_slots=1, fft_mode=0, disable_coarse=True, fft_placement=0, freqsync_method=0, taps=None):
    S = len(sigs)
    n = max(len(s) for s in sigs)
    buf = np.zeros((S, n), np.complex64)
    for i, s in enumerate(sigs):
        buf[i, :len(s)] = s
    ctx = pkg.Context(n_streams=S, keep_taps=True, n_subch_slots=n_slots, fft_mode=fft_mode, disable_coarse=disable_coarse,
                      fft_placement=fft_placement, freqsync_method=freqsync_method)
    d = ctx.dev(buf)
    res = [dict(info=[], fibs=[], crc=[], msc=[], rs=[], soft=[], sf=[]) for _ in range(S)]
    selected = False
    lens = [len(s) for s in sigs]
    for step in range(200):
        out = ctx.process(d, n, np.zeros(S, np.int64), n, msc_stride=3 * bitrate, sf_stride=15 * bitrate)
        r = out["results"]
        soft = ctx.read_tap(0)
        if taps is not None:
            taps.append((r.copy(), ctx.read_tap(1), ctx.read_tap(2), ctx.read_tap(3)))
        decoded = 0
        for i in range(S):
            if r["status"][i] != pkg.FRAME_DECODED:
                continue
            if r["next_pos"][i] > lens[i]:
                continue          # ran into the zero padding: the oracle stops before this frame too
            decoded += 1
            res[i]["info"].append((int(r["start_index"][i]), int(r["fine_corr"][i]), int(r["coarse_corr"][i])))
            res[i]["fibs"].append(out["fibs"][i].copy()); res[i]["crc"].append(int(r["fib_crc_mask"][i]))
            res[i]["soft"].append(soft[i].copy())
            nl = int(r["n_logical"][i][0])
            # logical frames are produced by the last nl CIFs of the frame
            for c in range(4 - nl, 4):
                res[i]["msc"].append(out["msc"][i, 0, c, :3 * bitrate].copy())
            for e in range(int(r["n_rs_events"][i][0])):
                res[i]["rs"].append(((int(r["rs_uncorr_mask"][i][0]) >> e) & 1, int(r["rs_corr"][i][0][e])))
            if r["sf_ready"][i][0]:
                res[i]["sf"].append((out["sf"][i, 0, :15 * bitrate].copy(), int(r["sf_au_count"][i][0]), int(r["sf_au_crc_mask"][i][0])))
        if not selected and decoded:
            ctx.select_subchannel(0, cu, bitrate, eep_profile_a=True, eep_level=level)
            selected = True
        if decoded == 0 and step > 3:
            break
    ctx.close()
    return res


def compare(res, orc, name):
    n = min(len(res["info"]), orc["frames"])
    assert n >= orc["frames"] - 1 and n > 5, (name, len(res["info"]), orc["frames"])
    for f in range(n):
        assert res["info"][f][0] == orc["info"][f]["start_index"], (name, f, res["info"][f], orc["info"][f])
        assert res["info"][f][1] == orc["info"][f]["fine"], (name, f, res["info"][f], orc["info"][f])
        if "soft" in orc and f < len(orc["soft"]):
            assert np.array_equal(res["soft"][f], orc["soft"][f]), (name, f, int((res["soft"][f] != orc["soft"][f]).sum()))
        ofib = orc["fibs"][12 * f: 12 * f + 12]
        assert np.array_equal(res["fibs"][f], ofib[:, 1:]), (name, f)
        assert res["crc"][f] == int(sum(int(o) << k for k, o in enumerate(ofib[:, 0]))), (name, f)
    msc = np.concatenate(res["msc"]) if res["msc"] else np.zeros(0, np.uint8)
    m = min(len(msc), len(orc["msc"]))
    assert m > 0 and np.array_equal(msc[:m], orc["msc"][:m]), name
    k = min(len(res["rs"]), len(orc["rs"]))
    assert k > 0 and [tuple(x) for x in orc["rs"][:k].tolist()] == [tuple(x) for x in res["rs"][:k]], name


def test_closed_loop_matches_oracle(oracle):
    pkg = load_pkg()
    prot = oracle.prot_eep(96, 1, 3)
    sigs = []
    tx0 = dabtx.DabTx(seed=0xDAB); sigs.append(tx0.frames(26))
    tx1 = dabtx.DabTx(seed=0xDAC); s1 = tx1.frames(26); sigs.append(dabtx.add_awgn(s1, 14.0, seed=0x5EED, signal_power=float(np.mean(np.abs(s1[3000:190000]) ** 2))))
    tx2 = dabtx.DabTx(seed=0xDAD); s2 = tx2.frames(26); sigs.append(np.concatenate([np.zeros(1000, np.complex64) + 1e-6, s2])[:len(s2)])   # time offset
    res = run_gpu(pkg, sigs)
    for i, sig in enumerate(sigs):
        orc = oracle.rx_run(sig, prot=prot, start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True, want_soft=30)
        compare(res[i], orc, f"stream{i}")
    # the clean stream must also reproduce what was transmitted
    lf = np.concatenate(tx0.logical)
    msc = np.concatenate(res[0]["msc"])
    hits = [k for k in range(0, 40) if np.array_equal(msc[:288], lf[k * 288:(k + 1) * 288])]
    assert hits and np.array_equal(msc, lf[hits[0] * 288: hits[0] * 288 + len(msc)])
    assert all(c == 0xFFF for c in res[0]["crc"])
    assert len(res[0]["sf"]) >= 2 and all(m == 0x3F and n == 6 for _, n, m in res[0]["sf"])
    sf_bytes = [s for s, _, _ in res[0]["sf"]]
    txsf = [bytes(s) for s in tx0.superframes]
    assert all(bytes(s) in txsf for s in sf_bytes)


def test_coarse_corrector_matches_oracle(oracle):
    """RadioReceiverOptions::disableCoarseCorrector = false: processPRS (PatternOfZeros) runs while the FIC success counter is
    low.  A clean stream must stay at coarse 0; a stream shifted by +2 carriers must be pulled back exactly like the oracle."""
    pkg = load_pkg()
    prot = oracle.prot_eep(96, 1, 3)
    tx0 = dabtx.DabTx(seed=0xC0A); s0 = tx0.frames(14)
    tx1 = dabtx.DabTx(seed=0xC0B); s1 = dabtx.freq_shift(tx1.frames(14), 2000.0)
    sigs = [s0, s1]
    res = run_gpu(pkg, sigs, disable_coarse=False)
    for i, sig in enumerate(sigs):
        orc = oracle.rx_run(sig, prot=prot, start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=False)
        n = min(len(res[i]["info"]), orc["frames"])
        assert n >= 8
        assert [x[2] for x in res[i]["info"][:n]] == [x["coarse"] for x in orc["info"][:n]], (i, res[i]["info"][:n], [(x["fine"], x["coarse"]) for x in orc["info"][:n]])
        assert all(abs(a[1] - b["fine"]) <= 1 for a, b in zip(res[i]["info"][:n], orc["info"][:n]))
        crc_o = [int(sum(int(o) << k for k, o in enumerate(orc["fibs"][12 * f: 12 * f + 12, 0]))) for f in range(n)]
        assert res[i]["crc"][:n] == crc_o
    assert all(c == 0 for _, _, c in res[0]["info"])
    assert res[1]["info"][-1][2] != 0 and res[1]["crc"][-1] == 0xFFF


def test_raw_sample_formats(oracle):
    """u8 / s16 RAW-file sample formats converted on the device (CRAWFile::convertSamples semantics) through the host-buffer
    path of dabb_process: FIBs must equal the oracle fed with the CPU-converted float samples."""
    pkg = load_pkg()
    tx = dabtx.DabTx(seed=0xF0F)
    sig = tx.frames(9)
    inter = np.stack([sig.real, sig.imag], axis=-1)
    u8 = np.clip(np.round(inter * 2.0 * 128.0 + 128.0), 0, 255).astype(np.uint8)
    f_u8 = ((u8.astype(np.float32) - 128.0) / 128.0).view(np.complex64).reshape(-1)
    s16 = np.clip(np.round(inter * 20000.0), -32768, 32767).astype(np.int16)
    f_s16 = s16.astype(np.float32).view(np.complex64).reshape(-1)
    s16_bytes = s16.astype(">i2").view(np.uint8).reshape(len(sig), 4)       # the reference's "s16le" reader takes the first byte as the high byte
    for fmt, raw, ref_sig in ((pkg.IQ_U8, u8, f_u8), (pkg.IQ_S16LE, s16_bytes, f_s16)):
        ctx = pkg.Context(n_streams=1)
        raw = np.ascontiguousarray(raw)
        fibs, crcs = [], []
        for step in range(9):
            out = ctx.process(raw, len(sig), np.zeros(1, np.int64), len(sig), iq_is_host=True, iq_format=fmt)
            r = out["results"]
            if r["status"][0] == pkg.FRAME_DECODED and r["next_pos"][0] <= len(sig):
                fibs.append(out["fibs"][0].copy()); crcs.append(int(r["fib_crc_mask"][0]))
        ctx.close()
        o = oracle.rx_run(ref_sig, disable_coarse=True)
        n = min(len(fibs), o["frames"])
        assert n >= 6
        for f in range(n):
            assert np.array_equal(fibs[f], o["fibs"][12 * f:12 * f + 12, 1:]) and crcs[f] == 0xFFF, (fmt, f)


def test_two_subchannel_slots_and_remove(oracle):
    """two sub-channels decoded at once (slot 0: the DAB+ EEP-3A service; slot 1: a 64 kbit/s UEP-3 DAB/MP2 sub-channel placed on
    the random-filled capacity units 200..247), then slot 1 removed.  Slot 1's logical frames must equal the oracle's
    time-de-interleaver + UEP Viterbi + dispersal applied to the softbits the GPU itself produced."""
    pkg = load_pkg()
    tx = dabtx.DabTx(seed=0x2222)
    sig = tx.frames(12)
    ctx = pkg.Context(n_streams=1, keep_taps=True, n_subch_slots=2)
    d = ctx.dev(sig.reshape(1, -1))
    n = len(sig)
    prot = oracle.prot_uep(64, 3)
    cu1, start1 = 48, 200
    cifs, got1, got0 = [], [], []
    selected = removed = False
    for step in range(12):
        out = ctx.process(d, n, np.zeros(1, np.int64), n, msc_stride=288)
        r = out["results"]
        if r["status"][0] != pkg.FRAME_DECODED or r["next_pos"][0] > n:
            continue
        if selected and not removed:
            soft = ctx.read_tap(0)[0]
            for c in range(4):
                cifs.append(soft[3 + 18 * c: 21 + 18 * c].reshape(-1)[start1 * 64:(start1 + cu1) * 64].copy())
            for c in range(4 - int(r["n_logical"][0][1]), 4):
                got1.append(out["msc"][0, 1, c, :192].copy())
        if selected:
            for c in range(4 - int(r["n_logical"][0][0]), 4):
                got0.append(out["msc"][0, 0, c, :288].copy())
            assert int(r["n_rs_events"][0][1]) == 0          # DAB (MP2) sub-channel: no superframe / RS handling
        if not selected:
            ctx.select_subchannel(0, 72, 96, eep_profile_a=True, eep_level=3, dabplus=True, slot=0)
            ctx.select_subchannel(start1, cu1, 64, short_form=True, uep_level=3, dabplus=False, slot=1)
            selected = True
        elif len(cifs) >= 28 and not removed:
            ctx.remove_subchannel(slot=1); removed = True
    ctx.close()
    de = oracle.deinterleave(np.stack(cifs))
    exp = [oracle.pack_bits(oracle.msc_deconvolve(prot, x[:prot.in_bits], True)) for x in de]
    assert len(got1) == len(exp) >= 8
    for a, b in zip(got1, exp):
        assert np.array_equal(a, b)
    lf = np.concatenate(tx.logical); g0 = np.concatenate(got0)
    hits = [k for k in range(40) if np.array_equal(g0[:288], lf[k * 288:(k + 1) * 288])]
    assert hits and np.array_equal(g0, lf[hits[0] * 288: hits[0] * 288 + len(g0)])


LOCKING_STREAMS = {0: [(0x51, 2000.0), (4, 1000.0), (8, -1000.0)], 1: [(3, -3000.0), (4, 1000.0), (7, 5000.0)], 2: [(4, 1000.0), (8, -1000.0)]}


@pytest.mark.parametrize("method", [0, 1, 2])
def test_coarse_methods_closed_loop(oracle, method):
    """all three FreqsyncMethods in the closed loop on frequency-shifted streams on which the estimator locks (most offsets do not
    lock with PatternOfZeros / CorrelatePRS - in the reference either: their arg() terms are truncated to integers, see
    oracle/dab_oracle.c): start index, coarse corrector and FIB CRC mask of every frame like the oracle, which is pinned to the
    reference RadioReceiver with the same option (tests/test_oracle_vs_ref.py); fine corrector within 1 Hz."""
    pkg = load_pkg()
    sigs = [dabtx.freq_shift(dabtx.DabTx(seed=sd).frames(12), hz) for sd, hz in LOCKING_STREAMS[method]]
    res = run_gpu(pkg, sigs, disable_coarse=False, freqsync_method=method)
    for i, sig in enumerate(sigs):
        orc = oracle.rx_run(sig, disable_coarse=False, freqsync_method=method)
        n = min(len(res[i]["info"]), orc["frames"])
        assert n >= 8
        crc_o = [int(sum(int(o) << k for k, o in enumerate(orc["fibs"][12 * f: 12 * f + 12, 0]))) for f in range(n)]
        assert all(c == 0xFFF for c in crc_o[-4:]) and orc["info"][n - 1]["coarse"] == int(LOCKING_STREAMS[method][i][1])
        msg = (i, res[i]["info"][:n], [(x["start_index"], x["fine"], x["coarse"]) for x in orc["info"][:n]])
        assert [x[2] for x in res[i]["info"][:n]] == [x["coarse"] for x in orc["info"][:n]], msg
        assert [x[0] for x in res[i]["info"][:n]] == [x["start_index"] for x in orc["info"][:n]], msg
        assert all(abs(a[1] - b["fine"]) <= 1 for a, b in zip(res[i]["info"][:n], orc["info"][:n])), msg
        assert res[i]["crc"][:n] == crc_o, msg


@pytest.mark.parametrize("placement", [1, 2])
def test_other_placements_closed_loop(oracle, placement):
    pkg = load_pkg()
    tx = dabtx.DabTx(seed=0x61); s0 = tx.frames(10)
    s1 = np.concatenate([np.zeros(777, np.complex64), dabtx.add_awgn(dabtx.DabTx(seed=0x62).frames(10), 12.0, seed=9)])
    sigs = [s0, s1]
    res = run_gpu(pkg, sigs, fft_placement=placement)
    prot = oracle.prot_eep(96, 1, 3)
    for i, sig in enumerate(sigs):
        orc = oracle.rx_run(sig, prot=prot, start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True, fft_placement=placement)
        compare(res[i], orc, f"placement{placement}/{i}")


def test_diagnostic_taps(oracle):
    """taps behind onNewImpulseResponse / onConstellationPoints / onNewNullSymbol: CIR = oracle findIndex CIR, constellation = every
    96th r1 of the oracle demap of the same frame, null symbol = the 2656 samples after the frame times the corrected oscillator"""
    pkg = load_pkg()
    TU, TS, TNULL = 2048, 2552, 2656
    s0 = dabtx.DabTx(seed=0x71).frames(8)
    s1 = dabtx.freq_shift(dabtx.DabTx(seed=0x72).frames(8), 137.0)           # fine corrector becomes non-zero
    sigs = [s0, s1]
    taps = []
    res = run_gpu(pkg, sigs, taps=taps)
    checked = 0
    for i, sig in enumerate(sigs):
        orc = oracle.rx_run(sig, disable_coarse=True)
        f = 0
        for (r, cir, con, nul) in taps:
            if r["status"][i] != pkg.FRAME_DECODED or r["next_pos"][i] > len(sig) or f >= orc["frames"]:
                continue
            info = orc["info"][f]; f += 1
            assert int(r["start_index"][i]) == info["start_index"]
            null_start = int(r["next_pos"][i]) - TNULL
            prs0 = null_start - 75 * TS - TU
            if i == 0:
                # NCO idle: the taps are plain functions of the input samples
                s_o, r1 = oracle.demod_frame(sig[prs0: prs0 + TU], sig[prs0 + TU: prs0 + TU + 75 * TS], True)
                assert np.array_equal(con[i].view(np.uint32), np.ascontiguousarray(r1[:, ::96]).view(np.uint32))
                assert np.array_equal(nul[i], sig[null_start: null_start + TNULL])
                _, c_o = oracle.find_index(sig[info["pos"]: info["pos"] + TU])
                assert np.array_equal(cir[i].view(np.uint32), c_o.view(np.uint32))
            else:
                # corrected oscillator: unit-modulus rotation whose phase falls by 2 pi (coarse + fine) / 2 048 000 per sample
                raw = sig[null_start: null_start + TNULL]
                rot = nul[i][100:] * np.conj(raw[100:])
                assert np.allclose(np.abs(nul[i]), np.abs(raw), rtol=1e-5, atol=1e-7)
                step = np.angle(np.sum(rot[1:] * np.conj(rot[:-1])))
                hz = int(r["fine_corr"][i]) + int(r["coarse_corr"][i])
                assert abs(step + 2 * np.pi * hz / 2048000.0) < 1e-6, (step, hz)
            checked += 1
    assert checked >= 10
