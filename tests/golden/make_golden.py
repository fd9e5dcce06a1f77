"""Generates tests/golden/golden_v1.npz from the UNMODIFIED reference (oracle/_ref/libwelle_ref.so, built from
/root/reference by oracle/Makefile, KISS-FFT option).  Run in the container that has /root/reference:
    python tests/golden/make_golden.py
The fixtures travel with the repository; tests/test_oracle_golden.py checks the C oracle against them on any machine."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dabtx  # noqa: E402
from oracle.bind import Ref  # noqa: E402

TU, TS, TF, TNULL = 2048, 2552, 196608, 2656


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def main():
    r = Ref()
    g = {}
    rng = np.random.default_rng(20260922)
    g["perm"] = r.perm_table(); g["prs"] = r.prs_table(); g["pcodes"] = r.pcodes()
    x = (rng.standard_normal(2048) + 1j * rng.standard_normal(2048)).astype(np.complex64)
    g["fft_in"] = x; g["fft_out"] = r.fft(x); g["ifft_out"] = r.fft(x, True)
    # viterbi
    g["vit_soft"] = rng.integers(-128, 128, (4, (768 + 6) * 4)).astype(np.int8)
    g["vit_out"] = np.stack([r.viterbi(s, 768) for s in g["vit_soft"]])
    # one synthetic frame: findIndex, demod hash, FIC
    tx = dabtx.DabTx(seed=0xDAB)
    iq = tx.frames(14)
    g["tx_sha"] = sha(iq)          # guards against a numpy RNG / FFT change silently altering the synthetic input
    base = 3 * TF + TNULL
    v = iq[base - 199: base - 199 + TU]
    idx, cir = r.find_index(v)
    g["find_index"] = np.array([idx]); g["cir_sha"] = sha(cir); g["cir_head"] = cir[690:720]
    st = base + 305
    soft, r1 = r.demod_frame(iq[st: st + TU], iq[st + TU: st + TU + 75 * TS], True)
    g["soft_sha"] = sha(soft); g["soft_head"] = soft[:4, :64].copy(); g["r1_sha"] = sha(r1)
    g["fic_soft"] = soft[:3].reshape(-1).copy()
    fb, ok = r.fic_decode(g["fic_soft"])
    g["fib_bytes"] = np.packbits(fb, axis=1); g["fib_ok"] = ok
    # EEP / UEP
    for name, (br, pa, lv) in {"eep96a3": (96, 1, 3), "eep32b2": (32, 0, 2)}.items():
        n = {"eep96a3": 4608, "eep32b2": 21 * 64}[name]
        s = rng.integers(-127, 128, n).astype(np.int8)
        g[name + "_soft"] = s; g[name + "_out"] = np.packbits(r.eep_deconvolve(br, pa, lv, s, True))
    s = rng.integers(-127, 128, 2236).astype(np.int8)
    g["uep48_3_soft"] = s; g["uep48_3_out"] = np.packbits(r.uep_deconvolve(48, 3, s, True))
    # RS
    cws, res = [], []
    for i in range(12):
        data = rng.integers(0, 256, 110).astype(np.uint8)
        cw = np.concatenate([data, r.rs_encode(data)])
        ne = [0, 1, 3, 5, 6, 9][i % 6]
        pos = rng.choice(120, ne, replace=False); cw[pos] ^= rng.integers(1, 256, ne).astype(np.uint8)
        cnt, out, p = r.rs_decode_codeword(cw)
        cws.append(cw); res.append(np.concatenate([[cnt & 0xFF], out]))
    g["rs_cw"] = np.stack(cws); g["rs_res"] = np.stack(res).astype(np.uint8)
    g["rs_parity_kat"] = r.rs_encode(np.arange(110, dtype=np.uint8))
    g["crc"] = np.array([r.crc_fire(np.arange(9, dtype=np.uint8)), r.crc_ccitt(np.frombuffer(b"123456789", np.uint8))])
    # closed loop
    e = r.e2e(iq, disable_coarse=True, select_at_fib=24, dump_path="/tmp/golden_e2e.msc")
    g["e2e_fibs"] = e["fibs"]; g["e2e_msc"] = e["msc"]; g["e2e_rs"] = e["rs"]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")
    make_sync_options(r, iq)


def sync_windows(iq):
    """the windows golden_sync_v1.npz is quoted on: (name, 2048 samples) in a fixed order"""
    base = 3 * TF + TNULL
    out = []
    for hz in (0, 2000, -3000, 7000):
        sh = dabtx.freq_shift(iq[3 * TF: 4 * TF], hz)
        nz = dabtx.add_awgn(sh, 6.0, seed=(hz & 0xFF) + 1)
        for d in (0, 100, 327, -200):
            out.append((f"clean{hz}/{d}", sh[TNULL + 504 - d: TNULL + 504 - d + TU]))
            out.append((f"noisy{hz}/{d}", nz[TNULL + 504 - d: TNULL + 504 - d + TU]))
    return out


def make_sync_options(r, iq):
    """non-default receiver options: the other two FFT placements, the three coarse frequency estimators (stage level through
    OFDMProcessor::processPRS, and in the closed loop with the coarse corrector on)"""
    g = {}
    wins = sync_windows(iq)
    g["win_sha"] = sha(np.stack([w for _, w in wins]))
    for pl in (1, 2):
        res = [r.find_index(w, {1: 0, 2: 1}[pl]) for _, w in wins]
        g[f"find_index_p{pl}"] = np.array([x[0] for x in res], np.int32)
        lim = TU if pl == 1 else 2040
        g[f"cir_sha_p{pl}"] = np.stack([sha(x[1][:lim]) for x in res])
    for m in (0, 1, 2):
        g[f"coarse_m{m}"] = np.array([r.process_prs(w, m) for _, w in wins], np.int32)
    tx = dabtx.DabTx(seed=0x51)
    sig = dabtx.freq_shift(tx.frames(12), 2000)
    g["loop_sha"] = sha(sig)
    for m in (0, 1, 2):
        e = r.e2e(sig, disable_coarse=False, select_at_fib=10 ** 9, dump_path="/tmp/golden_c.msc", freqsync_method=m)
        g[f"loop_fibs_m{m}"] = e["fibs"]
    for pl in (1, 2):
        e = r.e2e(iq[:10 * TF], disable_coarse=True, select_at_fib=10 ** 9, dump_path="/tmp/golden_p.msc", fft_placement=pl)
        g[f"loop_fibs_p{pl}"] = e["fibs"]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_sync_v1.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
