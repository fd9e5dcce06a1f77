"""The C oracle against golden vectors produced by the unmodified reference (tests/golden/make_golden.py).
Runs anywhere (no /root/reference needed): this is what keeps the oracle pinned on the GPU box."""
import hashlib
import os

import numpy as np
import pytest

import dabtx

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
TU, TS, TF, TNULL = 2048, 2552, 196608, 2656


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


@pytest.fixture(scope="module")
def iq():
    sig = dabtx.DabTx(seed=0xDAB).frames(14)
    if not np.array_equal(sha(sig), G["tx_sha"]):
        pytest.skip("synthetic transmitter output differs from the one the fixtures were made with (numpy version?)")
    return sig


def test_tables(oracle):
    assert np.array_equal(oracle.perm_table(), G["perm"])
    assert np.array_equal(oracle.prs_table().view(np.uint32), G["prs"].view(np.uint32))
    assert np.array_equal(oracle.pcodes(), G["pcodes"])


def test_fft(oracle):
    assert np.array_equal(oracle.fft(G["fft_in"]).view(np.uint32), G["fft_out"].view(np.uint32))
    assert np.array_equal(oracle.ifft_scaled(G["fft_in"]).view(np.uint32), G["ifft_out"].view(np.uint32))


def test_viterbi(oracle):
    for s, o in zip(G["vit_soft"], G["vit_out"]):
        assert np.array_equal(oracle.viterbi(s, 768), o)


def test_sync_and_demod(oracle, iq):
    base = 3 * TF + TNULL
    idx, cir = oracle.find_index(iq[base - 199: base - 199 + TU])
    assert idx == int(G["find_index"][0]) == 504
    assert np.array_equal(sha(cir), G["cir_sha"]) and np.array_equal(cir[690:720], G["cir_head"])
    st = base + 305
    soft, r1 = oracle.demod_frame(iq[st: st + TU], iq[st + TU: st + TU + 75 * TS], True)
    assert np.array_equal(soft[:4, :64], G["soft_head"])
    assert np.array_equal(sha(soft), G["soft_sha"]) and np.array_equal(sha(r1), G["r1_sha"])


def test_fic(oracle):
    fb, ok = oracle.fic_decode(G["fic_soft"])
    assert np.array_equal(np.packbits(fb, axis=1), G["fib_bytes"]) and np.array_equal(ok, G["fib_ok"]) and ok.all()


def test_msc_protection(oracle):
    p = oracle.prot_eep(96, 1, 3)
    assert np.array_equal(np.packbits(oracle.msc_deconvolve(p, G["eep96a3_soft"], True)), G["eep96a3_out"])
    p = oracle.prot_eep(32, 0, 2)
    assert np.array_equal(np.packbits(oracle.msc_deconvolve(p, G["eep32b2_soft"], True)), G["eep32b2_out"])
    p = oracle.prot_uep(48, 3)
    assert np.array_equal(np.packbits(oracle.msc_deconvolve(p, G["uep48_3_soft"], True)), G["uep48_3_out"])


def test_rs_crc(oracle):
    for cw, res in zip(G["rs_cw"], G["rs_res"]):
        cnt, out, _ = oracle.rs_decode_codeword(cw)
        assert (cnt & 0xFF) == res[0] and np.array_equal(out, res[1:])
    assert np.array_equal(oracle.rs_encode(np.arange(110, dtype=np.uint8)), G["rs_parity_kat"])
    assert oracle.crc_fire(np.arange(9, dtype=np.uint8)) == G["crc"][0] and oracle.crc_ccitt(np.frombuffer(b"123456789", np.uint8)) == G["crc"][1]


def test_closed_loop(oracle, iq):
    p = oracle.prot_eep(96, 1, 3)
    m = oracle.rx_run(iq, prot=p, start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True)
    n = len(G["e2e_fibs"])
    assert np.array_equal(m["fibs"][:n], G["e2e_fibs"])
    k = min(len(m["msc"]), len(G["e2e_msc"]))
    assert k > 0 and np.array_equal(m["msc"][:k], G["e2e_msc"][:k])
    j = min(len(m["rs"]), len(G["e2e_rs"]))
    assert np.array_equal(m["rs"][:j], G["e2e_rs"][:j])


def test_round_trips(oracle):
    """size-independent properties: encode -> corrupt -> decode"""
    rng = np.random.default_rng(3)
    for nb in (768, 2304):
        bits = rng.integers(0, 2, nb).astype(np.uint8)
        soft = ((oracle.conv_encode(bits).astype(np.int16) * 2 - 1) * 90).astype(np.int8)
        soft[rng.random(soft.size) < 0.4] = 0          # heavy puncturing / erasures
        assert np.array_equal(oracle.viterbi(soft, nb), bits)
    data = rng.integers(0, 256, 110).astype(np.uint8)
    cw = np.concatenate([data, oracle.rs_encode(data)])
    e = cw.copy(); e[[3, 40, 77, 101, 119]] ^= 0x5A
    cnt, out, pos = oracle.rs_decode_codeword(e)
    assert cnt == 5 and np.array_equal(out, cw) and sorted(p - 135 for p in pos[:5]) == [3, 40, 77, 101, 119]


# ---- non-default receiver options (golden_sync_v1.npz): FFT placements and coarse frequency estimators
GS = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_sync_v1.npz"))


@pytest.fixture(scope="module")
def sync_wins(iq):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    src = open(spec.origin).read()
    ns = {"dabtx": dabtx, "np": np, "TU": TU, "TS": TS, "TF": TF, "TNULL": TNULL}
    start = src.index("def sync_windows"); end = src.index("def make_sync_options")
    exec(src[start:end], ns)                      # the window list is defined once, next to the generator
    wins = ns["sync_windows"](iq)
    if not np.array_equal(sha(np.stack([w for _, w in wins])), GS["win_sha"]):
        pytest.skip("synthetic windows differ from the ones the fixtures were made with (numpy version?)")
    return wins


@pytest.mark.parametrize("placement", [1, 2])
def test_find_index_placements_golden(oracle, sync_wins, placement):
    lim = TU if placement == 1 else 2040
    for k, (name, w) in enumerate(sync_wins):
        idx, cir = oracle.find_index(w, placement)
        assert idx == int(GS[f"find_index_p{placement}"][k]), name
        assert np.array_equal(sha(cir[:lim]), GS[f"cir_sha_p{placement}"][k]), name


@pytest.mark.parametrize("method", [0, 1, 2])
def test_coarse_estimators_golden(oracle, sync_wins, method):
    got = np.array([oracle.coarse(w, method) for _, w in sync_wins], np.int32)
    assert np.array_equal(got, GS[f"coarse_m{method}"]), (got, GS[f"coarse_m{method}"])


@pytest.mark.parametrize("method", [0, 1, 2])
def test_closed_loop_coarse_golden(oracle, method):
    sig = dabtx.freq_shift(dabtx.DabTx(seed=0x51).frames(12), 2000)
    if not np.array_equal(sha(sig), GS["loop_sha"]):
        pytest.skip("synthetic stream differs from the fixture's")
    m = oracle.rx_run(sig, disable_coarse=False, freqsync_method=method)
    ref = GS[f"loop_fibs_m{method}"]
    n = min(len(ref), len(m["fibs"]))
    assert n >= 12 * 8 and np.array_equal(m["fibs"][:n], ref[:n])


@pytest.mark.parametrize("placement", [1, 2])
def test_closed_loop_placement_golden(oracle, iq, placement):
    m = oracle.rx_run(iq[:10 * TF], disable_coarse=True, fft_placement=placement)
    ref = GS[f"loop_fibs_p{placement}"]
    n = min(len(ref), len(m["fibs"]))
    assert n >= 12 * 7 and np.array_equal(m["fibs"][:n], ref[:n])
