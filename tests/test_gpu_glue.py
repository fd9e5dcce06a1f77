"""The host glue (RadioReceiver surface over the C ABI, welle.io_b200/host) driven like the reference's own harnesses
(file input, FIB dump, .msc dump, RS callbacks) against the oracle's closed-loop receiver."""
import os
import subprocess

import numpy as np
import pytest

import dabtx
from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_radio_receiver_glue(oracle, tmp_path):
    exe = os.path.join(ROOT, "welle.io_b200", "glue_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "welle.io_b200", "host")])
    tx = dabtx.DabTx(seed=0xBEEF)
    iq = tx.frames(22)
    f = tmp_path / "in.cf32"
    iq.tofile(f)
    out = subprocess.run([exe, str(f), str(tmp_path / "o"), "12"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    summary = dict(kv.split("=") for kv in out.stdout.split())
    fibs = np.fromfile(tmp_path / "o.fibs", np.uint8).reshape(-1, 33)
    msc = np.fromfile(tmp_path / "o.msc", np.uint8)
    rs = np.loadtxt(tmp_path / "o.rs", dtype=int).reshape(-1, 2)
    prot = oracle.prot_eep(96, 1, 3)
    o = oracle.rx_run(iq, prot=prot, start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True)
    n = min(len(fibs), len(o["fibs"]))
    assert n >= 12 * 18 and np.array_equal(fibs[:n], o["fibs"][:n]) and fibs[:, 0].all()
    m = min(len(msc), len(o["msc"]))
    assert m >= 288 * 40 and np.array_equal(msc[:m], o["msc"][:m])
    k = min(len(rs), len(o["rs"]))
    assert k >= 5 and np.array_equal(rs[:k], o["rs"][:k])
    assert summary["selected"] == "1" and int(summary["services"]) == 1 and int(summary["superframes"]) >= 5
    # one impulse response, one set of (L-1) K / 96 constellation points and one null symbol per decoded frame
    nfr = len(fibs) // 12
    assert int(summary["cirs"]) == int(summary["consts"]) == int(summary["nulls"]) == nfr and summary["tapsizes"] == "1"


def test_batch_decode_on_gpu(oracle, tmp_path):
    """welle.io_b200/batch_decode (native multi-file driver over the C ABI + the service database) on the real library: three
    recordings of different length, start offset and sub-channel protection decoded in lock-step; every .fic / .msc dump equals the
    oracle's for that recording (the same flow runs on the test double in tests/test_glue_mock.py)."""
    exe = os.path.join(ROOT, "welle.io_b200", "batch_decode")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "welle.io_b200", "host")])
    specs = [dict(seed=0x11, n=12, kw={}, pad=0), dict(seed=0x22, n=9, kw=dict(bitrate=64, level=2), pad=5000), dict(seed=0x33, n=14, kw={}, pad=777)]
    files, sigs = [], []
    for k, sp in enumerate(specs):
        tx = dabtx.DabTx(seed=sp["seed"], **sp["kw"])
        iq = np.concatenate([np.zeros(sp["pad"], np.complex64), tx.frames(sp["n"])])
        f = tmp_path / f"rec{k}.iq"
        iq.tofile(f); files.append(str(f)); sigs.append(iq)
    outdir = tmp_path / "out"; outdir.mkdir()
    run = subprocess.run([exe, "--out", str(outdir)] + files, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stderr
    lines = run.stdout.strip().splitlines()
    assert len(lines) == 3
    for k, sp in enumerate(specs):
        fibs = np.fromfile(outdir / f"rec{k}.iq.fic", np.uint8).reshape(-1, 33)
        msc = np.fromfile(outdir / f"rec{k}.iq.msc", np.uint8)
        br = sp["kw"].get("bitrate", 96); lv = sp["kw"].get("level", 3)
        prot = oracle.prot_eep(br, 1, lv)
        o_fic = oracle.rx_run(sigs[k], disable_coarse=True)
        n = min(len(fibs), len(o_fic["fibs"]))
        assert n >= 12 * (sp["n"] - 3) and np.array_equal(fibs[:n], o_fic["fibs"][:n]), k
        o = oracle.rx_run(sigs[k], prot=prot, start_cu=0, len_cu=dabtx.eep_cu(br, True, lv), select_after_frames=1, disable_coarse=True)
        m = min(len(msc), len(o["msc"]))
        assert m >= 3 * br * 8 and np.array_equal(msc[:m], o["msc"][:m]), k
        assert f"bitrate={br}" in lines[k] and "service=0x" in lines[k]


@pytest.mark.parametrize("fmt", ["cf32", "u8"])
def test_stock_crawfile_drives_the_glue_on_gpu(oracle, tmp_path, fmt):
    """the reference's own CRAWFile (unmodified raw_file.cpp compiled against the glue, oracle/Makefile: rawfile) feeding the B200
    backend like welle-cli does: FIB and .msc dumps equal the oracle's on the same recording (cf32 and the RTL-SDR style u8 format,
    which CRAWFile converts on the host before the glue sees it)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "rawfile_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/rawfile_test not built (needs /root/reference at build time)")
    tx = dabtx.DabTx(seed=0xC0DE)
    iq = tx.frames(16)
    if fmt == "cf32":
        f = tmp_path / "rec.cf32.iq"; iq.tofile(f); ref_sig = iq
    else:
        inter = np.stack([iq.real, iq.imag], axis=-1)
        u8 = np.clip(np.round(inter * 2.0 * 128.0 + 128.0), 0, 255).astype(np.uint8)
        f = tmp_path / "rec.u8.iq"; u8.tofile(f)
        ref_sig = ((u8.astype(np.float32) - 128.0) / 128.0).view(np.complex64).reshape(-1)
    out = subprocess.run([exe, str(f), str(tmp_path / "o"), "12"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    summary = dict(kv.split("=") for kv in out.stdout.split())
    fibs = np.fromfile(tmp_path / "o.fibs", np.uint8).reshape(-1, 33)
    msc = np.fromfile(tmp_path / "o.msc", np.uint8)
    o = oracle.rx_run(ref_sig, prot=oracle.prot_eep(96, 1, 3), start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True)
    n = min(len(fibs), len(o["fibs"]))
    assert n >= 12 * 12 and np.array_equal(fibs[:n], o["fibs"][:n]) and fibs[:n, 0].all()
    m = min(len(msc), len(o["msc"]))
    assert m >= 288 * 24 and np.array_equal(msc[:m], o["msc"][:m])
    assert summary["selected"] == "1" and int(summary["services"]) == 1 and int(summary["superframes"]) >= 3
    print(f"stock CRAWFile -> glue -> GPU ({fmt}): {summary}")


def test_scan_reports_signal_presence(tmp_path):
    """RadioReceiver::restart(doScan = true): onSignalPresence(true) once SyncOnPhase succeeds on a DAB signal, onSignalPresence(false)
    after the sixth entry into the unsynchronised state on noise (ofdm-processor.cpp:258-262,351-355)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "rawfile_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/rawfile_test not built (needs /root/reference at build time)")
    sig = dabtx.DabTx(seed=0x5CA).frames(8)
    f1 = tmp_path / "sig.cf32.iq"; sig.tofile(f1)
    rng = np.random.default_rng(5)
    noise = ((rng.standard_normal(8 * 196608) + 1j * rng.standard_normal(8 * 196608)) * 0.05).astype(np.complex64)
    f2 = tmp_path / "noise.cf32.iq"; noise.tofile(f2)
    s1 = dict(kv.split("=") for kv in subprocess.run([exe, str(f1), str(tmp_path / "a"), "12", "1", "1"], capture_output=True, text=True, timeout=600).stdout.split())
    s2 = dict(kv.split("=") for kv in subprocess.run([exe, str(f2), str(tmp_path / "b"), "12", "1", "1"], capture_output=True, text=True, timeout=600).stdout.split())
    assert s1["presence_true"] == "1" and s1["presence_false"] == "0" and int(s1["ok"]) > 0
    assert s2["presence_true"] == "0" and s2["presence_false"] == "1" and int(s2["fibs"]) == 0


def test_tii_spectra_tap_and_glue_measurements(oracle, ref, tmp_path):
    """RadioReceiverOptions::decodeTII.  (1) tap 4 of the library = fft::Forward of the frame's phase reference symbol and of the last
    T_u samples of the null symbol that follows it, bit for bit (oracle FFT = reference FFT); (2) the glue's onTIIMeasurement stream on a
    recording with two TII transmitters = the unmodified TIIDecoder fed with the same frames (positions: the reference's steady-state
    windows)."""
    from conftest import load_pkg
    pkg = load_pkg()
    TF, TU, TNULL = 196608, 2048, 2656
    tx = dabtx.DabTx(seed=0x711)
    tx.tii = [(4, 17, 23, 0.5), (11, 52, 140, 0.35)]
    sig = dabtx.add_awgn(tx.frames(16), 25.0, seed=4)
    # (1) the tap
    ctx = pkg.Context(n_streams=1, keep_taps=True)
    ctx.set_options(disable_coarse=True, decode_tii=True)
    d = ctx.dev(sig.reshape(1, -1))
    checked = 0
    for step in range(15):
        r = ctx.process(d, len(sig), np.zeros(1, np.int64), len(sig))["results"]
        if r["status"][0] != pkg.FRAME_DECODED or r["next_pos"][0] > len(sig):
            continue
        spec = ctx.read_tap(4)[0]
        null_start = int(r["next_pos"][0]) - TNULL
        prs0 = null_start - 75 * 2552 - TU
        assert np.array_equal(spec[0].view(np.uint32), oracle.fft(sig[prs0: prs0 + TU]).view(np.uint32))
        assert np.array_equal(spec[1].view(np.uint32), oracle.fft(sig[null_start + TNULL - TU: null_start + TNULL]).view(np.uint32))
        checked += 1
    ctx.close()
    assert checked >= 12
    # (2) through the glue
    exe = os.path.join(ROOT, "welle.io_b200", "glue_test")
    f = tmp_path / "tii.cf32"; sig.tofile(f)
    out = subprocess.run([exe, str(f), str(tmp_path / "t"), "12", "1", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = [tuple(float(x) for x in ln.split()) for ln in open(tmp_path / "t.tii").read().splitlines()]
    nfr = len(np.fromfile(tmp_path / "t.fibs", np.uint8)) // (33 * 12)
    nulls = np.stack([sig[(k + 1) * TF - 199: (k + 1) * TF - 199 + TNULL] for k in range(1, 1 + nfr)])
    prss = np.stack([sig[k * TF + 2961: k * TF + 2961 + TU] for k in range(1, 1 + nfr)])
    want = ref.tii_run(nulls, prss)
    print(f"TII through the glue: {nfr} frames, measurements {got}")
    assert len(want) >= 2 and got == want, (got, want)


def test_service_selection_from_the_controller_thread(oracle, tmp_path):
    """playSingleProgramme / removeServiceToDecode called from the controller thread while the worker thread is
    inside dabb_process.  glue_test's controller mode selects, removes and re-selects the service four times a few milliseconds apart
    during decoding; every dabb_* call of the glue is serialised by its context mutex.  The run must stay error-free, every FIB must
    equal the oracle's, and the .msc dump written after the last selection must be a contiguous run of the oracle's logical frames."""
    exe = os.path.join(ROOT, "welle.io_b200", "glue_test")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "welle.io_b200", "host")])
    tx = dabtx.DabTx(seed=0x2A9)
    iq = tx.frames(90)
    f = tmp_path / "in.cf32"; iq.tofile(f)
    out = subprocess.run([exe, str(f), str(tmp_path / "o"), "12", "1", "0", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "CUDA" not in out.stderr and "msg:" not in out.stderr, out.stderr
    summary = dict(kv.split("=") for kv in out.stdout.split())
    assert summary["zaps"] == "4" and summary["selected"] == "1", summary
    fibs = np.fromfile(tmp_path / "o.fibs", np.uint8).reshape(-1, 33)
    msc = np.fromfile(tmp_path / "o.msc", np.uint8)
    o = oracle.rx_run(iq, prot=oracle.prot_eep(96, 1, 3), start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True)
    n = min(len(fibs), len(o["fibs"]))
    assert n >= 12 * 85 and np.array_equal(fibs[:n], o["fibs"][:n]) and fibs[:n, 0].all()
    # the dump restarts at every selection; what it holds afterwards is the oracle's stream from some logical frame on
    assert len(msc) >= 288 * 4 * 20, (len(msc), summary)
    ref = o["msc"].tobytes()
    at = ref.find(msc[:288 * 2].tobytes())
    assert at >= 0 and at % 288 == 0
    m = min(len(msc), len(ref) - at)
    assert m >= 288 * 4 * 20 and msc[:m].tobytes() == ref[at:at + m]
    print(f"controller-thread selection: {summary['zaps']} selections, dump = oracle logical frames {at // 288} .. {(at + m) // 288}")
