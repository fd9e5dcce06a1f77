"""The host glue (welle.io_b200/host: RadioReceiver surface, worker loop, service database, callbacks, .msc dump) end to end on the CPU,
with a TEST DOUBLE of the C ABI (tests/mock_backend/mock_dab_b200.c: hands out the oracle's frame records) in place of libdab_b200.so.
The double is compiled into a temporary directory and only this test points the dynamic loader at it; the same flow runs against the real
GPU library in tests/test_gpu_glue.py."""
import os
import subprocess

import numpy as np

import dabtx
from conftest import ROOT


def test_radio_receiver_glue_with_mock_backend(oracle, tmp_path):
    exe = os.path.join(ROOT, "welle.io_b200", "glue_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "welle.io_b200", "host")])
    mock = tmp_path / "libdab_b200.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", str(mock), os.path.join(ROOT, "tests", "mock_backend", "mock_dab_b200.c"),
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    tx = dabtx.DabTx(seed=0xBEEF)
    iq = tx.frames(14)
    f = tmp_path / "in.cf32"
    iq.tofile(f)
    env = dict(os.environ, LD_LIBRARY_PATH=str(tmp_path), DABB_MOCK_IQ=str(f))
    out = subprocess.run([exe, str(f), str(tmp_path / "o"), "12"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    summary = dict(kv.split("=") for kv in out.stdout.split())
    fibs = np.fromfile(tmp_path / "o.fibs", np.uint8).reshape(-1, 33)
    msc = np.fromfile(tmp_path / "o.msc", np.uint8)
    prot = oracle.prot_eep(96, 1, 3)
    o = oracle.rx_run(iq, prot=prot, start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True)
    # every FIB of every frame reaches onFIBDecodeSuccess, bit for bit, and the service database built from them lists the one service
    n = min(len(fibs), len(o["fibs"]))
    assert n >= 12 * 10 and np.array_equal(fibs[:n], o["fibs"][:n]) and fibs[:, 0].all()
    assert summary["selected"] == "1" and int(summary["services"]) == 1 and int(summary["syncs"]) == 1
    # playSingleProgramme (from inside the 12th FIB callback) selected the sub-channel FIG 0/1 + 0/2 describe: the dump holds the logical frames
    m = min(len(msc), len(o["msc"]))
    assert m >= 288 * 20 and np.array_equal(msc[:m], o["msc"][:m])
    assert int(summary["logical_frames"]) == len(msc) // 288


def test_batch_decode_with_mock_backend(oracle, tmp_path):
    """welle.io_b200/batch_decode (native multi-file driver over the C ABI + the service database): three recordings of different length,
    start offset and sub-channel protection decoded in lock-step; every .fic / .msc dump equals the oracle's for that recording"""
    exe = os.path.join(ROOT, "welle.io_b200", "batch_decode")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "welle.io_b200", "host")])
    mock = tmp_path / "libdab_b200.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", str(mock), os.path.join(ROOT, "tests", "mock_backend", "mock_dab_b200.c"),
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    specs = [dict(seed=0x11, n=12, kw={}, pad=0), dict(seed=0x22, n=9, kw=dict(bitrate=64, level=2), pad=5000), dict(seed=0x33, n=14, kw={}, pad=777)]
    files, sigs, txs = [], [], []
    for k, sp in enumerate(specs):
        tx = dabtx.DabTx(seed=sp["seed"], **sp["kw"])
        iq = np.concatenate([np.zeros(sp["pad"], np.complex64), tx.frames(sp["n"])])
        f = tmp_path / f"rec{k}.iq"
        iq.tofile(f); files.append(str(f)); sigs.append(iq); txs.append(tx)
    outdir = tmp_path / "out"; outdir.mkdir()
    env = dict(os.environ, LD_LIBRARY_PATH=str(tmp_path), DABB_MOCK_IQ=":".join(files))
    run = subprocess.run([exe, "--out", str(outdir)] + files, capture_output=True, text=True, timeout=600, env=env)
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
        # the driver tunes as soon as FIG 0/1 + 0/2 have been seen twice: after the first decoded frame
        o = oracle.rx_run(sigs[k], prot=prot, start_cu=0, len_cu=dabtx.eep_cu(br, True, lv), select_after_frames=1, disable_coarse=True)
        m = min(len(msc), len(o["msc"]))
        assert m >= 3 * br * 8 and np.array_equal(msc[:m], o["msc"][:m]), k
        assert f"bitrate={br}" in lines[k] and "service=0x" in lines[k]


def _rawfile_exe():
    import pytest
    exe = os.path.join(ROOT, "oracle", "_ref", "rawfile_test")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "welle.io_b200", "host")])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "rawfile"])
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/rawfile_test not built (needs /root/reference)")
    return exe


def test_stock_crawfile_drives_the_glue_with_mock_backend(oracle, tmp_path):
    """The lower boundary with the reference's OWN input class: the unmodified src/input/raw_file.cpp (+ raw_file.h, virtual_input.h,
    various/ringbuffer.h), compiled against the glue's headers and linked with libwelle_b200_host.so (oracle/Makefile: rawfile), feeds
    RadioReceiver through CRAWFile(throttle=false, rewind=false) exactly like welle-cli.cpp:514-516,612-664: FIB dump and .msc dump equal
    the oracle's (which is pinned to the reference's own RadioReceiver on the same file)."""
    exe = _rawfile_exe()
    mock = tmp_path / "libdab_b200.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", str(mock), os.path.join(ROOT, "tests", "mock_backend", "mock_dab_b200.c"),
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    tx = dabtx.DabTx(seed=0xC0DE)
    iq = tx.frames(14)
    f = tmp_path / "rec.cf32.iq"
    iq.tofile(f)
    env = dict(os.environ, LD_LIBRARY_PATH=str(tmp_path), DABB_MOCK_IQ=str(f))
    out = subprocess.run([exe, str(f), str(tmp_path / "o"), "12"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    summary = dict(kv.split("=") for kv in out.stdout.split())
    fibs = np.fromfile(tmp_path / "o.fibs", np.uint8).reshape(-1, 33)
    msc = np.fromfile(tmp_path / "o.msc", np.uint8)
    o = oracle.rx_run(iq, prot=oracle.prot_eep(96, 1, 3), start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True)
    n = min(len(fibs), len(o["fibs"]))
    assert n >= 12 * 10 and np.array_equal(fibs[:n], o["fibs"][:n]) and fibs[:, 0].all()
    assert summary["selected"] == "1" and int(summary["services"]) == 1 and int(summary["syncs"]) == 1
    m = min(len(msc), len(o["msc"]))
    assert m >= 288 * 20 and np.array_equal(msc[:m], o["msc"][:m])
    # getReceiverStats().timeLastFCT0Frame is served (set when the ensemble was cleared at restart / by FIG 0/0 with CIF count 0)
    assert 0.0 <= float(summary["fct0_age_s"]) < 120.0


def test_service_selection_from_the_controller_thread_with_mock_backend(oracle, tmp_path):
    """the glue's controller-thread path (playSingleProgramme / removeServiceToDecode from another thread while the worker decodes; every
    dabb_* call serialised by the glue's context mutex) on the test double; the same flow runs on the GPU in tests/test_gpu_glue.py"""
    exe = os.path.join(ROOT, "welle.io_b200", "glue_test")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "welle.io_b200", "host")])
    mock = tmp_path / "libdab_b200.so"
    subprocess.check_call(["gcc", "-O1", "-shared", "-fPIC", "-o", str(mock), os.path.join(ROOT, "tests", "mock_backend", "mock_dab_b200.c"),
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    iq = dabtx.DabTx(seed=0x2A9).frames(44)
    f = tmp_path / "in.cf32"; iq.tofile(f)
    env = dict(os.environ, LD_LIBRARY_PATH=str(tmp_path), DABB_MOCK_IQ=str(f))
    out = subprocess.run([exe, str(f), str(tmp_path / "o"), "12", "1", "0", "1"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr
    summary = dict(kv.split("=") for kv in out.stdout.split())
    assert int(summary["zaps"]) >= 2 and summary["selected"] == "1", summary
    fibs = np.fromfile(tmp_path / "o.fibs", np.uint8).reshape(-1, 33)
    msc = np.fromfile(tmp_path / "o.msc", np.uint8)
    o = oracle.rx_run(iq, prot=oracle.prot_eep(96, 1, 3), start_cu=0, len_cu=72, select_after_frames=1, disable_coarse=True)
    n = min(len(fibs), len(o["fibs"]))
    assert n >= 12 * 40 and np.array_equal(fibs[:n], o["fibs"][:n])
    ref = o["msc"].tobytes()
    assert len(msc) >= 288 * 8
    at = ref.find(msc[:288 * 2].tobytes())
    m = min(len(msc), len(ref) - at)
    assert at >= 0 and at % 288 == 0 and m >= 288 * 8 and msc[:m].tobytes() == ref[at:at + m]
