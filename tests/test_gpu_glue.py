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
