"""diagnostic (GPU box): per-call receiver state across a sync loss, next to the oracle's frames"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dabtx
from conftest import load_pkg
from oracle.bind import Oracle
TF = 196608
pkg = load_pkg(); orc = Oracle()
s = dabtx.DabTx(seed=0x6B).frames(22)
a, b = s[:9 * TF + 40000], s[11 * TF - 777:]
iq = np.concatenate([a, np.full(int(1.3 * TF), 1e-5 + 0j, np.complex64), b]).astype(np.complex64)
ctx = pkg.Context(n_streams=1)
d = ctx.dev(iq.reshape(1, -1))
pos = 0
for call in range(40):
    r = ctx.process(d, len(iq), np.zeros(1, np.int64), len(iq))["results"]
    print(f"gpu call {call:2d}: status {int(r['status'][0])} idx {int(r['start_index'][0]):5d} next_pos {int(r['next_pos'][0]):8d} (+{int(r['next_pos'][0]) - pos:7d}) fine {int(r['fine_corr'][0]):4d} crc {int(r['fib_crc_mask'][0]):03x} slevel {float(r['slevel'][0]):.5f} acq_failed {int(r['acq_failed'][0])}")
    pos = int(r["next_pos"][0])
    if r["status"][0] == pkg.FRAME_NEED_SAMPLES:
        break
ctx.close()
o = orc.rx_run(iq, disable_coarse=True)
for f in range(o["frames"]):
    i = o["info"][f]
    crc = int(sum(int(x) << k for k, x in enumerate(o["fibs"][12 * f:12 * f + 12, 0])))
    print(f"orc frame {f:2d}: pos {i['pos']:8d} idx {i['start_index']:5d} prs_abs {i['pos'] + i['start_index']:8d} fine {i['fine']:4d} crc {crc:03x}")
