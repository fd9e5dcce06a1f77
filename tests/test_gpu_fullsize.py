"""GPU tests at BASELINE.json's sizes through size-independent properties (the oracle is too slow to check every unit):
encode -> noise -> decode round trips for the batched Viterbi at 32 768 codewords, an SNR sweep whose decoded bits
must equal the CPU oracle bit for bit on a sample of codewords (config[3]: "Viterbi BER curve matches CPU ref"), and
batch = 1024 frames through the full chain with 100 % FIC CRC pass rate (config[2])."""
import numpy as np
import pytest

import dabtx
from conftest import load_pkg

pytestmark = pytest.mark.gpu
TF = 196608


def test_viterbi_round_trip_32768_codewords(oracle):
    pkg = load_pkg()
    ctx = pkg.Context(n_streams=1)
    rng = np.random.default_rng(42)
    n, nb = 32768, 768
    bits = rng.integers(0, 2, (64, nb)).astype(np.uint8)
    enc = np.stack([oracle.conv_encode(b) for b in bits]).astype(np.int16) * 2 - 1          # 64 distinct codewords, tiled
    idx = rng.integers(0, 64, n)
    soft = enc[idx] * 60 + rng.integers(-40, 41, (n, enc.shape[1]), dtype=np.int16)          # high SNR: decoding must be error free
    soft[rng.random(soft.shape) < 0.25] = 0                                                  # punctured positions
    out = ctx.viterbi(np.clip(soft, -127, 127).astype(np.int8), nb)
    assert np.array_equal(out, bits[idx])
    ctx.close()


def test_viterbi_snr_sweep_matches_cpu(oracle):
    """S3: per-SNR decoded bits identical to the CPU on identical softbits; BER falls with SNR"""
    pkg = load_pkg()
    ctx = pkg.Context(n_streams=1)
    rng = np.random.default_rng(7)
    nb, per = 2304, 96
    ber = []
    for snr_db in range(0, 9):        # Es/N0 per coded bit, rate 1/4 mother code
        bits = rng.integers(0, 2, (per, nb)).astype(np.uint8)
        enc = np.stack([oracle.conv_encode(b) for b in bits]).astype(np.float32) * 2 - 1
        sigma = 10 ** (-snr_db / 20) / np.sqrt(2)
        soft = np.clip(np.round((enc + rng.standard_normal(enc.shape) * sigma * 2.5) * 32), -127, 127).astype(np.int8)
        out = ctx.viterbi(soft, nb)
        for i in range(0, per, 8):    # every 8th codeword against the oracle (bit exact)
            assert np.array_equal(out[i], oracle.viterbi(soft[i], nb)), (snr_db, i)
        ber.append(float((out != bits).mean()))
    assert ber[-1] == 0.0 and ber[0] >= ber[4] >= ber[-1]
    ctx.close()


def test_full_chain_batch_1024():
    """config[2]: 1024 independent streams, full chain, FIC CRC pass rate 100 %, MSC logical frames and RS clean"""
    pkg = load_pkg()
    S = 1024
    rings = [dabtx.periodic_ring(0x900 + i, 5)[1] for i in range(4)]
    buf_len = 6 * TF + 4096
    host = np.zeros((S, buf_len), np.complex64)
    rng = np.random.default_rng(1)
    for s in range(S):
        r = rings[s % 4]
        host[s, :5 * TF] = r
        host[s, 5 * TF:] = r[:buf_len - 5 * TF]
    host += ((rng.standard_normal(host.shape, dtype=np.float32) + 1j * rng.standard_normal(host.shape, dtype=np.float32)) * 0.005).astype(np.complex64)
    ctx = pkg.Context(n_streams=S)
    ctx.select_subchannel(0, 72, 96, eep_profile_a=True, eep_level=3)
    d = ctx.dev(host)
    tot_fib = ok_fib = logical = rs_unc = rs_att = 0
    for call in range(12):
        n = call + 1
        bs = np.zeros(S, np.int64) if call == 0 else np.full(S, 5 * TF * (n // 5), np.int64)
        out = ctx.process(d, buf_len, bs, buf_len, msc_stride=288)
        r = out["results"]
        assert (r["status"] == pkg.FRAME_DECODED).all(), (call, np.unique(r["status"], return_counts=True))
        tot_fib += 12 * S; ok_fib += int(sum(bin(int(m)).count("1") for m in r["fib_crc_mask"]))
        if call >= 8:
            logical += int(r["n_logical"][:, 0].sum()); rs_att += int(r["n_rs_events"][:, 0].sum()); rs_unc += int((r["rs_uncorr_mask"][:, 0] != 0).sum())
    assert ok_fib == tot_fib
    assert logical == 4 * S * 4 and rs_att > 0 and rs_unc == 0
    ctx.close()
