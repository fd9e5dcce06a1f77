"""CPU emulation of the CUDA per-thread code (the same .cuh headers compiled as plain C++): the FFT pass structure with its
shared-memory swizzle, and the packed-16-bit Viterbi with its rotating register layout, against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")


@pytest.fixture(scope="module")
def libs():
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", os.path.join(HERE, "libemul.so"), os.path.join(HERE, "emul_fft.cpp")])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", os.path.join(HERE, "libemulv.so"), os.path.join(HERE, "emul_viterbi.cpp")])
    return C.CDLL(os.path.join(HERE, "libemul.so")), C.CDLL(os.path.join(HERE, "libemulv.so"))


def test_fft_passes_bit_exact_and_conflict_free(libs, oracle):
    rng = np.random.default_rng(3)
    for inv in (0, 1):
        x = (rng.standard_normal(2048) + 1j * rng.standard_normal(2048)).astype(np.complex64); y = np.zeros_like(x)
        worst = libs[0].emul_fft2048(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), inv)
        assert np.array_equal(y.view(np.uint32), oracle.fft(x, bool(inv)).view(np.uint32))
        assert worst == 1, "shared-memory exchange pattern has bank conflicts"


def test_packed_viterbi(libs, oracle):
    rng = np.random.default_rng(11)
    for trial in range(16):
        nb = [768, 2304, 192, 24 * 384][trial % 4]
        soft = rng.integers(-128, 128, (nb + 6) * 4).astype(np.int8)
        if trial % 2:
            bits = rng.integers(0, 2, nb).astype(np.uint8)
            soft = np.clip((oracle.conv_encode(bits).astype(np.float32) * 2 - 1) * 30 + rng.standard_normal((nb + 6) * 4) * 60, -127, 127).astype(np.int8)
            soft[rng.random(soft.size) < 0.3] = 0
        out = np.zeros(nb, np.uint8)
        libs[1].emul_viterbi(nb, soft.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert np.array_equal(out, oracle.viterbi(soft, nb))


def test_oscillator_on_the_fly_equals_table_for_every_index():
    """the formula the device evaluates instead of reading the reference's 2 048 000-entry oscillator table (osc_factors.h), run with
    IEEE double arithmetic on the CPU: identical to the table for every index; three factors (the quarter turns) carry the table's value"""
    so = os.path.join(HERE, "libemulo.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "emul_osc.cpp")])
    lib = C.CDLL(so)
    patched = C.c_int()
    assert lib.emul_osc_mismatches(C.byref(patched)) == 0
    assert patched.value == 3
