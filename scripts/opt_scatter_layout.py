"""Offline search for the softbit staging layout of ofdm_demod_kernel.

The demap scatters one 16-bit (re | im << 8) store per carrier into shared memory at the carrier's position in the frequency
de-interleaved order - a pseudo-random permutation, so a warp's 32 stores hit some bank 3.29 times on average (ncu: 3.29 wavefronts per
store instruction; the model below reproduces that number).  The reader only needs every 16-byte chunk (8 consecutive positions) to stay
together, so (a) the 192 chunks may sit in any slot and (b) the two 8-byte halves of a chunk may be stored swapped (the reader swaps its
two output words back).  Simulated annealing over (slot permutation, swap bits), keeping the reader's 16-byte loads conflict-free,
minimises the wavefronts of the 52 store instructions (4 warps x 13 slots).  The result is pasted into csrc/tables.cpp
(kChunkSlot / kChunkSwap); build_host_tables() checks that it is a permutation and tests compare every softbit with the oracle.

    python scripts/opt_scatter_layout.py [iterations] [seed]
"""
import random
import sys

import numpy as np

TU, KC = 2048, 1536
invperm = -np.ones(TU, int)
pi = 0; n = 0
for i in range(TU):                    # frequency interleaver of Mode I (EN 300 401 14.6), as in csrc/tables.cpp
    if i > 0:
        pi = (13 * pi + 511) % TU
    if pi == TU // 2 or pi < 256 or pi > 256 + KC:
        continue
    carrier = pi - TU // 2
    invperm[carrier + TU if carrier < 0 else carrier] = n; n += 1
assert n == KC
slot_c = lambda s: s if s < 7 else s + 3          # bins owned by thread t: t + 128 c (csrc/ofdm_core.cuh)
instrs = []                                        # one entry per store instruction: the 32 lanes' logical positions (None = dummy)
for w in range(4):
    for s in range(13):
        instrs.append([int(invperm[32 * w + lane + 128 * slot_c(s)]) if invperm[32 * w + lane + 128 * slot_c(s)] >= 0 else None for lane in range(32)])


def wavefronts(sigma, swap):
    tot = 0
    for k, pos in enumerate(instrs):
        w = k // 13
        banks = {}
        for lane, p in enumerate(pos):
            word = (1536 + 32 * w + lane) // 2 if p is None else 4 * sigma[p // 8] + ((((p % 8) // 2) + 2 * swap[p // 8]) % 4)
            banks.setdefault(word % 32, set()).add(word)
        tot += max(len(v) for v in banks.values())
    return tot


def reader_ok(sigma):                              # thread t < 96 loads chunks t and t + 96: 8 lanes of a quarter warp -> 8 distinct bank groups
    return all(len(set(sigma[8 * q + i + 96 * h] % 8 for i in range(8))) == 8 for h in range(2) for q in range(12))


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 600000
    random.seed(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    sigma, swap = list(range(192)), [0] * 192
    print("identity layout:", wavefronts(sigma, swap) / len(instrs), "wavefronts per store instruction")
    cs = wavefronts(sigma, swap); bs, best = cs, (sigma[:], swap[:]); T = 2.0
    for it in range(iters):
        if random.random() < 0.5:
            i, j = random.randrange(192), random.randrange(192)
            if i == j:
                continue
            sigma[i], sigma[j] = sigma[j], sigma[i]
            ok = reader_ok(sigma)
            sc = wavefronts(sigma, swap) if ok else None
            if ok and (sc <= cs or random.random() < np.exp((cs - sc) / T)):
                cs = sc
            else:
                sigma[i], sigma[j] = sigma[j], sigma[i]
        else:
            i = random.randrange(192); swap[i] ^= 1
            sc = wavefronts(sigma, swap)
            if sc <= cs or random.random() < np.exp((cs - sc) / T):
                cs = sc
            else:
                swap[i] ^= 1
        if cs < bs:
            bs, best = cs, (sigma[:], swap[:])
        T = max(0.03, T * (0.001 ** (1.0 / iters)))
        if it % 100000 == 0:
            print(it, cs / len(instrs), bs / len(instrs), flush=True)
    print("best:", bs / len(instrs), "wavefronts per store instruction")
    print("kChunkSlot =", best[0]); print("kChunkSwap =", best[1])


if __name__ == "__main__":
    main()
