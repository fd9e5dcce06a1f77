"""Offline search for the softbit staging layout of ofdm_demod_kernel.

The demap scatters one 16-bit (re | im << 8) store per carrier into shared memory at the carrier's position in the frequency de-interleaved
order - a pseudo-random permutation, so a warp's 32 stores hit some bank 3.29 times on average (ncu: 3.29 wavefronts per instruction; this
model reproduces the number).  The reader only needs every 16-byte chunk (8 consecutive positions) to stay contiguous, so the 192 chunks
may sit anywhere: simulated annealing over chunk permutations, keeping the reader's 16-byte loads conflict-free, minimises the total
wavefront count of the 52 store instructions (4 warps x 13 slots).  The result is pasted into csrc/tables.cpp (kChunkSlot).
"""
import numpy as np, random, sys
TU, KC = 2048, 1536
invperm = -np.ones(TU, int)
pi = 0; n = 0
for i in range(TU):
    if i > 0: pi = (13 * pi + 511) % TU
    if pi == TU // 2 or pi < 256 or pi > 256 + KC: continue
    carrier = pi - TU // 2
    invperm[carrier + TU if carrier < 0 else carrier] = n; n += 1
assert n == KC
slot_c = lambda s: s if s < 7 else s + 3
# store instructions: (warp, slot) -> list of positions (or None for dummy)
instrs = []
for w in range(4):
    for s in range(13):
        pos = []
        for lane in range(32):
            t = 32 * w + lane
            iv = invperm[t + 128 * slot_c(s)]
            pos.append(int(iv) if iv >= 0 else None)
        instrs.append(pos)
def wavefronts(sigma):
    # address in 16-bit units: 8*sigma[pos//8] + pos%8 ; dummy: 1536 + t (kept unscrambled, beyond the chunks)
    tot = 0
    for k, pos in enumerate(instrs):
        w = k // 13
        banks = {}
        for lane, p in enumerate(pos):
            a = (1536 + 32 * w + lane) if p is None else 8 * sigma[p // 8] + (p % 8)
            word = a // 2
            banks.setdefault(word % 32, set()).add(word)
        tot += max(len(v) for v in banks.values())
    return tot
ident = list(range(192))
base = wavefronts(ident)
print("identity: avg wavefronts per store instr", base / len(instrs))
# reader: thread t < 96 reads chunks sigma-slot ... reader reads logical chunk c at slot sigma[c]; quarter-warp = 8 consecutive lanes reading chunks t..t+7 (h=0) / t+96.. (h=1): conflict-free iff slots distinct mod 8
def reader_conf(sigma):
    bad = 0
    for h in range(2):
        for q in range(12):
            sl = [sigma[8 * q + i + 96 * h] % 8 for i in range(8)]
            bad += 8 - len(set(sl))
    return bad
best = ident[:]; bs = base
cur = best[:]; cs = bs
T = 2.0
ITER = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
random.seed(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
for it in range(ITER):
    i, j = random.randrange(192), random.randrange(192)
    if i == j: continue
    cur[i], cur[j] = cur[j], cur[i]
    if reader_conf(cur) > 0:
        cur[i], cur[j] = cur[j], cur[i]; continue
    sc = wavefronts(cur)
    if sc <= cs or random.random() < np.exp((cs - sc) / T):
        cs = sc
        if sc < bs: bs = sc; best = cur[:]
    else:
        cur[i], cur[j] = cur[j], cur[i]
    T = max(0.03, T * (0.001 ** (1.0 / ITER)))
    if it % 20000 == 0: print(it, cs / len(instrs), bs / len(instrs), flush=True)
print("best avg", bs / len(instrs))
print(best)
