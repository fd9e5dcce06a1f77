"""N > 1 host logic on CPU: two gloo ranks partition 6 synthetic ensemble streams, get the work descriptor by broadcast,
decode their shards (CPU oracle standing in for the GPU) and all-reduce the counters; the totals must equal a
single-process run."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_pkg


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _decode_shard(ids, n_frames):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import dabtx
    from oracle.bind import Oracle
    o = Oracle()
    frames = fib_ok = fib_tot = 0
    for e in ids:
        iq = dabtx.DabTx(seed=0x500 + e).frames(n_frames)
        r = o.rx_run(iq, disable_coarse=True)
        frames += r["frames"]; fib_ok += int(r["fibs"][:, 0].sum()); fib_tot += len(r["fibs"])
    return [frames, fib_ok, fib_tot]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_pkg()
    dev = torch.device("cpu")
    desc = pkg.sharding.broadcast_descriptor([6, 4] if rank == 0 else [], dev)
    ids = pkg.sharding.partition(desc[0], world, rank)
    local = _decode_shard(ids, desc[1])
    total = pkg.sharding.reduce_counters(local, dev)
    tmax = pkg.sharding.max_over_ranks(float(rank + 1), dev)
    if rank == 0:
        q.put((desc, ids, total, tmax))
    dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    desc, ids, total, tmax = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert desc == [6, 4] and ids == [0, 2, 4] and tmax == 2.0
    single = _decode_shard(list(range(6)), 4)
    assert total == single and total[1] == total[2] > 0
    pkg = load_pkg()
    assert sorted(pkg.sharding.partition(10, 4, 1)) == [1, 5, 9]
    assert sorted(sum((pkg.sharding.partition(13, 4, r) for r in range(4)), [])) == list(range(13))
