"""CPU, world_size 2 over gloo: the N>1 host logic -- contiguous image sharding and the single
collective of the path (all-reduce of the balanced-loss partial sums)."""
import os
import socket

import numpy as np
import pytest

from dsrg_b200 import shard, synth
from oracle import loss_oracle


def test_shard_ranges_partition_the_batch():
    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            got = [shard.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _terms(probs, seeds):
    """What dsrg_seedloss_forward returns for a shard (float64 restatement)."""
    p, l = probs.astype(np.float64), seeds.astype(np.float64)
    s_bg = (l[:, 0] * np.log(p[:, 0])).sum(axis=(1, 2)) / np.maximum(l[:, 0].sum(axis=(1, 2)), 1e-4)
    s_fg = (l[:, 1:] * np.log(p[:, 1:])).sum(axis=(1, 2, 3)) / np.maximum(l[:, 1:].sum(axis=(1, 2, 3)), 1e-4)
    return np.array([s_bg.sum(), s_fg.sum()])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = synth.make_batch(5, 12, 9, cues="random", start=70)   # 5 images over 2 ranks: 3 + 2
        probs = np.clip(batch["probs"], 1e-4, None)
        lo, hi = shard.shard_range(5, rank, world)
        terms, n_global = shard.allreduce_loss_terms(_terms(probs[lo:hi], batch["cues"][lo:hi]), hi - lo)
        _, n2 = shard.allreduce_loss_terms(None, hi - lo)
        q.put((rank, shard.balanced_loss_from_terms(terms, n_global), n_global, n2, (lo, hi)))
    finally:
        dist.destroy_process_group()


def test_sharded_loss_equals_single_process_loss():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = synth.make_batch(5, 12, 9, cues="random", start=70)
    want = loss_oracle.balanced_seed_loss(np.clip(batch["probs"], 1e-4, None), batch["cues"])
    for rank, loss, n_global, n2, rng in res:
        assert n_global == 5 and n2 == 5
        assert abs(loss - want) < 1e-9 * max(1.0, abs(want))
    assert sorted(r[4] for r in res) == [(0, 3), (3, 5)]
