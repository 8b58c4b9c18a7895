"""Worker of tests/test_gpu_multi.py: one process per GPU under torchrun (NCCL).  Each rank runs the DSRG pass on
its contiguous shard of a seeded global batch; rank 0 also runs the whole batch alone and compares:
  * SRG-only seeds of the sharded run == single-GPU seeds, bitwise;
  * full pass (CRF -> SRG) seeds: equal up to threshold ties under the float-atomics run-to-run noise, and the CRF
    marginals within 2e-5;
  * balanced seeding loss after the NCCL all-reduce == the single-process loss (1e-6: the per-rank sums are
    accumulated in float64 on the device and handed over as float32);
  * every rank stayed on its own CUDA device (the drop-ins pick the caller's current device).
Prints one line "MGPU-OK ..." on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from dsrg_b200 import api, shard, synth, _lib

    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    assert L.dsrg_current_device() == local
    NB, H, W, M = 7, 41, 41, 21          # 7 images over `world` ranks: ragged shards
    batch = synth.make_batch(NB, H, W, cues="cam", image="smooth", start=300)
    params = api.crf_params(12.0)
    lo, hi = shard.shard_range(NB, rank, world)
    n = hi - lo
    eng = api.Engine(max(n, 1), H, W, M)   # device=None: this rank's current device
    assert eng.device == local
    sl = slice(lo, hi)
    probs = batch["probs"][sl].copy()
    seeds = eng.dsrg_forward_host(batch["labels"][sl], probs, batch["cues"][sl], batch["image"][sl], params, 0.99, 0.85)
    seeds_srg = eng.srg_host(batch["labels"][sl], batch["probs"][sl], batch["cues"][sl], 0.99, 0.85)
    terms = eng.seedloss_forward_host(np.clip(batch["probs"][sl], 1e-4, None), seeds)
    g_terms, n_global = shard.allreduce_loss_terms(terms, n)     # the path's only collective (NCCL)
    loss = shard.balanced_loss_from_terms(g_terms, n_global)
    assert torch.cuda.current_device() == local and L.dsrg_current_device() == local
    # gather the shards on rank 0 (test harness only: the product keeps outputs sharded)
    out = [None] * world
    dist.all_gather_object(out, (lo, hi, seeds, seeds_srg, probs))
    if rank == 0:
        full = api.Engine(NB, H, W, M)
        p1 = batch["probs"].copy()
        s1 = full.dsrg_forward_host(batch["labels"], p1, batch["cues"], batch["image"], params, 0.99, 0.85)
        s1_srg = full.srg_host(batch["labels"], batch["probs"], batch["cues"], 0.99, 0.85)
        t1 = full.seedloss_forward_host(np.clip(batch["probs"], 1e-4, None), s1)
        loss1 = shard.balanced_loss_from_terms(t1, NB)
        got = np.concatenate([o[2] for o in sorted(out, key=lambda o: o[0])])
        got_srg = np.concatenate([o[3] for o in sorted(out, key=lambda o: o[0])])
        got_p = np.concatenate([o[4] for o in sorted(out, key=lambda o: o[0])])
        assert np.array_equal(got_srg, s1_srg), "sharded SRG differs from the single-GPU SRG"
        assert np.array_equal(got_p, p1), "in-place clamp differs"
        mism = int((got != s1).sum())
        assert mism <= 4, "sharded full-pass seeds differ in %d values" % mism
        # the loss terms depend on the seeds; with identical seeds the two sums agree to float64 rounding
        tol = 1e-6 if mism == 0 else 1e-4   # the C ABI hands the two per-rank sums over as float32
        assert abs(loss - loss1) <= tol * max(1.0, abs(loss1)), (loss, loss1)
        assert n_global == NB
        print("MGPU-OK world=%d shards=%s seed_mismatch=%d loss=%.9f loss_1gpu=%.9f" %
              (world, [(o[0], o[1]) for o in sorted(out, key=lambda o: o[0])], mism, loss, loss1), flush=True)
        full.close()
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
