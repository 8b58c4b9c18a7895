"""Sweep the chunk schedule of the host-buffer pass (DSRG_B200_HOST_SCHEDULE) on the headline workload
(batch 64 @ 321x321x21): ms per dsrg_dsrg_forward_host call, median of `--reps`.
usage: python tools/sweep_schedule.py "8,24,32" "8,12,20,24" ..."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from dsrg_b200 import api, synth
    scheds = [a for a in sys.argv[1:] if not a.startswith("--")] or ["8,24,32"]
    reps = 7
    B, H, W, M = 64, 321, 321, 21
    ps = [synth.make_problem(i, H, W) for i in range(8)]
    batch = {k: np.stack([ps[i % 8][k] for i in range(B)]) for k in ("labels", "probs", "cues", "image")}
    eng = api.Engine(B, H, W, M)
    params = api.crf_params(1.0, 13, 10)
    hb = {}
    for k, dt in (("labels", np.float32), ("probs", np.float32), ("cues", np.float32), ("image", np.uint8)):
        hb[k] = api.pinned_empty(batch[k].shape, dt)
        hb[k][...] = batch[k]
    seeds = api.pinned_empty(batch["probs"].shape, np.float32)

    def step():
        eng.dsrg_forward_host(hb["labels"], hb["probs"], hb["cues"], hb["image"], params, 0.99, 0.85, seeds_out=seeds)
    os.environ.pop("DSRG_B200_HOST_SCHEDULE", None)
    for _ in range(3):
        step()
    for sc in scheds + [scheds[0]]:
        os.environ["DSRG_B200_HOST_SCHEDULE"] = sc
        step()
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step()
            ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort()
        print("%-22s median %.2f ms  min %.2f  max %.2f  -> %.0f images/s" % (sc, ts[len(ts) // 2], ts[0], ts[-1],
                                                                         1e3 * B / ts[len(ts) // 2]), flush=True)


if __name__ == "__main__":
    main()
