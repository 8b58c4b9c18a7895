"""Latency of the inference post-processing: predict_mask() tails of
training/tools/test-ms.py on VOC-sized images, one image per call, sizes changing from call to call --
through postprocess.predict_mask_ms (host numpy in, host label map out).  Not the headline metric
(that is bench.py); prints one JSON line.

usage: python tools/bench_infer.py [--images 40] [--cpu 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SIZES = [(375, 500), (500, 375), (333, 500), (500, 334), (366, 500), (281, 500), (500, 500), (375, 500)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=40)
    ap.add_argument("--cpu", type=int, default=2, help="images timed through the CPU oracle (0 = skip)")
    args = ap.parse_args()
    import torch
    from dsrg_b200 import pool, postprocess, synth
    cases = [synth.make_score_blobs(100 + i, H, W, (31, 41, 51)) for i, (H, W) in enumerate(SIZES)]
    for c in cases:                                   # warm-up: engine growth, lazy allocations
        postprocess.predict_mask_ms(c["image"], c["blobs"])
    eng = list(pool._ENGINES.values())[0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.images):
        c = cases[i % len(cases)]
        postprocess.predict_mask_ms(c["image"], c["blobs"])
    dt = time.perf_counter() - t0
    eng.take_launch_count()
    eng.profile(True)
    for c in cases:
        postprocess.predict_mask_ms(c["image"], c["blobs"])
    prof = eng.profile_read()
    eng.profile(False)
    launches = eng.take_launch_count()
    gpu_ms = sum(v[0] for v in prof.values()) / len(cases)
    out = {"metric": "images/s, predict_mask post-processing (3 score scales -> zoom/sum -> softmax -> CRF 10 it -> argmax), "
                     "one VOC-sized image per call, host in / host out",
           "value": args.images / dt, "ms_per_image": 1e3 * dt / args.images, "images": args.images,
           "sizes": SIZES, "engine_capacity": list(eng.capacity), "engine_device_MB": eng.device_bytes / 1e6,
           "kernel_ms_per_image": gpu_ms, "launches_per_image": launches / len(cases),
           "kernel_classes_ms_per_image": {k: v[0] / len(cases) for k, v in sorted(prof.items())}}
    if args.cpu:
        from oracle import post_oracle
        t0 = time.perf_counter()
        for c in cases[:args.cpu]:
            post_oracle.predict_mask_ms(c["image"], c["blobs"])
        cdt = (time.perf_counter() - t0) / args.cpu
        out["cpu_oracle"] = {"ms_per_image": 1e3 * cdt, "images": args.cpu, "cores": 1,
                             "kind": "port (scipy zoom + numpy + oracle/crf_oracle.c)"}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
