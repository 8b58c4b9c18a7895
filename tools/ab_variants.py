"""A/B harness for kernel tuning: build variant libraries here (CPU, nvcc cross-compiles), then time them
all in ONE gpurun call.

  here :  python tools/ab_variants.py build  base:  pad0:DSRG_ROW_PAD=0  ctas3:DSRG_TILE_CTAS=3,DSRG_MAXLOC_BI=256
  GPU  :  gpurun --timeout 900 -- 'python tools/ab_variants.py run base pad0 ctas3 --rounds 2'

`name:` (no defines) is the default library.  `run` executes `bench.py --no-cpu-baseline --no-e2e` for every
variant, `--rounds` times in interleaved order, and prints ms/step, images/s and the dominant kernel's mean
launch time -- numbers for decisions, not for reporting (the JSON lines are kept under gpurun_out/ab/).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def lib_path(name):
    from dsrg_b200 import build
    return build.LIB if name == "base" else os.path.join(build.LIBDIR, "libdsrg_b200_%s.so" % name)


def do_build(specs):
    from dsrg_b200 import build
    for spec in specs:
        name, _, defs = spec.partition(":")
        defines = [d for d in defs.split(",") if d]
        if name == "base":
            assert not defines, "`base` is the default build"
            print(build.build())
        else:
            print(build.build(defines=defines, out=lib_path(name)), defines)


def do_run(names, rounds, extra):
    out = os.path.join(ROOT, "gpurun_out", "ab")
    os.makedirs(out, exist_ok=True)
    res = {n: [] for n in names}
    for r in range(rounds):
        for n in names:
            env = dict(os.environ)
            env.pop("DSRG_B200_LIB", None)
            if n != "base":
                env["DSRG_B200_LIB"] = lib_path(n)
            p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-e2e", "--no-extras"] + extra,
                               env=env, capture_output=True, text=True)
            line = [l for l in p.stdout.splitlines() if l.startswith("{")]
            if p.returncode != 0 or not line:
                print("%-12s round %d FAILED rc=%d %s" % (n, r, p.returncode, p.stderr.strip().splitlines()[-1:]), flush=True)
                continue
            d = json.loads(line[-1])
            tag = "".join(a.strip("-")[:6] for a in extra if not a.isdigit())
            with open(os.path.join(out, "%s_%s%d.json" % (n, tag + "_" if tag else "", r)), "w") as f:
                f.write(line[-1] + "\n")
            res[n].append((d["ms_per_step"], d["value"], d["roofline"]["kernel"], d["roofline"]["avg_launch_ms"]))
            print("%-12s round %d  %.3f ms/step  %.0f images/s  %s %.4f ms/launch" % ((n, r) + res[n][-1]), flush=True)
    print("\nbest of %d:" % rounds)
    for n in names:
        if res[n]:
            b = min(res[n])
            print("  %-12s %.3f ms/step  %.0f images/s  %s %.4f ms/launch" % ((n,) + b))


def main():
    if len(sys.argv) < 3 or sys.argv[1] not in ("build", "run"):
        print(__doc__)
        sys.exit(2)
    if sys.argv[1] == "build":
        do_build(sys.argv[2:])
        return
    args = sys.argv[2:]
    rounds, extra, names = 2, [], []
    i = 0
    while i < len(args):
        if args[i] == "--rounds":
            rounds = int(args[i + 1]); i += 2
        elif args[i] == "--":
            extra = args[i + 1:]; break
        else:
            names.append(args[i]); i += 1
    do_run(names, rounds, extra)


if __name__ == "__main__":
    main()
