"""A/B harness for tuning runs (measurement tool; the numbers it prints are not bench values).

A variant is  NAME[:KEY=VALUE,...][:-DFLAG=1,...]  — environment for bench.py and/or extra nvcc defines.

  here (no GPU):   python tools/ab_bench.py --build  base  cc::-DMAPAD_COMPACT_CAND=1  shift::-DMAPAD_HEAP_SHIFT=1 \
                                                      both::-DMAPAD_COMPACT_CAND=1,-DMAPAD_HEAP_SHIFT=1  t3:MAPAD_POOL_THREADS=3072
                   (compiles one library per distinct set of defines into mapad_b200/variants/, which travels with gpurun)
  on the GPU box:  gpurun --timeout 1200 -- 'python tools/ab_bench.py --run base cc shift both t3 ...same specs...'

Each variant runs `bench.py --steps 16 --warmup 3 --no-cpu-baseline` with the resident pass only, the index files cached in
/tmp between runs, and MAPAD_TRACE=1 so that the lane timeline lands in gpurun_out/ab_<name>.err.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = os.path.join(ROOT, "mapad_b200", "variants")


def parse(spec):
    parts = spec.split(":")
    name = parts[0]
    env = dict(kv.split("=", 1) for kv in parts[1].split(",") if kv) if len(parts) > 1 else {}
    defs = [d for d in parts[2].split(",") if d] if len(parts) > 2 else []
    return name, env, defs


def lib_for(defs):
    if not defs:
        return None
    tag = "_".join(d.replace("-D", "").replace("=", "") for d in sorted(defs))
    return os.path.join(VARIANTS, "libmapad_%s.so" % tag)


def main():
    mode, specs = sys.argv[1], [parse(s) for s in sys.argv[2:]]
    if mode == "--build":
        sys.path.insert(0, ROOT)
        from mapad_b200 import build
        os.makedirs(VARIANTS, exist_ok=True)
        build.build()
        for defs in {tuple(sorted(d)) for _, _, d in specs if d}:
            out = lib_for(list(defs))
            print("building", out)
            build.build(out=out, extra=list(defs))
        return
    assert mode == "--run", mode
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    steps = os.environ.get("AB_STEPS", "16")
    rows = []
    for name, env, defs in specs:
        e = dict(os.environ, MAPAD_BENCH_INDEX_CACHE="/tmp/ab_index", MAPAD_BENCH_SKIP_E2E="1", MAPAD_BENCH_DISTINCT_CHUNKS="6", MAPAD_TRACE="1")
        e.update(env)
        lib = lib_for(defs)
        if lib:
            e["MAPAD_GPU_LIB"] = lib
        err = open(os.path.join(ROOT, "gpurun_out", "ab_%s.err" % name), "w")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "3", "--no-cpu-baseline", "--workload", os.environ.get("AB_WORKLOAD", "cfg3")],
                           env=e, cwd=ROOT, stdout=subprocess.PIPE, stderr=err, text=True, timeout=900)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        open(os.path.join(ROOT, "gpurun_out", "ab_%s.json" % name), "w").write(line + "\n")
        try:
            d = json.loads(line)
            done = d["config"].get("handle_done_s") or [0]
            rows.append((name, round(d["value"]), round(d["ms_per_step"], 1), int(d["config"].get("retry_lane_reads", 0)), done[0], done[-1]))
        except Exception as ex:  # noqa: BLE001
            rows.append((name, "failed: %s" % ex, r.returncode, 0, 0, 0))
    print("%-12s %10s %10s %8s %8s %8s" % ("variant", "reads/s", "ms/step", "retry", "first_s", "last_s"))
    for row in rows:
        print("%-12s %10s %10s %8s %8s %8s" % row)


if __name__ == "__main__":
    main()
