#!/bin/bash
set -u
mkdir -p gpurun_out
export MAPAD_BENCH_INDEX_CACHE=/tmp/cfg3idx MAPAD_BENCH_SKIP_E2E=1 MAPAD_BENCH_DISTINCT_CHUNKS=6 MAPAD_TRACE=1
B="python bench.py --steps 16 --warmup 3 --no-cpu-baseline"
timeout 400 $B > gpurun_out/tr_base.json 2> gpurun_out/tr_base.err
MAPAD_POOL_THREADS=4736 timeout 400 $B > gpurun_out/tr_t4736.json 2> gpurun_out/tr_t4736.err
for f in base t4736; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/tr_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), d["ms_per_step"], d["config"].get("retry_lane_reads"), d["config"].get("handle_done_s"))
except Exception as e:
    print("$f failed", e)
PY
done
