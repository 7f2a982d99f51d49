#!/bin/bash
set -u
mkdir -p gpurun_out
export MAPAD_BENCH_INDEX_CACHE=/tmp/cfg3idx MAPAD_BENCH_DISTINCT_CHUNKS=6 MAPAD_BENCH_SKIP_E2E=1
B="python bench.py --no-cpu-baseline"
MAPAD_POOL_THREADS=8192 timeout 400 $B > gpurun_out/t5_i32t8k.json 2> gpurun_out/t5_i32t8k.err
MAPAD_POOL_THREADS=14336 timeout 400 $B > gpurun_out/t5_i32t14k.json 2> gpurun_out/t5_i32t14k.err
MAPAD_BENCH_INFLIGHT=16 MAPAD_POOL_THREADS=8192 timeout 400 $B > gpurun_out/t5_i16t8k.json 2> gpurun_out/t5_i16t8k.err
for f in i32t8k i32t14k i16t8k; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t5_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), d["ms_per_step"], d["config"].get("retry_lane_reads"), d["config"].get("handle_done_s"))
except Exception as e:
    print("$f failed", e)
PY
done
