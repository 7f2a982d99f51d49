#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 ) > gpurun_out/t4_tests.log
export MAPAD_BENCH_INDEX_CACHE=/tmp/cfg3idx MAPAD_BENCH_DISTINCT_CHUNKS=6
MAPAD_BENCH_SKIP_E2E=1 MAPAD_TRACE=1 timeout 400 python bench.py --steps 16 --warmup 3 --no-cpu-baseline > gpurun_out/t4_tr16.json 2> gpurun_out/t4_tr16.err
MAPAD_BENCH_DISTINCT_CHUNKS=12 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/t4_full.json 2> gpurun_out/t4_full.err
cat gpurun_out/t4_tests.log
for f in tr16 full; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t4_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["e2e"]["value"] or 0), d["ms_per_step"], d["config"].get("retry_lane_reads"), d["config"].get("handle_done_s"))
except Exception as e:
    print("$f failed", e)
PY
done
