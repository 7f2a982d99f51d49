#!/bin/bash
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3 ) > gpurun_out/t3_tests.log
export MAPAD_BENCH_INDEX_CACHE=/tmp/cfg3idx MAPAD_BENCH_SKIP_E2E=1 MAPAD_BENCH_DISTINCT_CHUNKS=6 MAPAD_TRACE=1
B="python bench.py --steps 16 --warmup 3 --no-cpu-baseline"
timeout 400 $B > gpurun_out/t3_few.json 2> gpurun_out/t3_few.err
MAPAD_WARP_FEW=0 timeout 400 $B > gpurun_out/t3_nofew.json 2> gpurun_out/t3_nofew.err
cat gpurun_out/t3_tests.log
for f in few nofew; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t3_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), d["ms_per_step"], d["config"].get("retry_lane_reads"), d["config"].get("handle_done_s"))
except Exception as e:
    print("$f failed", e)
PY
done
