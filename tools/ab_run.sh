#!/bin/bash
# A/B tuning run on the GPU box: parity subset, latency probe, bench variants (resident pass only).
set -u
mkdir -p gpurun_out
export MAPAD_BENCH_INDEX_CACHE=/tmp/cfg3idx MAPAD_BENCH_SKIP_E2E=1 MAPAD_BENCH_DISTINCT_CHUNKS=6
B="python bench.py --steps 16 --warmup 3 --no-cpu-baseline"
( MAPAD_POOL_STAGES=64,512 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "simulated or retry or bench_reads" 2>&1 | tail -3 ) > gpurun_out/ab_tests.log
timeout 600 python tools/latency_probe.py > gpurun_out/ab_probe.log 2>&1
timeout 400 $B > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err
MAPAD_GPU_LIB=$PWD/mapad_b200/variants/libmapad_mb6.so timeout 400 $B > gpurun_out/ab_mb6.json 2> gpurun_out/ab_mb6.err
MAPAD_GPU_LIB=$PWD/mapad_b200/variants/libmapad_mb8.so MAPAD_POOL_THREADS=12288 timeout 400 $B > gpurun_out/ab_mb8.json 2> gpurun_out/ab_mb8.err
MAPAD_POOL_STAGES=16384 timeout 400 $B > gpurun_out/ab_st16k.json 2> gpurun_out/ab_st16k.err
tail -n 3 gpurun_out/ab_tests.log
cat gpurun_out/ab_probe.log | tail -n 20
for f in base mb6 mb8 st16k; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), d["ms_per_step"], d["config"].get("retry_lane_reads"), d["config"].get("handle_done_s"))
except Exception as e:
    print("$f failed", e)
PY
done
