#!/bin/bash
# round 2, GPU call D: admission control + occupancy variants (cfg3 A/B), cfg4 tuning run (resident pass only)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if ! timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "known_answers or bench_reads or retry_launch or limits or cfg5" > gpurun_out/d_canary.log 2>&1; then
  tail -40 gpurun_out/d_canary.log; echo "CANARY FAILED"; exit 1
fi
tail -2 gpurun_out/d_canary.log
( time AB_STEPS=8 MAPAD_BENCH_INFLIGHT=8 timeout 900 python tools/ab_bench.py --run d_base mb20:MAPAD_GROUPS_PER_SM=80:-DMAPAD_GROUP_MIN_BLOCKS=20 mb24:MAPAD_GROUPS_PER_SM=96:-DMAPAD_GROUP_MIN_BLOCKS=24 pf::-DMAPAD_TRICKLE_PREFETCH=1 d_g4:MAPAD_GROUP=4 ) > gpurun_out/d_ab.log 2>&1
tail -8 gpurun_out/d_ab.log
( time MAPAD_TRACE=1 MAPAD_BENCH_SKIP_E2E=1 timeout 1200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/d_bench_cfg4.json 2> gpurun_out/d_bench_cfg4.err
tail -c 2500 gpurun_out/d_bench_cfg4.json; tail -3 gpurun_out/d_bench_cfg4.err
