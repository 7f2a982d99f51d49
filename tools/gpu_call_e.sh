#!/bin/bash
# round 2, GPU call E: fair-share launches; G = 8 vs 4 (and per-launch share) on hg19-scale chunks; cfg3 check
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if ! timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "known_answers or bench_reads or retry_launch or limits or simulated" > gpurun_out/e_canary.log 2>&1; then
  tail -40 gpurun_out/e_canary.log; echo "CANARY FAILED"; exit 1
fi
tail -2 gpurun_out/e_canary.log
( time AB_STEPS=16 MAPAD_BENCH_INFLIGHT=16 timeout 600 python tools/ab_bench.py --run e_g8:MAPAD_GROUP=8 e_g4:MAPAD_GROUP=4 ) > gpurun_out/e_ab.log 2>&1
tail -4 gpurun_out/e_ab.log
( time MAPAD_TRACE=1 timeout 1500 python tools/probe_cfg4.py 12 25000 g8:MAPAD_GROUP=8 g4:MAPAD_GROUP=4 g8t43:MAPAD_GROUP=8,MAPAD_TOPL=43 ) > gpurun_out/e_probe.log 2> gpurun_out/e_probe.err
grep -v "^\[mapad" gpurun_out/e_probe.log | tail -5
