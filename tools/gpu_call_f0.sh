#!/bin/bash
# round 2, GPU call F0: hg19-scale probe of one-read-per-warp + launch share vs G = 8 full grids (16 chunks in flight)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if ! timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "known_answers or bench_reads or retry_launch or limits" > gpurun_out/f0_canary.log 2>&1; then
  tail -40 gpurun_out/f0_canary.log; echo "CANARY FAILED"; exit 1
fi
tail -2 gpurun_out/f0_canary.log
( time MAPAD_TRACE=1 timeout 1500 python tools/probe_cfg4.py 16 25000 g32share:MAPAD_GROUP=32 g32full:MAPAD_GROUP=32,MAPAD_LAUNCH_SHARE=1 g8full:MAPAD_GROUP=8,MAPAD_LAUNCH_SHARE=1 ) > gpurun_out/f0_probe.log 2> gpurun_out/f0_probe.err
grep -v "^\[mapad" gpurun_out/f0_probe.log | tail -5
