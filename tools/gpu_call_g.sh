#!/bin/bash
# round 2, GPU call G: latency hiding for the deep heaps of hg19-scale reads (one read per warp): occ prefetch, heap-line
# prefetch, 171 heap lines in shared memory, two reads per warp, 24 resident warps per SM.  16 chunks of 25 000 reads in flight.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if ! timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "known_answers or retry_launch or limits" > gpurun_out/g_canary.log 2>&1; then
  tail -40 gpurun_out/g_canary.log; echo "CANARY FAILED"; exit 1
fi
tail -2 gpurun_out/g_canary.log
( time MAPAD_TRACE=1 timeout 1000 python tools/probe_cfg4.py 16 25000 pf2t171:MAPAD_TRICKLE_PREFETCH=2,MAPAD_TOPL=171 pf3t171:MAPAD_TRICKLE_PREFETCH=3,MAPAD_TOPL=171 g16pf2:MAPAD_GROUP=16,MAPAD_TRICKLE_PREFETCH=2 w24pf2:MAPAD_GROUPS_PER_SM=24,MAPAD_TRICKLE_PREFETCH=2 ) > gpurun_out/g_probe.log 2> gpurun_out/g_probe.err
grep -v "^\[mapad" gpurun_out/g_probe.log | tail -6
