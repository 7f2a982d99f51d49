#!/bin/bash
# round 2, GPU call A: parity of the group kernel, cfg3 A/B of its variants, hg19-scale probe
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
# canary: one small parity test; stop early (and cheaply) if the library is broken
if ! timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "known_answers or bench_reads" > gpurun_out/a_canary.log 2>&1; then
  tail -40 gpurun_out/a_canary.log; echo "CANARY FAILED"; exit 1
fi
( time timeout 1500 python -X faulthandler -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
tail -5 gpurun_out/a_pytest.log
( time AB_STEPS=8 MAPAD_BENCH_INFLIGHT=8 timeout 1500 python tools/ab_bench.py --run g8:MAPAD_GROUP=8 g8t43:MAPAD_GROUP=8,MAPAD_TOPL=43 g1:MAPAD_GROUP=1 \
   g1t11:MAPAD_GROUP=1,MAPAD_TOPL=11,MAPAD_GROUPS_PER_SM=256 g32:MAPAD_GROUP=32 g4:MAPAD_GROUP=4 ) > gpurun_out/a_ab.log 2>&1
tail -12 gpurun_out/a_ab.log
( time MAPAD_TRACE=1 timeout 1300 python -X faulthandler tests/tools/run_cfg4.py 3.1e9 20000 1000 8,32,1 ) > gpurun_out/a_cfg4.log 2>&1
tail -8 gpurun_out/a_cfg4.log
