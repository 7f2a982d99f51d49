#!/bin/bash
# round 2, GPU call C: sharded device-wide pool + leaner step; cfg3 A/B, first full cfg4 bench line, profiles
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if ! timeout 600 python -X faulthandler -m pytest tests/test_gpu_parity.py -x -q -k "known_answers or bench_reads or retry_launch or limits" > gpurun_out/c_canary.log 2>&1; then
  tail -40 gpurun_out/c_canary.log; echo "CANARY FAILED"; exit 1
fi
tail -2 gpurun_out/c_canary.log
( time AB_STEPS=8 MAPAD_BENCH_INFLIGHT=8 timeout 900 python tools/ab_bench.py --run c_g8:MAPAD_GROUP=8 c_g4:MAPAD_GROUP=4 c_g1:MAPAD_GROUP=1 ) > gpurun_out/c_ab.log 2>&1
tail -6 gpurun_out/c_ab.log
( time MAPAD_TRACE=1 timeout 1500 python bench.py --steps 20 --warmup 3 ) > gpurun_out/c_bench_cfg4.json 2> gpurun_out/c_bench_cfg4.err
tail -c 3000 gpurun_out/c_bench_cfg4.json; tail -5 gpurun_out/c_bench_cfg4.err
NCU="ncu --set full --clock-control none --import-source on -k regex:k_search_group -c 1"
MAPAD_GROUP=8 timeout 600 $NCU -f -o gpurun_out/r2c_g8_cfg3 python tools/profile_saturated.py > gpurun_out/c_ncu_g8.log 2>&1
MAPAD_GROUP=8 MAPAD_PROFILE_ITERS=6000 MAPAD_PROFILE_LIMITS=20000,100000 timeout 600 $NCU -f -o gpurun_out/r2c_g8_limit python tools/profile_saturated.py 86 100 > gpurun_out/c_ncu_g8l.log 2>&1
for f in r2c_g8_cfg3 r2c_g8_limit; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/${f}_source.csv 2>/dev/null
done
cp mapad_b200/libmapad_gpu.so gpurun_out/c_libmapad_gpu.so
ls -la gpurun_out/r2c*
