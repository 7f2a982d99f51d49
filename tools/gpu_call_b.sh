#!/bin/bash
# round 2, GPU call B: full GPU test-suite, hg19-scale probe of the cooperative pop_min / prefetch, ncu captures of the saturated phase
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( time timeout 1500 python -X faulthandler -m pytest tests -m gpu -q ) > gpurun_out/b_pytest.log 2>&1
tail -6 gpurun_out/b_pytest.log
( time MAPAD_TRACE=1 timeout 1000 python -X faulthandler tests/tools/run_cfg4.py 3.1e9 20000 300 8,32 ) > gpurun_out/b_cfg4.log 2>&1
grep -v "mapad trace" gpurun_out/b_cfg4.log | tail -5
NCU="ncu --set full --clock-control none --import-source on -k regex:k_search_group -c 1"
MAPAD_GROUP=1 timeout 600 $NCU -f -o gpurun_out/r2_g1_cfg3 python tools/profile_saturated.py > gpurun_out/b_ncu_g1.log 2>&1
MAPAD_GROUP=8 timeout 600 $NCU -f -o gpurun_out/r2_g8_cfg3 python tools/profile_saturated.py > gpurun_out/b_ncu_g8.log 2>&1
MAPAD_GROUP=8 MAPAD_PROFILE_ITERS=6000 timeout 600 $NCU -f -o gpurun_out/r2_g8_heavy python tools/profile_saturated.py 86 100 > gpurun_out/b_ncu_g8h.log 2>&1
for f in r2_g1_cfg3 r2_g8_cfg3 r2_g8_heavy; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page source --csv > gpurun_out/${f}_source.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep gpurun_out/*_raw.csv
tail -3 gpurun_out/b_ncu_g8.log
