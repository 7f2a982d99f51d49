#!/bin/bash
# round 2, GPU call H (final): full GPU test-suite, smoke, the default bench line (cfg4) and the reference arm as the driver
# runs them, the ncu launch list of a short default-workload bench, ncu metrics of a saturated hg19-scale search launch,
# and — time permitting — a same-box A/B of 24 resident warps per SM.  Every step is bounded by its own timeout.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
BUDGET=${CALL_BUDGET_S:-1950}
left() { local want=$1; local rest=$((BUDGET - SECONDS - 20)); if [ $rest -lt 60 ]; then echo 0; elif [ $rest -lt $want ]; then echo $rest; else echo $want; fi; }
T=$(left 1200); [ $T -gt 0 ] && ( time timeout $T python -X faulthandler -m pytest tests -m gpu -q ) > gpurun_out/h_pytest.log 2>&1
tail -4 gpurun_out/h_pytest.log
T=$(left 300); [ $T -gt 0 ] && ( time timeout $T python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/h_smoke.log 2>&1
tail -4 gpurun_out/h_smoke.log
T=$(left 860); [ $T -gt 0 ] && ( time timeout $T python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/h_bench_cfg4.json 2> gpurun_out/h_bench_cfg4.err
tail -c 1800 gpurun_out/h_bench_cfg4.json; tail -4 gpurun_out/h_bench_cfg4.err
T=$(left 600); [ $T -gt 0 ] && ( time timeout $T python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/h_bench_ref.json 2> gpurun_out/h_bench_ref.err
tail -c 900 gpurun_out/h_bench_ref.json; tail -4 gpurun_out/h_bench_ref.err
# launch list of the step's kernels (serialised, cold cache: shares only), default workload with small chunks
T=$(left 420); [ $T -gt 0 ] && ( time timeout $T ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_(penalties|darray|search_group|epilogue)" -c 200 --csv \
    --log-file gpurun_out/h_launches_cfg4.csv python bench.py --steps 3 --warmup 1 --batch 500 --no-cpu-baseline ) > gpurun_out/h_launches_bench.log 2>&1
tail -3 gpurun_out/h_launches_bench.log | cut -c1-600
# DRAM traffic and issue metrics of ONE hg19-scale search launch (one read per warp), every warp busy for 40 000 expansions
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
T=$(left 480); [ $T -gt 0 ] && ( time PROFILE_WORKLOAD=cfg4 PROFILE_READS=25000 MAPAD_PROFILE_ITERS=40000 MAPAD_WS_BYTES=$((40<<30)) timeout $T ncu --metrics $M --clock-control none \
    -k regex:k_search_group -c 1 --csv --log-file gpurun_out/h_cfg4_saturated_metrics.csv python tools/profile_saturated.py 86 100 ) > gpurun_out/h_ncu_cfg4.log 2>&1
tail -3 gpurun_out/h_ncu_cfg4.log | cut -c1-400
# same-box tuning probe, 16 chunks of 25 000 reads as in calls F0 / G (base there: 184 s): 24 resident warps per SM; 11 heap
# lines in shared memory (more L1 for the pooled heap lines)
T=$(left 600); [ $T -gt 0 ] && ( time timeout $T python tools/probe_cfg4.py 16 25000 w24:MAPAD_GROUPS_PER_SM=24 t11:MAPAD_TOPL=11 ) > gpurun_out/h_probe.log 2> gpurun_out/h_probe.err
tail -3 gpurun_out/h_probe.log | cut -c1-700
