#!/bin/bash
# round 2, 2-GPU smoke of the multi-GPU paths after the device-wide pool / launch-share changes: the NCCL sharding test, the
# one-process two-GPU CLI test, and a short torchrun bench on cfg3 (weak and strong scaling lines)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( time timeout 400 python -X faulthandler -m pytest tests/test_gpu_multi.py -x -q ) > gpurun_out/m2_pytest.log 2>&1
tail -3 gpurun_out/m2_pytest.log
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload cfg3 --steps 4 --warmup 1 --no-cpu-baseline ) > gpurun_out/m2_bench_cfg3_weak.json 2> gpurun_out/m2_bench_cfg3_weak.err
tail -c 700 gpurun_out/m2_bench_cfg3_weak.json; tail -3 gpurun_out/m2_bench_cfg3_weak.err
