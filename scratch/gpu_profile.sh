#!/bin/bash
# (every profiler call is bounded: a kernel that polls for work of kernels not launched yet dead-locks under ncu's
# serialisation — that cost 25 GPU-minutes once)
# ncu evidence of one round: launch list of the bench command + one full-set capture per kernel
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -s 30 -c 11 -o gpurun_out/prof_step -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
