#!/bin/bash
# ncu evidence of round 2: launch list of the bench command + one full-set capture of every kernel (stage-wise, whole workload)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/ncu_b.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -o gpurun_out/prof_stages_r02 -f python scratch/stages_once.py C3 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
tail -2 gpurun_out/ncu_full.log
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r02_full.json 2> gpurun_out/bench_r02_full.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; echo "ref rc=$?"
