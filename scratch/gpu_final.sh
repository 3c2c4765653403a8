#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
python bench.py --config C1 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C1.json 2>> gpurun_out/bench.err; tail -c 200 gpurun_out/bench_C1.json
python bench.py --config C2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2.json 2>> gpurun_out/bench.err; tail -c 200 gpurun_out/bench_C2.json
python scratch/ragged_time.py 2>&1 | tail -1 | tee gpurun_out/ragged_C3.json
