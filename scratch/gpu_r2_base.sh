#!/bin/bash
# round-2 baseline: parity tests, chain profile, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python scratch/chain_prof.py C1 C2 C3 > gpurun_out/chain_prof.log 2>&1; tail -3 gpurun_out/chain_prof.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
