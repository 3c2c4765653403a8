#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
python scratch/memset_bw.py > gpurun_out/memset_bw.log 2>&1; cat gpurun_out/memset_bw.log
