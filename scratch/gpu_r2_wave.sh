#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wave_chain.py -x -q > gpurun_out/pytest_wave.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_wave.log
CHAIN_V=3 timeout 300 python scratch/chain_prof.py C1 C2 C3 > gpurun_out/wave_prof.log 2>&1; echo "prof rc=$?"; tail -5 gpurun_out/wave_prof.log
