#!/bin/bash
mkdir -p gpurun_out
MOSAIC_WAVE_DEBUG=1 timeout 300 python -u -m pytest tests/test_gpu_wave_chain.py -x -q > gpurun_out/pytest_wave.log 2>&1; echo "pytest rc=$?"; tail -45 gpurun_out/pytest_wave.log
CHAIN_V=3 timeout 120 python -u scratch/chain_prof.py C3 > gpurun_out/wave_prof.log 2>&1; echo "prof rc=$?"; tail -5 gpurun_out/wave_prof.log
