#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_chain.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_chain.log
timeout 300 python scratch/chain_prof.py C1 C2 C3 > gpurun_out/chain_prof.log 2>&1; tail -3 gpurun_out/chain_prof.log
