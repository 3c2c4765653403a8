#!/bin/bash
mkdir -p gpurun_out
CHAIN_V=1 timeout 200 python -u scratch/chain_prof.py C3 > gpurun_out/chain_prof_v1.log 2>&1; tail -3 gpurun_out/chain_prof_v1.log
