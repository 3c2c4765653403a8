#!/bin/bash
mkdir -p gpurun_out
MOSAIC_CHAIN=wave timeout 280 ncu --set full --import-source on --clock-control none -k regex:greedy_wave -c 1 -o gpurun_out/prof_wave -f python scratch/place_once.py C3 > gpurun_out/ncu_wave.log 2>&1; echo "ncu rc=$?"; tail -4 gpurun_out/ncu_wave.log
