#!/bin/bash
mkdir -p gpurun_out
timeout 280 ncu --set full --import-source on --clock-control none -k regex:lab_mean_cells -c 1 -o gpurun_out/prof_celllab -f python scratch/lab_once.py > gpurun_out/ncu_celllab.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_celllab.log
