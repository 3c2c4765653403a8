#!/bin/bash
mkdir -p gpurun_out
timeout 280 ncu --set full --import-source on --clock-control none -k regex:nn_topk -s 1 -c 1 -o gpurun_out/prof_topk -f python scratch/topk_time.py > gpurun_out/ncu_topk.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_topk.log
