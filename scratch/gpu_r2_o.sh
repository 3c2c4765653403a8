#!/bin/bash
mkdir -p gpurun_out
timeout 280 ncu --set full --import-source on --clock-control none -k regex:anneal_kernel -s 1 -c 1 -o gpurun_out/prof_anneal_r02 -f python scratch/anneal_once.py > gpurun_out/ncu_anneal.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_anneal.log
