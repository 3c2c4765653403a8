#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -m pytest tests/test_jpeg_codec.py tests/test_png_codec.py -q -k "cli or round_trips" 2>&1 | tail -2
timeout 280 ncu --set full --import-source on --clock-control none -k regex:nn_topk -s 1 -c 1 -o gpurun_out/prof_topk -f python scratch/topk_time.py > gpurun_out/ncu_topk.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_topk.log
