#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scratch/topk_time.py 2>&1 | tail -1
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -q -x -k "candidates" > gpurun_out/sanitizer_racecheck_topk_r02.log 2>&1; echo "racecheck topk rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_topk_r02.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c3.py tests/test_gpu_large.py -q -x 2>&1 | tail -2
