#!/bin/bash
timeout 200 python scratch/topk_time.py 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c3.py tests/test_gpu_large.py -q -x 2>&1 | tail -3
