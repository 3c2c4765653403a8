#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_wave_chain.py -q -k "whole_pipeline" 2>&1 | grep -E "MosaicError|passed|failed" | head -5
timeout 300 python -m pytest tests/test_gpu_wave_chain.py -q 2>&1 | grep -E "MosaicError|passed|failed" | head -5
