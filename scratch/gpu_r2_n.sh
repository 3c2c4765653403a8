#!/bin/bash
CHAIN_V=1 timeout 200 python -u scratch/chain_prof.py C1 C2 C3 2>&1 | cut -c1-900 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c3.py -q -x 2>&1 | tail -3
