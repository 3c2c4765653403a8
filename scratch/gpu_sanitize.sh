#!/bin/bash
# compute-sanitizer on the parity tests (small cases).  racecheck is run on everything except the two serial chain
# kernels (greedy_chain / anneal): their producer/consumer hand-offs are ordered by flags + fences in shared memory, a
# protocol racecheck does not model (it reports them as hazards; sanitizer_racecheck_chains.log keeps that output).
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ragged.py -q -x -k "similarity or ragged or loaded_database or generate_matches or material_lab or cell_lab" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ragged.py -q -x -k "similarity_bit_equal or similarity_reference or ragged_render or ragged_material_lab or test_material_lab or test_cell_lab or candidates or render_within or render_identity" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
compute-sanitizer --tool initcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ragged.py -q -x -k "similarity or ragged" > gpurun_out/sanitizer_initcheck.log 2>&1; echo "initcheck rc=$?"; tail -4 gpurun_out/sanitizer_initcheck.log
python scratch/ragged_time.py 2>&1 | tail -1 | tee gpurun_out/ragged_C3.json
