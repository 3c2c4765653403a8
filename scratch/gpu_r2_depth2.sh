#!/bin/bash
# ring depth of the placement chain as a template parameter: parity at the default (deepest that fits), the step at 256 / 128 / 64
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c3.py -m gpu -q -x 2>&1 | tail -2
for d in 256 128 64; do MOSAIC_CHAIN_DEPTH=$d timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_depth$d.json 2> gpurun_out/bench_depth$d.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench_depth$d.json').read().strip().splitlines()[-1])
print('depth $d', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'verified', d['verified']['all'], [(k['kernel'],k['ms']) for k in d['kernels'] if k['kernel'] in ('place','topk','render')])
PY
done
