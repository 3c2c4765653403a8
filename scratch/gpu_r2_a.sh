#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'collective_ms', 'exchange_path')})
    print('e2e', d['e2e']); print('verified', d['verified']); print('roofline', d['roofline'])
    print([(r['kernel'], round(r['frac'], 3)) for r in d['rooflines']])
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('no json', e)
PY
