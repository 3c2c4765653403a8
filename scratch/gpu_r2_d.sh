#!/bin/bash
mkdir -p gpurun_out
for BR in 0 1 2; do
  echo "== MOSAIC_BAND_ROUNDS=$BR"
  MOSAIC_BAND_ROUNDS=$BR MOSAIC_TRACE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-verify 2> gpurun_out/trace_$BR.err > gpurun_out/bench_br$BR.json
  grep "mosaic trace" gpurun_out/trace_$BR.err | tail -2
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_br$BR.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], [(k['kernel'], round(k['ms'],4)) for k in d.get('kernels',[]) if 'ms' in k])
PY
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
