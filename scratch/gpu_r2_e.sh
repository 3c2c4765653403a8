#!/bin/bash
mkdir -p gpurun_out
for S in 0 1 2; do
  echo "== MOSAIC_SCHED=$S"
  MOSAIC_SCHED=$S MOSAIC_TRACE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-verify 2> gpurun_out/trace_s$S.err > gpurun_out/bench_s$S.json
  grep "mosaic trace" gpurun_out/trace_s$S.err | sed -n 9,10p
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_s$S.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], [(k['kernel'], round(k['ms'],4)) for k in d.get('kernels',[]) if 'ms' in k])
PY
done
