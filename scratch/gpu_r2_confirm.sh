#!/bin/bash
# confirmation of the tree after the WebP codec went in: GPU tests, smoke, the bench line and the reference arm
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_confirm.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_confirm.log
timeout 200 python -m pytest tests/test_webp_codec.py -q -k cli 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r02_confirm.json 2> gpurun_out/bench_r02_confirm.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r02_confirm.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], 'verified', d['verified']['all'], 'cpu', d['cpu_baseline']['ms_per_step'])
PY
