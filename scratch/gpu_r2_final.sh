#!/bin/bash
# final evidence of round 2: smoke, launch list of the bench command, the full bench line (with the CPU baseline leg) and the reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/ncu_b.log 2>&1; echo "launch list rc=$?"
python bench.py --steps 30 --warmup 3 > gpurun_out/bench_r02_full.json 2> gpurun_out/bench_r02_full.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r02_full.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], 'verified', d['verified']['all'], 'cpu', d['cpu_baseline']['ms_per_step'])
PY
