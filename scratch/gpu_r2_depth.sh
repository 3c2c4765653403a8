#!/bin/bash
# ring depth of the placement chain (64 -> 128 slots): parity, chain alone, the step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c3.py -m gpu -q -x 2>&1 | tail -2
timeout 120 python scratch/chain_prof.py C3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in d if k in ('ms','ns_per_cell','cyc_wait_producer','cyc_wait_committer','cyc_total','cyc_per_generic','cyc_per_exact','cyc_per_fast')})"
for i in 1 2; do timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_depth.json 2> gpurun_out/bench_depth.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench_depth.json').read().strip().splitlines()[-1])
print('depth128', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'verified', d['verified']['all'], [(k['kernel'],k['ms']) for k in d['kernels']])
PY
done
