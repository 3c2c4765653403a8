#!/bin/bash
# committer look-ahead (L1 prefetch of the cold data of exhausted cells): parity, chain profile and bench with and without it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_c3.py -m gpu -q -x 2>&1 | tail -2
for la in 0 20 40; do echo "lookahead $la"; MOSAIC_CHAIN_LOOKAHEAD=$la timeout 120 python scratch/chain_prof.py C3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in d if k in ('ms','ns_per_cell','n_generic','cyc_per_generic','cyc_per_exact','n_exact','cyc_per_fast','fallbacks')})"; done
for la in 0 20 0 20; do MOSAIC_CHAIN_LOOKAHEAD=$la timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_la$la.json 2> gpurun_out/bench_la$la.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench_la$la.json').read().strip().splitlines()[-1])
print('lookahead $la', {k: round(d[k],4) for k in ('ms_per_step',)}, 'e2e', round(d['e2e']['ms_per_step'],3), 'verified', d['verified']['all'], [ (k['kernel'],k['ms']) for k in d['kernels'] if k['kernel'] in ('place',)])
PY
done
