#!/bin/bash
n=$1
mkdir -p gpurun_out
if [ $n -eq 4 ]; then timeout 600 python -m pytest tests/test_gpu_sharded.py -q -x 2>&1 | tail -3; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_C3_${n}gpu_r02.json 2> gpurun_out/scale_C3_${n}gpu_r02.err; echo "bench rc=$?"
python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/scale_C3_{n}gpu_r02.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'collective_ms')}, 'e2e ms', d['e2e']['ms_per_step'], 'verified', (d.get('verified') or {}).get('all'))
    print([(k['kernel'], k['ms']) for k in d['kernels']])
except Exception as e:
    print('no json', e); print(open(f'gpurun_out/scale_C3_{n}gpu_r02.err').read()[-1500:])
PY
