#!/bin/bash
# scaling run on an 8-GPU box: bench at N = 2, 4, 8 (+1)
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${n}gpu.json 2> gpurun_out/scale_${n}gpu.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_${n}gpu.json 2> gpurun_out/scale_${n}gpu.err
  fi
  echo "bench $n rc=$?"; tail -c 400 gpurun_out/scale_${n}gpu.err | grep -v OMP
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/scale_{n}gpu.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('n_gpus', 'value', 'ms_per_step', 'collective_ms', 'exchange_path')}, 'e2e ms', d['e2e']['ms_per_step'], 'verified', d['verified']['all'])
    print([(k['kernel'], k['ms']) for k in d['kernels']])
except Exception as e:
    print('no json', e)
PY
done
