#!/bin/bash
# C4 (similarity matrix, 50,000 materials) and C5 (10^6 cells, 100,000 materials) on N GPUs of one box
n=$1
mkdir -p gpurun_out
for cfg in C4 C5; do
  steps=5; [ $cfg = C5 ] && steps=3
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --config $cfg --gpus 1 --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/scale_${cfg}_${n}gpu.json 2> gpurun_out/scale_${cfg}_${n}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --config $cfg --gpus $n --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/scale_${cfg}_${n}gpu.json 2> gpurun_out/scale_${cfg}_${n}gpu.err
  fi
  echo "bench $cfg N=$n rc=$?"; tail -c 300 gpurun_out/scale_${cfg}_${n}gpu.err | grep -v OMP
  python - $cfg $n <<'PY'
import json, sys
cfg, n = sys.argv[1:3]
try:
    d = json.loads(open(f'gpurun_out/scale_{cfg}_{n}gpu.json').read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step', 'collective_ms')}, 'e2e ms', d['e2e']['ms_per_step'], 'verified', (d.get('verified') or {}).get('all'), d.get('with_allgather'))
    print([(k['kernel'], k['ms']) for k in d['kernels']])
except Exception as e:
    print('no json', e)
PY
done
