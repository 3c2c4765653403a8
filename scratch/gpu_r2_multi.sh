#!/bin/bash
# multi-GPU round: sharded parity test + bench at N = 2 (.. $1)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "pytest sharded rc=$?"; tail -15 gpurun_out/pytest_sharded.log
for n in $(seq 2 $N); do
  [ $n -eq 3 ] && continue; [ $n -gt 4 ] && [ $n -lt 8 ] && continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "bench $n rc=$?"; tail -c 1200 gpurun_out/bench_${n}gpu.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/bench_{n}gpu.json').read().strip().splitlines()[-1])
    print({k: d[k] for k in ('n_gpus', 'value', 'ms_per_step', 'gpu_launches', 'collective_ms', 'exchange_path')})
    print('e2e', d['e2e']); print('verified', {k: v for k, v in d['verified'].items() if k != 'how'})
    print([(k['kernel'], k['ms']) for k in d['kernels']])
except Exception as e:
    print('no json', e)
PY
done
