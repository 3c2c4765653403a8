#!/bin/bash
N=${1:-2}
MOSAIC_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 4 --warmup 3 --no-verify > gpurun_out/trace_${N}gpu.json 2> gpurun_out/trace_${N}gpu.err
grep "mosaic trace" gpurun_out/trace_${N}gpu.err | sed -n '5,8p;17,24p'
