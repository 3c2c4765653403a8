#!/bin/bash
mkdir -p gpurun_out
MOSAIC_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/trace2.json 2> gpurun_out/trace2.err; echo "rc=$?"
grep "mosaic trace rank 0" gpurun_out/trace2.err | tail -2
grep "mosaic trace rank 1" gpurun_out/trace2.err | tail -2
