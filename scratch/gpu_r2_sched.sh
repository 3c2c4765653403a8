#!/bin/bash
# what slows the chain inside the step (2.81 ms against ~2.55 alone)?  the schedule knobs, one bench line each
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-verify > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; python - <<PY
import json
d = json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), [(k['kernel'],k['ms']) for k in d['kernels']])
PY
}
run default A=1
run norender MOSAIC_SCHED=2
run rounds4 MOSAIC_BAND_ROUNDS=4
run rounds2 MOSAIC_BAND_ROUNDS=2
run unbanded MOSAIC_SCHED=1
