#!/bin/bash
# 8 ranks: which part of the step path produces the 3-19 ms stalls seen after ~10 queued steps?
set -u
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 30 --warmup 5 --quick --skip-e2e > gpurun_out/mg8d_$name.json 2> gpurun_out/mg8d_$name.err
  python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/mg8d_$name.json').read().strip().splitlines()[-1]); s=j['step_stats']
    print('$name', 'value', round(j['value'],1), 'ms', round(j['ms_per_step'],3), 'min', round(s['min_ms'],3), 'med', round(s['median_ms'],3), 'max', round(s['max_ms'],3), s['steps_ms'])
except Exception as e: print('$name failed', e)
PY
}
run default A=1
run static_outputs MVSTER_GRAPH_STATIC_OUTPUTS=1
run eager MVSTER_CUDA_GRAPH=0
run sleep_sync MVSTER_BENCH_SYNC=sleep
run no_overlap MVSTER_OVERLAP=0
run omp4 OMP_NUM_THREADS=4
