#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_x_fullsize_parity.py tests/test_gpu_parity.py tests/test_gpu_x_dataparallel.py tests/test_gpu_x_configs.py -m gpu -q > gpurun_out/c10_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c10_tests.log
for i in 1 2; do
timeout 400 python bench.py --quick --no-cpu-baseline > gpurun_out/c10_bench$i.json 2> gpurun_out/c10_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c10_bench.err
python -c "
import json; j=json.load(open('gpurun_out/c10_bench$i.json')); s=j['step_stats']; s.pop('steps_ms',None)
print(round(j['ms_per_step'],4), round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['e2e'].get('frames_in_flight'), s)"
done
MVSTER_CUDA_GRAPH=0 timeout 300 python bench.py --quick --no-cpu-baseline --skip-e2e > gpurun_out/c10_bench_eager.json 2>> gpurun_out/c10_bench.err; python -c "
import json; j=json.load(open('gpurun_out/c10_bench_eager.json')); print('eager', round(j['ms_per_step'],4))"
