#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_x_fullsize_parity.py tests/test_gpu_y_fpn.py tests/test_gpu_x_dataparallel.py -m gpu -q > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c9_tests.log
for i in 1 2; do
timeout 400 python bench.py > gpurun_out/c9_bench$i.json 2> gpurun_out/c9_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c9_bench.err
python -c "
import json; j=json.load(open('gpurun_out/c9_bench$i.json')); s=j['step_stats']; s.pop('steps_ms',None)
print(round(j['ms_per_step'],4), round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['e2e'].get('frames_in_flight'), 'serial', round(j['e2e']['serial_value'],1), s, 'parity', j['parity']['ok'])
print(j['roofline']['kernel'], round(j['roofline']['frac'],3), j['roofline']['traffic'], 'eager', round(j['gpu_eager_baseline']['tf32_on']['value'],1), round(j['gpu_eager_baseline']['tf32_off']['value'],1))"
done
