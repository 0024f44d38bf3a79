#!/bin/bash
# PDL on the conv_tc3 chain: correctness, then A/B of the bench step
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_conv.py tests/test_gpu_z_tc_cascade.py tests/test_gpu_y_fpn.py tests/test_gpu_x_fullsize_parity.py -m gpu -q -x > gpurun_out/c7_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/c7_tests.log
for pdl in 1 0 1 0; do
  MVSTER_TC3_PDL=$pdl timeout 300 python bench.py --quick --no-cpu-baseline > gpurun_out/c7_bench_pdl$pdl.json 2> gpurun_out/c7_bench.err; python -c "
import json; j=json.load(open('gpurun_out/c7_bench_pdl$pdl.json')); print('PDL=$pdl', round(j['ms_per_step'],4), round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['step_stats'])"
done
MVSTER_CUDA_GRAPH=0 timeout 300 python bench.py --quick --no-cpu-baseline > gpurun_out/c7_bench_nograph.json 2>> gpurun_out/c7_bench.err; python -c "
import json; j=json.load(open('gpurun_out/c7_bench_nograph.json')); print('eager launches, PDL=1', round(j['ms_per_step'],4), j['step_stats'])"
