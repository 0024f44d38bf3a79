#!/bin/bash
# round 2, GPU call 3: TMA-staged ET kernel after the footprint / rounding / interleave changes
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_w_et_tma.py -m gpu -q > gpurun_out/c3_tma_tests.log 2>&1; echo "tma tests rc=$?"; tail -12 gpurun_out/c3_tma_tests.log
timeout 300 python tools/et_ab.py > gpurun_out/c3_et_ab.json 2> gpurun_out/c3_et_ab.err; cat gpurun_out/c3_et_ab.err | tail -10
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_x_fullsize_parity.py tests/test_gpu_x_dataparallel.py tests/test_gpu_y_fpn.py tests/test_gpu_zzy_variants.py -m gpu -q > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c3_pytest.log
timeout 300 python bench.py --quick > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/c3_bench.err; python -c "
import json; j=json.load(open('gpurun_out/c3_bench.json')); print(j['ms_per_step'], j['value'], j['e2e']['value'], j['step_stats'], j['parity'], j['roofline']['kernel'], j['roofline']['frac'], [p['us'] for p in j['roofline']['per_stage']])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:et_fuse_tma -c 3 -o gpurun_out/c3_et_tma python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --quick > gpurun_out/c3_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/c3_ncu.log
