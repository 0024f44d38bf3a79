#!/bin/bash
# round 2, GPU call 2: TMA-staged ET kernel (tests, A/B, ncu), the tests call 1 did not reach, bench
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_w_et_tma.py -m gpu -x -q > gpurun_out/c2_tma_tests.log 2>&1; echo "tma tests rc=$?"; tail -15 gpurun_out/c2_tma_tests.log
timeout 300 python tools/et_ab.py > gpurun_out/c2_et_ab.json 2> gpurun_out/c2_et_ab.err; cat gpurun_out/c2_et_ab.err | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_x_dataparallel.py tests/test_gpu_y_fpn.py tests/test_gpu_z_tc_cascade.py tests/test_gpu_zz_fusion.py tests/test_gpu_zzy_variants.py tests/test_gpu_zzz_et_backward.py tests/test_gpu_zzzz_prefetch.py -m gpu -q > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c2_pytest.log
timeout 300 python bench.py --quick --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "bench rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/c2_bench.json')); print(j['ms_per_step'], j['value'], j['step_stats'], j['roofline']['kernel'], j['roofline']['frac'], [p['us'] for p in j['roofline']['per_stage']])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:et_fuse_tma -c 3 -o gpurun_out/c2_et_tma python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --quick > gpurun_out/c2_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/c2_ncu.log
