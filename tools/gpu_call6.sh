#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_w_et_tma.py tests/test_gpu_x_fullsize_parity.py -m gpu -q > gpurun_out/c6_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/c6_tests.log
timeout 300 python tools/et_ab.py > gpurun_out/c6_et_ab.json 2> gpurun_out/c6_et_ab.err; cat gpurun_out/c6_et_ab.err | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"et_fuse_tma|et_tile_boxes" -c 6 -o gpurun_out/c6_et_tma python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --quick > gpurun_out/c6_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/c6_ncu.log
