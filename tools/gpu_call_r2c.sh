#!/bin/bash
# round 2, call B2: TMA box-shape / depth / prefetch micro-benchmark, A/B of 6 staging boxes in conv_tc3 (variant library), the
# bf16 end-to-end tests
set -u
mkdir -p gpurun_out
tools/_build/tma_microbench 8 512 640 > gpurun_out/tma_microbench.txt 2>&1; tail -22 gpurun_out/tma_microbench.txt
for lib in "" tools/_build/libmvster_nf6.so; do
  echo "== lib: ${lib:-default}"
  for c in "v3 16 16 3 3 1 1 4 256 320 h16" "v3 16 16 1 3 1 5 1 512 640 h16" "v3 32 32 3 3 1 1 4 128 160 skip h16" "v3 8 16 1 3 2 1 4 512 640 h16"; do
    MVSTER_LIB_PATH=$lib timeout 90 python tests/tc_conv_check.py $c 2>&1 | tail -1 | python -c "import json,sys; r=json.loads(sys.stdin.read()); print(' '.join(r['case']).ljust(44), 'rel %.1e' % r['rel'], 'us %.1f' % r['us_tc'])"
  done
  for i in 1 2; do
    MVSTER_LIB_PATH=$lib timeout 300 python bench.py --quick --no-cpu-baseline --skip-e2e --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); s=j['step_stats']; print('bench ms/step %.4f median %.4f min %.4f' % (j['ms_per_step'], s['median_ms'], s['min_ms']))"
  done
done
timeout 300 python -m pytest tests/test_gpu_zzzzz_bf16.py -q -x -k "forward or rejects" 2>&1 | tail -8
