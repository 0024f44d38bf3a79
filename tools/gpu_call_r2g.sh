#!/bin/bash
# round 2, call B6: A/B of a deeper operand ring for the packed-operand kernels (variant library)
set -u
for lib in "" tools/_build/libmvster_deep.so; do
  echo "== lib: ${lib:-default}"
  for c in "v3 16 16 3 3 1 1 4 256 320 h16 p16" "v3 16 16 3 3 1 1 4 256 320 b16 p16" "v3 32 32 3 3 1 1 4 128 160 skip h16 p16" "v3 8 16 1 3 2 1 4 512 640 h16 p16" "d3 16 8 1 4 256 320 skip h16 p16f"; do
    MVSTER_LIB_PATH=$lib timeout 90 python tests/tc_conv_check.py $c 2>&1 | tail -1 | python -c "import json,sys; r=json.loads(sys.stdin.read()); print(' '.join(r['case']).ljust(48), 'rel %.1e' % r['rel'], 'us %.1f' % r['us_tc'])"
  done
  for i in 1 2; do
    MVSTER_LIB_PATH=$lib timeout 300 python bench.py --quick --no-cpu-baseline --skip-e2e --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); s=j['step_stats']; print('bench ms/step %.4f median %.4f min %.4f' % (j['ms_per_step'], s['median_ms'], s['min_ms']))"
  done
done
