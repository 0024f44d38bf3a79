#!/bin/bash
# round 2, call B4: first GPU contact of the packed fp16-pair operands (the fp32-faithful default), network-level equality, A/B
set -u
mkdir -p gpurun_out
while read -r c; do
  [ -z "$c" ] && continue
  timeout 60 python tests/tc_conv_check.py $c 2>&1 | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    r=json.loads(t); print(' '.join(r['case']).ljust(52), 'rel %.2e' % r['rel'], 'us %.1f' % r['us_tc'])
except Exception as e:
    print('NO RESULT:', t[-400:])
"
done <<'CASES'
v3 16 16 3 3 1 1 8 24 40 skip h16 p16f
v3 8 16 1 3 2 1 4 64 80 h16 p16f
v3 32 64 1 3 2 2 2 32 32 h16 p16
v3 64 64 3 3 1 1 4 64 80 skip h16 p16
v3 16 16 3 3 1 1 4 256 320 h16 p16
v3 16 16 3 3 1 1 4 256 320 h16
v3 32 32 3 3 1 1 4 128 160 skip h16 p16
v3 32 32 3 3 1 1 4 128 160 skip h16
d3 16 8 1 2 24 40 skip h16 p16f
d3 64 32 1 4 8 10 skip h16 p16
CASES
timeout 600 python -m pytest tests/test_gpu_zzzzz_bf16.py -q -x -k "fp16_pair" 2>&1 | tail -12
for p in 1 0; do
  for i in 1 2; do
    MVSTER_REG_PACKED=$p timeout 300 python bench.py --quick --no-cpu-baseline --skip-e2e --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); s=j['step_stats']; print('MVSTER_REG_PACKED=$p bench ms/step %.4f median %.4f min %.4f' % (j['ms_per_step'], s['median_ms'], s['min_ms']))"
  done
done
