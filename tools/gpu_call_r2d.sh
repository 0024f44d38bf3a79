#!/bin/bash
# round 2, call B3: first GPU contact of the packed-operand convolution (mvster_conv_tc3_pb16), then the network-level checks and the A/B
set -u
mkdir -p gpurun_out
while read -r c; do
  [ -z "$c" ] && continue
  timeout 60 python tests/tc_conv_check.py $c 2>&1 | tail -1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    r=json.loads(t); print(' '.join(r['case']).ljust(52), 'rel %.2e' % r['rel'], 'flip', r.get('flip_frac'), 'us %.1f' % r['us_tc'])
except Exception as e:
    print('NO RESULT:', t[-300:])
"
done <<'CASES'
v3 16 16 3 3 1 1 8 24 40 skip b16 p16f
v3 8 16 1 3 2 1 4 64 80 b16 p16f
v3 32 64 1 3 2 2 2 32 32 b16 p16f
v3 64 64 3 3 1 1 4 64 80 b16 p16f
v3 16 16 3 3 1 1 4 256 320 b16 p16
v3 16 16 3 3 1 1 4 256 320 b16
d3 16 8 1 2 24 40 skip b16 p16f
d3 64 32 1 4 8 10 skip b16 p16
CASES
timeout 600 python -m pytest tests/test_gpu_zzzzz_bf16.py -q -x -k "packed or reg2d or forward" 2>&1 | tail -12
timeout 200 python tools/bf16_ab.py 1152 1600 5 8 2>&1 | tail -2
timeout 200 python tools/bf16_ab.py 512 640 5 20 2>&1 | tail -2
