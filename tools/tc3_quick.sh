#!/bin/bash
# Quick correctness + timing sweep of conv_tc3 / deconv_tc3 (subset of tests/test_gpu_tc_conv.py), both arithmetics.
fail=0
while read -r c; do
  for h in "" "h16"; do
    out=$(python tests/tc_conv_check.py $c $h 2>&1 | tail -1)
    echo "$out" | python -c "
import json,sys
try:
    r=json.loads(sys.stdin.read()); ok = r['finite'] and r['rel'] < 1e-5
    print(('ok  ' if ok else 'FAIL'), ' '.join(r['case']).ljust(46), 'rel %.2e' % r['rel'], 'us %.1f' % r['us_tc'])
    sys.exit(0 if ok else 1)
except Exception as e:
    print('FAIL (no result)', e); sys.exit(1)
" || fail=1
  done
done <<'CASES'
v3 16 16 3 3 1 1 8 24 40
v3 64 64 3 3 1 1 4 64 80 skip
v3 4 8 1 3 1 1 4 32 48
v3 16 32 1 3 2 1 4 30 44
v3 32 64 1 5 2 1 1 64 80
v3 16 16 3 3 1 1 4 256 320
v3 16 16 1 3 1 5 1 512 640
v3 64 16 1 3 1 5 1 256 320 norelu
v3 32 32 3 3 1 1 4 128 160 skip
d3 64 32 1 4 8 10 skip
d3 32 16 1 4 128 160 skip
CASES
echo "tc3_quick: fail=$fail"
