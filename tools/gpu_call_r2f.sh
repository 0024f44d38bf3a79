#!/bin/bash
# round 2, call B5: A/B of streamed weights for one-group-per-CTA launches; ncu captures of the 16 -> 16 3x3x3 layer in its three
# forms (fp32 activations + converters | packed fp16 pair | packed bf16)
set -u
mkdir -p gpurun_out
for p in 1 0; do
  for i in 1 2; do
    MVSTER_TC3_STREAM_SMALL=$p timeout 300 python bench.py --quick --no-cpu-baseline --skip-e2e --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); s=j['step_stats']; print('MVSTER_TC3_STREAM_SMALL=$p bench ms/step %.4f median %.4f min %.4f' % (j['ms_per_step'], s['median_ms'], s['min_ms']))"
  done
done
i=0
for c in "v3 16 16 3 3 1 1 4 256 320 h16" "v3 16 16 3 3 1 1 4 256 320 h16 p16" "v3 16 16 3 3 1 1 4 256 320 b16 p16"; do
  i=$((i+1))
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:conv_tc3 -c 1 -f -o gpurun_out/g_conv2_$i python tests/tc_conv_check.py $c > gpurun_out/g_ncu_$i.log 2>&1; echo "ncu $i rc=$?"; tail -1 gpurun_out/g_ncu_$i.log | cut -c1-200
done
