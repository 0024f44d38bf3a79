#!/bin/bash
# Stage ablation of conv_tc3_kernel (MVSTER_TC3_DEBUG bits: 1 no epilogue traffic, 2 no conversion, 4 no MMAs, 8 no activation
# loads): which pipeline stage bounds a layer.  Results of ablated runs are garbage by design; only us_tc is read.
for c in "v3 16 16 1 3 1 5 1 512 640" "v3 16 16 3 3 1 1 4 256 320" "v3 32 32 3 3 1 1 4 128 160 skip"; do
  for h in "" "h16"; do
    line="$c $h:"
    for d in 0 1 2 4 8 7; do
      us=$(MVSTER_TC3_DEBUG=$d python tests/tc_conv_check.py $c $h 2>/dev/null | tail -1 | python -c "import json,sys; print(round(json.loads(sys.stdin.read())['us_tc'],1))")
      line="$line d$d=$us"
    done
    echo "$line"
  done
done
