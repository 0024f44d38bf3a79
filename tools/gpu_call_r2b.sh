#!/bin/bash
# round 2, call B1: first GPU contact of the bf16-storage path (tests), a regression subset of the fp32 tensor-core layers after the
# plan-header refactor, smoke, and the default bench line with the new bf16 leg
set -u
mkdir -p gpurun_out
for c in "v3 16 16 3 3 1 1 8 24 40 h16" "v3 8 16 1 5 2 2 1 64 96" "d3 64 32 1 4 8 10 skip h16" "v3 16 16 3 3 1 1 4 256 320 h16" "v3 64 64 3 3 1 1 4 64 80 skip h16"; do
  timeout 90 python tests/tc_conv_check.py $c 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_zzzzz_bf16.py -q -x > gpurun_out/b1_bf16_tests.log 2>&1; echo "bf16 pytest rc=$?"; tail -25 gpurun_out/b1_bf16_tests.log
grep b16 gpurun_out/tc_conv_report.jsonl | tail -12
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/b1_bench.json 2> gpurun_out/b1_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/b1_bench.err
python - <<'PY'
import json
j = json.load(open('gpurun_out/b1_bench.json'))
print(round(j['ms_per_step'], 4), round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'parity', j['parity'] and j['parity']['ok'])
print('roofline', round(j['roofline']['frac'], 3), 'tensor', round(j['roofline_tensor']['frac'], 4))
print('bf16 leg', json.dumps(j['extra'].get('bf16_storage'))[:3000])
PY
