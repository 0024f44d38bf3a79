#!/bin/bash
# round-2 wrap-up on one B200: the whole GPU suite, smoke, the bench line (with the bf16-storage leg), the reference arm
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f2_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/f2_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 500 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/f2_bench.err
python - <<'PY'
import json
j = json.load(open('gpurun_out/f2_bench.json'))
s = j['step_stats']; s.pop('steps_ms', None)
print(round(j['ms_per_step'], 4), round(j['value'], 1), 'e2e', round(j['e2e']['value'], 1), 'serial', round(j['e2e']['serial_value'], 1), s, 'parity', j['parity']['ok'], j['parity']['bad_frac_per_stage'], j['parity']['stage1_attn_abs'])
print(j['roofline']['kernel'], round(j['roofline']['frac'], 3), [round(p['us'], 1) for p in j['roofline']['per_stage']], 'all', round(j['roofline']['all_stages']['frac'], 3))
print('tensor', j['roofline_tensor']['kernel'], round(j['roofline_tensor']['us'], 1), round(j['roofline_tensor']['frac'], 4))
print('eager', j['gpu_eager_baseline']['tf32_on']['value'], j['gpu_eager_baseline']['tf32_off']['value'])
print('extra sizes', j['extra']['sizes'], 'no graph', j['extra']['no_cuda_graph'])
b = j['extra']['bf16_storage']
print('bf16', {k: v for k, v in b.items() if k not in ('roofline', 'parity')})
if 'roofline' in b:
    print('bf16 roofline', round(b['roofline']['frac'], 3), round(b['roofline']['all_stages']['frac'], 3), [round(p['us'], 1) for p in b['roofline']['per_stage']])
    print('bf16 parity', b.get('parity', {}).get('ok'), b.get('parity', {}).get('bad_frac_per_stage'), b.get('parity', {}).get('mean_attn_distance_to_bf16_oracle'))
print('launches', j['gpu_launches'], 'clocks', j['clocks'])
PY
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f2_bench_reference.json 2>> gpurun_out/f2_bench.err; tail -c 300 gpurun_out/f2_bench_reference.json; echo
