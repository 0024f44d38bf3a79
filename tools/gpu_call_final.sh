#!/bin/bash
# round-2 wrap-up on one B200: the whole GPU suite, the bench line, the launch list, the ncu capture of the default warp/ET kernel
# (-> profiles/et_traffic.json), the filter kernel's time
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/f_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench.err
python -c "
import json; j=json.load(open('gpurun_out/f_bench.json')); s=j['step_stats']; s.pop('steps_ms',None)
print(round(j['ms_per_step'],4), round(j['value'],1), 'e2e', round(j['e2e']['value'],1), s, 'parity', j['parity']['ok'], j['parity']['bad_frac_per_stage'])
print(j['roofline']['kernel'], round(j['roofline']['frac'],3), [round(p['us'],1) for p in j['roofline']['per_stage']], 'all', round(j['roofline']['all_stages']['frac'],3))
print('eager', j['gpu_eager_baseline']['tf32_on']['value'], j['gpu_eager_baseline']['tf32_off']['value'], 'extra', j['extra'])"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2>> gpurun_out/f_bench.err; tail -c 400 gpurun_out/f_bench_reference.json; echo
timeout 120 python tools/fusion_bench.py > gpurun_out/f_fusion_bench.json 2>> gpurun_out/f_bench.err; cat gpurun_out/f_fusion_bench.json | tail -40
MVSTER_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range --skip-e2e --quick > gpurun_out/f_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"et_fuse_tma" -c 3 -o gpurun_out/f_et_tma python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --quick > gpurun_out/f_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/f_ncu.log
