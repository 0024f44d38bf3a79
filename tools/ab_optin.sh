#!/bin/bash
# A/B of the opt-in kernel variants that were prepared without GPU time (functionally checked on the CPU emulation or by SASS
# inspection only): correctness on the GPU first, then the bench step with each switch.  Usage (on the B200 box):
#     bash tools/ab_optin.sh 2>&1 | tee gpurun_out/ab_optin.txt
set -u
echo "== correctness with the switches on"
timeout 400 python -m pytest tests/test_gpu_zzy_variants.py -m gpu -q 2>&1 | tail -3   # every variant against its default kernel
MVSTER_TC3_MERGE=1 timeout 400 python -m pytest tests/test_gpu_tc_conv.py -m gpu -q -x -k "v3" 2>&1 | tail -2
MVSTER_TC3_MERGE=1 MVSTER_FPN_GATHER=2 MVSTER_FPN_MERGE=2 MVSTER_CONV_FIRST=2 MVSTER_CONV0_PX4=1 timeout 300 python -m pytest tests/test_gpu_y_fpn.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
MVSTER_FPN_GATHER=3 MVSTER_FPN_MERGE=3 timeout 300 python -m pytest tests/test_gpu_y_fpn.py -m gpu -q -x 2>&1 | tail -2
echo "== kernel-level A/B (us per launch, GB/s on the algorithmic bytes, deviation from the default kernel)"
timeout 600 python tools/glue_ab.py > gpurun_out/glue_ab.json 2>gpurun_out/glue_ab.err; python -c "
import json
for r in json.load(open('gpurun_out/glue_ab.json'))['rows']: print('%-42s %-22s %8.2f us %8.1f GB/s  dev %.1e' % (r['kernel'], r['switch'], r['us'], r['GB/s'], r['max_dev_vs_default']))"
echo "== bench step (ms) per switch"
for sw in "" "MVSTER_TC3_MERGE=1" "MVSTER_FPN_GATHER=2" "MVSTER_FPN_GATHER=3" "MVSTER_FPN_MERGE=2" "MVSTER_FPN_MERGE=3" "MVSTER_CONV_FIRST=2" "MVSTER_CONV0_PX4=1" "MVSTER_FPN_GATHER=3 MVSTER_FPN_MERGE=3 MVSTER_CONV_FIRST=2 MVSTER_CONV0_PX4=1" "MVSTER_TC3_MERGE=1 MVSTER_FPN_GATHER=3 MVSTER_FPN_MERGE=3 MVSTER_CONV_FIRST=2 MVSTER_CONV0_PX4=1"; do
  out=$(env $sw timeout 200 python bench.py --no-cpu-baseline --quick 2>/dev/null)
  echo "$out" | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('%-70s %.4f ms  %.1f maps/s  e2e %.1f' % ('[$sw]', j['ms_per_step'], j['value'], j['e2e']['value']))"
done
echo "== training path: backward of the warp + ET kernel (parity first, then forward/backward time and held memory vs the PyTorch ops)"
timeout 600 python -m pytest tests/test_gpu_zzz_et_backward.py tests/test_gpu_zz_fusion.py -m gpu -q 2>&1 | tail -3
timeout 600 python tools/et_bwd_bench.py > gpurun_out/et_bwd_bench.json 2>gpurun_out/et_bwd_bench.err; tail -40 gpurun_out/et_bwd_bench.json
