#!/bin/bash
# diagnose the step-time outliers of call 7: same box, graph mode, PDL attribute on / off / compiled out, static outputs, GC on
set -u
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python bench.py --quick --no-cpu-baseline --skip-e2e > gpurun_out/c8_$name.json 2>> gpurun_out/c8.err
  python -c "
import json; j=json.load(open('gpurun_out/c8_$name.json')); s=j['step_stats']; print('$name', round(j['ms_per_step'],3), 'min', round(s['min_ms'],3), 'med', round(s['median_ms'],3), 'max', round(s['max_ms'],3), s['steps_ms'])"
}
run nopdl_lib MVSTER_LIB_PATH=$PWD/mvster_b200/lib_nopdl/libmvster_b200.so
run pdl1 MVSTER_TC3_PDL=1
run pdl0 MVSTER_TC3_PDL=0
run nopdl_lib2 MVSTER_LIB_PATH=$PWD/mvster_b200/lib_nopdl/libmvster_b200.so
run pdl1_static MVSTER_TC3_PDL=1 MVSTER_GRAPH_STATIC_OUTPUTS=1
run pdl1_nooverlap MVSTER_TC3_PDL=1 MVSTER_OVERLAP=0
run nopdl_nooverlap MVSTER_LIB_PATH=$PWD/mvster_b200/lib_nopdl/libmvster_b200.so MVSTER_OVERLAP=0
tail -3 gpurun_out/c8.err
