#!/bin/bash
# multi-GPU call (gpurun --gpus N): view-sharded forward == unsharded, nn.DataParallel replicas, bench.py --gpus N with the view_sharded / cfg5 legs
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/check_view_shard.py > gpurun_out/mg${N}_check_view_shard.txt 2>&1; echo "check_view_shard rc=$?"; tail -4 gpurun_out/mg${N}_check_view_shard.txt
timeout 300 python -m pytest tests/test_gpu_x_dataparallel.py -m gpu -q > gpurun_out/mg${N}_dataparallel.log 2>&1; echo "dataparallel rc=$?"; tail -4 gpurun_out/mg${N}_dataparallel.log
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/mg${N}_bench.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/mg${N}_bench.json').read().strip().splitlines()[-1])
    print('value', j['value'], 'ms', j['ms_per_step'], j['step_stats']); print(json.dumps(j.get('view_sharded'))); print(json.dumps(j.get('cfg5')))
except Exception as e: print('no bench line', e)
PY
[ "$N" -gt 2 ] && exit 0
MVSTER_SHARD_GRAPH=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/mg${N}_bench_nograph.json 2> gpurun_out/mg${N}_bench_nograph.err; echo "bench (eager shard) rc=$?"
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/mg${N}_bench_nograph.json').read().strip().splitlines()[-1])
    v=j.get('view_sharded'); print('eager-launch sharded: ms', v['ms_per_step'], 'unsharded', v['unsharded_ms_per_step'])
except Exception as e: print('no bench line', e)
PY
