#!/bin/bash
# ncu --set full of the CUDA-core kernels of one forward (FPN glue, stem, CUDA-core convolutions, head, hypotheses) - the launches
# profiles/r01_launches_final.md lists outside conv_tc3 / et_fuse - summarised with tools/ncu_summary.py.  On the B200 box:
#     bash tools/profile_glue.sh            # writes gpurun_out/glue.ncu-rep and gpurun_out/glue_ncu.md
set -eu
mkdir -p gpurun_out
MVSTER_CUDA_GRAPH=0 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'fpn_|conv_first|conv_px2|head_kernel|hypo_|upsample_|pose_kernel' -c 24 -f -o gpurun_out/glue \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range --skip-e2e > gpurun_out/glue_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/glue.ncu-rep > gpurun_out/glue_ncu.md
head -40 gpurun_out/glue_ncu.md
# the same launches with every prepared variant switched on (gather 3, four-pixel merge / stem / conv0): writes gpurun_out/glue_v2*
MVSTER_FPN_GATHER=3 MVSTER_FPN_MERGE=3 MVSTER_CONV_FIRST=2 MVSTER_CONV0_PX4=1 MVSTER_CUDA_GRAPH=0 \
    ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'fpn_|conv_first|conv0_px4|conv_px2|head_kernel' -c 16 -f -o gpurun_out/glue_v2 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range --skip-e2e > gpurun_out/glue_v2_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/glue_v2.ncu-rep > gpurun_out/glue_v2_ncu.md
head -40 gpurun_out/glue_v2_ncu.md
