#!/bin/bash
# ncu evidence for one warm hot-path step (second call of scripts/profile_step.py): launch list + --set full of every
# kernel of libdv_b200.so.  Usage: bash scripts/gpu_ncu_full.sh <tag>
TAG=${1:-r01b}
mkdir -p gpurun_out
K='gwc_volume|concat|softmax_regress|ddim_step|xstart|ensemble|volume_filter|att_softmax|filter_factor|upsample_regress'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python scripts/profile_step.py 8 > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 21 -c 21 \
  -o gpurun_out/full_step_$TAG -f python scripts/profile_step.py 8 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log; ls -la gpurun_out/ | grep $TAG
