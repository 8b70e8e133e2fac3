#!/bin/bash
# ncu --set full of one warm hot-path step (second call of scripts/profile_step.py): every kernel of libdv_b200.so.
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'gwc_volume|concat_volume|softmax_regress|ddim_step|xstart|ensemble|volume_filter' -s 19 -c 19 \
  -o gpurun_out/full_step -f python scripts/profile_step.py 8 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log; ls -la gpurun_out/
