#!/bin/bash
# compute-sanitizer over a subset of the GPU parity tests (memcheck: every kernel family; racecheck: the mbarrier / TMA
# pipelines) + an ncu capture of the fused DDIM step.  Usage: bash scripts/gpu_sanitize.sh
mkdir -p gpurun_out
SUB='weighted_producer or softmax_regress_full or gwc_volume_vs_oracle or ddim_trace or upsample_softmax or warp_golden or geo_lookup_golden or corr_volume_2sided_golden'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SUB" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize_memcheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_backward.py tests/test_gpu_sampler_igev.py -x -q > gpurun_out/sanitize_memcheck2.log 2>&1
echo "memcheck2 rc=$?" >> gpurun_out/sanitize_memcheck2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "weighted_producer or softmax_regress_full or gwc_volume_vs_oracle" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitize_racecheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ddim_step' -s 5 -c 2 -o gpurun_out/ddim_r01c -f python scripts/profile_step.py 8 > gpurun_out/ncu_ddim.log 2>&1
tail -4 gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_memcheck2.log gpurun_out/sanitize_racecheck.log; tail -2 gpurun_out/ncu_ddim.log
