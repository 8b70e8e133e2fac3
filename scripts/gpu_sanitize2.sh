#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in round 1 (packed geo pyramid, window lookup, k-chunked gwc,
# negative-shift split, quad backward, patch chain, context upsample, 3xTF32 all-pairs, bf16 stores) + racecheck on the
# new bulk-copy pipelines.  Usage: bash scripts/gpu_sanitize2.sh
mkdir -p gpurun_out
SUB='geo_pack or kitti15_geo or corr1d or corr_volume_2sided or bf16 or gwc_volume_golden'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "$SUB" > gpurun_out/sanitize2_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize2_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_f4.py tests/test_gpu_backward.py -x -q -m gpu > gpurun_out/sanitize2_memcheck_f4_bwd.log 2>&1
echo "memcheck f4+bwd rc=$?" >> gpurun_out/sanitize2_memcheck_f4_bwd.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -x -q -k "corr_volume_2sided_golden or gwc_volume_bwd or gwc_bwd or geo_pack" > gpurun_out/sanitize2_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitize2_racecheck.log
tail -n 4 gpurun_out/sanitize2_memcheck.log gpurun_out/sanitize2_memcheck_f4_bwd.log gpurun_out/sanitize2_racecheck.log
