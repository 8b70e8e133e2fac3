#!/bin/bash
# Round 2: compute-sanitizer over the kernels added this round.
#   memcheck: the whole backward / f4 / fuzz / sampler suites (ring gwc backward, negative-shift row kernels, warp backward,
#             fused backward, stream patch chain, softmax uncertainty vote, tcgen05 all-pairs, refinement-input assembly);
#   racecheck: the kernels that synchronise through shared memory without mbarriers — the cp.async ring of the gwc backward
#             (__syncthreads per stage), the patch-chain stream (__syncwarp only), the shuffle-based negative-shift rows.
# Usage: bash scripts/gpu_sanitize3.sh
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_backward.py tests/test_f4.py tests/test_gpu_fuzz.py tests/test_gpu_sampler.py tests/test_gpu_sampler_pcw.py tests/test_gpu_tier3.py -x -q -m gpu > gpurun_out/sanitize3_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitize3_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "corr1d or allpairs or warp or refine or corr_volume_2sided or geo" > gpurun_out/sanitize3_memcheck_parity.log 2>&1
echo "memcheck parity rc=$?" >> gpurun_out/sanitize3_memcheck_parity.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_backward.py tests/test_f4.py -x -q -m gpu -k "gwc_volume_backward or corr_volume_2sided_backward or patch_chain" > gpurun_out/sanitize3_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitize3_racecheck.log
tail -n 5 gpurun_out/sanitize3_memcheck.log gpurun_out/sanitize3_memcheck_parity.log gpurun_out/sanitize3_racecheck.log
