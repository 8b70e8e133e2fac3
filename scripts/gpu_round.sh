#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list and a full capture of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 8 > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gwc_volume_kernel|concat_stream_kernel|softmax_regress_tma_kernel|ddim_step_kernel|att_softmax|filter_factor|xstart|ensemble' -s 21 -c 10 -o gpurun_out/full_step -f python scripts/profile_step.py 8 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/smoke.log
