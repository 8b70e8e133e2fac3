#!/bin/bash
# 8-GPU visit: H2D ceiling of the box (1/2/4/8 ranks, with and without NUMA-local pinning), then the bench at N = 8.
# Usage (8x charged): gpurun --gpus 8 --timeout 900 -- bash scripts/gpu_scale8.sh
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" > gpurun_out/lscpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
: > gpurun_out/h2d_ceiling.jsonl
timeout 120 python scripts/h2d_ceiling.py >> gpurun_out/h2d_ceiling.jsonl 2>> gpurun_out/h2d.err
for n in 2 4 8; do
  timeout 180 $TR --nproc-per-node $n --master-port $((29600+n)) scripts/h2d_ceiling.py >> gpurun_out/h2d_ceiling.jsonl 2>> gpurun_out/h2d.err
  timeout 180 $TR --nproc-per-node $n --master-port $((29700+n)) scripts/h2d_ceiling.py --numa-local >> gpurun_out/h2d_ceiling.jsonl 2>> gpurun_out/h2d.err
done
timeout 180 $TR --nproc-per-node 8 --master-port 29811 scripts/h2d_ceiling.py --numa-local --streams 2 >> gpurun_out/h2d_ceiling.jsonl 2>> gpurun_out/h2d.err
for n in 2 4 8; do
  timeout 400 $TR --nproc-per-node $n --master-port $((29900+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-legs --no-graph --no-sustained \
    > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
python - <<'PY'
import json
for l in open('gpurun_out/h2d_ceiling.jsonl'):
    try: r = json.loads(l)
    except Exception: continue
    print(r['n_gpus'], 'numa_local' if r['numa_local'] else 'default', 'streams', r['streams'], 'aggregate', r['aggregate_gbs'], [x['gbs'] for x in r['ranks']])
for n in (2, 4, 8):
    try:
        r = json.load(open(f'gpurun_out/bench_n{n}.json'))
        print(n, r['value'], r['e2e']['value'], r['e2e'].get('h2d_GBs_per_gpu'), r.get('e2e_logits_boundary', {}).get('value'))
    except Exception as e:
        print(n, 'failed', e)
PY
