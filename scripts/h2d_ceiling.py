"""Host->device bandwidth ceiling of the box, per rank and aggregate (the e2e legs of bench.py are PCIe-bound).

    python scripts/h2d_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/h2d_ceiling.py [--numa-local]                        # N GPUs copying at the same time

Each rank copies a 1 GiB pinned host buffer to its GPU `--iters` times (cudaMemcpyAsync, one stream; `--streams 2` splits
every copy into two halves on two streams), all ranks started together behind a barrier, timed with CUDA events; rank 0
prints one JSON line with the per-rank GB/s, the aggregate, and the PCI / NUMA placement of every GPU (sysfs), so that a
plateau can be attributed to shared uplinks or to remote-socket host memory.  --numa-local pins the rank's threads to the
CPUs sysfs lists as local to its GPU BEFORE the pinned buffer is allocated (first-touch puts the pages on that node).
"""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffuvolume_b200.distributed import gpu_numa_info, bind_to_gpu_numa_node  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--numa-local", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    info = gpu_numa_info(local)
    bound = bind_to_gpu_numa_node(local) if args.numa_local else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n = args.mib << 20
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.fill_(rank + 1)                      # touch every page (first touch decides the NUMA node)
    dst = torch.empty(n, dtype=torch.uint8, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(args.streams)]
    parts = [(i * n // args.streams, (i + 1) * n // args.streams) for i in range(args.streams)]

    def copy_once():
        for st, (a, b) in zip(streams, parts):
            with torch.cuda.stream(st):
                dst[a:b].copy_(host[a:b], non_blocking=True)

    for _ in range(2):
        copy_once()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    ev0.record(cur)
    for st in streams:
        st.wait_event(ev0)
    for _ in range(args.iters):
        copy_once()
    for st in streams:
        cur.wait_stream(st)
    ev1.record(cur)
    torch.cuda.synchronize()
    gbs = args.iters * n / 1e9 / (ev0.elapsed_time(ev1) / 1e3)
    row = {"rank": rank, "gbs": round(gbs, 2), **info, "bound_cpus": bound}
    rows = [row]
    if dist is not None:
        rows = [None] * world
        dist.all_gather_object(rows, row)
    if rank == 0:
        topo = ""
        try:
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
        except Exception as e:  # noqa: BLE001
            topo = f"unavailable: {e!r}"
        print(json.dumps({"n_gpus": world, "numa_local": args.numa_local, "streams": args.streams, "mib": args.mib,
                          "aggregate_gbs": round(sum(r["gbs"] for r in rows), 1), "ranks": rows,
                          "host_cpus": os.cpu_count(), "topo": topo}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
