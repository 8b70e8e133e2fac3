"""Batch-sharded multi-GPU plumbing: one process per GPU, stereo pairs partitioned by index, no
collective on the data path, ONE all_reduce(SUM) of a small float64 vector at the end of a sweep.

The reference only has single-process nn.DataParallel (SceneFlow/main.py:67); its torch.distributed
helpers (SceneFlow/utils/experiment.py:155-190 `reduce_scalar_outputs`, utils/misc.py:20-41) are dead
code.  Metric semantics follow SceneFlow/utils/metrics.py:43-65 (per-image EPE / D1 / Thres, skipped
when the valid mask is almost empty, averaged over images — AverageMeterDict.mean,
utils/experiment.py:126-151).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

METRIC_KEYS = ("EPE", "D1", "Thres1", "Thres2", "Thres3")


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of `n_items` stereo pairs owned by `rank` (sizes differ by at most 1)."""
    assert 0 <= rank < world and n_items >= 0
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def per_image_metrics(d_est: torch.Tensor, d_gt: torch.Tensor, mask: torch.Tensor) -> Optional[Dict[str, float]]:
    """EPE_metric / D1_metric / Thres_metric for ONE image ([H,W] tensors) — SceneFlow/utils/metrics.py:43-65.
    Returns None when the image is skipped (metrics.py:32-33: mask coverage < 10 % of the gt>0 area)."""
    mask = mask.bool()
    denom = (d_gt > 0).float().mean()
    if mask.float().mean() / denom < 0.1:
        return None
    e = (d_gt[mask] - d_est[mask]).abs().double()
    g = d_gt[mask].abs().double()
    return {
        "EPE": float(e.mean()),
        "D1": float(((e > 3) & (e / g > 0.05)).double().mean()),
        "Thres1": float((e > 1.0).double().mean()),
        "Thres2": float((e > 2.0).double().mean()),
        "Thres3": float((e > 3.0).double().mean()),
    }


@dataclass
class MetricSums:
    """Per-rank running sums over the batches of its shard; `reduce()` makes them global.

    Averaging follows the reference exactly: `compute_metric_for_each_image` (SceneFlow/utils/metrics.py:21-41) returns
    the mean over the NON-SKIPPED images of a batch — and 0 for a batch whose images were all skipped — and
    `AverageMeterDict.mean` (utils/experiment.py:126-151) averages those per-batch values over the batches.  So the sums
    kept here are sums of per-batch means and the divisor is the number of batches (an all-skipped batch still counts).
    With batch size 1 (the reference's evaluation setting) this is the per-image mean; sharding the sweep over ranks changes
    the numbers only through which images share a batch."""
    sums: Dict[str, float] = field(default_factory=lambda: {k: 0.0 for k in METRIC_KEYS})
    n_batches: int = 0
    n_images: int = 0
    n_skipped: int = 0

    def update(self, d_est: torch.Tensor, d_gt: torch.Tensor, mask: torch.Tensor) -> None:
        """One batch: d_est, d_gt, mask are [B,H,W]."""
        assert d_est.dim() == 3 and d_est.shape == d_gt.shape == mask.shape
        per_image = []
        for i in range(d_est.shape[0]):
            m = per_image_metrics(d_est[i], d_gt[i], mask[i])
            if m is None:
                self.n_skipped += 1
                continue
            per_image.append(m)
            self.n_images += 1
        self.n_batches += 1
        if per_image:                      # else: the reference adds 0 for this batch
            for k in METRIC_KEYS:
                self.sums[k] += sum(m[k] for m in per_image) / len(per_image)

    def as_tensor(self, device) -> torch.Tensor:
        return torch.tensor([self.sums[k] for k in METRIC_KEYS] + [float(self.n_batches), float(self.n_images),
                                                                    float(self.n_skipped)],
                            dtype=torch.float64, device=device)

    def reduce(self, device=None, group=None) -> Dict[str, float]:
        """Global means.  With an initialised process group: one all_reduce(SUM) of 8 float64 values
        (NCCL over NVLink on GPUs, gloo in the CPU tests); otherwise the local values."""
        dev = device if device is not None else torch.device("cpu")
        v = self.as_tensor(dev)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        v = v.cpu()
        nk = len(METRIC_KEYS)
        n = max(float(v[nk]), 1.0)
        out = {k: float(v[i]) / n for i, k in enumerate(METRIC_KEYS)}
        out["n_batches"] = int(v[nk])
        out["n_images"] = int(v[nk + 1])
        out["n_skipped"] = int(v[nk + 2])
        return out


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group when world > 1."""
    import os
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


# --------------------------------------------------------------------------------------------
# Host placement of a rank: pinned staging buffers should live on the NUMA node its GPU hangs off
# --------------------------------------------------------------------------------------------
def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_info(device_index: int) -> Dict[str, object]:
    """PCI address, NUMA node and local CPU list of a CUDA device as sysfs reports them (empty values when sysfs does not
    expose the device, e.g. in a container without /sys/bus/pci)."""
    import os
    info: Dict[str, object] = {"pci": None, "numa_node": None, "local_cpus": None}
    try:
        props = torch.cuda.get_device_properties(device_index)
        pci = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        info["pci"] = pci
        base = f"/sys/bus/pci/devices/{pci}"
        if os.path.exists(f"{base}/numa_node"):
            info["numa_node"] = int(open(f"{base}/numa_node").read().strip())
        if os.path.exists(f"{base}/local_cpulist"):
            info["local_cpus"] = open(f"{base}/local_cpulist").read().strip()
    except Exception:  # noqa: BLE001 — placement is an optimisation, never a failure
        pass
    return info


def bind_to_gpu_numa_node(device_index: int) -> Optional[str]:
    """Restrict this process to the CPUs local to its GPU (sched_setaffinity), so that the pinned host buffers it allocates
    and first-touches afterwards land on that NUMA node and its H2D copies do not cross the socket interconnect.  Returns
    the CPU list applied, or None when sysfs has no placement for the device or the mask would be empty."""
    import os
    cpus_txt = gpu_numa_info(device_index).get("local_cpus")
    if not cpus_txt or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        allowed = os.sched_getaffinity(0)
        want = set(_parse_cpulist(str(cpus_txt))) & allowed
        if not want:
            return None
        os.sched_setaffinity(0, want)
        return str(cpus_txt)
    except OSError:
        return None
