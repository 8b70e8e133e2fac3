"""The N>1 path on CPU: world_size-2 gloo — shard the pairs, accumulate per-image metrics per rank, one
all_reduce(SUM); the result must equal the single-process sweep (mirrors AverageMeterDict.mean +
EPE/D1/Thres of SceneFlow/utils/metrics.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth


def _data(n=10, H=12, W=20):
    gt = synth.uniform((n, H, W), 1, dtype=np.float32) * np.float32(100)
    est = gt + synth.normal((n, H, W), 2) * np.float32(2.5)
    mask = (gt < 90) & (gt > 0)
    mask[3] = False           # an image the reference skips (mask coverage < 10 %)
    return torch.from_numpy(est), torch.from_numpy(gt), torch.from_numpy(mask)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from diffuvolume_b200.distributed import MetricSums, init_from_env, shard_range
    r, w, _ = init_from_env("gloo")
    est, gt, mask = _data()
    lo, hi = shard_range(est.shape[0], r, w)
    ms = MetricSums()
    for i in range(lo, hi):                         # batch size 1: the reference's evaluation setting
        ms.update(est[i:i + 1], gt[i:i + 1], mask[i:i + 1])
    out = ms.reduce()
    if r == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_metric_reduction_matches_single_process():
    from diffuvolume_b200.distributed import MetricSums
    est, gt, mask = _data()
    single = MetricSums()
    for i in range(est.shape[0]):
        single.update(est[i:i + 1], gt[i:i + 1], mask[i:i + 1])
    want = single.reduce()
    assert want["n_images"] == 9 and want["n_skipped"] == 1 and want["n_batches"] == 10
    # reference semantics for one image, spelled out
    e = (gt[0][mask[0]] - est[0][mask[0]]).abs()
    from diffuvolume_b200.distributed import per_image_metrics
    m0 = per_image_metrics(est[0], gt[0], mask[0])
    assert abs(m0["EPE"] - float(e.mean())) < 1e-6
    assert abs(m0["D1"] - float(((e > 3) & (e / gt[0][mask[0]].abs() > 0.05)).float().mean())) < 1e-6

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k, v in want.items():
        assert abs(got[k] - v) < 1e-9, (k, got[k], v)


def test_metric_sums_follow_the_reference_batch_rule():
    """compute_metric_for_each_image (SceneFlow/utils/metrics.py:21-41): mean over the non-skipped images of a batch, 0 for
    an all-skipped batch; AverageMeterDict.mean (utils/experiment.py:126-151): mean of those per-batch values."""
    from diffuvolume_b200.distributed import MetricSums, per_image_metrics
    est, gt, mask = _data()
    ms = MetricSums()
    ms.update(est[2:4], gt[2:4], mask[2:4])          # image 3 is skipped -> the batch value is image 2's
    ms.update(est[3:4], gt[3:4], mask[3:4])          # all skipped -> contributes 0, still counts as a batch
    ms.update(est[4:7], gt[4:7], mask[4:7])
    out = ms.reduce()
    m = [per_image_metrics(est[i], gt[i], mask[i]) for i in range(10)]
    want = (m[2]["EPE"] + 0.0 + (m[4]["EPE"] + m[5]["EPE"] + m[6]["EPE"]) / 3) / 3
    assert abs(out["EPE"] - want) < 1e-12 and out["n_batches"] == 3 and out["n_images"] == 4 and out["n_skipped"] == 2
