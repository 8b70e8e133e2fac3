"""gwc_volume_bwd / corr_volume_2sided_bwd alone (CUDA events, L2 flushed between launches); kernel variant through
DV_GWC_BWD_RING (0 = quad kernel with direct gradient loads, 1-4 = cp.async ring shapes)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from diffuvolume_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


B = 8
fl, fr = torch.randn(B, 320, 135, 240, device=dev), torch.randn(B, 320, 135, 240, device=dev)
gv = torch.randn(B, 40, 48, 135, 240, device=dev)
ms = timeit(lambda: ops.gwc_volume_bwd(gv, fl, fr, 40))
nbytes = B * (40 * 48 + 4 * 320) * 135 * 240 * 4
print(f"variant={os.environ.get('DV_GWC_BWD_RING', 'default')} gwc_bwd B=8 acv: {ms:.4f} ms {nbytes / ms / 1e6:.0f} GB/s")
del fl, fr, gv
Bc = 4
fl, fr = torch.randn(Bc, 32, 384, 1248, device=dev), torch.randn(Bc, 32, 384, 1248, device=dev)
gv = torch.randn(Bc, 1, 49, 384, 1248, device=dev)
ms = timeit(lambda: ops.gwc_volume_bwd(gv, fl, fr, 1, two_sided_maxdisp=24))
nbytes = Bc * (49 + 4 * 32) * 384 * 1248 * 4
print(f"variant={os.environ.get('DV_GWC_BWD_RING', 'default')} corr2_bwd B=4 pcw: {ms:.4f} ms {nbytes / ms / 1e6:.0f} GB/s")
