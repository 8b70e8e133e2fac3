"""Fused DDIM step: block-shape sweep (DV_DDIM_SHAPE) at the bench shapes."""
import os, sys, json
sys.path.insert(0, '.')
import torch
from diffuvolume_b200 import ops
from diffuvolume_b200.pipeline import DdimSchedule
dev = torch.device('cuda'); B, H, W, D = 8, 540, 960, 48; h, w = H // 4, W // 4
g = torch.Generator(device=dev); g.manual_seed(0)
rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
sch = DdimSchedule(); disp = torch.rand(B, H, W, device=dev) * 191; vote = (torch.rand(B, H, W, device=dev) > 0.5).float()
shift = rn(B, D) * 0.1
def timeit(f):
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[10]
for dt in (torch.float64, torch.float32):
    xt = rn(B, D, h, w, dt=dt); sn = rn(B, D, h, w, dt=dt); rz = torch.rand(B, D, h, w, device=dev, dtype=torch.float64)
    san, c, sg = sch.update_coefficients(999, 799)
    for shape in (0, 1, 2, 3):
        os.environ["DV_DDIM_SHAPE"] = str(shape)
        mask = torch.zeros(B, h, w, device=dev)
        f = lambda: ops.ddim_step(disp=disp, xt=xt, shift=shift, scale=1.0, sqrt_recip=sch.sqrt_recip(999), sqrt_recipm1=sch.sqrt_recipm1(999),
                                  last_step=False, vote=vote, mask=mask, sqrt_alpha_next=san, c=c, sigma=sg, step_noise=sn, renoise=rz,
                                  shift_next=shift, want_n_next=True)
        print(json.dumps({"dtype": str(dt), "shape": shape, "ms": round(timeit(f), 4)}), flush=True)
