"""softmax_regress variants: timing + agreement with the register-resident kernel (run on the GPU box)."""
import os, sys, json
sys.path.insert(0, '.')
import torch
from diffuvolume_b200 import ops
dev = torch.device('cuda')
B, H, W = 8, 540, 960
g = torch.Generator(device=dev); g.manual_seed(0)
cost = torch.randn(B, 192, H, W, generator=g, device=dev) * 4
used = torch.rand(B, H, W, device=dev) * 191
def run(full=True, **env):
    for k, v in env.items(): os.environ[k] = str(v)
    ens = torch.zeros(B, H, W, device=dev)
    f = (lambda: ops.softmax_regress(cost, used=used, want_unc=True, vote_thresholds=(1.0, 3.0), ens_acc=ens, ens_coef=0.2)) if full else (lambda: ops.softmax_regress(cost))
    for _ in range(3): r = f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[5], r
gb = B * (192 * H * W * 4 + 5 * H * W * 4) / 1e9
ms0, r0 = run(DV_SR_TMA=0)
print(json.dumps({"variant": "regs", "ms": round(ms0, 4), "GBs": round(gb / ms0 * 1e3, 1)}))
for name, env in [("shape0_16sl_2st_2cta", dict(DV_SR_TMA=1, DV_SR_SHAPE=0)), ("shape1_8sl_1st_3cta", dict(DV_SR_TMA=1, DV_SR_SHAPE=1)), ("shape2_8sl_2st_2cta", dict(DV_SR_TMA=1, DV_SR_SHAPE=2)), ("shape3_16sl_1st_3cta", dict(DV_SR_TMA=1, DV_SR_SHAPE=3))]:
    ms, r = run(**env)
    err = {k: float((r[k] - r0[k]).abs().max()) for k in ("disp", "unc", "vote")}
    print(json.dumps({"variant": name, "ms": round(ms, 4), "GBs": round(gb / ms * 1e3, 1), "maxdiff_vs_regs": err}))
    ms, r = run(full=False, **env)
    print(json.dumps({"variant": name + "_disp_only", "ms": round(ms, 4), "GBs": round(B * 192 * H * W * 4 / 1e9 / ms * 1e3, 1)}))
