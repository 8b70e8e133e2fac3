"""Fused trilinear-upsample + softmax-regress (f2) vs the unfused pair (torch upsample + dv softmax_regress)."""
import sys, json
sys.path.insert(0, '.')
import torch, torch.nn.functional as F
from diffuvolume_b200 import ops
dev = torch.device('cuda')
B, H, W, Dq = 8, 540, 960, 48
h, w = H // 4, W // 4
g = torch.Generator(device=dev); g.manual_seed(0)
cq = torch.randn(B, 1, Dq, h, w, generator=g, device=dev) * 4
used = torch.rand(B, H, W, device=dev) * 191
ens = torch.zeros(B, H, W, device=dev)
def timeit(f):
    for _ in range(3): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[5]
full = lambda: ops.upsample_softmax_regress(cq, (192, H, W), used=used, vote_thresholds=(1.0, 3.0), ens_acc=ens, ens_coef=0.2)
print(json.dumps({"fused_full_ms": round(timeit(full), 4), "fused_disp_only_ms": round(timeit(lambda: ops.upsample_softmax_regress(cq, (192, H, W))), 4)}))
up = lambda: F.interpolate(cq, size=(192, H, W), mode="trilinear")
print(json.dumps({"torch_trilinear_ms": round(timeit(up), 4)}))
cost = up()[:, 0]
r0 = ops.softmax_regress(cost, used=used, want_unc=True, vote_thresholds=(1.0, 3.0))
r1 = ops.upsample_softmax_regress(cq, (192, H, W), used=used, want_unc=True, vote_thresholds=(1.0, 3.0))
print(json.dumps({"maxdiff_disp": float((r0["disp"] - r1["disp"]).abs().max()), "maxdiff_unc": float((r0["unc"] - r1["unc"]).abs().max()),
                  "vote_mismatch_frac": float((r0["vote"] != r1["vote"]).float().mean())}))
