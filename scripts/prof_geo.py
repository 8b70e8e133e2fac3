"""ncu target: the IGEV lookups (origin / +noise) at B=8, 96x312."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from diffuvolume_b200 import ops
dev = torch.device('cuda', 0)
B, h, w, D, Cg = 8, 96, 312, 48, 8
g = torch.Generator(device=dev); g.manual_seed(3)
geo = torch.randn(B, Cg, D, h, w, generator=g, device=dev)
corr = torch.randn(B * h * w, 1, 1, w, generator=g, device=dev)
cp = [corr, ops.avgpool_w2(corr)]
pk = ops.geo_pack(geo, 2)
disp = torch.rand(B, 1, h, w, generator=g, device=dev) * 47
coords = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand(B, 1, h, w).contiguous()
noisy = torch.rand(B, D, h, w, generator=g, device=dev)
for _ in range(2):
    ops.geo_lookup_packed(pk, cp, disp, coords, None, 4)
    ops.geo_lookup_packed(pk, cp, disp, coords, noisy, 4)
torch.cuda.synchronize()
