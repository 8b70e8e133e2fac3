"""ncu target: PCWNet refinement kernels (warp, +-24 correlation) at B=8, 384x1248."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from diffuvolume_b200 import ops
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(3)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)
ru = lambda *s: torch.rand(*s, generator=g, device=dev)
B = 8
f1, f2 = rn(B, 32, 384, 1248), rn(B, 32, 384, 1248)
dsp = (torch.linspace(2, 90, 1248, device=dev).view(1, 1, 1, -1).expand(B, 1, 384, 1248) + ru(B, 1, 384, 1248)).contiguous()
for _ in range(2):
    w = ops.warp(f2, dsp)
    ops.corr_volume_2sided(f1, w, 24, 1)
torch.cuda.synchronize()
