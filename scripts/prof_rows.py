"""ncu target: one warm + one profiled launch of the secondary kernels at their BASELINE sizes."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
from diffuvolume_b200 import ops
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(3)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)
ru = lambda *s: torch.rand(*s, generator=g, device=dev)
B = 8
fl, fr = rn(B, 320, 135, 240), rn(B, 320, 135, 240)
gv = rn(B, 40, 48, 135, 240)
wp, wl = rn(40, 9), rn(40, 9)
f1, f2 = rn(4, 32, 384, 1248), rn(4, 32, 384, 1248)
dsp = (torch.linspace(2, 90, 1248, device=dev).view(1, 1, 1, -1).expand(4, 1, 384, 1248) + ru(4, 1, 384, 1248)).contiguous()
a1, a2 = rn(B, 96, 96, 312), rn(B, 96, 96, 312)
for _ in range(2):
    ops.gwc_volume_bwd(gv, fl, fr, 40)
    ops.acv_patch_volume(gv, wp, wl[:8], wl[8:24], wl[24:])
    ops.corr_volume_2sided(f1, f2, 24, 1)
    ops.warp(f2, dsp)
    ops.corr1d_allpairs(a1, a2)
torch.cuda.synchronize()
