import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
from diffuvolume_b200 import ops, kitti15
dev = torch.device('cuda',0)
g = torch.Generator(device=dev); g.manual_seed(1)
rn = lambda *s: torch.randn(*s, generator=g, device=dev)
B,h,w,D = 1,96,312,48
f1,f2,geo = rn(B,96,h,w), rn(B,96,h,w), rn(B,8,D,h,w)
fn = kitti15.Combined_Geo_Encoding_Volume(f1,f2,geo,num_levels=2,radius=4)
disp = torch.rand(B,1,h,w,device=dev)*40
coords = torch.arange(w,device=dev,dtype=torch.float32).view(1,1,1,w).expand(B,1,h,w).contiguous()
noisy = torch.rand(B,D,h,w,device=dev)
for name, call in (("lookup(noisy)", lambda: fn(disp,coords,noisy)), ("lookup(plain)", lambda: fn(disp,coords)),
                   ("gwc_volume B=1 igev", lambda: ops.gwc_volume(f1,f2,D,8)),
                   ("softmax_regress D=48", lambda: ops.softmax_regress(geo[:,0]))):
    for _ in range(20): call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 500
    for _ in range(n): call()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"{name:24s} host issue {1e6*(t1-t0)/n:6.1f} us/call   wall {1e6*(t2-t0)/n:6.1f} us/call")
