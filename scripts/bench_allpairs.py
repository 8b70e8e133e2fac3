import sys, torch, time
sys.path.insert(0, "/root/repo")
from diffuvolume_b200 import ops
B, C, H, W = 8, 96, 96, 312
f1 = torch.randn(B, C, H, W, device="cuda"); f2 = torch.randn(B, C, H, W, device="cuda")
for _ in range(3): ops.corr1d_allpairs(f1, f2, return_pooled=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): ops.corr1d_allpairs(f1, f2, return_pooled=True)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
byt = (2 * B * C * H * W + 1.5 * B * H * W * W) * 4
print(f"allpairs B=8 96x312: {ms:.4f} ms  {byt/1e9/(ms/1e3):.0f} GB/s")
