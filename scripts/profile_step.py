"""One warm-up step + one profiled step of the bench workload (for ncu)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import bench
from diffuvolume_b200.pipeline import AcvHotPath
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device('cuda', 0)
inp = bench.make_inputs(B, dev, seed=1234)
path = AcvHotPath()
for _ in range(2):
    path(**inp)
torch.cuda.synchronize()
