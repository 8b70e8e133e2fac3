import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from diffuvolume_b200 import ops
from oracle import dv_oracle as O
g = np.load('tests/golden/sceneflow.npz')
dq = g['trace.disp_q']
got = ops.xstart_from_disp(torch.from_numpy(dq).cuda(), 48, 1.0).cpu().numpy()
want = g['trace.asd']
bad = np.argwhere(np.abs(got - want) > 0)
print("n bad", len(bad))
for b, d, y, x in bad[:40]:
    print(b, d, y, x, "dq", repr(dq[b, y, x]), "got", got[b, d, y, x], "want", want[b, d, y, x])
