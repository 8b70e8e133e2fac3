import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
from diffuvolume_b200 import ops
from diffuvolume_b200.pipeline import AcvHotPath
from oracle import torch_port as P, dv_oracle as O
dev = torch.device("cuda", 0)
for fm, rm in (("regenerate","logits"),("volume","logits"),("regenerate","fused_upsample")):
    g = torch.Generator(device=dev); g.manual_seed(2024)
    B,H,W,D = 1,540,960,48; h,w = H//4, W//4
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
    ru = lambda *s, dt=torch.float32: torch.rand(*s, generator=g, device=dev, dtype=dt)
    fl,fr,cl,cr = rn(B,320,h,w),rn(B,320,h,w),rn(B,32,h,w),rn(B,32,h,w); att = rn(B,1,D,h,w)
    fused = rm == "fused_upsample"
    costs = [(rn(B,1,D,h,w) if fused else rn(B,192,H,W))*4.0 for _ in range(5)]
    used = ru(B,H,W)*191.0; disp_q = ru(B,h,w)*47.75
    shifts = [rn(B,D)*0.1 for _ in range(5)]
    sn = [rn(B,D,h,w,dt=torch.float32 if i==0 else torch.float64) for i in range(4)]
    rz = [ru(B,D,h,w,dt=torch.float64) for _ in range(4)]
    out = AcvHotPath(filter_mode=fm, regress_mode=rm)(fl,fr,cl,cr,att,costs,used,disp_q,shifts,sn,rz,keep_volumes=True)
    asd = ops.xstart_from_disp(disp_q, D, 1.0)
    wp,(wg,wi,wm) = P.hot_path_pair(fl,fr,cl,cr,att,costs,used,asd,shifts,sn,rz,sched=O.Schedule(),upsample_to=(192,H,W) if fused else None)
    wa = P.acv_volume(att, P.concat_volume(cl,cr,D,mask_left=False))
    print(fm, rm, "gwc rel %.2e" % float((out["gwc"]-wg).abs().max()/wg.abs().max()), "ac rel %.2e" % float((out["ac"]-wa).abs().max()/wa.abs().max()),
          "EPE %.2e px" % float((out["pred"]-wp).abs().mean()), "max %.2e" % float((out["pred"]-wp).abs().max()), "mask agree %.5f" % float((out["mask"]==wm).float().mean()))
